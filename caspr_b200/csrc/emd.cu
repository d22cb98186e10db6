// Approximate earth mover's distance: the second reconstruction metric of
// caspr/utils/evaluations.py:45-46, which the reference gets from utils/emd.py:11-12
// (emd_cuda.approxmatch_forward + matchcost_forward — the PyTorchEMD port of Fan et al.'s approxmatch; the extension
// is NOT under /root/reference, its published algorithm is restated in oracle/emd_oracle.py: parity unpinned).
//
// Per cloud pair: ten annealing levels (-4^7 ... -4^-1, 0); each level runs three all-pairs passes
//   1. ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d2(k,l)) remainR[l])
//   2. sumr[l]  = remainR[l] sum_k exp(level d2) ratioL[k]; ratioR[l] = min(remainR[l]/(sumr+1e-9), 1) remainR[l];
//      remainR[l] = max(0, remainR[l] - sumr)
//   3. w(k,l) = exp(level d2) ratioL[k] ratioR[l] is added to the match; remainL[k] = max(0, remainL[k] - sum_l w)
// The upstream kernel runs one CTA per pair and stores the (m x n) match matrix; here every pass is one launch over
// (pair, 256-point block) CTAs, the other cloud is staged through shared memory as float4 (x, y, z, weight), and the
// transport cost sum w(k,l) |p_k - q_l| is accumulated while the match is formed, so the match matrix never exists.
#include "common.cuh"

namespace {

constexpr int kEmdThreads = 256;

__device__ __forceinline__ float d2(float x1, float y1, float z1, const float4& q) {
  const float dx = q.x - x1, dy = q.y - y1, dz = q.z - z1;
  return dx * dx + dy * dy + dz * dz;
}

// mode 1: pass 1 (rows = cloud 1, other = cloud 2 with remainR) ; mode 2: pass 2 (rows = cloud 2, other = cloud 1 with
// ratioL) ; mode 3: pass 3 (rows = cloud 1, other = cloud 2 with ratioR) + cost
template <int MODE>
__global__ void __launch_bounds__(kEmdThreads)
emd_pass_kernel(const float* __restrict__ rows_xyz, int n_rows, const float* __restrict__ other_xyz, int n_other,
                const float* __restrict__ other_w, float level, float* __restrict__ remain_rows,
                float* __restrict__ ratio_rows, float* __restrict__ costpart) {
  __shared__ float4 buf[kEmdThreads];
  const int pair = blockIdx.y;
  const int k = blockIdx.x * kEmdThreads + threadIdx.x;
  const float* rx = rows_xyz + (size_t)pair * n_rows * 3;
  const float* ox = other_xyz + (size_t)pair * n_other * 3;
  const float* ow = other_w + (size_t)pair * n_other;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f, own = 0.f;
  if (k < n_rows) {
    x1 = rx[3 * k]; y1 = rx[3 * k + 1]; z1 = rx[3 * k + 2];
    if (MODE == 3) own = ratio_rows[(size_t)pair * n_rows + k];
  }
  float sum = MODE == 1 ? 1e-9f : 0.f;
  float cost = 0.f;
  for (int l0 = 0; l0 < n_other; l0 += kEmdThreads) {
    const int l = l0 + threadIdx.x;
    buf[threadIdx.x] = l < n_other ? make_float4(ox[3 * l], ox[3 * l + 1], ox[3 * l + 2], ow[l])
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lend = min(kEmdThreads, n_other - l0);
    for (int j = 0; j < lend; ++j) {
      const float4 q = buf[j];
      const float dd = d2(x1, y1, z1, q);
      float w = __expf(level * dd) * q.w;
      if (MODE == 3) {
        w *= own;
        cost = fmaf(w, sqrtf(dd), cost);
      }
      sum += w;
    }
    __syncthreads();
  }
  if (k < n_rows) {
    const size_t i = (size_t)pair * n_rows + k;
    if (MODE == 1) {
      ratio_rows[i] = remain_rows[i] / sum;
    } else if (MODE == 2) {
      const float r = remain_rows[i];
      const float sumr = sum * r;
      const float consumption = fminf(r / (sumr + 1e-9f), 1.0f);
      ratio_rows[i] = consumption * r;
      remain_rows[i] = fmaxf(0.0f, r - sumr);
    } else {
      remain_rows[i] = fmaxf(0.0f, remain_rows[i] - sum);
    }
  }
  if (MODE == 3) {
    // block partial of the transport cost, accumulated over the levels in the block's own slot (deterministic)
    __shared__ float red[kEmdThreads / 32];
    cost = warp_sum(k < n_rows ? cost : 0.f);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < kEmdThreads / 32; ++w) s += red[w];
      costpart[(size_t)pair * gridDim.x + blockIdx.x] += s;
    }
  }
}

__global__ void emd_init_kernel(float* __restrict__ remainL, float* __restrict__ remainR, int B, int n, int m,
                                float multiL, float multiR, float* __restrict__ costpart, int nparts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * n) remainL[i] = multiL;
  if (i < B * m) remainR[i] = multiR;
  if (i < B * nparts) costpart[i] = 0.f;
}

__global__ void emd_finish_kernel(const float* __restrict__ costpart, int B, int nparts, float* __restrict__ cost) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += costpart[(size_t)b * nparts + p];
  cost[b] = s;
}

}  // namespace

extern "C" size_t caspr_emd_workspace_bytes(int B, int n, int m) {
  if (B <= 0 || n <= 0 || m <= 0) return 0;
  return ((size_t)B * (2 * n + 2 * m) + (size_t)B * ceil_div(n, kEmdThreads)) * sizeof(float) + 1024;
}

extern "C" int caspr_emd(const float* xyz1, const float* xyz2, int B, int n, int m, float* cost, void* workspace,
                         size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(xyz1 && xyz2 && cost && workspace && B > 0 && B <= 65535 && n > 0 && m > 0);
  if (workspace_bytes < caspr_emd_workspace_bytes(B, n, m)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* remainL = (float*)workspace;
  float* remainR = remainL + (size_t)B * n;
  float* ratioL = remainR + (size_t)B * m;
  float* ratioR = ratioL + (size_t)B * n;
  float* costpart = ratioR + (size_t)B * m;
  const int nparts = ceil_div(n, kEmdThreads);
  // integer division as upstream: multiL = m/n, multiR = n/m on ints
  const float multiL = n >= m ? 1.f : (float)(m / n);
  const float multiR = n >= m ? (float)(n / m) : 1.f;
  const int tot = B * (n > m ? n : m);
  CASPR_COUNT(); emd_init_kernel<<<ceil_div(tot, 256), 256, 0, s>>>(remainL, remainR, B, n, m, multiL, multiR, costpart, nparts);
  const dim3 gl(ceil_div(n, kEmdThreads), B), gr(ceil_div(m, kEmdThreads), B);
  for (int j = 7; j >= -2; --j) {
    const float level = j == -2 ? 0.f : -powf(4.0f, (float)j);
    CASPR_COUNT(); emd_pass_kernel<1><<<gl, kEmdThreads, 0, s>>>(xyz1, n, xyz2, m, remainR, level, remainL, ratioL, nullptr);
    CASPR_COUNT(); emd_pass_kernel<2><<<gr, kEmdThreads, 0, s>>>(xyz2, m, xyz1, n, ratioL, level, remainR, ratioR, nullptr);
    CASPR_COUNT(); emd_pass_kernel<3><<<gl, kEmdThreads, 0, s>>>(xyz1, n, xyz2, m, ratioR, level, remainL, ratioL, costpart);
  }
  CASPR_COUNT(); emd_finish_kernel<<<ceil_div(B, 128), 128, 0, s>>>(costpart, B, nparts, cost);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
