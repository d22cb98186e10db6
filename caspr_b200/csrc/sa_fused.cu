// One scale of a set-abstraction level in ONE kernel: group gather -> three [Conv1d(k=1) -> GroupNorm(16) -> ReLU?]
// layers -> max over the ball.  Replaces, for the levels whose layers are at most 64 wide (SA levels 1-2),
// PointNet2GroupingLayer + PointNetFeatureExtractor of caspr/models/pointnet2.py:391-401,649-708 (Kaolin group gather,
// three Conv1d/GroupNorm launches, max) — previously group_points_kernel + three linear_kernel<GN_BALL> launches that
// moved every intermediate (balls*ns x C) tensor through HBM.
//
// One warp owns 32/NS balls, a lane owns one grouped row.  The row's activations stay in registers from the gather to
// the max; the weights sit transposed in shared memory ([k][c], read as broadcast float4); GroupNorm statistics
// (per ball: ns rows x C/16 channels) and the max-pool are shuffle reductions over the ball's NS lanes.  Statistics are
// two-pass and shifted by the ball's first row, so padded balls (ns copies of one point) normalise to exactly beta, like
// the reference's fp32 GroupNorm.
#include "common.cuh"

namespace {

constexpr int kSaMaxCin = 136;      // 3 + 131 input channels at most (levels 1-2 use 9 and 99)

template <int NS>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int o = NS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NS>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
  for (int o = NS / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// GroupNorm(16, C) over the NS rows of a ball (+ optional ReLU), in place on the lane's C activations
template <int NS, int C, bool RELU>
__device__ __forceinline__ void ball_groupnorm(float (&x)[C], const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int lane) {
  constexpr int CPG = C / 16;
  const float inv_m = 1.f / (float)(NS * CPG);
  const int first = lane & ~(NS - 1);                  // lane holding row 0 of this ball
#pragma unroll
  for (int g = 0; g < 16; ++g) {
    const float ref = __shfl_sync(0xffffffffu, x[g * CPG], first);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) s += x[g * CPG + c] - ref;
    const float mean = ref + seg_sum<NS>(s) * inv_m;
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const float d = x[g * CPG + c] - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = 1.f / sqrtf(seg_sum<NS>(v) * inv_m + eps);
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      const int ch = g * CPG + c;
      float y = fmaf((x[ch] - mean) * rstd, gamma[ch], beta[ch]);
      x[ch] = RELU ? fmaxf(y, 0.f) : y;
    }
  }
}

// out[c] = bias[c] + sum_k Wt[k][c] in[k]   (Wt in shared memory, [CIN][COUT])
template <int CIN, int COUT>
__device__ __forceinline__ void dense_regs(const float (&in)[CIN], const float* __restrict__ Wt,
                                           const float* __restrict__ bias, float (&out)[COUT]) {
#pragma unroll
  for (int c = 0; c < COUT; ++c) out[c] = bias[c];
#pragma unroll
  for (int k = 0; k < CIN; ++k) {
    const float xk = in[k];
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + k * COUT + 4 * c4);
      out[4 * c4] = fmaf(w.x, xk, out[4 * c4]);
      out[4 * c4 + 1] = fmaf(w.y, xk, out[4 * c4 + 1]);
      out[4 * c4 + 2] = fmaf(w.z, xk, out[4 * c4 + 2]);
      out[4 * c4 + 3] = fmaf(w.w, xk, out[4 * c4 + 3]);
    }
  }
}

struct SaLayer {
  const float *W, *b, *gamma, *beta;
};

template <int NS, int C1, int C2, int C3>
__global__ void __launch_bounds__(256, 2)
sa_fused_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const float* __restrict__ feat,
                int ld_feat, int C, const int32_t* __restrict__ idx, int N, int M, long long balls, SaLayer l1,
                SaLayer l2, SaLayer l3, float eps, float* __restrict__ out, int ld_out) {
  __shared__ __align__(16) float sW1[kSaMaxCin * C1];          // [k][c]
  __shared__ __align__(16) float sW2[C1 * C2];
  __shared__ __align__(16) float sW3[C2 * C3];
  __shared__ float sP[3 * (C1 + C2 + C3)];                     // bias, gamma, beta of the three layers
  const int cin = 3 + C;
  for (int i = threadIdx.x; i < cin * C1; i += blockDim.x) {
    const int k = i / C1, c = i - k * C1;
    sW1[i] = l1.W[c * cin + k];
  }
  for (int i = threadIdx.x; i < C1 * C2; i += blockDim.x) {
    const int k = i / C2, c = i - k * C2;
    sW2[i] = l2.W[c * C1 + k];
  }
  for (int i = threadIdx.x; i < C2 * C3; i += blockDim.x) {
    const int k = i / C3, c = i - k * C3;
    sW3[i] = l3.W[c * C2 + k];
  }
  float* sB1 = sP; float* sG1 = sB1 + C1; float* sE1 = sG1 + C1;
  float* sB2 = sE1 + C1; float* sG2 = sB2 + C2; float* sE2 = sG2 + C2;
  float* sB3 = sE2 + C2; float* sG3 = sB3 + C3; float* sE3 = sG3 + C3;
  for (int i = threadIdx.x; i < C1; i += blockDim.x) { sB1[i] = l1.b[i]; sG1[i] = l1.gamma[i]; sE1[i] = l1.beta[i]; }
  for (int i = threadIdx.x; i < C2; i += blockDim.x) { sB2[i] = l2.b[i]; sG2[i] = l2.gamma[i]; sE2[i] = l2.beta[i]; }
  for (int i = threadIdx.x; i < C3; i += blockDim.x) { sB3[i] = l3.b[i]; sG3[i] = l3.gamma[i]; sE3[i] = l3.beta[i]; }
  __syncthreads();

  constexpr int BPW = 32 / NS;                                  // balls per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / NS, row = lane % NS;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const bool vec4 = (C % 4 == 0) && (ld_feat % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
  for (long long wb = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wb * BPW < balls; wb += warps) {
    const long long ball = wb * BPW + sub;
    const bool live = ball < balls;
    const long long bl = live ? ball : balls - 1;               // idle lanes shadow the last ball (never stored)
    const int b = (int)(bl / M);
    const int src = idx[bl * NS + row];
    const float* p = xyz + ((size_t)b * N + src) * 3;
    const float* cpt = new_xyz + (size_t)bl * 3;
    // layer 1 with the gathered row streamed through: [dx, dy, dz | features]
    float h1[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) h1[c] = sB1[c];
    auto feed = [&](int k, float xk) {
#pragma unroll
      for (int c4 = 0; c4 < C1 / 4; ++c4) {
        const float4 w = *reinterpret_cast<const float4*>(sW1 + k * C1 + 4 * c4);
        h1[4 * c4] = fmaf(w.x, xk, h1[4 * c4]);
        h1[4 * c4 + 1] = fmaf(w.y, xk, h1[4 * c4 + 1]);
        h1[4 * c4 + 2] = fmaf(w.z, xk, h1[4 * c4 + 2]);
        h1[4 * c4 + 3] = fmaf(w.w, xk, h1[4 * c4 + 3]);
      }
    };
    feed(0, p[0] - cpt[0]);
    feed(1, p[1] - cpt[1]);
    feed(2, p[2] - cpt[2]);
    const float* fr = feat + ((size_t)b * N + src) * ld_feat;
    if (vec4) {
      // software pipeline: the next 16 bytes of the gathered row are in flight while the current ones are consumed
      float4 f = *reinterpret_cast<const float4*>(fr);
      for (int k = 0; k < C; k += 4) {
        const float4 nxt = k + 4 < C ? *reinterpret_cast<const float4*>(fr + k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        feed(3 + k, f.x); feed(4 + k, f.y); feed(5 + k, f.z); feed(6 + k, f.w);
        f = nxt;
      }
    } else {
      for (int k = 0; k < C; ++k) feed(3 + k, fr[k]);
    }
    ball_groupnorm<NS, C1, true>(h1, sG1, sE1, eps, lane);
    float h2[C2];
    dense_regs<C1, C2>(h1, sW2, sB2, h2);
    ball_groupnorm<NS, C2, true>(h2, sG2, sE2, eps, lane);
    float h3[C3];
    dense_regs<C2, C3>(h2, sW3, sB3, h3);
    ball_groupnorm<NS, C3, false>(h3, sG3, sE3, eps, lane);
    // max over the ball's rows; lane `row` stores the channels congruent to it
#pragma unroll
    for (int c = 0; c < C3; ++c) {
      const float m = seg_max<NS>(h3[c]);
      if (live && (c % NS) == row) out[(size_t)ball * ld_out + c] = m;
    }
  }
}

template <int NS, int C1, int C2, int C3>
int launch_sa(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C, const int32_t* idx, int N,
              int M, long long balls, const SaLayer& l1, const SaLayer& l2, const SaLayer& l3, float eps, float* out,
              int ld_out, cudaStream_t s) {
  const long long warps = (balls * NS + 31) / 32;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CASPR_COUNT(); sa_fused_kernel<NS, C1, C2, C3><<<(int)blocks, 256, 0, s>>>(xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls,
                                                                          l1, l2, l3, eps, out, ld_out);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace

extern "C" int caspr_sa_fused_supported(int ns, int Cin, int C1, int C2, int C3) {
  if (Cin < 3 || Cin > kSaMaxCin) return 0;
  return (ns == 16 && C1 == 16 && C2 == 16 && C3 == 32) || (ns == 32 && C1 == 32 && C2 == 32 && C3 == 64) ||
         (ns == 16 && C1 == 32 && C2 == 32 && C3 == 64);
}

extern "C" int caspr_sa_fused(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                              const int32_t* idx, int B, int N, int M, int ns,
                              const float* W1, const float* b1, const float* g1, const float* e1, int C1,
                              const float* W2, const float* b2, const float* g2, const float* e2, int C2,
                              const float* W3, const float* b3, const float* g3, const float* e3, int C3,
                              float eps, float* out, int ld_out, void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && idx && out && B > 0 && N > 0 && M > 0 && C >= 0 && (C == 0 || (feat && ld_feat >= C)));
  CASPR_REQUIRE(W1 && b1 && g1 && e1 && W2 && b2 && g2 && e2 && W3 && b3 && g3 && e3 && ld_out >= C3);
  CASPR_REQUIRE(caspr_sa_fused_supported(ns, 3 + C, C1, C2, C3));
  const SaLayer l1 = {W1, b1, g1, e1}, l2 = {W2, b2, g2, e2}, l3 = {W3, b3, g3, e3};
  const long long balls = (long long)B * M;
  cudaStream_t s = (cudaStream_t)stream;
  if (ns == 16 && C1 == 16)
    return launch_sa<16, 16, 16, 32>(xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls, l1, l2, l3, eps, out, ld_out, s);
  if (ns == 32)
    return launch_sa<32, 32, 32, 64>(xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls, l1, l2, l3, eps, out, ld_out, s);
  return launch_sa<16, 32, 32, 64>(xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls, l1, l2, l3, eps, out, ld_out, s);
}
