// Training-mode (forward with saved statistics + backward) operators of the TPointNet++ encoder, BASELINE
// config 5.  They replace what torch autograd does in the reference for the Conv1d(k=1) / GroupNorm / ReLU / max
// chains (pointnet2.py:637-642,677-699,471-481,207-212; pointnet.py:27-46; tpointnet2.py:59-62,99-112) and the
// backward of Kaolin's group-gather and three_interpolate Functions (pointnet2.py:391,519).
// All tensors are channels-last rows with a leading dimension, like the inference operators; exact fp32 SIMT.
#include <string.h>
#include "common.cuh"

namespace {

int grid_for(long long work, int per_block, int cap) {
  long long b = (work + per_block - 1) / per_block;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// rows of a sample are split into chunks so that chunks*samples CTAs fill the GPU (>= 32 rows per chunk)
int gn_chunks(int samples, int rps) {
  int c = (2 * 148 + samples - 1) / samples;
  const int maxc = (rps + 31) / 32;
  if (c > maxc) c = maxc;
  if (c < 1) c = 1;
  return c;
}

// ------------------------------------------------------------------------------ GroupNorm forward
// partial[(s*chunks + chunk)*C + c] = (mean, sum of squared deviations from it) of channel c over the chunk's rows.
// Two passes per channel: sum-of-squares minus squared mean cancels catastrophically on the (frequent) balls whose
// rows are copies of one point, where the rounding noise would then be divided by sqrt(eps).
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ X, int ldx, int rps, int C, int rows_per_chunk, float2* __restrict__ partial) {
  const int s = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(rps, r0 + rows_per_chunk);
  const float* base = X + (size_t)s * rps * ldx;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // sums are taken relative to the first row, so that identical rows give an exact mean
    const float x0 = r1 > r0 ? base[(size_t)r0 * ldx + c] : 0.f;
    float a = 0.f;
    for (int r = r0; r < r1; ++r) a += base[(size_t)r * ldx + c] - x0;
    const float mean = r1 > r0 ? x0 + a / (float)(r1 - r0) : 0.f;
    float m2 = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float d = base[(size_t)r * ldx + c] - mean;
      m2 = fmaf(d, d, m2);
    }
    partial[((size_t)s * chunks + chunk) * C + c] = make_float2(mean, m2);
  }
}

// mean / rstd per (sample, group): the (count, mean, M2) triples of the group's channels and chunks are merged in
// fp64 (M2 = sum M2_i + sum n_i (mean_i - mean)^2)
__global__ void __launch_bounds__(128)
gn_finish_kernel(const float2* __restrict__ partial, int samples, int chunks, int C, int groups, int rps,
                 int rows_per_chunk, float eps, float2* __restrict__ mean_rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= samples * groups) return;
  const int s = i / groups, g = i - s * groups;
  const int cpg = C / groups;
  double wsum = 0.0;
  for (int ch = 0; ch < chunks; ++ch) {
    const int n = min(rps, (ch + 1) * rows_per_chunk) - ch * rows_per_chunk;
    if (n <= 0) continue;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) wsum += (double)n * partial[((size_t)s * chunks + ch) * C + c].x;
  }
  const double m = (double)rps * cpg;
  const double mean = wsum / m;
  double m2 = 0.0;
  for (int ch = 0; ch < chunks; ++ch) {
    const int n = min(rps, (ch + 1) * rows_per_chunk) - ch * rows_per_chunk;
    if (n <= 0) continue;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float2 p = partial[((size_t)s * chunks + ch) * C + c];
      const double d = (double)p.x - mean;
      m2 += (double)p.y + (double)n * d * d;
    }
  }
  mean_rstd[i] = make_float2((float)mean, (float)(1.0 / sqrt(m2 / m + (double)eps)));
}

// The same for samples split into many chunks (clouds): one WARP per (sample, group), lanes stride over the
// (chunk, channel) partials - the thread-per-group loop above is a latency-bound serial walk there.
__global__ void __launch_bounds__(256)
gn_finish_warp_kernel(const float2* __restrict__ partial, int samples, int chunks, int C, int groups, int rps,
                      int rows_per_chunk, float eps, float2* __restrict__ mean_rstd) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= samples * groups) return;
  const int s = i / groups, g = i - s * groups;
  const int cpg = C / groups;
  const int items = chunks * cpg;
  double wsum = 0.0;
  for (int it = lane; it < items; it += 32) {
    const int ch = it / cpg, c = g * cpg + (it - ch * cpg);
    const int n = min(rps, (ch + 1) * rows_per_chunk) - ch * rows_per_chunk;
    if (n > 0) wsum += (double)n * partial[((size_t)s * chunks + ch) * C + c].x;
  }
  wsum = warp_sum_d(wsum);
  const double m = (double)rps * cpg;
  const double mean = wsum / m;
  double m2 = 0.0;
  for (int it = lane; it < items; it += 32) {
    const int ch = it / cpg, c = g * cpg + (it - ch * cpg);
    const int n = min(rps, (ch + 1) * rows_per_chunk) - ch * rows_per_chunk;
    if (n > 0) {
      const float2 p = partial[((size_t)s * chunks + ch) * C + c];
      const double d = (double)p.x - mean;
      m2 += (double)p.y + (double)n * d * d;
    }
  }
  m2 = warp_sum_d(m2);
  if (lane == 0) mean_rstd[i] = make_float2((float)mean, (float)(1.0 / sqrt(m2 / m + (double)eps)));
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ X, int ldx, const float2* __restrict__ mean_rstd, long long rows, int rps,
                int C, int groups, const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                float* __restrict__ Y, int ldy) {
  const int cpg = C / groups;
  const long long total = rows * C;
  auto body = [&](auto r, int c) {                     // r: int (fast path) or long long
    const float2 mr = mean_rstd[(long long)(r / rps) * groups + c / cpg];
    float v = fmaf((X[(long long)r * ldx + c] - mr.x) * mr.y, gamma[c], beta[c]);
    if (relu) v = fmaxf(v, 0.f);
    Y[(long long)r * ldy + c] = v;
  };
  if (total < (1ll << 31)) {          // 32-bit index arithmetic: the 64-bit divisions dominate this kernel otherwise
    const unsigned tot = (unsigned)total, stride = gridDim.x * blockDim.x, Cu = (unsigned)C;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
      const unsigned r = i / Cu;
      body((int)r, (int)(i - r * Cu));
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / C;
      body(r, (int)(i - r * C));
    }
  }
}

// max over the rows of each sample and the first row attaining it, in row splits
__global__ void __launch_bounds__(128)
rowmax_partial_kernel(const float* __restrict__ Y, int ldy, int rps, int C, int rows_per_split,
                      float* __restrict__ pmax, int32_t* __restrict__ parg) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int s = blockIdx.x, sp = blockIdx.z, nsplit = gridDim.z;
  if (c >= C) return;
  const int r0 = sp * rows_per_split, r1 = min(rps, r0 + rows_per_split);
  const float* base = Y + (size_t)s * rps * ldy + c;
  float best = -INFINITY;
  int arg = r0;
  for (int r = r0; r < r1; ++r) {
    const float v = base[(size_t)r * ldy];
    if (v > best) { best = v; arg = r; }
  }
  pmax[((size_t)s * nsplit + sp) * C + c] = best;
  parg[((size_t)s * nsplit + sp) * C + c] = arg;
}
__global__ void __launch_bounds__(128)
rowmax_finish_kernel(const float* __restrict__ pmax, const int32_t* __restrict__ parg, int nsplit, int C,
                     float* __restrict__ maxout, int ld_max, int32_t* __restrict__ argmax) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int s = blockIdx.x;
  if (c >= C) return;
  float best = -INFINITY;
  int arg = 0;
  for (int sp = 0; sp < nsplit; ++sp) {
    const float v = pmax[((size_t)s * nsplit + sp) * C + c];
    if (v > best || sp == 0) { best = v; arg = parg[((size_t)s * nsplit + sp) * C + c]; }
  }
  maxout[(size_t)s * ld_max + c] = best;
  argmax[(size_t)s * C + c] = arg;
}

// ----------------------------------------------------------------------------- GroupNorm backward
// Effective output cotangent of one element: dense part + max-pool routing, through the optional ReLU.
__device__ __forceinline__ float gn_dy_eff(const float* __restrict__ dY, int lddy, const float* __restrict__ dMax,
                                           int ld_dmax, const int32_t* __restrict__ argmax, long long r, int s,
                                           int r_in_sample, int c, int C, float y, int relu) {
  float d = dY ? dY[r * lddy + c] : 0.f;
  if (dMax && argmax[(size_t)s * C + c] == r_in_sample) d += dMax[(size_t)s * ld_dmax + c];
  if (relu && !(y > 0.f)) d = 0.f;
  return d;
}

// partial[(s*chunks+chunk)*C + c] = (sum dY, sum dY*xhat)
__global__ void __launch_bounds__(256)
gn_bwd_partial_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ dMax, int ld_dmax,
                      const int32_t* __restrict__ argmax, const float* __restrict__ X, int ldx,
                      const float2* __restrict__ mean_rstd, int rps, int C, int groups,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int relu, int rows_per_chunk,
                      float2* __restrict__ partial) {
  const int s = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(rps, r0 + rows_per_chunk);
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float2 mr = mean_rstd[s * groups + c / cpg];
    const float gm = gamma[c], bt = beta[c];
    float a = 0.f, b = 0.f;
    for (int r = r0; r < r1; ++r) {
      const long long row = (long long)s * rps + r;
      const float xh = (X[row * ldx + c] - mr.x) * mr.y;
      const float y = fmaf(xh, gm, bt);
      const float d = gn_dy_eff(dY, lddy, dMax, ld_dmax, argmax, row, s, r, c, C, y, relu);
      a += d;
      b = fmaf(d, xh, b);
    }
    partial[((size_t)s * chunks + chunk) * C + c] = make_float2(a, b);
  }
}

// gs[s*groups+g] = (sum_c gamma_c a_c, sum_c gamma_c b_c) / m
__global__ void __launch_bounds__(128)
gn_bwd_group_kernel(const float2* __restrict__ partial, int samples, int chunks, int C, int groups, int rps,
                    const float* __restrict__ gamma, float2* __restrict__ gs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= samples * groups) return;
  const int s = i / groups, g = i - s * groups;
  const int cpg = C / groups;
  double a = 0.0, b = 0.0;
  for (int ch = 0; ch < chunks; ++ch)
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float2 p = partial[((size_t)s * chunks + ch) * C + c];
      a += (double)gamma[c] * p.x;
      b += (double)gamma[c] * p.y;
    }
  const double m = (double)rps * cpg;
  gs[i] = make_float2((float)(a / m), (float)(b / m));
}

__global__ void __launch_bounds__(256)
gn_bwd_group_warp_kernel(const float2* __restrict__ partial, int samples, int chunks, int C, int groups, int rps,
                         const float* __restrict__ gamma, float2* __restrict__ gs) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= samples * groups) return;
  const int s = i / groups, g = i - s * groups;
  const int cpg = C / groups;
  const int items = chunks * cpg;
  double a = 0.0, b = 0.0;
  for (int it = lane; it < items; it += 32) {
    const int ch = it / cpg, c = g * cpg + (it - ch * cpg);
    const float2 p = partial[((size_t)s * chunks + ch) * C + c];
    a += (double)gamma[c] * p.x;
    b += (double)gamma[c] * p.y;
  }
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  const double m = (double)rps * cpg;
  if (lane == 0) gs[i] = make_float2((float)(a / m), (float)(b / m));
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ dMax, int ld_dmax,
                    const int32_t* __restrict__ argmax, const float* __restrict__ X, int ldx,
                    const float2* __restrict__ mean_rstd, const float2* __restrict__ gs, long long rows, int rps, int C,
                    int groups, const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                    float* __restrict__ dX, int lddx) {
  const int cpg = C / groups;
  const long long total = rows * C;
  auto body = [&](auto r, int c) {                     // r: int (fast path) or long long
    const int s = (int)(r / rps);
    const int g = c / cpg;
    const float2 mr = mean_rstd[s * groups + g];
    const float2 sg = gs[s * groups + g];
    const float xh = (X[(long long)r * ldx + c] - mr.x) * mr.y;
    const float y = fmaf(xh, gamma[c], beta[c]);
    const float d = gn_dy_eff(dY, lddy, dMax, ld_dmax, argmax, (long long)r, s, (int)(r - s * rps), c, C, y, relu);
    dX[(long long)r * lddx + c] = mr.y * (d * gamma[c] - sg.x - xh * sg.y);
  };
  if (total < (1ll << 31)) {
    const unsigned tot = (unsigned)total, stride = gridDim.x * blockDim.x, Cu = (unsigned)C;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
      const unsigned r = i / Cu;
      body((int)r, (int)(i - r * Cu));
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / C;
      body(r, (int)(i - r * C));
    }
  }
}

// ------------------------------------------------------------------------------- column sums
// partial[split][c] = sum over the split's rows of X[r][c]
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const float* __restrict__ X, int ldx, long long rows, int C, long long rows_per_split,
                      float* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long r0 = blockIdx.y * rows_per_split;
  const long long r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long long r = r0;
  for (; r + 3 < r1; r += 4) {
    a0 += X[r * ldx + c];
    a1 += X[(r + 1) * ldx + c];
    a2 += X[(r + 2) * ldx + c];
    a3 += X[(r + 3) * ldx + c];
  }
  for (; r < r1; ++r) a0 += X[r * ldx + c];
  partial[(size_t)blockIdx.y * C + c] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float* __restrict__ part, int nparts, size_t nelem, float* __restrict__ dst, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * nelem + i];
  dst[i] = accumulate ? dst[i] + s : s;
}

// ---------------------------------------------------------------------- linear weight gradient
// dW[o][i] = sum_r dY[r][o] * act(X[r][i]) over one split of the rows.  CTA tile 64 x 64, k-slab 16 rows,
// 256 threads, 4 x 4 per thread; bounds-checked scalar loads (any Cout / Cin / leading dimension).
constexpr int kLwTile = 64, kLwRows = 16;
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx, long long rows,
                    int Cout, int Cin, int relu_x, long long rows_per_split, float* __restrict__ part) {
  __shared__ float As[2][kLwRows][kLwTile + 4];
  __shared__ float Bs[2][kLwRows][kLwTile + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int o0 = blockIdx.x * kLwTile, i0 = blockIdx.y * kLwTile;
  const long long r_begin = blockIdx.z * rows_per_split;
  const long long r_end = r_begin + rows_per_split < rows ? r_begin + rows_per_split : rows;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // loader: 16 rows x 64 channels per operand = 1024 elements, 4 per thread: row lr, channels lc + 16*q
  const int lr = tid >> 4, lc = tid & 15;
  float ra[4], rb[4];
  auto load = [&](long long r0) {
    const long long r = r0 + lr;
    const bool ok = r < r_end;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = o0 + lc + 16 * q, i = i0 + lc + 16 * q;
      ra[q] = (ok && o < Cout) ? dY[r * lddy + o] : 0.f;
      float xv = (ok && i < Cin) ? X[r * ldx + i] : 0.f;
      if (relu_x) xv = fmaxf(xv, 0.f);
      rb[q] = xv;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      As[buf][lr][lc + 16 * q] = ra[q];
      Bs[buf][lr][lc + 16 * q] = rb[q];
    }
  };
  const long long span = r_end > r_begin ? r_end - r_begin : 0;
  const int nslab = (int)((span + kLwRows - 1) / kLwRows);
  if (nslab > 0) {
    load(r_begin);
    store(0);
  }
  __syncthreads();
  for (int sl = 0; sl < nslab; ++sl) {
    const int buf = sl & 1;
    if (sl + 1 < nslab) load(r_begin + (long long)(sl + 1) * kLwRows);
#pragma unroll
    for (int k = 0; k < kLwRows; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    if (sl + 1 < nslab) {
      store(buf ^ 1);
      __syncthreads();
    }
  }
  float* dst = part + (size_t)blockIdx.z * Cout * Cin;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int o = o0 + ty * 4 + p;
    if (o >= Cout) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + tx * 4 + q;
      if (i < Cin) dst[(size_t)o * Cin + i] = acc[p][q];
    }
  }
}

// ------------------------------------------------------------------------ gather backward
// dfeat[b][idx[b][m][k]][c] += dOut[(b*M+m)*ns + k][3 + c]
__global__ void __launch_bounds__(256)
group_points_bwd_kernel(const float* __restrict__ dOut, int ld_out, const int32_t* __restrict__ idx, int N, int M,
                        int C, int ns, long long total_rows, float* __restrict__ dfeat, int ld_feat) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < total_rows; row += stride) {
    const long long b = row / ((long long)M * ns);
    const int src = idx[row];
    float* dst = dfeat + ((size_t)b * N + src) * ld_feat;
    const float* g = dOut + row * ld_out + 3;
    for (int c = lane; c < C; c += 32) atomicAdd(dst + c, g[c]);
  }
}

// dprev[b][idx[row][k]][c] += w_k(row) * dOut[row][c]   (weights as in three_interp_concat_kernel)
__global__ void __launch_bounds__(256)
three_interp_bwd_kernel(const float* __restrict__ dOut, int ld_out, const int32_t* __restrict__ idx,
                        const float* __restrict__ dist, int n, int m, int Cp, long long total_rows,
                        float* __restrict__ dprev, int ld_prev) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < total_rows; row += stride) {
    const int b = (int)(row / n);
    const float d0 = dist[row * 3 + 0], d1 = dist[row * 3 + 1], d2 = dist[row * 3 + 2];
    const float v0 = __fdiv_rn(1.0f, __fadd_rn(d0, 1e-8f));
    const float v1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f));
    const float v2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f));
    const float tot = __fadd_rn(__fadd_rn(v0, v1), v2);
    const float w0 = __fdiv_rn(v0, tot), w1 = __fdiv_rn(v1, tot), w2 = __fdiv_rn(v2, tot);
    float* f0 = dprev + ((size_t)b * m + idx[row * 3 + 0]) * ld_prev;
    float* f1 = dprev + ((size_t)b * m + idx[row * 3 + 1]) * ld_prev;
    float* f2 = dprev + ((size_t)b * m + idx[row * 3 + 2]) * ld_prev;
    const float* g = dOut + row * ld_out;
    for (int c = lane; c < Cp; c += 32) {
      const float gv = g[c];
      atomicAdd(f0 + c, w0 * gv);
      atomicAdd(f1 + c, w1 * gv);
      atomicAdd(f2 + c, w2 * gv);
    }
  }
}

// dst[r][c] (op)= src[r][c]: mode 0 copy, 1 add; optional mask: keep only where ref[r][c] > 0
__global__ void __launch_bounds__(256)
rows_update_kernel(const float* __restrict__ src, int ld_src, long long rows, int C, int mode,
                   const float* __restrict__ ref, int ld_ref, float* __restrict__ dst, int ld_dst) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    float v = src[r * ld_src + c];
    if (ref && !(ref[r * ld_ref + c] > 0.f)) v = 0.f;
    dst[r * ld_dst + c] = mode ? dst[r * ld_dst + c] + v : v;
  }
}

__global__ void transpose2d_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = blockIdx.y * 32 + r;
    if (x < cols && y < rows) tile[r][threadIdx.x] = src[(size_t)y * cols + x];
  }
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int yo = blockIdx.x * 32 + r;
    if (xo < rows && yo < cols) dst[(size_t)yo * rows + xo] = tile[threadIdx.x][r];
  }
}

int wgrad_splits(long long rows, int Cout, int Cin) {
  const long long tiles = (long long)ceil_div(Cout, kLwTile) * ceil_div(Cin, kLwTile);
  long long s = (2 * 148 + tiles - 1) / tiles;
  const long long maxs = (rows + 4 * kLwRows - 1) / (4 * kLwRows);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}
int colsum_splits(long long rows, int C) {
  long long s = (2 * 148 + ceil_div(C, 128) - 1) / ceil_div(C, 128);
  const long long maxs = (rows + 63) / 64;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace

extern "C" size_t caspr_gn_workspace_bytes(int samples, int rows_per_sample, int C) {
  if (samples <= 0 || rows_per_sample <= 0 || C <= 0) return 0;
  const size_t chunks = gn_chunks(samples, rows_per_sample);
  // per-channel partials + (sample, group) pairs + column-sum partials of the parameter gradients
  return align_up((size_t)samples * chunks * C * sizeof(float2), 256) + align_up((size_t)samples * 64 * sizeof(float2), 256) +
         align_up((size_t)colsum_splits((long long)samples * chunks, 2 * C) * 2 * C * sizeof(float), 256) +
         align_up((size_t)2 * C * sizeof(float), 256);
}

extern "C" int caspr_gn_moments(const float* X, int ldx, int samples, int rows_per_sample, int C, int groups, float eps,
                              float* mean_rstd, void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(X && mean_rstd && workspace && samples > 0 && rows_per_sample > 0 && C > 0);
  CASPR_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && ldx >= C);
  if (workspace_bytes < caspr_gn_workspace_bytes(samples, rows_per_sample, C)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = gn_chunks(samples, rows_per_sample);
  const int rpc = ceil_div(rows_per_sample, chunks);
  float2* partial = (float2*)workspace;
  const int threads = C >= 256 ? 256 : (C + 31) / 32 * 32;
  CASPR_COUNT(); gn_partial_kernel<<<dim3(samples, chunks), threads, 0, s>>>(X, ldx, rows_per_sample, C, rpc, partial);
  if (chunks * (C / groups) >= 64) {
    CASPR_COUNT(); gn_finish_warp_kernel<<<ceil_div(samples * groups, 8), 256, 0, s>>>(
        partial, samples, chunks, C, groups, rows_per_sample, rpc, eps, (float2*)mean_rstd);
  } else {
    CASPR_COUNT(); gn_finish_kernel<<<ceil_div(samples * groups, 128), 128, 0, s>>>(
        partial, samples, chunks, C, groups, rows_per_sample, rpc, eps, (float2*)mean_rstd);
  }
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_gn_apply(const float* X, int ldx, const float* mean_rstd, int samples, int rows_per_sample, int C,
                              int groups, const float* gamma, const float* beta, int relu, float* Y, int ldy,
                              void* stream) {
  CASPR_REQUIRE(X && mean_rstd && gamma && beta && Y && samples > 0 && rows_per_sample > 0 && C > 0);
  CASPR_REQUIRE(groups > 0 && C % groups == 0 && ldx >= C && ldy >= C);
  const long long rows = (long long)samples * rows_per_sample;
  CASPR_COUNT(); gn_apply_kernel<<<grid_for(rows * C, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      X, ldx, (const float2*)mean_rstd, rows, rows_per_sample, C, groups, gamma, beta, relu, Y, ldy);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" size_t caspr_rowmax_workspace_bytes(int samples, int rows_per_sample, int C) {
  if (samples <= 0 || rows_per_sample <= 0 || C <= 0) return 0;
  const size_t nsplit = ceil_div(rows_per_sample, 256);
  return 2 * align_up((size_t)samples * nsplit * C * 4, 256);
}

extern "C" int caspr_rowmax(const float* Y, int ldy, int samples, int rows_per_sample, int C, float* maxout, int ld_max,
                            int32_t* argmax, void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(Y && maxout && argmax && workspace && samples > 0 && rows_per_sample > 0 && C > 0 && ldy >= C && ld_max >= C);
  if (workspace_bytes < caspr_rowmax_workspace_bytes(samples, rows_per_sample, C)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int nsplit = ceil_div(rows_per_sample, 256);
  const int rps_split = ceil_div(rows_per_sample, nsplit);
  float* pmax = (float*)workspace;
  int32_t* parg = (int32_t*)((char*)workspace + align_up((size_t)samples * nsplit * C * 4, 256));
  CASPR_REQUIRE(nsplit <= 65535);
  CASPR_COUNT(); rowmax_partial_kernel<<<dim3(samples, ceil_div(C, 128), nsplit), 128, 0, s>>>(Y, ldy, rows_per_sample, C,
                                                                                           rps_split, pmax, parg);
  CASPR_COUNT(); rowmax_finish_kernel<<<dim3(samples, ceil_div(C, 128)), 128, 0, s>>>(pmax, parg, nsplit, C, maxout, ld_max,
                                                                                  argmax);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_gn_backward(const float* dY, int lddy, const float* dMax, int ld_dmax, const int32_t* argmax,
                                 const float* X, int ldx, const float* mean_rstd, int samples, int rows_per_sample,
                                 int C, int groups, const float* gamma, const float* beta, int relu, float* dX,
                                 int lddx, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  CASPR_REQUIRE((dY || dMax) && X && mean_rstd && gamma && beta && dX && dgamma && dbeta && workspace);
  CASPR_REQUIRE(!dMax || argmax);
  CASPR_REQUIRE(samples > 0);
  CASPR_REQUIRE(rows_per_sample > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0);
  if (workspace_bytes < caspr_gn_workspace_bytes(samples, rows_per_sample, C)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int chunks = gn_chunks(samples, rows_per_sample);
  const int rpc = ceil_div(rows_per_sample, chunks);
  char* p = (char*)workspace;
  float2* partial = (float2*)p;
  p += align_up((size_t)samples * chunks * C * sizeof(float2), 256);
  float2* gs = (float2*)p;
  p += align_up((size_t)samples * 64 * sizeof(float2), 256);
  float* cpart = (float*)p;
  const long long prow = (long long)samples * chunks;
  const int csplit = colsum_splits(prow, 2 * C);
  p += align_up((size_t)csplit * 2 * C * sizeof(float), 256);
  float* csum = (float*)p;
  const int threads = C >= 256 ? 256 : (C + 31) / 32 * 32;
  CASPR_COUNT(); gn_bwd_partial_kernel<<<dim3(samples, chunks), threads, 0, s>>>(
      dY, lddy, dMax, ld_dmax, argmax, X, ldx, (const float2*)mean_rstd, rows_per_sample, C, groups, gamma, beta, relu,
      rpc, partial);
  if (chunks * (C / groups) >= 64) {
    CASPR_COUNT(); gn_bwd_group_warp_kernel<<<ceil_div(samples * groups, 8), 256, 0, s>>>(partial, samples, chunks, C, groups,
                                                                                         rows_per_sample, gamma, gs);
  } else {
    CASPR_COUNT(); gn_bwd_group_kernel<<<ceil_div(samples * groups, 128), 128, 0, s>>>(partial, samples, chunks, C, groups,
                                                                                       rows_per_sample, gamma, gs);
  }
  // parameter gradients: column sums of the partials viewed as (samples*chunks) x 2C interleaved (dbeta, dgamma)
  const long long rps_split = (prow + csplit - 1) / csplit;
  CASPR_COUNT(); colsum_partial_kernel<<<dim3(ceil_div(2 * C, 128), csplit), 128, 0, s>>>((const float*)partial, 2 * C, prow,
                                                                                      2 * C, rps_split, cpart);
  CASPR_COUNT(); sum_parts_kernel<<<ceil_div(2 * C, 256), 256, 0, s>>>(cpart, csplit, (size_t)2 * C, csum, 0);
  if (cudaMemcpy2DAsync(dbeta, 4, csum, 8, 4, C, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpy2DAsync(dgamma, 4, csum + 1, 8, 4, C, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  const long long rows = (long long)samples * rows_per_sample;
  CASPR_COUNT(); gn_bwd_apply_kernel<<<grid_for(rows * C, 256 * 4, 148 * 16), 256, 0, s>>>(
      dY, lddy, dMax, ld_dmax, argmax, X, ldx, (const float2*)mean_rstd, gs, rows, rows_per_sample, C, groups, gamma,
      beta, relu, dX, lddx);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" size_t caspr_linear_wgrad_workspace_bytes(long long rows, int Cout, int Cin) {
  if (rows <= 0 || Cout <= 0 || Cin <= 0) return 0;
  return align_up((size_t)wgrad_splits(rows, Cout, Cin) * Cout * Cin * 4, 256) +
         align_up((size_t)colsum_splits(rows, Cout) * Cout * 4, 256);
}

extern "C" int caspr_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, long long rows, int Cout, int Cin,
                                  int relu_x, float* dW, float* db, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  CASPR_REQUIRE(dY && X && dW && workspace && rows > 0 && Cout > 0 && Cin > 0 && lddy >= Cout && ldx >= Cin);
  if (workspace_bytes < caspr_linear_wgrad_workspace_bytes(rows, Cout, Cin)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int splits = wgrad_splits(rows, Cout, Cin);
  long long rps = (rows + splits - 1) / splits;
  rps = (rps + kLwRows - 1) / kLwRows * kLwRows;
  float* part = (float*)workspace;
  const size_t nelem = (size_t)Cout * Cin;
  CASPR_COUNT(); linear_wgrad_kernel<<<dim3(ceil_div(Cout, kLwTile), ceil_div(Cin, kLwTile), splits), 256, 0, s>>>(
      dY, lddy, X, ldx, rows, Cout, Cin, relu_x, rps, part);
  CASPR_COUNT(); sum_parts_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, s>>>(part, splits, nelem, dW, 0);
  if (db) {
    float* cpart = (float*)((char*)workspace + align_up((size_t)splits * nelem * 4, 256));
    const int cs = colsum_splits(rows, Cout);
    const long long crps = (rows + cs - 1) / cs;
    CASPR_COUNT(); colsum_partial_kernel<<<dim3(ceil_div(Cout, 128), cs), 128, 0, s>>>(dY, lddy, rows, Cout, crps, cpart);
    CASPR_COUNT(); sum_parts_kernel<<<ceil_div(Cout, 256), 256, 0, s>>>(cpart, cs, (size_t)Cout, db, 0);
  }
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" size_t caspr_colsum_workspace_bytes(long long rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  return align_up((size_t)colsum_splits(rows, C) * C * 4, 256);
}

/* out[c] (+)= sum_r X[r][c] */
extern "C" int caspr_colsum(const float* X, int ldx, long long rows, int C, float* out, int accumulate, void* workspace,
                            size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(X && out && workspace && rows > 0 && C > 0 && ldx >= C);
  if (workspace_bytes < caspr_colsum_workspace_bytes(rows, C)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int cs = colsum_splits(rows, C);
  const long long crps = (rows + cs - 1) / cs;
  CASPR_COUNT(); colsum_partial_kernel<<<dim3(ceil_div(C, 128), cs), 128, 0, s>>>(X, ldx, rows, C, crps, (float*)workspace);
  CASPR_COUNT(); sum_parts_kernel<<<ceil_div(C, 256), 256, 0, s>>>((const float*)workspace, cs, (size_t)C, out, accumulate);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_group_points_bwd(const float* dOut, int ld_out, const int32_t* idx, int B, int N, int M, int C,
                                      int ns, float* dfeat, int ld_feat, void* stream) {
  CASPR_REQUIRE(dOut && idx && dfeat && B > 0 && N > 0 && M > 0 && C > 0 && ns > 0 && ld_out >= 3 + C && ld_feat >= C);
  const long long rows = (long long)B * M * ns;
  CASPR_COUNT(); group_points_bwd_kernel<<<grid_for(rows, 8, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      dOut, ld_out, idx, N, M, C, ns, rows, dfeat, ld_feat);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_three_interp_bwd(const float* dOut, int ld_out, const int32_t* idx, const float* dist, int B, int n,
                                      int m, int Cp, float* dprev, int ld_prev, void* stream) {
  CASPR_REQUIRE(dOut && idx && dist && dprev && B > 0 && n > 0 && m > 0 && Cp > 0 && ld_out >= Cp && ld_prev >= Cp);
  const long long rows = (long long)B * n;
  CASPR_COUNT(); three_interp_bwd_kernel<<<grid_for(rows, 8, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      dOut, ld_out, idx, dist, n, m, Cp, rows, dprev, ld_prev);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_rows_update(const float* src, int ld_src, long long rows, int C, int accumulate, const float* relu_ref,
                                 int ld_ref, float* dst, int ld_dst, void* stream) {
  CASPR_REQUIRE(src && dst && rows > 0 && C > 0 && ld_src >= C && ld_dst >= C);
  CASPR_COUNT(); rows_update_kernel<<<grid_for(rows * C, 256 * 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, rows, C, accumulate, relu_ref, ld_ref, dst, ld_dst);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_transpose(const float* src, int rows, int cols, float* dst, void* stream) {
  CASPR_REQUIRE(src && dst && rows > 0 && cols > 0);
  CASPR_COUNT(); transpose2d_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      src, rows, cols, dst);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
