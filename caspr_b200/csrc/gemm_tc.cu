// caspr_linear_tc: the 1x1-conv / Linear layers of the TPointNet++ encoder on the tcgen05 tensor
// cores with fp32-grade accuracy (three fp16 products, fp32 accumulation in TMEM; tc_gemm.cuh).
// Replaces the large torch Conv1d(k=1) calls of caspr/models/tpointnet2.py:99-105 and
// pointnet2.py:471-481,207-212 (feature-propagation and final layers); the small per-ball layers
// stay on the exact-fp32 SIMT kernel (dense.cu).
//
// Per call: (1) split X and W into fp16 hi/lo planes, every row scaled by its own power of two that
// places the row's largest element in the upper fp16 range, K padded to 64, rows / channels padded to
// the tile (zero fill); (2) persistent TMA + tcgen05 GEMM; the epilogue undoes the row and channel
// scales exactly, adds the bias, applies the activation and stores fp32 rows with the caller's
// leading dimension.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

using tcg::kBK;
using tcg::kBM;
using tcg::kBN;

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == CASPR_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CASPR_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// GroupNorm folded into the operand split of the NEXT layer: x <- ReLU?((x - mean)*rstd*gamma + beta) with
// (mean, rstd) per (sample, group) from `tab` (written by gn_table_kernel from the statistics the previous
// GEMM's epilogue accumulated).  Saves the separate normalisation pass over the activation tensor.
struct NormFold {
  const float2* tab;     // [samples][C] (scale, shift): x <- x*scale + shift; nullptr = no folding
  int rows_per_sample, C, relu;
};

__device__ __forceinline__ float fold_apply(const NormFold& nf, const float2* tab_row, float v, int c) {
  const float2 ss = __ldg(tab_row + c);
  v = fmaf(v, ss.x, ss.y);
  return nf.relu ? fmaxf(v, 0.f) : v;
}

// (sum, sum of squares) in fp64 per (sample, group) -> per (sample, channel) scale = rstd*gamma and
// shift = beta - mean*rstd*gamma in fp32
__global__ void gn_table_kernel(const double* __restrict__ stats, int samples, int groups, int C, double count,
                                float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float2* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= samples * C) return;
  const int sample = i / C, c = i - sample * C;
  const int g = c / (C / groups);
  const double* st = stats + ((size_t)sample * groups + g) * 2;
  const double mean = st[0] / count;
  double var = st[1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double sc = rstd * (double)gamma[c];
  tab[i] = make_float2((float)sc, (float)((double)beta[c] - mean * sc));
}

// fp32 [rows][cols] (leading dimension ld) -> fp16 hi / lo planes [rows_pad][k_pad], zero padded, each
// row scaled by its own power of two 2^(14-e) (max|row| < 2^e, so the largest element lands in
// [2^13, 2^14)); inv_scale[row] = 2^(e-14) is undone exactly in the GEMM epilogue.  One warp per row:
// the row is read twice, the second time from L1/L2.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, int ld, long long rows, int cols, long long rows_pad, int k_pad,
                  int relu, __half2* __restrict__ hi, __half2* __restrict__ lo, float* __restrict__ inv_scale,
                  NormFold nf = NormFold()) {
  const int lane = threadIdx.x & 31;
  const int kp2 = k_pad / 2;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_pad; r += warps) {
    __half2* h = hi + r * kp2;
    __half2* l = lo + r * kp2;
    if (r >= rows) {
      const __half2 z = __floats2half2_rn(0.f, 0.f);
      for (int c2 = lane; c2 < kp2; c2 += 32) { h[c2] = z; l[c2] = z; }
      if (lane == 0) inv_scale[r] = 0.f;
      continue;
    }
    const float* xr = x + r * ld;
    const float2* tab_row = nf.tab ? nf.tab + (r / nf.rows_per_sample) * nf.C : nullptr;
    float m = 0.f;
    for (int c = lane; c < cols; c += 32) {
      float v = xr[c];
      if (tab_row) v = fold_apply(nf, tab_row, v, c);
      if (relu) v = fmaxf(v, 0.f);
      m = fmaxf(m, fabsf(v));
    }
    m = warp_max(m);
    float s = 1.f, inv = 1.f;
    if (m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);
      s = ldexpf(1.f, 14 - e);
      inv = ldexpf(1.f, e - 14);
    }
    if (lane == 0) inv_scale[r] = inv;
    for (int c2 = lane; c2 < kp2; c2 += 32) {
      const int c = 2 * c2;
      float a = c < cols ? xr[c] : 0.f;
      float b = c + 1 < cols ? xr[c + 1] : 0.f;
      if (tab_row) {
        if (c < cols) a = fold_apply(nf, tab_row, a, c);
        if (c + 1 < cols) b = fold_apply(nf, tab_row, b, c + 1);
      }
      if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      a *= s;
      b *= s;
      const __half2 hh = __floats2half2_rn(a, b);
      const float2 hf = __half22float2(hh);
      h[c2] = hh;
      l[c2] = __floats2half2_rn(a - hf.x, b - hf.y);
    }
  }
}

// Single-read variant of split_rows_kernel for rows of at most KCH*128 columns with 16-byte aligned rows: a warp keeps
// the whole (transformed) row in registers as KCH float4 per lane, so the row is read from memory ONCE with all its
// loads in flight together, and the planes are written as 8-byte pieces (256 contiguous bytes per warp and plane).
template <int KCH>
__global__ void __launch_bounds__(256)
split_rows_reg_kernel(const float* __restrict__ x, int ld, long long rows, int cols, long long rows_pad, int k_pad,
                      int relu, __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv_scale,
                      NormFold nf) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_pad; r += warps) {
    uint2* h = reinterpret_cast<uint2*>(hi + r * k_pad);
    uint2* l = reinterpret_cast<uint2*>(lo + r * k_pad);
    if (r >= rows) {
#pragma unroll
      for (int j = 0; j < KCH; ++j) {
        const int c = 4 * (lane + 32 * j);
        if (c < k_pad) { h[c / 4] = make_uint2(0u, 0u); l[c / 4] = make_uint2(0u, 0u); }
      }
      if (lane == 0) inv_scale[r] = 0.f;
      continue;
    }
    const float* xr = x + r * ld;
    const float2* tab_row = nf.tab ? nf.tab + (r / nf.rows_per_sample) * nf.C : nullptr;
    float v[KCH][4];
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = 4 * (lane + 32 * j);
      if (c + 3 < cols) {
        const float4 f = *reinterpret_cast<const float4*>(xr + c);
        v[j][0] = f.x; v[j][1] = f.y; v[j][2] = f.z; v[j][3] = f.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[j][u] = c + u < cols ? xr[c + u] : 0.f;
      }
    }
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = 4 * (lane + 32 * j);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a = v[j][u];
        if (c + u < cols) {
          if (tab_row) a = fold_apply(nf, tab_row, a, c + u);
          if (relu) a = fmaxf(a, 0.f);
        } else {
          a = 0.f;
        }
        v[j][u] = a;
        m = fmaxf(m, fabsf(a));
      }
    }
    m = warp_max(m);
    float s = 1.f, inv = 1.f;
    if (m > 0.f && m < 3.0e38f) {
      int e;
      frexpf(m, &e);
      s = ldexpf(1.f, 14 - e);
      inv = ldexpf(1.f, e - 14);
    }
    if (lane == 0) inv_scale[r] = inv;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int c = 4 * (lane + 32 * j);
      if (c >= k_pad) continue;
      uint32_t h0, l0, h1, l1;
      tcg::split2(v[j][0] * s, v[j][1] * s, h0, l0);
      tcg::split2(v[j][2] * s, v[j][3] * s, h1, l1);
      h[c / 4] = make_uint2(h0, h1);
      l[c / 4] = make_uint2(l0, l1);
    }
  }
}

// dispatch: register-resident rows when they fit and are aligned, the generic two-pass kernel otherwise
void launch_split_rows(const float* x, int ld, long long rows, int cols, long long rows_pad, int k_pad, int relu,
                       __half* hi, __half* lo, float* inv_scale, const NormFold& nf, int blocks, cudaStream_t s) {
  const bool aligned = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (k_pad % 4 == 0);
  const int kch = (cols + 127) / 128;
#define CASPR_SPLIT_REG(K)                                                                                     \
  split_rows_reg_kernel<K><<<blocks, 256, 0, s>>>(x, ld, rows, cols, rows_pad, k_pad, relu, hi, lo, inv_scale, nf)
  if (aligned && kch <= 1) CASPR_SPLIT_REG(1);
  else if (aligned && kch <= 2) CASPR_SPLIT_REG(2);
  else if (aligned && kch <= 4) CASPR_SPLIT_REG(4);
  else if (aligned && kch <= 8) CASPR_SPLIT_REG(8);
  else if (aligned && kch <= 13) CASPR_SPLIT_REG(13);
  else
    split_rows_kernel<<<blocks, 256, 0, s>>>(x, ld, rows, cols, rows_pad, k_pad, relu, (__half2*)hi, (__half2*)lo,
                                             inv_scale, nf);
#undef CASPR_SPLIT_REG
}

struct LinearEpilogue {
  const float* bias;
  int bias_rps;           // 0: one bias vector; > 0: bias is (samples, cout), row r uses bias row r / bias_rps
  const float* x_inv;     // per-row 1/scale of X
  const float* w_inv;     // per-output-channel 1/scale of W
  float* Y;
  int ldy;
  int rows, cout, act_out;
  int vec_ok;
  // optional GroupNorm statistics of the OUTPUT (before any normalisation): per (sample, group) sum and sum
  // of squares accumulated in fp64; requires rows_per_sample % 32 == 0 so a warp never straddles samples
  double* stats;
  int st_rows_per_sample, st_cpg, st_groups;
  // per-thread tile state
  long long row;
  int col0;
  float inv;

  __device__ __forceinline__ void setup(uint8_t*, const CUtensorMap*, const CUtensorMap*, int) {}
  __device__ __forceinline__ void tile_begin(int m_tile, int n_tile, int q, int lane) {
    if (stats && st_live) flush_stats();               // statistics of the previous tile
    row = (long long)m_tile * kBM + q * 32 + lane;
    col0 = n_tile * kBN;
    inv = x_inv[row];       // planes are padded to whole tiles, so the index is always valid
    bias_row = bias;
    if (bias && bias_rps > 0) bias_row = bias + ((row < rows ? row : rows - 1) / bias_rps) * (long long)cout;
    if (stats) {
      const long long warp_row = row - lane;
      st_live = warp_row < rows;
      st_row = stats + ((st_live ? warp_row : 0) / st_rows_per_sample) * st_groups * 2;
      st_g = col0 / st_cpg;
      st_next = (st_g + 1) * st_cpg;
      st_s = 0.f; st_q = 0.f;
    }
  }
  // Code size matters here: the epilogue warps of a GEMM with a short k-loop are instruction-fetch bound if this body
  // is large (the first version inlined the statistics flush 32 times: 8.7 k instructions, "no instruction" was the
  // top stall reason and a 128 x 256 tile took 36 us).  One vector path (cout % 4 == 0 is required), statistics with
  // at most one group boundary per chunk (channels per group >= 32 is required), activation outside the main loop.
  __device__ __forceinline__ void chunk(int chunk, uint32_t (&r)[32]) {
    const int c = col0 + chunk * 32;
    if (c >= cout) return;                              // warp-uniform
    const bool row_ok = row < rows;
    float v[32];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const bool ok = c + 4 * j4 < cout;                // warp-uniform; cout % 4 == 0
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        w4 = *reinterpret_cast<const float4*>(w_inv + c + j4 * 4);
        if (bias) b4 = *reinterpret_cast<const float4*>(bias_row + c + j4 * 4);
      }
      v[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), inv * w4.x, b4.x);
      v[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), inv * w4.y, b4.y);
      v[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), inv * w4.z, b4.z);
      v[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), inv * w4.w, b4.w);
    }
    if (stats && st_live) {
      // running (sum, sum of squares) of the current channel group; a group has >= 32 channels, so at most one
      // boundary falls into this chunk: columns before it extend the running group, the others start the next one.
      // Padded columns (>= cout) hold 0 (zero weight rows, w_inv = 0) and padded rows are masked.
      float s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = row_ok ? v[j] : 0.f;
        if (c + j >= st_next) { s1 += x; q1 = fmaf(x, x, q1); }
        else { st_s += x; st_q = fmaf(x, x, st_q); }
      }
      if (c + 32 >= st_next) {                          // warp-uniform: the running group ends inside this chunk
        flush_stats();
        st_s = s1; st_q = q1; ++st_g; st_next += st_cpg;
      }
    }
    if (!row_ok) return;
    if (act_out == CASPR_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    float* y = Y + row * ldy + c;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      if (c + 4 * j4 < cout)
        reinterpret_cast<float4*>(y)[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
  }
  const float* bias_row;
  // per-thread state of the running group
  double* st_row;
  float st_s, st_q;
  int st_g, st_next;
  bool st_live;
  __device__ __forceinline__ void flush_stats() {
    const float s = warp_sum(st_s), q = warp_sum(st_q);
    if ((threadIdx.x & 31) == 0 && st_g < st_groups) {
      atomicAdd(st_row + 2 * st_g, (double)s);
      atomicAdd(st_row + 2 * st_g + 1, (double)q);
    }
  }
  __device__ __forceinline__ void finish() {
    if (stats && st_live) flush_stats();
  }
};

struct Layout {
  long long rows_pad;
  int k_pad, cout_pad;
  size_t off_xhi, off_xlo, off_xinv, off_w, total;     // off_w: start of an embedded WeightLayout block
};

// Split weights (reusable across calls while the weights do not change): [w_inv | W_hi | W_lo]
struct WeightLayout {
  int k_pad, cout_pad;
  size_t off_winv, off_whi, off_wlo, total;
};

WeightLayout make_weight_layout(int cin, int cout) {
  WeightLayout l;
  l.k_pad = (cin + kBK - 1) / kBK * kBK;
  l.cout_pad = (cout + kBN - 1) / kBN * kBN;
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 1024); return r; };
  l.off_winv = take((size_t)l.cout_pad * 4);
  l.off_whi = take((size_t)l.cout_pad * l.k_pad * 2);
  l.off_wlo = take((size_t)l.cout_pad * l.k_pad * 2);
  l.total = p;
  return l;
}

Layout make_layout(long long rows, int cin, int cout) {
  Layout l;
  l.rows_pad = (rows + kBM - 1) / kBM * kBM;
  l.k_pad = (cin + kBK - 1) / kBK * kBK;
  l.cout_pad = (cout + kBN - 1) / kBN * kBN;
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 1024); return r; };
  l.off_xinv = take((size_t)l.rows_pad * 4);
  l.off_xhi = take((size_t)l.rows_pad * l.k_pad * 2);
  l.off_xlo = take((size_t)l.rows_pad * l.k_pad * 2);
  l.off_w = take(make_weight_layout(cin, cout).total);
  l.total = p;
  return l;
}


// ------------------------------------------------------------------ weight gradient (split-K, transposed operands)
// dW[o][i] = sum_r dY[r][o] X[r][i]: both operands are needed with the ROW index as the contraction (K) dimension, i.e.
// transposed.  colmax -> per-column power-of-two scale; transpose_split writes the fp16 hi/lo planes [C_pad][R_pad].
// Column maxima are kept as the bit patterns of non-negative floats (their unsigned order is the float order), so
// every producer can fold its part in with atomicMax; power-of-two scale 2^(14-e) for max|column| < 2^e.
__device__ __forceinline__ void scale_from_max_bits(unsigned bits, float& scale, float& inv) {
  const float m = __uint_as_float(bits);
  scale = 1.f;
  inv = 1.f;
  if (m > 0.f && m < 3.0e38f) {
    int e;
    frexpf(m, &e);
    scale = ldexpf(1.f, 14 - e);
    inv = ldexpf(1.f, e - 14);
  }
}

__global__ void __launch_bounds__(128)
colmax_kernel(const float* __restrict__ X, int ldx, long long rows, int C, long long rows_per_split, int relu,
              unsigned* __restrict__ cmax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long r0 = blockIdx.y * rows_per_split;
  const long long r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  // eight independent loads in flight per thread; with relu the negative values count as 0
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = 0.f;
  long long r = r0;
  for (; r + 7 < r1; r += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = X[(r + j) * ldx + c];
      m[j] = fmaxf(m[j], relu ? v : fabsf(v));
    }
  }
  for (; r < r1; ++r) {
    const float v = X[r * ldx + c];
    m[0] = fmaxf(m[0], relu ? v : fabsf(v));
  }
  const float mm = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
  atomic_max_nonneg(cmax + c, mm);
}
// planes[c][r] = split(X[r][c] * scale[c]) for a 64 (rows) x 64 (channels) tile through shared memory; zero padding
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ X, int ldx, long long rows, int C, long long rows_pad, int relu,
                       const unsigned* __restrict__ cmax, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  // 16 independent loads per thread in flight (rows rr0 + 4j, channel cc), then the shared-memory transpose
  const int cc = threadIdx.x & 63, rr0 = threadIdx.x >> 6;
  const int c = c0 + cc;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const long long r = r0 + rr0 + 4 * j;
    v[j] = (r < rows && c < C) ? __ldg(X + r * ldx + c) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) tile[rr0 + 4 * j][cc] = relu ? fmaxf(v[j], 0.f) : v[j];
  __syncthreads();
  // thread -> (row pair rp, channels ch0 + 8q): 64 channels x 32 pairs = 2048 half2 per plane, 8 per thread
  const int rp = threadIdx.x & 31, ch0 = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int ch = ch0 + 8 * q;
    float s, inv_unused;
    scale_from_max_bits(c0 + ch < C ? cmax[c0 + ch] : 0u, s, inv_unused);
    const float a = tile[2 * rp][ch] * s, b = tile[2 * rp + 1][ch] * s;
    const __half2 hh = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hh);
    const size_t off = ((size_t)(c0 + ch) * rows_pad + r0) / 2 + rp;
    reinterpret_cast<__half2*>(hi)[off] = hh;
    reinterpret_cast<__half2*>(lo)[off] = __floats2half2_rn(a - hf.x, b - hf.y);
  }
}

// partial[split][o][i] = acc * inv_dy[o] * inv_x[i]
struct WgradEpilogue {
  const unsigned* a_max;  // column maxima (bit patterns) of dY: per output channel o (rows of the A operand)
  const unsigned* w_max;  // column maxima of X: per input channel i
  float* part;            // [k_splits][cout][cin]
  int cout, cin, m_tiles;
  long long row;
  int col0, split;
  float inv;
  __device__ __forceinline__ void setup(uint8_t*, const CUtensorMap*, const CUtensorMap*, int) {}
  __device__ __forceinline__ void tile_begin(int m_tile_epi, int n_tile, int q, int lane) {
    split = m_tile_epi / m_tiles;
    row = (long long)(m_tile_epi - split * m_tiles) * kBM + q * 32 + lane;
    col0 = n_tile * kBN;
    float s_unused;
    scale_from_max_bits(row < cout ? a_max[row] : 0u, s_unused, inv);
  }
  __device__ __forceinline__ void chunk(int chunk, uint32_t (&r)[32]) {
    const int c = col0 + chunk * 32;
    if (c >= cin || row >= cout) return;
    float* y = part + ((size_t)split * cout + row) * cin + c;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < cin) {
        float s_unused, winv;
        scale_from_max_bits(w_max[c + j], s_unused, winv);
        y[j] = __uint_as_float(r[j]) * (inv * winv);
      }
  }
  __device__ __forceinline__ void finish() {}
};

struct WgradLayout {
  long long r_pad;          // contraction length padded to k_splits * k_chunks * 64
  int k_splits, k_chunks, cout_pad, cin_pad, colmax_splits;
  size_t off_amax, off_wmax, off_ahi, off_alo, off_whi, off_wlo, off_part, total;
};
WgradLayout make_wgrad_layout(long long rows, int cout, int cin) {
  WgradLayout l;
  l.cout_pad = (cout + kBM - 1) / kBM * kBM;
  l.cin_pad = (cin + kBN - 1) / kBN * kBN;
  const long long chunks_total = (rows + kBK - 1) / kBK;
  // at most 64 k-chunks (4096 rows) per split: bounds the truncation error of the fp32 accumulation in TMEM and
  // gives enough work items to fill the GPU
  long long splits = (chunks_total + 63) / 64;
  const long long tiles = (long long)(l.cout_pad / kBM) * (l.cin_pad / kBN);
  while (splits * tiles < 148 && splits < chunks_total) ++splits;
  {
    // fill whole waves of 148 CTAs: as many splits as fit into the number of waves the minimum needs
    const long long waves = (splits * tiles + 147) / 148;
    long long cand = waves * 148 / tiles;
    if (cand > chunks_total) cand = chunks_total;
    if (cand > splits) splits = cand;
  }
  l.k_splits = (int)splits;
  l.k_chunks = (int)((chunks_total + splits - 1) / splits);
  l.r_pad = (long long)l.k_splits * l.k_chunks * kBK;
  long long cs = (8 * 148) / ((cout > cin ? cout : cin) / 128 + 1);
  const long long maxs = (rows + 127) / 128;
  if (cs > maxs) cs = maxs;
  if (cs < 1) cs = 1;
  l.colmax_splits = (int)cs;
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 1024); return r; };
  l.off_amax = take((size_t)l.cout_pad * 4);
  l.off_wmax = take((size_t)l.cin_pad * 4);
  l.off_ahi = take((size_t)l.cout_pad * l.r_pad * 2);
  l.off_alo = take((size_t)l.cout_pad * l.r_pad * 2);
  l.off_whi = take((size_t)l.cin_pad * l.r_pad * 2);
  l.off_wlo = take((size_t)l.cin_pad * l.r_pad * 2);
  l.off_part = take((size_t)l.k_splits * cout * cin * 4);
  l.total = p;
  return l;
}

__global__ void __launch_bounds__(256)
wgrad_sum_parts_kernel(const float* __restrict__ part, int nparts, size_t nelem, float* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * nelem + i];
  dst[i] = s;
}

}  // namespace

extern "C" size_t caspr_linear_wgrad_tc_workspace_bytes(long long rows, int Cout, int Cin) {
  if (rows <= 0 || Cout <= 0 || Cin <= 0) return 0;
  return make_wgrad_layout(rows, Cout, Cin).total;
}

extern "C" int caspr_linear_wgrad_tc(const float* dY, int lddy, const float* X, int ldx, long long rows, int Cout,
                                     int Cin, int relu_x, float* dW, const uint32_t* dy_colmax,
                                     const uint32_t* x_colmax, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  CASPR_REQUIRE(dY && X && dW && workspace && rows > 0 && Cout > 0 && Cin > 0 && lddy >= Cout && ldx >= Cin);
  CASPR_REQUIRE(((uintptr_t)workspace & 1023) == 0);
  const WgradLayout l = make_wgrad_layout(rows, Cout, Cin);
  if (workspace_bytes < l.total) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  unsigned* amax = (unsigned*)(base + l.off_amax);
  unsigned* wmax = (unsigned*)(base + l.off_wmax);
  __half* ahi = (__half*)(base + l.off_ahi);
  __half* alo = (__half*)(base + l.off_alo);
  __half* whi = (__half*)(base + l.off_whi);
  __half* wlo = (__half*)(base + l.off_wlo);
  float* part = (float*)(base + l.off_part);
  const long long rps = (rows + l.colmax_splits - 1) / l.colmax_splits;
  CASPR_REQUIRE(l.r_pad / 64 < (1ll << 31));
  // column maxima: supplied by the producers of the operands, or one pass here
  if (!dy_colmax) {
    if (cudaMemsetAsync(amax, 0, (size_t)l.cout_pad * 4, s) != cudaSuccess) return CASPR_ELAUNCH;
    CASPR_COUNT(); colmax_kernel<<<dim3(ceil_div(Cout, 128), l.colmax_splits), 128, 0, s>>>(dY, lddy, rows, Cout, rps, 0, amax);
    dy_colmax = amax;
  }
  if (!x_colmax) {
    if (cudaMemsetAsync(wmax, 0, (size_t)l.cin_pad * 4, s) != cudaSuccess) return CASPR_ELAUNCH;
    CASPR_COUNT(); colmax_kernel<<<dim3(ceil_div(Cin, 128), l.colmax_splits), 128, 0, s>>>(X, ldx, rows, Cin, rps, relu_x, wmax);
    x_colmax = wmax;
  }
  // operand A = dY^T, operand W = X^T
  CASPR_COUNT(); transpose_split_kernel<<<dim3((unsigned)(l.r_pad / 64), l.cout_pad / 64), 256, 0, s>>>(
      dY, lddy, rows, Cout, l.r_pad, 0, dy_colmax, ahi, alo);
  CASPR_COUNT(); transpose_split_kernel<<<dim3((unsigned)(l.r_pad / 64), l.cin_pad / 64), 256, 0, s>>>(
      X, ldx, rows, Cin, l.r_pad, relu_x, x_colmax, whi, wlo);
  CASPR_CHECK_LAUNCH();
  CUtensorMap tm_ahi, tm_alo, tm_whi, tm_wlo;
  bool ok = true;
  ok &= caspr_make_tmap_f16(&tm_ahi, ahi, (uint64_t)l.cout_pad, (uint64_t)l.r_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_alo, alo, (uint64_t)l.cout_pad, (uint64_t)l.r_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_whi, whi, (uint64_t)l.cin_pad, (uint64_t)l.r_pad, tcg::w_box_rows());
  ok &= caspr_make_tmap_f16(&tm_wlo, wlo, (uint64_t)l.cin_pad, (uint64_t)l.r_pad, tcg::w_box_rows());
  if (!ok) return CASPR_ELAUNCH;
  int dev = 0, num_sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return CASPR_ELAUNCH;
  const int m_tiles = l.cout_pad / kBM, n_tiles = l.cin_pad / kBN;
  WgradEpilogue epi{};
  epi.a_max = dy_colmax; epi.w_max = x_colmax; epi.part = part; epi.cout = Cout; epi.cin = Cin; epi.m_tiles = m_tiles;
  CASPR_COUNT();
  if (tcg::launch_gemm(tm_ahi, tm_alo, tm_whi, tm_wlo, tm_ahi, tm_alo, 0, m_tiles, n_tiles, l.k_chunks, nullptr, epi,
                       l.k_splits, num_sms, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  const size_t nelem = (size_t)Cout * Cin;
  CASPR_COUNT(); wgrad_sum_parts_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, s>>>(part, l.k_splits, nelem, dW);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

namespace {
}  // namespace

extern "C" size_t caspr_linear_tc_workspace_bytes(int rows, int Cin, int Cout) {
  if (rows <= 0 || Cin <= 0 || Cout <= 0) return 0;
  return make_layout(rows, Cin, Cout).total;
}

extern "C" int caspr_gn_table(const double* stats, int samples, int groups, int rows_per_sample, int C, float eps,
                              const float* gamma, const float* beta, float* table, void* stream) {
  CASPR_REQUIRE(stats && table && gamma && beta && samples > 0 && groups > 0 && rows_per_sample > 0 && C % groups == 0);
  const int n = samples * C;
  const double count = (double)(C / groups) * (double)rows_per_sample;
  CASPR_COUNT(); gn_table_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(stats, samples, groups, C, count, eps,
                                                                                   gamma, beta, (float2*)table);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" size_t caspr_linear_tc_weight_bytes(int Cin, int Cout) {
  if (Cin <= 0 || Cout <= 0) return 0;
  return make_weight_layout(Cin, Cout).total;
}

extern "C" int caspr_linear_tc_prepare_weights(const float* W, int ldw, int Cin, int Cout, void* prepared,
                                               size_t prepared_bytes, void* stream) {
  CASPR_REQUIRE(W && prepared && Cin > 0 && Cout > 0 && ldw >= Cin && ((uintptr_t)prepared & 1023) == 0);
  const WeightLayout wl = make_weight_layout(Cin, Cout);
  if (prepared_bytes < wl.total) return CASPR_EWORKSPACE;
  char* base = (char*)prepared;
  CASPR_COUNT(); split_rows_kernel<<<ceil_div(wl.cout_pad, 8), 256, 0, (cudaStream_t)stream>>>(
      W, ldw, Cout, Cin, wl.cout_pad, wl.k_pad, 0, (__half2*)(base + wl.off_whi), (__half2*)(base + wl.off_wlo),
      (float*)(base + wl.off_winv));
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_linear_tc(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y,
                               int ldy, int rows, int Cin, int Cout, int act_in, int act_out,
                               const void* prepared_weights, const caspr_gn_fold* in_norm,
                               const caspr_gn_stats* out_stats, int bias_rows_per_sample, void* workspace,
                               size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(X && (W || prepared_weights) && Y && workspace && rows > 0 && Cin > 0 && Cout > 0);
  CASPR_REQUIRE(bias_rows_per_sample >= 0 && (bias_rows_per_sample == 0 || bias));
  // the epilogue stores float4 pieces and keeps one running GroupNorm group per chunk of 32 columns
  CASPR_REQUIRE(Cout % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)Y & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0));
  if (out_stats) CASPR_REQUIRE(Cout / out_stats->groups >= 32);
  CASPR_REQUIRE(act_out == CASPR_ACT_NONE || act_out == CASPR_ACT_RELU);     // sigmoid outputs: caspr_linear
  if (in_norm)
    CASPR_REQUIRE(in_norm->table && in_norm->rows_per_sample > 0);
  if (out_stats)
    CASPR_REQUIRE(out_stats->stats && out_stats->groups > 0 && Cout % out_stats->groups == 0 &&
                  out_stats->rows_per_sample > 0 && out_stats->rows_per_sample % 32 == 0 &&
                  rows % out_stats->rows_per_sample == 0);
  CASPR_REQUIRE(ldx >= Cin && (!W || ldw >= Cin) && ldy >= Cout);
  CASPR_REQUIRE(act_in == CASPR_ACT_NONE || act_in == CASPR_ACT_RELU);
  CASPR_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)prepared_weights & 1023) == 0);
  const Layout l = make_layout(rows, Cin, Cout);
  const WeightLayout wl = make_weight_layout(Cin, Cout);
  if (workspace_bytes < l.total) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  float* xinv = (float*)(base + l.off_xinv);
  __half* xhi = (__half*)(base + l.off_xhi);
  __half* xlo = (__half*)(base + l.off_xlo);
  const char* wbase = (const char*)prepared_weights;
  if (!wbase) {
    int rc = caspr_linear_tc_prepare_weights(W, ldw, Cin, Cout, base + l.off_w, wl.total, stream);
    if (rc) return rc;
    wbase = base + l.off_w;
  }
  const float* winv = (const float*)(wbase + wl.off_winv);
  const __half* whi = (const __half*)(wbase + wl.off_whi);
  const __half* wlo = (const __half*)(wbase + wl.off_wlo);

  const int nb = 148 * 8;
  NormFold nf = NormFold();
  if (in_norm) {
    nf.tab = (const float2*)in_norm->table; nf.rows_per_sample = in_norm->rows_per_sample; nf.C = Cin;
    nf.relu = in_norm->relu;
  }
  CASPR_COUNT(); launch_split_rows(X, ldx, rows, Cin, l.rows_pad, l.k_pad, act_in == CASPR_ACT_RELU, xhi, xlo, xinv, nf, nb, s);
  CASPR_CHECK_LAUNCH();
  if (out_stats) {
    const size_t n_stats = (size_t)(rows / out_stats->rows_per_sample) * out_stats->groups * 2;
    if (cudaMemsetAsync(out_stats->stats, 0, n_stats * sizeof(double), s) != cudaSuccess) return CASPR_ELAUNCH;
  }

  CUtensorMap tm_xhi, tm_xlo, tm_whi, tm_wlo;
  bool ok = true;
  ok &= caspr_make_tmap_f16(&tm_xhi, xhi, (uint64_t)l.rows_pad, (uint64_t)l.k_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_xlo, xlo, (uint64_t)l.rows_pad, (uint64_t)l.k_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_whi, whi, (uint64_t)l.cout_pad, (uint64_t)l.k_pad, tcg::w_box_rows());
  ok &= caspr_make_tmap_f16(&tm_wlo, wlo, (uint64_t)l.cout_pad, (uint64_t)l.k_pad, tcg::w_box_rows());
  if (!ok) return CASPR_ELAUNCH;
  int dev = 0, num_sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return CASPR_ELAUNCH;
  LinearEpilogue epi{};
  epi.bias = bias; epi.bias_rps = bias_rows_per_sample; epi.x_inv = xinv; epi.w_inv = winv; epi.Y = Y; epi.ldy = ldy; epi.rows = rows; epi.cout = Cout;
  epi.act_out = act_out;
  if (out_stats) {
    epi.stats = out_stats->stats; epi.st_rows_per_sample = out_stats->rows_per_sample;
    epi.st_groups = out_stats->groups; epi.st_cpg = Cout / out_stats->groups;
  }
  epi.vec_ok = (ldy % 4 == 0) && (((uintptr_t)Y & 15) == 0) && (!bias || ((uintptr_t)bias & 15) == 0);
  const int m_tiles = (int)(l.rows_pad / kBM), n_tiles = l.cout_pad / kBN;
  caspr_prof_begin(CASPR_PROF_LINEAR, s);
  CASPR_COUNT();
  const cudaError_t lerr = tcg::launch_gemm(tm_xhi, tm_xlo, tm_whi, tm_wlo, tm_xhi, tm_xlo, 0, m_tiles, n_tiles,
                                            l.k_pad / kBK, nullptr, epi, 1, num_sms, s);
  caspr_prof_end(CASPR_PROF_LINEAR, s);
  if (lerr != cudaSuccess) return CASPR_ELAUNCH;
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
