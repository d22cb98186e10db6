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
#include <algorithm>
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
  static constexpr int kEpilogueGroups = 2;       // each group stages through its own 16 KB half (one buffer)
  const float* bias;
  int bias_rps;           // 0: one bias vector; > 0: bias is (samples, cout), row r uses bias row r / bias_rps
  const float* x_inv;     // per-row 1/scale of X
  const float* w_inv;     // per-output-channel 1/scale of W
  float* Y;
  int ldy;
  int rows, cout, act_out;
  int vec_ok;
  // use_tma: the fp32 rows leave through shared memory and TMA (full 128-byte lines; the direct path issues eight
  // 16-byte stores per thread and chunk to 32 different lines each, which made the epilogue LSU-bound)
  int use_tma;
  uint8_t* stg;
  const CUtensorMap* tm_y;
  int etid, m_tile_cur, box_row, bar_id;
  // optional GroupNorm statistics of the OUTPUT (before any normalisation): per (sample, group) sum and sum
  // of squares accumulated in fp64; requires rows_per_sample % 32 == 0 so a warp never straddles samples
  double* stats;
  int st_rows_per_sample, st_cpg, st_groups;
  // optional per-(sample, channel) max / min of the output as ordered-uint keys ([samples][2][cout]); Y is then not
  // written (GroupNorm + max-pool consumers need nothing else, caspr_gn_max_from_extrema)
  unsigned* ext;
  // per-thread tile state
  long long row;
  int col0;
  float inv;

  __device__ __forceinline__ void setup(uint8_t* staging, const CUtensorMap* o_hi, const CUtensorMap*, int epi_tid) {
    stg = staging; tm_y = o_hi; etid = epi_tid;
    bar_id = 1 + ((((int)threadIdx.x >> 5) - 2) >> 2);           // named barrier of this epilogue warpgroup
  }
  __device__ __forceinline__ void tile_begin(int m_tile, int n_tile, int q, int lane) {
    if (stats && st_live) flush_stats();               // statistics of the previous tile
    m_tile_cur = m_tile; box_row = q * 32 + lane;
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
    if (ext) {
      // column extrema over the warp's 32 rows (one sample: rows_per_sample % 32 == 0), one redux per column and bound;
      // lane j keeps column j and issues the two atomics
      unsigned my_mx = 0u, my_mn = 0xffffffffu;
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const unsigned key = float_to_ordered(v[j]);
        const unsigned mx = __reduce_max_sync(0xffffffffu, row_ok ? key : 0u);
        const unsigned mn = __reduce_min_sync(0xffffffffu, row_ok ? key : 0xffffffffu);
        if (lane == j) { my_mx = mx; my_mn = mn; }
      }
      if (st_live && c + lane < cout) {
        unsigned* e = ext + (size_t)((row - lane) / st_rows_per_sample) * 2 * cout + c + lane;
        // the bounds only move outwards: a plain read that already shows a wider bound makes the atomic unnecessary,
        // which is the case for almost every tile after the first few (hundreds of tiles share one address)
        if (my_mx > __ldcg(e)) atomicMax(e, my_mx);
        if (my_mn < __ldcg(e + cout)) atomicMin(e + cout, my_mn);
      }
      return;
    }
    if (act_out == CASPR_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (use_tma) {
      // one staging buffer of [128 rows][128 B] per warpgroup (128-byte swizzle); rows / columns beyond the matrix are
      // clipped by TMA.  The other warpgroup works on the other accumulator meanwhile.
      uint8_t* buf = stg;
      if (etid == 0) tc::tma_store_wait_read();                    // the previous store has left `buf`
      tc::named_bar_sync(bar_id, 128);
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(buf + box_row * 128 + ((j4 ^ (box_row & 7)) << 4)) =
            make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      tc::fence_proxy_async_smem();
      tc::named_bar_sync(bar_id, 128);
      if (etid == 0) {
        tc::tma_store_2d(tm_y, buf, c, m_tile_cur * kBM);
        tc::tma_store_commit();
      }
      return;
    }
    if (!row_ok) return;
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
    if (use_tma && etid == 0) tc::tma_store_wait_all();
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
  static constexpr int kEpilogueGroups = 2;
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

// ---------------------------------------------------------------------------------------------------------------
// Per-ball MLP of a set-abstraction scale on the tensor cores (pointnet2.py:649-708 as used by SA levels 3-5):
//   [Conv1d(k=1) -> GroupNorm(16) over each BALL of ns rows -> ReLU] x 2 -> Conv1d -> GroupNorm -> max over the ball.
// A 128-row accumulator tile holds whole balls (ns = 16 or 32 consecutive rows) and a GroupNorm group is CPG = C/16
// consecutive channels, so the epilogue of each GEMM normalises in registers (two passes over the thread's values,
// ball reductions by warp shuffles) and emits the NEXT layer's fp16 hi / lo operand planes directly (fixed power-of-
// two scale: normalised activations are bounded) or, for the last layer, the max over the ball.  Replaces, per layer,
// the sequence  GEMM (fp32 rows out) -> per-ball GroupNorm kernel (read + write) -> operand split (read + write).
constexpr float kBallPlaneScale = 512.f;      // |GroupNorm output| <= |gamma| sqrt(ns*CPG) + |beta| << 65504 / 512

template <int CPG, bool LAST>
struct BallNormEpilogue {
  static constexpr int kEpilogueGroups = 2;
  static constexpr bool kReadsTmem = true;
  static constexpr int kChunk = (CPG == 6) ? 24 : 32;           // whole groups, a multiple of 8 columns (C = 96: 24)
  static_assert(kChunk % CPG == 0 && kChunk % 8 == 0, "a chunk holds whole GroupNorm groups");
  static constexpr int kGroups = kChunk / CPG;
  const float* bias;
  const float* x_inv;      // per-row 1/scale of the A planes, or nullptr: inv_const for every row
  float inv_const;
  const float* w_inv;
  const float* gamma;
  const float* beta;
  float eps;
  int ns;
  long long rows;
  int cout;
  __half* out_hi;          // !LAST: operand planes of the next layer [rows_pad][ld_out]
  __half* out_lo;
  int ld_out;
  float* maxout;           // LAST: (balls, cout) with row stride ld_max
  int ld_max;
  int* range_flag;
  // per-thread tile state
  long long row;
  int col0, lane;
  float inv;

  __device__ __forceinline__ void setup(uint8_t*, const CUtensorMap*, const CUtensorMap*, int) {}
  __device__ __forceinline__ void tile_begin(int m_tile, int n_tile, int q, int ln) {
    row = (long long)m_tile * kBM + q * 32 + ln;
    col0 = n_tile * kBN;
    lane = ln;
    inv = x_inv ? x_inv[row] : inv_const;
  }
  __device__ __forceinline__ float ball_sum(float v) const {
    for (int off = ns >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  __device__ __forceinline__ void tile(uint32_t taddr) {
    const bool row_ok = row < rows;                             // uniform over a ball
    const int ncols = min(kBN, cout - col0);
    const float inv_n = 1.f / (float)(ns * CPG);
    float range_max = 0.f;
#pragma unroll 1
    for (int c = 0; c < ncols; c += kChunk) {
      uint32_t r[kChunk];
#pragma unroll
      for (int i = 0; i < kChunk / 8; ++i) tc::tmem_ld_32x8(taddr + c + 8 * i, r + 8 * i);
      tc::tmem_ld_wait();
      float x[kChunk];
#pragma unroll
      for (int j4 = 0; j4 < kChunk / 4; ++j4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w_inv + col0 + c + j4 * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(bias + col0 + c + j4 * 4);
        x[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), inv * w4.x, b4.x);
        x[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), inv * w4.y, b4.y);
        x[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), inv * w4.z, b4.z);
        x[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), inv * w4.w, b4.w);
      }
      // GroupNorm per (ball, group): mean, then centred second moment (two passes over registers)
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < CPG; ++k) s += x[g * CPG + k];
        const float mean = ball_sum(s) * inv_n;
        float qv = 0.f;
#pragma unroll
        for (int k = 0; k < CPG; ++k) {
          const float d = x[g * CPG + k] - mean;
          x[g * CPG + k] = d;
          qv = fmaf(d, d, qv);
        }
        const float rstd = 1.f / sqrtf(ball_sum(qv) * inv_n + eps);
#pragma unroll
        for (int k = 0; k < CPG; ++k) x[g * CPG + k] *= rstd;
      }
#pragma unroll
      for (int j4 = 0; j4 < kChunk / 4; ++j4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + col0 + c + j4 * 4);
        const float4 e4 = *reinterpret_cast<const float4*>(beta + col0 + c + j4 * 4);
        x[4 * j4 + 0] = fmaf(x[4 * j4 + 0], g4.x, e4.x);
        x[4 * j4 + 1] = fmaf(x[4 * j4 + 1], g4.y, e4.y);
        x[4 * j4 + 2] = fmaf(x[4 * j4 + 2], g4.z, e4.z);
        x[4 * j4 + 3] = fmaf(x[4 * j4 + 3], g4.w, e4.w);
      }
      if (LAST) {
        // max over the ball's rows (no ReLU before it, pointnet2.py:693-698): integer warp reductions on an
        // order-preserving encoding; lane l of the ball keeps columns l (+ 16 when ns = 16)
        const unsigned mask = ns == 32 ? 0xffffffffu : (lane < 16 ? 0x0000ffffu : 0xffff0000u);
        const int l = lane & (ns - 1);
        float keep0 = 0.f, keep1 = 0.f;
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
          int e = __float_as_int(x[j]);
          e ^= (e >> 31) & 0x7fffffff;
          e = __reduce_max_sync(mask, e);
          e ^= (e >> 31) & 0x7fffffff;
          const float m = __int_as_float(e);
          if (j == l) keep0 = m;
          if (ns == 16 && j == l + 16) keep1 = m;
        }
        if (row_ok) {
          float* o = maxout + (row / ns) * ld_max + col0 + c;
          if (ns == 32) {
            if (l < kChunk) o[l] = keep0;
          } else {
            o[l] = keep0;
            if (l + 16 < kChunk) o[l + 16] = keep1;
          }
        }
      } else {
        if (row_ok) {
          __half* ph = out_hi + row * ld_out + col0 + c;
          __half* pl = out_lo + row * ld_out + col0 + c;
          uint32_t h[kChunk / 2], lo2[kChunk / 2];
#pragma unroll
          for (int u = 0; u < kChunk / 2; ++u) {
            const float a = fmaxf(x[2 * u], 0.f) * kBallPlaneScale;
            const float b = fmaxf(x[2 * u + 1], 0.f) * kBallPlaneScale;
            range_max = fmaxf(range_max, fmaxf(a, b));
            tcg::split2(a, b, h[u], lo2[u]);
          }
          if (kChunk == 32) {
            // 32-byte stores (st.global.v8.b32): half as many (instruction, line) pairs for the LSU as 16-byte stores
#pragma unroll
            for (int j16 = 0; j16 < kChunk / 16; ++j16) {
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ph + 16 * j16),
                           "r"(h[8 * j16]), "r"(h[8 * j16 + 1]), "r"(h[8 * j16 + 2]), "r"(h[8 * j16 + 3]),
                           "r"(h[8 * j16 + 4]), "r"(h[8 * j16 + 5]), "r"(h[8 * j16 + 6]), "r"(h[8 * j16 + 7]) : "memory");
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(pl + 16 * j16),
                           "r"(lo2[8 * j16]), "r"(lo2[8 * j16 + 1]), "r"(lo2[8 * j16 + 2]), "r"(lo2[8 * j16 + 3]),
                           "r"(lo2[8 * j16 + 4]), "r"(lo2[8 * j16 + 5]), "r"(lo2[8 * j16 + 6]), "r"(lo2[8 * j16 + 7]) : "memory");
            }
          } else {
#pragma unroll
            for (int j8 = 0; j8 < kChunk / 8; ++j8) {
              *reinterpret_cast<uint4*>(ph + 8 * j8) = make_uint4(h[4 * j8], h[4 * j8 + 1], h[4 * j8 + 2], h[4 * j8 + 3]);
              *reinterpret_cast<uint4*>(pl + 8 * j8) = make_uint4(lo2[4 * j8], lo2[4 * j8 + 1], lo2[4 * j8 + 2], lo2[4 * j8 + 3]);
            }
          }
        }
      }
    }
    if (!LAST && row_ok && col0 + kBN >= cout) {
      // zero the K padding of the next layer's operand (columns cout .. ld_out)
      for (int cc = cout; cc < ld_out; cc += 8) {
        *reinterpret_cast<uint4*>(out_hi + row * ld_out + cc) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(out_lo + row * ld_out + cc) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    if (!LAST && range_max > 65504.f) atomicOr(range_flag, 1);
  }
  __device__ __forceinline__ void chunk(int, uint32_t (&)[32]) {}
  __device__ __forceinline__ void finish() {}
};

struct SaMlpLayout {
  long long rows_pad;
  int kpad[3];
  size_t off_xinv, off_hi[3], off_lo[3], off_flag, total;
};
SaMlpLayout sa_mlp_layout(long long rows, int cin, int c1, int c2) {
  SaMlpLayout l;
  l.rows_pad = (rows + kBM - 1) / kBM * kBM;
  const int k[3] = {cin, c1, c2};
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 1024); return r; };
  l.off_xinv = take((size_t)l.rows_pad * 4);
  for (int i = 0; i < 3; ++i) {
    l.kpad[i] = (k[i] + kBK - 1) / kBK * kBK;
    l.off_hi[i] = take((size_t)l.rows_pad * l.kpad[i] * 2);
    l.off_lo[i] = take((size_t)l.rows_pad * l.kpad[i] * 2);
  }
  l.off_flag = take(256);
  l.total = p;
  return l;
}

template <int CPG, bool LAST>
int sa_mlp_layer(const __half* a_hi, const __half* a_lo, long long rows_pad, int k_pad, const float* x_inv, float inv_const,
                 const void* prepared, int cin, int cout, const float* bias, const float* gamma, const float* beta,
                 float eps, int ns, long long rows, __half* out_hi, __half* out_lo, int ld_out, float* maxout,
                 int ld_max, int* range_flag, int num_sms, cudaStream_t s) {
  const WeightLayout wl = make_weight_layout(cin, cout);
  if (wl.k_pad != k_pad) return CASPR_EINVAL;
  const char* wbase = (const char*)prepared;
  CUtensorMap tm_ahi, tm_alo, tm_whi, tm_wlo;
  bool ok = true;
  ok &= caspr_make_tmap_f16(&tm_ahi, a_hi, (uint64_t)rows_pad, (uint64_t)k_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_alo, a_lo, (uint64_t)rows_pad, (uint64_t)k_pad, kBM);
  ok &= caspr_make_tmap_f16(&tm_whi, wbase + wl.off_whi, (uint64_t)wl.cout_pad, (uint64_t)k_pad, tcg::w_box_rows());
  ok &= caspr_make_tmap_f16(&tm_wlo, wbase + wl.off_wlo, (uint64_t)wl.cout_pad, (uint64_t)k_pad, tcg::w_box_rows());
  if (!ok) return CASPR_ELAUNCH;
  BallNormEpilogue<CPG, LAST> epi{};
  epi.bias = bias; epi.x_inv = x_inv; epi.inv_const = inv_const; epi.w_inv = (const float*)(wbase + wl.off_winv);
  epi.gamma = gamma; epi.beta = beta; epi.eps = eps; epi.ns = ns; epi.rows = rows; epi.cout = cout;
  epi.out_hi = out_hi; epi.out_lo = out_lo; epi.ld_out = ld_out; epi.maxout = maxout; epi.ld_max = ld_max;
  epi.range_flag = range_flag;
  CASPR_COUNT();
  if (tcg::launch_gemm_pair(tm_ahi, tm_alo, tm_whi, tm_wlo, (int)(rows_pad / kBM), wl.cout_pad / kBN, k_pad / kBK, epi,
                            num_sms, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  return CASPR_OK;
}

template <bool LAST>
int sa_mlp_layer_dispatch(int cout, const __half* a_hi, const __half* a_lo, long long rows_pad, int k_pad,
                          const float* x_inv, float inv_const, const void* prepared, int cin, const float* bias,
                          const float* gamma, const float* beta, float eps, int ns, long long rows, __half* out_hi,
                          __half* out_lo, int ld_out, float* maxout, int ld_max, int* range_flag, int num_sms,
                          cudaStream_t s) {
#define CASPR_SA_LAYER(CPG)                                                                                          \
  return sa_mlp_layer<CPG, LAST>(a_hi, a_lo, rows_pad, k_pad, x_inv, inv_const, prepared, cin, cout, bias, gamma, beta, \
                                 eps, ns, rows, out_hi, out_lo, ld_out, maxout, ld_max, range_flag, num_sms, s)
  switch (cout) {
    case 64: CASPR_SA_LAYER(4);
    case 96: CASPR_SA_LAYER(6);
    case 128: CASPR_SA_LAYER(8);
    case 256: CASPR_SA_LAYER(16);
    case 512: CASPR_SA_LAYER(32);
    default: return CASPR_EINVAL;
  }
#undef CASPR_SA_LAYER
}

}  // namespace

extern "C" int caspr_sa_mlp_tc_supported(int ns, int Cin, int C1, int C2, int C3) {
  auto okc = [](int c) { return c == 64 || c == 96 || c == 128 || c == 256 || c == 512; };
  return (ns == 16 || ns == 32) && Cin >= 64 && okc(C1) && okc(C2) && okc(C3) && tcg::use_pair();
}

extern "C" size_t caspr_sa_mlp_tc_workspace_bytes(long long rows, int Cin, int C1, int C2) {
  if (rows <= 0 || Cin <= 0 || C1 <= 0 || C2 <= 0) return 0;
  return sa_mlp_layout(rows, Cin, C1, C2).total;
}

namespace {
// split_rows_kernel on VIRTUAL grouped rows: row (ball, j) = [xyz[idx] - centre | feat[idx]] (the layout of
// caspr_group_points, PointNet2GroupingLayer) is gathered on the fly, so the grouped tensor - 0.9 GB per encode for SA
// levels 3-5 at config 2 - is neither written nor read back.  One warp per grouped row, the row is gathered twice (the
// second time from L1 / L2), same per-row power-of-two scale as split_rows_kernel.
struct GroupGather {
  const float *xyz, *new_xyz, *feat;
  const int32_t* idx;
  int ld_feat, C, N, M, ns;
  // "delayed aggregation" of the first layer: P = feat . W1[:, 3:]^T per SOURCE point (ld ldp); W1 for its xyz columns
  const float* P;
  const float* W1;
  int ldp, ldw1;
};
__global__ void __launch_bounds__(256)
gather_split_rows_kernel(GroupGather gg, long long rows, long long rows_pad, int k_pad, __half2* __restrict__ hi,
                         __half2* __restrict__ lo, float* __restrict__ inv_scale) {
  const int lane = threadIdx.x & 31;
  const int kp2 = k_pad / 2, cols = 3 + gg.C;
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
    const long long ball = r / gg.ns;
    const long long src = (ball / gg.M) * gg.N + gg.idx[r];
    const float* px = gg.xyz + src * 3;
    const float* pc = gg.new_xyz + ball * 3;
    const float* pf = gg.feat + src * gg.ld_feat;
    auto elem = [&](int c) -> float { return c < 3 ? px[c] - pc[c] : pf[c - 3]; };
    float m = 0.f;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(elem(c)));
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
      const float a = (c < cols ? elem(c) : 0.f) * s;
      const float b = (c + 1 < cols ? elem(c + 1) : 0.f) * s;
      const __half2 hh = __floats2half2_rn(a, b);
      const float2 hf = __half22float2(hh);
      h[c2] = hh;
      l[c2] = __floats2half2_rn(a - hf.x, b - hf.y);
    }
  }
}

// First per-ball layer with the product taken BEFORE the gather: W1 . [xyz[idx] - centre | feat[idx]] + b =
// P[idx] + W1[:, :3] . (xyz[idx] - centre) + b with P = feat . W1[:, 3:]^T computed once per source point (every point is
// a member of ~24 balls at SA level 3, so the layer costs 1/24 of the grouped product and the gather moves C1 instead of
// 3 + C channels).  A thread owns a grouped row like the threads of BallNormEpilogue: per-ball GroupNorm by shuffles
// over the ball's 16 / 32 lanes, ReLU, fp16 hi / lo planes of the next layer with the fixed scale kBallPlaneScale.
template <int CPG>
__global__ void __launch_bounds__(256)
sa_first_layer_kernel(GroupGather gg, const float* __restrict__ bias, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, long long rows, int C1, __half* __restrict__ out_hi,
                      __half* __restrict__ out_lo, int ld_out, int* __restrict__ range_flag) {
  extern __shared__ float sp[];                                  // [wx | wy | wz | bias | gamma | beta] x C1
  float* swx = sp; float* swy = swx + C1; float* swz = swy + C1;
  float* sb = swz + C1; float* sg = sb + C1; float* se = sg + C1;
  for (int c = threadIdx.x; c < C1; c += blockDim.x) {
    swx[c] = gg.W1[(size_t)c * gg.ldw1]; swy[c] = gg.W1[(size_t)c * gg.ldw1 + 1]; swz[c] = gg.W1[(size_t)c * gg.ldw1 + 2];
    sb[c] = bias[c]; sg[c] = gamma[c]; se[c] = beta[c];
  }
  __syncthreads();
  constexpr int kChunk = 32, kGroups = kChunk / CPG;
  const int lane = threadIdx.x & 31, ns = gg.ns;
  const float inv_n = 1.f / (float)(ns * CPG);
  const long long tiles = (rows + 31) / 32, warps = (long long)gridDim.x * (blockDim.x >> 5);
  float range_max = 0.f;
  for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < tiles; t += warps) {
    const long long row = t * 32 + lane;
    const bool row_ok = row < rows;                              // uniform over a ball (rows % ns == 0)
    const long long r = row_ok ? row : rows - 1;
    const long long ball = r / ns;
    const long long src = (ball / gg.M) * gg.N + gg.idx[r];
    const float dx = gg.xyz[src * 3] - gg.new_xyz[ball * 3], dy = gg.xyz[src * 3 + 1] - gg.new_xyz[ball * 3 + 1],
                dz = gg.xyz[src * 3 + 2] - gg.new_xyz[ball * 3 + 2];
    const float* p = gg.P + src * gg.ldp;
#pragma unroll 1
    for (int c = 0; c < C1; c += kChunk) {
      float x[kChunk];
#pragma unroll
      for (int j4 = 0; j4 < kChunk / 4; ++j4) {
        const float4 v = *reinterpret_cast<const float4*>(p + c + 4 * j4);
        const float pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ch = c + 4 * j4 + k;
          x[4 * j4 + k] = fmaf(swz[ch], dz, fmaf(swy[ch], dy, fmaf(swx[ch], dx, pv[k]))) + sb[ch];
        }
      }
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        float sm = 0.f;
#pragma unroll
        for (int k = 0; k < CPG; ++k) sm += x[g * CPG + k];
        for (int off = ns >> 1; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
        const float mean = sm * inv_n;
        float qv = 0.f;
#pragma unroll
        for (int k = 0; k < CPG; ++k) {
          const float d = x[g * CPG + k] - mean;
          x[g * CPG + k] = d;
          qv = fmaf(d, d, qv);
        }
        for (int off = ns >> 1; off > 0; off >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, off);
        const float rstd = 1.f / sqrtf(qv * inv_n + eps);
#pragma unroll
        for (int k = 0; k < CPG; ++k) x[g * CPG + k] *= rstd;
      }
      if (row_ok) {
        uint32_t h[kChunk / 2], l2[kChunk / 2];
#pragma unroll
        for (int u = 0; u < kChunk / 2; ++u) {
          const float a = fmaxf(fmaf(x[2 * u], sg[c + 2 * u], se[c + 2 * u]), 0.f) * kBallPlaneScale;
          const float b = fmaxf(fmaf(x[2 * u + 1], sg[c + 2 * u + 1], se[c + 2 * u + 1]), 0.f) * kBallPlaneScale;
          range_max = fmaxf(range_max, fmaxf(a, b));
          tcg::split2(a, b, h[u], l2[u]);
        }
        __half* ph = out_hi + row * ld_out + c;
        __half* pl = out_lo + row * ld_out + c;
#pragma unroll
        for (int j16 = 0; j16 < kChunk / 16; ++j16) {
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ph + 16 * j16),
                       "r"(h[8 * j16]), "r"(h[8 * j16 + 1]), "r"(h[8 * j16 + 2]), "r"(h[8 * j16 + 3]),
                       "r"(h[8 * j16 + 4]), "r"(h[8 * j16 + 5]), "r"(h[8 * j16 + 6]), "r"(h[8 * j16 + 7]) : "memory");
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(pl + 16 * j16),
                       "r"(l2[8 * j16]), "r"(l2[8 * j16 + 1]), "r"(l2[8 * j16 + 2]), "r"(l2[8 * j16 + 3]),
                       "r"(l2[8 * j16 + 4]), "r"(l2[8 * j16 + 5]), "r"(l2[8 * j16 + 6]), "r"(l2[8 * j16 + 7]) : "memory");
        }
      }
    }
    if (row_ok) {
      for (int cc = C1; cc < ld_out; cc += 8) {                  // K padding of the next layer's operand
        *reinterpret_cast<uint4*>(out_hi + row * ld_out + cc) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(out_lo + row * ld_out + cc) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  if (range_max > 65504.f) atomicOr(range_flag, 1);
}

int sa_mlp_tc_impl(const float* X, int ldx, const GroupGather* gg, long long rows, int Cin, int ns,
                   const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                   const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                   const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                   float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes, void* stream);
}  // namespace

extern "C" size_t caspr_sa_mlp_tc_delayed_workspace_bytes(long long rows, int C1, int C2) {
  if (rows <= 0 || C1 <= 0 || C2 <= 0) return 0;
  return sa_mlp_layout(rows, 0, C1, C2).total;
}

extern "C" int caspr_sa_mlp_tc_delayed(const float* xyz, const float* new_xyz, const float* P, int ldp, const int32_t* idx,
                                       int B, int N, int M, int ns, const float* W1, int ldw1, const float* b1,
                                       const float* g1, const float* e1, int C1,
                                       const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                                       const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                                       float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && P && idx && W1 && B > 0 && N > 0 && M > 0 && ldp >= C1 && ldw1 >= 3);
  CASPR_REQUIRE(ldp % 4 == 0 && ((uintptr_t)P & 15) == 0 && (C1 == 64 || C1 == 128 || C1 == 256));
  GroupGather gg;
  memset(&gg, 0, sizeof(gg));
  gg.xyz = xyz; gg.new_xyz = new_xyz; gg.idx = idx; gg.N = N; gg.M = M; gg.ns = ns;
  gg.P = P; gg.ldp = ldp; gg.W1 = W1; gg.ldw1 = ldw1;
  return sa_mlp_tc_impl(nullptr, 0, &gg, (long long)B * M * ns, 64, ns, nullptr, b1, g1, e1, C1, prep2, b2, g2, e2, C2,
                        prep3, b3, g3, e3, C3, eps, maxout, ld_max, workspace, workspace_bytes, stream);
}

extern "C" int caspr_sa_mlp_tc_grouped(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                                       const int32_t* idx, int B, int N, int M, int ns,
                                       const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                                       const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                                       const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                                       float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && feat && idx && B > 0 && N > 0 && M > 0 && C > 0 && ld_feat >= C);
  GroupGather gg;
  memset(&gg, 0, sizeof(gg));
  gg.xyz = xyz; gg.new_xyz = new_xyz; gg.feat = feat; gg.idx = idx; gg.ld_feat = ld_feat; gg.C = C; gg.N = N; gg.M = M;
  gg.ns = ns;
  return sa_mlp_tc_impl(nullptr, 0, &gg, (long long)B * M * ns, 3 + C, ns, prep1, b1, g1, e1, C1, prep2, b2, g2, e2, C2,
                        prep3, b3, g3, e3, C3, eps, maxout, ld_max, workspace, workspace_bytes, stream);
}

extern "C" int caspr_sa_mlp_tc(const float* X, int ldx, long long rows, int Cin, int ns,
                               const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                               const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                               const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                               float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes,
                               void* stream) {
  CASPR_REQUIRE(X && ldx >= Cin);
  return sa_mlp_tc_impl(X, ldx, nullptr, rows, Cin, ns, prep1, b1, g1, e1, C1, prep2, b2, g2, e2, C2, prep3, b3, g3, e3,
                        C3, eps, maxout, ld_max, workspace, workspace_bytes, stream);
}

namespace {
int sa_mlp_tc_impl(const float* X, int ldx, const GroupGather* gg, long long rows, int Cin, int ns,
                   const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                   const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                   const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                   float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes, void* stream) {
  const bool delayed = gg && gg->P;
  CASPR_REQUIRE((prep1 || delayed) && prep2 && prep3 && b1 && b2 && b3 && g1 && g2 && g3 && e1 && e2 && e3 && maxout &&
                workspace);
  CASPR_REQUIRE(rows > 0 && rows % ns == 0 && ld_max >= C3);
  CASPR_REQUIRE(caspr_sa_mlp_tc_supported(ns, Cin, C1, C2, C3));
  CASPR_REQUIRE(((uintptr_t)workspace & 1023) == 0 && (((uintptr_t)prep1 | (uintptr_t)prep2 | (uintptr_t)prep3) & 1023) == 0);
  const SaMlpLayout l = sa_mlp_layout(rows, delayed ? 0 : Cin, C1, C2);
  if (workspace_bytes < l.total) return CASPR_EWORKSPACE;
  CASPR_REQUIRE(l.rows_pad / kBM < (1ll << 30));
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  float* xinv = (float*)(base + l.off_xinv);
  __half* hi[3];
  __half* lo[3];
  for (int i = 0; i < 3; ++i) { hi[i] = (__half*)(base + l.off_hi[i]); lo[i] = (__half*)(base + l.off_lo[i]); }
  int* flag = (int*)(base + l.off_flag);
  if (cudaMemsetAsync(flag, 0, sizeof(int), s) != cudaSuccess) return CASPR_ELAUNCH;
  int dev = 0, num_sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (delayed) {
    // a warp owns 32 rows; few rows (the coarse levels) get small CTAs so that every SM has work
    const long long tiles = (rows + 31) / 32;
    const int wpb = tiles >= 148 * 16 ? 8 : 2;
    const int blocks = (int)std::min<long long>((tiles + wpb - 1) / wpb, 148 * 8);
    const size_t smem = (size_t)6 * C1 * sizeof(float);
#define CASPR_SA_FIRST(CPG) \
    sa_first_layer_kernel<CPG><<<blocks, 32 * wpb, smem, s>>>(*gg, b1, g1, e1, eps, rows, C1, hi[1], lo[1], l.kpad[1], flag)
    CASPR_COUNT();
    if (C1 == 64) CASPR_SA_FIRST(4);
    else if (C1 == 128) CASPR_SA_FIRST(8);
    else CASPR_SA_FIRST(16);
#undef CASPR_SA_FIRST
    CASPR_CHECK_LAUNCH();
    int rc = sa_mlp_layer_dispatch<false>(C2, hi[1], lo[1], l.rows_pad, l.kpad[1], nullptr, 1.f / kBallPlaneScale, prep2, C1,
                                          b2, g2, e2, eps, ns, rows, hi[2], lo[2], l.kpad[2], nullptr, 0, flag, num_sms, s);
    if (rc) return rc;
    rc = sa_mlp_layer_dispatch<true>(C3, hi[2], lo[2], l.rows_pad, l.kpad[2], nullptr, 1.f / kBallPlaneScale, prep3, C2, b3,
                                     g3, e3, eps, ns, rows, nullptr, nullptr, 0, maxout, ld_max, flag, num_sms, s);
    if (rc) return rc;
    CASPR_CHECK_LAUNCH();
    return CASPR_OK;
  }
  // operand planes of the grouped input rows (per-row power-of-two scale)
  if (gg) {
    CASPR_COUNT(); gather_split_rows_kernel<<<148 * 8, 256, 0, s>>>(*gg, rows, l.rows_pad, l.kpad[0], (__half2*)hi[0],
                                                                     (__half2*)lo[0], xinv);
  } else {
    CASPR_COUNT(); launch_split_rows(X, ldx, rows, Cin, l.rows_pad, l.kpad[0], 0, hi[0], lo[0], xinv, NormFold(), 148 * 8, s);
  }
  CASPR_CHECK_LAUNCH();
  int rc = sa_mlp_layer_dispatch<false>(C1, hi[0], lo[0], l.rows_pad, l.kpad[0], xinv, 0.f, prep1, Cin, b1, g1, e1, eps, ns,
                                        rows, hi[1], lo[1], l.kpad[1], nullptr, 0, flag, num_sms, s);
  if (rc) return rc;
  rc = sa_mlp_layer_dispatch<false>(C2, hi[1], lo[1], l.rows_pad, l.kpad[1], nullptr, 1.f / kBallPlaneScale, prep2, C1, b2,
                                    g2, e2, eps, ns, rows, hi[2], lo[2], l.kpad[2], nullptr, 0, flag, num_sms, s);
  if (rc) return rc;
  rc = sa_mlp_layer_dispatch<true>(C3, hi[2], lo[2], l.rows_pad, l.kpad[2], nullptr, 1.f / kBallPlaneScale, prep3, C2, b3,
                                   g3, e3, eps, ns, rows, nullptr, nullptr, 0, maxout, ld_max, flag, num_sms, s);
  if (rc) return rc;
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
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

namespace {
__global__ void extrema_init_kernel(unsigned* ext, int samples, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < samples * 2 * C) ext[i] = ((i / C) & 1) ? 0xffffffffu : 0u;      // [sample][max | min][C]
}
// same arithmetic as groupnorm_apply_vec_kernel (dense.cu): mean / rstd from the fp64 sums, then
// (x - mean) * rstd * gamma + beta - evaluated at the channel's max (gamma >= 0) or min (gamma < 0)
__global__ void gn_max_from_extrema_kernel(const double* __restrict__ stats, const unsigned* __restrict__ ext, int samples,
                                           int rows_per_sample, int C, int groups, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, float eps, float* __restrict__ maxout,
                                           int ld_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= samples * C) return;
  const int sm = i / C, c = i - sm * C;
  const int cpg = C / groups, g = c / cpg;
  const double cnt = (double)cpg * (double)rows_per_sample;
  const double s = stats[((size_t)sm * groups + g) * 2], q = stats[((size_t)sm * groups + g) * 2 + 1];
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float mu = (float)mean, rs = (float)(1.0 / sqrt(var + (double)eps));
  const float ga = gamma[c], be = beta[c];
  const float hi = ordered_to_float(ext[(size_t)sm * 2 * C + c]), lo = ordered_to_float(ext[(size_t)sm * 2 * C + C + c]);
  const float a = fmaf(__fmul_rn(__fsub_rn(hi, mu), rs), ga, be), b = fmaf(__fmul_rn(__fsub_rn(lo, mu), rs), ga, be);
  maxout[(size_t)sm * ld_max + c] = fmaxf(a, b);
}
}  // namespace

extern "C" int caspr_gn_max_from_extrema(const double* stats, const unsigned* extrema, int samples, int rows_per_sample,
                                         int C, int groups, const float* gamma, const float* beta, float eps,
                                         float* maxout, int ld_max, void* stream) {
  CASPR_REQUIRE(stats && extrema && gamma && beta && maxout && samples > 0 && rows_per_sample > 0 && C > 0);
  CASPR_REQUIRE(groups > 0 && C % groups == 0 && ld_max >= C);
  CASPR_COUNT(); gn_max_from_extrema_kernel<<<ceil_div(samples * C, 256), 256, 0, (cudaStream_t)stream>>>(
      stats, extrema, samples, rows_per_sample, C, groups, gamma, beta, eps, maxout, ld_max);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_linear_tc(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y,
                               int ldy, int rows, int Cin, int Cout, int act_in, int act_out,
                               const void* prepared_weights, const caspr_gn_fold* in_norm,
                               const caspr_gn_stats* out_stats, int bias_rows_per_sample, void* workspace,
                               size_t workspace_bytes, void* stream) {
  const bool reduce_only = out_stats && out_stats->extrema;      // statistics + column extrema, no Y
  CASPR_REQUIRE(X && (W || prepared_weights) && (Y || reduce_only) && workspace && rows > 0 && Cin > 0 && Cout > 0);
  CASPR_REQUIRE(bias_rows_per_sample >= 0 && (bias_rows_per_sample == 0 || bias));
  // the epilogue stores float4 pieces and keeps one running GroupNorm group per chunk of 32 columns
  CASPR_REQUIRE(Cout % 4 == 0 && (!bias || ((uintptr_t)bias & 15) == 0));
  CASPR_REQUIRE(reduce_only || (ldy % 4 == 0 && ((uintptr_t)Y & 15) == 0 && ldy >= Cout));
  CASPR_REQUIRE(!reduce_only || act_out == CASPR_ACT_NONE);
  if (out_stats) CASPR_REQUIRE(Cout / out_stats->groups >= 32);
  CASPR_REQUIRE(act_out == CASPR_ACT_NONE || act_out == CASPR_ACT_RELU);     // sigmoid outputs: caspr_linear
  if (in_norm)
    CASPR_REQUIRE(in_norm->table && in_norm->rows_per_sample > 0);
  if (out_stats)
    CASPR_REQUIRE(out_stats->stats && out_stats->groups > 0 && Cout % out_stats->groups == 0 &&
                  out_stats->rows_per_sample > 0 && out_stats->rows_per_sample % 32 == 0 &&
                  rows % out_stats->rows_per_sample == 0);
  CASPR_REQUIRE(ldx >= Cin && (!W || ldw >= Cin));
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
    if (reduce_only) {
      const int samples = rows / out_stats->rows_per_sample;
      CASPR_COUNT(); extrema_init_kernel<<<ceil_div(samples * 2 * Cout, 256), 256, 0, s>>>(out_stats->extrema, samples, Cout);
      CASPR_CHECK_LAUNCH();
    }
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
  epi.ext = reduce_only ? out_stats->extrema : nullptr;
  epi.vec_ok = (ldy % 4 == 0) && (((uintptr_t)Y & 15) == 0) && (!bias || ((uintptr_t)bias & 15) == 0);
  const int m_tiles = (int)(l.rows_pad / kBM), n_tiles = l.cout_pad / kBN;
  // fp32 output through TMA (full lines); CASPR_LINEAR_TMA_STORE=0 keeps the direct 16-byte stores
  CUtensorMap tm_y = tm_xhi;
  {
    static int want = -1;
    if (want < 0) { const char* e = getenv("CASPR_LINEAR_TMA_STORE"); want = (e && e[0] == '0') ? 0 : 1; }
    epi.use_tma = want && !reduce_only &&
                  caspr_make_tmap_f32_box32(&tm_y, Y, (uint64_t)rows, (uint64_t)Cout, (uint64_t)ldy, kBM);
  }
  caspr_prof_begin(CASPR_PROF_LINEAR, s);
  CASPR_COUNT();
  const cudaError_t lerr = tcg::launch_gemm(tm_xhi, tm_xlo, tm_whi, tm_wlo, tm_y, tm_xlo, 0, m_tiles, n_tiles,
                                            l.k_pad / kBK, nullptr, epi, 1, num_sms, s);
  caspr_prof_end(CASPR_PROF_LINEAR, s);
  if (lerr != cudaSuccess) return CASPR_ELAUNCH;
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
