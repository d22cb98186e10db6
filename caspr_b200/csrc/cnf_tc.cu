// Tensor-core engine (CASPR_CNF_TC_FP16X3) of the CNF dynamics: the two H x H ConcatSquash
// layers of ODEnet (odefunc.py:98-105, diffeq_layers.py:83-90) over [activation ; tangent] rows
// as tcgen05 GEMMs with fp32-grade accuracy from three fp16 products
//     a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo        (a = a_hi + a_lo, w = w_hi + w_lo)
// accumulated in fp32 in TMEM.
//
// Data layout (HBM): activations live as two fp16 planes (hi, lo) of shape [rows][512], K-major,
// pre-scaled by 2^4; a 128-row tile holds 64 points, warp-quadrant q = rows 32q..32q+31 carries the
// activation rows of points 16q..16q+15 in lanes 0..15 and their tangent rows in lanes 16..31, so
// the epilogue pairs a point's two rows with one warp shuffle.  Weights are split once per solve
// into fp16 hi/lo planes scaled by a power of two chosen from max|W| (scales are undone exactly in
// the epilogue through the pre-scaled gate).
//
// Kernel: persistent, warp-specialised, one CTA per SM (192 threads):
//   warp 0   TMA producer  : per 64-wide k-chunk A_hi, A_lo (128x64) and W_hi, W_lo (256x64) tiles,
//                            128-byte swizzle, 2-stage mbarrier ring (96 KB per stage)
//   warp 1   MMA issuer    : 3 x tcgen05.mma (M=128, N=256, K=16) per k-step into one of two
//                            256-column TMEM accumulators; tcgen05.commit frees the smem stage /
//                            publishes the accumulator
//   warps 2-5 epilogue     : tcgen05.ld, ConcatSquash gate/bias, softplus and the tangent's chain
//                            rule, split to fp16 hi/lo (or fp32 for the last layer's input), stores
// so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "dopri5.cuh"
#include "cnf_state.cuh"
#include "cnf_tc.cuh"
#include "tc_gemm.cuh"

namespace cnf_tc {

namespace {

using tcg::kBM;
using tcg::kBN;
using tcg::split2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Epilogue of one H x H ConcatSquash layer: rows of a 128-row tile are 64 points, quadrant q holds the
// activation rows of points 16q..16q+15 in lanes 0-15 and their tangent rows in lanes 16-31.
template <bool OUT_F32>
struct CnfEpilogue {
  const float* gate;     // per-frame gate of this layer, pre-multiplied by 1/(act_scale*w_scale)
  const float* biasf;    // per-frame folded bias  b*gate + hyper_bias
  int ld_hyper;
  int n;                 // points
  int P;                 // points per frame
  __half* out_hi;        // OUT_F16: next layer's planes
  __half* out_lo;
  const float* W3;       // OUT_F32 (= last mid layer): the 3 x 512 weights of the output layer, which is
  float* acc6;           //   fused here: per-point partial sums [n][8] = {W3.h (3), W3.v (3), -, -}
  int* range_flag;
  // per-thread tile state
  const float* gp;
  const float* bp;
  size_t row_h;
  int pt, col0, is_v;
  bool live;
  float range_max;
  float pa[3], pv[3];    // fused output layer: partial dot products of this thread's columns
  // TMA-store staging (OUT_F16): hi box at staging, lo box 16 KB behind it
  uint8_t* stg;
  const CUtensorMap* tm_hi;
  const CUtensorMap* tm_lo;
  int etid, m_tile_cur, n_tile_cur, box_row;
  bool stores_in_flight;

  __device__ __forceinline__ void setup(uint8_t* staging, const CUtensorMap* o_hi, const CUtensorMap* o_lo, int epi_tid) {
    stg = staging; tm_hi = o_hi; tm_lo = o_lo; etid = epi_tid; stores_in_flight = false;
  }

  __device__ __forceinline__ void tile_begin(int m_tile, int n_tile, int q, int lane) {
    is_v = lane >> 4;
    const int pl = q * 16 + (lane & 15);
    pt = m_tile * 64 + pl;
    live = pt < n;
    const int f = (live ? pt : n - 1) / P;
    col0 = n_tile * kBN + is_v * 16;
    gp = gate + (size_t)f * ld_hyper + col0;
    bp = biasf + (size_t)f * ld_hyper + col0;
    row_h = (size_t)m_tile * kBM + q * 32 + (lane & 15);
    m_tile_cur = m_tile; n_tile_cur = n_tile; box_row = q * 32 + (lane & 15);
    if (OUT_F32) {
      if (pending) flush();                // previous tile's sums (kept until now to overlap the atomics)
      pa[0] = pa[1] = pa[2] = pv[0] = pv[1] = pv[2] = 0.f;
      pending = true;
    }
  }

  bool pending;
  int flush_pt;
  bool flush_live;
  __device__ __forceinline__ void flush() {
    // pair lanes L and L+16 hold the two column halves of the same point
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pa[c] += __shfl_xor_sync(0xffffffffu, pa[c], 16);
      pv[c] += __shfl_xor_sync(0xffffffffu, pv[c], 16);
    }
    if (flush_live && !flush_is_v) {
      float* a = acc6 + (size_t)flush_pt * 8;
      atomicAdd(a + 0, pa[0]); atomicAdd(a + 1, pa[1]); atomicAdd(a + 2, pa[2]);
      atomicAdd(a + 3, pv[0]); atomicAdd(a + 4, pv[1]); atomicAdd(a + 5, pv[2]);
    }
    pending = false;
  }
  int flush_is_v;

  __device__ __forceinline__ void chunk(int chunk, uint32_t (&r)[32]) {
    // lanes 0-15 keep columns 0-15 of the chunk, lanes 16-31 columns 16-31; each lane receives its
    // partner row (tangent resp. activation) for those columns
    float ah[16], av[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t send = is_v ? r[j] : r[16 + j];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 16);
      ah[j] = __uint_as_float(is_v ? recv : r[j]);
      av[j] = __uint_as_float(is_v ? r[16 + j] : recv);
    }
    float g[16], b[16];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 g4 = *reinterpret_cast<const float4*>(gp + chunk * 32 + j4 * 4);
      const float4 b4 = *reinterpret_cast<const float4*>(bp + chunk * 32 + j4 * 4);
      g[j4 * 4 + 0] = g4.x; g[j4 * 4 + 1] = g4.y; g[j4 * 4 + 2] = g4.z; g[j4 * 4 + 3] = g4.w;
      b[j4 * 4 + 0] = b4.x; b[j4 * 4 + 1] = b4.y; b[j4 * 4 + 2] = b4.z; b[j4 * 4 + 3] = b4.w;
    }
    float ho[16], vo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float pre = fmaf(ah[j], g[j], b[j]);
      // softplus (beta 1, threshold 20) and its derivative from the SFU approximations
      const float z = ex2_approx(fminf(pre, 40.f) * kLog2e);
      const float t = 1.f + z;
      const bool big = pre > 20.f;
      const float sp = big ? pre : kLn2 * lg2_approx(t);
      const float dsp = big ? 1.f : z * rcp_approx(t);
      ho[j] = sp;
      vo[j] = dsp * g[j] * av[j];
    }
    const int col = col0 + chunk * 32;
    if (OUT_F32) {
      // fused output layer (H -> 3): accumulate W3[c][col..col+15] . {h', v'} for this thread's columns
      flush_pt = pt; flush_live = live; flush_is_v = is_v;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4* w4 = reinterpret_cast<const float4*>(W3 + c * 512 + col);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w = __ldg(w4 + j4);
          pa[c] = fmaf(w.x, ho[4 * j4], pa[c]); pv[c] = fmaf(w.x, vo[4 * j4], pv[c]);
          pa[c] = fmaf(w.y, ho[4 * j4 + 1], pa[c]); pv[c] = fmaf(w.y, vo[4 * j4 + 1], pv[c]);
          pa[c] = fmaf(w.z, ho[4 * j4 + 2], pa[c]); pv[c] = fmaf(w.z, vo[4 * j4 + 2], pv[c]);
          pa[c] = fmaf(w.w, ho[4 * j4 + 3], pa[c]); pv[c] = fmaf(w.w, vo[4 * j4 + 3], pv[c]);
        }
      }
    } else {
      // Stage 64 columns (two chunks) of the hi / lo planes in shared memory in the 128-byte-swizzled
      // box layout of the tensor map, then one thread stores both boxes with TMA (full 128-byte lines
      // instead of 16-byte scattered stores).
      uint32_t hh[8], hl[8], vh[8], vl[8];
#pragma unroll
      for (int j2 = 0; j2 < 8; ++j2) {
        const float h0 = ho[2 * j2] * kActScale, h1 = ho[2 * j2 + 1] * kActScale;
        const float v0 = vo[2 * j2] * kActScale, v1 = vo[2 * j2 + 1] * kActScale;
        if (live) range_max = fmaxf(range_max, fmaxf(fmaxf(fabsf(h0), fabsf(h1)), fmaxf(fabsf(v0), fabsf(v1))));
        split2(h0, h1, hh[j2], hl[j2]);
        split2(v0, v1, vh[j2], vl[j2]);
      }
      const int odd = chunk & 1;
      if (!odd && stores_in_flight) {
        if (etid == 0) tc::tma_store_wait_read();          // previous boxes have left shared memory
        tc::named_bar_sync(1, 128);
      }
      // 16-byte chunk index inside the 128-byte box row, XOR-swizzled with the row
      const int cc0 = odd * 4 + is_v * 2;
      uint8_t* hi_box = stg;
      uint8_t* lo_box = stg + kBM * 128;
      const int rh = box_row, rv = box_row + 16;
      *reinterpret_cast<uint4*>(hi_box + rh * 128 + (((cc0) ^ (rh & 7)) << 4)) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      *reinterpret_cast<uint4*>(hi_box + rh * 128 + (((cc0 + 1) ^ (rh & 7)) << 4)) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
      *reinterpret_cast<uint4*>(lo_box + rh * 128 + (((cc0) ^ (rh & 7)) << 4)) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
      *reinterpret_cast<uint4*>(lo_box + rh * 128 + (((cc0 + 1) ^ (rh & 7)) << 4)) = make_uint4(hl[4], hl[5], hl[6], hl[7]);
      *reinterpret_cast<uint4*>(hi_box + rv * 128 + (((cc0) ^ (rv & 7)) << 4)) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
      *reinterpret_cast<uint4*>(hi_box + rv * 128 + (((cc0 + 1) ^ (rv & 7)) << 4)) = make_uint4(vh[4], vh[5], vh[6], vh[7]);
      *reinterpret_cast<uint4*>(lo_box + rv * 128 + (((cc0) ^ (rv & 7)) << 4)) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
      *reinterpret_cast<uint4*>(lo_box + rv * 128 + (((cc0 + 1) ^ (rv & 7)) << 4)) = make_uint4(vl[4], vl[5], vl[6], vl[7]);
      if (odd) {
        tc::fence_proxy_async_smem();
        tc::named_bar_sync(1, 128);
        if (etid == 0) {
          const int c_out = n_tile_cur * kBN + (chunk >> 1) * 64;
          tc::tma_store_2d(tm_hi, hi_box, c_out, m_tile_cur * kBM);
          tc::tma_store_2d(tm_lo, lo_box, c_out, m_tile_cur * kBM);
          tc::tma_store_commit();
        }
        stores_in_flight = true;
      }
    }
  }

  __device__ __forceinline__ void finish() {
    if (!OUT_F32 && stores_in_flight && etid == 0) tc::tma_store_wait_all();
    if (OUT_F32 && pending) flush();
    if (!OUT_F32 && range_max > 65504.f) atomicOr(range_flag, 1);
  }
};

// ------------------------------------------------------------------------- weight split
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ w, int nelem, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// scale = 2^(14 - e) with max|W| < 2^e  ->  max|W*scale| in [2^13, 2^14); inv = 1/(act_scale*scale)
__global__ void weight_scale_kernel(const unsigned* __restrict__ max_bits, float* __restrict__ scales, int layers) {
  const int l = threadIdx.x;
  if (l >= layers) return;
  const float m = __uint_as_float(max_bits[l]);
  float scale = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);                                   // m = f * 2^e, f in [0.5, 1)
    scale = ldexpf(1.f, 14 - e);
  }
  scales[2 * l] = scale;
  scales[2 * l + 1] = 1.f / (kActScale * scale);
}

__global__ void __launch_bounds__(256)
weight_split_kernel(const float* __restrict__ w, int nelem, const float* __restrict__ scales, int layer,
                    __half* __restrict__ hi, __half* __restrict__ lo) {
  const float s = scales[2 * layer];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += gridDim.x * blockDim.x) {
    const float x = w[i] * s;
    const __half h = __float2half_rn(x);
    hi[i] = h;
    lo[i] = __float2half_rn(x - __half2float(h));
  }
}

// gate columns of the two tensor-core layers get the exact power-of-two factor 1/(act_scale*w_scale)
__global__ void fill_col_scale_kernel(const float* __restrict__ scales, int H, int ctot, float* __restrict__ col_scale) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ctot) return;
  float s = 1.f;
  if (j >= H && j < 2 * H) s = scales[1];
  else if (j >= 2 * H && j < 3 * H) s = scales[3];
  col_scale[j] = s;
}

// Layer 0 (3 -> 512) for the tensor-core engine: same math as cnf_layer0_kernel (cnf.cu) but the
// outputs are written as scaled fp16 hi/lo planes in the tile/quadrant row order described above.
__global__ void __launch_bounds__(256)
cnf_tc_layer0_kernel(const float4* __restrict__ y0, const float4* __restrict__ kbuf, size_t kstride,
                     const float* __restrict__ e, const float* __restrict__ W0, int pt0, int n, int P, int stage,
                     const float* __restrict__ gate, const float* __restrict__ biasf, int ld_hyper,
                     const CnfState* __restrict__ st, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                     int* __restrict__ range_flag) {
  if (st->done) return;
  constexpr int H = 512;
  // no shared memory: the kernel must be able to co-reside with the persistent GEMM CTAs (which own almost all
  // of it) when the two halves of the point set are pipelined on two streams; the 6 KB of layer-0 weights are
  // read through L1, 24 contiguous floats (8 channels x 3) per lane and channel set
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float dt = (float)st->dt;
  float range_max = 0.f;
  for (int pt = pt0 + blockIdx.x * wpb + (threadIdx.x >> 5); pt < n; pt += gridDim.x * wpb) {      // points [pt0, n)
    float4 y = y0[pt];
    float ys[3] = {y.x, y.y, y.z};
    if (stage > 0) {
      float kx[6], ky[6], kz[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < stage) {
          float4 kv = kbuf[(size_t)j * kstride + pt];
          kx[j] = kv.x; ky[j] = kv.y; kz[j] = kv.z;
        } else {
          kx[j] = ky[j] = kz[j] = 0.f;
        }
      }
      ys[0] = dopri5::stage_combine(y.x, dt, kx, stage - 1);
      ys[1] = dopri5::stage_combine(y.y, dt, ky, stage - 1);
      ys[2] = dopri5::stage_combine(y.z, dt, kz, stage - 1);
    }
    const float e0 = e[3 * (size_t)pt], e1 = e[3 * (size_t)pt + 1], e2 = e[3 * (size_t)pt + 2];
    const int f = pt / P;
    const int pl = pt & 63;
    const size_t row_h = (size_t)(pt >> 6) * 128 + (pl >> 4) * 32 + (pl & 15);
    const size_t row_v = row_h + 16;
    // each lane produces two runs of 8 consecutive channels (j = set*256 + lane*8 + 0..7), so every
    // 16-byte store instruction of the warp covers 512 contiguous bytes of one plane row
#pragma unroll
    for (int set = 0; set < 2; ++set) {
      const int j0 = set * 256 + lane * 8;
      const float4* g4p = reinterpret_cast<const float4*>(gate + (size_t)f * ld_hyper + j0);
      const float4* b4p = reinterpret_cast<const float4*>(biasf + (size_t)f * ld_hyper + j0);
      float g[8], bf[8];
#pragma unroll
      for (int j4 = 0; j4 < 2; ++j4) {
        const float4 a = g4p[j4], b = b4p[j4];
        g[4 * j4] = a.x; g[4 * j4 + 1] = a.y; g[4 * j4 + 2] = a.z; g[4 * j4 + 3] = a.w;
        bf[4 * j4] = b.x; bf[4 * j4 + 1] = b.y; bf[4 * j4 + 2] = b.z; bf[4 * j4 + 3] = b.w;
      }
      float wl[24];
#pragma unroll
      for (int j4 = 0; j4 < 6; ++j4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(W0 + 3 * j0) + j4);
        wl[4 * j4] = w4.x; wl[4 * j4 + 1] = w4.y; wl[4 * j4 + 2] = w4.z; wl[4 * j4 + 3] = w4.w;
      }
      uint32_t hh[4], hl[4], vh[4], vl[4];
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        float hv[2], vv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int jj = j2 * 2 + u;
          const float w0 = wl[3 * jj], w1 = wl[3 * jj + 1], w2 = wl[3 * jj + 2];
          const float a = fmaf(w2, ys[2], fmaf(w1, ys[1], w0 * ys[0]));
          const float ta = fmaf(w2, e2, fmaf(w1, e1, w0 * e0));
          const float pre = fmaf(a, g[jj], bf[jj]);
          const float z = ex2_approx(fminf(pre, 40.f) * kLog2e);
          const float t = 1.f + z;
          const bool big = pre > 20.f;
          const float sp = big ? pre : kLn2 * lg2_approx(t);
          const float dsp = big ? 1.f : z * rcp_approx(t);
          hv[u] = sp * kActScale;
          vv[u] = dsp * g[jj] * ta * kActScale;
          range_max = fmaxf(range_max, fmaxf(fabsf(hv[u]), fabsf(vv[u])));
        }
        split2(hv[0], hv[1], hh[j2], hl[j2]);
        split2(vv[0], vv[1], vh[j2], vl[j2]);
      }
      *reinterpret_cast<uint4*>(out_hi + row_h * H + j0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      *reinterpret_cast<uint4*>(out_lo + row_h * H + j0) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
      *reinterpret_cast<uint4*>(out_hi + row_v * H + j0) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
      *reinterpret_cast<uint4*>(out_lo + row_v * H + j0) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
    }
  }
  if (range_max > 65504.f) atomicOr(range_flag, 1);
}

// Output layer, second half: k = sign * (W3.h * gate + biasf, -(e . (gate * W3.v))) from the per-point sums
// the last tensor-core layer accumulated; the sums are cleared for the next evaluation.
__global__ void __launch_bounds__(256)
cnf_tc_last_finish_kernel(float* __restrict__ acc6, const float* __restrict__ e, int pt0, int n, int P,
                          const float* __restrict__ gate, const float* __restrict__ biasf, int ld_hyper, int reverse,
                          const CnfState* __restrict__ st, float4* __restrict__ kout) {
  if (st->done) return;
  const int pt = pt0 + blockIdx.x * blockDim.x + threadIdx.x;          // points [pt0, n)
  if (pt >= n) return;
  float4* a4 = reinterpret_cast<float4*>(acc6 + (size_t)pt * 8);
  const float4 lo = a4[0], hi = a4[1];
  a4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  a4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int f = pt / P;
  const float* g = gate + (size_t)f * ld_hyper;
  const float* bf = biasf + (size_t)f * ld_hyper;
  const float dy0 = fmaf(lo.x, g[0], bf[0]);
  const float dy1 = fmaf(lo.y, g[1], bf[1]);
  const float dy2 = fmaf(lo.z, g[2], bf[2]);
  const float e0 = e[3 * (size_t)pt], e1 = e[3 * (size_t)pt + 1], e2 = e[3 * (size_t)pt + 2];
  const float div = (g[0] * lo.w) * e0 + (g[1] * hi.x) * e1 + (g[2] * hi.y) * e2;
  kout[pt] = reverse ? make_float4(-dy0, -dy1, -dy2, div) : make_float4(dy0, dy1, dy2, -div);
}

}  // namespace

size_t weights_workspace_bytes() {
  // 2 layers x (hi, lo) fp16 planes + scales + max bits + col scale
  return 2 * 2 * (size_t)512 * 512 * 2 + 1024;
}

int prepare_weights(const float* W1, const float* W2, const Weights& out, cudaStream_t s) {
  const int nelem = 512 * 512;
  if (cudaMemsetAsync(out.max_bits, 0, 2 * sizeof(unsigned), s) != cudaSuccess) return CASPR_ELAUNCH;
  const float* w[2] = {W1, W2};
  for (int l = 0; l < 2; ++l) {
    CASPR_COUNT(); absmax_kernel<<<64, 256, 0, s>>>(w[l], nelem, out.max_bits + l);
  }
  CASPR_COUNT(); weight_scale_kernel<<<1, 32, 0, s>>>(out.max_bits, out.scales, 2);
  for (int l = 0; l < 2; ++l) {
    CASPR_COUNT(); weight_split_kernel<<<128, 256, 0, s>>>(w[l], nelem, out.scales, l, out.hi[l], out.lo[l]);
  }
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int fill_col_scale(const Weights& w, int ctot, float* col_scale, cudaStream_t s) {
  CASPR_COUNT(); fill_col_scale_kernel<<<ceil_div(ctot, 256), 256, 0, s>>>(w.scales, 512, ctot, col_scale);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int make_plan(Plan& plan, const Weights& w, __half* a_hi, __half* a_lo, __half* b_hi, __half* b_lo, int n) {
  plan.n_tiles = (n + 63) / 64;
  const uint64_t rows = (uint64_t)plan.n_tiles * 128;
  bool ok = true;
  ok &= caspr_make_tmap_f16(&plan.tm_act[0][0], a_hi, rows, 512, kBM);
  ok &= caspr_make_tmap_f16(&plan.tm_act[0][1], a_lo, rows, 512, kBM);
  ok &= caspr_make_tmap_f16(&plan.tm_act[1][0], b_hi, rows, 512, kBM);
  ok &= caspr_make_tmap_f16(&plan.tm_act[1][1], b_lo, rows, 512, kBM);
  for (int l = 0; l < 2; ++l) {
    ok &= caspr_make_tmap_f16(&plan.tm_w[l][0], w.hi[l], 512, 512, tcg::w_box_rows());
    ok &= caspr_make_tmap_f16(&plan.tm_w[l][1], w.lo[l], 512, 512, tcg::w_box_rows());
  }
  plan.a_hi = a_hi; plan.a_lo = a_lo; plan.b_hi = b_hi; plan.b_lo = b_lo;
  if (!ok) return CASPR_ELAUNCH;
  return CASPR_OK;
}

int enqueue_layer0(const Plan& plan, const float4* y0, const float4* kbuf, size_t kstride, const float* e,
                   const float* W0, int pt0, int n, int P, int stage, const float* gate, const float* biasf,
                   int ld_hyper, const CnfState* st, int* range_flag, cudaStream_t s) {
  int blocks = (n - pt0 + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CASPR_COUNT(); cnf_tc_layer0_kernel<<<blocks, 256, 0, s>>>(y0, kbuf, kstride, e, W0, pt0, n, P, stage, gate, biasf,
                                                             ld_hyper, st, plan.a_hi, plan.a_lo, range_flag);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int enqueue_mid(const Plan& plan, int layer, int m_tile0, int m_tiles, const float* gate, const float* biasf,
                int ld_hyper, int n, int P, const CnfState* st, const float* W3, float* acc6, int* range_flag,
                int num_sms, cudaStream_t s) {
  const int* skip = &st->done;
  caspr_prof_begin(CASPR_PROF_CNF_FUSED_TC, s);
  CASPR_COUNT();
  cudaError_t err;
  if (layer == 0) {
    CnfEpilogue<false> epi{};
    epi.gate = gate; epi.biasf = biasf; epi.ld_hyper = ld_hyper; epi.n = n; epi.P = P;
    epi.out_hi = plan.b_hi; epi.out_lo = plan.b_lo; epi.range_flag = range_flag;
    err = tcg::launch_gemm(plan.tm_act[0][0], plan.tm_act[0][1], plan.tm_w[0][0], plan.tm_w[0][1], plan.tm_act[1][0],
                           plan.tm_act[1][1], m_tile0, m_tiles, 2, 512 / tcg::kBK, skip, epi, 1, num_sms, s);
  } else {
    CnfEpilogue<true> epi{};
    epi.gate = gate; epi.biasf = biasf; epi.ld_hyper = ld_hyper; epi.n = n; epi.P = P;
    epi.W3 = W3; epi.acc6 = acc6; epi.range_flag = range_flag; epi.pending = false;
    err = tcg::launch_gemm(plan.tm_act[1][0], plan.tm_act[1][1], plan.tm_w[1][0], plan.tm_w[1][1], plan.tm_act[1][0],
                           plan.tm_act[1][1], m_tile0, m_tiles, 2, 512 / tcg::kBK, skip, epi, 1, num_sms, s);
  }
  if (err != cudaSuccess) return CASPR_ELAUNCH;
  caspr_prof_end(CASPR_PROF_CNF_FUSED_TC, s);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int enqueue_last_finish(float* acc6, const float* e, int pt0, int n, int P, const float* gate, const float* biasf,
                        int ld_hyper, int reverse, const CnfState* st, float4* kout, cudaStream_t s) {
  CASPR_COUNT(); cnf_tc_last_finish_kernel<<<ceil_div(n - pt0, 256), 256, 0, s>>>(acc6, e, pt0, n, P, gate, biasf,
                                                                                  ld_hyper, reverse, st, kout);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace cnf_tc
