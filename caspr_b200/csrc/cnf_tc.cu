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
// Two ways to run one dynamics evaluation on this layout:
//   * cnf_fused_eval_kernel (default, further down): ONE launch for layer 0, both H x H layers, the output layer and
//     the divergence; persistent CTA pairs (cta_group::2), the planes between the layers stay in an L2-resident per-CTA
//     scratch.
//   * the round-1 four-kernel path (CASPR_CNF_FUSED=0): cnf_tc_layer0_kernel -> two launches of the shared GEMM skeleton
//     (tc_gemm.cuh) with CnfEpilogue<false> / CnfEpilogue<true> -> cnf_tc_last_finish_kernel; the planes make a round
//     trip through HBM between the launches.  The skeleton is persistent and warp-specialised: TMA producer warp, MMA
//     warp (3 x tcgen05.mma per k-step into one of two 256-column TMEM accumulators), four epilogue warps, so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
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


// ====================================================================================================================
// Fused dynamics evaluation: ONE launch per RK stage (north_star: "a fused RK solver that evaluates the dynamics MLP and
// accumulates the divergence trace ... in one launch per step").  Replaces the four launches
// layer 0 -> H x H GEMM -> H x H GEMM -> output finish, whose [activation ; tangent] fp16 planes made a round trip
// through HBM between them (2.6 GB per evaluation at config 2).
//
// Persistent CTA pairs (cta_group::2), one CTA per SM, 448 threads.  Every CTA streams its own 64-point tiles through
// the whole network; the planes between the layers live in a per-CTA scratch of 512 KB that is re-used in place every
// tile and therefore stays in the 126 MB L2 (148 x 512 KB = 76 MB):
//
//   warps 10-13  layer 0   : RK stage input from (y0, k_j), 3 -> H layer, softplus, tangent W0 e; writes the A planes
//                            of the tile as fp16 hi / lo rows into scratch SA, k-chunk by k-chunk (`sa_full[kc]`)
//   warp 0       producer  : TMA loads of A (SA for layer 1, SB for layer 2) and of this CTA's half of the W tile into
//                            a 2-stage ring (64 KB per stage)
//   warp 1       MMA       : leader CTA only; 3 x tcgen05.mma.cta_group::2 (M 256, N 256, K 16) per k-step
//   warps 2-5, 6-9         : two epilogue groups; BOTH work on every accumulator, group g on its columns 128g..128g+127
//                            (halves the latency of an item's epilogue, which the chase below depends on).  Layer-1
//                            items: gate / bias / softplus / tangent chain rule, fp16 hi / lo split, 16-byte stores
//                            into scratch SB; layer-2 items: the same followed by the fused H -> 3 output layer; the
//                            four partial sums of a tile meet in shared memory and group 1 writes k = (dy, -e.J.e).
//
// Work items of one CTA pair, issued in this order (tile i = the i-th tile pair of the cluster):
//   L1(i).n0, L1(i).n1, L2(i).n0, L2(i).n1, L1(i+1).n0, ...
// The layer-2 products chase the layer-1 epilogues through per-box flags (`sb_full[kc]`: the 64-column box kc of SB
// has landed): L2(i).n0 starts with the four k-chunks L1(i).n0 produced while the epilogue of L1(i).n1 is still
// writing the other four.  Keeping the re-use distance of a scratch line this short (SA and SB are single buffers of
// 256 KB per CTA, 76 MB in total) is what keeps them in L2: a first version that ran L2(i-1) one iteration behind
// L1(i) with a double-buffered SB (114 MB) had a 60 % L2 hit rate and still moved 2.0 GB through HBM per launch.
namespace fused {

constexpr int kThreads = 448;
constexpr int kStages = 3;
constexpr int kATile = tcg::kATile;                               // 16 KB
constexpr int kWTile = tcg::kPairWTile;                           // 16 KB (this CTA's half of the 256-channel W tile)
constexpr int kStageBytes = 2 * kATile + 2 * kWTile;              // 64 KB
constexpr int kSlotFloats = 4 * 64 * 6;                           // partial sums of the output layer: [group][n half][point][6]
constexpr int kSmemBytes = kStages * kStageBytes + kSlotFloats * 4 + 512 + 1024;

struct Params {
  const float4* y0;
  const float4* kbuf;
  size_t kstride;
  const float* e;
  const float* W0;
  const float* W3;
  const float* gate;       // this stage's slot: layer l at + l*H (l = 0..2), output layer at + 3H
  const float* biasf;
  int ld_hyper;
  int n, P, stage, reverse, n_tiles;
  const CnfState* st;
  float4* kout;
  int* range_flag;
  __half* sa_hi;           // scratch SA planes [gridDim.x * 128][512]
  __half* sa_lo;
  __half* sb_hi;           // scratch SB planes, same shape
  __half* sb_lo;
  int dbg_mode;            // timing experiments only (CASPR_CNF_FUSED_DEBUG=2: no SB stores, 3: no split / range)
  long long* debug;        // optional [gridDim.x][8] cycle counters of the producer / MMA threads (CASPR_CNF_FUSED_DEBUG)
};

// 32-byte global store (sm_100: st.global.v8.b32): halves the number of (instruction, line) pairs the LSU handles
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&a)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]),
               "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void softplus_pair(float pre, float& sp, float& dsp) {
  const float z = ex2_approx(fminf(pre, 40.f) * kLog2e);
  const float t = 1.f + z;
  const bool big = pre > 20.f;
  sp = big ? pre : kLn2 * lg2_approx(t);
  dsp = big ? 1.f : z * rcp_approx(t);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
cnf_fused_eval_kernel(const __grid_constant__ CUtensorMap tm_sa_hi, const __grid_constant__ CUtensorMap tm_sa_lo,
                      const __grid_constant__ CUtensorMap tm_sb_hi, const __grid_constant__ CUtensorMap tm_sb_lo,
                      const __grid_constant__ CUtensorMap tm_w1_hi, const __grid_constant__ CUtensorMap tm_w1_lo,
                      const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                      const Params p) {
  if (p.st->done) return;                   // uniform over the grid
  constexpr int H = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* slot = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(slot + kSlotFloats);
  uint64_t* full = bars;                     // [3]  leader's copy is live
  uint64_t* empty = bars + 3;                // [3]
  uint64_t* tfull = bars + 6;                // [2]
  uint64_t* tempty = bars + 8;               // [2]  leader's copy is live
  uint64_t* sa_full = bars + 10;             // [8]  layer-0 warps -> producer, per k-chunk
  uint64_t* sa_free = bars + 18;             //      MMA (commit) -> layer-0 warps
  uint64_t* sb_full = bars + 19;             // [8]  epilogue groups -> producer, per 64-column box of SB
  uint64_t* out_half = bars + 27;            //      n0 epilogue group -> n1 epilogue group
  uint64_t* out_free = bars + 28;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int pairs = (p.n_tiles + 1) >> 1;
  const int n_iter = cluster_id < pairs ? (pairs - cluster_id + n_clusters - 1) / n_clusters : 0;
  const int cta = blockIdx.x;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tm_sa_hi); tc::prefetch_tmap(&tm_sa_lo); tc::prefetch_tmap(&tm_sb_hi); tc::prefetch_tmap(&tm_sb_lo);
    tc::prefetch_tmap(&tm_w1_hi); tc::prefetch_tmap(&tm_w1_lo); tc::prefetch_tmap(&tm_w2_hi); tc::prefetch_tmap(&tm_w2_lo);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);            // the leader's arrive.expect_tx; the peer only contributes bytes
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tfull[s], 1);
      tc::mbar_init(&tempty[s], 512);        // both epilogue groups of both CTAs
    }
    for (int k = 0; k < 8; ++k) {
      tc::mbar_init(&sa_full[k], 128);
      tc::mbar_init(&sb_full[k], 128);       // every thread of the storing epilogue group arrives
    }
    tc::mbar_init(sa_free, 1);
    tc::mbar_init(out_half, 256);            // all partial sums of a tile are in `slot`
    tc::mbar_init(out_free, 128);            // group 1 has consumed them
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_pair(tmem_slot, 512);
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long dbg_empty = 0, dbg_dep = 0, dbg_dep2 = 0;
      const long long dbg_t0 = p.debug ? clock64() : 0;
      auto load_item = [&](const CUtensorMap* a_hi, const CUtensorMap* a_lo, const CUtensorMap* w_hi,
                           const CUtensorMap* w_lo, int a_row, int nh, const uint64_t* chunk_bars, uint32_t chunk_parity,
                           bool chunk_layout64) {
        for (int kk = 0; kk < 8; ++kk) {
          // layer-2 items take the k-chunks in the order their SB boxes are published: the two epilogue groups write
          // boxes (4, 5) and (6, 7) of the n1 half concurrently, so box 6 lands long before box 5
          const int kc = (chunk_bars == sb_full && kk >= 5 && kk <= 6) ? 11 - kk : kk;
          const long long c0 = p.debug ? clock64() : 0;
          tc::mbar_wait(&empty[stage], phase ^ 1);
          const long long c1 = p.debug ? clock64() : 0;
          if (chunk_bars) tc::mbar_wait(const_cast<uint64_t*>(&chunk_bars[kc]), chunk_parity);
          if (p.debug) { const long long c2 = clock64(); dbg_empty += c1 - c0; if (chunk_bars == sa_full) dbg_dep += c2 - c1; else dbg_dep2 += c2 - c1; }
          uint8_t* sb = smem + stage * kStageBytes;
          const uint32_t lead_full = tc::mapa_shared(&full[stage], 0);
          if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * kStageBytes);       // bytes of BOTH CTAs
          if (chunk_layout64) {
            // A from SB: two [128][32] sub-chunk boxes per plane (64-byte swizzle), sub-chunk s of this CTA at rows
            // (cta * 16 + s) * 128 of the [.][32] tensor
            const int r0 = (cta * 16 + 2 * kc) * kBM;
            tc::tma_load_2d_pair(sb, a_hi, lead_full, 0, r0);
            tc::tma_load_2d_pair(sb + kATile / 2, a_hi, lead_full, 0, r0 + kBM);
            tc::tma_load_2d_pair(sb + kATile, a_lo, lead_full, 0, r0);
            tc::tma_load_2d_pair(sb + kATile + kATile / 2, a_lo, lead_full, 0, r0 + kBM);
          } else {
            tc::tma_load_2d_pair(sb, a_hi, lead_full, kc * 64, a_row);
            tc::tma_load_2d_pair(sb + kATile, a_lo, lead_full, kc * 64, a_row);
          }
          tc::tma_load_2d_pair(sb + 2 * kATile, w_hi, lead_full, kc * 64, nh * kBN + (int)rank * (kBN / 2));
          tc::tma_load_2d_pair(sb + 2 * kATile + kWTile, w_lo, lead_full, kc * 64, nh * kBN + (int)rank * (kBN / 2));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      for (int i = 0; i < n_iter; ++i) {
        const uint32_t par = (uint32_t)(i & 1);
        load_item(&tm_sa_hi, &tm_sa_lo, &tm_w1_hi, &tm_w1_lo, cta * kBM, 0, sa_full, par, false);
        load_item(&tm_sa_hi, &tm_sa_lo, &tm_w1_hi, &tm_w1_lo, cta * kBM, 1, nullptr, 0, false);
        load_item(&tm_sb_hi, &tm_sb_lo, &tm_w2_hi, &tm_w2_lo, cta * kBM, 0, sb_full, par, true);
        load_item(&tm_sb_hi, &tm_sb_lo, &tm_w2_hi, &tm_w2_lo, cta * kBM, 1, nullptr, 0, true);
      }
      if (p.debug) {
        p.debug[cta * 24 + 0] = dbg_empty; p.debug[cta * 24 + 1] = dbg_dep; p.debug[cta * 24 + 2] = clock64() - dbg_t0;
        p.debug[cta * 24 + 3] = dbg_dep2;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(2 * kBM, kBN);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      long long dbg_tempty = 0, dbg_full = 0;
      const long long dbg_t0 = p.debug ? clock64() : 0;
      auto mma_item = [&](bool a_layout64) {
        const int buf = it & 1;
        const long long c0 = p.debug ? clock64() : 0;
        tc::mbar_wait(&tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        if (p.debug) dbg_tempty += clock64() - c0;
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kc = 0; kc < 8; ++kc) {
          const long long c1 = p.debug ? clock64() : 0;
          tc::mbar_wait(&full[stage], phase);
          if (p.debug) dbg_full += clock64() - c1;
          tc::fence_after_sync();
          const uint32_t sb = tc::smem_u32(smem + stage * kStageBytes);
          const uint64_t w_hi = tc::make_desc_k128(sb + 2 * kATile);
          const uint64_t w_lo = tc::make_desc_k128(sb + 2 * kATile + kWTile);
          if (a_layout64) {
            // A planes as two [128][32] sub-chunk boxes (64-byte swizzle): K steps 0-1 in the first, 2-3 in the second
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t off = (uint32_t)(ks >> 1) * (kATile / 2);
              const uint64_t a_hi = tc::make_desc_k64(sb + off) + (uint64_t)((ks & 1) * 2);
              const uint64_t a_lo = tc::make_desc_k64(sb + kATile + off) + (uint64_t)((ks & 1) * 2);
              const uint64_t adv = (uint64_t)(ks * 2);
              tc::umma_f16_ss_pair(d_tmem, a_hi, w_hi + adv, idesc, (kc | ks) != 0);
              tc::umma_f16_ss_pair(d_tmem, a_lo, w_hi + adv, idesc, 1);
              tc::umma_f16_ss_pair(d_tmem, a_hi, w_lo + adv, idesc, 1);
            }
          } else {
            const uint64_t a_hi = tc::make_desc_k128(sb);
            const uint64_t a_lo = tc::make_desc_k128(sb + kATile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              tc::umma_f16_ss_pair(d_tmem, a_hi + adv, w_hi + adv, idesc, (kc | ks) != 0);
              tc::umma_f16_ss_pair(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
              tc::umma_f16_ss_pair(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
            }
          }
          tc::umma_commit_pair(&empty[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit_pair(&tfull[buf], 3);
        ++it;
      };
      for (int i = 0; i < n_iter; ++i) {
        mma_item(false);
        mma_item(false);
        tc::umma_commit_pair(sa_free, 3);            // both layer-1 products of tile i have consumed SA
        mma_item(true);
        mma_item(true);
      }
      if (p.debug) {
        p.debug[cta * 24 + 4] = dbg_tempty; p.debug[cta * 24 + 5] = dbg_full; p.debug[cta * 24 + 6] = clock64() - dbg_t0;
      }
    }
  } else if (warp < 10) {
    // ---------------------------------------------------------------------------------------- epilogue groups 0 / 1
    const int grp = (warp - 2) >> 2;                              // owns accumulator `grp`
    const int q = warp & 3;                                       // TMEM lane quadrant
    const int etid = (int)threadIdx.x - 64 - grp * 128;
    const int is_v = lane >> 4;
    const int pl = q * 16 + (lane & 15);                          // point of this lane within the tile
    const int box_row = q * 32 + (lane & 15);
    float range_max = 0.f;
    int pending_box = -1;                                         // 64-column box of SB stored but not yet published
    int it = 0;                                                   // item counter (all roles count alike)
    for (int i = 0; i < n_iter; ++i) {
      for (int sub = 0; sub < 4; ++sub) {
        const bool layer2 = sub >= 2;
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        ++it;
        const int nh = sub & 1;
        const int ti = i;                                         // tile-pair index this item belongs to
        const int tile = 2 * (cluster_id + ti * n_clusters) + (int)rank;
        const int pt = tile * 64 + pl;
        const bool live = pt < p.n;
        const int f = (live ? pt : p.n - 1) / p.P;
        const int col0 = nh * kBN + is_v * 16;
        const float* gp = p.gate + (size_t)f * p.ld_hyper + (layer2 ? 2 * H : H) + col0;
        const float* bp = p.biasf + (size_t)f * p.ld_hyper + (layer2 ? 2 * H : H) + col0;
        const long long ec0 = (p.debug && etid == 0) ? clock64() : 0;
        tc::mbar_wait(&tfull[buf], acc_phase);
        const long long ec1 = (p.debug && etid == 0) ? clock64() : 0;
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN;
        float pa[3] = {0.f, 0.f, 0.f}, pv[3] = {0.f, 0.f, 0.f};
        const int sb_row = cta * kBM;
#pragma unroll 1
        for (int chunk = grp * 4; chunk < grp * 4 + 4; ++chunk) {
          // gate / bias of this chunk: issued before the accumulator read so that the L2 latency overlaps it
          float4 g4[4], b4[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            g4[j4] = *reinterpret_cast<const float4*>(gp + chunk * 32 + j4 * 4);
            b4[j4] = *reinterpret_cast<const float4*>(bp + chunk * 32 + j4 * 4);
          }
          uint32_t r[32];
          tc::tmem_ld_32x32(taddr + chunk * 32, r);
          tc::tmem_ld_wait();
          float ah[16], av[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t send = is_v ? r[j] : r[16 + j];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 16);
            ah[j] = __uint_as_float(is_v ? recv : r[j]);
            av[j] = __uint_as_float(is_v ? r[16 + j] : recv);
          }
          float ho[16], vo[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float g[4] = {g4[j4].x, g4[j4].y, g4[j4].z, g4[j4].w};
            const float b[4] = {b4[j4].x, b4[j4].y, b4[j4].z, b4[j4].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = j4 * 4 + u;
              float sp, dsp;
              if (p.dbg_mode == 4) { sp = fmaf(ah[j], g[u], b[u]); dsp = 1.f; }      // timing experiment: no MUFU
              else softplus_pair(fmaf(ah[j], g[u], b[u]), sp, dsp);
              ho[j] = sp;
              vo[j] = dsp * g[u] * av[j];
            }
          }
          if (layer2) {
            const int col = col0 + chunk * 32;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float4* w4 = reinterpret_cast<const float4*>(p.W3 + c * H + col);
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
            uint32_t hh[8], hl[8], vh[8], vl[8];
            if (p.dbg_mode == 3) {
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) {
                hh[j2] = __float_as_uint(ho[2 * j2]); hl[j2] = __float_as_uint(ho[2 * j2 + 1]);
                vh[j2] = __float_as_uint(vo[2 * j2]); vl[j2] = __float_as_uint(vo[2 * j2 + 1]);
              }
            } else {
#pragma unroll
            for (int j2 = 0; j2 < 8; ++j2) {
              const float h0 = ho[2 * j2] * kActScale, h1 = ho[2 * j2 + 1] * kActScale;
              const float v0 = vo[2 * j2] * kActScale, v1 = vo[2 * j2 + 1] * kActScale;
              if (live) range_max = fmaxf(range_max, fmaxf(fmaxf(fabsf(h0), fabsf(h1)), fmaxf(fabsf(v0), fabsf(v1))));
              split2(h0, h1, hh[j2], hl[j2]);
              split2(v0, v1, vh[j2], vl[j2]);
            }
            }
            // 16-byte stores straight into the SB planes of this CTA (the lines stay in L2 and are re-read by TMA a
            // few microseconds later).  A 64-column box is published (proxy fence + arrive) one chunk LATER, just
            // before the next chunk's stores: its own stores have been performed by then, so the fence does not stall
            // the warp (publishing right after the stores cost ~2.5 k cycles per chunk, measured).
            {
              if (pending_box >= 0) {
                fence_proxy_async_all();                           // generic-proxy stores -> visible to the TMA loads
                tc::mbar_arrive(&sb_full[pending_box]);
                pending_box = -1;
              }
              // SB layout [CTA][sub-chunk = nh*8 + chunk][row][32 fp16]: lanes L / L+16 hold the two 32-byte halves of
              // a row's 64 bytes, the warp's 16 h rows (then its 16 v rows) are one contiguous kilobyte
              const size_t sub_base = ((size_t)(cta * 16 + nh * 8 + chunk) * kBM + box_row) * 32 + is_v * 16;
              __half* sb_hi_p = p.sb_hi + sub_base;
              __half* sb_lo_p = p.sb_lo + sub_base;
              if (p.dbg_mode != 2) {
                st_global_v8(sb_hi_p, hh);
                st_global_v8(sb_lo_p, hl);
                st_global_v8(sb_hi_p + 16 * 32, vh);
                st_global_v8(sb_lo_p + 16 * 32, vl);
              }
              if (chunk & 1) pending_box = nh * 4 + (chunk >> 1);
            }
          }
        }
        if (!layer2) {
          // hand the accumulator back first (below), the last box of the item is published right after
        }
        if (p.debug && etid == 0) {
          const long long ec2 = clock64();
          p.debug[cta * 24 + 8 + grp * 8 + sub] += ec1 - ec0;          // wait for the accumulator, per item kind
          p.debug[cta * 24 + 8 + grp * 8 + 4 + sub] += ec2 - ec1;      // chunk loop, per item kind
        }
        // the accumulator is drained: hand it back to the MMA issuer (leader's barrier, both CTAs arrive)
        tc::fence_before_sync();
        tc::mbar_arrive_cluster(tc::mapa_shared(&tempty[buf], 0));
        if (pending_box >= 0) {                                    // last box of a layer-1 item
          fence_proxy_async_all();
          tc::mbar_arrive(&sb_full[pending_box]);
          pending_box = -1;
        }
        if (layer2) {
          // fused output layer: combine the two column halves of every point's lane pair, then the two n halves
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            pa[c] += __shfl_xor_sync(0xffffffffu, pa[c], 16);
            pv[c] += __shfl_xor_sync(0xffffffffu, pv[c], 16);
          }
          // slot[grp][nh][point][6]; the slots of a tile may be written once group 1 has consumed the previous tile's
          if (nh == 0) tc::mbar_wait(out_free, (uint32_t)((ti & 1) ^ 1));
          if (!is_v) {
            float* sl = slot + ((grp * 2 + nh) * 64 + pl) * 6;
            sl[0] = pa[0]; sl[1] = pa[1]; sl[2] = pa[2]; sl[3] = pv[0]; sl[4] = pv[1]; sl[5] = pv[2];
          }
          if (nh == 1) {
            tc::mbar_arrive(out_half);                             // 256 arrivals: both groups have written both halves
            if (grp == 1) {
              tc::mbar_wait(out_half, (uint32_t)(ti & 1));
              if (!is_v) {
                float a[6];
#pragma unroll
                for (int c = 0; c < 6; ++c)
                  a[c] = (slot[(0 * 64 + pl) * 6 + c] + slot[(1 * 64 + pl) * 6 + c]) +
                         (slot[(2 * 64 + pl) * 6 + c] + slot[(3 * 64 + pl) * 6 + c]);
                if (live) {
                  const float* g = p.gate + (size_t)f * p.ld_hyper + 3 * H;
                  const float* bf = p.biasf + (size_t)f * p.ld_hyper + 3 * H;
                  const float dy0 = fmaf(a[0], g[0], bf[0]);
                  const float dy1 = fmaf(a[1], g[1], bf[1]);
                  const float dy2 = fmaf(a[2], g[2], bf[2]);
                  const float e0 = p.e[3 * (size_t)pt], e1 = p.e[3 * (size_t)pt + 1], e2 = p.e[3 * (size_t)pt + 2];
                  const float div = (g[0] * a[3]) * e0 + (g[1] * a[4]) * e1 + (g[2] * a[5]) * e2;
                  p.kout[pt] = p.reverse ? make_float4(-dy0, -dy1, -dy2, div) : make_float4(dy0, dy1, dy2, -div);
                }
              }
              tc::mbar_arrive(out_free);
            }
          }
        }
      }
    }
    if (range_max > 65504.f) atomicOr(p.range_flag, 1);
  } else {
    // --------------------------------------------------------------------------------------------- layer-0 warps
    const int wl = warp - 10;                                     // 0..3: 16 points each
    const int cg = lane & 7;                                      // 8-channel group inside a 64-channel k-chunk
    const float dt = (float)p.st->dt;
    float range_max = 0.f;
    long long l0_wait = 0;
    for (int i = 0; i < n_iter; ++i) {
      const int tile = 2 * (cluster_id + i * n_clusters) + (int)rank;
      float ys[4][3], ev[4][3];
      int fr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int pt = tile * 64 + wl * 16 + r * 4 + (lane >> 3);
        if (pt >= p.n) pt = p.n - 1;                              // tail: recompute a valid point, nobody reads it
        const float4 y = p.y0[pt];
        ys[r][0] = y.x; ys[r][1] = y.y; ys[r][2] = y.z;
        if (p.stage > 0) {
          float kx[6], ky[6], kz[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (j < p.stage) {
              const float4 kv = p.kbuf[(size_t)j * p.kstride + pt];
              kx[j] = kv.x; ky[j] = kv.y; kz[j] = kv.z;
            } else {
              kx[j] = ky[j] = kz[j] = 0.f;
            }
          }
          ys[r][0] = dopri5::stage_combine(y.x, dt, kx, p.stage - 1);
          ys[r][1] = dopri5::stage_combine(y.y, dt, ky, p.stage - 1);
          ys[r][2] = dopri5::stage_combine(y.z, dt, kz, p.stage - 1);
        }
        ev[r][0] = p.e[3 * (size_t)pt]; ev[r][1] = p.e[3 * (size_t)pt + 1]; ev[r][2] = p.e[3 * (size_t)pt + 2];
        fr[r] = pt / p.P;
      }
      const long long l0c = (p.debug && threadIdx.x == 320) ? clock64() : 0;
      if (i > 0) tc::mbar_wait(sa_free, (uint32_t)((i - 1) & 1)); // layer-1 products of the previous tile are done
      if (p.debug && threadIdx.x == 320) l0_wait += clock64() - l0c;
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc) {
        const int j0 = kc * 64 + cg * 8;
        float wl0[24];
#pragma unroll
        for (int j4 = 0; j4 < 6; ++j4) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.W0 + 3 * j0) + j4);
          wl0[4 * j4] = w4.x; wl0[4 * j4 + 1] = w4.y; wl0[4 * j4 + 2] = w4.z; wl0[4 * j4 + 3] = w4.w;
        }
        float g[8], bf[8];
        int fr_loaded = -1;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r == 2 && kc > 0) {
            // publish the PREVIOUS k-chunk here, half a chunk of compute after its stores were issued: the proxy fence
            // (generic-proxy stores -> visible to the TMA loads) then finds them already performed and does not stall
            fence_proxy_async_all();
            tc::mbar_arrive(&sa_full[kc - 1]);
          }
          // gate / bias of this channel group: the same for every point of a frame, so they are fetched once per
          // k-chunk and only re-fetched for a point of another frame (tiles straddle frames when P % 64 != 0)
          if (r == 0 || fr[r] != fr_loaded) {
            const float4* g4p = reinterpret_cast<const float4*>(p.gate + (size_t)fr[r] * p.ld_hyper + j0);
            const float4* b4p = reinterpret_cast<const float4*>(p.biasf + (size_t)fr[r] * p.ld_hyper + j0);
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const float4 a = g4p[j4], b = b4p[j4];
              g[4 * j4] = a.x; g[4 * j4 + 1] = a.y; g[4 * j4 + 2] = a.z; g[4 * j4 + 3] = a.w;
              bf[4 * j4] = b.x; bf[4 * j4 + 1] = b.y; bf[4 * j4 + 2] = b.z; bf[4 * j4 + 3] = b.w;
            }
            fr_loaded = fr[r];
          }
          uint32_t hh[4], hl[4], vh[4], vl[4];
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            float hv[2], vv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int jj = j2 * 2 + u;
              const float w0 = wl0[3 * jj], w1 = wl0[3 * jj + 1], w2 = wl0[3 * jj + 2];
              const float a = fmaf(w2, ys[r][2], fmaf(w1, ys[r][1], w0 * ys[r][0]));
              const float ta = fmaf(w2, ev[r][2], fmaf(w1, ev[r][1], w0 * ev[r][0]));
              float sp, dsp;
              softplus_pair(fmaf(a, g[jj], bf[jj]), sp, dsp);
              hv[u] = sp * kActScale;
              vv[u] = dsp * g[jj] * ta * kActScale;
              range_max = fmaxf(range_max, fmaxf(fabsf(hv[u]), fabsf(vv[u])));
            }
            split2(hv[0], hv[1], hh[j2], hl[j2]);
            split2(vv[0], vv[1], vh[j2], vl[j2]);
          }
          const int plr = wl * 16 + r * 4 + (lane >> 3);
          const size_t row_h = (size_t)cta * kBM + (plr >> 4) * 32 + (plr & 15);
          const size_t row_v = row_h + 16;
          *reinterpret_cast<uint4*>(p.sa_hi + row_h * H + j0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(p.sa_lo + row_h * H + j0) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
          *reinterpret_cast<uint4*>(p.sa_hi + row_v * H + j0) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
          *reinterpret_cast<uint4*>(p.sa_lo + row_v * H + j0) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
        }
      }
      fence_proxy_async_all();
      tc::mbar_arrive(&sa_full[7]);
    }
    if (range_max > 65504.f) atomicOr(p.range_flag, 1);
    if (p.debug && threadIdx.x == 320) p.debug[cta * 24 + 7] = l0_wait;
  }
  tc::fence_before_sync();
  tc::cluster_sync_all();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace fused

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

int make_plan(Plan& plan, const Weights& w, __half* a_hi, __half* a_lo, __half* b_hi, __half* b_lo, int n,
              void* fused_scratch, int fused_grid) {
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
  plan.fused_grid = 0;
  if (fused_scratch && fused_grid > 0 && tcg::use_pair()) {
    // scratch layout: [SA hi | SA lo | SB hi | SB lo], one [128][512] fp16 tile per CTA and plane
    const size_t tile_bytes = (size_t)128 * 512 * 2;
    char* base = (char*)fused_scratch;
    plan.sa_hi = (__half*)base;
    plan.sa_lo = (__half*)(base + (size_t)fused_grid * tile_bytes);
    __half* sb_hi = (__half*)(base + (size_t)2 * fused_grid * tile_bytes);
    __half* sb_lo = (__half*)(base + (size_t)3 * fused_grid * tile_bytes);
    plan.sb_hi = sb_hi; plan.sb_lo = sb_lo;
    ok &= caspr_make_tmap_f16(&plan.tm_sa[0], plan.sa_hi, (uint64_t)fused_grid * 128, 512, kBM);
    ok &= caspr_make_tmap_f16(&plan.tm_sa[1], plan.sa_lo, (uint64_t)fused_grid * 128, 512, kBM);
    // SB is stored as [CTA][16 sub-chunks of 32 columns][128 rows][32 fp16]: an epilogue warp's 16 rows of a 32-column
    // chunk are ONE contiguous kilobyte (8 full lines per store instruction instead of 32 partial ones); TMA brings a
    // sub-chunk back as a [128][32] box with the 64-byte swizzle
    ok &= caspr_make_tmap_f16_box32(&plan.tm_sb[0], sb_hi, (uint64_t)fused_grid * 16 * 128, 32, kBM);
    ok &= caspr_make_tmap_f16_box32(&plan.tm_sb[1], sb_lo, (uint64_t)fused_grid * 16 * 128, 32, kBM);
    plan.fused_grid = fused_grid;
  }
  if (!ok) return CASPR_ELAUNCH;
  return CASPR_OK;
}

long long* g_debug_buf = nullptr;

bool fused_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CASPR_CNF_FUSED");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode == 1 && tcg::use_pair();
}

int enqueue_fused(const Plan& plan, const float4* y0, const float4* kbuf, size_t kstride, const float* e,
                  const float* W0, const float* W3, int n, int P, int stage, int reverse, const float* gate,
                  const float* biasf, int ld_hyper, const CnfState* st, float4* kout, int* range_flag,
                  cudaStream_t s) {
  if (plan.fused_grid <= 0) return CASPR_EINVAL;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fused::cnf_fused_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             fused::kSmemBytes) != cudaSuccess)
      return CASPR_ELAUNCH;
    attr_set = true;
  }
  fused::Params p;
  p.y0 = y0; p.kbuf = kbuf; p.kstride = kstride; p.e = e; p.W0 = W0; p.W3 = W3; p.gate = gate; p.biasf = biasf;
  p.ld_hyper = ld_hyper; p.n = n; p.P = P; p.stage = stage; p.reverse = reverse; p.n_tiles = (n + 63) / 64; p.st = st;
  p.kout = kout; p.range_flag = range_flag; p.sa_hi = plan.sa_hi; p.sa_lo = plan.sa_lo; p.sb_hi = plan.sb_hi; p.sb_lo = plan.sb_lo;
  p.debug = nullptr;
  p.dbg_mode = 0;
  {
    // CASPR_CNF_FUSED_DEBUG=1: cycle counters of the producer / MMA threads of the LAST launch, printed at exit by
    // tools/fused_debug.py through caspr_cnf_fused_debug_read
    static int dbg = -1;
    if (dbg < 0) { const char* e2 = getenv("CASPR_CNF_FUSED_DEBUG"); dbg = (e2 && e2[0] >= '1' && e2[0] <= '9') ? e2[0] - '0' : 0; }
    p.dbg_mode = dbg;
    if (dbg) {
      if (!g_debug_buf && cudaMalloc(&g_debug_buf, 148 * 24 * sizeof(long long)) != cudaSuccess) return CASPR_ELAUNCH;
      cudaMemsetAsync(g_debug_buf, 0, 148 * 24 * sizeof(long long), s);
      p.debug = g_debug_buf;
    }
  }
  caspr_prof_begin(CASPR_PROF_CNF_EVAL_FUSED, s);
  CASPR_COUNT();
  fused::cnf_fused_eval_kernel<<<plan.fused_grid, fused::kThreads, fused::kSmemBytes, s>>>(
      plan.tm_sa[0], plan.tm_sa[1], plan.tm_sb[0], plan.tm_sb[1], plan.tm_w[0][0], plan.tm_w[0][1], plan.tm_w[1][0],
      plan.tm_w[1][1], p);
  caspr_prof_end(CASPR_PROF_CNF_EVAL_FUSED, s);
  CASPR_CHECK_LAUNCH();
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

long long* fused_debug_buffer() { return g_debug_buf; }

}  // namespace cnf_tc
