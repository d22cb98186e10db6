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
#include "tc_common.cuh"

namespace cnf_tc {

namespace {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 2;
constexpr int kATile = kBM * kBK * 2;                    // 16 KB
constexpr int kWTile = kBN * kBK * 2;                    // 32 KB
constexpr int kStageBytes = 2 * kATile + 2 * kWTile;     // 96 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
constexpr int kThreads = 192;
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

// x (already multiplied by kActScale) -> fp16 hi and fp16 lo with hi + lo ~= x to ~22 bits
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

struct MidParams {
  const float* gate;     // per-frame gate of this layer, pre-multiplied by 1/(act_scale*w_scale)
  const float* biasf;    // per-frame folded bias  b*gate + hyper_bias
  int ld_hyper;
  int n;                 // points
  int P;                 // points per frame
  int n_tiles;           // 64-point tiles
  const CnfState* st;
  __half* out_hi;        // OUT_F16: next layer's planes
  __half* out_lo;
  float* out_h;          // OUT_F32: H, V [n][512]
  float* out_v;
  int* range_flag;
};

template <bool OUT_F32>
__global__ void __launch_bounds__(kThreads, 1)
cnf_tc_mid_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                  MidParams p) {
  if (p.st->done) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* tfull = bars + 2 * kStages;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * 2;                 // (64-point tile, 256-column half)

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tm_a_hi);
    tc::prefetch_tmap(&tm_a_lo);
    tc::prefetch_tmap(&tm_w_hi);
    tc::prefetch_tmap(&tm_w_lo);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tfull[b], 1);
      tc::mbar_init(&tempty[b], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile >> 1, nh = tile & 1;
        for (int kc = 0; kc < 512 / kBK; ++kc) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * kStageBytes;
          tc::mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tc::tma_load_2d(sb, &tm_a_hi, &full[stage], kc * kBK, m_tile * kBM);
          tc::tma_load_2d(sb + kATile, &tm_a_lo, &full[stage], kc * kBK, m_tile * kBM);
          tc::tma_load_2d(sb + 2 * kATile, &tm_w_hi, &full[stage], kc * kBK, nh * kBN);
          tc::tma_load_2d(sb + 2 * kATile + kWTile, &tm_w_lo, &full[stage], kc * kBK, nh * kBN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty[buf], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kc = 0; kc < 512 / kBK; ++kc) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t sb = tc::smem_u32(smem + stage * kStageBytes);
          const uint64_t a_hi = tc::make_desc_k128(sb);
          const uint64_t a_lo = tc::make_desc_k128(sb + kATile);
          const uint64_t w_hi = tc::make_desc_k128(sb + 2 * kATile);
          const uint64_t w_lo = tc::make_desc_k128(sb + 2 * kATile + kWTile);
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);              // 32 bytes per UMMA_K
            tc::umma_f16_ss(d_tmem, a_hi + adv, w_hi + adv, idesc, (kc | ks) != 0);
            tc::umma_f16_ss(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
            tc::umma_f16_ss(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
          }
          tc::umma_commit(&empty[stage]);                         // frees the smem stage when the MMAs finish
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull[buf]);                             // accumulator ready for the epilogue
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int q = warp & 3;                                       // TMEM lane quadrant of this warp
    const int is_v = lane >> 4;
    const int pl = q * 16 + (lane & 15);                          // point within the 64-point tile
    float range_max = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile >> 1, nh = tile & 1;
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int pt = m_tile * 64 + pl;
      const bool live = pt < p.n;
      const int f = (live ? pt : p.n - 1) / p.P;
      const float* gp = p.gate + (size_t)f * p.ld_hyper + nh * kBN + is_v * 16;
      const float* bp = p.biasf + (size_t)f * p.ld_hyper + nh * kBN + is_v * 16;
      tc::mbar_wait(&tfull[buf], acc_phase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN;
#pragma unroll 1
      for (int chunk = 0; chunk < kBN / 32; ++chunk) {
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + chunk * 32, r);
        tc::tmem_ld_wait();
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
        const int col = nh * kBN + chunk * 32 + is_v * 16;
        if (OUT_F32) {
          if (live) {
            float4* oh = reinterpret_cast<float4*>(p.out_h + (size_t)pt * 512 + col);
            float4* ov = reinterpret_cast<float4*>(p.out_v + (size_t)pt * 512 + col);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              oh[j4] = make_float4(ho[4 * j4], ho[4 * j4 + 1], ho[4 * j4 + 2], ho[4 * j4 + 3]);
              ov[j4] = make_float4(vo[4 * j4], vo[4 * j4 + 1], vo[4 * j4 + 2], vo[4 * j4 + 3]);
            }
          }
        } else {
          // rows of the next layer's planes: activation row of this point, tangent row 16 lanes on
          const size_t row_h = (size_t)m_tile * kBM + q * 32 + (lane & 15);
          const size_t row_v = row_h + 16;
          uint32_t hh[8], hl[8], vh[8], vl[8];
#pragma unroll
          for (int j2 = 0; j2 < 8; ++j2) {
            const float h0 = ho[2 * j2] * kActScale, h1 = ho[2 * j2 + 1] * kActScale;
            const float v0 = vo[2 * j2] * kActScale, v1 = vo[2 * j2 + 1] * kActScale;
            if (live) range_max = fmaxf(range_max, fmaxf(fmaxf(fabsf(h0), fabsf(h1)), fmaxf(fabsf(v0), fabsf(v1))));
            split2(h0, h1, hh[j2], hl[j2]);
            split2(v0, v1, vh[j2], vl[j2]);
          }
          uint4* d;
          d = reinterpret_cast<uint4*>(p.out_hi + row_h * 512 + col);
          d[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]); d[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          d = reinterpret_cast<uint4*>(p.out_lo + row_h * 512 + col);
          d[0] = make_uint4(hl[0], hl[1], hl[2], hl[3]); d[1] = make_uint4(hl[4], hl[5], hl[6], hl[7]);
          d = reinterpret_cast<uint4*>(p.out_hi + row_v * 512 + col);
          d[0] = make_uint4(vh[0], vh[1], vh[2], vh[3]); d[1] = make_uint4(vh[4], vh[5], vh[6], vh[7]);
          d = reinterpret_cast<uint4*>(p.out_lo + row_v * 512 + col);
          d[0] = make_uint4(vl[0], vl[1], vl[2], vl[3]); d[1] = make_uint4(vl[4], vl[5], vl[6], vl[7]);
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&tempty[buf]);
    }
    if (!OUT_F32 && range_max > 65504.f) atomicOr(p.range_flag, 1);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

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
                     const float* __restrict__ e, const float* __restrict__ W0, int n, int P, int stage,
                     const float* __restrict__ gate, const float* __restrict__ biasf, int ld_hyper,
                     const CnfState* __restrict__ st, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                     int* __restrict__ range_flag) {
  if (st->done) return;
  constexpr int H = 512;
  __shared__ float sW[H * 3];
  for (int i = threadIdx.x; i < H * 3; i += blockDim.x) sW[i] = W0[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float dt = (float)st->dt;
  float range_max = 0.f;
  for (int pt = blockIdx.x * wpb + (threadIdx.x >> 5); pt < n; pt += gridDim.x * wpb) {
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
    const float* g = gate + (size_t)f * ld_hyper;
    const float* bf = biasf + (size_t)f * ld_hyper;
    const int pl = pt & 63;
    const size_t row_h = (size_t)(pt >> 6) * 128 + (pl >> 4) * 32 + (pl & 15);
    const size_t row_v = row_h + 16;
    // each lane produces 16 consecutive channels: j = lane*16 .. lane*16+15
    uint32_t hh[8], hl[8], vh[8], vl[8];
#pragma unroll
    for (int j2 = 0; j2 < 8; ++j2) {
      float hv[2], vv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane * 16 + j2 * 2 + u;
        const float w0 = sW[3 * j], w1 = sW[3 * j + 1], w2 = sW[3 * j + 2];
        const float a = fmaf(w2, ys[2], fmaf(w1, ys[1], w0 * ys[0]));
        const float ta = fmaf(w2, e2, fmaf(w1, e1, w0 * e0));
        const float gj = g[j];
        const float pre = fmaf(a, gj, bf[j]);
        float sp, dsp;
        if (pre > 20.f) { sp = pre; dsp = 1.f; }
        else { const float z = expf(pre); sp = log1pf(z); dsp = __fdiv_rn(z, __fadd_rn(z, 1.f)); }
        hv[u] = sp * kActScale;
        vv[u] = dsp * gj * ta * kActScale;
        range_max = fmaxf(range_max, fmaxf(fabsf(hv[u]), fabsf(vv[u])));
      }
      split2(hv[0], hv[1], hh[j2], hl[j2]);
      split2(vv[0], vv[1], vh[j2], vl[j2]);
    }
    uint4* d;
    d = reinterpret_cast<uint4*>(out_hi + row_h * H + lane * 16);
    d[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]); d[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
    d = reinterpret_cast<uint4*>(out_lo + row_h * H + lane * 16);
    d[0] = make_uint4(hl[0], hl[1], hl[2], hl[3]); d[1] = make_uint4(hl[4], hl[5], hl[6], hl[7]);
    d = reinterpret_cast<uint4*>(out_hi + row_v * H + lane * 16);
    d[0] = make_uint4(vh[0], vh[1], vh[2], vh[3]); d[1] = make_uint4(vh[4], vh[5], vh[6], vh[7]);
    d = reinterpret_cast<uint4*>(out_lo + row_v * H + lane * 16);
    d[0] = make_uint4(vl[0], vl[1], vl[2], vl[3]); d[1] = make_uint4(vl[4], vl[5], vl[6], vl[7]);
  }
  if (range_max > 65504.f) atomicOr(range_flag, 1);
}

bool g_attr_set = false;

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
    ok &= caspr_make_tmap_f16(&plan.tm_w[l][0], w.hi[l], 512, 512, kBN);
    ok &= caspr_make_tmap_f16(&plan.tm_w[l][1], w.lo[l], 512, 512, kBN);
  }
  plan.a_hi = a_hi; plan.a_lo = a_lo; plan.b_hi = b_hi; plan.b_lo = b_lo;
  if (!ok) return CASPR_ELAUNCH;
  if (!g_attr_set) {
    if (cudaFuncSetAttribute(cnf_tc_mid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) !=
            cudaSuccess ||
        cudaFuncSetAttribute(cnf_tc_mid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) !=
            cudaSuccess)
      return CASPR_ELAUNCH;
    g_attr_set = true;
  }
  return CASPR_OK;
}

int enqueue_layer0(const Plan& plan, const float4* y0, const float4* kbuf, size_t kstride, const float* e,
                   const float* W0, int n, int P, int stage, const float* gate, const float* biasf, int ld_hyper,
                   const CnfState* st, int* range_flag, cudaStream_t s) {
  int blocks = (n + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CASPR_COUNT(); cnf_tc_layer0_kernel<<<blocks, 256, 0, s>>>(y0, kbuf, kstride, e, W0, n, P, stage, gate, biasf,
                                                             ld_hyper, st, plan.a_hi, plan.a_lo, range_flag);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int enqueue_mid(const Plan& plan, int layer, const float* gate, const float* biasf, int ld_hyper, int n, int P,
                const CnfState* st, float* out_h, float* out_v, int* range_flag, int num_sms, cudaStream_t s) {
  MidParams p;
  p.gate = gate; p.biasf = biasf; p.ld_hyper = ld_hyper; p.n = n; p.P = P; p.n_tiles = plan.n_tiles; p.st = st;
  p.out_hi = plan.b_hi; p.out_lo = plan.b_lo; p.out_h = out_h; p.out_v = out_v; p.range_flag = range_flag;
  int grid = plan.n_tiles * 2;
  if (grid > num_sms) grid = num_sms;
  caspr_prof_begin(CASPR_PROF_CNF_FUSED_TC, s);
  CASPR_COUNT();
  if (layer == 0)
    cnf_tc_mid_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(plan.tm_act[0][0], plan.tm_act[0][1], plan.tm_w[0][0],
                                                                plan.tm_w[0][1], p);
  else
    cnf_tc_mid_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(plan.tm_act[1][0], plan.tm_act[1][1], plan.tm_w[1][0],
                                                               plan.tm_w[1][1], p);
  caspr_prof_end(CASPR_PROF_CNF_FUSED_TC, s);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace cnf_tc
