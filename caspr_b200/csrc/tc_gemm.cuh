// Persistent warp-specialised tcgen05 GEMM skeleton with fp32-grade accuracy from three fp16
// products (a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo), shared by the CNF engine (cnf_tc.cu) and the
// encoder's dense layers (gemm_tc.cu).
//
//   D[128 x 256 tile] = sum_k  A[128 x K] . W[256 x K]^T       (both operands K-major fp16 planes)
// for the row tiles m_tile0 .. m_tile0 + m_tiles - 1 (a launch may cover a sub-range of the rows so that
// independent row ranges can be pipelined on different streams).
//
// One CTA per SM, 192 threads:
//   warp 0    TMA producer : per 64-wide k-chunk the A_hi, A_lo (128x64) and W_hi, W_lo (256x64)
//                            tiles, 128-byte swizzle, kStages-deep mbarrier ring (96 KB per stage)
//   warp 1    MMA issuer   : 3 x tcgen05.mma (M=128, N=256, K=16) per k-step into one of two
//                            256-column TMEM accumulators; tcgen05.commit releases the smem stage /
//                            publishes the accumulator
//   warps 2-5 epilogue     : tcgen05.ld 32 columns at a time and hand them to the Epilogue functor
// so the epilogue of tile i overlaps the MMAs of tile i+1.  Tiles are enumerated n-fastest so the
// CTAs working on the same rows run together and share the A tile through L2.
//
// Epilogue functor interface (all calls are made by the 128 epilogue threads):
//   void setup(uint8_t* staging, const CUtensorMap* out_hi, const CUtensorMap* out_lo, int epi_tid);
//        staging: 32 KB of shared memory (two 128 x 64 fp16 boxes) for epilogues that store through TMA
//        (tm_o_hi / tm_o_lo are kernel parameters; epilogues that store directly ignore them)
//   void tile_begin(int m_tile, int n_tile, int quadrant, int lane);
//   void chunk(int chunk_idx, uint32_t (&acc)[32]);   // fp32 bits of columns chunk*32 .. +31 of this thread's row
//   void finish();                                     // once, after the last tile
#pragma once
#include "tc_common.cuh"

namespace tcg {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 2;
constexpr int kATile = kBM * kBK * 2;                    // 16 KB
constexpr int kWTile = kBN * kBK * 2;                    // 32 KB
constexpr int kStageBytes = 2 * kATile + 2 * kWTile;     // 96 KB
constexpr int kStagingBytes = 2 * kBM * 64 * 2;          // epilogue staging: hi and lo boxes of 128 rows x 64 fp16
constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
constexpr int kThreads = 192;

template <class Epilogue>
__global__ void __launch_bounds__(kThreads, 1)
gemm_fp16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                   const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
                   int m_tile0, int m_tiles, int n_tiles, int k_chunks, const int* __restrict__ skip_flag,
                   Epilogue epi, int k_splits = 1) {
  if (skip_flag && *skip_flag) return;       // device-side "solve finished" flag (CNF solver)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* tfull = bars + 2 * kStages;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K (weight gradients): work item = (k split, row tile, column tile); split s covers the k-chunks
  // [s*k_chunks, (s+1)*k_chunks) and the epilogue sees it as row tile m_tile + s*m_tiles (partial outputs stacked).
  const int mn_tiles = m_tiles * n_tiles;
  const int total_tiles = mn_tiles * k_splits;

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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
        const int m_local = mn / n_tiles, n_tile = mn - m_local * n_tiles;
        const int m_tile = m_tile0 + m_local;
        const int kc0 = split * k_chunks;
        for (int kc = 0; kc < k_chunks; ++kc) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * kStageBytes;
          tc::mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tc::tma_load_2d(sb, &tm_a_hi, &full[stage], (kc0 + kc) * kBK, m_tile * kBM);
          tc::tma_load_2d(sb + kATile, &tm_a_lo, &full[stage], (kc0 + kc) * kBK, m_tile * kBM);
          tc::tma_load_2d(sb + 2 * kATile, &tm_w_hi, &full[stage], (kc0 + kc) * kBK, n_tile * kBN);
          tc::tma_load_2d(sb + 2 * kATile + kWTile, &tm_w_lo, &full[stage], (kc0 + kc) * kBK, n_tile * kBN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
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
        for (int kc = 0; kc < k_chunks; ++kc) {
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
          tc::umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&tfull[buf]);
      }
    }
  } else {
    const int q = warp & 3;                                       // TMEM lane quadrant of this warp
    epi.setup(staging, &tm_o_hi, &tm_o_lo, (int)threadIdx.x - 64);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int split = tile / mn_tiles, mn = tile - split * mn_tiles;
      const int m_local = mn / n_tiles, n_tile = mn - m_local * n_tiles;
      const int m_tile = m_tile0 + m_local + split * m_tiles;
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      epi.tile_begin(m_tile, n_tile, q, lane);
      tc::mbar_wait(&tfull[buf], acc_phase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN;
#pragma unroll 1
      for (int chunk = 0; chunk < kBN / 32; ++chunk) {
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + chunk * 32, r);
        tc::tmem_ld_wait();
        epi.chunk(chunk, r);
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&tempty[buf]);
    }
    epi.finish();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// x -> fp16 hi and fp16 lo with hi + lo ~= x to ~22 bits (two values at a time)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tcg
