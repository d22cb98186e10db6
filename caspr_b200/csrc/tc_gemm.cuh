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
#include <stdlib.h>
#include <type_traits>
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

// ------------------------------------------------------------------------------------------------- CTA-pair variant
// Same GEMM with `tcgen05.mma.cta_group::2` (M = 256 over the two SMs of a TPC): a cluster of two CTAs works on the row
// tiles (2p, 2p+1) of one column tile.  Every CTA loads its own A tile and only HALF of the W tile (128 of the 256
// output channels), so the bytes an SM pulls from L2 per k-chunk drop from 96 KB to 64 KB — the single-CTA kernel
// runs at the L2 -> SM delivery limit (148 x 96 KB per 0.83 us), not at the tensor-core limit — and the smaller stage
// leaves room for a 3-deep ring.  Roles per CTA as above; only the leader's warp 1 issues MMAs; smem stages and
// accumulators are released / published to both CTAs with multicast commits; the leader's `full` barrier collects
// the TMA bytes of both CTAs; both CTAs' epilogue threads arrive on the leader's `tempty`.
// Epilogue functors that define `static constexpr bool kReadsTmem = true` get `tile(taddr)` instead of eight
// `chunk()` calls (pair kernel only): they choose their own column grouping (e.g. 24 columns = 4 GroupNorm groups of 6).
template <class E, class = void>
struct epilogue_reads_tmem : std::false_type {};
template <class E>
struct epilogue_reads_tmem<E, std::enable_if_t<E::kReadsTmem>> : std::true_type {};

// Epilogue::kEpilogueGroups = 2: two epilogue warpgroups per CTA, group g drains the accumulator buffer g (every other work
// item) with its own half of the staging buffer.  For layers with a short k-loop the epilogue, not the MMA, bounds a
// tile (per-ball GroupNorm: 27 k cycles per tile against 1-2 k of MMA), and two groups halve that.
template <class E, class = void>
struct epilogue_groups : std::integral_constant<int, 1> {};
template <class E>
struct epilogue_groups<E, std::void_t<decltype(E::kEpilogueGroups)>> : std::integral_constant<int, E::kEpilogueGroups> {};
constexpr int kPairThreadsMax = 64 + 2 * 128;

constexpr int kPairStages = 3;
constexpr int kPairWTile = (kBN / 2) * kBK * 2;                      // 16 KB: this CTA's half of the W tile
constexpr int kPairStageBytes = 2 * kATile + 2 * kPairWTile;        // 64 KB
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kStagingBytes + 1024 + 256;

template <class Epilogue>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreadsMax, 1)
gemm_fp16x3_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                        const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                        const __grid_constant__ CUtensorMap tm_o_hi, const __grid_constant__ CUtensorMap tm_o_lo,
                        int m_tile0, int m_tiles, int n_tiles, int k_chunks, const int* __restrict__ skip_flag,
                        Epilogue epi, int k_splits = 1) {
  if (skip_flag && *skip_flag) return;       // uniform over the grid: both CTAs of a pair leave together
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + kPairStages * kPairStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* full = bars;                     // [kPairStages]  (leader's copy is the live one)
  uint64_t* empty = bars + kPairStages;      // [kPairStages]
  uint64_t* tfull = bars + 2 * kPairStages;  // [2]
  uint64_t* tempty = tfull + 2;              // [2]           (leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int mn_items = m_pairs * n_tiles;
  const int total_items = mn_items * k_splits;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tm_a_hi);
    tc::prefetch_tmap(&tm_a_lo);
    tc::prefetch_tmap(&tm_w_hi);
    tc::prefetch_tmap(&tm_w_lo);
    for (int s = 0; s < kPairStages; ++s) {
      tc::mbar_init(&full[s], 1);            // the leader's arrive.expect_tx; the peer only contributes bytes
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tfull[b], 1);
      tc::mbar_init(&tempty[b], 256);        // the epilogue threads of both CTAs
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_pair(tmem_slot, 512);
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < total_items; item += n_clusters) {
        const int split = item / mn_items, mn = item - split * mn_items;
        const int m_pair = mn / n_tiles, n_tile = mn - m_pair * n_tiles;
        int m_local = 2 * m_pair + (int)rank;
        if (m_local >= m_tiles) m_local = m_tiles - 1;            // odd tail: load a valid tile, result is discarded
        const int m_tile = m_tile0 + m_local;
        const int kc0 = split * k_chunks;
        for (int kc = 0; kc < k_chunks; ++kc) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * kPairStageBytes;
          const uint32_t lead_full = tc::mapa_shared(&full[stage], 0);
          if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * kPairStageBytes);   // bytes of BOTH CTAs
          tc::tma_load_2d_pair(sb, &tm_a_hi, lead_full, (kc0 + kc) * kBK, m_tile * kBM);
          tc::tma_load_2d_pair(sb + kATile, &tm_a_lo, lead_full, (kc0 + kc) * kBK, m_tile * kBM);
          tc::tma_load_2d_pair(sb + 2 * kATile, &tm_w_hi, lead_full, (kc0 + kc) * kBK,
                               n_tile * kBN + (int)rank * (kBN / 2));
          tc::tma_load_2d_pair(sb + 2 * kATile + kPairWTile, &tm_w_lo, lead_full, (kc0 + kc) * kBK,
                               n_tile * kBN + (int)rank * (kBN / 2));
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(2 * kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = cluster_id; item < total_items; item += n_clusters, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        tc::mbar_wait(&tempty[buf], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kc = 0; kc < k_chunks; ++kc) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint32_t sb = tc::smem_u32(smem + stage * kPairStageBytes);
          const uint64_t a_hi = tc::make_desc_k128(sb);
          const uint64_t a_lo = tc::make_desc_k128(sb + kATile);
          const uint64_t w_hi = tc::make_desc_k128(sb + 2 * kATile);
          const uint64_t w_lo = tc::make_desc_k128(sb + 2 * kATile + kPairWTile);
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);              // 32 bytes per UMMA_K
            tc::umma_f16_ss_pair(d_tmem, a_hi + adv, w_hi + adv, idesc, (kc | ks) != 0);
            tc::umma_f16_ss_pair(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
            tc::umma_f16_ss_pair(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
          }
          tc::umma_commit_pair(&empty[stage], 3);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit_pair(&tfull[buf], 3);
      }
    }
  } else {
    constexpr int G = epilogue_groups<Epilogue>::value;
    const int q = warp & 3;                                       // TMEM lane quadrant of this warp
    const int grp = (warp - 2) >> 2;                              // epilogue warpgroup (G == 1: only group 0 exists)
    epi.setup(staging + grp * (kStagingBytes / G), &tm_o_hi, &tm_o_lo, ((int)threadIdx.x - 64) & 127);
    int it = 0;
    for (int item = cluster_id; item < total_items; item += n_clusters, ++it) {
      if (G == 2 && (it & 1) != grp) continue;                    // group g owns accumulator buffer g
      const int split = item / mn_items, mn = item - split * mn_items;
      const int m_pair = mn / n_tiles, n_tile = mn - m_pair * n_tiles;
      const int m_local = 2 * m_pair + (int)rank;
      const bool valid = m_local < m_tiles;                       // warp-uniform
      const int m_tile = m_tile0 + m_local + split * m_tiles;
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (valid) epi.tile_begin(m_tile, n_tile, q, lane);
      tc::mbar_wait(&tfull[buf], acc_phase);
      tc::fence_after_sync();
      if (valid) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN;
        if constexpr (epilogue_reads_tmem<Epilogue>::value) {
          epi.tile(taddr);                                        // the functor issues its own tcgen05.ld
        } else {
#pragma unroll 1
          for (int chunk = 0; chunk < kBN / 32; ++chunk) {
            uint32_t r[32];
            tc::tmem_ld_32x32(taddr + chunk * 32, r);
            tc::tmem_ld_wait();
            epi.chunk(chunk, r);
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive_cluster(tc::mapa_shared(&tempty[buf], 0));
    }
    epi.finish();
  }
  tc::fence_before_sync();
  tc::cluster_sync_all();                    // the leader's MMAs read the peer's smem and write its TMEM until here
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------- host launcher
// CASPR_TC_PAIR=0 selects the single-CTA kernel (kept for comparison); default is the CTA-pair kernel.
inline bool use_pair() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CASPR_TC_PAIR");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  return mode == 1;
}
// rows of the W box the tensor maps of the W planes must be built with
inline uint32_t w_box_rows() { return use_pair() ? kBN / 2 : kBN; }

template <class Epilogue>
inline cudaError_t launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                               const CUtensorMap& w_lo, const CUtensorMap& o_hi, const CUtensorMap& o_lo, int m_tile0,
                               int m_tiles, int n_tiles, int k_chunks, const int* skip_flag, const Epilogue& epi,
                               int k_splits, int num_sms, cudaStream_t s) {
  static bool attr_set = false;           // per Epilogue instantiation
  if (use_pair()) {
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(gemm_fp16x3_pair_kernel<Epilogue>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kPairSmemBytes);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    long long clusters = (long long)((m_tiles + 1) / 2) * n_tiles * k_splits;
    if (clusters > num_sms / 2) clusters = num_sms / 2;
    gemm_fp16x3_pair_kernel<Epilogue><<<(int)(2 * clusters), 64 + 128 * epilogue_groups<Epilogue>::value, kPairSmemBytes,
                                        s>>>(a_hi, a_lo, w_hi, w_lo, o_hi, o_lo, m_tile0, m_tiles, n_tiles, k_chunks,
                                             skip_flag, epi, k_splits);
  } else {
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(gemm_fp16x3_kernel<Epilogue>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBytes);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    long long grid = (long long)m_tiles * n_tiles * k_splits;
    if (grid > num_sms) grid = num_sms;
    gemm_fp16x3_kernel<Epilogue><<<(int)grid, kThreads, kSmemBytes, s>>>(a_hi, a_lo, w_hi, w_lo, o_hi, o_lo, m_tile0,
                                                                        m_tiles, n_tiles, k_chunks, skip_flag, epi,
                                                                        k_splits);
  }
  return cudaGetLastError();
}

// pair kernel only (epilogues with kReadsTmem); returns cudaErrorNotSupported when CASPR_TC_PAIR=0
template <class Epilogue>
inline cudaError_t launch_gemm_pair(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                    const CUtensorMap& w_lo, int m_tiles, int n_tiles, int k_chunks, const Epilogue& epi,
                                    int num_sms, cudaStream_t s) {
  static bool attr_set = false;
  if (!use_pair()) return cudaErrorNotSupported;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_fp16x3_pair_kernel<Epilogue>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kPairSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  long long clusters = (long long)((m_tiles + 1) / 2) * n_tiles;
  if (clusters > num_sms / 2) clusters = num_sms / 2;
  gemm_fp16x3_pair_kernel<Epilogue><<<(int)(2 * clusters), 64 + 128 * epilogue_groups<Epilogue>::value, kPairSmemBytes,
                                      s>>>(a_hi, a_lo, w_hi, w_lo, a_hi, a_lo, 0, m_tiles, n_tiles, k_chunks, nullptr, epi,
                                           1);
  return cudaGetLastError();
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
