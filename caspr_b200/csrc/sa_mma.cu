// One scale of a set-abstraction level in ONE kernel, the three per-ball layers on the tensor cores.
//
// Same contract as sa_fused.cu (group gather -> three [Conv1d(k=1) -> GroupNorm(16) -> ReLU?] layers -> max over the
// ball; PointNet2GroupingLayer + PointNetFeatureExtractor of caspr/models/pointnet2.py:391-401,649-708), for the same
// shapes (SA levels 1-2: at most 64 channels wide, 16 or 32 rows per ball).  sa_fused.cu keeps a grouped row per lane
// and reads every weight as a shared-memory broadcast: 3.0 ms per encode at FMA pipe 30 %, L1/shared 77 %.
//
// Here a warp owns a 32-row tile (one 32-row ball or two 16-row balls) as two m16 tiles of
// mma.sync.m16n8k16 (fp16 inputs, fp32 accumulate).  fp32 accuracy comes from the 3-product split the tcgen05 GEMMs of
// this library use (x = hi + lo in fp16 after a power-of-two scale; hi.hi + lo.hi + hi.lo).  Tiles this small (32 x 64
// x 32) are far below a tcgen05 tile (128 x N x 16 per instruction, operands in shared memory), and the per-ball
// GroupNorm between layers wants the accumulators in registers, so the warp-level MMA is the fitting instruction.
//   * the accumulator fragment of layer i (row g / g+8, columns 8j+2t, 8j+2t+1) IS the A fragment of layer i+1 (two
//     n-tiles = one k16 chunk), so activations never leave registers and never move between lanes;
//   * a GroupNorm group (C/16 = 1, 2 or 4 consecutive channels x the ball's rows) lives in one column pair of one
//     n-tile: statistics are 3 (4) xor-shuffles over the lanes that share t, two-pass, summed pairwise so that padded
//     balls (ns copies of one point) normalise to exactly beta like the reference's fp32 GroupNorm;
//   * the weights sit in shared memory already in B-fragment order, hi and lo planes interleaved: one conflict-free
//     LDS.128 per (n-tile, k-chunk) feeds 6 MMAs;
//   * layer 1 is evaluated as (W x_ref + b) + W (x_row - x_ref), x_ref = the ball's first row: the constant term once
//     per ball in fp32 FMAs, the small differences on the tensor cores (see the comment in the kernel) - the split
//     products then carry less rounding noise into the per-ball GroupNorm than the reference's own fp32 GEMM does;
//   * layer 1 reads its K columns as [features | dx dy dz | 0 pad] (a permutation of the reference's [dx dy dz |
//     features], applied to the weight columns as well) so that gathered feature rows are read as aligned 8-byte pairs.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kSaMmaMaxCin = 99;

struct SaLayer {
  const float *W, *b, *gamma, *beta;
};

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x, y) -> packed fp16 pairs hi and lo with hi + lo = (x, y) to ~22 bits
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// power of two s with m*s in [512, 1024) and its inverse (exact to undo); m = 0 or denormal / huge values clamp
__device__ __forceinline__ void pow2_scale(float m, float& s, float& inv) {
  int e = (__float_as_int(m) >> 23) & 0xff;
  e = min(max(e, 40), 220);
  s = __int_as_float((263 - e) << 23);
  inv = __int_as_float((e - 9) << 23);
}

// three split products of one k16 chunk against NT n-tiles of a fragment-ordered weight block
template <int NT, int KC>
__device__ __forceinline__ void mma_chunk(float (&acc)[2][NT][4], const uint32_t (&ahi)[2][4], const uint32_t (&alo)[2][4],
                                          const uint4* __restrict__ frag, int c, int lane) {
  uint4 f[NT];                                                // {b0 hi, b1 hi, b0 lo, b1 lo}
#pragma unroll
  for (int j = 0; j < NT; ++j) f[j] = frag[(j * KC + c) * 32 + lane];
  // product-major order: the 2*NT accumulators are independent, so consecutive MMAs never wait on each other
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], alo[mt], f[j].x, f[j].y);
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], ahi[mt], f[j].z, f[j].w);
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) mma_f16(acc[mt][j], ahi[mt], f[j].x, f[j].y);
}

// acc * unscale + bias, then GroupNorm(16, 8*NT) over each ball of the tile (+ ReLU), in place.
// Fragment element e of acc[mt][j]: row 16*mt + g + 8*(e>>1), column 8*j + 2*t + (e&1).
template <int NS, int NT, int CTOT, bool RELU>
__device__ __forceinline__ void finish_layer(float (&a)[2][NT][4], float unscale, const float* __restrict__ bias,
                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                             int col0, int t, int bias_mt_stride = 0) {
  constexpr int CPG = CTOT / 16;
  static_assert(CPG == 1 || CPG == 2 || CPG == 4, "group = 1, 2 or 4 channels");
  constexpr int SLOTS = NS == 32 ? 1 : 2;             // balls per tile
  constexpr int SUBS = CPG == 1 ? 2 : 1;              // separate groups inside a thread's column pair
  constexpr float inv_cnt = 1.f / (float)(NS * CPG);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int col = col0 + 8 * j + 2 * t;
    const float2 ga = *reinterpret_cast<const float2*>(gamma + col);
    const float2 be = *reinterpret_cast<const float2*>(beta + col);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float2 bi = *reinterpret_cast<const float2*>(bias + mt * bias_mt_stride + col);
      a[mt][j][0] = fmaf(a[mt][j][0], unscale, bi.x);
      a[mt][j][1] = fmaf(a[mt][j][1], unscale, bi.y);
      a[mt][j][2] = fmaf(a[mt][j][2], unscale, bi.x);
      a[mt][j][3] = fmaf(a[mt][j][3], unscale, bi.y);
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int m0 = NS == 32 ? 0 : s, m1 = NS == 32 ? 2 : s + 1;
#pragma unroll
      for (int h = 0; h < SUBS; ++h) {
        float sum = 0.f;
#pragma unroll
        for (int mt = m0; mt < m1; ++mt) {
          const float p = CPG == 1 ? a[mt][j][h] + a[mt][j][h + 2]
                                   : (a[mt][j][0] + a[mt][j][1]) + (a[mt][j][2] + a[mt][j][3]);
          sum = mt == m0 ? p : sum + p;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (CPG == 4) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const float mean = sum * inv_cnt;
        float var = 0.f;
#pragma unroll
        for (int mt = m0; mt < m1; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (CPG == 1 && (e & 1) != h) continue;
            const float d = a[mt][j][e] - mean;
            var = fmaf(d, d, var);
          }
        }
        var += __shfl_xor_sync(0xffffffffu, var, 4);
        var += __shfl_xor_sync(0xffffffffu, var, 8);
        var += __shfl_xor_sync(0xffffffffu, var, 16);
        if (CPG == 4) var += __shfl_xor_sync(0xffffffffu, var, 1);
        const float rstd = 1.f / sqrtf(var * inv_cnt + eps);
#pragma unroll
        for (int mt = m0; mt < m1; ++mt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (CPG == 1 && (e & 1) != h) continue;
            const float y = fmaf((a[mt][j][e] - mean) * rstd, (e & 1) ? ga.y : ga.x, (e & 1) ? be.y : be.x);
            a[mt][j][e] = RELU ? fmaxf(y, 0.f) : y;
          }
        }
      }
    }
  }
}

// accumulator fragments (post GroupNorm + ReLU, all >= 0) -> scaled hi / lo A fragments of the next layer
template <int NT>
__device__ __forceinline__ float to_operand(const float (&a)[2][NT][4], uint32_t (&hi)[2][NT / 2][4],
                                            uint32_t (&lo)[2][NT / 2][4]) {
  float m = 0.f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) m = fmaxf(m, fabsf(a[mt][j][e]));
  m = warp_max(m);
  float s, inv;
  pow2_scale(m, s, inv);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int c = 0; c < NT / 2; ++c) {
      split2(a[mt][2 * c][0] * s, a[mt][2 * c][1] * s, hi[mt][c][0], lo[mt][c][0]);
      split2(a[mt][2 * c][2] * s, a[mt][2 * c][3] * s, hi[mt][c][1], lo[mt][c][1]);
      split2(a[mt][2 * c + 1][0] * s, a[mt][2 * c + 1][1] * s, hi[mt][c][2], lo[mt][c][2]);
      split2(a[mt][2 * c + 1][2] * s, a[mt][2 * c + 1][3] * s, hi[mt][c][3], lo[mt][c][3]);
    }
  }
  return inv;
}

// weight block (COUT x K source columns, row-major with row stride ldw) -> B fragments in shared memory.
// colmap(kk) gives the source column of operand column kk, or -1 for zero padding.
template <typename ColMap>
__device__ __forceinline__ void fill_fragments(uint4* __restrict__ frag, const float* __restrict__ W, int ldw, int NT,
                                               int KC, float scale, ColMap colmap) {
  for (int i = threadIdx.x; i < NT * KC * 32; i += blockDim.x) {
    const int ln = i & 31, jc = i >> 5;
    const int j = jc / KC, c = jc - j * KC;
    const int n = 8 * j + (ln >> 2), k0 = 16 * c + 2 * (ln & 3);
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int src = colmap(k0 + (q & 1) + 8 * (q >> 1));
      w[q] = src >= 0 ? W[(size_t)n * ldw + src] * scale : 0.f;
    }
    uint4 f;
    split2(w[0], w[1], f.x, f.z);
    split2(w[2], w[3], f.y, f.w);
    frag[i] = f;
  }
}

__device__ __forceinline__ float block_absmax(const float* __restrict__ W, int n, unsigned* slot) {
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(W[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
  return m;
}

template <int NS, int C1, int C2, int C3, int KC1>
__global__ void __launch_bounds__(256, 2)
sa_mma_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, const float* __restrict__ feat,
              int ld_feat, int C, const int32_t* __restrict__ idx, int N, int M, long long balls, SaLayer l1, SaLayer l2,
              SaLayer l3, float eps, const float* __restrict__ in_absmax, float* __restrict__ out, int ld_out) {
  constexpr int NT1 = C1 / 8, NT2 = C2 / 8, NT3 = C3 / 8, KC2 = C1 / 16, KC3 = C2 / 16;
  constexpr int NH = 4;                                        // layer 3 runs 4 n-tiles (32 channels) at a time
  static_assert(NT3 % NH == 0 && (32 % (C3 / 16)) == 0, "layer-3 groups must not straddle the n-tile blocks");
  __shared__ uint4 sF1[NT1 * KC1 * 32];
  __shared__ uint4 sF2[NT2 * KC2 * 32];
  __shared__ uint4 sF3[NT3 * KC3 * 32];
  __shared__ __align__(8) float sP[3 * (C1 + C2 + C3)];       // bias, gamma, beta of the three layers
  __shared__ unsigned sWmax[3];
  __shared__ float sW1t[kSaMmaMaxCin * C1];                   // fp32 layer-1 weights [k][c], reference's column order
  __shared__ __align__(8) float sH[8 * (32 / NS) * C1];       // per warp: layer-1 output of each ball's first row
  const int cin = 3 + C;
  for (int i = threadIdx.x; i < cin * C1; i += blockDim.x) {
    const int k = i / C1, c = i - k * C1;
    sW1t[i] = l1.W[c * cin + k];
  }
  if (threadIdx.x < 3) sWmax[threadIdx.x] = 0u;
  __syncthreads();
  block_absmax(l1.W, C1 * cin, &sWmax[0]);
  block_absmax(l2.W, C2 * C1, &sWmax[1]);
  block_absmax(l3.W, C3 * C2, &sWmax[2]);
  __syncthreads();
  float sw1, iw1, sw2, iw2, sw3, iw3;
  pow2_scale(__uint_as_float(sWmax[0]), sw1, iw1);
  pow2_scale(__uint_as_float(sWmax[1]), sw2, iw2);
  pow2_scale(__uint_as_float(sWmax[2]), sw3, iw3);
  fill_fragments(sF1, l1.W, cin, NT1, KC1, sw1, [=](int kk) { return kk < C ? 3 + kk : (kk < C + 3 ? kk - C : -1); });
  fill_fragments(sF2, l2.W, C1, NT2, KC2, sw2, [](int kk) { return kk; });
  fill_fragments(sF3, l3.W, C2, NT3, KC3, sw3, [](int kk) { return kk; });
  float* sB1 = sP; float* sG1 = sB1 + C1; float* sE1 = sG1 + C1;
  float* sB2 = sE1 + C1; float* sG2 = sB2 + C2; float* sE2 = sG2 + C2;
  float* sB3 = sE2 + C2; float* sG3 = sB3 + C3; float* sE3 = sG3 + C3;
  for (int i = threadIdx.x; i < C1; i += blockDim.x) { sB1[i] = l1.b[i]; sG1[i] = l1.gamma[i]; sE1[i] = l1.beta[i]; }
  for (int i = threadIdx.x; i < C2; i += blockDim.x) { sB2[i] = l2.b[i]; sG2[i] = l2.gamma[i]; sE2[i] = l2.beta[i]; }
  for (int i = threadIdx.x; i < C3; i += blockDim.x) { sB3[i] = l3.b[i]; sG3[i] = l3.gamma[i]; sE3[i] = l3.beta[i]; }
  __syncthreads();

  float sa1 = 1.f, ia1 = 1.f;
  if (in_absmax) pow2_scale(2.f * __ldg(in_absmax), sa1, ia1);   // operands are differences of two bounded entries
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long tiles = (balls * NS + 31) / 32;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const bool vec2 = (ld_feat % 2 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 7) == 0);
  const bool out2 = (ld_out % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);

  // source row (b*N + idx) of the thread's four rows of tile wt_: tile row g + 8q, q = 0..3 (m-tile q>>1, half q&1)
  auto gather_rows = [&](long long wt_, unsigned (&sr)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long rowid = wt_ * 32 + g + 8 * q;
      long long ball = rowid / NS;
      if (ball >= balls) ball = balls - 1;                     // idle rows shadow the last ball (never stored)
      sr[q] = (unsigned)(ball / M) * (unsigned)N + (unsigned)idx[ball * NS + (int)(rowid % NS)];
    }
  };
  const long long wt0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned srow[4] = {0u, 0u, 0u, 0u};
  if (wt0 < tiles) gather_rows(wt0, srow);
  for (long long wt = wt0; wt < tiles; wt += warps) {
    long long ballq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long ball = (wt * 32 + g + 8 * q) / NS;
      ballq[q] = ball < balls ? ball : balls - 1;
    }
    // the next tile's indices are fetched now: the dependent chain idx -> row address -> gather is off the critical path
    unsigned srow_next[4] = {0u, 0u, 0u, 0u};
    if (wt + warps < tiles) gather_rows(wt + warps, srow_next);
    // Layer 1 is split as  W x_r + b = (W x_ref + b) + W (x_r - x_ref)  with x_ref the ball's first row: the first
    // term once per ball in plain fp32 (lane = output channel), the second on the tensor cores.  Rows of a ball are
    // neighbours, so the differences are small and the per-ball GroupNorm - which divides by the spread of the ball,
    // 20-50x below the magnitude of the entries at radius 0.02-0.05 - no longer amplifies the rounding of the products
    // of the full-size entries.  A padded ball has all differences exactly 0.
    constexpr int BPT = 32 / NS;
    float* h0 = sH + (threadIdx.x >> 5) * BPT * C1;
    __syncwarp();
#pragma unroll
    for (int item = lane; item < BPT * C1; item += 32) {
      const int sl = item / C1, ch = item - sl * C1;
      long long ball = wt * BPT + sl;
      if (ball >= balls) ball = balls - 1;
      // row 0 of ball `sl` is tile row 16*sl*(NS==16): lane 0 holds it as q = 0 (or 2)
      const unsigned r_a = __shfl_sync(0xffffffffu, srow[0], 0);
      const unsigned r_b = NS == 32 ? r_a : __shfl_sync(0xffffffffu, srow[2], 0);
      const unsigned rs = sl == 0 ? r_a : r_b;
      float acc4[4] = {sB1[ch], 0.f, 0.f, 0.f};               // four chains: the sum is latency-, not throughput-bound
#pragma unroll
      for (int d = 0; d < 3; ++d)
        acc4[d + 1] = sW1t[d * C1 + ch] * (xyz[(size_t)rs * 3 + d] - new_xyz[ball * 3 + d]);
      const float* fr = feat + (size_t)rs * ld_feat;
      if (vec2 && C % 4 == 0 && ld_feat % 4 == 0 && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0)) {
#pragma unroll 4
        for (int k = 0; k < C; k += 4) {
          const float4 f = *reinterpret_cast<const float4*>(fr + k);
          acc4[0] = fmaf(sW1t[(3 + k) * C1 + ch], f.x, acc4[0]);
          acc4[1] = fmaf(sW1t[(4 + k) * C1 + ch], f.y, acc4[1]);
          acc4[2] = fmaf(sW1t[(5 + k) * C1 + ch], f.z, acc4[2]);
          acc4[3] = fmaf(sW1t[(6 + k) * C1 + ch], f.w, acc4[3]);
        }
      } else {
        for (int k = 0; k < C; ++k) acc4[k & 3] = fmaf(sW1t[(3 + k) * C1 + ch], fr[k], acc4[k & 3]);
      }
      const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      // GroupNorm is invariant to a shift of the whole group: take out the mean of the group's constant terms, so
      // that what is normalised (h0 - m0) + W (x_r - x_ref) is of the size of the ball's spread, not of the entries
      constexpr int CPG1 = C1 / 16;
      float m0 = acc;
      if (CPG1 >= 2) m0 += __shfl_xor_sync(0xffffffffu, m0, 1);
      if (CPG1 == 4) m0 += __shfl_xor_sync(0xffffffffu, m0, 2);
      h0[item] = acc - m0 * (1.f / (float)CPG1);
    }
    __syncwarp();
    // operand column kk of row q: [features | dx dy dz | 0]
    auto column = [&](int q, int kk) -> float {
      if (kk < C) return feat[(size_t)srow[q] * ld_feat + kk];
      const int d = kk - C;
      if (d < 3) return xyz[(size_t)srow[q] * 3 + d] - new_xyz[ballq[q] * 3 + d];
      return 0.f;
    };
    auto load_chunk = [&](int c, float (&v)[4][4]) {
      const int k0 = 16 * c + 2 * t;
      if (vec2 && 16 * c + 16 <= C) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* fr = feat + (size_t)srow[q] * ld_feat + k0;
          const float2 u = *reinterpret_cast<const float2*>(fr);
          const float2 w = *reinterpret_cast<const float2*>(fr + 8);
          v[q][0] = u.x; v[q][1] = u.y; v[q][2] = w.x; v[q][3] = w.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[q][0] = column(q, k0); v[q][1] = column(q, k0 + 1);
          v[q][2] = column(q, k0 + 8); v[q][3] = column(q, k0 + 9);
        }
      }
      // minus the ball's first row (tile row 0, or 16 for the second 16-row ball): lanes 0..3 hold it (g = 0)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float r0 = __shfl_sync(0xffffffffu, v[0][e], t);
        const float r1 = NS == 32 ? r0 : __shfl_sync(0xffffffffu, v[2][e], t);
        v[0][e] -= r0; v[1][e] -= r0; v[2][e] -= r1; v[3][e] -= r1;
      }
    };

    // ---- layer 1
    float a1[2][NT1][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < NT1; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) a1[mt][j][e] = 0.f;
    float cur[4][4];
    load_chunk(0, cur);
#pragma unroll
    for (int c = 0; c < KC1; ++c) {
      float nxt[4][4];
      if (c + 1 < KC1) load_chunk(c + 1, nxt);
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        split2(cur[2 * mt][0] * sa1, cur[2 * mt][1] * sa1, ahi[mt][0], alo[mt][0]);
        split2(cur[2 * mt + 1][0] * sa1, cur[2 * mt + 1][1] * sa1, ahi[mt][1], alo[mt][1]);
        split2(cur[2 * mt][2] * sa1, cur[2 * mt][3] * sa1, ahi[mt][2], alo[mt][2]);
        split2(cur[2 * mt + 1][2] * sa1, cur[2 * mt + 1][3] * sa1, ahi[mt][3], alo[mt][3]);
      }
      mma_chunk<NT1, KC1>(a1, ahi, alo, sF1, c, lane);
      if (c + 1 < KC1) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int e = 0; e < 4; ++e) cur[q][e] = nxt[q][e];
      }
    }
    finish_layer<NS, NT1, C1, true>(a1, ia1 * iw1, h0, sG1, sE1, eps, 0, t, NS == 32 ? 0 : C1);

    // ---- layer 2
    uint32_t h2[2][KC2][4], l2o[2][KC2][4];
    const float ia2 = to_operand<NT1>(a1, h2, l2o);
    float a2[2][NT2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < NT2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) a2[mt][j][e] = 0.f;
#pragma unroll
    for (int c = 0; c < KC2; ++c) {
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { ahi[mt][e] = h2[mt][c][e]; alo[mt][e] = l2o[mt][c][e]; }
      mma_chunk<NT2, KC2>(a2, ahi, alo, sF2, c, lane);
    }
    finish_layer<NS, NT2, C2, true>(a2, ia2 * iw2, sB2, sG2, sE2, eps, 0, t);

    // ---- layer 3, 32 output channels at a time: GroupNorm (no ReLU) and the max over the ball per block
    uint32_t h3[2][KC3][4], l3o[2][KC3][4];
    const float ia3 = to_operand<NT2>(a2, h3, l3o);
#pragma unroll 1
    for (int nb = 0; nb < NT3 / NH; ++nb) {
      float a3[2][NH][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < NH; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) a3[mt][j][e] = 0.f;
#pragma unroll
      for (int c = 0; c < KC3; ++c) {
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) { ahi[mt][e] = h3[mt][c][e]; alo[mt][e] = l3o[mt][c][e]; }
        mma_chunk<NH, KC3>(a3, ahi, alo, sF3 + nb * NH * KC3 * 32, c, lane);
      }
      finish_layer<NS, NH, C3, false>(a3, ia3 * iw3, sB3, sG3, sE3, eps, nb * NH * 8, t);
      // max over each ball's rows; afterwards every lane of a t-column holds it, lane g == j stores n-tile j
#pragma unroll
      for (int s = 0; s < (NS == 32 ? 1 : 2); ++s) {
        float2 mine = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NH; ++j) {
          float mx, my;
          if (NS == 32) {
            mx = fmaxf(fmaxf(a3[0][j][0], a3[0][j][2]), fmaxf(a3[1][j][0], a3[1][j][2]));
            my = fmaxf(fmaxf(a3[0][j][1], a3[0][j][3]), fmaxf(a3[1][j][1], a3[1][j][3]));
          } else {
            mx = fmaxf(a3[s][j][0], a3[s][j][2]);
            my = fmaxf(a3[s][j][1], a3[s][j][3]);
          }
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o));
          }
          if (g == j) mine = make_float2(mx, my);
        }
        const long long ball = NS == 32 ? wt : wt * 2 + s;
        if (g < NH && ball < balls) {
          float* o = out + (size_t)ball * ld_out + nb * NH * 8 + 8 * g + 2 * t;
          if (out2) *reinterpret_cast<float2*>(o) = mine;
          else { o[0] = mine.x; o[1] = mine.y; }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) srow[q] = srow_next[q];
  }
}

// max(|feat|, 2 max|xyz|) >= every entry of a gathered row [xyz[idx] - centre | feat[idx]]
__global__ void sa_absmax_kernel(const float* __restrict__ xyz, long long n_xyz, const float* __restrict__ feat,
                                 long long rows, int C, int ld_feat, unsigned* __restrict__ out) {
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = i0; i < n_xyz; i += stride) m = fmaxf(m, 2.f * fabsf(xyz[i]));
  const long long n_feat = rows * C;
  for (long long i = i0; i < n_feat; i += stride) {
    const long long r = i / C;
    m = fmaxf(m, fabsf(feat[r * ld_feat + (i - r * C)]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out, m);
}

template <int NS, int C1, int C2, int C3, int KC1>
int launch_sa_mma(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C, const int32_t* idx, int N,
                  int M, long long balls, const SaLayer& l1, const SaLayer& l2, const SaLayer& l3, float eps,
                  const float* in_absmax, float* out, int ld_out, cudaStream_t s) {
  const long long tiles = (balls * NS + 31) / 32;
  long long blocks = (tiles + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;                      // persistent: the weight fragments are built once per CTA
  CASPR_COUNT(); sa_mma_kernel<NS, C1, C2, C3, KC1><<<(int)blocks, 256, 0, s>>>(
      xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls, l1, l2, l3, eps, in_absmax, out, ld_out);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace

extern "C" int caspr_sa_mma_supported(int ns, int Cin, int C1, int C2, int C3) {
  if (Cin != 9 && Cin != 99) return 0;
  return (ns == 16 && Cin == 9 && C1 == 16 && C2 == 16 && C3 == 32) ||
         (ns == 32 && C1 == 32 && C2 == 32 && C3 == 64) ||
         (ns == 16 && Cin == 99 && C1 == 32 && C2 == 32 && C3 == 64);
}

extern "C" int caspr_sa_absmax(const float* xyz, const float* feat, int ld_feat, int C, int B, int N, float* absmax,
                               void* stream) {
  CASPR_REQUIRE(xyz && absmax && B > 0 && N > 0 && C >= 0 && (C == 0 || (feat && ld_feat >= C)));
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(absmax, 0, sizeof(float), s) != cudaSuccess) return CASPR_ELAUNCH;
  const long long rows = (long long)B * N;
  long long work = rows * (C > 3 ? C : 3);
  int blocks = (int)((work + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  CASPR_COUNT(); sa_absmax_kernel<<<blocks, 256, 0, s>>>(xyz, rows * 3, feat, rows, C, ld_feat,
                                                         reinterpret_cast<unsigned*>(absmax));
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_sa_mma(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                            const int32_t* idx, int B, int N, int M, int ns,
                            const float* W1, const float* b1, const float* g1, const float* e1, int C1,
                            const float* W2, const float* b2, const float* g2, const float* e2, int C2,
                            const float* W3, const float* b3, const float* g3, const float* e3, int C3,
                            float eps, const float* in_absmax, float* out, int ld_out, void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && idx && out && B > 0 && N > 0 && M > 0 && C > 0 && feat && ld_feat >= C);
  CASPR_REQUIRE(W1 && b1 && g1 && e1 && W2 && b2 && g2 && e2 && W3 && b3 && g3 && e3 && ld_out >= C3);
  CASPR_REQUIRE(caspr_sa_mma_supported(ns, 3 + C, C1, C2, C3));
  CASPR_REQUIRE((unsigned long long)B * N * (unsigned long long)(ld_feat > 3 ? ld_feat : 3) < (1ull << 32));
  const SaLayer l1 = {W1, b1, g1, e1}, l2 = {W2, b2, g2, e2}, l3 = {W3, b3, g3, e3};
  const long long balls = (long long)B * M;
  cudaStream_t s = (cudaStream_t)stream;
#define CASPR_SA_MMA(NS, A, Bc, Cc, KC) \
  return launch_sa_mma<NS, A, Bc, Cc, KC>(xyz, new_xyz, feat, ld_feat, C, idx, N, M, balls, l1, l2, l3, eps, in_absmax, \
                                          out, ld_out, s)
  if (C == 6) {
    if (ns == 16) CASPR_SA_MMA(16, 16, 16, 32, 1);
    CASPR_SA_MMA(32, 32, 32, 64, 1);
  }
  if (ns == 16) CASPR_SA_MMA(16, 32, 32, 64, 7);
  CASPR_SA_MMA(32, 32, 32, 64, 7);
#undef CASPR_SA_MMA
}
