// Dense per-point building blocks of the TPointNet++ encoder in fp32 SIMT:
// 1x1-conv / Linear (row-major "rows x channels" GEMM with strided operands so that concat
// buffers are written in place), GroupNorm over row-blocks with fused ReLU / max-pool, and a
// few layout kernels.  Replaces the torch Conv1d/GroupNorm/ReLU/max calls of
// caspr/models/pointnet2.py:637-642,677-699,471-481,207-212, pointnet.py:27-46 and
// tpointnet2.py:59-62,99-112.  The tensor-core (tcgen05) variant of the GEMM lives in
// gemm_tc.cu; this file is the exact-fp32 engine and the fallback-free baseline it is
// validated against.
#include "common.cuh"

namespace {

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CASPR_ACT_RELU) return fmaxf(v, 0.f);
  if (act == CASPR_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// ---------------------------------------------------------------------------- linear
// Y[r, n] = act_out(sum_k act_in(X[r,k]) * W[n,k] + bias[n]).
// 256 threads as 16x16; CTA tile BM x BN, k-slab 16; thread tile TM x TN with
// TM = BM/16, TN = BN/16.  Operands are staged k-major in shared memory so the inner loop
// reads are conflict-free broadcasts.
// Optional fused epilogue for the per-ball layers of a set-abstraction scale (pointnet2.py:637-642,
// 677-699 on (B'*M, C, ns)): GroupNorm(16, Cout) over each ball of `ns` consecutive rows, ReLU, and
// the max over the ball, all inside the CTA that computed the rows (a 128-row tile holds 8 or 4 whole
// balls; Cout <= BN so every group's channels are in the tile).
struct GnBallArgs {
  const float* gamma;
  const float* beta;
  float eps;
  int ns;          // rows per ball: 16 or 32
  int relu;
  int write_y;     // store the normalised rows to Y
  float* maxout;   // (balls, Cout) max over each ball, or nullptr
  int ld_max;
};

template <int BM, int BN, bool GN_BALL = false>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
              const float* __restrict__ bias, float* __restrict__ Y, int ldy, int rows, int Cin,
              int Cout, int act_in, int act_out, GnBallArgs gn = GnBallArgs()) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader mapping: each thread loads A elements (r = tid/16 + 16*i, k = tid%16)
  constexpr int A_PER_T = BM * BK / 256;
  constexpr int B_PER_T = BN * BK / 256;
  float ra[A_PER_T], rb[B_PER_T];
  const int lk = tid & 15, lr = tid >> 4;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER_T; ++i) {
      long long r = row0 + lr + 16 * i;
      int k = k0 + lk;
      float v = 0.f;
      if (r < rows && k < Cin) v = X[r * ldx + k];
      if (act_in == CASPR_ACT_RELU) v = fmaxf(v, 0.f);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER_T; ++i) {
      int n = col0 + lr + 16 * i;
      int k = k0 + lk;
      rb[i] = (n < Cout && k < Cin) ? W[(long long)n * ldw + k] : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER_T; ++i) As[buf][lk][lr + 16 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < B_PER_T; ++i) Bs[buf][lk][lr + 16 * i] = rb[i];
  };

  // thread (tx,ty) owns rows  (i/4)*64 + ty*4 + i%4  and columns colmap(j): groups of four
  // consecutive elements per thread so the shared-memory reads are conflict-free LDS.128.
  auto colmap = [&](int j) { return TN >= 4 ? (j / 4) * 64 + tx * 4 + (j & 3) : tx * TN + j; };
  auto rowmap = [&](int i) { return (i / 4) * 64 + ty * 4 + (i & 3); };

  const int nk = (Cin + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], bfrag[TN];
#pragma unroll
      for (int i4 = 0; i4 < TM / 4; ++i4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][i4 * 64 + ty * 4]);
        a[i4 * 4 + 0] = v.x; a[i4 * 4 + 1] = v.y; a[i4 * 4 + 2] = v.z; a[i4 * 4 + 3] = v.w;
      }
      if constexpr (TN >= 4) {
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
          float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][j4 * 64 + tx * 4]);
          bfrag[j4 * 4 + 0] = v.x; bfrag[j4 * 4 + 1] = v.y; bfrag[j4 * 4 + 2] = v.z; bfrag[j4 * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) bfrag[j] = Bs[buf][k][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bfrag[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  if constexpr (GN_BALL) {
    extern __shared__ float dyn_smem[];
    float* T = dyn_smem;                               // BM x (BN+1) pre-normalisation rows
    float* S = dyn_smem + BM * (BN + 1);               // (ball, group) -> mean, rstd
    constexpr int LD = BN + 1;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int c = colmap(j);
        acc[i][j] += (bias && c < Cout) ? bias[c] : 0.f;
        T[rowmap(i) * LD + c] = acc[i][j];
      }
    __syncthreads();
    const int ns = gn.ns;
    const int balls = BM / ns;                          // 8 (ns 16) or 4 (ns 32)
    const int tpp = 256 / (balls * 16);                 // threads per (ball, group): 2 or 4
    const int cpg = Cout / 16;
    const int gsz = ns * cpg;
    {
      const int pair = tid / tpp, sub = tid - pair * tpp;
      const int ball = pair >> 4, g = pair & 15;
      float s1 = 0.f;
      for (int e = sub; e < gsz; e += tpp) {
        const int r = ball * ns + e / cpg, c = g * cpg + e % cpg;
        s1 += T[r * LD + c];
      }
      for (int o = 1; o < tpp; o <<= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      const float mean = s1 / (float)gsz;
      float s2 = 0.f;
      for (int e = sub; e < gsz; e += tpp) {
        const int r = ball * ns + e / cpg, c = g * cpg + e % cpg;
        const float d = T[r * LD + c] - mean;
        s2 += d * d;
      }
      for (int o = 1; o < tpp; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      if (sub == 0) {
        S[2 * pair] = mean;
        S[2 * pair + 1] = 1.0f / sqrtf(s2 / (float)gsz + gn.eps);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int rl = rowmap(i);
      const long long r = row0 + rl;
      const int ball = rl / ns;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int c = colmap(j);
        if (c >= Cout) continue;
        const int pair = ball * 16 + c / cpg;
        float v = (acc[i][j] - S[2 * pair]) * S[2 * pair + 1] * gn.gamma[c] + gn.beta[c];
        if (gn.relu) v = fmaxf(v, 0.f);
        if (gn.write_y && r < rows) Y[r * ldy + c] = v;
        if (gn.maxout) T[rl * LD + c] = v;
      }
    }
    if (gn.maxout) {
      __syncthreads();
      for (int it = tid; it < balls * Cout; it += 256) {
        const int ball = it / Cout, c = it - ball * Cout;
        const long long gball = row0 / ns + ball;
        if (gball * ns >= rows) continue;
        float m = -3.0e38f;
        for (int rr = 0; rr < ns; ++rr) m = fmaxf(m, T[(ball * ns + rr) * LD + c]);
        gn.maxout[gball * gn.ld_max + c] = m;
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long r = row0 + rowmap(i);
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = col0 + colmap(j);
      if (n >= Cout) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      Y[r * ldy + n] = apply_act(v, act_out);
    }
  }
}

// ------------------------------------------------------------------------- GroupNorm
// Small samples (one CTA per sample, whole sample staged in shared memory): the per-ball
// GroupNorm of the set-abstraction MLPs (pointnet2.py:642 on (B'*M, C, ns)).
__global__ void __launch_bounds__(256)
groupnorm_small_kernel(float* __restrict__ X, int ldx, int rows_per_sample, int C, int groups,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       int relu, int write_back, float* __restrict__ maxout, int ld_max) {
  extern __shared__ float tile[];                 // rows_per_sample x C
  __shared__ float s_mean[64], s_rstd[64];
  const int sample = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = rows_per_sample;
  float* base = X + (size_t)sample * R * ldx;
  for (int i = tid; i < R * C; i += 256) {
    int r = i / C, c = i - r * C;
    tile[i] = base[(size_t)r * ldx + c];
  }
  __syncthreads();
  const int cpg = C / groups;
  const int gsz = cpg * R;
  for (int g = warp; g < groups; g += 8) {
    float s = 0.f;
    for (int i = lane; i < gsz; i += 32) {
      int r = i / cpg, c = g * cpg + (i - r * cpg);
      s += tile[r * C + c];
    }
    s = warp_sum(s);
    const float mean = s / (float)gsz;
    float v = 0.f;
    for (int i = lane; i < gsz; i += 32) {
      int r = i / cpg, c = g * cpg + (i - r * cpg);
      float d = tile[r * C + c] - mean;
      v += d * d;
    }
    v = warp_sum(v);
    if (lane == 0) {
      s_mean[g] = mean;
      s_rstd[g] = 1.0f / sqrtf(v / (float)gsz + eps);
    }
  }
  __syncthreads();
  if (write_back) {
    for (int i = tid; i < R * C; i += 256) {
      int r = i / C, c = i - r * C;
      int g = c / cpg;
      float v = (tile[i] - s_mean[g]) * s_rstd[g] * gamma[c] + beta[c];
      if (relu) v = fmaxf(v, 0.f);
      base[(size_t)r * ldx + c] = v;
    }
  }
  if (maxout) {
    for (int c = tid; c < C; c += 256) {
      int g = c / cpg;
      const float mu = s_mean[g], rs = s_rstd[g], ga = gamma[c], be = beta[c];
      float m = -3.0e38f;
      for (int r = 0; r < R; ++r) {
        float v = (tile[r * C + c] - mu) * rs * ga + be;
        if (relu) v = fmaxf(v, 0.f);
        m = fmaxf(m, v);
      }
      maxout[(size_t)sample * ld_max + c] = m;
    }
  }
}

// Tiny samples (rows x C <= 2048 floats): one WARP per sample, eight samples in flight per CTA.  Two
// lanes share a group's statistics (two-pass mean / variance like the kernel above); the tile is staged
// in the warp's slice of shared memory so global traffic stays one coalesced read and one write.
constexpr int kGnWarpTile = 2048;
__global__ void __launch_bounds__(256)
groupnorm_warp_kernel(float* __restrict__ X, int ldx, int samples, int rows_per_sample, int C, int groups,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      int relu, int write_back, float* __restrict__ maxout, int ld_max) {
  extern __shared__ float tiles[];                 // 8 x (R*C) + 8 x 2*groups
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = rows_per_sample;
  const int elems = R * C;
  float* tile = tiles + (size_t)warp * elems;
  float* stat = tiles + (size_t)8 * elems + warp * 2 * groups;     // mean[groups], rstd[groups]
  const int cpg = C / groups;
  const int gsz = cpg * R;
  for (int sample = blockIdx.x * 8 + warp; sample < samples; sample += gridDim.x * 8) {
    float* base = X + (size_t)sample * R * ldx;
    for (int i = lane; i < elems; i += 32) {
      const int r = i / C, c = i - r * C;
      tile[i] = base[(size_t)r * ldx + c];
    }
    __syncwarp();
    // lanes 2g and 2g+1 split the elements of group g (groups <= 16 per pass, loop for more)
    for (int g0 = 0; g0 < groups; g0 += 16) {
      const int g = g0 + (lane >> 1);
      const int half = lane & 1;
      float sum = 0.f;
      if (g < groups)
        for (int i = half; i < gsz; i += 2) {
          const int r = i / cpg, c = g * cpg + (i - r * cpg);
          sum += tile[r * C + c];
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const float mean = sum / (float)gsz;
      float var = 0.f;
      if (g < groups)
        for (int i = half; i < gsz; i += 2) {
          const int r = i / cpg, c = g * cpg + (i - r * cpg);
          const float d = tile[r * C + c] - mean;
          var += d * d;
        }
      var += __shfl_xor_sync(0xffffffffu, var, 1);
      if (g < groups && half == 0) {
        stat[g] = mean;
        stat[groups + g] = 1.0f / sqrtf(var / (float)gsz + eps);
      }
    }
    __syncwarp();
    if (write_back) {
      for (int i = lane; i < elems; i += 32) {
        const int r = i / C, c = i - r * C;
        const int g = c / cpg;
        float v = (tile[i] - stat[g]) * stat[groups + g] * gamma[c] + beta[c];
        if (relu) v = fmaxf(v, 0.f);
        base[(size_t)r * ldx + c] = v;
      }
    }
    if (maxout) {
      for (int c = lane; c < C; c += 32) {
        const int g = c / cpg;
        const float mu = stat[g], rs = stat[groups + g], ga = gamma[c], be = beta[c];
        float m = -3.0e38f;
        for (int r = 0; r < R; ++r) {
          float v = (tile[r * C + c] - mu) * rs * ga + be;
          if (relu) v = fmaxf(v, 0.f);
          m = fmaxf(m, v);
        }
        maxout[(size_t)sample * ld_max + c] = m;
      }
    }
    __syncwarp();
  }
}

// Large samples: pass 1 accumulates per-(sample, group) sum / sum of squares in fp64,
// pass 2 normalises (and max-pools through ordered-uint atomics).
constexpr int kGnRowsPerCta = 64;

__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ X, int ldx, int rows_per_sample, int C, int groups,
                       double* __restrict__ stats) {
  const int sample = blockIdx.y;
  const int r0 = blockIdx.x * kGnRowsPerCta;
  const int r1 = min(rows_per_sample, r0 + kGnRowsPerCta);
  const int cpg = C / groups;
  const float* base = X + (size_t)sample * rows_per_sample * ldx;
  __shared__ double sh_s[64], sh_q[64];
  for (int g = threadIdx.x; g < groups; g += 256) { sh_s[g] = 0.0; sh_q[g] = 0.0; }
  __syncthreads();
  // thread owns a fixed channel stripe so its group index is constant per c
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f, q = 0.f;
    for (int r = r0; r < r1; ++r) {
      float v = base[(size_t)r * ldx + c];
      s += v;
      q = fmaf(v, v, q);
    }
    atomicAdd(&sh_s[c / cpg], (double)s);
    atomicAdd(&sh_q[c / cpg], (double)q);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += 256) {
    atomicAdd(&stats[((size_t)sample * groups + g) * 2 + 0], sh_s[g]);
    atomicAdd(&stats[((size_t)sample * groups + g) * 2 + 1], sh_q[g]);
  }
}

__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(float* __restrict__ X, int ldx, int rows_per_sample, int C, int groups,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       int relu, int write_back, const double* __restrict__ stats,
                       unsigned* __restrict__ maxout_ordered, int ld_max) {
  const int sample = blockIdx.y;
  const int r0 = blockIdx.x * kGnRowsPerCta;
  const int r1 = min(rows_per_sample, r0 + kGnRowsPerCta);
  const int cpg = C / groups;
  const double cnt = (double)cpg * (double)rows_per_sample;
  float* base = X + (size_t)sample * rows_per_sample * ldx;
  __shared__ float s_mean[64], s_rstd[64];
  for (int g = threadIdx.x; g < groups; g += 256) {
    double s = stats[((size_t)sample * groups + g) * 2 + 0];
    double q = stats[((size_t)sample * groups + g) * 2 + 1];
    double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float mu = s_mean[g], rs = s_rstd[g], ga = gamma[c], be = beta[c];
    float m = -3.0e38f;
    for (int r = r0; r < r1; ++r) {
      float v = (base[(size_t)r * ldx + c] - mu) * rs * ga + be;
      if (relu) v = fmaxf(v, 0.f);
      if (write_back) base[(size_t)r * ldx + c] = v;
      m = fmaxf(m, v);
    }
    if (maxout_ordered) atomicMax(&maxout_ordered[(size_t)sample * ld_max + c], float_to_ordered(m));
  }
}

// float4 variants of the two passes (C/groups % 4 == 0, ldx % 4 == 0, 16-byte aligned rows): a warp
// walks one row at a time, lane L owns the channel quads L, L+32, ... whose group index does not depend
// on the row, so statistics accumulate in registers and every load / store is a full 128-byte line.
constexpr int kGnMaxQuadsPerLane = 16;          // C <= 2048

// PROJ: the normalised row additionally feeds a narrow linear layer (<= 4 outputs, e.g. the encoder head's
// 1600 -> 4 T-NOCS regression, tpointnet2.py:108-113) while it is still in registers: out = act(W . relu(v) + b),
// one warp reduction per row, so the normalised activations are never written back nor read a second time.
struct GnProject {
  const float* W;        // (P, C) row-major, 16-byte aligned
  const float* bias;     // (P) or nullptr
  float* out;            // rows x ld_out
  int P, ld_out, act;
};

// NQ: channel quads per lane (ceil(C / 128)); a template parameter so that the per-lane arrays are exactly as long as
// the layer needs (C = 512: 4 quads, 40 registers and full occupancy instead of the 128 of the generic 16-quad version)
template <bool PROJ, int NQ>
__global__ void __launch_bounds__(256, 2)
groupnorm_apply_vec_kernel(float* __restrict__ X, int ldx, int rows_per_sample, int C, int groups,
                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                           int relu, int write_back, const double* __restrict__ stats,
                           unsigned* __restrict__ maxout_ordered, int ld_max, GnProject pj) {
  const int sample = blockIdx.y;
  const int r0 = blockIdx.x * kGnRowsPerCta;
  const int r1 = min(rows_per_sample, r0 + kGnRowsPerCta);
  const int cpg = C / groups, Q = C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double cnt = (double)cpg * (double)rows_per_sample;
  float* base = X + (size_t)sample * rows_per_sample * ldx;
  __shared__ float s_mean[64], s_rstd[64];
  for (int g = threadIdx.x; g < groups; g += 256) {
    double s = stats[((size_t)sample * groups + g) * 2 + 0];
    double q = stats[((size_t)sample * groups + g) * 2 + 1];
    double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  // per-quad scale / shift: v*sc + sh with sc = rstd*gamma, sh = beta - mean*rstd*gamma would change the
  // rounding of (x - mean)*rstd*gamma + beta; keep the reference's operation order instead.
  // (mean, rstd) per channel quad sit in shared memory, not in per-lane arrays: the registers go to the row itself,
  // whose NQ loads are all issued before the first one is consumed (the kernel is otherwise latency-bound: 1.1 TB/s
  // with one load in flight per dependent chain).
  __shared__ float2 s_quad[128 * 4];                              // Q <= 512
  for (int qi = threadIdx.x; qi < Q; qi += 256) {
    const int g = (qi * 4) / cpg;
    s_quad[qi] = make_float2(s_mean[g], s_rstd[g]);
  }
  __syncthreads();
  float4 mx[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) mx[k] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* w4 = reinterpret_cast<const float4*>(pj.W);
  for (int r = r0 + warp; r < r1; r += 8) {
    float4* row = reinterpret_cast<float4*>(base + (size_t)r * ldx);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float4 vv[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int qi = lane + 32 * k;
      vv[k] = qi < Q ? row[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int qi = lane + 32 * k;
      if (qi < Q) {
        float4 v = vv[k];
        const float2 mr = s_quad[qi];
        const float4 ga = __ldg(g4 + qi), be = __ldg(b4 + qi);
        v.x = (v.x - mr.x) * mr.y * ga.x + be.x;
        v.y = (v.y - mr.x) * mr.y * ga.y + be.y;
        v.z = (v.z - mr.x) * mr.y * ga.z + be.z;
        v.w = (v.w - mr.x) * mr.y * ga.w + be.w;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (write_back) row[qi] = v;
        mx[k].x = fmaxf(mx[k].x, v.x); mx[k].y = fmaxf(mx[k].y, v.y);
        mx[k].z = fmaxf(mx[k].z, v.z); mx[k].w = fmaxf(mx[k].w, v.w);
        if (PROJ) {
          const float px = fmaxf(v.x, 0.f), py = fmaxf(v.y, 0.f), pz = fmaxf(v.z, 0.f), pw = fmaxf(v.w, 0.f);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            if (o < pj.P) {
              const float4 w = __ldg(w4 + (size_t)o * Q + qi);
              acc[o] = fmaf(px, w.x, fmaf(py, w.y, fmaf(pz, w.z, fmaf(pw, w.w, acc[o]))));
            }
          }
        }
      }
    }
    if (PROJ) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
      }
      if (lane < pj.P) {
        float t = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        if (pj.bias) t += __ldg(pj.bias + lane);
        pj.out[((size_t)sample * rows_per_sample + r) * pj.ld_out + lane] = apply_act(t, pj.act);
      }
    }
  }
  if (maxout_ordered) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int qi = lane + 32 * k;
      if (qi < Q && r0 + warp < r1) {
        // the running maximum only grows: a plain read that already shows a larger key makes the atomic unnecessary
        // (1280 CTAs x 8 warps share each address at the head's shape; almost all of them lose)
        unsigned* o = maxout_ordered + (size_t)sample * ld_max + qi * 4;
        const unsigned kx = float_to_ordered(mx[k].x), ky = float_to_ordered(mx[k].y),
                       kz = float_to_ordered(mx[k].z), kw = float_to_ordered(mx[k].w);
        if (kx > __ldcg(o + 0)) atomicMax(o + 0, kx);
        if (ky > __ldcg(o + 1)) atomicMax(o + 1, ky);
        if (kz > __ldcg(o + 2)) atomicMax(o + 2, kz);
        if (kw > __ldcg(o + 3)) atomicMax(o + 3, kw);
      }
    }
  }
}

__global__ void fill_u32_kernel(unsigned* p, int samples, int C, int ld, unsigned v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < samples * C) p[(size_t)(i / C) * ld + (i % C)] = v;
}
__global__ void decode_ordered_kernel(unsigned* p, int samples, int C, int ld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < samples * C) {
    size_t o = (size_t)(i / C) * ld + (i % C);
    reinterpret_cast<float*>(p)[o] = ordered_to_float(p[o]);
  }
}

// ----------------------------------------------------------------------- layout ops
__global__ void augment_xyz_kernel(const float* __restrict__ x4, int rows, float* __restrict__ out9) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float4 p = reinterpret_cast<const float4*>(x4)[r];
  float* o = out9 + (size_t)r * 9;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
  o[3] = __fmul_rn(p.x, p.x); o[4] = __fmul_rn(p.y, p.y); o[5] = __fmul_rn(p.z, p.z);
  o[6] = __fmul_rn(p.x, p.z); o[7] = __fmul_rn(p.x, p.y); o[8] = __fmul_rn(p.z, p.y);
}
__global__ void strip_time_kernel(const float* __restrict__ x4, int rows, float* __restrict__ xyz3) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float4 p = reinterpret_cast<const float4*>(x4)[r];
  float* o = xyz3 + (size_t)r * 3;
  o[0] = p.x; o[1] = p.y; o[2] = p.z;
}
__global__ void broadcast_rows_kernel(const float* __restrict__ src, int ld_src, int rows_per_sample,
                                      int C, long long total, float* __restrict__ dst, int ld_dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    long long row = i / C;
    int c = (int)(i - row * C);
    long long s = row / rows_per_sample;
    dst[row * ld_dst + c] = src[s * ld_src + c];
  }
}

template <bool PROJ>
void launch_gn_apply_vec(dim3 grid, cudaStream_t s, float* X, int ldx, int rows_per_sample, int C, int groups,
                         const float* gamma, const float* beta, float eps, int relu, int write_back, const double* stats,
                         unsigned* mo, int ld_max, GnProject pj) {
  const int nq = ceil_div(C, 128);
#define CASPR_GN_VEC(NQ) \
  groupnorm_apply_vec_kernel<PROJ, NQ><<<grid, 256, 0, s>>>(X, ldx, rows_per_sample, C, groups, gamma, beta, eps, relu, \
                                                            write_back, stats, mo, ld_max, pj)
  if (nq <= 4) CASPR_GN_VEC(4);
  else if (nq <= 8) CASPR_GN_VEC(8);
  else if (nq <= 13) CASPR_GN_VEC(13);
  else CASPR_GN_VEC(16);
#undef CASPR_GN_VEC
}

}  // namespace

namespace {
// A handful of rows against a wide layer (per-sequence vectors: the head's global-feature bias, 8 x 1024 -> 1600): the
// tiled kernel would run 13 CTAs for 120 us.  One warp per output channel streams its weight row once, the input rows
// sit in shared memory.
constexpr int kFewRows = 16;
__global__ void __launch_bounds__(256)
linear_fewrows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                      const float* __restrict__ bias, float* __restrict__ Y, int ldy, int rows, int Cin, int Cout,
                      int act_in, int act_out) {
  extern __shared__ float sx[];                                  // rows x Cin
  for (int i = threadIdx.x; i < rows * Cin; i += blockDim.x) {
    const int r = i / Cin, k = i - r * Cin;
    float v = X[(size_t)r * ldx + k];
    sx[i] = act_in == CASPR_ACT_RELU ? fmaxf(v, 0.f) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= Cout) return;
  float acc[kFewRows];
#pragma unroll
  for (int r = 0; r < kFewRows; ++r) acc[r] = 0.f;
  const float* w = W + (size_t)c * ldw;
  for (int k = lane; k < Cin; k += 32) {
    const float wk = __ldg(w + k);
#pragma unroll
    for (int r = 0; r < kFewRows; ++r)
      if (r < rows) acc[r] = fmaf(wk, sx[r * Cin + k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kFewRows; ++r) {
    if (r < rows) {                                              // warp-uniform
      const float v = warp_sum(acc[r]) + (bias ? bias[c] : 0.f);
      if (lane == 0) Y[(size_t)r * ldy + c] = apply_act(v, act_out);
    }
  }
}
}  // namespace

extern "C" int caspr_linear(const float* X, int ldx, const float* W, int ldw, const float* bias,
                            float* Y, int ldy, int rows, int Cin, int Cout, int act_in, int act_out,
                            void* stream) {
  CASPR_REQUIRE(X && W && Y && rows > 0 && Cin > 0 && Cout > 0);
  CASPR_REQUIRE(ldx >= Cin && ldw >= Cin && ldy >= Cout);
  CASPR_REQUIRE(act_in == CASPR_ACT_NONE || act_in == CASPR_ACT_RELU);
  cudaStream_t s = (cudaStream_t)stream;
  if (rows <= kFewRows && Cout >= 64 && (size_t)rows * Cin * sizeof(float) <= 48 * 1024) {
    CASPR_COUNT(); linear_fewrows_kernel<<<ceil_div(Cout, 8), 256, (size_t)rows * Cin * sizeof(float), s>>>(
        X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, act_in, act_out);
    CASPR_CHECK_LAUNCH();
    return CASPR_OK;
  }
  if (Cout > 64) {
    dim3 grid(ceil_div(rows, 128), ceil_div(Cout, 128));
    CASPR_COUNT(); linear_kernel<128, 128><<<grid, 256, 0, s>>>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, act_in, act_out);
  } else if (Cout > 32) {
    dim3 grid(ceil_div(rows, 128), ceil_div(Cout, 64));
    CASPR_COUNT(); linear_kernel<128, 64><<<grid, 256, 0, s>>>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, act_in, act_out);
  } else if (Cout > 16) {
    dim3 grid(ceil_div(rows, 128), ceil_div(Cout, 32));
    CASPR_COUNT(); linear_kernel<128, 32><<<grid, 256, 0, s>>>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, act_in, act_out);
  } else {
    dim3 grid(ceil_div(rows, 128), 1);
    CASPR_COUNT(); linear_kernel<128, 16><<<grid, 256, 0, s>>>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, act_in, act_out);
  }
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

namespace {
template <int BN>
int launch_linear_gn_ball(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy,
                          int rows, int Cin, int Cout, const GnBallArgs& gn, cudaStream_t s) {
  const size_t smem = ((size_t)128 * (BN + 1) + 256) * sizeof(float);
  static bool attr_set = false;            // static + dynamic shared memory exceeds the 48 KB default for BN = 64
  if (!attr_set) {
    if (cudaFuncSetAttribute(linear_kernel<128, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return CASPR_ELAUNCH;
    attr_set = true;
  }
  CASPR_COUNT(); linear_kernel<128, BN, true><<<ceil_div(rows, 128), 256, smem, s>>>(
      X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, CASPR_ACT_NONE, CASPR_ACT_NONE, gn);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
}  // namespace

extern "C" int caspr_linear_gn_ball(const float* X, int ldx, const float* W, int ldw, const float* bias,
                                    const float* gamma, const float* beta, float eps, int rows, int Cin, int Cout,
                                    int ns, int relu, float* Y, int ldy, float* maxout, int ld_max, void* stream) {
  CASPR_REQUIRE(X && W && gamma && beta && rows > 0 && Cin > 0 && Cout > 0 && (Y || maxout));
  CASPR_REQUIRE(ldx >= Cin && ldw >= Cin && (!Y || ldy >= Cout) && (!maxout || ld_max >= Cout));
  CASPR_REQUIRE((ns == 16 || ns == 32) && rows % ns == 0 && Cout % 16 == 0 && Cout <= 64);
  GnBallArgs gn;
  gn.gamma = gamma; gn.beta = beta; gn.eps = eps; gn.ns = ns; gn.relu = relu; gn.write_y = Y != nullptr;
  gn.maxout = maxout; gn.ld_max = ld_max;
  cudaStream_t s = (cudaStream_t)stream;
  if (Cout <= 16) return launch_linear_gn_ball<16>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, gn, s);
  if (Cout <= 32) return launch_linear_gn_ball<32>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, gn, s);
  return launch_linear_gn_ball<64>(X, ldx, W, ldw, bias, Y, ldy, rows, Cin, Cout, gn, s);
}

extern "C" int caspr_groupnorm(float* X, int ldx, int samples, int rows_per_sample, int C, int groups,
                               const float* gamma, const float* beta, float eps, int relu,
                               int write_back, float* maxout, int ld_max, double* stats_ws,
                               int stats_ready, void* stream) {
  CASPR_REQUIRE(X && gamma && beta && samples > 0 && rows_per_sample > 0 && C > 0);
  CASPR_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && ldx >= C);
  CASPR_REQUIRE(write_back || maxout);
  CASPR_REQUIRE(!maxout || ld_max >= C);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t tile_bytes = (size_t)rows_per_sample * C * sizeof(float);
  if (!stats_ready && rows_per_sample * C <= kGnWarpTile) {
    const size_t smem = 8 * tile_bytes + 8 * 2 * groups * sizeof(float);
    if (smem > 48 * 1024) {
      if (cudaFuncSetAttribute(groupnorm_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               80 * 1024) != cudaSuccess)
        return CASPR_EINVAL;
    }
    int blocks = ceil_div(samples, 8);
    if (blocks > 148 * 6) blocks = 148 * 6;
    CASPR_COUNT(); groupnorm_warp_kernel<<<blocks, 256, smem, s>>>(X, ldx, samples, rows_per_sample, C, groups,
                                                                   gamma, beta, eps, relu, write_back, maxout, ld_max);
    CASPR_CHECK_LAUNCH();
    return CASPR_OK;
  }
  if (!stats_ready && rows_per_sample <= 64 && tile_bytes <= 96 * 1024) {
    if (tile_bytes > 48 * 1024) {
      if (cudaFuncSetAttribute(groupnorm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               96 * 1024) != cudaSuccess)
        return CASPR_EINVAL;
    }
    CASPR_COUNT(); groupnorm_small_kernel<<<samples, 256, tile_bytes, s>>>(X, ldx, rows_per_sample, C, groups, gamma,
                                                            beta, eps, relu, write_back, maxout, ld_max);
    CASPR_CHECK_LAUNCH();
    return CASPR_OK;
  }
  CASPR_REQUIRE(stats_ws != nullptr);
  if (!stats_ready && cudaMemsetAsync(stats_ws, 0, (size_t)samples * groups * 2 * sizeof(double), s) != cudaSuccess)
    return CASPR_ELAUNCH;
  dim3 grid(ceil_div(rows_per_sample, kGnRowsPerCta), samples);
  const bool vec = (C / groups) % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)X & 15) == 0 &&
                   ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0 && C <= 128 * kGnMaxQuadsPerLane;
  if (!stats_ready) {
    CASPR_COUNT(); groupnorm_stats_kernel<<<grid, 256, 0, s>>>(X, ldx, rows_per_sample, C, groups, stats_ws);
  }
  CASPR_CHECK_LAUNCH();
  unsigned* mo = reinterpret_cast<unsigned*>(maxout);
  if (mo) {
    CASPR_COUNT(); fill_u32_kernel<<<ceil_div(samples * C, 256), 256, 0, s>>>(mo, samples, C, ld_max, 0u);
    CASPR_CHECK_LAUNCH();
  }
  if (vec) {
    CASPR_COUNT(); launch_gn_apply_vec<false>(grid, s, X, ldx, rows_per_sample, C, groups, gamma, beta, eps, relu, write_back,
                                              stats_ws, mo, ld_max, GnProject{});
  } else {
    CASPR_COUNT(); groupnorm_apply_kernel<<<grid, 256, 0, s>>>(X, ldx, rows_per_sample, C, groups, gamma, beta, eps, relu,
                                               write_back, stats_ws, mo, ld_max);
  }
  CASPR_CHECK_LAUNCH();
  if (mo) {
    CASPR_COUNT(); decode_ordered_kernel<<<ceil_div(samples * C, 256), 256, 0, s>>>(mo, samples, C, ld_max);
    CASPR_CHECK_LAUNCH();
  }
  return CASPR_OK;
}

extern "C" int caspr_groupnorm_project(const float* X, int ldx, int samples, int rows_per_sample, int C, int groups,
                                       const float* gamma, const float* beta, float eps, float* maxout, int ld_max,
                                       const double* stats, const float* W, const float* bias, int P, int act,
                                       float* out, int ld_out, void* stream) {
  CASPR_REQUIRE(X && gamma && beta && stats && W && out && samples > 0 && rows_per_sample > 0 && C > 0);
  CASPR_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && ldx >= C && P >= 1 && P <= 4 && ld_out >= P);
  CASPR_REQUIRE(!maxout || ld_max >= C);
  CASPR_REQUIRE((C / groups) % 4 == 0 && ldx % 4 == 0 && C <= 128 * kGnMaxQuadsPerLane);
  CASPR_REQUIRE((((uintptr_t)X | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)W) & 15) == 0);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(ceil_div(rows_per_sample, kGnRowsPerCta), samples);
  unsigned* mo = reinterpret_cast<unsigned*>(maxout);
  if (mo) {
    CASPR_COUNT(); fill_u32_kernel<<<ceil_div(samples * C, 256), 256, 0, s>>>(mo, samples, C, ld_max, 0u);
    CASPR_CHECK_LAUNCH();
  }
  GnProject pj;
  pj.W = W; pj.bias = bias; pj.out = out; pj.P = P; pj.ld_out = ld_out; pj.act = act;
  CASPR_COUNT(); launch_gn_apply_vec<true>(grid, s, const_cast<float*>(X), ldx, rows_per_sample, C, groups, gamma, beta, eps,
                                           0, 0, stats, mo, ld_max, pj);
  CASPR_CHECK_LAUNCH();
  if (mo) {
    CASPR_COUNT(); decode_ordered_kernel<<<ceil_div(samples * C, 256), 256, 0, s>>>(mo, samples, C, ld_max);
    CASPR_CHECK_LAUNCH();
  }
  return CASPR_OK;
}

extern "C" int caspr_augment_xyz(const float* x4, int rows, float* out9, void* stream) {
  CASPR_REQUIRE(x4 && out9 && rows > 0 && ((uintptr_t)x4 & 15) == 0);
  CASPR_COUNT(); augment_xyz_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(x4, rows, out9);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_strip_time(const float* x4, int rows, float* xyz3, void* stream) {
  CASPR_REQUIRE(x4 && xyz3 && rows > 0 && ((uintptr_t)x4 & 15) == 0);
  CASPR_COUNT(); strip_time_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(x4, rows, xyz3);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_broadcast_rows(const float* src, int ld_src, int samples, int rows_per_sample,
                                    int C, float* dst, int ld_dst, void* stream) {
  CASPR_REQUIRE(src && dst && samples > 0 && rows_per_sample > 0 && C > 0 && ld_src >= C && ld_dst >= C);
  long long total = (long long)samples * rows_per_sample * C;
  int blocks = (int)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  CASPR_COUNT(); broadcast_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows_per_sample, C, total,
                                                                   dst, ld_dst);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
