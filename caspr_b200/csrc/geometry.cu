// Geometry kernels of the TPointNet++ encoder: farthest-point sampling, ball query,
// grouping, three-nearest-neighbour search and inverse-distance interpolation.
//
// They replace the Kaolin CUDA ops the reference imports at
// caspr/models/pointnet2.py:7-10 (call sites :384-387, :391, :514, :519).  None of these
// is HBM-bound: a 2048-point cloud is 24 KB and lives in shared memory / registers for the
// whole kernel; FPS is bound by its chain of M dependent arg-max reductions, ball query and
// three_nn by the per-centre scan.  Index results are bit-exact w.r.t. the canonical
// arithmetic declared in oracle/pointnet2_ops.py (see sqdist_canonical in common.cuh).
#include "common.cuh"

namespace {

constexpr float kFpsSkipMag = 1e-3f;   // upstream `if (mag <= 1e-3) continue;`
constexpr float kFpsInit = 1e10f;

// One CTA per cloud.  Thread t owns points t, t+THREADS, ... (PPT of them) in registers
// together with their running min-distance; the cloud is mirrored in shared memory (SoA)
// only to fetch the coordinates of the last pick.  Each of the M-1 dependent iterations
// costs one distance update per owned point, an arg-max over (distance bits, lowest index) as two
// redux.sync per stage and ONE __syncthreads (the cross-warp stage is double-buffered).
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS)
fps_kernel(const float* __restrict__ xyz, int N, int M, int32_t* __restrict__ idx,
           float* __restrict__ new_xyz) {
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + N;
  float* sz = sy + N;
  int* spick = reinterpret_cast<int*>(sz + N);                 // M picks
  __shared__ unsigned long long swarp[2][THREADS / 32];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + (size_t)b * N * 3;
  // coalesced load of the AoS cloud into SoA shared memory
  for (int i = tid; i < N * 3; i += THREADS) {
    float v = p[i];
    int k = i / 3, c = i - 3 * k;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], temp[PPT];
  bool live[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    int k = i * THREADS + tid;
    bool in = k < N;
    px[i] = in ? sx[k] : 0.f;
    py[i] = in ? sy[k] : 0.f;
    pz[i] = in ? sz[k] : 0.f;
    float mag = __fadd_rn(__fadd_rn(__fmul_rn(px[i], px[i]), __fmul_rn(py[i], py[i])),
                          __fmul_rn(pz[i], pz[i]));
    live[i] = in && !(mag <= kFpsSkipMag);
    temp[i] = kFpsInit;
  }

  int last = 0;
  if (tid == 0) spick[0] = 0;
  for (int j = 1; j < M; ++j) {
    const float lx = sx[last], ly = sy[last], lz = sz[last];
    // arg-max with ties to the lowest index = max of the 64-bit key (distance bits | ~index), taken as two 32-bit
    // warp reductions (redux.sync): the largest distance, then the lowest index among the lanes that hold it.
    // Distances are >= 0, so their bit patterns order like the values.  0xffffffff = "no candidate".
    unsigned bd = 0u, bk = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (live[i]) {
        float d = sqdist_canonical(px[i], py[i], pz[i], lx, ly, lz);
        float d2 = fminf(d, temp[i]);
        temp[i] = d2;
        const unsigned k = (unsigned)(i * THREADS + tid), db = __float_as_uint(d2);
        if (db > bd || (db == bd && k < bk)) { bd = db; bk = k; }
      }
    }
    unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    unsigned wk = __reduce_min_sync(0xffffffffu, bd == wd ? bk : 0xffffffffu);
    if (THREADS > 32) {
      if (lane == 0) swarp[j & 1][warp] = ((unsigned long long)wd << 32) | wk;
      __syncthreads();
      const unsigned long long v = lane < THREADS / 32 ? swarp[j & 1][lane] : 0xffffffffull;
      const unsigned vd = (unsigned)(v >> 32), vk = (unsigned)v;
      wd = __reduce_max_sync(0xffffffffu, vd);
      wk = __reduce_min_sync(0xffffffffu, vd == wd ? vk : 0xffffffffu);
    }
    last = wk == 0xffffffffu ? 0 : (int)wk;
    if (tid == 0) spick[j] = last;
  }
  __syncthreads();
  for (int j = tid; j < M; j += THREADS) {
    int k = spick[j];
    idx[(size_t)b * M + j] = k;
    if (new_xyz) {
      float* o = new_xyz + ((size_t)b * M + j) * 3;
      o[0] = sx[k];
      o[1] = sy[k];
      o[2] = sz[k];
    }
  }
}

template <int THREADS, int PPT>
int launch_fps(const float* xyz, int B, int N, int M, int32_t* idx, float* new_xyz, cudaStream_t s) {
  size_t smem = (size_t)N * 3 * sizeof(float) + (size_t)M * sizeof(int);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(fps_kernel<THREADS, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return CASPR_EINVAL;
  }
  CASPR_COUNT(); fps_kernel<THREADS, PPT><<<B, THREADS, smem, s>>>(xyz, N, M, idx, new_xyz);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

// Warp per centre, cloud in shared memory.  One scan in index order serves both radii of
// the set-abstraction level (hits of the small ball are a subset of the large one's):
// __ballot_sync + prefix popcount assigns output slots in index order, the scan stops as
// soon as both lists are full, and the tail is padded with the first hit.
constexpr int kBqThreads = 256;
constexpr int kBqCentresPerCta = 64;

__global__ void __launch_bounds__(kBqThreads)
ball_query2_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int M,
                   float r0sq, int ns0, int32_t* __restrict__ idx0,
                   float r1sq, int ns1, int32_t* __restrict__ idx1) {
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + N;
  float* sz = sy + N;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + (size_t)b * N * 3;
  for (int i = tid; i < N * 3; i += kBqThreads) {
    float v = p[i];
    int k = i / 3, c = i - 3 * k;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
  }
  __syncthreads();
  const unsigned lt_mask = (1u << lane) - 1u;
  const int c_end = min(M, (int)(blockIdx.x + 1) * kBqCentresPerCta);
  for (int c = blockIdx.x * kBqCentresPerCta + warp; c < c_end; c += kBqThreads / 32) {
    const float* q = new_xyz + ((size_t)b * M + c) * 3;
    const float cx = q[0], cy = q[1], cz = q[2];
    int32_t* o0 = idx0 ? idx0 + ((size_t)b * M + c) * ns0 : nullptr;
    int32_t* o1 = idx1 ? idx1 + ((size_t)b * M + c) * ns1 : nullptr;
    int cnt0 = o0 ? 0 : ns0, cnt1 = o1 ? 0 : ns1;
    int first0 = 0, first1 = 0;
    for (int base = 0; base < N; base += 32) {
      const int k = base + lane;
      float d2 = 3.0e38f;
      if (k < N) d2 = sqdist_canonical(cx, cy, cz, sx[k], sy[k], sz[k]);
      const bool h1 = d2 < r1sq, h0 = d2 < r0sq;
      const unsigned m1 = __ballot_sync(0xffffffffu, h1);
      const unsigned m0 = __ballot_sync(0xffffffffu, h0);
      if ((m0 | m1) == 0u) continue;
      if (o1 && cnt1 < ns1 && m1) {
        if (cnt1 == 0) first1 = base + __ffs(m1) - 1;
        int pos = cnt1 + __popc(m1 & lt_mask);
        if (h1 && pos < ns1) o1[pos] = k;
        cnt1 += __popc(m1);
      }
      if (o0 && cnt0 < ns0 && m0) {
        if (cnt0 == 0) first0 = base + __ffs(m0) - 1;
        int pos = cnt0 + __popc(m0 & lt_mask);
        if (h0 && pos < ns0) o0[pos] = k;
        cnt0 += __popc(m0);
      }
      if (cnt0 >= ns0 && cnt1 >= ns1) break;
    }
    if (o0) for (int s = min(cnt0, ns0) + lane; s < ns0; s += 32) o0[s] = first0;
    if (o1) for (int s = min(cnt1, ns1) + lane; s < ns1; s += 32) o1[s] = first1;
  }
}

// Warp per grouped row: out[(b,c,s)] = [xyz[idx]-centre | feat[idx][0:C]].
__global__ void __launch_bounds__(256)
group_points_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                    const float* __restrict__ feat, int ld_feat, const int32_t* __restrict__ idx,
                    int N, int M, int C, int ns, long long total_rows,
                    float* __restrict__ out, int ld_out) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < total_rows; row += stride) {
    const long long bc = row / ns;            // (b*M + c)
    const int b = (int)(bc / M);
    const int k = idx[row];
    float* o = out + row * (long long)ld_out;
    if (lane < 3) {
      float v = xyz[((size_t)b * N + k) * 3 + lane];
      float ctr = new_xyz[bc * 3 + lane];
      o[lane] = __fsub_rn(v, ctr);
    }
    if (C > 0) {
      const float* f = feat + ((size_t)b * N + k) * ld_feat;
      for (int c = lane; c < C; c += 32) o[3 + c] = f[c];
    }
  }
}

constexpr int kNnThreads = 256;

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, int n, int m,
                float* __restrict__ dist, int32_t* __restrict__ idx) {
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + m;
  float* sz = sy + m;
  const int b = blockIdx.y;
  const float* p = known + (size_t)b * m * 3;
  for (int i = threadIdx.x; i < m * 3; i += kNnThreads) {
    float v = p[i];
    int k = i / 3, c = i - 3 * k;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
  }
  __syncthreads();
  const int i = blockIdx.x * kNnThreads + threadIdx.x;
  if (i >= n) return;
  const float* u = unknown + ((size_t)b * n + i) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  float b1 = 3.0e38f, b2 = 3.0e38f, b3 = 3.0e38f;     // larger than any finite squared distance
  int i1 = 0, i2 = 0, i3 = 0;
  // oracle semantics: stable sort by (d2, index); with fewer than 3 known points the
  // remaining slots keep index order as well (m >= 3 in every reference configuration)
  for (int k = 0; k < m; ++k) {
    float d = sqdist_canonical(ux, uy, uz, sx[k], sy[k], sz[k]);
    if (d < b1) {
      b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
    } else if (d < b2) {
      b3 = b2; i3 = i2; b2 = d; i2 = k;
    } else if (d < b3) {
      b3 = d; i3 = k;
    }
  }
  float* od = dist + ((size_t)b * n + i) * 3;
  int32_t* oi = idx + ((size_t)b * n + i) * 3;
  od[0] = __fsqrt_rn(b1); od[1] = __fsqrt_rn(b2); od[2] = __fsqrt_rn(b3);
  oi[0] = i1; oi[1] = i2; oi[2] = i3;
}

// Warp per output row: inverse-distance weights (pointnet2.py:516-518), 3-point
// interpolation in slot order (oracle three_interpolate) and the skip-feature concat.
__global__ void __launch_bounds__(256)
three_interp_concat_kernel(const float* __restrict__ feat_prev, int ld_prev,
                           const int32_t* __restrict__ idx, const float* __restrict__ dist,
                           const float* __restrict__ skip, int ld_skip, int n, int m, int Cp, int Cs,
                           long long total_rows, float* __restrict__ out, int ld_out) {
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
    const float* f0 = feat_prev + ((size_t)b * m + idx[row * 3 + 0]) * ld_prev;
    const float* f1 = feat_prev + ((size_t)b * m + idx[row * 3 + 1]) * ld_prev;
    const float* f2 = feat_prev + ((size_t)b * m + idx[row * 3 + 2]) * ld_prev;
    float* o = out + row * (long long)ld_out;
    for (int c = lane; c < Cp; c += 32)
      o[c] = __fadd_rn(__fadd_rn(__fmul_rn(f0[c], w0), __fmul_rn(f1[c], w1)), __fmul_rn(f2[c], w2));
    if (Cs > 0) {
      const float* sk = skip + row * (long long)ld_skip;
      for (int c = lane; c < Cs; c += 32) o[Cp + c] = sk[c];
    }
  }
}

}  // namespace

extern "C" int caspr_fps(const float* xyz, int B, int N, int M, int32_t* idx, float* new_xyz,
                         void* stream) {
  CASPR_REQUIRE(xyz && idx && B > 0 && N > 0 && M > 0);
  cudaStream_t s = (cudaStream_t)stream;
  if (N <= 64) return launch_fps<64, 1>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 256) return launch_fps<128, 2>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 512) return launch_fps<256, 2>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 1024) return launch_fps<512, 2>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 2048) return launch_fps<512, 4>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 4096) return launch_fps<512, 8>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 8192) return launch_fps<1024, 8>(xyz, B, N, M, idx, new_xyz, s);
  if (N <= 16384) return launch_fps<1024, 16>(xyz, B, N, M, idx, new_xyz, s);
  return CASPR_EINVAL;
}

extern "C" int caspr_ball_query2(const float* xyz, const float* new_xyz, int B, int N, int M,
                                 float r0, int ns0, int32_t* idx0, float r1, int ns1, int32_t* idx1,
                                 void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && B > 0 && N > 0 && M > 0 && ns0 > 0 && ns1 > 0 && r0 <= r1);
  CASPR_REQUIRE(N <= 16384);
  size_t smem = (size_t)N * 3 * sizeof(float);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(ball_query2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return CASPR_EINVAL;
  }
  const float r0sq = r0 * r0, r1sq = r1 * r1;     // fp32, as upstream / the oracle
  dim3 grid(ceil_div(M, kBqCentresPerCta), B);
  CASPR_COUNT(); ball_query2_kernel<<<grid, kBqThreads, smem, (cudaStream_t)stream>>>(
      xyz, new_xyz, N, M, r0sq, ns0, idx0, r1sq, ns1, idx1);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_group_points(const float* xyz, const float* new_xyz, const float* feat,
                                  int ld_feat, const int32_t* idx, int B, int N, int M, int C, int ns,
                                  float* out, int ld_out, void* stream) {
  CASPR_REQUIRE(xyz && new_xyz && idx && out && B > 0 && N > 0 && M > 0 && ns > 0 && C >= 0);
  CASPR_REQUIRE((C == 0 || (feat && ld_feat >= C)) && ld_out >= 3 + C);
  long long rows = (long long)B * M * ns;
  int blocks = (int)((rows + 7) / 8 < 148LL * 16 ? (rows + 7) / 8 : 148LL * 16);
  CASPR_COUNT(); group_points_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(xyz, new_xyz, feat, ld_feat, idx, N, M, C, ns,
                                                                  rows, out, ld_out);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_three_nn(const float* unknown, const float* known, int B, int n, int m,
                              float* dist, int32_t* idx, void* stream) {
  CASPR_REQUIRE(unknown && known && dist && idx && B > 0 && n > 0 && m >= 3 && m <= 16384);
  size_t smem = (size_t)m * 3 * sizeof(float);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return CASPR_EINVAL;
  }
  dim3 grid(ceil_div(n, kNnThreads), B);
  CASPR_COUNT(); three_nn_kernel<<<grid, kNnThreads, smem, (cudaStream_t)stream>>>(unknown, known, n, m, dist, idx);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

extern "C" int caspr_three_interp_concat(const float* feat_prev, int ld_prev, const int32_t* idx,
                                         const float* dist, const float* skip, int ld_skip, int B,
                                         int n, int m, int Cp, int Cs, float* out, int ld_out,
                                         void* stream) {
  CASPR_REQUIRE(feat_prev && idx && dist && out && B > 0 && n > 0 && m > 0 && Cp > 0 && Cs >= 0);
  CASPR_REQUIRE((Cs == 0 || skip) && ld_out >= Cp + Cs && ld_prev >= Cp && (Cs == 0 || ld_skip >= Cs));
  long long rows = (long long)B * n;
  int blocks = (int)((rows + 7) / 8 < 148LL * 16 ? (rows + 7) / 8 : 148LL * 16);
  CASPR_COUNT(); three_interp_concat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      feat_prev, ld_prev, idx, dist, skip, ld_skip, n, m, Cp, Cs, rows, out, ld_out);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
