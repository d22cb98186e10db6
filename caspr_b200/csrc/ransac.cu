// Correspondence-RANSAC rigid pose from T-NOCS predictions (second half).
//
// Replaces the per-frame CPU loop of utils/evaluations.py:360-380:
//   o3d.registration.registration_ransac_based_on_correspondence(source = predicted NOCS - 0.5, target = input points,
//       corres = identity, max_correspondence_distance = 0.015, TransformationEstimationPointToPoint(False),
//       ransac_n = 4, RANSACConvergenceCriteria(50000, 5000))
// open3d is NOT under /root/reference and is unpinned (the `o3d.registration` namespace dates it <= 0.10), so this is a
// restatement of its published algorithm (parity unpinned): for each of min(max_iteration, max_validation) = 5000
// hypotheses draw 4 correspondences, fit the least-squares rigid transform (Kabsch / Umeyama without scale), score it
// by the number of correspondences closer than the threshold (fitness) and their RMSE, keep the best (higher fitness,
// then lower RMSE, then the earlier hypothesis); no final refit in that version (optional here).  open3d draws the
// samples with C rand(); here the caller supplies them (counter-based device RNG), so the oracle scores the SAME
// hypotheses.
//
// One launch for all frames: grid (ceil(H/256), frames), a thread owns one hypothesis.  The frame's correspondences are
// staged once in shared memory (48 KB at N = 2048); every thread fits its transform in registers (Horn's quaternion
// form of the Kabsch problem, 4x4 symmetric Jacobi eigen-solve in double) and scans all N correspondences with
// broadcast shared-memory reads.  Declared arithmetic of the inlier test (the oracle follows it bit for bit): R, t
// rounded to fp32; p' = ((r0*x + r1*y) + r2*z) + t and d2 = ((dx*dx)+(dy*dy))+(dz*dz) with every product and sum
// rounded separately (no FMA contraction); inlier iff d2 < thr*thr (fp32).
#include "common.cuh"

namespace {

struct Hyp {
  float R[9], t[3];
  int ok;
};

// Largest-eigenvalue eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations.
__device__ void jacobi4_max(double A[4][4], double q[4]) {
  double V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int r = p + 1; r < 4; ++r) off += A[p][r] * A[p][r];
    if (off < 1e-30) break;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int r = p + 1; r < 4; ++r) {
        const double apq = A[p][r];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[r][r] - A[p][p]) / (2.0 * apq);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akr = A[k][r];
          A[k][p] = c * akp - s * akr;
          A[k][r] = s * akp + c * akr;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], ark = A[r][k];
          A[p][k] = c * apk - s * ark;
          A[r][k] = s * apk + c * ark;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkr = V[k][r];
          V[k][p] = c * vkp - s * vkr;
          V[k][r] = s * vkp + c * vkr;
        }
      }
    }
  }
  int best = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (A[i][i] > A[best][best]) best = i;
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

// Least-squares rigid transform dst ~ R src + t of n correspondences (Horn 1987; equals Kabsch / Umeyama without
// scaling whenever the optimum is unique).
__device__ void fit_rigid(const double (*s)[3], const double (*d)[3], int n, double R[9], double t[3]) {
  double cs[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) { cs[c] += s[i][c]; cd[c] += d[i][c]; }
#pragma unroll
  for (int c = 0; c < 3; ++c) { cs[c] /= n; cd[c] /= n; }
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};          // S[a][b] = sum (s_a - cs_a)(d_b - cd_b)
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) S[a][b] += (s[i][a] - cs[a]) * (d[i][b] - cd[b]);
  double N[4][4];
  N[0][0] = S[0][0] + S[1][1] + S[2][2];
  N[0][1] = S[1][2] - S[2][1];
  N[0][2] = S[2][0] - S[0][2];
  N[0][3] = S[0][1] - S[1][0];
  N[1][1] = S[0][0] - S[1][1] - S[2][2];
  N[1][2] = S[0][1] + S[1][0];
  N[1][3] = S[2][0] + S[0][2];
  N[2][2] = -S[0][0] + S[1][1] - S[2][2];
  N[2][3] = S[1][2] + S[2][1];
  N[3][3] = -S[0][0] - S[1][1] + S[2][2];
#pragma unroll
  for (int i = 1; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) N[i][j] = N[j][i];
  double q[4];
  jacobi4_max(N, q);
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / nq, x = q[1] / nq, y = q[2] / nq, z = q[3] / nq;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
#pragma unroll
  for (int a = 0; a < 3; ++a) t[a] = cd[a] - (R[3 * a] * cs[0] + R[3 * a + 1] * cs[1] + R[3 * a + 2] * cs[2]);
}

__device__ __forceinline__ float sqdist_after(const float* R, const float* t, float sx, float sy, float sz, float dx,
                                              float dy, float dz) {
  const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], sx), __fmul_rn(R[1], sy)), __fmul_rn(R[2], sz)), t[0]);
  const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], sx), __fmul_rn(R[4], sy)), __fmul_rn(R[5], sz)), t[1]);
  const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], sx), __fmul_rn(R[7], sy)), __fmul_rn(R[8], sz)), t[2]);
  const float ex = __fsub_rn(px, dx), ey = __fsub_rn(py, dy), ez = __fsub_rn(pz, dz);
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

// grid (ceil(H/256), frames).  dynamic shared memory: 6*N floats (src x,y,z | dst x,y,z as separate arrays).
__global__ void __launch_bounds__(256)
ransac_hypotheses_kernel(const float* __restrict__ src, const float* __restrict__ dst, const int32_t* __restrict__ samples,
                         int N, int H, float thr2, int32_t* __restrict__ counts, double* __restrict__ err2,
                         float* __restrict__ Rt) {
  extern __shared__ float sm[];
  float *sx = sm, *sy = sm + N, *sz = sm + 2 * N, *dx = sm + 3 * N, *dy = sm + 4 * N, *dz = sm + 5 * N;
  const int frame = blockIdx.y;
  const float* ps = src + (size_t)frame * N * 3;
  const float* pd = dst + (size_t)frame * N * 3;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) {
    const int k = i / 3, c = i - 3 * k;
    const float a = ps[i], b = pd[i];
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = a;
    (c == 0 ? dx : (c == 1 ? dy : dz))[k] = b;
  }
  __syncthreads();
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int32_t* smp = samples + ((size_t)frame * H + h) * 4;
  double s4[4][3], d4[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int k = smp[j];
    k = k < 0 ? 0 : (k >= N ? N - 1 : k);
    s4[j][0] = sx[k]; s4[j][1] = sy[k]; s4[j][2] = sz[k];
    d4[j][0] = dx[k]; d4[j][1] = dy[k]; d4[j][2] = dz[k];
  }
  double Rd[9], td[3];
  fit_rigid(s4, d4, 4, Rd, td);
  float R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = (float)td[i];
  int cnt = 0;
  double e2 = 0.0;
  for (int k = 0; k < N; ++k) {
    const float d2 = sqdist_after(R, t, sx[k], sy[k], sz[k], dx[k], dy[k], dz[k]);
    if (d2 < thr2) { ++cnt; e2 += (double)d2; }
  }
  const size_t o = (size_t)frame * H + h;
  counts[o] = cnt;
  err2[o] = e2;
  float* out = Rt + o * 12;
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) out[9 + i] = t[i];
}

// One CTA per frame: best hypothesis = highest inlier count, then lowest RMSE (= err2 / count), then lowest index;
// optional least-squares refit on the inliers of the winner (newer open3d versions do this).
__global__ void __launch_bounds__(256)
ransac_select_kernel(const float* __restrict__ src, const float* __restrict__ dst, int N, int H, float thr2,
                     const int32_t* __restrict__ counts, const double* __restrict__ err2, const float* __restrict__ Rt,
                     int refine, float* __restrict__ R_out, float* __restrict__ t_out, int32_t* __restrict__ best_out,
                     float* __restrict__ fitness_out, float* __restrict__ rmse_out) {
  __shared__ int s_cnt[256], s_idx[256];
  __shared__ double s_mse[256];
  __shared__ double s_acc[256][8];
  const int frame = blockIdx.x;
  const int32_t* c = counts + (size_t)frame * H;
  const double* e = err2 + (size_t)frame * H;
  int bc = -1, bi = 0x7fffffff;
  double bm = 0.0;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const int cc = c[h];
    const double mse = cc > 0 ? e[h] / (double)cc : 0.0;
    if (cc > bc || (cc == bc && (mse < bm || (mse == bm && h < bi)))) { bc = cc; bm = mse; bi = h; }
  }
  s_cnt[threadIdx.x] = bc; s_mse[threadIdx.x] = bm; s_idx[threadIdx.x] = bi;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      const int oc = s_cnt[threadIdx.x + w], oi = s_idx[threadIdx.x + w];
      const double om = s_mse[threadIdx.x + w];
      const int mc = s_cnt[threadIdx.x], mi = s_idx[threadIdx.x];
      const double mm = s_mse[threadIdx.x];
      if (oc > mc || (oc == mc && (om < mm || (om == mm && oi < mi)))) {
        s_cnt[threadIdx.x] = oc; s_mse[threadIdx.x] = om; s_idx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const int best = s_idx[0], cnt = s_cnt[0];
  const float* rt = Rt + ((size_t)frame * H + best) * 12;
  if (!refine || cnt < 3) {
    if (threadIdx.x < 9) R_out[frame * 9 + threadIdx.x] = rt[threadIdx.x];
    if (threadIdx.x < 3) t_out[frame * 3 + threadIdx.x] = rt[9 + threadIdx.x];
    if (threadIdx.x == 0) {
      best_out[frame] = best;
      fitness_out[frame] = (float)cnt / (float)N;
      rmse_out[frame] = cnt > 0 ? (float)sqrt(s_mse[0]) : 0.f;
    }
    return;
  }
  // refit on the winner's inliers: centroids and the 3x3 cross-covariance by two block reductions in double
  float R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = rt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = rt[9 + i];
  const float* ps = src + (size_t)frame * N * 3;
  const float* pd = dst + (size_t)frame * N * 3;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const float d2 = sqdist_after(R, t, ps[3 * k], ps[3 * k + 1], ps[3 * k + 2], pd[3 * k], pd[3 * k + 1], pd[3 * k + 2]);
    if (d2 < thr2) {
      a[0] += ps[3 * k]; a[1] += ps[3 * k + 1]; a[2] += ps[3 * k + 2];
      a[3] += pd[3 * k]; a[4] += pd[3 * k + 1]; a[5] += pd[3 * k + 2];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_acc[threadIdx.x][i] = a[i];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
#pragma unroll
      for (int i = 0; i < 6; ++i) s_acc[threadIdx.x][i] += s_acc[threadIdx.x + w][i];
    __syncthreads();
  }
  double cen[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) cen[i] = s_acc[0][i] / (double)cnt;
  __syncthreads();
  __shared__ double s_cov[256][9];
  double cv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const float d2 = sqdist_after(R, t, ps[3 * k], ps[3 * k + 1], ps[3 * k + 2], pd[3 * k], pd[3 * k + 1], pd[3 * k + 2]);
    if (d2 < thr2) {
      const double s0 = ps[3 * k] - cen[0], s1 = ps[3 * k + 1] - cen[1], s2 = ps[3 * k + 2] - cen[2];
      const double d0 = pd[3 * k] - cen[3], d1 = pd[3 * k + 1] - cen[4], dd2 = pd[3 * k + 2] - cen[5];
      cv[0] += s0 * d0; cv[1] += s0 * d1; cv[2] += s0 * dd2;
      cv[3] += s1 * d0; cv[4] += s1 * d1; cv[5] += s1 * dd2;
      cv[6] += s2 * d0; cv[7] += s2 * d1; cv[8] += s2 * dd2;
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) s_cov[threadIdx.x][i] = cv[i];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
#pragma unroll
      for (int i = 0; i < 9; ++i) s_cov[threadIdx.x][i] += s_cov[threadIdx.x + w][i];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // reuse fit_rigid's quaternion solve on the accumulated covariance: two synthetic "points" are not needed, the
    // covariance enters directly
    double S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) S[i][j] = s_cov[0][3 * i + j];
    double Nm[4][4];
    Nm[0][0] = S[0][0] + S[1][1] + S[2][2];
    Nm[0][1] = S[1][2] - S[2][1];
    Nm[0][2] = S[2][0] - S[0][2];
    Nm[0][3] = S[0][1] - S[1][0];
    Nm[1][1] = S[0][0] - S[1][1] - S[2][2];
    Nm[1][2] = S[0][1] + S[1][0];
    Nm[1][3] = S[2][0] + S[0][2];
    Nm[2][2] = -S[0][0] + S[1][1] - S[2][2];
    Nm[2][3] = S[1][2] + S[2][1];
    Nm[3][3] = -S[0][0] - S[1][1] + S[2][2];
    for (int i = 1; i < 4; ++i)
      for (int j = 0; j < i; ++j) Nm[i][j] = Nm[j][i];
    double q[4];
    jacobi4_max(Nm, q);
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] / nq, x = q[1] / nq, y = q[2] / nq, z = q[3] / nq;
    double Rr[9];
    Rr[0] = 1 - 2 * (y * y + z * z); Rr[1] = 2 * (x * y - w * z);     Rr[2] = 2 * (x * z + w * y);
    Rr[3] = 2 * (x * y + w * z);     Rr[4] = 1 - 2 * (x * x + z * z); Rr[5] = 2 * (y * z - w * x);
    Rr[6] = 2 * (x * z - w * y);     Rr[7] = 2 * (y * z + w * x);     Rr[8] = 1 - 2 * (x * x + y * y);
    for (int i = 0; i < 9; ++i) R_out[frame * 9 + i] = (float)Rr[i];
    for (int i = 0; i < 3; ++i)
      t_out[frame * 3 + i] = (float)(cen[3 + i] - (Rr[3 * i] * cen[0] + Rr[3 * i + 1] * cen[1] + Rr[3 * i + 2] * cen[2]));
    best_out[frame] = best;
    fitness_out[frame] = (float)cnt / (float)N;
    rmse_out[frame] = (float)sqrt(s_mse[0]);
  }
}

struct RansacLayout {
  size_t off_counts, off_err2, off_rt, total;
};
RansacLayout ransac_layout(int frames, int H) {
  RansacLayout l;
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 256); return r; };
  l.off_counts = take((size_t)frames * H * 4);
  l.off_err2 = take((size_t)frames * H * 8);
  l.off_rt = take((size_t)frames * H * 12 * 4);
  l.total = p;
  return l;
}

}  // namespace

extern "C" size_t caspr_ransac_pose_workspace_bytes(int frames, int hypotheses) {
  if (frames <= 0 || hypotheses <= 0) return 0;
  return ransac_layout(frames, hypotheses).total;
}

extern "C" int caspr_ransac_pose(const float* src, const float* dst, const int32_t* samples, int frames, int N,
                                 int hypotheses, float max_distance, int refine, float* R_out, float* t_out,
                                 int32_t* best_out, float* fitness_out, float* rmse_out, int32_t* counts_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(src && dst && samples && R_out && t_out && best_out && fitness_out && rmse_out && workspace);
  CASPR_REQUIRE(frames > 0 && N >= 4 && hypotheses > 0 && max_distance > 0.f);
  CASPR_REQUIRE((size_t)N * 6 * sizeof(float) <= 200 * 1024);           // the frame's correspondences live in smem
  CASPR_REQUIRE(((uintptr_t)workspace & 255) == 0);
  const RansacLayout l = ransac_layout(frames, hypotheses);
  if (workspace_bytes < l.total) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  int32_t* counts = (int32_t*)(base + l.off_counts);
  double* err2 = (double*)(base + l.off_err2);
  float* Rt = (float*)(base + l.off_rt);
  const size_t smem = (size_t)N * 6 * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(ransac_hypotheses_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return CASPR_ELAUNCH;
    smem_set = smem;
  }
  const float thr2 = max_distance * max_distance;
  CASPR_COUNT(); ransac_hypotheses_kernel<<<dim3(ceil_div(hypotheses, 256), frames), 256, smem, s>>>(
      src, dst, samples, N, hypotheses, thr2, counts, err2, Rt);
  CASPR_COUNT(); ransac_select_kernel<<<frames, 256, 0, s>>>(src, dst, N, hypotheses, thr2, counts, err2, Rt, refine, R_out,
                                                            t_out, best_out, fitness_out, rmse_out);
  CASPR_CHECK_LAUNCH();
  if (counts_out &&
      cudaMemcpyAsync(counts_out, counts, (size_t)frames * hypotheses * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  return CASPR_OK;
}
