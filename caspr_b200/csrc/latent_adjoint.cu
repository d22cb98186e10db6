// Adjoint backward of the latent ODE (training, BASELINE config 5).
//
// Replaces torchdiffeq 0.0.1's OdeintAdjointMethod.backward for the call at
// caspr/models/latent_ode_model.py:98 (restated in oracle/odeint001.py::_AdjointMethod) and the autograd VJP
// through DynamicsNet (latent_ode_model.py:129-147: Linear tanh Linear tanh Linear tanh Linear).
//
// For every output interval [t_{i-1}, t_i], last to first, the augmented state (z, adj_z, adj_params) is
// integrated backwards with dopri5 from z = the saved forward solution at t_i.  The tolerances are scalars here
// (rtol = atol = 1e-3), so ALL augmented tensors take part in step control.  adj_t is left out: DynamicsNet
// ignores t, its derivative is identically zero, its error ratio is zero, and the reference discards the time
// gradients of this solve.
// The state is tiny (B x 64) and the work is latency-bound: one evaluation is 13 small launches (4 forward
// layers, 4 data-gradient layers on transposed weights, 4 outer-product weight gradients, 1 stage combine).
#include <string.h>
#include "rk_flat.cuh"

namespace {

enum { kEpiNone = 0, kEpiTanh = 1, kEpiTanhGrad = 2 };

// y[b][j] = epi(scale * (x[b][:] . W[j][:] + bias[j])); one warp per output channel j, rows in tiles of 8.
// kEpiTanhGrad multiplies by (1 - aux[b][j]^2) (cotangent through h = tanh(.)).
__global__ void __launch_bounds__(256)
lat_linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                  int B, int K, int N, int epi, float scale, const float* __restrict__ aux, int ld_aux,
                  const CnfState* __restrict__ st, float* __restrict__ y, int ldy) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= N) return;
  const float* w = W + (size_t)j * K;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int k = lane; k < K; k += 32) {
      const float wv = w[k];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (b0 + r < B) acc[r] = fmaf(wv, x[(size_t)(b0 + r) * ldx + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bj = bias ? bias[j] : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (b0 + r >= B) break;
        float v = scale * (acc[r] + bj);
        if (epi == kEpiTanh) v = tanhf(v);
        if (epi == kEpiTanhGrad) {
          const float h = aux[(size_t)(b0 + r) * ld_aux + j];
          v = v * (1.f - h * h);
        }
        y[(size_t)(b0 + r) * ldy + j] = v;
      }
    }
  }
}

// dW[j][k] = sum_b d[b][j] x[b][k] ; db[j] = sum_b d[b][j].  Block j, thread k.
__global__ void __launch_bounds__(512)
lat_wgrad_kernel(const float* __restrict__ d, int ldd, const float* __restrict__ x, int ldx, int B, int K,
                 const CnfState* __restrict__ st, float* __restrict__ dW, float* __restrict__ db) {
  if (st->done) return;
  const int j = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(d[(size_t)b * ldd + j], x[(size_t)b * ldx + k], acc);
    dW[(size_t)j * K + k] = acc;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += d[(size_t)b * ldd + j];
    db[j] = s;
  }
}

__global__ void lat_transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = blockIdx.y * 32 + r;
    if (x < cols && y < rows) tile[r][threadIdx.x] = src[(size_t)y * cols + x];
  }
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int yo = blockIdx.x * 32 + r;
    if (xo < rows && yo < cols) dst[(size_t)yo * rows + xo] = tile[threadIdx.x][r];
  }
}

// y0[z] = zs_i ; y0[adj] = (first ? 0 : adj_out) + g_i
__global__ void lat_segment_init_kernel(const float* __restrict__ zs_i, const float* __restrict__ g_i,
                                        const float* __restrict__ adj_prev, int nz, float* __restrict__ y0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nz) return;
  y0[i] = zs_i[i];
  y0[nz + i] = (adj_prev ? adj_prev[i] : 0.f) + g_i[i];
}
__global__ void lat_add_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

__global__ void lat_clear_done_kernel(CnfState* st) { st->done = 0; }

struct LatLayout {
  size_t W[4], b[4], total;
};
LatLayout lat_layout(int D, int H) {
  LatLayout L;
  const size_t din[4] = {(size_t)D, (size_t)H, (size_t)H, (size_t)H};
  const size_t dout[4] = {(size_t)H, (size_t)H, (size_t)H, (size_t)D};
  size_t off = 0;
  for (int l = 0; l < 4; ++l) {
    L.W[l] = off; off += dout[l] * din[l];
    L.b[l] = off; off += dout[l];
  }
  L.total = off;
  return L;
}

struct LatWorkspace {
  CnfState* st;
  double* sums;
  float *Y0, *kY, *Yout, *ys;      // flat state [z | adj_z | params]
  float *h1, *h2, *h3, *d2, *d1, *d0;
  float* Wt[4];
  size_t ny, par_off, bytes;
};
LatWorkspace lat_carve(void* base, int B, int D, int H) {
  LatWorkspace w;
  memset(&w, 0, sizeof(w));
  const LatLayout pl = lat_layout(D, H);
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
  w.st = (CnfState*)take(sizeof(CnfState));
  w.sums = (double*)take(8 * 8);
  w.par_off = (size_t)2 * B * D;
  w.ny = (w.par_off + pl.total + 3) / 4 * 4;
  w.Y0 = (float*)take(w.ny * 4);
  w.kY = (float*)take(7 * w.ny * 4);
  w.Yout = (float*)take(w.ny * 4);
  w.ys = (float*)take((size_t)2 * B * D * 4);
  float** acts[] = {&w.h1, &w.h2, &w.h3, &w.d2, &w.d1, &w.d0};
  for (float** a : acts) *a = (float*)take((size_t)B * H * 4);
  const size_t wsz[4] = {(size_t)H * D, (size_t)H * H, (size_t)H * H, (size_t)D * H};
  for (int l = 0; l < 4; ++l) w.Wt[l] = (float*)take(wsz[l] * 4);
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

struct LatNet {
  const float* W[4];
  const float* b[4];
};

int lat_blocks(long long work, int per_block, int cap) {
  long long b = (work + per_block - 1) / per_block;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// one augmented evaluation at RK stage `stage`: k[z] = -f(z_s), k[adj] = d(adj_s . f)/dz, k[params] = d(adj_s . f)/dtheta
int lat_aug_eval(const LatWorkspace& w, const LatNet& net, int B, int D, int H, int stage, cudaStream_t s) {
  const LatLayout pl = lat_layout(D, H);
  const size_t nz = (size_t)B * D;
  float* k = w.kY + (size_t)stage * w.ny;
  float* kpar = k + w.par_off;
  CASPR_COUNT(); flat_stage_kernel<<<lat_blocks(2 * nz, 256, 64), 256, 0, s>>>(w.Y0, w.kY, w.ny, 2 * nz, stage, w.st, w.ys);
  const float* zs = w.ys;
  const float* adj = w.ys + nz;
  auto lin = [&](const float* x, int ldx, const float* W, const float* bias, int K, int N, int epi, float scale,
                 const float* aux, int ld_aux, float* y, int ldy) {
    CASPR_COUNT(); lat_linear_kernel<<<ceil_div(N, 8), 256, 0, s>>>(x, ldx, W, bias, B, K, N, epi, scale, aux, ld_aux,
                                                                   w.st, y, ldy);
  };
  // forward
  lin(zs, D, net.W[0], net.b[0], D, H, kEpiTanh, 1.f, nullptr, 0, w.h1, H);
  lin(w.h1, H, net.W[1], net.b[1], H, H, kEpiTanh, 1.f, nullptr, 0, w.h2, H);
  lin(w.h2, H, net.W[2], net.b[2], H, H, kEpiTanh, 1.f, nullptr, 0, w.h3, H);
  lin(w.h3, H, net.W[3], net.b[3], H, D, kEpiNone, -1.f, nullptr, 0, k, D);          // reversed time: -f
  // backward: cotangent of the output = adj_s
  CASPR_COUNT(); lat_wgrad_kernel<<<D, 512, 0, s>>>(adj, D, w.h3, H, B, H, w.st, kpar + pl.W[3], kpar + pl.b[3]);
  lin(adj, D, w.Wt[3], nullptr, D, H, kEpiTanhGrad, 1.f, w.h3, H, w.d2, H);
  CASPR_COUNT(); lat_wgrad_kernel<<<H, 512, 0, s>>>(w.d2, H, w.h2, H, B, H, w.st, kpar + pl.W[2], kpar + pl.b[2]);
  lin(w.d2, H, w.Wt[2], nullptr, H, H, kEpiTanhGrad, 1.f, w.h2, H, w.d1, H);
  CASPR_COUNT(); lat_wgrad_kernel<<<H, 512, 0, s>>>(w.d1, H, w.h1, H, B, H, w.st, kpar + pl.W[1], kpar + pl.b[1]);
  lin(w.d1, H, w.Wt[1], nullptr, H, H, kEpiTanhGrad, 1.f, w.h1, H, w.d0, H);
  CASPR_COUNT(); lat_wgrad_kernel<<<H, 64, 0, s>>>(w.d0, H, zs, D, B, D, w.st, kpar + pl.W[0], kpar + pl.b[0]);
  lin(w.d0, H, w.Wt[0], nullptr, H, D, kEpiNone, 1.f, nullptr, 0, k + nz, D);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace

extern "C" size_t caspr_latent_ode_param_count(int D, int H) {
  if (D <= 0 || H <= 0) return 0;
  return lat_layout(D, H).total;
}

extern "C" size_t caspr_latent_ode_adjoint_workspace_bytes(int B, int D, int H) {
  if (B <= 0 || D <= 0 || H <= 0) return 0;
  return lat_carve(nullptr, B, D, H).bytes;
}

extern "C" int caspr_latent_ode_adjoint(const float* zs, const float* gzs, int B, int D, int H, const float* W0,
                                        const float* b0, const float* W1, const float* b1, const float* W2,
                                        const float* b2, const float* W3, const float* b3, const double* h_times,
                                        int nT, float rtol, float atol, float* gz0, float* gparams, int32_t* info,
                                        int32_t* h_info, void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(zs && gzs && W0 && b0 && W1 && b1 && W2 && b2 && W3 && b3 && h_times && gz0 && gparams);
  CASPR_REQUIRE(info && h_info && workspace && B > 0 && D > 0 && H > 0 && nT >= 1 && nT <= 64);
  CASPR_REQUIRE(((uintptr_t)workspace & 255) == 0);
  for (int i = 1; i < nT; ++i) CASPR_REQUIRE(h_times[i] > h_times[i - 1]);
  if (workspace_bytes < caspr_latent_ode_adjoint_workspace_bytes(B, D, H)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  LatWorkspace w = lat_carve(workspace, B, D, H);
  const LatLayout pl = lat_layout(D, H);
  const LatNet net = {{W0, W1, W2, W3}, {b0, b1, b2, b3}};
  const int nz = B * D;
  FlatSegs segs;
  memset(&segs, 0, sizeof(segs));
  segs.nseg = 3;
  segs.begin[0] = 0; segs.end[0] = nz;
  segs.begin[1] = nz; segs.end[1] = 2 * (size_t)nz;
  segs.begin[2] = w.par_off; segs.end[2] = w.par_off + pl.total;

  {
    const dim3 tb(32, 8);
    const int rows[4] = {H, H, H, D}, cols[4] = {D, H, H, H};
    for (int l = 0; l < 4; ++l) {
      const dim3 tg(ceil_div(cols[l], 32), ceil_div(rows[l], 32));
      CASPR_COUNT(); lat_transpose_kernel<<<tg, tb, 0, s>>>(net.W[l], rows[l], cols[l], w.Wt[l]);
    }
    CASPR_CHECK_LAUNCH();
  }
  if (cudaMemsetAsync(w.Y0, 0, w.ny * sizeof(float), s) != cudaSuccess ||
      cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), s) != cudaSuccess)
    return CASPR_ELAUNCH;
  CnfState hst;
  memset(&hst, 0, sizeof(hst));
  if (cudaMemcpyAsync(w.st, &hst, sizeof(CnfState), cudaMemcpyHostToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;

  const int fb = lat_blocks((long long)w.ny, 256, 148 * 4);
  const int kMaxSteps = 100000;
  int step_id = 0, status = CASPR_OK;
  bool have_adj = false;
  for (int i = nT - 1; i >= 1 && status == CASPR_OK; --i) {
    CASPR_COUNT(); lat_segment_init_kernel<<<ceil_div(nz, 256), 256, 0, s>>>(
        zs + (size_t)i * nz, gzs + (size_t)i * nz, have_adj ? w.Yout + nz : nullptr, nz, w.Y0);
    if (have_adj) {       // carry the parameter adjoint over from the previous interval
      if (cudaMemcpyAsync(w.Y0 + w.par_off, w.Yout + w.par_off, pl.total * sizeof(float), cudaMemcpyDeviceToDevice,
                          s) != cudaSuccess)
        return CASPR_ELAUNCH;
    }
    // stage-0 evaluation with done = 0, then the first step size
    CASPR_COUNT(); lat_clear_done_kernel<<<1, 1, 0, s>>>(w.st);
    int rc = lat_aug_eval(w, net, B, D, H, 0, s);
    if (rc) return rc;
    CASPR_COUNT(); flat_init_norm_kernel<<<fb, 256, 0, s>>>(w.Y0, w.kY, segs, rtol, atol, w.sums);
    CASPR_COUNT(); flat_init_controller_kernel<<<1, 1, 0, s>>>(w.st, segs, w.sums, -h_times[i], -h_times[i - 1]);
    CASPR_CHECK_LAUNCH();
    for (;;) {
      for (int stage = 1; stage <= 6; ++stage) {
        rc = lat_aug_eval(w, net, B, D, H, stage, s);
        if (rc) return rc;
      }
      CASPR_COUNT(); flat_error_kernel<<<fb, 256, 0, s>>>(w.Y0, w.kY, w.ny, segs, rtol, atol, w.st, w.sums);
      CASPR_COUNT(); flat_controller_kernel<<<1, 1, 0, s>>>(w.st, segs, w.sums, step_id);
      CASPR_COUNT(); flat_finalize_kernel<<<fb, 256, 0, s>>>(w.Y0, w.kY, w.ny, w.ny, step_id, w.st, w.Yout);
      CASPR_CHECK_LAUNCH();
      ++step_id;
      if (cudaMemcpyAsync(&hst, w.st, sizeof(CnfState), cudaMemcpyDeviceToHost, s) != cudaSuccess) return CASPR_ELAUNCH;
      if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
      if (hst.done) break;
      if (step_id >= kMaxSteps) { hst.status = CASPR_ESOLVER_MAXSTEPS; break; }
    }
    status = hst.status;
    have_adj = true;
  }
  if (status == CASPR_OK) {
    if (have_adj) {
      CASPR_COUNT(); lat_add_kernel<<<ceil_div(nz, 256), 256, 0, s>>>(w.Yout + nz, gzs, nz, gz0);
      CASPR_CHECK_LAUNCH();
      if (cudaMemcpyAsync(gparams, w.Yout + w.par_off, pl.total * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return CASPR_ELAUNCH;
    } else {             // a single time point: the output is z0 itself
      if (cudaMemcpyAsync(gz0, gzs, (size_t)nz * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
          cudaMemsetAsync(gparams, 0, pl.total * sizeof(float), s) != cudaSuccess)
        return CASPR_ELAUNCH;
    }
  }
  int32_t out_info[8] = {status, hst.nfe, hst.accepted, hst.rejected, hst.done, 0, 0, step_id};
  for (int i = 0; i < 8; ++i) h_info[i] = out_info[i];
  if (cudaMemcpyAsync(info, h_info, 8 * sizeof(int32_t), cudaMemcpyHostToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  return status;
}
