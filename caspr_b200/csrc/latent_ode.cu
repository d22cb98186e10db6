// Latent ODE: the whole adaptive dopri5 solve of the 64-d dynamic latent for ALL sequences of
// the batch in ONE persistent CTA (no host round-trips; torchdiffeq's controller is
// batch-global, so the batch cannot be split across independent solvers).
//
// Replaces LatentODE.forward / ODESolver / DynamicsNet of caspr/models/latent_ode_model.py:45-70,
// 76-99,129-147 and torchdiffeq 0.0.1's odeint_adjoint(dopri5) (oracle/odeint001.py).
// The dynamics are exact fp32 (the solver runs at rtol=atol=1e-3 and its accept/reject
// decisions must track the oracle's), weights stream from L2 (2.4 MB, resident).
#include "common.cuh"
#include "dopri5.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kBChunk = 8;

struct LatentParams {
  const float* W[4];
  const float* b[4];
  int B, D, H;
};

// out[b][j] = act(sum_k W[j][k] * in[b][k] + bias[j]) for all b: warp per output row j, lanes
// stride over k (coalesced weight reads), batch handled in register chunks of 8.
__device__ void dense_layer(const float* W, const float* bias, const float* in, float* out, int B,
                            int K, int N, bool do_tanh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < N; j += kWarps) {
    const float* wrow = W + (size_t)j * K;
    for (int b0 = 0; b0 < B; b0 += kBChunk) {
      float acc[kBChunk];
#pragma unroll
      for (int i = 0; i < kBChunk; ++i) acc[i] = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float w = wrow[k];
#pragma unroll
        for (int i = 0; i < kBChunk; ++i)
          if (b0 + i < B) acc[i] = fmaf(w, in[(size_t)(b0 + i) * K + k], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < kBChunk; ++i) {
        float v = warp_sum(acc[i]);
        if (lane == 0 && b0 + i < B) {
          v += bias[j];
          out[(size_t)(b0 + i) * N + j] = do_tanh ? tanhf(v) : v;
        }
      }
    }
  }
}

// f(z): Linear(D,H) tanh Linear(H,H) tanh Linear(H,H) tanh Linear(H,D)   (latent_ode_model.py:129-147)
__device__ void dynamics(const LatentParams& p, const float* z, float* dz, float* h0, float* h1) {
  dense_layer(p.W[0], p.b[0], z, h0, p.B, p.D, p.H, true);
  __syncthreads();
  dense_layer(p.W[1], p.b[1], h0, h1, p.B, p.H, p.H, true);
  __syncthreads();
  dense_layer(p.W[2], p.b[2], h1, h0, p.B, p.H, p.H, true);
  __syncthreads();
  dense_layer(p.W[3], p.b[3], h0, dz, p.B, p.H, p.D, false);
  __syncthreads();
}

__device__ double block_sum(double v, double* red) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = lane < kWarps ? red[lane] : 0.0;
  t = warp_sum_d(t);
  return t;     // every thread holds the block total
}

__global__ void __launch_bounds__(kThreads, 1)
latent_ode_kernel(LatentParams p, const float* z0, const double* times, int nT, float rtol,
                  float atol, float* out, int32_t* info, float* ws, int max_steps) {
  __shared__ double red[kWarps];
  __shared__ double s_t0, s_t1, s_dt;
  __shared__ int s_flag;
  const int n = p.B * p.D;
  const int tid = threadIdx.x;
  float* y0 = ws;                    // state at the start of the current step
  float* k = y0 + n;                 // 7 stage derivatives
  float* ys = k + 7 * (size_t)n;     // stage input / y1
  float* ymid = ys + n;
  float* yprev = ymid + n;           // y0 of the last ACCEPTED step (dense output)
  float* f0prev = yprev + n;
  float* y1acc = f0prev + n;
  float* f1acc = y1acc + n;
  float* h0 = f1acc + n;
  float* h1 = h0 + (size_t)p.B * p.H;

  int nfe = 0, accepted = 0, rejected = 0, status = CASPR_OK;
  for (int i = tid; i < n; i += kThreads) {
    float v = z0[i];
    y0[i] = v;
    out[i] = v;                                       // solution[0] = y0
  }
  __syncthreads();
  dynamics(p, y0, k, h0, h1);                         // f0
  nfe++;

  // ---- _select_initial_step(order 4), fp32 like torch
  double q0 = 0.0, q1 = 0.0;
  for (int i = tid; i < n; i += kThreads) {
    float sc = __fadd_rn(atol, __fmul_rn(fabsf(y0[i]), rtol));
    float a = __fdiv_rn(y0[i], sc), b = __fdiv_rn(k[i], sc);
    q0 += (double)a * a;
    q1 += (double)b * b;
  }
  q0 = block_sum(q0, red);
  q1 = block_sum(q1, red);
  const float rn = sqrtf((float)n);
  const float d0 = (float)sqrt(q0) / rn, d1 = (float)sqrt(q1) / rn;
  float hh0;
  if ((double)d0 < 1e-5 || (double)d1 < 1e-5) hh0 = 1e-6f;
  else hh0 = __fmul_rn(0.01f, __fdiv_rn(d0, d1));
  for (int i = tid; i < n; i += kThreads) ys[i] = __fadd_rn(y0[i], __fmul_rn(hh0, k[i]));
  __syncthreads();
  dynamics(p, ys, k + n, h0, h1);                     // f1 probe (stored in k[1], overwritten later)
  nfe++;
  double q2 = 0.0;
  for (int i = tid; i < n; i += kThreads) {
    float sc = __fadd_rn(atol, __fmul_rn(fabsf(y0[i]), rtol));
    float c = __fdiv_rn(__fsub_rn(k[n + i], k[i]), sc);
    q2 += (double)c * c;
  }
  q2 = block_sum(q2, red);
  const float d2 = __fdiv_rn((float)sqrt(q2) / rn, hh0);
  float hh1;
  if ((double)d1 <= 1e-15 && (double)d2 <= 1e-15) hh1 = fmaxf(1e-6f, __fmul_rn(hh0, 1e-3f));
  else hh1 = powf(__fdiv_rn(0.01f, fmaxf(d1, d2)), 1.0f / 5.0f);
  if (tid == 0) {
    s_t0 = times[0];
    s_t1 = times[0];
    s_dt = (double)fminf(__fmul_rn(100.f, hh0), hh1);
  }
  __syncthreads();

  for (int io = 1; io < nT && status == CASPR_OK; ++io) {
    const double t_out = times[io];
    int steps_here = 0;
    while (t_out > s_t1) {
      if (steps_here++ >= max_steps) { status = CASPR_ESOLVER_MAXSTEPS; break; }
      const double t0 = s_t1, dt = s_dt;
      if (!(t0 + dt > t0)) { status = CASPR_ESOLVER_DT; break; }
      // non-finite check on y0
      int bad = 0;
      for (int i = tid; i < n; i += kThreads) bad |= !isfinite(y0[i]);
      bad = __syncthreads_or(bad);
      if (bad) { status = CASPR_ESOLVER_NONFINITE; break; }
      const float dtf = (float)dt;
      for (int s = 0; s < 6; ++s) {
        for (int i = tid; i < n; i += kThreads) {
          float kv[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) kv[j] = j <= s ? k[(size_t)j * n + i] : 0.f;
          ys[i] = dopri5::stage_combine(y0[i], dtf, kv, s);
        }
        __syncthreads();
        dynamics(p, ys, k + (size_t)(s + 1) * n, h0, h1);
        nfe++;
      }
      // error ratio (ys holds y1)
      double q = 0.0;
      for (int i = tid; i < n; i += kThreads) {
        float kv[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) kv[j] = k[(size_t)j * n + i];
        float err = dopri5::weighted7(dtf, dopri5::kCErr, kv);
        float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y0[i]), fabsf(ys[i]))));
        float r = __fdiv_rn(err, tol);
        q += (double)__fmul_rn(r, r);
      }
      q = block_sum(q, red);
      const float ratio = (float)(q / (double)n);
      const bool accept = ratio <= 1.f;
      if (accept) {
        for (int i = tid; i < n; i += kThreads) {
          float kv[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) kv[j] = k[(size_t)j * n + i];
          ymid[i] = __fadd_rn(y0[i], dopri5::weighted7(dtf, dopri5::kCMid, kv));
          yprev[i] = y0[i];
          f0prev[i] = kv[0];
          y1acc[i] = ys[i];
          f1acc[i] = kv[6];
          y0[i] = ys[i];
          k[i] = kv[6];                                // FSAL
        }
        accepted++;
      } else {
        rejected++;
      }
      __syncthreads();
      if (tid == 0) {
        s_t0 = t0;
        s_t1 = accept ? t0 + dt : t0;
        s_dt = dopri5::optimal_step(dt, ratio);
        if (accept) s_flag = __float_as_int(dtf);      // dt of the step the dense output belongs to
      }
      __syncthreads();
    }
    if (status != CASPR_OK) break;
    // dense output at t_out (abscissa formed in fp32 like _interp_evaluate)
    const float t0f = (float)s_t0, t1f = (float)s_t1, tf = (float)t_out;
    const float x = __fdiv_rn(__fsub_rn(tf, t0f), __fsub_rn(t1f, t0f));
    const float dts = __int_as_float(s_flag);
    for (int i = tid; i < n; i += kThreads)
      out[(size_t)io * n + i] =
          dopri5::interp_eval(yprev[i], y1acc[i], ymid[i], f0prev[i], f1acc[i], dts, x);
    __syncthreads();
  }
  if (tid == 0) {
    info[INFO_STATUS] = status;
    info[INFO_NFE] = nfe;
    info[INFO_ACCEPTED] = accepted;
    info[INFO_REJECTED] = rejected;
    info[INFO_DONE] = 1;
  }
}

}  // namespace

extern "C" size_t caspr_latent_ode_workspace_bytes(int B, int D, int H) {
  size_t n = (size_t)B * D;
  return (n * 15 + (size_t)B * H * 2) * sizeof(float) + 8 * sizeof(double) * 64;
}

extern "C" int caspr_latent_ode_solve(const float* z0, int B, int D, int H, const float* W0,
                                      const float* b0, const float* W1, const float* b1,
                                      const float* W2, const float* b2, const float* W3,
                                      const float* b3, const double* h_times, int nT, float rtol,
                                      float atol, float* out, int32_t* info, int32_t* h_info,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(z0 && W0 && b0 && W1 && b1 && W2 && b2 && W3 && b3 && h_times && out && info && h_info);
  CASPR_REQUIRE(B > 0 && D > 0 && H > 0 && nT >= 1 && nT <= 64);
  for (int i = 1; i < nT; ++i) CASPR_REQUIRE(h_times[i] > h_times[i - 1]);
  if (workspace_bytes < caspr_latent_ode_workspace_bytes(B, D, H)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  // time grid lives at the tail of the workspace (64 doubles reserved)
  size_t n = (size_t)B * D;
  float* ws = (float*)workspace;
  double* d_times = (double*)((char*)workspace + align_up((n * 15 + (size_t)B * H * 2) * sizeof(float), 8));
  if (cudaMemcpyAsync(d_times, h_times, nT * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  LatentParams p;
  p.W[0] = W0; p.W[1] = W1; p.W[2] = W2; p.W[3] = W3;
  p.b[0] = b0; p.b[1] = b1; p.b[2] = b2; p.b[3] = b3;
  p.B = B; p.D = D; p.H = H;
  CASPR_COUNT(); latent_ode_kernel<<<1, kThreads, 0, s>>>(p, z0, d_times, nT, rtol, atol, out, info, ws, 100000);
  CASPR_CHECK_LAUNCH();
  if (cudaMemcpyAsync(h_info, info, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  return h_info[INFO_STATUS];
}
