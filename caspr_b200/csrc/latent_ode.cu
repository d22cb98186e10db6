// Latent ODE: the whole adaptive dopri5 solve of the 64-d dynamic latent for ALL sequences of the
// batch in ONE cooperative kernel launch (no host round-trips; torchdiffeq's controller is
// batch-global, so the batch cannot be split across independent solvers).
//
// Replaces LatentODE.forward / ODESolver / DynamicsNet of caspr/models/latent_ode_model.py:45-70,
// 76-99,129-147 and torchdiffeq 0.0.1's odeint_adjoint(dopri5) (oracle/odeint001.py).
//
// Structure: G = 64 co-resident CTAs.  Each CTA keeps its slice of the four weight matrices
// (H/G output rows of every layer, 37 KB for D=64, H=512) in shared memory for the whole solve, so a
// dynamics evaluation costs four small dot-product phases separated by grid barriers; the layer
// activations (B x H) travel through L2.  The RK bookkeeping on the tiny (B x D) state — stage
// combination, error ratio, accept/reject, step size, dense output — is replicated in every CTA
// with identical arithmetic, which keeps all CTAs in lock-step without extra barriers.
// The dynamics are exact fp32 (the solver runs at rtol=atol=1e-3 and its accept/reject decisions
// must track the oracle's).
#include "common.cuh"
#include "dopri5.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCtas = 64;

struct LatentParams {
  const float* W[4];
  const float* b[4];
  int B, D, H, G;
  float* act0;          // (B,H) shared between CTAs
  float* act1;          // (B,H)
  float* kglob;         // (7,B,D): stage derivatives, written by slices
  float* priv;          // G private state blocks
  unsigned* barrier;    // monotonically increasing arrival counter
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Grid-wide barrier (all CTAs are co-resident: cooperative launch).  `epoch` counts the barriers
// this CTA has passed; the counter never resets.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned nblocks, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = epoch * nblocks;
    while (ld_volatile_u32(counter) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// out[b][row0 + r] = act(sum_k Wslice[r][k] * in[b][k] + bias) for the CTA's rows; `in` is in shared
// memory (B x K), Wslice in shared memory (rows x K).  One warp per (row, batch element) item.
__device__ void dense_slice(const float* __restrict__ Ws, const float* __restrict__ bias_s, const float* __restrict__ in_s,
                            float* __restrict__ out, int ld_out, int row0, int rows, int B, int K, bool do_tanh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int items = rows * B;
  for (int it = warp; it < items; it += kWarps) {
    const int r = it / B, b = it - r * B;
    const float* w = Ws + (size_t)r * K;
    const float* x = in_s + (size_t)b * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w[k], x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + bias_s[r];
      out[(size_t)b * ld_out + row0 + r] = do_tanh ? tanhf(v) : v;
    }
  }
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
latent_ode_kernel(LatentParams p, const float* __restrict__ z0, const double* __restrict__ times, int nT, float rtol,
                  float atol, float* __restrict__ out, int32_t* __restrict__ info, int max_steps) {
  extern __shared__ float smem[];
  __shared__ double red[kWarps];
  __shared__ double s_t0, s_t1, s_dt;
  __shared__ float s_dts;
  const int B = p.B, D = p.D, H = p.H, G = p.G;
  const int n = B * D;
  const int tid = threadIdx.x;
  const int cta = blockIdx.x;
  // rows of the hidden layers / of the output layer owned by this CTA
  const int rh = (H + G - 1) / G, rd = (D + G - 1) / G;
  const int h0r = min(cta * rh, H), h1r = min(h0r + rh, H);
  const int d0r = min(cta * rd, D), d1r = min(d0r + rd, D);
  const int nh = h1r - h0r, nd = d1r - d0r;
  // shared memory carve-up: weight slices, biases, activation staging buffer
  float* sW0 = smem;                          // nh x D
  float* sW1 = sW0 + (size_t)rh * D;          // nh x H
  float* sW2 = sW1 + (size_t)rh * H;          // nh x H
  float* sW3 = sW2 + (size_t)rh * H;          // nd x H
  float* sb0 = sW3 + (size_t)rd * H;
  float* sb1 = sb0 + rh;
  float* sb2 = sb1 + rh;
  float* sb3 = sb2 + rh;
  float* sact = sb3 + rd;                     // B x max(H, D)
  for (int i = tid; i < nh * D; i += kThreads) sW0[i] = p.W[0][(size_t)h0r * D + i];
  for (int i = tid; i < nh * H; i += kThreads) {
    sW1[i] = p.W[1][(size_t)h0r * H + i];
    sW2[i] = p.W[2][(size_t)h0r * H + i];
  }
  for (int i = tid; i < nd * H; i += kThreads) sW3[i] = p.W[3][(size_t)d0r * H + i];
  for (int i = tid; i < nh; i += kThreads) {
    sb0[i] = p.b[0][h0r + i];
    sb1[i] = p.b[1][h0r + i];
    sb2[i] = p.b[2][h0r + i];
  }
  for (int i = tid; i < nd; i += kThreads) sb3[i] = p.b[3][d0r + i];
  __syncthreads();

  // private (per-CTA, identical content in every CTA) state block
  float* y0 = p.priv + (size_t)cta * 14 * n;
  float* k = y0 + n;                 // 7 stage derivatives
  float* ys = k + 7 * (size_t)n;     // stage input / y1
  float* ymid = ys + n;
  float* yprev = ymid + n;           // y0 of the last ACCEPTED step (dense output)
  float* f0prev = yprev + n;
  float* y1acc = f0prev + n;
  float* f1acc = y1acc + n;
  unsigned epoch = 0;

  // f(z): Linear(D,H) tanh Linear(H,H) tanh Linear(H,H) tanh Linear(H,D)   (latent_ode_model.py:129-147)
  auto dynamics = [&](const float* z, int slot) {
    for (int i = tid; i < n; i += kThreads) sact[i] = z[i];
    __syncthreads();
    dense_slice(sW0, sb0, sact, p.act0, H, h0r, nh, B, D, true);
    grid_barrier(p.barrier, G, epoch);
    for (int i = tid; i < B * H; i += kThreads) sact[i] = __ldcg(p.act0 + i);
    __syncthreads();
    dense_slice(sW1, sb1, sact, p.act1, H, h0r, nh, B, H, true);
    grid_barrier(p.barrier, G, epoch);
    for (int i = tid; i < B * H; i += kThreads) sact[i] = __ldcg(p.act1 + i);
    __syncthreads();
    dense_slice(sW2, sb2, sact, p.act0, H, h0r, nh, B, H, true);
    grid_barrier(p.barrier, G, epoch);
    for (int i = tid; i < B * H; i += kThreads) sact[i] = __ldcg(p.act0 + i);
    __syncthreads();
    float* kg = p.kglob + (size_t)slot * n;
    dense_slice(sW3, sb3, sact, kg, D, d0r, nd, B, H, false);
    grid_barrier(p.barrier, G, epoch);
    for (int i = tid; i < n; i += kThreads) k[(size_t)slot * n + i] = __ldcg(kg + i);
    __syncthreads();
  };

  int nfe = 0, accepted = 0, rejected = 0, status = CASPR_OK;
  for (int i = tid; i < n; i += kThreads) {
    float v = z0[i];
    y0[i] = v;
    if (cta == 0) out[i] = v;                         // solution[0] = y0
  }
  __syncthreads();
  dynamics(y0, 0);                                    // f0
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
  dynamics(ys, 1);                                    // f1 probe (slot 1 is overwritten by the first step)
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
    s_dts = 0.f;
  }
  __syncthreads();

  for (int io = 1; io < nT && status == CASPR_OK; ++io) {
    const double t_out = times[io];
    int steps_here = 0;
    while (t_out > s_t1) {
      if (steps_here++ >= max_steps) { status = CASPR_ESOLVER_MAXSTEPS; break; }
      const double t0 = s_t1, dt = s_dt;
      if (!(t0 + dt > t0)) { status = CASPR_ESOLVER_DT; break; }
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
        dynamics(ys, s + 1);
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
        if (accept) s_dts = dtf;                       // dt of the step the dense output belongs to
      }
      __syncthreads();
    }
    if (status != CASPR_OK) break;
    // dense output at t_out (abscissa formed in fp32 like _interp_evaluate)
    const float t0f = (float)s_t0, t1f = (float)s_t1, tf = (float)t_out;
    const float x = __fdiv_rn(__fsub_rn(tf, t0f), __fsub_rn(t1f, t0f));
    const float dts = s_dts;
    if (cta == 0)
      for (int i = tid; i < n; i += kThreads)
        out[(size_t)io * n + i] =
            dopri5::interp_eval(yprev[i], y1acc[i], ymid[i], f0prev[i], f1acc[i], dts, x);
    __syncthreads();
  }
  if (cta == 0 && tid == 0) {
    info[INFO_STATUS] = status;
    info[INFO_NFE] = nfe;
    info[INFO_ACCEPTED] = accepted;
    info[INFO_REJECTED] = rejected;
    info[INFO_DONE] = 1;
  }
}

struct Sizes {
  int G;
  size_t smem_bytes;
  size_t off_act0, off_act1, off_kglob, off_priv, off_barrier, off_times, total;
};

Sizes make_sizes(int B, int D, int H) {
  Sizes z;
  z.G = kMaxCtas;
  const int rh = (H + z.G - 1) / z.G, rd = (D + z.G - 1) / z.G;
  const int kmax = H > D ? H : D;
  z.smem_bytes = ((size_t)rh * D + 2 * (size_t)rh * H + (size_t)rd * H + 3 * rh + rd + (size_t)B * kmax) * sizeof(float);
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t r = p; p += align_up(bytes, 256); return r; };
  const size_t n = (size_t)B * D;
  z.off_barrier = take(256);
  z.off_times = take(64 * sizeof(double));
  z.off_act0 = take((size_t)B * H * 4);
  z.off_act1 = take((size_t)B * H * 4);
  z.off_kglob = take(7 * n * 4);
  z.off_priv = take((size_t)z.G * 14 * n * 4);
  z.total = p;
  return z;
}

}  // namespace

extern "C" size_t caspr_latent_ode_workspace_bytes(int B, int D, int H) {
  if (B <= 0 || D <= 0 || H <= 0) return 0;
  return make_sizes(B, D, H).total;
}

extern "C" int caspr_latent_ode_solve(const float* z0, int B, int D, int H, const float* W0,
                                      const float* b0, const float* W1, const float* b1,
                                      const float* W2, const float* b2, const float* W3,
                                      const float* b3, const double* h_times, int nT, float rtol,
                                      float atol, float* out, int32_t* info, int32_t* h_info,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(z0 && W0 && b0 && W1 && b1 && W2 && b2 && W3 && b3 && h_times && out && info && h_info);
  CASPR_REQUIRE(workspace && B > 0 && D > 0 && H > 0 && nT >= 1 && nT <= 64);
  CASPR_REQUIRE(((uintptr_t)workspace & 255) == 0);
  for (int i = 1; i < nT; ++i) CASPR_REQUIRE(h_times[i] > h_times[i - 1]);
  const Sizes z = make_sizes(B, D, H);
  CASPR_REQUIRE(z.smem_bytes <= 200 * 1024);            // B x max(H,D) staging + weight slices must fit
  if (workspace_bytes < z.total) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  double* d_times = (double*)(base + z.off_times);
  if (cudaMemsetAsync(base + z.off_barrier, 0, 256, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaMemcpyAsync(d_times, h_times, nT * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  LatentParams p;
  p.W[0] = W0; p.W[1] = W1; p.W[2] = W2; p.W[3] = W3;
  p.b[0] = b0; p.b[1] = b1; p.b[2] = b2; p.b[3] = b3;
  p.B = B; p.D = D; p.H = H; p.G = z.G;
  p.act0 = (float*)(base + z.off_act0);
  p.act1 = (float*)(base + z.off_act1);
  p.kglob = (float*)(base + z.off_kglob);
  p.priv = (float*)(base + z.off_priv);
  p.barrier = (unsigned*)(base + z.off_barrier);
  if (z.smem_bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(latent_ode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)z.smem_bytes) != cudaSuccess)
      return CASPR_EINVAL;
  }
  int max_steps = 100000;
  void* args[] = {&p, (void*)&z0, (void*)&d_times, &nT, &rtol, &atol, (void*)&out, (void*)&info, &max_steps};
  CASPR_COUNT();
  if (cudaLaunchCooperativeKernel((const void*)latent_ode_kernel, dim3(z.G), dim3(kThreads), args, z.smem_bytes, s) !=
      cudaSuccess)
    return CASPR_ELAUNCH;
  if (cudaMemcpyAsync(h_info, info, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  return h_info[INFO_STATUS];
}
