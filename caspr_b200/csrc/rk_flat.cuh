// dopri5 bookkeeping over a flat fp32 state vector made of up to four controlled segments (torchdiffeq 0.0.1
// semantics: one mean error ratio per state tensor, accept iff all <= 1).  Used by the latent-ODE adjoint, whose
// scalar tolerances put every augmented tensor under step control (latent_ode_model.py:98 via
// oracle/odeint001.py::_AdjointMethod).  Shares CnfState with the CNF solver.
#pragma once
#include "common.cuh"
#include "dopri5.cuh"
#include "cnf_state.cuh"

namespace {

struct FlatSegs {
  unsigned long long begin[4], end[4];     // element ranges of the controlled tensors
  int nseg;
};

__device__ __forceinline__ void flat_block_add(double v, double* dst) {
  __shared__ double s_w[32];
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    atomicAdd(dst, t);
  }
}

// ys[i] = y0[i] + sum_j (dt*beta[s][j]) k_j[i] for the leading `nelem` entries (the part the dynamics read)
__global__ void __launch_bounds__(256)
flat_stage_kernel(const float* __restrict__ y0, const float* __restrict__ k, size_t kstride, size_t nelem, int stage,
                  const CnfState* __restrict__ st, float* __restrict__ ys) {
  if (st->done) return;
  const float dt = (float)st->dt;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += (size_t)gridDim.x * blockDim.x) {
    float kc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) kc[j] = j < stage ? k[(size_t)j * kstride + i] : 0.f;
    ys[i] = stage == 0 ? y0[i] : dopri5::stage_combine(y0[i], dt, kc, stage - 1);
  }
}

// sums[s] += sum over segment s of (k0 / (atol + |y0| rtol))^2   (initial-step heuristic, d1)
__global__ void __launch_bounds__(256)
flat_init_norm_kernel(const float* __restrict__ y0, const float* __restrict__ k0, FlatSegs segs, float rtol, float atol,
                      double* sums) {
  for (int s = 0; s < segs.nseg; ++s) {
    double acc = 0.0;
    for (size_t i = segs.begin[s] + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.end[s];
         i += (size_t)gridDim.x * blockDim.x) {
      const float sc = __fadd_rn(atol, __fmul_rn(fabsf(y0[i]), rtol));
      const float r = __fdiv_rn(k0[i], sc);
      acc += (double)r * (double)r;
    }
    flat_block_add(acc, sums + s);
  }
}

// first step = (0.01 / max d1)^(1/5): the form _select_initial_step reduces to when one state tensor has zero
// dynamics (see cnf_init_controller_kernel); t bookkeeping reset for a new integration interval.
__global__ void flat_init_controller_kernel(CnfState* st, FlatSegs segs, double* sums, double t_start, double t_end) {
  float d1 = 0.f;
  for (int s = 0; s < segs.nseg; ++s) {
    const float d = (float)sqrt(sums[s]) / sqrtf((float)(segs.end[s] - segs.begin[s]));
    d1 = fmaxf(d1, d);
    sums[s] = 0.0;
  }
  const float dt = ((double)d1 < 1e-5) ? 1e-6f : powf(__fdiv_rn(0.01f, d1), 1.0f / 5.0f);
  st->t = t_start;
  st->t_prev = t_start;
  st->t_end = t_end;
  st->dt = (double)dt;
  st->nfe += 3;
  st->status = CASPR_OK;
  st->fin_step = -1;
  st->accept = 0;
  st->done = (t_end > t_start) ? 0 : 1;
  st->first_dt = dt;
}

// y1 = y0 + sum_j (dt*c_sol[j]) k_j ; err = sum_j (dt*c_err[j]) k_j ; per-segment sums of (err/tol)^2
__global__ void __launch_bounds__(256)
flat_error_kernel(const float* __restrict__ y0, const float* __restrict__ k, size_t kstride, FlatSegs segs, float rtol,
                  float atol, const CnfState* __restrict__ st, double* sums) {
  if (st->done) return;
  const float dt = (float)st->dt;
  for (int s = 0; s < segs.nseg; ++s) {
    double acc = 0.0;
    for (size_t i = segs.begin[s] + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < segs.end[s];
         i += (size_t)gridDim.x * blockDim.x) {
      float kc[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) kc[j] = k[(size_t)j * kstride + i];
      const float y = y0[i];
      const float y1 = dopri5::stage_combine(y, dt, kc, 5);
      const float err = dopri5::weighted7(dt, dopri5::kCErr, kc);
      const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y), fabsf(y1))));
      const float r = __fdiv_rn(err, tol);
      acc += (double)__fmul_rn(r, r);
    }
    flat_block_add(acc, sums + s);
  }
}

__global__ void flat_controller_kernel(CnfState* st, FlatSegs segs, double* sums, int step_id) {
  if (st->done) return;
  bool accept = true, isnan_ = false;
  float ratio = 0.f;
  for (int s = 0; s < segs.nseg; ++s) {
    const float r = (float)(sums[s] / (double)(segs.end[s] - segs.begin[s]));
    sums[s] = 0.0;
    if (!(r <= 1.f)) accept = false;
    if (r != r) isnan_ = true;
    ratio = fmaxf(ratio, r);
  }
  st->nfe += 6;
  const double t0 = st->t, dt = st->dt;
  st->t_prev = t0;
  st->dt_prev = (float)dt;
  st->accept = accept ? 1 : 0;
  st->fin_step = step_id;
  if (accept) { st->t = t0 + dt; st->accepted++; } else { st->rejected++; }
  if (isnan_) {
    st->status = CASPR_ESOLVER_NONFINITE;
    st->done = 1;
    return;
  }
  const double dt_next = dopri5::optimal_step(dt, ratio);
  st->dt = dt_next;
  if (accept && !(st->t_end > st->t)) {
    st->done = 1;
  } else if (!(st->t + dt_next > st->t)) {
    st->status = CASPR_ESOLVER_DT;
    st->done = 1;
  }
}

// Accepted step: y0 <- y1 and FSAL shift k_0 <- k_6; on the step that passes t_end the dense-output value at
// t_end goes to `out` instead (torchdiffeq _interp_fit / _interp_evaluate).
__global__ void __launch_bounds__(256)
flat_finalize_kernel(float* __restrict__ y0, float* __restrict__ k, size_t kstride, size_t nelem, int step_id,
                     const CnfState* __restrict__ st, float* __restrict__ out) {
  if (st->fin_step != step_id || !st->accept) return;
  const int finished = st->done && st->status == CASPR_OK;
  const float dt = st->dt_prev;
  float xq = 0.f;
  if (finished) {
    const float t0f = (float)st->t_prev, t1f = (float)st->t, tf = (float)st->t_end;
    xq = __fdiv_rn(__fsub_rn(tf, t0f), __fsub_rn(t1f, t0f));
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += (size_t)gridDim.x * blockDim.x) {
    float kc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) kc[j] = k[(size_t)j * kstride + i];
    const float y = y0[i];
    const float y1 = dopri5::stage_combine(y, dt, kc, 5);
    if (finished) {
      const float ymid = __fadd_rn(y, dopri5::weighted7(dt, dopri5::kCMid, kc));
      out[i] = dopri5::interp_eval(y, y1, ymid, kc[0], kc[6], dt, xq);
    } else {
      y0[i] = y1;
      k[i] = kc[6];
    }
  }
}

}  // namespace
