// Dormand-Prince 5(4) constants and the scalar step-size controller of torchdiffeq 0.0.1
// (restated in oracle/odeint001.py; call sites in the reference:
// caspr/models/latent_ode_model.py:98 and caspr/models/cnf.py:102-119).
#pragma once
#include <math.h>

namespace dopri5 {

// Python doubles of the tableau rounded to fp32: torch multiplies the fp32 0-d tensor `dt` by
// each Python float in the state dtype.
#define DP_F(x) ((float)(x))
static __device__ __constant__ const float kAlpha[6] = {DP_F(1.0 / 5), DP_F(3.0 / 10), DP_F(4.0 / 5),
                                                 DP_F(8.0 / 9), 1.f, 1.f};
static __device__ __constant__ const float kBeta[6][6] = {
    {DP_F(1.0 / 5), 0, 0, 0, 0, 0},
    {DP_F(3.0 / 40), DP_F(9.0 / 40), 0, 0, 0, 0},
    {DP_F(44.0 / 45), DP_F(-56.0 / 15), DP_F(32.0 / 9), 0, 0, 0},
    {DP_F(19372.0 / 6561), DP_F(-25360.0 / 2187), DP_F(64448.0 / 6561), DP_F(-212.0 / 729), 0, 0},
    {DP_F(9017.0 / 3168), DP_F(-355.0 / 33), DP_F(46732.0 / 5247), DP_F(49.0 / 176),
     DP_F(-5103.0 / 18656), 0},
    {DP_F(35.0 / 384), 0.f, DP_F(500.0 / 1113), DP_F(125.0 / 192), DP_F(-2187.0 / 6784),
     DP_F(11.0 / 84)},
};
static __device__ __constant__ const float kCErr[7] = {
    DP_F(35.0 / 384 - 1951.0 / 21600),
    0.f,
    DP_F(500.0 / 1113 - 22642.0 / 50085),
    DP_F(125.0 / 192 - 451.0 / 720),
    DP_F(-2187.0 / 6784 - -12231.0 / 42400),
    DP_F(11.0 / 84 - 649.0 / 6300),
    DP_F(-1.0 / 60.0),
};
static __device__ __constant__ const float kCMid[7] = {
    DP_F(6025192743.0 / 30085553152.0 / 2),
    0.f,
    DP_F(51252292925.0 / 65400821598.0 / 2),
    DP_F(-2691868925.0 / 45128329728.0 / 2),
    DP_F(187940372067.0 / 1594534317056.0 / 2),
    DP_F(-1776094331.0 / 19743644256.0 / 2),
    DP_F(11237099.0 / 235043384.0 / 2),
};
#undef DP_F

// number of beta coefficients used by stage s (0-based): s+1
// y_stage = y0 + sum_{j<=s} (dt*beta[s][j]) * k[j], summed left to right from 0 with every
// product and sum rounded separately (torch evaluates them as separate fp32 kernels).
__device__ __forceinline__ float stage_combine(float y0, float dt, const float* kv, int s) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j <= s) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(dt, kBeta[s][j]), kv[j]));
  return __fadd_rn(y0, acc);
}
__device__ __forceinline__ float weighted7(float dt, const float* c, const float* kv) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(dt, c[j]), kv[j]));
  return acc;
}

// _optimal_step_size(last_step, mean_error_ratio, safety=.9, ifactor=10, dfactor=.2, order=5)
__device__ __forceinline__ double optimal_step(double dt, float ratio) {
  if (ratio == 0.f) return dt * 10.0;
  double dfactor = ratio < 1.f ? 1.0 : 0.2;
  double er = (double)sqrtf(ratio);
  double factor = fmax(1.0 / 10.0, fmin(pow(er, 1.0 / 5.0) / 0.9, 1.0 / dfactor));
  return dt / factor;
}

// Quartic dense output of a step (torchdiffeq _interp_fit + _interp_evaluate), x in [0,1].
__device__ __forceinline__ float interp_eval(float y0, float y1, float ymid, float f0, float f1,
                                             float dt, float x) {
  // a,b,c,d,e exactly as _interp_fit builds them (left-to-right sums from 0)
  float a = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(-2.f, dt), f0),
                                                      __fmul_rn(__fmul_rn(2.f, dt), f1)),
                                            __fmul_rn(-8.f, y0)),
                                  __fmul_rn(-8.f, y1)),
                        __fmul_rn(16.f, ymid));
  float b = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(5.f, dt), f0),
                                                      __fmul_rn(__fmul_rn(-3.f, dt), f1)),
                                            __fmul_rn(18.f, y0)),
                                  __fmul_rn(14.f, y1)),
                        __fmul_rn(-32.f, ymid));
  float c = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(-4.f, dt), f0),
                                                      __fmul_rn(dt, f1)),
                                            __fmul_rn(-11.f, y0)),
                                  __fmul_rn(-5.f, y1)),
                        __fmul_rn(16.f, ymid));
  float d = __fmul_rn(dt, f0);
  float x2 = __fmul_rn(x, x), x3 = __fmul_rn(x2, x), x4 = __fmul_rn(x3, x);
  float r = __fmul_rn(a, x4);
  r = __fadd_rn(r, __fmul_rn(b, x3));
  r = __fadd_rn(r, __fmul_rn(c, x2));
  r = __fadd_rn(r, __fmul_rn(d, x));
  r = __fadd_rn(r, y0);
  return r;
}

}  // namespace dopri5

// layout of the 8-int info block every solver writes
enum { INFO_STATUS = 0, INFO_NFE = 1, INFO_ACCEPTED = 2, INFO_REJECTED = 3, INFO_DONE = 4,
       INFO_SKIP = 5, INFO_AUX0 = 6, INFO_AUX1 = 7 };
