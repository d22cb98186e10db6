// Kernels and workspace layout of the CNF dopri5 solver, shared by the forward/reverse flow (cnf.cu) and the
// adjoint backward (cnf_adjoint.cu).  Everything lives in an anonymous namespace: each including translation
// unit gets its own copy.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "dopri5.cuh"
#include "cnf_state.cuh"
#include "cnf_tc.cuh"

namespace {

constexpr int kMaxHidden = 512;

// softplus (beta=1, threshold=20) and its derivative, as torch.nn.Softplus / its backward
// (odefunc.py:54, torch SoftplusBackward: z/(z+1) with z = exp(x), 1 above the threshold).
__device__ __forceinline__ void softplus_and_grad(float x, float& sp, float& dsp) {
  if (x > 20.f) {
    sp = x;
    dsp = 1.f;
  } else {
    float z = expf(x);
    sp = log1pf(z);
    dsp = __fdiv_rn(z, __fadd_rn(z, 1.f));
  }
}

// ------------------------------------------------------------------ MovingBatchNorm helpers
struct MbnDev {
  float w[3], b[3], mean[3], var[3];
  int present;
};

__device__ __forceinline__ void mbn_forward(const MbnDev& m, float* x, float& logdet) {
  logdet = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float lv = logf(m.var[c] + 1e-4f);                            // normalization.py:70
    float y = (x[c] - m.mean[c]) * expf(-0.5f * lv);
    x[c] = y * expf(m.w[c]) + m.b[c];                             // :74
    logdet += -0.5f * lv + m.w[c];                                // :103-108
  }
}
__device__ __forceinline__ void mbn_reverse(const MbnDev& m, float* y, float& logdet) {
  logdet = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float lv = logf(m.var[c] + 1e-4f);
    float v = (y[c] - m.b[c]) * expf(-m.w[c]);                    // :92
    y[c] = v * expf(0.5f * lv) + m.mean[c];                       // :94
    logdet += -0.5f * lv + m.w[c];
  }
}

// ------------------------------------------------------------------------------- prepare
// y0 = MBN_pre(x_in); logp0 = logp_in -/+ logdet (or 0).  State layout: float4 (x,y,z,logp).
__global__ void __launch_bounds__(256)
cnf_init_state_kernel(const float* __restrict__ x_in, const float* __restrict__ logp_in, int n,
                      MbnDev pre, int reverse, float4* __restrict__ y0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[3] = {x_in[3 * (size_t)i], x_in[3 * (size_t)i + 1], x_in[3 * (size_t)i + 2]};
  float lp = logp_in ? logp_in[i] : 0.f;
  if (pre.present) {
    float ld;
    if (reverse) { mbn_reverse(pre, x, ld); if (logp_in) lp = lp + ld; }
    else { mbn_forward(pre, x, ld); if (logp_in) lp = lp - ld; }
  }
  y0[i] = make_float4(x[0], x[1], x[2], lp);
}

// stage time exactly as torchdiffeq forms it: ti = t0.to(fp32) + alpha_i * dt.to(fp32); the dynamics see -ti
// when integrating backwards (odeint001.odeint negates time).
__device__ __forceinline__ float cnf_stage_time(const CnfState* st, int stage, int reverse) {
  float ti = (float)st->t;
  if (stage > 0) ti = __fadd_rn(ti, __fmul_rn(dopri5::kAlpha[stage - 1], (float)st->dt));
  return reverse ? -ti : ti;
}

// gate[f][j] = sigmoid(Gc[f][j] + wg_t[j]*t), biasf[f][j] = b[j]*gate + (Bc[f][j] + wb_t[j]*t)
// for the concatenated output channels j of the four layers (H,H,H,3).
__global__ void __launch_bounds__(256)
cnf_hyper_stage_kernel(const float* __restrict__ Gc, const float* __restrict__ Bc,
                       const float* __restrict__ wg_t, const float* __restrict__ wb_t,
                       const float* __restrict__ lbias, int frames, int ctot, int ld, int stage_first,
                       int reverse, const CnfState* __restrict__ st, const float* __restrict__ col_scale,
                       float* __restrict__ gate_all, float* __restrict__ biasf_all) {
  if (st->done) return;
  // blockIdx.y selects the RK stage; every stage has its own slot of (frames x ld) gates / biases so that the
  // stages of a step can be evaluated by independent streams
  const int stage = stage_first + blockIdx.y;
  float* gate = gate_all + (size_t)stage * frames * ld;
  float* biasf = biasf_all + (size_t)stage * frames * ld;
  // stage time exactly as torchdiffeq forms it: ti = t0.to(fp32) + alpha_i * dt.to(fp32); the
  // dynamics see -ti when integrating backwards (odeint001.odeint negates time).
  const float t = cnf_stage_time(st, stage, reverse);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)frames * ld;
  if (i >= total) return;
  const int j = (int)(i % ld);
  if (j >= ctot) return;
  const float g = 1.f / (1.f + expf(-(Gc[i] + wg_t[j] * t)));
  // the tensor-core engine folds its exact power-of-two operand scales into the gate it reads
  gate[i] = col_scale ? g * col_scale[j] : g;
  biasf[i] = lbias[j] * g + (Bc[i] + wb_t[j] * t);
}

// Layer 0 (3 -> H) with the RK stage input formed on the fly.  One warp per point.
//   ys = y0 + sum_j (dt*beta[s][j]) k_j   (xyz only: the dynamics do not depend on logp)
//   pre = (W0 ys)*gate + biasf ; h = softplus(pre) ; v = softplus'(pre) * gate * (W0 e)
__global__ void __launch_bounds__(256)
cnf_layer0_kernel(const float4* __restrict__ y0, const float4* __restrict__ kbuf, size_t kstride,
                  const float* __restrict__ e, const float* __restrict__ W0, int H, int n, int P,
                  int stage, const float* __restrict__ gate, const float* __restrict__ biasf,
                  int ld_hyper, const CnfState* __restrict__ st, float* __restrict__ Hout,
                  float* __restrict__ Vout, unsigned* __restrict__ colmax = nullptr) {
  if (st->done) return;
  __shared__ float sW[kMaxHidden * 3];
  for (int i = threadIdx.x; i < H * 3; i += blockDim.x) sW[i] = W0[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float dt = (float)st->dt;
  // optional per-channel max |h|, |v| over all points (operand scales of the tensor-core weight gradient)
  float cmx[kMaxHidden / 32];
#pragma unroll
  for (int i = 0; i < kMaxHidden / 32; ++i) cmx[i] = 0.f;
  for (int pt = blockIdx.x * wpb + (threadIdx.x >> 5); pt < n; pt += gridDim.x * wpb) {
    float4 y = y0[pt];
    float ys[3] = {y.x, y.y, y.z};
    if (stage > 0) {
      float kx[6], ky[6], kz[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < stage) {
          float4 kv = kbuf[(size_t)j * kstride + pt];
          kx[j] = kv.x; ky[j] = kv.y; kz[j] = kv.z;
        } else {
          kx[j] = ky[j] = kz[j] = 0.f;
        }
      }
      ys[0] = dopri5::stage_combine(y.x, dt, kx, stage - 1);
      ys[1] = dopri5::stage_combine(y.y, dt, ky, stage - 1);
      ys[2] = dopri5::stage_combine(y.z, dt, kz, stage - 1);
    }
    const float e0 = e[3 * (size_t)pt], e1 = e[3 * (size_t)pt + 1], e2 = e[3 * (size_t)pt + 2];
    const int f = pt / P;
    const float* g = gate + (size_t)f * ld_hyper;
    const float* bf = biasf + (size_t)f * ld_hyper;
    float* ho = Hout + (size_t)pt * H;
    float* vo = Vout + (size_t)pt * H;
#pragma unroll
    for (int i = 0; i < kMaxHidden / 32; ++i) {
      const int j = lane + 32 * i;
      if (j >= H) break;
      const float w0 = sW[3 * j], w1 = sW[3 * j + 1], w2 = sW[3 * j + 2];
      const float a = fmaf(w2, ys[2], fmaf(w1, ys[1], w0 * ys[0]));
      const float ta = fmaf(w2, e2, fmaf(w1, e1, w0 * e0));
      const float gj = g[j];
      const float pre = fmaf(a, gj, bf[j]);
      float sp, dsp;
      softplus_and_grad(pre, sp, dsp);
      const float tv = dsp * gj * ta;
      ho[j] = sp;
      vo[j] = tv;
      cmx[i] = fmaxf(cmx[i], fmaxf(fabsf(sp), fabsf(tv)));
    }
  }
  if (colmax) {
#pragma unroll
    for (int i = 0; i < kMaxHidden / 32; ++i) {
      const int j = lane + 32 * i;
      if (j < H) atomic_max_nonneg(colmax + j, cmx[i]);
    }
  }
}

// Mid layers (H -> H): [h ; v] rows through the same weights, fp32 SIMT.
// CTA tile: 64 points x 128 output channels, k-slab 16, 256 threads as 16 (channels) x 16
// (points); thread tile 4 points x 8 channels x {h, v}.
constexpr int kMidBM = 64, kMidBN = 128, kMidBK = 16;

// MODE 0: forward layer (gate/bias, softplus, tangent chain rule).  MODE 1: the same, and the raw products
// (W h, W v) are also stored (Araw, Vraw) for the adjoint backward.  MODE 2: plain product (Hout, Vout) =
// (Hin, Vin) . W^T, used by the adjoint with the transposed weights.
enum { kMidForward = 0, kMidForwardKeepRaw = 1, kMidPlain = 2 };
template <int MODE>
__global__ void __launch_bounds__(256, 2)
cnf_mid_layer_kernel(const float* __restrict__ Hin, const float* __restrict__ Vin,
                     const float* __restrict__ W, int H, int n, int P,
                     const float* __restrict__ gate, const float* __restrict__ biasf, int ld_hyper,
                     const CnfState* __restrict__ st, float* __restrict__ Hout,
                     float* __restrict__ Vout, float* __restrict__ Araw, float* __restrict__ Vraw) {
  if (st->done) return;
  __shared__ __align__(16) float Ah[2][kMidBK][kMidBM + 4];
  __shared__ __align__(16) float Av[2][kMidBK][kMidBM + 4];
  __shared__ __align__(16) float Bs[2][kMidBK][kMidBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int pt0 = blockIdx.x * kMidBM;
  const int col0 = blockIdx.y * kMidBN;

  float acc_h[4][8], acc_v[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_h[i][j] = 0.f; acc_v[i][j] = 0.f; }

  // loaders: A tiles 64 rows x 16 k as one float4 (along k) per thread; W tile 128 rows x 16 k
  // as two float4 per thread.
  const int a_r = tid >> 2, a_kq = tid & 3;
  float4 rah, rav, rb0, rb1;
  auto load_tiles = [&](int k0) {
    const int pt = pt0 + a_r;
    if (pt < n) {
      rah = *reinterpret_cast<const float4*>(Hin + (size_t)pt * H + k0 + a_kq * 4);
      rav = *reinterpret_cast<const float4*>(Vin + (size_t)pt * H + k0 + a_kq * 4);
    } else {
      rah = make_float4(0.f, 0.f, 0.f, 0.f);
      rav = rah;
    }
    rb0 = *reinterpret_cast<const float4*>(W + (size_t)(col0 + a_r) * H + k0 + a_kq * 4);
    rb1 = *reinterpret_cast<const float4*>(W + (size_t)(col0 + 64 + a_r) * H + k0 + a_kq * 4);
  };
  auto store_tiles = [&](int buf) {
    const int kk = a_kq * 4;
    Ah[buf][kk + 0][a_r] = rah.x; Ah[buf][kk + 1][a_r] = rah.y;
    Ah[buf][kk + 2][a_r] = rah.z; Ah[buf][kk + 3][a_r] = rah.w;
    Av[buf][kk + 0][a_r] = rav.x; Av[buf][kk + 1][a_r] = rav.y;
    Av[buf][kk + 2][a_r] = rav.z; Av[buf][kk + 3][a_r] = rav.w;
    Bs[buf][kk + 0][a_r] = rb0.x; Bs[buf][kk + 1][a_r] = rb0.y;
    Bs[buf][kk + 2][a_r] = rb0.z; Bs[buf][kk + 3][a_r] = rb0.w;
    Bs[buf][kk + 0][64 + a_r] = rb1.x; Bs[buf][kk + 1][64 + a_r] = rb1.y;
    Bs[buf][kk + 2][64 + a_r] = rb1.z; Bs[buf][kk + 3][64 + a_r] = rb1.w;
  };

  const int nk = H / kMidBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * kMidBK);
#pragma unroll
    for (int k = 0; k < kMidBK; ++k) {
      const float4 ah = *reinterpret_cast<const float4*>(&Ah[buf][k][ty * 4]);
      const float4 av = *reinterpret_cast<const float4*>(&Av[buf][k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a_h[4] = {ah.x, ah.y, ah.z, ah.w};
      const float a_v[4] = {av.x, av.y, av.z, av.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc_h[i][j] = fmaf(a_h[i], b[j], acc_h[i][j]);
          acc_v[i][j] = fmaf(a_v[i], b[j], acc_v[i][j]);
        }
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue: ConcatSquash gate/bias, softplus and the tangent's chain rule
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pt = pt0 + ty * 4 + i;
    if (pt >= n) continue;
    if (MODE == kMidPlain) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = half * 64 + tx * 4;
        *reinterpret_cast<float4*>(Hout + (size_t)pt * H + col0 + c) =
            make_float4(acc_h[i][half * 4], acc_h[i][half * 4 + 1], acc_h[i][half * 4 + 2], acc_h[i][half * 4 + 3]);
        *reinterpret_cast<float4*>(Vout + (size_t)pt * H + col0 + c) =
            make_float4(acc_v[i][half * 4], acc_v[i][half * 4 + 1], acc_v[i][half * 4 + 2], acc_v[i][half * 4 + 3]);
      }
      continue;
    }
    const int f = pt / P;
    const float* g = gate + (size_t)f * ld_hyper + col0;
    const float* bf = biasf + (size_t)f * ld_hyper + col0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = half * 64 + tx * 4;
      if (MODE == kMidForwardKeepRaw) {
        *reinterpret_cast<float4*>(Araw + (size_t)pt * H + col0 + c) =
            make_float4(acc_h[i][half * 4], acc_h[i][half * 4 + 1], acc_h[i][half * 4 + 2], acc_h[i][half * 4 + 3]);
        *reinterpret_cast<float4*>(Vraw + (size_t)pt * H + col0 + c) =
            make_float4(acc_v[i][half * 4], acc_v[i][half * 4 + 1], acc_v[i][half * 4 + 2], acc_v[i][half * 4 + 3]);
      }
      const float4 g4 = *reinterpret_cast<const float4*>(g + c);
      const float4 b4 = *reinterpret_cast<const float4*>(bf + c);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      float ho[4], vo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pre = fmaf(acc_h[i][half * 4 + j], gg[j], bb[j]);
        float sp, dsp;
        softplus_and_grad(pre, sp, dsp);
        ho[j] = sp;
        vo[j] = dsp * gg[j] * acc_v[i][half * 4 + j];
      }
      *reinterpret_cast<float4*>(Hout + (size_t)pt * H + col0 + c) = make_float4(ho[0], ho[1], ho[2], ho[3]);
      *reinterpret_cast<float4*>(Vout + (size_t)pt * H + col0 + c) = make_float4(vo[0], vo[1], vo[2], vo[3]);
    }
  }
}

// Last layer (H -> 3) + divergence.  One warp per point.
//   dy_c = (W3_c . h)*gate_c + biasf_c ; tang_c = gate_c * (W3_c . v) ; div = sum_c e_c tang_c
//   k = sign * (dy, -div)     (sign = -1 when integrating backwards: odeint negates f)
__global__ void __launch_bounds__(256)
cnf_last_layer_kernel(const float* __restrict__ Hin, const float* __restrict__ Vin,
                      const float* __restrict__ W3, int H, int n, int P,
                      const float* __restrict__ e, const float* __restrict__ gate,
                      const float* __restrict__ biasf, int ld_hyper, int reverse,
                      const CnfState* __restrict__ st, float4* __restrict__ kout,
                      float* __restrict__ raw8 = nullptr) {
  if (st->done) return;
  __shared__ float sW[kMaxHidden * 3];
  for (int i = threadIdx.x; i < H * 3; i += blockDim.x) sW[i] = W3[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int pt = blockIdx.x * wpb + (threadIdx.x >> 5); pt < n; pt += gridDim.x * wpb) {
    const float* h = Hin + (size_t)pt * H;
    const float* v = Vin + (size_t)pt * H;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 h4 = *reinterpret_cast<const float4*>(h + k);
      const float4 v4 = *reinterpret_cast<const float4*>(v + k);
      const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float w0 = sW[k + q], w1 = sW[H + k + q], w2 = sW[2 * H + k + q];
        a0 = fmaf(w0, hh[q], a0); a1 = fmaf(w1, hh[q], a1); a2 = fmaf(w2, hh[q], a2);
        t0 = fmaf(w0, vv[q], t0); t1 = fmaf(w1, vv[q], t1); t2 = fmaf(w2, vv[q], t2);
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (lane == 0) {
      const int f = pt / P;
      const float* g = gate + (size_t)f * ld_hyper;
      const float* bf = biasf + (size_t)f * ld_hyper;
      const float dy0 = fmaf(a0, g[0], bf[0]);
      const float dy1 = fmaf(a1, g[1], bf[1]);
      const float dy2 = fmaf(a2, g[2], bf[2]);
      const float e0 = e[3 * (size_t)pt], e1 = e[3 * (size_t)pt + 1], e2 = e[3 * (size_t)pt + 2];
      const float div = (g[0] * t0) * e0 + (g[1] * t1) * e1 + (g[2] * t2) * e2;
      kout[pt] = reverse ? make_float4(-dy0, -dy1, -dy2, div) : make_float4(dy0, dy1, dy2, -div);
      if (raw8) {                                   // raw products W3.h, W3.v for the adjoint backward
        *reinterpret_cast<float4*>(raw8 + 8 * (size_t)pt) = make_float4(a0, a1, a2, 0.f);
        *reinterpret_cast<float4*>(raw8 + 8 * (size_t)pt + 4) = make_float4(t0, t1, t2, 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------------- reductions / control
__device__ __forceinline__ void block_accumulate2(double a, double b, double* dst_a, double* dst_b) {
  __shared__ double sa[32], sb[32];
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? sa[lane] : 0.0;
    b = lane < nw ? sb[lane] : 0.0;
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    if (lane == 0) { atomicAdd(dst_a, a); atomicAdd(dst_b, b); }
  }
}

// _select_initial_step pieces that survive the zero-dynamics context quirk (odeint001.py
// header): sum (f0/scale)^2 per state tensor, scale = atol + |y0|*rtol.  Also the d0 sums.
__global__ void __launch_bounds__(256)
cnf_init_norm_kernel(const float4* __restrict__ y0, const float4* __restrict__ k0, int n, float rtol,
                     float atol, CnfState* st) {
  double sx = 0.0, sl = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 y = y0[i], f = k0[i];
    const float yy[4] = {y.x, y.y, y.z, y.w}, ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float sc = __fadd_rn(atol, __fmul_rn(fabsf(yy[c]), rtol));
      const float r = __fdiv_rn(ff[c], sc);
      const double r2 = (double)r * (double)r;
      if (c < 3) sx += r2; else sl += r2;
    }
  }
  block_accumulate2(sx, sl, &st->sum_x, &st->sum_l);
}

__global__ void cnf_init_controller_kernel(CnfState* st, int n, float t_start, float t_end) {
  // d1 = rms(f0/scale) per tensor, in fp32 like torch
  const float d1x = (float)sqrt(st->sum_x) / sqrtf((float)n * 3.f);
  const float d1l = (float)sqrt(st->sum_l) / sqrtf((float)n);
  const float d1 = fmaxf(d1x, d1l);
  float dt;
  if ((double)d1 < 1e-5) {
    // h0 = 1e-6; the probe f1 is then a real evaluation the oracle uses.  The CNF dynamics are
    // never this flat in practice; fall back to the oracle's lower bound min(100*h0, h1>=1e-6).
    dt = 1e-6f;
  } else {
    // h0 = 0.01*max(d0/d1) = +inf through the context tensor (d1_ctx = 0), d2 is 0 or NaN and
    // never wins Python's max(): dt = min(100*h0, h1) = h1 = (0.01/max(d1))^(1/5).
    dt = powf(__fdiv_rn(0.01f, d1), 1.0f / 5.0f);
  }
  st->t = (double)t_start;
  st->t_end = (double)t_end;
  st->dt = (double)dt;
  st->sum_x = 0.0;
  st->sum_l = 0.0;
  st->nfe = 2;
  st->accepted = 0;
  st->rejected = 0;
  st->status = CASPR_OK;
  st->fin_step = -1;
  st->done = (st->t_end > st->t) ? 0 : 1;
  st->first_dt = dt;
}

// y1 = y0 + sum_j (dt*beta[5][j]) k_j ; err = sum_j (dt*c_err[j]) k_j ; ratio sums.
__global__ void __launch_bounds__(256)
cnf_error_kernel(const float4* __restrict__ y0, const float4* __restrict__ kbuf, size_t kstride,
                 int n, float rtol, float atol, CnfState* st, float4* __restrict__ y1out) {
  if (st->done) return;
  const float dt = (float)st->dt;
  double sx = 0.0, sl = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 y = y0[i];
    float4 kv[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) kv[j] = kbuf[(size_t)j * kstride + i];
    const float yy[4] = {y.x, y.y, y.z, y.w};
    float y1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float kc[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) kc[j] = c == 0 ? kv[j].x : (c == 1 ? kv[j].y : (c == 2 ? kv[j].z : kv[j].w));
      y1[c] = dopri5::stage_combine(yy[c], dt, kc, 5);
      const float err = dopri5::weighted7(dt, dopri5::kCErr, kc);
      const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(yy[c]), fabsf(y1[c]))));
      const float r = __fdiv_rn(err, tol);
      const double r2 = (double)__fmul_rn(r, r);
      if (c < 3) sx += r2; else sl += r2;
    }
    y1out[i] = make_float4(y1[0], y1[1], y1[2], y1[3]);
  }
  block_accumulate2(sx, sl, &st->sum_x, &st->sum_l);
}

__global__ void cnf_controller_kernel(CnfState* st, int n, int step_id) {
  if (st->done) return;
  const float rx = (float)(st->sum_x / ((double)n * 3.0));
  const float rl = (float)(st->sum_l / (double)n);
  st->sum_x = 0.0;
  st->sum_l = 0.0;
  st->nfe += 6;
  // accept iff every ratio <= 1 (NaN rejects); the context tensor's ratio is 0
  const bool accept = (rx <= 1.f) && (rl <= 1.f);
  const float ratio = (rx != rx || rl != rl) ? __int_as_float(0x7fc00000) : fmaxf(rx, rl);
  const double t0 = st->t, dt = st->dt;
  st->t_prev = t0;
  st->dt_prev = (float)dt;
  st->accept = accept ? 1 : 0;
  st->fin_step = step_id;
  if (accept) { st->t = t0 + dt; st->accepted++; } else { st->rejected++; }
  if (ratio != ratio) {                              // non-finite state (torchdiffeq asserts)
    st->status = CASPR_ESOLVER_NONFINITE;
    st->done = 1;
    return;
  }
  const double dt_next = dopri5::optimal_step(dt, ratio);
  st->dt = dt_next;
  if (accept && !(st->t_end > st->t)) {
    st->done = 1;
  } else if (!(st->t + dt_next > st->t)) {          // torchdiffeq: "underflow in dt"
    st->status = CASPR_ESOLVER_DT;
    st->done = 1;
  }
}

// After the controller: on an accepted step shift (y0,k0) <- (y1,k6) (FSAL); on the accepted step
// that passed t_end evaluate the quartic dense output at t_end, apply the MovingBatchNorm
// post-transform and write the result.
__global__ void __launch_bounds__(256)
cnf_finalize_kernel(float4* __restrict__ y0, float4* __restrict__ kbuf, size_t kstride,
                    const float4* __restrict__ y1buf, int n, int step_id, const CnfState* __restrict__ st,
                    MbnDev post, int reverse, int have_logp, float* __restrict__ x_out,
                    float* __restrict__ logp_out) {
  if (st->fin_step != step_id || !st->accept) return;
  const int finished = st->done && st->status == CASPR_OK;
  const float dt = st->dt_prev;
  float xq = 0.f;
  if (finished) {
    const float t0f = (float)st->t_prev, t1f = (float)st->t, tf = (float)st->t_end;
    xq = __fdiv_rn(__fsub_rn(tf, t0f), __fsub_rn(t1f, t0f));
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 y1 = y1buf[i];
    const float4 k6 = kbuf[6 * kstride + i];
    if (finished) {
      const float4 y = y0[i];
      float4 kv[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) kv[j] = kbuf[(size_t)j * kstride + i];
      const float yy[4] = {y.x, y.y, y.z, y.w};
      const float y1v[4] = {y1.x, y1.y, y1.z, y1.w};
      float out[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float kc[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) kc[j] = c == 0 ? kv[j].x : (c == 1 ? kv[j].y : (c == 2 ? kv[j].z : kv[j].w));
        const float ymid = __fadd_rn(yy[c], dopri5::weighted7(dt, dopri5::kCMid, kc));
        out[c] = dopri5::interp_eval(yy[c], y1v[c], ymid, kc[0], kc[6], dt, xq);
      }
      float lp = out[3];
      if (post.present) {
        float ld;
        if (reverse) { mbn_reverse(post, out, ld); lp = lp + ld; }
        else { mbn_forward(post, out, ld); lp = lp - ld; }
      }
      x_out[3 * (size_t)i] = out[0];
      x_out[3 * (size_t)i + 1] = out[1];
      x_out[3 * (size_t)i + 2] = out[2];
      if (logp_out && have_logp) logp_out[i] = lp;
    } else {
      y0[i] = y1;
      kbuf[i] = k6;
    }
  }
}

// Degenerate solve (t_end <= t_start): output = post(pre(x)).
__global__ void __launch_bounds__(256)
cnf_passthrough_kernel(const float4* __restrict__ y0, int n, MbnDev post, int reverse, int have_logp,
                       float* __restrict__ x_out, float* __restrict__ logp_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 y = y0[i];
  float out[3] = {y.x, y.y, y.z};
  float lp = y.w;
  if (post.present) {
    float ld;
    if (reverse) { mbn_reverse(post, out, ld); lp += ld; } else { mbn_forward(post, out, ld); lp -= ld; }
  }
  x_out[3 * (size_t)i] = out[0]; x_out[3 * (size_t)i + 1] = out[1]; x_out[3 * (size_t)i + 2] = out[2];
  if (logp_out && have_logp) logp_out[i] = lp;
}

// Hyper-network hoist: out[f][j] = W[j][1:1+C] . ctx[f] (+ bias[j]) for a handful of frames.  A skinny GEMM
// (frames x C x D with frames ~ 80).  Each CTA owns 16 output channels (two per warp, weight rows kept in
// registers) for one tile of 8 frames staged in shared memory (grid = channel blocks x frame tiles).
constexpr int kHoistMaxC = 2048;
constexpr int kHoistFrames = 8;
constexpr int kHoistChPerCta = 16;
__global__ void __launch_bounds__(256)
cnf_hyper_hoist_kernel(const float* __restrict__ ctx, int C, const float* __restrict__ W, int ldw,
                       const float* __restrict__ bias, int frames, int D, float* __restrict__ out, int ld_out) {
  extern __shared__ float s_ctx[];                               // kHoistFrames x C
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * kHoistChPerCta + warp * 2;
  float w0[kHoistMaxC / 32], w1[kHoistMaxC / 32];
#pragma unroll
  for (int i = 0; i < kHoistMaxC / 32; ++i) {
    const int k = lane + 32 * i;
    w0[i] = (k < C && j0 < D) ? W[(size_t)j0 * ldw + 1 + k] : 0.f;          // column 0 multiplies t
    w1[i] = (k < C && j0 + 1 < D) ? W[(size_t)(j0 + 1) * ldw + 1 + k] : 0.f;
  }
  const float b0 = (bias && j0 < D) ? bias[j0] : 0.f;
  const float b1 = (bias && j0 + 1 < D) ? bias[j0 + 1] : 0.f;
  {
    const int f0 = blockIdx.y * kHoistFrames;                    // one tile of frames per CTA
    const int nf = min(kHoistFrames, frames - f0);
    for (int i = threadIdx.x; i < nf * C; i += 256) s_ctx[i] = ctx[(size_t)f0 * C + i];
    __syncthreads();
    for (int f = 0; f < nf; ++f) {
      const float* c = s_ctx + f * C;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < kHoistMaxC / 32; ++i) {
        const int k = lane + 32 * i;
        if (k < C) {
          const float cv = c[k];
          a0 = fmaf(w0[i], cv, a0);
          a1 = fmaf(w1[i], cv, a1);
        }
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane == 0) {
        if (j0 < D) out[(size_t)(f0 + f) * ld_out + j0] = a0 + b0;
        if (j0 + 1 < D) out[(size_t)(f0 + f) * ld_out + j0 + 1] = a1 + b1;
      }
    }
  }
}

__global__ void gather_col0_kernel(const float* __restrict__ W, int ldw, int rows, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) out[i] = W[(size_t)i * ldw];
}
__global__ void copy_f32_kernel(const float* __restrict__ src, int nelem, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nelem) dst[i] = src[i];
}

// ----------------------------------------------------------------------------- workspace
// leading dimension of the per-frame hyper arrays: 3H+3 channels padded to a float4 multiple
inline int hyper_ld(int H) { return (3 * H + 3 + 3) / 4 * 4; }

struct CnfWorkspace {
  CnfState* st;
  float *Gc, *Bc, *gate, *biasf;       // frames x ctot
  float *wg_t, *wb_t, *lbias;          // ctot
  float4 *y0, *y1, *kbuf;              // n, n, 7n
  float *Ha, *Va, *Hb, *Vb;            // n_pad x H each (the tensor-core engine reuses them as fp16 planes)
  float* col_scale;                    // ctot (tensor-core engine)
  float* acc6;                         // n x 8 partial sums of the fused output layer (tensor-core engine)
  int* range_flag;
  cnf_tc::Weights tcw;
  void* fused_scratch;                 // per-CTA L2-resident planes of the fused evaluation kernel
  int fused_grid;
  size_t bytes;
};

CnfWorkspace carve(void* base, int frames, int pts, int H) {
  CnfWorkspace w;
  const size_t n = (size_t)frames * pts;
  const size_t n_pad = (n + 63) / 64 * 64;
  const size_t ctot = hyper_ld(H);
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
  w.st = (CnfState*)take(sizeof(CnfState));
  w.Gc = (float*)take(frames * ctot * 4);
  w.Bc = (float*)take(frames * ctot * 4);
  w.gate = (float*)take(7 * frames * ctot * 4);        // one slot per RK stage (0 = f at the step start)
  w.biasf = (float*)take(7 * frames * ctot * 4);
  w.wg_t = (float*)take(ctot * 4);
  w.wb_t = (float*)take(ctot * 4);
  w.lbias = (float*)take(ctot * 4);
  w.y0 = (float4*)take(n * 16);
  w.y1 = (float4*)take(n * 16);
  w.kbuf = (float4*)take(7 * n * 16);
  w.Ha = (float*)take(n_pad * H * 4);
  w.Va = (float*)take(n_pad * H * 4);
  w.Hb = (float*)take(n_pad * H * 4);
  w.Vb = (float*)take(n_pad * H * 4);
  w.col_scale = (float*)take(ctot * 4);
  w.acc6 = (float*)take(n * 8 * 4);
  w.range_flag = (int*)take(256);
  for (int l = 0; l < 2; ++l) {
    w.tcw.hi[l] = (__half*)take((size_t)512 * 512 * 2);
    w.tcw.lo[l] = (__half*)take((size_t)512 * 512 * 2);
  }
  w.tcw.scales = (float*)take(256);
  w.tcw.max_bits = (unsigned*)take(256);
  // sized for the largest device this library targets (148 SMs); a smaller point set needs fewer CTAs
  w.fused_grid = cnf_tc::fused_grid_for((int)n, 148);
  w.fused_scratch = take(cnf_tc::fused_scratch_bytes(w.fused_grid));
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

MbnDev load_mbn(const caspr_mbn_params* m, const float* h) {
  MbnDev d;
  d.present = m ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    d.w[c] = m ? h[c] : 0.f;
    d.b[c] = m ? h[3 + c] : 0.f;
    d.mean[c] = m ? h[6 + c] : 0.f;
    d.var[c] = m ? h[9 + c] : 1.f;
  }
  return d;
}

int blocks_for(long long work, int per_block, int cap) {
  long long b = (work + per_block - 1) / per_block;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// Hoist the context part of the hyper-linears: Gc = Wg[:,1:].c + bg ; Bc = Wb[:,1:].c
int prepare_hyper(const CnfWorkspace& w, const caspr_cnf_weights* cw, const float* ctx, int frames,
                  cudaStream_t s) {
  const int H = cw->hidden, C = cw->ctx_dim;
  const int ctot = hyper_ld(H);
  int off = 0;
  for (int l = 0; l < 4; ++l) {
    const int D = l < 3 ? H : 3;
    if (C <= kHoistMaxC) {
      const size_t smem = (size_t)kHoistFrames * C * sizeof(float);
      static bool attr_set = false;
      if (!attr_set) {
        if (cudaFuncSetAttribute(cnf_hyper_hoist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kHoistFrames * kHoistMaxC * (int)sizeof(float)) != cudaSuccess)
          return CASPR_ELAUNCH;
        attr_set = true;
      }
      const dim3 grid(ceil_div(D, kHoistChPerCta), ceil_div(frames, kHoistFrames));
      CASPR_COUNT(); cnf_hyper_hoist_kernel<<<grid, 256, smem, s>>>(ctx, C, cw->Wgate[l], C + 1, cw->bgate[l], frames, D,
                                                                   w.Gc + off, ctot);
      CASPR_COUNT(); cnf_hyper_hoist_kernel<<<grid, 256, smem, s>>>(ctx, C, cw->Wbias[l], C + 1, nullptr, frames, D,
                                                                   w.Bc + off, ctot);
    } else {
      int rc = caspr_linear(ctx, C, cw->Wgate[l] + 1, C + 1, cw->bgate[l], w.Gc + off, ctot, frames, C, D,
                            CASPR_ACT_NONE, CASPR_ACT_NONE, s);
      if (rc) return rc;
      rc = caspr_linear(ctx, C, cw->Wbias[l] + 1, C + 1, nullptr, w.Bc + off, ctot, frames, C, D,
                        CASPR_ACT_NONE, CASPR_ACT_NONE, s);
      if (rc) return rc;
    }
    CASPR_COUNT(); gather_col0_kernel<<<ceil_div(D, 128), 128, 0, s>>>(cw->Wgate[l], C + 1, D, w.wg_t + off);
    CASPR_COUNT(); gather_col0_kernel<<<ceil_div(D, 128), 128, 0, s>>>(cw->Wbias[l], C + 1, D, w.wb_t + off);
    CASPR_COUNT(); copy_f32_kernel<<<ceil_div(D, 128), 128, 0, s>>>(cw->b[l], D, w.lbias + off);
    CASPR_CHECK_LAUNCH();
    off += D;
  }
  return CASPR_OK;
}

// Side stream used to pipeline the two halves of the point set (tensor-core engine): while one half runs its
// MMA-bound GEMMs, the other half's HBM-bound layer-0 kernel shares the SMs.  Created once per process.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
SideStream& side_stream() {
  static SideStream ss;
  if (!ss.ok) {
    ss.ok = cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) == cudaSuccess;
  }
  return ss;
}

// tensor-core engine: layers 0..3 of one dynamics evaluation for the points [pt0, pt1) into kbuf[stage]
int enqueue_feval_tc_range(const CnfWorkspace& w, const caspr_cnf_weights* cw, const float* e, int frames, int pts,
                           int stage, int reverse, const cnf_tc::Plan* plan, int num_sms, int pt0, int pt1,
                           cudaStream_t s) {
  const int H = cw->hidden;
  const int n = frames * pts;
  const int ctot = hyper_ld(H);
  const float* gate = w.gate + (size_t)stage * frames * ctot;
  const float* biasf = w.biasf + (size_t)stage * frames * ctot;
  const int m0 = pt0 / 64, m1 = (pt1 + 63) / 64;
  int rc = cnf_tc::enqueue_layer0(*plan, w.y0, w.kbuf, (size_t)n, e, cw->W[0], pt0, pt1, pts, stage, gate, biasf,
                                  ctot, w.st, w.range_flag, s);
  if (rc) return rc;
  rc = cnf_tc::enqueue_mid(*plan, 0, m0, m1 - m0, gate + H, biasf + H, ctot, pt1, pts, w.st, nullptr, nullptr,
                           w.range_flag, num_sms, s);
  if (rc) return rc;
  rc = cnf_tc::enqueue_mid(*plan, 1, m0, m1 - m0, gate + 2 * H, biasf + 2 * H, ctot, pt1, pts, w.st, cw->W[3], w.acc6,
                           w.range_flag, num_sms, s);
  if (rc) return rc;
  return cnf_tc::enqueue_last_finish(w.acc6, e, pt0, pt1, pts, gate + 3 * H, biasf + 3 * H, ctot, reverse, w.st,
                                     w.kbuf + (size_t)stage * n, s);
}

// Dynamics evaluations for the RK stages [stage_first, stage_last] into kbuf[stage] (stage 0 = f at the step
// start / f0).  Stage inputs depend only on the same point's earlier stages, so on the tensor-core engine the two
// halves of the point set run all stages independently on two streams and meet again before the error norm.
int enqueue_stages(const CnfWorkspace& w, const caspr_cnf_weights* cw, const float* e, int frames, int pts,
                   int stage_first, int stage_last, int reverse, int engine, const cnf_tc::Plan* plan, int num_sms,
                   cudaStream_t s) {
  const int H = cw->hidden;
  const int n = frames * pts;
  const int ctot = hyper_ld(H);
  const long long tot = (long long)frames * ctot;
  const bool use_tc = engine == CASPR_CNF_TC_FP16X3;
  const dim3 hgrid((unsigned)((tot + 255) / 256), (unsigned)(stage_last - stage_first + 1));
  CASPR_COUNT(); cnf_hyper_stage_kernel<<<hgrid, 256, 0, s>>>(
      w.Gc, w.Bc, w.wg_t, w.wb_t, w.lbias, frames, 3 * H + 3, ctot, stage_first, reverse, w.st,
      use_tc ? w.col_scale : nullptr, w.gate, w.biasf);
  CASPR_CHECK_LAUNCH();
  if (use_tc && plan->fused_grid > 0) {
    // ONE launch per dynamics evaluation: layer 0, both H x H layers, the output layer and the divergence
    for (int stage = stage_first; stage <= stage_last; ++stage) {
      const float* gate = w.gate + (size_t)stage * frames * ctot;
      const float* biasf = w.biasf + (size_t)stage * frames * ctot;
      int rc = cnf_tc::enqueue_fused(*plan, w.y0, w.kbuf, (size_t)n, e, cw->W[0], cw->W[3], n, pts, stage, reverse, gate,
                                     biasf, ctot, w.st, w.kbuf + (size_t)stage * n, w.range_flag, s);
      if (rc) return rc;
    }
    return CASPR_OK;
  }
  if (use_tc) {
    const int n_tiles = (n + 63) / 64;
    // Optional (CASPR_CNF_PIPELINE_HALVES=1): measured on B200 at 163 840 points the decode span drops from 41.1 to
    // 38.8 ms, but kernels of the two streams then queue behind each other, which makes the per-kernel CUDA-event
    // durations (bench.py's roofline) meaningless; off by default.
    const char* env = getenv("CASPR_CNF_PIPELINE_HALVES");
    const bool want_split = env && env[0] == '1';
    SideStream dummy;
    SideStream& ss = want_split ? side_stream() : dummy;
    const bool split = want_split && ss.ok && n_tiles >= 2 * num_sms;       // each half keeps every SM busy
    const int pt_split = split ? (n_tiles / 2) * 64 : n;
    if (split) {
      if (cudaEventRecord(ss.fork, s) != cudaSuccess || cudaStreamWaitEvent(ss.stream, ss.fork, 0) != cudaSuccess)
        return CASPR_ELAUNCH;
    }
    for (int stage = stage_first; stage <= stage_last; ++stage) {
      int rc = enqueue_feval_tc_range(w, cw, e, frames, pts, stage, reverse, plan, num_sms, 0, pt_split, s);
      if (rc) return rc;
      if (split) {
        rc = enqueue_feval_tc_range(w, cw, e, frames, pts, stage, reverse, plan, num_sms, pt_split, n, ss.stream);
        if (rc) return rc;
      }
    }
    if (split) {
      if (cudaEventRecord(ss.join, ss.stream) != cudaSuccess || cudaStreamWaitEvent(s, ss.join, 0) != cudaSuccess)
        return CASPR_ELAUNCH;
    }
    return CASPR_OK;
  }
  for (int stage = stage_first; stage <= stage_last; ++stage) {
    const float* gate = w.gate + (size_t)stage * frames * ctot;
    const float* biasf = w.biasf + (size_t)stage * frames * ctot;
    CASPR_COUNT(); cnf_layer0_kernel<<<blocks_for(n, 8, 148 * 16), 256, 0, s>>>(
        w.y0, w.kbuf, (size_t)n, e, cw->W[0], H, n, pts, stage, gate, biasf, ctot, w.st, w.Ha, w.Va);
    dim3 grid(ceil_div(n, kMidBM), H / kMidBN);
    caspr_prof_begin(CASPR_PROF_CNF_MID_SIMT, s);
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidForward><<<grid, 256, 0, s>>>(w.Ha, w.Va, cw->W[1], H, n, pts, gate + H, biasf + H,
                                              ctot, w.st, w.Hb, w.Vb, nullptr, nullptr);
    caspr_prof_end(CASPR_PROF_CNF_MID_SIMT, s);
    caspr_prof_begin(CASPR_PROF_CNF_MID_SIMT, s);
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidForward><<<grid, 256, 0, s>>>(w.Hb, w.Vb, cw->W[2], H, n, pts, gate + 2 * H,
                                              biasf + 2 * H, ctot, w.st, w.Ha, w.Va, nullptr, nullptr);
    caspr_prof_end(CASPR_PROF_CNF_MID_SIMT, s);
    CASPR_COUNT(); cnf_last_layer_kernel<<<blocks_for(n, 8, 148 * 16), 256, 0, s>>>(
        w.Ha, w.Va, cw->W[3], H, n, pts, e, gate + 3 * H, biasf + 3 * H, ctot, reverse, w.st,
        w.kbuf + (size_t)stage * n);
    CASPR_CHECK_LAUNCH();
  }
  return CASPR_OK;
}

// Engine set-up shared by caspr_cnf_flow / caspr_cnf_feval: weight split + tensor maps.
int prepare_engine(const CnfWorkspace& w, const caspr_cnf_weights* cw, int n, int engine, cnf_tc::Plan* plan,
                   int* num_sms, cudaStream_t s) {
  *num_sms = 148;
  if (engine != CASPR_CNF_TC_FP16X3) return CASPR_OK;
  if (cw->hidden != 512) return CASPR_EINVAL;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (cudaMemsetAsync(w.range_flag, 0, sizeof(int), s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaMemsetAsync(w.acc6, 0, (size_t)n * 8 * sizeof(float), s) != cudaSuccess) return CASPR_ELAUNCH;
  int rc = cnf_tc::prepare_weights(cw->W[1], cw->W[2], w.tcw, s);
  if (rc) return rc;
  rc = cnf_tc::fill_col_scale(w.tcw, hyper_ld(cw->hidden), w.col_scale, s);
  if (rc) return rc;
  int fgrid = cnf_tc::fused_grid_for(n, *num_sms);
  if (fgrid > w.fused_grid) fgrid = w.fused_grid;
  {
    // Experiment (CASPR_CNF_L2_PERSIST=1): mark the per-CTA scratch of the fused kernel as persisting in L2 so that its
    // dirty lines are not written back to HBM between the tile that writes them and the tile that overwrites them.
    static int persist = -1;
    if (persist < 0) { const char* e2 = getenv("CASPR_CNF_L2_PERSIST"); persist = (e2 && e2[0] == '1') ? 1 : 0; }
    if (persist && cnf_tc::fused_enabled()) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
        size_t bytes = cnf_tc::fused_scratch_bytes(fgrid);
        size_t set_aside = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
        static size_t limit_set = 0;                 // the set-aside is configured once per process (it is a device-wide
        if (set_aside > limit_set) {                 // reconfiguration); lines only persist while a window is active
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
          limit_set = set_aside;
        }
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = w.fused_scratch;
        attr.accessPolicyWindow.num_bytes = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
        static bool printed = false;
        if (!printed) {
          fprintf(stderr, "[caspr] L2 persistence: set-aside %zu MB (max %d MB), window %zu MB (max %d MB)\n", set_aside >> 20,
                  prop.persistingL2CacheMaxSize >> 20, (size_t)attr.accessPolicyWindow.num_bytes >> 20,
                  prop.accessPolicyMaxWindowSize >> 20);
          printed = true;
        }
      }
    }
  }
  return cnf_tc::make_plan(*plan, w.tcw, (__half*)w.Ha, (__half*)w.Va, (__half*)w.Hb, (__half*)w.Vb, n,
                           cnf_tc::fused_enabled() ? w.fused_scratch : nullptr, fgrid);
}

// end of a solve: close the access-policy window and hand the persisting lines back to normal use
void release_l2_persistence(cudaStream_t s) {
  const char* e2 = getenv("CASPR_CNF_L2_PERSIST");
  if (!(e2 && e2[0] == '1') || !cnf_tc::fused_enabled()) return;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  attr.accessPolicyWindow.num_bytes = 0;
  attr.accessPolicyWindow.hitRatio = 0.f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaCtxResetPersistingL2Cache();
}

bool weights_ok(const caspr_cnf_weights* cw) {
  if (!cw) return false;
  for (int l = 0; l < 4; ++l)
    if (!cw->W[l] || !cw->b[l] || !cw->Wgate[l] || !cw->bgate[l] || !cw->Wbias[l]) return false;
  return cw->hidden > 0 && cw->hidden <= kMaxHidden && cw->hidden % kMidBN == 0 && cw->ctx_dim > 0;
}

}  // namespace
