// Adjoint backward of the conditional CNF block (training, BASELINE config 5).
//
// Replaces torchdiffeq 0.0.1's OdeintAdjointMethod.backward for the call at caspr/models/cnf.py:102-111
// (restated in oracle/odeint001.py::_AdjointMethod) together with the autograd VJP it takes through
// ODEfunc.forward / divergence_approx / ODEnet / ConcatSquashLinear (odefunc.py:13-31,98-105,119-142,
// diffeq_layers.py:83-90).
//
// The augmented state (x, logp, ctx, adj_x, adj_logp, adj_ctx, adj_t, adj_params) is integrated from
// t1 = sqrt_end_time^2 back to 0 with dopri5.  As in the reference, the tolerance lists have three entries, so
// only (x, logp, ctx) take part in step control; the adjoint tensors ride along, but all eight tensors enter
// the initial-step heuristic.  One augmented evaluation =
//   forward  : hyper gates -> layer 0 -> two H x H layers (raw products kept) -> output layer (dy, div)
//   backward : reverse-mode sweep over the (activation, tangent) pair of every layer with cotangents
//              u3 = adj_x, v3 = -adj_logp * e:  two H x H data-gradient products, two H x H weight-gradient
//              products (split over the points, reduced deterministically), per-frame gate / bias cotangents,
//              then the hyper-network gradients (weights, context, time).
// Engines: CASPR_CNF_SIMT_FP32 runs everything as exact fp32 SIMT; CASPR_CNF_TC_FP16X3 sends the forward and the
// data-gradient H x H products through caspr_linear_tc and the weight gradients through the split-K
// caspr_linear_wgrad_tc (all six products of an evaluation on the tcgen05 fp16x3 GEMM of gemm_tc.cu).
#include "cnf_kernels.cuh"
#include "rk_flat.cuh"

namespace {

// ------------------------------------------------------------------------------ parameter layout
// Flat order = ODEfunc.parameters(): per layer _layer.weight, _layer.bias, _hyper_bias.weight,
// _hyper_gate.weight, _hyper_gate.bias (diffeq_layers.py:79-81).
struct ParamLayout {
  size_t W[4], b[4], Wb[4], Wg[4], bg[4], total;
};
ParamLayout param_layout(int H, int C) {
  ParamLayout L;
  size_t off = 0;
  for (int l = 0; l < 4; ++l) {
    const size_t D = l < 3 ? H : 3, Din = l == 0 ? 3 : H;
    L.W[l] = off;  off += D * Din;
    L.b[l] = off;  off += D;
    L.Wb[l] = off; off += D * (C + 1);
    L.Wg[l] = off; off += D * (C + 1);
    L.bg[l] = off; off += D;
  }
  L.total = off;
  return L;
}
struct ParamOffsets {      // by-value kernel argument
  unsigned W[4], b[4], Wb[4], Wg[4], bg[4];
};

// softplus' and softplus'' (beta 1, threshold 20: torch's softplus_backward / softplus_double_backward)
__device__ __forceinline__ void softplus_d12(float x, float& d1, float& d2) {
  if (x > 20.f) {
    d1 = 1.f;
    d2 = 0.f;
  } else {
    const float z = expf(x);
    d1 = __fdiv_rn(z, __fadd_rn(z, 1.f));
    d2 = d1 * (1.f - d1);
  }
}

// Stage input of a float4 state: y0 + sum_j (dt*beta[s][j]) k_j for all four components.
__device__ __forceinline__ float4 stage_input4(const float4* __restrict__ y0, const float4* __restrict__ kbuf,
                                               size_t kstride, int pt, int stage, float dt) {
  const float4 y = y0[pt];
  if (stage == 0) return y;
  float kx[6], ky[6], kz[6], kw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (j < stage) {
      const float4 kv = kbuf[(size_t)j * kstride + pt];
      kx[j] = kv.x; ky[j] = kv.y; kz[j] = kv.z; kw[j] = kv.w;
    } else {
      kx[j] = ky[j] = kz[j] = kw[j] = 0.f;
    }
  }
  return make_float4(dopri5::stage_combine(y.x, dt, kx, stage - 1), dopri5::stage_combine(y.y, dt, ky, stage - 1),
                     dopri5::stage_combine(y.z, dt, kz, stage - 1), dopri5::stage_combine(y.w, dt, kw, stage - 1));
}

// From the cotangents (hb, hdb) of a layer's outputs h = softplus(u), hd = softplus'(u) * v with
// u = a*g + bf, v = g*ad:  ubar, vbar and the contributions to the gate / bias cotangents.
__device__ __forceinline__ void through_activation(float hb, float hdb, float a, float ad, float g, float bf,
                                                   float blayer, float& ubar, float& vbar, float& gsum,
                                                   float& bsum) {
  const float u = fmaf(a, g, bf);
  const float v = g * ad;
  float d1, d2;
  softplus_d12(u, d1, d2);
  ubar = hb * d1 + hdb * d2 * v;
  vbar = hdb * d1;
  gsum += ubar * (a + blayer) + vbar * ad;
  bsum += ubar;
}

// Tensor-core engine: the H x H products come back raw from caspr_linear_tc; this applies the ConcatSquash gate /
// bias, softplus and the tangent's chain rule (the epilogue of cnf_mid_layer_kernel<kMidForward>).
__global__ void __launch_bounds__(256)
adj_act_kernel(const float* __restrict__ A, const float* __restrict__ Ad, const float* __restrict__ gate,
               const float* __restrict__ biasf, int ld, int H, int P, long long total4,
               const CnfState* __restrict__ st, float* __restrict__ Hn, float* __restrict__ Vn,
               unsigned* __restrict__ colmax) {
  if (st->done) return;
  const int h4 = H / 4;
  // the grid stride (gridDim.x * 256) is a multiple of H/4, so a thread always works on the same four channels
  float cmx[4] = {0.f, 0.f, 0.f, 0.f};
  int cch = -1;
  // total4 = n * H/4 < 2^31 (n < 2^30 / ... checked by the caller): 32-bit index arithmetic
  const unsigned tot = (unsigned)total4, stride = gridDim.x * blockDim.x, h4u = (unsigned)h4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
    const unsigned ptu = i / h4u;
    const long long pt = ptu;
    const int c = (int)(i - ptu * h4u) * 4;
    const int f = (int)(ptu / (unsigned)P);
    const float4 a4 = *reinterpret_cast<const float4*>(A + pt * H + c);
    const float4 d4 = *reinterpret_cast<const float4*>(Ad + pt * H + c);
    const float4 g4 = *reinterpret_cast<const float4*>(gate + (size_t)f * ld + c);
    const float4 b4 = *reinterpret_cast<const float4*>(biasf + (size_t)f * ld + c);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, d[4] = {d4.x, d4.y, d4.z, d4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    float ho[4], vo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sp, dsp;
      softplus_and_grad(fmaf(a[j], g[j], b[j]), sp, dsp);
      ho[j] = sp;
      vo[j] = dsp * g[j] * d[j];
      cmx[j] = fmaxf(cmx[j], fmaxf(fabsf(ho[j]), fabsf(vo[j])));
    }
    cch = c;
    *reinterpret_cast<float4*>(Hn + pt * H + c) = make_float4(ho[0], ho[1], ho[2], ho[3]);
    *reinterpret_cast<float4*>(Vn + pt * H + c) = make_float4(vo[0], vo[1], vo[2], vo[3]);
  }
  if (colmax && cch >= 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      atomic_max_nonneg(colmax + cch + j, cmx[j]);
  }
}

constexpr int kChunkMax = 64;       // points per chunk of the thread-per-channel kernels

// Output layer backward + conversion to the cotangents of layer 2's pre-activations.
// grid (chunks, frames), block H threads (thread j = hidden channel j).
__global__ void __launch_bounds__(512)
adj_bwd_last_kernel(const float4* __restrict__ adj0, const float4* __restrict__ kadj, size_t kstride, int stage,
                    const CnfState* __restrict__ st, const float* __restrict__ e, const float* __restrict__ W3,
                    const float* __restrict__ raw8, const float* __restrict__ A2, const float* __restrict__ Ad2,
                    const float* __restrict__ H3, const float* __restrict__ V3, const float* __restrict__ gate,
                    const float* __restrict__ biasf, const float* __restrict__ lbias, int ld, int H, int P, int L,
                    float* __restrict__ Ab, float* __restrict__ Av, float* __restrict__ gpart,
                    float* __restrict__ bpart, float* __restrict__ w3part, unsigned* __restrict__ colmax) {
  if (st->done) return;
  __shared__ float s_ab[kChunkMax][8];      // g3*ubar3 [0..2], g3*vbar3 [4..6]
  __shared__ float s_g3[kChunkMax][8];      // gate cotangent terms [0..2], ubar3 [4..6]
  const int f = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int q0 = chunk * L;
  const int npts = min(L, P - q0);
  const int j = threadIdx.x;
  const float dt = (float)st->dt;
  const float* g = gate + (size_t)f * ld;
  const float* bf = biasf + (size_t)f * ld;
  if (j < npts) {
    const int pt = f * P + q0 + j;
    const float4 a = stage_input4(adj0, kadj, kstride, pt, stage, dt);
    const float ub[3] = {a.x, a.y, a.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float vb = -a.w * e[3 * (size_t)pt + c];
      const float g3 = g[3 * H + c];
      s_ab[j][c] = g3 * ub[c];
      s_ab[j][4 + c] = g3 * vb;
      s_g3[j][c] = ub[c] * (raw8[8 * (size_t)pt + c] + lbias[3 * H + c]) + vb * raw8[8 * (size_t)pt + 4 + c];
      s_g3[j][4 + c] = ub[c];
    }
  }
  __syncthreads();
  const float w0 = W3[j], w1 = W3[H + j], w2 = W3[2 * H + j];
  const float gj = g[2 * H + j], bfj = bf[2 * H + j], bl = lbias[2 * H + j];
  float gsum = 0.f, bsum = 0.f, wa0 = 0.f, wa1 = 0.f, wa2 = 0.f, cmx = 0.f;
  for (int q = 0; q < npts; ++q) {
    const size_t row = (size_t)(f * P + q0 + q) * H + j;
    const float ab0 = s_ab[q][0], ab1 = s_ab[q][1], ab2 = s_ab[q][2];
    const float av0 = s_ab[q][4], av1 = s_ab[q][5], av2 = s_ab[q][6];
    const float hb = fmaf(w2, ab2, fmaf(w1, ab1, w0 * ab0));
    const float hdb = fmaf(w2, av2, fmaf(w1, av1, w0 * av0));
    float ubar, vbar;
    through_activation(hb, hdb, A2[row], Ad2[row], gj, bfj, bl, ubar, vbar, gsum, bsum);
    Ab[row] = gj * ubar;
    Av[row] = gj * vbar;
    cmx = fmaxf(cmx, fmaxf(fabsf(gj * ubar), fabsf(gj * vbar)));
    const float h3 = H3[row], v3 = V3[row];
    wa0 = fmaf(ab0, h3, fmaf(av0, v3, wa0));
    wa1 = fmaf(ab1, h3, fmaf(av1, v3, wa1));
    wa2 = fmaf(ab2, h3, fmaf(av2, v3, wa2));
  }
  const size_t part = (size_t)f * nchunk + chunk;
  gpart[part * ld + 2 * H + j] = gsum;
  bpart[part * ld + 2 * H + j] = bsum;
  w3part[part * 3 * H + j] = wa0;
  w3part[part * 3 * H + H + j] = wa1;
  w3part[part * 3 * H + 2 * H + j] = wa2;
  if (colmax) atomic_max_nonneg(colmax + j, cmx);
  if (j < 3) {
    float sg = 0.f, sb = 0.f;
    for (int q = 0; q < npts; ++q) { sg += s_g3[q][j]; sb += s_g3[q][4 + j]; }
    gpart[part * ld + 3 * H + j] = sg;
    bpart[part * ld + 3 * H + j] = sb;
  }
}

// (Gh, Gv) = cotangents of layer (lp+1)'s inputs h, hd  ->  (Ab, Av) = g*ubar, g*vbar of layer lp (lp = 1).
__global__ void __launch_bounds__(512)
adj_bwd_mid_kernel(const float* __restrict__ Gh, const float* __restrict__ Gv, const float* __restrict__ A,
                   const float* __restrict__ Ad, const float* __restrict__ gate, const float* __restrict__ biasf,
                   const float* __restrict__ lbias, int ld, int H, int P, int L, int lp,
                   const CnfState* __restrict__ st, float* __restrict__ Ab, float* __restrict__ Av,
                   float* __restrict__ gpart, float* __restrict__ bpart, unsigned* __restrict__ colmax) {
  if (st->done) return;
  const int f = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int q0 = chunk * L;
  const int npts = min(L, P - q0);
  const int j = threadIdx.x;
  const float gj = gate[(size_t)f * ld + lp * H + j], bfj = biasf[(size_t)f * ld + lp * H + j];
  const float bl = lbias[lp * H + j];
  float gsum = 0.f, bsum = 0.f, cmx = 0.f;
  for (int q = 0; q < npts; ++q) {
    const size_t row = (size_t)(f * P + q0 + q) * H + j;
    float ubar, vbar;
    through_activation(Gh[row], Gv[row], A[row], Ad[row], gj, bfj, bl, ubar, vbar, gsum, bsum);
    Ab[row] = gj * ubar;
    Av[row] = gj * vbar;
    cmx = fmaxf(cmx, fmaxf(fabsf(gj * ubar), fabsf(gj * vbar)));
  }
  const size_t part = (size_t)f * nchunk + chunk;
  gpart[part * ld + lp * H + j] = gsum;
  bpart[part * ld + lp * H + j] = bsum;
  if (colmax) atomic_max_nonneg(colmax + j, cmx);
}

// Layer 0 backward: (Gh, Gv) = cotangents of h1, hd1 -> gradient wrt the stage input (k of adj_x), W0 gradient
// partials, gate / bias cotangents of layer 0.
__global__ void __launch_bounds__(512)
adj_bwd_layer0_kernel(const float4* __restrict__ y0, const float4* __restrict__ kbuf, size_t kstride, int stage,
                      const CnfState* __restrict__ st, const float* __restrict__ e, const float* __restrict__ W0,
                      const float* __restrict__ Gh, const float* __restrict__ Gv, const float* __restrict__ gate,
                      const float* __restrict__ biasf, const float* __restrict__ lbias, int ld, int H, int P, int L,
                      float4* __restrict__ kadj_out, float* __restrict__ gpart, float* __restrict__ bpart,
                      float* __restrict__ w0part) {
  if (st->done) return;
  __shared__ float s_ys[kChunkMax][4];
  __shared__ float s_e[kChunkMax][4];
  __shared__ float s_vx[16][kChunkMax][3];
  const int f = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int q0 = chunk * L;
  const int npts = min(L, P - q0);
  const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
  const float dt = (float)st->dt;
  if (j < npts) {
    const int pt = f * P + q0 + j;
    const float4 ys = stage_input4(y0, kbuf, kstride, pt, stage, dt);
    s_ys[j][0] = ys.x; s_ys[j][1] = ys.y; s_ys[j][2] = ys.z;
    s_e[j][0] = e[3 * (size_t)pt]; s_e[j][1] = e[3 * (size_t)pt + 1]; s_e[j][2] = e[3 * (size_t)pt + 2];
  }
  __syncthreads();
  const float w0 = W0[3 * j], w1 = W0[3 * j + 1], w2 = W0[3 * j + 2];
  const float gj = gate[(size_t)f * ld + j], bfj = biasf[(size_t)f * ld + j], bl = lbias[j];
  float gsum = 0.f, bsum = 0.f;
  float wacc[3] = {0.f, 0.f, 0.f};
  for (int q = 0; q < npts; ++q) {
    const size_t row = (size_t)(f * P + q0 + q) * H + j;
    const float y_0 = s_ys[q][0], y_1 = s_ys[q][1], y_2 = s_ys[q][2];
    const float e_0 = s_e[q][0], e_1 = s_e[q][1], e_2 = s_e[q][2];
    const float a = fmaf(w2, y_2, fmaf(w1, y_1, w0 * y_0));
    const float ad = fmaf(w2, e_2, fmaf(w1, e_1, w0 * e_0));
    float ubar, vbar;
    through_activation(Gh[row], Gv[row], a, ad, gj, bfj, bl, ubar, vbar, gsum, bsum);
    const float ab = gj * ubar, av = gj * vbar;
    wacc[0] = fmaf(ab, y_0, fmaf(av, e_0, wacc[0]));
    wacc[1] = fmaf(ab, y_1, fmaf(av, e_1, wacc[1]));
    wacc[2] = fmaf(ab, y_2, fmaf(av, e_2, wacc[2]));
    const float vx = warp_sum(w0 * ab), vy = warp_sum(w1 * ab), vz = warp_sum(w2 * ab);
    if (lane == 0) { s_vx[warp][q][0] = vx; s_vx[warp][q][1] = vy; s_vx[warp][q][2] = vz; }
  }
  const size_t part = (size_t)f * nchunk + chunk;
  gpart[part * ld + j] = gsum;
  bpart[part * ld + j] = bsum;
  w0part[part * 3 * H + 3 * j] = wacc[0];
  w0part[part * 3 * H + 3 * j + 1] = wacc[1];
  w0part[part * 3 * H + 3 * j + 2] = wacc[2];
  __syncthreads();
  if (j < npts) {
    const int nw = blockDim.x >> 5;
    float v[3] = {0.f, 0.f, 0.f};
    for (int w = 0; w < nw; ++w) { v[0] += s_vx[w][j][0]; v[1] += s_vx[w][j][1]; v[2] += s_vx[w][j][2]; }
    kadj_out[f * P + q0 + j] = make_float4(v[0], v[1], v[2], 0.f);        // d(adj_logp)/dt = 0
  }
}

// Weight gradient of an H x H layer: C[o][i] = sum_p Ab[p][o] Hin[p][i] + Av[p][o] Vin[p][i] over one split of
// the points.  CTA tile 128 x 128, k-slab = 8 points x {activation, tangent}; 256 threads, 8 x 8 per thread.
constexpr int kWgTile = 128, kWgPts = 8;
__global__ void __launch_bounds__(256, 2)
adj_wgrad_kernel(const float* __restrict__ Ab, const float* __restrict__ Av, const float* __restrict__ Hin,
                 const float* __restrict__ Vin, int H, int n, int pts_per_split, const CnfState* __restrict__ st,
                 float* __restrict__ part) {
  if (st->done) return;
  __shared__ __align__(16) float As[2][2 * kWgPts][kWgTile];
  __shared__ __align__(16) float Bs[2][2 * kWgPts][kWgTile];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int o0 = blockIdx.x * kWgTile, i0 = blockIdx.y * kWgTile;
  const int p_begin = blockIdx.z * pts_per_split;
  const int p_end = min(n, p_begin + pts_per_split);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) acc[i][jn] = 0.f;
  const int lr = tid >> 5, lc = (tid & 31) * 4;          // loader: point lr of the slab, 4 channels at lc
  float4 ra0, ra1, rb0, rb1;
  auto load = [&](int p0) {
    const int p = p0 + lr;
    if (p < p_end) {
      ra0 = *reinterpret_cast<const float4*>(Ab + (size_t)p * H + o0 + lc);
      ra1 = *reinterpret_cast<const float4*>(Av + (size_t)p * H + o0 + lc);
      rb0 = *reinterpret_cast<const float4*>(Hin + (size_t)p * H + i0 + lc);
      rb1 = *reinterpret_cast<const float4*>(Vin + (size_t)p * H + i0 + lc);
    } else {
      ra0 = ra1 = rb0 = rb1 = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = ra0;
    *reinterpret_cast<float4*>(&As[buf][kWgPts + lr][lc]) = ra1;
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = rb0;
    *reinterpret_cast<float4*>(&Bs[buf][kWgPts + lr][lc]) = rb1;
  };
  const int nslab = (max(p_end - p_begin, 0) + kWgPts - 1) / kWgPts;
  if (nslab > 0) {
    load(p_begin);
    store(0);
  }
  __syncthreads();
  for (int sl = 0; sl < nslab; ++sl) {
    const int buf = sl & 1;
    if (sl + 1 < nslab) load(p_begin + (sl + 1) * kWgPts);
#pragma unroll
    for (int k = 0; k < 2 * kWgPts; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jn = 0; jn < 8; ++jn) acc[i][jn] = fmaf(a[i], b[jn], acc[i][jn]);
    }
    if (sl + 1 < nslab) {
      store(buf ^ 1);
      __syncthreads();
    }
  }
  float* dst = part + (size_t)blockIdx.z * H * H;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = o0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = i0 + half * 64 + tx * 4;
      *reinterpret_cast<float4*>(dst + (size_t)o * H + c) =
          make_float4(acc[i][half * 4], acc[i][half * 4 + 1], acc[i][half * 4 + 2], acc[i][half * 4 + 3]);
    }
  }
}

// dst[i] = sum_s part[s][i]
__global__ void __launch_bounds__(256)
adj_reduce_parts_kernel(const float* __restrict__ part, int nparts, size_t nelem, const CnfState* __restrict__ st,
                        float* __restrict__ dst) {
  if (st->done) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nelem) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * nelem + i];
  dst[i] = s;
}

// Per (frame, channel): sum the chunk partials, Ghat = gbar * g(1-g) (through the sigmoid), Bhat = bbar.
__global__ void __launch_bounds__(256)
adj_hyper_reduce_kernel(const float* __restrict__ gpart, const float* __restrict__ bpart, int nchunk, int frames,
                        int ctot, int ld, const float* __restrict__ gate, const CnfState* __restrict__ st,
                        float* __restrict__ Ghat, float* __restrict__ Bhat) {
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * ld) return;
  const int f = i / ld, j = i - f * ld;
  float gb = 0.f, bb = 0.f;
  if (j < ctot) {
    for (int c = 0; c < nchunk; ++c) {
      gb += gpart[((size_t)f * nchunk + c) * ld + j];
      bb += bpart[((size_t)f * nchunk + c) * ld + j];
    }
    const float g = gate[i];
    gb = gb * g * (1.f - g);
  }
  Ghat[i] = gb;
  Bhat[i] = bb;
}

// Per channel j of the concatenated layers: sums over frames -> layer bias, hyper-gate bias, the time columns of
// the two hyper weights; W0 / W3 partial reduction.
__global__ void __launch_bounds__(256)
adj_hyper_small_kernel(const float* __restrict__ Ghat, const float* __restrict__ Bhat, const float* __restrict__ gate,
                       int frames, int ctot, int ld, int H, int C, const CnfState* __restrict__ st, int stage,
                       ParamOffsets po, const float* __restrict__ w0part, const float* __restrict__ w3part,
                       int nparts, float* __restrict__ kpar) {
  if (st->done) return;
  const float t = cnf_stage_time(st, stage, 1);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ctot) {
    const int l = min(j / H, 3), jj = j - l * H;
    float sG = 0.f, sB = 0.f, sgb = 0.f;
    for (int f = 0; f < frames; ++f) {
      const float G = Ghat[(size_t)f * ld + j], B = Bhat[(size_t)f * ld + j];
      sG += G;
      sB += B;
      sgb = fmaf(gate[(size_t)f * ld + j], B, sgb);
    }
    kpar[po.Wg[l] + (size_t)jj * (C + 1)] = sG * t;
    kpar[po.bg[l] + jj] = sG;
    kpar[po.Wb[l] + (size_t)jj * (C + 1)] = sB * t;
    kpar[po.b[l] + jj] = sgb;
  }
  // W0 (H,3) and W3 (3,H): 3H elements each, reduced over the (frame, chunk) partials
  if (j < 3 * H) {
    float s0 = 0.f, s3 = 0.f;
#pragma unroll 8
    for (int p = 0; p < nparts; ++p) {
      s0 += w0part[(size_t)p * 3 * H + j];
      s3 += w3part[(size_t)p * 3 * H + j];
    }
    kpar[po.W[0] + j] = s0;
    kpar[po.W[3] + j] = s3;
  }
}

// Context gradient, one layer per blockIdx.z: cpart[l][f][k] = sum_jj Wg_l[jj][1+k] Ghat[f][off+jj] + Wb_l[jj][1+k] Bhat[f][off+jj]
constexpr int kCtxFrames = 8;
struct HyperPtrs {
  const float* Wg[4];
  const float* Wb[4];
};
__global__ void __launch_bounds__(128)
adj_ctx_bwd_kernel(HyperPtrs hp, const float* __restrict__ Ghat, const float* __restrict__ Bhat, int frames, int ld,
                   int H, int C, const CnfState* __restrict__ st, float* __restrict__ cpart) {
  if (st->done) return;
  __shared__ float sG[kCtxFrames][kMaxHidden];
  __shared__ float sB[kCtxFrames][kMaxHidden];
  const int l = blockIdx.z;
  const int D = l < 3 ? H : 3;
  const int off = l * H;
  const int f0 = blockIdx.y * kCtxFrames;
  const int nf = min(kCtxFrames, frames - f0);
  for (int i = threadIdx.x; i < kCtxFrames * D; i += blockDim.x) {
    const int ff = i / D, jj = i - ff * D;
    sG[ff][jj] = ff < nf ? Ghat[(size_t)(f0 + ff) * ld + off + jj] : 0.f;
    sB[ff][jj] = ff < nf ? Bhat[(size_t)(f0 + ff) * ld + off + jj] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= C) return;
  const float* wg = hp.Wg[l] + 1 + k;
  const float* wb = hp.Wb[l] + 1 + k;
  float acc[kCtxFrames];
#pragma unroll
  for (int ff = 0; ff < kCtxFrames; ++ff) acc[ff] = 0.f;
  // eight weight rows (2 x 8 independent loads) in flight per thread: the loop is latency-bound otherwise
  int jj = 0;
  for (; jj + 7 < D; jj += 8) {
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a[u] = __ldg(wg + (size_t)(jj + u) * (C + 1));
      b[u] = __ldg(wb + (size_t)(jj + u) * (C + 1));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int ff = 0; ff < kCtxFrames; ++ff)
        acc[ff] = fmaf(a[u], sG[ff][jj + u], fmaf(b[u], sB[ff][jj + u], acc[ff]));
  }
  for (; jj < D; ++jj) {
    const float a = wg[(size_t)jj * (C + 1)], b = wb[(size_t)jj * (C + 1)];
#pragma unroll
    for (int ff = 0; ff < kCtxFrames; ++ff) acc[ff] = fmaf(a, sG[ff][jj], fmaf(b, sB[ff][jj], acc[ff]));
  }
  for (int ff = 0; ff < nf; ++ff) cpart[((size_t)l * frames + f0 + ff) * C + k] = acc[ff];
}

// k[ctx slot] = sum of the four layer partials.  k[t slot] = 0: the reference detaches the time inside
// ODEfunc.forward (odefunc.py:121, `t.clone().detach()`), so the adjoint's vjp_t is identically zero and adj_time
// stays at -dL/dt1.
__global__ void __launch_bounds__(256)
adj_ctx_sum_kernel(const float* __restrict__ cpart, size_t nelem, const CnfState* __restrict__ st,
                   float* __restrict__ kctx, float* __restrict__ kt) {
  if (st->done) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nelem) kctx[i] = (cpart[i] + cpart[nelem + i]) + (cpart[2 * nelem + i] + cpart[3 * nelem + i]);
  if (i == 0) { kt[0] = 0.f; kt[1] = 0.f; kt[2] = 0.f; kt[3] = 0.f; }
}

// Hyper-weight gradients, context columns: Wg_l[jj][1+k] = sum_f Ghat[f][j] ctx[f][k]; Wb likewise with Bhat.
constexpr int kHwCh = 4;
__global__ void __launch_bounds__(256)
adj_hyper_wgrad_kernel(const float* __restrict__ Ghat, const float* __restrict__ Bhat, const float* __restrict__ ctx,
                       int frames, int ctot, int ld, int H, int C, const CnfState* __restrict__ st, ParamOffsets po,
                       float* __restrict__ kpar) {
  if (st->done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * kHwCh;
  if (k >= C) return;
  float ag[kHwCh], ab[kHwCh];
#pragma unroll
  for (int q = 0; q < kHwCh; ++q) { ag[q] = 0.f; ab[q] = 0.f; }
  for (int f = 0; f < frames; ++f) {
    const float c = ctx[(size_t)f * C + k];
#pragma unroll
    for (int q = 0; q < kHwCh; ++q) {
      const int j = min(j0 + q, ctot - 1);
      ag[q] = fmaf(Ghat[(size_t)f * ld + j], c, ag[q]);
      ab[q] = fmaf(Bhat[(size_t)f * ld + j], c, ab[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < kHwCh; ++q) {
    const int j = j0 + q;
    if (j >= ctot) break;
    const int l = min(j / H, 3), jj = j - l * H;
    kpar[po.Wg[l] + (size_t)jj * (C + 1) + 1 + k] = ag[q];
    kpar[po.Wb[l] + (size_t)jj * (C + 1) + 1 + k] = ab[q];
  }
}

// sum over a flat range of (k0 / (atol + |y0| rtol))^2 -> dst (double), for the initial-step heuristic
__global__ void __launch_bounds__(256)
adj_sumsq_kernel(const float* __restrict__ y0, const float* __restrict__ k0, size_t nelem, float rtol, float atol,
                 double* dst) {
  __shared__ double s_w[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += (size_t)gridDim.x * blockDim.x) {
    const float sc = __fadd_rn(atol, __fmul_rn(fabsf(y0[i]), rtol));
    const float r = __fdiv_rn(k0[i], sc);
    s += (double)r * (double)r;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    atomicAdd(dst, t);
  }
}

// dLd_cur_t = sum f(t1, y1) . grad_out  (odeint001._AdjointMethod.backward): k0 holds -f (reversed time).
__global__ void __launch_bounds__(256)
adj_time_dot_kernel(const float4* __restrict__ k0, const float4* __restrict__ adj0, int n, double* dst) {
  __shared__ double s_w[8];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 k = k0[i], a = adj0[i];
    s -= (double)k.x * a.x + (double)k.y * a.y + (double)k.z * a.z + (double)k.w * a.w;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    atomicAdd(dst, t);
  }
}

// sums: [0] dLd_cur_t, [1] adj_x, [2] adj_ctx, [3] adj_t, [4] adj_params.  Sets adj_time(t1) = -dLd_cur_t.
__global__ void adj_set_time_kernel(const double* sums, float* u0_t, float* gtimes) {
  const float d = (float)sums[0];
  u0_t[0] = -d;
  gtimes[1] = d;
}

// _select_initial_step over all eight state tensors (rtol[0], atol[0] for every one of them); see
// cnf_init_controller_kernel for the zero-dynamics quirk that reduces it to (0.01 / max d1)^(1/5).
__global__ void adj_init_controller_kernel(CnfState* st, int n, double n_ctx, double n_par, const double* sums,
                                           float t_start, float t_end) {
  const float d1x = (float)sqrt(st->sum_x) / sqrtf((float)n * 3.f);
  const float d1l = (float)sqrt(st->sum_l) / sqrtf((float)n);
  const float d1ax = (float)sqrt(sums[1]) / sqrtf((float)n * 3.f);
  const float d1c = (float)sqrt(sums[2]) / sqrtf((float)n_ctx);
  const float d1t = (float)sqrt(sums[3]);
  const float d1p = (float)sqrt(sums[4]) / sqrtf((float)n_par);
  const float d1 = fmaxf(fmaxf(fmaxf(d1x, d1l), fmaxf(d1ax, d1c)), fmaxf(d1t, d1p));
  float dt;
  if ((double)d1 < 1e-5) dt = 1e-6f;
  else dt = powf(__fdiv_rn(0.01f, d1), 1.0f / 5.0f);
  st->t = (double)t_start;
  st->t_end = (double)t_end;
  st->dt = (double)dt;
  st->sum_x = 0.0;
  st->sum_l = 0.0;
  st->nfe = 3;                  // func(t1, y1) for dL/dt1, f0 and the heuristic's probe evaluation
  st->accepted = 0;
  st->rejected = 0;
  st->status = CASPR_OK;
  st->fin_step = -1;
  st->done = (st->t_end > st->t) ? 0 : 1;
  st->first_dt = dt;
}

__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
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

__global__ void unpack4_kernel(const float4* __restrict__ src, int n, float* __restrict__ xyz, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = src[i];
  xyz[3 * (size_t)i] = v.x; xyz[3 * (size_t)i + 1] = v.y; xyz[3 * (size_t)i + 2] = v.z;
  w[i] = v.w;
}

// ----------------------------------------------------------------------------------- workspace
struct AdjWorkspace {
  CnfWorkspace base;                 // st, hyper arrays, (x, logp) state and its stage derivatives, Ha/Va (h1, v1)
  float4 *adj0, *kadj, *adj_out, *state_out;
  float *U0, *kU, *Uout;             // [ctx | t (4) | params]
  size_t nu, par_off, t_off;
  float *A1, *Ad1, *H2, *V2, *A2, *Ad2, *H3, *V3, *raw8;
  float *Gh, *Gv, *Ab, *Av;
  float *W1t, *W2t;
  char* tc_ws;                       // caspr_linear_tc operand planes for a (2n x H) activation / tangent stack
  size_t tc_ws_bytes, prep_bytes;
  char* prep[4];                     // fp16 hi/lo planes of W1, W2, W1^T, W2^T
  char* wg_ws;                       // caspr_linear_wgrad_tc workspace
  size_t wg_ws_bytes;
  unsigned* cm;                      // 4 x H column maxima (bit patterns): [h1;v1], [h2;v2], cotangents of layer 2 / 1
  float *gpart, *bpart, *w0part, *w3part, *wpart, *Ghat, *Bhat, *cpart;
  double* sums;
  float* scratch_x;                  // x(t0), logp(t0) written by the state finalize (not returned)
  int nchunk, L, nsplit, pts_per_split;
  size_t bytes;
};

AdjWorkspace adj_carve(void* base, int frames, int pts, int H, int C) {
  AdjWorkspace w;
  memset(&w, 0, sizeof(w));
  const size_t n = (size_t)frames * pts;
  const size_t ctot = hyper_ld(H);
  const ParamLayout pl = param_layout(H, C);
  // chunks of points per frame for the thread-per-channel kernels: <= 64 points each, and a CTA count that fills whole
  // waves - these kernels run 512 threads at 34-42 registers, i.e. 3 CTAs per SM = 444 slots, and 640 CTAs (40 frames
  // x 16 chunks of 64 points) would be one full wave plus a 44 % one
  const int slots = 148 * 3;
  int nchunk = (2 * 148 + frames - 1) / frames;
  if (nchunk < (pts + kChunkMax - 1) / kChunkMax) nchunk = (pts + kChunkMax - 1) / kChunkMax;
  {
    const int waves = (frames * nchunk + slots - 1) / slots;
    const int fill = waves * slots / frames;                  // chunks per frame that fill `waves` waves
    if (fill > nchunk) nchunk = fill;
  }
  if (nchunk > pts) nchunk = pts;
  w.L = (pts + nchunk - 1) / nchunk;
  w.nchunk = (pts + w.L - 1) / w.L;
  w.nsplit = 19;
  w.pts_per_split = (int)(((n + w.nsplit - 1) / w.nsplit + kWgPts - 1) / kWgPts * kWgPts);
  w.nsplit = (int)((n + w.pts_per_split - 1) / w.pts_per_split);
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
  CnfWorkspace& b = w.base;
  b.st = (CnfState*)take(sizeof(CnfState));
  b.Gc = (float*)take(frames * ctot * 4);
  b.Bc = (float*)take(frames * ctot * 4);
  b.gate = (float*)take(7 * frames * ctot * 4);
  b.biasf = (float*)take(7 * frames * ctot * 4);
  b.wg_t = (float*)take(ctot * 4);
  b.wb_t = (float*)take(ctot * 4);
  b.lbias = (float*)take(ctot * 4);
  b.y0 = (float4*)take(n * 16);
  b.y1 = (float4*)take(n * 16);
  b.kbuf = (float4*)take(7 * n * 16);
  // activation and tangent of a layer are stored back to back, so that [h ; v] is ONE (2n x H) row-major operand
  auto take_pair = [&](float** a, float** v) {
    *a = (float*)take((size_t)2 * n * H * 4);
    *v = *a + n * H;
  };
  take_pair(&b.Ha, &b.Va);
  w.adj0 = (float4*)take(n * 16);
  w.kadj = (float4*)take(7 * n * 16);
  w.adj_out = (float4*)take(n * 16);
  w.state_out = (float4*)take(n * 16);
  w.t_off = (size_t)frames * C;
  w.par_off = w.t_off + 4;
  w.nu = (w.par_off + pl.total + 3) / 4 * 4;
  w.U0 = (float*)take(w.nu * 4);
  w.kU = (float*)take(7 * w.nu * 4);
  w.Uout = (float*)take(w.nu * 4);
  take_pair(&w.A1, &w.Ad1);
  take_pair(&w.H2, &w.V2);
  take_pair(&w.A2, &w.Ad2);
  take_pair(&w.H3, &w.V3);
  take_pair(&w.Gh, &w.Gv);
  take_pair(&w.Ab, &w.Av);
  auto take1k = [&](size_t bytes) { p = (char*)align_up((size_t)p, 1024); char* r = p; p += align_up(bytes, 1024); return r; };
  w.tc_ws_bytes = caspr_linear_tc_workspace_bytes((int)(2 * n), H, H);
  w.tc_ws = take1k(w.tc_ws_bytes);
  w.prep_bytes = caspr_linear_tc_weight_bytes(H, H);
  for (int i = 0; i < 4; ++i) w.prep[i] = take1k(w.prep_bytes);
  w.wg_ws_bytes = caspr_linear_wgrad_tc_workspace_bytes((long long)(2 * n), H, H);
  w.wg_ws = take1k(w.wg_ws_bytes);
  w.cm = (unsigned*)take((size_t)4 * H * 4);
  w.raw8 = (float*)take(n * 8 * 4);
  w.W1t = (float*)take((size_t)H * H * 4);
  w.W2t = (float*)take((size_t)H * H * 4);
  const size_t nparts = (size_t)frames * w.nchunk;
  w.gpart = (float*)take(nparts * ctot * 4);
  w.bpart = (float*)take(nparts * ctot * 4);
  w.w0part = (float*)take(nparts * 3 * H * 4);
  w.w3part = (float*)take(nparts * 3 * H * 4);
  w.wpart = (float*)take((size_t)w.nsplit * H * H * 4);
  w.Ghat = (float*)take(frames * ctot * 4);
  w.Bhat = (float*)take(frames * ctot * 4);
  w.cpart = (float*)take((size_t)4 * frames * C * 4);
  w.sums = (double*)take(8 * 8);
  w.scratch_x = (float*)take(n * 16);
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

// One augmented evaluation at RK stage `stage` (hyper gates of that stage must be in place).
int enqueue_aug_eval(const AdjWorkspace& w, const caspr_cnf_weights* cw, const float* e, const float* ctx, int frames,
                     int pts, int stage, int engine, cudaStream_t s) {
  const int H = cw->hidden, C = cw->ctx_dim;
  const int n = frames * pts;
  const int ctot = hyper_ld(H);
  const CnfWorkspace& b = w.base;
  const float* gate = b.gate + (size_t)stage * frames * ctot;
  const float* biasf = b.biasf + (size_t)stage * frames * ctot;
  const ParamLayout pl = param_layout(H, C);
  ParamOffsets po;
  for (int l = 0; l < 4; ++l) {
    po.W[l] = (unsigned)pl.W[l]; po.b[l] = (unsigned)pl.b[l]; po.Wb[l] = (unsigned)pl.Wb[l];
    po.Wg[l] = (unsigned)pl.Wg[l]; po.bg[l] = (unsigned)pl.bg[l];
  }
  float* kU = w.kU + (size_t)stage * w.nu;
  float* kpar = kU + w.par_off;
  // ---- forward, keeping what the backward sweep needs
  const bool tc = engine == CASPR_CNF_TC_FP16X3;
  unsigned *cm_hv1 = nullptr, *cm_hv2 = nullptr, *cm_ab2 = nullptr, *cm_ab1 = nullptr;
  if (tc) {
    if (cudaMemsetAsync(w.cm, 0, (size_t)4 * H * 4, s) != cudaSuccess) return CASPR_ELAUNCH;
    cm_hv1 = w.cm; cm_hv2 = w.cm + H; cm_ab2 = w.cm + 2 * H; cm_ab1 = w.cm + 3 * H;
  }
  CASPR_COUNT(); cnf_layer0_kernel<<<blocks_for(n, 8, 148 * 16), 256, 0, s>>>(
      b.y0, b.kbuf, (size_t)n, e, cw->W[0], H, n, pts, stage, gate, biasf, ctot, b.st, b.Ha, b.Va, cm_hv1);
  const dim3 ggrid(ceil_div(n, kMidBM), H / kMidBN);
  const long long total4 = (long long)n * H / 4;
  if (tc && total4 >= (1ll << 31)) return CASPR_EINVAL;
  const int agrid = blocks_for(total4, 256 * 2, 148 * 16);
  // (2n x H) . W^T on the tcgen05 fp16x3 GEMM: rows [0,n) = activations, rows [n,2n) = tangents
  auto tc_product = [&](const float* X, int widx, float* Y) {
    return caspr_linear_tc(X, H, nullptr, H, nullptr, Y, H, 2 * n, H, H, CASPR_ACT_NONE, CASPR_ACT_NONE, w.prep[widx],
                           nullptr, nullptr, 0, w.tc_ws, w.tc_ws_bytes, s);
  };
  if (tc) {
    int rc = tc_product(b.Ha, 0, w.A1);
    if (rc) return rc;
    CASPR_COUNT(); adj_act_kernel<<<agrid, 256, 0, s>>>(w.A1, w.Ad1, gate + H, biasf + H, ctot, H, pts, total4, b.st, w.H2, w.V2, cm_hv2);
    rc = tc_product(w.H2, 1, w.A2);
    if (rc) return rc;
    CASPR_COUNT(); adj_act_kernel<<<agrid, 256, 0, s>>>(w.A2, w.Ad2, gate + 2 * H, biasf + 2 * H, ctot, H, pts, total4, b.st, w.H3, w.V3, nullptr);
  } else {
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidForwardKeepRaw><<<ggrid, 256, 0, s>>>(
        b.Ha, b.Va, cw->W[1], H, n, pts, gate + H, biasf + H, ctot, b.st, w.H2, w.V2, w.A1, w.Ad1);
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidForwardKeepRaw><<<ggrid, 256, 0, s>>>(
        w.H2, w.V2, cw->W[2], H, n, pts, gate + 2 * H, biasf + 2 * H, ctot, b.st, w.H3, w.V3, w.A2, w.Ad2);
  }
  CASPR_COUNT(); cnf_last_layer_kernel<<<blocks_for(n, 8, 148 * 16), 256, 0, s>>>(
      w.H3, w.V3, cw->W[3], H, n, pts, e, gate + 3 * H, biasf + 3 * H, ctot, 1, b.st, b.kbuf + (size_t)stage * n,
      w.raw8);
  CASPR_CHECK_LAUNCH();
  // ---- backward sweep
  const dim3 egrid(w.nchunk, frames);
  const dim3 wgrid(H / kWgTile, H / kWgTile, w.nsplit);
  const size_t hh = (size_t)H * H;
  CASPR_COUNT(); adj_bwd_last_kernel<<<egrid, H, 0, s>>>(
      w.adj0, w.kadj, (size_t)n, stage, b.st, e, cw->W[3], w.raw8, w.A2, w.Ad2, w.H3, w.V3, gate, biasf, b.lbias,
      ctot, H, pts, w.L, w.Ab, w.Av, w.gpart, w.bpart, w.w3part, cm_ab2);
  // weight gradient of an H x H layer: [Ab ; Av]^T . [h ; v] over the 2n stacked rows
  auto wgrad = [&](const float* Hin, const float* Vin, float* dst, const unsigned* cm_dy, const unsigned* cm_x) {
    if (tc)
      return caspr_linear_wgrad_tc(w.Ab, H, Hin, H, 2ll * n, H, H, 0, dst, cm_dy, cm_x, w.wg_ws, w.wg_ws_bytes, s);
    CASPR_COUNT(); adj_wgrad_kernel<<<wgrid, 256, 0, s>>>(w.Ab, w.Av, Hin, Vin, H, n, w.pts_per_split, b.st, w.wpart);
    CASPR_COUNT(); adj_reduce_parts_kernel<<<(unsigned)((hh + 255) / 256), 256, 0, s>>>(w.wpart, w.nsplit, hh, b.st, dst);
    return (int)CASPR_OK;
  };
  {
    const int rc = wgrad(w.H2, w.V2, kpar + pl.W[2], cm_ab2, cm_hv2);
    if (rc) return rc;
  }
  if (tc) {
    const int rc = tc_product(w.Ab, 3, w.Gh);
    if (rc) return rc;
  } else {
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidPlain><<<ggrid, 256, 0, s>>>(
        w.Ab, w.Av, w.W2t, H, n, pts, nullptr, nullptr, ctot, b.st, w.Gh, w.Gv, nullptr, nullptr);
  }
  CASPR_COUNT(); adj_bwd_mid_kernel<<<egrid, H, 0, s>>>(w.Gh, w.Gv, w.A1, w.Ad1, gate, biasf, b.lbias, ctot, H, pts,
                                                        w.L, 1, b.st, w.Ab, w.Av, w.gpart, w.bpart, cm_ab1);
  {
    const int rc = wgrad(b.Ha, b.Va, kpar + pl.W[1], cm_ab1, cm_hv1);
    if (rc) return rc;
  }
  if (tc) {
    const int rc = tc_product(w.Ab, 2, w.Gh);
    if (rc) return rc;
  } else {
    CASPR_COUNT(); cnf_mid_layer_kernel<kMidPlain><<<ggrid, 256, 0, s>>>(
        w.Ab, w.Av, w.W1t, H, n, pts, nullptr, nullptr, ctot, b.st, w.Gh, w.Gv, nullptr, nullptr);
  }
  CASPR_COUNT(); adj_bwd_layer0_kernel<<<egrid, H, 0, s>>>(
      b.y0, b.kbuf, (size_t)n, stage, b.st, e, cw->W[0], w.Gh, w.Gv, gate, biasf, b.lbias, ctot, H, pts, w.L,
      w.kadj + (size_t)stage * n, w.gpart, w.bpart, w.w0part);
  CASPR_CHECK_LAUNCH();
  // ---- hyper networks
  const int c3 = 3 * H + 3;
  CASPR_COUNT(); adj_hyper_reduce_kernel<<<ceil_div(frames * ctot, 256), 256, 0, s>>>(
      w.gpart, w.bpart, w.nchunk, frames, c3, ctot, gate, b.st, w.Ghat, w.Bhat);
  const int nsmall = ceil_div(c3, 64);         // small blocks: the kernel is a latency-bound column reduction
  CASPR_COUNT(); adj_hyper_small_kernel<<<nsmall, 64, 0, s>>>(
      w.Ghat, w.Bhat, gate, frames, c3, ctot, H, C, b.st, stage, po, w.w0part, w.w3part, frames * w.nchunk, kpar);
  HyperPtrs hp;
  for (int l = 0; l < 4; ++l) { hp.Wg[l] = cw->Wgate[l]; hp.Wb[l] = cw->Wbias[l]; }
  const dim3 cgrid(ceil_div(C, 128), ceil_div(frames, kCtxFrames), 4);
  CASPR_COUNT(); adj_ctx_bwd_kernel<<<cgrid, 128, 0, s>>>(hp, w.Ghat, w.Bhat, frames, ctot, H, C, b.st, w.cpart);
  const size_t nctx = (size_t)frames * C;
  CASPR_COUNT(); adj_ctx_sum_kernel<<<(unsigned)((nctx + 255) / 256), 256, 0, s>>>(w.cpart, nctx, b.st, kU,
                                                                                    kU + w.t_off);
  const dim3 hgrid(ceil_div(C, 256), ceil_div(c3, kHwCh));
  CASPR_COUNT(); adj_hyper_wgrad_kernel<<<hgrid, 256, 0, s>>>(w.Ghat, w.Bhat, ctx, frames, c3, ctot, H, C, b.st, po,
                                                              kpar);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

int enqueue_hyper_stages(const AdjWorkspace& w, int frames, int H, int stage_first, int stage_last, cudaStream_t s) {
  const int ctot = hyper_ld(H);
  const long long tot = (long long)frames * ctot;
  const dim3 hgrid((unsigned)((tot + 255) / 256), (unsigned)(stage_last - stage_first + 1));
  CASPR_COUNT(); cnf_hyper_stage_kernel<<<hgrid, 256, 0, s>>>(
      w.base.Gc, w.base.Bc, w.base.wg_t, w.base.wb_t, w.base.lbias, frames, 3 * H + 3, ctot, stage_first, 1,
      w.base.st, nullptr, w.base.gate, w.base.biasf);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

}  // namespace

extern "C" size_t caspr_cnf_param_count(int hidden, int ctx_dim) {
  if (hidden <= 0 || ctx_dim <= 0) return 0;
  return param_layout(hidden, ctx_dim).total;
}

extern "C" size_t caspr_cnf_adjoint_workspace_bytes(int frames, int pts, int hidden, int ctx_dim) {
  if (frames <= 0 || pts <= 0 || hidden <= 0 || ctx_dim <= 0) return 0;
  return adj_carve(nullptr, frames, pts, hidden, ctx_dim).bytes;
}

extern "C" int caspr_cnf_adjoint(const float* x1, const float* logp1, const float* gx1, const float* glogp1,
                                 const float* e, const float* ctx, int frames, int pts,
                                 const caspr_cnf_weights* cw, float end_time, float rtol, float atol, int engine,
                                 float* gx0, float* glogp0, float* gctx, float* gparams, float* gtimes,
                                 int32_t* info, int32_t* h_info, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  CASPR_REQUIRE(x1 && logp1 && gx1 && glogp1 && e && ctx && gx0 && glogp0 && gctx && gparams && gtimes);
  CASPR_REQUIRE(info && h_info && workspace);
  CASPR_REQUIRE(frames > 0 && pts > 0 && (long long)frames * pts < (1ll << 30));
  CASPR_REQUIRE(weights_ok(cw) && cw->hidden <= 512);
  CASPR_REQUIRE(end_time > 0.f);
  CASPR_REQUIRE(engine == CASPR_CNF_SIMT_FP32 || engine == CASPR_CNF_TC_FP16X3);
  CASPR_REQUIRE(((uintptr_t)workspace & 1023) == 0);
  const int H = cw->hidden, C = cw->ctx_dim;
  if (workspace_bytes < caspr_cnf_adjoint_workspace_bytes(frames, pts, H, C)) return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = frames * pts;
  AdjWorkspace w = adj_carve(workspace, frames, pts, H, C);
  const CnfWorkspace& b = w.base;
  const ParamLayout pl = param_layout(H, C);

  int rc = prepare_hyper(b, cw, ctx, frames, s);
  if (rc) return rc;
  {
    const dim3 tg(ceil_div(H, 32), ceil_div(H, 32)), tb(32, 8);
    CASPR_COUNT(); transpose_kernel<<<tg, tb, 0, s>>>(cw->W[1], H, H, w.W1t);
    CASPR_COUNT(); transpose_kernel<<<tg, tb, 0, s>>>(cw->W[2], H, H, w.W2t);
    CASPR_CHECK_LAUNCH();
  }
  if (engine == CASPR_CNF_TC_FP16X3) {
    const float* mats[4] = {cw->W[1], cw->W[2], w.W1t, w.W2t};
    for (int i = 0; i < 4; ++i) {
      rc = caspr_linear_tc_prepare_weights(mats[i], H, H, H, w.prep[i], w.prep_bytes, s);
      if (rc) return rc;
    }
  }
  const MbnDev none = load_mbn(nullptr, nullptr);
  CASPR_COUNT(); cnf_init_state_kernel<<<ceil_div(n, 256), 256, 0, s>>>(x1, logp1, n, none, 0, b.y0);
  CASPR_COUNT(); cnf_init_state_kernel<<<ceil_div(n, 256), 256, 0, s>>>(gx1, glogp1, n, none, 0, w.adj0);
  CASPR_CHECK_LAUNCH();
  if (cudaMemsetAsync(w.U0, 0, w.nu * sizeof(float), s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), s) != cudaSuccess) return CASPR_ELAUNCH;
  // reversed time: integrate -aug(-t) over [-t1, 0]
  const float t_start = -end_time, t_stop = 0.f;
  {
    CnfState h0;
    memset(&h0, 0, sizeof(h0));
    h0.t = (double)t_start;
    if (cudaMemcpyAsync(b.st, &h0, sizeof(CnfState), cudaMemcpyHostToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
    if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  }
  rc = enqueue_hyper_stages(w, frames, H, 0, 0, s);
  if (rc) return rc;
  rc = enqueue_aug_eval(w, cw, e, ctx, frames, pts, 0, engine, s);
  if (rc) return rc;
  const int eb = blocks_for(n, 256, 148 * 8);
  // dL/dt1 and adj_time(t1) = -dL/dt1 (needs f(t1, y1) = -k0)
  CASPR_COUNT(); adj_time_dot_kernel<<<eb, 256, 0, s>>>(b.kbuf, w.adj0, n, w.sums);
  CASPR_COUNT(); adj_set_time_kernel<<<1, 1, 0, s>>>(w.sums, w.U0 + w.t_off, gtimes);
  // initial-step heuristic over all state tensors
  CASPR_COUNT(); cnf_init_norm_kernel<<<eb, 256, 0, s>>>(b.y0, b.kbuf, n, rtol, atol, b.st);
  CASPR_COUNT(); adj_sumsq_kernel<<<eb, 256, 0, s>>>((const float*)w.adj0, (const float*)w.kadj, (size_t)n * 4, rtol,
                                                     atol, w.sums + 1);
  CASPR_COUNT(); adj_sumsq_kernel<<<blocks_for((long long)frames * C, 256, 148 * 8), 256, 0, s>>>(
      w.U0, w.kU, (size_t)frames * C, rtol, atol, w.sums + 2);
  CASPR_COUNT(); adj_sumsq_kernel<<<1, 256, 0, s>>>(w.U0 + w.t_off, w.kU + w.t_off, 1, rtol, atol, w.sums + 3);
  CASPR_COUNT(); adj_sumsq_kernel<<<blocks_for((long long)pl.total, 256, 148 * 8), 256, 0, s>>>(
      w.U0 + w.par_off, w.kU + w.par_off, pl.total, rtol, atol, w.sums + 4);
  CASPR_COUNT(); adj_init_controller_kernel<<<1, 1, 0, s>>>(b.st, n, (double)frames * C, (double)pl.total, w.sums,
                                                           t_start, t_stop);
  CASPR_CHECK_LAUNCH();

  const MbnDev post = none;
  const int kMaxSteps = 100000;
  int step_id = 0;
  CnfState hst;
  const int fb4 = blocks_for((long long)n * 4, 256, 148 * 8);
  const int fbu = blocks_for((long long)w.nu, 256, 148 * 8);
  for (;;) {
    rc = enqueue_hyper_stages(w, frames, H, 1, 6, s);
    if (rc) return rc;
    for (int stage = 1; stage <= 6; ++stage) {
      rc = enqueue_aug_eval(w, cw, e, ctx, frames, pts, stage, engine, s);
      if (rc) return rc;
    }
    CASPR_COUNT(); cnf_error_kernel<<<eb, 256, 0, s>>>(b.y0, b.kbuf, (size_t)n, n, rtol, atol, b.st, b.y1);
    CASPR_COUNT(); cnf_controller_kernel<<<1, 1, 0, s>>>(b.st, n, step_id);
    CASPR_COUNT(); cnf_finalize_kernel<<<eb, 256, 0, s>>>(b.y0, b.kbuf, (size_t)n, b.y1, n, step_id, b.st, post, 1, 1,
                                                         w.scratch_x, w.scratch_x + (size_t)3 * n);
    CASPR_COUNT(); flat_finalize_kernel<<<fb4, 256, 0, s>>>((float*)w.adj0, (float*)w.kadj, (size_t)n * 4,
                                                               (size_t)n * 4, step_id, b.st, (float*)w.adj_out);
    CASPR_COUNT(); flat_finalize_kernel<<<fbu, 256, 0, s>>>(w.U0, w.kU, w.nu, w.nu, step_id, b.st, w.Uout);
    CASPR_CHECK_LAUNCH();
    ++step_id;
    if (cudaMemcpyAsync(&hst, b.st, sizeof(CnfState), cudaMemcpyDeviceToHost, s) != cudaSuccess) return CASPR_ELAUNCH;
    if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
    if (hst.done) break;
    if (step_id >= kMaxSteps) { hst.status = CASPR_ESOLVER_MAXSTEPS; break; }
  }
  if (hst.status == CASPR_OK) {
    CASPR_COUNT(); unpack4_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w.adj_out, n, gx0, glogp0);
    CASPR_CHECK_LAUNCH();
    if (cudaMemcpyAsync(gctx, w.Uout, (size_t)frames * C * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(gparams, w.Uout + w.par_off, pl.total * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(gtimes, w.Uout + w.t_off, sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return CASPR_ELAUNCH;
  }
  int32_t first_dt_bits;
  memcpy(&first_dt_bits, &hst.first_dt, 4);
  int32_t out_info[8] = {hst.status, hst.nfe, hst.accepted, hst.rejected, hst.done, 0, first_dt_bits, step_id};
  for (int i = 0; i < 8; ++i) h_info[i] = out_info[i];
  if (cudaMemcpyAsync(info, h_info, 8 * sizeof(int32_t), cudaMemcpyHostToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  return hst.status;
}
