// Conditional continuous-normalizing-flow decoder: MovingBatchNorm -> CNF(dopri5) -> MovingBatchNorm.
//
// Replaces SequentialFlow / CNF / ODEfunc / ODEnet / ConcatSquashLinear / MovingBatchNorm1d of
// caspr/models/cnf.py:33-48,70-128, odefunc.py:13-31,98-105,119-142, diffeq_layers.py:76-90,
// normalization.py:59-108 and torchdiffeq 0.0.1's dopri5 for the (x, logp, context) state
// (restated in oracle/odeint001.py).
//
// Structure of one solve (everything on the device; the host only enqueues and polls `info`):
//   prepare : hoist the context part of the 8 hyper-linears (W[:,1:].c, once per solve),
//             MovingBatchNorm pre-transform, f0 = f(t_start, y0), first step size
//   step    : 6 dynamics evaluations (stage input formed on the fly from y0 and k_j),
//             error-ratio reduction, 1-thread controller (accept / dt / done), finalize
//             (FSAL shift on accept; dense-output interpolation + MovingBatchNorm post-transform
//             on the step that passes t_end).
// One dynamics evaluation = hyper_stage (gate/bias for this stage time) -> layer0 (3->H, writes
// activations h and tangents v) -> two mid layers (HxH GEMM over [h ; v] rows) -> last layer
// (H->3 + divergence e.J.e).  The divergence is carried in FORWARD mode (tangent v = J_l..J_0 e),
// mathematically equal to the reference's VJP (e^T J).e (odefunc.py:13-26).
//
// This file holds the exact-fp32 SIMT engine (CASPR_CNF_SIMT_FP32).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "dopri5.cuh"
#include "cnf_state.cuh"
#include "cnf_tc.cuh"

#include "cnf_kernels.cuh"


extern "C" size_t caspr_cnf_workspace_bytes(int frames, int pts, int hidden, int ctx_dim, int engine) {
  (void)ctx_dim;
  (void)engine;
  if (frames <= 0 || pts <= 0 || hidden <= 0) return 0;
  return carve(nullptr, frames, pts, hidden).bytes;
}

namespace {

// Lock-step step control over ranks: sum the two error-ratio accumulators (CnfState.sum_x, sum_l: adjacent doubles)
// over all ranks before the controller looks at them.  The library stays free of NCCL: the sums are staged in a
// caller-owned device buffer and the caller's hook performs the all-reduce on `stream`.
int sync_sums(const caspr_cnf_sync* sync, CnfState* st, cudaStream_t s) {
  if (!sync) return CASPR_OK;
  if (cudaMemcpyAsync(sync->stage, &st->sum_x, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (sync->allreduce_sum(sync->stage, 2, sync->user, (void*)s) != 0) return CASPR_ELAUNCH;
  if (cudaMemcpyAsync(&st->sum_x, sync->stage, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  return CASPR_OK;
}

int cnf_flow_impl(const float* x_in, const float* logp_in, const float* e, const float* ctx,
                  int frames, int pts, const caspr_cnf_weights* cw,
                  const caspr_mbn_params* mbn0, const caspr_mbn_params* mbn2,
                  float end_time, int reverse, float rtol, float atol, int engine,
                  float* x_out, float* logp_out, int32_t* info, int32_t* h_info,
                  void* workspace, size_t workspace_bytes, void* stream, const caspr_cnf_sync* sync) {
  CASPR_REQUIRE(x_in && e && ctx && x_out && info && h_info && workspace);
  CASPR_REQUIRE(frames > 0 && pts > 0 && (long long)frames * pts < (1ll << 30));
  CASPR_REQUIRE(weights_ok(cw));
  CASPR_REQUIRE(engine == CASPR_CNF_SIMT_FP32 || engine == CASPR_CNF_TC_FP16X3);
  CASPR_REQUIRE(((uintptr_t)workspace & 255) == 0);
  if (workspace_bytes < caspr_cnf_workspace_bytes(frames, pts, cw->hidden, cw->ctx_dim, engine))
    return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = frames * pts;
  CnfWorkspace w = carve(workspace, frames, pts, cw->hidden);

  // MovingBatchNorm parameters are 12 floats per layer: fetch them once (the only D2H besides info)
  float h_mbn[2][12];
  const caspr_mbn_params* mm[2] = {mbn0, mbn2};
  for (int i = 0; i < 2; ++i) {
    if (!mm[i]) continue;
    CASPR_REQUIRE(mm[i]->weight && mm[i]->bias && mm[i]->running_mean && mm[i]->running_var);
    const float* src[4] = {mm[i]->weight, mm[i]->bias, mm[i]->running_mean, mm[i]->running_var};
    for (int q = 0; q < 4; ++q)
      if (cudaMemcpyAsync(&h_mbn[i][3 * q], src[q], 3 * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess)
        return CASPR_ELAUNCH;
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  // forward: chain[0] before, chain[2] after; reverse: chain[2]^-1 before, chain[0]^-1 after
  const MbnDev pre = reverse ? load_mbn(mbn2, h_mbn[1]) : load_mbn(mbn0, h_mbn[0]);
  const MbnDev post = reverse ? load_mbn(mbn0, h_mbn[0]) : load_mbn(mbn2, h_mbn[1]);

  int rc = prepare_hyper(w, cw, ctx, frames, s);
  if (rc) return rc;
  cnf_tc::Plan plan;
  int num_sms = 148;
  rc = prepare_engine(w, cw, n, engine, &plan, &num_sms, s);
  if (rc) return rc;
  const int eb = blocks_for(n, 256, 148 * 8);
  CASPR_COUNT(); cnf_init_state_kernel<<<ceil_div(n, 256), 256, 0, s>>>(x_in, logp_in, n, pre, reverse, w.y0);
  CASPR_CHECK_LAUNCH();
  // odeint001.odeint: decreasing times are integrated as -f(-t) over [-T, 0]
  const float t_start = reverse ? -end_time : 0.f;
  const float t_stop = reverse ? 0.f : end_time;
  // f0 (stage 0 reads st->t for the stage time: set it first)
  {
    CnfState h0;
    memset(&h0, 0, sizeof(h0));
    h0.t = (double)t_start;
    if (cudaMemcpyAsync(w.st, &h0, sizeof(CnfState), cudaMemcpyHostToDevice, s) != cudaSuccess)
      return CASPR_ELAUNCH;
    if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;   // h0 is a stack object
  }
  rc = enqueue_stages(w, cw, e, frames, pts, 0, 0, reverse, engine, &plan, num_sms, s);
  if (rc) return rc;
  // number of points the controller's means run over: this rank's, or all ranks' in lock-step mode
  const int n_ctrl = sync ? (int)sync->n_global : n;
  CASPR_COUNT(); cnf_init_norm_kernel<<<eb, 256, 0, s>>>(w.y0, w.kbuf, n, rtol, atol, w.st);
  rc = sync_sums(sync, w.st, s);
  if (rc) return rc;
  CASPR_COUNT(); cnf_init_controller_kernel<<<1, 1, 0, s>>>(w.st, n_ctrl, t_start, t_stop);
  CASPR_CHECK_LAUNCH();

  const int have_logp = logp_in != nullptr;
  if (!(t_stop > t_start)) {
    CASPR_COUNT(); cnf_passthrough_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w.y0, n, post, reverse, have_logp, x_out, logp_out);
    CASPR_CHECK_LAUNCH();
  }
  // Steps enqueued between two polls of the solver state.  Every kernel of a step exits at once when the solve has
  // finished, so enqueueing one speculative step costs ~10 empty launches at the end of the solve and hides the
  // launch + synchronisation gap (~60 us) of every other step.  Lock-step mode polls every step (its hook is a
  // collective all ranks must call the same number of times).
  const int kBatch = sync ? 1 : 2, kMaxSteps = 100000;
  int step_id = 0;
  CnfState hst;
  for (;;) {
    for (int b = 0; b < kBatch; ++b, ++step_id) {
      rc = enqueue_stages(w, cw, e, frames, pts, 1, 6, reverse, engine, &plan, num_sms, s);
      if (rc) return rc;
      CASPR_COUNT(); cnf_error_kernel<<<eb, 256, 0, s>>>(w.y0, w.kbuf, (size_t)n, n, rtol, atol, w.st, w.y1);
      rc = sync_sums(sync, w.st, s);
      if (rc) return rc;
      CASPR_COUNT(); cnf_controller_kernel<<<1, 1, 0, s>>>(w.st, n_ctrl, step_id);
      CASPR_COUNT(); cnf_finalize_kernel<<<eb, 256, 0, s>>>(w.y0, w.kbuf, (size_t)n, w.y1, n, step_id, w.st, post, reverse,
                                             have_logp, x_out, logp_out);
      CASPR_CHECK_LAUNCH();
    }
    if (cudaMemcpyAsync(&hst, w.st, sizeof(CnfState), cudaMemcpyDeviceToHost, s) != cudaSuccess)
      return CASPR_ELAUNCH;
    if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
    if (hst.done) break;
    if (step_id >= kMaxSteps) { hst.status = CASPR_ESOLVER_MAXSTEPS; break; }
  }
  if (engine == CASPR_CNF_TC_FP16X3) {
    release_l2_persistence(s);
    int h_range = 0;
    if (cudaMemcpyAsync(&h_range, w.range_flag, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      return CASPR_ELAUNCH;
    if (h_range) hst.status = CASPR_ERANGE;      // an fp16 operand overflowed: results are not trustworthy
  }
  int32_t first_dt_bits;
  memcpy(&first_dt_bits, &hst.first_dt, 4);
  int32_t out_info[8] = {hst.status, hst.nfe, hst.accepted, hst.rejected, hst.done, 0,
                         first_dt_bits, step_id};
  for (int i = 0; i < 8; ++i) h_info[i] = out_info[i];
  if (cudaMemcpyAsync(info, h_info, 8 * sizeof(int32_t), cudaMemcpyHostToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  return hst.status;
}

}  // namespace

extern "C" int caspr_cnf_flow(const float* x_in, const float* logp_in, const float* e, const float* ctx,
                              int frames, int pts, const caspr_cnf_weights* cw,
                              const caspr_mbn_params* mbn0, const caspr_mbn_params* mbn2,
                              float end_time, int reverse, float rtol, float atol, int engine,
                              float* x_out, float* logp_out, int32_t* info, int32_t* h_info,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return cnf_flow_impl(x_in, logp_in, e, ctx, frames, pts, cw, mbn0, mbn2, end_time, reverse, rtol, atol, engine,
                       x_out, logp_out, info, h_info, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int caspr_cnf_flow_lockstep(const float* x_in, const float* logp_in, const float* e, const float* ctx,
                                       int frames, int pts, const caspr_cnf_weights* cw,
                                       const caspr_mbn_params* mbn0, const caspr_mbn_params* mbn2,
                                       float end_time, int reverse, float rtol, float atol, int engine,
                                       float* x_out, float* logp_out, int32_t* info, int32_t* h_info,
                                       void* workspace, size_t workspace_bytes, void* stream,
                                       const caspr_cnf_sync* sync) {
  CASPR_REQUIRE(sync && sync->allreduce_sum && sync->stage && sync->n_global >= (long long)frames * pts &&
                sync->n_global < (1ll << 30));
  return cnf_flow_impl(x_in, logp_in, e, ctx, frames, pts, cw, mbn0, mbn2, end_time, reverse, rtol, atol, engine,
                       x_out, logp_out, info, h_info, workspace, workspace_bytes, stream, sync);
}

extern "C" int caspr_cnf_feval(const float* y, const float* e, const float* ctx, int frames, int pts,
                               const caspr_cnf_weights* cw, float t, int engine, float* dy,
                               float* neg_div, void* workspace, size_t workspace_bytes, void* stream) {
  CASPR_REQUIRE(y && e && ctx && dy && neg_div && workspace);
  CASPR_REQUIRE(frames > 0 && pts > 0 && weights_ok(cw));
  CASPR_REQUIRE(engine == CASPR_CNF_SIMT_FP32 || engine == CASPR_CNF_TC_FP16X3);
  CASPR_REQUIRE(((uintptr_t)workspace & 255) == 0);
  if (workspace_bytes < caspr_cnf_workspace_bytes(frames, pts, cw->hidden, cw->ctx_dim, engine))
    return CASPR_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = frames * pts;
  CnfWorkspace w = carve(workspace, frames, pts, cw->hidden);
  CnfState h0;
  memset(&h0, 0, sizeof(h0));
  h0.t = (double)t;
  if (cudaMemcpyAsync(w.st, &h0, sizeof(CnfState), cudaMemcpyHostToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaStreamSynchronize(s) != cudaSuccess) return CASPR_ELAUNCH;
  int rc = prepare_hyper(w, cw, ctx, frames, s);
  if (rc) return rc;
  cnf_tc::Plan plan;
  int num_sms = 148;
  rc = prepare_engine(w, cw, n, engine, &plan, &num_sms, s);
  if (rc) return rc;
  MbnDev none = load_mbn(nullptr, nullptr);
  CASPR_COUNT(); cnf_init_state_kernel<<<ceil_div(n, 256), 256, 0, s>>>(y, nullptr, n, none, 0, w.y0);
  CASPR_CHECK_LAUNCH();
  rc = enqueue_stages(w, cw, e, frames, pts, 0, 0, 0, engine, &plan, num_sms, s);
  if (rc) return rc;
  // unpack k0 -> dy (n,3), neg_div (n)
  if (cudaMemcpy2DAsync(dy, 12, w.kbuf, 16, 12, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaMemcpy2DAsync(neg_div, 4, (const char*)w.kbuf + 12, 16, 4, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return CASPR_ELAUNCH;
  return CASPR_OK;
}

// Profiling aid: with CASPR_CNF_FUSED_DEBUG=1 the fused evaluation kernel records, per CTA, the cycles its TMA producer
// and MMA threads spent waiting ([0] stage free, [1] layer-0 / layer-1 output ready, [2] producer total, [4] accumulator
// free, [5] operands landed, [6] MMA thread total) during its LAST launch.  Slots [8 + 8g + k] / [8 + 8g + 4 + k]: cycles epilogue group g waited for / worked on items of
// kind k (L1.n0, L1.n1, L2.n0, L2.n1).  24 counters per CTA; copies count (<= 148*24) counters to `out`.
extern "C" int caspr_cnf_fused_debug_read(long long* out, int count) {
  CASPR_REQUIRE(out && count > 0 && count <= 148 * 24);
  long long* buf = cnf_tc::fused_debug_buffer();
  if (!buf) return CASPR_EINVAL;
  if (cudaDeviceSynchronize() != cudaSuccess) return CASPR_ELAUNCH;
  if (cudaMemcpy(out, buf, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return CASPR_ELAUNCH;
  return CASPR_OK;
}
