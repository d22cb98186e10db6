/*
 * caspr_b200.h — C-ABI of libcaspr_b200.so (hand-written sm_100a CUDA for the
 * CaSPR reconstruction hot path).
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name starts with `h_` (host);
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never
 *     allocates device memory, never frees and never keeps a pointer after return;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); only the
 *     two *_solve entry points synchronise that stream (documented there);
 *   - returns 0 on success, a negative caspr_status otherwise; never throws/aborts;
 *   - fp32 data, int32 indices; "rows x channels" (channels-last) activations.
 *
 * File:line citations name the reference interface each entry point replaces
 * (paths relative to /root/reference/caspr/models unless stated).
 */
#ifndef CASPR_B200_H_
#define CASPR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CASPR_OK = 0,
  CASPR_EINVAL = -1,        /* bad shape / null pointer / unsupported size      */
  CASPR_ELAUNCH = -2,       /* cudaGetLastError() != cudaSuccess after a launch */
  CASPR_EWORKSPACE = -3,    /* workspace too small                              */
  CASPR_ESOLVER_DT = -4,    /* torchdiffeq's "underflow in dt" assertion         */
  CASPR_ESOLVER_NONFINITE = -5, /* torchdiffeq's "non-finite values in state"    */
  CASPR_ESOLVER_MAXSTEPS = -6,  /* max_num_steps exceeded                        */
  CASPR_ERANGE = -7         /* split-precision operand left the fp16 range      */
} caspr_status;

/* Library / build identification. */
int caspr_version(void);                    /* e.g. 100 = 0.1.0                       */
const char* caspr_build_arch(void);         /* "sm_100a"                              */
const char* caspr_status_string(int status);

/* Number of kernels the library has launched in this process (bench.py: gpu_launches). */
unsigned long long caspr_launch_count(void);
/* Kernels replayed from a CUDA graph are launched by the driver, not through the entry points: the
 * host adds the number of kernel nodes of a replayed graph here. */
void caspr_launch_count_add(unsigned long long n);

/* CUDA-event timing of the dominant kernels, measured on the stream they are launched on.
 * caspr_profile_enable(1) clears the records and starts recording one event pair around every
 * launch of the kernels below; caspr_profile_read waits for the recorded events and returns the
 * summed device time and the number of launches recorded. */
enum { CASPR_PROF_CNF_MID_SIMT = 0, CASPR_PROF_CNF_FUSED_TC = 1, CASPR_PROF_LINEAR = 2, CASPR_PROF_FPS = 3,
       CASPR_PROF_CNF_EVAL_FUSED = 4 /* cnf_fused_eval_kernel: one whole dynamics evaluation per launch */ };
void caspr_profile_enable(int on);
int caspr_profile_read(int kernel_id, double* total_ms, long long* launches);

/* ------------------------------------------------------------------ geometry
 * Replaces the Kaolin ops imported at pointnet2.py:7-10.  Arithmetic is the
 * declared canonical one of oracle/pointnet2_ops.py: d2 = ((dx*dx)+(dy*dy))+(dz*dz)
 * in fp32 without FMA contraction, ties -> lowest index, FPS skips |p|^2 <= 1e-3. */

/* furthest_point_sampling + fps_gather_by_index (pointnet2.py:384-387).
 * xyz (B,N,3) -> idx (B,M) int32 and, if new_xyz != NULL, new_xyz (B,M,3). */
int caspr_fps(const float* xyz, int B, int N, int M, int32_t* idx, float* new_xyz, void* stream);

/* Ball query of PointNet2GroupingLayer for BOTH scales of one set-abstraction level
 * in one scan (pointnet2.py:340-342,391): for each centre the first ns points in index
 * order with d2 < r^2, unfilled slots = first hit.  r0 <= r1 required.
 * idx0 (B,M,ns0), idx1 (B,M,ns1) int32; either may be NULL. */
int caspr_ball_query2(const float* xyz, const float* new_xyz, int B, int N, int M,
                      float r0, int ns0, int32_t* idx0, float r1, int ns1, int32_t* idx1,
                      void* stream);

/* group_gather_by_index x2 + centre subtraction + channel concat (pointnet2.py:391-398).
 * feat is channels-last (B,N,C) with row stride ld_feat floats (C may be 0, feat NULL).
 * out rows = B*M*ns, each row [dx,dy,dz | feat(C)], row stride ld_out floats (>= 3+C). */
int caspr_group_points(const float* xyz, const float* new_xyz, const float* feat, int ld_feat,
                       const int32_t* idx, int B, int N, int M, int C, int ns, float* out, int ld_out,
                       void* stream);

/* three_nn (pointnet2.py:514): unknown (B,n,3), known (B,m,3) -> dist (B,n,3) Euclidean,
 * idx (B,n,3) int32; strict '<' running best-3 over ascending index. */
int caspr_three_nn(const float* unknown, const float* known, int B, int n, int m,
                   float* dist, int32_t* idx, void* stream);

/* inverse-distance weights + three_interpolate + skip concat (pointnet2.py:516-523).
 * feat_prev (B,m,Cp) channels-last with row stride ld_prev; skip (B,n,Cs) row stride ld_skip
 * (Cs may be 0); out (B,n,Cp+Cs) row stride ld_out = [interp | skip]. */
int caspr_three_interp_concat(const float* feat_prev, int ld_prev, const int32_t* idx, const float* dist,
                              const float* skip, int ld_skip, int B, int n, int m, int Cp, int Cs,
                              float* out, int ld_out, void* stream);

/* ------------------------------------------------------------------ dense ops
 * Replace the Conv1d(k=1)/Linear/GroupNorm/ReLU/max calls of pointnet2.py:637-642,
 * 677-699,471-481,207-212; pointnet.py:27-46; tpointnet2.py:59-62,99-112. */

enum { CASPR_ACT_NONE = 0, CASPR_ACT_RELU = 1, CASPR_ACT_SIGMOID = 2 };

/* Y[r, 0:Cout] = act_out( act_in(X[r, 0:Cin]) . W[0:Cout, 0:Cin]^T + bias ),  fp32 SIMT.
 * X row stride ldx, W row stride ldw, Y row stride ldy (floats). bias may be NULL. */
int caspr_linear(const float* X, int ldx, const float* W, int ldw, const float* bias,
                 float* Y, int ldy, int rows, int Cin, int Cout, int act_in, int act_out,
                 void* stream);

/* One per-ball layer of a set-abstraction scale in a single kernel (pointnet2.py:637-642,677-699 applied to
 * (B'*M, C, ns)): Y = ReLU?(GroupNorm(16, Cout)(X.W^T + b)) with the statistics taken over each ball of `ns`
 * (16 or 32) consecutive rows, and optionally maxout[ball, :] = max over the ball's rows of that result
 * (the max-pool of pointnet2.py:698).  Cout <= 64, Cout % 16 == 0, rows % ns == 0.  Y and/or maxout. */
int caspr_linear_gn_ball(const float* X, int ldx, const float* W, int ldw, const float* bias,
                         const float* gamma, const float* beta, float eps, int rows, int Cin, int Cout,
                         int ns, int relu, float* Y, int ldy, float* maxout, int ld_max, void* stream);

/* Per-ball MLP of a set-abstraction scale on the tensor cores (pointnet2.py:649-708 as used by SA levels 3-5):
 * X (rows, Cin) grouped rows (caspr_group_points), balls of `ns` (16 or 32) consecutive rows ->
 * [Conv1d(k=1) + GroupNorm(16) per ball + ReLU] x 2 -> Conv1d + GroupNorm -> max over the ball -> maxout (rows/ns, C3).
 * The GroupNorm of every layer runs in the epilogue of its tcgen05 GEMM (a 128-row accumulator tile holds whole
 * balls), which emits the next layer's fp16 operand planes directly.  prep1..3: caspr_linear_tc_prepare_weights
 * blocks of the three weight matrices.  Widths in {64, 96, 128, 256, 512}, Cin >= 64 (caspr_sa_mlp_tc_supported). */
int caspr_sa_mlp_tc_supported(int ns, int Cin, int C1, int C2, int C3);
size_t caspr_sa_mlp_tc_workspace_bytes(long long rows, int Cin, int C1, int C2);
int caspr_sa_mlp_tc(const float* X, int ldx, long long rows, int Cin, int ns,
                    const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                    const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                    const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                    float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes, void* stream);

/* The same with the group gather of caspr_group_points folded into the operand split: the grouped rows
 * [xyz[idx] - centre | feat[idx]] (B*M*ns x (3 + C)) are formed on the fly and never exist in memory.  Arguments as
 * caspr_group_points (xyz (B,N,3), new_xyz (B,M,3), feat (B,N,C) with row stride ld_feat, idx (B,M,ns)); workspace as
 * caspr_sa_mlp_tc_workspace_bytes(B*M*ns, 3 + C, C1, C2). */
int caspr_sa_mlp_tc_grouped(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                            const int32_t* idx, int B, int N, int M, int ns,
                            const void* prep1, const float* b1, const float* g1, const float* e1, int C1,
                            const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                            const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                            float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes, void* stream);

/* The same with the first layer's product taken BEFORE the gather ("delayed aggregation"):
 *   W1 . [xyz[idx] - centre | feat[idx]] + b1 = P[idx] + W1[:, :3] . (xyz[idx] - centre) + b1,
 * P (B*N, C1) = feat . W1[:, 3:]^T computed by the caller once per source point (caspr_linear / caspr_linear_tc; row
 * stride ldp, 16-byte aligned rows).  One kernel gathers P, adds the xyz term, applies the per-ball GroupNorm + ReLU and
 * writes the second layer's operand planes; layers 2 and 3 as in caspr_sa_mlp_tc.  W1: the first layer's weight
 * (C1, 3 + C) with row stride ldw1 (only its first three columns are read).  C1 in {64, 128, 256}. */
size_t caspr_sa_mlp_tc_delayed_workspace_bytes(long long rows, int C1, int C2);
int caspr_sa_mlp_tc_delayed(const float* xyz, const float* new_xyz, const float* P, int ldp, const int32_t* idx,
                            int B, int N, int M, int ns, const float* W1, int ldw1, const float* b1, const float* g1,
                            const float* e1, int C1,
                            const void* prep2, const float* b2, const float* g2, const float* e2, int C2,
                            const void* prep3, const float* b3, const float* g3, const float* e3, int C3,
                            float eps, float* maxout, int ld_max, void* workspace, size_t workspace_bytes, void* stream);

/* One whole scale of a set-abstraction level in a single kernel (pointnet2.py:391-401,649-708): group gather
 * ([xyz[idx]-centre | feat[idx]], caspr_group_points) -> three layers Conv1d(k=1) + GroupNorm(16) (+ReLU after the
 * first two, pointnet2.py:693) with per-ball statistics -> max over the ball's ns rows.  Activations never leave
 * registers.  Supported shapes (caspr_sa_fused_supported): ns 16 or 32, layer widths (16,16,32), (32,32,64), at most
 * 136 input channels - the four scales of SA levels 1-2.  feat (B,N,C) channels-last with row stride ld_feat (C may be
 * 0), idx (B,M,ns) from caspr_ball_query2, weights row-major (Cout,Cin), out (B*M, C3) with row stride ld_out. */
int caspr_sa_fused_supported(int ns, int Cin, int C1, int C2, int C3);
int caspr_sa_fused(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                   const int32_t* idx, int B, int N, int M, int ns,
                   const float* W1, const float* b1, const float* g1, const float* e1, int C1,
                   const float* W2, const float* b2, const float* g2, const float* e2, int C2,
                   const float* W3, const float* b3, const float* g3, const float* e3, int C3,
                   float eps, float* out, int ld_out, void* stream);

/* The same scale on the tensor cores (warp-level mma.sync m16n8k16, fp16 hi/lo 3-product split, fp32 accumulate; a
 * warp owns a 32-row tile and keeps every activation in accumulator / operand fragments).  Shapes of
 * caspr_sa_mma_supported: the four scales of SA levels 1-2 (Cin = 3 + C = 9 or 99).  in_absmax: device scalar
 * >= every |entry| of the gathered rows (caspr_sa_absmax computes max(|feat|, 2 max|xyz|)), from which the operand
 * scale is taken; NULL = unscaled.  Requires B*N*max(ld_feat,3) < 2^32. */
int caspr_sa_mma_supported(int ns, int Cin, int C1, int C2, int C3);
int caspr_sa_absmax(const float* xyz, const float* feat, int ld_feat, int C, int B, int N, float* absmax,
                    void* stream);
int caspr_sa_mma(const float* xyz, const float* new_xyz, const float* feat, int ld_feat, int C,
                 const int32_t* idx, int B, int N, int M, int ns,
                 const float* W1, const float* b1, const float* g1, const float* e1, int C1,
                 const float* W2, const float* b2, const float* g2, const float* e2, int C2,
                 const float* W3, const float* b3, const float* g3, const float* e3, int C3,
                 float eps, const float* in_absmax, float* out, int ld_out, void* stream);

/* Same contract on the tcgen05 tensor cores: 3-product fp16 split ("fp16x3",
 * X_hi.W_hi + X_lo.W_hi + X_hi.W_lo, fp32 accumulate in TMEM), operands scaled per call by powers of
 * two taken from max|X| and max|W| (undone exactly in the epilogue).  Meant for the large layers
 * (feature propagation, final layers, the 1600-wide head); any Cin / Cout / rows are accepted.
 * workspace: caspr_linear_tc_workspace_bytes(rows, Cin, Cout) bytes, 1024-byte aligned. */
/* GroupNorm folded around the tensor-core linear (saves the separate statistics and normalisation passes
 * of a Conv1d -> GroupNorm -> ReLU -> Conv1d chain, pointnet2.py:471-481, tpointnet2.py:99-100):
 *   out_stats : the epilogue accumulates per (sample, group) the sum and the sum of squares of the OUTPUT
 *               rows (fp64 [samples][groups][2], zeroed by the call); rows_per_sample % 32 == 0.
 *   caspr_gn_table turns them into per (sample, channel) fp32 pairs (scale, shift) = (rstd*gamma,
 *               beta - mean*rstd*gamma), [samples][C][2].
 *   in_norm   : the operand split of the next layer applies ReLU?(x*scale + shift) on the fly.
 * caspr_groupnorm(..., stats_ready = 1) consumes the same statistics when the normalised tensor itself
 * (or its max-pool) is needed. */
typedef struct {
  const float* table;        /* [samples][C] (scale, shift) from caspr_gn_table */
  int rows_per_sample, relu;
} caspr_gn_fold;
typedef struct {
  double* stats;             /* [samples][groups][2] */
  int rows_per_sample, groups;
  /* Optional (may be NULL): [samples][2][Cout] order-preserving uint keys of the per-channel MAX and MIN of the output
   * over each sample's rows (initialised by the call).  When set, Y is NOT written (may be NULL): a GroupNorm followed
   * by a max-pool only needs the statistics and these extrema - the affine map x -> (x - mean) * rstd * gamma + beta is
   * monotone per channel, so its maximum over the rows is its value at the max (gamma >= 0) or the min (gamma < 0) of x,
   * bit for bit: caspr_gn_max_from_extrema.  The activation itself never exists in memory (pointnet.py:40-42). */
  unsigned* extrema;
} caspr_gn_stats;
/* maxout[s, c] = max over the rows of sample s of GroupNorm(x)[row, c], from the statistics and extrema above. */
int caspr_gn_max_from_extrema(const double* stats, const unsigned* extrema, int samples, int rows_per_sample, int C,
                              int groups, const float* gamma, const float* beta, float eps, float* maxout,
                              int ld_max, void* stream);
int caspr_gn_table(const double* stats, int samples, int groups, int rows_per_sample, int C, float eps,
                   const float* gamma, const float* beta, float* table, void* stream);

size_t caspr_linear_tc_workspace_bytes(int rows, int Cin, int Cout);
/* Optional: split the weights once (caspr_linear_tc_weight_bytes bytes, 1024-byte aligned) and pass the
 * block as `prepared_weights` for as long as W does not change; with prepared_weights == NULL the split
 * is redone inside every call (W must then be non-NULL). */
size_t caspr_linear_tc_weight_bytes(int Cin, int Cout);
int caspr_linear_tc_prepare_weights(const float* W, int ldw, int Cin, int Cout, void* prepared,
                                    size_t prepared_bytes, void* stream);
int caspr_linear_tc(const float* X, int ldx, const float* W, int ldw, const float* bias,
                    float* Y, int ldy, int rows, int Cin, int Cout, int act_in, int act_out,
                    const void* prepared_weights, const caspr_gn_fold* in_norm, const caspr_gn_stats* out_stats,
                    int bias_rows_per_sample, void* workspace, size_t workspace_bytes, void* stream);

/* GroupNorm over samples of `rows_per_sample` consecutive rows (torch.nn.GroupNorm(groups, C)
 * on (samples, C, rows_per_sample)), eps as given; optional ReLU; optional max over the rows
 * of each sample AFTER normalisation (and after ReLU if set) into maxout (samples, C) with
 * row stride ld_max.  If write_back == 0 X is left untouched (only maxout is produced).
 * stats_ws: >= samples*groups*2 doubles (used when a sample spans several CTAs); with stats_ready != 0 it
 * already holds the (sum, sum of squares) pairs (caspr_linear_tc out_stats) and the statistics pass is skipped. */
int caspr_groupnorm(float* X, int ldx, int samples, int rows_per_sample, int C, int groups,
                    const float* gamma, const float* beta, float eps, int relu,
                    int write_back, float* maxout, int ld_max, double* stats_ws, int stats_ready,
                    void* stream);

/* The encoder head's tail in one pass (tpointnet2.py:104-113): GroupNorm with ready statistics (as above,
 * stats_ready = 1, no ReLU, X left untouched), max over the rows of each sample into maxout (may be NULL), and a
 * narrow linear layer on the ReLU of the normalised row while it is in registers:
 *   out[row, 0..P) = act(W (P, C) . relu(gn(X[row])) + bias),  1 <= P <= 4, act = CASPR_ACT_*.
 * Requires (C/groups) % 4 == 0, ldx % 4 == 0, C <= 2048 and 16-byte aligned X, gamma, beta, W. */
int caspr_groupnorm_project(const float* X, int ldx, int samples, int rows_per_sample, int C, int groups,
                            const float* gamma, const float* beta, float eps, float* maxout, int ld_max,
                            const double* stats, const float* W, const float* bias, int P, int act,
                            float* out, int ld_out, void* stream);

/* tpointnet2.py:79-90: x (R,4) rows [x,y,z,t] -> out (R,9) rows [x,y,z,x2,y2,z2,xz,xy,yz]. */
int caspr_augment_xyz(const float* x4, int rows, float* out9, void* stream);

/* xyz (R,3) contiguous from x (R,4). */
int caspr_strip_time(const float* x4, int rows, float* xyz3, void* stream);

/* pointnet.py:44-46 / tpointnet2.py:96: dst[s*rows_per_sample + r, 0:C] = src[s, 0:C]. */
int caspr_broadcast_rows(const float* src, int ld_src, int samples, int rows_per_sample, int C,
                         float* dst, int ld_dst, void* stream);

/* ---------------------------------------------------------------- latent ODE
 * Replaces LatentODE.forward -> ODESolver -> torchdiffeq.odeint_adjoint(dopri5)
 * (latent_ode_model.py:45-70,98) with DynamicsNet (latent_ode_model.py:129-147):
 * Linear(D,H) tanh Linear(H,H) tanh Linear(H,H) tanh Linear(H,D).
 * One persistent CTA runs the whole adaptive solve (torchdiffeq 0.0.1 semantics,
 * oracle/odeint001.py) for all B sequences: the step controller is batch-global.
 * z0 (B,D); h_times: nT float64 HOST values, strictly increasing, times[0] = start;
 * out (nT,B,D); info (8 int32, device): [status, nfe, accepted steps, rejected steps,..].
 * Synchronises `stream` before returning (reads info) and maps solver failures to
 * CASPR_ESOLVER_*.  h_info (8 int32, host) receives a copy of info. */
size_t caspr_latent_ode_workspace_bytes(int B, int D, int H);
int caspr_latent_ode_solve(const float* z0, int B, int D, int H,
                           const float* W0, const float* b0, const float* W1, const float* b1,
                           const float* W2, const float* b2, const float* W3, const float* b3,
                           const double* h_times, int nT, float rtol, float atol,
                           float* out, int32_t* info, int32_t* h_info,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------- CNF
 * Replaces SequentialFlow/CNF/ODEfunc/ODEnet/ConcatSquashLinear/MovingBatchNorm1d
 * (cnf.py:33-48,70-128; odefunc.py:13-31,98-105,119-142; diffeq_layers.py:76-90;
 * normalization.py:59-108) and torchdiffeq's dopri5 for the 3-tuple state. */

typedef struct {
  /* ODEnet main weights, row-major (out,in): W1 (H,3) W2 (H,H) W3 (H,H) W4 (3,H) */
  const float* W[4];
  const float* b[4];
  /* hyper nets, row-major (out, 1+ctx): column 0 multiplies t */
  const float* Wgate[4];
  const float* bgate[4];
  const float* Wbias[4];
  int hidden;            /* H = 512 */
  int ctx_dim;           /* 1600    */
} caspr_cnf_weights;

typedef struct {
  /* MovingBatchNorm1d (eval): chain[0] (data side) and chain[2] (base side): 3 floats each */
  const float* weight;
  const float* bias;
  const float* running_mean;
  const float* running_var;
} caspr_mbn_params;

enum { CASPR_CNF_SIMT_FP32 = 0, CASPR_CNF_TC_FP16X3 = 1 };

/* Workspace for a solve over `frames` contexts x `pts` points each. */
size_t caspr_cnf_workspace_bytes(int frames, int pts, int hidden, int ctx_dim, int engine);

/* One full flow evaluation x -> chain (forward: MBN0, CNF 0->T, MBN2; reverse: MBN2^-1,
 * CNF T->0, MBN0^-1), dopri5 rtol/atol on (x, logp), context carried as a zero-dynamics state.
 *   x_in (frames,pts,3); logp_in (frames,pts) or NULL (treated as zeros, cnf.py:71-74);
 *   e (frames,pts,3) Hutchinson noise (odefunc.py:127-128), fixed for the whole solve;
 *   ctx (frames,ctx_dim); end_time = sqrt_end_time^2 (cnf.py:89-91);
 *   x_out (frames,pts,3); logp_out (frames,pts) or NULL.
 * info/h_info: 8 int32 [status, nfe, accepted, rejected, ...].  Synchronises `stream`
 * (polls the device-side controller every few steps). */
int caspr_cnf_flow(const float* x_in, const float* logp_in, const float* e, const float* ctx,
                   int frames, int pts, const caspr_cnf_weights* w,
                   const caspr_mbn_params* mbn0, const caspr_mbn_params* mbn2,
                   float end_time, int reverse, float rtol, float atol, int engine,
                   float* x_out, float* logp_out, int32_t* info, int32_t* h_info,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Lock-step step control for a batch sharded over ranks: torchdiffeq's controller takes the mean
 * error ratio over the WHOLE state tensor, so an unsharded batch shares one step sequence.  With this variant every
 * rank sums its two error-ratio accumulators with all other ranks before each accept / reject decision (and in the
 * initial-step heuristic), which reproduces the unsharded step sequence.  The library does not link NCCL: the two
 * doubles are staged in `stage` (device, caller-owned) and `allreduce_sum(stage, 2, user, stream)` must enqueue an
 * in-place sum all-reduce on `stream` and return 0.  n_global = total number of points over all ranks. */
typedef struct {
  long long n_global;
  double* stage;
  int (*allreduce_sum)(double* device_values, int count, void* user, void* stream);
  void* user;
} caspr_cnf_sync;
int caspr_cnf_flow_lockstep(const float* x_in, const float* logp_in, const float* e, const float* ctx,
                            int frames, int pts, const caspr_cnf_weights* w,
                            const caspr_mbn_params* mbn0, const caspr_mbn_params* mbn2,
                            float end_time, int reverse, float rtol, float atol, int engine,
                            float* x_out, float* logp_out, int32_t* info, int32_t* h_info,
                            void* workspace, size_t workspace_bytes, void* stream, const caspr_cnf_sync* sync);

/* One dynamics evaluation (dy, -div) = ODEfunc(t, (y, logp, ctx)) for testing/profiling:
 * y (frames,pts,3), e (frames,pts,3) -> dy (frames,pts,3), neg_div (frames,pts). */
int caspr_cnf_feval(const float* y, const float* e, const float* ctx, int frames, int pts,
                    const caspr_cnf_weights* w, float t, int engine,
                    float* dy, float* neg_div, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------ CNF training (adjoint backward)
 * Replaces torchdiffeq 0.0.1's OdeintAdjointMethod.backward for the CNF block (call site cnf.py:102-111) and the
 * autograd VJP through ODEfunc / divergence_approx / ODEnet (odefunc.py:13-31,98-105,119-142): integrates the
 * augmented state (x, logp, ctx, adj_x, adj_logp, adj_ctx, adj_t, adj_params) from t1 = end_time back to 0 with
 * dopri5; step control sees (x, logp, ctx) only (three-entry tolerance lists, cnf.py:80-81), the initial-step
 * heuristic sees all eight tensors.
 *   x1 (frames,pts,3), logp1 (frames,pts): the block's outputs at t1;  gx1, glogp1: dL/d(x1), dL/d(logp1);
 *   e, ctx, w, rtol, atol: as in the forward caspr_cnf_flow call (MovingBatchNorm layers are NOT part of this call);
 *   gx0 (frames,pts,3), glogp0 (frames,pts), gctx (frames,ctx_dim): dL/d(x0), dL/d(logp0), dL/d(ctx);
 *   gparams: caspr_cnf_param_count(hidden, ctx_dim) floats in ODEfunc.parameters() order — per layer
 *            _layer.weight, _layer.bias, _hyper_bias.weight, _hyper_gate.weight, _hyper_gate.bias;
 *   gtimes (2 floats, device): dL/dt0, dL/dt1 (dL/d sqrt_end_time = 2 sqrt_end_time gtimes[1], cnf.py:89-91);
 *   engine: CASPR_CNF_SIMT_FP32 (exact fp32) or CASPR_CNF_TC_FP16X3 (all six H x H products of an evaluation —
 *           forward, data gradient, weight gradient — on the tcgen05 fp16x3 GEMM); workspace 1024-byte aligned.
 * Synchronises `stream` once per attempted step (polls the device-side controller). */
size_t caspr_cnf_param_count(int hidden, int ctx_dim);
size_t caspr_cnf_adjoint_workspace_bytes(int frames, int pts, int hidden, int ctx_dim);
int caspr_cnf_adjoint(const float* x1, const float* logp1, const float* gx1, const float* glogp1,
                      const float* e, const float* ctx, int frames, int pts,
                      const caspr_cnf_weights* w, float end_time, float rtol, float atol, int engine,
                      float* gx0, float* glogp0, float* gctx, float* gparams, float* gtimes,
                      int32_t* info, int32_t* h_info, void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------- latent ODE training (adjoint backward)
 * Replaces OdeintAdjointMethod.backward for latent_ode_model.py:98 and the VJP through DynamicsNet
 * (latent_ode_model.py:129-147).  zs (nT,B,D): the forward solution (caspr_latent_ode_solve's `out`);
 * gzs (nT,B,D): dL/d(zs); h_times as in the forward call.  Returns gz0 (B,D) = dL/d(z0) and gparams:
 * caspr_latent_ode_param_count(D,H) floats in DynamicsNet.parameters() order (W0,b0,W1,b1,W2,b2,W3,b3).
 * Scalar tolerances: every augmented tensor (z, adj_z, adj_params) is under step control, one dopri5 solve per
 * output interval, last to first.  Synchronises `stream` once per attempted step. */
size_t caspr_latent_ode_param_count(int D, int H);
size_t caspr_latent_ode_adjoint_workspace_bytes(int B, int D, int H);
int caspr_latent_ode_adjoint(const float* zs, const float* gzs, int B, int D, int H,
                             const float* W0, const float* b0, const float* W1, const float* b1,
                             const float* W2, const float* b2, const float* W3, const float* b3,
                             const double* h_times, int nT, float rtol, float atol,
                             float* gz0, float* gparams, int32_t* info, int32_t* h_info,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------- encoder training operators (forward + backward)
 * What torch autograd does in the reference for the Conv1d(k=1) / GroupNorm / ReLU / max chains
 * (pointnet2.py:637-642,677-699,471-481,207-212; pointnet.py:27-46; tpointnet2.py:59-62,99-112) and for Kaolin's
 * group-gather / three_interpolate Functions (pointnet2.py:391,519).  Rows x channels fp32, leading dimensions. */

/* GroupNorm statistics kept for the backward: mean_rstd [samples][groups] (mean, 1/sqrt(var+eps)) pairs. */
size_t caspr_gn_workspace_bytes(int samples, int rows_per_sample, int C);
int caspr_gn_moments(const float* X, int ldx, int samples, int rows_per_sample, int C, int groups, float eps,
                   float* mean_rstd, void* workspace, size_t workspace_bytes, void* stream);
/* Y = ReLU?((X - mean) * rstd * gamma + beta), out of place (X is kept for the backward). */
int caspr_gn_apply(const float* X, int ldx, const float* mean_rstd, int samples, int rows_per_sample, int C,
                   int groups, const float* gamma, const float* beta, int relu, float* Y, int ldy, void* stream);
/* max over the rows of each sample and the (first) row index attaining it: maxout (samples,C), argmax (samples,C). */
size_t caspr_rowmax_workspace_bytes(int samples, int rows_per_sample, int C);
int caspr_rowmax(const float* Y, int ldy, int samples, int rows_per_sample, int C, float* maxout, int ld_max,
                 int32_t* argmax, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of [GroupNorm -> ReLU? -> (dense consumer and/or max-pool)]: the output cotangent is
 * dY (rows,C; may be NULL) + dMax (samples,C; may be NULL) routed to argmax rows.  Writes dX, dgamma, dbeta. */
int caspr_gn_backward(const float* dY, int lddy, const float* dMax, int ld_dmax, const int32_t* argmax,
                      const float* X, int ldx, const float* mean_rstd, int samples, int rows_per_sample,
                      int C, int groups, const float* gamma, const float* beta, int relu, float* dX,
                      int lddx, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                      void* stream);
/* Conv1d(k=1)/Linear parameter gradients: dW (Cout,Cin) = dY^T . act(X), db (Cout) = column sums of dY (db may be
 * NULL); relu_x applies ReLU to X on the fly (tpointnet2.py:105).  Deterministic split-row reduction. */
size_t caspr_linear_wgrad_workspace_bytes(long long rows, int Cout, int Cin);
int caspr_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, long long rows, int Cout, int Cin,
                       int relu_x, float* dW, float* db, void* workspace, size_t workspace_bytes, void* stream);
/* The same dW on the tensor cores: both operands are transposed into fp16 hi/lo planes with the ROW index as the
 * contraction dimension (per-channel power-of-two scales), split-K tcgen05 fp16x3 GEMM (at most 4096 rows per
 * split), ordered reduction of the partials.  Meant for Cout, Cin >= 64; workspace 1024-byte aligned.
 * dy_colmax / x_colmax: optional per-channel max |.| of the operands as float bit patterns (a producer kernel can
 * maintain them with atomicMax); NULL = computed here with one extra pass over the operand. */
size_t caspr_linear_wgrad_tc_workspace_bytes(long long rows, int Cout, int Cin);
int caspr_linear_wgrad_tc(const float* dY, int lddy, const float* X, int ldx, long long rows, int Cout, int Cin,
                          int relu_x, float* dW, const uint32_t* dy_colmax, const uint32_t* x_colmax,
                          void* workspace, size_t workspace_bytes, void* stream);
/* out[c] (+)= sum_r X[r][c] (backward of the repeat at pointnet.py:44). */
size_t caspr_colsum_workspace_bytes(long long rows, int C);
int caspr_colsum(const float* X, int ldx, long long rows, int C, float* out, int accumulate, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Backward of caspr_group_points wrt the features: dfeat[b][idx][c] += dOut[row][3+c] (atomic adds). */
int caspr_group_points_bwd(const float* dOut, int ld_out, const int32_t* idx, int B, int N, int M, int C,
                           int ns, float* dfeat, int ld_feat, void* stream);
/* Backward of the interpolation part of caspr_three_interp_concat: dprev[b][idx_k][c] += w_k dOut[row][c]. */
int caspr_three_interp_bwd(const float* dOut, int ld_out, const int32_t* idx, const float* dist, int B, int n,
                           int m, int Cp, float* dprev, int ld_prev, void* stream);
/* dst (+)= src, optionally only where relu_ref > 0 (gradient through a ReLU whose output is relu_ref). */
int caspr_rows_update(const float* src, int ld_src, long long rows, int C, int accumulate, const float* relu_ref,
                      int ld_ref, float* dst, int ld_dst, void* stream);
/* dst (cols,rows) = src (rows,cols)^T, contiguous. */
int caspr_transpose(const float* src, int rows, int cols, float* dst, void* stream);

/* ------------------------------------------------------------ input pipeline -> device (next row)
 * Batch assembly of data/caspr_dataset.py:148-208 (load_seq_path) and :288-325 (DynamicPCLDataset.__getitem__) for
 * frames the host has already decoded: nocs / depth are the raw float64 points of all frames of all B sequences
 * back to back ((total,3) each; depth = nocs where a frame has no depth data, :174-176), frame_off (B*Tfull+1) the
 * point offsets of the frames, n_valid[b] the number of frames before the first blank one (:183-186; later frames
 * stay zero).  steps (B,T): chosen time steps (sorted, as :296); pts (B,Tp,N), Tp = 1 or T: chosen point indices into
 * the frame padded to expected_num_pts by cycling its points (:188-195).  Outputs (B,T,N,4) float32:
 * input_out = [depth xyz | max_timestamp*step/(Tfull-1)], output_out = [nocs xyz | step/(Tfull-1)], time stamps formed
 * in float64 like numpy, optionally shifted so each item starts at 0 (:319-322). */
int caspr_assemble_batch(const double* nocs, const double* depth, const long long* frame_off,
                         const int32_t* n_valid, int B, int Tfull, int expected_num_pts,
                         const int32_t* steps, int T, const int32_t* pts, int Tp, int N,
                         double max_timestamp, int shift_time_to_zero, float* input_out,
                         float* output_out, void* stream);

/* -------------------------------------------------------------------- metric
 * Symmetric squared-NN Chamfer distance (reference utils/evaluations.py:40-43 via
 * tk3dv ChamferDistance): a (B,P,3), b (B,Q,3) -> d_ab (B,P) min sq dist a->b, d_ba (B,Q). */
int caspr_chamfer(const float* a, const float* b, int B, int P, int Q,
                  float* d_ab, float* d_ba, void* stream);

/* T-NOCS regression error (utils/evaluations.py:243-254, test_tnocs_regression): pred, gt (frames,N,4) rows [x,y,z,t]
 * -> space[frame] = mean_i |pred_xyz - gt_xyz|_2, time_err[frame] = mean_i |pred_t - gt_t|. */
int caspr_tnocs_error(const float* pred, const float* gt, int frames, int N, float* space, float* time_err,
                      void* stream);

/* Approximate earth mover's distance (reference utils/emd.py:11-12 -> emd_cuda approxmatch_forward + matchcost_forward,
 * evaluations.py:45-46): xyz1 (B,n,3), xyz2 (B,m,3) -> cost (B) = sum of match(k,l) |p_k - q_l| after the ten
 * annealing levels of the published approxmatch algorithm (oracle/emd_oracle.py).  The match matrix is never stored. */
size_t caspr_emd_workspace_bytes(int B, int n, int m);
int caspr_emd(const float* xyz1, const float* xyz2, int B, int n, int m, float* cost, void* workspace,
              size_t workspace_bytes, void* stream);

/* Profiling aid for the fused CNF evaluation kernel (CASPR_CNF_FUSED_DEBUG=1): per-CTA cycle counters of its TMA
 * producer and MMA threads for the last launch, 24 counters per CTA (see csrc/cnf.cu); synchronises the device. */
int caspr_cnf_fused_debug_read(long long* out, int count);

/* Correspondence-RANSAC rigid pose (reference utils/evaluations.py:360-380: open3d
 * registration_ransac_based_on_correspondence with identity correspondences, ransac_n = 4, threshold 0.015,
 * RANSACConvergenceCriteria(50000, 5000), TransformationEstimationPointToPoint(False)).  open3d is absent from the
 * reference tree and unpinned: this restates its published algorithm (oracle/ransac_oracle.py).
 * src, dst (frames,N,3): corresponding points (src = predicted NOCS - 0.5, dst = observed points); samples
 * (frames,H,4) int32: the four correspondences of every hypothesis (open3d draws them with C rand(); here the caller
 * supplies them).  Per frame: every hypothesis is fitted (least-squares rigid transform dst ~ R src + t) and scored
 * (inliers = correspondences with |R src + t - dst| < max_distance); the best one (most inliers, then lowest inlier
 * RMSE, then lowest index) is returned: R_out (frames,9) row-major, t_out (frames,3), best_out (frames) hypothesis
 * index, fitness_out = inliers / N, rmse_out; counts_out (frames,H) inlier counts of all hypotheses or NULL.
 * refine != 0 refits the transform on the winner's inliers (newer open3d versions). */
size_t caspr_ransac_pose_workspace_bytes(int frames, int hypotheses);
int caspr_ransac_pose(const float* src, const float* dst, const int32_t* samples, int frames, int N, int hypotheses,
                      float max_distance, int refine, float* R_out, float* t_out, int32_t* best_out,
                      float* fitness_out, float* rmse_out, int32_t* counts_out, void* workspace,
                      size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* CASPR_B200_H_ */
