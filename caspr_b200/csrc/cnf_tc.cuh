// Interface between the CNF solver (cnf.cu) and its tensor-core engine (cnf_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include "cnf_state.cuh"

namespace cnf_tc {

constexpr float kActScale = 16.f;     // activations / tangents are stored as fp16(x * 2^4)

struct Weights {
  __half* hi[2];       // W1, W2 split planes [512][512], scaled by scales[2*l]
  __half* lo[2];
  float* scales;       // [l] = {w_scale, 1/(act_scale*w_scale)}
  unsigned* max_bits;  // scratch: max|W| bits per layer
};

struct Plan {
  CUtensorMap tm_act[2][2];   // [buffer A/B][hi/lo]
  CUtensorMap tm_w[2][2];     // [layer][hi/lo]
  __half *a_hi, *a_lo, *b_hi, *b_lo;
  int n_tiles;
  // fused evaluation kernel: per-CTA scratch planes (SA, SB: one 128-row tile per CTA) and their tensor maps
  CUtensorMap tm_sa[2], tm_sb[2];     // [hi/lo]
  __half *sa_hi, *sa_lo, *sb_hi, *sb_lo;
  int fused_grid;                     // CTAs of the fused kernel (0: fused path not available)
};

// number of CTAs the fused kernel uses for n points on a device with num_sms SMs (even: CTA pairs)
inline int fused_grid_for(int n, int num_sms) {
  const int pairs = ((n + 63) / 64 + 1) / 2;
  const int clusters = pairs < num_sms / 2 ? pairs : num_sms / 2;
  return 2 * clusters;
}
// bytes of scratch the fused kernel needs for `grid` CTAs: SA (hi, lo) + SB (hi, lo), one [128][512] fp16 tile each
inline size_t fused_scratch_bytes(int grid) { return (size_t)grid * 128 * 512 * 2 * 4; }

size_t weights_workspace_bytes();
int prepare_weights(const float* W1, const float* W2, const Weights& out, cudaStream_t s);
int fill_col_scale(const Weights& w, int ctot, float* col_scale, cudaStream_t s);
int make_plan(Plan& plan, const Weights& w, __half* a_hi, __half* a_lo, __half* b_hi, __half* b_lo, int n,
              void* fused_scratch = nullptr, int fused_grid = 0);
// whether the fused single-launch evaluation is used (CASPR_CNF_FUSED=0 selects the four-kernel path)
bool fused_enabled();
long long* fused_debug_buffer();
// One dynamics evaluation (all four layers) for ALL points into kout, one launch.
int enqueue_fused(const Plan& plan, const float4* y0, const float4* kbuf, size_t kstride, const float* e,
                  const float* W0, const float* W3, int n, int P, int stage, int reverse, const float* gate,
                  const float* biasf, int ld_hyper, const CnfState* st, float4* kout, int* range_flag,
                  cudaStream_t s);
// The three stages below work on the point range [pt0, n_end) (tile aligned: pt0 % 64 == 0) resp. the row tiles
// [m_tile0, m_tile0 + m_tiles), so that two halves of the point set can be pipelined on two streams.
int enqueue_layer0(const Plan& plan, const float4* y0, const float4* kbuf, size_t kstride, const float* e,
                   const float* W0, int pt0, int n_end, int P, int stage, const float* gate, const float* biasf,
                   int ld_hyper, const CnfState* st, int* range_flag, cudaStream_t s);
// layer 0: planes A -> planes B;  layer 1: planes B -> fused output layer (W3 (3,512)) accumulated into acc6 [n][8]
int enqueue_mid(const Plan& plan, int layer, int m_tile0, int m_tiles, const float* gate, const float* biasf,
                int ld_hyper, int n, int P, const CnfState* st, const float* W3, float* acc6, int* range_flag,
                int num_sms, cudaStream_t s);
int enqueue_last_finish(float* acc6, const float* e, int pt0, int n_end, int P, const float* gate, const float* biasf,
                        int ld_hyper, int reverse, const CnfState* st, float4* kout, cudaStream_t s);

}  // namespace cnf_tc
