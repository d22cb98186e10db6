// Input pipeline -> device: the per-item array work of the reference's loader done on the GPU
// for a whole batch at once.  Replaces, for already-decoded frames, caspr/data/caspr_dataset.py:148-208
// (load_seq_path: pad short frames by cycling their points, stop at a blank frame, append NOCS / world time stamps)
// and :288-325 (DynamicPCLDataset.__getitem__: time-step and point sub-sampling, shift_time_to_zero, float32 cast).
// The host only decodes the .npz files and uploads the raw float64 points once (pinned, ragged); HBM-bound gather.
#include "common.cuh"

namespace {

// one thread per output point (b, t, i): gathers 3 coordinates from each of the two clouds, forms both time stamps in
// float64 exactly as numpy does, casts to float32.
__global__ void __launch_bounds__(256)
assemble_batch_kernel(const double* __restrict__ nocs, const double* __restrict__ depth,
                      const long long* __restrict__ frame_off, const int* __restrict__ n_valid, int B, int Tfull,
                      int expected_num_pts, const int* __restrict__ steps, int T, const int* __restrict__ pts, int Tp,
                      int N, double max_timestamp, int shift_time_to_zero, float* __restrict__ input_out,
                      float* __restrict__ output_out) {
  const long long total = (long long)B * T * N;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(o % N);
    const int t = (int)((o / N) % T);
    const int b = (int)(o / ((long long)N * T));
    const int step = steps[b * T + t];
    const int p = pts[((long long)b * Tp + (Tp == 1 ? 0 : t)) * N + i];        // index into the padded frame
    const double step_size = Tfull == 1 ? 0.0 : 1.0 / (double)(Tfull - 1);     // caspr_dataset.py:155-158
    const int valid = n_valid[b];                                              // frames before the first blank one
    float in4[4] = {0.f, 0.f, 0.f, 0.f}, out4[4] = {0.f, 0.f, 0.f, 0.f};
    double t_nocs = 0.0, t_world = 0.0;
    if (step < valid) {
      const long long f0 = frame_off[(long long)b * Tfull + step];
      const long long cnt = frame_off[(long long)b * Tfull + step + 1] - f0;
      // frames longer than expected_num_pts cannot be stored by the reference (shape error); short ones are padded by
      // cycling: padded[p] = frame[p mod cnt]  (:188-195)
      const long long src = f0 + (cnt >= expected_num_pts ? p : p % cnt);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out4[c] = (float)nocs[3 * src + c];
        in4[c] = (float)depth[3 * src + c];
      }
      t_nocs = (1.0 * step_size) * (double)step;                               // :200
      t_world = ((max_timestamp * 1.0) * step_size) * (double)step;            // :204
    }
    if (shift_time_to_zero) {                                                  // :319-322: minus the item's smallest stamp
      double mn_nocs = 1e300, mn_world = 1e300;
      for (int tt = 0; tt < T; ++tt) {
        const int s = steps[b * T + tt];
        const double a = s < valid ? (1.0 * step_size) * (double)s : 0.0;
        const double w = s < valid ? ((max_timestamp * 1.0) * step_size) * (double)s : 0.0;
        mn_nocs = a < mn_nocs ? a : mn_nocs;
        mn_world = w < mn_world ? w : mn_world;
      }
      t_nocs -= mn_nocs;
      t_world -= mn_world;
    }
    out4[3] = (float)t_nocs;
    in4[3] = (float)t_world;
    reinterpret_cast<float4*>(input_out)[o] = make_float4(in4[0], in4[1], in4[2], in4[3]);
    reinterpret_cast<float4*>(output_out)[o] = make_float4(out4[0], out4[1], out4[2], out4[3]);
  }
}

}  // namespace

extern "C" int caspr_assemble_batch(const double* nocs, const double* depth, const long long* frame_off,
                                    const int32_t* n_valid, int B, int Tfull, int expected_num_pts,
                                    const int32_t* steps, int T, const int32_t* pts, int Tp, int N,
                                    double max_timestamp, int shift_time_to_zero, float* input_out,
                                    float* output_out, void* stream) {
  CASPR_REQUIRE(nocs && depth && frame_off && n_valid && steps && pts && input_out && output_out);
  CASPR_REQUIRE(B > 0 && Tfull > 0 && T > 0 && N > 0 && expected_num_pts >= N && (Tp == 1 || Tp == T));
  CASPR_REQUIRE((((uintptr_t)input_out | (uintptr_t)output_out) & 15) == 0);
  const long long total = (long long)B * T * N;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  CASPR_COUNT(); assemble_batch_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      nocs, depth, frame_off, n_valid, B, Tfull, expected_num_pts, steps, T, pts, Tp, N, max_timestamp,
      shift_time_to_zero, input_out, output_out);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
