// Library identification and the Chamfer-distance metric kernel.
#include "common.cuh"

extern "C" int caspr_version(void) { return 100; }
extern "C" const char* caspr_build_arch(void) { return "sm_100a"; }
extern "C" const char* caspr_status_string(int status) {
  switch (status) {
    case CASPR_OK: return "ok";
    case CASPR_EINVAL: return "invalid argument (shape / null pointer / unsupported size)";
    case CASPR_ELAUNCH: return "CUDA launch or runtime error";
    case CASPR_EWORKSPACE: return "workspace too small";
    case CASPR_ESOLVER_DT: return "underflow in dt";
    case CASPR_ESOLVER_NONFINITE: return "non-finite values in state `y`";
    case CASPR_ESOLVER_MAXSTEPS: return "max_num_steps exceeded";
    case CASPR_ERANGE: return "split-precision operand left the fp16 range";
    default: return "unknown status";
  }
}

namespace {

// Squared nearest-neighbour distance from every point of `a` to the cloud `b` (one direction of
// the Chamfer distance, reference utils/evaluations.py:40-43 via tk3dv's ChamferDistance).
// grid (ceil(P/256), B); the target cloud streams through shared memory in tiles of 1024 points.
constexpr int kChTile = 1024;
__global__ void __launch_bounds__(256)
chamfer_dir_kernel(const float* __restrict__ a, const float* __restrict__ b, int P, int Q,
                   float* __restrict__ d) {
  __shared__ float sx[kChTile], sy[kChTile], sz[kChTile];
  const int batch = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float* pa = a + ((size_t)batch * P + (i < P ? i : 0)) * 3;
  const float ax = pa[0], ay = pa[1], az = pa[2];
  float best = 3.0e38f;
  for (int q0 = 0; q0 < Q; q0 += kChTile) {
    const int cnt = min(kChTile, Q - q0);
    __syncthreads();
    const float* pb = b + ((size_t)batch * Q + q0) * 3;
    for (int t = threadIdx.x; t < cnt * 3; t += 256) {
      float v = pb[t];
      int k = t / 3, c = t - 3 * k;
      (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
    }
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      float dx = ax - sx[k], dy = ay - sy[k], dz = az - sz[k];
      best = fminf(best, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
  }
  if (i < P) d[(size_t)batch * P + i] = best;
}

}  // namespace

extern "C" int caspr_chamfer(const float* a, const float* b, int B, int P, int Q, float* d_ab,
                             float* d_ba, void* stream) {
  CASPR_REQUIRE(a && b && B > 0 && P > 0 && Q > 0 && (d_ab || d_ba));
  cudaStream_t s = (cudaStream_t)stream;
  if (d_ab) chamfer_dir_kernel<<<dim3(ceil_div(P, 256), B), 256, 0, s>>>(a, b, P, Q, d_ab);
  if (d_ba) chamfer_dir_kernel<<<dim3(ceil_div(Q, 256), B), 256, 0, s>>>(b, a, Q, P, d_ba);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
