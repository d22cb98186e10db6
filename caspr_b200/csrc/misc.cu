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

// ------------------------------------------------------------------ launch count / profiling
unsigned long long g_caspr_launches = 0;
extern "C" unsigned long long caspr_launch_count(void) { return g_caspr_launches; }
extern "C" void caspr_launch_count_add(unsigned long long n) { g_caspr_launches += n; }

namespace {
constexpr int kProfKernels = 8;
constexpr int kProfMaxPairs = 16384;
struct ProfSlot {
  cudaEvent_t beg[kProfMaxPairs], end[kProfMaxPairs];
  int created = 0, used = 0;
  long long dropped = 0;
};
ProfSlot* g_prof[kProfKernels] = {nullptr};
int g_prof_on = 0;
}  // namespace

void caspr_prof_begin(int id, cudaStream_t s) {
  if (!g_prof_on || id < 0 || id >= kProfKernels) return;
  if (!g_prof[id]) g_prof[id] = new ProfSlot();
  ProfSlot* p = g_prof[id];
  if (p->used >= kProfMaxPairs) { p->dropped++; return; }
  if (p->used >= p->created) {
    cudaEventCreate(&p->beg[p->created]);
    cudaEventCreate(&p->end[p->created]);
    p->created++;
  }
  cudaEventRecord(p->beg[p->used], s);
}
void caspr_prof_end(int id, cudaStream_t s) {
  if (!g_prof_on || id < 0 || id >= kProfKernels || !g_prof[id]) return;
  ProfSlot* p = g_prof[id];
  if (p->used >= kProfMaxPairs) return;
  cudaEventRecord(p->end[p->used], s);
  p->used++;
}

extern "C" void caspr_profile_enable(int on) {
  g_prof_on = on ? 1 : 0;
  if (on)
    for (int i = 0; i < kProfKernels; ++i)
      if (g_prof[i]) { g_prof[i]->used = 0; g_prof[i]->dropped = 0; }
}

extern "C" int caspr_profile_read(int kernel_id, double* total_ms, long long* launches) {
  CASPR_REQUIRE(kernel_id >= 0 && kernel_id < kProfKernels && total_ms && launches);
  *total_ms = 0.0;
  *launches = 0;
  ProfSlot* p = g_prof[kernel_id];
  if (!p) return CASPR_OK;
  for (int i = 0; i < p->used; ++i) {
    if (cudaEventSynchronize(p->end[i]) != cudaSuccess) return CASPR_ELAUNCH;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->beg[i], p->end[i]) != cudaSuccess) return CASPR_ELAUNCH;
    *total_ms += (double)ms;
  }
  *launches = p->used;
  return CASPR_OK;
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
  if (d_ab) { CASPR_COUNT(); chamfer_dir_kernel<<<dim3(ceil_div(P, 256), B), 256, 0, s>>>(a, b, P, Q, d_ab); }
  if (d_ba) { CASPR_COUNT(); chamfer_dir_kernel<<<dim3(ceil_div(Q, 256), B), 256, 0, s>>>(b, a, Q, P, d_ba); }
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}

// ------------------------------------------------------------------ T-NOCS regression error (next row 8f.4, first half)
// utils/evaluations.py:243-254 (test_tnocs_regression): per frame, mean over the points of the L2 distance between the
// predicted and the ground-truth NOCS position, and of the absolute time difference.  One CTA per frame.
namespace {
__global__ void __launch_bounds__(256)
tnocs_error_kernel(const float4* __restrict__ pred, const float4* __restrict__ gt, int N, float* __restrict__ space,
                   float* __restrict__ time_err) {
  __shared__ float s_a[8], s_b[8];
  const size_t base = (size_t)blockIdx.x * N;
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float4 p = pred[base + i], g = gt[base + i];
    const float dx = p.x - g.x, dy = p.y - g.y, dz = p.z - g.z;
    a += sqrtf(dx * dx + dy * dy + dz * dz);
    b += fabsf(p.w - g.w);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += s_a[w]; tb += s_b[w]; }
    space[blockIdx.x] = ta / (float)N;
    time_err[blockIdx.x] = tb / (float)N;
  }
}
}  // namespace

extern "C" int caspr_tnocs_error(const float* pred, const float* gt, int frames, int N, float* space, float* time_err,
                                 void* stream) {
  CASPR_REQUIRE(pred && gt && space && time_err && frames > 0 && N > 0);
  CASPR_REQUIRE((((uintptr_t)pred | (uintptr_t)gt) & 15) == 0);
  CASPR_COUNT(); tnocs_error_kernel<<<frames, 256, 0, (cudaStream_t)stream>>>((const float4*)pred, (const float4*)gt, N, space, time_err);
  CASPR_CHECK_LAUNCH();
  return CASPR_OK;
}
