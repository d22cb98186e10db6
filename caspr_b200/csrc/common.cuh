// Shared helpers for libcaspr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/caspr_b200.h"

#define CASPR_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return CASPR_ELAUNCH;              \
  } while (0)

#define CASPR_REQUIRE(cond)             \
  do {                                  \
    if (!(cond)) return CASPR_EINVAL;   \
  } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
extern unsigned long long g_caspr_launches;
#define CASPR_COUNT() (++g_caspr_launches)

// optional CUDA-event timing of selected kernels (caspr_profile_* in misc.cu)
void caspr_prof_begin(int kernel_id, cudaStream_t s);
void caspr_prof_end(int kernel_id, cudaStream_t s);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Canonical squared distance of the oracle (oracle/pointnet2_ops.py::_sqdist_f32):
// ((dx*dx)+(dy*dy))+(dz*dz), every operation rounded separately — the *_rn intrinsics are
// never contracted into FMAs by nvcc, which is what makes the indices bit-exact.
__device__ __forceinline__ float sqdist_canonical(float ax, float ay, float az,
                                                  float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Running maximum of non-negative floats kept as their bit patterns (unsigned order == float order).  The value only
// grows, so a plain read that already shows a larger maximum makes the atomic unnecessary: almost all of them are.
__device__ __forceinline__ void atomic_max_nonneg(unsigned* p, float v) {
  const unsigned b = __float_as_uint(v);
  if (v > 0.f && b > __ldcg(p)) atomicMax(p, b);
}

// Order-preserving float <-> uint mapping for atomicMax on floats.
__device__ __forceinline__ unsigned float_to_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
