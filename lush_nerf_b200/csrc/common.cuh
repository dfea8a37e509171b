// Shared host/device helpers for the lushnerf_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/lushnerf_b200.h"

namespace lush {

// thread-local error string + global launch counter (abi.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define LUSH_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::lush::set_error(__VA_ARGS__);      \
      return LUSH_EINVAL;                  \
    }                                      \
  } while (0)

#define LUSH_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    ::lush::count_launch();                                       \
    int _rc = ::lush::check_cuda(cudaGetLastError(), name);       \
    if (_rc) return _rc;                                          \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) 128-bit / 32-bit loads that do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// 1/(1+exp(-x)): __frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to the IEEE division 1.f/y
// without the division routine's slow-path branches
__device__ __forceinline__ float sigmoidf_(float x) { return __frcp_rn(1.f + expf(-x)); }

}  // namespace lush
