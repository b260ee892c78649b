// Shared helpers for the corenet_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "corenet_b200.h"

void crn_set_error(const char* fmt, ...);
void crn_count_launches(int n);
// bit 13 of crn_set_flags: single-pass TF32 (only the hi x hi product of the 3xTF32 split is issued)
static inline int crn_single_pass();
int crn_get_flags();              // debug switches, see crn_set_flags   // bumps the process-wide kernel launch counter

#define CRN_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      crn_set_error(__VA_ARGS__);         \
      return CRN_ERR_BAD_ARG;             \
    }                                     \
  } while (0)

#define CRN_LAUNCH_CHECK(name)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      crn_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return CRN_ERR_LAUNCH;                                                \
    }                                                                       \
    crn_count_launches(1);                                                  \
  } while (0)

static inline int crn_single_pass() { return (crn_get_flags() >> 13) & 1; }

static inline cudaStream_t crn_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int64_t crn_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// Double-precision atomic add to global memory (native on sm_60+).
__device__ __forceinline__ void atomic_add_f64(double* p, double v) { atomicAdd(p, v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
