// common.cuh -- shared host/device helpers for libpvb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/pvb200.h"

namespace pvb {

// ---- error state (per calling thread) -----------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();
int sm_count();

#define PVB_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      pvb::set_error(__VA_ARGS__);             \
      return PVB200_ERR_INVALID;               \
    }                                          \
  } while (0)

#define PVB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      pvb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PVB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors
#define PVB_LAUNCHED(name)                                                               \
  do {                                                                                   \
    pvb::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      pvb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));           \
      return PVB200_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

static inline cudaStream_t as_stream(pvb200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

// ---- device helpers -------------------------------------------------------------------------------
// The reference normalisation (netcdf_dataset.py:96-101): two separately rounded fp32 ops.  The
// intrinsics forbid FMA contraction / reciprocal substitution, so the result is bit-identical.
__device__ __forceinline__ float sat_norm(int16_t v, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(static_cast<float>(v), mean), stdv);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pvb
