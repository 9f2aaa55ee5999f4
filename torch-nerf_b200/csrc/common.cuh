// Shared helpers for libnerf_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nerf_b200.h"

namespace nerf {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(nerf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define NERF_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::nerf::set_error(__VA_ARGS__);    \
      return NERF_ERR_ARG;               \
    }                                    \
  } while (0)

#define NERF_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::nerf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return NERF_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define NERF_LAUNCH_CHECK() NERF_CUDA(cudaGetLastError())

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// number of SMs of the current device (cached)
int sm_count();

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nerf
