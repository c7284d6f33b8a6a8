// Shared helpers for libampconv.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ampconv.h"

namespace ampconv {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;   // kernels launched by this library (bench.py's gpu_launches)

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return AMPCONV_ERR_CUDA;
}

#define AMPCONV_CUDA_TRY(expr)                                   \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return ::ampconv::cuda_fail(_e);      \
  } while (0)

#define AMPCONV_CHECK_LAUNCH()                                   \
  do {                                                           \
    ++::ampconv::g_launch_count;                                 \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::ampconv::cuda_fail(_e);      \
  } while (0)

#define AMPCONV_REQUIRE(cond)                                    \
  do {                                                           \
    if (!(cond)) return AMPCONV_ERR_INVALID_ARGUMENT;            \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

}  // namespace ampconv
