// common.cuh -- shared helpers for the sm_100a kernels of libupsnet_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/upsnet_b200.h"

#define UPS_CHECK_LAUNCH()                         \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

#define UPS_CUDA(call)                             \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace ups {

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline int conv_out_size(int in, int pad, int dil, int k, int stride) {
  return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

}  // namespace ups
