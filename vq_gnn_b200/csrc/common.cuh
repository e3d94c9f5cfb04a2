// Shared helpers for libvqgnn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqgnn.h"

namespace vqgnn {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define VQ_CHECK_ARG(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      ::vqgnn::set_error(__VA_ARGS__);           \
      return VQGNN_ERR_ARG;                      \
    }                                            \
  } while (0)

#define VQ_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::vqgnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return VQGNN_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define VQ_LAUNCH_CHECK()                 \
  do {                                    \
    VQ_CUDA(cudaPeekAtLastError());       \
    ::vqgnn::count_launch();              \
  } while (0)

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
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// float atomic max through the ordered-int trick (works for any sign, no NaNs expected)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace vqgnn
