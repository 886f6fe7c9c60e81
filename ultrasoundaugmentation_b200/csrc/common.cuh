// Shared helpers for libusaug.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "usaug.h"

namespace usaug {

void set_error(const char* fmt, ...);

// Launch-error check: never synchronises, only reads the sticky launch status.
#define USAUG_CHECK_LAUNCH(what)                                            \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      ::usaug::set_error("%s: %s", what, cudaGetErrorString(e__));          \
      return USAUG_E_CUDA;                                                  \
    }                                                                       \
  } while (0)

#define USAUG_CUDA(call)                                                    \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      ::usaug::set_error("%s: %s", #call, cudaGetErrorString(e__));         \
      return USAUG_E_CUDA;                                                  \
    }                                                                       \
  } while (0)

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Carves 256-byte aligned sub-buffers out of the caller's workspace.
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T));
    return r;
  }
};

int device_sm_count();

#ifdef __CUDACC__
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Butterfly sum: every lane ends with the same bits (fixed tree => deterministic).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace usaug
