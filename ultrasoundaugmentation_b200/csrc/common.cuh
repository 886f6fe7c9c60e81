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

// ---- bounds-checked debug build (-DUSAUG_BOUNDS_CHECK -> lib/libusaug_dbg.so) ---------------------------------------
// compute-sanitizer is not available on the GPU pool, so the library carries its own checks:
//   * every sub-buffer carved out of a workspace is followed by a 256-byte red zone, filled with a canary before the
//     kernels of a call run and verified after them (the debug build synchronises at the end of a call; the product build
//     never does): a kernel that writes past one of its planes turns the call into USAUG_E_BOUNDS;
//   * USAUG_DEV_ASSERT(cond) guards the index computations of the kernels (row offsets, ring slots, queue slots,
//     shared-memory tiles, gather rows): a violated one prints file:line and traps, which surfaces as a CUDA error.
// Both are compiled out of the product library.
#ifdef USAUG_BOUNDS_CHECK
constexpr size_t kRedZone = 256;
#define USAUG_DEV_ASSERT(cond)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      printf("usaug bounds check failed: %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
      __trap();                                                                             \
    }                                                                                       \
  } while (0)
#else
constexpr size_t kRedZone = 0;
#define USAUG_DEV_ASSERT(cond) do { } while (0)
#endif
constexpr unsigned char kCanary = 0xA5;

// Carves 256-byte aligned sub-buffers out of the caller's workspace.
struct Carver {
  char* base;
  size_t off;
  size_t zones[64];     // debug build: offsets of the red zones
  int nzones;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0), nzones(0) {}
  template <typename T>
  T* take(size_t n) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T));
    if (kRedZone && n) {
      if (nzones < 64) zones[nzones++] = off;
      off += kRedZone;
    }
    return r;
  }
  // debug build: fill the red zones (before the first kernel of the call) ...
  int arm(cudaStream_t stream) {
    for (int i = 0; i < nzones && base; ++i) USAUG_CUDA(cudaMemsetAsync(base + zones[i], kCanary, kRedZone, stream));
    return USAUG_OK;
  }
  // ... and check them after the last one (synchronises; debug build only)
  int verify(cudaStream_t stream, const char* what) {
    if (!kRedZone || !base) return USAUG_OK;
    USAUG_CUDA(cudaStreamSynchronize(stream));
    unsigned char host[256];
    for (int i = 0; i < nzones; ++i) {
      USAUG_CUDA(cudaMemcpy(host, base + zones[i], kRedZone, cudaMemcpyDeviceToHost));
      for (size_t k = 0; k < kRedZone; ++k)
        if (host[k] != kCanary) {
          set_error("%s: red zone %d of the workspace (offset %zu) was overwritten at byte %zu", what, i, zones[i], k);
          return USAUG_E_BOUNDS;
        }
    }
    return USAUG_OK;
  }
};

int device_sm_count();

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A short chain of small kernels (scan -> profile -> warp) pays a launch gap at every boundary.  A kernel
// launched with launch_pdl may start while its predecessor on the stream drains: every block of the
// predecessor calls pdl_trigger() when it starts, the dependent kernel runs whatever does not need the
// predecessor's results and then pdl_wait()s -- which returns once the predecessor grid has completed and
// its memory is visible.  Kernels launched normally treat both calls as no-ops.
// RULE (measured on B200, CUDA 12.9 / driver 580): only use this for a kernel that is followed, inside the same
// C-ABI call, by a NORMALLY launched kernel.  A small host-to-device copy of pageable memory enqueued after a
// PDL-launched kernel was observed to land before that kernel had finished (torch hands the parameter
// buffers of one call to the next, so late blocks read the next call's parameters): the last kernel of every
// call is therefore launched normally, which restores plain stream order for whatever the caller does next.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
