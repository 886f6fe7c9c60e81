// Version / error plumbing of the C ABI (include/usaug.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace usaug {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

}  // namespace usaug

extern "C" int usaug_version(void) { return USAUG_VERSION; }
extern "C" const char* usaug_last_error(void) { return usaug::g_err; }

extern "C" int usaug_upload(const void* host, void* dev, size_t bytes, void* stream) {
  if (bytes == 0) return USAUG_OK;
  if (!host || !dev) {
    usaug::set_error("usaug_upload: null pointer");
    return USAUG_E_INVALID_ARG;
  }
  USAUG_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return USAUG_OK;
}
