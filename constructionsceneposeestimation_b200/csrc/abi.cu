// Library-level entry points of libcspe.so: version, thread-local error string, device info.
#include <stdarg.h>
#include <string.h>

#include "cspe_common.cuh"

namespace cspe {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// SM count of the current device (immutable attribute; cached per device ordinal).
int sm_count() {
  static int cache[64] = {0};
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return -1;
  if (dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (dev < 64) cache[dev] = n;
  return n;
}

}  // namespace cspe

extern "C" int cspe_version(void) { return CSPE_ABI_VERSION; }

extern "C" const char* cspe_last_error(void) { return cspe::g_error; }

extern "C" int cspe_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0) {
    cspe::set_error("cspe_device_info: no current CUDA device (%s)", cudaGetErrorString(e));
    return CSPE_ERR_NO_DEVICE;
  }
  int n = 0, maj = 0, min = 0;
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return CSPE_OK;
}
