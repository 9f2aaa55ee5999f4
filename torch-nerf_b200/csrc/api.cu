// Library-level entry points: version, error string, device info.
#include <stdarg.h>

#include "common.cuh"

namespace nerf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // cached per device: a process may drive several GPUs (the value is written once per device and is idempotent,
  // so concurrent first calls are harmless)
  constexpr int kMaxDev = 64;
  static int cached[kMaxDev] = {};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < kMaxDev && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < kMaxDev) cached[dev] = n;
  return n;
}

}  // namespace nerf

extern "C" {

int nerf_version(void) { return 100; }

const char* nerf_last_error(void) { return nerf::g_err; }

int nerf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  NERF_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  NERF_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  NERF_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  NERF_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return NERF_OK;
}

}  // extern "C"
