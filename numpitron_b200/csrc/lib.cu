// Library plumbing: error text, version, device query.
#include <stdarg.h>

#include "common.cuh"

namespace npt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}

}  // namespace npt

extern "C" {

const char* npt_last_error(void) { return npt::g_err; }

int npt_version(void) { return 100; }

int npt_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  NPT_CHECK_CUDA(cudaGetDevice(&dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}

}  // extern "C"
