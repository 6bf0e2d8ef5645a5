// Library plumbing: error text, version, device query.
#include <stdlib.h>
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

static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }

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

static int g_sm_margin = 0;
int sm_margin() { return g_sm_margin; }

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("NUMPITRON_PDL");
    cached = (e && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

}  // namespace npt

extern "C" {

// SMs that persistent GEMM grids leave free (for a collective that is in flight on another stream).
int npt_set_sm_margin(int32_t sms) {
  NPT_REQUIRE(sms >= 0 && sms < 128, "set_sm_margin: out of range");
  npt::g_sm_margin = sms;
  return 0;
}

const char* npt_last_error(void) { return npt::g_err; }

int npt_version(void) { return 100; }

unsigned long long npt_launch_count(void) { return npt::g_launches; }

int npt_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  NPT_CHECK_CUDA(cudaGetDevice(&dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
  NPT_CHECK_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}


// Pinned host staging + async copies, owned by the caller.  Kept out of PyTorch's caching host
// allocator on purpose: the copies are issued inside CUDA-graph capture.
int npt_pinned_alloc(void** out, size_t bytes) {
  NPT_REQUIRE(out != nullptr && bytes > 0, "pinned_alloc: bad arguments");
  NPT_CHECK_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 0;
}

int npt_pinned_free(void* p) {
  if (p) NPT_CHECK_CUDA(cudaFreeHost(p));
  return 0;
}

int npt_memcpy_async(void* dst, const void* src, size_t bytes, int32_t kind, void* stream) {
  NPT_REQUIRE(kind == 1 || kind == 2 || kind == 3, "memcpy_async: kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
  if (bytes == 0) return 0;
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  NPT_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream));
  return 0;
}

int npt_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                       void* stream) {
  if (width_bytes == 0 || rows == 0) return 0;
  NPT_REQUIRE(dst_pitch >= width_bytes && src_pitch >= width_bytes, "memcpy2d_async: pitch smaller than the row width");
  NPT_CHECK_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
