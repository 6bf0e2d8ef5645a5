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

// Two-rank device-side barrier on the compute stream (tensor parallelism over peer memory, distributed/peer.py).
// One thread: epoch = ++(*counter) (local), publish it into the peer's flag, wait until the peer's epoch shows up in
// ours.  Launched with programmatic dependent launch: it lets its successor start its private prologue at once
// and itself waits for its predecessor's memory (the tiles / rows that kernel pushed over NVLink) before it signals,
// so no kernel boundary drains the SMs the way a stream-ordered barrier kernel of another library does.
}  // extern "C" (kernel below)
namespace npt {
__global__ void __launch_bounds__(32) peer_barrier_kernel(unsigned long long* peer_flag, volatile unsigned long long* own_flag,
                                                           unsigned long long* counter) {
  pdl_trigger();
  pdl_wait();   // predecessor complete, its writes (local and peer) performed
  if (threadIdx.x == 0) {
    const unsigned long long epoch = *counter + 1;
    *counter = epoch;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(epoch) : "memory");
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(own_flag) : "memory");
    } while (seen < epoch);
    __threadfence_system();
  }
}
}  // namespace npt
extern "C" {

int npt_peer_barrier(void* peer_flag, void* own_flag, void* counter, void* stream) {
  NPT_REQUIRE(peer_flag && own_flag && counter, "peer_barrier: null pointer");
  NPT_CHECK_CUDA(npt::launch_pdl(npt::peer_barrier_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (unsigned long long*)peer_flag,
                                 (volatile unsigned long long*)own_flag, (unsigned long long*)counter));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
