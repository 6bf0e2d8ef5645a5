// Shared helpers for the numpitron_b200 kernels (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/numpitron_b200.h"

namespace npt {

void set_error(const char* fmt, ...);

#define NPT_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      npt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define NPT_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      npt::set_error(__VA_ARGS__);            \
      return 2;                               \
    }                                         \
  } while (0)

void count_launch();
#define NPT_LAUNCH_CHECK()                 \
  do {                                     \
    npt::count_launch();                   \
    NPT_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

int sm_count();
int sm_margin();  // SMs persistent GEMMs leave free while a collective overlaps them

// ---- programmatic dependent launch ----------------------------------------------------------------
// The training step is ~200 short kernels in one stream / CUDA graph; each pays its launch latency
// and prologue after the previous kernel has drained.  Kernels launched through launch_pdl() may be
// scheduled while their predecessor is still running: they do their private set-up (barrier init,
// TMEM allocation, index math), then pdl_wait() blocks until EVERY prerequisite grid has completed
// and its writes are visible — nothing before pdl_wait() may read or write global memory.
// pdl_trigger() tells the scheduler this CTA no longer minds the dependent grid being launched (the
// dependent starts once all CTAs have triggered or exited).  NUMPITRON_PDL=0 disables the attribute;
// the device-side calls are then no-ops.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Block-wide reductions for blocks of up to 1024 threads.  `smem` needs 32 floats.
__device__ __forceinline__ float block_sum(float v, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? smem[lane] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? smem[lane] : -INFINITY;
  return warp_max(t);
}

// 128-bit streaming accesses (read-once / write-once data: keep it out of L1).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st_bf16x4(void* p, const float4& v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace npt
