// Fused multi-tensor Adam — replaces optimizer/adam.py:15-45 (adam_update) for every parameter
// of the model in ONE launch.
//
// Bit-exact with the reference's NumPy float32 arithmetic: every product/sum/quotient below is
// a separately rounded IEEE operation (no FMA contraction), sqrt is IEEE, and the six scalars
// arrive already rounded to float32 the way NumPy's scalar promotion rounds them.  The
// reference never advances `timestep`, so c1 = 1-beta1 and c2 = 1-beta2.
// HBM-bound: 28 B per parameter (read p,g,m,v; write p,m,v) + 2 B for the bf16 shadow.
#include "common.cuh"

namespace npt {

constexpr int kAdamThreads = 256;
constexpr int64_t kAdamChunk = 256 * 4 * 8;  // elements per block

struct AdamScalars {
  float lr, b1, omb1, b2, omb2, c1, c2, eps;
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamScalars& s) {
  m = __fadd_rn(__fmul_rn(s.b1, m), __fmul_rn(s.omb1, g));
  v = __fadd_rn(__fmul_rn(s.b2, v), __fmul_rn(s.omb2, __fmul_rn(g, g)));
  const float mh = __fdiv_rn(m, s.c1);
  const float vh = __fdiv_rn(v, s.c2);
  const float upd = __fdiv_rn(__fmul_rn(s.lr, mh), __fadd_rn(__fsqrt_rn(vh), s.eps));
  p = __fsub_rn(p, upd);
  return p;
}

// gradient element(s) i of tensor t: plain, or the sum of its split-K partials in split order — the same
// additions, starting from 0, as splitk_reduce_*_kernel (gemm.cu), so nothing changes bit-wise
__device__ __forceinline__ float4 adam_grad4(const npt_adam_tensor& t, int64_t i4) {
  if (t.g_splits <= 1) return ld_stream(reinterpret_cast<const float4*>(t.g) + i4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* g0 = reinterpret_cast<const float4*>(t.g) + i4;
  const int64_t st4 = t.g_split_stride >> 2;  // the vector path requires a stride that is a multiple of 4
  int64_t s = 0;
  for (; s + 16 <= t.g_splits; s += 16) {  // sixteen independent loads in flight (the LayerNorm partial stacks are ~300 deep)
    float4 a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = ld_stream(g0 + (s + k) * st4);
#pragma unroll
    for (int k = 0; k < 16; ++k) { acc.x += a[k].x; acc.y += a[k].y; acc.z += a[k].z; acc.w += a[k].w; }
  }
  for (; s + 4 <= t.g_splits; s += 4) {  // additions in split order throughout
    float4 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = ld_stream(g0 + (s + k) * st4);
#pragma unroll
    for (int k = 0; k < 4; ++k) { acc.x += a[k].x; acc.y += a[k].y; acc.z += a[k].z; acc.w += a[k].w; }
  }
  for (; s < t.g_splits; ++s) {
    const float4 a = ld_stream(g0 + s * st4);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  return acc;
}
__device__ __forceinline__ float adam_grad1(const npt_adam_tensor& t, int64_t i) {
  if (t.g_splits <= 1) return t.g[i];
  float acc = 0.f;
  for (int64_t s = 0; s < t.g_splits; ++s) acc += t.g[s * t.g_split_stride + i];
  return acc;
}

__global__ void __launch_bounds__(kAdamThreads)
adam_multi_kernel(const npt_adam_tensor* __restrict__ tensors, int n_tensors, AdamScalars s) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  // locate this block's tensor and chunk (n_tensors is small: a linear walk is fine)
  int64_t chunk = blockIdx.x;
  int ti = 0;
  for (; ti < n_tensors; ++ti) {
    const int64_t nchunks = (tensors[ti].n + kAdamChunk - 1) / kAdamChunk;
    if (chunk < nchunks) break;
    chunk -= nchunks;
  }
  if (ti >= n_tensors) return;
  const npt_adam_tensor t = tensors[ti];
  const int64_t lo = chunk * kAdamChunk;
  int64_t hi = lo + kAdamChunk;
  if (hi > t.n) hi = t.n;
  __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(t.p_bf16);
  const bool vec = (t.g_splits <= 1 || (t.g_split_stride & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                     reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15u) == 0 &&
                   (!pb || (t.shadow_cols == t.shadow_ld && (reinterpret_cast<uintptr_t>(pb) & 7u) == 0));
  if (vec) {
    const int64_t v_lo = lo >> 2, v_hi = hi >> 2;  // lo is a multiple of 4
    for (int64_t i = v_lo + threadIdx.x; i < v_hi; i += kAdamThreads) {
      float4 p = reinterpret_cast<float4*>(t.p)[i];
      const float4 g = adam_grad4(t, i);
      float4 m = reinterpret_cast<float4*>(t.m)[i];
      float4 v = reinterpret_cast<float4*>(t.v)[i];
      adam_one(p.x, g.x, m.x, v.x, s);
      adam_one(p.y, g.y, m.y, v.y, s);
      adam_one(p.z, g.z, m.z, v.z, s);
      adam_one(p.w, g.w, m.w, v.w, s);
      reinterpret_cast<float4*>(t.p)[i] = p;
      reinterpret_cast<float4*>(t.m)[i] = m;
      reinterpret_cast<float4*>(t.v)[i] = v;
      if (pb) st_bf16x4(pb + i * 4, p);
    }
    for (int64_t i = (v_hi << 2) + threadIdx.x; i < hi; i += kAdamThreads) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_one(p, adam_grad1(t, i), m, v, s);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (pb) pb[i] = __float2bfloat16_rn(p);
    }
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += kAdamThreads) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_one(p, adam_grad1(t, i), m, v, s);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (pb) {
        const int64_t r = i / t.shadow_cols, c = i - r * t.shadow_cols;
        pb[r * t.shadow_ld + c] = __float2bfloat16_rn(p);
      }
    }
  }
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_adam_step(const npt_adam_tensor* tensors_dev, const npt_adam_tensor* tensors_host,
                  int32_t n_tensors, float lr, float b1, float omb1, float b2, float omb2, float c1,
                  float c2, float eps, void* stream) {
  NPT_REQUIRE(n_tensors >= 0, "adam_step: negative tensor count");
  if (n_tensors == 0) return 0;
  NPT_REQUIRE(tensors_dev && tensors_host, "adam_step: null tensor table");
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    NPT_REQUIRE(tensors_host[i].n >= 0, "adam_step: tensor %d has negative size", i);
    NPT_REQUIRE(!tensors_host[i].p_bf16 || tensors_host[i].shadow_cols > 0,
                "adam_step: tensor %d has a shadow but no shadow_cols", i);
    blocks += (tensors_host[i].n + kAdamChunk - 1) / kAdamChunk;
  }
  if (blocks == 0) return 0;
  AdamScalars s{lr, b1, omb1, b2, omb2, c1, c2, eps};
  NPT_CHECK_CUDA(launch_pdl(adam_multi_kernel, dim3((unsigned)blocks), dim3(kAdamThreads), 0, (cudaStream_t)stream, tensors_dev, n_tensors, s));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
