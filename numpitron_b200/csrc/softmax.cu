// Causal-mask softmax over attention scores, forward and backward (stand-alone form).
// Replaces nn/attention.py:48-51 (scale, mask, softmax) and nn/attention.py:86-87 +
// nn/activation.py:22-31 (softmax backward, mask, scale).  The backward uses the compact form
// P*(dP - sum_j dP*P) of the reference's (B,H,S,S,S) Jacobian product.
//
// One warp per score row, row cached in registers, warp-shuffle reductions; only the lower
// triangle is read, the full row is written (zeros above the diagonal feed the P.V GEMM).
// HBM-bound: fwd ~ (S+1)/2*4 read + S*4 write per row; bwd reads P and dP triangles.
#include "common.cuh"

namespace npt {

constexpr int kSmWarps = 8;

template <int NPL>  // elements per lane; S <= NPL*32.  NPL == 0: uncached generic path.
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_causal_fwd_kernel(const float* __restrict__ scores, float* __restrict__ p,
                          __nv_bfloat16* __restrict__ pb, int64_t rows, int S, float divisor) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int i = (int)(row % S);  // query position inside its (S x S) matrix
  const float* sr = scores + row * S;
  float* pr = p ? p + row * S : nullptr;
  __nv_bfloat16* br = pb ? pb + row * S : nullptr;
  const int valid = i + 1;
  if (NPL > 0) {
    float x[NPL > 0 ? NPL : 1];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int j = lane + k * 32;
      x[k] = (j < valid) ? __fdiv_rn(sr[j], divisor) : -INFINITY;
      mx = fmaxf(mx, x[k]);
    }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int j = lane + k * 32;
      x[k] = (j < valid) ? expf(x[k] - mx) : 0.f;
      s += x[k];
    }
    s = warp_sum(s);
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int j = lane + k * 32;
      if (j < S) {
        const float o = __fdiv_rn(x[k], s);
        if (pr) pr[j] = o;
        if (br) br[j] = __float2bfloat16_rn(o);
      }
    }
  } else {
    float mx = -INFINITY;
    for (int j = lane; j < valid; j += 32) mx = fmaxf(mx, __fdiv_rn(sr[j], divisor));
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < valid; j += 32) s += expf(__fdiv_rn(sr[j], divisor) - mx);
    s = warp_sum(s);
    for (int j = lane; j < S; j += 32) {
      const float o = (j < valid) ? __fdiv_rn(expf(__fdiv_rn(sr[j], divisor) - mx), s) : 0.f;
      if (pr) pr[j] = o;
      if (br) br[j] = __float2bfloat16_rn(o);
    }
  }
}

template <int NPL>
__global__ void __launch_bounds__(kSmWarps * 32)
softmax_causal_bwd_kernel(const void* __restrict__ pv, int p_is_bf16, const float* __restrict__ dp,
                          float* __restrict__ ds, __nv_bfloat16* __restrict__ dsb, int64_t rows, int S,
                          float divisor) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kSmWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int i = (int)(row % S);
  const int valid = i + 1;
  const float* pr = reinterpret_cast<const float*>(pv) + row * S;
  const __nv_bfloat16* prb = reinterpret_cast<const __nv_bfloat16*>(pv) + row * S;
  const float* dr = dp + row * S;
  float* outr = ds ? ds + row * S : nullptr;
  __nv_bfloat16* br = dsb ? dsb + row * S : nullptr;
  if (NPL > 0) {
    float a[NPL > 0 ? NPL : 1], g[NPL > 0 ? NPL : 1];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int j = lane + k * 32;
      a[k] = (j < valid) ? (p_is_bf16 ? __bfloat162float(prb[j]) : pr[j]) : 0.f;
      g[k] = (j < valid) ? dr[j] : 0.f;
      dot += a[k] * g[k];
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const int j = lane + k * 32;
      if (j < S) {
        const float o = (j < valid) ? __fdiv_rn(a[k] * (g[k] - dot), divisor) : 0.f;
        if (outr) outr[j] = o;
        if (br) br[j] = __float2bfloat16_rn(o);
      }
    }
  } else {
    float dot = 0.f;
    for (int j = lane; j < valid; j += 32) dot += (p_is_bf16 ? __bfloat162float(prb[j]) : pr[j]) * dr[j];
    dot = warp_sum(dot);
    for (int j = lane; j < S; j += 32) {
      const float o = (j < valid) ? __fdiv_rn((p_is_bf16 ? __bfloat162float(prb[j]) : pr[j]) * (dr[j] - dot), divisor) : 0.f;
      if (outr) outr[j] = o;
      if (br) br[j] = __float2bfloat16_rn(o);
    }
  }
}

}  // namespace npt

using namespace npt;

#define NPT_SM_DISPATCH(KERNEL, ...)                                                              \
  do {                                                                                            \
    const unsigned grid = (unsigned)cdiv(rows, kSmWarps);                                         \
    if (S <= 64) KERNEL<2><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                          \
    else if (S <= 128) KERNEL<4><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                    \
    else if (S <= 256) KERNEL<8><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                    \
    else if (S <= 512) KERNEL<16><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                   \
    else if (S <= 1024) KERNEL<32><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                  \
    else KERNEL<0><<<grid, kSmWarps * 32, 0, st>>>(__VA_ARGS__);                                  \
  } while (0)

extern "C" {

int npt_softmax_causal_fwd(const float* scores, float* p, void* p_bf16, int64_t n_mats, int32_t S,
                           float divisor, void* stream) {
  NPT_REQUIRE(n_mats > 0 && S > 0, "softmax_causal_fwd: bad shape");
  NPT_REQUIRE(p || p_bf16, "softmax_causal_fwd: no output");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n_mats * S;
  NPT_SM_DISPATCH(softmax_causal_fwd_kernel, scores, p, (__nv_bfloat16*)p_bf16, rows, S, divisor);
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_softmax_causal_bwd(const void* p, int32_t p_is_bf16, const float* dp, float* ds, void* ds_bf16,
                           int64_t n_mats, int32_t S, float divisor, void* stream) {
  NPT_REQUIRE(n_mats > 0 && S > 0, "softmax_causal_bwd: bad shape");
  NPT_REQUIRE(ds || ds_bf16, "softmax_causal_bwd: no output");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n_mats * S;
  NPT_SM_DISPATCH(softmax_causal_bwd_kernel, p, p_is_bf16, dp, ds, (__nv_bfloat16*)ds_bf16, rows, S, divisor);
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
