// Shared GEMM epilogue: alpha, bias, ReLU, ReLU-mask, fp32 and/or bf16 stores.
#pragma once
#include "common.cuh"

namespace npt {

struct GemmEpilogue {
  float* C; int64_t ldc, sco, sci;
  __nv_bfloat16* Cb; int64_t ldcb, scbo, scbi;
  float alpha;
  const float* bias;
  int relu;
  const void* mask; int mask_bf16; int64_t ldm, smo, smi;
  int M, N, batch_inner;
  int c_vec4, cb_vec4, mask_vec4, bias_vec4;  // host-checked alignment of 4-wide accesses
  uint32_t* bits_out; const uint32_t* bits_in; int64_t ld_bits;  // packed ReLU mask (tensor-core path)
};

__device__ __forceinline__ float epi_value(const GemmEpilogue& e, int64_t mask_off, int n, float acc) {
  float v = acc * e.alpha;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.mask) {
    const float mk = e.mask_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.mask)[mask_off])
                                 : reinterpret_cast<const float*>(e.mask)[mask_off];
    v = mk > 0.f ? v : 0.f;
  }
  return v;
}

// One element.
__device__ __forceinline__ void epi_store1(const GemmEpilogue& e, int b, int m, int n, float acc) {
  if (m >= e.M || n >= e.N) return;
  const int bo = b / e.batch_inner, bi = b - bo * e.batch_inner;
  const float v = epi_value(e, bo * e.smo + bi * e.smi + (int64_t)m * e.ldm + n, n, acc);
  if (e.C) e.C[bo * e.sco + bi * e.sci + (int64_t)m * e.ldc + n] = v;
  if (e.Cb) e.Cb[bo * e.scbo + bi * e.scbi + (int64_t)m * e.ldcb + n] = __float2bfloat16_rn(v);
}

// Four consecutive columns n..n+3 (n % 4 == 0).  Falls back to scalars at the edges / when unaligned.
__device__ __forceinline__ void epi_store4(const GemmEpilogue& e, int b, int m, int n, float4 acc) {
  if (m >= e.M || n >= e.N) return;
  if (n + 3 >= e.N) {
    epi_store1(e, b, m, n, acc.x);
    epi_store1(e, b, m, n + 1, acc.y);
    epi_store1(e, b, m, n + 2, acc.z);
    return;  // n+3 >= N: the 4th is out of range
  }
  const int bo = b / e.batch_inner, bi = b - bo * e.batch_inner;
  float4 v = make_float4(acc.x * e.alpha, acc.y * e.alpha, acc.z * e.alpha, acc.w * e.alpha);
  if (e.bias) {
    if (e.bias_vec4) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n));
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    } else {
      v.x += __ldg(e.bias + n); v.y += __ldg(e.bias + n + 1);
      v.z += __ldg(e.bias + n + 2); v.w += __ldg(e.bias + n + 3);
    }
  }
  if (e.relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  if (e.mask) {
    const int64_t off = bo * e.smo + bi * e.smi + (int64_t)m * e.ldm + n;
    float4 mk;
    if (e.mask_bf16) {
      const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(e.mask) + off;
      if (e.mask_vec4) {
        const uint2 u = *reinterpret_cast<const uint2*>(mp);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        mk = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
      } else {
        mk = make_float4(__bfloat162float(mp[0]), __bfloat162float(mp[1]), __bfloat162float(mp[2]),
                         __bfloat162float(mp[3]));
      }
    } else {
      const float* mp = reinterpret_cast<const float*>(e.mask) + off;
      if (e.mask_vec4) mk = *reinterpret_cast<const float4*>(mp);
      else mk = make_float4(mp[0], mp[1], mp[2], mp[3]);
    }
    v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
    v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
  }
  if (e.C) {
    float* cp = e.C + bo * e.sco + bi * e.sci + (int64_t)m * e.ldc + n;
    if (e.c_vec4) *reinterpret_cast<float4*>(cp) = v;
    else { cp[0] = v.x; cp[1] = v.y; cp[2] = v.z; cp[3] = v.w; }
  }
  if (e.Cb) {
    __nv_bfloat16* cp = e.Cb + bo * e.scbo + bi * e.scbi + (int64_t)m * e.ldcb + n;
    if (e.cb_vec4) st_bf16x4(cp, v);
    else {
      cp[0] = __float2bfloat16_rn(v.x); cp[1] = __float2bfloat16_rn(v.y);
      cp[2] = __float2bfloat16_rn(v.z); cp[3] = __float2bfloat16_rn(v.w);
    }
  }
}

// Provided by gemm.cu
GemmEpilogue make_epilogue(const npt_gemm_args& a);
int gemm_simt_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st);
int gemm_tc_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st);
int gemm_tc_supported(const npt_gemm_args& a);  // 0 = ok, else sets error
int gemm_simt_splits(const npt_gemm_args& a);
int gemm_tc_splits(const npt_gemm_args& a);
// 3xTF32 tensor-core kernel (gemm_tf32.cu)
bool gemm_tf32_supported(const npt_gemm_args& a);
int gemm_tf32_splits(const npt_gemm_args& a);
int gemm_tf32_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st);

}  // namespace npt
