// npt_gemm: the one strided-batched GEMM entry point behind Linear / Attention / OutputEmbedding
// (nn/linear.py:44-60, nn/attention.py:47-52,81-92, nn/embedding.py:92-102 of the reference).
// Dispatches to the fp32 FFMA kernel (gemm_simt.cu) or the tcgen05 kernels (gemm_tc.cu) and
// owns the deterministic split-K reduction.
#include "gemm_common.cuh"

namespace npt {

GemmEpilogue make_epilogue(const npt_gemm_args& a) {
  GemmEpilogue e;
  e.C = a.C; e.ldc = a.ldc; e.sco = a.stride_c_outer; e.sci = a.stride_c_inner;
  e.Cb = (__nv_bfloat16*)a.C_bf16; e.ldcb = a.ldcb; e.scbo = a.stride_cb_outer; e.scbi = a.stride_cb_inner;
  e.alpha = a.alpha;
  e.bias = a.bias;
  e.relu = a.relu;
  e.mask = a.mask; e.mask_bf16 = a.mask_is_bf16; e.ldm = a.ldmask; e.smo = a.stride_mask_outer; e.smi = a.stride_mask_inner;
  e.M = a.M; e.N = a.N; e.batch_inner = a.batch_inner;
  e.c_vec4 = a.C && aligned16(a.C) && a.ldc % 4 == 0 && a.stride_c_outer % 4 == 0 && a.stride_c_inner % 4 == 0;
  e.cb_vec4 = a.C_bf16 && aligned8(a.C_bf16) && a.ldcb % 4 == 0 && a.stride_cb_outer % 4 == 0 && a.stride_cb_inner % 4 == 0;
  if (a.mask) {
    const bool al = a.mask_is_bf16 ? aligned8(a.mask) : aligned16(a.mask);
    e.mask_vec4 = al && a.ldmask % 4 == 0 && a.stride_mask_outer % 4 == 0 && a.stride_mask_inner % 4 == 0;
  } else {
    e.mask_vec4 = 0;
  }
  e.bias_vec4 = a.bias && aligned16(a.bias);
  e.bits_out = a.relu_bits_out; e.bits_in = a.mask_bits; e.ld_bits = a.ld_bits;
  return e;
}

// partial layout: [split][batch][M][N] fp32, contiguous.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, GemmEpilogue e, int splits, int batch) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t mn = (int64_t)e.M * e.N;
  const int64_t total4 = ((int64_t)batch * mn + 3) / 4;
  const bool n4 = (e.N % 4) == 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (int64_t)gridDim.x * blockDim.x) {
    if (n4) {
      const int64_t i = q * 4;
      const int b = (int)(i / mn);
      const int64_t r = i - (int64_t)b * mn;
      const int m = (int)(r / e.N), n = (int)(r - (int64_t)m * e.N);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < splits; ++s) {
        const float4 p = *reinterpret_cast<const float4*>(partial + ((int64_t)s * batch) * mn + i);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      epi_store4(e, b, m, n, acc);
    } else {
      for (int k = 0; k < 4; ++k) {
        const int64_t i = q * 4 + k;
        if (i >= (int64_t)batch * mn) break;
        const int b = (int)(i / mn);
        const int64_t r = i - (int64_t)b * mn;
        const int m = (int)(r / e.N), n = (int)(r - (int64_t)m * e.N);
        float acc = 0.f;
        const float* p0 = partial + i;
        const int64_t stride = (int64_t)batch * mn;
        int s = 0;
        for (; s + 8 <= splits; s += 8) {   // eight independent loads in flight, additions in split order
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcs(p0 + (int64_t)(s + u) * stride);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; s < splits; ++s) acc += __ldcs(p0 + (int64_t)s * stride);
        epi_store1(e, b, m, n, acc);
      }
    }
  }
}

// N % 4 == 0: one thread per float4 of the output, indexed (column quad, row, batch) by the grid — no
// divisions (the generic kernel above spends two 64-bit divisions per element on them), four
// independent partial loads in flight per thread, the additions in split order 0..S-1 as before.
__global__ void __launch_bounds__(256)
splitk_reduce_vec_kernel(const float* __restrict__ partial, GemmEpilogue e, int splits, int batch) {
  pdl_wait();
  pdl_trigger();
  const int n = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int m = blockIdx.y * 8 + threadIdx.y;
  const int b = blockIdx.z;
  if (n >= e.N || m >= e.M) return;
  const int64_t mn = (int64_t)e.M * e.N;
  const int64_t stride = (int64_t)batch * mn;
  const float* p0 = partial + (int64_t)b * mn + (int64_t)m * e.N + n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  for (; s + 4 <= splits; s += 4) {
    const float4 a0 = __ldcs(reinterpret_cast<const float4*>(p0 + (int64_t)s * stride));
    const float4 a1 = __ldcs(reinterpret_cast<const float4*>(p0 + (int64_t)(s + 1) * stride));
    const float4 a2 = __ldcs(reinterpret_cast<const float4*>(p0 + (int64_t)(s + 2) * stride));
    const float4 a3 = __ldcs(reinterpret_cast<const float4*>(p0 + (int64_t)(s + 3) * stride));
    acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
    acc.x += a1.x; acc.y += a1.y; acc.z += a1.z; acc.w += a1.w;
    acc.x += a2.x; acc.y += a2.y; acc.z += a2.z; acc.w += a2.w;
    acc.x += a3.x; acc.y += a3.y; acc.z += a3.z; acc.w += a3.w;
  }
  for (; s < splits; ++s) {
    const float4 a0 = __ldcs(reinterpret_cast<const float4*>(p0 + (int64_t)s * stride));
    acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
  }
  epi_store4(e, b, m, n, acc);
}

// NPT_GEMM_TF32X3_TC runs on the tensor cores when TMA can address the operands, else on the FFMA kernel
static bool use_simt(const npt_gemm_args& a) {
  return a.mode == NPT_GEMM_FP32_SIMT || (a.mode == NPT_GEMM_TF32X3_TC && !gemm_tf32_supported(a));
}

static int choose_splits(const npt_gemm_args& a) {
  if (a.C_bf16_peer) return 1;   // the peer store is part of the tile epilogue: no split-K (small-batch launches would pick it)
  if (use_simt(a)) return gemm_simt_splits(a);
  if (a.mode == NPT_GEMM_TF32X3_TC) return gemm_tf32_splits(a);
  return gemm_tc_splits(a);
}

}  // namespace npt

using namespace npt;

extern "C" {

size_t npt_gemm_workspace_bytes(const npt_gemm_args* a) {
  if (!a) return 0;
  const int splits = choose_splits(*a);
  if (splits <= 1) return 0;
  return (size_t)splits * (size_t)a->batch * (size_t)a->M * (size_t)a->N * sizeof(float);
}

int32_t npt_gemm_splits(const npt_gemm_args* args) {
  if (!args) return 1;
  return (int32_t)choose_splits(*args);
}

int npt_gemm(const npt_gemm_args* args, void* stream) {
  NPT_REQUIRE(args != nullptr, "gemm: null args");
  const npt_gemm_args& a = *args;
  NPT_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  NPT_REQUIRE(a.batch >= 1 && a.batch_inner >= 1, "gemm: bad batch %d/%d", a.batch, a.batch_inner);
  NPT_REQUIRE(a.A && a.B, "gemm: null operand");
  NPT_REQUIRE(a.C || a.C_bf16, "gemm: no output");
  NPT_REQUIRE(a.lda >= (a.trans_a ? a.M : a.K), "gemm: lda too small");
  NPT_REQUIRE(a.ldb >= (a.trans_b ? a.K : a.N), "gemm: ldb too small");
  NPT_REQUIRE(!a.C || a.ldc >= a.N, "gemm: ldc too small");
  NPT_REQUIRE(!a.C_bf16 || a.ldcb >= a.N, "gemm: ldcb too small");
  NPT_REQUIRE(!a.mask || a.ldmask >= a.N, "gemm: ldmask too small");
  if (a.relu_bits_out || a.mask_bits) {
    NPT_REQUIRE(a.mode == NPT_GEMM_BF16_TC && a.batch == 1, "gemm: packed ReLU bits need the bf16 tensor-core mode and batch == 1");
    NPT_REQUIRE(a.ld_bits >= a.M, "gemm: ld_bits (row pitch of the word-major bit matrix) must be >= M");
    NPT_REQUIRE(!a.relu_bits_out || a.relu, "gemm: relu_bits_out needs relu = 1");
  }
  cudaStream_t st = (cudaStream_t)stream;
  NPT_REQUIRE(a.mode >= NPT_GEMM_FP32_SIMT && a.mode <= NPT_GEMM_TF32X3_TC, "gemm: unknown mode %d", a.mode);
  if (a.mode == NPT_GEMM_BF16_TC) {
    if (int rc = gemm_tc_supported(a)) return rc;
  } else {
    NPT_REQUIRE(!a.C_bf16_peer && !a.relu_bits_out && !a.mask_bits, "gemm: peer stores / packed ReLU bits need the bf16 tensor-core mode");
  }
  const GemmEpilogue epi = make_epilogue(a);
  const int splits = choose_splits(a);
  NPT_REQUIRE(!a.C_bf16_peer || (a.mode == NPT_GEMM_BF16_TC && splits == 1),
              "gemm: C_bf16_peer needs the bf16 tensor-core mode and a launch without split-K");
  float* partial = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * a.batch * (size_t)a.M * a.N * sizeof(float);
    NPT_REQUIRE(a.workspace && a.workspace_bytes >= need,
                "gemm: split-K needs %zu workspace bytes, got %zu", need, a.workspace_bytes);
    NPT_REQUIRE(aligned16(a.workspace), "gemm: workspace must be 16-byte aligned");
    partial = (float*)a.workspace;
  }
  int rc = use_simt(a) ? gemm_simt_launch(a, epi, splits, partial, st)
           : a.mode == NPT_GEMM_TF32X3_TC ? gemm_tf32_launch(a, epi, splits, partial, st)
                                          : gemm_tc_launch(a, epi, splits, partial, st);
  if (rc) return rc;
  if (splits > 1 && !a.skip_splitk_reduce) {
    if (a.N % 4 == 0 && aligned16(partial) && a.batch <= 65535 && cdiv(a.M, 8) <= 65535) {
      dim3 grid((unsigned)cdiv(a.N / 4, 32), (unsigned)cdiv(a.M, 8), (unsigned)a.batch);
      NPT_CHECK_CUDA(launch_pdl(splitk_reduce_vec_kernel, grid, dim3(32, 8), 0, st, partial, epi, splits, a.batch));
      NPT_LAUNCH_CHECK();
      return 0;
    }
    const int64_t total4 = ((int64_t)a.batch * a.M * a.N + 3) / 4;
    int64_t blocks = cdiv(total4, 256);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    NPT_CHECK_CUDA(launch_pdl(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, st, partial, epi, splits, a.batch));
    NPT_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
