// Vocab-parallel softmax cross-entropy — replaces nn/loss.py:6-50.
//
// The reference interleaves three TP all-reduces (row max, picked logit, sum-exp) with the
// arithmetic, so the kernel is split at exactly those points (stage1/2/3); `fused` is the
// tp == 1 case in one pass.  Integer label math (range mask, shift) is exact; the float part
// uses expf/logf/IEEE division so it tracks NumPy's float32 path to ~1 ulp.
// HBM-bound: algorithmic bytes T*(8*V_local + 6) for the fused form.
#include "common.cuh"

namespace npt {

enum { CE_FUSED = 0, CE_STAGE1 = 1, CE_STAGE2 = 2, CE_STAGE3 = 3 };

template <bool BLOCK>
struct Group {
  float* red;
  __device__ __forceinline__ int tid() const { return BLOCK ? threadIdx.x : (threadIdx.x & 31); }
  __device__ __forceinline__ int size() const { return BLOCK ? blockDim.x : 32; }
  __device__ __forceinline__ float sum(float v) const { return BLOCK ? block_sum(v, red) : warp_sum(v); }
  __device__ __forceinline__ float max(float v) const { return BLOCK ? block_max(v, red) : warp_max(v); }
};

// Walks one logits row 128 bits at a time.  V_local is odd for every BASELINE vocabulary (65, 6283, 25129,
// 50257), so rows start at every alignment: scalar elements up to the first 16-byte boundary, float4 body, scalar
// tail.  f(v, x) is called for every element (v = column).
template <typename F>
__device__ __forceinline__ void ce_row_read(const float* __restrict__ lr, int V, int tid, int n, F f) {
  if (V < 128) {   // a row of the Shakespeare vocabulary is 65 floats: 32 lanes x scalar loads beat 15 lanes x float4
    for (int v = tid; v < V; v += n) f(v, lr[v]);
    return;
  }
  int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(lr) & 15u)) & 15u) >> 2);
  if (head > V) head = V;
  for (int v = tid; v < head; v += n) f(v, lr[v]);
  const int nvec = (V - head) >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(lr + head);
  for (int i = tid; i < nvec; i += n) {
    const float4 x = p4[i];
    const int v = head + 4 * i;
    f(v, x.x); f(v + 1, x.y); f(v + 2, x.z); f(v + 3, x.w);
  }
  for (int v = head + 4 * nvec + tid; v < V; v += n) f(v, lr[v]);
}

template <bool BLOCK, int STAGE>
__global__ void __launch_bounds__(BLOCK ? 1024 : 256)
ce_kernel(const float* __restrict__ logits, int64_t ld, const uint16_t* __restrict__ labels,
          float* __restrict__ rowmax, float* __restrict__ picked, float* __restrict__ sumexp,
          float* __restrict__ loss, float* __restrict__ grad, int64_t ldg,
          __nv_bfloat16* __restrict__ gradb, int64_t ldgb, int64_t T, int V, int chunk_start,
          int chunk_end) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float red[32];
  Group<BLOCK> g{red};
  const int64_t row = BLOCK ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;  // BLOCK: whole block exits together; warp mode has no block syncs
  const float* lr = logits + row * ld;
  const int tid = g.tid(), n = g.size();

  float m;
  if (STAGE == CE_FUSED || STAGE == CE_STAGE1) {
    float mx = -INFINITY;
    ce_row_read(lr, V, tid, n, [&](int, float x) { mx = fmaxf(mx, x); });
    m = g.max(mx);
    if (STAGE == CE_STAGE1) {
      if (tid == 0) rowmax[row] = m;
      return;
    }
  } else {
    m = rowmax[row];
  }

  int label = 0;
  bool inside = false;
  if (STAGE != CE_STAGE1) {
    const int lab = labels[row];
    inside = !(lab < chunk_start || lab >= chunk_end);
    label = inside ? lab - chunk_start : 0;  // labels[mask] = 0 (loss.py:17-18)
  }

  float pk, se;
  if (STAGE == CE_FUSED || STAGE == CE_STAGE2) {
    float s = 0.f;
    ce_row_read(lr, V, tid, n, [&](int, float x) { s += expf(x - m); });
    se = g.sum(s);
    pk = inside ? (lr[label] - m) : 0.f;
    if (STAGE == CE_STAGE2) {
      if (tid == 0) {
        picked[row] = pk;
        sumexp[row] = se;
      }
      return;
    }
  } else {
    pk = picked[row];
    se = sumexp[row];
  }

  // STAGE3 / FUSED tail
  if (tid == 0) loss[row] = __fadd_rn(-pk, logf(se));
  float* gr = grad + row * ldg;
  __nv_bfloat16* gb = gradb ? gradb + row * ldgb : nullptr;
  auto prob = [&](int v, float x) {
    float p = __fdiv_rn(expf(x - m), se);
    if (inside && v == label) p = __fsub_rn(p, 1.0f);
    return p;
  };
  if (V >= 128 && (reinterpret_cast<uintptr_t>(gr) & 15u) == (reinterpret_cast<uintptr_t>(lr) & 15u)) {
    // the gradient row has the alignment of the logits row (same pitch): 128-bit loads and stores
    int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(lr) & 15u)) & 15u) >> 2);
    if (head > V) head = V;
    for (int v = tid; v < head; v += n) {
      const float p = prob(v, lr[v]);
      gr[v] = p;
      if (gb) gb[v] = __float2bfloat16_rn(p);
    }
    const int nvec = (V - head) >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(lr + head);
    float4* g4 = reinterpret_cast<float4*>(gr + head);
    for (int i = tid; i < nvec; i += n) {
      const float4 x = p4[i];
      const int v = head + 4 * i;
      float4 p;
      p.x = prob(v, x.x); p.y = prob(v + 1, x.y); p.z = prob(v + 2, x.z); p.w = prob(v + 3, x.w);
      g4[i] = p;
      if (gb) {
        if ((head & 1) == 0) {   // 2 * (head + 4 i) bytes: 4-byte aligned
          reinterpret_cast<uint32_t*>(gb + v)[0] = pack_bf16(p.x, p.y);
          reinterpret_cast<uint32_t*>(gb + v)[1] = pack_bf16(p.z, p.w);
        } else {
          gb[v] = __float2bfloat16_rn(p.x);
          *reinterpret_cast<uint32_t*>(gb + v + 1) = pack_bf16(p.y, p.z);
          gb[v + 3] = __float2bfloat16_rn(p.w);
        }
      }
    }
    for (int v = head + 4 * nvec + tid; v < V; v += n) {
      const float p = prob(v, lr[v]);
      gr[v] = p;
      if (gb) gb[v] = __float2bfloat16_rn(p);
    }
  } else {
    for (int v = tid; v < V; v += n) {
      const float p = prob(v, lr[v]);
      gr[v] = p;
      if (gb) gb[v] = __float2bfloat16_rn(p);
    }
  }
  if (gb)  // zero the shadow's padding columns so padded GEMM operands stay clean
    for (int v = V + tid; v < ldgb; v += n) gb[v] = __float2bfloat16_rn(0.f);
}

template <int STAGE>
static int launch_ce(const float* logits, int64_t ld, const uint16_t* labels, float* rowmax,
                     float* picked, float* sumexp, float* loss, float* grad, int64_t ldg,
                     void* gradb, int64_t ldgb, int64_t T, int V, int cs, int ce, cudaStream_t st) {
  if (V <= 2048) {
    const int warps = 8;
    NPT_CHECK_CUDA(launch_pdl(ce_kernel<false, STAGE>, dim3((unsigned)cdiv(T, warps)), dim3(warps * 32), 0, st, 
        logits, ld, labels, rowmax, picked, sumexp, loss, grad, ldg, (__nv_bfloat16*)gradb, ldgb, T, V, cs, ce));
  } else {
    // one row per block; wide rows get fat blocks so that few enough rows are in flight for the later passes of the
    // fused form to find theirs in L2 (1184 rows x 201 KB at 256 threads did not fit the 126 MB)
    const int threads = V >= 16384 ? 1024 : 256;   // (512 threads at V = 6283 measured slower than 256: 503 vs 470 us)
    NPT_CHECK_CUDA(launch_pdl(ce_kernel<true, STAGE>, dim3((unsigned)T), dim3(threads), 0, st, 
        logits, ld, labels, rowmax, picked, sumexp, loss, grad, ldg, (__nv_bfloat16*)gradb, ldgb, T, V, cs, ce));
  }
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_softmax_ce_stage1(const float* logits, int64_t ld, float* rowmax, int64_t T, int32_t V_local, void* stream) {
  NPT_REQUIRE(T > 0 && V_local > 0 && ld >= V_local, "softmax_ce_stage1: bad shape");
  return launch_ce<CE_STAGE1>(logits, ld, nullptr, rowmax, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0,
                              T, V_local, 0, 0, (cudaStream_t)stream);
}

int npt_softmax_ce_stage2(const float* logits, int64_t ld, const uint16_t* labels, const float* rowmax,
                          float* picked, float* sumexp, int64_t T, int32_t V_local, int32_t chunk_start,
                          int32_t chunk_end, void* stream) {
  NPT_REQUIRE(T > 0 && V_local > 0 && ld >= V_local, "softmax_ce_stage2: bad shape");
  return launch_ce<CE_STAGE2>(logits, ld, labels, (float*)rowmax, picked, sumexp, nullptr, nullptr, 0, nullptr, 0,
                              T, V_local, chunk_start, chunk_end, (cudaStream_t)stream);
}

int npt_softmax_ce_stage3(const float* logits, int64_t ld, const uint16_t* labels, const float* rowmax,
                          const float* picked, const float* sumexp, float* loss, float* grad, int64_t ldg,
                          void* grad_bf16, int64_t ldgb, int64_t T, int32_t V_local, int32_t chunk_start,
                          int32_t chunk_end, void* stream) {
  NPT_REQUIRE(T > 0 && V_local > 0 && ld >= V_local && ldg >= V_local, "softmax_ce_stage3: bad shape");
  NPT_REQUIRE(!grad_bf16 || ldgb >= V_local, "softmax_ce_stage3: bf16 pitch too small");
  return launch_ce<CE_STAGE3>(logits, ld, labels, (float*)rowmax, (float*)picked, (float*)sumexp, loss, grad, ldg,
                              grad_bf16, ldgb, T, V_local, chunk_start, chunk_end, (cudaStream_t)stream);
}

int npt_softmax_ce_fused(const float* logits, int64_t ld, const uint16_t* labels, float* loss, float* grad,
                         int64_t ldg, void* grad_bf16, int64_t ldgb, int64_t T, int32_t V_local,
                         int32_t chunk_start, int32_t chunk_end, void* stream) {
  NPT_REQUIRE(T > 0 && V_local > 0 && ld >= V_local && ldg >= V_local, "softmax_ce_fused: bad shape");
  NPT_REQUIRE(!grad_bf16 || ldgb >= V_local, "softmax_ce_fused: bf16 pitch too small");
  return launch_ce<CE_FUSED>(logits, ld, labels, nullptr, nullptr, nullptr, loss, grad, ldg, grad_bf16, ldgb,
                             T, V_local, chunk_start, chunk_end, (cudaStream_t)stream);
}

}  // extern "C"
