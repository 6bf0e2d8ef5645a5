// Vocab-parallel embedding gather and the bit-exact ordered scatter-add.
// Replaces nn/embedding.py:44-83 (InputEmbedding.forward/backward) and :125-128 (PositionalEncoding).
//
// Integer/index work is exact by construction.  The scatter-add reproduces
//   np.add.at(grad.T, tok[~mask], d_out[~mask])
// bit for bit: for every vocab column the rows of d_out are added ONE BY ONE in increasing
// flattened (b,s) order onto the value already in grad.  No float atomics:
//   1) integer histogram of in-chunk tokens          (int atomics: order-independent result)
//   2) exclusive scan -> segment offsets
//   3) stable rank of every token inside its segment
//   4) order[offset+rank] = t
//   5) one block per vocab column walks its segment in order.
#include "common.cuh"

namespace npt {

// ---- gather ---------------------------------------------------------------------------------
// One warp per token; lanes stride over d.  E is (d, ldE): column reads are strided (the
// reference's layout), the writes are coalesced.
__global__ void __launch_bounds__(256)
embedding_gather_kernel(const uint16_t* __restrict__ tokens, const float* __restrict__ E, int64_t ldE,
                        const float* __restrict__ pos, float* __restrict__ out,
                        __nv_bfloat16* __restrict__ outb, int64_t T, int S, int d, int chunk_start,
                        int chunk_end, int masked) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const int tok = tokens[t];
  const bool outside = (tok < chunk_start) || (tok >= chunk_end);
  const int local = tok - chunk_start;
  const float* prow = pos ? pos + (int64_t)(t % S) * d : nullptr;
  for (int c = lane; c < d; c += 32) {
    float v = outside ? 0.f : __ldg(E + (int64_t)c * ldE + local);
    if (prow) v = __fadd_rn(__ldg(prow + c), v);  // encoding + inputs (embedding.py:127)
    out[t * d + c] = v;
    if (outb) outb[t * d + c] = __float2bfloat16_rn(v);
  }
  (void)masked;
}

__global__ void __launch_bounds__(256)
add_pos_kernel(const float* x, const float* __restrict__ pos, float* out, __nv_bfloat16* __restrict__ xb,
               int64_t T, int S, int d) {
  const int64_t n = T * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / d;
    const int c = (int)(i - t * d);
    const float v = __fadd_rn(pos[(t % S) * d + c], x[i]);
    out[i] = v;
    if (xb) xb[i] = __float2bfloat16_rn(v);
  }
}

// ---- scatter-add ------------------------------------------------------------------------------
__device__ __forceinline__ int local_id(int tok, int chunk_start, int chunk_end) {
  return (tok < chunk_start || tok >= chunk_end) ? -1 : tok - chunk_start;
}

__global__ void __launch_bounds__(256)
sa_count_kernel(const uint16_t* __restrict__ tokens, int* __restrict__ counts, int64_t T,
                int chunk_start, int chunk_end) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int v = local_id(tokens[t], chunk_start, chunk_end);
  if (v >= 0) atomicAdd(counts + v, 1);
}

// Single block exclusive scan of counts[0..V) into offsets[0..V].
__global__ void __launch_bounds__(1024)
sa_scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets, int V) {
  __shared__ int part[1024];
  const int per = (V + 1023) / 1024;
  const int lo = threadIdx.x * per;
  int hi = lo + per;
  if (hi > V) hi = V;
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    int add = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;  // exclusive prefix of this thread's chunk
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
  if (threadIdx.x == 1023) offsets[V] = part[1023];
}

// rank[t] = #{t' < t : tok[t'] == tok[t]}; then order[offsets[v] + rank] = t.
constexpr int kRankTile = 2048;
__global__ void __launch_bounds__(256)
sa_rank_fill_kernel(const uint16_t* __restrict__ tokens, const int* __restrict__ offsets,
                    int* __restrict__ order, int64_t T, int chunk_start, int chunk_end) {
  __shared__ uint16_t tile[kRankTile];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t_block_end = min((int64_t)(blockIdx.x + 1) * blockDim.x, T);  // exclusive
  const int mytok = (t < T) ? tokens[t] : -1;
  int rank = 0;
  for (int64_t base = 0; base < t_block_end; base += kRankTile) {
    const int64_t lim = min((int64_t)kRankTile, t_block_end - base);
    __syncthreads();
    for (int i = threadIdx.x; i < lim; i += blockDim.x) tile[i] = tokens[base + i];
    __syncthreads();
    if (t < T) {
      const int64_t upto = min(lim, t - base);  // only t' < t
      for (int64_t i = 0; i < upto; ++i) rank += (tile[i] == mytok);
    }
  }
  if (t < T) {
    const int v = local_id(mytok, chunk_start, chunk_end);
    if (v >= 0) order[offsets[v] + rank] = (int)t;
  }
}

// One block per vocab column; threads over d.  Sequential fp32 adds in segment order.
__global__ void __launch_bounds__(128)
sa_accumulate_kernel(const int* __restrict__ offsets, const int* __restrict__ order,
                     const float* __restrict__ d_out, float* __restrict__ grad, int64_t ldG, int d) {
  const int v = blockIdx.x;
  const int beg = offsets[v], end = offsets[v + 1];
  if (beg == end) return;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = grad[(int64_t)c * ldG + v];
    int i = beg;
    for (; i + 4 <= end; i += 4) {  // loads are independent of the adds: issue 4 at a time
      const float a0 = d_out[(int64_t)order[i] * d + c];
      const float a1 = d_out[(int64_t)order[i + 1] * d + c];
      const float a2 = d_out[(int64_t)order[i + 2] * d + c];
      const float a3 = d_out[(int64_t)order[i + 3] * d + c];
      acc = __fadd_rn(acc, a0);
      acc = __fadd_rn(acc, a1);
      acc = __fadd_rn(acc, a2);
      acc = __fadd_rn(acc, a3);
    }
    for (; i < end; ++i) acc = __fadd_rn(acc, d_out[(int64_t)order[i] * d + c]);
    grad[(int64_t)c * ldG + v] = acc;
  }
}

static size_t sa_align(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace npt

using namespace npt;

extern "C" {

int npt_embedding_gather(const uint16_t* tokens, const float* E, int64_t ldE, const float* pos,
                         float* out, void* out_bf16, int64_t T, int32_t S, int32_t d,
                         int32_t chunk_start, int32_t chunk_end, int32_t masked, void* stream) {
  NPT_REQUIRE(T >= 0 && S > 0 && d > 0 && chunk_end >= chunk_start, "embedding_gather: bad shape");
  if (T == 0) return 0;
  const int warps = 8;
  embedding_gather_kernel<<<(unsigned)cdiv(T, warps), warps * 32, 0, (cudaStream_t)stream>>>(
      tokens, E, ldE, pos, out, (__nv_bfloat16*)out_bf16, T, S, d, chunk_start, chunk_end, masked);
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_add_pos(const float* x, const float* pos, float* out, void* x_bf16, int64_t T, int32_t S, int32_t d, void* stream) {
  if (T <= 0) return 0;
  int64_t blocks = cdiv(T * d, 256);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  add_pos_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, pos, out, (__nv_bfloat16*)x_bf16, T, S, d);
  NPT_LAUNCH_CHECK();
  return 0;
}

size_t npt_embedding_scatter_add_workspace_bytes(int64_t T, int32_t V_local) {
  return sa_align((size_t)(V_local + 1) * 4) * 2 + sa_align((size_t)T * 4);
}

int npt_embedding_scatter_add(const uint16_t* tokens, const float* d_out, float* grad, int64_t ldG,
                              int64_t T, int32_t d, int32_t V_local, int32_t chunk_start,
                              int32_t chunk_end, int32_t masked, void* workspace,
                              size_t workspace_bytes, void* stream) {
  NPT_REQUIRE(T > 0 && d > 0 && V_local > 0, "scatter_add: bad shape");
  NPT_REQUIRE(chunk_end - chunk_start <= V_local, "scatter_add: chunk wider than V_local");
  NPT_REQUIRE(T < (int64_t)1 << 31, "scatter_add: T too large");
  NPT_REQUIRE(workspace && workspace_bytes >= npt_embedding_scatter_add_workspace_bytes(T, V_local),
              "scatter_add: workspace too small");
  (void)masked;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int* counts = (int*)ws;
  int* offsets = (int*)(ws + sa_align((size_t)(V_local + 1) * 4));
  int* order = (int*)(ws + 2 * sa_align((size_t)(V_local + 1) * 4));
  NPT_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)(V_local + 1) * 4, st));
  sa_count_kernel<<<(unsigned)cdiv(T, 256), 256, 0, st>>>(tokens, counts, T, chunk_start, chunk_end);
  NPT_LAUNCH_CHECK();
  sa_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, V_local);
  NPT_LAUNCH_CHECK();
  sa_rank_fill_kernel<<<(unsigned)cdiv(T, 256), 256, 0, st>>>(tokens, offsets, order, T, chunk_start, chunk_end);
  NPT_LAUNCH_CHECK();
  sa_accumulate_kernel<<<(unsigned)V_local, 128, 0, st>>>(offsets, order, d_out, grad, ldG, d);
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
