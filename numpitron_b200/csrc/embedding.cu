// Vocab-parallel embedding gather and the bit-exact ordered scatter-add.
// Replaces nn/embedding.py:44-83 (InputEmbedding.forward/backward) and :125-128 (PositionalEncoding).
//
// Integer/index work is exact by construction.  The scatter-add reproduces
//   np.add.at(grad.T, tok[~mask], d_out[~mask])
// bit for bit: for every vocab column the rows of d_out are added ONE BY ONE in increasing
// flattened (b,s) order onto the value already in grad.  No float atomics:
//   1) per 256-token chunk: integer histogram table[chunk][v] and the stable rank of every token
//      among the equal tokens of its own chunk (256^2/2 comparisons in shared memory);
//   2) column scan of the table over chunks (-> tokens of earlier chunks) and scan over v (-> segment
//      offsets);
//   3) order[offset[v] + earlier[chunk][v] + rank] = t;
//   4) one block per (vocab column, 128-wide slice of d) walks its segment in order.
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
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
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
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t n = T * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / d;
    const int c = (int)(i - t * d);
    const float v = __fadd_rn(pos[(t % S) * d + c], x[i]);
    out[i] = v;
    if (xb) xb[i] = __float2bfloat16_rn(v);
  }
}

// d % 4 == 0, 16-byte aligned buffers: a warp per token row, 128-bit accesses, no per-element division
__global__ void __launch_bounds__(256)
add_pos_vec_kernel(const float* x, const float* __restrict__ pos, float* out, __nv_bfloat16* __restrict__ xb,
                   int64_t T, int S, int d) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const float4* xr = reinterpret_cast<const float4*>(x + t * d);
  const float4* pr = reinterpret_cast<const float4*>(pos + (int64_t)(t % S) * d);
  float4* orow = reinterpret_cast<float4*>(out + t * d);
  for (int c = lane; c < (d >> 2); c += 32) {
    const float4 a = xr[c], p = __ldg(pr + c);
    float4 v;
    v.x = __fadd_rn(p.x, a.x); v.y = __fadd_rn(p.y, a.y); v.z = __fadd_rn(p.z, a.z); v.w = __fadd_rn(p.w, a.w);
    orow[c] = v;
    if (xb) st_bf16x4(xb + t * d + (int64_t)c * 4, v);
  }
}

// ---- scatter-add ------------------------------------------------------------------------------
constexpr int kSaChunk = 256;

__device__ __forceinline__ int local_id(int tok, int chunk_start, int chunk_end) {
  return (tok < chunk_start || tok >= chunk_end) ? -1 : tok - chunk_start;
}

// Per chunk of 256 tokens: rank[t] = #{t' < t in the same chunk with the same token};
// table[chunk][v] = number of in-chunk tokens with local id v (table pre-zeroed).
__global__ void __launch_bounds__(kSaChunk)
sa_chunk_kernel(const uint16_t* __restrict__ tokens, int* __restrict__ table, int* __restrict__ rank,
                int64_t T, int V, int chunk_start, int chunk_end) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ int ids[kSaChunk];
  const int64_t t = (int64_t)blockIdx.x * kSaChunk + threadIdx.x;
  const int v = (t < T) ? local_id(tokens[t], chunk_start, chunk_end) : -1;
  ids[threadIdx.x] = v;
  __syncthreads();
  if (v < 0) return;
  int r = 0;
  for (int i = 0; i < (int)threadIdx.x; ++i) r += (ids[i] == v);
  rank[t] = r;
  atomicAdd(table + (size_t)blockIdx.x * V + v, 1);  // integer: result independent of order
}

// Column scan over chunks, in place: table[c][v] <- sum_{c' < c} table[c'][v]; counts[v] <- total.
__global__ void __launch_bounds__(256)
sa_colscan_kernel(int* __restrict__ table, int* __restrict__ counts, int nchunks, int V) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int run = 0;
  int c = 0;
  for (; c + 16 <= nchunks; c += 16) {   // sixteen independent loads, then the running sum (the loop was one load per step)
    int n[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) n[u] = table[(size_t)(c + u) * V + v];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      table[(size_t)(c + u) * V + v] = run;
      run += n[u];
    }
  }
  for (; c < nchunks; ++c) {
    const int n = table[(size_t)c * V + v];
    table[(size_t)c * V + v] = run;
    run += n;
  }
  counts[v] = run;
}

// Single block exclusive scan of counts[0..V) into offsets[0..V].
__global__ void __launch_bounds__(1024)
sa_scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets, int V) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ int part[1024];
  const int per = (V + 1023) / 1024;
  const int lo = threadIdx.x * per;
  int hi = lo + per;
  if (hi > V) hi = V;
  int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan over the 1024 partials
    int add = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  int run = part[threadIdx.x] - s;
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
  if (threadIdx.x == 1023) offsets[V] = part[1023];
}

__global__ void __launch_bounds__(256)
sa_fill_kernel(const uint16_t* __restrict__ tokens, const int* __restrict__ table, const int* __restrict__ rank,
               const int* __restrict__ offsets, int* __restrict__ order, int64_t T, int V, int chunk_start,
               int chunk_end) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int v = local_id(tokens[t], chunk_start, chunk_end);
  if (v < 0) return;
  const int64_t c = t / kSaChunk;
  order[offsets[v] + table[(size_t)c * V + v] + rank[t]] = (int)t;
}

// grid (V, ceil(d/128)); sequential fp32 adds in segment order, 16 loads in flight per thread.
__global__ void __launch_bounds__(128)
sa_accumulate_kernel(const int* __restrict__ offsets, const int* __restrict__ order,
                     const float* __restrict__ d_out, float* __restrict__ grad, int64_t ldG, int d) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int v = blockIdx.x;
  const int beg = offsets[v], end = offsets[v + 1];
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (beg == end || c >= d) return;
  float acc = grad[(int64_t)c * ldG + v];
  int i = beg;
  for (; i + 16 <= end; i += 16) {   // sixteen loads in flight per thread: the chain of additions is latency-bound
    int idx[16];
    float a[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) idx[u] = __ldg(order + i + u);
#pragma unroll
    for (int u = 0; u < 16; ++u) a[u] = __ldg(d_out + (int64_t)idx[u] * d + c);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, a[u]);  // strictly in order
  }
  for (; i + 4 <= end; i += 4) {
    int idx[4];
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) idx[u] = __ldg(order + i + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = __ldg(d_out + (int64_t)idx[u] * d + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = __fadd_rn(acc, a[u]);
  }
  for (; i < end; ++i) acc = __fadd_rn(acc, __ldg(d_out + (int64_t)__ldg(order + i) * d + c));
  grad[(int64_t)c * ldG + v] = acc;
}

static size_t sa_align(size_t x) { return (x + 255) & ~(size_t)255; }

struct SaLayout {
  size_t table, counts, offsets, rank, order, total;
  SaLayout(int64_t T, int V) {
    const size_t nchunks = (size_t)cdiv(T, kSaChunk);
    table = 0;
    counts = table + sa_align(nchunks * V * 4);
    offsets = counts + sa_align((size_t)(V + 1) * 4);
    rank = offsets + sa_align((size_t)(V + 1) * 4);
    order = rank + sa_align((size_t)T * 4);
    total = order + sa_align((size_t)T * 4);
  }
};


// ---------------------------------------------------------------------------------------
// Token windows of the data loader (data/loader.py:49-54): inputs = data[start + 0..S-1],
// labels = data[start + 1 .. start + S], for B random starts, from a token file that was staged in
// HBM once.  Pure index work on uint16: bit-exact by construction.
__global__ void __launch_bounds__(256)
window_gather_kernel(const uint16_t* __restrict__ tokens, const int64_t* __restrict__ starts,
                     uint16_t* __restrict__ x, uint16_t* __restrict__ y, int B, int S) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * S) return;
  const int b = (int)(i / S), s = (int)(i - (int64_t)b * S);
  const int64_t src = starts[b] + s;
  x[i] = tokens[src];
  y[i] = tokens[src + 1];
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_embedding_gather(const uint16_t* tokens, const float* E, int64_t ldE, const float* pos,
                         float* out, void* out_bf16, int64_t T, int32_t S, int32_t d,
                         int32_t chunk_start, int32_t chunk_end, int32_t masked, void* stream) {
  NPT_REQUIRE(T >= 0 && S > 0 && d > 0 && chunk_end >= chunk_start, "embedding_gather: bad shape");
  if (T == 0) return 0;
  const int warps = 8;
  NPT_CHECK_CUDA(launch_pdl(embedding_gather_kernel, dim3((unsigned)cdiv(T, warps)), dim3(warps * 32), 0, (cudaStream_t)stream, 
      tokens, E, ldE, pos, out, (__nv_bfloat16*)out_bf16, T, S, d, chunk_start, chunk_end, masked));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_add_pos(const float* x, const float* pos, float* out, void* x_bf16, int64_t T, int32_t S, int32_t d, void* stream) {
  if (T <= 0) return 0;
  if (d % 4 == 0 && aligned16(x) && aligned16(pos) && aligned16(out) && (!x_bf16 || aligned8(x_bf16))) {
    NPT_CHECK_CUDA(launch_pdl(add_pos_vec_kernel, dim3((unsigned)cdiv(T, 8)), dim3(256), 0, (cudaStream_t)stream, x, pos, out,
                              (__nv_bfloat16*)x_bf16, T, S, d));
    NPT_LAUNCH_CHECK();
    return 0;
  }
  int64_t blocks = cdiv(T * d, 256);
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  NPT_CHECK_CUDA(launch_pdl(add_pos_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, pos, out, (__nv_bfloat16*)x_bf16, T, S, d));
  NPT_LAUNCH_CHECK();
  return 0;
}

size_t npt_embedding_scatter_add_workspace_bytes(int64_t T, int32_t V_local) {
  return SaLayout(T, V_local).total;
}

int npt_embedding_scatter_add(const uint16_t* tokens, const float* d_out, float* grad, int64_t ldG,
                              int64_t T, int32_t d, int32_t V_local, int32_t chunk_start,
                              int32_t chunk_end, int32_t masked, void* workspace,
                              size_t workspace_bytes, void* stream) {
  NPT_REQUIRE(T > 0 && d > 0 && V_local > 0, "scatter_add: bad shape");
  NPT_REQUIRE(chunk_end - chunk_start <= V_local, "scatter_add: chunk wider than V_local");
  NPT_REQUIRE(T < (int64_t)1 << 31, "scatter_add: T too large");
  const SaLayout L(T, V_local);
  NPT_REQUIRE(workspace && workspace_bytes >= L.total, "scatter_add: workspace too small (%zu < %zu)", workspace_bytes, L.total);
  (void)masked;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int* table = (int*)(ws + L.table);
  int* counts = (int*)(ws + L.counts);
  int* offsets = (int*)(ws + L.offsets);
  int* rank = (int*)(ws + L.rank);
  int* order = (int*)(ws + L.order);
  const int nchunks = (int)cdiv(T, kSaChunk);
  NPT_CHECK_CUDA(cudaMemsetAsync(table, 0, (size_t)nchunks * V_local * 4, st));
  NPT_CHECK_CUDA(launch_pdl(sa_chunk_kernel, dim3(nchunks), dim3(kSaChunk), 0, st, tokens, table, rank, T, V_local, chunk_start, chunk_end));
  NPT_LAUNCH_CHECK();
  NPT_CHECK_CUDA(launch_pdl(sa_colscan_kernel, dim3((unsigned)cdiv(V_local, 256)), dim3(256), 0, st, table, counts, nchunks, V_local));
  NPT_LAUNCH_CHECK();
  NPT_CHECK_CUDA(launch_pdl(sa_scan_kernel, dim3(1), dim3(1024), 0, st, counts, offsets, V_local));
  NPT_LAUNCH_CHECK();
  NPT_CHECK_CUDA(launch_pdl(sa_fill_kernel, dim3((unsigned)cdiv(T, 256)), dim3(256), 0, st, tokens, table, rank, offsets, order, T, V_local, chunk_start, chunk_end));
  NPT_LAUNCH_CHECK();
  dim3 grid((unsigned)V_local, (unsigned)cdiv(d, 128));
  NPT_CHECK_CUDA(launch_pdl(sa_accumulate_kernel, dim3(grid), dim3(128), 0, st, offsets, order, d_out, grad, ldG, d));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_window_gather(const uint16_t* tokens, int64_t n_tokens, const int64_t* starts, uint16_t* x, uint16_t* y,
                      int32_t B, int32_t S, void* stream) {
  NPT_REQUIRE(B >= 0 && S > 0 && n_tokens >= (int64_t)S + 2, "window_gather: bad shape B=%d S=%d n_tokens=%lld", B, S,
              (long long)n_tokens);
  if (B == 0) return 0;
  const int64_t n = (int64_t)B * S;
  NPT_CHECK_CUDA(launch_pdl(window_gather_kernel, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, (cudaStream_t)stream, tokens,
                            starts, x, y, B, S));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
