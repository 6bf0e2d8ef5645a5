// Stand-alone elementwise / small reduction kernels:
//   ReLU fwd/bwd (nn/activation.py:38-43), fp32->bf16 cast, column sum (linear.py:57-58),
//   loss mean (train_shakespeare.py:21).
// All are HBM-bound streaming kernels: 128-bit accesses, grid sized in multiples of the SM count.
#include "common.cuh"

namespace npt {

constexpr int kEwThreads = 256;

static unsigned stream_grid(int64_t nvec) {
  int64_t want = cdiv(nvec, kEwThreads);
  int64_t cap = (int64_t)sm_count() * 8;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

__global__ void __launch_bounds__(kEwThreads)
relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                int64_t n, bool vec) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t nv = n >> 2;
    for (int64_t i = tid; i < nv; i += stride) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      st_stream(reinterpret_cast<float4*>(y) + i, v);
      if (yb) st_bf16x4(yb + i * 4, v);
    }
    for (int64_t i = (nv << 2) + tid; i < n; i += stride) {
      const float v = fmaxf(x[i], 0.f);
      y[i] = v;
      if (yb) yb[i] = __float2bfloat16_rn(v);
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const float v = fmaxf(x[i], 0.f);
      y[i] = v;
      if (yb) yb[i] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void __launch_bounds__(kEwThreads)
relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                __nv_bfloat16* __restrict__ dxb, int64_t n, bool vec) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t nv = n >> 2;
    for (int64_t i = tid; i < nv; i += stride) {
      const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
      float4 g = ld_stream(reinterpret_cast<const float4*>(dy) + i);
      g.x = a.x > 0.f ? g.x : 0.f * g.x; g.y = a.y > 0.f ? g.y : 0.f * g.y;
      g.z = a.z > 0.f ? g.z : 0.f * g.z; g.w = a.w > 0.f ? g.w : 0.f * g.w;
      st_stream(reinterpret_cast<float4*>(dx) + i, g);
      if (dxb) st_bf16x4(dxb + i * 4, g);
    }
    for (int64_t i = (nv << 2) + tid; i < n; i += stride) {
      const float g = x[i] > 0.f ? dy[i] : 0.f * dy[i];
      dx[i] = g;
      if (dxb) dxb[i] = __float2bfloat16_rn(g);
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const float g = x[i] > 0.f ? dy[i] : 0.f * dy[i];
      dx[i] = g;
      if (dxb) dxb[i] = __float2bfloat16_rn(g);
    }
  }
}

// rows x cols fp32 (pitch ld_src) -> bf16 (pitch ld_dst), zero-filling [cols, ld_dst).
__global__ void __launch_bounds__(kEwThreads)
cast_bf16_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                 int64_t ld_dst, int64_t rows, int64_t cols, bool vec) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {  // contiguous on both sides, multiple of 4, aligned
    const int64_t nv = (rows * cols) >> 2;
    for (int64_t i = tid; i < nv; i += stride)
      st_bf16x4(dst + i * 4, ld_stream(reinterpret_cast<const float4*>(src) + i));
  } else {
    const int64_t n = rows * ld_dst;
    for (int64_t i = tid; i < n; i += stride) {
      const int64_t r = i / ld_dst, c = i - r * ld_dst;
      dst[i] = __float2bfloat16_rn(c < cols ? src[r * ld_src + c] : 0.f);
    }
  }
}

// Column sums, stage 1: block b sums rows [b*rpb, (b+1)*rpb) for a 32-column strip.
// Threads are (32 columns) x (8 row lanes); coalesced 128 B reads per warp.
__global__ void __launch_bounds__(256)
colsum_stage1_kernel(const void* __restrict__ xv, int x_is_bf16, int64_t ld, float* __restrict__ partial,
                     int64_t rows, int cols, int64_t rows_per_block) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const float* x = reinterpret_cast<const float*>(xv);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv);
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (col < cols)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += x_is_bf16 ? __bfloat162float(xb[r * ld + col]) : x[r * ld + col];
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][cx];
    partial[(size_t)blockIdx.y * cols + col] = t;
  }
}
// Vectorised stage 1: each thread owns 4 adjacent columns (16-byte fp32 / 8-byte bf16 loads),
// block = 32 column groups (128 columns) x 8 row lanes, 8 rows in flight per lane.  The last block to
// finish a 128-column strip (ticket counter per strip) adds the strip's partial rows in fixed order
// and writes the result, so the sum does not depend on which block that is and no second launch is
// needed.  The counters return to zero for the next launch (one stream at a time, like the workspace).
__device__ unsigned int g_colsum_tickets[1024];

template <bool BF16>
__global__ void __launch_bounds__(256)
colsum_stage1_vec_kernel(const void* __restrict__ xv, int64_t ld, float* __restrict__ partial, float* __restrict__ out,
                         int64_t rows, int cols, int64_t rows_per_block) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float4 sm[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int64_t r) -> float4 {
    if (BF16) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(xv) + r * ld + col));
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
      return make_float4(a.x, a.y, b.x, b.y);
    }
    return ld_stream(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv) + r * ld + col));
  };
  if (col < cols) {
    int64_t r = r0 + ry;
    for (; r + 56 < r1; r += 64) {  // 8 rows in flight per thread: the kernel is bound by bytes in flight
      const float4 a = load(r), b = load(r + 8), c = load(r + 16), e = load(r + 24);
      const float4 f = load(r + 32), g = load(r + 40), h = load(r + 48), k = load(r + 56);
      acc.x += ((a.x + b.x) + (c.x + e.x)) + ((f.x + g.x) + (h.x + k.x));
      acc.y += ((a.y + b.y) + (c.y + e.y)) + ((f.y + g.y) + (h.y + k.y));
      acc.z += ((a.z + b.z) + (c.z + e.z)) + ((f.z + g.z) + (h.z + k.z));
      acc.w += ((a.w + b.w) + (c.w + e.w)) + ((f.w + g.w) + (h.w + k.w));
    }
    for (; r + 24 < r1; r += 32) {
      const float4 a = load(r), b = load(r + 8), c = load(r + 16), e = load(r + 24);
      acc.x += (a.x + b.x) + (c.x + e.x); acc.y += (a.y + b.y) + (c.y + e.y);
      acc.z += (a.z + b.z) + (c.z + e.z); acc.w += (a.w + b.w) + (c.w + e.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = load(r);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float4 t = sm[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) { t.x += sm[k][cx].x; t.y += sm[k][cx].y; t.z += sm[k][cx].z; t.w += sm[k][cx].w; }
    *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * cols + col) = t;
  }
  // ---- last block of the strip: final sum over the gridDim.y partial rows ----
  __shared__ unsigned int ticket;
  __threadfence();   // this block's partial row is visible device-wide before its ticket is drawn
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(&g_colsum_tickets[blockIdx.x], 1u);
  __syncthreads();
  if (ticket != gridDim.y - 1) return;
  __threadfence();
  acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < cols)
    for (unsigned b = ry; b < gridDim.y; b += 8) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)b * cols + col));
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float4 t = sm[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) { t.x += sm[k][cx].x; t.y += sm[k][cx].y; t.z += sm[k][cx].z; t.w += sm[k][cx].w; }
    *reinterpret_cast<float4*>(out + col) = t;
  }
  if (threadIdx.x == 0) g_colsum_tickets[blockIdx.x] = 0;
}

// Stage 2: 32 columns x 32 row-lanes per block (independent loads), fixed-order smem tree.
__global__ void __launch_bounds__(1024)
colsum_stage2_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int cols) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float t = 0.f;
  if (c < cols)
    for (int b = ry; b < nparts; b += 32) t += partial[(size_t)b * cols + c];
  sm[ry][cx] = t;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += sm[k][cx];
    out[c] = s;
  }
}

static int colsum_parts(int64_t rows, int cols) {
  const int strips = (int)cdiv(cols, 128);
  int64_t want = cdiv((int64_t)sm_count() * 4, strips);
  int64_t maxp = cdiv(rows, 64);
  if (maxp < 1) maxp = 1;
  if (want > maxp) want = maxp;
  if (want < 1) want = 1;
  return (int)want;
}

// mean of n floats: single block, fixed-order tree -> deterministic.
__global__ void __launch_bounds__(1024)
mean_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = t / (float)n;
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_relu_fwd(const float* x, float* y, void* y_bf16, int64_t n, void* stream) {
  if (n <= 0) return 0;
  const bool vec = aligned16(x) && aligned16(y) && (!y_bf16 || aligned8(y_bf16));
  NPT_CHECK_CUDA(launch_pdl(relu_fwd_kernel, dim3(stream_grid(n / 4 + 1)), dim3(kEwThreads), 0, (cudaStream_t)stream, 
      x, y, (__nv_bfloat16*)y_bf16, n, vec));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_relu_bwd(const float* x, const float* dy, float* dx, void* dx_bf16, int64_t n, void* stream) {
  if (n <= 0) return 0;
  const bool vec = aligned16(x) && aligned16(dy) && aligned16(dx) && (!dx_bf16 || aligned8(dx_bf16));
  NPT_CHECK_CUDA(launch_pdl(relu_bwd_kernel, dim3(stream_grid(n / 4 + 1)), dim3(kEwThreads), 0, (cudaStream_t)stream, 
      x, dy, dx, (__nv_bfloat16*)dx_bf16, n, vec));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                      int64_t cols, void* stream) {
  NPT_REQUIRE(ld_src >= cols && ld_dst >= cols, "cast: leading dimension smaller than cols");
  if (rows <= 0 || cols <= 0) return 0;
  const bool vec = ld_src == cols && ld_dst == cols && ((rows * cols) % 4 == 0) && aligned16(src) && aligned8(dst);
  const int64_t work = vec ? (rows * cols) / 4 : rows * ld_dst;
  NPT_CHECK_CUDA(launch_pdl(cast_bf16_kernel, dim3(stream_grid(work)), dim3(kEwThreads), 0, (cudaStream_t)stream, 
      src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, vec));
  NPT_LAUNCH_CHECK();
  return 0;
}

size_t npt_colsum_workspace_bytes(int64_t rows, int32_t cols) {
  return (size_t)colsum_parts(rows, cols) * (size_t)cols * sizeof(float);
}

int npt_colsum(const void* x, int32_t x_is_bf16, int64_t ld, float* out, int64_t rows, int32_t cols, void* workspace,
               size_t workspace_bytes, void* stream) {
  NPT_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "colsum: bad shape");
  NPT_REQUIRE(workspace && workspace_bytes >= npt_colsum_workspace_bytes(rows, cols),
              "colsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = colsum_parts(rows, cols);
  const int64_t rpb = cdiv(rows, parts);
  const bool vec = cols % 4 == 0 && ld % 4 == 0 && aligned16(workspace) && aligned16(out) && cols <= 1024 * 128 &&
                   (x_is_bf16 ? aligned8(x) : aligned16(x));
  if (vec) {  // single launch: the last block of each strip finishes the sum
    dim3 grid((unsigned)cdiv(cols, 128), (unsigned)parts);
    if (x_is_bf16) NPT_CHECK_CUDA(launch_pdl(colsum_stage1_vec_kernel<true>, dim3(grid), dim3(256), 0, st, x, ld, (float*)workspace, out, rows, cols, rpb));
    else NPT_CHECK_CUDA(launch_pdl(colsum_stage1_vec_kernel<false>, dim3(grid), dim3(256), 0, st, x, ld, (float*)workspace, out, rows, cols, rpb));
    NPT_LAUNCH_CHECK();
    return 0;
  } else {
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)parts);
    NPT_CHECK_CUDA(launch_pdl(colsum_stage1_kernel, dim3(grid), dim3(256), 0, st, x, x_is_bf16, ld, (float*)workspace, rows, cols, rpb));
  }
  NPT_LAUNCH_CHECK();
  NPT_CHECK_CUDA(launch_pdl(colsum_stage2_kernel, dim3((unsigned)cdiv(cols, 32)), dim3(1024), 0, st, (const float*)workspace, out, parts, cols));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_mean(const float* x, float* out, int64_t n, void* stream) {
  NPT_REQUIRE(n > 0, "mean: n must be positive");
  NPT_CHECK_CUDA(launch_pdl(mean_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, x, out, n));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
