// LayerNorm forward/backward — replaces nn/normalization.py:22-52 of the reference.
//
// HBM-bound row kernels: one warp owns one row, the row lives in registers (128-bit loads),
// mean/variance are two-pass (the model's rows ride on a mean ~30-70x their std, SURVEY Q12,
// so E[x^2]-E[x]^2 is not accurate enough), reductions are warp shuffles.
// Algorithmic bytes: fwd 8 B/elem (+4 with residual, +2 with the bf16 shadow), bwd 12 B/elem.
#include "common.cuh"

namespace npt {

constexpr int kLnWarps = 8;      // rows per block
constexpr int kLnMaxVec = 8;     // float4 per lane -> d <= 1024 on the register path

// ---------------------------------------------------------------------------------------
// EXACT: every operation rounded separately like NumPy (IEEE division per element) — the fp32 parity
// mode.  !EXACT (a bf16 output is requested, i.e. the bf16 throughput mode): one reciprocal per row and a
// multiply per element; the IEEE divisions were ~40 % of the instructions these kernels issue.
// MODE 0 = EXACT (fp32 input), 1 = fast (fp32 input), 2 = fast, bf16 input, 3 = 2 + the other tensor-parallel
// rank's partial is added while reading (x — and dy in the backward — come
// from a GEMM epilogue as bf16 alone: half the bytes to write there, to all-reduce under tensor
// parallelism and to read here).
__device__ __forceinline__ float4 ld_bf16x4_stream(const void* p) {
  const uint2 u = __ldcs(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <bool INBF>
__device__ __forceinline__ float4 ln_load4(const void* base, int64_t row, int d, int c) {
  if (INBF) return ld_bf16x4_stream(reinterpret_cast<const __nv_bfloat16*>(base) + row * d + (int64_t)c * 4);
  return ld_stream(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + row * d) + c);
}

// Tensor parallelism over 2 ranks without an all-reduce kernel: `peer` is the other rank's partial
// tensor (a peer-memory pointer over NVLink).  The sum is rounded to bf16 — exactly what a bf16
// all-reduce would have left in memory — so both ranks, and the forward and backward passes, see
// identical values.
// PEER is a template parameter (MODE 3), not a run-time test on the pointer: with the test the
// compiler serialised the loads of a row behind it and every LayerNorm launch got 3-4 us slower.
template <bool INBF, bool PEER>
__device__ __forceinline__ float4 ln_load4_sum(const void* base, const void* peer, int64_t row, int d, int c) {
  float4 a = ln_load4<INBF>(base, row, d, c);
  if (INBF && PEER) {
    const float4 b = ln_load4<INBF>(peer, row, d, c);
    a.x = __bfloat162float(__float2bfloat16_rn(a.x + b.x)); a.y = __bfloat162float(__float2bfloat16_rn(a.y + b.y));
    a.z = __bfloat162float(__float2bfloat16_rn(a.z + b.z)); a.w = __bfloat162float(__float2bfloat16_rn(a.w + b.w));
  }
  return a;
}

template <int NV, int MODE>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_warp_kernel(const void* __restrict__ x, const void* __restrict__ x_peer, __nv_bfloat16* __restrict__ x_sum,
                   const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ residual,
                   float* __restrict__ y, __nv_bfloat16* __restrict__ yb, __nv_bfloat16* __restrict__ yb_peer,
                   float* __restrict__ mean_out, float* __restrict__ var_out,
                   int64_t rows, int d, float eps) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + warp;
  if (row >= rows) return;
  constexpr bool EXACT = MODE == 0, INBF = MODE >= 2, PEER = MODE == 3;
  const int nvec = d >> 2;  // float4 per row
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = ln_load4_sum<INBF, PEER>(x, x_peer, row, d, c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (PEER && x_sum) {  // the reduced input, kept for the backward (after the loads: nothing between them)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) st_bf16x4(x_sum + row * d + (int64_t)c * 4, v[i]);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float var = warp_sum(q) / (float)d;
  const float denom = sqrtf(var + eps);
  const float rden = __frcp_rn(denom);
  if (lane == 0) {
    mean_out[row] = mean;
    var_out[row] = var;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* r4 = residual ? reinterpret_cast<const float4*>(residual + row * d) : nullptr;
  float4* yr = reinterpret_cast<float4*>(y + row * d);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      // gamma * ((x - mean) / sqrt(var+eps)) + beta, each op rounded separately like NumPy.
      if (EXACT) {
        o.x = __fadd_rn(__fmul_rn(g.x, __fdiv_rn(v[i].x - mean, denom)), b.x);
        o.y = __fadd_rn(__fmul_rn(g.y, __fdiv_rn(v[i].y - mean, denom)), b.y);
        o.z = __fadd_rn(__fmul_rn(g.z, __fdiv_rn(v[i].z - mean, denom)), b.z);
        o.w = __fadd_rn(__fmul_rn(g.w, __fdiv_rn(v[i].w - mean, denom)), b.w);
      } else {
        o.x = g.x * ((v[i].x - mean) * rden) + b.x;
        o.y = g.y * ((v[i].y - mean) * rden) + b.y;
        o.z = g.z * ((v[i].z - mean) * rden) + b.z;
        o.w = g.w * ((v[i].w - mean) * rden) + b.w;
      }
      if (r4) {
        const float4 r = ld_stream(r4 + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st_stream(yr + c, o);
      if (yb) st_bf16x4(yb + row * d + (int64_t)c * 4, o);
      if (PEER && yb_peer) st_bf16x4(yb_peer + row * d + (int64_t)c * 4, o);   // all-gather by push (row-sharded LayerNorm)
    }
  }
}

// Wide rows (1024 < d <= 8192, d % 4 == 0): one block of 256 threads per row, the row in registers (NV float4
// per thread), block-wide two-pass statistics.  Same arithmetic as the warp kernel (EXACT = NumPy's rounding).
template <int NV, bool EXACT>
__global__ void __launch_bounds__(256)
ln_fwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ residual, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                   float* __restrict__ mean_out, float* __restrict__ var_out, int64_t rows, int d, float eps) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int nvec = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec) {
      v[i] = ld_stream(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = block_sum(s, red) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
      q += (a * a + b * b) + (e * e + f * f);
    }
  }
  const float var = block_sum(q, red) / (float)d;
  const float denom = sqrtf(var + eps);
  const float rden = __frcp_rn(denom);
  if (threadIdx.x == 0) {
    mean_out[row] = mean;
    var_out[row] = var;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* r4 = residual ? reinterpret_cast<const float4*>(residual + row * d) : nullptr;
  float4* yr = reinterpret_cast<float4*>(y + row * d);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      if (EXACT) {
        o.x = __fadd_rn(__fmul_rn(g.x, __fdiv_rn(v[i].x - mean, denom)), b.x);
        o.y = __fadd_rn(__fmul_rn(g.y, __fdiv_rn(v[i].y - mean, denom)), b.y);
        o.z = __fadd_rn(__fmul_rn(g.z, __fdiv_rn(v[i].z - mean, denom)), b.z);
        o.w = __fadd_rn(__fmul_rn(g.w, __fdiv_rn(v[i].w - mean, denom)), b.w);
      } else {
        o.x = g.x * ((v[i].x - mean) * rden) + b.x;
        o.y = g.y * ((v[i].y - mean) * rden) + b.y;
        o.z = g.z * ((v[i].z - mean) * rden) + b.z;
        o.w = g.w * ((v[i].w - mean) * rden) + b.w;
      }
      if (r4) {
        const float4 r = ld_stream(r4 + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st_stream(yr + c, o);
      if (yb) st_bf16x4(yb + row * d + (int64_t)c * 4, o);
    }
  }
}

// Wide rows, backward: a block walks rows (grid stride); every thread owns the same 4 * NV columns in every row, so
// the dgamma / dbeta partials live in its registers and go straight to partial[block][2][d] at the end.
template <int NV, bool EXACT>
__global__ void __launch_bounds__(256)
ln_bwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ mean_in,
                   const float* __restrict__ var_in, const float* __restrict__ dy, float* __restrict__ dx,
                   __nv_bfloat16* __restrict__ dxb, float* __restrict__ partial, int64_t rows, int d, float eps) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int nvec = d >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    g[i] = (c < nvec) ? __ldg(g4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_d = 1.0f / (float)d;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = mean_in[row], var = var_in[row];
    float denom = sqrtf(var + eps), stdv = sqrtf(var);   // inputs.std(): no eps (normalization.py:48-50)
    if (!EXACT) { denom = __frcp_rn(denom); stdv = __frcp_rn(stdv); }
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
    const float4* dr = reinterpret_cast<const float4*>(dy + row * d);
    float4 xh[NV], w[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < nvec) { xh[i] = ld_stream(xr + c); w[i] = ld_stream(dr + c); }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < nvec) {
        float4& h = xh[i];
        if (EXACT) {
          h.x = __fdiv_rn(h.x - mean, denom); h.y = __fdiv_rn(h.y - mean, denom);
          h.z = __fdiv_rn(h.z - mean, denom); h.w = __fdiv_rn(h.w - mean, denom);
        } else {
          h.x = (h.x - mean) * denom; h.y = (h.y - mean) * denom; h.z = (h.z - mean) * denom; h.w = (h.w - mean) * denom;
        }
        const float4 dd = w[i];
        ag[i].x += dd.x * h.x; ag[i].y += dd.y * h.y; ag[i].z += dd.z * h.z; ag[i].w += dd.w * h.w;
        ab[i].x += dd.x; ab[i].y += dd.y; ab[i].z += dd.z; ab[i].w += dd.w;
        w[i].x = g[i].x * dd.x; w[i].y = g[i].y * dd.y; w[i].z = g[i].z * dd.z; w[i].w = g[i].w * dd.w;
        s1 += (h.x * w[i].x + h.y * w[i].y) + (h.z * w[i].z + h.w * w[i].w);
        s2 += (w[i].x + w[i].y) + (w[i].z + w[i].w);
      }
    }
    const float c1 = block_sum(s1, red) * inv_d;
    const float c2 = block_sum(s2, red) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = threadIdx.x + i * 256;
      if (c < nvec) {
        float4 o;
        if (EXACT) {
          o.x = __fdiv_rn(w[i].x - c1 * xh[i].x - c2, stdv); o.y = __fdiv_rn(w[i].y - c1 * xh[i].y - c2, stdv);
          o.z = __fdiv_rn(w[i].z - c1 * xh[i].z - c2, stdv); o.w = __fdiv_rn(w[i].w - c1 * xh[i].w - c2, stdv);
        } else {
          o.x = (w[i].x - c1 * xh[i].x - c2) * stdv; o.y = (w[i].y - c1 * xh[i].y - c2) * stdv;
          o.z = (w[i].z - c1 * xh[i].z - c2) * stdv; o.w = (w[i].w - c1 * xh[i].w - c2) * stdv;
        }
        if (dx) st_stream(reinterpret_cast<float4*>(dx + row * d) + c, o);
        if (dxb) st_bf16x4(dxb + row * d + (int64_t)c * 4, o);
      }
    }
  }
  float* pg = partial + (size_t)blockIdx.x * 2 * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < nvec) {
      reinterpret_cast<float4*>(pg)[c] = ag[i];
      reinterpret_cast<float4*>(pg + d)[c] = ab[i];
    }
  }
}

// Any d: one block per row, three passes over the (L1/L2-resident) row.
__global__ void __launch_bounds__(256)
ln_fwd_block_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ residual,
                    float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                    float* __restrict__ mean_out, float* __restrict__ var_out,
                    int64_t rows, int d, float eps) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += xr[c];
  const float mean = block_sum(s, red) / (float)d;
  float q = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float a = xr[c] - mean;
    q += a * a;
  }
  const float var = block_sum(q, red) / (float)d;
  const float denom = sqrtf(var + eps);
  if (threadIdx.x == 0) {
    mean_out[row] = mean;
    var_out[row] = var;
  }
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float o = __fadd_rn(__fmul_rn(gamma[c], __fdiv_rn(xr[c] - mean, denom)), beta[c]);
    if (residual) o += residual[row * d + c];
    y[row * d + c] = o;
    if (yb) yb[row * d + c] = __float2bfloat16_rn(o);
  }
}

// ---------------------------------------------------------------------------------------
// Backward stage 1: dx per row + per-block partial dgamma/dbeta.
// DXSUM: also accumulate the column sums of dx (third slice of the partials) — dx is the dy of the
// Linear layer that precedes this LayerNorm, so this is that layer's bias gradient (nn/linear.py:57-58)
// without a separate pass over dx.  The running sums live in shared memory (one slice per warp, each
// lane owns its columns), not in registers: at 126 registers two blocks fit an SM, at 138 only one.
template <int NV, int MODE, bool DXSUM>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_bwd_warp_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ var_in,
                   const void* __restrict__ dy, const void* __restrict__ dy_peer, float* __restrict__ dx,
                   __nv_bfloat16* __restrict__ dxb, __nv_bfloat16* __restrict__ dxb_peer, float* __restrict__ partial,
                   float* __restrict__ partial_peer, int64_t rows, int d, float eps) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  extern __shared__ float sm[];  // [kLnWarps][NSL][d]
  constexpr int NSL = DXSUM ? 3 : 2;
  constexpr bool EXACT = MODE == 0, INBF = MODE >= 2, PEER = MODE == 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4* dsum = reinterpret_cast<float4*>(sm + ((size_t)warp * NSL + 2) * d);  // DXSUM: this warp's running sums
  if (DXSUM) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + i * 32 < nvec) dsum[lane + i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    g[i] = (c < nvec) ? __ldg(g4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_d = 1.0f / (float)d;
  // Two rows per warp iteration: all loads of both rows are issued before either row's
  // reductions, which doubles the bytes in flight per warp (the kernel is latency-bound otherwise).
  const int64_t stride = (int64_t)gridDim.x * kLnWarps;
  for (int64_t row0 = (int64_t)blockIdx.x * kLnWarps + warp; row0 < rows; row0 += 2 * stride) {
    float4 xv[2][NV], dv[2][NV];
    float mean[2], denom[2], stdv[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t row = row0 + u * stride;
      live[u] = row < rows;
      if (live[u]) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + i * 32;
          if (c < nvec) { xv[u][i] = ln_load4<INBF>(x, row, d, c); dv[u][i] = ln_load4_sum<INBF, PEER>(dy, dy_peer, row, d, c); }
        }
        mean[u] = mean_in[row];
        const float var = var_in[row];
        denom[u] = sqrtf(var + eps);
        stdv[u] = sqrtf(var);  // inputs.std(): no eps (normalization.py:48-50)
        if (!EXACT) { denom[u] = __frcp_rn(denom[u]); stdv[u] = __frcp_rn(stdv[u]); }  // reciprocals
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!live[u]) continue;  // warp-uniform
      const int64_t row = row0 + u * stride;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          float4& xh = xv[u][i];  // becomes x-hat in place
          const float4 g4v = g[i], dd = dv[u][i];
          if (EXACT) {
            xh.x = __fdiv_rn(xh.x - mean[u], denom[u]); xh.y = __fdiv_rn(xh.y - mean[u], denom[u]);
            xh.z = __fdiv_rn(xh.z - mean[u], denom[u]); xh.w = __fdiv_rn(xh.w - mean[u], denom[u]);
          } else {
            xh.x = (xh.x - mean[u]) * denom[u]; xh.y = (xh.y - mean[u]) * denom[u];
            xh.z = (xh.z - mean[u]) * denom[u]; xh.w = (xh.w - mean[u]) * denom[u];
          }
          ag[i].x += dd.x * xh.x; ag[i].y += dd.y * xh.y; ag[i].z += dd.z * xh.z; ag[i].w += dd.w * xh.w;
          ab[i].x += dd.x; ab[i].y += dd.y; ab[i].z += dd.z; ab[i].w += dd.w;
          float4& w = dv[u][i];   // becomes gamma*dy in place
          w.x = g4v.x * dd.x; w.y = g4v.y * dd.y; w.z = g4v.z * dd.z; w.w = g4v.w * dd.w;
          s1 += (xh.x * w.x + xh.y * w.y) + (xh.z * w.z + xh.w * w.w);
          s2 += (w.x + w.y) + (w.z + w.w);
        }
      }
      const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
      float4* outr = reinterpret_cast<float4*>(dx + row * d);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          const float4 xh = xv[u][i], w = dv[u][i];
          float4 o;
          if (EXACT) {
            o.x = __fdiv_rn(w.x - c1 * xh.x - c2, stdv[u]);
            o.y = __fdiv_rn(w.y - c1 * xh.y - c2, stdv[u]);
            o.z = __fdiv_rn(w.z - c1 * xh.z - c2, stdv[u]);
            o.w = __fdiv_rn(w.w - c1 * xh.w - c2, stdv[u]);
          } else {
            o.x = (w.x - c1 * xh.x - c2) * stdv[u];
            o.y = (w.y - c1 * xh.y - c2) * stdv[u];
            o.z = (w.z - c1 * xh.z - c2) * stdv[u];
            o.w = (w.w - c1 * xh.w - c2) * stdv[u];
          }
          if (dx) st_stream(outr + c, o);
          if (dxb) st_bf16x4(dxb + row * d + (int64_t)c * 4, o);
          if (PEER && dxb_peer) st_bf16x4(dxb_peer + row * d + (int64_t)c * 4, o);   // all-gather by push
          if (DXSUM) {
            float4 t = dsum[c];
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            dsum[c] = t;
          }
        }
      }
    }
  }
  // combine the block's warps: smem [warp][NSL][d]
  float* mine = sm + (size_t)warp * NSL * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      reinterpret_cast<float4*>(mine)[c] = ag[i];
      reinterpret_cast<float4*>(mine + d)[c] = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NSL * d; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kLnWarps; ++w2) t += sm[(size_t)w2 * NSL * d + c];
    partial[(size_t)blockIdx.x * NSL * d + c] = t;
    if (PEER && partial_peer) partial_peer[(size_t)blockIdx.x * NSL * d + c] = t;
  }
}

// ---------------------------------------------------------------------------------------
// Backward, bf16 inputs (the bf16 mode's step path): STREAMING variant.  The register kernel above keeps two
// rows per warp in flight in registers (126 registers, 16 warps per SM) and issues no loads while it reduces
// and computes: 2.6 TB/s.  Here every warp owns a ring of R row slots in shared memory that one lane fills
// with 1-D bulk async copies (cp.async.bulk, completion on an mbarrier per slot): the loads of the next R
// rows are always in flight, independent of what the warp is computing, and cost no registers.
// Per-warp running sums of dx (DXSUM) stay in shared memory; the warps' dgamma / dbeta partials are added
// into one [NSL][d] block partial in warp order (deterministic).
__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 ln_lds_bf16x4(uint32_t addr) {
  uint2 u;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "r"(addr) : "memory");
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int NV, bool PEER, bool DXSUM>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_bwd_stream_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ var_in,
                     const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy_peer,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb, __nv_bfloat16* __restrict__ dxb_peer,
                     float* __restrict__ partial, float* __restrict__ partial_peer, int64_t rows, int d, float eps, int R) {
  constexpr int NT = PEER ? 3 : 2;       // row tensors per slot: x, dy (, the other rank's dy)
  constexpr int NSL = DXSUM ? 3 : 2;
  constexpr bool TWO_PASS = NV >= 5;     // rows wider than 512 columns: two passes over the slot (see below)
  extern __shared__ __align__(16) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  const uint32_t row_bytes = (uint32_t)d * 2;
  // layout: [warp][R][NT][d] bf16 | [warp][d] fp32 (DXSUM) | [NSL][d] fp32 (block partial) | [warp][R] mbarriers
  uint8_t* ring = ln_smem + (size_t)warp * R * NT * row_bytes;
  float* dsum_base = reinterpret_cast<float*>(ln_smem + (size_t)kLnWarps * R * NT * row_bytes);
  float4* dsum = reinterpret_cast<float4*>(dsum_base + (size_t)warp * d);
  float* comb = dsum_base + (DXSUM ? (size_t)kLnWarps * d : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(comb + (size_t)NSL * d) + warp * R;
  if (lane == 0)
    for (int s2 = 0; s2 < R; ++s2)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ln_smem_u32(bars + s2)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  if (DXSUM) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + i * 32 < nvec) dsum[lane + i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    g[i] = (c < nvec) ? __ldg(g4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();  // programmatic dependent launch: everything above is private set-up
  pdl_trigger();
  const int64_t stride = (int64_t)gridDim.x * kLnWarps;
  const int64_t first = (int64_t)blockIdx.x * kLnWarps + warp;
  auto issue = [&](int slot, int64_t row) {   // lane 0 only
    const uint32_t bar = ln_smem_u32(bars + slot);
    const uint32_t dst = ln_smem_u32(ring + (size_t)slot * NT * row_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NT * row_bytes) : "memory");
    ln_bulk_row(dst, x + row * d, row_bytes, bar);
    ln_bulk_row(dst + row_bytes, dy + row * d, row_bytes, bar);
    if (PEER) ln_bulk_row(dst + 2 * row_bytes, dy_peer + row * d, row_bytes, bar);
  };
  if (lane == 0)
    for (int s2 = 0; s2 < R; ++s2)
      if (first + s2 * stride < rows) issue(s2, first + s2 * stride);
  const float inv_d = 1.0f / (float)d;
  int slot = 0;
  uint32_t phase = 0;
  for (int64_t row = first; row < rows; row += stride) {
    const float mean = mean_in[row], var = var_in[row];
    const float rden = __frcp_rn(sqrtf(var + eps)), rstd = __frcp_rn(sqrtf(var));  // std without eps (normalization.py:48-50)
    {  // wait for this slot's bytes
      const uint32_t bar = ln_smem_u32(bars + slot);
      uint32_t done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      } while (!done);
    }
    const uint32_t base = ln_smem_u32(ring + (size_t)slot * NT * row_bytes);
    // x-hat and gamma * dy of 4 columns, from the row slot
    auto load4 = [&](int c, float4& xv, float4& dd, float4& wv) {
      xv = ln_lds_bf16x4(base + c * 8);
      dd = ln_lds_bf16x4(base + row_bytes + c * 8);
      if (PEER) {
        const float4 pp = ln_lds_bf16x4(base + 2 * row_bytes + c * 8);
        dd.x = __bfloat162float(__float2bfloat16_rn(dd.x + pp.x)); dd.y = __bfloat162float(__float2bfloat16_rn(dd.y + pp.y));
        dd.z = __bfloat162float(__float2bfloat16_rn(dd.z + pp.z)); dd.w = __bfloat162float(__float2bfloat16_rn(dd.w + pp.w));
      }
      xv.x = (xv.x - mean) * rden; xv.y = (xv.y - mean) * rden; xv.z = (xv.z - mean) * rden; xv.w = (xv.w - mean) * rden;
      const float4 gv = TWO_PASS ? __ldg(g4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      wv = dd;
      if (TWO_PASS) { wv.x *= gv.x; wv.y *= gv.y; wv.z *= gv.z; wv.w *= gv.w; }
    };
    auto emit4 = [&](int c, const float4& xv, const float4& wv, float c1, float c2) {
      float4 o;
      o.x = (wv.x - c1 * xv.x - c2) * rstd;
      o.y = (wv.y - c1 * xv.y - c2) * rstd;
      o.z = (wv.z - c1 * xv.z - c2) * rstd;
      o.w = (wv.w - c1 * xv.w - c2) * rstd;
      if (dx) st_stream(reinterpret_cast<float4*>(dx + row * d) + c, o);
      if (dxb) st_bf16x4(dxb + row * d + (int64_t)c * 4, o);
      if (PEER && dxb_peer) st_bf16x4(dxb_peer + row * d + (int64_t)c * 4, o);   // all-gather by push
      if (DXSUM) {
        float4 t = dsum[c];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        dsum[c] = t;
      }
    };
    float s1 = 0.f, s2 = 0.f;
    if (TWO_PASS) {
      // Wide rows (NV >= 5): holding x-hat and gamma * dy of the whole row next to the dgamma / dbeta accumulators
      // costs 157+ registers = one block per SM.  The row sits in shared memory anyway: pass 1 accumulates, pass 2
      // reads it again for dx; gamma comes from L1 instead of registers.  ~100 registers, two blocks per SM.
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          float4 xv, dd, wv;
          load4(c, xv, dd, wv);
          ag[i].x += dd.x * xv.x; ag[i].y += dd.y * xv.y; ag[i].z += dd.z * xv.z; ag[i].w += dd.w * xv.w;
          ab[i].x += dd.x; ab[i].y += dd.y; ab[i].z += dd.z; ab[i].w += dd.w;
          s1 += (xv.x * wv.x + xv.y * wv.y) + (xv.z * wv.z + xv.w * wv.w);
          s2 += (wv.x + wv.y) + (wv.z + wv.w);
        }
      }
      const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          float4 xv, dd, wv;
          load4(c, xv, dd, wv);
          emit4(c, xv, wv, c1, c2);
        }
      }
      __syncwarp();   // every lane is done with the slot: it can be refilled
      if (lane == 0 && row + (int64_t)R * stride < rows) issue(slot, row + (int64_t)R * stride);
      if (++slot == R) { slot = 0; phase ^= 1; }
    } else {
      float4 xh[NV], w[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          float4 xv, dd, wv;
          load4(c, xv, dd, wv);
          ag[i].x += dd.x * xv.x; ag[i].y += dd.y * xv.y; ag[i].z += dd.z * xv.z; ag[i].w += dd.w * xv.w;
          ab[i].x += dd.x; ab[i].y += dd.y; ab[i].z += dd.z; ab[i].w += dd.w;
          const float4 gv = g[i];
          wv.x = gv.x * dd.x; wv.y = gv.y * dd.y; wv.z = gv.z * dd.z; wv.w = gv.w * dd.w;   // gamma * dy
          s1 += (xv.x * wv.x + xv.y * wv.y) + (xv.z * wv.z + xv.w * wv.w);
          s2 += (wv.x + wv.y) + (wv.z + wv.w);
          xh[i] = xv; w[i] = wv;
        }
      }
      const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
      // every lane's shared-memory reads fed the shuffles above: the slot can be refilled
      if (lane == 0 && row + (int64_t)R * stride < rows) issue(slot, row + (int64_t)R * stride);
      if (++slot == R) { slot = 0; phase ^= 1; }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) emit4(c, xh[i], w[i], c1, c2);
      }
    }
  }
  // block partial = sum over the warps, in warp order
  for (int w2 = 0; w2 < kLnWarps; ++w2) {
    if (warp == w2) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
          float4* cg = reinterpret_cast<float4*>(comb) + c;
          float4* cb = reinterpret_cast<float4*>(comb + d) + c;
          float4 a = ag[i], b = ab[i];
          if (w2 > 0) { const float4 pa = *cg, pb = *cb; a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w; b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w; }
          *cg = a; *cb = b;
          if (DXSUM) {
            float4* cs = reinterpret_cast<float4*>(comb + 2 * d) + c;
            float4 t = dsum[c];
            if (w2 > 0) { const float4 ps = *cs; t.x += ps.x; t.y += ps.y; t.z += ps.z; t.w += ps.w; }
            *cs = t;
          }
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < NSL * d; c += blockDim.x) {
    partial[(size_t)blockIdx.x * NSL * d + c] = comb[c];
    if (PEER && partial_peer) partial_peer[(size_t)blockIdx.x * NSL * d + c] = comb[c];
  }
}

// Any d: block per row-group, scalar loops.  partial layout identical.
__global__ void __launch_bounds__(256)
ln_bwd_block_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ mean_in, const float* __restrict__ var_in,
                    const float* __restrict__ dy, float* __restrict__ dx,
                    __nv_bfloat16* __restrict__ dxb, float* __restrict__ partial,
                    int64_t rows, int d, float eps) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  __shared__ float red[32];
  // per-thread column accumulators live in `partial` directly (each block owns its slice,
  // each column is touched by exactly one thread of the block -> no race).
  float* pg = partial + (size_t)blockIdx.x * 2 * d;
  float* pb = pg + d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { pg[c] = 0.f; pb[c] = 0.f; }
  const float inv_d = 1.0f / (float)d;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = mean_in[row], var = var_in[row];
    const float denom = sqrtf(var + eps), stdv = sqrtf(var);
    const float* xr = x + row * d;
    const float* dr = dy + row * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float xh = __fdiv_rn(xr[c] - mean, denom), w = gamma[c] * dr[c];
      s1 += xh * w;
      s2 += w;
      pg[c] += dr[c] * xh;
      pb[c] += dr[c];
    }
    const float c1 = block_sum(s1, red) * inv_d;
    const float c2 = block_sum(s2, red) * inv_d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      const float xh = __fdiv_rn(xr[c] - mean, denom), w = gamma[c] * dr[c];
      const float o = __fdiv_rn(w - c1 * xh - c2, stdv);
      if (dx) dx[row * d + c] = o;
      if (dxb) dxb[row * d + c] = __float2bfloat16_rn(o);
    }
  }
}

// Stage 2: out[c] = ((0 + partial[0][c]) + partial[1][c]) + ...  — block order, one thread per column, sixteen
// loads in flight.  The same additions in the same order as the fused Adam kernel performs when the partials
// are handed to it un-reduced (npt_adam_tensor.g_splits), so both routes give identical bits.
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out0,
                       float* __restrict__ out1, float* __restrict__ out2, int nsl, int nblocks, int d) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, then let the successor set up
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nsl * d) return;
  const size_t pitch = (size_t)nsl * d;
  const float* src = partial + c;
  float acc = 0.f;
  int b = 0;
  for (; b + 16 <= nblocks; b += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = src[(size_t)(b + k) * pitch];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += v[k];
  }
  for (; b < nblocks; ++b) acc += src[(size_t)b * pitch];
  if (c < d) out0[c] = acc; else if (c < 2 * d) out1[c - d] = acc; else out2[c - 2 * d] = acc;
}

static bool ln_bwd_force_register() {   // NUMPITRON_LN_BWD=register: the register-resident kernel at every width (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("NUMPITRON_LN_BWD"); v = (e && e[0] == 'r') ? 1 : 0; }
  return v == 1;
}

static int ln_bwd_blocks(int64_t rows) {
  int64_t want = cdiv(rows, 2 * kLnWarps);
  int64_t cap = (int64_t)sm_count() * 2;
  return (int)(want < cap ? want : cap);
}

template <int NV>
static int launch_fwd(const void* x, bool in_bf16, const void* xp, void* xs, const float* g, const float* b, const float* r, float* y,
                      void* yb, void* ybp, float* mean, float* var, int64_t rows, int d, float eps,
                      cudaStream_t st) {
  const int64_t grid = cdiv(rows, kLnWarps);
  if (in_bf16 && xp)
    NPT_CHECK_CUDA(launch_pdl(ln_fwd_warp_kernel<NV, 3>, dim3((unsigned)grid), dim3(kLnWarps * 32), 0, st,
        x, xp, (__nv_bfloat16*)xs, g, b, r, y, (__nv_bfloat16*)yb, (__nv_bfloat16*)ybp, mean, var, rows, d, eps));
  else if (in_bf16)
    NPT_CHECK_CUDA(launch_pdl(ln_fwd_warp_kernel<NV, 2>, dim3((unsigned)grid), dim3(kLnWarps * 32), 0, st,
        x, xp, (__nv_bfloat16*)xs, g, b, r, y, (__nv_bfloat16*)yb, (__nv_bfloat16*)ybp, mean, var, rows, d, eps));
  else if (yb)  // bf16 throughput mode
    NPT_CHECK_CUDA(launch_pdl(ln_fwd_warp_kernel<NV, 1>, dim3((unsigned)grid), dim3(kLnWarps * 32), 0, st,
        x, xp, (__nv_bfloat16*)xs, g, b, r, y, (__nv_bfloat16*)yb, (__nv_bfloat16*)ybp, mean, var, rows, d, eps));
  else
    NPT_CHECK_CUDA(launch_pdl(ln_fwd_warp_kernel<NV, 0>, dim3((unsigned)grid), dim3(kLnWarps * 32), 0, st,
        x, xp, (__nv_bfloat16*)xs, g, b, r, y, (__nv_bfloat16*)yb, (__nv_bfloat16*)ybp, mean, var, rows, d, eps));
  NPT_LAUNCH_CHECK();
  return 0;
}

struct LnPeerOut { void* dxb_peer; float* partial_peer; };   // row-sharded mode: second (peer-GPU) destinations

template <int NV, int MODE, bool DXSUM>
static int launch_bwd_k(const void* x, const float* g, const float* mean, const float* var, const void* dy, const void* dyp, float* dx,
                        void* dxb, float* partial, int nblocks, int64_t rows, int d, float eps, const LnPeerOut& po, cudaStream_t st) {
  const size_t smem = (size_t)kLnWarps * (DXSUM ? 3 : 2) * d * sizeof(float);
  auto kern = ln_bwd_warp_kernel<NV, MODE, DXSUM>;
  NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NPT_CHECK_CUDA(launch_pdl(kern, dim3(nblocks), dim3(kLnWarps * 32), smem, st, x, g, mean, var, dy, dyp, dx,
                            (__nv_bfloat16*)dxb, (__nv_bfloat16*)po.dxb_peer, partial, po.partial_peer, rows, d, eps));
  NPT_LAUNCH_CHECK();
  return 0;
}

static int ln_stream_blocks(int64_t rows) {
  const int64_t want = cdiv(rows, kLnWarps), cap = (int64_t)sm_count() * 2;
  return (int)(want < cap ? want : cap);
}
static size_t ln_stream_smem(int d, int R, bool peer, bool dxsum) {
  return (size_t)kLnWarps * R * (peer ? 3 : 2) * d * 2 + (dxsum ? (size_t)kLnWarps * d * 4 : 0) + (size_t)(dxsum ? 3 : 2) * d * 4 +
         (size_t)kLnWarps * R * 8;
}
static int ln_stream_slots(int d, bool peer) {
  // about 6 KB of rows in flight per warp (48 KB per block), at least 2 slots
  const int per = (peer ? 3 : 2) * d * 2;
  int R = 6144 / per;
  return R < 2 ? 2 : (R > 6 ? 6 : R);
}

template <int NV, bool PEER, bool DXSUM>
static int launch_bwd_stream_k(const void* x, const float* g, const float* mean, const float* var, const void* dy, const void* dyp,
                               float* dx, void* dxb, float* partial, int* nblocks, int64_t rows, int d, float eps, const LnPeerOut& po, cudaStream_t st) {
  const int R = ln_stream_slots(d, PEER);
  const size_t smem = ln_stream_smem(d, R, PEER, DXSUM);
  auto kern = ln_bwd_stream_kernel<NV, PEER, DXSUM>;
  NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  *nblocks = ln_stream_blocks(rows);   // the partial workspace holds 2 blocks per SM
  NPT_CHECK_CUDA(launch_pdl(kern, dim3(*nblocks), dim3(kLnWarps * 32), smem, st, (const __nv_bfloat16*)x, g, mean, var,
                            (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dyp, dx, (__nv_bfloat16*)dxb, (__nv_bfloat16*)po.dxb_peer, partial,
                            po.partial_peer, rows, d, eps, R));
  NPT_LAUNCH_CHECK();
  return 0;
}

template <int NV>
static int launch_bwd_stream(const void* x, const float* g, const float* mean, const float* var, const void* dy, const void* dyp,
                             float* dx, void* dxb, float* partial, bool dxsum, int* nblocks, int64_t rows, int d, float eps, const LnPeerOut& po, cudaStream_t st) {
  if (dyp) return dxsum ? launch_bwd_stream_k<NV, true, true>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st)
                        : launch_bwd_stream_k<NV, true, false>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st);
  return dxsum ? launch_bwd_stream_k<NV, false, true>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st)
               : launch_bwd_stream_k<NV, false, false>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st);
}

template <int NV>
static int launch_bwd(const void* x, const float* g, const float* mean, const float* var,
                      const void* dy, const void* dyp, bool in_bf16, float* dx, void* dxb, float* partial, bool dxsum, int nblocks,
                      int64_t rows, int d, float eps, const LnPeerOut& po, cudaStream_t st) {
#define NPT_LN_BWD(MODE)                                                                                        \
  return dxsum ? launch_bwd_k<NV, MODE, true>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st) \
               : launch_bwd_k<NV, MODE, false>(x, g, mean, var, dy, dyp, dx, dxb, partial, nblocks, rows, d, eps, po, st)
  if (in_bf16 && dyp) { NPT_LN_BWD(3); }
  if (in_bf16) { NPT_LN_BWD(2); }
  if (dxb) { NPT_LN_BWD(1); }  // bf16 throughput mode
  NPT_LN_BWD(0);
#undef NPT_LN_BWD
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_layernorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual,
                      float* y, void* y_bf16, float* mean, float* var, int64_t rows, int32_t d,
                      float eps, void* stream) {
  return npt_layernorm_fwd_v2(x, 0, nullptr, nullptr, gamma, beta, residual, y, y_bf16, mean, var, rows, d, eps, stream);
}

int npt_layernorm_bf16_input_supported(int32_t d) { return (d % 4 == 0 && d <= kLnMaxVec * 128) ? 1 : 0; }

int npt_layernorm_fwd_v2(const void* x, int32_t x_is_bf16, const void* x_peer, void* x_sum_bf16, const float* gamma,
                         const float* beta, const float* residual,
                         float* y, void* y_bf16, float* mean, float* var, int64_t rows, int32_t d,
                         float eps, void* stream) {
  return npt_layernorm_fwd_v3(x, x_is_bf16, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, nullptr, mean, var, rows, d, eps, stream);
}

int npt_layernorm_fwd_v3(const void* x, int32_t x_is_bf16, const void* x_peer, void* x_sum_bf16, const float* gamma,
                         const float* beta, const float* residual,
                         float* y, void* y_bf16, void* y_bf16_peer, float* mean, float* var, int64_t rows, int32_t d,
                         float eps, void* stream) {
  NPT_REQUIRE(!y_bf16_peer || (x_peer && y_bf16 && aligned8(y_bf16_peer)), "layernorm_fwd: y_bf16_peer needs x_peer and y_bf16");
  NPT_REQUIRE(rows >= 0 && d > 0, "layernorm_fwd: bad shape rows=%lld d=%d", (long long)rows, d);
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (d % 4 == 0) && d <= kLnMaxVec * 128 && (x_is_bf16 ? aligned8(x) : aligned16(x)) && aligned16(y) &&
                   aligned16(gamma) && aligned16(beta) && (!residual || aligned16(residual)) &&
                   (!y_bf16 || aligned8(y_bf16));
  NPT_REQUIRE(!x_peer || (x_is_bf16 && aligned8(x_peer) && (!x_sum_bf16 || aligned8(x_sum_bf16))),
              "layernorm_fwd: a peer partial needs bf16 inputs and 8-byte aligned buffers");
  NPT_REQUIRE(!x_is_bf16 || vec, "layernorm_fwd: a bf16 input needs d %% 4 == 0, d <= %d and aligned buffers", kLnMaxVec * 128);
  if (vec) {
    const int nv = (int)cdiv(d, 128);
    const bool ib = x_is_bf16 != 0;
    switch (nv) {
      case 1: return launch_fwd<1>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
      case 2: return launch_fwd<2>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
      case 3: return launch_fwd<3>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
      case 4: return launch_fwd<4>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
      case 5: case 6: return launch_fwd<6>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
      default: return launch_fwd<8>(x, ib, x_peer, x_sum_bf16, gamma, beta, residual, y, y_bf16, y_bf16_peer, mean, var, rows, d, eps, st);
    }
  }
  if (d % 4 == 0 && d <= 8192 && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) &&
      (!residual || aligned16(residual)) && (!y_bf16 || aligned8(y_bf16))) {   // wide rows: block per row, row in registers
    const int nv = (int)cdiv(d, 1024);
    const bool exact = y_bf16 == nullptr;   // a bf16 copy is requested only in the bf16 throughput mode
#define NPT_LN_WIDE(NVV)                                                                                                     \
    if (exact) NPT_CHECK_CUDA(launch_pdl(ln_fwd_wide_kernel<NVV, true>, dim3((unsigned)rows), dim3(256), 0, st, (const float*)x, gamma, \
                                         beta, residual, y, (__nv_bfloat16*)y_bf16, mean, var, rows, d, eps));              \
    else       NPT_CHECK_CUDA(launch_pdl(ln_fwd_wide_kernel<NVV, false>, dim3((unsigned)rows), dim3(256), 0, st, (const float*)x, gamma, \
                                         beta, residual, y, (__nv_bfloat16*)y_bf16, mean, var, rows, d, eps))
    if (nv <= 2) { NPT_LN_WIDE(2); } else if (nv <= 4) { NPT_LN_WIDE(4); } else { NPT_LN_WIDE(8); }
#undef NPT_LN_WIDE
    NPT_LAUNCH_CHECK();
    return 0;
  }
  NPT_CHECK_CUDA(launch_pdl(ln_fwd_block_kernel, dim3((unsigned)rows), dim3(256), 0, st, (const float*)x, gamma, beta, residual, y,
                            (__nv_bfloat16*)y_bf16, mean, var, rows, d, eps));
  NPT_LAUNCH_CHECK();
  return 0;
}

size_t npt_layernorm_bwd_workspace_bytes(int64_t rows, int32_t d) {
  // sized for the larger of the two grid choices
  int64_t nb = (int64_t)sm_count() * 2;
  return (size_t)nb * 3 * (size_t)d * sizeof(float);  // room for the optional dx column sums
}

int npt_layernorm_bwd(const float* x, const float* gamma, const float* mean, const float* var,
                      const float* dy, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                      int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return npt_layernorm_bwd_v2(x, dy, 0, nullptr, gamma, mean, var, dx, dx_bf16, dgamma, dbeta, nullptr, rows, d, eps, workspace,
                              workspace_bytes, stream);
}

int npt_layernorm_bwd_colsum_supported(int32_t d) { return (d % 4 == 0 && d <= kLnMaxVec * 128) ? 1 : 0; }

int npt_layernorm_bwd_v2(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer, const float* gamma, const float* mean,
                         const float* var, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum,
                         int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return npt_layernorm_bwd_v3(x, dy, inputs_are_bf16, dy_peer, gamma, mean, var, dx, dx_bf16, dgamma, dbeta, dx_colsum, rows, d, eps,
                              workspace, workspace_bytes, 0, nullptr, stream);
}

int npt_layernorm_bwd_v3(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer, const float* gamma, const float* mean,
                         const float* var, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum,
                         int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes,
                         int32_t defer_reduce, int32_t* n_partials, void* stream) {
  return npt_layernorm_bwd_v4(x, dy, inputs_are_bf16, dy_peer, gamma, mean, var, dx, dx_bf16, nullptr, dgamma, dbeta, dx_colsum, rows, d,
                              eps, workspace, workspace_bytes, nullptr, defer_reduce, n_partials, stream);
}

// Which first-stage kernel runs, and with how many blocks (= per-block partials left in the workspace).
static bool ln_bwd_streams(int32_t d, bool ib) {
  return (d % 4 == 0) && d <= kLnMaxVec * 128 && ib && d > 512 && d % 8 == 0 && !ln_bwd_force_register();
}
int32_t npt_layernorm_bwd_partials(int64_t rows, int32_t d, int32_t inputs_are_bf16) {
  if (rows <= 0 || d <= 0) return 0;
  if (ln_bwd_streams(d, inputs_are_bf16 != 0)) return ln_stream_blocks(rows);
  if (d % 4 == 0 && d <= kLnMaxVec * 128) return ln_bwd_blocks(rows);
  const int64_t cap = (int64_t)sm_count() * 2;
  return (int32_t)(rows < cap ? rows : cap);
}

int npt_layernorm_bwd_reduce(const float* partials, int32_t n_partials, int32_t n_slices, int32_t d, float* dgamma, float* dbeta,
                             float* dx_colsum, void* stream) {
  NPT_REQUIRE(partials && n_partials > 0 && (n_slices == 2 || n_slices == 3) && d > 0 && dgamma && dbeta && (n_slices == 2 || dx_colsum),
              "layernorm_bwd_reduce: bad arguments");
  NPT_CHECK_CUDA(launch_pdl(reduce_partials_kernel, dim3((unsigned)cdiv(n_slices * (int64_t)d, 128)), dim3(128), 0, (cudaStream_t)stream,
                            partials, dgamma, dbeta, dx_colsum, (int)n_slices, (int)n_partials, (int)d));
  NPT_LAUNCH_CHECK();
  return 0;
}

int npt_layernorm_bwd_v4(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer, const float* gamma, const float* mean,
                         const float* var, float* dx, void* dx_bf16, void* dx_bf16_peer, float* dgamma, float* dbeta, float* dx_colsum,
                         int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes, void* workspace_peer,
                         int32_t defer_reduce, int32_t* n_partials, void* stream) {
  NPT_REQUIRE(rows > 0 && d > 0, "layernorm_bwd: bad shape rows=%lld d=%d", (long long)rows, d);
  NPT_REQUIRE((!dx_bf16_peer && !workspace_peer) || (dy_peer && dx_bf16 && defer_reduce),
              "layernorm_bwd: peer outputs need dy_peer, dx_bf16 and defer_reduce (the caller reduces both ranks' partials)");
  const LnPeerOut po{dx_bf16_peer, (float*)workspace_peer};
  NPT_REQUIRE(dx || dx_bf16, "layernorm_bwd: no output");
  const bool ib = inputs_are_bf16 != 0;
  {
    const size_t need = (size_t)npt_layernorm_bwd_partials(rows, d, inputs_are_bf16) * (dx_colsum ? 3 : 2) * (size_t)d * sizeof(float);
    NPT_REQUIRE(workspace && workspace_bytes >= need, "layernorm_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  int nblocks;
  const bool vec = (d % 4 == 0) && d <= kLnMaxVec * 128 && (ib ? aligned8(x) && aligned8(dy) : aligned16(x) && aligned16(dy)) &&
                   (!dx || aligned16(dx)) && aligned16(gamma) && (!dx_bf16 || aligned8(dx_bf16)) &&
                   (size_t)kLnWarps * 3 * d * sizeof(float) <= 200 * 1024;
  const bool dxsum = dx_colsum != nullptr;
  NPT_REQUIRE(!dy_peer || (ib && aligned8(dy_peer)), "layernorm_bwd: a peer partial needs bf16 inputs and an 8-byte aligned buffer");
  NPT_REQUIRE(!(dxsum || ib) || vec,
              "layernorm_bwd: fused dx column sums / bf16 inputs need d %% 4 == 0, d <= %d and aligned buffers", kLnMaxVec * 128);
  // measured (benchmarks/kernbench.py): rows of <= 512 columns fit the register kernel's two-rows-in-flight scheme
  // (18 us vs 22 us at 16384 x 384); wider rows run 1.6-2x faster on the streaming kernel (32768 x 768: 76 us vs 145 us)
  const bool stream_ok = vec && ln_bwd_streams(d, ib) && aligned16(x) && aligned16(dy) && (!dy_peer || aligned16(dy_peer));
  if (stream_ok) {
    const int nv = (int)cdiv(d, 128);
    int rc;
    switch (nv) {
      case 1: rc = launch_bwd_stream<1>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
      case 2: rc = launch_bwd_stream<2>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
      case 3: rc = launch_bwd_stream<3>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
      case 4: rc = launch_bwd_stream<4>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
      case 5: case 6: rc = launch_bwd_stream<6>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
      default: rc = launch_bwd_stream<8>(x, gamma, mean, var, dy, dy_peer, dx, dx_bf16, partial, dxsum, &nblocks, rows, d, eps, po, st); break;
    }
    if (rc) return rc;
  } else if (vec) {
    nblocks = ln_bwd_blocks(rows);
    const int nv = (int)cdiv(d, 128);
    int rc;
    switch (nv) {
      case 1: rc = launch_bwd<1>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
      case 2: rc = launch_bwd<2>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
      case 3: rc = launch_bwd<3>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
      case 4: rc = launch_bwd<4>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
      case 5: case 6: rc = launch_bwd<6>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
      default: rc = launch_bwd<8>(x, gamma, mean, var, dy, dy_peer, ib, dx, dx_bf16, partial, dxsum, nblocks, rows, d, eps, po, st); break;
    }
    if (rc) return rc;
  } else {
    int64_t cap = (int64_t)sm_count() * 2;
    nblocks = (int)(rows < cap ? rows : cap);
    if (d % 4 == 0 && d <= 8192 && aligned16(x) && aligned16(dy) && aligned16(gamma) && (!dx || aligned16(dx)) &&
        (!dx_bf16 || aligned8(dx_bf16))) {   // wide rows: column-owning threads, partials in registers
      const int nv = (int)cdiv(d, 1024);
      const bool exact = dx_bf16 == nullptr;
#define NPT_LN_WIDE(NVV)                                                                                                       \
      if (exact) NPT_CHECK_CUDA(launch_pdl(ln_bwd_wide_kernel<NVV, true>, dim3(nblocks), dim3(256), 0, st, (const float*)x, gamma, mean, \
                                           var, (const float*)dy, dx, (__nv_bfloat16*)dx_bf16, partial, rows, d, eps));         \
      else       NPT_CHECK_CUDA(launch_pdl(ln_bwd_wide_kernel<NVV, false>, dim3(nblocks), dim3(256), 0, st, (const float*)x, gamma, mean, \
                                           var, (const float*)dy, dx, (__nv_bfloat16*)dx_bf16, partial, rows, d, eps))
      if (nv <= 2) { NPT_LN_WIDE(2); } else if (nv <= 4) { NPT_LN_WIDE(4); } else { NPT_LN_WIDE(8); }
#undef NPT_LN_WIDE
    } else {
      NPT_CHECK_CUDA(launch_pdl(ln_bwd_block_kernel, dim3(nblocks), dim3(256), 0, st, (const float*)x, gamma, mean, var,
                                (const float*)dy, dx, (__nv_bfloat16*)dx_bf16, partial, rows, d, eps));
    }
    NPT_LAUNCH_CHECK();
  }
  const int nsl = dxsum ? 3 : 2;
  if (n_partials) *n_partials = nblocks;
  if (defer_reduce) return 0;   // the caller hands partial[nblocks][nsl][d] to Adam (g_splits = nblocks, stride nsl * d)
  NPT_CHECK_CUDA(launch_pdl(reduce_partials_kernel, dim3((unsigned)cdiv(nsl * (int64_t)d, 128)), dim3(128), 0, st, partial,
                            dgamma, dbeta, dx_colsum, nsl, nblocks, d));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
