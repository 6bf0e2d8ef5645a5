// Fused causal attention for sm_100a (bf16 mode): forward and backward without ever writing the
// (B,h,S,S) score / probability matrices to HBM.  Replaces, for head_dim 64 or 96 and S % 128 == 0,
//   forward : q@k^T / sqrt(d_model) -> causal mask -> softmax -> p@v        nn/attention.py:43-52
//   backward: dV = p^T@dO, dP = dO@v^T, softmax backward, mask, scale,
//             dQ = dS@k, dK = dS^T@q, concat([dq,dk,dv])                      nn/attention.py:81-98
// operating directly on the (B,S,h,3,Dh) QKV-projection layout ([q_h|k_h|v_h] per head, Q4) and
// the merged (B,S,h*Dh) head output.  The score scale is the caller's `divisor` (= sqrt(d_model),
// NOT sqrt(head_dim): Q3).
//
// All three kernels are PERSISTENT (one CTA per SM walks a static list of work items) and share one
// shape: warp 0 = TMA producer, warp 1 = tcgen05 MMA issuer, warps 2-5 = one thread per query (or
// key) row = TMEM lane doing the softmax arithmetic.  128x128 score tiles live in TMEM; P / dS go
// through 128B-swizzled shared memory as the A operand of the second GEMM — read K-major (P@V, dS@K)
// or MN-major (P^T@dO, dS^T@Q) from the same bytes; Q / dO / K / V tiles are likewise read K-major
// or MN-major from the same TMA box.  Loads run two tiles ahead and the MMA warp issues the next
// score tile before the current tile's second GEMM, so TMA, tensor core and the softmax threads
// overlap across tiles and across work items.
//   fwd : item = (b, head, 128 queries); steps = key tiles 0..diag (online softmax, fp32 running
//         max / sum, O accumulated in registers); writes O (bf16) and the row log-sum-exp.
//   dkv : item = (b, head, 128 keys); steps = query tiles diag..last; dK, dV accumulate in TMEM.
//   dq  : item = (b, head, 128 queries); steps = key tiles 0..diag; dQ accumulates in TMEM.
// No atomics: every output element has exactly one writer and a fixed summation order.
//
// head_dim is a template parameter.  A {128 rows x Dh} operand tile is NCH column chunks of CH columns, each
// chunk one TMA box with rows of CH * 2 bytes = one swizzle row: Dh = 64 -> one 128-byte-swizzled chunk;
// Dh = 96 (GPT-2 small under tp = 8: ONE local head of 96, nn/attention.py:37-45) -> three
// 64-byte-swizzled chunks of 32 columns.  The same bytes are read K-major (contraction over Dh: chunk by
// chunk, 16 columns per MMA) or MN-major (Dh is the N extent: one MMA spans the chunks, LBO = chunk pitch).
#include "tc_ptx.cuh"

namespace npt {
using namespace tc;

constexpr int AT_THREADS = 192;
constexpr int AT_PTILE = 128 * 128 * 2;  // a 128 x 128 bf16 P / dS tile: two 64-column halves, 128-byte swizzle

template <int DH>
struct AttnCfg {
  static_assert(DH == 64 || DH == 96, "fused attention: head_dim 64 or 96");
  static constexpr int CH = (DH % 64 == 0) ? 64 : 32;      // columns per chunk (= TMA box width)
  static constexpr int NCH = DH / CH;
  static constexpr int RB = CH * 2;                          // bytes per swizzle row
  static constexpr int CHUNK = 128 * RB;                     // bytes per chunk
  static constexpr int TILE = NCH * CHUNK;                   // 128 x DH bf16
  static constexpr uint32_t LAYOUT = (CH == 64) ? 2u : 4u;   // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static constexpr uint32_t SBO = 8 * RB;                    // 8-row group pitch
};
// K-major read of a {128 x DH} tile, 16-column step kk of the contraction over DH
template <int DH>
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int kk) {
  using C = AttnCfg<DH>;
  const int e = kk * 16;
  return make_smem_desc_l(tile + (e / C::CH) * C::CHUNK, 16, C::SBO, C::LAYOUT) + (uint64_t)(((e % C::CH) * 2) >> 4);
}
// MN-major read (DH = the N extent, rows = the contraction index), 16-row step kk
template <int DH>
__device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int kk) {
  using C = AttnCfg<DH>;
  return make_smem_desc_l(tile, C::CHUNK, C::SBO, C::LAYOUT) + (uint64_t)(kk * ((16 * C::RB) >> 4));
}
template <int DH>
__device__ __forceinline__ void load_tile(const CUtensorMap* map, uint64_t* bar, uint8_t* dst, int col, int row) {
  using C = AttnCfg<DH>;
#pragma unroll
  for (int c = 0; c < C::NCH; ++c) tma_load_4d(map, bar, dst + c * C::CHUNK, col + c * C::CH, row, 0, 0);
}

__device__ __forceinline__ uint32_t pack2(float a, float b) { return pack_bf16(a, b); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Write 32 consecutive bf16 values (cols c*32 .. c*32+31 of a 128-wide row) of `row` into a
// [2 halves][128 rows][128 B] swizzled tile (half = 64 columns); `tile` is a shared-space address.
__device__ __forceinline__ void store_row_chunk(uint32_t tile, int row, int c, const float (&v)[32]) {
  const uint32_t base = tile + (c >> 1) * 16384;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + swz128(row, (c & 1) * 4 + q)),
                 "r"(pack2(v[8 * q], v[8 * q + 1])), "r"(pack2(v[8 * q + 2], v[8 * q + 3])),
                 "r"(pack2(v[8 * q + 4], v[8 * q + 5])), "r"(pack2(v[8 * q + 6], v[8 * q + 7])) : "memory");
  }
}

// Walks the (item, step) sequence of one persistent CTA.  Items are (bh, tile) pairs, tile fastest.
struct Cursor {
  int w, total, stride, nqt;
  int bh, tile;       // current item
  int j, n_steps;     // step inside the item
  uint32_t t, wi;     // global step / item counters (barrier phases)
  bool desc;          // tile order inside a bh: descending (heavy query tiles first) or ascending
  bool key_items;     // steps of an item: (nqt - tile) query tiles [dkv] or (tile + 1) key tiles [fwd, dq]
  __device__ void load() {
    // Heaviest tiles of ALL (b, head) pairs first, then the next lighter ones (longest-processing-
    // time order): a static stride-gridDim walk then gives every CTA the same mix of heavy and light
    // items instead of pinning the 2-step items on the even CTAs.
    const int n_bh = total / nqt;
    const int r = w / n_bh;
    bh = w - r * n_bh;
    tile = desc ? nqt - 1 - r : r;
    n_steps = key_items ? nqt - tile : tile + 1;
    j = 0;
  }
  __device__ void init(int first, int total_, int stride_, int nqt_, bool desc_, bool key_items_) {
    w = first; total = total_; stride = stride_; nqt = nqt_; desc = desc_; key_items = key_items_;
    t = 0; wi = 0;
    if (valid()) load();
  }
  __device__ bool valid() const { return w < total; }
  __device__ bool first_step() const { return j == 0; }
  __device__ bool last_step() const { return j == n_steps - 1; }
  __device__ void next() {
    ++t;
    if (++j == n_steps) {
      w += stride;
      ++wi;
      if (valid()) load();
    }
  }
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int DH>
struct AttnFwdSmem {
  static constexpr int T = AttnCfg<DH>::TILE;
  static constexpr int Q = 0, K = 2 * T, V = 4 * T, P = 6 * T, BAR = 6 * T + 2 * AT_PTILE;
  static constexpr int TOTAL = BAR + 24 * 8 + 16 + 1024;
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out,
                float* __restrict__ lse, int B, int S, int h, float inv_div) {
  using L = AttnFwdSmem<DH>;
  constexpr int AT_TILE = L::T;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t *q_full = bars, *q_empty = bars + 2, *k_full = bars + 4, *v_full = bars + 6, *kv_empty = bars + 8,
           *s_full = bars + 10, *s_empty = bars + 12, *p_full = bars + 14, *pv_done = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = S / 128;
  const int total = B * h * nqt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full + s, 1); mbar_init(q_empty + s, 1);
      mbar_init(k_full + s, 1); mbar_init(v_full + s, 1); mbar_init(kv_empty + s, 1);
      mbar_init(s_full + s, 1); mbar_init(s_empty + s, 4);
      mbar_init(p_full + s, 4); mbar_init(pv_done + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();     // set-up above is private; global memory is touched only from here on
  pdl_trigger();  // persistent grid: the next kernel may be scheduled behind this one
  // TMEM columns: S[2] at 0 / 128, O[2] at 256 / 256 + DH

  if (warp == 0) {
    if (lane == 0) {
      Cursor c;
      for (c.init(blockIdx.x, total, gridDim.x, nqt, true, false); c.valid(); c.next()) {
        const int b = c.bh / h, hh = c.bh - b * h;
        const int col_q = hh * 3 * DH, row0 = b * S;
        if (c.first_step()) {
          const int qb = c.wi & 1;
          mbar_wait(q_empty + qb, ((c.wi >> 1) & 1) ^ 1);
          mbar_expect_tx(q_full + qb, AT_TILE);
          load_tile<DH>(&tmQKV, q_full + qb, smem + L::Q + qb * AT_TILE, col_q, row0 + c.tile * 128);
        }
        const int st = c.t & 1;
        mbar_wait(kv_empty + st, ((c.t >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full + st, AT_TILE);
        load_tile<DH>(&tmQKV, k_full + st, smem + L::K + st * AT_TILE, col_q + DH, row0 + c.j * 128);
        mbar_expect_tx(v_full + st, AT_TILE);
        load_tile<DH>(&tmQKV, v_full + st, smem + L::V + st * AT_TILE, col_q + 2 * DH, row0 + c.j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, false, false);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc(DH, false, true);    // O = P V   : P K-major, V MN-major
      auto issue_qk = [&](const Cursor& c) {
        const int st = c.t & 1, ph = (c.t >> 1) & 1, qb = c.wi & 1;
        if (c.first_step()) mbar_wait(q_full + qb, (c.wi >> 1) & 1);
        mbar_wait(k_full + st, ph);
        mbar_wait(s_empty + st, ph ^ 1);  // softmax threads have drained this S buffer (two steps ago)
        fence_after();
        const uint32_t tq = smem_u32(smem + L::Q + qb * AT_TILE), tk = smem_u32(smem + L::K + st * AT_TILE);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem + st * 128, desc_k<DH>(tq, k), desc_k<DH>(tk, k), idesc_s, k > 0);
        umma_commit(s_full + st);
        if (c.last_step()) umma_commit(q_empty + qb);
      };
      Cursor cq, cp;
      cq.init(blockIdx.x, total, gridDim.x, nqt, true, false);
      cp = cq;
      if (cq.valid()) { issue_qk(cq); cq.next(); }
      for (; cp.valid(); cp.next()) {
        if (cq.valid()) { issue_qk(cq); cq.next(); }  // next score tile runs while this tile's softmax is in flight
        const int st = cp.t & 1, ph = (cp.t >> 1) & 1;
        mbar_wait(v_full + st, ph);
        mbar_wait(p_full + st, ph);
        fence_after();
        const uint32_t sp = smem_u32(smem + L::P + st * AT_PTILE);
        const uint32_t tv = smem_u32(smem + L::V + st * AT_TILE);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 128 keys = 8 x 16
          const uint64_t dp = make_smem_desc(sp + (k >> 2) * 16384, 16, 1024) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + 256 + st * DH, dp, desc_mn<DH>(tv, k), idesc_o, k > 0);
        }
        umma_commit(pv_done + st);
        umma_commit(kv_empty + st);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o[DH];
    const float k2 = inv_div * 1.4426950408889634f;  // log2(e) / sqrt(d_model)
    Cursor c;
    for (c.init(blockIdx.x, total, gridDim.x, nqt, true, false); c.valid(); c.next()) {
      const int st = c.t & 1, ph = (c.t >> 1) & 1;
      const bool diag = c.last_step();  // the last key tile of a query tile is its diagonal tile
      if (c.first_step()) {
        m_run = -INFINITY; l_run = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) o[d] = 0.f;
      }
      const uint32_t tS = tmem + st * 128 + lane_addr, tO = tmem + 256 + st * DH + lane_addr;
      const uint32_t sP = smem_u32(smem + L::P + st * AT_PTILE);
      mbar_wait(s_full + st, ph);
      fence_after();
      // The whole 128-wide score row goes to registers in one shot, so the TMEM buffer is handed back
      // to the MMA warp before any arithmetic.  Softmax runs in the log2 domain: x2 = s * (log2e / div),
      // p = 2^(x2 - m2): one FFMA + one MUFU.EX2 per element; masking only on the diagonal tile.
      // On the diagonal tile the causal mask is warp-uniform per 32-column chunk: this warp's rows are
      // quarter*32 .. +31, so chunks below `quarter` are fully visible, chunk `quarter` is the triangle
      // (column t visible to lane >= t) and the chunks above it are fully masked — those are neither
      // loaded nor exponentiated, only zero-filled in P.  (At S = 256 two of every three key tiles are
      // diagonal ones; masking every element cost 8.5 instructions per score against 3.5.)
      const int n_live = diag ? quarter + 1 : 4;  // chunks with any visible column (warp-uniform)
      float m_new, alpha, sum = 0.f;
      if constexpr (DH <= 64) {
        uint32_t r[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          if (cc < n_live) tmem_ld_32x32(tS + cc * 32, r[cc]);
        tmem_ld_wait();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + st);
        if (diag) {  // the triangle: masked scores become -inf once (2^-inf = 0), the math below is mask-free
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            if (cc == quarter) {
#pragma unroll
              for (int t = 0; t < 32; ++t) r[cc][t] = (t <= lane) ? r[cc][t] : 0xff800000u;
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          if (cc < n_live) {
#pragma unroll
            for (int t = 0; t < 32; ++t) mx = fmaxf(mx, __uint_as_float(r[cc][t]));
          }
        m_new = fmaxf(m_run, mx * k2);  // k2 > 0, so max commutes with the scaling
        alpha = c.first_step() ? 0.f : ex2f(m_run - m_new);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float p[32];
          if (cc < n_live) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              p[t] = ex2f(fmaf(__uint_as_float(r[cc][t]), k2, -m_new));
              sum += p[t];
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) p[t] = 0.f;
          }
          store_row_chunk(sP, row, cc, p);
        }
      } else {
        // Dh = 96: the 96 output accumulators leave no room for the whole score row in registers; two passes
        // over the TMEM buffer (row max, then the exponentials), 32 columns at a time; the buffer goes back
        // to the MMA warp after the second pass.
        float mx = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < n_live; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(tS + cc * 32, r);
          tmem_ld_wait();
          const bool tri = diag && cc == quarter;
#pragma unroll
          for (int t = 0; t < 32; ++t) mx = fmaxf(mx, (tri && t > lane) ? -INFINITY : __uint_as_float(r[t]));
        }
        m_new = fmaxf(m_run, mx * k2);
        alpha = c.first_step() ? 0.f : ex2f(m_run - m_new);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          float p[32];
          if (cc < n_live) {
            uint32_t r[32];
            tmem_ld_32x32(tS + cc * 32, r);
            tmem_ld_wait();
            const bool tri = diag && cc == quarter;
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              p[t] = (tri && t > lane) ? 0.f : ex2f(fmaf(__uint_as_float(r[t]), k2, -m_new));
              sum += p[t];
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) p[t] = 0.f;
          }
          store_row_chunk(sP, row, cc, p);
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty + st);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      fence_before();          // TMEM reads of O two steps ago are complete
      fence_proxy_async();     // P writes visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + st);
      mbar_wait(pv_done + st, ph);
      fence_after();
#pragma unroll
      for (int cc = 0; cc < DH / 32; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(tO + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o[cc * 32 + t] = o[cc * 32 + t] * alpha + __uint_as_float(r[t]);
      }
      if (c.last_step()) {
        const int b = c.bh / h, hh = c.bh - b * h;
        const float inv_l = 1.0f / l_run;
        const int64_t q_glob = (int64_t)b * S + c.tile * 128 + row;
        __nv_bfloat16* orow = out + (q_glob * h + hh) * DH;
#pragma unroll
        for (int q = 0; q < DH / 8; ++q) {
          uint4 u;
          u.x = pack2(o[8 * q] * inv_l, o[8 * q + 1] * inv_l);
          u.y = pack2(o[8 * q + 2] * inv_l, o[8 * q + 3] * inv_l);
          u.z = pack2(o[8 * q + 4] * inv_l, o[8 * q + 5] * inv_l);
          u.w = pack2(o[8 * q + 6] * inv_l, o[8 * q + 7] * inv_l);
          *reinterpret_cast<uint4*>(orow + 8 * q) = u;
        }
        lse[((int64_t)b * h + hh) * S + c.tile * 128 + row] = m_run * 0.6931471805599453f + __logf(l_run);  // natural log
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: D[b,h,s] = sum_d dO*O   (8 lanes per row of Dh)
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                     float* __restrict__ D, int64_t rows /* B*S*h */, int S, int h) {
  pdl_wait();
  pdl_trigger();
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int part = threadIdx.x & 7;
  constexpr int PER = DH / 8;   // elements per lane: 8 or 12 (8-byte aligned)
  float acc = 0.f;
  if (r < rows) {
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
      const uint2 a = *reinterpret_cast<const uint2*>(o + r * DH + part * PER + i * 4);
      const uint2 g = *reinterpret_cast<const uint2*>(d_o + r * DH + part * PER + i * 4);
      const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* pg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 x = __bfloat1622float2(pa[j]), y = __bfloat1622float2(pg[j]);
        acc += x.x * y.x + x.y * y.y;
      }
    }
  }
  acc += __shfl_xor_sync(kFull, acc, 1);
  acc += __shfl_xor_sync(kFull, acc, 2);
  acc += __shfl_xor_sync(kFull, acc, 4);
  if (r < rows && part == 0) {
    // row r = (b*S + s)*h + hh  ->  D[(b*h + hh)*S + s]
    const int hh = (int)(r % h);
    const int64_t bs = r / h;
    const int64_t b = bs / S, s = bs - b * S;
    D[(b * h + hh) * S + s] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// backward, shared row arithmetic: p = exp(s/div - lse), dS = p*(dP - D)/div, both to smem tiles.
// `wait_tiles` is polled (once) right before the first shared-memory store.
// ------------------------------------------------------------------------------------------------
template <bool WRITE_P>
__device__ __forceinline__ void bwd_rows(uint32_t tS, uint32_t tdP, int row, bool diag, float lse_r, float D_r,
                                         float inv_div, uint32_t sP, uint32_t sdS, uint64_t* wait_tiles,
                                         uint32_t wait_parity) {
  const float k2 = inv_div * 1.4426950408889634f, lse2 = lse_r * 1.4426950408889634f;  // log2 domain
  // diagonal tile: per 32-column chunk the causal mask is warp-uniform (rows quarter*32 .. +31): chunks
  // below `quarter` are fully visible, chunk `quarter` is the triangle, the chunks above are all zero
  const int quarter = row >> 5, lane = row & 31;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float p[32], ds[32];
    if (diag && c > quarter) {
#pragma unroll
      for (int t = 0; t < 32; ++t) { p[t] = 0.f; ds[t] = 0.f; }
    } else {
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tS + c * 32, rs);
      tmem_ld_32x32(tdP + c * 32, rp);
      tmem_ld_wait();
      if (diag && c == quarter) {
#pragma unroll
        for (int t = 0; t < 32; ++t) rs[t] = (t <= lane) ? rs[t] : 0xff800000u;  // -inf: 2^-inf = 0
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        p[t] = ex2f(fmaf(__uint_as_float(rs[t]), k2, -lse2));
        ds[t] = p[t] * (__uint_as_float(rp[t]) - D_r) * inv_div;
      }
    }
    if (c == 0 && wait_tiles) mbar_wait(wait_tiles, wait_parity);  // previous P / dS tiles consumed by the MMAs
    if (WRITE_P) store_row_chunk(sP, row, c, p);
    store_row_chunk(sdS, row, c, ds);
  }
}

template <int DH>
__device__ __forceinline__ void store_acc_row(uint32_t taddr, __nv_bfloat16* dst) {
#pragma unroll
  for (int c = 0; c < DH / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      u.x = pack2(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1]));
      u.y = pack2(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3]));
      u.z = pack2(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5]));
      u.w = pack2(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7]));
      *reinterpret_cast<uint4*>(dst + c * 32 + 8 * q) = u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward dK / dV : item = (b, head, key tile j); steps = query tiles i = j .. nqt-1
// ------------------------------------------------------------------------------------------------
template <int DH>
struct AttnDkvSmem {
  static constexpr int T = AttnCfg<DH>::TILE;
  static constexpr int KVB = (DH <= 64) ? 2 : 1;   // K / V buffers (per item): double-buffered while they fit
  static constexpr int K = 0, V = KVB * T, Q = 2 * KVB * T, DO = Q + 2 * T, P = DO + 2 * T, DS = P + AT_PTILE,
                       BAR = DS + AT_PTILE;
  static constexpr int TOTAL = BAR + 24 * 8 + 16 + 1024;
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const float* __restrict__ lse, const float* __restrict__ D, __nv_bfloat16* __restrict__ dqkv,
                    int B, int S, int h, float inv_div) {
  using L = AttnDkvSmem<DH>;
  constexpr int AT_TILE = L::T, KVB = L::KVB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t *kv_full = bars, *kv_empty = bars + 2, *qdo_full = bars + 4, *qdo_empty = bars + 6, *s_full = bars + 8,
           *s_empty = bars + 9, *pds_full = bars + 10, *acc_done = bars + 11, *out_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = S / 128;
  const int total = B * h * nqt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1);
      mbar_init(qdo_full + s, 1); mbar_init(qdo_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(pds_full, 4);
    mbar_init(acc_done, 1);
    mbar_init(out_done, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();     // set-up above is private; global memory is touched only from here on
  pdl_trigger();  // persistent grid: the next kernel may be scheduled behind this one
  const uint32_t tmem_S = tmem, tmem_dP = tmem + 128, tmem_dK = tmem + 256, tmem_dV = tmem + 256 + DH;

  if (warp == 0) {
    if (lane == 0) {
      Cursor c;
      for (c.init(blockIdx.x, total, gridDim.x, nqt, false, true); c.valid(); c.next()) {
        const int b = c.bh / h, hh = c.bh - b * h;
        const int col_q = hh * 3 * DH, row0 = b * S;
        if (c.first_step()) {
          const int kb = c.wi % KVB;
          mbar_wait(kv_empty + kb, ((c.wi / KVB) & 1) ^ 1);
          mbar_expect_tx(kv_full + kb, 2 * AT_TILE);
          load_tile<DH>(&tmQKV, kv_full + kb, smem + L::K + kb * AT_TILE, col_q + DH, row0 + c.tile * 128);
          load_tile<DH>(&tmQKV, kv_full + kb, smem + L::V + kb * AT_TILE, col_q + 2 * DH, row0 + c.tile * 128);
        }
        const int st = c.t & 1, i = c.tile + c.j;
        mbar_wait(qdo_empty + st, ((c.t >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + st, 2 * AT_TILE);
        load_tile<DH>(&tmQKV, qdo_full + st, smem + L::Q + st * AT_TILE, col_q, row0 + i * 128);
        load_tile<DH>(&tmDO, qdo_full + st, smem + L::DO + st * AT_TILE, hh * DH, row0 + i * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, false, false);  // S = Q K^T, dP = dO V^T
      constexpr uint32_t idesc_a = make_idesc(DH, true, true);     // dV = P^T dO, dK = dS^T Q : both MN-major
      const uint32_t sp = smem_u32(smem + L::P), sds = smem_u32(smem + L::DS);
      auto issue_sdp = [&](const Cursor& c) {
        const int st = c.t & 1, kb = c.wi % KVB;
        if (c.first_step()) mbar_wait(kv_full + kb, (c.wi / KVB) & 1);
        mbar_wait(qdo_full + st, (c.t >> 1) & 1);
        mbar_wait(s_empty, (c.t & 1) ^ 1);  // threads have read S / dP of the previous step
        fence_after();
        const uint32_t tq = smem_u32(smem + L::Q + st * AT_TILE), tk = smem_u32(smem + L::K + kb * AT_TILE);
        const uint32_t tdo = smem_u32(smem + L::DO + st * AT_TILE), tv = smem_u32(smem + L::V + kb * AT_TILE);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_S, desc_k<DH>(tq, k), desc_k<DH>(tk, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_dP, desc_k<DH>(tdo, k), desc_k<DH>(tv, k), idesc_s, k > 0);
        umma_commit(s_full);
      };
      Cursor cs, ca;
      cs.init(blockIdx.x, total, gridDim.x, nqt, false, true);
      ca = cs;
      if (cs.valid()) { issue_sdp(cs); cs.next(); }
      for (; ca.valid(); ca.next()) {
        const int st = ca.t & 1, kb = ca.wi % KVB;
        mbar_wait(pds_full, ca.t & 1);               // P / dS of this step are in shared memory (S / dP drained)
        // next score tiles first: the threads are waiting for them.  With a single K / V buffer (Dh = 96) the first
        // step of the NEXT item needs that buffer, which is released by the MMAs below: issue it after them.
        const bool early = cs.valid() && (KVB > 1 || !cs.first_step());
        if (early) { issue_sdp(cs); cs.next(); }
        if (ca.first_step()) mbar_wait(out_done, (ca.wi & 1) ^ 1);  // previous item's dK / dV were read out
        fence_after();
        // A = P^T / dS^T : MN-major (M = keys: two 64-key halves 16 KB apart; K = queries: 16 rows = 2048 B)
        // B = dO / Q tile : MN-major (N = Dh, K = queries: 16 rows = 2048 B)
        const uint64_t dpt = make_smem_desc(sp, 16384, 1024), ddst = make_smem_desc(sds, 16384, 1024);
        const uint32_t tdo = smem_u32(smem + L::DO + st * AT_TILE), tq = smem_u32(smem + L::Q + st * AT_TILE);
        const bool acc0 = !ca.first_step();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_dV, dpt + (uint64_t)(k * 128), desc_mn<DH>(tdo, k), idesc_a, (acc0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_dK, ddst + (uint64_t)(k * 128), desc_mn<DH>(tq, k), idesc_a, (acc0 || k > 0));
        umma_commit(acc_done);
        umma_commit(qdo_empty + st);
        if (ca.last_step()) umma_commit(kv_empty + kb);
        if (!early && cs.valid()) { issue_sdp(cs); cs.next(); }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t sP = smem_u32(smem + L::P), sdS = smem_u32(smem + L::DS);
    Cursor c;
    for (c.init(blockIdx.x, total, gridDim.x, nqt, false, true); c.valid(); c.next()) {
      const int b = c.bh / h, hh = c.bh - b * h;
      const int i = c.tile + c.j;
      const int64_t stat = ((int64_t)b * h + hh) * S + i * 128 + row;
      const float lse_r = lse[stat], D_r = D[stat];
      mbar_wait(s_full, c.t & 1);
      fence_after();
      // P / dS tiles are free once the previous step's dV / dK MMAs retired
      bwd_rows<true>(tmem_S + lane_addr, tmem_dP + lane_addr, row, c.first_step(), lse_r, D_r, inv_div, sP, sdS,
                     c.t > 0 ? acc_done : nullptr, (c.t - 1) & 1);
      fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(pds_full); }
      if (c.last_step()) {
        mbar_wait(acc_done, c.t & 1);
        fence_after();
        // rows of dK / dV are KEYS of tile j
        const int64_t k_glob = (int64_t)b * S + c.tile * 128 + row;
        __nv_bfloat16* base = dqkv + (k_glob * h + hh) * 3 * DH;
        store_acc_row<DH>(tmem_dK + lane_addr, base + DH);
        store_acc_row<DH>(tmem_dV + lane_addr, base + 2 * DH);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_done);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward dQ : item = (b, head, query tile i); steps = key tiles j = 0 .. i
// ------------------------------------------------------------------------------------------------
template <int DH>
struct AttnDqSmem {
  static constexpr int T = AttnCfg<DH>::TILE;
  static constexpr int Q = 0, DO = 2 * T, K = 4 * T, V = 6 * T, DS = 8 * T, BAR = 8 * T + AT_PTILE;
  static constexpr int TOTAL = BAR + 24 * 8 + 16 + 1024;
};

template <int DH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse, const float* __restrict__ D, __nv_bfloat16* __restrict__ dqkv,
                   int B, int S, int h, float inv_div) {
  using L = AttnDqSmem<DH>;
  constexpr int AT_TILE = L::T;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t *qdo_full = bars, *qdo_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 6, *s_full = bars + 8,
           *s_empty = bars + 9, *ds_full = bars + 10, *acc_done = bars + 11, *out_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = S / 128;
  const int total = B * h * nqt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qdo_full + s, 1); mbar_init(qdo_empty + s, 1);
      mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(ds_full, 4);
    mbar_init(acc_done, 1);
    mbar_init(out_done, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();     // set-up above is private; global memory is touched only from here on
  pdl_trigger();  // persistent grid: the next kernel may be scheduled behind this one
  const uint32_t tmem_S = tmem, tmem_dP = tmem + 128, tmem_dQ = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      Cursor c;
      for (c.init(blockIdx.x, total, gridDim.x, nqt, true, false); c.valid(); c.next()) {
        const int b = c.bh / h, hh = c.bh - b * h;
        const int col_q = hh * 3 * DH, row0 = b * S;
        if (c.first_step()) {
          const int qb = c.wi & 1;
          mbar_wait(qdo_empty + qb, ((c.wi >> 1) & 1) ^ 1);
          mbar_expect_tx(qdo_full + qb, 2 * AT_TILE);
          load_tile<DH>(&tmQKV, qdo_full + qb, smem + L::Q + qb * AT_TILE, col_q, row0 + c.tile * 128);
          load_tile<DH>(&tmDO, qdo_full + qb, smem + L::DO + qb * AT_TILE, hh * DH, row0 + c.tile * 128);
        }
        const int st = c.t & 1;
        mbar_wait(kv_empty + st, ((c.t >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full + st, 2 * AT_TILE);
        load_tile<DH>(&tmQKV, kv_full + st, smem + L::K + st * AT_TILE, col_q + DH, row0 + c.j * 128);
        load_tile<DH>(&tmQKV, kv_full + st, smem + L::V + st * AT_TILE, col_q + 2 * DH, row0 + c.j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, false, false);
      constexpr uint32_t idesc_q = make_idesc(DH, false, true);  // dQ = dS K : dS K-major, K tile MN-major
      const uint32_t sds = smem_u32(smem + L::DS);
      auto issue_sdp = [&](const Cursor& c) {
        const int st = c.t & 1, qb = c.wi & 1;
        if (c.first_step()) mbar_wait(qdo_full + qb, (c.wi >> 1) & 1);
        mbar_wait(kv_full + st, (c.t >> 1) & 1);
        mbar_wait(s_empty, (c.t & 1) ^ 1);
        fence_after();
        const uint32_t tq = smem_u32(smem + L::Q + qb * AT_TILE), tk = smem_u32(smem + L::K + st * AT_TILE);
        const uint32_t tdo = smem_u32(smem + L::DO + qb * AT_TILE), tv = smem_u32(smem + L::V + st * AT_TILE);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_S, desc_k<DH>(tq, k), desc_k<DH>(tk, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_dP, desc_k<DH>(tdo, k), desc_k<DH>(tv, k), idesc_s, k > 0);
        umma_commit(s_full);
      };
      Cursor cs, ca;
      cs.init(blockIdx.x, total, gridDim.x, nqt, true, false);
      ca = cs;
      if (cs.valid()) { issue_sdp(cs); cs.next(); }
      for (; ca.valid(); ca.next()) {
        const int st = ca.t & 1, qb = ca.wi & 1;
        mbar_wait(ds_full, ca.t & 1);
        if (cs.valid()) { issue_sdp(cs); cs.next(); }
        if (ca.first_step()) mbar_wait(out_done, (ca.wi & 1) ^ 1);
        fence_after();
        const uint32_t tk = smem_u32(smem + L::K + st * AT_TILE);
        const bool acc0 = !ca.first_step();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dds = make_smem_desc(sds + (k >> 2) * 16384, 16, 1024) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem_dQ, dds, desc_mn<DH>(tk, k), idesc_q, (acc0 || k > 0));
        }
        umma_commit(acc_done);
        umma_commit(kv_empty + st);
        if (ca.last_step()) umma_commit(qdo_empty + qb);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t sdS = smem_u32(smem + L::DS);
    float lse_r = 0.f, D_r = 0.f;
    Cursor c;
    for (c.init(blockIdx.x, total, gridDim.x, nqt, true, false); c.valid(); c.next()) {
      const int b = c.bh / h, hh = c.bh - b * h;
      if (c.first_step()) {
        const int64_t stat = ((int64_t)b * h + hh) * S + c.tile * 128 + row;
        lse_r = lse[stat];
        D_r = D[stat];
      }
      mbar_wait(s_full, c.t & 1);
      fence_after();
      bwd_rows<false>(tmem_S + lane_addr, tmem_dP + lane_addr, row, c.last_step(), lse_r, D_r, inv_div, 0u, sdS,
                      c.t > 0 ? acc_done : nullptr, (c.t - 1) & 1);
      fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(ds_full); }
      if (c.last_step()) {
        mbar_wait(acc_done, c.t & 1);
        fence_after();
        const int64_t q_glob = (int64_t)b * S + c.tile * 128 + row;
        store_acc_row<DH>(tmem_dQ + lane_addr, dqkv + (q_glob * h + hh) * 3 * DH);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_done);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, S = 256 (two 128-row tiles), Dh = 64: ONE kernel for dQ, dK and dV.
// The split dkv / dq kernels above recompute S and dP (and the exponentials, the expensive part: MUFU-bound) for
// every (query tile, key tile) pair twice, once per kernel.  With two tiles per sequence the whole causal
// problem of a (batch, head) fits one CTA: item = (b, head); steps = the three live pairs (i, j) =
// (0,0), (1,0), (1,1); per pair S, dP -> P, dS (once) -> dV_j += P^T dO_i, dK_j += dS^T Q_i, dQ_i += dS K_j.
// TMEM: S | dP | dK | dV | dQ_0 | dQ_1 = 128 + 128 + 4 x 64 = 512 columns.  Every operand tile is loaded once per
// item (8 tiles, 128 KB) and released as soon as its last pair is done, so the next item's loads run under
// this item's later pairs.  15 tile GEMMs and 3 softmax-backward tiles per item instead of 21 and 6.
// ------------------------------------------------------------------------------------------------
struct AttnBwd2Smem {
  static constexpr int T = AttnCfg<64>::TILE;   // 16 KB
  // tile ids: 0 Q0, 1 Q1, 2 dO0, 3 dO1, 4 K0, 5 K1, 6 V0, 7 V1
  static constexpr int TILES = 0, P = 8 * T, DS = P + AT_PTILE, BAR = DS + AT_PTILE;
  static constexpr int TOTAL = BAR + 32 * 8 + 16 + 1024;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_s256_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const float* __restrict__ lse, const __nv_bfloat16* __restrict__ o_fwd, const __nv_bfloat16* __restrict__ d_o,
                     __nv_bfloat16* __restrict__ dqkv, int B, int h, float inv_div) {
  constexpr int DH = 64, S = 256;
  using L = AttnBwd2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
  uint64_t *full = bars, *empty = bars + 8, *s_full = bars + 16, *s_empty = bars + 17, *pds_full = bars + 18,
           *acc_done = bars + 19, *kv_out_done = bars + 20, *out_done = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = B * h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 8; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(pds_full, 4);
    mbar_init(acc_done, 1);
    mbar_init(kv_out_done, 4);
    mbar_init(out_done, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_S = tmem, tmem_dP = tmem + 128, tmem_dK = tmem + 256, tmem_dV = tmem + 320, tmem_dQ0 = tmem + 384,
                 tmem_dQ1 = tmem + 448;
  auto tile = [&](int id) { return smem + L::TILES + id * L::T; };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t wi = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++wi) {
        const int b = w / h, hh = w - b * h;
        const int col_q = hh * 3 * DH, row0 = b * S;
        const uint32_t ph = (wi & 1) ^ 1;
        // in the order the pairs need them: (0,0): Q0 K0 dO0 V0; (1,0): Q1 dO1; (1,1): K1 V1
        auto load = [&](int id, const CUtensorMap* map, int col, int row) {
          mbar_wait(empty + id, ph);
          mbar_expect_tx(full + id, L::T);
          load_tile<DH>(map, full + id, tile(id), col, row);
        };
        load(0, &tmQKV, col_q, row0);
        load(4, &tmQKV, col_q + DH, row0);
        load(2, &tmDO, hh * DH, row0);
        load(6, &tmQKV, col_q + 2 * DH, row0);
        load(1, &tmQKV, col_q, row0 + 128);
        load(3, &tmDO, hh * DH, row0 + 128);
        load(5, &tmQKV, col_q + DH, row0 + 128);
        load(7, &tmQKV, col_q + 2 * DH, row0 + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, false, false);  // S = Q K^T, dP = dO V^T
      constexpr uint32_t idesc_a = make_idesc(DH, true, true);     // dV = P^T dO, dK = dS^T Q
      constexpr uint32_t idesc_q = make_idesc(DH, false, true);    // dQ = dS K
      const uint32_t sp = smem_u32(smem + L::P), sds = smem_u32(smem + L::DS);
      const int n_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t n_steps = (uint32_t)n_items * 3;
      // pair p of an item: query tile i, key tile j
      auto issue_sdp = [&](uint32_t t) {
        const uint32_t wi = t / 3, p = t - wi * 3, ph = wi & 1;
        const int i = p == 0 ? 0 : 1, j = p == 2 ? 1 : 0;
        if (p == 0) { mbar_wait(full + 0, ph); mbar_wait(full + 4, ph); mbar_wait(full + 2, ph); mbar_wait(full + 6, ph); }
        if (p == 1) { mbar_wait(full + 1, ph); mbar_wait(full + 3, ph); }
        if (p == 2) { mbar_wait(full + 5, ph); mbar_wait(full + 7, ph); }
        mbar_wait(s_empty, (t & 1) ^ 1);   // the threads have read S / dP of the previous step
        fence_after();
        const uint32_t tq = smem_u32(tile(i)), tdo = smem_u32(tile(2 + i)), tk = smem_u32(tile(4 + j)), tv = smem_u32(tile(6 + j));
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_S, desc_k<DH>(tq, k), desc_k<DH>(tk, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_dP, desc_k<DH>(tdo, k), desc_k<DH>(tv, k), idesc_s, k > 0);
        umma_commit(s_full);
      };
      if (n_steps > 0) issue_sdp(0);
      for (uint32_t t = 0; t < n_steps; ++t) {
        const uint32_t wi = t / 3, p = t - wi * 3;
        const int i = p == 0 ? 0 : 1, j = p == 2 ? 1 : 0;
        mbar_wait(pds_full, t & 1);                 // P / dS of this step are in shared memory (S / dP drained)
        if (t + 1 < n_steps) issue_sdp(t + 1);      // next score tiles first: the threads are waiting for them
        if (p == 0) mbar_wait(out_done, (wi & 1) ^ 1);   // the previous item's accumulators were read out
        if (p == 2) mbar_wait(kv_out_done, wi & 1);      // dK_0 / dV_0 were read out: the columns take dK_1 / dV_1
        fence_after();
        const uint64_t dpt = make_smem_desc(sp, 16384, 1024), ddst = make_smem_desc(sds, 16384, 1024);
        const uint32_t tq = smem_u32(tile(i)), tdo = smem_u32(tile(2 + i)), tk = smem_u32(tile(4 + j));
        const uint32_t tdq = i == 0 ? tmem_dQ0 : tmem_dQ1;
        const bool acc_kv = p == 1, acc_q = p == 2;     // (1,0) adds to dK_0 / dV_0; (1,1) adds to dQ_1
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_dV, dpt + (uint64_t)(k * 128), desc_mn<DH>(tdo, k), idesc_a, (acc_kv || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_dK, ddst + (uint64_t)(k * 128), desc_mn<DH>(tq, k), idesc_a, (acc_kv || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dds = make_smem_desc(sds + (k >> 2) * 16384, 16, 1024) + (uint64_t)((k & 3) * 2);
          umma_bf16(tdq, dds, desc_mn<DH>(tk, k), idesc_q, (acc_q || k > 0));
        }
        umma_commit(acc_done);
        // operand tiles whose last pair this was go back to the producer
        if (p == 0) { umma_commit(empty + 0); umma_commit(empty + 2); }
        if (p == 1) { umma_commit(empty + 4); umma_commit(empty + 6); }
        if (p == 2) { umma_commit(empty + 1); umma_commit(empty + 3); umma_commit(empty + 5); umma_commit(empty + 7); }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t sP = smem_u32(smem + L::P), sdS = smem_u32(smem + L::DS);
    uint32_t t = 0;
    // The accumulators a pair completes are read out at the START of the next step, behind the acc_done wait that
    // step needs anyway before it overwrites the P / dS tiles — never as a wait of their own.
    int pend_p = -1, pend_b = 0, pend_hh = 0;
    auto read_out = [&]() {   // after acc_done of the pending step was observed
      const int i = pend_p == 0 ? 0 : 1, j = pend_p == 2 ? 1 : 0;
      const int64_t q_glob = (int64_t)pend_b * S + i * 128 + row, k_glob = (int64_t)pend_b * S + j * 128 + row;
      if (pend_p == 0) {
        store_acc_row<DH>(tmem_dQ0 + lane_addr, dqkv + (q_glob * h + pend_hh) * 3 * DH);
      } else {
        __nv_bfloat16* kbase = dqkv + (k_glob * h + pend_hh) * 3 * DH;
        store_acc_row<DH>(tmem_dK + lane_addr, kbase + DH);
        store_acc_row<DH>(tmem_dV + lane_addr, kbase + 2 * DH);
        if (pend_p == 2) store_acc_row<DH>(tmem_dQ1 + lane_addr, dqkv + (q_glob * h + pend_hh) * 3 * DH);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pend_p == 1 ? kv_out_done : out_done);
      }
    };
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int b = w / h, hh = w - b * h;
      float lse_r = 0.f, D_r = 0.f;
#pragma unroll 1
      for (int p = 0; p < 3; ++p, ++t) {
        const int i = p == 0 ? 0 : 1, j = p == 2 ? 1 : 0;
        uint4 ov[8], gv[8];
        if (p < 2) {   // first pair of query tile i: its row statistics; D = sum_d dO * O of the row is computed here
          // (the split kernels take it from attn_bwd_prep_kernel: one launch per layer less).  The 2 x 128 bytes are
          // requested before the waits below and used after them.
          const int64_t stat = ((int64_t)b * h + hh) * S + i * 128 + row;
          lse_r = lse[stat];
          const int64_t q_glob = (int64_t)b * S + i * 128 + row;
          const uint4* po = reinterpret_cast<const uint4*>(o_fwd + (q_glob * h + hh) * DH);
          const uint4* pg = reinterpret_cast<const uint4*>(d_o + (q_glob * h + hh) * DH);
#pragma unroll
          for (int k = 0; k < 8; ++k) { ov[k] = __ldg(po + k); gv[k] = __ldg(pg + k); }
        }
        if (pend_p >= 0) {   // previous step: its accumulate MMAs retired -> P / dS tiles are free, its outputs complete
          mbar_wait(acc_done, (t - 1) & 1);
          fence_after();
          read_out();
        }
        if (p < 2) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ov[k]);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&gv[k]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 x = __bfloat1622float2(pa[u]), y = __bfloat1622float2(pb[u]);
              acc += x.x * y.x + x.y * y.y;
            }
          }
          D_r = acc;
        }
        mbar_wait(s_full, t & 1);
        fence_after();
        bwd_rows<true>(tmem_S + lane_addr, tmem_dP + lane_addr, row, i == j, lse_r, D_r, inv_div, sP, sdS, nullptr, 0);
        fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(pds_full); }
        pend_p = p; pend_b = b; pend_hh = hh;
      }
    }
    if (pend_p >= 0) {
      mbar_wait(acc_done, (t - 1) & 1);
      fence_after();
      read_out();
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<512>(tmem);
  }
}

static int check_shape(const char* who, int B, int S, int h, int Dh) {
  NPT_REQUIRE(B > 0 && h > 0, "%s: bad shape", who);
  NPT_REQUIRE(Dh == 64 || Dh == 96, "%s: fused attention supports head_dim 64 and 96 (got %d); use the GEMM + softmax path", who, Dh);
  NPT_REQUIRE(S > 0 && S % 128 == 0, "%s: fused attention needs seq_len %% 128 == 0 (got %d)", who, S);
  return 0;
}

static bool attn_bwd_force_split() {   // NUMPITRON_ATTN_BWD=split: the dkv + dq kernels at every shape (A/B measurements)
  static int v = -1;
  if (v < 0) { const char* e = getenv("NUMPITRON_ATTN_BWD"); v = (e && e[0] == 's') ? 1 : 0; }
  return v == 1;
}

static int persistent_grid(int items) {
  const int sms = sm_count();
  return items < sms ? items : sms;
}

template <int DH>
static int attention_fwd_t(const void* qkv_bf16, void* out_bf16, float* lse, int B, int S, int h, float divisor, cudaStream_t st) {
  using C = AttnCfg<DH>;
  CUtensorMap tm;
  if (int rc = make_map_bf16_2d(&tm, qkv_bf16, (int64_t)h * 3 * DH, (int64_t)B * S, (int64_t)h * 3 * DH, C::CH, 128, C::CH == 64 ? 128 : 64)) return rc;
  auto kern = attn_fwd_kernel<DH>;
  static bool attr = false;
  if (!attr) {
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnFwdSmem<DH>::TOTAL));
    attr = true;
  }
  const int grid = persistent_grid(B * h * (S / 128));
  NPT_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(AT_THREADS), AttnFwdSmem<DH>::TOTAL, st, tm,
                            (__nv_bfloat16*)out_bf16, lse, B, S, h, 1.0f / divisor));
  NPT_LAUNCH_CHECK();
  return 0;
}

template <int DH>
static int attention_bwd_t(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse, void* dqkv_bf16,
                           int B, int S, int h, float divisor, float* D, cudaStream_t st) {
  using C = AttnCfg<DH>;
  const int64_t rows = (int64_t)B * S * h;
  CUtensorMap tmQKV, tmDO;
  const int sw = C::CH == 64 ? 128 : 64;
  if (int rc = make_map_bf16_2d(&tmQKV, qkv_bf16, (int64_t)h * 3 * DH, (int64_t)B * S, (int64_t)h * 3 * DH, C::CH, 128, sw)) return rc;
  if (int rc = make_map_bf16_2d(&tmDO, d_out_bf16, (int64_t)h * DH, (int64_t)B * S, (int64_t)h * DH, C::CH, 128, sw)) return rc;
  const float inv_div = 1.0f / divisor;
  if (DH == 64 && S == 256 && !attn_bwd_force_split()) {   // both tiles of a sequence in one CTA: dQ, dK, dV in one pass
    static bool attr2 = false;
    if (!attr2) {
      NPT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_s256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwd2Smem::TOTAL));
      attr2 = true;
    }
    NPT_CHECK_CUDA(launch_pdl(attn_bwd_s256_kernel, dim3(persistent_grid(B * h)), dim3(AT_THREADS), AttnBwd2Smem::TOTAL, st, tmQKV,
                              tmDO, lse, (const __nv_bfloat16*)out_bf16, (const __nv_bfloat16*)d_out_bf16, (__nv_bfloat16*)dqkv_bf16,
                              B, h, inv_div));
    NPT_LAUNCH_CHECK();
    return 0;
  }
  NPT_CHECK_CUDA(launch_pdl(attn_bwd_prep_kernel<DH>, dim3((unsigned)cdiv(rows * 8, 256)), dim3(256), 0, st,
                            (const __nv_bfloat16*)out_bf16, (const __nv_bfloat16*)d_out_bf16, D, rows, S, h));
  NPT_LAUNCH_CHECK();
  auto kdkv = attn_bwd_dkv_kernel<DH>;
  auto kdq = attn_bwd_dq_kernel<DH>;
  static bool attr = false;
  if (!attr) {
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kdkv, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDkvSmem<DH>::TOTAL));
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kdq, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDqSmem<DH>::TOTAL));
    attr = true;
  }
  const int grid = persistent_grid(B * h * (S / 128));
  NPT_CHECK_CUDA(launch_pdl(kdkv, dim3(grid), dim3(AT_THREADS), AttnDkvSmem<DH>::TOTAL, st, tmQKV, tmDO, lse, (const float*)D,
                            (__nv_bfloat16*)dqkv_bf16, B, S, h, inv_div));
  NPT_LAUNCH_CHECK();
  NPT_CHECK_CUDA(launch_pdl(kdq, dim3(grid), dim3(AT_THREADS), AttnDqSmem<DH>::TOTAL, st, tmQKV, tmDO, lse, (const float*)D,
                            (__nv_bfloat16*)dqkv_bf16, B, S, h, inv_div));
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace npt

using namespace npt;

extern "C" {

int npt_attention_supported(int32_t S, int32_t Dh) { return ((Dh == 64 || Dh == 96) && S > 0 && S % 128 == 0) ? 1 : 0; }

int npt_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int32_t B, int32_t S, int32_t h, int32_t Dh,
                      float divisor, void* stream) {
  if (int rc = check_shape("attention_fwd", B, S, h, Dh)) return rc;
  NPT_REQUIRE(aligned16(qkv_bf16) && aligned16(out_bf16), "attention_fwd: buffers must be 16-byte aligned");
  if (Dh == 96) return attention_fwd_t<96>(qkv_bf16, out_bf16, lse, B, S, h, divisor, (cudaStream_t)stream);
  return attention_fwd_t<64>(qkv_bf16, out_bf16, lse, B, S, h, divisor, (cudaStream_t)stream);
}

size_t npt_attention_bwd_workspace_bytes(int32_t B, int32_t S, int32_t h) { return (size_t)B * S * h * sizeof(float); }

int npt_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                      void* dqkv_bf16, int32_t B, int32_t S, int32_t h, int32_t Dh, float divisor, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (int rc = check_shape("attention_bwd", B, S, h, Dh)) return rc;
  NPT_REQUIRE(workspace && workspace_bytes >= npt_attention_bwd_workspace_bytes(B, S, h), "attention_bwd: workspace too small");
  NPT_REQUIRE(aligned16(qkv_bf16) && aligned16(out_bf16) && aligned16(d_out_bf16) && aligned16(dqkv_bf16),
              "attention_bwd: buffers must be 16-byte aligned");
  if (Dh == 96)
    return attention_bwd_t<96>(qkv_bf16, out_bf16, d_out_bf16, lse, dqkv_bf16, B, S, h, divisor, (float*)workspace, (cudaStream_t)stream);
  return attention_bwd_t<64>(qkv_bf16, out_bf16, d_out_bf16, lse, dqkv_bf16, B, S, h, divisor, (float*)workspace, (cudaStream_t)stream);
}

}  // extern "C"
