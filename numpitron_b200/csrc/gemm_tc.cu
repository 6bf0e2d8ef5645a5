// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma with the
// fp32 accumulator in TMEM -> tcgen05.ld epilogue (alpha / bias / ReLU / ReLU-mask) -> swizzled
// shared-memory staging -> TMA store (fp32 and/or bf16).  Serves every dense contraction of the
// training step in the bf16 throughput mode (nn/linear.py:44,55,60; nn/attention.py:48,52,81-92;
// nn/embedding.py:92,99,102).
//
// Layout support (no operand is ever transposed in memory):
//   A K-major  (trans_a=0, stored M x K)   one TMA box {64 k, 128 m}
//   A MN-major (trans_a=1, stored K x M)   two TMA boxes {64 m, 64 k}
//   B K-major  (trans_b=1, stored N x K)   one TMA box {64 k, BN n}
//   B MN-major (trans_b=0, stored K x N)   BN/64 TMA boxes {64 n, 64 k}
// Batches ride in tensor-map dimensions 2 and 3 (inner / outer batch index), which is how the
// per-head [q|k|v] column layout of the QKV projection is addressed in place.
//
// Persistent: one CTA per SM loops over 128 x BN output tiles (n fastest, so the A panel of a
// row of tiles stays in L2).  6 warps: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2-5 = epilogue (one TMEM lane quarter each).  The accumulator is
// double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of
// tile i+1.  Split-K partials go to a workspace (also by TMA store) and are reduced in fixed
// order by gemm.cu (deterministic).
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"

namespace npt {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Multicast form: the box lands at the same shared-memory offset in every CTA of `mask`, and the
// bytes are credited to the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6, %7}], [%2], %3;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Explicit shared-space accesses (the staging pointers come from an integer round-up, so the
// compiler would otherwise emit slower generic LD/ST for them).
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64u(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): SWIZZLE_128B, version 1.
//   K-major : LBO unused (1), SBO = 1024 B between 8-row groups.
//   MN-major: LBO = bytes between 64-element MN chunks, SBO = 1024 B between 8-k groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7) | (1u << 10)        // a_format = b_format = BF16
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

template <int BN>
struct TcSmem {
  // smem budget: ~178 KB of pipeline stages next to 48 KB of epilogue staging
  static constexpr int STAGES = BN >= 256 ? 3 : BN >= 192 ? 4 : BN >= 128 ? 5 : 6;
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;  // power of two >= 2*BN
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  // per epilogue warp: 2 x (32 rows x 128 B) fp32 staging + 2 x (32 rows x 64 B) bf16 staging
  static constexpr int EPI_F32 = 32 * 128, EPI_BF16 = 32 * 64;
  static constexpr int EPI_WARP_BYTES = 2 * EPI_F32 + 2 * EPI_BF16;
  static constexpr int EPI_OFFSET = PIPE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + 4 * EPI_WARP_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
};

struct TcParams {
  int M, N, K, batch, splits, kb_per_split;
  int m_tiles, n_tiles, total_tiles;
  int tma_store;  // 1: staged TMA-store epilogue; 0: direct register stores (unaligned outputs)
  int raw;        // 1: split-K partial (no epilogue math), stored through map C as {N, M, batch, split}
};

// CL = 2: CTA pairs (thread-block clusters) work on two vertically adjacent tiles (m, n) / (m+1, n);
// each CTA fetches half of the shared B tile and TMA-multicasts it into both CTAs' shared memory,
// which halves the L2 -> SM traffic of the weight operand (the dominant term when M >> N).
template <int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCb,
               const TcParams p, const GemmEpilogue epi, float* __restrict__ partial) {
  using L = TcSmem<BN>;
  constexpr int TC_STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (p.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) {
      if (epi.C || p.raw) tma_prefetch_desc(&tmC);
      if (epi.Cb && !p.raw) tma_prefetch_desc(&tmCb);
    }
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, CL);  // every CTA of the cluster must have consumed the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar + a, 1);
      mbar_init(tmem_empty_bar + a, 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<L::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();  // peers' barriers are initialised before any multicast reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  // work units: one tile (CL = 1) or one vertical pair of tiles (CL = 2) per step of the persistent loop
  const int unit_first = blockIdx.x / CL, unit_stride = gridDim.x / CL, units = p.total_tiles / CL;

  // tile -> (n_tile fastest, m_tile, z = batch*splits)
  auto decode = [&](int t, int& m0, int& n0, int& b, int& split, int& bi, int& bo, int& kb_begin, int& nkb) {
    const int nt = t % p.n_tiles;
    const int r = t / p.n_tiles;
    const int mrows = p.m_tiles / CL;          // rows of units
    const int mt = (r % mrows) * CL + cta_rank;
    const int z = r / mrows;
    b = z / p.splits;
    split = z - b * p.splits;
    bo = b / epi.batch_inner;
    bi = b - bo * epi.batch_inner;
    m0 = mt * TC_BM;
    n0 = nt * BN;
    kb_begin = split * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > kblocks) kb_end = kblocks;
    nkb = kb_end - kb_begin;  // host guarantees >= 1
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles
      for (int t = unit_first; t < units; t += unit_stride) {
        int m0, n0, b, split, bi, bo, kb_begin, nkb;
        decode(t, m0, n0, b, split, bi, bo, kb_begin, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(empty_bar + s, ((it / TC_STAGES) & 1) ^ 1);
          mbar_expect_tx(full_bar + s, L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const int k0 = (kb_begin + i) * TC_BK;
          if (A_MN) {
            tma_load_4d(&tmA, full_bar + s, sa, m0, k0, bi, bo);
            tma_load_4d(&tmA, full_bar + s, sa + 64 * TC_BK * 2, m0 + 64, k0, bi, bo);
          } else {
            tma_load_4d(&tmA, full_bar + s, sa, k0, m0, bi, bo);
          }
          if (CL > 1) {
            // this CTA fetches its half of the B tile and multicasts it to both CTAs of the pair
            if (B_MN) {  // boxes {64 n, 32 k}: k rows [rank*32, rank*32+32) of every 64-wide n chunk
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_4d_mc(&tmB, full_bar + s, sb + c * 64 * TC_BK * 2 + cta_rank * 32 * 128, n0 + c * 64,
                               k0 + cta_rank * 32, bi, bo, (uint16_t)3);
            } else {     // box {64 k, BN/2 n}: n rows [rank*BN/2, ...)
              tma_load_4d_mc(&tmB, full_bar + s, sb + cta_rank * (BN / 2) * 128, k0, n0 + cta_rank * (BN / 2), bi, bo,
                             (uint16_t)3);
            }
          } else if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_4d(&tmB, full_bar + s, sb + c * 64 * TC_BK * 2, n0 + c * 64, k0, bi, bo);
          } else {
            tma_load_4d(&tmB, full_bar + s, sb, k0, n0, bi, bo);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      uint32_t it = 0, tile_i = 0;
      for (int t = unit_first; t < units; t += unit_stride, ++tile_i) {
        int m0, n0, b, split, bi, bo, kb_begin, nkb;
        decode(t, m0, n0, b, split, bi, bo, kb_begin, nkb);
        const uint32_t acc = tile_i & 1;
        mbar_wait(tmem_empty_bar + acc, ((tile_i >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(full_bar + s, (it / TC_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          // K-major: 16 k = 32 B inside the 128 B swizzle row.  MN-major: 16 k = two 8-k groups = 2048 B.
          const uint64_t da = A_MN ? make_smem_desc(sa, 64 * TC_BK * 2, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb, 64 * TC_BK * 2, 1024) : make_smem_desc(sb, 16, 1024);
          constexpr uint32_t adv_a = (A_MN ? 2048 : 32) >> 4, adv_b = (B_MN ? 2048 : 32) >> 4;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(d_tmem, da + (uint64_t)(k * adv_a), db + (uint64_t)(k * adv_b), idesc, (i > 0 || k > 0) ? 1u : 0u);
          // frees this smem stage (in both CTAs of a pair) once the MMAs above retire
          if (CL > 1) umma_commit_mc(empty_bar + s, (uint16_t)3); else umma_commit(empty_bar + s);
        }
        umma_commit(tmem_full_bar + acc);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int quarter = warp & 3;
    uint8_t* stage_base = smem + L::EPI_OFFSET + (warp - 2) * L::EPI_WARP_BYTES;
    const bool want_f32 = p.raw || epi.C != nullptr;
    const bool want_bf16 = !p.raw && epi.Cb != nullptr;
    const bool masked = !p.raw && epi.mask != nullptr;
    uint32_t tile_i = 0, store_i = 0;  // store_i counts committed TMA-store groups (staging buffer parity)
    for (int t = unit_first; t < units; t += unit_stride, ++tile_i) {
      int m0, n0, b, split, bi, bo, kb_begin, nkb;
      decode(t, m0, n0, b, split, bi, bo, kb_begin, nkb);
      const uint32_t acc = tile_i & 1;
      const int mrow0 = m0 + quarter * 32;  // first row of this warp's 32-row band
      const int m = mrow0 + lane;
      // ReLU-mask prefetch (bf16 mask, aligned, block fully inside the matrix): eight 8-byte loads per
      // lane, issued back to back one chunk ahead of their use.  lane -> (row i*4 + lane/8, cols (lane%8)*4..+3).
      uint2 mnext[8];
      bool next_fast = false;
      auto prefetch_mask = [&](int cc) {
        const int nbm = n0 + cc * 32;
        next_fast = epi.mask_bf16 && epi.mask_vec4 && (mrow0 + 32 <= p.M) && (nbm + 32 <= p.N);
        if (next_fast) {
          const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(epi.mask) + bo * epi.smo + bi * epi.smi +
                                    (int64_t)(mrow0 + (lane >> 3)) * epi.ldm + nbm + (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) mnext[i] = __ldg(reinterpret_cast<const uint2*>(mp + (int64_t)(i * 4) * epi.ldm));
        }
      };
      if (masked && p.tma_store) prefetch_mask(0);
      // packed ReLU mask: one 32-bit word per (row, 32-column chunk); the whole tile's words for this
      // lane's row are fetched here, before the wait on the accumulator.
      uint32_t mwords[BN / 32];
      if (epi.bits_in) {
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc)
          mwords[cc] = (m < p.M && n0 + cc * 32 < p.N) ? __ldg(epi.bits_in + (int64_t)m * epi.ld_bits + (n0 >> 5) + cc) : 0u;
      }
      mbar_wait(tmem_full_bar + acc, (tile_i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        // ReLU-mask bits for this warp's 32x32 block, fetched first so the global-load latency hides
        // behind the TMEM load and the epilogue math: bit (4*i + e) <-> row i*4 + lane/8, col (lane%8)*4 + e.
        uint32_t mbits = 0;
        uint2 mraw[8];         // bf16 fast path: 4 mask values per 8-byte load
#pragma unroll
        for (int i = 0; i < 8; ++i) mraw[i] = mnext[i];
        const bool mask_fast = next_fast;
        if (masked && p.tma_store) {
          const int nbm = n0 + c * 32;
          if (c + 1 < BN / 32) prefetch_mask(c + 1);  // next chunk's mask bits fly while this chunk is processed
          if (mask_fast) {
          } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int gm = mrow0 + i * 4 + (lane >> 3), gn = nbm + (lane & 7) * 4;
              if (gm < p.M) {
                const int64_t off = bo * epi.smo + bi * epi.smi + (int64_t)gm * epi.ldm + gn;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (gn + e < p.N) {
                    const float mv = epi.mask_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(epi.mask)[off + e])
                                                   : reinterpret_cast<const float*>(epi.mask)[off + e];
                    mbits |= (mv > 0.f ? 1u : 0u) << (4 * i + e);
                  }
              }
            }
          }
        }
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {  // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar + acc);
        }
        const int nb = n0 + c * 32;
        if (!p.tma_store) {
          // ---- direct stores (outputs whose pitch/base TMA cannot address, e.g. N = 65) ----
          if (m < p.M && nb < p.N) {
            if (p.raw) {
              float* prow = partial + (((int64_t)split * p.batch + b) * p.M + m) * p.N;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) prow[nb + j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                epi_store4(epi, b, m, nb + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
            }
          }
          continue;
        }
        if (mrow0 >= p.M || nb >= p.N) continue;  // whole 32x32 block is outside (warp-uniform)
        // ---- staged TMA-store epilogue ----
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (!p.raw) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= epi.alpha;
          if (epi.bias) {
            if (epi.bias_vec4 && nb + 32 <= p.N) {  // eight 16-byte broadcast loads instead of 32 scalar ones
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(epi.bias + nb) + j);
                v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (nb + j < p.N) ? __ldg(epi.bias + nb + j) : 0.f;
            }
          }
          if (epi.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (epi.bits_in) {  // packed ReLU mask: this lane's row, this chunk's 32 columns = one word
            const uint32_t w = mwords[c];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
          }
          if (epi.bits_out && m < p.M) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
            epi.bits_out[(int64_t)m * epi.ld_bits + (nb >> 5)] = w;
          }
        }
        // two staging slots per warp (a deeper ring — 3 fp32 / 6 bf16 slots — was measured: no gain)
        const uint32_t bufsel = store_i & 1;
        ++store_i;
        uint8_t* s32 = stage_base + bufsel * L::EPI_F32;
        uint8_t* s16 = stage_base + 2 * L::EPI_F32 + bufsel * L::EPI_BF16;
        const uint32_t s32a = smem_u32(s32), s16a = smem_u32(s16);
        // the TMA store that last read this buffer (two chunks ago) must have finished reading it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        if (want_f32 || masked) {
#pragma unroll
          for (int j = 0; j < 8; ++j)  // SWIZZLE_128B: 16-byte chunk j of row r lands at chunk j ^ (r & 7)
            sts128(s32a + lane * 128 + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (masked) {
          if (mask_fast) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&mraw[i].x);
              const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&mraw[i].y);
              mbits |= ((__low2float(lo) > 0.f ? 1u : 0u) | (__high2float(lo) > 0.f ? 2u : 0u) |
                        (__low2float(hi) > 0.f ? 4u : 0u) | (__high2float(hi) > 0.f ? 8u : 0u)) << (4 * i);
            }
          }
          // second pass in a coalesced layout: lane -> (row = i*4 + lane/8, 16-byte chunk = lane%8)
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3), ch = lane & 7;
            const uint32_t sp = s32a + row * 128 + ((ch ^ (row & 7)) << 4);
            float4 x = lds128(sp);
            const uint32_t mb = mbits >> (4 * i);
            x.x = (mb & 1u) ? x.x : 0.f; x.y = (mb & 2u) ? x.y : 0.f;
            x.z = (mb & 4u) ? x.z : 0.f; x.w = (mb & 8u) ? x.w : 0.f;
            if (want_f32) sts128(sp, x.x, x.y, x.z, x.w);
            if (want_bf16) {  // SWIZZLE_64B: 16-byte chunk q of row r lands at chunk q ^ ((r >> 1) & 3)
              uint2 u;
              u.x = pack_bf16(x.x, x.y);
              u.y = pack_bf16(x.z, x.w);
              sts64u(s16a + row * 64 + (((ch >> 1) ^ ((row >> 1) & 3)) << 4) + ((ch & 1) << 3), u.x, u.y);
            }
          }
        } else if (want_bf16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(v[8 * q], v[8 * q + 1]);
            u.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
            u.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]);
            u.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
            sts128u(s16a + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), u.x, u.y, u.z, u.w);
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (p.raw) {
            tma_store_4d(&tmC, s32, nb, mrow0, b, split);
          } else {
            if (want_f32) tma_store_4d(&tmC, s32, nb, mrow0, bi, bo);
            if (want_bf16) tma_store_4d(&tmCb, s16, nb, mrow0, bi, bo);
          }
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();  // global writes complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();  // no CTA leaves while its peer can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<L::TMEM_COLS>(tmem_base);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 4-D tensor map: dims {inner (contiguous), rows, batch_inner, batch_outer}; strides in elements.
static int make_map(CUtensorMap* map, CUtensorMapDataType dt, int es, CUtensorMapSwizzle sw, const void* base,
                    int64_t inner, int64_t rows, int64_t ld, int64_t n_bi, int64_t s_bi, int64_t n_bo, int64_t s_bo,
                    int box_inner, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NPT_REQUIRE(enc != nullptr, "gemm(tc): cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)n_bi, (cuuint64_t)n_bo};
  const int64_t natural = ((rows * ld * es + 15) / 16) * 16;
  cuuint64_t strides[3] = {(cuuint64_t)(ld * es), (cuuint64_t)(n_bi > 1 ? s_bi * es : natural),
                           (cuuint64_t)(n_bo > 1 ? s_bo * es : natural)};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NPT_REQUIRE(r == CUDA_SUCCESS, "gemm(tc): cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld rows=%lld ld=%lld)",
              (int)r, (long long)inner, (long long)rows, (long long)ld);
  return 0;
}
static int make_map_bf16(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int64_t n_bi,
                         int64_t s_bi, int64_t n_bo, int64_t s_bo, int box_inner, int box_rows) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, base, inner, rows, ld, n_bi, s_bi,
                  n_bo, s_bo, box_inner, box_rows);
}

int gemm_tc_supported(const npt_gemm_args& a) {
  NPT_REQUIRE(a.mode == NPT_GEMM_BF16_TC, "gemm: mode %d is not implemented (use 0 = fp32 or 1 = bf16 tcgen05)", a.mode);
  NPT_REQUIRE(aligned16(a.A) && aligned16(a.B), "gemm(tc): operand base pointers must be 16-byte aligned");
  NPT_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm(tc): bf16 leading dimensions must be multiples of 8 (lda=%lld ldb=%lld)",
              (long long)a.lda, (long long)a.ldb);
  NPT_REQUIRE(a.stride_a_outer % 8 == 0 && a.stride_a_inner % 8 == 0 && a.stride_b_outer % 8 == 0 && a.stride_b_inner % 8 == 0,
              "gemm(tc): batch strides must be multiples of 8 elements");
  return 0;
}

// Tile width: minimise the operand bytes a row of tiles pulls through L2 (each tile loads 128 rows
// of A and BN rows of B per k-block), ties broken towards less padded MMA work.
static int tc_bn(const npt_gemm_args& a) {
  if (a.N <= 64) return 64;
  const int cand[3] = {128, 192, 256};
  int best = 128;
  int64_t best_cost = -1;
  for (int i = 0; i < 3; ++i) {
    const int64_t nt = cdiv(a.N, cand[i]);
    const int64_t cost = nt * (cand[i] + 128) * 1024 + nt * cand[i];
    if (best_cost < 0 || cost < best_cost) { best = cand[i]; best_cost = cost; }
  }
  return best;
}

int gemm_tc_splits(const npt_gemm_args& a) {
  if (a.relu_bits_out || a.mask_bits) return 1;  // the packed-mask epilogue runs in the GEMM kernel itself
  const int bn = tc_bn(a);
  const int64_t tiles = cdiv(a.M, TC_BM) * cdiv(a.N, bn) * a.batch;
  const int64_t kblocks = cdiv(a.K, TC_BK);
  const int64_t target = (int64_t)sm_count();
  if (tiles * 2 > target || kblocks < 8) return 1;
  int64_t s = (target + tiles - 1) / tiles;
  if (s > kblocks / 4) s = kblocks / 4;  // >= 4 k-blocks (256 k) per split
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  const int64_t kps = cdiv(kblocks, s);
  return (int)cdiv(kblocks, kps);  // no empty split
}

template <int BN, bool A_MN, bool B_MN, int CL>
static int launch_tc_cl(const TcParams& p, const GemmEpilogue& epi, float* partial, const CUtensorMap& tmA,
                        const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCb, cudaStream_t st) {
  using L = TcSmem<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  int grid = sm_count() - sm_margin();
  if (grid < 1) grid = 1;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (CL == 1) {
    kern<<<grid, TC_THREADS, L::TOTAL, st>>>(tmA, tmB, tmC, tmCb, p, epi, partial);
  } else {
    grid -= grid % CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmCb, p, epi, partial));
  }
  NPT_LAUNCH_CHECK();
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const TcParams& p, const GemmEpilogue& epi, float* partial, const CUtensorMap& tmA,
                     const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCb, int cl, cudaStream_t st) {
  if (cl == 2) return launch_tc_cl<BN, A_MN, B_MN, 2>(p, epi, partial, tmA, tmB, tmC, tmCb, st);
  return launch_tc_cl<BN, A_MN, B_MN, 1>(p, epi, partial, tmA, tmB, tmC, tmCb, st);
}

// CTA pairs with a multicast B tile.  MEASURED (profiles/README.md): on B200 a 2-CTA multicast does not
// lower the time of these GEMMs (16384x1536x384: 40.0 us vs 39.0 us without; 8192^3: 907 vs 927 us) —
// consistent with the microarchitecture notes (multicast ~ unicast for clusters <= 4) — so the path
// is OFF by default and kept, tested, behind NUMPITRON_GEMM_CLUSTER=1 for larger-cluster work.
static int tc_cluster(const npt_gemm_args& a, int bn, int splits) {
  const char* on = getenv("NUMPITRON_GEMM_CLUSTER");
  if (!on || on[0] != '1' || bn < 128) return 1;
  const int64_t m_tiles = cdiv(a.M, TC_BM);
  if (m_tiles % 2 != 0) return 1;
  const int64_t units = m_tiles / 2 * cdiv(a.N, bn) * a.batch * splits;
  return units >= sm_count() / 2 ? 2 : 1;
}

int gemm_tc_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st) {
  const int bn = tc_bn(a);
  const int64_t n_bi = a.batch_inner, n_bo = cdiv(a.batch, a.batch_inner);
  CUtensorMap tmA, tmB, tmC, tmCb;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmCb, 0, sizeof(tmCb));
  int rc;
  if (a.trans_a) rc = make_map_bf16(&tmA, a.A, a.M, a.K, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 64, 64);
  else           rc = make_map_bf16(&tmA, a.A, a.K, a.M, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 64, TC_BM);
  if (rc) return rc;
  const int cl = tc_cluster(a, bn, splits);  // 2: each CTA of a pair loads half of the B tile (multicast)
  if (a.trans_b) rc = make_map_bf16(&tmB, a.B, a.K, a.N, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 64, bn / cl);
  else           rc = make_map_bf16(&tmB, a.B, a.N, a.K, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 64, 64 / cl);
  if (rc) return rc;

  TcParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch; p.splits = splits;
  const int kblocks = (int)cdiv(a.K, TC_BK);
  p.kb_per_split = (int)cdiv(kblocks, splits);
  p.m_tiles = (int)cdiv(a.M, TC_BM);
  p.n_tiles = (int)cdiv(a.N, bn);
  const int64_t total = (int64_t)p.m_tiles * p.n_tiles * a.batch * splits;
  NPT_REQUIRE(total < ((int64_t)1 << 31), "gemm(tc): too many tiles");
  p.total_tiles = (int)total;
  p.raw = splits > 1;
  NPT_REQUIRE(!(p.raw && (a.relu_bits_out || a.mask_bits)), "gemm(tc): packed ReLU bits are not supported with split-K");
  // Can the outputs be written by TMA?  (16-byte aligned base, pitches and batch strides)
  if (p.raw) {
    p.tma_store = (a.N % 4 == 0) && aligned16(partial);
    if (p.tma_store) {
      rc = make_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B, partial, a.N, a.M, a.N, a.batch,
                    (int64_t)a.M * a.N, splits, (int64_t)a.batch * a.M * a.N, 32, 32);
      if (rc) return rc;
    }
  } else {
    const bool c_ok = !a.C || (aligned16(a.C) && a.ldc % 4 == 0 && a.stride_c_outer % 4 == 0 && a.stride_c_inner % 4 == 0);
    const bool cb_ok = !a.C_bf16 || (aligned16(a.C_bf16) && a.ldcb % 8 == 0 && a.stride_cb_outer % 8 == 0 &&
                                     a.stride_cb_inner % 8 == 0);
    p.tma_store = c_ok && cb_ok;
    NPT_REQUIRE(p.tma_store || !(a.relu_bits_out || a.mask_bits),
                "gemm(tc): packed ReLU bits need TMA-storable outputs (16-byte aligned base and pitches)");
    if (p.tma_store && a.C) {
      rc = make_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, CU_TENSOR_MAP_SWIZZLE_128B, a.C, a.N, a.M, a.ldc, n_bi,
                    a.stride_c_inner, n_bo, a.stride_c_outer, 32, 32);
      if (rc) return rc;
    }
    if (p.tma_store && a.C_bf16) {
      rc = make_map(&tmCb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B, a.C_bf16, a.N, a.M, a.ldcb, n_bi,
                    a.stride_cb_inner, n_bo, a.stride_cb_outer, 32, 32);
      if (rc) return rc;
    }
  }
  const bool a_mn = a.trans_a != 0, b_mn = a.trans_b == 0;
#define NPT_TC_CASE(BNV)                                                                            \
  if (a_mn && b_mn) return launch_tc<BNV, true, true>(p, epi, partial, tmA, tmB, tmC, tmCb, cl, st);     \
  if (a_mn && !b_mn) return launch_tc<BNV, true, false>(p, epi, partial, tmA, tmB, tmC, tmCb, cl, st);   \
  if (!a_mn && b_mn) return launch_tc<BNV, false, true>(p, epi, partial, tmA, tmB, tmC, tmCb, cl, st);   \
  return launch_tc<BNV, false, false>(p, epi, partial, tmA, tmB, tmC, tmCb, cl, st);
  if (bn == 64) { NPT_TC_CASE(64) }
  if (bn == 192) { NPT_TC_CASE(192) }
  if (bn == 256) { NPT_TC_CASE(256) }
  NPT_TC_CASE(128)
#undef NPT_TC_CASE
}

}  // namespace npt
