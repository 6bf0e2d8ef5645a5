// Device side of the bf16 tensor-core GEMM (see gemm_tc.cu for the description): PTX wrappers, the
// shared-memory layout, the kernel template and its launcher.  Included by gemm_tc.cu (host side)
// and by gemm_tc_bn*.cu, which instantiate the kernels of one tile width each so that they
// compile in parallel.
#pragma once
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
// CTA-pair form (cta_group::2): the box lands in THIS CTA's shared memory, the bytes are credited to
// the mbarrier at the same offset in the pair's leader (the even-ranked CTA): clearing bit 24 of a
// shared::cluster address selects the leader's copy.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrives (once the MMAs issued so far retire) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// plain arrive on the mbarrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
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

// CL = 2: one warp of EACH CTA of the pair executes these (same warp index, same dst offset)
template <int COLS, int CL>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  if (CL == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int COLS, int CL>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  if (CL == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
  else         asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// CTA-pair MMA: M = 256 (128 rows from each CTA's A tile), N columns of B split between the two
// CTAs' shared memory, the accumulator rows in each CTA's own TMEM.  Issued by the leader only.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
// Same wait, but it names the registers of the load it completes as read-write operands, so that no
// use of them can be scheduled above it while another tcgen05.ld (into other registers) is in flight.
__device__ __forceinline__ void tmem_ld_wait_for(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

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
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, M = 128 (one CTA) or
// 256 (CTA pair), N = BN.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7) | (1u << 10)        // a_format = b_format = BF16
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// CL = 2 (CTA pair, cta_group::2): each CTA stages its own 128 rows of A and HALF of the B tile.
template <int BN, int CL>
struct TcSmem {
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;  // power of two >= 2*BN
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = (BN / CL) * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // smem budget: 227 KB - 64 KB of epilogue staging - barriers / alignment slack
  static constexpr int STAGES_FIT = (232448 - 65536 - 1024 - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;  // CL=1: 3/4/5/6 for BN = 256/192/128/64
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  // general epilogue, per warp (4 warps): 2 x (32 rows x 128 B) fp32 staging + 2 x (32 rows x 64 B) bf16 staging
  static constexpr int EPI_F32 = 32 * 128, EPI_BF16 = 32 * 64;
  static constexpr int EPI_WARP_BYTES = 2 * EPI_F32 + 2 * EPI_BF16;
  // fast epilogue, per warp (8 warps): 2 slots of 32 rows x 128 B (32 fp32 or 64 bf16 columns)
  static constexpr int FAST_SLOT_BYTES = 32 * 128;
  static constexpr int FAST_WARP_BYTES = 2 * FAST_SLOT_BYTES;
  static constexpr int EPI_BYTES = 8 * FAST_WARP_BYTES;  // >= 4 * EPI_WARP_BYTES
  static constexpr int EPI_OFFSET = PIPE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
};

// Position of a role's current work unit; see `decode` in the kernel.
struct TileCursor {
  int nt, mr, z;          // n tile, row of units, batch*split index
  int d_nt, d_mr, d_z;    // what one stride adds
  int n_tiles, mrows;
  int zc, b, split, bo, bi;  // cached decomposition of z
  __device__ __forceinline__ void init(int first, int stride, int n_tiles_, int mrows_) {
    n_tiles = n_tiles_; mrows = mrows_;
    nt = first % n_tiles;
    const int r = first / n_tiles;
    mr = r % mrows; z = r / mrows;
    d_nt = stride % n_tiles;
    const int dr = stride / n_tiles;
    d_mr = dr % mrows; d_z = dr / mrows;
    zc = -1; b = split = bo = bi = 0;
  }
  __device__ __forceinline__ void next() {
    nt += d_nt;
    const int carry = nt >= n_tiles ? 1 : 0;
    nt -= carry ? n_tiles : 0;
    mr += d_mr + carry;
    z += d_z;
    if (mr >= mrows) { mr -= mrows; ++z; }
  }
};

struct TcParams {
  int M, N, K, batch, splits, kb_per_split;
  int m_tiles, n_tiles, total_tiles;
  int tma_store;  // 1: staged TMA-store epilogue; 0: direct register stores (unaligned outputs)
  int raw;        // 1: split-K partial (no epilogue math), stored through map C as {N, M, batch, split}
  int peer_store; // 1: the bf16 output is also stored through map P (a peer GPU's buffer); 2: rows in [peer_lo, peer_hi) go to P only
  int peer_lo, peer_hi;
#ifdef NPT_GEMM_DEBUG
  int dbg;        // timing experiments only: 1 = no epilogue stores, 2 = no MMA issue, 4 = no TMA loads
  long long* trace;  // [cta][64] clock64() stamps (slot 0/1: entry / after prologue, then 7 per tile)
#endif
};
#ifdef NPT_GEMM_DEBUG
#define NPT_DBG(p, bit) ((p).dbg & (bit))
#define NPT_TRACE(p, slot) do { if ((p).trace && (slot) < 64) (p).trace[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define NPT_DBG(p, bit) 0
#define NPT_TRACE(p, slot) do { } while (0)
#endif

// CL = 2: CTA pairs (2-CTA clusters, tcgen05 cta_group::2) work on two vertically adjacent tiles
// (m, n) / (m+1, n) as ONE 256 x BN MMA: each CTA stages its own 128 rows of A and half of the B
// tile, the pair's leader (even rank) issues the MMAs, each CTA's TMEM receives its own 128 rows of
// the accumulator and each CTA runs its own epilogue.  What this buys is shared-memory fill
// bandwidth: a 1-CTA 128x256 tile needs 48 KB per 64-deep k-block, ~750 cycles of the SM's
// ~64 B/clk L2->SM port against ~590 cycles of MMA (measured: loads alone 0.375 us / k-block, MMAs
// alone 0.31 us); as a pair each CTA needs 32 KB (~500 cycles), so the tensor pipe is the limit.
// (A 2-CTA TMA multicast of B was tried first and does not help: the bytes still enter both SMs.)
//
// EPI selects the epilogue.  EPI_GENERAL (6 warps) handles everything: direct stores for outputs TMA
// cannot address, tensor ReLU masks, fp32 + bf16 outputs at once.  Every other value is the fast
// epilogue (10 warps: two epilogue warps per TMEM lane quarter, alternating 64-column pairs) for
// single-output launches through TMA stores.  A clock64() trace of the general epilogue on the
// 16384x1536x384 GEMM showed ~1000 cycles per 32-column chunk against ~460 cycles of MMA work per
// chunk — ~400 issued instructions once every optional step was if-converted, plus ~600 cycles of
// exposed latencies (store-slot wait, fence, TMA issue) on a warp that runs alone on its scheduler.
// The fast epilogue therefore compiles the optional steps in or out (EPI bit mask; EPI_DYNAMIC
// keeps them as run-time tests for combinations outside the instantiated list), keeps the next
// chunk's tcgen05.ld in flight, stores bf16 in 64-column boxes, and doubles the warps.
enum : int {
  EPI_GENERAL = -1, EPI_F32 = 1, EPI_BF16 = 2, EPI_BIAS = 4, EPI_RELU = 8, EPI_BITS_IN = 16, EPI_BITS_OUT = 32,
  EPI_ALPHA = 64, EPI_DYNAMIC = 128
};
constexpr int TC_THREADS_FAST = 320;

template <int BN, bool A_MN, bool B_MN, int CL, int EPI>
__global__ void __launch_bounds__(EPI == EPI_GENERAL ? TC_THREADS : TC_THREADS_FAST, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCb,
               const __grid_constant__ CUtensorMap tmP, const TcParams p, const GemmEpilogue epi,
               float* __restrict__ partial) {
  static_assert(CL == 1 || EPI != EPI_GENERAL, "CTA pairs run the fast epilogue only");
  using L = TcSmem<BN, CL>;
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

  if (threadIdx.x == 0) NPT_TRACE(p, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) {
      if (epi.C || p.raw) tma_prefetch_desc(&tmC);
      if (epi.Cb && !p.raw) tma_prefetch_desc(&tmCb);
      if (p.peer_store) tma_prefetch_desc(&tmP);
    }
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);   // one commit (CL = 2: the leader's, delivered to both CTAs)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar + a, 1);
      mbar_init(tmem_empty_bar + a, (EPI == EPI_GENERAL ? 4 : 8) * CL);  // one arrive per epilogue warp (of the pair)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<L::TMEM_COLS, CL>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();  // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above is private set-up; from here on global memory is read and written, so the
  // preceding kernels must have completed.  The grid is persistent (no CTA waits to be scheduled),
  // so the next kernel may be launched right away and do its own set-up behind this one.
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) NPT_TRACE(p, 1);
  const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  // work units: one tile (CL = 1) or one vertical pair of tiles (CL = 2) per step of the persistent loop
  const int unit_first = blockIdx.x / CL, unit_stride = gridDim.x / CL, units = p.total_tiles / CL;

  // unit -> (n_tile fastest, row of units, z = batch*splits).  Every role walks the same sequence
  // t = unit_first, unit_first + unit_stride, ...; the cursor advances by additions only — integer
  // divisions by run-time values cost a single thread hundreds of cycles per tile, which showed up
  // as ~700 idle cycles of the MMA issuer between tiles in the clock64() trace.
  auto decode = [&](TileCursor& c, int& m0, int& n0, int& b, int& split, int& bi, int& bo, int& kb_begin, int& nkb) {
    if (c.z != c.zc) {  // batch / split decomposition: only when z changes
      c.zc = c.z;
      c.b = c.z / p.splits;
      c.split = c.z - c.b * p.splits;
      c.bo = c.b / epi.batch_inner;
      c.bi = c.b - c.bo * epi.batch_inner;
    }
    b = c.b; split = c.split; bo = c.bo; bi = c.bi;
    m0 = (c.mr * CL + cta_rank) * TC_BM;
    n0 = c.nt * BN;
    kb_begin = split * p.kb_per_split;
    int kb_end = kb_begin + p.kb_per_split;
    if (kb_end > kblocks) kb_end = kblocks;
    nkb = kb_end - kb_begin;  // host guarantees >= 1
  };
  TileCursor cur;
  cur.init(unit_first, unit_stride, p.n_tiles, p.m_tiles / CL);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles
      int trace_tile = 0;
      for (int t = unit_first; t < units; t += unit_stride, cur.next(), ++trace_tile) {
        int m0, n0, b, split, bi, bo, kb_begin, nkb;
        decode(cur, m0, n0, b, split, bi, bo, kb_begin, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(empty_bar + s, ((it / TC_STAGES) & 1) ^ 1);
          if (i == 0) NPT_TRACE(p, 2 + trace_tile * 7 + 0);
          if (NPT_DBG(p, 4)) { mbar_arrive(full_bar + s); continue; }
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const int k0 = (kb_begin + i) * TC_BK;
          if (CL > 1) {
            // both CTAs of the pair load (own A rows, own half of B) and credit the LEADER's barrier
            if (cta_rank == 0) mbar_expect_tx(full_bar + s, 2 * L::STAGE_BYTES);
            if (A_MN) {
              tma_load_4d_2sm(&tmA, full_bar + s, sa, m0, k0, bi, bo);
              tma_load_4d_2sm(&tmA, full_bar + s, sa + 64 * TC_BK * 2, m0 + 64, k0, bi, bo);
            } else {
              tma_load_4d_2sm(&tmA, full_bar + s, sa, k0, m0, bi, bo);
            }
            const int nh = n0 + cta_rank * (BN / 2);  // this CTA's half of the tile's columns
            if (B_MN) {
#pragma unroll
              for (int c = 0; c < BN / 128; ++c)
                tma_load_4d_2sm(&tmB, full_bar + s, sb + c * 64 * TC_BK * 2, nh + c * 64, k0, bi, bo);
            } else {
              tma_load_4d_2sm(&tmB, full_bar + s, sb, k0, nh, bi, bo);
            }
            continue;
          }
          mbar_expect_tx(full_bar + s, L::STAGE_BYTES);
          if (A_MN) {
            tma_load_4d(&tmA, full_bar + s, sa, m0, k0, bi, bo);
            tma_load_4d(&tmA, full_bar + s, sa + 64 * TC_BK * 2, m0 + 64, k0, bi, bo);
          } else {
            tma_load_4d(&tmA, full_bar + s, sa, k0, m0, bi, bo);
          }
          if (B_MN) {
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
    // ===== MMA issuer (CL = 2: the pair's leader only) =====
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM * CL, BN, A_MN, B_MN);
      uint32_t it = 0, tile_i = 0;
      for (int t = unit_first; t < units; t += unit_stride, cur.next(), ++tile_i) {
        int m0, n0, b, split, bi, bo, kb_begin, nkb;
        decode(cur, m0, n0, b, split, bi, bo, kb_begin, nkb);
        const uint32_t acc = tile_i & 1;
        mbar_wait(tmem_empty_bar + acc, ((tile_i >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        NPT_TRACE(p, 2 + tile_i * 7 + 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(full_bar + s, (it / TC_STAGES) & 1);
          tc_fence_after();
          if (i == 0) NPT_TRACE(p, 2 + tile_i * 7 + 2);
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          // K-major: 16 k = 32 B inside the 128 B swizzle row.  MN-major: 16 k = two 8-k groups = 2048 B.
          const uint64_t da = A_MN ? make_smem_desc(sa, 64 * TC_BK * 2, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb, 64 * TC_BK * 2, 1024) : make_smem_desc(sb, 16, 1024);
          constexpr uint32_t adv_a = (A_MN ? 2048 : 32) >> 4, adv_b = (B_MN ? 2048 : 32) >> 4;
          if (!NPT_DBG(p, 2))
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if (CL > 1) umma_bf16_2sm(d_tmem, da + (uint64_t)(k * adv_a), db + (uint64_t)(k * adv_b), idesc, (i > 0 || k > 0) ? 1u : 0u);
            else        umma_bf16(d_tmem, da + (uint64_t)(k * adv_a), db + (uint64_t)(k * adv_b), idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above retire
          if (CL > 1) umma_commit_2sm(empty_bar + s, (uint16_t)3); else umma_commit(empty_bar + s);
        }
        if (CL > 1) umma_commit_2sm(tmem_full_bar + acc, (uint16_t)3); else umma_commit(tmem_full_bar + acc);  // accumulator complete
        NPT_TRACE(p, 2 + tile_i * 7 + 3);
      }
    }
  } else if constexpr (EPI != EPI_GENERAL) {
    // ===== epilogue (fast path): warps 2..9; warp % 4 = TMEM lane quarter, (warp - 2) / 4 = which
    // of the tile's 64-column pairs (even / odd) the warp handles =====
    constexpr bool DYN = (EPI & EPI_DYNAMIC) != 0;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int PAIRS = BN / 64;           // 64-column pairs per tile
    constexpr int MAXP = (PAIRS + 1) / 2;    // pairs per warp (upper bound)
    uint8_t* stage_base = smem + L::EPI_OFFSET + (warp - 2) * L::FAST_WARP_BYTES;
    const uint32_t stage_a = smem_u32(stage_base);
    const bool raw = p.raw != 0;
    // one output kind per launch: fp32 (32-column chunks, one TMA store each) or bf16 (the two chunks
    // of a pair share one 128-byte-row staging slot and one TMA store)
    const bool out_bf16 = DYN ? (!raw && epi.Cb != nullptr) : (EPI & EPI_BF16) != 0;
    const bool use_alpha = DYN ? !raw : (EPI & EPI_ALPHA) != 0;
    const bool use_bias = DYN ? (!raw && epi.bias != nullptr) : (EPI & EPI_BIAS) != 0;
    const bool use_relu = DYN ? (!raw && epi.relu) : (EPI & EPI_RELU) != 0;
    const bool use_bits_in = DYN ? (!raw && epi.bits_in != nullptr) : (EPI & EPI_BITS_IN) != 0;
    const bool use_bits_out = DYN ? (!raw && epi.bits_out != nullptr) : (EPI & EPI_BITS_OUT) != 0;
    const float alpha = epi.alpha;
    const float* bias = epi.bias;
    const uint32_t* bits_in = epi.bits_in;
    uint32_t* bits_out = epi.bits_out;
    const int64_t ld_bits = epi.ld_bits;
    const int M = p.M, N = p.N;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t swz = (uint32_t)(lane & 7);
    const CUtensorMap* out_map = out_bf16 ? &tmCb : &tmC;
    uint32_t tile_i = 0, store_i = 0;  // store_i counts committed TMA-store groups (staging slot parity)
    for (int t = unit_first; t < units; t += unit_stride, cur.next(), ++tile_i) {
      int m0, n0, b, split, bi, bo, kb_begin, nkb;
      decode(cur, m0, n0, b, split, bi, bo, kb_begin, nkb);
      const uint32_t acc = tile_i & 1;
      const int mrow0 = m0 + quarter * 32;  // first row of this warp's 32-row band
      const int m = mrow0 + lane;
      const int c2 = raw ? b : bi, c3 = raw ? split : bo;  // batch coordinates of the TMA store
      // pairs of this warp that hold any column of the matrix (warp-uniform)
      int npairs = 0;
#pragma unroll
      for (int pp = 0; pp < MAXP; ++pp)
        if (half + 2 * pp < PAIRS && n0 + (half + 2 * pp) * 64 < N && mrow0 < M) npairs = pp + 1;
      // fetched before the wait on the accumulator: this lane's packed ReLU-mask words (one per
      // 32-column chunk).  (The bias is read with warp-uniform 16-byte loads inside the chunk: the
      // whole vector stays in L1 for the life of the kernel.)
      uint32_t mwords[MAXP][2];
      if (use_bits_in) {
#pragma unroll
        for (int pp = 0; pp < MAXP; ++pp)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int nb = n0 + (half + 2 * pp) * 64 + h * 32;
            mwords[pp][h] = (pp < npairs && m < M && nb < N) ? __ldg(bits_in + (int64_t)(nb >> 5) * ld_bits + m) : 0u;
          }
      }
      mbar_wait(tmem_full_bar + acc, (tile_i >> 1) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) NPT_TRACE(p, 2 + tile_i * 7 + 4);
      const uint32_t taddr = lane_taddr + acc * BN + (uint32_t)(half * 64);

      auto release_acc = [&]() {  // this warp has read everything it needs from the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL > 1) mbar_arrive_cluster(tmem_empty_bar + acc, 0); else mbar_arrive(tmem_empty_bar + acc);
        }
        if (warp == 2 && lane == 0) NPT_TRACE(p, 2 + tile_i * 7 + 5);
      };
      // one 32-column chunk: epilogue math on this lane's row, then staging (+ TMA store)
      auto chunk = [&](uint32_t (&r)[32], const int pp, const int h) {
        const int nb = n0 + (half + 2 * pp) * 64 + h * 32;
        const bool inside = nb < N;  // warp-uniform; only the second chunk of a pair can be outside
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (use_alpha) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= alpha;
        }
        if (use_bias) {  // warp-uniform addresses: broadcast loads, L1 hits after the first tile row
          if (nb + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + nb) + j);
              v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (nb + j < N) ? __ldg(bias + nb + j) : 0.f;
          }
        }
        if (use_relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (use_bits_in) {  // packed ReLU mask: this lane's row, this chunk's 32 columns = one word
          const uint32_t w = mwords[pp][h];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
        }
        if (use_bits_out) {
          uint32_t w = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
          if (inside && m < M) bits_out[(int64_t)(nb >> 5) * ld_bits + m] = w;  // word-major: a warp's 32 rows are one 128-byte line
        }
        if (NPT_DBG(p, 1)) {
          float acc_dbg = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc_dbg += v[j];
          if (acc_dbg == 123.456f) epi.C[0] = 1.f;  // keeps the math alive
          return;
        }
        const uint32_t slot = stage_a + (store_i & 1) * L::FAST_SLOT_BYTES;
        const bool opens = !out_bf16 || h == 0;   // first write into the slot
        const bool closes = !out_bf16 || h == 1;  // slot complete: store it
        if (opens) {
          // the TMA store that last read this slot (two stores ago) must have finished reading it
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
        // SWIZZLE_128B staging: 16-byte chunk j of row r lands at chunk j ^ (r & 7)
        if (!out_bf16) {
          if (inside) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(slot + lane * 128 + (((uint32_t)j ^ swz) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else if (inside) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128u(slot + lane * 128 + (((uint32_t)(h * 4 + q) ^ swz) << 4), pack_bf16(v[8 * q], v[8 * q + 1]),
                    pack_bf16(v[8 * q + 2], v[8 * q + 3]), pack_bf16(v[8 * q + 4], v[8 * q + 5]),
                    pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        }
        if (closes && (inside || out_bf16)) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            // tensor parallelism: the staged tile also (peer_store 1) or instead (2: rows the peer owns) goes to
            // the peer GPU (posted NVLink writes)
            const bool peer_only = p.peer_store == 2 && mrow0 >= p.peer_lo && mrow0 < p.peer_hi;
            if (!peer_only) tma_store_4d(out_map, stage_base + (store_i & 1) * L::FAST_SLOT_BYTES, out_bf16 ? nb - 32 : nb, mrow0, c2, c3);
            if (p.peer_store == 1 || peer_only) tma_store_4d(&tmP, stage_base + (store_i & 1) * L::FAST_SLOT_BYTES, nb - 32, mrow0, c2, c3);
            tma_store_commit();
          }
          ++store_i;
        }
      };

      if (npairs == 0) {
        release_acc();
      } else {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(taddr, ra);
#pragma unroll
        for (int pp = 0; pp < MAXP; ++pp) {
          if (pp < npairs) {  // warp-uniform
            tmem_ld_wait_for(ra);
            tmem_ld_32x32(taddr + (uint32_t)(pp * 128 + 32), rb);
            chunk(ra, pp, 0);
            tmem_ld_wait_for(rb);
            if (pp + 1 < npairs) tmem_ld_32x32(taddr + (uint32_t)((pp + 1) * 128), ra);
            else release_acc();
            chunk(rb, pp, 1);
          }
        }
      }
      if (warp == 2 && lane == 0) NPT_TRACE(p, 2 + tile_i * 7 + 6);
    }
    if (lane == 0) tma_store_wait_all();  // global writes complete before the CTA retires
  } else if (warp < 6) {
    // ===== epilogue (general path): warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int quarter = warp & 3;
    uint8_t* stage_base = smem + L::EPI_OFFSET + (warp - 2) * L::EPI_WARP_BYTES;
    const bool want_f32 = p.raw || epi.C != nullptr;
    const bool want_bf16 = !p.raw && epi.Cb != nullptr;
    const bool masked = !p.raw && epi.mask != nullptr;
    uint32_t tile_i = 0, store_i = 0;  // store_i counts committed TMA-store groups (staging buffer parity)
    for (int t = unit_first; t < units; t += unit_stride, cur.next(), ++tile_i) {
      int m0, n0, b, split, bi, bo, kb_begin, nkb;
      decode(cur, m0, n0, b, split, bi, bo, kb_begin, nkb);
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
          mwords[cc] = (m < p.M && n0 + cc * 32 < p.N) ? __ldg(epi.bits_in + (int64_t)((n0 >> 5) + cc) * epi.ld_bits + m) : 0u;
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
            epi.bits_out[(int64_t)(nb >> 5) * epi.ld_bits + m] = w;
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
  if (CL > 1) cluster_sync();  // no CTA leaves while its peer can still signal it / read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<L::TMEM_COLS, CL>(tmem_base);
  }
}


template <int BN, bool A_MN, bool B_MN, int CL, int EPI>
static int launch_tc_cl(const TcParams& p, const GemmEpilogue& epi, float* partial, const CUtensorMap& tmA,
                        const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCb, const CUtensorMap& tmP, cudaStream_t st) {
  using L = TcSmem<BN, CL>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL, EPI>;
  constexpr int threads = EPI == EPI_GENERAL ? TC_THREADS : TC_THREADS_FAST;
  static bool attr_set = false;
  if (!attr_set) {
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  int grid = sm_count() - sm_margin();
  if (grid < 1) grid = 1;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (CL == 1) {
    NPT_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(threads), L::TOTAL, st, tmA, tmB, tmC, tmCb, tmP, p, epi, partial));
  } else {
    grid -= grid % CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    NPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmCb, tmP, p, epi, partial));
  }
  NPT_LAUNCH_CHECK();
  return 0;
}

// The epilogue combinations of the training step (nn/linear.py, nn/mlp.py, nn/attention.py in bf16
// mode) get their own kernels; any other fast-path combination runs the EPI_DYNAMIC kernel.
#define NPT_TC_FAST_KINDS(X)                                                                        \
  X(EPI_F32) X(EPI_BF16) X(EPI_F32 | EPI_BIAS) X(EPI_BF16 | EPI_BIAS)                                \
  X(EPI_BF16 | EPI_BIAS | EPI_RELU | EPI_BITS_OUT) X(EPI_BF16 | EPI_BITS_IN)

inline int tc_epi_kind(int flags) {
#define NPT_X(K) if (flags == (K)) return (K);
  NPT_TC_FAST_KINDS(NPT_X)
#undef NPT_X
  return EPI_DYNAMIC;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const TcParams& p, const GemmEpilogue& epi, float* partial, const CUtensorMap& tmA,
                     const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmCb, const CUtensorMap& tmP, int cl,
                     int kind, cudaStream_t st) {
  if (kind == EPI_GENERAL) return launch_tc_cl<BN, A_MN, B_MN, 1, EPI_GENERAL>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, st);
  if constexpr (BN >= 128 && !(BN == 192 && B_MN)) {  // CTA pairs (cta_group::2)
    if (cl == 2) {
#define NPT_X(K) if (kind == (K)) return launch_tc_cl<BN, A_MN, B_MN, 2, (K)>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, st);
      NPT_TC_FAST_KINDS(NPT_X)
#undef NPT_X
      return launch_tc_cl<BN, A_MN, B_MN, 2, EPI_DYNAMIC>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, st);
    }
  }
  NPT_REQUIRE(cl == 1, "gemm(tc): no CTA-pair kernel for this tile shape");
#define NPT_X(K) if (kind == (K)) return launch_tc_cl<BN, A_MN, B_MN, 1, (K)>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, st);
  NPT_TC_FAST_KINDS(NPT_X)
#undef NPT_X
  return launch_tc_cl<BN, A_MN, B_MN, 1, EPI_DYNAMIC>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, st);
}

// One tile width per translation unit (gemm_tc_bn*.cu instantiate this).
template <int BN>
int launch_tc_bn(const TcParams& p, const GemmEpilogue& epi, float* partial, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const CUtensorMap& tmC, const CUtensorMap& tmCb, const CUtensorMap& tmP, bool a_mn, bool b_mn, int cl, int kind,
                 cudaStream_t st) {
  if (a_mn && b_mn) return launch_tc<BN, true, true>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, cl, kind, st);
  if (a_mn && !b_mn) return launch_tc<BN, true, false>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, cl, kind, st);
  if (!a_mn && b_mn) return launch_tc<BN, false, true>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, cl, kind, st);
  return launch_tc<BN, false, false>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, cl, kind, st);
}

}  // namespace npt
