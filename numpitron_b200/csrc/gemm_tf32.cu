// fp32-accurate GEMM on the sm_100a tensor cores: 3xTF32 (NPT_GEMM_TF32X3_TC).
//
// tcgen05 has no fp32 input kind; `kind::tf32` reads fp32 words and uses their top 19 bits.  Every fp32
// operand x is therefore split as x = hi + lo with hi = x & 0xffffe000 (exactly a TF32 number) and
// lo = x - hi (exact in fp32, |lo| <= 2^-10 |x|), and the product is accumulated as
//     A*B  ~=  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi            (the dropped A_lo*B_lo term is ~2^-20 relative)
// three tcgen05.mma per k-step into fp32 TMEM accumulators (main term and cross terms apart, added in the epilogue).  The split is done INSIDE the kernel:
// TMA brings the raw fp32 tiles into 128B-swizzled shared memory, four "converter" warps rewrite each tile
// in place as `hi` and write `lo` into a second tile at the same (swizzled) offsets — the transform is
// element-wise, so the layout never has to be decoded — and only then is the stage handed to the MMA warp.
// Nothing is pre-split in HBM and the operands are read once.
//
// Serves the contractions of the fp32-accurate mode (nn/linear.py:44,55,60; nn/attention.py:48,52,81-92;
// nn/embedding.py:92,99,102) in all four operand layouts without transposes:
//   A K-major  (trans_a=0, stored M x K): one TMA box {32 k, 128 m}
//   A MN-major (trans_a=1, stored K x M): four boxes {32 m, 32 k}
//   B K-major  (trans_b=1, stored N x K): one box {32 k, BN n}
//   B MN-major (trans_b=0, stored K x N): BN/32 boxes {32 n, 32 k}
// The tensor core always sees K-major SWIZZLE_128B tiles: a k-block is 32 fp32 = one 128-byte swizzle row, so a
// {32 mn, 32 k} box of an MN-major operand is exactly the 4 KB that rows [32c, 32c+32) of the K-major tile occupy.
// The converter warp that owns the box reads all of it into registers, transposes 4x4 blocks there and writes
// hi / lo back K-major — in place, with a lane -> block assignment that is bank-conflict-free for both the
// swizzled reads and the swizzled writes.  (kind::tf32 reads MN-major operands only in the 32-byte-atom flavour
// of the 128-byte swizzle; since every element passes through registers for the split anyway, the transpose
// there costs nothing extra and keeps one descriptor format.)
// One CTA per (tile, batch, split): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 =
// converters during the main loop and epilogue (alpha / bias / ReLU / mask through the shared GemmEpilogue)
// afterwards.  Split-K partials go to the workspace and are reduced in fixed order by gemm.cu.
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace npt {
using namespace tc;

constexpr int TF_BM = 128;
constexpr int TF_BK = 32;           // fp32 elements = 128 bytes = one swizzle row
constexpr int TF_THREADS = 192;
constexpr int TF_A_BYTES = TF_BM * TF_BK * 4;   // 16 KB

template <int BN>
struct Tf32Smem {
  static constexpr int B_BYTES = BN * TF_BK * 4;
  static constexpr int STAGE_BYTES = 2 * (TF_A_BYTES + B_BYTES);   // hi + lo of both operands
  static constexpr int STAGES = (232448 - 2048) / STAGE_BYTES > 4 ? 4 : (232448 - 2048) / STAGE_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (3 * STAGES + 1) * 8 + 16 + 1024;
};

struct Tf32Params {
  int M, N, K, batch, splits, kb_per_split, raw;
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Instruction descriptor: tf32 x tf32 -> f32, M = 128, N = BN.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TF_BM >> 4) << 24);
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// One warp, one 4 KB block (32 rows x 128 B, SWIZZLE_128B) of a K-major operand: element-wise split, the
// layout never has to be decoded.
__device__ __forceinline__ void split_block(uint32_t hi_blk, uint32_t lo_blk, int lane) {
  float4 v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = lds128(hi_blk + (r * 32 + lane) * 16);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float hx = tf32_hi(v[r].x), hy = tf32_hi(v[r].y), hz = tf32_hi(v[r].z), hw = tf32_hi(v[r].w);
    sts128(hi_blk + (r * 32 + lane) * 16, hx, hy, hz, hw);
    sts128(lo_blk + (r * 32 + lane) * 16, v[r].x - hx, v[r].y - hy, v[r].z - hz, v[r].w - hw);
  }
}
// One warp, one {32 mn, 32 k} box of an MN-major operand (row = k, 16-byte chunk g = mn 4g..4g+3, stored at
// chunk g ^ (k & 7)) -> rows mn of the K-major tile (chunk c = k 4c..4c+3, stored at chunk c ^ (mn & 7)), split
// into hi / lo.  Lane (q, t) = (lane / 8, lane % 8) owns the 4x4 blocks (c, g) = (t, (t + q + 4p) % 8), p = 0, 1:
// within a quarter-warp the eight physical chunks are distinct for the reads (g ^ (4 (t & 1) + i)) and for the
// writes (t ^ (4 (g & 1) + j)).  Everything is read before anything is written (in place).
__device__ __forceinline__ void split_block_transposed(uint32_t hi_blk, uint32_t lo_blk, int lane) {
  const int t = lane & 7, q = lane >> 3;
  float v[2][4][4];   // [pass][k i][mn j]
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int g = (t + q + 4 * p) & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 4 * t + i;
      const float4 x = lds128(hi_blk + k * 128 + ((g ^ (k & 7)) << 4));
      v[p][i][0] = x.x; v[p][i][1] = x.y; v[p][i][2] = x.z; v[p][i][3] = x.w;
    }
  }
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int g = (t + q + 4 * p) & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mn = 4 * g + j;
      const uint32_t off = mn * 128 + ((t ^ (mn & 7)) << 4);
      const float h0 = tf32_hi(v[p][0][j]), h1 = tf32_hi(v[p][1][j]), h2 = tf32_hi(v[p][2][j]), h3 = tf32_hi(v[p][3][j]);
      sts128(hi_blk + off, h0, h1, h2, h3);
      sts128(lo_blk + off, v[p][0][j] - h0, v[p][1][j] - h1, v[p][2][j] - h2, v[p][3][j] - h3);
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TF_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Tf32Params p,
                 const GemmEpilogue epi, float* __restrict__ partial) {
  using L = Tf32Smem<BN>;
  constexpr int ST = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);   // TMA bytes landed
  uint64_t* conv_bar = full_bar + ST;                                       // hi / lo tiles written
  uint64_t* empty_bar = conv_bar + ST;                                      // MMAs retired
  uint64_t* acc_bar = empty_bar + ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < ST; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(conv_bar + s, 4);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);   // columns [0, BN): A_hi*B_hi, [BN, 2 BN): the two cross terms
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TF_BM;
  const int z = blockIdx.z;
  const int b = z / p.splits, split = z - b * p.splits;
  const int bo = b / epi.batch_inner, bi = b - bo * epi.batch_inner;
  const int kblocks = (p.K + TF_BK - 1) / TF_BK;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > kblocks) kb_end = kblocks;
  const int nkb = kb_end - kb_begin;   // host guarantees >= 1

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % ST;
        mbar_wait(empty_bar + s, ((i / ST) & 1) ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;                       // A (raw -> hi)
        uint8_t* sb = sa + 2 * TF_A_BYTES;                             // B (raw -> hi); the lo tiles follow each
        const int k0 = (kb_begin + i) * TF_BK;
        mbar_expect_tx(full_bar + s, TF_A_BYTES + L::B_BYTES);
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < TF_BM / 32; ++c) tma_load_4d(&tmA, full_bar + s, sa + c * 4096, m0 + c * 32, k0, bi, bo);
        } else {
          tma_load_4d(&tmA, full_bar + s, sa, k0, m0, bi, bo);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_4d(&tmB, full_bar + s, sb + c * 4096, n0 + c * 32, k0, bi, bo);
        } else {
          tma_load_4d(&tmB, full_bar + s, sb, k0, n0, bi, bo);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      // one MMA covers K = 8 fp32 = 32 B inside the 128-byte swizzle row
      constexpr uint32_t adv_a = 32 >> 4, adv_b = 32 >> 4;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % ST;
        mbar_wait(conv_bar + s, (i / ST) & 1);
        fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + 2 * TF_A_BYTES;
        const uint64_t a_hi = make_smem_desc(sa, 16, 1024), a_lo = make_smem_desc(sa + TF_A_BYTES, 16, 1024);
        const uint64_t b_hi = make_smem_desc(sb, 16, 1024), b_lo = make_smem_desc(sb + L::B_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < TF_BK / 8; ++k) {
          const uint64_t oa = (uint64_t)(k * adv_a), ob = (uint64_t)(k * adv_b);
          // The tensor core adds into the TMEM accumulator with truncation: one error of up to an ulp OF THE
          // ACCUMULATOR per MMA.  The cross terms are 2^-11 of the main term, so they get their own accumulator
          // (their truncation errors scale with it) and the main one sees a third of the additions.
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          umma_tf32(tmem + BN, a_lo + oa, b_hi + ob, idesc, acc);
          umma_tf32(tmem + BN, a_hi + oa, b_lo + ob, idesc, 1u);
          umma_tf32(tmem, a_hi + oa, b_hi + ob, idesc, acc);
        }
        umma_commit(empty_bar + s);
      }
      umma_commit(acc_bar);
    }
  } else {
    // ===== converters: raw fp32 tile -> (hi in place, lo in the twin tile) =====
    const int cw = warp - 2;   // 0..3
    for (int i = 0; i < nkb; ++i) {
      const int s = i % ST;
      mbar_wait(full_bar + s, (i / ST) & 1);
      const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
      const uint32_t sb = sa + 2 * TF_A_BYTES;
      // 4 KB blocks: A has 4, B has BN / 32; warp cw takes blocks cw, cw + 4, ...
      if (A_MN) split_block_transposed(sa + cw * 4096, sa + TF_A_BYTES + cw * 4096, lane);
      else      split_block(sa + cw * 4096, sa + TF_A_BYTES + cw * 4096, lane);
#pragma unroll
      for (int blk = cw; blk < BN / 32; blk += 4) {
        if (B_MN) split_block_transposed(sb + blk * 4096, sb + L::B_BYTES + blk * 4096, lane);
        else      split_block(sb + blk * 4096, sb + L::B_BYTES + blk * 4096, lane);
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar + s);
    }
    // ===== epilogue: warp % 4 = TMEM lane quarter =====
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    mbar_wait(acc_bar, 0);
    fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;   // warp-uniform
      uint32_t r[32], rx[32];
      tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(BN + c * 32), rx);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rx[j]));
      if (m >= p.M) continue;
      if (p.raw) {
        float* prow = partial + (((int64_t)split * p.batch + b) * p.M + m) * p.N;
        if ((p.N & 3) == 0 && nb + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(prow + nb + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < p.N) prow[nb + j] = __uint_as_float(r[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          epi_store4(epi, b, m, nb + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                    __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    tmem_dealloc<2 * BN>(tmem);
  }
}

// 4-D fp32 tensor map {inner, rows, batch_inner, batch_outer}; 32 fp32 per 128-byte swizzle row.
static int make_map_f32(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int64_t n_bi, int64_t s_bi,
                        int64_t n_bo, int64_t s_bo, int box_inner, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NPT_REQUIRE(enc != nullptr, "gemm(tf32): cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)n_bi, (cuuint64_t)n_bo};
  const int64_t natural = ((rows * ld * 4 + 15) / 16) * 16;
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 4), (cuuint64_t)(n_bi > 1 ? s_bi * 4 : natural), (cuuint64_t)(n_bo > 1 ? s_bo * 4 : natural)};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NPT_REQUIRE(r == CUDA_SUCCESS, "gemm(tf32): cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld rows=%lld ld=%lld)", (int)r,
              (long long)inner, (long long)rows, (long long)ld);
  return 0;
}

// Can TMA address the operands?  (16-byte aligned bases, row pitches and batch strides.)  When it cannot —
// e.g. the (d, 65) embedding of the Shakespeare vocabulary: row pitch 260 bytes — npt_gemm runs the FFMA kernel.
bool gemm_tf32_supported(const npt_gemm_args& a) {
  return aligned16(a.A) && aligned16(a.B) && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.stride_a_outer % 4 == 0 &&
         a.stride_a_inner % 4 == 0 && a.stride_b_outer % 4 == 0 && a.stride_b_inner % 4 == 0 && !a.C_bf16 && !a.relu_bits_out &&
         !a.mask_bits && !a.C_bf16_peer && (int64_t)a.batch * 64 <= 65535 && cdiv(a.M, TF_BM) <= 65535;
}

static int tf32_bn(const npt_gemm_args& a) { return a.N <= 64 ? 64 : 128; }

// Split-K serves two purposes here: filling the machine when there are few output tiles (weight gradients), and
// bounding the length of one tensor-core accumulation chain.  tcgen05 adds every MMA into the fp32 TMEM
// accumulator with truncation, not round-to-nearest, so the error of a chain grows linearly with its length
// (measured: 1.3e-5 of max|C| after 1638 k, 3 MMAs per 8 k); the partials of the splits are added in fp32
// round-to-nearest by the fixed-order reduction.  Chains are capped at 32 k-blocks (1024 k).
constexpr int TF_MAX_CHAIN_KB = 32;
int gemm_tf32_splits(const npt_gemm_args& a) {
  const int64_t tiles = cdiv(a.M, TF_BM) * cdiv(a.N, tf32_bn(a)) * a.batch;
  const int64_t kblocks = cdiv(a.K, TF_BK);
  const int64_t target = (int64_t)sm_count();
  int64_t s = 1;
  if (tiles * 2 <= target && kblocks >= 32) {
    s = target / tiles;
    if (s > kblocks / 16) s = kblocks / 16;   // >= 16 k-blocks (512 k) per split
  }
  if (s < cdiv(kblocks, TF_MAX_CHAIN_KB)) s = cdiv(kblocks, TF_MAX_CHAIN_KB);
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  const int64_t kps = cdiv(kblocks, s);
  return (int)cdiv(kblocks, kps);            // no empty split
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const Tf32Params& p, const GemmEpilogue& epi, float* partial,
                       dim3 grid, cudaStream_t st) {
  using L = Tf32Smem<BN>;
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    NPT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  NPT_CHECK_CUDA(launch_pdl(kern, grid, dim3(TF_THREADS), L::TOTAL, st, tmA, tmB, p, epi, partial));
  NPT_LAUNCH_CHECK();
  return 0;
}

template <int BN>
static int launch_tf32_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const Tf32Params& p, const GemmEpilogue& epi, float* partial,
                          dim3 grid, bool a_mn, bool b_mn, cudaStream_t st) {
  if (a_mn && b_mn) return launch_tf32<BN, true, true>(tmA, tmB, p, epi, partial, grid, st);
  if (a_mn && !b_mn) return launch_tf32<BN, true, false>(tmA, tmB, p, epi, partial, grid, st);
  if (!a_mn && b_mn) return launch_tf32<BN, false, true>(tmA, tmB, p, epi, partial, grid, st);
  return launch_tf32<BN, false, false>(tmA, tmB, p, epi, partial, grid, st);
}

int gemm_tf32_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st) {
  const int bn = tf32_bn(a);
  const int64_t n_bi = a.batch_inner, n_bo = cdiv(a.batch, a.batch_inner);
  CUtensorMap tmA, tmB;
  int rc;
  if (a.trans_a) rc = make_map_f32(&tmA, a.A, a.M, a.K, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 32, 32);
  else           rc = make_map_f32(&tmA, a.A, a.K, a.M, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 32, TF_BM);
  if (rc) return rc;
  if (a.trans_b) rc = make_map_f32(&tmB, a.B, a.K, a.N, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 32, bn);
  else           rc = make_map_f32(&tmB, a.B, a.N, a.K, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 32, 32);
  if (rc) return rc;
  Tf32Params p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch; p.splits = splits;
  p.kb_per_split = (int)cdiv(cdiv(a.K, TF_BK), splits);
  p.raw = splits > 1;
  NPT_REQUIRE((int64_t)a.batch * splits <= 65535, "gemm(tf32): batch*splits %lld exceeds grid.z", (long long)a.batch * splits);
  dim3 grid((unsigned)cdiv(a.N, bn), (unsigned)cdiv(a.M, TF_BM), (unsigned)(a.batch * splits));
  const bool a_mn = a.trans_a != 0, b_mn = a.trans_b == 0;
  if (bn == 64) return launch_tf32_bn<64>(tmA, tmB, p, epi, partial, grid, a_mn, b_mn, st);
  return launch_tf32_bn<128>(tmA, tmB, p, epi, partial, grid, a_mn, b_mn, st);
}

}  // namespace npt
