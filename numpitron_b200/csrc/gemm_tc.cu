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
// Persistent: one CTA per SM (or one CTA pair per SM pair, tcgen05 cta_group::2) loops over 128 x BN
// (256 x BN) output tiles, n fastest, so the A panel of a row of tiles stays in L2.  Warp 0 = TMA
// producer, warp 1 = TMEM owner + MMA issuer (one elected lane), then the epilogue warps: 8 in the
// fast epilogue (two per TMEM lane quarter), 4 in the general one.  The accumulator is
// double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of
// tile i+1.  Split-K partials go to a workspace (also by TMA store) and are added in fixed order by
// gemm.cu's reduce kernel or directly by the consumer (Adam), deterministically either way.
// This file is the host side: tile / pair / split-K choice, tensor maps, dispatch.  The kernel is
// in gemm_tc_kernel.cuh, instantiated per tile width in gemm_tc_bn*.cu.
#include "gemm_tc_kernel.cuh"

namespace npt {

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

// Split-K factor for launches with too few tiles to fill the GPU (the weight-gradient GEMMs: a few
// hundred rows and columns, K = all the tokens).  Picks the split count with the shortest modelled
// makespan: waves x (k-blocks per unit + the unit's fixed cost) + the reduce pass over the partials.
// What matters most is the wave count: 18 tiles x 9 splits = 162 units is TWO waves on 148 SMs,
// 18 x 8 = 144 is one (the first version rounded the split count up and ran two waves on three of
// the four weight gradients of a transformer block).
int gemm_tc_splits(const npt_gemm_args& a) {
  if (a.relu_bits_out || a.mask_bits) return 1;  // the packed-mask epilogue runs in the GEMM kernel itself
  const int bn = tc_bn(a);
  const int64_t tiles = cdiv(a.M, TC_BM) * cdiv(a.N, bn) * a.batch;
  const int64_t kblocks = cdiv(a.K, TC_BK);
  const int64_t target = (int64_t)sm_count();
  if (tiles * 2 > target || kblocks < 8) return 1;
  const double kblock_us = 0.0014 * (128 + bn);                       // shared-memory fill at ~64 B/clk
  const double unit_us = 1.5;                                          // accumulator drain + partial store
  const double reduce_us = (double)a.M * a.N * a.batch * 4.0 / 3.0e6;  // per split: partials re-read at ~3 TB/s
  int64_t best = 1;
  double best_cost = 0.0;
  for (int64_t s = 1; s <= 64 && s <= kblocks / 4; ++s) {               // >= 4 k-blocks (256 k) per split
    const int64_t kps = cdiv(kblocks, s);
    const int64_t s_eff = cdiv(kblocks, kps);                          // no empty split
    if (s_eff != s) continue;
    const int64_t waves = cdiv(tiles * s, target);
    const double cost = (double)waves * ((double)kps * kblock_us + unit_us) + (s > 1 ? 3.0 + (double)s * reduce_us : 0.0);
    if (s == 1 || cost < best_cost) { best = s; best_cost = cost; }
  }
  return (int)best;
}

// Can this launch use the fast epilogue (gemm_tc_kernel.cuh)?  One output kind, written by TMA, no
// tensor mask, bias readable 16 bytes at a time.
static bool tc_fast_ok(const npt_gemm_args& a) {
  if (a.mask || (a.C != nullptr) == (a.C_bf16 != nullptr)) return false;
  if (a.bias && !aligned16(a.bias)) return false;
  if (a.C) return aligned16(a.C) && a.ldc % 4 == 0 && a.stride_c_outer % 4 == 0 && a.stride_c_inner % 4 == 0;
  return aligned16(a.C_bf16) && a.ldcb % 8 == 0 && a.stride_cb_outer % 8 == 0 && a.stride_cb_inner % 8 == 0;
}

// Tile width and CTA grouping of a launch without split-K.  Model: the launch takes
// waves x (time of one k-block), where a k-block costs the larger of its MMA time (~2.3 cycles per
// tile column) and its shared-memory fill time at ~64 B/clk (128 rows of A + BN / cl rows of B, 128
// bytes each); a CTA pair (cl = 2, tcgen05 cta_group::2) halves the B rows each SM has to pull in.
// NUMPITRON_GEMM_PAIR=0 turns the pairs off (A/B comparison in the microbenchmarks).
struct TcConfig { int bn, cl; };
static TcConfig tc_config(const npt_gemm_args& a) {
  TcConfig best = {tc_bn(a), 1};
  const int64_t m_tiles = cdiv(a.M, TC_BM);
  const int sms = sm_count();
  // measured (profiles/README.md): pairs win from N ~ 1024 up (16384x1536x384: 24.3 vs 26.8 us, 8192^3: 756 vs
  // 925 us) and lose on narrow outputs, where a launch is a couple of tiles per CTA and fixed costs dominate
  static int min_n = -1;
  if (min_n < 0) { const char* e = getenv("NUMPITRON_GEMM_PAIR_MINN"); min_n = e ? atoi(e) : 1024; }
  if (a.N < min_n || m_tiles % 2 != 0 || !tc_fast_ok(a)) return best;
  if (a.relu_bits_out) return best;  // bias + ReLU + bit packing: the pair's epilogue becomes the limit (27.8 vs 25.7 us)
  if (m_tiles * cdiv(a.N, best.bn) * a.batch * 2 <= sms) return best;  // split-K territory: single CTAs
  const char* env = getenv("NUMPITRON_GEMM_PAIR");
  if (env && env[0] == '0') return best;
  auto cost = [&](int bn, int cl) {
    const int64_t units = m_tiles / cl * cdiv(a.N, bn) * a.batch;
    const int64_t waves = cdiv(units, (int64_t)(sms / cl));
    const double mma = 2.3 * bn, fill = 2.0 * (128 + bn / cl);
    return (double)waves * (mma > fill ? mma : fill);
  };
  double best_cost = cost(best.bn, 1);
  const int cand[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    if (cand[i] == 192 && !a.trans_b) continue;  // half of a 192-wide MN-major B tile is not a whole 64-column box
    const double c = cost(cand[i], 2);
    if (c < best_cost * 0.97) { best_cost = c; best.bn = cand[i]; best.cl = 2; }
  }
  return best;
}

#ifdef NPT_GEMM_DEBUG
static long long* g_trace_buf = nullptr;
extern "C" int npt_debug_gemm_trace(long long* host_out) {  // 256 x 64 stamps of the last traced launch
  if (!g_trace_buf) return -1;
  return (int)cudaMemcpy(host_out, g_trace_buf, 256 * 64 * sizeof(long long), cudaMemcpyDeviceToHost);
}
#endif

#define NPT_TC_EXTERN(BNV)                                                                                  \
  extern template int launch_tc_bn<BNV>(const TcParams&, const GemmEpilogue&, float*, const CUtensorMap&,      \
                                        const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, bool, \
                                        bool, int, int, cudaStream_t);
NPT_TC_EXTERN(64) NPT_TC_EXTERN(128) NPT_TC_EXTERN(192) NPT_TC_EXTERN(256)
#undef NPT_TC_EXTERN

int gemm_tc_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st) {
  TcConfig cfg = {tc_bn(a), 1};
  if (splits == 1) cfg = tc_config(a);
  const int bn = cfg.bn, cl = cfg.cl;
  const int64_t n_bi = a.batch_inner, n_bo = cdiv(a.batch, a.batch_inner);
  CUtensorMap tmA, tmB, tmC, tmCb, tmP;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmCb, 0, sizeof(tmCb));
  memset(&tmP, 0, sizeof(tmP));
  int rc;
  if (a.trans_a) rc = make_map_bf16(&tmA, a.A, a.M, a.K, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 64, 64);
  else           rc = make_map_bf16(&tmA, a.A, a.K, a.M, a.lda, n_bi, a.stride_a_inner, n_bo, a.stride_a_outer, 64, TC_BM);
  if (rc) return rc;
  // B: one box of bn / cl rows (K-major), or 64 x 64 boxes (MN-major); a CTA of a pair loads half of the tile
  if (a.trans_b) rc = make_map_bf16(&tmB, a.B, a.K, a.N, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 64, bn / cl);
  else           rc = make_map_bf16(&tmB, a.B, a.N, a.K, a.ldb, n_bi, a.stride_b_inner, n_bo, a.stride_b_outer, 64, 64);
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
  p.peer_store = 0;
  p.peer_lo = p.peer_hi = 0;
#ifdef NPT_GEMM_DEBUG
  {
    const char* d = getenv("NUMPITRON_GEMM_DBG");
    p.dbg = d ? atoi(d) : 0;
    static long long* trace_buf = nullptr;
    if (!trace_buf) cudaMalloc(&trace_buf, 256 * 64 * sizeof(long long));
    g_trace_buf = trace_buf;
    p.trace = getenv("NUMPITRON_GEMM_TRACE") ? trace_buf : nullptr;
    if (p.trace) cudaMemsetAsync(trace_buf, 0, 256 * 64 * sizeof(long long), st);
  }
#endif
  NPT_REQUIRE(!(p.raw && (a.relu_bits_out || a.mask_bits)), "gemm(tc): packed ReLU bits are not supported with split-K");
  // Can the outputs be written by TMA?  (16-byte aligned base, pitches and batch strides)
  int kind = EPI_GENERAL;
  if (p.raw) {
    p.tma_store = (a.N % 4 == 0) && aligned16(partial);
    if (p.tma_store) kind = EPI_F32;
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
    if (tc_fast_ok(a)) {
      const int flags = (a.C ? EPI_F32 : EPI_BF16) | (a.alpha != 1.f ? EPI_ALPHA : 0) | (a.bias ? EPI_BIAS : 0) |
                        (a.relu ? EPI_RELU : 0) | (a.mask_bits ? EPI_BITS_IN : 0) | (a.relu_bits_out ? EPI_BITS_OUT : 0);
      kind = tc_epi_kind(flags);
    }
    if (p.tma_store && a.C_bf16) {
      if (kind == EPI_GENERAL)  // 32-column boxes, 64-byte rows
        rc = make_map(&tmCb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_64B, a.C_bf16, a.N, a.M, a.ldcb, n_bi,
                      a.stride_cb_inner, n_bo, a.stride_cb_outer, 32, 32);
      else                      // 64-column boxes, 128-byte rows
        rc = make_map(&tmCb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, a.C_bf16, a.N, a.M, a.ldcb, n_bi,
                      a.stride_cb_inner, n_bo, a.stride_cb_outer, 64, 32);
      if (rc) return rc;
    }
  }
  if (a.C_bf16_peer) {  // second (peer-GPU) destination of the bf16 output: fast epilogue only
    NPT_REQUIRE(!p.raw && kind != EPI_GENERAL && a.C_bf16 && !a.C && aligned16(a.C_bf16_peer),
                "gemm(tc): C_bf16_peer needs a single bf16 output that TMA can store, no split-K, 16-byte aligned buffers");
    rc = make_map(&tmP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, CU_TENSOR_MAP_SWIZZLE_128B, a.C_bf16_peer, a.N, a.M, a.ldcb, n_bi,
                  a.stride_cb_inner, n_bo, a.stride_cb_outer, 64, 32);
    if (rc) return rc;
    p.peer_store = 1;
    if (a.peer_rows_hi > a.peer_rows_lo) {
      NPT_REQUIRE(a.peer_rows_lo % TC_BM == 0 && a.peer_rows_hi % TC_BM == 0 && a.batch == 1,
                  "gemm(tc): peer_rows bounds must be multiples of %d (got %d, %d) and the launch unbatched", TC_BM, a.peer_rows_lo, a.peer_rows_hi);
      p.peer_store = 2;
      p.peer_lo = a.peer_rows_lo;
      p.peer_hi = a.peer_rows_hi;
    }
  }
  NPT_REQUIRE(cl == 1 || kind != EPI_GENERAL, "gemm(tc): internal error, CTA pairs need the fast epilogue");
  const bool a_mn = a.trans_a != 0, b_mn = a.trans_b == 0;
  if (bn == 64) return launch_tc_bn<64>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, a_mn, b_mn, cl, kind, st);
  if (bn == 192) return launch_tc_bn<192>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, a_mn, b_mn, cl, kind, st);
  if (bn == 256) return launch_tc_bn<256>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, a_mn, b_mn, cl, kind, st);
  return launch_tc_bn<128>(p, epi, partial, tmA, tmB, tmC, tmCb, tmP, a_mn, b_mn, cl, kind, st);
}

}  // namespace npt
