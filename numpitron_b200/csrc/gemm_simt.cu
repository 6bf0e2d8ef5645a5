// fp32 FFMA GEMM (the fp32-accurate mode): register-tiled, double-buffered through shared
// memory, any strides / transposes / batch, optional split-K.  Products are accumulated in
// fp32 in plain k order, which keeps it within ~1e-6 relative of NumPy's float32 einsum/matmul
// (nn/linear.py:44,55,60; nn/attention.py:48,52,81-92).
#include "gemm_common.cuh"

namespace npt {

template <int BM, int BN, int BK>
struct SimtTile {
  static constexpr int TM = BM / 16, TN = BN / 16;  // 16x16 threads
  static constexpr int THREADS = 256;
  static constexpr int A_PER_T = BM * BK / THREADS, B_PER_T = BN * BK / THREADS;
};

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int64_t lda, int64_t sao, int64_t sai,
                 const float* __restrict__ B, int64_t ldb, int64_t sbo, int64_t sbi, int trans_a,
                 int trans_b, int M, int N, int K, int batch, int splits, int ktiles_per_split,
                 GemmEpilogue epi, float* __restrict__ partial) {
  using T = SimtTile<BM, BN, BK>;
  constexpr int HM = BM / 2, HN = BN / 2;      // the tile is split in halves so that
  constexpr int QM = T::TM / 2, QN = T::TN / 2;  // 128-bit smem reads are conflict free
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bz = blockIdx.z;
  const int b = bz / splits, split = bz - b * splits;
  const int bo = b / epi.batch_inner, bi = b - bo * epi.batch_inner;
  const float* Ab = A + bo * sao + bi * sai;
  const float* Bb = B + bo * sbo + bi * sbi;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ktiles = (K + BK - 1) / BK;
  const int kt_begin = split * ktiles_per_split;
  int kt_end = kt_begin + ktiles_per_split;
  if (kt_end > ktiles) kt_end = ktiles;

  float acc[T::TM][T::TN];
#pragma unroll
  for (int i = 0; i < T::TM; ++i)
#pragma unroll
    for (int j = 0; j < T::TN; ++j) acc[i][j] = 0.f;

  float ra[T::A_PER_T], rb[T::B_PER_T];

  auto load_global = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < T::A_PER_T; ++i) {
      const int e = tid + i * T::THREADS;
      int m, k;
      if (trans_a) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = trans_a ? __ldg(Ab + (int64_t)gk * lda + gm) : __ldg(Ab + (int64_t)gm * lda + gk);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < T::B_PER_T; ++i) {
      const int e = tid + i * T::THREADS;
      int n, k;
      if (trans_b) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = trans_b ? __ldg(Bb + (int64_t)gn * ldb + gk) : __ldg(Bb + (int64_t)gk * ldb + gn);
      rb[i] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < T::A_PER_T; ++i) {
      const int e = tid + i * T::THREADS;
      int m, k;
      if (trans_a) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < T::B_PER_T; ++i) {
      const int e = tid + i * T::THREADS;
      int n, k;
      if (trans_b) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      Bs[buf][k][n] = rb[i];
    }
  };

  if (kt_begin < kt_end) {
    load_global(kt_begin);
    store_smem(0);
  }
  __syncthreads();
  int buf = 0;
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    const bool more = kt + 1 < kt_end;
    if (more) load_global(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[T::TM], bb[T::TN];
#pragma unroll
      for (int i = 0; i < QM; ++i) {
        a[i] = As[buf][k][ty * QM + i];
        a[QM + i] = As[buf][k][HM + ty * QM + i];
      }
#pragma unroll
      for (int j = 0; j < QN; ++j) {
        bb[j] = Bs[buf][k][tx * QN + j];
        bb[QN + j] = Bs[buf][k][HN + tx * QN + j];
      }
#pragma unroll
      for (int i = 0; i < T::TM; ++i)
#pragma unroll
        for (int j = 0; j < T::TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // epilogue: rows {ty*QM+i, HM+ty*QM+i}, cols {tx*QN+j, HN+tx*QN+j}
#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < QM; ++i) {
      const int m = m0 + ih * HM + ty * QM + i;
      if (m >= M) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int nb = n0 + jh * HN + tx * QN;
        if (splits > 1) {
          float* prow = partial + (((int64_t)split * batch + b) * M + m) * N;
#pragma unroll
          for (int j = 0; j < QN; ++j)
            if (nb + j < N) prow[nb + j] = acc[ih * QM + i][jh * QN + j];
        } else {
          if (QN == 4 && (epi.N % 4) == 0) {
            epi_store4(epi, b, m, nb, make_float4(acc[ih * QM + i][jh * QN + 0], acc[ih * QM + i][jh * QN + 1],
                                                  acc[ih * QM + i][jh * QN + 2], acc[ih * QM + i][jh * QN + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < QN; ++j) epi_store1(epi, b, m, nb + j, acc[ih * QM + i][jh * QN + j]);
          }
        }
      }
    }
}

static bool simt_big_tile(const npt_gemm_args& a) {
  // 128x128 tiles once there is enough work to fill the machine with them.
  const int64_t tiles = cdiv(a.M, 128) * cdiv(a.N, 128) * a.batch;
  return a.M >= 96 && a.N >= 96 && tiles >= sm_count();
}

int gemm_simt_splits(const npt_gemm_args& a) {
  const int bm = simt_big_tile(a) ? 128 : 64;
  const int64_t tiles = cdiv(a.M, bm) * cdiv(a.N, bm) * a.batch;
  const int64_t ktiles = cdiv(a.K, 16);
  const int64_t target = (int64_t)sm_count() * 2;
  if (tiles >= target || ktiles < 16) return 1;
  int64_t s = target / tiles;
  if (s > ktiles / 8) s = ktiles / 8;  // at least 8 k-tiles (128 k) per split
  if (s > 32) s = 32;
  return s < 1 ? 1 : (int)s;
}

int gemm_simt_launch(const npt_gemm_args& a, const GemmEpilogue& epi, int splits, float* partial, cudaStream_t st) {
  const int ktiles = (int)cdiv(a.K, 16);
  const int kps = (int)cdiv(ktiles, splits);
  NPT_REQUIRE((int64_t)a.batch * splits <= 65535, "gemm(simt): batch*splits %lld exceeds grid.z", (long long)a.batch * splits);
  const float* A = (const float*)a.A;
  const float* B = (const float*)a.B;
  if (simt_big_tile(a)) {
    dim3 grid((unsigned)cdiv(a.N, 128), (unsigned)cdiv(a.M, 128), (unsigned)(a.batch * splits));
    gemm_simt_kernel<128, 128, 16><<<grid, 256, 0, st>>>(A, a.lda, a.stride_a_outer, a.stride_a_inner, B, a.ldb,
                                                         a.stride_b_outer, a.stride_b_inner, a.trans_a, a.trans_b,
                                                         a.M, a.N, a.K, a.batch, splits, kps, epi, partial);
  } else {
    dim3 grid((unsigned)cdiv(a.N, 64), (unsigned)cdiv(a.M, 64), (unsigned)(a.batch * splits));
    gemm_simt_kernel<64, 64, 16><<<grid, 256, 0, st>>>(A, a.lda, a.stride_a_outer, a.stride_a_inner, B, a.ldb,
                                                       a.stride_b_outer, a.stride_b_inner, a.trans_a, a.trans_b,
                                                       a.M, a.N, a.K, a.batch, splits, kps, epi, partial);
  }
  NPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace npt
