/*
 * numpitron_b200 — C ABI of the B200 (sm_100a) kernels behind NuMPItron's layer API.
 *
 * The reference (lweitkamp/numpitron) is pure Python/NumPy and has no FFI of its own
 * (SURVEY.md §8b); the boundary these entry points sit behind is the body of each layer's
 * forward/backward/step.  Every declaration cites the reference expression(s) it replaces
 * (paths relative to /root/reference/src/numpitron/).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; npt_last_error() gives the text;
 *   - every data pointer is a DEVICE pointer owned by the caller; nothing is allocated or
 *     freed behind the caller's back (scratch is passed in as workspace pointer + size);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - matrices are row-major with explicit leading dimensions counted in ELEMENTS;
 *   - fp32 unless a name says bf16 (bf16 = __nv_bfloat16, 2 bytes); tokens/labels are uint16
 *     exactly as the reference's memmap (data/loader.py:32).
 */
#ifndef NUMPITRON_B200_H
#define NUMPITRON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------- */
const char* npt_last_error(void);
int npt_version(void);
/* Number of kernel launches this library has issued in this process (monotonic). */
unsigned long long npt_launch_count(void);
/* Number of SMs the persistent GEMM grids leave free from now on (0 = use all).  The tensor-parallel
 * backward raises it while an NCCL all-reduce (distributed/all_reduce.py:23) is in flight on another
 * stream, so the collective's CTAs can be scheduled next to the dW GEMM. */
int npt_set_sm_margin(int32_t sms);
/* Fills SM count and compute capability of the current device. */
int npt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Pinned host staging buffers and async copies for the step's inputs (uint16 tokens, as
 * data/loader.py:48-54 yields them) and its scalar result.  kind: 1 = H2D, 2 = D2H, 3 = D2D. */
int npt_pinned_alloc(void** out, size_t bytes);
int npt_pinned_free(void* p);
int npt_memcpy_async(void* dst, const void* src, size_t bytes, int32_t kind, void* stream);
/* Device-side barrier of two tensor-parallel ranks on `stream`: `peer_flag` = the flag word in the OTHER rank's
 * memory (peer pointer), `own_flag` = the same word in local memory, `counter` = a local 8-byte epoch counter
 * (all zero-initialised; both ranks must execute the same number of barriers).  Replaces the implicit
 * synchronisation of the reference's blocking Allreduce (distributed/all_reduce.py:23) where the exchange itself
 * is done by the kernels on either side (npt_gemm_args.C_bf16_peer, npt_layernorm_*_v3/_v4). */
int npt_peer_barrier(void* peer_flag, void* own_flag, void* counter, void* stream);
/* device-to-device copy of `rows` rows of `width_bytes` bytes between pitched buffers (re-assembling the vocabulary
 * shards of the tied embedding into one table: nn/embedding.py:19-37 shards it over the vocabulary axis) */
int npt_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                       void* stream);

/* ---- dense contractions ---------------------------------------------------------------
 * One strided-batched GEMM serves every contraction on the path:
 *   Linear.forward   einsum("bsd,dh->bsh") + bias            nn/linear.py:44-47
 *   Linear.backward  einsum("bsd,bsh->dh"), einsum("bsh,dh->bsd")   nn/linear.py:55,60
 *   Attention        q@k^T, p@v and their backward products  nn/attention.py:47-52, 81-92
 *   OutputEmbedding  x@E, einsum("bsd,bsv->dv"), d@E^T       nn/embedding.py:92,99,102
 *
 *   C[b] = alpha * op(A[b]) * op(B[b])   (+ bias[n])  (ReLU)  (* (mask[b] > 0))
 * op(A) is M x K, op(B) is K x N.  trans_a = 0: A stored (M,K); 1: stored (K,M).
 * trans_b = 0: B stored (K,N); 1: stored (N,K).
 * Batch index i in [0, batch) maps to (i / batch_inner, i % batch_inner) and offsets
 * X + outer*stride_X_outer + inner*stride_X_inner (elements) — this is how the per-head
 * [q_h | k_h | v_h] column layout of the QKV projection (attention.py:43-45) is addressed
 * without any transpose/copy.
 */
enum { NPT_GEMM_FP32_SIMT = 0,   /* fp32 FFMA, the fp32-accurate mode            */
       NPT_GEMM_BF16_TC   = 1,   /* bf16 operands, tcgen05.mma + TMEM, fp32 accum */
       NPT_GEMM_TF32X3_TC = 2 }; /* fp32 operands, 3xTF32 split on tcgen05       */

typedef struct npt_gemm_args {
  int32_t mode;            /* NPT_GEMM_* */
  int32_t trans_a, trans_b;
  int32_t M, N, K;
  int32_t batch, batch_inner;          /* batch >= 1, batch_inner >= 1 */
  const void* A; int64_t lda, stride_a_outer, stride_a_inner;  /* fp32 (mode 0,2) or bf16 (mode 1) */
  const void* B; int64_t ldb, stride_b_outer, stride_b_inner;
  float* C;        int64_t ldc,  stride_c_outer,  stride_c_inner;   /* fp32 out, may be NULL  */
  void*  C_bf16;   int64_t ldcb, stride_cb_outer, stride_cb_inner;  /* bf16 out, may be NULL  */
  float alpha;
  const float* bias;       /* [N] or NULL; added after alpha                        */
  int32_t relu;            /* 1: C = max(C, 0) after bias (MLP column linear, mlp.py:38-39) */
  const void* mask;        /* NULL or tensor shaped like C: C *= (mask > 0)  (ReLU.backward, activation.py:42-43) */
  int32_t mask_is_bf16;    /* dtype of mask */
  int64_t ldmask, stride_mask_outer, stride_mask_inner;
  void* workspace; size_t workspace_bytes;   /* split-K partials; see npt_gemm_workspace_bytes */
  /* Packed ReLU mask (bf16 tensor-core mode, batch == 1): bit (n % 32) of word [n / 32][m], i.e. the
   * words are stored word-major with row pitch ld_bits >= M (32 consecutive rows = one 128-byte line).
   * relu_bits_out: written by a relu=1 GEMM (bit = output > 0); mask_bits: consumed like `mask`
   * by the backward GEMM — 1/16 of the bytes of a bf16 mask and no extra pass in the epilogue. */
  uint32_t* relu_bits_out; const uint32_t* mask_bits; int64_t ld_bits;
  /* Optional second destination of the bf16 output (same ldcb / strides as C_bf16; bf16 tensor-core mode,
   * single-output launches): every tile is also TMA-stored there.  Under tensor parallelism this is a buffer
   * in the OTHER rank's memory (NVLink P2P), so the partial result of a row-parallel layer (nn/attention.py:56,
   * nn/mlp.py:34) lands next to the peer's own partial while the GEMM is still running, and the LayerNorm
   * that follows adds the two from local memory (npt_layernorm_fwd_v2, x_peer). */
  void* C_bf16_peer;
  /* peer_rows_hi > peer_rows_lo: SPLIT store instead of a double store — tiles whose rows lie in
   * [peer_rows_lo, peer_rows_hi) go to C_bf16_peer ONLY, all others to C_bf16 only (both bounds multiples of 128).
   * That is a reduce-scatter by push: each rank ends up with both partials of the rows it owns and normalises
   * only those (npt_layernorm_fwd_v3 / _bwd_v4), instead of every rank reducing and normalising all rows. */
  int32_t peer_rows_lo, peer_rows_hi;
  /* 1: when the launch is split along K (npt_gemm_splits() > 1) leave the fp32 partials
   * [split][batch][M][N] in `workspace` and do NOT reduce them into C — the consumer adds them in split
   * order (npt_adam_tensor.g_splits: the weight gradients of nn/linear.py:55 are only ever read by Adam,
   * so the reduce pass, one launch per weight, disappears).  No effect on launches that are not split. */
  int32_t skip_splitk_reduce;
} npt_gemm_args;

size_t npt_gemm_workspace_bytes(const npt_gemm_args* args);
int32_t npt_gemm_splits(const npt_gemm_args* args);   /* split-K factor npt_gemm will use for these arguments */
int npt_gemm(const npt_gemm_args* args, void* stream);

/* fp32 -> bf16 copy of a rows x cols matrix; columns [cols, ld_dst) of dst are zero-filled. */
int npt_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst,
                      int64_t rows, int64_t cols, void* stream);
/* The hi/lo TF32 split used by NPT_GEMM_TF32X3_TC is done inside the GEMM kernel. */

/* ---- LayerNorm — nn/normalization.py:22-52 ----------------------------------------------
 * fwd: y = gamma*(x-mean)/sqrt(var+eps)+beta (+ residual: TransformerBlock.forward's
 *      `norm(sub(x)) + x`, nn/transformer_block.py:20-21).  mean/var (population) are saved.
 * bwd: dgamma = sum dy*xhat, dbeta = sum dy, dx = (g*dy - c1*xhat - c2) / sqrt(var)  — the
 *      final division uses std WITHOUT eps exactly as normalization.py:45-50 does.
 *      partial: workspace of npt_layernorm_bwd_workspace_bytes(rows,d) bytes.  dx or dx_bf16 may be NULL.
 */
int npt_layernorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual,
                      float* y, void* y_bf16, float* mean, float* var,
                      int64_t rows, int32_t d, float eps, void* stream);
size_t npt_layernorm_bwd_workspace_bytes(int64_t rows, int32_t d);
int npt_layernorm_bwd(const float* x, const float* gamma, const float* mean, const float* var,
                      const float* dy, float* dx, void* dx_bf16, float* dgamma, float* dbeta,
                      int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes,
                      void* stream);
/* v2 entry points.  Additions over the calls above:
 *  - x (and, in the backward, dy) may be bf16 (x_is_bf16 / inputs_are_bf16 = 1): in the bf16 throughput
 *    mode the sub-layer output that LayerNorm normalises, and the dX that reaches its backward, come out
 *    of a GEMM epilogue as bf16 alone — half the bytes to store, to all-reduce across tensor-parallel
 *    ranks (nn/attention.py:57-58,101-102, nn/mlp.py:35-36,52-53) and to read here.  Needs
 *    npt_layernorm_bf16_input_supported(d) (d % 4 == 0, d <= 1024).
 *  - dx_colsum[d] (NULL: not wanted) = column sums of dx.  dx is the dy of the Linear layer in front of
 *    this LayerNorm (TransformerBlock.backward, nn/transformer_block.py:27-30), so this is that layer's
 *    bias gradient (nn/linear.py:57-58) without a second pass over dx.  Needs
 *    npt_layernorm_bwd_colsum_supported(d).
 *  - x_peer / dy_peer (NULL: none; bf16 inputs only): the OTHER tensor-parallel rank's partial of the same
 *    tensor, a peer-memory pointer (NVLink P2P).  The kernel normalises round_bf16(x + x_peer), i.e. what
 *    the reference's Allreduce (nn/attention.py:57-58, nn/mlp.py:35-36; backward: attention.py:101-102,
 *    mlp.py:52-53) would have produced, without an all-reduce kernel in between; x_sum_bf16 (optional)
 *    receives that reduced input for the backward pass.  tp = 2 only; the caller synchronises the two
 *    ranks (both partials complete and visible) before the launch. */
int npt_layernorm_bf16_input_supported(int32_t d);
int npt_layernorm_bwd_colsum_supported(int32_t d);
int npt_layernorm_fwd_v2(const void* x, int32_t x_is_bf16, const void* x_peer, void* x_sum_bf16,
                         const float* gamma, const float* beta, const float* residual, float* y,
                         void* y_bf16, float* mean, float* var, int64_t rows, int32_t d, float eps,
                         void* stream);
int npt_layernorm_bwd_v2(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer,
                         const float* gamma, const float* mean, const float* var, float* dx,
                         void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int64_t rows,
                         int32_t d, float eps, void* workspace, size_t workspace_bytes, void* stream);
/* v3 = v2 plus: `n_partials` (host pointer, may be NULL) receives the number of per-block partials the first
 * stage left in `workspace`, laid out [n_partials][2 or 3][d] (dgamma, dbeta, and dx_colsum when requested).
 * defer_reduce != 0: the second stage is skipped and dgamma / dbeta / dx_colsum are not written; the caller
 * keeps `workspace` alive and hands the partials to npt_adam_step (npt_adam_tensor.g_splits = n_partials,
 * g_split_stride = (2 or 3) * d), which adds them in block order — the additions of the second stage, so the
 * parameter update is bit-identical (these gradients, nn/normalization.py:38-45 and nn/linear.py:57-58, are
 * only ever read by the optimizer). */
int npt_layernorm_bwd_v3(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer,
                         const float* gamma, const float* mean, const float* var, float* dx,
                         void* dx_bf16, float* dgamma, float* dbeta, float* dx_colsum, int64_t rows,
                         int32_t d, float eps, void* workspace, size_t workspace_bytes,
                         int32_t defer_reduce, int32_t* n_partials, void* stream);

/* Row-sharded LayerNorm under tensor parallelism over 2 ranks (tp = 2, bf16 mode).  The GEMM in front stored its
 * partial split by rows (npt_gemm_args.peer_rows_*), so this rank holds BOTH partials of the rows it owns: it
 * normalises only those (the pointers passed are those of its first row, `rows` = its row count) and pushes its half
 * of every tensor the next kernels need in full into the other rank's memory — forward: y_bf16_peer; backward:
 * dx_bf16_peer and the per-block dgamma / dbeta / colsum partials (workspace_peer; each rank's stack of
 * npt_layernorm_bwd_partials() blocks sits behind the other's in one [2][n][2 or 3][d] buffer that both hold
 * after the caller's barrier).  fwd_v3 / bwd_v4 = fwd_v2 / bwd_v3 plus those destinations (NULL = none). */
int npt_layernorm_fwd_v3(const void* x, int32_t x_is_bf16, const void* x_peer, void* x_sum_bf16,
                         const float* gamma, const float* beta, const float* residual, float* y,
                         void* y_bf16, void* y_bf16_peer, float* mean, float* var, int64_t rows, int32_t d,
                         float eps, void* stream);
int npt_layernorm_bwd_v4(const void* x, const void* dy, int32_t inputs_are_bf16, const void* dy_peer,
                         const float* gamma, const float* mean, const float* var, float* dx,
                         void* dx_bf16, void* dx_bf16_peer, float* dgamma, float* dbeta, float* dx_colsum,
                         int64_t rows, int32_t d, float eps, void* workspace, size_t workspace_bytes,
                         void* workspace_peer, int32_t defer_reduce, int32_t* n_partials, void* stream);
/* number of per-block partials the first stage of npt_layernorm_bwd_* leaves for these arguments */
int32_t npt_layernorm_bwd_partials(int64_t rows, int32_t d, int32_t inputs_are_bf16);
/* second stage alone: dgamma / dbeta (/ dx_colsum, n_slices = 3) = block-order sums of partials[n_partials][n_slices][d] */
int npt_layernorm_bwd_reduce(const float* partials, int32_t n_partials, int32_t n_slices, int32_t d,
                             float* dgamma, float* dbeta, float* dx_colsum, void* stream);

/* ---- ReLU — nn/activation.py:38-43 (stand-alone; the GEMM epilogue fuses both) ---------- */
int npt_relu_fwd(const float* x, float* y, void* y_bf16, int64_t n, void* stream);
int npt_relu_bwd(const float* x, const float* dy, float* dx, void* dx_bf16, int64_t n, void* stream);

/* Column sum over rows: db = dy.sum(axis=(0,1)) — nn/linear.py:57-58.  Deterministic two-stage. */
size_t npt_colsum_workspace_bytes(int64_t rows, int32_t cols);
int npt_colsum(const void* x, int32_t x_is_bf16, int64_t ld, float* out, int64_t rows, int32_t cols,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- causal softmax over attention scores — nn/attention.py:48-51,87; nn/activation.py:6-31
 * fwd: P = softmax(where(j<=i, S/divisor, -inf)) row-wise over (rows_total = B*h*S) x S.
 * bwd: dS = where(j<=i, P*(dP - sum_j dP*P), 0) / divisor   (compact form of the Jacobian).
 * Either output (fp32 / bf16) may be NULL; the backward reads P as fp32 or as the bf16 copy.
 */
int npt_softmax_causal_fwd(const float* scores, float* p, void* p_bf16, int64_t n_mats, int32_t S,
                           float divisor, void* stream);
int npt_softmax_causal_bwd(const void* p, int32_t p_is_bf16, const float* dp, float* ds, void* ds_bf16,
                           int64_t n_mats, int32_t S, float divisor, void* stream);

/* ---- fused causal attention (bf16 mode; P never leaves the SM) --------------------------------------
 * Replaces nn/attention.py:43-52 (forward) and :81-98 (backward) for head_dim 64 or 96 (96 = the single local head of
 * GPT-2 small under tp = 8, attention.py:37-45) and S % 128 == 0; other shapes use npt_gemm + npt_softmax_causal_*.  qkv is the QKV projection output
 * (B,S,h,3,Dh) bf16 exactly as attention.py:43-45 lays it out ([q_h|k_h|v_h] per head); `out` /
 * `d_out` are the merged heads (B,S,h*Dh) bf16; lse is the per-row log-sum-exp (B,h,S) fp32 saved
 * for the backward; dqkv (B,S,h,3,Dh) bf16 = concat([dq,dk,dv],-1) merged back (attention.py:94-98).
 * divisor = sqrt(d_model) (attention.py:48).  Deterministic (no atomics).  The workspace of the backward holds
 * D = sum_d dO * O per row; for S = 256, head_dim 64 one kernel computes dQ, dK and dV (and D) in a single pass.
 */
int npt_attention_supported(int32_t S, int32_t Dh);
int npt_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int32_t B, int32_t S, int32_t h,
                      int32_t Dh, float divisor, void* stream);
size_t npt_attention_bwd_workspace_bytes(int32_t B, int32_t S, int32_t h);
int npt_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                      void* dqkv_bf16, int32_t B, int32_t S, int32_t h, int32_t Dh, float divisor,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- embeddings — nn/embedding.py:44-83,125-128 ---------------------------------------------
 * gather: out[t,:] = (tok in [chunk_start,chunk_end) ? E[:, tok-chunk_start] : 0) (+ pos[t % S,:]).
 *         E is (d, ldE>=V_local) row-major as the reference stores it.  `masked`=0 reproduces the
 *         un-scattered path (no range mask).  pos may be NULL (added after the TP all-reduce).
 * scatter_add: np.add.at(grad.T, tok[~mask], d_out[~mask]) in flattened (b,s) order, i.e. for
 *         every vocab column the contributions are added one by one in increasing t, starting
 *         from the value already in grad — bit-identical to NumPy.  No float atomics.
 */
int npt_embedding_gather(const uint16_t* tokens, const float* E, int64_t ldE, const float* pos,
                         float* out, void* out_bf16, int64_t T, int32_t S, int32_t d,
                         int32_t chunk_start, int32_t chunk_end, int32_t masked, void* stream);
/* out[t,:] = pos[t % S,:] + x[t,:]  (PositionalEncoding.forward, embedding.py:125-128); out may alias x. */
int npt_add_pos(const float* x, const float* pos, float* out, void* out_bf16, int64_t T, int32_t S, int32_t d, void* stream);
/* Data-loader windows (data/loader.py:49-54), SURVEY 8(f)2: the uint16 token file lives in HBM; for each of
 * the B start indices (drawn on the host from the loader's NumPy generator, so the stream is the
 * reference's) x[b,:] = tokens[start_b + 0..S-1], y[b,:] = tokens[start_b + 1..S].  start_b + S < n_tokens. */
int npt_window_gather(const uint16_t* tokens, int64_t n_tokens, const int64_t* starts, uint16_t* x, uint16_t* y,
                      int32_t B, int32_t S, void* stream);
size_t npt_embedding_scatter_add_workspace_bytes(int64_t T, int32_t V_local);
int npt_embedding_scatter_add(const uint16_t* tokens, const float* d_out, float* grad, int64_t ldG,
                              int64_t T, int32_t d, int32_t V_local, int32_t chunk_start,
                              int32_t chunk_end, int32_t masked, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- vocab-parallel softmax cross-entropy — nn/loss.py:6-50 -----------------------------------
 * Split at the three TP all-reduces (loss.py:23-26, 34-35, 41-42):
 *   stage1: rowmax[t] = max_v logits[t,v]                         -> all-reduce(max)
 *   stage2: picked[t] = in-chunk ? logits[t,label]-rowmax : 0 ;  sumexp[t] = sum_v exp(logits-rowmax)
 *                                                                  -> all-reduce(sum) x2
 *   stage3: loss[t] = -picked + log(sumexp);  grad = exp(logits-rowmax)/sumexp - onehot*(in-chunk)
 * `fused` does all three for tp == 1 in one pass.  The gradient is NOT divided by B*S (loss.py:47-48).
 */
int npt_softmax_ce_stage1(const float* logits, int64_t ld, float* rowmax, int64_t T, int32_t V_local, void* stream);
int npt_softmax_ce_stage2(const float* logits, int64_t ld, const uint16_t* labels, const float* rowmax,
                          float* picked, float* sumexp, int64_t T, int32_t V_local,
                          int32_t chunk_start, int32_t chunk_end, void* stream);
int npt_softmax_ce_stage3(const float* logits, int64_t ld, const uint16_t* labels, const float* rowmax,
                          const float* picked, const float* sumexp, float* loss, float* grad,
                          int64_t ldg, void* grad_bf16, int64_t ldgb, int64_t T, int32_t V_local,
                          int32_t chunk_start, int32_t chunk_end, void* stream);
int npt_softmax_ce_fused(const float* logits, int64_t ld, const uint16_t* labels, float* loss,
                         float* grad, int64_t ldg, void* grad_bf16, int64_t ldgb, int64_t T,
                         int32_t V_local, int32_t chunk_start, int32_t chunk_end, void* stream);
/* mean of n floats into out[0] (train_step's `loss.mean()`, train_shakespeare.py:21); deterministic. */
int npt_mean(const float* x, float* out, int64_t n, void* stream);

/* ---- Adam — optimizer/adam.py:31-45 ----------------------------------------------------------------
 * m = b1*m + omb1*g ; v = b2*v + omb2*g*g ; p -= (lr*(m/c1)) / (sqrt(v/c2) + eps)
 * with the six scalars rounded to float32 by the caller exactly as NumPy's scalar promotion
 * does (timestep never advances in the reference, so c1 = 1-beta1, c2 = 1-beta2).  Products and
 * sums are individually rounded (no FMA contraction) so the update is bit-identical to NumPy.
 * `tensors` is a device array of npt_adam_tensor; one launch updates all of them.
 */
typedef struct npt_adam_tensor {
  float* p; const float* g; float* m; float* v;
  void* p_bf16;            /* optional bf16 shadow of p refreshed in the same pass (may be NULL) */
  int64_t n;               /* elements */
  int64_t shadow_cols, shadow_ld;  /* if p_bf16: p is rows x shadow_cols, shadow row pitch shadow_ld */
  /* g_splits > 1: g holds split-K partials, gradient[i] = ((0 + g[i]) + g[stride + i]) + ... in split order
   * (the exact additions of the split-K reduce kernel, so the update is bit-identical to reducing first). */
  int64_t g_splits, g_split_stride;
} npt_adam_tensor;
int npt_adam_step(const npt_adam_tensor* tensors_dev, const npt_adam_tensor* tensors_host,
                  int32_t n_tensors, float lr, float b1, float omb1, float b2, float omb2,
                  float c1, float c2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NUMPITRON_B200_H */
