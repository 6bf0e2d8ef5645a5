"""Thin Python wrappers: torch buffers in, one C-ABI call each, torch buffers out.

Nothing here computes: every function marshals pointers/sizes into `libnumpitron_b200.so`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import runtime as rt
from ._lib import AdamTensor, GemmArgs, check

_F32, _BF16, _U16 = torch.float32, torch.bfloat16, torch.uint16


def _lib_():
    rt.require_gpu("kernel call")
    return _lib.load()


def _p(t, off=0):
    return rt.ptr(t, off)


# ---- per-launch timing (bench.py's roofline legs, benchmarks/microbench.py) -------------------------------
PROFILE = None   # set to a list: every op below then appends dict(name, bound, work, events) per C-ABI call


def _nbytes(*tensors) -> int:
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


class _timed:
    """CUDA events around one C-ABI call, on the launching stream.  `work` = algorithmic bytes (bound "hbm")
    or FLOPs (bound "tensor") of the call."""

    __slots__ = ("name", "bound", "work", "info", "e0")

    def __init__(self, name: str, bound: str, work: float, **info):
        self.name, self.bound, self.work, self.info = name, bound, work, info

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append(dict(name=self.name, bound=self.bound, work=float(self.work), events=(self.e0, e1), **self.info))
        return False


def set_sm_margin(sms: int) -> None:
    """Leave `sms` SMs out of the persistent GEMM grids (room for an in-flight NCCL collective)."""
    check(_lib_().npt_set_sm_margin(int(sms)), "npt_set_sm_margin")


# ---- GEMM ----------------------------------------------------------------------------------------
def gemm(A, B, M, N, K, *, lda, ldb, trans_a=False, trans_b=False, a_off=0, b_off=0,
         a_strides=(0, 0), b_strides=(0, 0), C_out=None, ldc=0, c_off=0, c_strides=(0, 0),
         Cb_out=None, ldcb=0, cb_off=0, cb_strides=(0, 0), batch=1, batch_inner=1, alpha=1.0,
         bias=None, relu=False, mask=None, ldmask=0, mask_off=0, mask_strides=(0, 0), mode=None,
         relu_bits_out=None, mask_bits=None, Cb_peer=0, skip_reduce: bool = False):
    """C[b] = alpha * op(A[b]) op(B[b]) (+bias)(ReLU)(*mask>0) — see npt_gemm in the header.

    strides are (outer, inner) in elements.  A/B must be fp32 for mode 0 and bf16 for mode 1.
    `skip_reduce`: if the launch is split along K, the fp32 partials stay in a fresh (splits, M, N) tensor,
    C is not written, and that tensor is returned (tagged `_npt_splits`) for a consumer that adds them
    itself (Adam); otherwise None is returned.
    """
    lib = _lib_()
    mode = rt.gemm_mode() if mode is None else mode
    want = _BF16 if mode == 1 else _F32
    if A.dtype != want or B.dtype != want:
        raise TypeError(f"gemm mode {mode} needs {want} operands, got {A.dtype}/{B.dtype}")
    g = GemmArgs()
    g.mode, g.trans_a, g.trans_b = mode, int(trans_a), int(trans_b)
    g.M, g.N, g.K = M, N, K
    g.batch, g.batch_inner = batch, batch_inner
    g.A, g.lda, g.stride_a_outer, g.stride_a_inner = _p(A, a_off), lda, a_strides[0], a_strides[1]
    g.B, g.ldb, g.stride_b_outer, g.stride_b_inner = _p(B, b_off), ldb, b_strides[0], b_strides[1]
    if C_out is not None:
        assert C_out.dtype == _F32
        g.C, g.ldc, g.stride_c_outer, g.stride_c_inner = _p(C_out, c_off), ldc, c_strides[0], c_strides[1]
    if Cb_out is not None:
        assert Cb_out.dtype == _BF16
        g.C_bf16, g.ldcb, g.stride_cb_outer, g.stride_cb_inner = _p(Cb_out, cb_off), ldcb, cb_strides[0], cb_strides[1]
    if Cb_peer:
        assert Cb_out is not None and C_out is None
        if isinstance(Cb_peer, tuple):   # (address, first row, end row): rows in that range go to the peer ONLY
            Cb_peer, g.peer_rows_lo, g.peer_rows_hi = Cb_peer
        g.C_bf16_peer = Cb_peer + 2 * cb_off
    g.alpha = alpha
    g.bias = _p(bias)
    g.relu = int(relu)
    if mask is not None:
        g.mask, g.mask_is_bf16 = _p(mask, mask_off), int(mask.dtype == _BF16)
        g.ldmask, g.stride_mask_outer, g.stride_mask_inner = ldmask, mask_strides[0], mask_strides[1]
    if relu_bits_out is not None or mask_bits is not None:
        bits = relu_bits_out if relu_bits_out is not None else mask_bits
        assert bits.dtype == torch.int32 and bits.shape[0] * 32 >= N and bits.shape[-1] >= M  # word-major [N/32][M]
        g.relu_bits_out, g.mask_bits, g.ld_bits = _p(relu_bits_out), _p(mask_bits), bits.shape[-1]
    need = lib.npt_gemm_workspace_bytes(C.byref(g))
    partials = None
    if need:
        splits = int(lib.npt_gemm_splits(C.byref(g))) if (skip_reduce and batch == 1) else 1
        if splits > 1:
            partials = rt.empty((splits, M, N))       # owned by the caller from here on (not the shared scratch)
            partials._npt_splits = splits
            assert partials.numel() * 4 >= need
            g.workspace, g.workspace_bytes, g.skip_splitk_reduce = partials.data_ptr(), partials.numel() * 4, 1
        else:
            ws = rt.workspace(need)
            g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    with _timed("gemm", "tensor", 2.0 * M * N * K * batch, M=M, N=N, K=K, batch=batch,
                trans=(int(trans_a), int(trans_b)), mode=mode, splitk=bool(need), args=g):
        check(lib.npt_gemm(C.byref(g), rt.stream()), "npt_gemm")
    return partials


def cast_bf16(t: torch.Tensor, pad_cols: bool = False) -> torch.Tensor:
    """bf16 copy of a contiguous fp32 tensor; with pad_cols the last dim is padded to a multiple of 8."""
    lib = _lib_()
    assert t.dtype == _F32 and t.is_contiguous()
    cols = t.shape[-1]
    rows = t.numel() // cols
    ld = rt.pad8(cols) if pad_cols else cols
    out = rt.empty((*t.shape[:-1], ld), _BF16)
    check(lib.npt_cast_f32_bf16(_p(t), cols, _p(out), ld, rows, cols, rt.stream()), "npt_cast_f32_bf16")
    return out


# ---- LayerNorm -----------------------------------------------------------------------------------
def layernorm_bf16_input_supported(d: int) -> bool:
    return bool(_lib_().npt_layernorm_bf16_input_supported(int(d)))


def layernorm_fwd(x, gamma, beta, eps, residual=None, want_bf16=False, x_peer: int = 0):
    """x: fp32, or bf16 (a sub-layer output that came out of its GEMM epilogue as bf16 alone).
    `x_peer`: device pointer of the other tensor-parallel rank's partial of x (peer memory); the kernel
    then normalises round_bf16(x + x_peer) and the reduced input is returned as a 4th value."""
    lib = _lib_()
    d = x.shape[-1]
    rows = x.numel() // d
    y = rt.empty(x.shape)
    yb = rt.empty(x.shape, _BF16) if want_bf16 else None
    mean, var = rt.empty((rows,)), rt.empty((rows,))
    x_sum = rt.empty(x.shape, _BF16) if x_peer else None
    with _timed("layernorm_fwd", "hbm", _nbytes(x, residual, y, yb, mean, var, x_sum) + (_nbytes(x) if x_peer else 0)):
        check(lib.npt_layernorm_fwd_v2(_p(x), int(x.dtype == _BF16), x_peer or None, _p(x_sum), _p(gamma), _p(beta),
                                       _p(residual), _p(y), _p(yb), _p(mean), _p(var), rows, d, eps, rt.stream()),
              "npt_layernorm_fwd_v2")
    if yb is not None:
        rt.attach_bf16(y, yb)
    if x_peer:
        return y, mean, var, x_sum
    return y, mean, var


def layernorm_bwd(x, gamma, mean, var, dy, eps, want_bf16=False, bf16_only=False, want_colsum=False, dy_peer: int = 0):
    """Returns (dx, dgamma, dbeta); with bf16_only dx is a bf16 tensor (it only feeds GEMMs).
    `want_colsum`: the column sums of dx (the bias gradient of the Linear in front of this LayerNorm) are
    computed in the same pass and attached to the returned dx as `_npt_colsum` (nn.functional.bias_grad
    picks them up)."""
    lib = _lib_()
    d = x.shape[-1]
    rows = x.numel() // d
    dx = None if bf16_only else rt.empty(x.shape)
    dxb = rt.empty(x.shape, _BF16) if (want_bf16 or bf16_only) else None
    want_cs = bool(want_colsum and lib.npt_layernorm_bwd_colsum_supported(d))
    need = lib.npt_layernorm_bwd_workspace_bytes(rows, d)
    # x (the saved LayerNorm input) and dy are read as one dtype: both bf16 when either is
    in_bf16 = x.dtype == _BF16 or dy.dtype == _BF16
    if in_bf16 and x.dtype != _BF16:
        x = cast_bf16(x)
    if in_bf16 and dy.dtype != _BF16:
        dy = cast_bf16(dy)
    assert not dy_peer or in_bf16
    # Inside GraphedTrainStep nothing but Adam reads dgamma / dbeta / the bias gradient: the per-block partials
    # go to it un-reduced (one launch less per LayerNorm; same additions in the same order).
    defer = rt.defer_wgrad()
    if defer:
        ws = rt.empty((need,), torch.uint8)   # lives until the optimizer step
        dgamma = dbeta = colsum_out = None
    else:
        ws = rt.workspace(need)
        dgamma, dbeta = rt.empty((d,)), rt.empty((d,))
        colsum_out = rt.empty((d,)) if want_cs else None
    n_part = C.c_int32(0)
    with _timed("layernorm_bwd", "hbm", _nbytes(x, dy, dx, dxb, mean, var) + (_nbytes(dy) if dy_peer else 0)):
        check(lib.npt_layernorm_bwd_v3(_p(x), _p(dy), int(in_bf16), dy_peer or None, _p(gamma), _p(mean), _p(var), _p(dx), _p(dxb),
                                       _p(dgamma), _p(dbeta), (_p(ws) if want_cs else None) if defer else _p(colsum_out),  # non-NULL = wanted
                                       rows, d, eps, ws.data_ptr(), ws.numel(),
                                       int(defer), C.addressof(n_part) if defer else None, rt.stream()), "npt_layernorm_bwd_v3")
    if defer:
        nsl, n = (3 if want_cs else 2), int(n_part.value)
        part = ws.view(torch.float32)

        def slice_(k):   # partial[:, k, :] as a stack of n split partials, pitch nsl * d
            t = part.as_strided((n, d), (nsl * d, 1), k * d)
            t._npt_splits, t._npt_split_stride, t._npt_keep = n, nsl * d, ws
            return t

        dgamma, dbeta = slice_(0), slice_(1)
        colsum_out = slice_(2) if want_cs else None
    out = dxb if bf16_only else dx
    if not bf16_only and dxb is not None:
        rt.attach_bf16(dx, dxb)
    if colsum_out is not None:
        out._npt_colsum = colsum_out
    return out, dgamma, dbeta


def layernorm_fwd_rows(x, x_recv: int, gamma, beta, eps, residual, rows, px):
    """Row-sharded LayerNorm forward under tp = 2 (distributed/peer.py): `x` holds this rank's partial and
    `x_recv` (local address) the other rank's, both valid for the rows this rank owns, rows = (first, count).
    Only those rows are normalised; the bf16 result is written into a peer-visible buffer AND into the other
    rank's copy of it, so after the caller's barrier both ranks hold every row of it.  Returns
    (y fp32 [own rows valid] with the full bf16 tensor attached, mean, var, reduced input [own rows valid])."""
    lib = _lib_()
    r0, n = rows
    d = x.shape[-1]
    T = x.numel() // d
    assert x.dtype == _BF16 and x.is_contiguous()
    y = rt.empty(x.shape)
    yb, yb_peer = px.arena(x.shape, _BF16)
    mean, var = rt.empty((T,)), rt.empty((T,))
    x_sum = rt.empty(x.shape, _BF16)
    e2, e4 = r0 * d * 2, r0 * d * 4
    with _timed("layernorm_fwd", "hbm", n * d * (2 + 2 + 2 + 4 + 2 + 2 + (4 if residual is not None else 0)) + 8 * n):
        check(lib.npt_layernorm_fwd_v3(x.data_ptr() + e2, 1, x_recv + e2, x_sum.data_ptr() + e2, _p(gamma), _p(beta),
                                       (residual.data_ptr() + e4) if residual is not None else None, y.data_ptr() + e4,
                                       yb.data_ptr() + e2, yb_peer + e2, mean.data_ptr() + 4 * r0, var.data_ptr() + 4 * r0,
                                       n, d, eps, rt.stream()), "npt_layernorm_fwd_v3")
    rt.attach_bf16(y, yb)
    return y, mean, var, x_sum


def layernorm_bwd_rows(x_sum, gamma, mean, var, dy, dy_recv: int, eps, rows, px, want_colsum: bool):
    """Row-sharded LayerNorm backward (see layernorm_fwd_rows): dx (bf16) of the own rows goes to a peer-visible
    buffer and to the other rank's copy; the per-block dgamma / dbeta / column-sum partials go into this rank's
    half of a [2][n][slices][d] stack on both ranks.  Returns (dx [complete after the caller's barrier],
    finish) — `finish()` must be called AFTER that barrier: it returns (dgamma, dbeta, colsum or None), either
    reduced over both ranks' partials or, inside GraphedTrainStep, as the un-reduced stacks for Adam."""
    lib = _lib_()
    r0, n = rows
    d = x_sum.shape[-1]
    assert dy.dtype == _BF16 and x_sum.dtype == _BF16
    dxb, dxb_peer = px.arena(x_sum.shape, _BF16)
    nsl = 3 if want_colsum else 2
    nb = int(lib.npt_layernorm_bwd_partials(n, d, 1))
    stacks, stacks_peer = px.arena((2, nb, nsl, d), _F32)
    mine = px.rank * nb * nsl * d * 4
    e2 = r0 * d * 2
    with _timed("layernorm_bwd", "hbm", n * d * (2 + 2 + 2 + 2 + 2) + 8 * n):
        check(lib.npt_layernorm_bwd_v4(x_sum.data_ptr() + e2, dy.data_ptr() + e2, 1, dy_recv + e2, _p(gamma), mean.data_ptr() + 4 * r0,
                                       var.data_ptr() + 4 * r0, None, dxb.data_ptr() + e2, dxb_peer + e2, None, None,
                                       stacks.data_ptr() if want_colsum else None,   # non-NULL = column sums wanted
                                       n, d, eps, stacks.data_ptr() + mine, nb * nsl * d * 4, stacks_peer + mine, 1, None,
                                       rt.stream()), "npt_layernorm_bwd_v4")

    def finish():
        if rt.defer_wgrad():
            def slice_(k):
                t = stacks.view(-1).as_strided((2 * nb, d), (nsl * d, 1), k * d)
                t._npt_splits, t._npt_split_stride, t._npt_keep = 2 * nb, nsl * d, stacks
                return t
            return slice_(0), slice_(1), (slice_(2) if want_colsum else None)
        dgamma, dbeta = rt.empty((d,)), rt.empty((d,))
        cs = rt.empty((d,)) if want_colsum else None
        check(lib.npt_layernorm_bwd_reduce(stacks.data_ptr(), 2 * nb, nsl, d, _p(dgamma), _p(dbeta), _p(cs), rt.stream()),
              "npt_layernorm_bwd_reduce")
        return dgamma, dbeta, cs

    return dxb, finish


# ---- ReLU / column sum ---------------------------------------------------------------------------
def relu_fwd(x, want_bf16=False):
    lib = _lib_()
    y = rt.empty(x.shape)
    yb = rt.empty(x.shape, _BF16) if want_bf16 else None
    check(lib.npt_relu_fwd(_p(x), _p(y), _p(yb), x.numel(), rt.stream()), "npt_relu_fwd")
    if yb is not None:
        rt.attach_bf16(y, yb)
    return y


def relu_bwd(x, dy, want_bf16=False):
    lib = _lib_()
    dx = rt.empty(x.shape)
    dxb = rt.empty(x.shape, _BF16) if want_bf16 else None
    check(lib.npt_relu_bwd(_p(x), _p(dy), _p(dx), _p(dxb), x.numel(), rt.stream()), "npt_relu_bwd")
    if dxb is not None:
        rt.attach_bf16(dx, dxb)
    return dx


def colsum(x2d_rows: int, cols: int, x: torch.Tensor, ld: int | None = None) -> torch.Tensor:
    lib = _lib_()
    ld = cols if ld is None else ld
    out = rt.empty((cols,))
    need = lib.npt_colsum_workspace_bytes(x2d_rows, cols)
    ws = rt.workspace(need)
    with _timed("colsum", "hbm", x2d_rows * cols * x.element_size()):
        check(lib.npt_colsum(_p(x), int(x.dtype == _BF16), ld, _p(out), x2d_rows, cols, ws.data_ptr(), ws.numel(),
                             rt.stream()), "npt_colsum")
    return out


# ---- causal softmax ----------------------------------------------------------------------------------
def softmax_causal_fwd(scores, divisor, want_bf16=False, bf16_only=False):
    lib = _lib_()
    S = scores.shape[-1]
    n_mats = scores.numel() // (S * S)
    p = None if bf16_only else rt.empty(scores.shape)
    pb = rt.empty(scores.shape, _BF16) if (want_bf16 or bf16_only) else None
    with _timed("softmax_causal_fwd", "hbm", _nbytes(scores) // 2 + (_nbytes(p, pb))):
        check(lib.npt_softmax_causal_fwd(_p(scores), _p(p), _p(pb), n_mats, S, divisor, rt.stream()), "npt_softmax_causal_fwd")
    if bf16_only:
        return pb
    if pb is not None:
        rt.attach_bf16(p, pb)
    return p


def softmax_causal_bwd(p, dp, divisor, want_bf16=False, bf16_only=False):
    """p may be the fp32 probabilities or their bf16 copy."""
    lib = _lib_()
    S = p.shape[-1]
    n_mats = p.numel() // (S * S)
    ds = None if bf16_only else rt.empty(p.shape)
    dsb = rt.empty(p.shape, _BF16) if (want_bf16 or bf16_only) else None
    with _timed("softmax_causal_bwd", "hbm", (_nbytes(p, dp)) // 2 + _nbytes(ds, dsb)):
        check(lib.npt_softmax_causal_bwd(_p(p), int(p.dtype == _BF16), _p(dp), _p(ds), _p(dsb), n_mats, S, divisor,
                                         rt.stream()), "npt_softmax_causal_bwd")
    if bf16_only:
        return dsb
    if dsb is not None:
        rt.attach_bf16(ds, dsb)
    return ds


# ---- fused attention (bf16 mode) ------------------------------------------------------------------------
def attention_supported(S: int, Dh: int) -> bool:
    return bool(_lib.load().npt_attention_supported(S, Dh))


def attention_fwd(qkv_b, B, S, h, Dh, divisor):
    """qkv_b (B,S,h*3*Dh) bf16 -> (out (B,S,h*Dh) bf16, lse (B,h,S) fp32)."""
    lib = _lib_()
    assert qkv_b.dtype == _BF16 and qkv_b.is_contiguous()
    out = rt.empty((B, S, h * Dh), _BF16)
    lse = rt.empty((B, h, S))
    with _timed("attention_fwd", "tensor", 2.0 * B * h * S * S * Dh):
        check(lib.npt_attention_fwd(_p(qkv_b), _p(out), _p(lse), B, S, h, Dh, divisor, rt.stream()), "npt_attention_fwd")
    return out, lse


def attention_bwd(qkv_b, out_b, d_out_b, lse, B, S, h, Dh, divisor):
    """-> dqkv (B,S,h*3*Dh) bf16."""
    lib = _lib_()
    assert d_out_b.dtype == _BF16 and d_out_b.is_contiguous()
    dqkv = rt.empty((B, S, h * 3 * Dh), _BF16)
    need = lib.npt_attention_bwd_workspace_bytes(B, S, h)
    ws = rt.workspace(need)
    with _timed("attention_bwd", "tensor", 5.0 * B * h * S * S * Dh):
        check(lib.npt_attention_bwd(_p(qkv_b), _p(out_b), _p(d_out_b), _p(lse), _p(dqkv), B, S, h, Dh, divisor,
                                    ws.data_ptr(), ws.numel(), rt.stream()), "npt_attention_bwd")
    return dqkv


# ---- embeddings ----------------------------------------------------------------------------------------
def embedding_gather(tokens, E, pos, S, chunk_start, chunk_end, masked, want_bf16=False):
    lib = _lib_()
    assert tokens.dtype == _U16 and E.dtype == _F32
    d, ldE = E.shape[0], E.stride(0)
    T = tokens.numel()
    out = rt.empty((*tokens.shape, d))
    ob = rt.empty((*tokens.shape, d), _BF16) if want_bf16 else None
    with _timed("embedding_gather", "hbm", _nbytes(tokens, out, ob) + T * d * 4):
        check(lib.npt_embedding_gather(_p(tokens), _p(E), ldE, _p(pos), _p(out), _p(ob), T, S, d, chunk_start, chunk_end,
                                       int(masked), rt.stream()), "npt_embedding_gather")
    if ob is not None:
        rt.attach_bf16(out, ob)
    return out


def add_pos(x, pos, S, want_bf16=False):
    """pos[:S] + x into a new buffer (x is left untouched, as in the reference)."""
    lib = _lib_()
    d = x.shape[-1]
    T = x.numel() // d
    out = rt.empty(x.shape)
    ob = rt.empty(x.shape, _BF16) if want_bf16 else None
    with _timed("add_pos", "hbm", _nbytes(x, out, ob)):
        check(lib.npt_add_pos(_p(x), _p(pos), _p(out), _p(ob), T, S, d, rt.stream()), "npt_add_pos")
    if ob is not None:
        rt.attach_bf16(out, ob)
    return out


def embedding_scatter_add(tokens, d_out, grad, chunk_start, chunk_end, masked):
    """In place on `grad` (d, V_local)."""
    lib = _lib_()
    assert tokens.dtype == _U16
    d, V_local, ldG = grad.shape[0], grad.shape[1], grad.stride(0)
    T = tokens.numel()
    need = lib.npt_embedding_scatter_add_workspace_bytes(T, V_local)
    ws = rt.workspace(need)
    with _timed("embedding_scatter_add", "hbm", _nbytes(tokens, d_out) + 2 * _nbytes(grad)):
        check(lib.npt_embedding_scatter_add(_p(tokens), _p(d_out), _p(grad), ldG, T, d, V_local, chunk_start, chunk_end,
                                            int(masked), ws.data_ptr(), ws.numel(), rt.stream()), "npt_embedding_scatter_add")
    return grad


# ---- loss ------------------------------------------------------------------------------------------------
def softmax_ce_fused(logits, labels, chunk_start, chunk_end, want_bf16=False):
    lib = _lib_()
    V = logits.shape[-1]
    T = logits.numel() // V
    loss = rt.empty(labels.shape)
    grad = rt.empty(logits.shape)
    ldgb = rt.pad8(V)
    gb = rt.empty((*logits.shape[:-1], ldgb), _BF16) if want_bf16 else None
    with _timed("softmax_ce_fused", "hbm", _nbytes(logits, labels, loss, grad, gb)):
        check(lib.npt_softmax_ce_fused(_p(logits), V, _p(labels), _p(loss), _p(grad), V, _p(gb), ldgb, T, V, chunk_start,
                                       chunk_end, rt.stream()), "npt_softmax_ce_fused")
    if gb is not None:
        rt.attach_bf16(grad, gb)
    return loss, grad


def softmax_ce_stage1(logits):
    lib = _lib_()
    V = logits.shape[-1]
    T = logits.numel() // V
    rowmax = rt.empty((T,))
    with _timed("softmax_ce_stage1", "hbm", _nbytes(logits, rowmax)):
        check(lib.npt_softmax_ce_stage1(_p(logits), V, _p(rowmax), T, V, rt.stream()), "npt_softmax_ce_stage1")
    return rowmax


def softmax_ce_stage2(logits, labels, rowmax, chunk_start, chunk_end):
    lib = _lib_()
    V = logits.shape[-1]
    T = logits.numel() // V
    picked, sumexp = rt.empty((T,)), rt.empty((T,))
    with _timed("softmax_ce_stage2", "hbm", _nbytes(logits, labels, rowmax, picked, sumexp)):
        check(lib.npt_softmax_ce_stage2(_p(logits), V, _p(labels), _p(rowmax), _p(picked), _p(sumexp), T, V, chunk_start,
                                        chunk_end, rt.stream()), "npt_softmax_ce_stage2")
    return picked, sumexp


def softmax_ce_stage3(logits, labels, rowmax, picked, sumexp, chunk_start, chunk_end, want_bf16=False):
    lib = _lib_()
    V = logits.shape[-1]
    T = logits.numel() // V
    loss = rt.empty(labels.shape)
    grad = rt.empty(logits.shape)
    ldgb = rt.pad8(V)
    gb = rt.empty((*logits.shape[:-1], ldgb), _BF16) if want_bf16 else None
    with _timed("softmax_ce_stage3", "hbm", _nbytes(logits, labels, rowmax, picked, sumexp, loss, grad, gb)):
        check(lib.npt_softmax_ce_stage3(_p(logits), V, _p(labels), _p(rowmax), _p(picked), _p(sumexp), _p(loss), _p(grad), V,
                                        _p(gb), ldgb, T, V, chunk_start, chunk_end, rt.stream()), "npt_softmax_ce_stage3")
    if gb is not None:
        rt.attach_bf16(grad, gb)
    return loss, grad


def mean(x):
    lib = _lib_()
    out = rt.empty((1,))
    check(lib.npt_mean(_p(x), _p(out), x.numel(), rt.stream()), "npt_mean")
    return out


# ---- pinned staging ---------------------------------------------------------------------------------
class PinnedBuffer:
    """Page-locked host bytes owned by the library (not by torch's caching host allocator, whose
    event bookkeeping breaks when the copies are issued under CUDA-graph capture)."""

    def __init__(self, nbytes: int):
        import numpy as np

        lib = _lib_()
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib.npt_pinned_alloc(C.byref(p), self.nbytes), "npt_pinned_alloc")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def view(self, dtype, shape):
        return self.array.view(dtype).reshape(shape)

    def to_device(self, dst: torch.Tensor, nbytes: int | None = None):
        check(_lib_().npt_memcpy_async(dst.data_ptr(), self.ptr, self.nbytes if nbytes is None else nbytes, 1,
                                       rt.stream()), "npt_memcpy_async")

    def from_device(self, src: torch.Tensor, nbytes: int | None = None):
        check(_lib_().npt_memcpy_async(self.ptr, src.data_ptr(), self.nbytes if nbytes is None else nbytes, 2,
                                       rt.stream()), "npt_memcpy_async")

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                _lib.load().npt_pinned_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


# ---- Adam ------------------------------------------------------------------------------------------------
class AdamTable:
    """Host (pinned) + device copies of the npt_adam_tensor table for one optimizer."""

    def __init__(self, capacity: int):
        self.capacity = max(int(capacity), 1)
        self.host = (AdamTensor * self.capacity)()
        self.pinned = PinnedBuffer(C.sizeof(self.host))
        self.dev = torch.empty(C.sizeof(self.host), dtype=torch.uint8, device=rt.device())
        self.n = 0
        self.key = None
        self.keep = None
        self.nbytes = 0
        self.upload_done = None   # event recorded after the last upload, on the stream that issued it

    def fill(self, entries):
        """entries: list of (p, g, m, v, shadow_or_None).  Re-uploads only when a pointer changed."""
        key = tuple((_p(p), _p(g), _p(m), _p(v), _p(sh), getattr(g, "_npt_splits", 1), getattr(g, "_npt_split_stride", 0), p.numel())
                    for p, g, m, v, sh in entries)
        if key == self.key:
            return
        assert len(entries) <= self.capacity
        if self.upload_done is not None and not torch.cuda.is_current_stream_capturing():
            self.upload_done.synchronize()  # the earlier upload (whatever stream issued it) has read `pinned`
        C.memset(self.host, 0, C.sizeof(self.host))
        for i, (p, g, m, v, sh) in enumerate(entries):
            t = self.host[i]
            t.p, t.g, t.m, t.v = _p(p), _p(g), _p(m), _p(v)
            t.n = p.numel()
            t.g_splits = getattr(g, "_npt_splits", 1)   # split-K partials left by a weight-gradient GEMM
            t.g_split_stride = getattr(g, "_npt_split_stride", p.numel())   # LayerNorm block partials: pitch (2 or 3) * d
            if sh is not None:
                t.p_bf16 = _p(sh)
                t.shadow_cols = p.shape[-1]
                t.shadow_ld = sh.shape[-1]
        C.memmove(self.pinned.ptr, self.host, C.sizeof(self.host))
        self.pinned.to_device(self.dev)  # async memcpy node: safe under graph capture
        if not torch.cuda.is_current_stream_capturing():
            self.upload_done = torch.cuda.Event()
            self.upload_done.record()
        self.n, self.key, self.keep = len(entries), key, entries
        # algorithmic bytes of one step over this table: read p, g (every split-K partial), m, v; write p, m, v
        # (+ the bf16 shadow) = 28 B / parameter for a plain gradient
        self.nbytes = sum(p.numel() * (24 + 4 * getattr(g, "_npt_splits", 1) + (2 if sh is not None else 0))
                          for p, g, m, v, sh in entries)


def adam_step(table: AdamTable, lr, b1, omb1, b2, omb2, c1, c2, eps):
    lib = _lib_()
    with _timed("adam_step", "hbm", table.nbytes):
        check(lib.npt_adam_step(table.dev.data_ptr(), table.host, table.n, lr, b1, omb1, b2, omb2, c1, c2, eps, rt.stream()),
              "npt_adam_step")
