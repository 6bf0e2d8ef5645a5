"""GEMM-level helpers shared by Linear / MLP / Attention / OutputEmbedding.

Each helper is one `npt_gemm` call (plus a bf16 cast of an operand when no shadow exists yet).
fp32 mode -> the FFMA kernel on fp32 operands; bf16 mode -> the tcgen05 kernel on bf16 operands
with fp32 accumulation.  In bf16 mode the residual stream, everything LayerNorm / softmax /
loss / Adam touch and all weight gradients stay fp32; tensors that only ever feed
another GEMM (`out="bf16"`) are stored as bf16 alone, which halves their HBM/L2 traffic.
"""
from __future__ import annotations

import torch

from .. import ops
from .. import runtime as rt

_BF16 = torch.bfloat16


def _act(x: torch.Tensor):
    """(operand tensor, leading dimension) of an activation for the current precision."""
    if rt.is_bf16():
        if x.dtype == _BF16:
            return x, x.shape[-1]
        sh = getattr(x, "_npt_bf16", None)
        if sh is None:
            sh = ops.cast_bf16(x, pad_cols=True)
            rt.attach_bf16(x, sh)
        return sh, sh.shape[-1]
    return x, x.shape[-1]


def _weight(layer, name: str):
    p = getattr(layer, name).data
    if rt.is_bf16():
        sh = layer.shadow(name)
        return sh, sh.shape[-1]
    return p, p.stride(0)


def _outputs(shape, out: str):
    """(fp32 buffer or None, bf16 buffer or None) for out in {"f32", "bf16", "both"}."""
    n = shape[-1]
    if not rt.is_bf16() or n % 8 != 0:
        return rt.empty(shape), None
    y = rt.empty(shape) if out in ("f32", "both") else None
    yb = rt.empty(shape, _BF16) if out in ("bf16", "both") else None
    return y, yb


def _result(y, yb):
    if y is None:
        return yb
    if yb is not None:
        rt.attach_bf16(y, yb)
    return y


def relu_bits_supported(d_out: int) -> bool:
    """Packed ReLU-mask words are produced by the tensor-core epilogue (bf16 mode, TMA-storable N)."""
    return rt.is_bf16() and d_out % 8 == 0


def linear_forward(x, layer, wname="weight", bias=None, relu=False, out="f32", relu_bits=None, out_tensor=None,
                   peer_out: int = 0):
    """y = x @ W (+ b) (ReLU) — nn/linear.py:44-47.  x (..., d_in), W (d_in, d_out).
    `relu_bits`: optional int32 (ceil(d_out/32), T) buffer that receives the packed (y > 0) mask."""
    W = getattr(layer, wname).data
    d_in, d_out = W.shape
    T = x.numel() // d_in
    A, lda = _act(x)
    B, ldb = _weight(layer, wname)
    y, yb = _outputs((*x.shape[:-1], d_out), out) if out_tensor is None else (None, out_tensor)  # bf16 peer slot
    ops.gemm(A, B, T, d_out, d_in, lda=lda, ldb=ldb, C_out=y, ldc=d_out, Cb_out=yb, ldcb=d_out, bias=bias, relu=relu,
             relu_bits_out=relu_bits, Cb_peer=peer_out)
    return _result(y, yb)


def linear_backward_weight(x, dy, may_defer: bool = False):
    """dW = x^T @ dy over all tokens — einsum("bsd,bsh->dh"), nn/linear.py:55.
    `may_defer` (and rt.deferred_wgrad active): a split-K launch leaves its partials for Adam to add, and
    the returned gradient is that (splits, d_in, d_out) tensor tagged `_npt_splits`."""
    d_in, d_out = x.shape[-1], dy.shape[-1]
    T = x.numel() // d_in
    A, lda = _act(x)
    B, ldb = _act(dy)
    dW = rt.empty((d_in, d_out))
    partials = ops.gemm(A, B, d_in, d_out, T, lda=lda, ldb=ldb, trans_a=True, C_out=dW, ldc=d_out,
                        skip_reduce=may_defer and rt.defer_wgrad())
    return dW if partials is None else partials


def linear_backward_input(dy, layer, wname="weight", mask=None, out="f32", out_tensor=None, peer_out: int = 0):
    """dx = dy @ W^T (optionally * (mask > 0): the fused ReLU.backward) — nn/linear.py:60."""
    W = getattr(layer, wname).data
    d_in, d_out = W.shape
    T = dy.numel() // d_out
    A, lda = _act(dy)
    B, ldb = _weight(layer, wname)
    dx, dxb = _outputs((*dy.shape[:-1], d_in), out) if out_tensor is None else (None, out_tensor)  # bf16 peer slot
    mk, ldm, bits = None, 0, None
    if mask is not None:
        if mask.dtype == torch.int32:
            bits = mask  # packed (hidden > 0) words written by the forward GEMM
        else:
            mk, ldm = mask, mask.shape[-1]
    ops.gemm(A, B, T, d_in, d_out, lda=lda, ldb=ldb, trans_b=True, C_out=dx, ldc=d_in, Cb_out=dxb, ldcb=d_in,
             mask=mk, ldmask=ldm, mask_bits=bits, Cb_peer=peer_out)
    return _result(dx, dxb)


def bias_grad(dy):
    """db = dy.sum(axis=(0,1)) — nn/linear.py:57-58 (reads fp32 or bf16).  When dy came out of
    LayerNorm.backward the sums were already computed there (ops.layernorm_bwd, want_colsum)."""
    fused = getattr(dy, "_npt_colsum", None)
    if fused is not None:
        return fused
    n = dy.shape[-1]
    return ops.colsum(dy.numel() // n, n, dy)
