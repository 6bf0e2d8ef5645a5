"""Causal multi-head attention — drop-in for the reference's nn/attention.py.

Faithful to the reference's arithmetic: the QKV projection's columns are laid out
per head as [q_h | k_h | v_h]; the score scale is 1/sqrt(d_model) (FULL model width); the local
head count is n_heads // tp_size with no divisibility check.  Heads are never transposed or
copied: every product is a strided-batched GEMM that addresses the (B,S,h,3,Dh) projection
output in place and writes straight into the merged (B,S,h*Dh) / (B,S,h,3,Dh) layouts.
In bf16 mode q/k/v, P, dS and the merged head outputs only feed tensor-core GEMMs and are kept
in bf16 alone; scores and dP (softmax inputs) stay fp32.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import distributed as npdist
from ..distributed import peer
from .. import ops
from .. import runtime as rt
from .activation import Softmax
from .linear import Linear
from .model import Model

_BF16 = torch.bfloat16


class Attention(Model):
    def __init__(self, d_model: int, n_heads: int, d_hidden: int, **kwargs):
        super().__init__()
        self.add_layer("qkv_projection", Linear(d_model, n_heads * d_hidden * 3, weight_shard_axis=1,
                                                use_bias=False, **kwargs))
        self.add_layer("out_projection", Linear(n_heads * d_hidden, d_model, weight_shard_axis=0,
                                                use_bias=False, **kwargs))
        self.add_layer("softmax", Softmax(axis=-1))
        self.add_settings({"d_model": d_model, "n_heads": n_heads, "d_hidden": d_hidden})

    def _local_heads(self) -> int:
        return self.n_heads if not self.is_scattered else self.n_heads // npdist.tensor_parallel_size()

    @staticmethod
    def _operand(t):
        """The tensor a GEMM reads: itself (fp32 mode / already bf16) or its bf16 copy."""
        if not rt.is_bf16() or t.dtype == _BF16:
            return t
        sh = getattr(t, "_npt_bf16", None)
        if sh is None:
            sh = ops.cast_bf16(t, pad_cols=True)
            rt.attach_bf16(t, sh)
        return sh

    def forward(self, inputs, out: str = "f32"):
        """`out`: dtype of the returned tensor in bf16 mode ("bf16" when it only feeds a LayerNorm)."""
        batch_size, seq_len, d_model = inputs.shape
        h = self._local_heads()
        bf16 = rt.is_bf16()
        qkv = self.qkv_projection.forward(inputs, out="bf16")              # (B,S,h*3*Dh)
        ld = qkv.shape[-1]
        dh3 = ld // h
        dh = dh3 // 3
        if bf16:
            assert seq_len % 8 == 0 and dh % 8 == 0, "bf16 mode needs seq_len and head_dim to be multiples of 8"
        divisor = float(np.float32(d_model**0.5))                           # attention.py:48
        if bf16 and ops.attention_supported(seq_len, dh):
            # fused tcgen05 attention: scores / probabilities never leave the SM
            att_b, lse = ops.attention_fwd(qkv, batch_size, seq_len, h, dh, divisor)
            self.ctx |= {"qkv": qkv, "att": att_b, "lse": lse, "divisor": divisor,
                         "shape": (batch_size, seq_len, h, dh)}
            return self._project_out(att_b, inputs.shape, out)
        nb = batch_size * h
        q_op = self._operand(qkv)
        head = (seq_len * ld, dh3)                                          # (batch, head) strides in qkv
        mat = (h * seq_len * seq_len, seq_len * seq_len)                    # ... in (B,h,S,S)
        merged = (seq_len * h * dh, dh)                                     # ... in (B,S,h*Dh)

        scores = rt.empty((batch_size, h, seq_len, seq_len))               # q @ k^T
        ops.gemm(q_op, q_op, seq_len, seq_len, dh, lda=ld, ldb=ld, trans_b=True, a_off=0, b_off=dh,
                 a_strides=head, b_strides=head, C_out=scores, ldc=seq_len, c_strides=mat,
                 batch=nb, batch_inner=h)
        attention_weights = self.softmax.forward(scores, divisor=divisor)   # scale, mask, softmax
        p_op = self._operand(attention_weights)

        att = None if bf16 else rt.empty((batch_size, seq_len, h * dh))     # p @ v, heads merged in place
        att_b = rt.empty((batch_size, seq_len, h * dh), _BF16) if bf16 else None
        ops.gemm(p_op, q_op, seq_len, dh, seq_len, lda=seq_len, ldb=ld, b_off=2 * dh, a_strides=mat,
                 b_strides=head, C_out=att, ldc=h * dh, c_strides=merged,
                 Cb_out=att_b, ldcb=h * dh, cb_strides=merged, batch=nb, batch_inner=h)
        outputs = self.out_projection.forward(att_b if bf16 else att, out="bf16" if out == "peer" else out)

        self.ctx |= {"qkv": qkv, "attention_weights": attention_weights, "shape": (batch_size, seq_len, h, dh)}
        if self.is_scattered and npdist.tensor_parallel_size() > 1:
            npdist.all_reduce(outputs, group=npdist.tensor_parallel_group())
        return outputs

    def _project_out(self, att, out_shape, out: str):
        """Output projection (row-parallel) + the tensor-parallel reduction of its result
        (attention.py:57-58): an NCCL all-reduce, or — out == "peer" — no reduction here at all: the partial
        goes to a peer-visible slot and the LayerNorm that follows adds the other rank's (distributed/peer.py)."""
        tp = self.is_scattered and npdist.tensor_parallel_size() > 1
        px = peer.exchange(out_shape) if (tp and out == "peer") else None
        if px is not None:
            slot, send_ptr, recv_ptr = px.slot("fwd", out_shape)
            n_rows = att.numel() // att.shape[-1]
            rows = px.rows(n_rows)   # row-sharded LayerNorm: each tile is stored once, on the rank that owns its rows
            outputs = self.out_projection.forward(att, out="bf16", out_tensor=slot,
                                                  peer_out=(send_ptr, *px.peer_rows(n_rows)) if rows else send_ptr)
            outputs._npt_peer, outputs._npt_rows = (px, recv_ptr), rows
            return outputs
        outputs = self.out_projection.forward(att, out="bf16" if out == "peer" else out)
        if tp:
            npdist.all_reduce(outputs, group=npdist.tensor_parallel_group())
        return outputs

    def backward(self, d_out, out: str = "f32"):
        batch_size, seq_len, h, dh = self.ctx.pop("shape")
        qkv = self.ctx.pop("qkv")
        if "lse" in self.ctx:  # fused path
            att_b, lse, divisor = self.ctx.pop("att"), self.ctx.pop("lse"), self.ctx.pop("divisor")
            d_att = self.out_projection.backward(d_out, out="bf16")
            d_qkv = ops.attention_bwd(qkv, att_b, d_att, lse, batch_size, seq_len, h, dh, divisor)
            return self._qkv_backward(d_qkv, out)
        p = self.ctx.pop("attention_weights")
        bf16 = rt.is_bf16()
        ld = h * 3 * dh
        nb = batch_size * h
        head = (seq_len * ld, 3 * dh)
        mat = (h * seq_len * seq_len, seq_len * seq_len)
        merged = (seq_len * h * dh, dh)

        d_att = self.out_projection.backward(d_out, out="bf16")             # (B,S,h*Dh)
        qkv_op, p_op, datt_op = self._operand(qkv), self._operand(p), self._operand(d_att)
        d_qkv = None if bf16 else rt.empty((batch_size, seq_len, ld))
        d_qkv_b = rt.empty((batch_size, seq_len, ld), _BF16) if bf16 else None

        def into_qkv(col):
            return dict(C_out=d_qkv, ldc=ld, c_off=col, c_strides=head,
                        Cb_out=d_qkv_b, ldcb=ld, cb_off=col, cb_strides=head)

        # dV = P^T @ dAtt                                                    (attention.py:81-83)
        ops.gemm(p_op, datt_op, seq_len, dh, seq_len, lda=seq_len, ldb=h * dh, trans_a=True, a_strides=mat,
                 b_strides=merged, batch=nb, batch_inner=h, **into_qkv(2 * dh))
        # dP = dAtt @ V^T                                                    (attention.py:84)
        d_p = rt.empty((batch_size, h, seq_len, seq_len))
        ops.gemm(datt_op, qkv_op, seq_len, seq_len, dh, lda=h * dh, ldb=ld, trans_b=True, b_off=2 * dh,
                 a_strides=merged, b_strides=head, C_out=d_p, ldc=seq_len, c_strides=mat, batch=nb, batch_inner=h)
        # dS = softmax backward, masked, / sqrt(d_model)                     (attention.py:86-87)
        d_s = self.softmax.backward(d_p)
        ds_op = self._operand(d_s)
        # dQ = dS @ K ; dK = dS^T @ Q                                        (attention.py:89-92)
        ops.gemm(ds_op, qkv_op, seq_len, dh, seq_len, lda=seq_len, ldb=ld, b_off=dh, a_strides=mat, b_strides=head,
                 batch=nb, batch_inner=h, **into_qkv(0))
        ops.gemm(ds_op, qkv_op, seq_len, dh, seq_len, lda=seq_len, ldb=ld, trans_a=True, b_off=0, a_strides=mat,
                 b_strides=head, batch=nb, batch_inner=h, **into_qkv(dh))
        return self._qkv_backward(d_qkv_b if bf16 else d_qkv, out)

    def _qkv_backward(self, d_qkv, out: str = "f32"):
        """QKV-projection backward; under TP the all-reduce of dX (attention.py:101-102) is started as
        soon as dX is enqueued and overlaps the dW GEMM."""
        tp = self.is_scattered and npdist.tensor_parallel_size() > 1
        if tp and out == "peer":
            d_model = self.qkv_projection.weight.data.shape[0]
            px = peer.exchange((*d_qkv.shape[:-1], d_model))
            if px is not None:
                slot, send_ptr, recv_ptr = px.slot("bwd", (*d_qkv.shape[:-1], d_model))
                n_rows = d_qkv.numel() // d_qkv.shape[-1]
                rows = px.rows(n_rows)
                d_in = self.qkv_projection.backward(d_qkv, out="bf16", out_tensor=slot,
                                                    peer_out=(send_ptr, *px.peer_rows(n_rows)) if rows else send_ptr)
                d_in._npt_peer, d_in._npt_rows = (px, recv_ptr), rows
                return d_in
        if out == "peer":
            out = "bf16"
        if tp:
            pending = []
            d_in = self.qkv_projection.backward(
                d_qkv, out=out, on_dx=lambda dx: pending.append(npdist.all_reduce_async(dx, group=npdist.tensor_parallel_group())))
            pending[0].wait()
            return d_in
        return self.qkv_projection.backward(d_qkv, out=out)
