"""MLP: column-parallel Linear -> ReLU -> row-parallel Linear -> TP all-reduce.
Drop-in for the reference's nn/mlp.py; the ReLU lives in the GEMM epilogues."""
from __future__ import annotations

import torch

from .. import distributed as npdist
from ..distributed import peer
from .. import runtime as rt
from . import functional as F
from .activation import ReLU
from .linear import Linear
from .model import Model


class MLP(Model):
    def __init__(self, d_in: int, d_hidden: int, d_out: int, **kwargs):
        super().__init__()
        self.add_settings({"d_in": d_in, "d_hidden": d_hidden, "d_out": d_out})
        self.add_layer("column_linear", Linear(d_in=d_in, d_out=d_hidden, weight_shard_axis=1, bias_shard_axis=0, **kwargs))
        self.add_layer("row_linear", Linear(d_in=d_hidden, d_out=d_out, weight_shard_axis=0, **kwargs))
        self.add_layer("relu", ReLU())
        # Only TP rank 0 adds the row-linear bias (mlp.py:35); the parameter exists everywhere.
        self.row_linear.use_bias = npdist.tensor_parallel_rank() == 0

    def forward(self, inputs, out: str = "f32"):
        # `out`: dtype of the returned tensor in bf16 mode ("bf16" when it only feeds a LayerNorm).
        # max(0, x@W1+b1) in one GEMM epilogue; the post-ReLU tensor doubles as the backward mask
        # because relu(a) > 0  <=>  a > 0.  It only feeds GEMMs, so bf16 mode keeps it in bf16 alone.
        d_hidden = self.column_linear.weight.data.shape[1]
        bits = None
        if F.relu_bits_supported(d_hidden):
            # 1 bit per element (hidden > 0), packed by the forward epilogue: the backward reads these
            # 3 MB instead of the 50 MB bf16 activation
            bits = rt.empty(((d_hidden + 31) // 32, inputs.numel() // inputs.shape[-1]), torch.int32)  # word-major
        hidden = self.column_linear.forward(inputs, relu=True, out="bf16", relu_bits=bits)
        self.ctx["relu_mask"] = bits if bits is not None else hidden
        if self.is_scattered and npdist.tensor_parallel_size() > 1:
            px = peer.exchange(inputs.shape) if out == "peer" else None
            if px is not None:
                # no all-reduce: the partial goes to a peer-visible slot and the LayerNorm that follows adds
                # the other rank's partial while it reads (distributed/peer.py)
                slot, send_ptr, recv_ptr = px.slot("fwd", (*inputs.shape[:-1], self.row_linear.weight.data.shape[1]))
                n_rows = hidden.numel() // hidden.shape[-1]
                rows = px.rows(n_rows)   # row-sharded LayerNorm (distributed/peer.py)
                outputs = self.row_linear.forward(hidden, out="bf16", out_tensor=slot,
                                                  peer_out=(send_ptr, *px.peer_rows(n_rows)) if rows else send_ptr)
                outputs._npt_peer, outputs._npt_rows = (px, recv_ptr), rows
                return outputs
            outputs = self.row_linear.forward(hidden, out="bf16" if out == "peer" else out)
            npdist.all_reduce(outputs, group=npdist.tensor_parallel_group())
            return outputs
        return self.row_linear.forward(hidden, out="bf16" if out == "peer" else out)

    def backward(self, d_out, out: str = "f32"):
        d_out = self.row_linear.backward(d_out, mask=self.ctx.pop("relu_mask"), out="bf16")  # (dy@W2^T) * (a>0)
        if out == "peer" and not (self.is_scattered and npdist.tensor_parallel_size() > 1):
            out = "bf16"
        if self.is_scattered and npdist.tensor_parallel_size() > 1:
            px = peer.exchange(d_out.shape) if out == "peer" else None
            if px is not None:
                slot, send_ptr, recv_ptr = px.slot("bwd", (*d_out.shape[:-1], self.column_linear.weight.data.shape[0]))
                n_rows = d_out.numel() // d_out.shape[-1]
                rows = px.rows(n_rows)
                d_in = self.column_linear.backward(d_out, out="bf16", out_tensor=slot,
                                                   peer_out=(send_ptr, *px.peer_rows(n_rows)) if rows else send_ptr)
                d_in._npt_peer, d_in._npt_rows = (px, recv_ptr), rows
                return d_in
            if out == "peer":
                out = "bf16"
            # all-reduce of dX (mlp.py:52-53) overlapped with the column-linear dW / db kernels
            pending = []
            d_out = self.column_linear.backward(
                d_out, out=out, on_dx=lambda dx: pending.append(npdist.all_reduce_async(dx, group=npdist.tensor_parallel_group())))
            pending[0].wait()
            return d_out
        return self.column_linear.backward(d_out, out=out)
