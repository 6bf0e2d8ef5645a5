"""Linear — drop-in for the reference's nn/linear.py (same constructor, forward, backward)."""
from __future__ import annotations

import os

from .. import ops
from . import functional as F
from .core import Layer, weight_init_fn

# SMs the dW GEMM leaves free while the tensor-parallel all-reduce of dX is in flight.  Measured on
# 2 x B200 (profiles/README.md): 0, 16 and 32 give the same step time, so the default is 0.
TP_OVERLAP_SM_MARGIN = int(os.environ.get("NUMPITRON_TP_SM_MARGIN", "0"))


class Linear(Layer):
    def __init__(self, d_in: int, d_out: int, use_bias: bool = True, weight_shard_axis: int | None = None,
                 bias_shard_axis: int | None = None, weight_init: str = "scaled_normal", bias_init: str = "zeros",
                 **kwargs):
        super().__init__()
        self.add_settings({"d_in": d_in, "d_out": d_out, "use_bias": use_bias,
                           "weight_shard_axis": weight_shard_axis, "bias_shard_axis": bias_shard_axis,
                           "weight_init": weight_init, "bias_init": bias_init})
        scale = kwargs.pop("scale", 0.006)  # nn/linear.py:31
        self.add_parameter("weight", weight_init_fn(weight_init, **kwargs, scale=scale)((d_in, d_out)),
                           weight_shard_axis)
        if self.use_bias:
            self.add_parameter("bias", weight_init_fn(bias_init, **kwargs)((d_out,)), bias_shard_axis)

    # `relu`, `mask` and `out` are the hooks MLP / Attention use to fuse ReLU into this layer's GEMM
    # epilogues and to keep GEMM-to-GEMM tensors in bf16 only (bf16 mode).
    def forward(self, inputs, relu: bool = False, out: str = "f32", relu_bits=None, out_tensor=None, peer_out: int = 0):
        """outputs = inputs @ W (+ b) — one GEMM with the bias (and ReLU) in its epilogue.
        `out_tensor` / `peer_out`: a preallocated bf16 buffer for the result and the address in the other
        tensor-parallel rank's memory where every tile is stored as well (distributed/peer.py)."""
        bias = self.bias.data if self.use_bias else None
        outputs = F.linear_forward(inputs, self, "weight", bias=bias, relu=relu, out=out, relu_bits=relu_bits,
                                   out_tensor=out_tensor, peer_out=peer_out)
        self.ctx["inputs"] = inputs
        return outputs

    def backward(self, d_out, mask=None, out: str = "f32", on_dx=None, out_tensor=None, peer_out: int = 0):
        """dW = x^T dy (replaces, never accumulates), db = sum dy, returns dx = dy W^T.
        `on_dx(dx)` (optional) is called as soon as dx is enqueued — the tensor-parallel callers
        start their all-reduce there so it overlaps the dW / db kernels."""
        inputs = self.ctx.pop("inputs")
        d_in = F.linear_backward_input(d_out, self, "weight", mask=mask, out=out, out_tensor=out_tensor,
                                       peer_out=peer_out)
        margin = TP_OVERLAP_SM_MARGIN if on_dx is not None else 0
        if on_dx is not None:
            on_dx(d_in)
        if margin:
            ops.set_sm_margin(margin)  # leave SMs for the all-reduce next to the dW GEMM
        self.update_parameter("weight", gradient=F.linear_backward_weight(inputs, d_out, may_defer=True))
        if self.use_bias:
            self.update_parameter("bias", gradient=F.bias_grad(d_out))
        if margin:
            ops.set_sm_margin(0)
        return d_in
