"""Linear — drop-in for the reference's nn/linear.py (same constructor, forward, backward)."""
from __future__ import annotations

from . import functional as F
from .core import Layer, weight_init_fn


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
    def forward(self, inputs, relu: bool = False, out: str = "f32", relu_bits=None):
        """outputs = inputs @ W (+ b) — one GEMM with the bias (and ReLU) in its epilogue."""
        bias = self.bias.data if self.use_bias else None
        outputs = F.linear_forward(inputs, self, "weight", bias=bias, relu=relu, out=out, relu_bits=relu_bits)
        self.ctx["inputs"] = inputs
        return outputs

    def backward(self, d_out, mask=None, out: str = "f32", on_dx=None):
        """dW = x^T dy (replaces, never accumulates), db = sum dy, returns dx = dy W^T.
        `on_dx(dx)` (optional) is called as soon as dx is enqueued — the tensor-parallel callers
        start their all-reduce there so it overlaps the dW / db kernels."""
        inputs = self.ctx.pop("inputs")
        d_in = F.linear_backward_input(d_out, self, "weight", mask=mask, out=out)
        if on_dx is not None:
            on_dx(d_in)
        self.update_parameter("weight", gradient=F.linear_backward_weight(inputs, d_out))
        if self.use_bias:
            self.update_parameter("bias", gradient=F.bias_grad(d_out))
        return d_in
