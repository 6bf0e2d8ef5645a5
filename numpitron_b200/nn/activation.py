"""Softmax / ReLU layers — drop-in for the reference's nn/activation.py."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .. import runtime as rt
from .core import Layer


def softmax(inputs, axis: int = -1):
    """Row softmax of a (..., S, S) causal score tensor (entries above the diagonal are treated
    as -inf, which is how the training step calls it: attention.py:48-51)."""
    assert axis in (-1, inputs.ndim - 1)
    if isinstance(inputs, np.ndarray):
        return rt.to_numpy(ops.softmax_causal_fwd(rt.from_numpy(inputs.astype(np.float32)), 1.0))
    return ops.softmax_causal_fwd(inputs, 1.0)


class Softmax(Layer):
    """Causal-masked softmax over the last axis; `divisor` carries attention's 1/sqrt(d_model)."""

    def __init__(self, axis: int = -1, **kwargs):
        super().__init__()
        self.add_setting("axis", axis)

    def forward(self, inputs, divisor: float = 1.0):
        # bf16 mode: the probabilities only feed tensor-core GEMMs (and their own backward)
        outputs = ops.softmax_causal_fwd(inputs, divisor, bf16_only=rt.is_bf16())
        self.ctx["outputs"] = outputs
        self.ctx["divisor"] = divisor
        return outputs

    def backward(self, d_out):
        """P*(dP - sum dP*P) (masked, / divisor): the compact form of activation.py:22-31."""
        outputs = self.ctx.pop("outputs")
        return ops.softmax_causal_bwd(outputs, d_out, self.ctx.pop("divisor"), bf16_only=rt.is_bf16())


class ReLU(Layer):
    def __init__(self, **kwargs):
        super().__init__()

    def forward(self, inputs):
        self.ctx["inputs"] = inputs
        return ops.relu_fwd(inputs, want_bf16=rt.is_bf16())

    def backward(self, d_out):
        return ops.relu_bwd(self.ctx.pop("inputs"), d_out, want_bf16=rt.is_bf16())
