"""TransformerBlock — drop-in for the reference's nn/transformer_block.py (quirks Q1/Q2 kept)."""
from __future__ import annotations

from .attention import Attention
from .mlp import MLP
from .model import Model
from .normalization import LayerNorm


class TransformerBlock(Model):
    def __init__(self, d_model: int, n_heads: int, **kwargs):
        super().__init__()
        self.add_settings({"d_model": d_model, "n_heads": n_heads})
        self.add_layer("norm1", LayerNorm(d_model, **kwargs))
        self.add_layer("attention", Attention(d_model, n_heads, d_model // n_heads, **kwargs))
        self.add_layer("norm2", LayerNorm(d_model, **kwargs))
        self.add_layer("mlp", MLP(d_model, d_model * 4, d_model, **kwargs))

    def forward(self, inputs):
        # LayerNorm is applied to the SUB-LAYER OUTPUT, then the residual is added (Q1); the add is
        # fused into the LayerNorm kernel.
        outputs = self.norm1.forward(self.attention(inputs), residual=inputs)
        outputs = self.norm2.forward(self.mlp(outputs), residual=outputs)
        return outputs

    def backward(self, d_out):
        # The residual gradient is dropped, exactly as the reference does (Q2).
        d_out = self.mlp.backward(self.norm2.backward(d_out, gemm_only=True))
        d_out = self.attention.backward(self.norm1.backward(d_out, gemm_only=True))
        return d_out
