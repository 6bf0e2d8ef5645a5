"""TransformerBlock — drop-in for the reference's nn/transformer_block.py (quirks Q1/Q2 kept)."""
from __future__ import annotations

from .. import ops
from ..distributed import peer
from .. import runtime as rt
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

    # Set by Transformer.backward on the first block: its input gradient feeds the embedding layers,
    # which read fp32.
    input_grad_f32 = False

    def _sub_dtype(self) -> str:
        """bf16 mode: the sub-layer outputs LayerNorm normalises, and the dX that reaches a LayerNorm's
        backward, leave their GEMM as bf16 alone (the residual stream itself stays fp32): half the bytes to
        store, to all-reduce across tensor-parallel ranks and for LayerNorm to read."""
        if not (rt.is_bf16() and ops.layernorm_bf16_input_supported(self.d_model)) or rt.ln_inputs_fp32():
            return "f32"
        # tp = 2: the partial results are exchanged through peer memory and reduced inside LayerNorm
        return "peer" if peer.enabled() else "bf16"

    def forward(self, inputs):
        # LayerNorm is applied to the SUB-LAYER OUTPUT, then the residual is added (Q1); the add is
        # fused into the LayerNorm kernel.
        sub = self._sub_dtype()
        outputs = self.norm1.forward(self.attention.forward(inputs, out=sub), residual=inputs)
        outputs = self.norm2.forward(self.mlp.forward(outputs, out=sub), residual=outputs)
        return outputs

    def backward(self, d_out):
        # The residual gradient is dropped, exactly as the reference does (Q2).
        sub = self._sub_dtype()
        d_out = self.mlp.backward(self.norm2.backward(d_out, gemm_only=True), out=sub)
        d_out = self.attention.backward(self.norm1.backward(d_out, gemm_only=True),
                                        out="f32" if self.input_grad_f32 else sub)
        return d_out
