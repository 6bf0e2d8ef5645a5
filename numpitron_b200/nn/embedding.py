"""Vocab-parallel input embedding, tied output embedding, sinusoidal positions.
Drop-in for the reference's nn/embedding.py, including its chunk-bound behaviour."""
from __future__ import annotations

import numpy as np

from .. import distributed as npdist
from .. import ops
from .. import runtime as rt
from . import functional as F
from .core import Layer, weight_init_fn


class InputEmbedding(Layer):
    def __init__(self, d_model: int, vocab_size: int, weight_init: str = "scaled_normal", **kwargs):
        super().__init__()
        self.add_settings({"d_model": d_model, "vocab_size": vocab_size})
        # Stored (d_model, vocab) like the reference (embedding.py:19-23), sharded over vocab.
        self.add_parameter("embedding", weight_init_fn(weight_init, **kwargs)((d_model, vocab_size)), shard_axis=1)
        npdist.set_var("embedding_chunk_start", 0)
        npdist.set_var("embedding_chunk_end", self.embedding.data.shape[1] - 1)  # sic: vocab-1 (Q7)

    def scatter(self, src: int = 0):
        total_vocab_size = self.embedding.data.shape[1]
        super().scatter(src)
        tp = npdist.tensor_parallel_size()
        if tp > 1:  # embedding.py:32-37 (raises when vocab % tp == 0, like the reference: Q8)
            chunk_size = total_vocab_size // tp
            chunks = np.arange(0, total_vocab_size, chunk_size)
            chunks[1:] += total_vocab_size % tp
            npdist.set_var("embedding_chunk_start", int(chunks[npdist.tensor_parallel_rank()]))
            npdist.set_var("embedding_chunk_end", int(chunks[npdist.tensor_parallel_rank() + 1]))

    def all_gather(self):
        super().all_gather()
        npdist.set_var("embedding_chunk_start", 0)
        npdist.set_var("embedding_chunk_end", self.embedding.data.shape[1] - 1)

    def _bounds(self):
        if self.is_scattered:
            return npdist.get_var("embedding_chunk_start"), npdist.get_var("embedding_chunk_end")
        return 0, self.embedding.data.shape[1]  # un-scattered: the mask is computed but unused

    def forward(self, inputs, pos=None):
        """inputs (B,S) uint16 tokens -> (B,S,d); out-of-chunk tokens give zero rows.
        `pos`: the positional-encoding table; when given (Transformer.forward at tp = 1, where no all-reduce
        sits between the two layers) `encoding + x` of PositionalEncoding.forward (embedding.py:125-128) is
        done by the same kernel, with the same single fp32 addition."""
        lo, hi = self._bounds()
        seq_len = inputs.shape[-1]
        outputs = ops.embedding_gather(inputs, self.embedding.data, pos, seq_len, lo, hi, self.is_scattered,
                                       want_bf16=pos is not None and rt.is_bf16())
        if self.is_scattered and npdist.tensor_parallel_size() > 1:
            npdist.all_reduce(outputs, group=npdist.tensor_parallel_group())
        self.ctx["inputs"] = inputs
        self.ctx["bounds"] = (lo, hi)
        return outputs

    def backward(self, d_out):
        """Ordered scatter-add onto the gradient left by OutputEmbedding.backward (Q14)."""
        gradient = self.embedding.gradient
        if gradient is None:
            gradient = rt.zeros(tuple(self.embedding.data.shape))
        lo, hi = self.ctx.pop("bounds")
        ops.embedding_scatter_add(self.ctx.pop("inputs"), d_out, gradient, lo, hi, self.is_scattered)
        self.update_parameter("embedding", gradient=gradient)
        return None


class OutputEmbedding(Layer):
    def __init__(self, input_embedding: Layer, **kwargs):
        super().__init__()
        self.input_embedding = input_embedding

    def forward(self, inputs):
        outputs = F.linear_forward(inputs, self.input_embedding, "embedding")  # x @ E
        self.ctx["inputs"] = inputs
        return outputs

    def backward(self, d_out):
        self.input_embedding.update_parameter(
            "embedding", gradient=F.linear_backward_weight(self.ctx.pop("inputs"), d_out))
        # d @ E^T.  In bf16 mode this dX is read by the last block's LayerNorm backward, which takes bf16
        # (the value it would otherwise be cast to): let the GEMM epilogue write bf16 directly.
        d_model = self.input_embedding.embedding.data.shape[0]
        out = "bf16" if rt.is_bf16() and ops.layernorm_bf16_input_supported(d_model) else "f32"
        d_out = F.linear_backward_input(d_out, self.input_embedding, "embedding", out=out)
        if self.is_scattered and npdist.tensor_parallel_size() > 1:
            npdist.all_reduce(d_out, group=npdist.tensor_parallel_group())
        return d_out

    @classmethod
    def from_dict(cls, layer_dict: dict[str, dict]):
        return cls(None)


class PositionalEncoding(Layer):
    def __init__(self, d_model: int, seq_len: int, **kwargs):
        super().__init__()
        self.add_settings({"d_model": d_model, "seq_len": seq_len})
        # Built on the host exactly as the reference does (float64 math, float32 storage).
        pos = np.expand_dims(np.arange(0, seq_len), -1)
        _2i = np.arange(d_model, step=2) / d_model
        encoding = np.zeros((seq_len, d_model), dtype=np.float32)
        encoding[:, 0::2] = np.sin(pos / (10000**_2i))
        encoding[:, 1::2] = np.cos(pos / (10000**_2i))
        self.encoding = rt.from_numpy(encoding)

    def forward(self, inputs):
        seq_len = inputs.shape[1]
        return ops.add_pos(inputs, self.encoding, seq_len, want_bf16=rt.is_bf16())
