"""Vocab-parallel input embedding, tied output embedding, sinusoidal positions.
Drop-in for the reference's nn/embedding.py, including its chunk-bound behaviour."""
from __future__ import annotations

import numpy as np

from .. import distributed as npdist
from ..distributed import peer
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
            # every rank's bounds and shard width (np.array_split), for the small-vocabulary path of forward()
            self._chunks = [int(c) for c in chunks[:tp + 1]]
            self._widths = [total_vocab_size // tp + (1 if r < total_vocab_size % tp else 0) for r in range(tp)]

    def all_gather(self):
        super().all_gather()
        npdist.set_var("embedding_chunk_start", 0)
        npdist.set_var("embedding_chunk_end", self.embedding.data.shape[1] - 1)

    def _bounds(self):
        if self.is_scattered:
            return npdist.get_var("embedding_chunk_start"), npdist.get_var("embedding_chunk_end")
        return 0, self.embedding.data.shape[1]  # un-scattered: the mask is computed but unused

    def gathers_full_table(self, n_tokens: int) -> bool:
        """Tensor parallelism with a vocabulary much smaller than the batch (the Shakespeare shape: 65 tokens,
        32768 rows): instead of all-reducing the (B,S,d) fp32 result of the masked per-shard gathers
        (embedding.py:61-62, 50 MB per step), the shards themselves are all-gathered (d x V floats) and every rank
        gathers from the whole table.  Same bits: each token lies in exactly one rank's chunk, so the
        reference's sum has one non-zero term.  Only when the reference's chunk bounds coincide with the shard
        widths (they do whenever vocab % tp <= 1)."""
        tp = npdist.tensor_parallel_size()
        if not (self.is_scattered and tp > 1 and rt.has_gpu() and getattr(self, "_chunks", None)):
            return False
        consistent = all(self._chunks[r + 1] - self._chunks[r] == self._widths[r] for r in range(tp)) and \
            self._chunks[tp] == sum(self._widths)
        return consistent and 8 * sum(self._widths) <= n_tokens

    def _full_table(self):
        """(d, V) table re-assembled from the shards: own shard -> padded staging row block, one small NCCL
        all-gather, tp pitched copies into place."""
        import torch
        import torch.distributed as dist

        lib = __import__("numpitron_b200._lib", fromlist=["load"]).load()
        tp, me = npdist.tensor_parallel_size(), npdist.tensor_parallel_rank()
        E = self.embedding.data
        d, wmax, V = E.shape[0], max(self._widths), sum(self._widths)
        stage = rt.zeros((d, wmax)) if wmax != E.shape[1] else E
        if stage is not E:
            ops.check(lib.npt_memcpy2d_async(stage.data_ptr(), wmax * 4, E.data_ptr(), E.stride(0) * 4, E.shape[1] * 4, d, rt.stream()),
                      "npt_memcpy2d_async")
        gathered = rt.empty((tp, d, wmax))
        dist.all_gather_into_tensor(gathered, stage, group=npdist.tensor_parallel_group().pg)
        full = rt.empty((d, V))
        for r in range(tp):
            ops.check(lib.npt_memcpy2d_async(full.data_ptr() + self._chunks[r] * 4, V * 4, gathered[r].data_ptr(), wmax * 4,
                                             self._widths[r] * 4, d, rt.stream()), "npt_memcpy2d_async")
        return full

    def forward(self, inputs, pos=None):
        """inputs (B,S) uint16 tokens -> (B,S,d); out-of-chunk tokens give zero rows.
        `pos`: the positional-encoding table; when given (Transformer.forward at tp = 1, where no all-reduce
        sits between the two layers) `encoding + x` of PositionalEncoding.forward (embedding.py:125-128) is
        done by the same kernel, with the same single fp32 addition."""
        lo, hi = self._bounds()
        seq_len = inputs.shape[-1]
        if self.gathers_full_table(inputs.numel()):
            full = self._full_table()
            outputs = ops.embedding_gather(inputs, full, pos, seq_len, 0, full.shape[1], False,
                                           want_bf16=pos is not None and rt.is_bf16())
            self.ctx["inputs"] = inputs
            self.ctx["bounds"] = (lo, hi)
            return outputs
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
        out = "bf16" if rt.is_bf16() and ops.layernorm_bf16_input_supported(d_model) and not rt.ln_inputs_fp32() else "f32"
        tp = self.is_scattered and npdist.tensor_parallel_size() > 1
        if tp and out == "bf16" and peer.enabled():
            # tp = 2: no all-reduce kernel — the partial goes to the other rank's memory from the GEMM epilogue and
            # the last block's LayerNorm backward adds the two partials while it reads (distributed/peer.py)
            shape = (*d_out.shape[:-1], d_model)
            px = peer.exchange(shape)
            if px is not None:
                slot, send_ptr, recv_ptr = px.slot("bwd", shape)
                n_rows = d_out.numel() // d_out.shape[-1]
                rows = px.rows(n_rows)
                d_in = F.linear_backward_input(d_out, self.input_embedding, "embedding", out="bf16", out_tensor=slot,
                                               peer_out=(send_ptr, *px.peer_rows(n_rows)) if rows else send_ptr)
                d_in._npt_peer, d_in._npt_rows = (px, recv_ptr), rows
                return d_in
        d_out = F.linear_backward_input(d_out, self.input_embedding, "embedding", out=out)
        if tp:
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
