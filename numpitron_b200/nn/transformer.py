"""Transformer — drop-in for the reference's nn/transformer.py."""
from __future__ import annotations

from functools import partial

from .. import distributed as npdist
from ..distributed import peer
from .. import runtime as rt
from .embedding import InputEmbedding, OutputEmbedding, PositionalEncoding
from .model import Model
from .transformer_block import TransformerBlock


def set_dp_sync(mode: str) -> None:
    rt.set_dp_sync(mode)


def _gradients_below(layer):
    """Every gradient held by the (nested) layers of a Model, in definition order."""
    out = []
    for sub in layer.layers.values():
        if isinstance(sub, Model):
            out += _gradients_below(sub)
        out += [p.gradient for p in sub.parameters.values() if p.gradient is not None]
    return out


class Transformer(Model):
    def __init__(self, d_model: int, n_heads: int, n_layers: int, vocab_size: int, seq_len: int, **kwargs):
        super().__init__()
        self.add_settings({"d_model": d_model, "n_heads": n_heads, "n_layers": n_layers,
                           "vocab_size": vocab_size, "seq_len": seq_len})
        block = partial(TransformerBlock, d_model, n_heads, **kwargs)
        self.add_layer("input_embedding", InputEmbedding(d_model, vocab_size, **kwargs))
        self.add_layer("positional_encoding", PositionalEncoding(d_model, seq_len, **kwargs))
        for layer in range(n_layers):
            self.add_layer(f"block_{layer}", block())
        self.add_layer("output_embedding", OutputEmbedding(self.input_embedding))  # tied

    def forward(self, inputs):
        """Model.forward (nn/model.py:15-18).  Without tensor parallelism nothing sits between the embedding
        gather and the positional add, so both run in one kernel (same arithmetic, one pass less)."""
        if npdist.tensor_parallel_size() > 1:
            # tp = 2 peer exchange: size / rewind the per-step arena of the row-sharded LayerNorms
            peer.prepare(inputs.numel(), self.d_model, 2 * self.n_layers)
        if npdist.tensor_parallel_size() > 1 and not self.input_embedding.gathers_full_table(inputs.numel()):
            return super().forward(inputs)
        layers = list(self.layers.values())
        outputs = self.input_embedding.forward(inputs, pos=self.positional_encoding.encoding)
        for layer in layers[2:]:
            outputs = layer(outputs)
        return outputs

    def backward(self, d_out):
        """Reverse pass; after each TOP-LEVEL layer its own parameters are all-reduced (SUM) over the
        data-parallel group — which in practice is only the tied embedding gradient, because a
        TransformerBlock's own `parameters` dict is empty (reference: transformer.py:48-59, Q11).

        `rt.set_dp_sync("all")` (non-reference): the gradients of every parameter below a
        block are summed over the group as well — one coalesced bucket per block, started asynchronously as
        soon as the block's backward is enqueued, so the transfers run under the backward pass of the blocks
        below; all buckets are waited for before the optimizer may run."""
        blocks = [layer for layer in self.layers.values() if isinstance(layer, TransformerBlock)]
        if blocks:
            blocks[0].input_grad_f32 = True  # its dX goes to the embedding layers (fp32 readers)
        dp = npdist.data_parallel_size() > 1
        sync_all = dp and rt.dp_sync() == "all"
        pending = []
        for layer in list(self.layers.values())[::-1]:
            d_out = layer.backward(d_out)
            if dp:
                if sync_all and isinstance(layer, Model):
                    pending.append(npdist.all_reduce_bucket_async(_gradients_below(layer), group=npdist.data_parallel_group()))
                for parameter in layer.parameters.values():
                    npdist.all_reduce(parameter.gradient, group=npdist.data_parallel_group())
        for handle in pending:
            handle.wait()
        return d_out

    @classmethod
    def from_dict(cls, model_dict: dict[str, dict]):
        transformer = super().from_dict(model_dict)
        transformer.output_embedding.input_embedding = transformer.input_embedding  # tie
        return transformer
