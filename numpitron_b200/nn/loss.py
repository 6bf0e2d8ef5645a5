"""Vocab-parallel softmax cross-entropy — drop-in for the reference's nn/loss.py."""
from __future__ import annotations

import numpy as np

from .. import distributed as npdist
from .. import ops
from .. import runtime as rt
from .core import _host_to_device


def softmax_cross_entropy(inputs, labels):
    """(loss (B,S), d_logits (B,S,V_local)).  Per-token loss, un-normalised gradient (Q9).

    The label range mask uses the process-global chunk bounds set by InputEmbedding, exactly as
    the reference does (loss.py:12-18) — including the vocab-1 quirk at tp == 1 (Q7)."""
    host = isinstance(inputs, np.ndarray)
    if host:
        inputs = _host_to_device(inputs)
    if isinstance(labels, np.ndarray):
        labels = _host_to_device(labels)
    chunk_start = npdist.get_var("embedding_chunk_start")
    chunk_end = npdist.get_var("embedding_chunk_end")
    want_bf16 = rt.is_bf16()
    if npdist.tensor_parallel_size() > 1:
        group = npdist.tensor_parallel_group()
        rowmax = ops.softmax_ce_stage1(inputs)
        npdist.all_reduce(rowmax, op="max", group=group)
        picked, sumexp = ops.softmax_ce_stage2(inputs, labels, rowmax, chunk_start, chunk_end)
        npdist.all_reduce(picked, group=group)
        npdist.all_reduce(sumexp, group=group)
        loss, grad = ops.softmax_ce_stage3(inputs, labels, rowmax, picked, sumexp, chunk_start, chunk_end, want_bf16)
    else:
        loss, grad = ops.softmax_ce_fused(inputs, labels, chunk_start, chunk_end, want_bf16)
    if host:
        return rt.to_numpy(loss), rt.to_numpy(grad)
    return loss, grad
