"""Sampling — the generation loop of the reference's sample_shakespeare.py.

`basic_sample` and `greedy_sample` keep the reference's signatures and arithmetic
(sample_shakespeare.py:13-49): the last position's logits are gathered across the tensor-parallel
group, divided by the temperature, soft-maxed in float64 on the host and one multinomial draw is taken
from the caller's NumPy generator — so with the same generator the drawn tokens are the reference's.
The device does what it is good at, the forward pass; the 17…50 257 logits of one position are a
few hundred bytes to kilobytes of D2H per token.  `generate` is the loop of `generate()` (:67-75)
without the tokenizer / CLI around it: context window of `seq_len` tokens, full forward, sample.
(No KV cache, like the reference: the whole window is recomputed per token.)
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator

from . import distributed as npdist
from .runtime import to_numpy


def _softmax64(x: np.ndarray) -> np.ndarray:
    """nn/activation.py:6-9 on a float64 vector."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def _last_logits(logits) -> np.ndarray:
    assert logits.shape[0] == 1, "Only single batch size supported"
    return np.ascontiguousarray(to_numpy(logits[0, -1, :]), dtype=np.float32)


def basic_sample(logits, temperature: float, rng: Generator, vocab_size: int) -> int:
    """Softmax-with-temperature sampling of the last position (sample_shakespeare.py:13-25)."""
    local = _last_logits(logits)
    logits_full = np.empty(vocab_size, dtype=np.float32)
    npdist.all_gather(local, logits_full, group=npdist.tensor_parallel_group())
    probabilities = _softmax64((logits_full / temperature).astype(np.float64))
    return int(rng.multinomial(1, probabilities).argmax())


def greedy_sample(logits) -> int:
    """Arg-max over the vocabulary shards (sample_shakespeare.py:28-49)."""
    local = _last_logits(logits)
    rank_max_index = int(local.argmax())
    rank_max_value = local[rank_max_index]
    rank_max_index += npdist.tensor_parallel_rank() * len(local)
    tp = npdist.tensor_parallel_size()
    max_indices = np.zeros(tp, dtype=np.float32)
    max_values = np.zeros(tp, dtype=np.float32)
    max_values[npdist.tensor_parallel_rank()] = rank_max_value
    max_indices[npdist.tensor_parallel_rank()] = rank_max_index
    npdist.all_reduce(max_values, group=npdist.tensor_parallel_group())
    npdist.all_reduce(max_indices, group=npdist.tensor_parallel_group())
    return int(max_indices[max_values.argmax()])


def generate(transformer, tokens: list[int], num_samples: int, temperature: float | None, rng: Generator | None,
             vocab_size: int) -> list[int]:
    """The loop of sample_shakespeare.py:67-75; `temperature=None` = greedy.  Returns the sampled tokens."""
    tokens = list(tokens)
    out = []
    for _ in range(num_samples):
        tokens = tokens[-transformer.seq_len:]
        logits = transformer(np.asarray([tokens], dtype=np.uint16))
        nxt = greedy_sample(logits) if temperature is None else basic_sample(logits, temperature, rng, vocab_size)
        tokens.append(nxt)
        out.append(nxt)
    return out
