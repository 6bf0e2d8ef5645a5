"""Synthetic token stream + batch draw: what the benchmark and the multi-GPU parity
tests feed the training step when no dataset file exists.

`draw_batch` is the reference loader's draw (`data/loader.py:48-54`): `batch` start indices from the
caller's NumPy generator, inputs = the windows, labels = the same windows shifted by one token.
`tests/test_oracle.py` pins both functions to the oracle's (which are pinned to the reference loader).
"""
from __future__ import annotations

import numpy as np


def synthetic_tokens(vocab: int, n: int = 2_000_000, seed: int = 1234) -> np.ndarray:
    """Uniform uint16 tokens over [0, vocab) — deliberately includes token vocab-1 (quirk Q7)."""
    return np.random.default_rng(seed).integers(0, vocab, n).astype(np.uint16)


def draw_batch(data: np.ndarray, rng: np.random.Generator, batch: int, seq_len: int):
    """One global batch: (inputs, labels), each (batch, seq_len) uint16."""
    start = rng.integers(0, len(data) - 1 - seq_len, size=batch)[:, None]
    offs = np.arange(seq_len)
    return data[start + offs], data[start + 1 + offs]


def local_rows(a: np.ndarray, dp: int, dp_rank: int) -> np.ndarray:
    """The rows of a global batch that DP rank `dp_rank` trains on (`loader.py:56-80`: Scatterv along the
    batch axis == np.array_split)."""
    return np.ascontiguousarray(np.array_split(a, dp, axis=0)[dp_rank])
