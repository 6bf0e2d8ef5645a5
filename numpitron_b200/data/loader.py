"""DataLoader — drop-in for the reference's data/loader.py with the token file resident in HBM.

The reference memory-maps the uint16 token file and, per batch, draws `batch_size` start indices from
its NumPy generator, fancy-indexes two `(batch, seq_len)` windows on the host (inputs, and labels =
the same windows shifted by one, loader.py:49-54) and, under data parallelism, scatters the batch
from DP rank 0 (loader.py:56-78).  At several million tokens/s that host work and the 2 x B x S x 2
bytes of H2D per step are on the critical path.  Here the file is copied to the GPU once; per batch
only the start indices (8 bytes each) are drawn on the host — from the same generator, with the same
call, so the random stream and therefore every batch is the reference's, bit for bit — and one small
kernel (`npt_window_gather`) cuts the windows out on the device.  Under data parallelism the start
indices of DP rank 0 are broadcast (what the reference's scatter of the batch amounts to) and every
rank gathers its own `batch_size // dp` rows.

`iter_epoch()` yields device tensors (uint16, shape (B_local, seq_len)) that `train_step` /
`GraphedTrainStep.load_resident` take as they are.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from numpy.random import Generator

from .. import _lib
from .. import distributed as npdist
from .. import ops
from .. import runtime as rt


class DataLoader:
    def __init__(self, dataset_path: str | Path, seq_len: int, batch_size: int, rng: Generator):
        rt.require_gpu("DataLoader")
        self.dataset_path = dataset_path
        self.rng = rng
        host = np.fromfile(self.dataset_path, dtype=np.uint16)   # the reference memmaps the same bytes
        self.n_tokens = len(host)
        self.batches_per_epoch = self.n_tokens // (seq_len * batch_size)   # loader.py:33
        self.batch_size = batch_size
        self.seq_len = seq_len
        # staged once; uint16 has no torch arithmetic, it is only ever indexed by our kernels
        self.tokens = torch.from_numpy(host.view(np.int16).copy()).to(rt.device()).view(torch.uint16)

    def _starts(self) -> np.ndarray:
        """The reference's draw (loader.py:49-51), same generator call, same dtype."""
        return self.rng.integers(0, self.n_tokens - 1 - self.seq_len, size=self.batch_size)

    def next_batch(self):
        """One (inputs, labels) pair on the device."""
        lib = _lib.load()
        starts = np.ascontiguousarray(self._starts(), dtype=np.int64)
        dp = npdist.data_parallel_size()
        if dp > 1:
            # the reference scatters DP rank 0's batch; its rows are determined by rank 0's start indices
            npdist.broadcast(starts, src=0, group=npdist.data_parallel_group())
            local = self.batch_size // dp
            r = npdist.data_parallel_rank()
            starts = np.ascontiguousarray(starts[r * local:(r + 1) * local])
        b = len(starts)
        starts_dev = torch.from_numpy(starts).to(rt.device(), non_blocking=False)
        x = torch.empty((b, self.seq_len), dtype=torch.uint16, device=rt.device())
        y = torch.empty((b, self.seq_len), dtype=torch.uint16, device=rt.device())
        ops.check(lib.npt_window_gather(self.tokens.data_ptr(), self.n_tokens, starts_dev.data_ptr(), x.data_ptr(),
                                        y.data_ptr(), b, self.seq_len, rt.stream()), "npt_window_gather")
        return x, y

    def iter_epoch(self):
        """One pass over the dataset: `batches_per_epoch + 1` batches, like the reference's
        `while batch_idx <= self.batches_per_epoch` (loader.py:47)."""
        def _iter_epoch():
            batch_idx = 0
            while batch_idx <= self.batches_per_epoch:
                yield self.next_batch()
                batch_idx += 1

        # the reference returns the tqdm bar itself (loader.py:84-91): its training loop calls
        # set_description() / refresh() on it (train_shakespeare.py:103-106)
        import tqdm

        return tqdm.tqdm(_iter_epoch(), leave=False, total=self.batches_per_epoch, disable=npdist.world_rank() != 0)
