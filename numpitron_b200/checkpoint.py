"""Checkpoint interop with the reference.

The reference's `checkpoint()` (`train_shakespeare.py:30-41`) gathers the tensor-parallel shards, writes

    {"model_parameters": model.to_dict(), "optimizer_parameters": optimizer.to_dict(),
     "epoch": epoch, "validation_loss": loss}

with `np.save(path, state, allow_pickle=True)` and scatters again; `train()` resumes from such a file
with `Transformer.from_dict` / `Adam.from_dict` (`train_shakespeare.py:87-93`).  The two functions below
are those two code paths for the B200 layers: every array in the file is a host NumPy float32 array in
the reference's nested-dict layout, so a file written here loads in the reference and a file written by
the reference (tests/golden/tiny_checkpoint.npy) trains on here — `tests/test_train_gpu.py` checks both
directions against the reference's own resumed step.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .nn import Transformer
from .optimizer import Adam


def save_checkpoint(model, optimizer, epoch: int, loss: float, save_path) -> None:
    """`checkpoint(model, optimizer, epoch, loss, save_path)` of train_shakespeare.py:30-41."""
    model.all_gather()
    optimizer.all_gather()
    state = {
        "model_parameters": model.to_dict(),
        "optimizer_parameters": optimizer.to_dict(),
        "epoch": epoch,
        "validation_loss": loss,
    }
    np.save(save_path, state, allow_pickle=True)
    model.scatter()
    optimizer.scatter()


def load_checkpoint(save_path):
    """The resume branch of train_shakespeare.py:87-93: returns (transformer, optimizer, epoch,
    validation_loss); like there, the caller scatters both afterwards."""
    path = Path(save_path)
    state = np.load(path, allow_pickle=True)[()]
    transformer = Transformer.from_dict(state["model_parameters"])
    optimizer = Adam.from_dict(state["optimizer_parameters"], transformer)
    return transformer, optimizer, state["epoch"], state["validation_loss"]
