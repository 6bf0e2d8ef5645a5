"""train_step / validation_step — the hot loop body of the reference (train_shakespeare.py:14-27),
plus a CUDA-graph wrapper that replays the whole step (H2D of the tokens included) as one launch."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from . import runtime as rt
from .nn import softmax_cross_entropy
from .nn.core import _host_to_device


def _tokens(a):
    return _host_to_device(a) if isinstance(a, np.ndarray) else a


def train_step(transformer, optimizer, x, y) -> float:
    """forward -> softmax-cross-entropy -> backward -> Adam; returns loss.mean() as a float."""
    return float(train_step_device(transformer, optimizer, _tokens(x), _tokens(y)).item())


def train_step_device(transformer, optimizer, x, y) -> torch.Tensor:
    """Same, but the mean loss stays on the device (1-element tensor)."""
    y_hat = transformer(x)
    loss, d_out = softmax_cross_entropy(y_hat, y)
    transformer.backward(d_out)
    optimizer.step()
    return ops.mean(loss)


def validation_step(transformer, x, y) -> float:
    y_hat = transformer(_tokens(x))
    loss, _ = softmax_cross_entropy(y_hat, _tokens(y))
    return float(ops.mean(loss).item())


class GraphedTrainStep:
    """Captures one full train step into a CUDA graph and replays it.

    `step(x, y)` takes HOST uint16 arrays: they are copied into pinned staging buffers, the graph
    (H2D copies + every kernel of the step + the D2H copy of the mean loss) is launched once, and
    the loss is read back.  Launch-bound inner loops (~150 kernels per step at the Shakespeare
    shape) become a single graph launch.
    """

    def __init__(self, transformer, optimizer, batch: int, seq_len: int, warmup: int = 2):
        rt.require_gpu("GraphedTrainStep")
        self.transformer, self.optimizer = transformer, optimizer
        self.x_host = torch.empty((batch, seq_len), dtype=torch.uint16).pin_memory()
        self.y_host = torch.empty((batch, seq_len), dtype=torch.uint16).pin_memory()
        self.loss_host = torch.zeros((1,), dtype=torch.float32).pin_memory()
        self.x_dev = rt.empty((batch, seq_len), torch.uint16)
        self.y_dev = rt.empty((batch, seq_len), torch.uint16)
        self.graph = None
        self.warmup = warmup
        self._eager_steps = 0
        self.stream = torch.cuda.Stream()

    def _body(self):
        self.x_dev.copy_(self.x_host, non_blocking=True)
        self.y_dev.copy_(self.y_host, non_blocking=True)
        loss = train_step_device(self.transformer, self.optimizer, self.x_dev, self.y_dev)
        self.loss_host.copy_(loss, non_blocking=True)

    def step_async(self, x: np.ndarray, y: np.ndarray) -> None:
        self.x_host.numpy()[...] = x
        self.y_host.numpy()[...] = y
        if self._eager_steps < self.warmup:
            with torch.cuda.stream(self.stream):
                self._body()
            self._eager_steps += 1
            return
        if self.graph is None:
            self.stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                self._body()
            self.graph = g
        with torch.cuda.stream(self.stream):
            self.graph.replay()

    def step(self, x: np.ndarray, y: np.ndarray) -> float:
        self.step_async(x, y)
        self.stream.synchronize()
        return float(self.loss_host[0])
