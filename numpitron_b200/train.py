"""train_step / validation_step — the hot loop body of the reference (train_shakespeare.py:14-27),
plus a CUDA-graph wrapper that replays the whole step (H2D of the tokens included) as one launch."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from . import runtime as rt
from .nn import softmax_cross_entropy
from .nn.core import _host_to_device, storage_generation


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
    shape) become a single graph launch.  `step_resident()` replays the same step on token
    buffers that already live in HBM (no host copies in the graph).
    """

    def __init__(self, transformer, optimizer, batch: int, seq_len: int, warmup: int = 2):
        rt.require_gpu("GraphedTrainStep")
        self.transformer, self.optimizer = transformer, optimizer
        n = batch * seq_len * 2
        self._x_pin, self._y_pin, self._loss_pin = ops.PinnedBuffer(n), ops.PinnedBuffer(n), ops.PinnedBuffer(4)
        self.x_host = self._x_pin.view(np.uint16, (batch, seq_len))
        self.y_host = self._y_pin.view(np.uint16, (batch, seq_len))
        self.loss_host = self._loss_pin.view(np.float32, (1,))
        self.h2d_bytes_per_step, self.d2h_bytes_per_step = 2 * n, 4
        self.x_dev = rt.empty((batch, seq_len), torch.uint16)
        self.y_dev = rt.empty((batch, seq_len), torch.uint16)
        self.loss_dev = rt.zeros((1,))
        self.graph = None            # H2D + step + D2H
        self.graph_resident = None   # step only, tokens already in x_dev / y_dev
        self.launches_per_step = None
        self.warmup = warmup
        self._eager_steps = 0
        self.stream = torch.cuda.Stream()
        # parameters / shadows were produced on the caller's stream: order the step stream after it
        self.stream.wait_stream(torch.cuda.current_stream())
        self._generation = storage_generation()

    # -- bodies -----------------------------------------------------------------------------------
    def _compute(self):
        # nothing reads the weight gradients between backward and Adam here: skip their split-K reduce
        with rt.deferred_wgrad():
            loss = train_step_device(self.transformer, self.optimizer, self.x_dev, self.y_dev)
        ops.check(_lib.load().npt_memcpy_async(self.loss_dev.data_ptr(), loss.data_ptr(), 4, 3, rt.stream()),
                  "npt_memcpy_async")

    def _body(self):
        self._x_pin.to_device(self.x_dev)
        self._y_pin.to_device(self.y_dev)
        self._compute()
        self._loss_pin.from_device(self.loss_dev)

    def _check_storage(self) -> None:
        """The captured graphs hold raw pointers of every parameter, Adam moment and bf16 shadow.  When any
        of that storage was replaced since the capture (checkpointing gathers and re-scatters the model and
        the optimizer: train_shakespeare.py:30-41), the graphs would keep training the orphaned buffers —
        drop them, run the eager warm-up steps again on the new storage and re-capture."""
        gen = storage_generation()
        if gen != self._generation:
            self._generation = gen
            self.graph = self.graph_resident = None
            self._eager_steps = 0
            self.stream.wait_stream(torch.cuda.current_stream())   # the new buffers were filled on the caller's stream

    def _warm(self, body) -> bool:
        self._check_storage()
        if self._eager_steps < self.warmup:
            with torch.cuda.stream(self.stream):
                body()
            self._eager_steps += 1
            return True
        return False

    def _capture(self, body):
        self.stream.synchronize()
        lib = _lib.load()
        before = lib.npt_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            body()
        self.launches_per_step = int(lib.npt_launch_count() - before)
        return g

    # -- host-buffer API (what a user calls) ----------------------------------------------------------
    def step_async(self, x: np.ndarray, y: np.ndarray) -> None:
        self.x_host[...] = x
        self.y_host[...] = y
        if self._warm(self._body):
            return
        if self.graph is None:
            self.graph = self._capture(self._body)
        with torch.cuda.stream(self.stream):
            self.graph.replay()

    def step(self, x: np.ndarray, y: np.ndarray) -> float:
        self.step_async(x, y)
        self.stream.synchronize()
        return float(self.loss_host[0])

    # -- device-resident API (tokens already in HBM) -----------------------------------------------------
    def load_resident(self, x_dev: torch.Tensor, y_dev: torch.Tensor) -> None:
        """Device-to-device copy of a resident batch into the graph's token buffers (stream-ordered)."""
        lib = _lib.load()
        # x_dev / y_dev were produced on the caller's stream (DataLoader.next_batch launches its gather there):
        # the copies below must wait for it, and the caching allocator must not hand the source buffers to
        # someone else while the step stream still reads them
        self.stream.wait_stream(torch.cuda.current_stream())
        x_dev.record_stream(self.stream)
        y_dev.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            n = self.x_dev.numel() * 2
            ops.check(lib.npt_memcpy_async(self.x_dev.data_ptr(), x_dev.data_ptr(), n, 3, rt.stream()), "memcpy")
            ops.check(lib.npt_memcpy_async(self.y_dev.data_ptr(), y_dev.data_ptr(), n, 3, rt.stream()), "memcpy")

    def step_resident_async(self) -> None:
        if self._warm(self._compute):
            return
        if self.graph_resident is None:
            self.graph_resident = self._capture(self._compute)
        with torch.cuda.stream(self.stream):
            self.graph_resident.replay()
