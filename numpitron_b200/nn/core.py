"""Layer base class, Parameter and initialisers — same protocol as the reference's nn/core.py.

Differences that are invisible through the API: `Parameter.data` / `.gradient` are device
buffers (torch tensors on cuda:LOCAL_RANK) and every forward/backward body is a C-ABI call into
the sm_100a kernels.  NumPy in -> NumPy out at the outermost call; device buffers flow between
layers.
"""
from __future__ import annotations

import functools
from abc import abstractmethod
from dataclasses import dataclass, replace

import numpy as np
import torch
from numpy.random import Generator

from .. import distributed as npdist
from .. import runtime as rt


def _to_f32_tensor(a: np.ndarray) -> torch.Tensor:
    return rt.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def weight_init_fn(name: str = "default", rng: Generator | None = None, scale: float = 1.0, **kwargs):
    """Initialisers with the reference's names and RNG consumption (nn/core.py:10-45).

    Note "scaled_normal" is uniform [0,1) * scale drawn in float64 and cast to float32."""

    def zeros(shape):
        return _to_f32_tensor(np.zeros(shape))

    def ones(shape):
        return _to_f32_tensor(np.ones(shape))

    def in_projection(shape):
        assert rng is not None
        in_, out_ = shape
        return _to_f32_tensor(rng.random((in_, out_)) * 3 / (in_ * out_))

    def scaled_normal(shape):
        assert rng is not None
        return _to_f32_tensor(rng.random(shape) * scale)

    def init_for_testing(shape):
        assert rng is not None
        return _to_f32_tensor((rng.random(shape) + 1) / max(shape))

    return {"zeros": zeros, "ones": ones, "in_projection": in_projection,
            "scaled_normal": scaled_normal, "init_for_testing": init_for_testing}[name]


# Bumped whenever a parameter's STORAGE is replaced (scatter / all_gather / from_dict / a new `data`
# through update_parameter).  A captured CUDA graph holds raw device pointers of every parameter, Adam
# moment and bf16 shadow: train.GraphedTrainStep compares this counter before every replay and re-captures
# when it moved (Adam.step updates in place and does not count).
_STORAGE_GENERATION = [0]


def storage_generation() -> int:
    return _STORAGE_GENERATION[0]


def bump_storage_generation() -> None:
    _STORAGE_GENERATION[0] += 1


@dataclass(frozen=True)
class Parameter:
    data: torch.Tensor
    gradient: torch.Tensor | None
    shard_axis: int | None = None


def _host_to_device(x):
    """NumPy -> device buffer: integers become uint16 tokens, floats become float32."""
    if np.issubdtype(x.dtype, np.integer):
        if x.size and (x.min() < 0 or x.max() > 0xFFFF):
            raise ValueError("token ids must fit uint16 (the reference's memmap dtype)")
        return rt.from_numpy(np.ascontiguousarray(x, dtype=np.uint16))
    return rt.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


def host_boundary(fn):
    """NumPy in -> NumPy out; device buffer in -> device buffer out."""

    @functools.wraps(fn)
    def wrapped(self, x, *args, **kwargs):
        if isinstance(x, np.ndarray):
            out = fn(self, _host_to_device(x), *args, **kwargs)
            return rt.to_numpy(out) if isinstance(out, torch.Tensor) else out
        return fn(self, x, *args, **kwargs)

    wrapped._npt_boundary = True
    return wrapped


class Layer:
    def __init__(self, **kwargs):
        self.ctx = {}  # Backward pass variables.
        self.parameters = {}  # Anything requiring a gradient.
        self.settings = {}  # Anything that is required for initializing itself.
        self.is_scattered = False
        self._shadows = {}  # bf16 shadows of parameters (bf16 mode only)

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        for name in ("forward", "backward"):
            fn = cls.__dict__.get(name)
            if fn is not None and not getattr(fn, "_npt_boundary", False):
                setattr(cls, name, host_boundary(fn))

    @abstractmethod
    def forward(self):
        ...

    def backward(self, d_out):
        return d_out

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def add_parameter(self, name, data, shard_axis: int | None = None) -> None:
        """Add a parameter: a device buffer that also has a gradient."""
        if isinstance(data, np.ndarray):
            data = _to_f32_tensor(data)
        if shard_axis is not None:
            assert shard_axis <= len(data.shape), f"Cannot shard {shard_axis} on {len(data.shape)} dims."
        setattr(self, name, Parameter(data=data, gradient=None, shard_axis=shard_axis))
        self.parameters[name] = getattr(self, name)
        self._shadows.pop(name, None)
        bump_storage_generation()

    def add_setting(self, name, value):
        setattr(self, name, value)
        self.settings[name] = value

    def add_settings(self, settings: dict):
        for key, value in settings.items():
            self.add_setting(key, value)

    def update_parameter(self, name: str, **updates):
        current_parameter = getattr(self, name)
        for key in ("data", "gradient"):
            if isinstance(updates.get(key), np.ndarray):
                updates[key] = _to_f32_tensor(updates[key])
        if "data" in updates and updates["data"] is not current_parameter.data:
            self._shadows.pop(name, None)
            bump_storage_generation()
        setattr(self, name, replace(current_parameter, **updates))
        self.parameters[name] = getattr(self, name)

    # -- bf16 shadow of a 2-D parameter (rows x pad8(cols)), kept fresh by Adam ----------------
    def shadow(self, name: str) -> torch.Tensor:
        sh = self._shadows.get(name)
        if sh is None:
            from .. import ops

            sh = ops.cast_bf16(getattr(self, name).data, pad_cols=True)
            self._shadows[name] = sh
        return sh

    def to_dict(self):
        """Dictionary representation (NumPy arrays, the reference's checkpoint layout)."""
        parameters = {
            name: {"data": rt.to_numpy(p.data), "gradient": rt.to_numpy(p.gradient), "shard_axis": p.shard_axis}
            for name, p in ((n, getattr(self, n)) for n in self.parameters)
        }
        return {"settings": self.settings, "parameters": parameters}

    @classmethod
    def from_dict(cls, layer_dict: dict[str, dict]):
        settings, parameters = layer_dict["settings"], layer_dict["parameters"]
        settings |= {"weight_init": "zeros", "bias_init": "zeros"}
        layer = cls(**settings)
        for name, parameter in parameters.items():
            layer.add_parameter(name, parameter["data"], parameter["shard_axis"])
        return layer

    def scatter(self, src: int = 0):
        """Scatter this layer in place along each parameter's shard axis (np.array_split pieces)."""
        assert not self.is_scattered, "Cannot scatter an already scattered layer."
        tp, tp_rank = npdist.tensor_parallel_size(), npdist.tensor_parallel_rank()
        for name, parameter in self.parameters.items():
            if parameter.shard_axis is None:
                continue
            update = {}
            new_shape = list(parameter.data.shape)
            new_shape[parameter.shard_axis] = _split_sizes(new_shape[parameter.shard_axis], tp)[tp_rank]
            new_data = torch.zeros(new_shape, dtype=parameter.data.dtype, device=parameter.data.device)
            npdist.scatter(parameter.data, new_data, parameter.shard_axis, src, npdist.tensor_parallel_group())
            update["data"] = new_data
            if parameter.gradient is not None:
                new_gradient = torch.zeros(new_shape, dtype=parameter.gradient.dtype, device=parameter.gradient.device)
                npdist.scatter(parameter.gradient, new_gradient, parameter.shard_axis, src,
                               npdist.tensor_parallel_group())
                update["gradient"] = new_gradient
            self.update_parameter(name, **update)
        self.is_scattered = True

    def all_gather(self):
        """All-gather this layer in place."""
        tp_group = npdist.tensor_parallel_group()
        for name, parameter in self.parameters.items():
            if parameter.shard_axis is None:
                continue
            update = {}
            new_shape = list(parameter.data.shape)
            shard_dim = torch.tensor([new_shape[parameter.shard_axis]], dtype=torch.int64,
                                     device=parameter.data.device)
            npdist.all_reduce(shard_dim, group=tp_group)
            new_shape[parameter.shard_axis] = int(shard_dim.item())
            new_data = torch.zeros(new_shape, dtype=parameter.data.dtype, device=parameter.data.device)
            npdist.all_gather(parameter.data, new_data, parameter.shard_axis, tp_group)
            update["data"] = new_data
            if parameter.gradient is not None:
                new_gradient = torch.zeros(new_shape, dtype=parameter.gradient.dtype, device=parameter.gradient.device)
                npdist.all_gather(parameter.gradient, new_gradient, parameter.shard_axis, tp_group)
                update["gradient"] = new_gradient
            self.update_parameter(name, **update)
        self.is_scattered = False


def _split_sizes(size: int, k: int):
    base, extra = divmod(size, k)
    return [base + 1] * extra + [base] * (k - extra)
