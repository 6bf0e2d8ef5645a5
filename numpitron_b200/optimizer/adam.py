"""Adam — drop-in for the reference's optimizer/adam.py (same constructor, step(), state tree).

One fused multi-tensor kernel launch updates every parameter that has a gradient.  The
reference's behaviour is kept exactly: `timestep` never advances, so the bias
corrections are the constants 1-beta; eps sits outside the sqrt; parameters whose gradient is
None are skipped; gradients are reset to None after the update.  The scalars are rounded to
float32 the way NumPy rounds Python scalars against float32 arrays, which makes the update
bit-identical to the reference's.
"""
from __future__ import annotations

import numpy as np

from .. import ops
from .. import runtime as rt
from ..nn.core import Layer, weight_init_fn
from ..nn.model import Model


class AdamParams(Layer):
    def __init__(self, shape: tuple[int, ...], shard_axis: int | None = None, **kwargs):
        super().__init__()
        self.add_parameter("momentum", weight_init_fn("zeros")(tuple(shape)), shard_axis)
        self.add_parameter("velocity", weight_init_fn("zeros")(tuple(shape)), shard_axis)


def _f32(x: float) -> float:
    return float(np.float32(x))


class Adam:
    def __init__(self, model: Model, learning_rate: float, beta1: float = 0.99, beta2: float = 0.999,
                 eps: float = 1e-7):
        self.model = model
        self.learning_rate = learning_rate
        self.beta1 = beta1
        self.beta2 = beta2
        self.eps = eps
        self.timestep = 1
        self._table = None          # the table eager steps recycle
        self._tables = {}           # pointer-key -> table (captured graphs keep theirs forever)
        self._spare_tables = []
        self.init()

    def init(self):
        self.layers = self.model.get_layers()  # references to each layer

        def _init_params(layer: dict) -> dict:
            params = {}
            for n, l in layer.items():
                if isinstance(l, Layer):
                    params[n] = {name: AdamParams(p.data.shape, p.shard_axis) for name, p in l.parameters.items()}
                elif isinstance(l, dict):
                    params[n] = _init_params(l)
            return params

        self.parameters = _init_params(self.layers)

    # -- one fused launch over the whole tree ---------------------------------------------------
    def _collect(self, layer, params, out):
        if isinstance(layer, Layer):
            for name, state in params.items():
                p = layer.parameters[name]
                if p.gradient is None:  # "possible footgun" of the reference: silently skipped
                    continue
                out.append((layer, name, state))
            return
        for key in params:
            self._collect(layer[key], params[key], out)

    def step(self, layer=None, params=None):
        if layer is None and params is None:
            layer, params = self.layers, self.parameters
        todo = []
        self._collect(layer, params, todo)
        if not todo:
            return
        entries = []
        for lyr, name, state in todo:
            p = lyr.parameters[name]
            g = p.gradient
            if not g.is_contiguous() and not hasattr(g, "_npt_split_stride"):   # (strided stacks of partials are read in place)
                g = g.contiguous()
            shadow = lyr._shadows.get(name) if rt.is_bf16() else None
            entries.append((p.data, g, state.momentum.data, state.velocity.data, shadow))
        # blocks are scheduled in table order: the tensors whose gradient is a deep stack of partials (LayerNorm /
        # bias gradients handed over un-reduced) take longest per element, so they go first and overlap the bulk
        entries.sort(key=lambda e: -getattr(e[1], "_npt_splits", 1))
        table = self._table_for(entries)
        t = self.timestep
        ops.adam_step(table,_f32(self.learning_rate), _f32(self.beta1), _f32(1 - self.beta1),
                      _f32(self.beta2), _f32(1 - self.beta2), _f32(1 - self.beta1**t), _f32(1 - self.beta2**t),
                      _f32(self.eps))
        for lyr, name, state in todo:
            # data / state were updated in place; the gradient is consumed.
            p = lyr.parameters[name]
            setattr(lyr, name, type(p)(data=p.data, gradient=None, shard_axis=p.shard_axis))
            lyr.parameters[name] = getattr(lyr, name)

    def _table_for(self, entries):
        """Pointer table for this set of buffers.  Eager steps recycle one table; every CUDA-graph
        capture takes its own pre-allocated table, which is never touched again (a replayed graph
        reads it)."""
        import torch

        # the gradient may be a stack of split-K partials at the address an earlier plain gradient had: the
        # split count and the element count are part of the identity of an entry
        key = tuple((rt.ptr(a), rt.ptr(b), rt.ptr(c), rt.ptr(d), rt.ptr(e), getattr(b, "_npt_splits", 1),
                     getattr(b, "_npt_split_stride", 0), a.numel()) for a, b, c, d, e in entries)
        cap = max(len(entries), self._count_parameters())
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            while len(self._spare_tables) < 3:  # pinned allocations are illegal during capture
                self._spare_tables.append(ops.AdamTable(cap))
        table = self._tables.get(key)
        if table is not None:
            return table
        if capturing:
            if not self._spare_tables:
                raise RuntimeError("Adam: run one eager step before capturing more CUDA graphs")
            table = self._spare_tables.pop()
        else:
            if self._table is None or self._table.capacity < len(entries):
                self._table = ops.AdamTable(cap)
            table = self._table
            self._tables = {k: v for k, v in self._tables.items() if v is not table}
        table.fill(entries)
        self._tables[key] = table
        return table

    def _count_parameters(self) -> int:
        def walk(params):
            return sum(1 if isinstance(v, AdamParams) else walk(v) for v in params.values())

        return walk(self.parameters)

    def scatter(self, params: dict = None):
        if params is None:
            params = self.parameters
        for key, value in params.items():
            if isinstance(value, AdamParams):
                value.scatter()
            else:
                self.scatter(value)
        self._drop_tables()

    def all_gather(self, params: dict = None):
        if params is None:
            params = self.parameters
        for key, value in params.items():
            if isinstance(value, AdamParams):
                value.all_gather()
            else:
                self.all_gather(params=value)
        self._drop_tables()

    def _drop_tables(self):
        """Parameter / moment storage was replaced: every pointer table is stale.  Captured graphs that
        still reference one are invalidated through nn.core.storage_generation() (GraphedTrainStep
        re-captures), so nothing keeps the old buffers alive."""
        self._table = None
        self._tables = {}

    def to_dict(self):
        def _to_dict(params) -> dict:
            out = {}
            for key, value in params.items():
                if isinstance(value, dict):
                    out[key] = _to_dict(value)
                elif isinstance(value, Layer):
                    out[key] = value.to_dict()
            return out

        return {"parameters": _to_dict(self.parameters), "timestep": self.timestep,
                "learning_rate": self.learning_rate, "beta1": self.beta1, "beta2": self.beta2, "eps": self.eps}

    @classmethod
    def from_dict(cls, state: dict, model: Model):
        parameters = state.pop("parameters")
        timestep = state.pop("timestep")
        optimizer = cls(model, **state)

        def recursive(something):
            out = {}
            for key, value in something.items():
                if "settings" in value and "parameters" in value:
                    momentum = value["parameters"]["momentum"]
                    velocity = value["parameters"]["velocity"]
                    out[key] = AdamParams(momentum["data"].shape, momentum["shard_axis"])
                    out[key].update_parameter("momentum", data=momentum["data"])
                    out[key].update_parameter("velocity", data=velocity["data"])
                else:
                    out[key] = recursive(value)
            return out

        optimizer.parameters = recursive(parameters)
        optimizer.timestep = timestep
        return optimizer
