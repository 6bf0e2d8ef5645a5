"""tp x dp (x pp) world setup and collectives — the drop-in for `numpitron.distributed`.

Same surface and semantics as the reference (`distributed/__init__.py:35-166`,
`all_reduce.py:5-23`, `scatter.py`, `all_gather.py`), but one process per GPU and NCCL over
NVLink instead of mpi4py: `torch.distributed` is the plumbing (backend "nccl" when a GPU is
present, "gloo" for the CPU-side tests).  Rank layout is the reference's
`arange(world).reshape(pp, dp, tp)`: TP groups are consecutive ranks, DP groups are the ranks
with equal `rank % tp`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from .. import runtime as rt


class Comm:
    """A communicator: ordered tuple of world ranks (+ the torch process group carrying it)."""

    def __init__(self, ranks, pg=None):
        self.ranks = tuple(int(r) for r in ranks)
        self.pg = pg

    def Get_size(self) -> int:
        return len(self.ranks)

    def Get_rank(self) -> int:
        return self.ranks.index(_world_rank())

    def __repr__(self):
        return f"Comm(ranks={self.ranks})"


@dataclass
class ParallelState:
    """Holds all communication groups (reference: distributed/__init__.py:19-32)."""

    world_group: Comm = None
    tensor_parallel_group: Comm = None
    pipeline_parallel_group: Comm = None
    data_parallel_group: Comm = None
    model_parallel_group: Comm = None
    shared_variables: dict = field(default_factory=dict)


_GROUP_CACHE: dict = {}


def _env_world() -> tuple[int, int]:
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))


def _ensure_process_group() -> None:
    world, rank = _env_world()
    if world == 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if rt.has_gpu():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=rt.device())
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)


def _world_rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def _world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def _make_world() -> Comm:
    _ensure_process_group()
    return Comm(range(_world_size()), dist.group.WORLD if dist.is_initialized() else None)


PARALLEL_STATE = ParallelState()


def _state() -> ParallelState:
    if PARALLEL_STATE.world_group is None:
        PARALLEL_STATE.world_group = _make_world()
    return PARALLEL_STATE


def set_var(name: str, value) -> None:
    """Process-global shared variable (the vocab chunk bounds live here, embedding.py:24-25)."""
    _state().shared_variables[name] = value


def get_var(name: str):
    return _state().shared_variables.get(name, None)


def world_size() -> int:
    return _state().world_group.Get_size()


def world_rank() -> int:
    return _state().world_group.Get_rank()


def world_group() -> Comm:
    return _state().world_group


def _need(group, what):
    assert group is not None, f"{what} group is not initiated."
    return group


def _default_init() -> ParallelState:
    """The reference initialises the groups as a side effect of importing nn/transformer.py
    (`npdist.init(tp_size=npdist.world_size())`, nn/transformer.py:15), so its layers can be
    built without an explicit init().  Same default here, but lazily — on the first use of a tensor / data /
    pipeline parallel accessor — so that importing the package never opens a process group."""
    st = _state()
    if st.tensor_parallel_group is None:
        init(tp_size=world_size())
    return st


def tensor_parallel_size() -> int:
    return _need(_default_init().tensor_parallel_group, "Tensor Parallel").Get_size()


def tensor_parallel_rank() -> int:
    return _need(_default_init().tensor_parallel_group, "Tensor Parallel").Get_rank()


def tensor_parallel_group() -> Comm:
    return _need(_default_init().tensor_parallel_group, "Tensor Parallel")


def pipeline_parallel_size() -> int:
    return _need(_default_init().pipeline_parallel_group, "Pipeline Parallel").Get_size()


def pipeline_parallel_rank() -> int:
    return _need(_default_init().pipeline_parallel_group, "Pipeline Parallel").Get_rank()


def pipeline_parallel_group() -> Comm:
    return _need(_default_init().pipeline_parallel_group, "Pipeline Parallel")


def data_parallel_size() -> int:
    return _need(_default_init().data_parallel_group, "Data Parallel").Get_size()


def data_parallel_rank() -> int:
    return _need(_default_init().data_parallel_group, "Data Parallel").Get_rank()


def data_parallel_group() -> Comm:
    return _need(_default_init().data_parallel_group, "Data Parallel")


def add_group(all_groups_ranks) -> Comm | None:
    """Create EVERY group of a partition (all ranks must take part in each new_group call) and
    return the one containing this rank (reference: distributed/__init__.py:125-131)."""
    mine = None
    me = _world_rank()
    for group_ranks in all_groups_ranks:
        ranks = tuple(sorted(int(r) for r in group_ranks))
        if ranks not in _GROUP_CACHE:
            if not dist.is_initialized() or len(ranks) == _world_size():
                pg = dist.group.WORLD if dist.is_initialized() else None
            elif len(ranks) == 1:
                pg = None  # size-1 groups never communicate
            else:
                pg = dist.new_group(list(ranks))
            _GROUP_CACHE[ranks] = pg
        if me in ranks:
            mine = Comm(ranks, _GROUP_CACHE[ranks])
    return mine


def init(tp_size: int = 1, pp_size: int = 1, dp_size: int = 1) -> None:
    """Build the tensor / pipeline / data parallel groups (reference: distributed/__init__.py:134-166).

    May be called repeatedly (the reference re-inits after import)."""
    st = _state()
    assert tp_size * pp_size * dp_size == world_size(), (
        "The total world size must match the product of parallelization "
        f"strategies. {world_size()=}, {tp_size=}, {pp_size=}, {dp_size=}"
    )
    ranks = np.arange(0, world_size()).reshape(pp_size, dp_size, tp_size)
    st.tensor_parallel_group = add_group(ranks.transpose((0, 1, 2)).reshape(-1, tp_size))
    st.pipeline_parallel_group = add_group(ranks.transpose((1, 2, 0)).reshape(-1, pp_size))
    st.data_parallel_group = add_group(ranks.transpose((2, 0, 1)).reshape(-1, dp_size))
    st.model_parallel_group = add_group([ranks[:, dp_rank, :].reshape(-1) for dp_rank in range(dp_size)])


from .collectives import all_gather, all_reduce, all_reduce_async, all_reduce_bucket_async, broadcast, scatter  # noqa: E402,F401
