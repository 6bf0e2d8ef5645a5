"""Collectives with the reference's call signatures, carried by torch.distributed (NCCL/gloo)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .. import runtime as rt

_OPS = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX}


def _default_group():
    from . import world_group

    return world_group()


def _backend_is_nccl() -> bool:
    return dist.is_initialized() and dist.get_backend() == "nccl"


def _stage(t):
    """Return (tensor usable by the backend, copy-back function)."""
    if isinstance(t, np.ndarray):
        host = torch.from_numpy(t)
        if _backend_is_nccl():
            dev = host.to(rt.device())
            return dev, lambda: host.copy_(dev.cpu())
        return host, lambda: None
    if _backend_is_nccl() and not t.is_cuda:
        dev = t.to(rt.device())
        return dev, lambda: t.copy_(dev.cpu())
    return t, lambda: None


def all_reduce(tensor, op: str = "sum", group=None) -> None:
    """In-place all-reduce, op in {"sum","max"} (reference: distributed/all_reduce.py:5-23).

    This is the only collective on the training-step path (call sites C1-C10 of SURVEY §2.1)."""
    group = group or _default_group()
    if group.Get_size() == 1:
        return None
    buf, back = _stage(tensor)
    dist.all_reduce(buf, op=_OPS[op], group=group.pg)
    back()
    if isinstance(tensor, torch.Tensor):
        rt.drop_bf16(tensor)
    return None


class _Done:
    def wait(self):
        return None


def all_reduce_async(tensor: torch.Tensor, op: str = "sum", group=None):
    """Start an in-place all-reduce of a device buffer and return a handle with `.wait()`.

    Extension used inside the layers (not part of the reference API): the backward TP all-reduce of
    dX (call sites C4/C5) is started as soon as dX exists and overlaps the dW / db kernels, which do
    not depend on it.  `.wait()` orders the current stream after the collective (no host sync)."""
    group = group or _default_group()
    if group.Get_size() == 1:
        return _Done()
    work = dist.all_reduce(tensor, op=_OPS[op], group=group.pg, async_op=True)
    rt.drop_bf16(tensor)
    return work


def all_reduce_bucket_async(tensors, op: str = "sum", group=None):
    """One bucket = one coalesced all-reduce of several device buffers (a single ncclGroup: one launch, not one
    per tensor), started asynchronously; returns a handle with `.wait()`.  Used by the labelled non-reference
    data-parallel mode `dp_sync="all"` (nn/transformer.py): one bucket per TransformerBlock, issued as soon as
    the block's backward is enqueued, so it overlaps the backward of the blocks below it."""
    group = group or _default_group()
    tensors = [t for t in tensors if t is not None]
    if group.Get_size() == 1 or not tensors:
        return _Done()
    from torch.distributed.distributed_c10d import _coalescing_manager

    with _coalescing_manager(group=group.pg, async_ops=True) as cm:
        for t in tensors:
            dist.all_reduce(t, op=_OPS[op], group=group.pg)
    for t in tensors:
        rt.drop_bf16(t)
    return cm


def broadcast(tensor, src: int = 0, group=None) -> None:
    group = group or _default_group()
    if group.Get_size() == 1:
        return None
    buf, back = _stage(tensor)
    dist.broadcast(buf, src=group.ranks[src], group=group.pg)
    back()
    return None


def scatter(source_tensor, destination_tensor, axis: int, src: int = 0, group=None) -> None:
    """Split `source_tensor` (taken from rank `src` of the group) along `axis` with np.array_split
    semantics and place this rank's piece in `destination_tensor` (reference: distributed/scatter.py)."""
    group = group or _default_group()
    n, me = group.Get_size(), group.Get_rank()
    is_np = isinstance(source_tensor, np.ndarray)
    split = np.array_split if is_np else (lambda a, k, axis: torch.tensor_split(a, _split_points(a.shape[axis], k), dim=axis))
    if n == 1:
        piece = split(source_tensor, 1, axis=axis)[0]
        if is_np:
            np.copyto(destination_tensor, piece)
        else:
            destination_tensor.copy_(piece)
        return None
    # Every rank holds a same-shaped source; only src's values may be used: broadcast src's
    # copy, then slice locally (equivalent to Scatterv, uses only broadcast).
    full = source_tensor.copy() if is_np else source_tensor.clone()
    broadcast(full, src=src, group=group)
    piece = split(full, n, axis=axis)[me]
    if is_np:
        np.copyto(destination_tensor, piece)
    else:
        destination_tensor.copy_(piece)
    return None


def _split_points(size: int, k: int):
    """Indices at which np.array_split cuts an axis of length `size` into k pieces."""
    base, extra = divmod(size, k)
    sizes = [base + 1] * extra + [base] * (k - extra)
    return list(np.cumsum(sizes)[:-1])


def all_gather(source_tensor, destination_tensor, axis: int = -1, group=None) -> None:
    """Concatenate every rank's `source_tensor` along `axis` into `destination_tensor`
    (uneven np.array_split pieces allowed; reference: distributed/all_gather.py)."""
    group = group or _default_group()
    n = group.Get_size()
    is_np = isinstance(source_tensor, np.ndarray)
    if n == 1:
        if is_np:
            np.copyto(destination_tensor, source_tensor)
        else:
            destination_tensor.copy_(source_tensor)
        return None
    dst = torch.from_numpy(destination_tensor) if is_np else destination_tensor
    pieces = list(torch.tensor_split(dst, _split_points(dst.shape[axis], n), dim=axis))
    me = group.Get_rank()
    for r in range(n):
        buf = (torch.from_numpy(np.ascontiguousarray(source_tensor)) if is_np else source_tensor.contiguous()).clone() \
            if r == me else torch.empty(pieces[r].shape, dtype=dst.dtype, device=pieces[r].device)
        if _backend_is_nccl() and not buf.is_cuda:
            dev = buf.to(rt.device())
            dist.broadcast(dev, src=group.ranks[r], group=group.pg)
            buf = dev.cpu()
        else:
            dist.broadcast(buf, src=group.ranks[r], group=group.pg)
        pieces[r].copy_(buf)
    return None
