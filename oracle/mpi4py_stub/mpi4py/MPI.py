"""Minimal `mpi4py.MPI` replacement for running the reference as an oracle.

TEST INFRASTRUCTURE ONLY.  Surface = what `/root/reference/src/numpitron/distributed/*.py`
uses: COMM_WORLD, Comm(comm), .group.Incl, .Create, Get_size/Get_rank,
Allreduce(IN_PLACE, buf, op), Scatterv, Allgatherv, Gatherv, Reduce, bcast,
Send/Recv, Alltoallv, and the constants IN_PLACE SUM MAX FLOAT UINT16_T.
"""
from __future__ import annotations

import os

import numpy as np

IN_PLACE = "IN_PLACE"
SUM = "SUM"
MAX = "MAX"
FLOAT = "FLOAT"
UINT16_T = "UINT16_T"
DOUBLE = "DOUBLE"

_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
_RANK = int(os.environ.get("RANK", "0"))
_dist = None


def _ensure_dist():
    """Lazily bring up gloo when more than one rank exists."""
    global _dist
    if _WORLD == 1:
        return None
    if _dist is None:
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            dist.init_process_group("gloo", rank=_RANK, world_size=_WORLD)
        _dist = dist
    return _dist


class Op(str):
    pass


class Group:
    def __init__(self, ranks):
        self.ranks = tuple(int(r) for r in ranks)

    def Incl(self, ranks):
        return Group([self.ranks[i] for i in ranks])

    def Get_size(self):
        return len(self.ranks)


class Intracomm:
    """A communicator = ordered tuple of world ranks (+ a gloo process group)."""

    def __init__(self, ranks, pg=None):
        self.ranks = tuple(int(r) for r in ranks)
        self._pg = pg

    # -- identity -----------------------------------------------------------------
    @property
    def group(self):
        return Group(self.ranks)

    def Get_size(self):
        return len(self.ranks)

    def Get_rank(self):
        return self.ranks.index(_RANK)

    def Create(self, group: Group):
        """Called by each rank with ITS OWN group only (distributed/__init__.py:125-131).

        torch.distributed.new_group must be entered by every rank for every group in
        the same order, so gather all rank tuples first and create them sorted.
        """
        dist = _ensure_dist()
        if dist is None:
            return Intracomm(group.ranks)
        mine = tuple(group.ranks)
        everyone = [None] * _WORLD
        dist.all_gather_object(everyone, mine)
        made = None
        for ranks in sorted(set(everyone)):
            pg = dist.new_group(list(ranks))
            if ranks == mine:
                made = Intracomm(ranks, pg)
        return made

    # -- helpers --------------------------------------------------------------------
    def _torch(self, arr):
        import torch

        a = np.ascontiguousarray(arr)
        if a.dtype == np.uint16:  # gloo has no uint16 collectives
            return torch.from_numpy(a.astype(np.int32))
        return torch.from_numpy(a)

    # -- collectives ------------------------------------------------------------------
    def Allreduce(self, sendbuf, recvbuf, op=SUM):
        assert sendbuf is IN_PLACE or sendbuf == IN_PLACE
        if self.Get_size() == 1:
            return
        dist = _ensure_dist()
        t = self._torch(recvbuf)
        ropt = dist.ReduceOp.SUM if op == SUM else dist.ReduceOp.MAX
        dist.all_reduce(t, op=ropt, group=self._pg)
        np.copyto(recvbuf, t.numpy().astype(recvbuf.dtype).reshape(recvbuf.shape))

    def Scatterv(self, sendspec, recvbuf, root=0):
        buf, counts, displs, _ = sendspec
        me = self.Get_rank()
        if self.Get_size() == 1:
            c, d = int(counts[0]), int(displs[0])
            np.copyto(recvbuf, np.asarray(buf).reshape(-1)[d : d + c].reshape(recvbuf.shape))
            return
        dist = _ensure_dist()
        import torch

        wide = recvbuf.dtype == np.uint16
        dt = np.int32 if wide else recvbuf.dtype
        out = torch.from_numpy(np.zeros(int(counts[me]), dtype=dt))
        scatter_list = None
        if me == root:  # only the root's data may be used
            flat = np.asarray(buf).reshape(-1).astype(dt)
            scatter_list = [
                torch.from_numpy(np.ascontiguousarray(flat[int(d) : int(d) + int(c)]))
                for c, d in zip(counts, displs)
            ]
        # gloo scatter needs equal sizes → send/recv pairs instead.
        if me == root:
            for r, piece in enumerate(scatter_list):
                if r == me:
                    out.copy_(piece)
                else:
                    dist.send(piece, dst=self.ranks[r], group=self._pg)
        else:
            dist.recv(out, src=self.ranks[root], group=self._pg)
        np.copyto(recvbuf, out.numpy().astype(recvbuf.dtype).reshape(recvbuf.shape))

    def Allgatherv(self, sendbuf, recvspec):
        recv, counts, displs, _ = recvspec
        me = self.Get_rank()
        flat = np.ascontiguousarray(sendbuf).reshape(-1)
        if self.Get_size() == 1:
            recv[int(displs[0]) : int(displs[0]) + int(counts[0])] = flat
            return
        dist = _ensure_dist()
        import torch

        for r in range(self.Get_size()):
            c, d = int(counts[r]), int(displs[r])
            piece = torch.from_numpy(flat.copy()) if r == me else torch.from_numpy(
                np.zeros(c, dtype=flat.dtype))
            dist.broadcast(piece, src=self.ranks[r], group=self._pg)
            recv[d : d + c] = piece.numpy()

    def Gatherv(self, sendbuf, recvspec, root=0):
        me = self.Get_rank()
        if recvspec is None:
            recvspec = [np.empty(0, dtype=np.asarray(sendbuf).dtype), [0] * self.Get_size(),
                        [0] * self.Get_size(), None]
        recv, counts, displs, _ = recvspec
        flat = np.ascontiguousarray(sendbuf).reshape(-1)
        if self.Get_size() == 1:
            recv[int(displs[0]) : int(displs[0]) + int(counts[0])] = flat
            return
        dist = _ensure_dist()
        import torch

        if me == root:
            for r in range(self.Get_size()):
                c, d = int(counts[r]), int(displs[r])
                if r == me:
                    recv[d : d + c] = flat
                else:
                    piece = torch.from_numpy(np.zeros(c, dtype=flat.dtype))
                    dist.recv(piece, src=self.ranks[r], group=self._pg)
                    recv[d : d + c] = piece.numpy()
        else:
            dist.send(torch.from_numpy(flat.copy()), dst=self.ranks[root], group=self._pg)

    def Reduce(self, sendbuf, recvbuf, op=SUM, root=0):
        if self.Get_size() == 1:
            if not (sendbuf is IN_PLACE or (isinstance(sendbuf, str) and sendbuf == IN_PLACE)):
                np.copyto(recvbuf, sendbuf)
            return
        dist = _ensure_dist()
        src = recvbuf if (isinstance(sendbuf, str) and sendbuf == IN_PLACE) else sendbuf
        t = self._torch(src).clone()
        ropt = dist.ReduceOp.SUM if op == SUM else dist.ReduceOp.MAX
        dist.reduce(t, dst=self.ranks[root], op=ropt, group=self._pg)
        if self.Get_rank() == root and recvbuf is not None:
            np.copyto(recvbuf, t.numpy().astype(recvbuf.dtype).reshape(recvbuf.shape))

    def bcast(self, obj, root=0):
        if self.Get_size() == 1:
            return obj
        dist = _ensure_dist()
        box = [obj]
        dist.broadcast_object_list(box, src=self.ranks[root], group=self._pg)
        return box[0]

    def Bcast(self, buf, root=0):
        if self.Get_size() == 1:
            return
        dist = _ensure_dist()
        t = self._torch(buf)
        dist.broadcast(t, src=self.ranks[root], group=self._pg)
        np.copyto(buf, t.numpy().astype(buf.dtype).reshape(buf.shape))

    def Send(self, buf, dest=0, tag=0):
        dist = _ensure_dist()
        if dist is None:
            raise RuntimeError("Send needs a peer (world size 1)")
        dist.send(self._torch(buf), dst=self.ranks[dest], group=self._pg)

    def Recv(self, buf, source=0, tag=0):
        dist = _ensure_dist()
        if dist is None:
            raise RuntimeError("Recv needs a peer (world size 1)")
        t = self._torch(buf)
        dist.recv(t, src=self.ranks[source], group=self._pg)
        np.copyto(buf, t.numpy().astype(buf.dtype).reshape(buf.shape))

    def Barrier(self):
        dist = _ensure_dist()
        if dist is not None:
            dist.barrier(group=self._pg)


def Comm(comm: Intracomm) -> Intracomm:
    return comm


class _World(Intracomm):
    def __init__(self):
        super().__init__(range(_WORLD), None)

    def Allreduce(self, *a, **k):
        _ensure_dist()
        return super().Allreduce(*a, **k)


COMM_WORLD = _World()
