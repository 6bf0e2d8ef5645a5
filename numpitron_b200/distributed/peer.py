"""Peer-memory exchange for tensor parallelism over two GPUs.

The reference all-reduces the output of every row-parallel layer (forward: nn/attention.py:57-58,
nn/mlp.py:35-36) and the dX of every column-parallel layer (backward: attention.py:101-102,
mlp.py:52-53); in a TransformerBlock each of those tensors is consumed by exactly one kernel, the
LayerNorm forward / backward that follows.  With tp = 2 on one NVLink box the all-reduce kernel is
dropped: the GEMM epilogue TMA-stores every tile of its partial twice — into its own buffer and into
a buffer in the OTHER rank's memory (symmetric memory, NVLink P2P; `npt_gemm_args.C_bf16_peer`), so the
exchange rides on posted NVLink writes while the GEMM is still computing.  The two ranks then meet at
a device-side barrier and the LayerNorm kernel reads `round_bf16(mine + peer's)` from local memory
(`npt_layernorm_*_v2`, x_peer / dy_peer).  That replaces a 71 us NCCL all-reduce of 25 MiB (measured,
369 GB/s) plus LayerNorm's re-read of the result.  (A first version let LayerNorm *read* the peer's
buffer over NVLink: P2P loads are latency-bound, the kernel got several times slower and most of the
gain was lost.)

Slots: per direction two "own" and two "receive" slots, used alternately.  Rank A's GEMM writes B's
receive slot s for exchange i; B reads it after barrier i and A writes it again two exchanges later,
after barrier i+1, which B only reaches once that read (earlier in B's stream) is done — so one
barrier per exchange suffices.

Row-sharded LayerNorm (NUMPITRON_TP_ROWSPLIT=0 turns it off).  Reducing on both ranks means both ranks also
normalise every row: LayerNorm, which is memory-bound, costs twice what it costs on one GPU.  When the rows
split into two multiples of 128, the GEMM stores each tile ONCE — tiles of the rows this rank owns locally, the
others into the peer's receive slot (`npt_gemm_args.peer_rows_*`: a reduce-scatter by push) — LayerNorm forward /
backward run on the own half of the rows only and write the bf16 tensor the next GEMMs read (y, or dx) into a
peer-visible ARENA buffer and into the peer's copy of it (an all-gather by push), followed by a second barrier.
Same NVLink bytes as before, half the LayerNorm work and half the epilogue stores.  The per-block partials of
dgamma / dbeta / bias gradient are pushed the same way; both ranks add the two stacks in the same order, so the
replicated parameters stay bit-identical.  The arena is a bump allocator that `step_begin()` rewinds at the
start of every forward pass: the same sequence of requests gets the same addresses every step (CUDA graphs).

NUMPITRON_TP_PEER=0 disables all of this and keeps the NCCL all-reduces.
"""
from __future__ import annotations

import os
import sys

import torch

from .. import runtime as rt

_state = {"px": None, "failed": False, "arena_bytes": 0}


class PeerExchange:
    def __init__(self, pg, slot_bytes: int, arena_bytes: int = 0):
        import torch.distributed._symmetric_memory as symm

        self.slot_bytes = (slot_bytes + 255) // 256 * 256
        self.arena_bytes = (arena_bytes + 255) // 256 * 256
        self.arena_base, self.arena_pos = 8 * self.slot_bytes, 0
        self.flag_off = 8 * self.slot_bytes + self.arena_bytes
        self.buf = symm.empty(self.flag_off + 256, dtype=torch.uint8, device=rt.device())
        self.hdl = symm.rendezvous(self.buf, pg)
        assert self.hdl.world_size == 2
        self.rank = self.hdl.rank
        self.peer_base = int(self.hdl.buffer_ptrs[1 - self.rank])
        self.counters = {"fwd": 0, "bwd": 0}
        # own device-side barrier (npt_peer_barrier): one flag word per rank, written by the peer; epoch counter local
        self.buf[self.flag_off:].zero_()
        self.epoch = torch.zeros(1, dtype=torch.int64, device=rt.device())
        torch.cuda.synchronize()
        self.hdl.barrier(channel=0)      # both flag words are zero before anybody signals
        torch.cuda.synchronize()
        self.own_barrier = os.environ.get("NUMPITRON_TP_BARRIER", "torch") == "npt"   # measured: 3.89 ms (npt) vs 3.83 ms (torch) per N=2 step

    def slot(self, direction: str, shape, dtype=torch.bfloat16):
        """Next exchange of `direction`: (own: local tensor the GEMM writes its partial to,
        send_ptr: address in the PEER's memory where the same partial must also be stored,
        recv_ptr: address in LOCAL memory where the peer stores its partial)."""
        i = (0 if direction == "fwd" else 4) + (self.counters[direction] & 1)
        self.counters[direction] += 1
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        assert nbytes <= self.slot_bytes, "peer slot too small"
        own_off, recv_off = i * self.slot_bytes, (i + 2) * self.slot_bytes
        own = self.buf[own_off:own_off + nbytes].view(dtype).view(*shape)
        return own, self.peer_base + recv_off, self.buf.data_ptr() + recv_off

    # -- row-sharded LayerNorm ------------------------------------------------------------------------
    def step_begin(self):
        self.arena_pos = 0

    def rows(self, n_rows: int):
        """(first row, row count) this rank owns, or None when the rows do not split into two multiples of 128
        (the GEMM routes whole 128-row tiles) or no arena was configured."""
        half = n_rows // 2
        if self.arena_bytes == 0 or n_rows % 2 or half % 128 or os.environ.get("NUMPITRON_TP_ROWSPLIT", "1") == "0":
            return None
        return self.rank * half, half

    def peer_rows(self, n_rows: int):
        half = n_rows // 2
        return (1 - self.rank) * half, (2 - self.rank) * half

    def arena(self, shape, dtype):
        """Next peer-visible buffer of the step: (local tensor, address of the same buffer in the peer's memory)."""
        n = 1
        for k in shape:
            n *= int(k)
        nbytes = (n * torch.empty((), dtype=dtype).element_size() + 255) // 256 * 256
        assert self.arena_pos + nbytes <= self.arena_bytes, "peer arena too small (peer.prepare sizes it from the model)"
        off = self.arena_base + self.arena_pos
        self.arena_pos += nbytes
        return self.buf[off:off + nbytes].view(dtype)[:n].view(*shape), self.peer_base + off

    def barrier(self):
        """Device-side barrier of the two ranks on the current stream: both partials are complete."""
        if self.own_barrier:
            from .. import _lib

            rc = _lib.load().npt_peer_barrier(self.peer_base + self.flag_off, self.buf.data_ptr() + self.flag_off,
                                              self.epoch.data_ptr(), rt.stream())
            if rc:
                raise _lib.NptError("npt_peer_barrier failed")
        else:
            self.hdl.barrier(channel=0)


def enabled() -> bool:
    from . import tensor_parallel_size

    return (not _state["failed"] and os.environ.get("NUMPITRON_TP_PEER", "1") != "0" and rt.is_bf16()
            and not rt.ln_inputs_fp32() and tensor_parallel_size() == 2 and torch.cuda.is_available())


def exchange(shape, dtype=torch.bfloat16):
    """The process-wide PeerExchange, or None when unavailable.  Created on first use and re-created
    (collectively: every tensor-parallel rank runs the same layer code on the same shapes, so all of them
    get here in the same call) when a tensor larger than its slots comes along — a bigger batch, a
    validation shape.  Whether the exchange is usable is agreed on by the whole group: a rank whose
    allocation or rendezvous failed must not fall back to NCCL alone while its peer waits at the barrier."""
    if not enabled():
        return None
    n = 1
    for s in shape:
        n *= int(s)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    px = _state["px"]
    if px is not None and nbytes <= px.slot_bytes and px.arena_bytes >= _state["arena_bytes"]:
        return px
    import torch.distributed as dist

    from . import tensor_parallel_group

    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("peer exchange would have to grow during CUDA-graph capture; run a warm-up step first")
    pg = tensor_parallel_group().pg
    new, err = None, None
    try:
        torch.cuda.synchronize()   # nothing still reads the slots that are about to be released
        new = PeerExchange(pg, max(nbytes, px.slot_bytes if px is not None else 0), _state["arena_bytes"])
    except Exception as e:  # no P2P / symmetric memory on this system
        err = e
    ok = torch.tensor([0 if new is None else 1], dtype=torch.int32, device=rt.device())
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=pg)
    if int(ok.item()) == 0:
        _state["failed"], _state["px"] = True, None
        print(f"[numpitron_b200] peer exchange unavailable on at least one rank ({type(err).__name__ if err else 'peer'}: "
              f"{err}); using NCCL all-reduce", file=sys.stderr, flush=True)   # stderr: bench.py's stdout is one JSON line
        return None
    _state["px"] = new
    return new


def prepare(n_rows: int, d_model: int, n_norms: int):
    """Called at the start of a forward pass of a model with `n_norms` LayerNorms over (n_rows, d_model)
    activations: sizes the arena for one step of row-sharded LayerNorm (forward: one bf16 tensor per norm, kept
    for the backward pass; backward: one bf16 tensor and one stack of partials per norm) and rewinds it."""
    if not enabled():
        return None
    act = (n_rows * d_model * 2 + 255) // 256 * 256
    part = (2 * 2 * 148 * 2 * 3 * d_model * 4 + 255) // 256 * 256   # [2 ranks][<= 2 blocks per SM, margin x2][3][d] fp32
    need = n_norms * (2 * act + part) + (1 << 20)
    if need > _state["arena_bytes"]:
        _state["arena_bytes"] = need
    px = exchange((n_rows, d_model))
    if px is not None:
        px.step_begin()
    return px


def active() -> bool:
    """True when the tensor-parallel exchange really runs through peer memory (an exchange exists)."""
    return _state["px"] is not None


def reset():
    _state["px"] = None
    _state["failed"] = False
