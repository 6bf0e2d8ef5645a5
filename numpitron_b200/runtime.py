"""Device plumbing: buffers, stream, precision mode, scratch space.

PyTorch is used ONLY as the owner of device memory / streams / process groups.  Every
arithmetic operation of the training step goes through `numpitron_b200._lib` into the
hand-written sm_100a kernels; there is no CPU or torch fallback for compute.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib

FP32 = "fp32"   # fp32-accurate mode: 3xTF32 on tcgen05 (NPT_GEMM_TF32X3_TC); NUMPITRON_FP32_GEMM=simt -> fp32 FFMA
BF16 = "bf16"   # bf16 operands on tcgen05, fp32 accumulate / residual stream / norms / loss / Adam
_GEMM_MODE = {FP32: 0 if os.environ.get("NUMPITRON_FP32_GEMM", "tf32x3") == "simt" else 2, BF16: 1}

_state = {"precision": os.environ.get("NUMPITRON_PRECISION", FP32), "workspace": None, "device": None}


# ---- precision -----------------------------------------------------------------------------
def set_precision(mode: str) -> None:
    if mode not in _GEMM_MODE:
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {mode!r}")
    _state["precision"] = mode


def precision() -> str:
    return _state["precision"]


def gemm_mode() -> int:
    return _GEMM_MODE[_state["precision"]]


def is_bf16() -> bool:
    return _state["precision"] == BF16


def ln_inputs_fp32() -> bool:
    """NUMPITRON_LN_INPUT=fp32 (bf16 mode, an accuracy / speed A-B knob): the two tensors per sub-layer that sit between
    a GEMM and a LayerNorm — the sub-layer output it normalises and the dX that reaches its backward — stay fp32
    instead of leaving the GEMM epilogue as bf16.  Their rows ride on a mean 30-70x their std, so bf16
    storage costs ~10 % of the normalised value per element; fp32 storage costs 12.6 MB more to write and to read
    per tensor at the bench shape.  Off by default; not available with the tp = 2 peer exchange (which moves bf16)."""
    return os.environ.get("NUMPITRON_LN_INPUT", "bf16") == "fp32"


# ---- device ---------------------------------------------------------------------------------
def device() -> torch.device:
    """cuda:<LOCAL_RANK> when a GPU is present, else cpu (host-side setup/tests only)."""
    if _state["device"] is None:
        if torch.cuda.is_available():
            idx = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
            torch.cuda.set_device(idx)
            _state["device"] = torch.device("cuda", idx)
        else:
            _state["device"] = torch.device("cpu")
    return _state["device"]


def has_gpu() -> bool:
    return device().type == "cuda"


def require_gpu(what: str) -> None:
    if not has_gpu():
        raise _lib.NptError(
            f"{what}: numpitron_b200 computes only on a CUDA device (sm_100a kernels); "
            "there is no CPU fallback.")


def stream() -> int:
    """Raw cudaStream_t of torch's current stream (kernels are enqueued there)."""
    return torch.cuda.current_stream().cuda_stream


# ---- buffers ----------------------------------------------------------------------------------
def empty(shape, dtype=torch.float32) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=torch.float32) -> torch.Tensor:
    return torch.zeros(shape, dtype=dtype, device=device())


def from_numpy(a: np.ndarray) -> torch.Tensor:
    """Host array -> device buffer (no dtype change; uint16 stays uint16)."""
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.copy() if not a.flags.writeable else a).to(device())


def to_numpy(t) -> np.ndarray:
    if t is None or isinstance(t, np.ndarray):
        return t
    if t.dtype == torch.bfloat16:  # bf16-only activations / gradients of the bf16 mode: NumPy has no bf16
        t = t.float()
    return t.detach().cpu().numpy()


def as_device(x, dtype=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = from_numpy(x)
    if x.device != device():
        x = x.to(device())
    if dtype is not None and x.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {x.dtype}")
    return x


def ptr(t, offset_elems: int = 0) -> int:
    if t is None:
        return 0
    return t.data_ptr() + offset_elems * t.element_size()


class deferred_wgrad:
    """Context manager: inside it, Linear.backward leaves the split-K partials of its weight-gradient GEMM
    un-reduced and hands them to Adam, which adds them while it reads (same additions, same order: the
    update is bit-identical).  Used by GraphedTrainStep, where nothing else looks at those gradients
    between backward and the optimizer step; `parameter.gradient` is then a (splits, d_in, d_out) tensor
    tagged `_npt_splits`."""

    def __enter__(self):
        self.prev = _state.get("defer_wgrad", False)
        _state["defer_wgrad"] = True
        return self

    def __exit__(self, *exc):
        _state["defer_wgrad"] = self.prev
        return False


def defer_wgrad() -> bool:
    # dp_sync="all" all-reduces every weight gradient over the data-parallel group: it needs the reduced
    # gradient, not the stack of split-K partials
    return bool(_state.get("defer_wgrad", False)) and _state.get("dp_sync", "reference") == "reference"


# ---- data-parallel gradient synchronisation mode ---------------------------------------------------------
def set_dp_sync(mode: str) -> None:
    """"reference" (default): only the tied embedding gradient is all-reduced over the data-parallel group,
    exactly what the reference does (nn/transformer.py:48-59, SURVEY Q11 — block weights diverge between
    replicas).  "all": a labelled NON-reference mode — every gradient is summed over the group, one bucket
    per TransformerBlock, overlapped with the backward pass of the blocks below."""
    if mode not in ("reference", "all"):
        raise ValueError(f"dp_sync must be 'reference' or 'all', got {mode!r}")
    _state["dp_sync"] = mode


def dp_sync() -> str:
    return _state.get("dp_sync", "reference")


def workspace(nbytes: int) -> torch.Tensor:
    """A scratch buffer of at least `nbytes` (stream-ordered reuse; grown on demand)."""
    nbytes = max(int(nbytes), 256)
    ws = _state["workspace"]
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.NptError("workspace would have to grow during CUDA-graph capture; run a warm-up step first")
        ws = torch.empty(max(nbytes, 64 << 20), dtype=torch.uint8, device=device())
        _state["workspace"] = ws
    return ws


# ---- bf16 shadows -----------------------------------------------------------------------------
def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def attach_bf16(t: torch.Tensor, shadow: torch.Tensor) -> None:
    t._npt_bf16 = shadow


def drop_bf16(t: torch.Tensor) -> None:
    if hasattr(t, "_npt_bf16"):
        del t._npt_bf16


def bf16_of(t: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a contiguous fp32 tensor whose last dim is a multiple of 8 (cached on it)."""
    sh = getattr(t, "_npt_bf16", None)
    if sh is None:
        from . import ops

        sh = ops.cast_bf16(t)
        t._npt_bf16 = sh
    return sh
