"""numpitron_b200 — NuMPItron's transformer training step on B200 (sm_100a).

Same layer API as lweitkamp/numpitron (`nn`, `optimizer`, `distributed`); every
forward/backward/step body is a ctypes call into hand-written CUDA (csrc/, C ABI in
include/numpitron_b200.h).  There is no CPU fallback.
"""
# flake8: noqa
from . import runtime
from .runtime import precision, set_precision
from . import distributed
from . import nn
from . import optimizer
from .train import train_step, validation_step

__all__ = ["nn", "optimizer", "distributed", "runtime", "set_precision", "precision", "train_step",
           "validation_step"]
