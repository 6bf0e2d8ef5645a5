"""tp x dp parity on real GPUs (one process per GPU, NCCL): skipped unless the box has enough
devices.  Run during development with `gpurun --gpus 2|4 -- python -m pytest tests/test_multigpu.py -m gpu`."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
WORKER = Path(__file__).resolve().parent / "_nccl_worker.py"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(worker, args, world, timeout=150, extra_env=None):
    """One process per GPU; fail fast: as soon as one rank dies the others are stuck in a collective -> kill them."""
    import time

    def once():
        port = _free_port()
        procs = []
        for r in range(world):
            env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port), **(extra_env or {}))
            procs.append(subprocess.Popen([sys.executable, str(worker), *map(str, args)], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        deadline = time.time() + timeout
        failed = None
        while time.time() < deadline:
            codes = [p.poll() for p in procs]
            if any(c not in (None, 0) for c in codes):
                failed = [i for i, c in enumerate(codes) if c not in (None, 0)][0]
                break
            if all(c == 0 for c in codes):
                break
            time.sleep(0.5)
        for p in procs:
            if p.poll() is None:
                p.kill()
        return procs, [p.communicate()[0] for p in procs], failed

    procs, outs, failed = once()
    if failed is not None and "EADDRINUSE" in outs[failed]:  # the probed rendezvous port was taken meanwhile
        procs, outs, failed = once()
    assert failed is None, f"rank {failed} failed:\n{outs[failed][-3000:]}"
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} did not finish:\n{out[-3000:]}"
        assert f"rank {r} ok" in out
    return outs


def _need_gpus(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2)])
def test_tiny_train_steps_vs_reference_world(tp, dp, precision):
    _need_gpus(tp * dp)
    _run_world(WORKER, [tp, dp, precision], tp * dp)


CONFIG_WORKER = Path(__file__).resolve().parent / "_nccl_config_worker.py"
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("precision,rowsplit", [("bf16", "1"), ("bf16", "0"), ("fp32", "1")])
@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2)])
def test_c1_shape_vs_oracle_world(tp, dp, precision, rowsplit, tmp_path):
    """BASELINE configs[1] shape (L6 d384 H6 ctx256 vocab65) under tp2 (x dp2): 3 training steps, every rank
    against the oracle's emulated world, computed here (a few seconds).  In bf16 mode this is the path
    SCALE times: GEMM-epilogue peer push + device barrier + reduction inside LayerNorm — with the rows of every
    LayerNorm split between the two ranks (rowsplit 1: reduce-scatter / all-gather by push, distributed/peer.py)
    or both ranks normalising all rows (rowsplit 0)."""
    _need_gpus(tp * dp)
    sys.path.insert(0, str(GOLDEN))
    import make_config_golden as mk

    vec = tmp_path / "config_c1.npz"
    mk.main("c1", spec=dict(tp=tp, dp=dp, global_batch=2 * dp, steps=3), out_path=vec)
    outs = _run_world(CONFIG_WORKER, ["c1", vec, precision], tp * dp, timeout=240, extra_env={"NUMPITRON_TP_ROWSPLIT": rowsplit})
    if precision == "bf16":
        assert "path peer-push" in outs[0] or os.environ.get("NUMPITRON_TP_PEER") == "0", outs[0][-500:]
        if rowsplit == "1" and os.environ.get("NUMPITRON_TP_PEER") != "0":
            assert "rowsplit 12" in outs[0], outs[0][-500:]   # every LayerNorm of the 6 blocks ran row-sharded


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("config", ["c3", "c4"])
def test_baseline_config_vs_oracle_world(config, precision):
    """BASELINE configs[2] (GPT-2 small, tp=8: one local head of dim 96, V_local 6283/6282) and configs[3]
    (GPT-2 medium, tp2 x dp4: the 102.9 MB data-parallel all-reduce of the tied embedding gradient) at a
    reduced batch against the committed oracle vectors (tests/golden/config_c*.npz)."""
    _need_gpus(8)
    _run_world(CONFIG_WORKER, [config, GOLDEN / f"config_{config}.npz", precision], 8, timeout=600)
