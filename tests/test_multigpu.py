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


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tp,dp", [(2, 1), (2, 2)])
def test_tiny_train_steps_vs_reference_world(tp, dp, precision):
    world = tp * dp
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import time

    def run_world():
        port = _free_port()
        procs = []
        for r in range(world):
            env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, str(WORKER), str(tp), str(dp), precision], env=env,
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        # fail fast: as soon as one rank dies the others are stuck in a collective -> kill them
        deadline = time.time() + 150
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

    procs, outs, failed = run_world()
    if failed is not None and "EADDRINUSE" in outs[failed]:  # the probed rendezvous port was taken meanwhile
        procs, outs, failed = run_world()
    assert failed is None, f"rank {failed} failed:\n{outs[failed][-3000:]}"
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} did not finish:\n{out[-3000:]}"
        assert f"rank {r} ok" in out
