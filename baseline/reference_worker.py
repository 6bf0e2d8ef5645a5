"""One rank of the reference arm (`bench.py --impl reference`): the UNMODIFIED reference package from
baseline/_ref (pip-installed from /root/reference by baseline/install_reference.sh) running its own
training step on the host CPU, one OS process per rank, joined by oracle/mpi4py_stub (gloo) because the
image has neither mpi4py nor mpirun.  Nothing of numpitron_b200's compute is on this path: model, layers,
loss, optimizer, collectives and data loader are the reference's.

The loop body is the reference's `train_step` (train_shakespeare.py:14-21 — the script is not part of the
installable package, so its five lines are restated here) inside the epoch loop of train_shakespeare.py:
98-106 (DataLoader.iter_epoch over a synthetic uint16 token file).

    reference_worker.py <json spec>     (RANK / WORLD_SIZE / MASTER_* from the environment)
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle" / "mpi4py_stub"))
sys.path.insert(0, str(ROOT / "baseline" / "_ref"))

import numpy as np  # noqa: E402


def main(spec):
    from numpitron import distributed as npdist
    from numpitron.data import DataLoader
    from numpitron.nn import Transformer, softmax_cross_entropy
    from numpitron.optimizer import Adam

    def train_step(transformer, optimizer, x, y):       # train_shakespeare.py:14-21
        y_hat = transformer(x)
        loss, d_out = softmax_cross_entropy(y_hat, y)
        transformer.backward(d_out)
        optimizer.step()
        return loss.mean()

    npdist.init(tp_size=spec["tp"], dp_size=spec["dp"])                      # train_shakespeare.py:45-47
    rng = np.random.default_rng(42)
    transformer = Transformer(**spec["model"], rng=rng)                       # :52-61
    optimizer = Adam(model=transformer, learning_rate=spec["lr"], beta1=spec["beta1"], beta2=spec["beta2"])
    loader = DataLoader(spec["tokens_path"], spec["model"]["seq_len"], spec["global_batch"], rng)
    transformer.scatter()                                                     # :95-96
    optimizer.scatter()
    times, losses = [], []
    t_start = time.perf_counter()
    for i, (x, y) in enumerate(loader.iter_epoch()):
        t0 = time.perf_counter()
        loss = train_step(transformer, optimizer, x, y)
        dt = time.perf_counter() - t0
        if i >= spec["warmup"]:
            times.append(dt)
            losses.append(float(loss))
        # every rank takes the same decision: the clock is rank 0's, shared through an all-reduce (max)
        stop = np.array([1.0 if (len(times) >= spec["steps"] or
                                 (times and time.perf_counter() - t_start > spec["budget_s"])) else 0.0], dtype=np.float32)
        npdist.all_reduce(stop, op="max")
        if stop[0] > 0:
            break
    print("REFERENCE_RESULT " + json.dumps({"rank": npdist.world_rank(), "times": times, "losses": losses}), flush=True)


if __name__ == "__main__":
    main(json.loads(sys.argv[1]))
