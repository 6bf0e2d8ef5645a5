"""One rank (one GPU) of the BASELINE-config parity tests: the model of a workload (numpitron_b200/workloads.py)
trained for a few eager steps under tp x dp, every rank compared with the oracle's emulated world
(vectors: an .npz written by tests/golden/make_config_golden.py — committed for c3 / c4, produced on the fly
by the test for the small c1 world).

    _nccl_config_worker.py <config> <vectors.npz> <precision>
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import numpitron_b200 as nb  # noqa: E402
from _record import record, rel_err  # noqa: E402
from numpitron_b200 import distributed as npdist  # noqa: E402
from numpitron_b200 import runtime as rt  # noqa: E402
from numpitron_b200.data.synthetic import local_rows  # noqa: E402
from numpitron_b200.distributed import peer  # noqa: E402
from numpitron_b200.nn import Transformer, softmax_cross_entropy  # noqa: E402
from numpitron_b200.optimizer import Adam  # noqa: E402
from numpitron_b200.workloads import BETA1, BETA2, LR, WORKLOADS  # noqa: E402


def main(config, vectors, precision):
    rank = int(os.environ["RANK"])
    Z = np.load(vectors)
    tp, dp, steps = int(Z["tp"]), int(Z["dp"]), int(Z["steps"])
    rs, lc, ers, ec = int(Z["row_step"]), int(Z["logit_cols"]), int(Z["e_row_step"]), int(Z["e_cols"])
    nb.set_precision(precision)
    npdist.init(tp_size=tp, dp_size=dp)
    # norm-wise bounds: fp32-accurate mode rel 1e-4 per layer / 1e-3 on the loss (north star); bf16 mode 2e-2 on
    # the loss, operand-rounding-limited on everything else
    # bf16 mode, dE (the embedding gradient as it reaches Adam): it has crossed every block's two LayerNorm backward
    # passes, whose bf16-STORED inputs ride on a mean 30-70x their std (Q12), so one bf16 ulp of x is ~10 % of the
    # normalised value.  Measured norm-wise error of this slice: 0.3 after ONE block (tests/test_layers_gpu.py),
    # 0.41-0.77 after the 12 blocks of c3, 0.46-0.48 after the 24 of c4 — the individual gradient entries of the
    # first layers are noise-dominated in this mode, while their sums over the batch (what the loss sees) are not:
    # the losses below agree to 1e-4..1e-3.  The north-star bar of bf16 mode is the loss (2e-2); for dE the test
    # asks for the direction (cosine) only.  fp32 mode is held to 2e-3 element-wise.
    tol_loss, tol_logits, tol_dE = (1e-3, 2e-4, 2e-3) if precision == "fp32" else (2e-2, 3e-2, None)
    model = Transformer(**WORKLOADS[config]["model"], rng=np.random.default_rng(42))
    opt = Adam(model, LR, BETA1, BETA2)
    model.scatter()
    opt.scatter()
    dr = npdist.data_parallel_rank()
    errs = {}
    for i in range(steps):
        x, y = rt.from_numpy(local_rows(Z[f"x{i}"], dp, dr)), rt.from_numpy(local_rows(Z[f"y{i}"], dp, dr))
        logits = model(x)
        loss, d = softmax_cross_entropy(logits, y)
        model.backward(d)
        if i == 0:
            lg = rt.to_numpy(logits)
            lg = lg.reshape(-1, lg.shape[-1])
            errs["logits0"] = rel_err(lg[::rs, :lc], Z[f"rank{rank}/logits0"])
            errs["logits0_rowsum"] = rel_err(lg.sum(-1, dtype=np.float64), Z[f"rank{rank}/logits0_rowsum"].reshape(-1))
            dE = rt.to_numpy(model.input_embedding.embedding.gradient)
            want_dE = Z[f"rank{rank}/dE0"].astype(np.float64)
            errs["dE0"] = rel_err(dE[::ers, :ec], Z[f"rank{rank}/dE0"])
            errs["dE0_cos"] = float((dE[::ers, :ec] * want_dE).sum() /
                                    max(np.linalg.norm(dE[::ers, :ec]) * np.linalg.norm(want_dE), 1e-300))
        opt.step()
        want = Z[f"loss{i}/dp{dr}"]
        got = rt.to_numpy(loss)
        errs[f"loss{i}"] = rel_err(got, want)
        errs[f"loss{i}_mean"] = float(abs(got.mean() - want.mean()) / abs(want.mean()))
    E = rt.to_numpy(model.input_embedding.embedding.data)[::ers, :ec]
    errs["E_final"] = rel_err(E, Z[f"rank{rank}/E_final"])
    errs["E_final_mean_abs_diff"] = float(np.abs(E - Z[f"rank{rank}/E_final"]).mean())
    path = "single" if tp == 1 else ("peer-push" if peer.active() else "nccl")
    n_rowsplit = sum(1 for b in model.layers.values() for ln in (getattr(b, "norm1", None), getattr(b, "norm2", None))
                     if ln is not None and getattr(ln, "ran_row_sharded", False))
    path += f" rowsplit {n_rowsplit}"
    record(f"config_parity[{config}-{precision}]", tp=tp, dp=dp, path=path, **errs)
    print(f"rank {rank} errors {errs} path {path}", flush=True)
    assert errs["logits0"] <= tol_logits and errs["logits0_rowsum"] <= tol_logits, errs
    assert (errs["dE0"] <= tol_dE) if tol_dE is not None else (errs["dE0_cos"] >= 0.5), errs
    for i in range(steps):
        assert errs[f"loss{i}_mean"] <= tol_loss and errs[f"loss{i}"] <= 5 * tol_loss, errs
    # Adam's first updates are ~ lr * sign(g): single weights may sit a few lr apart when a tiny gradient's sign
    # differs; on average they must agree much better
    assert errs["E_final"] <= 1e-3 and errs["E_final_mean_abs_diff"] <= 0.3 * LR * steps, errs
    torch.cuda.synchronize()
    print(f"rank {rank} ok", flush=True)
    sys.stdout.flush()
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
