"""One rank (one GPU) of the multi-GPU parity test: the tiny model trained for 3 steps under
tp x dp, compared with the golden vectors recorded from the reference run as tp*dp processes."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import numpitron_b200 as nb  # noqa: E402
from numpitron_b200 import distributed as npdist  # noqa: E402
from numpitron_b200.nn import Transformer, softmax_cross_entropy  # noqa: E402
from numpitron_b200.optimizer import Adam  # noqa: E402
from numpitron_b200.runtime import to_numpy  # noqa: E402
from numpitron_b200.train import GraphedTrainStep  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
TINY = dict(d_model=32, n_heads=4, n_layers=2, vocab_size=17, seq_len=16)


def load(name):
    with np.load(GOLDEN / name) as z:
        return {k: z[k] for k in z.files}


def close(got, want, rtol, what):
    got, want = to_numpy(got), np.asarray(want)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * np.abs(want).max(), err_msg=what)


def main(tp, dp, precision):
    rank = int(os.environ["RANK"])
    nb.set_precision(precision)
    npdist.init(tp_size=tp, dp_size=dp)
    rtol = 2e-5 if precision == "fp32" else 3e-2
    W, G = load(f"tiny_tp{tp}_dp{dp}.npz"), load("layers_tp1.npz")
    model = Transformer(**TINY, rng=np.random.default_rng(42))
    opt = Adam(model, 1e-4, 0.9, 0.99)
    model.scatter()
    opt.scatter()
    for i in range(3):
        x, y = G[f"tiny/x{i}"], G[f"tiny/y{i}"]
        if dp > 1:
            x = np.ascontiguousarray(np.array_split(x, dp, axis=0)[npdist.data_parallel_rank()])
            y = np.ascontiguousarray(np.array_split(y, dp, axis=0)[npdist.data_parallel_rank()])
        logits = model(x)
        loss, d = softmax_cross_entropy(logits, y)
        model.backward(d)
        if i == 0:
            close(logits, W[f"rank{rank}/logits0"], rtol, "logits0")
            close(model.input_embedding.embedding.gradient, W[f"rank{rank}/dE0"], 10 * rtol, "dE0")
        opt.step()
        close(loss, W[f"rank{rank}/loss{i}"], rtol, f"loss{i}")
    close(model.input_embedding.embedding.data, W[f"rank{rank}/E_final"], rtol, "E_final")
    # Adam's update is lr*m/(sqrt(v)+eps) ~ lr*sign(g) in the first steps: with bf16 operands a tiny
    # gradient whose sign flips moves that weight by up to 2*lr per step the other way, so after 3 steps
    # single weights may sit up to 6*lr from the reference; on average they must agree far better.
    w_got = to_numpy(model.block_0.mlp.column_linear.weight.data)
    w_want = W[f"rank{rank}/w1_final_block0"]
    if precision == "fp32":
        # 3xTF32 contractions differ from the reference's by ~1e-6 relative in a gradient; through Adam's
        # m / (sqrt(v) + eps) that is a fraction of a percent of one lr step on single weights (measured 2.6e-7)
        assert np.abs(w_got - w_want).max() <= 1e-6, np.abs(w_got - w_want).max()
    else:
        diff = np.abs(w_got - w_want)
        assert diff.max() <= 6.1e-4 and diff.mean() <= 5e-5, (diff.max(), diff.mean())

    # the CUDA-graph replay (NCCL all-reduces captured inside) gives the eager numbers
    model2 = Transformer(**TINY, rng=np.random.default_rng(42))
    opt2 = Adam(model2, 1e-4, 0.9, 0.99)
    model2.scatter()
    opt2.scatter()
    lb = 4 // dp
    g = GraphedTrainStep(model2, opt2, lb, 16, warmup=1)
    for i in range(3):
        x, y = G[f"tiny/x{i}"], G[f"tiny/y{i}"]
        if dp > 1:
            x = np.ascontiguousarray(np.array_split(x, dp, axis=0)[npdist.data_parallel_rank()])
            y = np.ascontiguousarray(np.array_split(y, dp, axis=0)[npdist.data_parallel_rank()])
        got = g.step(x, y)
        want = W[f"rank{rank}/loss{i}"].mean()
        assert abs(got - want) <= rtol * abs(want), (i, got, want)
    assert g.graph is not None
    close(model2.input_embedding.embedding.data, W[f"rank{rank}/E_final"], rtol, "graph E_final")
    torch.cuda.synchronize()
    print(f"rank {rank} ok", flush=True)
    sys.stdout.flush()
    # orderly teardown: the captured graphs hold NCCL kernels and references into the symmetric-memory
    # exchange — release them, quiesce the device, then leave the process group together
    import gc

    from numpitron_b200.distributed import peer

    del g, model2, opt2, model, opt
    peer.reset()
    gc.collect()
    torch.cuda.synchronize()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
