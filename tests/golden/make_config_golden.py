"""Golden vectors for the multi-GPU parity of BASELINE configs[2] (c3: GPT-2 small, tp=8) and configs[3]
(c4: GPT-2 medium, tp=2 x dp=4): 3 (c3) / 2 (c4) training steps of the CPU oracle's emulated tp x dp world
(fast mode: BLAS contractions + compact softmax backward, pinned to the literal reference by
tests/test_oracle.py) at a reduced batch.  The oracle needs minutes per step at these shapes, so the
vectors are generated here once and committed; `tests/test_multigpu.py` and `bench.py`'s parity leg read them.

    python tests/golden/make_config_golden.py c3|c4

Kept small: the per-token losses of every step, and fixed slices of the step-0 logits, of the embedding
gradient as it reaches Adam at step 0, and of the embedding after the last step, for every rank.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from numpitron_b200.data.synthetic import draw_batch, synthetic_tokens  # noqa: E402
from numpitron_b200.workloads import BETA1, BETA2, LR, WORKLOADS  # noqa: E402
from oracle import numpitron_oracle as O  # noqa: E402

SPEC = {"c3": dict(tp=8, dp=1, global_batch=1, steps=3), "c4": dict(tp=2, dp=4, global_batch=4, steps=2)}
ROW_STEP, LOGIT_COLS, E_ROW_STEP, E_COLS = 16, 64, 4, 128


def main(cfg, spec=None, out_path=None):
    w, sp = WORKLOADS[cfg], (spec or SPEC[cfg])
    m = w["model"]
    tp, dp = sp["tp"], sp["dp"]
    orc = O.OracleTransformer(m["d_model"], m["n_heads"], m["n_layers"], m["vocab_size"], m["seq_len"],
                              rng=np.random.default_rng(42), tp=tp, dp=dp, literal=False, lr=LR, beta1=BETA1, beta2=BETA2)
    orc.keep = True
    data = synthetic_tokens(m["vocab_size"])
    brng = np.random.default_rng(99)
    out = {"tp": tp, "dp": dp, "global_batch": sp["global_batch"], "steps": sp["steps"],
           "row_step": ROW_STEP, "logit_cols": LOGIT_COLS, "e_row_step": E_ROW_STEP, "e_cols": E_COLS}
    for i in range(sp["steps"]):
        x, y = draw_batch(data, brng, sp["global_batch"], m["seq_len"])
        out[f"x{i}"], out[f"y{i}"] = x, y
        t0 = time.time()
        res = orc.forward_loss(x, y)   # per-token losses (train_step returns only the means)
        losses = [loss for (loss, _, _) in res]
        # same arithmetic again inside train_step would double the cost: finish the step by hand
        for row, (loss, ctx, dl) in zip(orc.ranks, res):
            orc._backward_replica(row, ctx, dl)
        if dp > 1:
            for tr in range(tp):
                tot = orc.ranks[0][tr].grads["embedding"].copy()
                for dr in range(1, dp):
                    tot = tot + orc.ranks[dr][tr].grads["embedding"]
                for dr in range(dp):
                    orc.ranks[dr][tr].grads["embedding"] = tot.copy()
        for dr in range(dp):
            out[f"loss{i}/dp{dr}"] = losses[dr]
            for tr in range(tp):
                r = dr * tp + tr
                if i == 0:
                    lg = orc.last_logits[dr][tr]
                    out[f"rank{r}/logits0"] = lg.reshape(-1, lg.shape[-1])[::ROW_STEP, :LOGIT_COLS].copy()
                    out[f"rank{r}/logits0_rowsum"] = lg.sum(-1, dtype=np.float64).astype(np.float32)
                    out[f"rank{r}/dE0"] = orc.ranks[dr][tr].grads["embedding"][::E_ROW_STEP, :E_COLS].copy()
        for row in orc.ranks:
            for st in row:
                for name, g in st.grads.items():
                    st.params[name], st.m[name], st.v[name] = O.adam_update(
                        st.params[name], g, st.m[name], st.v[name], orc.lr, orc.beta1, orc.beta2, orc.eps)
                st.grads = {}
        print(cfg, "step", i, [float(l.mean()) for l in losses], f"{time.time() - t0:.0f}s", flush=True)
    for dr in range(dp):
        for tr in range(tp):
            out[f"rank{dr * tp + tr}/E_final"] = orc.ranks[dr][tr].params["embedding"][::E_ROW_STEP, :E_COLS].copy()
    np.savez_compressed(out_path or Path(__file__).resolve().parent / f"config_{cfg}.npz", **out)


if __name__ == "__main__":
    for c in sys.argv[1:]:
        main(c)
