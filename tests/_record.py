"""Measured parity errors of the GPU tests, appended as JSON lines to gpurun_out/parity_measured.jsonl so a
`gpurun` call brings them back (copied to profiles/ per round)."""
import json
import os
from pathlib import Path

OUT = Path(__file__).resolve().parents[1] / "gpurun_out" / "parity_measured.jsonl"


def record(test: str, **values) -> None:
    try:
        OUT.parent.mkdir(exist_ok=True)
        with open(OUT, "a") as f:
            f.write(json.dumps({"test": test, "rank": int(os.environ.get("RANK", "0")), **values}) + "\n")
    except OSError:
        pass


def rel_err(got, want) -> float:
    """Norm-wise error: max |got - want| / max |want| (the reference does not meet an element-wise
    rtol against itself)."""
    import numpy as np

    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
