"""The BASELINE.json workloads (configs[0..3]) as data: model shape, tp x dp partition per GPU count,
batch sizes.  Shared by bench.py, the multi-GPU parity tests and the golden-vector generator."""
from __future__ import annotations

LR, BETA1, BETA2 = 1e-4, 0.9, 0.99   # train_shakespeare.py:160-162 (CLI defaults)

WORKLOADS = {
    # configs[0] / configs[1]: Shakespeare-char GPT; N = 1: tp1 dp1, N >= 2: tp2 x dp(N/2)
    "c1": dict(
        name="Shakespeare-char GPT L6 d384 H6 ctx256 vocab65 (BASELINE configs[1] shape)",
        model=dict(d_model=384, n_heads=6, n_layers=6, vocab_size=65, seq_len=256),
        seqs_per_gpu=64, sweep=None),
    # configs[2]: GPT-2 small, tp = all GPUs (8 in BASELINE), bf16; per-replica batch 8 / 16 / 32
    "c3": dict(
        name="GPT-2 small shape L12 d768 H12 ctx1024 vocab50257 (BASELINE configs[2]), tp = all ranks",
        model=dict(d_model=768, n_heads=12, n_layers=12, vocab_size=50257, seq_len=1024),
        global_batch=32, sweep=(8, 16, 32)),
    # configs[3]: GPT-2 medium, tp2 x dp(N/2), global batch sweep 4 .. 64
    "c4": dict(
        name="GPT-2 medium shape L24 d1024 H16 ctx1024 vocab50257 (BASELINE configs[3]), tp2 x dp(N/2)",
        model=dict(d_model=1024, n_heads=16, n_layers=24, vocab_size=50257, seq_len=1024),
        global_batch=64, sweep=(4, 8, 16, 32, 64)),
}


def partition(config: str, n_gpus: int, tp: int | None = None) -> tuple[int, int]:
    """(tp, dp) of a workload on n_gpus ranks — the reference's world is arange(n).reshape(pp=1, dp, tp).
    `tp` overrides the BASELINE partition (development runs on fewer GPUs)."""
    if tp:
        assert n_gpus % tp == 0
        return tp, n_gpus // tp
    if config == "c3":
        return n_gpus, 1
    tp = 1 if n_gpus == 1 else 2
    return tp, max(n_gpus // tp, 1)


def flop_per_token(model: dict, causal_skip: bool = False) -> float:
    """Dense FLOPs of one training step per token: 3 x (L (24 d^2 + 4 S d) + 2 d V), SURVEY §8(d);
    with causal_skip the attention term counts the lower triangle only (2 S d)."""
    d, L, S, V = model["d_model"], model["n_layers"], model["seq_len"], model["vocab_size"]
    att = (2 if causal_skip else 4) * S * d
    return 3.0 * (L * (24 * d * d + att) + 2 * d * V)
