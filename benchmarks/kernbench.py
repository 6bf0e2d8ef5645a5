#!/usr/bin/env python
"""Step-path kernel timings at the BASELINE shapes (the variants the bf16 training step actually launches:
bf16-input LayerNorm, fused softmax-CE, the tied-embedding gather / scatter-add, fused attention at head_dim
64 and 96), one JSON line per kernel, each against its roofline.  Used for A/B runs while a kernel is being
changed (`--only ln` etc.); `benchmarks/microbench.py` is the BASELINE configs[4] sweep.

    python benchmarks/kernbench.py [--only ln,ce,emb,attn,colsum] [--shape c1|c3|c4]

Timing: CUDA events on the launching stream, 3 warm-ups, 256 MB memset between timed iterations (L2 flush).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from numpitron_b200 import ops  # noqa: E402
from numpitron_b200 import runtime as rt  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = PEAKS.get("hbm_gbs", 6452.8)
TENSOR = PEAKS.get("bf16_tflops", 1650.2)
BF16 = torch.bfloat16
# rows = tokens per step per GPU at the bench batch; (d, V_local, heads, head_dim, seq)
SHAPES = {"c1": dict(T=16384, d=384, V=65, h=6, Dh=64, S=256, B=64),
          "c1tp2": dict(T=32768, d=384, V=33, h=3, Dh=64, S=256, B=128, tp=2),   # tp2 rank of the N >= 2 bench partition
          "c3": dict(T=32768, d=768, V=6283, h=1, Dh=96, S=1024, B=32),      # tp = 8 rank
          "c3tp1": dict(T=8192, d=768, V=50257, h=12, Dh=64, S=1024, B=8),
          "c4": dict(T=16384, d=1024, V=25129, h=8, Dh=64, S=1024, B=16)}    # tp2 x dp4 rank
_flush = None


def timeit(fn, iters=10, flush=True, reps=1):
    """`reps` > 1: that many back-to-back launches per timed region (launch queue kept full), time per launch."""
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters / reps * 1e-3


def timeit_graph(fn, reps=20, iters=5):
    """`reps` launches captured into one CUDA graph (programmatic dependent launch edges kept, no host
    overhead between launches), replayed `iters` times: time per launch as it is inside the step's graph."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        s.synchronize()
    return e0.elapsed_time(e1) / iters / reps * 1e-3


def emit(op, shape, s, flops=None, bytes_=None, **extra):
    line = {"op": op, "shape": shape, "us": round(s * 1e6, 2)}
    if flops is not None:
        tf = flops / s / 1e12
        line |= {"bound": "tensor", "achieved": round(tf, 1), "peak": TENSOR, "unit": "TFLOP/s", "frac": round(tf / TENSOR, 4)}
    else:
        gb = bytes_ / s / 1e9
        line |= {"bound": "hbm", "achieved": round(gb, 1), "peak": HBM, "unit": "GB/s", "frac": round(gb / HBM, 4)}
    print(json.dumps(line | extra), flush=True)


def ln(sh, name):
    T, d = sh["T"], sh["d"]
    x = (1.5 + 0.03 * torch.randn(T, d, device="cuda")).to(BF16)
    res = torch.randn(T, d, device="cuda")
    g, b = torch.rand(d, device="cuda"), torch.zeros(d, device="cuda")
    dy = torch.randn(T, d, device="cuda").to(BF16)
    _, mean, var = ops.layernorm_fwd(x, g, b, 1e-5)
    s = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-5, residual=res, want_bf16=True))
    emit("layernorm_fwd(bf16 in, +residual, fp32 + bf16 out)", name, s, bytes_=T * d * (2 + 4 + 4 + 2))
    s = timeit(lambda: ops.layernorm_bwd(x, g, mean, var, dy, 1e-5, bf16_only=True, want_colsum=True))
    emit("layernorm_bwd(bf16 in, bf16 out, +colsum)", name, s, bytes_=T * d * 6)
    with rt.deferred_wgrad():
        s = timeit(lambda: ops.layernorm_bwd(x, g, mean, var, dy, 1e-5, bf16_only=True, want_colsum=True))
    emit("layernorm_bwd(same, partials left to Adam)", name, s, bytes_=T * d * 6)


def ce(sh, name):
    T, V = sh["T"], sh["V"]
    logits = torch.randn(T, V, device="cuda") * 3
    labels = torch.randint(0, V, (T,), device="cuda").to(torch.int32).to(torch.uint16)
    s = timeit(lambda: ops.softmax_ce_fused(logits, labels, 0, V))
    emit("softmax_ce_fused", name, s, bytes_=T * (8 * V + 6))


def emb(sh, name):
    T, d, V = sh["T"], sh["d"], sh["V"]
    tok = torch.randint(0, V, (T,), device="cuda").to(torch.int32).to(torch.uint16).reshape(1, T)
    E = torch.randn(d, V, device="cuda")
    dout = torch.randn(1, T, d, device="cuda")
    grad = torch.zeros(d, V, device="cuda")
    s = timeit(lambda: ops.embedding_gather(tok, E, None, T, 0, V, True))
    emit("embedding_gather", name, s, bytes_=T * (2 + 8 * d))
    s = timeit(lambda: ops.embedding_scatter_add(tok, dout, grad, 0, V, True))
    emit("embedding_scatter_add(ordered, bit-exact)", name, s, bytes_=T * (2 + 4 * d) + 8 * d * min(V, T))


def colsum(sh, name):
    T, n = sh["T"], 4 * sh["d"]
    x = torch.randn(T, n, device="cuda").to(BF16)
    s = timeit(lambda: ops.colsum(T, n, x))
    emit("colsum(bf16, 4d columns)", name, s, bytes_=T * n * 2)


def attn(sh, name):
    B, S, h, Dh, d = sh["B"], sh["S"], sh["h"], sh["Dh"], sh["d"]
    if not ops.attention_supported(S, Dh):
        return
    qkv = torch.randn(B, S, h * 3 * Dh, device="cuda").to(BF16)
    do = torch.randn(B, S, h * Dh, device="cuda").to(BF16)
    div = float(np.float32(d**0.5))
    out, lse = ops.attention_fwd(qkv, B, S, h, Dh, div)
    causal = 0.5 * (1 + 128 / S)
    f = 4.0 * B * h * S * S * Dh * causal
    s = timeit_graph(lambda: ops.attention_fwd(qkv, B, S, h, Dh, div))
    emit("attention_fwd", name, s, flops=f, note="causal tiles only")
    s = timeit_graph(lambda: ops.attention_bwd(qkv, out, do, lse, B, S, h, Dh, div))
    emit("attention_bwd", name, s, flops=2.5 * f, note="5 algorithmic tile GEMMs, causal tiles only")


def gemm(sh, name):
    """The twelve Linear GEMMs of one transformer block (forward NN, dX NT, dW TN), bf16 in / bf16 out.
    `tp` in the shape: the per-rank shapes under tensor parallelism (column-parallel N / tp, row-parallel K / tp)."""
    T, d, t = sh["T"], sh["d"], sh.get("tp", 1)
    tot_us, tot_fl = 0.0, 0.0
    for M, N, K, ta, tb, what in [(T, 3 * d // t, d, 0, 0, "qkv fwd"), (T, d, d // t, 0, 0, "out fwd"), (T, 4 * d // t, d, 0, 0, "up fwd"),
                                  (T, d, 4 * d // t, 0, 0, "down fwd"), (T, d, 3 * d // t, 0, 1, "qkv dX"), (T, d // t, d, 0, 1, "out dX"),
                                  (T, d, 4 * d // t, 0, 1, "up dX"), (T, 4 * d // t, d, 0, 1, "down dX"), (d, 3 * d // t, T, 1, 0, "qkv dW"),
                                  (d // t, d, T, 1, 0, "out dW"), (d, 4 * d // t, T, 1, 0, "up dW"), (4 * d // t, d, T, 1, 0, "down dW")]:
        a = torch.randn((K, M) if ta else (M, K), device="cuda").to(BF16)
        b = torch.randn((N, K) if tb else (K, N), device="cuda").to(BF16)
        dw = what.endswith("dW")
        c = rt.empty((M, N)) if dw else rt.empty((M, N), BF16)
        kw = dict(C_out=c, ldc=N) if dw else dict(Cb_out=c, ldcb=N)
        s = timeit_graph(lambda: ops.gemm(a, b, M, N, K, lda=a.shape[1], ldb=b.shape[1], trans_a=bool(ta), trans_b=bool(tb), mode=1, **kw))
        emit(f"gemm {what} {M}x{N}x{K}", name, s, flops=2.0 * M * N * K, timing="20 launches in one CUDA graph, operands L2-warm")
        tot_us += s * 1e6
        tot_fl += 2.0 * M * N * K
    emit("gemm: the 12 Linear GEMMs of a block", name, tot_us * 1e-6, flops=tot_fl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="ln,ce,emb,colsum,attn")
    ap.add_argument("--shape", default="c1,c3,c3tp1,c4")
    a = ap.parse_args()
    rt.set_precision("bf16")
    fns = {"ln": ln, "ce": ce, "emb": emb, "colsum": colsum, "attn": attn, "gemm": gemm}
    for name in a.shape.split(","):
        for k in a.only.split(","):
            fns[k](SHAPES[name], name)


if __name__ == "__main__":
    main()
