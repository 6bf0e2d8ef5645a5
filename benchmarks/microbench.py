#!/usr/bin/env python
"""Per-layer micro-benchmarks (BASELINE.json configs[4]): Linear, Attention, LayerNorm, softmax-CE,
Adam across d_model 384-4096 and seq 256-4096, each against its roofline (measured peaks from
MEASURED_PEAKS.json: bf16 tensor burst for GEMM-shaped work, HBM copy bandwidth for the rest).

    python benchmarks/microbench.py [--quick] > profiles/r01_microbench.jsonl

Timing: CUDA events on the launching stream, 3 warm-ups, L2 flushed (256 MB memset) between timed
iterations for the memory-bound kernels; one JSON line per (op, shape).
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

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else \
    {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TENSOR = PEAKS["hbm_gbs"], PEAKS["bf16_tflops"]
BF16 = torch.bfloat16
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=10, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters * 1e-3  # seconds


def emit(op, shape, seconds, flops=None, bytes_=None, **extra):
    line = {"op": op, "shape": shape, "us": seconds * 1e6}
    if flops is not None:
        tf = flops / seconds / 1e12
        line |= {"bound": "tensor", "achieved": tf, "peak": TENSOR, "unit": "TFLOP/s", "frac": tf / TENSOR}
    else:
        gb = bytes_ / seconds / 1e9
        line |= {"bound": "hbm", "achieved": gb, "peak": HBM, "unit": "GB/s", "frac": gb / HBM}
    line |= extra
    print(json.dumps(line), flush=True)


def bench_linear(T, d, mode):
    dt = BF16 if mode == 1 else torch.float32
    x = torch.randn(T, d, device="cuda").to(dt)
    w = (torch.randn(d, d, device="cuda") * 0.02).to(dt)
    dy = torch.randn(T, d, device="cuda").to(dt)
    y, dx, dw = rt.empty((T, d)), rt.empty((T, d)), rt.empty((d, d))
    flops = 2.0 * T * d * d
    name = {0: "linear_fp32_ffma", 1: "linear_bf16", 2: "linear_fp32_3xtf32"}[mode]
    s = timeit(lambda: ops.gemm(x, w, T, d, d, lda=d, ldb=d, C_out=y, ldc=d, mode=mode), flush=False)
    emit(name + "_fwd_NN", {"T": T, "d": d}, s, flops=flops)
    s = timeit(lambda: ops.gemm(dy, w, T, d, d, lda=d, ldb=d, trans_b=True, C_out=dx, ldc=d, mode=mode), flush=False)
    emit(name + "_dX_NT", {"T": T, "d": d}, s, flops=flops)
    s = timeit(lambda: ops.gemm(x, dy, d, d, T, lda=d, ldb=d, trans_a=True, C_out=dw, ldc=d, mode=mode), flush=False)
    emit(name + "_dW_TN", {"T": T, "d": d}, s, flops=flops)


def bench_attention(B, S, d):
    h, Dh = d // 64, 64
    if not ops.attention_supported(S, Dh):
        return
    qkv = torch.randn(B, S, h * 3 * Dh, device="cuda").to(BF16)
    do = torch.randn(B, S, h * Dh, device="cuda").to(BF16)
    div = float(np.float32(d**0.5))
    out, lse = ops.attention_fwd(qkv, B, S, h, Dh, div)
    causal = 0.5 * (1 + 128 / S)          # fraction of 128x128 tiles on or below the diagonal
    f_fwd = 4.0 * B * h * S * S * Dh * causal
    s = timeit(lambda: ops.attention_fwd(qkv, B, S, h, Dh, div), flush=False)
    emit("attention_fused_fwd", {"B": B, "S": S, "d": d, "heads": h}, s, flops=f_fwd, note="causal tiles only")
    s = timeit(lambda: ops.attention_bwd(qkv, out, do, lse, B, S, h, Dh, div), flush=False)
    emit("attention_fused_bwd", {"B": B, "S": S, "d": d, "heads": h}, s, flops=3.5 * f_fwd,
         note="7 tile GEMMs (S and dP are recomputed in the dQ kernel), causal tiles only")


def bench_layernorm(T, d):
    x = (1.5 + 0.03 * torch.randn(T, d, device="cuda"))
    res = torch.randn(T, d, device="cuda")
    g, b = torch.rand(d, device="cuda"), torch.zeros(d, device="cuda")
    dy = torch.randn(T, d, device="cuda")
    _, mean, var = ops.layernorm_fwd(x, g, b, 1e-5)
    s = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-5, residual=res, want_bf16=True))
    emit("layernorm_fwd(+residual,+bf16 copy)", {"T": T, "d": d}, s, bytes_=T * d * (4 + 4 + 4 + 2))
    s = timeit(lambda: ops.layernorm_bwd(x, g, mean, var, dy, 1e-5))
    emit("layernorm_bwd", {"T": T, "d": d}, s, bytes_=T * d * 12)


def bench_loss(T, V):
    logits = torch.randn(T, V, device="cuda") * 3
    labels = torch.randint(0, V, (T,), device="cuda").to(torch.int32).to(torch.uint16)
    s = timeit(lambda: ops.softmax_ce_fused(logits, labels, 0, V))
    emit("softmax_ce_fused", {"T": T, "V": V}, s, bytes_=T * (8 * V + 6))


def bench_adam(n):
    p, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda") * 1e-3
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    table = ops.AdamTable(1)
    table.fill([(p, g, m, v, None)])
    f = lambda x: float(np.float32(x))  # noqa: E731
    s = timeit(lambda: ops.adam_step(table, f(1e-4), f(0.9), f(0.1), f(0.99), f(0.01), f(0.1), f(0.01), f(1e-7)))
    emit("adam", {"params": n}, s, bytes_=28 * n)


def bench_embedding(T, d, V):
    tok = torch.randint(0, V, (T,), device="cuda").to(torch.int32).to(torch.uint16).reshape(1, T)
    E = torch.randn(d, V, device="cuda")
    dout = torch.randn(1, T, d, device="cuda")
    grad = torch.zeros(d, V, device="cuda")
    s = timeit(lambda: ops.embedding_gather(tok, E, None, T, 0, V, True))
    emit("embedding_gather", {"T": T, "d": d, "V": V}, s, bytes_=T * (2 + 8 * d))
    s = timeit(lambda: ops.embedding_scatter_add(tok, dout, grad, 0, V, True))
    emit("embedding_scatter_add(ordered, bit-exact)", {"T": T, "d": d, "V": V}, s, bytes_=T * (2 + 4 * d) + 8 * d * min(V, T))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    ds = [384, 1024, 4096] if a.quick else [384, 768, 1024, 2048, 4096]
    seqs = [256, 1024] if a.quick else [256, 1024, 4096]
    for d in ds:
        for S in seqs:
            bench_linear(8 * S, d, 1)
    for d in [384, 1024]:
        bench_linear(8 * 256, d, 0)
    for d in ds:                       # the fp32-accurate mode: 3xTF32 on the tensor cores
        bench_linear(8 * 1024, d, 2)
    for d in ds:
        for S in seqs:
            bench_attention(8, S, d)
    for d in ds:
        bench_layernorm(8 * 1024, d)
    bench_loss(8192, 65)
    bench_loss(8192, 50257)
    for n in (10_662_528, 123_614_976) + (() if a.quick else (353_674_240,)):
        bench_adam(n)
    bench_embedding(16384, 384, 65)
    bench_embedding(8192, 768, 50257)


if __name__ == "__main__":
    main()
