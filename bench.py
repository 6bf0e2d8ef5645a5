#!/usr/bin/env python
"""Headline benchmark: train tokens/s of the NuMPItron training step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1] shape): Shakespeare-char GPT — 6 layers, d_model 384, 6 heads,
ctx 256, vocab 65 — synthetic uint16 tokens, random-init weights in the reference's rng order.
N = 1: tp=1 dp=1.  N >= 2: tp=2, dp=N/2 (the reference's partitioning).  Weak scaling: 64
sequences (16,384 tokens) per GPU per step, global batch 64*N.

One JSON line on stdout (rank 0).  `value` = whole-job tokens/s with the token batches already
resident in HBM (CUDA-graph replay of the full step: forward, loss, backward, Adam); `e2e` = the
same through the public host-buffer API (`GraphedTrainStep.step(x, y)`: pinned H2D of the uint16
tokens + graph + D2H of the mean loss, synchronised every step).  `roofline` = the tcgen05 GEMM
kernel family, timed launch by launch with CUDA events in one extra eager step.  `cpu_baseline`
= the CPU oracle (a NumPy port of the reference's algorithm) on a bounded sample, rank 0, N = 1.
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

def gemm_traffic(precision: str, n_gpus: int):
    """Average DRAM bytes per GEMM launch of the step, from the committed ncu capture (bf16, 1 GPU only)."""
    f = ROOT / "profiles" / "r01_gemm_dram_traffic.json"
    if precision != "bf16" or n_gpus != 1 or not f.exists():
        return None
    return json.loads(f.read_text())["avg_dram_bytes_per_launch"]


MODEL = dict(d_model=384, n_heads=6, n_layers=6, vocab_size=65, seq_len=256)
SEQS_PER_GPU = 64
LR, BETA1, BETA2 = 1e-4, 0.9, 0.99
FLOP_PER_TOKEN_FULL = 70.93e6      # 3 * (L*(24 d^2 + 4 S d) + 2 d V), SURVEY §8(d)
METRIC, UNIT = "train tokens/s", "tokens/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's algorithm; test infrastructure used as a baseline only)
# ---------------------------------------------------------------------------------------------------
def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_tokens_per_s(batch: int, steps: int, warmup: int, literal: bool, budget_s: float):
    """tokens/s of the oracle's train_step on `batch` x 256 tokens per step; stops at the time budget."""
    from oracle import numpitron_oracle as O

    model = O.OracleTransformer(MODEL["d_model"], MODEL["n_heads"], MODEL["n_layers"], MODEL["vocab_size"],
                                MODEL["seq_len"], rng=np.random.default_rng(42), literal=literal,
                                lr=LR, beta1=BETA1, beta2=BETA2)
    data = O.synthetic_tokens(MODEL["vocab_size"])
    brng = np.random.default_rng(99)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        x, y = O.draw_batch(data, brng, batch, MODEL["seq_len"])
        t0 = time.perf_counter()
        model.train_step(x, y)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and times:
            break
    med = statistics.median(times)
    return batch * MODEL["seq_len"] / med, med, len(times)


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU algorithm (literal oracle port: non-BLAS einsum
    contractions and the (B,H,S,S,S) softmax Jacobian, exactly what nn/linear.py and
    nn/activation.py do) on the box's host cores.  The Python reference itself cannot travel to the
    GPU box (/root/reference does not exist there); the port is pinned bit-exact to it by
    tests/test_oracle.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 1
    tps, med, done = cpu_oracle_tokens_per_s(batch, args.steps, min(args.warmup, 1), literal=True, budget_s=200.0)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "fp32") | {"reference_sample": f"{batch} x 256 tokens per step"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"literal oracle port, {batch} sequence x 256 tokens per step, {done} timed steps "
                                   f"(einsum contractions are single-threaded in NumPy; {cores} BLAS threads available)"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, precision: str):
    tp = 1 if n_gpus == 1 else 2
    dp = max(n_gpus // tp, 1)
    return {
        "workload": "Shakespeare-char GPT L6 d384 H6 ctx256 vocab65 (BASELINE configs[1] shape), full train step "
                    "(fwd + softmax-CE + bwd + Adam)",
        "tp": tp, "dp": dp, "parallelism": f"tp{tp}xdp{dp}",
        "global_batch": SEQS_PER_GPU * n_gpus, "seq_len": MODEL["seq_len"],
        "tokens_per_step": SEQS_PER_GPU * n_gpus * MODEL["seq_len"],
        "precision_mode": precision,
        "l2": "no explicit flush: one step streams > 2 GB of activations/attention matrices, far beyond the 126 MB L2",
    }


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.1]
        rows = inside if inside else [r for _, r in self.rows]   # very short runs: fall back to the warm-up samples
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_in_timed_region": len(inside)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import numpitron_b200 as nb
    from numpitron_b200 import _lib, ops
    from numpitron_b200 import distributed as npdist
    from numpitron_b200 import runtime as rt
    from numpitron_b200.nn import Transformer
    from numpitron_b200.optimizer import Adam
    from numpitron_b200.train import GraphedTrainStep, train_step_device
    from oracle import numpitron_oracle as O  # synthetic token stream helper only (not timed)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: numpitron_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    n_gpus = args.gpus
    assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    cfg = workload_config(n_gpus, args.precision)
    tp, dp = cfg["tp"], cfg["dp"]
    nb.set_precision(args.precision)
    npdist.init(tp_size=tp, dp_size=dp)
    dev = rt.device()

    rng = np.random.default_rng(42)
    model = Transformer(**MODEL, rng=rng)
    opt = Adam(model, LR, BETA1, BETA2)
    model.scatter()
    opt.scatter()

    global_batch, S = cfg["global_batch"], MODEL["seq_len"]
    local_batch = global_batch // dp
    data = O.synthetic_tokens(MODEL["vocab_size"])
    brng = np.random.default_rng(99)   # identical on every rank: lock-step draws (data/loader.py:48-54)
    dp_rank = npdist.data_parallel_rank()

    def draw():
        x, y = O.draw_batch(data, brng, global_batch, S)
        return (np.ascontiguousarray(np.array_split(x, dp, axis=0)[dp_rank]),
                np.ascontiguousarray(np.array_split(y, dp, axis=0)[dp_rank]))

    K, W = args.steps, max(args.warmup, 3)
    step = GraphedTrainStep(model, opt, local_batch, S, warmup=2)
    host_batches = [draw() for _ in range(8)]
    dev_batches = [(rt.from_numpy(x), rt.from_numpy(y)) for x, y in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg (the `value`) -------------------------------------------------------
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    for i in range(W + 3):   # eager warm-ups, capture, graph warm-ups
        step.load_resident(*dev_batches[i % 8])
        step.step_resident_async()
    step.stream.synchronize()
    barrier()
    sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(step.stream):
        e0.record()
    for i in range(K):
        step.load_resident(*dev_batches[i % 8])
        step.step_resident_async()
    with torch.cuda.stream(step.stream):
        e1.record()
    barrier()
    sampler.mark_end()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = step.launches_per_step
    tokens_per_step = cfg["tokens_per_step"]
    value = tokens_per_step * K / (ms_total * 1e-3)
    last_loss = float(step.loss_dev.item())

    # ---- end-to-end leg through the host-buffer API ---------------------------------------------------
    for i in range(3):
        step.step(*host_batches[i % 8])
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(step.stream):
        e2.record()
    for i in range(K):
        loss = step.step(*host_batches[i % 8])   # H2D + graph + D2H + sync, every step
    with torch.cuda.stream(step.stream):
        e3.record()
    barrier()
    wall = time.perf_counter() - t0
    ms_e2e = max_over_ranks(max(e2.elapsed_time(e3), 0.0))
    e2e_value = tokens_per_step * K / (ms_e2e * 1e-3)

    # ---- roofline leg: every GEMM launch of one eager step, CUDA events on the launching stream -----
    roof = None
    if rank == 0 or world > 1:
        ops.GEMM_PROFILE = []
        with torch.cuda.stream(step.stream):
            for _ in range(2):
                ops.GEMM_PROFILE.clear()
                # Park the GPU for ~8 ms first so the host runs ahead and every launch of the step is
                # already queued when the GPU reaches it: the event pairs then bracket pure kernel
                # time (no host-latency gaps), with the step's real cache state.
                torch.cuda._sleep(16_000_000)
                train_step_device(model, opt, *dev_batches[0])
        step.stream.synchronize()
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        tot_ms = sum(p["events"][0].elapsed_time(p["events"][1]) for p in prof)
        tot_flops = sum(p["flops"] for p in prof)
        pk = peaks()
        tc = args.precision == "bf16"
        achieved = tot_flops / (tot_ms * 1e-3) / 1e12
        peak = pk["bf16_burst"]
        roof = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": gemm_traffic(args.precision, n_gpus),
            "kernel": "gemm_tc_kernel (tcgen05.mma + TMA)" if tc else "gemm_simt_kernel (fp32 FFMA)",
            "launches": len(prof), "ms_per_step_in_kernel": tot_ms, "flop_per_step": tot_flops,
            "peak_source": f"{pk['src']} bf16 burst (kernel timed alone, launch by launch)",
            "share_of_step": tot_ms / (ms_total / K),
        }
    if rank != 0:
        return

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        tps, med, done = cpu_oracle_tokens_per_s(1, 2, 1, literal=True, budget_s=40.0)
        ftps, fmed, fdone = cpu_oracle_tokens_per_s(4, 2, 1, literal=False, budget_s=20.0)
        cores = blas_threads()
        cpu = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"literal oracle port of the reference algorithm: 1 sequence x 256 tokens per step, {done} timed "
                         f"steps after 1 warm-up ({med:.1f} s/step; einsum parts single-threaded, {cores} BLAS threads)",
               "fast_oracle_value": ftps,
               "fast_oracle_sample": f"BLAS matmul + compact softmax backward, 4 x 256 tokens per step, {fdone} steps"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": cfg,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes_per_step * world,
                "d2h_bytes_per_step": step.d2h_bytes_per_step * world, "ms_per_step": ms_e2e / K,
                "wall_ms_per_step": wall * 1e3 / K},
        "gpu_launches": int(launches_per_step) * K,
        "gpu_launches_per_step": int(launches_per_step),
        "roofline": roof,
        "cpu_baseline": cpu,
        "model_flop_per_token": FLOP_PER_TOKEN_FULL,
        "model_tflops": value * FLOP_PER_TOKEN_FULL / 1e12,
        "final_loss": last_loss, "e2e_last_loss": loss,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
