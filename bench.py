#!/usr/bin/env python
"""Headline benchmark: train tokens/s of the NuMPItron training step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c3|c4] [--precision bf16|fp32]
                    [--global-batch B] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workloads (numpitron_b200/workloads.py; synthetic uint16 tokens, random-init weights in the reference's rng order):
  c1 (default, the headline): BASELINE configs[0]/[1] — Shakespeare-char GPT L6 d384 H6 ctx256 vocab65.
      N = 1: tp1 dp1.  N >= 2: tp2 x dp(N/2), the reference's partitioning.  Weak scaling: 64 sequences
      (16,384 tokens) per GPU per step.
  c3: BASELINE configs[2] — GPT-2 small shape, tp = N (8 in BASELINE), bf16 mode, per-replica batch 8/16/32.
  c4: BASELINE configs[3] — GPT-2 medium shape, tp2 x dp(N/2), global batch sweep 4..64.

One JSON line on stdout (rank 0).  `value` = whole-job tokens/s with the token batches already resident in
HBM (CUDA-graph replay of the full step: forward, loss, backward, Adam); `e2e` = the same through the public
host-buffer API (`GraphedTrainStep.step(x, y)`: pinned H2D of the uint16 tokens + graph + D2H of the mean loss,
synchronised every step).  `parity` = the first training steps on a small batch, run before the timed region on
the path that is timed (peer push / NCCL), against the CPU oracle's emulated tp x dp world.  `roofline` = the
tcgen05 GEMM kernel family and `roofline_hbm` = the memory-bound kernels, both timed launch by launch with CUDA
events in one extra eager step.  `fp32_mode` (c1, N = 1) = a few steps of the fp32-accurate mode.  `cpu_baseline`
= the unmodified reference on a bounded sample, rank 0, N = 1.  `--impl reference` times that CPU path alone, as
N OS processes partitioned tp x dp like the GPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from numpitron_b200.data.synthetic import draw_batch, local_rows, synthetic_tokens  # noqa: E402
from numpitron_b200.workloads import BETA1, BETA2, LR, WORKLOADS, flop_per_token, partition  # noqa: E402

METRIC, UNIT = "train tokens/s", "tokens/s"
GEMM_TRAFFIC_FILE = ROOT / "profiles" / "r02_gemm_dram_traffic.json"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


def gemm_traffic(config: str, precision: str, n_gpus: int):
    """Average DRAM bytes per GEMM launch of the step from the committed ncu capture of this command
    (c1, bf16, 1 GPU only) — (value, source)."""
    if config != "c1" or precision != "bf16" or n_gpus != 1 or not GEMM_TRAFFIC_FILE.exists():
        return None, None
    j = json.loads(GEMM_TRAFFIC_FILE.read_text())
    return j["avg_dram_bytes_per_launch"], f"ncu dram__bytes_read+write per gemm_tc launch, committed capture {GEMM_TRAFFIC_FILE.name}"


def workload_config(config: str, n_gpus: int, precision: str, global_batch: int | None = None, tp: int | None = None):
    w = WORKLOADS[config]
    tp, dp = partition(config, n_gpus, tp)
    if global_batch is None:
        global_batch = w["seqs_per_gpu"] * n_gpus if config == "c1" else w["global_batch"]
    S = w["model"]["seq_len"]
    return {
        "workload": f"{w['name']}, full train step (fwd + softmax-CE + bwd + Adam)", "config": config,
        "tp": tp, "dp": dp, "parallelism": f"tp{tp}xdp{dp}",
        "global_batch": global_batch, "seq_len": S, "tokens_per_step": global_batch * S,
        "precision_mode": precision,
        "l2": "no explicit flush: one step streams > 2 GB of activations, far beyond the 126 MB L2",
    }


# ---------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (baseline/_ref) on the host cores, one OS process per rank
# ---------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_reference_world(config: str, n_ranks: int, steps: int, warmup: int, budget_s: float, seqs_per_replica: int = 1):
    """Runs baseline/reference_worker.py as `n_ranks` processes (tp x dp as the GPU arm partitions `config`).
    Returns dict(value tokens/s, ms_per_step, steps, cores, threads_per_rank, sample, losses) or raises."""
    if not (ROOT / "baseline" / "_ref" / "numpitron").exists():
        raise RuntimeError("baseline/_ref is missing (run baseline/install_reference.sh where /root/reference exists)")
    w = WORKLOADS[config]
    tp, dp = partition(config, n_ranks)
    cores = os.cpu_count() or 1
    threads = max(cores // n_ranks, 1)
    global_batch = seqs_per_replica * dp
    with tempfile.TemporaryDirectory() as tmp:
        tok = Path(tmp) / "tokens.bin"
        synthetic_tokens(w["model"]["vocab_size"]).tofile(tok)     # the uint16 file the reference memmaps (loader.py:32)
        spec = dict(tp=tp, dp=dp, model=w["model"], lr=LR, beta1=BETA1, beta2=BETA2, tokens_path=str(tok),
                    global_batch=global_batch, warmup=warmup, steps=steps, budget_s=budget_s)
        port = _free_port()
        procs = []
        for r in range(n_ranks):
            env = dict(os.environ, WORLD_SIZE=str(n_ranks), RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                       OPENBLAS_NUM_THREADS=str(threads), OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
            # under torchrun the parent's environment describes ITS world: drop all of it — in particular
            # TORCHELASTIC_USE_AGENT_STORE, which would make rank 0 of this CPU world wait for the elastic agent's
            # store on the new port instead of starting its own (every rank then hangs in the rendezvous)
            for k in list(env):
                if k.startswith("TORCHELASTIC_") or k in ("LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE",
                                                          "ROLE_RANK", "ROLE_WORLD_SIZE", "ROLE_NAME", "TORCH_NCCL_ASYNC_ERROR_HANDLING"):
                    env.pop(k, None)
            procs.append(subprocess.Popen([sys.executable, str(ROOT / "baseline" / "reference_worker.py"), json.dumps(spec)],
                                          env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        results, deadline = [], time.time() + budget_s * 3 + 300
        for p in procs:
            try:
                out, err = p.communicate(timeout=max(deadline - time.time(), 1))
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise RuntimeError("reference world timed out")
            lines = [ln for ln in out.splitlines() if ln.startswith("REFERENCE_RESULT ")]
            if p.returncode != 0 or not lines:
                for q in procs:
                    if q.poll() is None:
                        q.kill()
                raise RuntimeError(f"reference rank failed (rc={p.returncode}): {err[-800:]}")
            results.append(json.loads(lines[-1][len("REFERENCE_RESULT "):]))
    n = min(len(r["times"]) for r in results)
    step_s = statistics.median([max(r["times"][i] for r in results) for i in range(n)])   # max over ranks, median of steps
    S = w["model"]["seq_len"]
    return dict(value=global_batch * S / step_s, ms_per_step=step_s * 1e3, steps=n, cores=cores, threads_per_rank=threads,
                global_batch=global_batch, tp=tp, dp=dp, losses=results[0]["losses"],
                sample=f"unmodified reference (baseline/_ref, numpy {np.__version__}) as {n_ranks} OS process(es) tp{tp}xdp{dp} over "
                       f"oracle/mpi4py_stub (gloo; the image has no mpirun/mpi4py), {threads} BLAS thread(s) per rank on {cores} host "
                       f"cores; {seqs_per_replica} sequence(s) x {S} tokens per replica per step (global batch {global_batch}), "
                       f"{n} timed step(s) after {warmup} warm-up, median of the per-step max over ranks")


def run_reference_arm(args):
    """`--impl reference`.  Under torchrun only rank 0 works (it spawns its own N-process CPU world)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = workload_config(args.config, args.gpus, "fp32")
    warm = min(args.warmup, 1)
    try:
        r = run_reference_world(args.config, args.gpus, args.steps, warm, budget_s=150.0)
    except Exception as e:   # the reference package did not travel / cannot run: say so, exit 0
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:300]}), flush=True)
        return
    cfg |= {"global_batch": r["global_batch"], "tokens_per_step": r["global_batch"] * cfg["seq_len"],
            "reference_sample": f"{r['global_batch']} sequence(s) per step instead of the GPU arm's {cfg['global_batch']} "
                                f"(the literal reference needs ~6 s per sequence per step at this shape); same model, "
                                f"partition tp{r['tp']}xdp{r['dp']}, seeds, token stream and optimizer settings"}
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "losses": r["losses"][:5],
    }
    print(json.dumps(line), flush=True)


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
# parity leg: the checker (CPU oracle) against the path that is about to be timed
# ---------------------------------------------------------------------------------------------------
def parity_expected(config: str, tp: int, dp: int, n_steps: int):
    """(global batches [(x, y)], per-step list of per-replica per-token losses, source string).
    c1: the oracle's emulated tp x dp world, run here on the host (seconds).  c3 / c4: the same oracle needs
    minutes per step, its vectors are committed (tests/golden/config_*.npz, make_config_golden.py)."""
    m = WORKLOADS[config]["model"]
    if config == "c1":
        from oracle import numpitron_oracle as O   # test infrastructure, used here as the checker only

        orc = O.OracleTransformer(m["d_model"], m["n_heads"], m["n_layers"], m["vocab_size"], m["seq_len"],
                                  rng=np.random.default_rng(42), tp=tp, dp=dp, literal=False, lr=LR, beta1=BETA1, beta2=BETA2)
        data = synthetic_tokens(m["vocab_size"])
        brng = np.random.default_rng(7)
        batches, losses = [], []
        for _ in range(n_steps):
            x, y = draw_batch(data, brng, 2 * dp, m["seq_len"])
            batches.append((x, y))
            res = orc.forward_loss(x, y)
            losses.append([loss for (loss, _, _) in res])
            orc.train_step(x, y)   # advance the oracle (recomputes the forward; cheap at this shape)
        return batches, losses, "oracle/numpitron_oracle.py (fast mode), emulated world, computed in this run"
    f = ROOT / "tests" / "golden" / f"config_{config}.npz"
    if not f.exists():
        return None, None, f"{f.name} missing"
    z = np.load(f)
    if int(z["tp"]) != tp or int(z["dp"]) != dp:
        return None, None, f"{f.name} holds tp{int(z['tp'])}xdp{int(z['dp'])}, this run is tp{tp}xdp{dp}"
    n = min(n_steps, int(z["steps"]))
    batches = [(z[f"x{i}"], z[f"y{i}"]) for i in range(n)]
    losses = [[z[f"loss{i}/dp{d}"] for d in range(dp)] for i in range(n)]
    return batches, losses, f"tests/golden/{f.name} (oracle fast mode, generated by tests/golden/make_config_golden.py)"


def run_parity(config, tp, dp, precision, model_args, n_steps=2):
    """Trains a fresh model for `n_steps` eager steps on the small parity batches and compares every
    per-token loss with the oracle's.  Collective on all ranks; returns the sub-object on every rank."""
    import torch
    import torch.distributed as dist

    from numpitron_b200 import distributed as npdist
    from numpitron_b200 import runtime as rt
    from numpitron_b200.distributed import peer
    from numpitron_b200.nn import Transformer, softmax_cross_entropy
    from numpitron_b200.optimizer import Adam

    tol = 2e-2 if precision == "bf16" else 1e-3
    rank = int(os.environ.get("RANK", "0"))
    info = [None]
    if rank == 0:
        t0 = time.time()
        batches, losses, src = parity_expected(config, tp, dp, n_steps)
        info[0] = (batches, losses, src, time.time() - t0)
    if dist.is_initialized():
        dist.broadcast_object_list(info, src=0)
    batches, losses, src, oracle_s = info[0]
    if batches is None:
        return {"checked": False, "why": src}
    model = Transformer(**model_args, rng=np.random.default_rng(42))
    opt = Adam(model, LR, BETA1, BETA2)
    model.scatter()
    opt.scatter()
    dr = npdist.data_parallel_rank()
    worst_tok, worst_mean, got_means, want_means = 0.0, 0.0, [], []
    for (x, y), want in zip(batches, losses):
        xl, yl = rt.from_numpy(local_rows(x, dp, dr)), rt.from_numpy(local_rows(y, dp, dr))
        logits = model(xl)
        loss, d = softmax_cross_entropy(logits, yl)
        model.backward(d)
        opt.step()
        got = rt.to_numpy(loss)
        w = np.asarray(want[dr], dtype=np.float32).reshape(got.shape)
        worst_tok = max(worst_tok, float(np.abs(got - w).max() / np.abs(w).max()))
        worst_mean = max(worst_mean, float(abs(got.mean() - w.mean()) / abs(w.mean())))
        got_means.append(float(got.mean()))
        want_means.append(float(w.mean()))
    t = torch.tensor([worst_tok, worst_mean], device=rt.device())
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst_tok, worst_mean = float(t[0]), float(t[1])
    path = "single GPU" if tp == 1 else ("peer-push (GEMM epilogue -> NVLink peer memory, reduced inside LayerNorm) + NCCL"
                                         if peer.active() else "nccl all-reduce")
    del model, opt
    torch.cuda.empty_cache()
    return {"checked": True, "max_rel_err": worst_mean, "max_rel_err_per_token": worst_tok, "tolerance": tol,
            "ok": bool(worst_mean <= tol), "quantity": f"loss of the first {len(batches)} training steps (fwd + loss + bwd + Adam "
            f"each), per-replica batch {batches[0][0].shape[0] // dp} sequence(s), every rank, max over ranks; max_rel_err = "
            f"|mean loss - oracle| / oracle, per_token = max |loss - oracle| / max |oracle|",
            "loss_rank0": got_means, "oracle_rank0": want_means, "path": path, "oracle": src, "oracle_seconds": round(oracle_s, 1)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def profile_eager_step(step, model, opt, batch, ms_per_step, precision):
    """One extra eager step with CUDA events around every C-ABI call -> (roofline of the GEMM family,
    roofline_hbm list of the memory-bound kernels, attention summary)."""
    import torch

    from numpitron_b200 import ops
    from numpitron_b200 import runtime as rt
    from numpitron_b200.train import train_step_device

    ops.PROFILE = []
    with torch.cuda.stream(step.stream):
        for _ in range(2):
            ops.PROFILE.clear()
            # Park the GPU for ~8 ms first so the host runs ahead and every launch of the step is already
            # queued when the GPU reaches it: the event pairs then bracket pure kernel time (no host-latency
            # gaps), with the step's real cache state.
            torch.cuda._sleep(16_000_000)
            with rt.deferred_wgrad():   # the launches of the graphed step (split-K / LayerNorm partials go to Adam un-reduced)
                train_step_device(model, opt, *batch)
    step.stream.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    for p in prof:
        p["ms"] = p["events"][0].elapsed_time(p["events"][1])
    pk = peaks()

    def family(name):
        return [p for p in prof if p["name"] == name]

    g = family("gemm")
    eager_ms, tot_flops = sum(p["ms"] for p in g), sum(p["work"] for p in g)
    # The event pairs of the eager step serialise every launch (an event record sits between two kernels, so the
    # programmatic-dependent-launch overlap of the real step is off and each pair includes the launch latency).
    # The GEMM family is therefore timed again the way it runs inside the step: the same launches — same
    # argument blocks, same buffers — captured back to back into one CUDA graph and replayed between two events.
    lib = __import__("numpitron_b200._lib", fromlist=["load"]).load()
    import ctypes as C
    with torch.cuda.stream(step.stream):
        def replay_gemms():
            for p in g:
                ops.check(lib.npt_gemm(C.byref(p["args"]), torch.cuda.current_stream().cuda_stream), "npt_gemm")
        replay_gemms()
        step.stream.synchronize()
        # capture_begin / capture_end directly: the torch.cuda.graph context empties the caching allocator first,
        # which would unmap the (already released, still cached) buffers these launches point at
        graph = torch.cuda.CUDAGraph()
        graph.capture_begin()
        replay_gemms()
        graph.capture_end()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_rep = 5
        e0.record()
        for _ in range(n_rep):
            graph.replay()
        e1.record()
        step.stream.synchronize()
        tot_ms = e0.elapsed_time(e1) / n_rep
        del graph
    achieved = tot_flops / (tot_ms * 1e-3) / 1e12
    tc = precision == "bf16"
    roof = {
        "bound": "tensor", "achieved": achieved, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_burst"],
        "kernel": "gemm_tc_kernel (tcgen05.mma + TMA)" if tc else "gemm_tf32_kernel (3xTF32 tcgen05.mma) / gemm_simt_kernel",
        "launches": len(g), "ms_per_step_in_kernel": tot_ms, "flop_per_step": tot_flops,
        "timing": f"the step's {len(g)} GEMM launches (same arguments and buffers) replayed back to back as one CUDA graph, "
                  f"CUDA events around {n_rep} replays on the launching stream",
        "peak_source": f"{pk['src']} bf16 burst",
        "share_of_step": tot_ms / ms_per_step,
        "eager_event_pairs": {"ms_per_step_in_kernel": eager_ms, "frac": tot_flops / (eager_ms * 1e-3) / 1e12 / pk["bf16_burst"],
                              "note": "one event pair per launch in an eager step: launches serialised, launch latency included"},
    }
    hbm = []
    for name in sorted({p["name"] for p in prof if p["bound"] == "hbm"}):
        f = family(name)
        ms, by = sum(p["ms"] for p in f), sum(p["work"] for p in f)
        hbm.append({"kernel": name, "launches": len(f), "bytes": by, "us": ms * 1e3, "GBps": by / (ms * 1e-3) / 1e9 if ms else None,
                    "frac": by / (ms * 1e-3) / 1e9 / pk["hbm"] if ms else None, "share_of_step": ms / ms_per_step})
    hbm.sort(key=lambda r: -r["us"])
    att = {}
    for name in ("attention_fwd", "attention_bwd"):
        f = family(name)
        if f:
            ms, fl = sum(p["ms"] for p in f), sum(p["work"] for p in f)
            att[name] = {"launches": len(f), "us": ms * 1e3, "TFLOPs": fl / (ms * 1e-3) / 1e12,
                         "frac": fl / (ms * 1e-3) / 1e12 / pk["bf16_burst"], "share_of_step": ms / ms_per_step,
                         "flops": "causal (lower triangle only)"}
    return roof, hbm, att


def run_ours(args):
    import torch
    import torch.distributed as dist

    import numpitron_b200 as nb
    from numpitron_b200 import distributed as npdist
    from numpitron_b200 import runtime as rt
    from numpitron_b200.nn import Transformer
    from numpitron_b200.optimizer import Adam
    from numpitron_b200.train import GraphedTrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: numpitron_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    n_gpus = args.gpus
    assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    w = WORKLOADS[args.config]
    MODEL = w["model"]
    cfg = workload_config(args.config, n_gpus, args.precision, args.global_batch, args.tp)
    tp, dp = cfg["tp"], cfg["dp"]
    nb.set_precision(args.precision)
    npdist.init(tp_size=tp, dp_size=dp)
    if args.dp_sync != "reference":
        from numpitron_b200.nn import transformer as _tr

        _tr.set_dp_sync(args.dp_sync)
        cfg["dp_sync"] = args.dp_sync + " (non-reference: every gradient all-reduced over the DP group)"
    dev = rt.device()
    S = MODEL["seq_len"]
    dp_rank = npdist.data_parallel_rank()

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

    # ---- parity leg (before anything is timed) ----------------------------------------------------------
    parity = None if args.no_parity else run_parity(args.config, tp, dp, args.precision, MODEL)

    model = Transformer(**MODEL, rng=np.random.default_rng(42))
    opt = Adam(model, LR, BETA1, BETA2)
    model.scatter()
    opt.scatter()
    data = synthetic_tokens(MODEL["vocab_size"])
    K, W = args.steps, max(args.warmup, 3)

    def measure(global_batch: int, with_e2e: bool, sampler=None):
        """value / e2e legs for one global batch -> dict."""
        local_batch = global_batch // dp
        brng = np.random.default_rng(99)   # identical on every rank: lock-step draws (data/loader.py:48-54)
        host_batches = []
        for _ in range(8):
            x, y = draw_batch(data, brng, global_batch, S)
            host_batches.append((local_rows(x, dp, dp_rank), local_rows(y, dp, dp_rank)))
        dev_batches = [(rt.from_numpy(x), rt.from_numpy(y)) for x, y in host_batches]
        step = GraphedTrainStep(model, opt, local_batch, S, warmup=2)
        for i in range(W + 3):   # eager warm-ups, capture, graph warm-ups
            step.load_resident(*dev_batches[i % 8])
            step.step_resident_async()
        step.stream.synchronize()
        barrier()
        if sampler:
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
        if sampler:
            sampler.mark_end()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        tokens = global_batch * S
        out = dict(global_batch=global_batch, tokens_per_step=tokens, ms_per_step=ms_total / K,
                   value=tokens * K / (ms_total * 1e-3), launches_per_step=step.launches_per_step,
                   final_loss=float(step.loss_dev.item()), step=step, dev_batches=dev_batches)
        if with_e2e:
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
            out["e2e"] = {"value": tokens * K / (ms_e2e * 1e-3), "unit": UNIT,
                          "h2d_bytes_per_step": step.h2d_bytes_per_step * world, "d2h_bytes_per_step": step.d2h_bytes_per_step * world,
                          "ms_per_step": ms_e2e / K, "wall_ms_per_step": wall * 1e3 / K}
            out["e2e_last_loss"] = loss
        return out

    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    sweep = []
    if args.sweep and w["sweep"]:
        for gb in [b for b in w["sweep"] if b != cfg["global_batch"] and b % dp == 0]:
            r = measure(gb, with_e2e=False)
            sweep.append({"global_batch": gb, "value": r["value"], "ms_per_step": r["ms_per_step"]})
            del r
            torch.cuda.empty_cache()
    main = measure(cfg["global_batch"], with_e2e=True, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None
    if sweep:
        sweep.append({"global_batch": cfg["global_batch"], "value": main["value"], "ms_per_step": main["ms_per_step"]})
        sweep.sort(key=lambda r: r["global_batch"])
    step = main["step"]

    # ---- roofline legs: every launch of one eager step, CUDA events on the launching stream ---------------
    roof, hbm, att = profile_eager_step(step, model, opt, main["dev_batches"][0], main["ms_per_step"], args.precision)
    barrier()   # tp = 2: the replayed GEMMs push tiles into the peer's memory — nobody moves on (or tears down) before both are done
    roof["traffic"], roof["traffic_source"] = gemm_traffic(args.config, args.precision, n_gpus)

    # ---- fp32-accurate mode (the reference's own precision), a few steps, c1 / N = 1 only -------------------
    fp32_mode = None
    if args.config == "c1" and n_gpus == 1 and args.precision == "bf16" and not args.no_fp32:
        del step
        main.pop("step")
        nb.set_precision("fp32")
        model32 = Transformer(**MODEL, rng=np.random.default_rng(42))
        opt32 = Adam(model32, LR, BETA1, BETA2)
        model32.scatter()
        opt32.scatter()
        st32 = GraphedTrainStep(model32, opt32, cfg["global_batch"], S, warmup=2)
        for i in range(5):
            st32.load_resident(*main["dev_batches"][i % 8])
            st32.step_resident_async()
        st32.stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n32 = 10
        with torch.cuda.stream(st32.stream):
            e0.record()
        for i in range(n32):
            st32.load_resident(*main["dev_batches"][i % 8])
            st32.step_resident_async()
        with torch.cuda.stream(st32.stream):
            e1.record()
        st32.stream.synchronize()
        ms32 = e0.elapsed_time(e1) / n32
        r32, _, _ = profile_eager_step(st32, model32, opt32, main["dev_batches"][0], ms32, "fp32")
        fp32_mode = {"value": cfg["tokens_per_step"] / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "steps": n32,
                     "gemm": "3xTF32 on tcgen05 (kind::tf32, hi/lo split, fp32 TMEM accumulate)" if rt.gemm_mode() == 2 else "fp32 FFMA",
                     "roofline": r32, "final_loss": float(st32.loss_dev.item())}
        nb.set_precision(args.precision)
    if rank != 0:
        return

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only): the unmodified reference ----------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline and args.config == "c1":
        try:
            r = run_reference_world(args.config, 1, steps=2, warmup=1, budget_s=25.0)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"[:300]}

    fpt = flop_per_token(MODEL)
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak" if args.config == "c1" else "strong",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": cfg,
        "clocks": clocks, "e2e": main["e2e"],
        "gpu_launches": int(main["launches_per_step"]) * K, "gpu_launches_per_step": int(main["launches_per_step"]),
        "parity": parity, "roofline": roof, "roofline_hbm": hbm, "attention": att, "fp32_mode": fp32_mode,
        "cpu_baseline": cpu, "model_flop_per_token": fpt, "model_tflops": main["value"] * fpt / 1e12,
        "final_loss": main["final_loss"], "e2e_last_loss": main["e2e_last_loss"],
    }
    if sweep:
        line["sweep"] = sweep
    if dp > 1:
        # data-parallel gradient traffic per rank and step: the reference all-reduces the tied
        # embedding gradient only (nn/transformer.py:52-57, Q11); dp_sync = "all" sums every gradient
        def local_params(layer):
            n = sum(int(p.data.numel()) for p in layer.parameters.values())
            return n + sum(local_params(sub) for sub in getattr(layer, "layers", {}).values())
        emb = int(model.input_embedding.embedding.data.numel())
        line["dp_allreduce"] = {"mode": args.dp_sync, "group_size": dp,
                                "message_bytes_reference_mode": 4 * emb, "message_bytes_all_mode": 4 * local_params(model),
                                "note": "fp32 gradients of this rank's shards; reference mode = 4 * d_model * V_local"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--config", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--global-batch", type=int, default=None)
    ap.add_argument("--tp", type=int, default=None, help="override the BASELINE partition: tp x (N / tp)")
    ap.add_argument("--sweep", action="store_true", help="c3 / c4: also time the other batch sizes of the sweep")
    ap.add_argument("--dp-sync", default="reference", choices=["reference", "all"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"c1": 200, "c3": 20, "c4": 10}[args.config]
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
