"""Build `numpitron_b200/lib/libnumpitron_b200.so` from `csrc/*.cu` with nvcc for sm_100a.

    python -m numpitron_b200.build [--force] [--verbose]

The library is built IN-TREE so it travels to the GPU box with the repo snapshot; it is
git-ignored.  No torch, no pybind: a plain C-ABI shared object (see include/numpitron_b200.h).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libnumpitron_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v" if os.environ.get("NPT_PTXAS_V") else "-O3"]
FLAGS += os.environ.get("NPT_NVCC_EXTRA", "").split()  # e.g. -DNPT_GEMM_DEBUG for the timing experiments


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "numpitron_b200.h"]
    jobs = []
    for src in sources():
        obj = OBJDIR / (src.stem + ".o")
        if force or _stale(obj, [src, *headers]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *ARCH, *FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(compile_one, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {src.name}\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [OBJDIR / (s.stem + ".o") for s in sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
