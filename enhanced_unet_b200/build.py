"""Build libeunet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

``python -m enhanced_unet_b200.build`` or ``build_library()``.  Objects are rebuilt only when a
source or header is newer than the object.  The .so is git-ignored but travels with the repo
snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libeunet_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["core.cu", "metrics.cu", "elementwise.cu", "loss.cu", "tail.cu", "optim.cu", "conv_direct.cu", "conv_tc.cu", "conv_halo.cu", "conv_halo2.cu", "conv_wgrad_halo.cu", "conv_dgrad_few.cu", "conv_tail_bwd.cu", "tail_out_tma.cu", "bn_apply_tma.cu",
           "mask.cu", "fusion.cu", "probe.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "eunet.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(run, jobs):
                if verbose or r.returncode != 0 or ptxas_info:
                    sys.stderr.write(f"[nvcc {src}]\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {src}")
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    path = build_library(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(path)
