"""Builds librustpotter_b200.so (the C-ABI library: CUDA kernels for sm_100a + host side) in-tree.

nvcc cross-compiles without a GPU. Objects are cached under rustpotter_b200/_obj and rebuilt when a
source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librustpotter_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -ffp-contract=off: host-side table construction and audio filters must follow the reference's
# un-fused f32 operation order. Device code keeps FMA contraction except where the kernels use
# explicit __f*_rn intrinsics. No --use_fast_math (denormal mel energies must survive).
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
              "-Xptxas", "-v", "-I", INCLUDE]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(INCLUDE, "rustpotter_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr = _headers_mtime()
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr, os.path.getmtime(__file__))
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, "-x", "cu", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = obj + ".log"
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
