"""In-tree build of libedgpu.so (hand-written sm_100a CUDA + the C ABI of include/edgpu.h).

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to the GPU
box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["abi.cu", "sector.cu", "hxv.cu", "csr.cu", "vecops.cu", "lanczos.cu", "eigs.cu", "packed.cu", "comm.cu", "extra.cu", "orbs.cu"]
LIB = os.path.join(HERE, "libedgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(CSRC, "edgpu_internal.cuh"), os.path.join(CSRC, "trlan.hpp"),
            os.path.join(HERE, "..", "include", "edgpu.h")]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-ccbin", "/usr/bin/g++", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
