"""Builds smart_image_processing_b200/libdocscan.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m smart_image_processing_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libdocscan.so")

CU_SOURCES = ["ctx.cu", "pointwise.cu", "scalars.cu", "blur.cu", "tcblur.cu", "morph.cu", "adaptive.cu", "warp.cu", "resize.cu", "deskew.cu", "synth.cu", "capi.cu"]
CPP_SOURCES = ["hostmath.cpp"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                       # every fused multiply-add on the path is written explicitly
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC,
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-fno-fast-math", "-I", INCLUDE]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc05.cuh"), os.path.join(INCLUDE, "docscan.h"), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for s in CU_SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    for s in CPP_SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append(["g++"] + CXX_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {cmd[-3]}")

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-Xlinker", "--exclude-libs=ALL"])
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
