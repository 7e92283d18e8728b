"""Builds libb200seg.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python segmentation-pipeline_b200/build.py [--force]

The library is compiled for B200 only (`-gencode arch=compute_100a,code=sm_100a`), with -lineinfo so that
ncu source pages map to the .cu files.  nvcc cross-compiles without a GPU; the resulting .so is git-ignored
but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200seg.so")
SOURCES = ["common.cu", "hbm_kernels.cu", "norm_kernels.cu", "loss_kernels.cu", "ccl_kernels.cu", "train_kernels.cu", "conv_direct.cu", "conv_tc.cu"]
HEADERS = ["common.cuh", "ptx_sm100.cuh", os.path.join("..", "..", "include", "b200seg.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC, *FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=4) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(objdir, os.path.basename(s) + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or not _newer(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
