"""Build the in-tree CUDA library (lcqpow_b200/lib/liblcqp_cuda.so) for sm_100a with nvcc.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a
GPU.  ``python -m lcqpow_b200.build`` or ``__graft_entry__.build()`` call :func:`build`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblcqp_cuda.so")
# the same sources with -DLCQP_COUNT_WORK: the active-set kernel counts its fp64 multiply-adds and L2 bytes
# (lcqp_cuda_last_work); bench.py runs it on a sub-batch, outside the timed region, for roofline.work
LIB_WORK = os.path.join(LIBDIR, "liblcqp_cuda_work.so")
SOURCES = [os.path.join(CSRC, "lcqp_cabi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("lcqp_device.cuh", "lcqp_pas.cuh", "lcqp_osqp.cuh", "lcqp_osqp_impl.inc", "lcqp_sparse_host.hpp")] + [os.path.join(HERE, "..", "include", "lcqp_cuda.h")]
# LCQP_NO_ASSUME: the address-space hints (__builtin_assume(__isShared(p))) of lcqp_device.cuh are off -- with
# them nvcc 12.9 miscompiles the solver for sm_100a (garbage return values of the non-inlined device functions;
# seen on the B200, gpurun_out/dbg1.log vs dbg2.log of round 1).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-DLCQP_NO_ASSUME", "-diag-suppress", "20054", "-diag-suppress", "128", "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def up_to_date() -> bool:
    for lib in (LIB, LIB_WORK):
        if not os.path.exists(lib):
            return False
        t = os.path.getmtime(lib)
        if not all(os.path.getmtime(d) <= t for d in DEPS):
            return False
    return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    cmd_work = [_nvcc()] + NVCC_FLAGS + ["-DLCQP_COUNT_WORK", "-o", LIB_WORK] + SOURCES
    procs = [subprocess.Popen(c, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in (cmd, cmd_work)]
    for pr, name in zip(procs, ("liblcqp_cuda.so", "liblcqp_cuda_work.so")):
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed building " + name)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
