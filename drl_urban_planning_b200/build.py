"""Build the CUDA/C++ shared library in-tree with nvcc (sm_100a only).

    python -m drl_urban_planning_b200.build [--force] [--verbose]

The library has no torch dependency: plain CUDA runtime + C ABI (include/upb200.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libupb200.so")
SOURCES = ["upb200.cu", "pack_host.cpp", "errors.cpp"]
DEPS = SOURCES + ["sgnn_kernel.cuh", "mlp_kernel.cuh", "optim_kernels.cuh", "layout.h", "blob.h", "errors.h",
                  os.path.join("..", "..", "include", "upb200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


PYPTR = os.path.join(HERE, "_upb_pyptr.so")


def build_pyptr(force: bool = False) -> str:
    """Small CPython helper (host glue: pointer table of the state lists).  gcc only; optional at run time."""
    import sysconfig
    src = os.path.join(CSRC, "pyptr.c")
    if not force and os.path.exists(PYPTR) and os.path.getmtime(PYPTR) >= os.path.getmtime(src):
        return PYPTR
    import numpy
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
           "-o", PYPTR, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return PYPTR


LIB_BF16 = os.path.join(HERE, "libupb200_bf16.so")


def build_bf16_variant(force: bool = False) -> str:
    """The labelled NON-PARITY variant with bf16-rounded, single-pass tensor-core tiles (bench.py --tiles bf16)."""
    if not force and os.path.exists(LIB_BF16) and all(
            os.path.getmtime(os.path.join(CSRC, d)) <= os.path.getmtime(LIB_BF16) for d in DEPS):
        return LIB_BF16
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DUPB_TILE_BF16",
           "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-o", LIB_BF16] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return LIB_BF16


def build_library(force: bool = False, verbose: bool = False) -> str:
    build_pyptr(force)
    if not force and not stale():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
