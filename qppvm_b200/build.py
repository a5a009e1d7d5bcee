"""Builds the in-tree native library (sm_100a only): ``qppvm_b200/libqppvm_b200.so``."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libqppvm_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "qppvm_capi.cu"), os.path.join(HERE, "csrc", "qppvm_multi.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "shapes.def"), os.path.join(HERE, "csrc", "qp_kernel.cuh"), os.path.join(HERE, "csrc", "rbd_kernel.cuh"),
                  os.path.join(HERE, "..", "include", "qppvm_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
