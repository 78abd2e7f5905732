"""Builds the CUDA library in-tree for sm_100a with nvcc (no JIT cache, no torch headers).

    python -m otg_b200.build            # or: python optimal-trajectory-.../build.py

The resulting `libfrenet_b200.so` sits next to this file; it is git-ignored but travels to the
GPU box with the repository snapshot.  `__graft_entry__.build()` calls `build()`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SOURCES = [os.path.join(PKG_DIR, "csrc", "frenet_kernels.cu")]
HEADERS = [os.path.join(ROOT, "include", "frenet_b200.h")]
LIB_PATH = os.path.join(PKG_DIR, "libfrenet_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the library is missing or older than its sources; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
