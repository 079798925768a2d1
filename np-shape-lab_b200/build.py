"""Builds ``libnpsl.so`` (the C-ABI library of include/npsl.h) in-tree for sm_100a.

    python -m np_shape_lab_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the development container; the built
library travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnpsl.so")
SOURCES = ["npsl.cu"]
import glob

# every source, header and table under csrc/ plus the public header: a stale library is never picked up silently
DEPENDS = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                 glob.glob(os.path.join(CSRC, "*.inc"))) + [os.path.join(ROOT, "include", "npsl.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=false",  # explicit fma() only: see mesh_kernels.cuh
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(exe):
        raise RuntimeError("nvcc not found; libnpsl.so cannot be built")
    return exe


def stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPENDS:
        p = d if os.path.isabs(d) else os.path.join(CSRC, d)
        if os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        sys.stderr.write(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libnpsl.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
