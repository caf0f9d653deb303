"""Builds libconan_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m conan_b200.build [--force]
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libconan_b200.so")
STAMP = OUT + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# precise math only: no --use_fast_math anywhere on this path (erf GELU, tanh, exp must match torch)
CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
         "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map")]


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(open(os.path.join(CSRC, "exports.map"), "rb").read())
    h.update(" ".join(f for f in FLAGS if "version-script" not in f).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    if not os.path.exists(NVCC):
        if os.path.exists(OUT):
            return OUT          # box without a toolchain: use the artefact that travelled with the tree
        raise RuntimeError("nvcc not found and no prebuilt libconan_b200.so")
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cmd = [NVCC] + FLAGS + srcs + ["-o", OUT]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print("built", OUT)
