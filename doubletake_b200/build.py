"""Build the C-ABI shared library IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

    python -m doubletake_b200.build [--force] [-v]

Output: doubletake_b200/libdoubletake_b200.so (git-ignored, but it travels with gpurun snapshots).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdoubletake_b200.so")
SOURCES = ["capi.cu", "cost_volume.cu", "cost_volume_tc.cu", "cost_volume_tch.cu", "conv_simt.cu", "conv_tc.cu", "conv_tch.cu", "conv_graph.cu", "tsdf.cu", "encoder.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "doubletake_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("DTB200_NVCC_EXTRA", "").split()  # development switches, e.g. -DDTB200_RAW_BIG
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcuda"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
