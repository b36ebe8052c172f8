"""Builds the in-tree CUDA library (sm_100a only) with nvcc."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librecom_b200.so")
SOURCES = ["recom_abi.cu"]
DEPS = ["recom_abi.cu", "recom_kernels.cu", "recom_stats.cu", "recom_device.cuh",
        "../../include/recom_b200.h", "../../include/recom_rng.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
        "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", OUT,
    ] + (["-Xptxas", "-v"] if verbose else []) + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
