"""Builds the in-tree CUDA library (sm_100a only) with nvcc."""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librecom_b200.so")          # 8-bit district labels (k <= 255)
OUT_L16 = os.path.join(HERE, "librecom_b200_l16.so")  # the same sources with 16-bit labels
SOURCES = ["recom_abi.cu"]
DEPS = ["recom_abi.cu", "recom_kernels.cu", "recom_stats.cu", "recom_device.cuh",
        "../../include/recom_b200.h", "../../include/recom_rng.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def lib_path(label_bits: int = 8) -> str:
    if label_bits not in (8, 16):
        raise ValueError("label_bits must be 8 or 16")
    return OUT if label_bits == 8 else OUT_L16


def up_to_date(label_bits: int = 8) -> bool:
    out = lib_path(label_bits)
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False, label_bits=(8, 16)) -> str:
    """Compiles the library once per label width (include/recom_b200.h: RC_LABEL_BITS);
    the two compilations run side by side.  Returns the path of the 8-bit build."""
    widths = (label_bits,) if isinstance(label_bits, int) else tuple(label_bits)
    procs = []
    for bits in widths:
        if not force and up_to_date(bits):
            continue
        cmd = [
            nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
            "-std=c++17", "-Xcompiler", "-fPIC", "-shared", f"-DRC_LABEL_BITS={bits}", "-o", lib_path(bits),
        ] + (["-Xptxas", "-v"] if verbose and bits == 8 else []) + SOURCES
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
