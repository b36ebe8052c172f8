"""Shared-memory carve-up of the chain-step kernel (rc::smem_layout, recom_device.cuh):
arrays that are live together never overlap, the spill split keeps exactly the membership
bitmap in shared memory.  Host-only: tests/native/layout_check.cu is compiled with nvcc and run
on the CPU (no device code is launched)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_smem_layout_invariants(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "layout_check")
    subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(HERE, "native", "layout_check.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.startswith("ok "), out.stdout
