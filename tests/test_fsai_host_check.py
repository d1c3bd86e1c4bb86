"""CPU check of the FSAI fast path's index maps: the phase bodies of vface_b200/csrc/vf_fsai_fast.cuh are
__host__ __device__, tests/csrc/fsai_host_check.cu emulates one CTA thread by thread and compares with a
naive double-precision DFT form of combine_fft_high_low (scripts/face_swap_utils.py:425-464).
Compiles for the host only (no GPU, no device code is run)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_fsai_fast_path_index_maps_on_cpu(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fsai_host_check")
    src = os.path.join(ROOT, "tests", "csrc", "fsai_host_check.cu")
    res = subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "all ok" in run.stdout
