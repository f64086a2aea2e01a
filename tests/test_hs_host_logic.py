"""Planning code of the half-split path that runs on the host (no GPU): the parity-class decomposition of a
strided dgrad (``hs_make_parity``, csrc/hs_gemm.cuh) must enumerate exactly the (tap, source pixel) pairs of the
generic rule ``source = (dest + pad - tap) / stride`` (the rule the CPU oracle's conv transpose follows), and the
power-of-two slot scales must put the slot maximum into [2^14, 2^15).  The checker is ``tests/host/hs_host_test.cu``
(host code of the same header the kernels are built from), compiled with nvcc and executed on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_classes_and_scales(tmp_path):
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "hs_host_test")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "hs_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASS" in out.stdout
