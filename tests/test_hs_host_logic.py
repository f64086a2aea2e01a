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


def test_any_upper_bound_is_a_valid_plane_scale():
    """The half-split planes are floating point: choosing the power-of-two scale from an UPPER BOUND of the slot
    maximum that is loose by a factor L costs no mantissa bits, it only lifts the absolute floor to ~L * 2^-40 of
    the maximum (DESIGN.md §3.1, §4 'Next' (2): what lets BN kernels - and later GEMM epilogues - write operand
    planes before the exact maximum of what they produce is known).  numpy restatement of hs_split_kernel's
    arithmetic: hi = rn16(s*x), lo = rn16(s*x - hi)."""
    import numpy as np

    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1 << 18) * np.exp(2 * rng.standard_normal(1 << 18))).astype(np.float32)  # heavy tails
    mx = float(np.abs(x).max())
    e = int(np.floor(np.log2(mx)))
    for log2_loose, max_bound, rms_bound in [(0, -24.0, -31.0), (9, -24.0, -30.0), (13, -24.0, -26.5)]:
        s = np.float32(2.0 ** (14 - e - log2_loose))  # scaled maximum in [2^14, 2^15) / L
        v = x * s
        hi = v.astype(np.float16)
        lo = (v - hi.astype(np.float32)).astype(np.float16)
        assert np.isfinite(hi).all() and np.isfinite(lo).all()
        err = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) / float(s) - x.astype(np.float64)) / mx
        assert np.log2(err.max()) < max_bound, (log2_loose, np.log2(err.max()))
        assert np.log2(np.sqrt((err ** 2).mean())) < rms_bound, (log2_loose, np.log2(np.sqrt((err ** 2).mean())))
