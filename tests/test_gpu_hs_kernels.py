"""Half-split tcgen05 kernels (csrc/hs_gemm.cuh) against a CPU double-precision restatement of their contract.

The checker lives in ``tools/hs_selftest.cu`` (it links the real kernels, no engine, no torch): gather GEMM
(forward with both segments, C % 64 != 0, stride-1 / stride-2 dgrad incl. the parity-class decomposition, ragged
shapes, accumulate, bias, > 148 tiles), the N-stacked shared-activation kernel and the multi-slot wgrad (1..8 slots, several splits, TMEM flush) must
match the reference to 2e-6 of the slot maximum (1.5e-5 / 5e-5 for the two long accumulation-chain cases, which
bound the tensor core's truncating accumulator).  ``__graft_entry__.build()`` compiles the binary in-tree.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "hs_selftest")


@pytest.mark.gpu
@pytest.mark.parametrize("producers", ["tma", "cp.async"])
def test_half_split_kernels_match_cpu_reference(producers):
    if not os.path.exists(BIN):
        import __graft_entry__ as ge

        ge.build()
    # argv: which (1 gather | 2 wgrad | 32 N-stacked gather), debug knob, producer mode (1 = TMA, 0 = cp.async only)
    out = subprocess.run([BIN, "35", "0", "1" if producers == "tma" else "0"], capture_output=True, text=True,
                         timeout=600)
    print(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ALL PASS" in out.stdout
    assert out.stdout.count("PASS") >= 24
