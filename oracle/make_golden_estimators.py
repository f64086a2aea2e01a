"""Generate tests/golden/estimators.npz by running the REFERENCE's randomised estimators and its Neumann inverse
(/root/reference, read-only) on small seeded dense matrices.  TEST INFRASTRUCTURE.
Run: python oracle/make_golden_estimators.py

Every entry is (seed -> estimate): the estimators draw their probes from torch's global CPU generator, so the
tests re-seed, call the engine-side function on the same matrix and must reproduce the number.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import numpy as np
import torch

from curvlinops import (NeumannInverseLinearOperator, hutchinson_diag, hutchinson_squared_fro,  # noqa: E402
                        hutchinson_trace, hutchpp_trace, xdiag, xtrace)
from curvlinops.diag import DiagonalLinearOperator  # noqa: E402
from curvlinops.examples import TensorLinearOperator  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_default_dtype(torch.float64)


def matrices():
    g = torch.Generator().manual_seed(1234)
    B = torch.randn(48, 48, generator=g)
    decay = torch.logspace(0, -3, 48)
    Qm, _ = torch.linalg.qr(B)
    return {"general": torch.rand(48, 48, generator=g),
            "psd_decay": (Qm * decay) @ Qm.T,
            "wide": torch.randn(20, 48, generator=g)}


out = {}
mats = matrices()
for name, M in mats.items():
    out[f"M_{name}"] = M.numpy()

cases = []
for mat in ["general", "psd_decay"]:
    A = TensorLinearOperator(mats[mat])
    for dist in ["rademacher", "normal"]:
        for fn, nm in [(hutchinson_trace, 12), (hutchpp_trace, 12), (xtrace, 12), (hutchinson_diag, 12),
                       (hutchinson_squared_fro, 12)]:
            torch.manual_seed(7)
            out[f"{fn.__name__}|{mat}|{dist}|{nm}"] = fn(A, nm, distribution=dist).numpy()
    torch.manual_seed(7)
    out[f"xdiag|{mat}|rademacher|12"] = xdiag(A, 12).numpy()
torch.manual_seed(7)
out["hutchinson_squared_fro|wide|rademacher|8"] = hutchinson_squared_fro(mats["wide"], 8).numpy()

# Neumann series, plain / scaled / Jacobi-preconditioned, K = 3 right-hand sides
g = torch.Generator().manual_seed(99)
S = mats["psd_decay"] + 0.5 * torch.eye(48)
rhs = torch.randn(48, 3, generator=g)
out["neumann_rhs"] = rhs.numpy()
out["neumann_S"] = S.numpy()
Sop = TensorLinearOperator(S)
out["neumann|plain|30"] = (NeumannInverseLinearOperator(Sop, num_terms=30) @ rhs).numpy()
out["neumann|scale0.7|50"] = (NeumannInverseLinearOperator(Sop, num_terms=50, scale=0.7) @ rhs).numpy()
jac = DiagonalLinearOperator([S.diag().reciprocal()])
out["neumann|jacobi|25"] = (NeumannInverseLinearOperator(Sop, num_terms=25, preconditioner=jac.__matmul__)
                            @ rhs).numpy()

np.savez_compressed(os.path.join(GOLDEN, "estimators.npz"), **out)
for k, v in out.items():
    if not k.startswith(("M_", "neumann_")):
        print(k, v.shape, float(np.abs(v).max()))
