"""Checks of the downstream consumers (inverses, randomised estimators) that only need an operator, the dense
matrix it represents and the parameter shapes; shared by the CPU tests (dense fp32 operator) and the GPU tests
(engine GGN), so that the GPU run exercises exactly the logic the CPU suite has already validated."""
import warnings

import torch

from curvlinops_b200.dense import DiagonalLinearOperator, IdentityLinearOperator
from curvlinops_b200.estimators import hutchinson_diag, hutchinson_trace, hutchpp_trace, xdiag, xtrace
from curvlinops_b200.inverse import CGInverseLinearOperator, NeumannInverseLinearOperator


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.double().cpu()
    return ((got - want).norm() / want.norm()).item()


def check_damped_inverses(op, dense, shapes, preconditioner=None, delta_rel=0.1, tol=5e-3):
    """(op + delta I)^-1 by CG (plain, Jacobi- and optionally ``preconditioner``-preconditioned) and by a
    Jacobi-preconditioned Neumann series vs the dense float64 solve; fp32 operators reach ~cond * 1e-6."""
    dev, dt = op.device, op.dtype
    dense = dense.double().cpu()
    P = dense.shape[0]
    delta = delta_rel * dense.diag().mean().item()
    damped = op + delta * IdentityLinearOperator(shapes, dev, dt)
    X = torch.rand(P, 3, generator=torch.Generator().manual_seed(0)).to(device=dev, dtype=dt)
    want = torch.linalg.solve(dense + delta * torch.eye(P, dtype=torch.float64), X.double().cpu())
    jacobi = DiagonalLinearOperator([(dense.diag() + delta).reciprocal().to(device=dev, dtype=dt)])
    iters = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # fp32 products may stall above the tolerance
        for tag, pre in [("plain", None), ("jacobi", jacobi.__matmul__), ("given", preconditioner)]:
            if tag == "given" and pre is None:
                continue
            inv = CGInverseLinearOperator(damped, eps=0, tolerance=1e-4, max_iter=300, preconditioner=pre)
            got = inv @ X
            assert got.shape == X.shape and got.device == X.device
            err = rel_err(got, want)
            assert err < tol, f"CG[{tag}] relative error {err:.3e}"
            iters[tag] = inv.last_info[0]
        # list format in, list format out
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        Xl = [x.reshape(*s, 3) for x, s in zip(X.split(sizes), shapes)]
        Yl = CGInverseLinearOperator(damped, eps=0, tolerance=1e-4, max_iter=300) @ Xl
        assert [tuple(y.shape) for y in Yl] == [(*s, 3) for s in shapes]
        err = rel_err(torch.cat([y.reshape(-1, 3) for y in Yl]), want)
        assert err < tol, f"CG[list] relative error {err:.3e}"
    # truncated Neumann / Richardson series vs the same polynomial evaluated densely in float64
    Dd = dense + delta * torch.eye(P, dtype=torch.float64)
    alpha = 1.0 / (torch.linalg.eigvalsh(dense)[-1].item() + delta)
    term = X.double().cpu()
    series = term.clone()
    for _ in range(20):
        term = term - alpha * (Dd @ term)
        series += term
    Y = NeumannInverseLinearOperator(damped, num_terms=20, scale=alpha) @ X
    err = rel_err(Y, alpha * series)
    assert err < 1e-3, f"Neumann series relative error {err:.3e}"
    return iters


def check_estimators(op, dense, seed=0, rtol=2e-3):
    """Same seed -> same probes for the operator and for its dense matrix (a tensor on the same device), so
    the estimates must agree to the accuracy of the products."""
    dense = dense.to(device=op.device, dtype=op.dtype)
    scale = dense.diag().abs().max().item()
    for fn, n in [(hutchinson_trace, 6), (hutchpp_trace, 6), (xtrace, 6), (hutchinson_diag, 6), (xdiag, 6)]:
        torch.manual_seed(seed)
        got = fn(op, n)
        torch.manual_seed(seed)
        want = fn(dense, n)
        assert got.shape == want.shape
        atol = rtol * (scale if got.ndim else abs(want.item()))
        torch.testing.assert_close(got.double().cpu(), want.double().cpu(), rtol=rtol, atol=atol,
                                   msg=lambda m, f=fn: f"{f.__name__}: {m}")
