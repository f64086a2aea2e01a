"""Matrix-free inverses on dense CPU operators.

* Neumann: parity with the reference's ``NeumannInverseLinearOperator`` on seeded inputs
  (``tests/golden/estimators.npz``, from ``oracle/make_golden_estimators.py``) and the reference's toy cases
  (``test/test_inverse.py:169-262``).
* CG / LSMR: product with the inverse vs the dense inverse, as the reference tests them
  (``test/test_inverse.py:29-94``); CG parity with GPyTorch's routine itself is unpinned (dependency absent).
"""
import os
import warnings

import numpy as np
import pytest
import torch

from curvlinops_b200.dense import DiagonalLinearOperator, IdentityLinearOperator, TensorLinearOperator
from curvlinops_b200.inverse import (CGInverseLinearOperator, LSMRInverseLinearOperator,
                                     NeumannInverseLinearOperator, batched_cg)
from curvlinops_b200.linop import PyTorchLinearOperator

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimators.npz"))
f64 = torch.float64


def _spd(n, cond=1e3, seed=0, dtype=f64):
    g = torch.Generator().manual_seed(seed)
    Q, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=f64))
    d = torch.logspace(0, -np.log10(cond), n, dtype=f64)
    return ((Q * d) @ Q.T).to(dtype)


class TwoBlock(PyTorchLinearOperator):
    """Dense SPD matrix acting on a two-tensor parameter space ([3, 4] and [5])."""

    SELF_ADJOINT = True

    def __init__(self, M):
        super().__init__([(3, 4), (5,)], [(3, 4), (5,)])
        self.M = M

    device = property(lambda self: self.M.device)
    dtype = property(lambda self: self.M.dtype)

    def _matmat(self, X):
        K = X[0].shape[-1]
        Y = self.M @ torch.cat([x.reshape(-1, K) for x in X])
        return [Y[:12].reshape(3, 4, K), Y[12:].reshape(5, K)]


# ---- Neumann -------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["neumann|plain|30", "neumann|scale0.7|50", "neumann|jacobi|25"])
def test_neumann_matches_reference(key):
    S = torch.from_numpy(GOLD["neumann_S"])
    rhs = torch.from_numpy(GOLD["neumann_rhs"])
    _, variant, terms = key.split("|")
    kw = {"num_terms": int(terms)}
    if variant.startswith("scale"):
        kw["scale"] = float(variant[5:])
    if variant == "jacobi":
        kw["preconditioner"] = DiagonalLinearOperator([S.diag().reciprocal()]).__matmul__
    got = NeumannInverseLinearOperator(TensorLinearOperator(S), **kw) @ rhs
    np.testing.assert_allclose(got.numpy(), GOLD[key], rtol=1e-10, atol=1e-12)


def test_neumann_toy_and_preconditioners():
    """The reference's toy cases: divergence is reported, scaling / Jacobi / Gauss-Seidel converge."""
    A = torch.tensor([[5.0, 1.0, 1.0], [1.0, 4.0, 1.0], [1.0, 1.0, 3.0]], dtype=f64)
    op, inv, theta = TensorLinearOperator(A), torch.linalg.inv(A), 0.3
    x = torch.eye(3, dtype=f64)
    with pytest.raises(ValueError, match="Detected NaNs after application of"):
        NeumannInverseLinearOperator(op, num_terms=1000) @ x
    tols = dict(rtol=1e-3, atol=1e-5)
    assert not torch.allclose(NeumannInverseLinearOperator(op, num_terms=20, scale=theta) @ x, inv, **tols)
    torch.testing.assert_close(NeumannInverseLinearOperator(op, num_terms=100, scale=theta) @ x, inv, **tols)
    richardson = IdentityLinearOperator(op._in_shape, A.device, A.dtype) * theta
    jacobi = DiagonalLinearOperator([A.diag().reciprocal()])
    gauss_seidel = TensorLinearOperator(torch.linalg.inv(A.tril()))
    for terms, P in [(100, richardson), (20, jacobi), (20, gauss_seidel)]:
        N = NeumannInverseLinearOperator(op, num_terms=terms, preconditioner=P.__matmul__)
        torch.testing.assert_close(N @ x, inv, **tols)
        torch.testing.assert_close(N @ x[:, 0], inv[:, 0], **tols)
        # X @ A^-1 goes through the adjoint series (adjoint preconditioner of a bound __matmul__)
        torch.testing.assert_close(x @ N, inv, **tols)
    with pytest.raises(NotImplementedError, match="bound PyTorchLinearOperator.__matmul__"):
        NeumannInverseLinearOperator(op, preconditioner=lambda v: v).adjoint()

    B = torch.tensor([[0.0, 1 / 2, 1 / 4], [5 / 7, 0.0, 1 / 7], [3 / 10, 3 / 5, 0.0]], dtype=f64) + torch.eye(3, dtype=f64)
    torch.testing.assert_close(NeumannInverseLinearOperator(TensorLinearOperator(B), num_terms=1000) @ x,
                               torch.linalg.inv(B), **tols)


# ---- CG ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precondition", [False, True], ids=["", "jacobi"])
def test_cg_inverse_matches_dense_inverse(precondition):
    M = _spd(60, cond=1e3)
    A = TensorLinearOperator(M)
    P = DiagonalLinearOperator([M.diag().reciprocal()]).__matmul__ if precondition else None
    inv_op = CGInverseLinearOperator(A, eps=0, tolerance=1e-10, preconditioner=P)
    inv = torch.linalg.inv(M)
    X = torch.randn(60, 5, dtype=f64, generator=torch.Generator().manual_seed(3))
    torch.testing.assert_close(inv_op @ X, inv @ X, rtol=5e-5, atol=5e-8)
    torch.testing.assert_close(inv_op @ X[:, 0], inv @ X[:, 0], rtol=5e-5, atol=5e-8)
    torch.testing.assert_close(X.T @ inv_op, X.T @ inv, rtol=5e-5, atol=5e-8)
    torch.testing.assert_close(inv_op @ X, inv_op @ X, rtol=0, atol=0)  # consecutive products agree
    its, res = inv_op.last_info
    assert 10 <= its <= 1000 and res < 1e-10


def test_cg_list_format_and_damped_sum():
    """Tensor-list operator + damping * identity, the combination the reference inverts (GGN + delta I)."""
    M = _spd(17, cond=1e4, seed=5)
    delta = 1e-2
    op = TwoBlock(M) + delta * IdentityLinearOperator([(3, 4), (5,)], M.device, M.dtype)
    inv_op = CGInverseLinearOperator(op, tolerance=1e-9, eps=0)
    g = torch.Generator().manual_seed(0)
    X = [torch.randn(3, 4, 2, dtype=f64, generator=g), torch.randn(5, 2, dtype=f64, generator=g)]
    Y = inv_op @ X
    assert [tuple(y.shape) for y in Y] == [(3, 4, 2), (5, 2)]
    flat = torch.cat([x.reshape(-1, 2) for x in X])
    want = torch.linalg.solve(M + delta * torch.eye(17, dtype=f64), flat)
    torch.testing.assert_close(torch.cat([y.reshape(-1, 2) for y in Y]), want, rtol=1e-6, atol=1e-9)


def test_cg_exact_preconditioner_zero_columns_initial_guess():
    M = _spd(30, cond=1e2, seed=2)
    inv = torch.linalg.inv(M)
    B = torch.randn(30, 3, dtype=f64, generator=torch.Generator().manual_seed(1))
    B[:, 1] = 0.0  # a zero right-hand side stays zero (no 0/0)
    X, (its, res) = batched_cg(lambda V: M @ V, B, tolerance=1e-12, preconditioner=lambda R: inv @ R,
                               return_info=True)
    torch.testing.assert_close(X, inv @ B, rtol=1e-9, atol=1e-12)
    assert bool((X[:, 1] == 0).all())
    assert its <= 11  # exact preconditioner: converged at once, loop leaves at the first permitted test
    # the exact solution as initial guess: nothing to do
    X2, (its2, _) = batched_cg(lambda V: M @ V, B, tolerance=1e-12, initial_guess=inv @ B, return_info=True,
                               stop_updating_after=1e-8)
    torch.testing.assert_close(X2, inv @ B, rtol=1e-9, atol=1e-12)
    assert its2 == 0


def test_cg_float32_accuracy_level():
    """What the fp32 engine can expect: relative error ~ cond * eps_fp32."""
    M = _spd(200, cond=1e3, seed=4, dtype=torch.float32)
    X = torch.randn(200, 8, generator=torch.Generator().manual_seed(1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        Y = CGInverseLinearOperator(TensorLinearOperator(M), tolerance=1e-5, max_iter=400) @ X
    want = torch.linalg.solve(M.double(), X.double())
    assert ((Y.double() - want).norm() / want.norm()) < 5e-3


def test_cg_warns_when_not_converged_and_rejects_unknown_arguments():
    M = _spd(40, cond=1e6, seed=7)
    with pytest.warns(RuntimeWarning, match="CG terminated in 3 iterations"):
        CGInverseLinearOperator(TensorLinearOperator(M), max_iter=3, tolerance=1e-12) @ torch.ones(40, dtype=f64)
    with pytest.raises(TypeError, match="Unknown CG hyperparameter"):
        CGInverseLinearOperator(TensorLinearOperator(M), maxiter=3)
    with pytest.raises(NotImplementedError, match="n_tridiag"):
        CGInverseLinearOperator(TensorLinearOperator(M), n_tridiag=2)
    with pytest.raises(ValueError, match="must be square to form an inverse"):
        CGInverseLinearOperator(TensorLinearOperator(torch.zeros(3, 4, dtype=f64)))
    # reference docstring example (inverse.py:84-106)
    A = torch.tensor([[4.0, 1.0, 0.0], [1.0, 3.0, 1.0], [0.0, 1.0, 2.0]])
    b = torch.tensor([1.0, 2.0, 3.0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        x = CGInverseLinearOperator(TensorLinearOperator(A), max_iter=3, max_tridiag_iter=3, tolerance=1e-7) @ b
        xp = CGInverseLinearOperator(TensorLinearOperator(A), max_iter=3, max_tridiag_iter=3, tolerance=1e-7,
                                     preconditioner=DiagonalLinearOperator([A.diag().reciprocal()]).__matmul__) @ b
    torch.testing.assert_close(x, torch.linalg.solve(A, b), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xp, torch.linalg.solve(A, b), rtol=1e-4, atol=1e-5)


# ---- LSMR ----------------------------------------------------------------------------------------
def test_lsmr_inverse_matches_dense_inverse():
    M = _spd(25, cond=1e2, seed=9) + 0.3 * torch.randn(25, 25, dtype=f64,
                                                         generator=torch.Generator().manual_seed(4)) / 25
    inv_op = LSMRInverseLinearOperator(TensorLinearOperator(M), atol=0, btol=0, maxiter=100)
    X = torch.randn(25, 2, dtype=f64, generator=torch.Generator().manual_seed(5))
    torch.testing.assert_close(inv_op @ X, torch.linalg.solve(M, X), rtol=1e-6, atol=1e-9)
    assert len(inv_op._lsmr_info) == 2
    torch.testing.assert_close(X.T @ inv_op, X.T @ torch.linalg.inv(M), rtol=1e-6, atol=1e-9)


# ---- helper operators ----------------------------------------------------------------------------
def test_dense_helper_operators():
    d = [torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=f64), torch.tensor([5.0], dtype=f64)]
    D = DiagonalLinearOperator(d)
    assert D.shape == (5, 5) and D.SELF_ADJOINT and D.dtype == f64 and D.device.type == "cpu"
    x = torch.arange(5, dtype=f64)
    torch.testing.assert_close(D @ x, torch.tensor([0.0, 2.0, 6.0, 12.0, 20.0], dtype=f64))
    assert isinstance(D + D, DiagonalLinearOperator) and isinstance(D @ D, DiagonalLinearOperator)
    torch.testing.assert_close((D @ D) @ x, (D @ (D @ x)))
    torch.testing.assert_close((2 * D) @ x, 2 * (D @ x))
    torch.testing.assert_close(D.inverse(0.5) @ x, x / (torch.tensor([1.0, 2, 3, 4, 5], dtype=f64) + 0.5))
    eye = IdentityLinearOperator([(2, 2), (1,)], "cpu", f64)
    torch.testing.assert_close(eye @ x, x)
    torch.testing.assert_close((D + 0.5 * eye) @ x, D @ x + 0.5 * x)
    T = TensorLinearOperator(torch.arange(6, dtype=f64).reshape(2, 3))
    assert T.shape == (2, 3) and not T.SELF_ADJOINT
    torch.testing.assert_close(T.adjoint() @ torch.ones(2, dtype=f64), torch.tensor([3.0, 5.0, 7.0], dtype=f64))
    with pytest.raises(ValueError, match="must be 2D"):
        TensorLinearOperator(torch.zeros(3))
    with pytest.raises(RuntimeError, match="Expected single dtype"):
        DiagonalLinearOperator([torch.zeros(2), torch.zeros(2, dtype=f64)]).dtype


# ---- the checks the GPU suite runs on the engine's GGN, here on a dense fp32 operator ------------------
def test_consumer_checks_on_dense_fp32_operator():
    from tests.consumer_checks import check_damped_inverses, check_estimators

    M = _spd(17, cond=1e3, seed=11, dtype=torch.float32)
    op = TwoBlock(M)
    exact = TensorLinearOperator(torch.linalg.inv(M.double() + 0.1 * M.double().diag().mean()
                                                  * torch.eye(17, dtype=f64)).float())
    iters = check_damped_inverses(op, M, [(3, 4), (5,)], preconditioner=exact.__matmul__)
    assert iters["given"] <= iters["plain"]
    check_estimators(op, M)
