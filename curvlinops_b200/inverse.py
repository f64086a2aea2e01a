"""Matrix-free inverses on top of the engine's products: conjugate gradients, truncated Neumann series, LSMR.

SURVEY §8(f) row 3: the main consumer of fast GGN products and of the KFAC inverse (as preconditioner).
Interface of the reference's ``curvlinops/inverse.py`` (``CGInverseLinearOperator`` :54-141,
``LSMRInverseLinearOperator`` :144-211, ``NeumannInverseLinearOperator`` :214-391): same constructors, same
errors, results in the caller's format.

Everything iterates on ONE flat ``[P, K]`` matrix (the engine's native layout, K minor) that stays on the
operator's device; all ``K`` right-hand sides advance together, so every iteration is a single batched
engine product.  Per-column step sizes live in ``[1, K]`` device tensors; the only host read per CG
iteration is the stopping test.

CG: the reference delegates to ``linear_operator.utils.linear_cg`` (GPyTorch's batched preconditioned CG;
dependency ``linear-operator>=0.2.0``, unpinned and NOT vendored under the reference tree, not installed
here), so this is a restatement of that routine's published behaviour, **parity unpinned** against its
code: right-hand sides are normalised per column, converged columns are frozen (``stop_updating_after``),
divisions are guarded by ``eps`` (a column whose ``pᵀAp`` falls below ``eps`` takes no further steps, which
caps the attainable accuracy — the reference's tests pass ``eps=0`` for accurate solves), and the loop stops
once, after at least 10 iterations, the mean residual norm of the normalised systems is below ``tolerance``.  The parity anchor is the reference's own tests of
this class (``test/test_inverse.py:29-166``: product with the inverse vs the dense inverse), which
``tests/test_inverse_cpu.py`` / ``tests/test_gpu_z_consumers.py`` re-run.
"""

from __future__ import annotations

import warnings
from typing import Callable

import numpy
import torch
from torch import Tensor

from .linop import PyTorchLinearOperator


class _InverseLinearOperator(PyTorchLinearOperator):
    """Inverse of a square operator; device / dtype follow the operator (reference ``inverse.py:15-51``)."""

    def __init__(self, A: PyTorchLinearOperator):
        if A._in_shape != A._out_shape:
            raise ValueError("Input linear operator must be square to form an inverse."
                             + f"Got {A._in_shape} != {A._out_shape}.")
        super().__init__(A._in_shape, A._out_shape)
        self._A = A

    device = property(lambda self: self._A.device)
    dtype = property(lambda self: self._A.dtype)

    # flat [P, K] <-> list format
    @staticmethod
    def _flatten(X: list[Tensor]) -> Tensor:
        return torch.cat([x.flatten(end_dim=-2) for x in X])

    def _unflatten(self, Y: Tensor) -> list[Tensor]:
        K = Y.shape[1]
        return [y.reshape(*s, K) for y, s in zip(Y.split(self._out_shape_flat), self._out_shape)]


_CG_DEFAULTS = dict(max_iter=1000, tolerance=1.0, eps=1e-10, stop_updating_after=1e-10,
                    max_tridiag_iter=20, n_tridiag=0, initial_guess=None, preconditioner=None)


def batched_cg(matmul: Callable[[Tensor], Tensor], rhs: Tensor, *, max_iter: int = 1000,
               tolerance: float = 1.0, eps: float = 1e-10, stop_updating_after: float = 1e-10,
               initial_guess: Tensor | None = None,
               preconditioner: Callable[[Tensor], Tensor] | None = None,
               return_info: bool = False):
    """Solve ``A x_k = b_k`` for all columns of ``rhs`` (``[P, K]``) with preconditioned CG.

    ``matmul`` applies the symmetric positive-definite operator to a ``[P, K]`` matrix, ``preconditioner``
    (optional) an approximation of its inverse.  See the module docstring for the stopping rules.  With
    ``return_info`` also returns ``(iterations, mean residual norm of the normalised systems)``.
    """
    if rhs.ndim != 2:
        raise ValueError(f"rhs must be a [P, K] matrix. Got {tuple(rhs.shape)}.")
    if not torch.is_floating_point(rhs):
        raise ValueError(f"rhs must be a floating point tensor. Got {rhs.dtype}.")
    precond = preconditioner if preconditioner is not None else (lambda R: R)

    # normalise every system; columns that are (numerically) zero keep norm 1 and solve to zero
    b_norm = torch.linalg.vector_norm(rhs, dim=0, keepdim=True)
    b_zero = b_norm <= eps
    b_norm = b_norm.masked_fill(b_zero, 1.0)
    B = rhs / b_norm

    if initial_guess is None:
        x = torch.zeros_like(B)
        r = B.clone()
    else:
        if initial_guess.shape != rhs.shape:
            raise ValueError(f"initial_guess must have shape {tuple(rhs.shape)}. Got {tuple(initial_guess.shape)}.")
        x = initial_guess / b_norm
        r = B - matmul(x)
    if bool(torch.isnan(r).any()):
        raise RuntimeError("NaNs encountered when trying to perform matrix-vector multiplication")

    r_norm = torch.linalg.vector_norm(r, dim=0, keepdim=True)
    frozen = r_norm < stop_updating_after
    iters, mean_res, reached = 0, float(r_norm.mean()), False
    if not bool(frozen.all()) and max_iter > 0:
        z = precond(r)
        p = z.clone()
        rz = (r * z).sum(0, keepdim=True)
        min_iters = min(10, max_iter - 1)
        for k in range(max_iter):
            Ap = matmul(p)
            pAp = (p * Ap).sum(0, keepdim=True)
            tiny = pAp <= eps  # breakdown / non-positive curvature: no step for this column
            alpha = (rz / pAp.masked_fill(tiny, 1.0)).masked_fill(tiny | frozen, 0.0)
            x.addcmul_(p, alpha)
            r.addcmul_(Ap, alpha, value=-1.0)
            z = precond(r)
            rz_new = (r * z).sum(0, keepdim=True)
            small = rz <= eps
            beta = (rz_new / rz.masked_fill(small, 1.0)).masked_fill(small, 0.0)
            p.mul_(beta).add_(z)
            rz = rz_new

            r_norm = torch.linalg.vector_norm(r, dim=0, keepdim=True).masked_fill(b_zero, 0.0)
            frozen = r_norm < stop_updating_after
            iters = k + 1
            if k >= min_iters:
                mean_res = float(r_norm.mean())  # the one host read of the iteration
                if mean_res < tolerance:
                    reached = True
                    break
        if not reached:
            mean_res = float(r_norm.mean())
            warnings.warn(f"CG terminated in {iters} iterations with average residual norm {mean_res} which is "
                          f"larger than the tolerance of {tolerance}. Consider raising max_iter or a "
                          "preconditioner.", RuntimeWarning, stacklevel=2)
    x = x * b_norm
    return (x, (iters, mean_res)) if return_info else x


class CGInverseLinearOperator(_InverseLinearOperator):
    """``A^-1`` by batched preconditioned conjugate gradients; ``A`` symmetric positive definite.

    Keyword arguments (reference ``inverse.py:69-112``): ``max_iter`` (1000), ``tolerance`` (1.0, on the mean
    residual norm of the normalised systems; the reference inherits GPyTorch's training default, pass a small
    value for an accurate solve), ``eps`` (1e-10), ``stop_updating_after`` (1e-10), ``initial_guess``
    (flat ``[P, K]``), ``preconditioner`` (callable on flat matrices, e.g. ``KFAC.inverse(...).__matmul__``).
    ``max_tridiag_iter`` is accepted for compatibility (Lanczos tridiagonalisation is not produced),
    ``n_tridiag`` must be 0.
    """

    def __init__(self, A: PyTorchLinearOperator, **cg_hyperparameters):
        super().__init__(A)
        unknown = set(cg_hyperparameters) - set(_CG_DEFAULTS)
        if unknown:
            raise TypeError(f"Unknown CG hyperparameter(s) {sorted(unknown)}. Supported: {sorted(_CG_DEFAULTS)}.")
        if cg_hyperparameters.get("n_tridiag", 0):
            raise NotImplementedError("n_tridiag > 0 (tridiagonal matrices from CG) is not supported.")
        self._cg_hyperparameters = cg_hyperparameters
        self.SELF_ADJOINT = A.SELF_ADJOINT
        self.last_info: tuple[int, float] | None = None  # (iterations, mean residual) of the last product

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        hp = {k: v for k, v in self._cg_hyperparameters.items() if k not in ("max_tridiag_iter", "n_tridiag")}
        Y, self.last_info = batched_cg(self._A.__matmul__, self._flatten(X), return_info=True, **hp)
        return self._unflatten(Y)

    def _adjoint(self) -> "CGInverseLinearOperator":
        return CGInverseLinearOperator(self._A.adjoint(), **self._cg_hyperparameters)


class NeumannInverseLinearOperator(_InverseLinearOperator):
    r"""``A^-1 ≈ α Σ_{k=0}^{K} (I − α P A)^k P`` (truncated Neumann / preconditioned Richardson iteration).

    Converges iff the eigenvalues of ``α P A`` lie in (0, 2) (reference ``inverse.py:214-391``).
    ``preconditioner`` acts on flat ``[P, K]`` matrices (e.g. a bound ``PyTorchLinearOperator.__matmul__``).
    """

    def __init__(self, A: PyTorchLinearOperator, num_terms: int = 100, scale: float = 1.0,
                 check_nan: bool = True, preconditioner: Callable[[Tensor], Tensor] | None = None):
        super().__init__(A)
        self._num_terms = num_terms
        self._scale = scale
        self._check_nan = check_nan
        self._preconditioner = preconditioner
        self.SELF_ADJOINT = A.SELF_ADJOINT and preconditioner is None

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        P = self._preconditioner
        B = self._flatten(X)
        if P is not None:
            B = P(B)
        acc = B.clone()   # running sum of the series
        term = B.clone()  # (I - alpha P A)^k P b
        for idx in range(self._num_terms):
            step = self._A @ term
            if P is not None:
                step = P(step)
            term.sub_(step, alpha=self._scale)
            acc.add_(term)
            if self._check_nan and bool(torch.isnan(acc).any()):
                raise ValueError(f"Detected NaNs after application of {idx}-th term."
                                 + " This is probably because the Neumann series is non-convergent."
                                 + " Try decreasing `scale` and read the comment on convergence.")
        return self._unflatten(acc.mul_(self._scale))

    def _adjoint(self) -> "NeumannInverseLinearOperator":
        preconditioner = None
        if self._preconditioner is not None:
            owner = getattr(self._preconditioner, "__self__", None)
            if not isinstance(owner, PyTorchLinearOperator):
                raise NotImplementedError("Adjoint with a preconditioner is only supported when the "
                                          "preconditioner is a bound PyTorchLinearOperator.__matmul__ method.")
            preconditioner = owner.adjoint().__matmul__
        return NeumannInverseLinearOperator(self._A.adjoint(), num_terms=self._num_terms, scale=self._scale,
                                            check_nan=self._check_nan, preconditioner=preconditioner)


class LSMRInverseLinearOperator(_InverseLinearOperator):
    """``A^-1`` by SciPy's LSMR, one column at a time over the operator's SciPy bridge.

    As in the reference (``inverse.py:144-211``) the Krylov recurrences run in SciPy on the host; each of
    their matrix-vector products is an engine product through ``to_scipy()``'s pinned staging buffers.
    ``_lsmr_info`` holds SciPy's diagnostics (istop, itn, normr, ...) of the last product.
    """

    def __init__(self, A: PyTorchLinearOperator, **lsmr_hyperparameters):
        super().__init__(A)
        self._A_scipy = A.to_scipy()
        self._lsmr_hyperparameters = lsmr_hyperparameters
        self.SELF_ADJOINT = A.SELF_ADJOINT

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        from scipy.sparse.linalg import lsmr

        B = self._flatten(X).cpu().numpy().astype(self._A_scipy.dtype)
        solved = [lsmr(self._A_scipy, b, **self._lsmr_hyperparameters) for b in B.T]
        self._lsmr_info = [s[1:] for s in solved]
        Y = torch.as_tensor(numpy.column_stack([s[0] for s in solved]), device=self.device, dtype=self.dtype)
        return self._unflatten(Y)

    def _adjoint(self) -> "LSMRInverseLinearOperator":
        return LSMRInverseLinearOperator(self._A.adjoint(), **self._lsmr_hyperparameters)
