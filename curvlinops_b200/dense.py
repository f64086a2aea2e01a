"""Explicit (dense / diagonal / identity) operators used next to the matrix-free ones.

These are the small building blocks the reference's inverse and estimator tests combine with the curvature
operators: damping terms ``delta * I`` added to a GGN, Jacobi preconditioners, explicit matrices wrapped as
operators (reference ``curvlinops/diag.py:11-163``, ``curvlinops/examples/__init__.py:64-150,217-247``).  They
hold ordinary tensors on the operator's device; nothing here touches the engine.
"""

from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor

from .linop import PyTorchLinearOperator


def _single(values, what: str):
    found = set(values)
    if len(found) != 1:
        raise RuntimeError(f"Expected single {what}, got {found}.")
    return found.pop()


class TensorLinearOperator(PyTorchLinearOperator):
    """A 2-d tensor as a linear operator (reference ``examples/__init__.py:64-150``)."""

    def __init__(self, A: Tensor):
        if A.ndim != 2:
            raise ValueError(f"Input tensor must be 2D. Got {A.ndim}D.")
        super().__init__([(A.shape[1],)], [(A.shape[0],)])
        self._A = A
        self.SELF_ADJOINT = A.shape[0] == A.shape[1] and bool(torch.equal(A, A.conj().T))

    device = property(lambda self: self._A.device)
    dtype = property(lambda self: self._A.dtype)

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        (x,) = X
        return [self._A @ x]

    def _adjoint(self) -> "TensorLinearOperator":
        return TensorLinearOperator(self._A.conj().T)

    def trace(self) -> Tensor:
        return self._A.trace()

    def det(self) -> Tensor:
        return torch.linalg.det(self._A)

    def logdet(self) -> Tensor:
        return torch.logdet(self._A)

    def frobenius_norm(self) -> Tensor:
        return torch.linalg.matrix_norm(self._A)


class DiagonalLinearOperator(PyTorchLinearOperator):
    """Diagonal matrix given block-wise in tensor-list format (reference ``diag.py:11-163``).

    Sums, products and scalings of two diagonal operators stay diagonal, so a damped Jacobi preconditioner
    ``(D + delta I)^-1`` is one tensor list rather than an operator tree.
    """

    def __init__(self, diagonal: Sequence[Tensor]):
        diagonal = list(diagonal)
        shapes = [tuple(d.shape) for d in diagonal]
        super().__init__(shapes, shapes)
        self._diagonal = diagonal
        self.SELF_ADJOINT = all(not d.is_complex() or bool(torch.allclose(d.conj(), d)) for d in diagonal)

    @property
    def device(self) -> torch.device:
        return _single((d.device for d in self._diagonal), "device")

    @property
    def dtype(self) -> torch.dtype:
        return _single((d.dtype for d in self._diagonal), "dtype")

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        return [d.unsqueeze(-1) * x for d, x in zip(self._diagonal, X)]

    def _adjoint(self) -> "DiagonalLinearOperator":
        return DiagonalLinearOperator([d.conj() for d in self._diagonal])

    def inverse(self, damping: float) -> "DiagonalLinearOperator":
        """``(D + damping I)^-1``."""
        return DiagonalLinearOperator([1.0 / (d + damping) for d in self._diagonal])

    def _pairs_with(self, other) -> bool:
        return isinstance(other, DiagonalLinearOperator) and self._in_shape == other._in_shape

    def __add__(self, other):
        if self._pairs_with(other):
            return DiagonalLinearOperator([a + b for a, b in zip(self._diagonal, other._diagonal)])
        return super().__add__(other)

    def __matmul__(self, other):
        if self._pairs_with(other):
            return DiagonalLinearOperator([a * b for a, b in zip(self._diagonal, other._diagonal)])
        return super().__matmul__(other)

    def __mul__(self, scalar):
        return DiagonalLinearOperator([d * scalar for d in self._diagonal])

    __rmul__ = __mul__


class IdentityLinearOperator(DiagonalLinearOperator):
    """Identity on a tensor-list space; the diagonal is a stride-0 view of a single one per block
    (reference ``examples/__init__.py:217-247``)."""

    SELF_ADJOINT = True

    def __init__(self, shape: Sequence[Sequence[int]], device, dtype):
        ones = [torch.ones((1,) * len(s), device=device, dtype=dtype).expand(*s) if len(s)
                else torch.ones((), device=device, dtype=dtype) for s in shape]
        super().__init__(ones)
        self.SELF_ADJOINT = True

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        return X
