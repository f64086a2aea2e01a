"""Matrix-free linear operator interface (tensor / tensor-list / NumPy formats, sum, scale, chain).

Behavioural mirror of the reference's ``PyTorchLinearOperator`` and its three composition classes
(``curvlinops/_torch_base.py:33-814``): same public methods, same accepted input formats, same
``ValueError`` messages (the reference's tests regex-match them,
``test/test__torch_base.py:44-102,256-260``).  Format rules:

* ``A @ X``: ``X`` is a tensor ``[N]`` / ``[N, K]`` or a list ``[*N_i]`` / ``[*N_i, K]`` (trailing ``K``),
* ``X @ A``: ``X`` is ``[M]`` / ``[K, M]`` or a list ``[*M_i]`` / ``[K, *M_i]`` (leading ``K``), evaluated
  through the adjoint,
* the result comes back in the caller's format.
"""

from __future__ import annotations

from typing import Callable, Iterator, Sequence

import numpy
import torch
from scipy.sparse.linalg import LinearOperator as _ScipyLinearOperator
from torch import Size, Tensor


def report_allclose(t1, t2, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """``allclose`` that prints a short mismatch summary (role of ``utils.py:173-215``)."""
    t1 = torch.as_tensor(t1)
    t2 = torch.as_tensor(t2, device=t1.device)
    ok = bool(torch.allclose(t1, t2, rtol=rtol, atol=atol))
    if not ok:
        bad = ~torch.isclose(t1, t2, rtol=rtol, atol=atol)
        diff = (t1 - t2).abs()
        print(f"allclose failed: {int(bad.sum())}/{t1.numel()} entries differ "
              f"(max abs err {diff.max().item():.3e}, max |a| {t1.abs().max().item():.3e}, "
              f"max |b| {t2.abs().max().item():.3e}; rtol={rtol}, atol={atol}).")
    return ok


def _same_spaces(a: "PyTorchLinearOperator", b: "PyTorchLinearOperator") -> None:
    if a._in_shape != b._in_shape or a._out_shape != b._out_shape:
        raise ValueError(
            f"Shape mismatch: expected in_shape={a._in_shape}, out_shape={a._out_shape}, "
            f"got in_shape={b._in_shape}, out_shape={b._out_shape}."
        )


def _same_device_dtype(a, b) -> None:
    if a.device != b.device:
        raise ValueError(f"Device mismatch: expected {a.device}, got {b.device}.")
    if a.dtype != b.dtype:
        raise ValueError(f"Dtype mismatch: expected {a.dtype}, got {b.dtype}.")


class PyTorchLinearOperator:
    """Base class: subclasses implement ``_matmat`` (list in, list out, trailing ``K``) and, unless
    ``SELF_ADJOINT``, ``_adjoint``; ``device`` / ``dtype`` are needed for ``to_scipy``."""

    SELF_ADJOINT: bool = False

    def __init__(self, in_shape: Sequence[Sequence[int]], out_shape: Sequence[Sequence[int]]):
        if not in_shape or not out_shape:
            raise ValueError(f"In- {in_shape} and output shapes {out_shape} must be non-empty.")
        self._in_shape = [Size(s) for s in in_shape]
        self._out_shape = [Size(s) for s in out_shape]
        self._in_shape_flat = [s.numel() for s in self._in_shape]
        self._out_shape_flat = [s.numel() for s in self._out_shape]
        self.shape = (sum(self._out_shape_flat), sum(self._in_shape_flat))

    # ---- to be provided by subclasses ----------------------------------------------------------
    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        raise NotImplementedError

    def _adjoint(self) -> "PyTorchLinearOperator":
        raise NotImplementedError

    @property
    def device(self) -> torch.device:
        raise NotImplementedError

    @property
    def dtype(self) -> torch.dtype:
        raise NotImplementedError

    def adjoint(self) -> "PyTorchLinearOperator":
        return self if self.SELF_ADJOINT else self._adjoint()

    # ---- format handling -----------------------------------------------------------------------
    @staticmethod
    def _to_list(X, shapes: list[Size], leading: bool):
        """-> (list with explicit K axis, was_list, was_vector, K)."""
        total = sum(s.numel() for s in shapes)
        if isinstance(X, Tensor):
            fixed = -1 if leading else 0
            if X.ndim not in (1, 2) or X.shape[fixed] != total:
                want = f"({total},) or " + (f"(K, {total})" if leading else f"({total}, K)")
                raise ValueError(f"Input tensor must have shape {want}, with K arbitrary. Got {X.shape}.")
            vec = X.ndim == 1
            K = 1 if vec else X.shape[0 if leading else 1]
            parts = X.split([s.numel() for s in shapes], dim=fixed)
            out = [p.reshape(K, *s) if leading else p.reshape(*s, K) for p, s in zip(parts, shapes)]
            return out, False, vec, K
        if isinstance(X, list) and all(isinstance(x, Tensor) for x in X):
            if len(X) != len(shapes):
                raise ValueError(f"Input list must have {len(shapes)} tensors. Got {len(X)}.")
            kdim = 0 if leading else -1
            if all(x.shape == s for x, s in zip(X, shapes)):
                return [x.unsqueeze(kdim) for x in X], True, True, 1
            inner = [(x.shape[1:] if leading else x.shape[:-1]) if x.ndim == len(s) + 1 else None
                     for x, s in zip(X, shapes)]
            ks = {x.shape[kdim] for x in X if x.ndim > 0}
            if all(i == s for i, s in zip(inner, shapes)) and len(ks) == 1:
                return list(X), True, False, ks.pop()
            raise ValueError(
                f"Input list must contain tensors with shapes {shapes} and optional "
                f"{'leading' if leading else 'trailing'} dimension for the matrix columns. "
                f"Got {[x.shape for x in X]}."
            )
        raise ValueError(f"Input must be tensor or list of tensors. Got {type(X)}.")

    @staticmethod
    def _from_list(Y: list[Tensor], shapes: list[Size], leading: bool, was_list: bool, was_vec: bool,
                   K: int):
        if len(Y) != len(shapes):
            raise ValueError(f"Output tensor list must have {len(shapes)} tensors. Got {len(Y)}.")
        want = [(K, *s) if leading else (*s, K) for s in shapes]
        if any(tuple(y.shape) != w for y, w in zip(Y, want)):
            raise ValueError(
                f"Output tensors must have shapes {shapes} and additional "
                f"{'leading' if leading else 'trailing'} dimension of {K}. Got {[y.shape for y in Y]}."
            )
        kdim = 0 if leading else -1
        if was_list:
            return [y.squeeze(kdim) for y in Y] if was_vec else Y
        flat = torch.cat([y.reshape(K, s.numel()) if leading else y.reshape(s.numel(), K)
                          for y, s in zip(Y, shapes)], dim=1 if leading else 0)
        return flat.squeeze(kdim) if was_vec else flat

    # ---- products ------------------------------------------------------------------------------
    def __matmul__(self, X):
        if isinstance(X, PyTorchLinearOperator):
            lhs = tuple(self) if isinstance(self, _ChainPyTorchLinearOperator) else (self,)
            rhs = tuple(X) if isinstance(X, _ChainPyTorchLinearOperator) else (X,)
            return _ChainPyTorchLinearOperator(*lhs, *rhs)
        flat = getattr(self, "_matmat_flat", None)
        if (flat is not None and isinstance(X, Tensor) and X.ndim in (1, 2) and X.shape[0] == self.shape[1]
                and X.device == self.device and X.dtype == self.dtype):
            # operators that work on the flat [P, K] layout skip the per-parameter split / re-concatenation
            Y = flat(X if X.ndim == 2 else X.unsqueeze(1))
            return Y.squeeze(1) if X.ndim == 1 else Y
        Xl, was_list, was_vec, K = self._to_list(X, self._in_shape, leading=False)
        return self._from_list(self._matmat(Xl), self._out_shape, False, was_list, was_vec, K)

    def __rmatmul__(self, X):
        Xl, was_list, was_vec, K = self._to_list(X, self._out_shape, leading=True)
        # X @ A = (A^H @ X^H)^H
        XH = [x.conj().movedim(0, -1) for x in Xl]
        YH = self.adjoint()._matmat(XH)
        Y = [y.conj().movedim(-1, 0) for y in YH]
        return self._from_list(Y, self._in_shape, True, was_list, was_vec, K)

    # ---- algebra -------------------------------------------------------------------------------
    def __add__(self, other):
        return _SumPyTorchLinearOperator(self, other)

    def __sub__(self, other):
        return self + (-1.0 * other)

    def __mul__(self, scalar):
        return _ScalePyTorchLinearOperator(self, scalar)

    __rmul__ = __mul__

    def __truediv__(self, scalar):
        return self * (1.0 / scalar)

    # ---- SciPy bridge (reference _torch_base.py:491-516, 560-592) ---------------------------------
    def to_scipy(self, dtype=None) -> _ScipyLinearOperator:
        fwd = self._numpy_bridge(self.__matmul__)
        AH = self.adjoint()
        bwd = AH._numpy_bridge(AH.__matmul__)
        return _ScipyLinearOperator(self.shape, matvec=fwd, rmatvec=bwd, matmat=fwd, rmatmat=bwd,
                                    dtype=numpy.dtype(dtype) if dtype is None else dtype)

    def _numpy_bridge(self, f: Callable[[Tensor], Tensor]):
        dev, dt = self.device, self.dtype

        staging: dict = {}  # pinned host buffers, reused across matvecs (ARPACK calls this in a loop)

        def call(X: numpy.ndarray) -> numpy.ndarray:
            Y = f(torch.as_tensor(X, dtype=dt, device=dev))
            if Y.dtype == torch.bfloat16:  # NumPy has no bf16
                Y = Y.float()
            Y = Y.detach()
            if Y.device.type == "cuda":
                key = (tuple(Y.shape), Y.dtype)
                if key not in staging:
                    staging.clear()
                    staging[key] = torch.empty(Y.shape, dtype=Y.dtype, pin_memory=True)
                host = staging[key]
                host.copy_(Y, non_blocking=True)
                torch.cuda.current_stream(Y.device).synchronize()
                return host.numpy().astype(X.dtype)  # astype copies out of the staging buffer
            return Y.cpu().numpy().astype(X.dtype)

        return call

    def _check_deterministic_matvec(self, rtol: float = 1e-5, atol: float = 1e-8):
        """Two products with the same vector must agree (reference ``_torch_base.py:542-558``)."""
        v = torch.rand(self.shape[1], device=self.device, dtype=self.dtype)
        if not report_allclose(self @ v, self @ v, rtol=rtol, atol=atol):
            raise RuntimeError("Check for deterministic matvec failed.")


class _SumPyTorchLinearOperator(PyTorchLinearOperator):
    """``A + B``."""

    def __init__(self, A: PyTorchLinearOperator, B: PyTorchLinearOperator):
        _same_spaces(A, B)
        _same_device_dtype(A, B)
        super().__init__(A._in_shape, A._out_shape)
        self._A, self._B = A, B
        self.SELF_ADJOINT = A.SELF_ADJOINT and B.SELF_ADJOINT

    def _matmat(self, X):
        return [a + b for a, b in zip(self._A._matmat(X), self._B._matmat(X))]

    def _adjoint(self):
        return _SumPyTorchLinearOperator(self._A.adjoint(), self._B.adjoint())

    device = property(lambda self: self._A.device)
    dtype = property(lambda self: self._A.dtype)


class _ScalePyTorchLinearOperator(PyTorchLinearOperator):
    """``s * A``."""

    def __init__(self, A: PyTorchLinearOperator, scalar):
        super().__init__(A._in_shape, A._out_shape)
        self._A, self._scalar = A, scalar
        self.SELF_ADJOINT = A.SELF_ADJOINT

    def _matmat(self, X):
        return [self._scalar * y for y in self._A._matmat(X)]

    def _adjoint(self):
        return _ScalePyTorchLinearOperator(self._A.adjoint(), self._scalar)

    device = property(lambda self: self._A.device)
    dtype = property(lambda self: self._A.dtype)


class _ChainPyTorchLinearOperator(PyTorchLinearOperator):
    """``A @ B @ C @ ...`` applied right-to-left."""

    def __init__(self, *operators: PyTorchLinearOperator):
        if len(operators) < 2:
            raise ValueError(f"Need at least 2 operators, got {len(operators)}.")
        for left, right in zip(operators[:-1], operators[1:]):
            if left._in_shape != right._out_shape:
                raise ValueError(
                    f"Shape mismatch: input shape {left._in_shape} does not match output shape "
                    f"{right._out_shape}."
                )
            _same_device_dtype(left, right)
        self._operators = list(operators)
        super().__init__(operators[-1]._in_shape, operators[0]._out_shape)

    device = property(lambda self: self._operators[0].device)
    dtype = property(lambda self: self._operators[0].dtype)

    def _matmat(self, X):
        for op in reversed(self._operators):
            X = op._matmat(X)
        return X

    def _adjoint(self):
        return _ChainPyTorchLinearOperator(*(op.adjoint() for op in reversed(self._operators)))

    def __iter__(self) -> Iterator[PyTorchLinearOperator]:
        return iter(self._operators)

    def __len__(self) -> int:
        return len(self._operators)

    def __getitem__(self, index: int) -> PyTorchLinearOperator:
        return self._operators[index]

    def __setitem__(self, index: int, value: PyTorchLinearOperator):
        old = self._operators[index]
        _same_spaces(old, value)
        _same_device_dtype(old, value)
        self._operators[index] = value
