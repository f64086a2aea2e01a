"""Structured operators KFAC/EKFAC are assembled from: Kronecker products, eigen-decomposed
operators, block diagonals and the parameter-space <-> canonical-space converters.

Interfaces follow the reference (``curvlinops/kronecker.py:42-373``, ``curvlinops/eigh.py:12-177``,
``curvlinops/blockdiagonal.py:19-189``, ``curvlinops/kfac_utils.py:208-398``).  The products
``(S1 (x) S2) X`` and ``Q diag(lambda) Q^T X`` run in the CUDA library (``curv_kron_apply`` /
``curv_eigh_apply`` / ``curv_gemm``), the converters are pure data movement.  Factor *preparation*
(damped Cholesky inverse, ``eigh``) calls ``torch.linalg`` (cuSOLVER): a library call, stated in
DESIGN.md; it happens once per operator, not per product.
"""

from __future__ import annotations

from collections.abc import Iterator
from math import prod, sqrt
from warnings import warn

import torch
from torch import Size, Tensor

from . import _capi as capi
from .linop import PyTorchLinearOperator


def _same_meta(old: Tensor, new: Tensor) -> None:
    if old.shape != new.shape:
        raise ValueError(f"Shape mismatch: expected {old.shape}, got {new.shape}.")
    if old.device != new.device:
        raise ValueError(f"Device mismatch: expected {old.device}, got {new.device}.")
    if old.dtype != new.dtype:
        raise ValueError(f"Dtype mismatch: expected {old.dtype}, got {new.dtype}.")


def _one(values, what):
    vals = set(values)
    if len(vals) != 1:
        raise RuntimeError(f"Could not infer {what}: found {vals}.")
    return vals.pop()


def with_fp32_master(v32: Tensor, dtype: torch.dtype) -> Tensor:
    """Factor in the operator's dtype.  For bf16 operators the tensor handed out is bf16 (the reference's factors have
    the parameter dtype, ``kfac_hooks.py:350-353``) but it carries the fp32 matrix it was rounded from: the engine's
    contract for bf16 is fp32 factors / fp32 Cholesky and eigh (the reference cannot factorise bf16 matrices at all,
    ``kronecker.py:356-373``), so products and inverses of these operators read the master."""
    if dtype == v32.dtype:
        return v32
    t = v32.to(dtype)
    t._curv_fp32 = v32
    return t


def _fp32_master(t: Tensor) -> Tensor | None:
    m = getattr(t, "_curv_fp32", None)
    return m if isinstance(m, Tensor) and m.shape == t.shape and m.device == t.device else None


def _cuda_f32(t: Tensor, what: str) -> Tensor:
    if t.device.type != "cuda":
        raise RuntimeError(f"curvlinops_b200 applies {what} on CUDA devices only (no CPU fallback); got {t.device}.")
    m = _fp32_master(t)
    return (t if m is None else m).to(torch.float32).contiguous()


#: blocks with both sides at least this wide run on the tensor-core apply (``curv_kron_apply_tc``), smaller ones on the
#: fp32 SIMT kernels
TENSOR_CORE_MIN_DIM = 32


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def ensure_all_square(*mats) -> None:
    for m in mats:
        if len(m.shape) != 2 or m.shape[0] != m.shape[1]:
            raise ValueError(f"Expected square matrices, got shape {tuple(m.shape)}.")


def dense_matmul(A: Tensor, X: Tensor, transpose_a: bool = False) -> Tensor:
    """``op(A) @ X`` for 2-d fp32 CUDA tensors through ``curv_gemm`` (hand-written kernel, no cuBLAS)."""
    A32, X32 = _cuda_f32(A, "dense products"), _cuda_f32(X, "dense products")
    M = A32.shape[1] if transpose_a else A32.shape[0]
    Kd = A32.shape[0] if transpose_a else A32.shape[1]
    N = X32.shape[1]
    Y = torch.empty(M, N, device=X32.device, dtype=torch.float32)
    capi.check(capi.lib().curv_gemm(int(transpose_a), 0, M, N, Kd, 1.0, A32.data_ptr(), A32.shape[1],
                                    X32.data_ptr(), N, 0.0, Y.data_ptr(), N, _stream(X32)))
    return Y.to(X.dtype)


def batched_matmul_tn(A: Tensor, B: Tensor) -> Tensor:
    """``C[n] = A[n]^T @ B[n]`` for fp32 CUDA tensors ``A [N, S, P]``, ``B [N, S, Q]`` -> ``[N, P, Q]`` in ONE launch of the
    hand-written batched kernel (``curv_gemm_batched``; no cuBLAS)."""
    A32, B32 = _cuda_f32(A, "batched products"), _cuda_f32(B, "batched products")
    N, S, P = A32.shape
    Q = B32.shape[2]
    out = torch.empty(N, P, Q, device=A32.device, dtype=torch.float32)
    for n0 in range(0, N, 65535):
        nb = min(65535, N - n0)
        capi.check(capi.lib().curv_gemm_batched(1, 0, P, Q, S, 1.0, A32[n0:].data_ptr(), P, S * P, B32[n0:].data_ptr(), Q,
                                                S * Q, 0.0, out[n0:].data_ptr(), Q, P * Q, nb, _stream(A32)))
    return out


class KroneckerProductLinearOperator(PyTorchLinearOperator):
    r"""``S_1 \otimes S_2`` (one or two factors) acting on ``vec`` of a row-major ``[d_1, d_2]`` matrix."""

    def __init__(self, *factors: Tensor):
        if len(factors) == 0:
            raise ValueError("At least one factor must be provided.")
        for i, f in enumerate(factors):
            if f.ndim != 2:
                raise ValueError(f"Factor {i} must be a 2D tensor, got shape {f.shape}.")
        if len(factors) > 2:
            raise NotImplementedError("The B200 engine applies Kronecker products of one or two factors.")
        self._factors = list(factors)
        super().__init__([(prod(S.shape[1] for S in factors),)], [(prod(S.shape[0] for S in factors),)])

    def __iter__(self) -> Iterator[Tensor]:
        return iter(self._factors)

    def __len__(self) -> int:
        return len(self._factors)

    def __getitem__(self, index: int) -> Tensor:
        return self._factors[index]

    def __setitem__(self, index: int, value: Tensor):
        _same_meta(self._factors[index], value)
        self._factors[index] = value

    device = property(lambda self: _one((f.device for f in self._factors), "device"))
    dtype = property(lambda self: _one((f.dtype for f in self._factors), "dtype"))

    def _apply(self, x: Tensor, transpose: bool) -> Tensor:
        """``x`` is ``[D_in, K]``; returns ``[D_out, K]`` (einsum 'abZ,Aa,Bb->ABZ' of the reference)."""
        fs = [self._f32(i, transpose) for i in range(len(self._factors))]
        K = x.shape[-1]
        x32 = _cuda_f32(x, "Kronecker products")
        if len(fs) == 1:
            return dense_matmul(fs[0], x32).to(x.dtype)
        G, A = fs
        if G.shape[0] != G.shape[1] or A.shape[0] != A.shape[1]:
            # rectangular factors: two plain products  Y = G (X A^T)
            d_out_in, d_in_in = G.shape[1], A.shape[1]
            X3 = x32.reshape(d_out_in, d_in_in, K)
            T = dense_matmul(A, X3.permute(1, 0, 2).reshape(d_in_in, -1)).reshape(A.shape[0], d_out_in, K)
            Y = dense_matmul(G, T.permute(1, 0, 2).reshape(d_out_in, -1))
            return Y.reshape(-1, K).to(x.dtype)
        d_out, d_in = G.shape[0], A.shape[0]
        Y = torch.empty(d_out * d_in, K, device=x32.device, dtype=torch.float32)
        if min(d_out, d_in) >= TENSOR_CORE_MIN_DIM:
            # blocks of real networks: both contractions on the tcgen05 kernel, operands as fp16 hi/lo planes
            # (fp32-grade) for bf16 operators too: rows of a gradient covariance sum to ~0 against smooth vectors, and
            # bf16-rounded operands (measured on ResNet-18: 4e-2 of the result) do not survive that cancellation.
            # The operand forms of the two factors are built once and cached (they only change with the factors).
            fws, ready = self._factor_operands(transpose, G, A)
            for k0 in range(0, K, 8):
                kk = min(8, K - k0)
                xk = x32 if kk == K else x32[:, k0:k0 + kk].contiguous()
                yk = Y if kk == K else torch.empty(d_out * d_in, kk, device=x32.device, dtype=torch.float32)
                nbytes = capi.lib().curv_kron_apply_tc_workspace(d_out, d_in, kk)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x32.device)
                capi.check(capi.lib().curv_kron_apply_tc(G.data_ptr(), A.data_ptr(), d_out, d_in, kk, xk.data_ptr(),
                                                         yk.data_ptr(), fws.data_ptr(), fws.numel(), int(ready),
                                                         ws.data_ptr(), nbytes, _stream(x32)))
                ready = True
                if kk != K:
                    Y[:, k0:k0 + kk] = yk
            return Y.to(x.dtype)
        tmp = torch.empty_like(Y)
        capi.check(capi.lib().curv_kron_apply(G.data_ptr(), A.data_ptr(), d_out, d_in, K, x32.data_ptr(),
                                              Y.data_ptr(), tmp.data_ptr(), _stream(x32)))
        return Y.to(x.dtype)

    def _f32(self, index: int, transpose: bool) -> Tensor:
        """Factor ``index`` (its fp32 master for bf16 operators), transposed for the adjoint product, as a contiguous
        fp32 CUDA matrix; cached per factor object / version so that repeated products see the same buffer."""
        f = self._factors[index]
        key = (id(f), f._version)
        cache = self.__dict__.setdefault("_f32_cache", {})
        hit = cache.get((index, transpose))
        if hit is None or hit[0] != key:
            m = _fp32_master(f)
            base = f if m is None else m
            hit = (key, _cuda_f32(base.mH if transpose else base, "Kronecker products"))
            cache[(index, transpose)] = hit
        return hit[1]

    def _factor_operands(self, adjoint: bool, G: Tensor, A: Tensor):
        """Device buffer with the tensor-core operand forms of the two factors (``curv_kron_apply_tc`` fills it on first
        use) and whether it is current: keyed on the factor objects and their version counters."""
        key = tuple((id(f), f._version) for f in self._factors) + (G.data_ptr(), A.data_ptr())
        cache = self.__dict__.setdefault("_tc_cache", {})
        hit = cache.get(adjoint)
        if hit is not None and hit[0] == key:
            return hit[1], True
        nbytes = capi.lib().curv_kron_apply_tc_factor_bytes(G.shape[0], A.shape[0])
        buf = torch.empty(nbytes, dtype=torch.uint8, device=G.device)
        cache[adjoint] = (key, buf, G, A)  # G / A kept alive: their addresses are part of the key
        return buf, False

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        (x,) = X
        return [self._apply(x, transpose=False)]

    def _adjoint_matmat(self, X: list[Tensor]) -> list[Tensor]:
        (x,) = X
        return [self._apply(x, transpose=True)]

    def _adjoint(self) -> "KroneckerProductLinearOperator":
        return KroneckerProductLinearOperator(*[f.mH for f in self._factors])

    # ---- properties (host-side math on the small factors) ---------------------------------------
    def trace(self) -> Tensor:
        ensure_all_square(*self._factors)
        return torch.stack([S.trace() for S in self._factors]).prod()

    def det(self) -> Tensor:
        ensure_all_square(*self._factors)
        dim = prod(S.shape[0] for S in self._factors)
        return torch.stack([S.det() ** (dim // S.shape[0]) for S in self._factors]).prod()

    def logdet(self) -> Tensor:
        ensure_all_square(*self._factors)
        dim = prod(S.shape[0] for S in self._factors)
        return torch.stack([(dim // S.shape[0]) * S.logdet() for S in self._factors]).sum()

    def frobenius_norm(self) -> Tensor:
        return torch.stack([torch.linalg.matrix_norm(S) for S in self._factors]).prod()

    # ---- inverse (reference kronecker.py:250-373) ------------------------------------------------
    def inverse(self, damping: float = 0.0, use_heuristic_damping: bool = False, min_damping: float = 1e-8,
                use_exact_damping: bool = False, retry_double_precision: bool = True):
        ensure_all_square(*self._factors)
        if use_heuristic_damping and use_exact_damping:
            raise ValueError("Either use heuristic damping or exact damping, not both.")
        if use_heuristic_damping and len(self._factors) > 2:
            raise ValueError(
                f"Heuristic damping only implemented for at most two factors. Got {len(self._factors)}"
            )
        if use_exact_damping:  # (S1 (x) S2 + damping I)^-1 through the factors' eigendecompositions
            evals, evecs = zip(*[torch.linalg.eigh(S) for S in self._factors])
            lam = evals[0]
            for e in evals[1:]:
                lam = torch.kron(lam, e)
            return EighDecomposedLinearOperator(lam, KroneckerProductLinearOperator(*evecs)).inverse(
                damping=damping)
        if use_heuristic_damping and len(self._factors) == 1:
            dampings = (max(damping, min_damping),)
        elif use_heuristic_damping:  # Martens & Grosse 2015, section 6.3
            S1, S2 = self._factors
            m1, m2 = S1.diag().mean(), S2.diag().mean()
            if m1 < 0 or m2 < 0:
                raise RuntimeError("Negative mean eigenvalue detected")
            pi = (m2 / m1).sqrt()
            dampings = (max(sqrt(damping) / pi, min_damping), max(sqrt(damping) * pi, min_damping))
        else:
            dampings = (damping,) * len(self._factors)
        return KroneckerProductLinearOperator(*[
            self._damped_cholesky_inverse(S, d, retry_double_precision)
            for S, d in zip(self._factors, dampings)
        ])

    @staticmethod
    def _damped_cholesky_inverse(A: Tensor, damping, retry_double_precision: bool) -> Tensor:
        if A.dtype in (torch.bfloat16, torch.float16):  # factorise in fp32 (from the fp32 master when there is one)
            m = _fp32_master(A)
            inv32 = KroneckerProductLinearOperator._damped_cholesky_inverse(
                A.to(torch.float32) if m is None else m, damping, retry_double_precision)
            return with_fp32_master(inv32, A.dtype)

        def chol(M: Tensor) -> Tensor:
            return torch.linalg.cholesky(torch.diagonal_scatter(M, M.diag() + damping))

        try:
            L = chol(A)
        except RuntimeError as error:
            if not retry_double_precision or A.dtype == torch.float64:
                raise error
            warn(f"Failed to compute Cholesky decomposition in {A.dtype} precision with error {error}. "
                 "Retrying in double precision...", stacklevel=2)
            L = chol(A.to(torch.float64)).to(A.dtype)
        return torch.cholesky_inverse(L).contiguous()  # (LAPACK-ordered result: row-major once, not on every product)


class EighDecomposedLinearOperator(PyTorchLinearOperator):
    r"""``Q diag(lambda) Q^T`` with dense ``Q`` or a Kronecker product of eigenvector matrices."""

    SELF_ADJOINT = True

    def __init__(self, eigenvalues: Tensor, eigenvectors):
        if eigenvalues.ndim != 1:
            raise ValueError(f"Eigenvalues must be 1D, got shape {eigenvalues.shape}.")
        if len(eigenvectors.shape) != 2:
            raise ValueError(f"Eigenvectors must be 2D, got shape {eigenvectors.shape}.")
        if eigenvectors.shape[0] != eigenvectors.shape[1]:
            raise ValueError(f"Eigenvectors must be square, got shape {eigenvectors.shape}.")
        if eigenvalues.shape[0] != eigenvectors.shape[0]:
            raise ValueError(
                f"Incompatible shapes: eigenvalues {eigenvalues.shape}, eigenvectors {eigenvectors.shape}."
            )
        self._eigenvalues, self._eigenvectors = eigenvalues, eigenvectors
        n = eigenvalues.shape[0]
        super().__init__([(n,)], [(n,)])

    @property
    def eigenvalues(self) -> Tensor:
        return self._eigenvalues

    @eigenvalues.setter
    def eigenvalues(self, value: Tensor):
        _same_meta(self._eigenvalues, value)
        self._eigenvalues = value

    @property
    def eigenvectors(self):
        return self._eigenvectors

    device = property(lambda self: _one([self._eigenvalues.device, self._eigenvectors.device], "device"))
    dtype = property(lambda self: _one([self._eigenvalues.dtype, self._eigenvectors.dtype], "dtype"))

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        (x,) = X
        Q, lam = self._eigenvectors, self._eigenvalues
        K = x.shape[-1]
        x32 = _cuda_f32(x, "eigen-decomposed operators")
        lam32 = _cuda_f32(lam, "eigen-decomposed operators")
        if isinstance(Q, Tensor):
            QTx = dense_matmul(Q, x32, transpose_a=True)
            return [dense_matmul(Q, lam32.unsqueeze(1) * QTx).to(x.dtype)]
        fs = [_cuda_f32(f, "eigen-decomposed operators") for f in Q]
        if len(fs) == 1:
            QTx = dense_matmul(fs[0], x32, transpose_a=True)
            return [dense_matmul(fs[0], lam32.unsqueeze(1) * QTx).to(x.dtype)]
        Qg, Qa = fs
        d_out, d_in = Qg.shape[0], Qa.shape[0]
        if (isinstance(Q, KroneckerProductLinearOperator) and min(d_out, d_in) >= TENSOR_CORE_MIN_DIM
                and Qg.shape[0] == Qg.shape[1] and Qa.shape[0] == Qa.shape[1]):
            # the two rotations of eigh.py:98-104 on the tensor-core Kronecker apply: (Qg (x) Qa)^T x, scale, back
            rot = Q._apply(x32, transpose=True)
            return [Q._apply(lam32.unsqueeze(1) * rot, transpose=False).to(x.dtype)]
        Y = torch.empty(d_out * d_in, K, device=x32.device, dtype=torch.float32)
        t1, t2 = torch.empty_like(Y), torch.empty_like(Y)
        capi.check(capi.lib().curv_eigh_apply(Qg.data_ptr(), Qa.data_ptr(), lam32.data_ptr(), 0.0, 0, d_out,
                                              d_in, K, x32.data_ptr(), Y.data_ptr(), t1.data_ptr(),
                                              t2.data_ptr(), _stream(x32)))
        return [Y.to(x.dtype)]

    def trace(self) -> Tensor:
        return self._eigenvalues.sum()

    def det(self) -> Tensor:
        return self._eigenvalues.prod()

    def logdet(self) -> Tensor:
        return self._eigenvalues.log().sum()

    def frobenius_norm(self) -> Tensor:
        return self._eigenvalues.norm(p="fro")

    def inverse(self, damping: float = 0.0) -> "EighDecomposedLinearOperator":
        return EighDecomposedLinearOperator(1.0 / (self._eigenvalues + damping), self._eigenvectors)


class BlockDiagonalLinearOperator(PyTorchLinearOperator):
    """Block diagonal of linear operators; each block acts on its own slice of the tensor list."""

    def __init__(self, blocks: list[PyTorchLinearOperator]):
        if not blocks:
            raise ValueError("At least one block must be provided.")
        self._blocks = blocks
        super().__init__([tuple(s) for B in blocks for s in B._in_shape],
                         [tuple(s) for B in blocks for s in B._out_shape])
        self.SELF_ADJOINT = all(B.SELF_ADJOINT for B in blocks)

    def __iter__(self) -> Iterator[PyTorchLinearOperator]:
        return iter(self._blocks)

    def __len__(self) -> int:
        return len(self._blocks)

    def __getitem__(self, index: int) -> PyTorchLinearOperator:
        return self._blocks[index]

    def __setitem__(self, index: int, value: PyTorchLinearOperator):
        old = self._blocks[index]
        if old._in_shape != value._in_shape or old._out_shape != value._out_shape:
            raise ValueError(
                f"Shape mismatch: expected in_shape={old._in_shape}, out_shape={old._out_shape}, "
                f"got in_shape={value._in_shape}, out_shape={value._out_shape}."
            )
        if old.device != value.device:
            raise ValueError(f"Device mismatch: expected {old.device}, got {value.device}.")
        if old.dtype != value.dtype:
            raise ValueError(f"Dtype mismatch: expected {old.dtype}, got {value.dtype}.")
        self._blocks[index] = value

    def _matmat(self, X: list[Tensor]) -> list[Tensor]:
        out, pos = [], 0
        for B in self._blocks:
            n = len(B._in_shape)
            out.extend(B._matmat(X[pos:pos + n]))
            pos += n
        if pos != len(X):
            raise ValueError(f"List to be split has length {len(X)}, but blocks consume {pos} entries.")
        return out

    def _adjoint(self) -> "BlockDiagonalLinearOperator":
        return BlockDiagonalLinearOperator([B.adjoint() for B in self._blocks])

    device = property(lambda self: _one((B.device for B in self._blocks), "device"))
    dtype = property(lambda self: _one((B.dtype for B in self._blocks), "dtype"))

    def trace(self) -> Tensor:
        return torch.stack([B.trace() for B in self._blocks]).sum()

    def det(self) -> Tensor:
        return torch.stack([B.det() for B in self._blocks]).prod()

    def logdet(self) -> Tensor:
        return torch.stack([B.logdet() for B in self._blocks]).sum()

    def frobenius_norm(self) -> Tensor:
        return torch.stack([B.frobenius_norm() ** 2 for B in self._blocks]).sum().sqrt()


class _Canonicalization(PyTorchLinearOperator):
    """Shared logic of the converters between parameter space (one tensor per parameter, ``params``
    order) and KFAC's canonical space (one vector per group: ``vec_rowmajor([W.flatten(1) | b])``)."""

    def __init__(self, param_shapes: dict[str, Size], param_groups: list[dict[str, str]], device, dtype):
        self._param_shapes = {n: Size(s) for n, s in param_shapes.items()}
        self._param_groups = param_groups
        self._device, self._dtype = device, dtype
        self._pos = {n: i for i, n in enumerate(param_shapes)}
        param_space = [tuple(s) for s in self._param_shapes.values()]
        canonical = []
        for group in param_groups:
            if "W" in group and "b" in group:
                ws = self._param_shapes[group["W"]]
                canonical.append((ws.numel() + ws[0],))
            else:
                canonical.extend((self._param_shapes[n].numel(),) for n in group.values())
        self._spaces = (param_space, canonical)
        in_shape, out_shape = self._orient(param_space, canonical)
        super().__init__(in_shape, out_shape)

    device = property(lambda self: self._device)
    dtype = property(lambda self: self._dtype)


class ToCanonicalLinearOperator(_Canonicalization):
    """Parameter space -> canonical space (the ``P^T`` of ``KFAC = P K P^T``)."""

    @staticmethod
    def _orient(param_space, canonical):
        return param_space, canonical

    def _matmat(self, M: list[Tensor]) -> list[Tensor]:
        out = []
        for group in self._param_groups:
            if "W" in group and "b" in group:
                W, b = M[self._pos[group["W"]]], M[self._pos[group["b"]]]
                K = W.shape[-1]
                joined = torch.cat([W.reshape(W.shape[0], -1, K), b.unsqueeze(1)], dim=1)
                out.append(joined.reshape(-1, K))
            else:
                out.extend(M[self._pos[n]].reshape(-1, M[self._pos[n]].shape[-1]) for n in group.values())
        return out

    def _adjoint(self) -> "FromCanonicalLinearOperator":
        return FromCanonicalLinearOperator(self._param_shapes, self._param_groups, self._device, self._dtype)


class FromCanonicalLinearOperator(_Canonicalization):
    """Canonical space -> parameter space (the ``P`` of ``KFAC = P K P^T``)."""

    @staticmethod
    def _orient(param_space, canonical):
        return canonical, param_space

    def _matmat(self, M: list[Tensor]) -> list[Tensor]:
        out: list = [None] * len(self._param_shapes)
        (K,) = {m.shape[-1] for m in M}
        it = iter(M)
        used = 0
        for group in self._param_groups:
            if "W" in group and "b" in group:
                ws = self._param_shapes[group["W"]]
                rows, cols = ws[0], ws.numel() // ws[0]
                joined = next(it).reshape(rows, cols + 1, K)
                out[self._pos[group["W"]]] = joined[:, :cols].reshape(*ws, K)
                out[self._pos[group["b"]]] = joined[:, cols].reshape(rows, K)
                used += 1
            else:
                for n in group.values():
                    out[self._pos[n]] = next(it).reshape(*self._param_shapes[n], K)
                    used += 1
        if any(o is None for o in out) or used != len(M):
            raise RuntimeError("Mismatch in number of processed parameters.")
        return out

    def _adjoint(self) -> ToCanonicalLinearOperator:
        return ToCanonicalLinearOperator(self._param_shapes, self._param_groups, self._device, self._dtype)
