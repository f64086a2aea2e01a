"""On-device Lanczos for extremal eigenpairs of a symmetric operator.

``fast_lanczos`` mirrors the reference's routine of the same name (``curvlinops/papyan2020traces/spectrum.py:
413-475``: algorithm 2 of Papyan 2020, no re-orthogonalisation, returns the eigen-decomposition of the tridiagonal
matrix).  ``lanczos_eigsh`` is the driver the reference obtains from SciPy (``eigsh(A.to_scipy(), k)``,
``docs/examples/basic_usage/example_eigenvalues.py:65-77``) but keeps every vector on the device: no NumPy round
trip per matvec, full re-orthogonalisation of the (few) Lanczos vectors, Ritz pairs returned as tensors.

All arithmetic of the operator product runs in the engine; the vector algebra between products (dot products,
axpys on ``[P]`` vectors) is a handful of library calls per iteration, negligible next to a product.
"""

from __future__ import annotations

import torch
from torch import Tensor

from .linop import PyTorchLinearOperator


def _three_term_recurrence(A: PyTorchLinearOperator, steps: int, start: Tensor) -> tuple[Tensor, Tensor]:
    """Diagonal / off-diagonal of the Lanczos tridiagonal after ``steps`` products, keeping only the two most
    recent basis vectors (no re-orthogonalisation)."""
    diag = torch.empty(steps, device=A.device, dtype=A.dtype)
    off = torch.empty(max(steps - 1, 0), device=A.device, dtype=A.dtype)
    q_old, q = torch.zeros_like(start), start / torch.linalg.vector_norm(start)
    b_old = torch.zeros((), device=A.device, dtype=A.dtype)
    for j in range(steps):
        r = A @ q
        if j > 0:
            r = r - b_old * q_old
        diag[j] = torch.dot(r, q)
        r = r - diag[j] * q
        if j + 1 < steps:
            b_old = torch.linalg.vector_norm(r)
            off[j] = b_old
            q_old, q = q, r / b_old
    return diag, off


def fast_lanczos(A: PyTorchLinearOperator, ncv: int, use_eigh_tridiagonal: bool = False) -> tuple[Tensor, Tensor]:
    """Eigen-decomposition ``(evals, evecs)`` of the ``ncv x ncv`` Lanczos tridiagonal of ``A`` started from a
    standard-normal vector, without re-orthogonalisation - the contract of the reference's routine of the same name
    (``spectrum.py:413-475``; ``evecs[:, i]`` belongs to ``evals[i]``).  ``use_eigh_tridiagonal`` selects SciPy's
    tridiagonal solver (faster, less stable) instead of a dense ``eigh``."""
    start = torch.randn(A.shape[1], device=A.device, dtype=A.dtype)
    diag, off = _three_term_recurrence(A, ncv, start)
    if use_eigh_tridiagonal:
        from scipy.linalg import eigh_tridiagonal

        w, U = eigh_tridiagonal(diag.cpu().numpy(), off.cpu().numpy())
        return torch.as_tensor(w, device=A.device, dtype=A.dtype), torch.as_tensor(U, device=A.device, dtype=A.dtype)
    T = torch.diag(diag) + torch.diag(off, 1) + torch.diag(off, -1)
    return torch.linalg.eigh(T)


def _device_reorth(Q: Tensor, maxiter: int):
    """Two rounds of ``w -= Q[:m]^T (Q[:m] w)`` on the library's streaming kernels (``csrc/lanczos.cuh``): torch's
    matrix-vector products on [m, n] with m ~ 30 and n in the tens of millions ran at a few percent of the memory
    roofline (31 ms of a 57 ms Lanczos step on ResNet-50, against 26 ms for the GGN product itself)."""
    from . import _capi as capi

    L = capi.lib()
    n = Q.shape[1]
    ws = torch.empty(max(1, L.curv_lanczos_reorth_workspace(maxiter, n) // 4), dtype=torch.float32, device=Q.device)

    def run(Qc: Tensor, m: int, w: Tensor) -> None:
        assert w.is_contiguous() and Qc.stride(1) == 1
        with torch.cuda.device(Qc.device):
            capi.check(L.curv_lanczos_reorth(Qc.data_ptr(), Qc.stride(0), m, w.data_ptr(), n, 2, None, ws.data_ptr(),
                                             ws.numel() * 4, torch.cuda.current_stream(Qc.device).cuda_stream))

    return run


def lanczos_eigsh(A: PyTorchLinearOperator, k: int = 6, which: str = "LA", ncv: int | None = None,
                  tol: float = 1e-6, maxiter: int | None = None, v0: Tensor | None = None,
                  return_info: bool = False):
    """``k`` extremal eigenpairs of the symmetric operator ``A`` with all vectors resident on ``A.device``.

    Plain Lanczos with full re-orthogonalisation, grown until the residual estimates ``|beta_m s_mi|`` of the ``k``
    wanted Ritz pairs fall below ``tol * max|theta|`` (or ``maxiter`` products).  ``which``: ``"LA"`` largest
    algebraic, ``"SA"`` smallest algebraic, ``"LM"`` largest magnitude.  Returns ``(evals [k] ascending,
    evecs [P, k])`` like ``scipy.sparse.linalg.eigsh``; with ``return_info`` also the number of products.
    """
    if which not in ("LA", "SA", "LM"):
        raise ValueError(f"which must be 'LA', 'SA' or 'LM', got {which!r}.")
    device, dtype = A.device, A.dtype
    # the vector algebra runs in fp32 for low-precision operators (bf16 Lanczos vectors lose orthogonality at once);
    # the operator is handed its own dtype, as the SciPy bridge does (``_torch_base.py:560-592``)
    work = torch.float32 if dtype in (torch.bfloat16, torch.float16) else dtype
    n = A.shape[1]
    if not 0 < k < n:
        raise ValueError(f"k must satisfy 0 < k < {n}, got {k}.")
    ncv = min(n, max(2 * k + 1, 20)) if ncv is None else min(n, ncv)
    maxiter = min(n, 10 * ncv) if maxiter is None else min(n, maxiter)
    if v0 is None:  # seeded start vector, drawn on the operator's device (25 M normals take 0.4 s on the host)
        gen = torch.Generator(device=device).manual_seed(0)
        v0 = torch.randn(n, generator=gen, device=device, dtype=work if device.type == "cuda" else torch.float64)
    v0 = v0.to(device=device, dtype=work)
    # Lanczos vectors (rows), grown in blocks of ncv rows: maxiter x n up front would be tens of GB for a ResNet
    Q = torch.empty(min(maxiter, ncv) + 1, n, device=device, dtype=work)
    Q[0] = v0 / torch.linalg.vector_norm(v0)
    alphas, betas = [], []
    evals = S = None
    m = 0
    converged = False
    reorth = _device_reorth(Q, maxiter) if device.type == "cuda" and work == torch.float32 else None
    for m in range(1, maxiter + 1):
        q = Q[m - 1]
        w = (A @ q.to(dtype)).to(work)
        a = torch.dot(w, q)
        w = w - a * q - (betas[-1] * Q[m - 2] if betas else 0.0)
        # full re-orthogonalisation against all previous vectors (twice is enough)
        if reorth is not None:
            reorth(Q, m, w)
        else:
            for _ in range(2):
                w = w - Q[:m].T @ (Q[:m] @ w)
        b = torch.linalg.vector_norm(w)
        alphas.append(a)
        check = m >= ncv and (m == maxiter or (m - ncv) % max(1, k // 2) == 0)
        breakdown = float(b) <= 1e-12 * max(1.0, abs(float(a)))
        if check or breakdown or m == maxiter:
            al = torch.stack(alphas).double().cpu()
            be = torch.stack(betas).double().cpu() if betas else torch.zeros(0, dtype=torch.float64)
            T = torch.diag_embed(al) + torch.diag_embed(be, offset=1) + torch.diag_embed(be, offset=-1)
            theta, S = torch.linalg.eigh(T)
            order = {"LA": torch.argsort(theta, descending=True),
                     "SA": torch.argsort(theta),
                     "LM": torch.argsort(theta.abs(), descending=True)}[which][:k]
            resid = (float(b) * S[-1, order]).abs()
            evals, sel = theta[order], order
            converged = breakdown or bool((resid <= tol * theta.abs().max()).all())
            if converged or m == maxiter:
                break
        betas.append(b)
        if m >= Q.shape[0]:
            Q = torch.cat([Q, torch.empty(min(ncv, maxiter + 1 - Q.shape[0]), n, device=device, dtype=work)])
        Q[m] = w / b
    if not converged:
        import warnings

        warnings.warn(f"lanczos_eigsh: residual test not met after {m} products (tol={tol}); returning the current "
                      "Ritz pairs.", stacklevel=2)
    order = torch.argsort(evals)
    vecs = (Q[:m].T @ S[:, sel[order]].to(device=device, dtype=work)).to(dtype)
    out = (evals[order].to(device=device, dtype=dtype), vecs)
    return (*out, m) if return_info else out
