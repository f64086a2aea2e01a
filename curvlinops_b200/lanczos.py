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


def fast_lanczos(A: PyTorchLinearOperator, ncv: int, use_eigh_tridiagonal: bool = False) -> tuple[Tensor, Tensor]:
    """Lanczos iterations without re-orthogonalisation; returns ``(evals, evecs)`` of the tridiagonal matrix
    (same contract and random start ``randn(dim)`` as the reference, ``spectrum.py:413-475``)."""
    device, dtype = A.device, A.dtype
    alphas = torch.zeros(ncv, device=device, dtype=dtype)
    betas = torch.zeros(ncv - 1, device=device, dtype=dtype)
    dim = A.shape[1]
    v, v_prev = None, None
    for m in range(ncv):
        if m == 0:
            v = torch.randn(dim, device=device, dtype=dtype)
            v /= torch.linalg.vector_norm(v)
            v_next = A @ v
        else:
            v_next = A @ v - betas[m - 1] * v_prev
        alphas[m] = (v_next * v).sum()
        v_next -= alphas[m] * v
        if m != ncv - 1:
            betas[m] = torch.linalg.vector_norm(v_next)
            v_next /= betas[m]
            v_prev = v
            v = v_next
    if use_eigh_tridiagonal:
        from scipy.linalg import eigh_tridiagonal

        ev, evec = eigh_tridiagonal(alphas.detach().cpu().numpy(), betas.detach().cpu().numpy())
        return (torch.as_tensor(ev, device=device, dtype=dtype), torch.as_tensor(evec, device=device, dtype=dtype))
    T = torch.diag_embed(alphas) + torch.diag_embed(betas, offset=1) + torch.diag_embed(betas, offset=-1)
    return torch.linalg.eigh(T)


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
    n = A.shape[1]
    if not 0 < k < n:
        raise ValueError(f"k must satisfy 0 < k < {n}, got {k}.")
    ncv = min(n, max(2 * k + 1, 20)) if ncv is None else min(n, ncv)
    maxiter = min(n, 10 * ncv) if maxiter is None else min(n, maxiter)
    if v0 is None:
        gen = torch.Generator(device="cpu").manual_seed(0)
        v0 = torch.randn(n, generator=gen, dtype=torch.float64).to(device=device, dtype=dtype)
    Q = torch.empty(maxiter + 1, n, device=device, dtype=dtype)  # Lanczos vectors (rows)
    Q[0] = v0 / torch.linalg.vector_norm(v0)
    alphas, betas = [], []
    evals = S = None
    m = 0
    for m in range(1, maxiter + 1):
        q = Q[m - 1]
        w = A @ q
        a = torch.dot(w, q)
        w = w - a * q - (betas[-1] * Q[m - 2] if betas else 0.0)
        # full re-orthogonalisation against all previous vectors (twice is enough)
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
            if breakdown or bool((resid <= tol * theta.abs().max()).all()) or m == maxiter:
                break
        betas.append(b)
        Q[m] = w / b
    order = torch.argsort(evals)
    vecs = (Q[:m].T.double() @ S[:, sel[order]].to(device)).to(dtype)
    out = (evals[order].to(device=device, dtype=dtype), vecs)
    return (*out, m) if return_info else out
