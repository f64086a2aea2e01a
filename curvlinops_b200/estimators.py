"""Randomised trace / diagonal / Frobenius-norm estimators driven by batched engine products.

SURVEY §8(f) row 4.  Same functions, arguments, errors and — for a given torch RNG state — the same probe
vectors as the reference (``curvlinops/sampling.py:6-56``, ``trace/hutchinson.py:10-75``,
``trace/meyer2020hutch.py:15-102``, ``trace/epperly2024xtrace.py:16-101``, ``diagonal/hutchinson.py:10-76``,
``diagonal/epperly2024xtrace.py:16-89``, ``norm/hutchinson.py:9-73``), pinned by
``tests/golden/estimators.npz`` (made from the reference by ``oracle/make_golden_estimators.py``).

All probe vectors of one estimate form ONE ``[P, N]`` matrix, i.e. one engine call of ``N`` columns (the
engine processes them in groups of its slot count); everything after the products is small dense algebra on
``[P, N]`` / ``[N, N]`` tensors on the operator's device.  The leave-one-out loops of XTrace / XDiag are
evaluated for all probes at once instead of one probe per Python iteration.
"""

from __future__ import annotations

import torch
from torch import Tensor


# ---- probe vectors (reference sampling.py) -------------------------------------------------------
def rademacher(dim: int, device, dtype) -> Tensor:
    """i.i.d. ±1 entries, drawn exactly like the reference (``bernoulli_(0.5)*2-1``)."""
    return torch.empty(dim, device=device, dtype=dtype).bernoulli_(0.5).mul_(2).sub_(1)


def normal(dim: int, device, dtype) -> Tensor:
    return torch.randn(dim, device=device, dtype=dtype)


def random_vector(dim: int, distribution: str, device, dtype) -> Tensor:
    if distribution == "rademacher":
        return rademacher(dim, device, dtype)
    if distribution == "normal":
        return normal(dim, device, dtype)
    raise ValueError(f"Unknown distribution {distribution!r}.")


def _probes(dim: int, num: int, distribution: str, A) -> Tensor:
    """``[dim, num]`` probe matrix; vectors are drawn one after the other so that the RNG stream, and with it
    the estimate, equals the reference's for the same seed."""
    return torch.column_stack([random_vector(dim, distribution, A.device, A.dtype) for _ in range(num)])


# ---- argument checks (reference utils.py:218-264) -------------------------------------------------
def _square_dim(A) -> int:
    if len(A.shape) != 2 or A.shape[0] != A.shape[1]:
        raise ValueError(f"Operator must be square. Got shape {A.shape}.")
    return A.shape[0]


def _fewer_matvecs_than_dim(A, num_matvecs: int) -> None:
    if any(num_matvecs >= d for d in A.shape):
        raise ValueError(f"num_matvecs ({num_matvecs}) must be less than A's size ({A.shape}).")


def _divisible(num: int, divisor: int, name: str) -> None:
    if num % divisor != 0:
        raise ValueError(f"{name} ({num}) must be divisible by {divisor}.")


# ---- trace ---------------------------------------------------------------------------------------
def hutchinson_trace(A, num_matvecs: int, distribution: str = "rademacher") -> Tensor:
    """Girard–Hutchinson: ``mean_n v_nᵀ A v_n``."""
    dim = _square_dim(A)
    _fewer_matvecs_than_dim(A, num_matvecs)
    G = _probes(dim, num_matvecs, distribution, A)
    return (G * (A @ G)).sum() / num_matvecs


def hutchpp_trace(A, num_matvecs: int, distribution: str = "rademacher") -> Tensor:
    """Hutch++ (Meyer et al. 2020): exact trace on ``range(A S)`` + Hutchinson on its complement; a third
    of the products for each of sketch, subspace trace and complement."""
    dim = _square_dim(A)
    _fewer_matvecs_than_dim(A, num_matvecs)
    _divisible(num_matvecs, 3, "num_matvecs")
    N = num_matvecs // 3
    Q, _ = torch.linalg.qr(A @ _probes(dim, N, distribution, A))
    in_subspace = (Q * (A @ Q)).sum()
    G = _probes(dim, N, distribution, A)
    AG = A @ (G - Q @ (Q.T @ G))
    AG = AG - Q @ (Q.T @ AG)
    return in_subspace + (G * AG).sum() / N


def _leave_one_out_directions(R: Tensor) -> Tensor:
    """Columns ``s_i`` with ``Q_i Q_iᵀ = Q (I − s_i s_iᵀ) Qᵀ``, where ``Q_i`` is the basis had probe ``i`` been
    left out of ``qr(A W) = Q R`` (Epperly et al. 2024, §2.1): normalised columns of ``R⁻ᵀ``."""
    Rinv_T = torch.linalg.inv(R.T)
    return Rinv_T / torch.linalg.vector_norm(Rinv_T, dim=0, keepdim=True)


def xtrace(A, num_matvecs: int, distribution: str = "rademacher") -> Tensor:
    """XTrace (Epperly, Tropp, Webber 2024): Hutch++ made exchangeable — every probe serves once as the
    Hutchinson probe on the complement of the basis built from all the others; the estimates are averaged."""
    dim = _square_dim(A)
    _fewer_matvecs_than_dim(A, num_matvecs)
    _divisible(num_matvecs, 2, "num_matvecs")
    n = num_matvecs // 2
    W = _probes(dim, n, distribution, A)
    AW = A @ W
    Q, R = torch.linalg.qr(AW)
    AQ = A @ Q
    S = _leave_one_out_directions(R)

    H = Q.T @ AQ                      # [n, n]
    # tr(Q_iᵀ A Q_i) = tr(H) − s_iᵀ H s_i
    traces = H.trace() - ((H @ S) * S).sum(0)

    # Hutchinson term of probe i on the complement of Q_i:  w_iᵀ (I − Q_i Q_iᵀ) A (I − Q_i Q_iᵀ) w_i,
    # all probes at once; deflate(V)[:, i] = (I − s_i s_iᵀ) V[:, i]
    def deflate(V: Tensor) -> Tensor:
        return V - S * (S * V).sum(0, keepdim=True)

    A_P_W = AW - AQ @ deflate(Q.T @ W)
    PT_A_P_W = A_P_W - Q @ deflate(Q.T @ A_P_W)
    traces = traces + (W * PT_A_P_W).sum(0)
    return traces.mean()


# ---- diagonal ------------------------------------------------------------------------------------
def hutchinson_diag(A, num_matvecs: int, distribution: str = "rademacher") -> Tensor:
    """Bekas et al. 2007: ``mean_n v_n ⊙ A v_n``."""
    dim = _square_dim(A)
    _fewer_matvecs_than_dim(A, num_matvecs)
    G = _probes(dim, num_matvecs, distribution, A)
    return (G * (A @ G)).sum(1) / num_matvecs


def xdiag(A, num_matvecs: int) -> Tensor:
    """XDiag (Epperly, Tropp, Webber 2024): the exchangeable Diag++; Rademacher probes.  Uses ``Qᵀ A``,
    i.e. products with the adjoint."""
    dim = _square_dim(A)
    _fewer_matvecs_than_dim(A, num_matvecs)
    _divisible(num_matvecs, 2, "num_matvecs")
    n = num_matvecs // 2
    W = _probes(dim, n, "rademacher", A)
    AW = A @ W
    Q, R = torch.linalg.qr(AW)
    QT_A = Q.T @ A                    # [n, P]
    S = _leave_one_out_directions(R)

    # mean_i diag(Q_i Q_iᵀ A) = diag(Q Qᵀ A) − diag(Q S Sᵀ Qᵀ A)/n
    diagonal = (Q * QT_A.T).sum(1) - ((Q @ S) * (QT_A.T @ S)).sum(1) / n

    # + mean_i  w_i ⊙ (I − Q_i Q_iᵀ) A w_i / w_i²
    C = QT_A @ W                      # [n, n], column i = Qᵀ A w_i
    C = C - S * (S * C).sum(0, keepdim=True)
    comp = AW - Q @ C
    return diagonal + (W * comp / W**2).sum(1) / n


# ---- norm ----------------------------------------------------------------------------------------
def hutchinson_squared_fro(A, num_matvecs: int, distribution: str = "rademacher") -> Tensor:
    """``‖A‖_F² = tr(AᵀA) ≈ mean_n ‖A v_n‖²`` (with ``Aᵀ`` when the matrix is wider than tall)."""
    if len(A.shape) != 2:
        raise ValueError(f"A must be a matrix. Got shape {A.shape}.")
    dim = min(A.shape)
    if num_matvecs >= dim:
        raise ValueError(f"num_matvecs ({num_matvecs}) must be less than the minimum dimension of A.")
    if A.shape[1] > A.shape[0]:
        A = A.T if isinstance(A, Tensor) else A.adjoint()
    G = _probes(dim, num_matvecs, distribution, A)
    AG = A @ G
    return (AG**2 / num_matvecs).sum()
