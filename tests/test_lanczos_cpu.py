"""Host logic of the on-device Lanczos drivers on a dense CPU operator (no GPU, no engine): the reference's
``fast_lanczos`` contract (tridiagonal eigen-decomposition) and ``lanczos_eigsh`` vs ``torch.linalg.eigvalsh``."""
import pytest
import torch

from curvlinops_b200.lanczos import fast_lanczos, lanczos_eigsh
from curvlinops_b200.linop import PyTorchLinearOperator


class Dense(PyTorchLinearOperator):
    SELF_ADJOINT = True

    def __init__(self, M):
        super().__init__([(M.shape[1],)], [(M.shape[0],)])
        self.M = M

    @property
    def device(self):
        return self.M.device

    @property
    def dtype(self):
        return self.M.dtype

    def _matmat(self, X):
        return [self.M @ X[0]]


def _spd(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    B = torch.randn(n, n, generator=g, dtype=torch.float64)
    d = torch.logspace(-3, 2, n, dtype=torch.float64)
    Q, _ = torch.linalg.qr(B)
    return (Q * d) @ Q.T


@pytest.mark.parametrize("which", ["LA", "SA", "LM"])
def test_lanczos_eigsh_matches_dense(which):
    M = _spd(120)
    A = Dense(M)
    ev, vec, nprod = lanczos_eigsh(A, k=5, which=which, tol=1e-10, return_info=True)
    full = torch.linalg.eigvalsh(M)
    want = {"LA": full[-5:], "SA": full[:5], "LM": full[-5:]}[which]
    torch.testing.assert_close(ev, want, rtol=1e-8, atol=1e-10)
    resid = (M @ vec - vec * ev).norm(dim=0)
    assert bool((resid <= 1e-6 * full.abs().max()).all()), resid
    assert nprod <= 120


def test_fast_lanczos_contract():
    torch.manual_seed(0)
    M = _spd(60, seed=1)
    ev, evec = fast_lanczos(Dense(M), ncv=60)
    assert ev.shape == (60,) and evec.shape == (60, 60)
    torch.testing.assert_close(ev[-1], torch.linalg.eigvalsh(M)[-1], rtol=1e-8, atol=1e-10)
    ev2, _ = fast_lanczos(Dense(M), ncv=30, use_eigh_tridiagonal=True)
    torch.testing.assert_close(ev2[-1], torch.linalg.eigvalsh(M)[-1], rtol=1e-6, atol=1e-8)


def test_lanczos_argument_errors():
    A = Dense(_spd(10))
    with pytest.raises(ValueError):
        lanczos_eigsh(A, k=10)
    with pytest.raises(ValueError):
        lanczos_eigsh(A, k=2, which="XX")


def test_lanczos_basis_grows_in_blocks_and_warns_without_convergence():
    """The Lanczos basis is allocated ncv + 1 rows at a time (not maxiter + 1 up front), low-precision operators run
    their vector algebra in fp32, and a run that stops at maxiter without meeting the residual test says so."""
    import warnings

    M = _spd(60, seed=3)
    A = Dense(M)
    ref = torch.linalg.eigvalsh(M)[-3:]
    evals, evecs, m = lanczos_eigsh(A, k=3, ncv=8, maxiter=60, tol=1e-10, return_info=True)
    assert m > 9  # went past the first block of the basis
    torch.testing.assert_close(evals, ref, rtol=1e-8, atol=1e-10)
    torch.testing.assert_close(M @ evecs, evecs * evals, rtol=1e-6, atol=1e-8)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        lanczos_eigsh(A, k=3, ncv=4, maxiter=4, tol=0.0)
    assert any("residual test not met" in str(x.message) for x in w)
    # bf16 operator: fp32 Lanczos vectors, bf16 results
    Ab = Dense(M.to(torch.bfloat16))
    eb, vb = lanczos_eigsh(Ab, k=2, ncv=20, tol=1e-3)
    assert eb.dtype == torch.bfloat16 and vb.dtype == torch.bfloat16
    assert ((eb.double() - ref[-2:]).abs() / ref[-1] < 3e-2).all()
