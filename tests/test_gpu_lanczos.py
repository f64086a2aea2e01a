"""The library's Lanczos re-orthogonalisation kernels (``csrc/lanczos.cuh``) against a float64 torch restatement of the
same two Gram-Schmidt rounds, on ragged shapes (n not a multiple of 4, m not a multiple of the row group, one row, a
strided Q), bit-wise repeatability, and `lanczos_eigsh` on a CUDA operator with a known spectrum."""
import pytest
import torch

from curvlinops_b200 import _capi as capi
from curvlinops_b200.lanczos import lanczos_eigsh
from curvlinops_b200.linop import PyTorchLinearOperator

pytestmark = pytest.mark.gpu


def _reorth(Q, m, w, rounds=2, coeff=None):
    L = capi.lib()
    n = Q.shape[1]
    ws = torch.empty(L.curv_lanczos_reorth_workspace(m, n) // 4 + 1, dtype=torch.float32, device=Q.device)
    capi.check(L.curv_lanczos_reorth(Q.data_ptr(), Q.stride(0), m, w.data_ptr(), n, rounds,
                                     None if coeff is None else coeff.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                     torch.cuda.current_stream().cuda_stream))


@pytest.mark.parametrize("m,n,ld", [(1, 1000, None), (3, 4099, None), (7, 70001, None), (30, 1 << 20, None),
                                    (13, 300000, 300004), (5, 300001, 300003), (64, 50000, None)])
def test_reorth_matches_float64(m, n, ld):
    torch.manual_seed(m * 7 + n)
    dev = torch.device("cuda")
    ld = n if ld is None else ld
    Qs = torch.randn(m + 2, ld, device=dev)
    Q = Qs[:, :n]
    Q[:m] = torch.linalg.qr(Q[:m].double().T).Q.T.float()  # orthonormal rows, as in the solver
    w = torch.randn(n, device=dev)
    ref = w.double()
    tot = torch.zeros(m, dtype=torch.float64, device=dev)
    for _ in range(2):
        c = Q[:m].double() @ ref
        tot += c
        ref = ref - Q[:m].double().T @ c
    got, coeff = w.clone(), torch.full((m,), 7.0, device=dev)
    _reorth(Q, m, got, coeff=coeff)
    assert (got.double() - ref).abs().max() <= 2e-6 * w.abs().max()
    assert (coeff.double() - tot).abs().max() <= 1e-5 * max(1.0, tot.abs().max().item())
    assert (Q[:m] @ got).abs().max() <= 1e-5 * got.norm()  # orthogonal to every previous vector
    assert not torch.isnan(got).any()
    again = w.clone()
    _reorth(Q, m, again)
    assert torch.equal(again, got)  # fixed summation order


def test_reorth_rejects_bad_arguments():
    Q = torch.randn(4, 64, device="cuda")
    w = torch.randn(64, device="cuda")
    L = capi.lib()
    assert L.curv_lanczos_reorth(Q.data_ptr(), 64, 4, w.data_ptr(), 64, 2, None, None, 0, None) != 0
    ws = torch.empty(1, device="cuda")
    assert L.curv_lanczos_reorth(Q.data_ptr(), 64, 4, w.data_ptr(), 64, 2, None, ws.data_ptr(), 4, None) != 0
    assert b"workspace" in L.curv_last_error()


class _Diag(PyTorchLinearOperator):
    SELF_ADJOINT = True

    def __init__(self, d):
        super().__init__([tuple(d.shape)], [tuple(d.shape)])
        self._d = d

    @property
    def device(self):
        return self._d.device

    @property
    def dtype(self):
        return self._d.dtype

    def _matmat(self, X):
        return [self._d.unsqueeze(-1) * X[0]]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lanczos_eigsh_on_cuda_operator(dtype):
    n = 20011
    d = torch.linspace(0.0, 1.0, n, device="cuda")
    d[-4:] = torch.tensor([3.0, 5.0, 8.0, 13.0], device="cuda")
    A = _Diag(d.to(dtype))
    evals, evecs = lanczos_eigsh(A, k=4, which="LA", tol=1e-6 if dtype == torch.float32 else 1e-3)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert torch.allclose(evals.float(), torch.tensor([3.0, 5.0, 8.0, 13.0], device="cuda"), rtol=tol)
    assert evecs.dtype == dtype and evecs.shape == (n, 4)
    # eigenvectors of a diagonal matrix: unit vectors on the last four coordinates
    assert (evecs.float()[-4:].abs() - torch.eye(4, device="cuda")).abs().max() < (1e-3 if dtype == torch.float32 else 5e-2)
