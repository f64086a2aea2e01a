"""Full-size checks on BASELINE.json configs[1] (ResNet-18, 3x224x224, GGN): parity against the oracle's float64
restatement evaluated on the GPU at a reduced batch (the oracle needs ~20 GB at B = 128), and size-independent
properties of the product at the full batch of 128 with 8 columns: symmetry, linearity, bit-wise repeatability.
Tolerance: BASELINE north_star rtol 1e-4 (fp32), relative to the largest entry of the result."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator
from oracle import curvature_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(batch, seed=0):
    import torchvision

    torch.manual_seed(seed)
    dev = torch.device("cuda")
    model = torchvision.models.resnet18().eval().to(dev)
    X = torch.rand(batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (batch,), device=dev)
    return model, X, y


def test_resnet18_ggn_matches_float64_oracle():
    import torchvision

    model, X, y = _problem(16)
    dev = X.device
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, 2, device=dev)
    loss = torch.nn.CrossEntropyLoss()
    m64 = torchvision.models.resnet18().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    Vl = [v.reshape(*p.shape, 2).double() for v, p in zip(V.split([p.numel() for p in p64.values()]), p64.values())]
    ref = torch.cat([r.reshape(-1, 2) for r in orc.ggn_matmat(m64, loss, p64, [(X.double(), y)], Vl)])
    G = GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False)
    got = (G @ V).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"ResNet-18 GGN, B=16, K=2: max|err|/max|ref| = {err:.3e}")
    assert err < 1e-4, err


def test_resnet18_full_batch_properties():
    model, X, y = _problem(128)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False,
                          num_data=128)
    gen = torch.Generator(device="cuda").manual_seed(5)
    V = torch.rand(P, 8, device="cuda", generator=gen)
    GV = G @ V
    scale = GV.abs().max().item()
    assert torch.isfinite(GV).all() and scale > 0
    # repeatability: no floating-point atomics, fixed reduction orders -> bit-identical (eager and graph replay)
    assert torch.equal(GV, G @ V)
    assert torch.equal(GV, G @ V)
    # symmetry of the GGN: V^T (G V) is a symmetric 8 x 8 matrix
    S = (V.double().T @ GV.double())
    assert (S - S.T).abs().max().item() <= 1e-4 * S.abs().max().item()
    # linearity: G (V c) = (G V) c for a mixing matrix c (columns recombined)
    c = torch.rand(8, 8, device="cuda", generator=gen)
    lhs = G @ (V @ c)
    rhs = GV @ c
    assert (lhs - rhs).abs().max().item() <= 1e-4 * rhs.abs().max().item()
    # positive semi-definiteness (up to rounding): diagonal of V^T G V is non-negative
    assert (torch.diagonal(S) >= -1e-6 * S.abs().max()).all()
