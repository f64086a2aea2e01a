"""bf16 operators (BASELINE.json configs C3-C5 are bf16; the reference runs bf16 products with bf16 results,
``curvlinops/_torch_base.py:586-589``).  The engine computes the contractions as ONE bf16 ``tcgen05.mma`` per
product (fp32 accumulation in tensor memory) and keeps everything between the contractions in fp32.

Oracle: the float64 restatement on IDENTICAL bf16-representable inputs (parameters, data and V rounded to bf16
first; SURVEY 8c: the fp32/fp64 reference on bf16-representable inputs is the bf16 oracle).  Tolerance: north_star's
rtol = 1e-2 for bf16, with atol = 1e-2 max|ref| (entries far below the matrix scale carry the bf16 rounding of the
large ones)."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator, HessianLinearOperator, _capi as capi
from oracle import curvature_oracle as orc
from tests.golden_utils import flat, load_case, split_like

pytestmark = pytest.mark.gpu
RTOL = 1e-2


def _bf16_case(name):
    """(bf16 model on cuda, float64 model with the same rounded values, loss, bf16 data, float64 data, V)"""
    m64, loss, data64, fx = load_case(name, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        for t in list(m64.parameters()) + list(m64.buffers()):
            if t.is_floating_point():
                t.copy_(t.to(torch.bfloat16).double())
    data64 = [(X.to(torch.bfloat16).double(), (y.to(torch.bfloat16).double() if y.is_floating_point() else y))
              for X, y in data64]
    mb, _, _, _ = load_case(name, dtype=torch.float64, device="cuda")
    mb.load_state_dict(m64.state_dict())
    mb = mb.to(torch.bfloat16)
    datab = [(X.to(torch.bfloat16), (y.to(torch.bfloat16) if y.is_floating_point() else y)) for X, y in data64]
    V = fx["V"].to(torch.bfloat16).cuda()
    return mb, m64, loss, datab, data64, V


def _check(got, ref, what, tol=RTOL):
    assert got.dtype == torch.bfloat16  # bf16 operator: bf16 result, like the reference
    got, ref = got.double().cpu(), ref.double().cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    viol = (~torch.isclose(got, ref, rtol=tol, atol=tol * scale)).double().mean().item()
    print(f"{what}: max|err|/max|ref| = {err:.3e}, violations of isclose(rtol {tol:g}, atol {tol:g} max) = {viol:.2e}")
    assert viol == 0.0 and err < tol, (what, err, viol)


@pytest.mark.parametrize("mode", [2, 1])  # 2: every contraction on the tcgen05 bf16 kernels, 1: size-based default
@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "mlp_c1_mse_mean", "miniresnet_ce_mean", "cnn_bias_ce_mean"])
def test_bf16_ggn_matches_float64_oracle_on_bf16_inputs(name, mode):
    mb, m64, loss, datab, data64, V = _bf16_case(name)
    p64 = dict(m64.named_parameters())
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V.double(), p64)))
    old = capi.lib().curv_set_tensor_core_mode(mode)
    try:
        pb = dict(mb.named_parameters())
        G = GGNLinearOperator(mb, loss, pb, datab, check_deterministic=False)
        assert G.dtype == torch.bfloat16
        got = G @ V
        assert torch.equal(got, G @ V)  # deterministic
    finally:
        capi.lib().curv_set_tensor_core_mode(old)
    _check(got, ref, f"bf16 GGN {name} mode {mode}")


@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "miniresnet_ce_mean"])
def test_bf16_hessian_and_mc(name):
    mb, m64, loss, datab, data64, V = _bf16_case(name)
    p64, pb = dict(m64.named_parameters()), dict(mb.named_parameters())
    ref = flat(orc.hessian_matmat(m64, loss, p64, data64, split_like(V.double(), p64)))
    _check(HessianLinearOperator(mb, loss, pb, datab, check_deterministic=False) @ V, ref, f"bf16 Hessian {name}")
    # MC-GGN: the oracle (CPU, float64) and the engine use the SAME would-be gradients (drawn from the oracle's
    # stream after manual_seed, handed to the engine), so the comparison is exact rather than in expectation
    mc, pc = m64.cpu(), None
    pc = dict(mc.named_parameters())
    datac = [(X.cpu(), y.cpu()) for X, y in data64]
    gs = []
    with torch.random.fork_rng():
        torch.manual_seed(7)
        for X, _ in datac:
            gs.append(orc.mc_grad_outputs(loss, mc(X).detach(), 2).float())
    refmc = flat(orc.ggn_matmat(mc, loss, pc, datac, split_like(V.double().cpu(), pc), mc_samples=2, seed=7))
    G = GGNLinearOperator(mb, loss, pb, datab, check_deterministic=False, mc_samples=2, seed=7)
    G._mc_grad_override = gs
    old = capi.lib().curv_set_tensor_core_mode(2)
    try:
        got = G @ V
    finally:
        capi.lib().curv_set_tensor_core_mode(old)
    # rank-2 loss Hessian: the inner products <g_m, J v> cancel, which amplifies the bf16 rounding of the operands
    # (any bf16 evaluation, the reference's included, carries it) -- 3e-2 here, 1e-2 for the full-rank kinds above
    _check(got, refmc, f"bf16 MC-GGN {name}", tol=3e-2)


def test_bf16_resnet18_matches_float64_oracle():
    """ResNet-18 at 224 x 224 (the C2 / C3 / C5 model family), B = 16, K = 4, bf16 operator, default kernel choice."""
    import torchvision

    torch.manual_seed(0)
    dev = torch.device("cuda")
    mb = torchvision.models.resnet18().eval().to(dev).to(torch.bfloat16)
    X = torch.rand(16, 3, 224, 224, device=dev).to(torch.bfloat16)
    y = torch.randint(0, 1000, (16,), device=dev)
    pb = dict(mb.named_parameters())
    sizes = [p.numel() for p in pb.values()]
    V = torch.rand(sum(sizes), 4, device=dev).to(torch.bfloat16)
    loss = torch.nn.CrossEntropyLoss()
    G = GGNLinearOperator(mb, loss, pb, [(X, y)], check_deterministic=False)
    got = G @ V
    m64 = torchvision.models.resnet18().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in mb.state_dict().items()})
    p64 = dict(m64.named_parameters())
    ref = flat(orc.ggn_matmat(m64, loss, p64, [(X.double(), y)], split_like(V.double(), p64)))
    _check(got, ref, "bf16 GGN ResNet-18 B=16 K=4")
