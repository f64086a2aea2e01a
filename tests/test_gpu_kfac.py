"""GPU parity tests of the KFAC / EKFAC operators against reference-generated fixtures.
Tolerance: rtol 1e-4 (fp32 engine vs float64 reference), atol 1e-5 * max|ref|; the damped inverse
amplifies rounding by the factors' condition number, hence rtol 1e-3 there."""
import pytest
import torch

from curvlinops_b200 import (EKFACLinearOperator, KFACLinearOperator, KroneckerProductLinearOperator,
                             EighDecomposedLinearOperator)
from curvlinops_b200.kfac import KFACComputer
from oracle import curvature_oracle as orc
from tests.golden_utils import load_case

pytestmark = pytest.mark.gpu
# kfac_tokens: Linear layers shared over the T positions of [B, T, D] sequences (KFAC-expand weight sharing)
CASES = ["kfac_mlp", "kfac_cnn", "kfac_tokens"]
# kfac_tied: one Conv2d and one Linear used twice each; fixture from the reference's make_fx backend (usages concatenated
# along the weight-sharing axis).  KFAC only: the eigenvalue correction rejects tied weights.
KFAC_CASES = CASES + ["kfac_tied"]


def close(got, ref, rtol=1e-4, atol_scale=1e-5):
    got, ref = got.detach().double().cpu(), ref.double().cpu()
    atol = atol_scale * ref.abs().max().item()
    assert torch.allclose(got, ref, rtol=rtol, atol=atol), \
        f"max abs err {(got - ref).abs().max():.3e} vs max|ref| {ref.abs().max():.3e}"


def setup(name):
    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    pnames = [str(s) for s in fx["param_names"]]
    params = {n: p for n, p in model.named_parameters() if n in pnames}
    assert list(params) == pnames
    return model, loss, data, fx, params


@pytest.mark.parametrize("name", KFAC_CASES)
@pytest.mark.parametrize("sep", [False, True])
def test_kfac_type2(name, sep):
    model, loss, data, fx, params = setup(name)
    tag = "sep" if sep else "joint"
    Kop = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", separate_weight_and_bias=sep,
                             check_deterministic=False)
    P, K, PT = Kop
    for bi, block in enumerate(K):
        for fi, fac in enumerate(block):
            close(fac, fx[f"factor_type2_{tag}_{bi}_{fi}"])
    v = fx["v"].float().cuda()
    close(Kop @ v, fx[f"kfac_type2_{tag}"])
    close(Kop.inverse(damping=float(fx["damping"])) @ v, fx[f"kfacinv_type2_{tag}"], rtol=1e-3)
    # properties vs the dense Kronecker products
    dense = torch.block_diag(*[torch.kron(*list(b)) if len(b) == 2 else b[0] for b in K]).double()
    torch.testing.assert_close(Kop.trace().double(), dense.trace(), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(Kop.frobenius_norm().double(), dense.norm(), rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("name", KFAC_CASES)
@pytest.mark.parametrize("ft", ["mc", "empirical"])
def test_kfac_sampled_and_empirical(name, ft, monkeypatch):
    """MC: the engine is handed the would-be gradients the reference drew (same generator stream)."""
    model, loss, data, fx, params = setup(name)
    if ft == "mc":
        cpu_model, _, cpu_data, _ = load_case(name)
        gos = orc.kfac_grad_outputs(cpu_model, loss, cpu_data, "mc", mc_samples=2, seed=77)
        monkeypatch.setattr(KFACComputer, "_TEST_GRAD_OUTPUTS", [g.float() for g in gos])
    Kop = KFACLinearOperator(model, loss, params, data, fisher_type=ft, mc_samples=2 if ft == "mc" else 1,
                             seed=77, separate_weight_and_bias=False, check_deterministic=False)
    v = fx["v"].float().cuda()
    close(Kop @ v, fx[f"kfac_{ft}_joint"])
    close(Kop.inverse(damping=float(fx["damping"])) @ v, fx[f"kfacinv_{ft}_joint"], rtol=1e-3)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("sep", [False, True])
def test_ekfac_type2(name, sep):
    model, loss, data, fx, params = setup(name)
    tag = "sep" if sep else "joint"
    E = EKFACLinearOperator(model, loss, params, data, fisher_type="type-2", separate_weight_and_bias=sep,
                            check_deterministic=False)
    v = fx["v"].float().cuda()
    close(E @ v, fx[f"ekfac_type2_{tag}"], rtol=1e-3)
    # fp32 eigenvectors of near-degenerate factors: entries that are tiny relative to the matrix scale carry
    # an absolute error ~1e-5 * max|ref| after the damped inversion
    close(E.inverse(damping=float(fx["damping"])) @ v, fx[f"ekfacinv_type2_{tag}"], rtol=1e-3, atol_scale=1e-4)


def test_kronecker_and_eigh_operators_vs_dense():
    torch.manual_seed(0)
    G, A = torch.rand(5, 5, device="cuda"), torch.rand(7, 7, device="cuda")
    G, A = G @ G.T + torch.eye(5, device="cuda"), A @ A.T + torch.eye(7, device="cuda")
    Kop = KroneckerProductLinearOperator(G, A)
    X = torch.rand(35, 3, device="cuda")
    dense = torch.kron(G, A)
    torch.testing.assert_close(Kop @ X, dense @ X, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(X.T @ Kop, X.T @ dense, rtol=1e-4, atol=1e-5)
    for kw in ({"damping": 0.1}, {"damping": 0.1, "use_heuristic_damping": True},
               {"damping": 0.1, "use_exact_damping": True}):
        inv = Kop.inverse(**kw) @ X
        if kw.get("use_exact_damping"):
            ref = torch.linalg.solve(dense + 0.1 * torch.eye(35, device="cuda"), X)
        elif kw.get("use_heuristic_damping"):
            pi = (A.diag().mean() / G.diag().mean()).sqrt()
            d1, d2 = 0.1 ** 0.5 / pi, 0.1 ** 0.5 * pi
            ref = torch.kron(torch.linalg.inv(G + d1 * torch.eye(5, device="cuda")),
                             torch.linalg.inv(A + d2 * torch.eye(7, device="cuda"))) @ X
        else:
            ref = torch.kron(torch.linalg.inv(G + 0.1 * torch.eye(5, device="cuda")),
                             torch.linalg.inv(A + 0.1 * torch.eye(7, device="cuda"))) @ X
        torch.testing.assert_close(inv, ref, rtol=1e-3, atol=1e-4)
    lam, Q = torch.linalg.eigh(dense)
    Eop = EighDecomposedLinearOperator(lam, Q)
    torch.testing.assert_close(Eop @ X, dense @ X, rtol=1e-3, atol=1e-4)
    with pytest.raises(ValueError, match="Eigenvalues must be 1D"):
        EighDecomposedLinearOperator(lam.unsqueeze(0), Q)
    with pytest.raises(ValueError, match="Eigenvectors must be square"):
        EighDecomposedLinearOperator(lam, Q[:, :3])


def test_unsupported_params_raise():
    model, loss, data, fx, _ = setup("kfac_cnn")
    from oracle.models import MiniResNet
    net = MiniResNet().cuda().eval()
    X, y = torch.rand(2, 3, 32, 32, device="cuda"), torch.randint(0, 10, (2,), device="cuda")
    with pytest.raises(NotImplementedError, match="not in supported layers"):
        KFACLinearOperator(net, loss, dict(net.named_parameters()), [(X, y)], check_deterministic=False)
    with pytest.raises(ValueError, match="Invalid fisher_type"):
        KFACLinearOperator(model, loss, dict(model.named_parameters()), data, fisher_type="nope")


@pytest.mark.parametrize("name", KFAC_CASES)
def test_kfac_type2_on_tensor_core_kernels(name):
    """Same factors with every contraction forced onto the tcgen05 kernels (half-split forward / dgrad sweeps,
    3xTF32 Gram matrices): exercises the tensor-core paths at the fixtures' small shapes."""
    from curvlinops_b200 import _capi as capi

    model, loss, data, fx, params = setup(name)
    old = capi.lib().curv_set_tensor_core_mode(2)
    try:
        Kop = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", separate_weight_and_bias=False,
                                 check_deterministic=False)
        P, K, PT = Kop
        for bi, block in enumerate(K):
            for fi, fac in enumerate(block):
                close(fac, fx[f"factor_type2_joint_{bi}_{fi}"])
        close(Kop @ fx["v"].float().cuda(), fx["kfac_type2_joint"])
    finally:
        capi.lib().curv_set_tensor_core_mode(old)


@pytest.mark.parametrize("sep", [False, True])
def test_ekfac_device_correction_matches_host_path(sep, monkeypatch):
    """The eigenvalue correction on the tensor-core kernels (csrc/ekfac.cuh: rotations as GEMMs, per-example contraction
    split at the example boundaries) against the host-orchestrated path of round 1 (F.unfold + dense products) on a
    small residual CNN: conv / linear groups with and without joint bias, bias-only groups, several positions."""
    from curvlinops_b200.kfac import EKFACComputer
    from oracle.models import MiniResNet, randomize_bn_

    torch.manual_seed(0)
    model = MiniResNet().eval()
    randomize_bn_(model, torch.Generator().manual_seed(0))
    model = model.cuda()
    mods = dict(model.named_modules())
    names = [n for n, m in mods.items() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    params = {f"{n}.{pn}": p for n in names for pn, p in mods[n].named_parameters(recurse=False)}
    X, y = torch.rand(6, 3, 32, 32, device="cuda"), torch.randint(0, 10, (6,), device="cuda")
    loss = torch.nn.CrossEntropyLoss()
    res = {}
    for device_path in (True, False):
        monkeypatch.setattr(EKFACComputer, "DEVICE_CORRECTION", device_path)
        comp = EKFACComputer(model, loss, params, [(X, y)], fisher_type="type-2", separate_weight_and_bias=sep,
                             check_deterministic=False)
        res[device_path] = comp.compute()
    QA1, QG1, lam_dev, mapping = res[True]
    _, _, lam_host, _ = res[False]
    assert set(lam_dev) == set(lam_host)
    for k in lam_dev:
        assert lam_dev[k].shape == lam_host[k].shape, k
        close(lam_dev[k], lam_host[k], rtol=1e-4, atol_scale=1e-5)


def test_ekfac_rejects_tied_weights():
    model, loss, data, fx, params = setup("kfac_tied")
    with pytest.raises(NotImplementedError, match="Weight tying"):
        EKFACLinearOperator(model, loss, params, data, fisher_type="type-2", check_deterministic=False)


@pytest.mark.parametrize("sep", [False, True])
def test_kfac_reduce_on_token_model(sep, monkeypatch):
    """KFAC-reduce (inputs averaged, output gradients summed over the T positions) against the reference's own factors
    and products (fixture kfac_tokens_reduce); type-2, empirical, and MC with the reference's draws."""
    model, loss, data, fx, params = setup("kfac_tokens_reduce")
    tag = "sep" if sep else "joint"
    kw = dict(separate_weight_and_bias=sep, check_deterministic=False, kfac_approx="reduce")
    Kop = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", **kw)
    _, K, _ = Kop
    for bi, block in enumerate(K):
        for fi, fac in enumerate(block):
            close(fac, fx[f"factor_type2_{tag}_{bi}_{fi}"])
    v = fx["v"].float().cuda()
    close(Kop @ v, fx[f"kfac_type2_{tag}"])
    close(Kop.inverse(damping=float(fx["damping"])) @ v, fx[f"kfacinv_type2_{tag}"], rtol=1e-3)
    close(KFACLinearOperator(model, loss, params, data, fisher_type="empirical", **kw) @ v, fx[f"kfac_empirical_{tag}"])
    cpu_model, _, cpu_data, _ = load_case("kfac_tokens_reduce")
    gos = orc.kfac_grad_outputs(cpu_model, loss, cpu_data, "mc", mc_samples=2, seed=77)
    monkeypatch.setattr(KFACComputer, "_TEST_GRAD_OUTPUTS", [g.float() for g in gos])
    close(KFACLinearOperator(model, loss, params, data, fisher_type="mc", mc_samples=2, seed=77, **kw) @ v,
          fx[f"kfac_mc_{tag}"])
