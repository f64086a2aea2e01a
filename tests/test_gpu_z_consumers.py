"""GPU tests of the consumers of the engine's products: CG / Neumann inverses of the
damped GGN, with a KFAC inverse as preconditioner, and the randomised trace / diagonal estimators.  The checks
themselves live in ``tests/consumer_checks.py`` and are validated on the CPU against dense operators; here the
operator is the engine's GGN (fp32) and the dense matrix is that same GGN applied to the identity.

Also here: the BCEWithLogitsLoss and the Sigmoid / Tanh parity cases (all green on the B200 since round 1's final run)."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator, KFACLinearOperator
from tests.consumer_checks import check_damped_inverses, check_estimators
from tests.golden_utils import load_case

pytestmark = pytest.mark.gpu


def _ggn(name):
    model, loss, data, _ = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    dense = G @ torch.eye(G.shape[1], device=G.device, dtype=G.dtype)
    return model, loss, data, params, G, dense


@pytest.mark.parametrize("name", ["kfac_mlp"])
def test_inverses_of_damped_ggn(name):
    model, loss, data, params, G, dense = _ggn(name)
    shapes = [tuple(p.shape) for p in params.values()]
    delta = 0.1 * dense.diag().mean().item()
    kfac_inv = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", check_deterministic=False
                                  ).inverse(damping=delta, use_exact_damping=True)  # (G (x) A + delta I)^-1 per block
    iters = check_damped_inverses(G, dense, shapes, preconditioner=kfac_inv.__matmul__, delta_rel=0.1)
    assert all(1 <= n <= 300 for n in iters.values()), iters


@pytest.mark.parametrize("name", ["kfac_mlp"])
def test_estimators_on_ggn(name):
    *_, G, dense = _ggn(name)
    check_estimators(G, dense)


# ---- BCEWithLogitsLoss (fixtures of oracle/make_golden_bce.py): the third loss of the reference's test matrix;
# Sigmoid / Tanh activations under MSELoss(sum) (oracle/make_golden_act.py): second-order terms of the Hessian R-op
# plain CNN with conv biases, stride-2 un-padded conv, flatten -> Linear over a 3x3 map (oracle/make_golden_cnn.py)
BCE_CASES = ["mlp_bce_mean", "mlp_bce_sum", "mlp_sigmoid_tanh_mse_sum", "cnn_bias_ce_mean"]


def _parity(got, ref, rtol=1e-4):
    """rtol 1e-4 (BASELINE tolerance, fp32 engine vs float64 reference), atol 1e-5 * max|ref|."""
    got, ref = got.detach().double().cpu(), ref.double().cpu()
    assert torch.allclose(got, ref, rtol=rtol, atol=1e-5 * ref.abs().max().item()), \
        f"max abs err {(got - ref).abs().max():.3e} vs max|ref| {ref.abs().max():.3e}"


@pytest.mark.parametrize("name", BCE_CASES)
def test_bce_curvature_matches_reference_golden(name):
    from curvlinops_b200 import EFLinearOperator, HessianLinearOperator

    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    V = fx["V"].float().cuda()
    for cls, key in [(GGNLinearOperator, "ggn"), (HessianLinearOperator, "hessian"), (EFLinearOperator, "ef")]:
        op = cls(model, loss, params, data, check_deterministic=False)
        _parity(op @ V, fx[key])


@pytest.mark.parametrize("name", BCE_CASES)
@pytest.mark.parametrize("M", [1, 3])
def test_bce_mc_ggn_with_reference_samples(name, M):
    """MC-GGN handed the would-be gradients the reference drew (same seed => same CPU stream)."""
    from oracle import curvature_oracle as orc

    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    cpu_model = load_case(name)[0]
    gs = []
    with torch.random.fork_rng():
        torch.manual_seed(1234)
        for X, _ in data:
            gs.append(orc.mc_grad_outputs(loss, cpu_model(X.double().cpu()).detach(), M).float())
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=M, seed=1234)
    G._mc_grad_override = gs
    _parity(G @ fx["V"].float().cuda(), fx[f"ggn_mc{M}"])


# ---- dict-like inputs (reference test/cases.py:36-60,148-168): same products as with the tensor itself --------
def test_dict_like_inputs_match_tensor_inputs():
    from collections import UserDict

    from curvlinops_b200 import HessianLinearOperator

    model, loss, data, fx = load_case("mlp_c1_ce_mean", dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())

    class OnDict(torch.nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, batch):
            return self.net(batch["x"].to(next(self.parameters()).device))

    wrapped = OnDict(model)
    wparams = dict(wrapped.named_parameters())
    ddata = [(UserDict({"x": X.cpu(), "note": "kept"}), y.cpu()) for X, y in data]  # host-resident mappings
    V = fx["V"].float().cuda()
    for cls in (GGNLinearOperator, HessianLinearOperator):
        ref = cls(model, loss, params, data, check_deterministic=False) @ V
        got = cls(wrapped, loss, wparams, ddata, check_deterministic=True,
                  batch_size_fn=lambda b: b["x"].shape[0]) @ V
        assert torch.equal(got, ref)
