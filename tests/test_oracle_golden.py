"""Pins the CPU oracle (oracle/curvature_oracle.py) to fixtures produced by the reference."""
import pytest
import torch

from oracle import curvature_oracle as orc
from tests.golden_utils import flat, load_case, split_like

CASES = ["mlp_c1_ce_mean", "mlp_c1_ce_sum", "mlp_c1_mse_mean", "miniresnet_ce_mean"]
# fixtures of oracle/make_golden_bce.py / make_golden_act.py / make_golden_cnn.py (they carry their own "ef")
BCE_CASES = ["mlp_bce_mean", "mlp_bce_sum", "mlp_sigmoid_tanh_mse_sum", "cnn_bias_ce_mean"]


@pytest.mark.parametrize("name", CASES + BCE_CASES)
def test_ggn_and_hessian_match_reference(name):
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    torch.testing.assert_close(flat(orc.ggn_matmat(model, loss, params, data, V)), fx["ggn"],
                               rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(flat(orc.hessian_matmat(model, loss, params, data, V)),
                               fx["hessian"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES + BCE_CASES)
@pytest.mark.parametrize("M", [1, 3])
def test_mc_ggn_matches_reference_stream(name, M):
    """Same seed => same global-RNG stream as the reference (curvlinops/ggn.py:337-341)."""
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    got = flat(orc.ggn_matmat(model, loss, params, data, V, mc_samples=M, seed=1234))
    torch.testing.assert_close(got, fx[f"ggn_mc{M}"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_empirical_fisher_matches_reference(name):
    """Fixture tests/golden/ef.npz = the reference's EFLinearOperator on the same parameters / data / V
    (oracle/make_golden_ef.py)."""
    import numpy as np
    import os

    from tests.golden_utils import GOLDEN

    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "ef.npz"))[name])
    torch.testing.assert_close(flat(orc.ef_matmat(model, loss, params, data, V)), ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", BCE_CASES)
def test_empirical_fisher_bce_matches_reference(name):
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    torch.testing.assert_close(flat(orc.ef_matmat(model, loss, params, data, V)), fx["ef"], rtol=1e-9, atol=1e-12)


# fixtures of oracle/make_golden_ln.py: LayerNorm / GELU networks, 2-d inputs and token sequences [B, T, D]
LN_CASES = ["mlp_ln_gelu_ce_mean", "token_mlp_ce_mean", "mlp_gelu_mse_sum"]


@pytest.mark.parametrize("name", LN_CASES)
def test_layernorm_gelu_cases_match_reference(name):
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    torch.testing.assert_close(flat(orc.ggn_matmat(model, loss, params, data, V)), fx["ggn"], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(flat(orc.ef_matmat(model, loss, params, data, V)), fx["ef"], rtol=1e-9, atol=1e-12)
    if "hessian" in fx:
        torch.testing.assert_close(flat(orc.hessian_matmat(model, loss, params, data, V)), fx["hessian"],
                                   rtol=1e-9, atol=1e-12)
    for M in (1, 3):
        got = flat(orc.ggn_matmat(model, loss, params, data, V, mc_samples=M, seed=1234))
        torch.testing.assert_close(got, fx[f"ggn_mc{M}"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["ggn_diag_mlp_ce_mean", "ggn_diag_mlp_mse_sum", "ggn_diag_cnn_ce_mean"])
def test_ggn_diagonal_matches_reference(name):
    """oracle.ggn_diagonal vs the reference's GGNDiagonalLinearOperator (fixtures of oracle/make_golden_diag.py)."""
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    got = torch.cat([d.reshape(-1) for d in orc.ggn_diagonal(model, loss, params, data)])
    torch.testing.assert_close(got, fx["diag"], rtol=1e-9, atol=1e-13)


# fixtures of oracle/make_golden_attention.py: nn.MultiheadAttention blocks on token sequences, torchvision's
# VisionTransformer at toy size (the reference ran them through torch.func on the math attention path)
ATTN_CASES = ["transformer_block_ce_mean", "mini_vit_ce_mean"]


@pytest.mark.parametrize("name", ATTN_CASES)
def test_attention_cases_match_reference(name):
    from torch.nn.attention import SDPBackend, sdpa_kernel

    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    V = split_like(fx["V"], params)
    with sdpa_kernel(SDPBackend.MATH):  # the fused CPU attention kernel has no double backward
        torch.testing.assert_close(flat(orc.ggn_matmat(model, loss, params, data, V)), fx["ggn"], rtol=1e-9, atol=1e-12)
        torch.testing.assert_close(flat(orc.ef_matmat(model, loss, params, data, V)), fx["ef"], rtol=1e-9, atol=1e-12)
        got = flat(orc.ggn_matmat(model, loss, params, data, V, mc_samples=3, seed=1234))
    torch.testing.assert_close(got, fx["ggn_mc3"], rtol=1e-9, atol=1e-12)
