"""Multi-head self-attention (CURV_OP_ATTENTION, csrc/attention.cuh): GGN / MC-GGN / Jacobian products of transformer
encoder blocks (nn.MultiheadAttention, batch_first) against the oracle's float64 restatement evaluated on the GPU
(torch's math attention path, differentiable twice), fp32 (rtol 1e-4) and bf16 operators (1e-2), several heads / layers,
two unequal mini-batches, a parameter subset that leaves the attention projections constant."""
import pytest
import torch
from torch import nn

from curvlinops_b200 import GGNLinearOperator, JacobianLinearOperator, TransposedJacobianLinearOperator
from oracle import curvature_oracle as orc
from oracle.models import TransformerBlock
from tests.golden_utils import flat, split_like

pytestmark = pytest.mark.gpu


def _setup(dim, heads, layers, T, dtype=torch.float32, seed=0):
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    model = TransformerBlock(dim=dim, heads=heads, hidden=2 * dim, layers=layers).to(dev).eval()
    with torch.no_grad():  # default init leaves the attention nearly uniform: sharpen it
        for blk in model.blocks:
            blk.attn.in_proj_weight.mul_(3.0)
            blk.attn.in_proj_bias.normal_(0.0, 0.3)
    model = model.to(dtype)
    data = [(torch.randn(5, T, dim, device=dev).to(dtype), torch.randint(0, 5, (5,), device=dev)),
            (torch.randn(3, T, dim, device=dev).to(dtype), torch.randint(0, 5, (3,), device=dev))]
    m64 = TransformerBlock(dim=dim, heads=heads, hidden=2 * dim, layers=layers).to(dev).double().eval()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    data64 = [(X.double(), y) for X, y in data]
    return model, m64, data, data64


def _close(got, ref, tol):
    got, ref = got.double(), ref.double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert torch.allclose(got, ref, rtol=tol, atol=tol * 0.1 * ref.abs().max().item()), err
    return err


@pytest.mark.parametrize("dim,heads,layers,T", [(16, 2, 1, 7), (32, 4, 2, 13), (64, 1, 1, 70)])
def test_ggn_matches_float64_oracle(dim, heads, layers, T):
    model, m64, data, data64 = _setup(dim, heads, layers, T)
    loss = nn.CrossEntropyLoss()
    params, p64 = dict(model.named_parameters()), dict(m64.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, 3, device="cuda", dtype=torch.float64)
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V, p64)))
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    print("attention GGN err", _close(G @ V.float(), ref, 1e-4))
    # symmetry of the operator on the same vectors
    VtGV = V.float().T @ (G @ V.float())
    torch.testing.assert_close(VtGV, VtGV.T, rtol=1e-4, atol=1e-6 * VtGV.abs().max().item())


def test_jacobian_and_its_transpose():
    model, m64, data, data64 = _setup(16, 2, 1, 7, seed=1)
    params, p64 = dict(model.named_parameters()), dict(m64.named_parameters())
    J = JacobianLinearOperator(model, params, data, check_deterministic=False)
    JT = TransposedJacobianLinearOperator(model, params, data, check_deterministic=False)
    torch.manual_seed(2)
    v = torch.rand(J.shape[1], 2, device="cuda")
    w = torch.rand(J.shape[0], 2, device="cuda")
    Jv, JTw = J @ v, JT @ w
    # adjoint identity <w, J v> = <J^T w, v> and J v against float64 autograd
    torch.testing.assert_close((w * Jv).sum(0), (JTw * v).sum(0), rtol=1e-4, atol=1e-5)
    f_fn = orc._as_callable(m64)
    cols = split_like(v.double(), p64)
    ref = torch.cat([torch.stack([orc.jacobian_vector_product(f_fn, p64, X, [c[..., k] for c in cols])[1]
                                  for k in range(2)], dim=-1) for X, _ in data64]).reshape(-1, 2)
    _close(Jv, ref, 1e-4)


def test_parameter_subset_and_mc():
    model, m64, data, data64 = _setup(16, 2, 2, 9, seed=3)
    loss = nn.CrossEntropyLoss()
    names = ["blocks.1.fc1.weight", "blocks.0.attn.out_proj.weight", "blocks.0.ln1.bias", "head.bias"]
    allp, all64 = dict(model.named_parameters()), dict(m64.named_parameters())
    params, p64 = {n: allp[n] for n in names}, {n: all64[n] for n in names}
    V = torch.rand(sum(p.numel() for p in params.values()), 2, device="cuda", dtype=torch.float64)
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V, p64)))
    _close(GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V.float(), ref, 1e-4)
    # MC-GGN: positive semi-definite, deterministic for a seed, right scale in expectation (coarse)
    G = GGNLinearOperator(model, loss, allp, data, check_deterministic=False, mc_samples=4, seed=7)
    v = torch.rand(G.shape[1], device="cuda")
    a, b = G @ v, G @ v
    assert torch.equal(a, b) and torch.dot(v, a) >= 0


def test_bf16_operator():
    model, m64, data, data64 = _setup(32, 4, 1, 13, dtype=torch.bfloat16, seed=4)
    loss = nn.CrossEntropyLoss()
    params, p64 = dict(model.named_parameters()), dict(m64.named_parameters())
    V = torch.rand(sum(p.numel() for p in params.values()), 2, device="cuda").to(torch.bfloat16)
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V.double(), p64)))
    got = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V
    assert got.dtype == torch.bfloat16
    err = (got.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, err


def _mini_vit(dtype=torch.float32, seed=0, layers=2):
    from torchvision.models.vision_transformer import VisionTransformer

    torch.manual_seed(seed)
    kw = dict(image_size=32, patch_size=8, num_layers=layers, num_heads=2, hidden_dim=32, mlp_dim=64, num_classes=5)
    model = VisionTransformer(**kw).cuda().eval()
    with torch.no_grad():  # torchvision zero-initialises the head; give every parameter a generic value
        model.heads.head.weight.normal_(0.0, 0.3)
        model.heads.head.bias.normal_(0.0, 0.1)
        model.class_token.normal_(0.0, 0.5)
        model.encoder.pos_embedding.normal_(0.0, 0.5)
        for blk in model.encoder.layers:
            blk.self_attention.in_proj_weight.mul_(3.0)
            blk.self_attention.in_proj_bias.normal_(0.0, 0.3)
    model = model.to(dtype)
    m64 = VisionTransformer(**kw).cuda().double().eval()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    data = [(torch.rand(4, 3, 32, 32, device="cuda").to(dtype), torch.randint(0, 5, (4,), device="cuda")),
            (torch.rand(3, 3, 32, 32, device="cuda").to(dtype), torch.randint(0, 5, (3,), device="cuda"))]
    return model, m64, data, [(X.double(), y) for X, y in data]


def test_vision_transformer_ggn_matches_float64_oracle():
    """torchvision's VisionTransformer end to end: patch convolution read as tokens, class token, position embedding,
    encoder blocks with attention, class-token read-out; every parameter (incl. class_token / pos_embedding)."""
    model, m64, data, data64 = _mini_vit()
    loss = nn.CrossEntropyLoss()
    params, p64 = dict(model.named_parameters()), dict(m64.named_parameters())
    V = torch.rand(sum(p.numel() for p in params.values()), 3, device="cuda", dtype=torch.float64)
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V, p64)))
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    got = G @ V.float()
    print("ViT GGN err", _close(got, ref, 1e-4))
    # per-parameter check of the two broadcast parameters
    o = 0
    for n, p in params.items():
        if n in ("class_token", "encoder.pos_embedding"):
            blk = slice(o, o + p.numel())
            assert ref[blk].abs().max() > 0
            _close(got[blk], ref[blk], 1e-4)
        o += p.numel()


def test_vision_transformer_bf16_mc_and_subset():
    model, m64, data, data64 = _mini_vit(dtype=torch.bfloat16, seed=1)
    loss = nn.CrossEntropyLoss()
    params, p64 = dict(model.named_parameters()), dict(m64.named_parameters())
    V = torch.rand(sum(p.numel() for p in params.values()), 2, device="cuda").to(torch.bfloat16)
    ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V.double(), p64)))
    got = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V
    err = (got.double() - ref).abs().max().item() / ref.abs().max().item()
    # bf16 operators round the attention operands (q, k, v, the softmax matrix and its tangents) to bf16 for the
    # tensor-core products; the attention of this model is deliberately sharpened (in_proj x 3)
    assert got.dtype == torch.bfloat16 and err < 4e-2, err
    # class token and position embedding constant, encoder weights only; MC-GGN runs and is repeatable
    sub = {n: p for n, p in params.items() if "mlp" in n or "in_proj" in n}
    G = GGNLinearOperator(model, loss, sub, data, check_deterministic=False, mc_samples=2, seed=11)
    v = torch.rand(G.shape[1], device="cuda").to(torch.bfloat16)
    assert torch.equal(G @ v, G @ v)


def test_kfac_factors_of_vision_transformer_layers():
    """KFAC (type-2, joint weight + bias) of the Linear / Conv2d layers of a vision transformer -- the patch convolution,
    the MLP layers (weights shared over the 17 tokens) and the head -- with the attention, LayerNorm, class token and
    position embedding as constants in between: factors vs the oracle's hook-based float64 restatement on the GPU."""
    from curvlinops_b200 import KFACLinearOperator

    model, m64, data, data64 = _mini_vit(seed=5, layers=1)
    loss = nn.CrossEntropyLoss()
    layers = ["conv_proj", "encoder.layers.encoder_layer_0.mlp.0", "encoder.layers.encoder_layer_0.mlp.3", "heads.head"]
    params = {f"{n}.{r}": dict(model.named_parameters())[f"{n}.{r}"] for n in layers for r in ("weight", "bias")}
    A64, G64 = orc.kfac_factors(m64, loss, layers, data64, fisher_type="type2", joint_bias=True)
    K = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", separate_weight_and_bias=False,
                           check_deterministic=False)
    _, blocks, _ = K
    for n, blk in zip(layers, blocks):
        Gf, Af = list(blk)
        _close(Af, A64[n], 1e-4)
        _close(Gf, G64[n], 1e-4)


@pytest.mark.parametrize("name", ["transformer_block_ce_mean", "mini_vit_ce_mean"])
def test_attention_fixtures_generated_by_the_reference(name):
    """GGN and empirical Fisher against the REFERENCE's own outputs (tests/golden, oracle/make_golden_attention.py), and
    the MC-GGN with the would-be gradients of the reference's RNG stream handed to the engine."""
    from curvlinops_b200 import EFLinearOperator
    from tests.golden_utils import load_case

    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    V = fx["V"].float().cuda()
    _close(GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V, fx["ggn"].cuda(), 1e-4)
    _close(EFLinearOperator(model, loss, params, data, check_deterministic=False) @ V, fx["ef"].cuda(), 1e-4)
    cpu_model, _, cpu_data, _ = load_case(name)
    with torch.random.fork_rng():
        torch.manual_seed(1234)  # the reference's stream: one multinomial per mini-batch after manual_seed(seed)
        gs = [orc.mc_grad_outputs(loss, cpu_model(X).detach(), 3) for X, _ in cpu_data]
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=3, seed=1234)
    G._mc_grad_override = [g.float() for g in gs]
    _close(G @ V, fx["ggn_mc3"].cuda(), 1e-4)
