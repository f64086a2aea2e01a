"""KFAC at the shapes of BASELINE.json configs[2] (C3): ResNet-18 at 3 x 224 x 224, Conv2d / Linear parameters only,
joint weight+bias groups, MC Fisher with one sample, damped inverse 1e-3 -- as fp32 and as bf16 operator.

Oracle: ``oracle.curvature_oracle.kfac_factors`` in float64 on the GPU with the SAME would-be gradients (drawn once
here and handed to both sides), then the Kronecker apply ``G [W|b] A^T`` and the damped-inverse apply in float64
(for bf16 the oracle is the reference arithmetic in higher precision on identical bf16-representable
inputs -- the reference itself cannot invert in bf16, ``kronecker.py:356-373``).  B = 16 keeps the float64 patch
matrices of the oracle small; the factor kernels see the full C3 layer shapes (21 groups, A up to 4608^2).
Tolerances: north_star rtol 1e-4 (fp32) / 1e-2 (bf16) relative to the largest entry for the A factors and the apply;
for the G factors that bar or, if larger, the worst error torch autograd in the same dtype (the reference's
arithmetic) makes on a G factor of the same problem (measured: 9e-4 in strict fp32, 1.7e-2 in bf16; the engine stays
at 1.8e-4 / 1.3e-2); the damped inverse amplifies the factor rounding by the condition number, hence 50x there."""
import pytest
import torch

from curvlinops_b200 import KFACLinearOperator
from curvlinops_b200.kfac import KFACComputer
from oracle import curvature_oracle as orc

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kfac_resnet18_c3_shapes(dtype, monkeypatch):
    import torchvision

    torch.manual_seed(0)
    dev = torch.device("cuda")
    B = 16
    model = torchvision.models.resnet18().eval().to(dev)
    X = torch.rand(B, 3, 224, 224, device=dev)
    if dtype == torch.bfloat16:
        model, X = model.to(dtype), X.to(dtype)
    y = torch.randint(0, 1000, (B,), device=dev)
    m64 = torchvision.models.resnet18().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    X64 = X.double()
    mods = dict(model.named_modules())
    layer_names = [n for n, m in mods.items() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    params = {f"{n}.{pn}": p for n in layer_names for pn, p in mods[n].named_parameters(recurse=False)}
    loss = torch.nn.CrossEntropyLoss()
    # one would-be gradient per datum, shared by both sides
    with torch.no_grad():
        p = torch.softmax(m64(X64), 1)
    yhat = p.multinomial(1)
    gos = (p.unsqueeze(1) - torch.nn.functional.one_hot(yhat, 1000).double()).permute(1, 0, 2).contiguous()
    monkeypatch.setattr(KFACComputer, "_TEST_GRAD_OUTPUTS", [gos.float()])
    import os

    from curvlinops_b200 import _capi as capi

    diag_mode = os.environ.get("CURV_KFAC_DIAG_MODE")  # manual A/B runs of the Gram kernels (e.g. 0x4001)
    old_mode = capi.lib().curv_set_tensor_core_mode(int(diag_mode, 0)) if diag_mode else None
    try:
        Kop = KFACLinearOperator(model, loss, params, [(X, y)], fisher_type="mc", mc_samples=1,
                                 separate_weight_and_bias=False, check_deterministic=False, num_data=B)
    finally:
        if old_mode is not None:
            capi.lib().curv_set_tensor_core_mode(old_mode)
    A64, G64 = orc.kfac_factors(m64, loss, layer_names, [(X64, y)], n_data=B, fisher_type="mc", joint_bias=True,
                                grad_outputs=[gos])
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    _, K, _ = Kop
    assert len(K) == len(layer_names) == 21
    # how far the SAME computation through torch autograd in the operator's dtype (the reference's arithmetic: cuDNN /
    # cuBLAS kernels, strict fp32 or bf16) lands from the float64 factors: G sums squares of back-propagated vectors
    # over few rows, so single ReLU masks that flip between precisions and the rounding of the sweeps show up at the
    # 1e-4 (fp32) / 1e-2 (bf16) level for ANY evaluation in that dtype
    old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        Ad, Gd = orc.kfac_factors(model, loss, layer_names, [(X, y)], n_data=B, fisher_type="mc", joint_bias=True,
                                  grad_outputs=[gos.to(dtype)])
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
    rows = []
    for name, block in zip(layer_names, K):
        Gf, Af = list(block)
        assert Af.shape == A64[name].shape and Gf.shape == G64[name].shape
        rows.append((name, _rel(Gf, G64[name]), _rel(Af, A64[name]), _rel(Gd[name], G64[name]),
                     _rel(Ad[name], A64[name])))
    print(f"  {'layer':24s} engine G  engine A  | torch-{str(dtype).split('.')[-1]} autograd G, A   (max|err|/max|ref| vs float64)")
    for name, eg, ea, rg, ra in rows:
        print(f"  {name:24s} {eg:.2e}  {ea:.2e}  | {rg:.2e}  {ra:.2e}")
    worst = max(max(r[1], r[2]) for r in rows)
    print(f"{dtype}: 42 factors, worst engine error {worst:.3e}, worst torch-autograd error "
          f"{max(max(r[3], r[4]) for r in rows):.3e}")
    # A factors (one forward pass): the tolerance as is.  G factors: the tolerance, or the worst error torch autograd in
    # this dtype makes on any G of this very problem (per-factor comparisons are noisy: which masks flip is chance)
    g_bar = max(tol, max(r[3] for r in rows))
    bad = [r for r in rows if r[1] >= g_bar or r[2] >= tol]
    assert not bad, bad
    # apply and damped-inverse apply on a random vector, per layer in float64: (i) with the ENGINE's factors (isolates
    # the apply kernels and the damped inversion: the plain tolerance), (ii) with the oracle's factors (end to end: the
    # G-factor bar above carries over to the product)
    P = sum(p.numel() for p in params.values())
    v = torch.rand(P, device=dev).to(dtype)
    eye = lambda n: torch.eye(n, device=dev, dtype=torch.float64)
    master = lambda t: getattr(t, "_curv_fp32", t).double()

    Kinv = Kop.inverse(damping=1e-3)
    got, got_inv = Kop @ v, Kinv @ v
    assert got.dtype == dtype
    own = {name: tuple(master(f) for f in block) for name, block in zip(layer_names, K)}
    own_inv = {name: tuple(master(f) for f in block) for name, block in zip(layer_names, Kinv[1])}

    def apply64(factors):
        parts, o = [], 0
        for name in layer_names:
            m = mods[name]
            W = v[o:o + m.weight.numel()].double().reshape(m.weight.shape[0], -1)
            o += m.weight.numel()
            if m.bias is not None:
                W = torch.cat([W, v[o:o + m.bias.numel()].double().unsqueeze(1)], 1)
                o += m.bias.numel()
            Gx, Ax = factors[name]
            R = Gx @ W @ Ax.T
            parts += [R[:, :-1].reshape(-1), R[:, -1]] if m.bias is not None else [R.reshape(-1)]
        return torch.cat(parts)

    # (i) the apply kernels alone: float64 apply of the engine's own (inverse) factors
    r_own, o = apply64(own), 0
    for name in layer_names:  # per layer, relative to the layer's own largest entry
        n = sum(p.numel() for p in mods[name].parameters(recurse=False))
        print(f"  apply {name:24s} {_rel(got[o:o + n], r_own[o:o + n]):.2e}")
        o += n
    e_apply, e_inv = _rel(got, r_own), _rel(got_inv, apply64(own_inv))
    print(f"{dtype}: apply kernels vs float64 apply of the engine's factors: KFAC {e_apply:.3e}, inverse {e_inv:.3e}")
    assert e_apply < tol and e_inv < tol, (e_apply, e_inv)
    # (ii) the damped inversion (fp32 Cholesky, like the reference: condition number max eig / 1e-3 ~ 1e6) and
    # (iii) everything end to end against the oracle's factors
    inv64 = lambda t: torch.linalg.inv(t + 1e-3 * eye(t.shape[0]))
    e_chol = _rel(got_inv, apply64({n: (inv64(g), inv64(a)) for n, (g, a) in own.items()}))
    oracle = {name: (G64[name], A64[name]) for name in layer_names}
    e_apply = _rel(got, apply64(oracle))
    e_inv = _rel(got_inv, apply64({n: (inv64(g), inv64(a)) for n, (g, a) in oracle.items()}))
    print(f"{dtype}: inverse apply vs float64 inversion of the engine's factors {e_chol:.3e}; end to end vs the oracle's "
          f"factors: KFAC apply {e_apply:.3e}, inverse(1e-3) apply {e_inv:.3e}")
    assert e_apply < 2 * g_bar and e_chol < 0.2 and e_inv < 0.2, (e_apply, e_chol, e_inv)
