"""Full-size checks on BASELINE.json configs[1] (ResNet-18, 128 x 3 x 224 x 224, GGN @ 8 vectors, fp32):

* parity of the WHOLE configuration (B = 128, K = 8) against the oracle's float64 restatement evaluated on the GPU
  (the oracle loops over sample chunks -- the GGN is a sum over samples -- and over columns), reported as max-norm
  error, per-parameter max-norm error and the fraction of entries violating the elementwise
  ``isclose(rtol=1e-4, atol=1e-5 max|ref|)`` of north_star's fp32 tolerance; the same statistics are printed for the
  unmodified reference (oracle/_ref) run in strict fp32 on the same GPU, i.e. how far the reference's own autograd
  path is from the float64 truth, and the engine is compared with that fp32 reference result directly;
* size-independent properties at the full batch: symmetry, linearity, bit-wise repeatability."""
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


def _stats(got, ref, sizes, names):
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    viol = (~torch.isclose(got, ref, rtol=1e-4, atol=1e-5 * scale)).double().mean().item()
    viol6 = (~torch.isclose(got, ref, rtol=1e-4, atol=1e-6 * scale)).double().mean().item()
    per, o = [], 0
    for n, sz in zip(names, sizes):
        r = ref[o:o + sz].abs().max().item()
        per.append((err[o:o + sz].max().item() / max(r, 1e-300), n))
        o += sz
    return err.max().item() / scale, viol, viol6, max(per)


def test_resnet18_c2_full_configuration_matches_float64_oracle():
    """B = 128, K = 8: the configuration the headline number is measured on."""
    import torchvision

    Bn, K = 128, 8
    model, X, y = _problem(Bn)
    dev = X.device
    params = dict(model.named_parameters())
    sizes = [p.numel() for p in params.values()]
    P = sum(sizes)
    V = torch.rand(P, K, device=dev)
    loss = torch.nn.CrossEntropyLoss()
    G = GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False, num_data=Bn)
    got = (G @ V).double()
    del G
    torch.cuda.empty_cache()
    m64 = torchvision.models.resnet18().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    Vl = [v.reshape(*p.shape, K).double() for v, p in zip(V.split(sizes), p64.values())]
    chunks = [(X[i:i + 32].double(), y[i:i + 32]) for i in range(0, Bn, 32)]
    ref = torch.cat([r.reshape(-1, K) for r in orc.ggn_matmat(m64, loss, p64, chunks, Vl, n_data=Bn)])
    del m64, p64, Vl, chunks
    torch.cuda.empty_cache()
    e, viol, viol6, worst = _stats(got, ref, sizes, list(params))
    print(f"engine    vs fp64 oracle: max|err|/max|ref| = {e:.3e}, violations of isclose(rtol 1e-4, atol 1e-5 max) = "
          f"{viol:.3e} (atol 1e-6 max: {viol6:.3e}), worst parameter {worst[1]}: {worst[0]:.3e}")
    assert e < 1e-4, e
    assert viol == 0.0, viol
    assert worst[0] < 5e-4, worst
    # the reference's own fp32 autograd path on this GPU (strict fp32: no TF32), same inputs
    try:
        from oracle.build_ref import import_reference

        refpkg = import_reference()
    except ImportError as exc:
        pytest.skip(f"fp32 reference leg skipped: {exc}")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        Gr = refpkg.GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False, num_data=Bn)
        r32 = (Gr @ V).double()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    e2, v2, v26, w2 = _stats(r32, ref, sizes, list(params))
    print(f"reference (fp32, torch CUDA) vs fp64 oracle: max|err|/max|ref| = {e2:.3e}, violations = {v2:.3e} "
          f"(atol 1e-6 max: {v26:.3e}), worst parameter {w2[1]}: {w2[0]:.3e}")
    scale = r32.abs().max().item()
    assert torch.allclose(got, r32, rtol=1e-4, atol=1e-5 * scale), (got - r32).abs().max().item() / scale


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


@pytest.mark.parametrize("batch,K", [(16, 2), (8, 8)])
def test_resnet18_hessian_matches_float64_oracle(batch, K):
    """The Hessian R-op on the half-split tensor-core kernels (forward, both dgrad terms, both wgrad terms) at
    ResNet-18 layer shapes against the oracle's float64 forward-over-reverse restatement on the GPU; K = 8 fills all
    K + 1 cotangent slots.  Also: the same product on the 3xTF32 kernels (mode bit 0x8000) agrees."""
    import torchvision

    from curvlinops_b200 import HessianLinearOperator, _capi as capi

    model, X, y = _problem(batch)
    dev = X.device
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, K, device=dev)
    loss = torch.nn.CrossEntropyLoss()
    m64 = torchvision.models.resnet18().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    ref = []
    for k0 in range(0, K, 2):  # two columns at a time bounds the double-backward graph memory
        Vl = [v[..., k0:k0 + 2].reshape(*p.shape, -1).double()
              for v, p in zip(V.split([p.numel() for p in p64.values()]), p64.values())]
        ref.append(torch.cat([r.reshape(r.shape[0] if r.ndim == 1 else -1, r.shape[-1]).reshape(-1, r.shape[-1])
                              for r in orc.hessian_matmat(m64, loss, p64, [(X.double(), y)], Vl)]))
    ref = torch.cat(ref, dim=1)
    H = HessianLinearOperator(model, loss, params, [(X, y)], check_deterministic=False)
    got = (H @ V).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    old = capi.lib().curv_set_tensor_core_mode(1 | 0x8000)
    try:
        got3 = (H @ V).double()
    finally:
        capi.lib().curv_set_tensor_core_mode(old)
    err3 = (got3 - ref).abs().max().item() / ref.abs().max().item()
    print(f"ResNet-18 Hessian, B={batch}, K={K}: max|err|/max|ref| = {err:.3e} (half-split), {err3:.3e} (3xTF32)")
    assert err < 1e-4, err
    assert err3 < 5e-4, err3  # the round-1 kernels (A/B switch only): TF32 splits of badly scaled cotangents
