"""GPU parity tests: engine (through the C ABI) vs the reference-generated golden fixtures and vs
the CPU oracle.  Tolerance (BASELINE.json north_star): rtol = 1e-4 for fp32; atol = 1e-5 * max|ref|
absorbs entries that are ~0 relative to the matrix scale (the fixtures are float64)."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator, HessianLinearOperator
from oracle import curvature_oracle as orc
from tests.golden_utils import flat, load_case, split_like

pytestmark = pytest.mark.gpu
CASES = ["mlp_c1_ce_mean", "mlp_c1_ce_sum", "mlp_c1_mse_mean", "miniresnet_ce_mean"]
RTOL = 1e-4


def assert_parity(got, ref, params=None, rtol=RTOL):
    got, ref = got.detach().double().cpu(), ref.double().cpu()
    atol = 1e-5 * ref.abs().max().item()
    ok = torch.allclose(got, ref, rtol=rtol, atol=atol)
    if not ok and params is not None:  # per-parameter report to localise a failing layer
        o = 0
        for n, p in params.items():
            g, r = got[o:o + p.numel()], ref[o:o + p.numel()]
            print(f"{n:32s} max|ref|={r.abs().max():.3e} max|err|={(g - r).abs().max():.3e}")
            o += p.numel()
    assert ok, f"max abs err {(got - ref).abs().max():.3e} vs max|ref| {ref.abs().max():.3e}"


def _setup(name):
    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    return model, loss, data, fx, dict(model.named_parameters())


@pytest.mark.parametrize("name", CASES)
def test_ggn_matches_reference_golden(name):
    model, loss, data, fx, params = _setup(name)
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(G @ fx["V"].float().cuda(), fx["ggn"], params)


@pytest.mark.parametrize("name", CASES)
def test_hessian_matches_reference_golden(name):
    model, loss, data, fx, params = _setup(name)
    H = HessianLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(H @ fx["V"].float().cuda(), fx["hessian"], params)


@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "miniresnet_ce_mean"])
def test_formats_and_determinism(name):
    model, loss, data, fx, params = _setup(name)
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=True)  # runs the probes
    V = fx["V"].float().cuda()
    ref = G @ V
    assert torch.equal(ref, G @ V), "two matmats must be bit-identical"
    # vector, list and numpy formats, left multiplication (self-adjoint)
    torch.testing.assert_close(G @ V[:, 0], ref[:, 0], rtol=1e-5, atol=1e-8)
    lst = G @ split_like(V, params)
    torch.testing.assert_close(flat(lst), ref, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close((V.T @ G).T, ref, rtol=1e-5, atol=1e-8)
    got_np = G.to_scipy() @ V.double().cpu().numpy()
    torch.testing.assert_close(torch.from_numpy(got_np), ref.double().cpu(), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "mlp_c1_mse_mean", "miniresnet_ce_mean"])
@pytest.mark.parametrize("M", [1, 3])
def test_mc_ggn_with_reference_samples(name, M):
    """MC-GGN with the would-be gradients the reference drew (same seed => same stream on CPU);
    the engine is handed those samples, so the comparison is exact rather than in expectation."""
    model, loss, data, fx, params = _setup(name)
    cpu_model = load_case(name)[0]
    gs = []
    with torch.random.fork_rng():
        torch.manual_seed(1234)
        for X, _ in data:
            gs.append(orc.mc_grad_outputs(loss, cpu_model(X.double().cpu()).detach(), M).float())
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=M, seed=1234)
    G._mc_grad_override = gs
    assert_parity(G @ fx["V"].float().cuda(), fx[f"ggn_mc{M}"], params)


def test_mc_ggn_own_sampler_is_seeded_and_unbiased():
    model, loss, data, fx, params = _setup("mlp_c1_ce_mean")
    V = fx["V"].float().cuda()[:, :1]
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=2, seed=5)
    a, b = G @ V, G @ V
    assert torch.equal(a, b)
    state = torch.cuda.get_rng_state()
    G @ V
    assert torch.equal(state, torch.cuda.get_rng_state()), "global RNG must not be advanced"
    acc = torch.zeros_like(a)
    reps = 300
    for s in range(reps):
        G._seed = s
        acc += G @ V
    exact = fx["ggn"][:, :1].float().cuda()
    rel = (acc / reps - exact).norm() / exact.norm()
    assert rel < 0.1, rel


def test_param_subset_and_order():
    model, loss, data, fx, _ = _setup("miniresnet_ce_mean")
    names = ["fc.bias", "layer2.0.conv1.weight", "bn1.weight", "layer1.0.bn2.bias", "conv1.weight"]
    params = {n: dict(model.named_parameters())[n] for n in names}
    cpu_model = load_case("miniresnet_ce_mean")[0]
    p64 = {n: dict(cpu_model.named_parameters())[n] for n in names}
    data64 = [(X.double().cpu(), y.cpu()) for X, y in data]
    torch.manual_seed(0)
    V = torch.rand(sum(p.numel() for p in params.values()), 2, dtype=torch.float64)
    ref = flat(orc.ggn_matmat(cpu_model, loss, p64, data64, split_like(V, p64)))
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(G @ V.float().cuda(), ref, params)
    refh = flat(orc.hessian_matmat(cpu_model, loss, p64, data64, split_like(V, p64)))
    H = HessianLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(H @ V.float().cuda(), refh, params)


def test_wide_matrix_is_chunked():
    """K larger than the engine's column chunk (multiplying onto many columns, test/utils.py:153)."""
    model, loss, data, fx, params = _setup("mlp_c1_ce_mean")
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    torch.manual_seed(1)
    V = torch.rand(G.shape[1], 19, device="cuda")
    wide = G @ V
    for k in (0, 7, 8, 18):
        torch.testing.assert_close(wide[:, k], G @ V[:, k], rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# tcgen05 (3xTF32) contraction kernels: forced onto every layer, must agree with the fp32 SIMT kernels
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def force_tensor_cores():
    from curvlinops_b200 import _capi as capi

    old = capi.lib().curv_set_tensor_core_mode(2)
    yield
    capi.lib().curv_set_tensor_core_mode(old)


@pytest.mark.parametrize("name", CASES)
def test_tcgen05_path_matches_reference_golden(name, force_tensor_cores):
    model, loss, data, fx, params = _setup(name)
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(G @ fx["V"].float().cuda(), fx["ggn"], params)
    H = HessianLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(H @ fx["V"].float().cuda(), fx["hessian"], params)


def test_tcgen05_vs_simt_on_wide_convnet():
    """Layers large enough for full 128x128 tiles, several K chunks, stride-2 dgrad, 2 segments."""
    from curvlinops_b200 import _capi as capi
    from torch import nn

    torch.manual_seed(0)
    model = nn.Sequential(
        nn.Conv2d(3, 64, 3, 1, 1), nn.ReLU(), nn.Conv2d(64, 160, 3, 2, 1), nn.ReLU(),
        nn.Conv2d(160, 96, 1, 1, 0, bias=False), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
        nn.Linear(96, 10)).cuda().eval()
    params = dict(model.named_parameters())
    X, y = torch.rand(6, 3, 24, 24, device="cuda"), torch.randint(0, 10, (6,), device="cuda")
    G = GGNLinearOperator(model, nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False)
    V = torch.rand(G.shape[1], 3, device="cuda")
    old = capi.lib().curv_set_tensor_core_mode(0)
    try:
        ref = G @ V
        capi.lib().curv_set_tensor_core_mode(2)
        got = G @ V
        assert torch.equal(got, G @ V), "tcgen05 path must be run-to-run deterministic"
    finally:
        capi.lib().curv_set_tensor_core_mode(old)
    err = (got - ref).abs().max() / ref.abs().max()
    assert err < 2e-5, err


def test_jacobian_operators_match_oracle():
    """J and J^T (reference curvlinops/jacobian.py) against the oracle's double-vjp / vjp restatement."""
    from curvlinops_b200 import JacobianLinearOperator, TransposedJacobianLinearOperator

    model, loss, data, fx, params = _setup("miniresnet_ce_mean")
    cpu_model, _, cpu_data, _ = load_case("miniresnet_ce_mean")
    p64 = dict(cpu_model.named_parameters())
    J = JacobianLinearOperator(model, params, data, check_deterministic=False)
    N = sum(X.shape[0] for X, _ in data)
    assert J.shape == (N * 10, sum(p.numel() for p in params.values()))
    V = fx["V"][:, :2]
    got = J @ V.float().cuda()
    f_fn = orc._as_callable(cpu_model)
    ref = torch.cat([torch.stack([orc.jacobian_vector_product(f_fn, p64, X, [v[..., k] for v in split_like(V, p64)])[1]
                                  for k in range(2)], dim=-1) for X, _ in cpu_data]).reshape(-1, 2)
    assert_parity(got, ref)
    JT = J.adjoint()
    assert isinstance(JT, TransposedJacobianLinearOperator)
    torch.manual_seed(0)
    W = torch.rand(N * 10, 2, dtype=torch.float64)
    refT = torch.zeros(J.shape[1], 2, dtype=torch.float64)
    pos = 0
    for X, _ in cpu_data:
        for k in range(2):
            w = W[pos * 10:(pos + X.shape[0]) * 10, k].reshape(X.shape[0], 10)
            refT[:, k] += torch.cat([g.flatten() for g in orc.transposed_jacobian_vector_product(f_fn, p64, X, w)])
        pos += X.shape[0]
    assert_parity(JT @ W.float().cuda(), refT, params)


@pytest.mark.parametrize("name", CASES)
def test_empirical_fisher_matches_reference_golden(name):
    """EFLinearOperator vs the reference's own EFLinearOperator output (tests/golden/ef.npz)."""
    import os

    import numpy as np

    from curvlinops_b200 import EFLinearOperator
    from tests.golden_utils import GOLDEN

    model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "ef.npz"))[name])
    E = EFLinearOperator(model, loss, params, data, check_deterministic=False)
    got = (E @ fx["V"].float().cuda()).double().cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-4, err  # BASELINE tolerance: rtol 1e-4 (fp32 engine vs float64 reference)


@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "miniresnet_ce_mean"])
@pytest.mark.parametrize("op", ["ggn", "hessian"])
def test_matmat_pinned_streams_and_matches(name, op):
    """Host-resident V / result with pipelined upload and download (curv_matmat_batch_sync, lazy per-node
    preparation): same kernels in the same order per layer => bit-identical to the resident product."""
    model, loss, data, fx, params = _setup(name)
    cls = GGNLinearOperator if op == "ggn" else HessianLinearOperator
    A = cls(model, loss, params, data, check_deterministic=False)
    V = fx["V"].float()
    ref = (A @ V.cuda()).cpu()
    Vh = V.clone().pin_memory()
    # tiny buckets: every parameter its own bucket => exercises the per-parameter events
    got = A.matmat_pinned(Vh, bucket_bytes=1)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    got2 = A.matmat_pinned(Vh)  # default bucket size (single bucket here)
    torch.cuda.synchronize()
    assert torch.equal(got2, ref)
    assert_parity(got, fx["ggn" if op == "ggn" else "hessian"], params)


def test_on_device_lanczos_matches_scipy_eigsh():
    """``lanczos_eigsh`` (all vectors on the device) vs ``scipy.sparse.linalg.eigsh`` through ``to_scipy()`` on the
    same GGN operator (reference usage: docs/examples/basic_usage/example_eigenvalues.py:65-77)."""
    from scipy.sparse.linalg import eigsh

    from curvlinops_b200 import lanczos_eigsh

    model, loss, data, fx, params = _setup("mlp_c1_ce_mean")
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    ev, vec, nprod = lanczos_eigsh(G, k=3, which="LA", tol=1e-5, return_info=True)
    ref = eigsh(G.to_scipy(), k=3, which="LA", tol=1e-6, return_eigenvectors=False)
    torch.testing.assert_close(ev.double().cpu(), torch.from_numpy(ref).double().sort().values, rtol=2e-4, atol=1e-6)
    resid = ((G @ vec) - vec * ev).norm(dim=0) / ev.abs().max()
    assert bool((resid < 1e-3).all()), resid
    assert vec.device.type == "cuda" and nprod < 200


def test_parameter_space_glue_variants_are_bit_identical():
    """Building the tangent-weight images straight from the columns of V computes exactly what the two-step path
    (packed fp32 copy, then image; mode bit 0x2000) computes: same scales, same roundings."""
    from curvlinops_b200 import _capi as capi

    model, loss, data, fx, params = _setup("miniresnet_ce_mean")
    V = fx["V"].float().cuda()
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    old = capi.lib().curv_set_tensor_core_mode(2)
    try:
        new = G @ V
        capi.lib().curv_set_tensor_core_mode(2 | 0x2000)
        ref = G @ V
    finally:
        capi.lib().curv_set_tensor_core_mode(old)
    assert torch.equal(new, ref)
    assert_parity(new, fx["ggn"], params)


@pytest.mark.parametrize("name", ["mlp_c1_ce_mean", "mlp_c1_mse_mean", "mlp_bce_mean"])
def test_mc_draws_are_keyed_on_the_global_sample_index(name):
    """A rank that holds samples lo..hi of a mini-batch draws exactly the would-be gradients a single process draws
    for those samples (Monte-Carlo products must not depend on the number of ranks)."""
    model, loss, data, fx, params = _setup(name)
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=3)
    X = data[0][0]
    B = X.shape[0]
    torch.manual_seed(5)
    full = G._engine.mc_grad_outputs(X, 3)
    for lo, hi in ((0, 2), (2, B), (1, B - 1)):
        torch.manual_seed(5)
        part = G._engine.mc_grad_outputs(X[lo:hi], 3, shard=(lo, hi, B))
        assert torch.equal(part, full[lo:hi])


def test_mc_sampler_draws_from_the_softmax():
    """The engine's own sampler (inverse CDF on uniforms keyed on the global sample index): the empirical label
    frequencies of 40 000 draws per datum match the softmax within 5 standard errors."""
    model, loss, data, fx, params = _setup("mlp_c1_ce_mean")
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=1)
    X = data[0][0][:4]
    M = 40000
    p = torch.softmax(G._engine.predict(X), 1)
    torch.manual_seed(11)
    g = G._engine.mc_grad_outputs(X, M) * (M ** 0.5)       # [B, M, C] = p - onehot(yhat)
    freq = (p.unsqueeze(1) - g).mean(1)                    # mean of the one-hot draws
    se = (p * (1 - p) / M).sqrt()
    assert ((freq - p).abs() <= 5 * se + 1e-6).all()


class _TiedNet(torch.nn.Module):
    """One Linear and one Conv2d used twice each (weight tying): the engine reads the same columns of V for both
    usages and accumulates both gradients into the same rows of the result."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(4, 4, 3, padding=1)
        self.stem = torch.nn.Conv2d(3, 4, 3, padding=1)
        self.fc = torch.nn.Linear(4, 4)
        self.head = torch.nn.Linear(4, 5)

    def forward(self, x):
        x = torch.tanh(self.stem(x))
        x = torch.tanh(self.conv(x))
        x = torch.tanh(self.conv(x))
        x = torch.nn.functional.adaptive_avg_pool2d(x, 1).flatten(1)
        x = torch.tanh(self.fc(x))
        x = torch.tanh(self.fc(x))
        return self.head(x)


@pytest.mark.parametrize("op", ["ggn", "hessian"])
def test_weight_tying_matches_float64_oracle(op):
    torch.manual_seed(3)
    model = _TiedNet().cuda().eval()
    data = [(torch.rand(6, 3, 8, 8, device="cuda"), torch.randint(0, 5, (6,), device="cuda")),
            (torch.rand(3, 3, 8, 8, device="cuda"), torch.randint(0, 5, (3,), device="cuda"))]
    loss = torch.nn.CrossEntropyLoss()
    params = dict(model.named_parameters())
    m64 = _TiedNet().double().eval()
    m64.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    data64 = [(X.double().cpu(), y.cpu()) for X, y in data]
    V = torch.rand(sum(p.numel() for p in params.values()), 3, dtype=torch.float64)
    if op == "ggn":
        ref = flat(orc.ggn_matmat(m64, loss, p64, data64, split_like(V, p64)))
        A = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    else:
        ref = flat(orc.hessian_matmat(m64, loss, p64, data64, split_like(V, p64)))
        A = HessianLinearOperator(model, loss, params, data, check_deterministic=False)
    assert_parity(A @ V.float().cuda(), ref, params)


def test_mc_would_be_gradients_are_cached_without_changing_results():
    """The MC draws of a mini-batch are kept across products (same seed -> same draws): results are bit-identical to an
    operator without the cache, also when only part of the batches hit (the stream stays aligned), and a parameter
    update invalidates the entries."""
    model, loss, data, fx, params = _setup("mlp_c1_ce_mean")
    V = fx["V"].float().cuda()[:, :2]
    mk = lambda: GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=3, seed=5)
    G = mk()
    a = G @ V                        # fills the cache (two mini-batches)
    assert len(G._mc_cache) == 2
    assert torch.equal(G @ V, a) and torch.equal(mk() @ V, a)
    G._mc_cache.pop(0)               # batch 0 misses, batch 1 hits
    assert torch.equal(G @ V, a)
    with torch.no_grad():
        next(iter(params.values())).add_(0.01)
    b = G @ V
    assert not torch.equal(a, b) and torch.equal(b, mk() @ V)
