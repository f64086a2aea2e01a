"""The two-sided Kronecker apply on the tcgen05 kernels (``curv_kron_apply_tc``, reference ``kronecker.py:141-153``)
against float64 on random blocks at the shapes of real layers, including non-symmetric factors, the adjoint product,
K > 8 (chunked), repeated products (cached factor operands), and a factor whose rows sum to zero (a gradient covariance under the
softmax: the product then cancels to far below its terms and every fp32 evaluation loses digits -- the strict-fp32
cuBLAS product is printed as the yardstick).  Tolerance: 1e-4 of the largest entry (fp32), 2x the yardstick under
cancellation."""
import pytest
import torch

from curvlinops_b200 import KroneckerProductLinearOperator

pytestmark = pytest.mark.gpu


def _case(d_out, d_in, K, sym, centered, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    G = torch.randn(d_out, d_out, device="cuda", generator=g)
    A = torch.randn(d_in, d_in, device="cuda", generator=g)
    if sym:
        G, A = G @ G.T / d_out, A @ A.T / d_in
    if centered:  # rows (and, if symmetric, columns) sum to zero
        G = G - G.mean(1, keepdim=True)
        if sym:
            G = G - G.mean(0, keepdim=True)
            G = (G + G.T) / 2
    X = torch.rand(d_out * d_in, K, device="cuda", generator=g)
    ref = torch.einsum("abz,Aa,Bb->ABz", X.double().reshape(d_out, d_in, K), G.double(), A.double()).reshape(-1, K)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        lib = torch.einsum("abz,Aa,Bb->ABz", X.reshape(d_out, d_in, K), G, A).reshape(-1, K)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    got = KroneckerProductLinearOperator(G.to(dtype), A.to(dtype)) @ X.to(dtype)
    scale = ref.abs().max()
    return ((got.double() - ref).abs().max() / scale).item(), ((lib.double() - ref).abs().max() / scale).item()


@pytest.mark.parametrize("d_out,d_in,K", [(64, 148, 1), (64, 576, 3), (128, 1152, 8), (512, 4608, 1), (1000, 513, 2),
                                          (256, 2304, 11)])
@pytest.mark.parametrize("sym", [True, False])
def test_kron_apply_matches_float64(d_out, d_in, K, sym):
    e, lib = _case(d_out, d_in, K, sym, centered=False)
    print(f"[{d_out} x {d_in}, K={K}, sym={sym}] engine {e:.2e}, strict-fp32 cuBLAS einsum {lib:.2e}")
    assert e < 1e-4, e


@pytest.mark.parametrize("d_out,d_in,K", [(1000, 513, 1), (512, 1152, 4)])
def test_kron_apply_under_cancellation(d_out, d_in, K):
    e, lib = _case(d_out, d_in, K, sym=True, centered=True)
    print(f"[{d_out} x {d_in}, K={K}, rows of G sum to zero] engine {e:.2e}, strict-fp32 cuBLAS einsum {lib:.2e}")
    assert e < max(1e-4, 2 * lib), (e, lib)


def _kfac_like(which, seed=1):
    """fc block of a softmax classifier under MC Fisher: G = sum of 16 outer products of (p - onehot) (rank 16, rows
    sum to zero, entries spanning 1e5), A = second moment of [features, 1] with a few dominant features."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    d_out, d_in = 1000, 513
    p = torch.softmax(torch.randn(16, d_out, device="cuda", generator=g) * 2, 1)
    go = (p - torch.nn.functional.one_hot(p.multinomial(1, generator=g)[:, 0], d_out)) / 16
    G = go.T @ go * 16
    f = torch.relu(torch.randn(16, d_in - 1, device="cuda", generator=g)) * torch.logspace(-2, 0.7, d_in - 1, device="cuda")
    a = torch.cat([f, torch.ones(16, 1, device="cuda")], 1)
    A = a.T @ a / 16
    G, A = (G + G.T) / 2, (A + A.T) / 2
    if which == "G":
        A = torch.randn(d_in, d_in, device="cuda", generator=g); A = A @ A.T / d_in
    if which == "A":
        G = torch.randn(d_out, d_out, device="cuda", generator=g); G = G @ G.T / d_out
    X = torch.rand(d_out * d_in, 1, device="cuda", generator=g)
    ref = (G.double() @ X.double().reshape(d_out, d_in) @ A.double().T).reshape(-1, 1)
    got = KroneckerProductLinearOperator(G, A) @ X
    return ((got.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("which", ["G", "A", "both"])
def test_kron_apply_on_kfac_like_factors(which):
    e = _kfac_like(which)
    print(f"[fc block, KFAC-like {which}] engine {e:.2e}")
    assert e < 1e-4, e


def test_kron_adjoint_and_cached_factor_operands():
    g = torch.Generator(device="cuda").manual_seed(3)
    G, A = torch.randn(96, 96, device="cuda", generator=g), torch.randn(200, 200, device="cuda", generator=g)
    X = torch.rand(96 * 200, 2, device="cuda", generator=g)
    op = KroneckerProductLinearOperator(G, A)
    dense = torch.kron(G.double(), A.double())
    ref = dense @ X.double()
    atol = 1e-5 * ref.abs().max().item()
    for _ in range(2):  # second product: factor operands come from the cache
        torch.testing.assert_close((op @ X).double(), ref, rtol=1e-4, atol=atol)
        torch.testing.assert_close((X.T @ op).double(), X.double().T @ dense, rtol=1e-4, atol=atol)
    assert op._tc_cache[False][0][:2] == tuple((id(f), f._version) for f in op)
    op[0] = G * 2.0  # a replaced factor invalidates the cached operands
    torch.testing.assert_close((op @ X).double(), 2.0 * ref, rtol=1e-4, atol=2 * atol)
