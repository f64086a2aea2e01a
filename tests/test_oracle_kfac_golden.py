"""Pins the oracle's KFAC conventions (factors, Kronecker apply, damped inverse) to reference fixtures."""
import pytest
import torch

from oracle import curvature_oracle as orc
from tests.golden_utils import load_case

CASES = ["kfac_mlp", "kfac_cnn"]


def layer_names(model, param_names):
    return list(dict.fromkeys(n.rsplit(".", 1)[0] for n in param_names))


def kfac_dense_apply(model, A, G, names, v, joint, inverse_damping=None):
    """(P K P^T) v (or with K^-1) from per-layer factors, reference canonical layout (kfac_utils.py:280-385)."""
    mods = dict(model.named_modules())
    out, o = [], 0
    for n in names:
        m = mods[n]
        w, b = m.weight, m.bias
        K = v.shape[1]
        Wv = v[o:o + w.numel()].reshape(w.shape[0], -1, K); o += w.numel()
        bv = None
        if b is not None:
            bv = v[o:o + b.numel()]; o += b.numel()
        Gm, Am = G[n], A[n]
        if joint and b is not None:
            comb = torch.cat([Wv, bv.unsqueeze(1)], 1)
            if inverse_damping is not None:
                Gm, Am = orc.damped_inverse(Gm, inverse_damping), orc.damped_inverse(Am, inverse_damping)
            res = orc.kron_apply(Gm, Am, comb)
            out += [res[:, :-1].reshape(-1, K), res[:, -1]]
        else:
            Aw = Am[:-1, :-1] if (b is not None and Am.shape[0] == Wv.shape[1] + 1) else Am
            if inverse_damping is not None:
                Gi, Ai = orc.damped_inverse(Gm, inverse_damping), orc.damped_inverse(Aw, inverse_damping)
            else:
                Gi, Ai = Gm, Aw
            out.append(orc.kron_apply(Gi, Ai, Wv).reshape(-1, K))
            if b is not None:
                out.append(Gi @ bv)
    return torch.cat(out)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("ft", ["type2", "mc"])
def test_factors_and_apply(name, ft):
    model, loss, data, fx = load_case(name)
    pnames = [str(s) for s in fx["param_names"]]
    names = layer_names(model, pnames)
    A, G = orc.kfac_factors(model, loss, names, data, fisher_type=ft, mc_samples=2, seed=77, joint_bias=True)
    if ft == "type2":
        for bi, n in enumerate(names):
            torch.testing.assert_close(G[n], fx[f"factor_type2_joint_{bi}_0"], rtol=1e-9, atol=1e-12)
            torch.testing.assert_close(A[n], fx[f"factor_type2_joint_{bi}_1"], rtol=1e-9, atol=1e-12)
    v = fx["v"]
    got = kfac_dense_apply(model, A, G, names, v, joint=True)
    torch.testing.assert_close(got, fx[f"kfac_{ft}_joint"], rtol=1e-8, atol=1e-11)
    goti = kfac_dense_apply(model, A, G, names, v, joint=True, inverse_damping=float(fx["damping"]))
    torch.testing.assert_close(goti, fx[f"kfacinv_{ft}_joint"], rtol=1e-6, atol=1e-9)
    # separate weight / bias groups: A without the ones column
    gots = kfac_dense_apply(model, A, G, names, v, joint=False)
    torch.testing.assert_close(gots, fx[f"kfac_{ft}_sep"], rtol=1e-8, atol=1e-11)
