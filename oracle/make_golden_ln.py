"""Generate tests/golden/{mlp_ln_gelu_ce_mean, token_mlp_ce_mean, mlp_gelu_mse_sum}.npz by running the REFERENCE
(/root/reference, read-only): LayerNorm / GELU networks (2-d inputs and token sequences [B, T, D]), two unequal
mini-batches.  GGN, MC-GGN (reference RNG stream) and empirical Fisher for all; the Hessian for the GELU MLP (the
engine's R-op covers GELU, not LayerNorm).  TEST INFRASTRUCTURE.  Run: python oracle/make_golden_ln.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from curvlinops import EFLinearOperator, GGNLinearOperator, HessianLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import TokenMLP, mlp_gelu, mlp_ln_gelu  # noqa: E402

torch.set_default_dtype(torch.float64)


def run(name, model, data, loss, hessian):
    model = model.eval()
    with torch.no_grad():  # non-trivial LayerNorm parameters
        for m in model.modules():
            if isinstance(m, nn.LayerNorm):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, 3, generator=torch.Generator().manual_seed(1))
    extra = {"V": V}
    extra["ggn"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V
    extra["ef"] = EFLinearOperator(model, loss, params, data, check_deterministic=False) @ V
    if hessian:
        extra["hessian"] = HessianLinearOperator(model, loss, params, data, check_deterministic=False) @ V
    for M in (1, 3):
        extra[f"ggn_mc{M}"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=M,
                                                seed=1234) @ V
    save(name, model, data, extra)


torch.manual_seed(31)
run("mlp_ln_gelu_ce_mean", mlp_ln_gelu(), [(torch.randn(7, 16), torch.randint(0, 6, (7,))),
                                           (torch.randn(4, 16), torch.randint(0, 6, (4,)))], nn.CrossEntropyLoss(), False)
torch.manual_seed(32)
run("token_mlp_ce_mean", TokenMLP(), [(torch.randn(5, 9, 12), torch.randint(0, 5, (5,))),
                                      (torch.randn(3, 9, 12), torch.randint(0, 5, (3,)))], nn.CrossEntropyLoss(), False)
torch.manual_seed(33)
run("mlp_gelu_mse_sum", mlp_gelu(), [(torch.randn(6, 16), torch.randn(6, 6)), (torch.randn(4, 16), torch.randn(4, 6))],
    nn.MSELoss(reduction="sum"), True)
