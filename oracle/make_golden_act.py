"""Generate tests/golden/mlp_sigmoid_tanh_mse_sum.npz by running the REFERENCE (/root/reference, read-only):
GGN, Hessian (non-zero second-order terms through Sigmoid / Tanh), MC-GGN and empirical Fisher of an MLP with
smooth activations under MSELoss(reduction="sum"), two unequal mini-batches.  TEST INFRASTRUCTURE.
Run: python oracle/make_golden_act.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from curvlinops import EFLinearOperator, GGNLinearOperator, HessianLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import mlp_smooth  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.manual_seed(21)
model = mlp_smooth().eval()
data = [(torch.randn(6, 16), torch.randn(6, 6)), (torch.randn(4, 16), torch.randn(4, 6))]
loss = nn.MSELoss(reduction="sum")
params = dict(model.named_parameters())
P = sum(p.numel() for p in params.values())
V = torch.rand(P, 3, generator=torch.Generator().manual_seed(1))
extra = {"V": V}
extra["ggn"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V
extra["hessian"] = HessianLinearOperator(model, loss, params, data, check_deterministic=False) @ V
extra["ef"] = EFLinearOperator(model, loss, params, data, check_deterministic=False) @ V
for M in (1, 3):
    extra[f"ggn_mc{M}"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False,
                                            mc_samples=M, seed=1234) @ V
save("mlp_sigmoid_tanh_mse_sum", model, data, extra)
