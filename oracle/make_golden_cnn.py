"""Generate tests/golden/cnn_bias_ce_mean.npz by running the REFERENCE (/root/reference, read-only): GGN, Hessian,
MC-GGN and empirical Fisher of a plain CNN with conv biases, a stride-2 un-padded conv and flatten -> Linear over a
3x3 feature map (the engine lowers that Linear to a 'valid' convolution), two unequal mini-batches.
TEST INFRASTRUCTURE.  Run: python oracle/make_golden_cnn.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from curvlinops import EFLinearOperator, GGNLinearOperator, HessianLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import ConvNetBias  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.manual_seed(31)
model = ConvNetBias().eval()
data = [(torch.rand(4, 3, 8, 8), torch.randint(0, 5, (4,))), (torch.rand(3, 3, 8, 8), torch.randint(0, 5, (3,)))]
loss = nn.CrossEntropyLoss()
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
save("cnn_bias_ce_mean", model, data, extra)
