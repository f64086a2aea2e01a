"""Generate tests/golden/mlp_bce_{mean,sum}.npz by running the REFERENCE (/root/reference, read-only):
GGN, Hessian, MC-GGN (1 and 3 samples) and empirical Fisher of a small MLP under BCEWithLogitsLoss with 0/1
targets, two unequal mini-batches (the reference's own BCE cases: test/cases.py, binary_classification_targets).
TEST INFRASTRUCTURE.  Run: python oracle/make_golden_bce.py  (separate from make_golden.py so that the committed
fixtures of the other cases are not rewritten)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from curvlinops import EFLinearOperator, GGNLinearOperator, HessianLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import mlp_c1  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.manual_seed(11)
model = mlp_c1(classes=6, width=16).eval()
data = [(torch.randn(7, 16), torch.randint(0, 2, (7, 6)).double()),
        (torch.randn(5, 16), torch.randint(0, 2, (5, 6)).double())]
for red in ("mean", "sum"):
    loss = nn.BCEWithLogitsLoss(reduction=red)
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
    save(f"mlp_bce_{red}", model, data, extra)
