"""Generate tests/golden/ggn_diag_*.npz by running the REFERENCE (/root/reference, read-only): the exact GGN diagonal
(``GGNDiagonalLinearOperator(mc_samples=0)``, curvlinops/ggn_diagonal.py) of an MLP (CE-mean, MSE-sum) and of a CNN
with conv biases, two unequal mini-batches each.  Also stored: the diagonal of the dense GGN obtained from
``GGNLinearOperator @ I`` (the two must agree - checked here).  TEST INFRASTRUCTURE.  Run: python oracle/make_golden_diag.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from curvlinops import GGNDiagonalLinearOperator, GGNLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import ConvNetBias, mlp_c1  # noqa: E402

torch.set_default_dtype(torch.float64)


class _ConvNetAnyRank(ConvNetBias):
    """ConvNetBias (same parameters / state dict) whose flatten also takes un-batched inputs: the reference's
    diagonal computer vmaps over data points."""

    def forward(self, x):
        x = torch.relu(self.c2(torch.relu(self.c1(x))))
        return self.fc(torch.flatten(x, -3))


def case(name, model, data, loss, seed):
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    D = GGNDiagonalLinearOperator(model, loss, params, data, check_deterministic=False)
    diag = D @ torch.ones(P)
    dense = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ torch.eye(P)
    assert torch.allclose(diag, dense.diagonal(), rtol=1e-9, atol=1e-12), (diag - dense.diagonal()).abs().max()
    save(name, model, data, {"diag": diag})
    print(name, P, float(diag.abs().max()))


torch.manual_seed(41)
case("ggn_diag_mlp_ce_mean", mlp_c1(classes=4, width=12).eval(),
     [(torch.rand(5, 12), torch.randint(0, 4, (5,))), (torch.rand(3, 12), torch.randint(0, 4, (3,)))],
     nn.CrossEntropyLoss(), 41)
torch.manual_seed(42)
case("ggn_diag_mlp_mse_sum", mlp_c1(classes=4, width=12).eval(),
     [(torch.rand(5, 12), torch.rand(5, 4)), (torch.rand(3, 12), torch.rand(3, 4))], nn.MSELoss(reduction="sum"), 42)
torch.manual_seed(43)
case("ggn_diag_cnn_ce_mean", _ConvNetAnyRank().eval(),
     [(torch.rand(4, 3, 8, 8), torch.randint(0, 5, (4,))), (torch.rand(3, 3, 8, 8), torch.randint(0, 5, (3,)))],
     nn.CrossEntropyLoss(), 43)
