"""Generate tests/golden/kfac_tokens.npz by running the REFERENCE (/root/reference, read-only): KFAC / EKFAC (expand
setting) of a token model -- Linear layers applied to [B, T, D] sequences share their weights over the T positions
(reference test/test_kfac.py weight-sharing cases; kfac_math.py:47-203), LayerNorm / GELU / residual in between, mean
over tokens before the head.  Same contents as the other kfac_* fixtures.  TEST INFRASTRUCTURE.
Run: python oracle/make_golden_kfac_tokens.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from oracle.make_golden import kfac_cases  # noqa: E402
from oracle.models import TokenMLP  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.manual_seed(51)
kfac_cases("kfac_tokens", TokenMLP().eval(),
           [(torch.rand(4, 7, 12), torch.randint(0, 5, (4,))), (torch.rand(3, 7, 12), torch.randint(0, 5, (3,)))],
           nn.CrossEntropyLoss())

# KFAC-reduce on the same model / data: inputs averaged and output gradients summed over the T positions before the
# outer products (kfac_math.py:47-170); Linear layers only, so the reference needs no einconv
torch.manual_seed(51)
kfac_cases("kfac_tokens_reduce", TokenMLP().eval(),
           [(torch.rand(4, 7, 12), torch.randint(0, 5, (4,))), (torch.rand(3, 7, 12), torch.randint(0, 5, (3,)))],
           nn.CrossEntropyLoss(), ekfac=False, kfac_approx="reduce")
