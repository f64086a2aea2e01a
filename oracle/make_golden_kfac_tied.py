"""Generate tests/golden/kfac_tied.npz by running the REFERENCE (/root/reference, read-only): KFAC (expand) of a network
whose Conv2d and Linear are used twice each, through the reference's make_fx backend -- the backend that supports weight
tying by concatenating the usages along the weight-sharing axis (computers/io_collector/groups.py:123-168); the two
usages of the convolution see different numbers of positions (8 x 8 and 4 x 4).  No EKFAC entries (the engine's
eigenvalue correction does not take tied weights).  TEST INFRASTRUCTURE.  Run: python oracle/make_golden_kfac_tied.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn

from oracle.make_golden import kfac_cases  # noqa: E402
from oracle.models import TiedNet  # noqa: E402

torch.set_default_dtype(torch.float64)
torch.manual_seed(61)
kfac_cases("kfac_tied", TiedNet().eval(),
           [(torch.rand(4, 3, 8, 8), torch.randint(0, 5, (4,))), (torch.rand(3, 3, 8, 8), torch.randint(0, 5, (3,)))],
           nn.CrossEntropyLoss(), backend="make_fx", ekfac=False)
