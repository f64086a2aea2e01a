"""Generate tests/golden/ef.npz by running the REFERENCE's EFLinearOperator (/root/reference, read-only) on the
parameters / data / V of the existing curvature fixtures.  TEST INFRASTRUCTURE.  Run: python oracle/make_golden_ef.py
(kept separate from make_golden.py so that the committed fixtures of the other operators are not rewritten)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import numpy as np
import torch

from curvlinops import EFLinearOperator  # noqa: E402
from tests.golden_utils import GOLDEN, load_case  # noqa: E402

torch.set_default_dtype(torch.float64)
out = {}
for name in ["mlp_c1_ce_mean", "mlp_c1_ce_sum", "mlp_c1_mse_mean", "miniresnet_ce_mean"]:
    model, loss, data, fx = load_case(name)
    params = dict(model.named_parameters())
    E = EFLinearOperator(model, loss, params, data, check_deterministic=False)
    out[name] = (E @ fx["V"]).detach().numpy()
    print(name, out[name].shape, float(np.abs(out[name]).max()))
np.savez_compressed(os.path.join(GOLDEN, "ef.npz"), **out)
