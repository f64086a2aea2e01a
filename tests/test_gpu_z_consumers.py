"""GPU tests of the consumers of the engine's products: CG / Neumann inverses of the
damped GGN, with a KFAC inverse as preconditioner, and the randomised trace / diagonal estimators.  The checks
themselves live in ``tests/consumer_checks.py`` and are validated on the CPU against dense operators; here the
operator is the engine's GGN (fp32) and the dense matrix is that same GGN applied to the identity."""
import pytest
import torch

from curvlinops_b200 import GGNLinearOperator, KFACLinearOperator
from tests.consumer_checks import check_damped_inverses, check_estimators
from tests.golden_utils import load_case

pytestmark = pytest.mark.gpu


def _ggn(name):
    model, loss, data, _ = load_case(name, dtype=torch.float32, device="cuda")
    params = dict(model.named_parameters())
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    dense = G @ torch.eye(G.shape[1], device=G.device, dtype=G.dtype)
    return model, loss, data, params, G, dense


@pytest.mark.parametrize("name", ["kfac_mlp"])
def test_inverses_of_damped_ggn(name):
    model, loss, data, params, G, dense = _ggn(name)
    shapes = [tuple(p.shape) for p in params.values()]
    delta = 0.1 * dense.diag().mean().item()
    kfac_inv = KFACLinearOperator(model, loss, params, data, fisher_type="type-2", check_deterministic=False
                                  ).inverse(damping=delta, use_exact_damping=True)  # (G (x) A + delta I)^-1 per block
    iters = check_damped_inverses(G, dense, shapes, preconditioner=kfac_inv.__matmul__, delta_rel=0.1)
    assert all(1 <= n <= 300 for n in iters.values()), iters


@pytest.mark.parametrize("name", ["kfac_mlp"])
def test_estimators_on_ggn(name):
    *_, G, dense = _ggn(name)
    check_estimators(G, dense)
