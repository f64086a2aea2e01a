"""Generate tests/golden/*.npz by running the REFERENCE (/root/reference, read-only) in the
build container.  TEST INFRASTRUCTURE.  Run:  python oracle/make_golden.py

Needs the two stub packages of oracle/stubs (absent third-party imports off the hot path).
The fixtures hold parameters, data, V and the reference's result, all float64.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import numpy as np
import torch
from torch import nn

from curvlinops import (EKFACLinearOperator, GGNLinearOperator, HessianLinearOperator,  # noqa: E402
                        KFACLinearOperator)
from oracle.models import ConvNetBias, MiniResNet, mlp_c1, randomize_bn_  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_default_dtype(torch.float64)


def save(name, model, data, extra):
    d = {f"param::{k}": v.detach().numpy() for k, v in model.state_dict().items()}
    for i, (X, y) in enumerate(data):
        d[f"X{i}"], d[f"y{i}"] = X.numpy(), y.numpy()
    d["n_batches"] = np.array(len(data))
    d.update({k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in extra.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, {k: getattr(v, "shape", v) for k, v in extra.items()})


def curvature_cases(name, model, data, loss, K):
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    g = torch.Generator().manual_seed(1)
    V = torch.rand(P, K, generator=g)
    extra = {"V": V}
    G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
    extra["ggn"] = G @ V
    H = HessianLinearOperator(model, loss, params, data, check_deterministic=False)
    extra["hessian"] = H @ V
    for M in (1, 3):
        MC = GGNLinearOperator(model, loss, params, data, check_deterministic=False,
                               mc_samples=M, seed=1234)
        extra[f"ggn_mc{M}"] = MC @ V
    save(name, model, data, extra)


def kfac_cases(name, model, data, loss, damping=1e-2, backend="hooks", ekfac=True, kfac_approx="expand"):
    params = {n: p for n, p in model.named_parameters()
              if isinstance(dict(model.named_modules())[n.rsplit(".", 1)[0]], (nn.Linear, nn.Conv2d))}
    P = sum(p.numel() for p in params.values())
    g = torch.Generator().manual_seed(2)
    v = torch.rand(P, 2, generator=g)
    extra = {"v": v, "damping": np.array(damping), "param_names": np.array(list(params.keys()))}
    for ft in ("type-2", "mc", "empirical"):
        for sep in (False, True):
            tag = f"{ft.replace('-', '')}_{'sep' if sep else 'joint'}"
            Kop = KFACLinearOperator(model, loss, params, data, check_deterministic=False,
                                     fisher_type=ft, mc_samples=(2 if ft == 'mc' else 1), seed=77,
                                     separate_weight_and_bias=sep, backend=backend, kfac_approx=kfac_approx)
            extra[f"kfac_{tag}"] = Kop @ v
            extra[f"kfacinv_{tag}"] = Kop.inverse(damping=damping) @ v
            if ft == "type-2":
                _, Kb, _ = Kop
                for bi, blk in enumerate(Kb):
                    for fi, fac in enumerate(blk):
                        extra[f"factor_{tag}_{bi}_{fi}"] = fac
                if not ekfac:
                    continue
                E = EKFACLinearOperator(model, loss, params, data, check_deterministic=False,
                                        fisher_type=ft, separate_weight_and_bias=sep, backend=backend,
                                        kfac_approx=kfac_approx)
                extra[f"ekfac_{tag}"] = E @ v
                extra[f"ekfacinv_{tag}"] = E.inverse(damping=damping) @ v
    save(name, model, data, extra)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    # C1: MLP, CE mean, two unequal batches
    model = mlp_c1().eval()
    data = [(torch.randn(20, 64), torch.randint(0, 10, (20,))),
            (torch.randn(12, 64), torch.randint(0, 10, (12,)))]
    curvature_cases("mlp_c1_ce_mean", model, data, nn.CrossEntropyLoss(), K=3)
    curvature_cases("mlp_c1_ce_sum", model, data, nn.CrossEntropyLoss(reduction="sum"), K=2)
    # MSE
    data_mse = [(torch.randn(9, 64), torch.randn(9, 10)), (torch.randn(7, 64), torch.randn(7, 10))]
    curvature_cases("mlp_c1_mse_mean", model, data_mse, nn.MSELoss(), K=2)
    # mini ResNet (conv/bn-eval/relu/maxpool/residual/avgpool/fc)
    gen = torch.Generator().manual_seed(3)
    net = MiniResNet().eval()
    randomize_bn_(net, gen)
    data_r = [(torch.rand(5, 3, 32, 32), torch.randint(0, 10, (5,))),
              (torch.rand(3, 3, 32, 32), torch.randint(0, 10, (3,)))]
    curvature_cases("miniresnet_ce_mean", net, data_r, nn.CrossEntropyLoss(), K=3)
    # KFAC / EKFAC
    kfac_cases("kfac_mlp", mlp_c1(classes=4, width=12).eval(),
               [(torch.randn(6, 12), torch.randint(0, 4, (6,))),
                (torch.randn(4, 12), torch.randint(0, 4, (4,)))], nn.CrossEntropyLoss())
    kfac_cases("kfac_cnn", ConvNetBias().eval(),
               [(torch.rand(4, 3, 8, 8), torch.randint(0, 5, (4,))),
                (torch.rand(3, 3, 8, 8), torch.randint(0, 5, (3,)))], nn.CrossEntropyLoss())


if __name__ == "__main__":
    main()
