"""Helpers to load the reference-generated fixtures of tests/golden (see oracle/make_golden.py)."""
import os

import numpy as np
import torch
from torch import nn

from oracle.models import ConvNetBias, MiniResNet, TiedNet, TokenMLP, TransformerBlock, mini_vit, mlp_c1, mlp_gelu, mlp_ln_gelu, mlp_smooth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BUILDERS = {
    "mlp_c1_ce_mean": (lambda: mlp_c1(), lambda: nn.CrossEntropyLoss()),
    "mlp_c1_ce_sum": (lambda: mlp_c1(), lambda: nn.CrossEntropyLoss(reduction="sum")),
    "mlp_c1_mse_mean": (lambda: mlp_c1(), lambda: nn.MSELoss()),
    "miniresnet_ce_mean": (lambda: MiniResNet(), lambda: nn.CrossEntropyLoss()),
    "mlp_bce_mean": (lambda: mlp_c1(classes=6, width=16), lambda: nn.BCEWithLogitsLoss()),
    "mlp_bce_sum": (lambda: mlp_c1(classes=6, width=16), lambda: nn.BCEWithLogitsLoss(reduction="sum")),
    "mlp_sigmoid_tanh_mse_sum": (lambda: mlp_smooth(), lambda: nn.MSELoss(reduction="sum")),
    "cnn_bias_ce_mean": (lambda: ConvNetBias(), lambda: nn.CrossEntropyLoss()),
    "mlp_ln_gelu_ce_mean": (lambda: mlp_ln_gelu(), lambda: nn.CrossEntropyLoss()),
    "token_mlp_ce_mean": (lambda: TokenMLP(), lambda: nn.CrossEntropyLoss()),
    "mlp_gelu_mse_sum": (lambda: mlp_gelu(), lambda: nn.MSELoss(reduction="sum")),
    "ggn_diag_mlp_ce_mean": (lambda: mlp_c1(classes=4, width=12), lambda: nn.CrossEntropyLoss()),
    "ggn_diag_mlp_mse_sum": (lambda: mlp_c1(classes=4, width=12), lambda: nn.MSELoss(reduction="sum")),
    "ggn_diag_cnn_ce_mean": (lambda: ConvNetBias(), lambda: nn.CrossEntropyLoss()),
    "kfac_mlp": (lambda: mlp_c1(classes=4, width=12), lambda: nn.CrossEntropyLoss()),
    "kfac_tokens": (lambda: TokenMLP(), lambda: nn.CrossEntropyLoss()),
    "kfac_tokens_reduce": (lambda: TokenMLP(), lambda: nn.CrossEntropyLoss()),
    "transformer_block_ce_mean": (lambda: TransformerBlock(dim=16, heads=2, hidden=32, layers=2),
                                  lambda: nn.CrossEntropyLoss()),
    "mini_vit_ce_mean": (lambda: mini_vit(), lambda: nn.CrossEntropyLoss()),
    "kfac_tied": (lambda: TiedNet(), lambda: nn.CrossEntropyLoss()),
    "kfac_cnn": (lambda: ConvNetBias(), lambda: nn.CrossEntropyLoss()),
}


def load_case(name, dtype=torch.float64, device="cpu"):
    """Returns (model, loss, data, fixture dict of tensors/arrays)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    model = BUILDERS[name][0]().to(torch.float64)
    sd = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    model.load_state_dict(sd)
    model = model.eval().to(dtype).to(device)
    data = []
    for i in range(int(z["n_batches"])):
        X = torch.from_numpy(z[f"X{i}"]).to(dtype).to(device)
        y = torch.from_numpy(z[f"y{i}"])
        y = (y.to(dtype) if y.is_floating_point() else y).to(device)
        data.append((X, y))
    fx = {}
    for k in z.files:
        if k.startswith("param::") or k[0] in "Xy" and k[1:].isdigit() or k == "n_batches":
            continue
        a = z[k]
        fx[k] = torch.from_numpy(a) if a.dtype.kind == "f" and a.ndim > 0 else a
    return model, BUILDERS[name][1](), data, fx


def split_like(V, params):
    """[P, K] -> list of [*shape, K] (the reference's tensor-list format)."""
    sizes = [p.numel() for p in params.values()]
    return [v.reshape(*p.shape, V.shape[1]) for v, p in zip(V.split(sizes), params.values())]


def flat(Vlist):
    return torch.cat([v.reshape(-1, v.shape[-1]) for v in Vlist])
