"""Generate tests/golden/{transformer_block_ce_mean, mini_vit_ce_mean}.npz by running the REFERENCE (/root/reference,
read-only): GGN, MC-GGN (reference RNG stream) and empirical Fisher of (i) a pre-norm transformer encoder block on token
sequences (nn.MultiheadAttention, batch_first) and (ii) torchvision's VisionTransformer at toy size (patch convolution,
class token, position embedding, two encoder layers, class-token read-out), two unequal mini-batches each.  The reference
differentiates attention through torch.func, which needs the math attention path (SURVEY 8c recipe;
docs/examples/basic_usage/benchmark_utils.py:91-98).  TEST INFRASTRUCTURE.  Run: python oracle/make_golden_attention.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference", os.path.dirname(HERE)]

import torch
from torch import nn
from torch.nn.attention import SDPBackend, sdpa_kernel

from curvlinops import EFLinearOperator, GGNLinearOperator  # noqa: E402
from oracle.make_golden import save  # noqa: E402
from oracle.models import TransformerBlock, mini_vit  # noqa: E402

torch.set_default_dtype(torch.float64)


def run(name, model, data, loss):
    model = model.train()  # dropout p = 0: train() only steers nn.MultiheadAttention off its fused inference path
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    V = torch.rand(P, 3, generator=torch.Generator().manual_seed(1))
    extra = {"V": V}
    with sdpa_kernel(SDPBackend.MATH):
        extra["ggn"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False) @ V
        extra["ef"] = EFLinearOperator(model, loss, params, data, check_deterministic=False) @ V
        for M in (1, 3):
            extra[f"ggn_mc{M}"] = GGNLinearOperator(model, loss, params, data, check_deterministic=False, mc_samples=M,
                                                    seed=1234) @ V
    save(name, model, data, extra)


torch.manual_seed(71)
run("transformer_block_ce_mean", TransformerBlock(dim=16, heads=2, hidden=32, layers=2, sharpen=True),
    [(torch.randn(5, 7, 16), torch.randint(0, 5, (5,))), (torch.randn(3, 7, 16), torch.randint(0, 5, (3,)))],
    nn.CrossEntropyLoss())
torch.manual_seed(72)
run("mini_vit_ce_mean", mini_vit(),
    [(torch.rand(4, 3, 32, 32), torch.randint(0, 5, (4,))), (torch.rand(3, 3, 32, 32), torch.randint(0, 5, (3,)))],
    nn.CrossEntropyLoss())
