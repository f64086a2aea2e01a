"""Diagonal of the generalized Gauss-Newton matrix (reference ``curvlinops/ggn_diagonal.py:12-93`` and
``curvlinops/computers/ggn_diagonal.py:21-232``).

``diag(G) = w * sum_{n, v} (J_n^T g'_{n,v})^2`` with ``g'_{n,v}`` the columns of the loss Hessian's square root
(exact, ``mc_samples = 0``) or ``mc_samples`` sampled would-be gradients per datum (MC), ``w = 1/N`` for mean
reductions and 1 for sums.  The reference vmaps a VJP over data points and square-root columns; here it is the
per-example contraction of the EKFAC eigenvalue correction (``csrc/ekfac.cuh``) without the rotations: one
backward sweep per square-root column, the weight-gradient contraction split at the example boundaries so that its
split-K partials ARE the per-example gradients, then a square-and-sum finish.  Like the Kronecker-factored operators it
covers the parameters of ``Linear`` / ``Conv2d`` layers (any other parameter raises ``NotImplementedError``).
"""
from __future__ import annotations

from collections.abc import Callable, Iterable

import torch
from torch import Tensor
from torch.nn import Module

from .dense import DiagonalLinearOperator
from .kfac import EKFACComputer, FisherType


class GGNDiagonalComputer(EKFACComputer):
    """``compute() -> {parameter name: diagonal block shaped like the parameter}``."""

    def __init__(self, model_func: Module, loss_func, params: dict[str, Tensor], data: Iterable, progressbar: bool = False,
                 check_deterministic: bool = True, num_data: int | None = None, batch_size_fn: Callable | None = None,
                 mc_samples: int = 0, seed: int = 2_147_483_647):
        if mc_samples < 0:
            raise ValueError(f"mc_samples must be non-negative, got {mc_samples}.")
        if mc_samples > 0:
            self.FIXED_DATA_ORDER = True
        super().__init__(model_func, loss_func, params, data, progressbar=progressbar,
                         check_deterministic=check_deterministic, seed=seed,
                         fisher_type=FisherType.TYPE2 if mc_samples == 0 else FisherType.MC,
                         mc_samples=max(1, mc_samples), separate_weight_and_bias=False, num_data=num_data,
                         batch_size_fn=batch_size_fn)

    def compute(self) -> dict[str, Tensor]:
        self._engine._check_supported()
        lam = self._eigenvalue_correction(None, None, self._mapping, identity=True)
        out: dict[str, Tensor] = {}
        dt = self.dtype
        for group in self._mapping:
            block = lam[tuple(group.values())]
            if "W" in group:
                w = self._params[group["W"]]
                block = block.reshape(w.shape[0], -1)
                out[group["W"]] = block[:, : w[0].numel()].reshape(w.shape).to(dt)
                if "b" in group:
                    out[group["b"]] = block[:, -1].contiguous().to(dt)
            else:
                out[group["b"]] = block.reshape(-1).to(dt)
        return {n: out[n] for n in self._params}


class GGNDiagonalLinearOperator(DiagonalLinearOperator):
    r"""``diag(G)`` as a diagonal operator in the parameter space (same constructor as the reference's)."""

    def __init__(self, model_func: Module, loss_func, params: dict[str, Tensor], data: Iterable, progressbar: bool = False,
                 check_deterministic: bool = True, num_data: int | None = None, batch_size_fn: Callable | None = None,
                 mc_samples: int = 0, seed: int = 2_147_483_647):
        diag = GGNDiagonalComputer(model_func, loss_func, params, data, progressbar=progressbar,
                                   check_deterministic=check_deterministic, num_data=num_data,
                                   batch_size_fn=batch_size_fn, mc_samples=mc_samples, seed=seed).compute()
        super().__init__(list(diag.values()))
