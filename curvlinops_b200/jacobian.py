"""Jacobian and transposed-Jacobian operators on the engine's forward+Jv / backward+J^T sweeps.

Interfaces of the reference's ``JacobianLinearOperator`` / ``TransposedJacobianLinearOperator``
(``curvlinops/jacobian.py:108-358``); the per-mini-batch products are ``CURV_KIND_JVP`` / ``CURV_KIND_VJP``
of ``curv_matmat_batch`` instead of vmapped ``torch.func.jvp`` / ``vjp`` (``jacobian.py:47,92``).
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import _capi as capi
from .curvature import CurvatureLinearOperator
from .linop import PyTorchLinearOperator


class _JacobianBase(CurvatureLinearOperator):
    SELF_ADJOINT = False
    FIXED_DATA_ORDER = True
    _matmat_flat = None  # not a square parameter-space operator: no flat [P, K] fast path

    def __init__(self, model_func, params, data, progressbar=False, check_deterministic=True, num_data=None,
                 batch_size_fn=None):
        if not isinstance(next(iter(data))[0], Tensor):
            raise NotImplementedError("The Jacobian operators of the B200 engine need tensor inputs X.")
        self._ctor_args = dict(progressbar=progressbar, batch_size_fn=batch_size_fn)
        super().__init__(model_func, None, params, data, progressbar=progressbar,
                         check_deterministic=check_deterministic, num_data=num_data, batch_size_fn=batch_size_fn)
        # the operator maps parameter space <-> the stacked predictions [N_data, C]
        X0 = next(iter(self._data))[0]
        C = self._engine.predict(X0.to(self.device)).shape[1]
        p_shapes = [tuple(p.shape) for p in self._params.values()]
        in_shape, out_shape = self._orient(p_shapes, [(self._N_data, C)])
        PyTorchLinearOperator.__init__(self, in_shape, out_shape)

    def _check_deterministic_matvec(self, *a, **k):
        return None  # shapes are only final after __init__; the data/prediction probes already ran

    def _other(self, cls):
        return cls(self._model_func, self._params, self._data, check_deterministic=False, num_data=self._N_data,
                   **self._ctor_args)


class JacobianLinearOperator(_JacobianBase):
    """``J``: parameter space -> predictions of the whole data set, ``[N_data * C, P]``."""

    @staticmethod
    def _orient(p_shapes, pred_shape):
        return p_shapes, pred_shape

    def _matmat(self, M: list[Tensor]) -> list[Tensor]:
        V = self._flat_matrix(M)
        K = V.shape[1]
        outs = []
        for X, _ in self._loop_over_data(desc="_matmat"):
            if not isinstance(X, Tensor):
                raise NotImplementedError("The B200 engine needs tensor inputs X.")
            C = self._out_shape[0][1]
            out = torch.zeros(X.shape[0], C, K, device=V.device, dtype=torch.float32)
            self._engine.matmat_batch(capi.KIND_JVP, X, None, V, out, 1.0)
            outs.append(out)
        return [torch.cat(outs).to(M[0].dtype)]

    def _adjoint(self):
        return self._other(TransposedJacobianLinearOperator)


class TransposedJacobianLinearOperator(_JacobianBase):
    """``J^T``: stacked prediction-space vectors ``[N_data, C]`` -> parameter space."""

    @staticmethod
    def _orient(p_shapes, pred_shape):
        return pred_shape, p_shapes

    def _matmat(self, M: list[Tensor]) -> list[Tensor]:
        (W,) = M  # [N_data, C, K]
        K = W.shape[-1]
        P = sum(p.numel() for p in self._params.values())
        out = torch.zeros(P, K, device=W.device, dtype=torch.float32)
        pos = 0
        for X, _ in self._loop_over_data(desc="_matmat"):
            B = self._batch_size_fn(X)
            Wb = W[pos:pos + B].to(torch.float32).contiguous()
            self._engine.matmat_batch(capi.KIND_VJP, X, None, Wb, out, 1.0)
            pos += B
        parts = out.split([p.numel() for p in self._params.values()])
        return [o.reshape(*p.shape, K).to(W.dtype) for o, p in zip(parts, self._params.values())]

    def _adjoint(self):
        return self._other(JacobianLinearOperator)
