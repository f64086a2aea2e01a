"""curvlinops_b200: B200-native engine for the curvature-matvec hot path of f-dangel/curvlinops.

Drop-in operator classes (same names / constructor arguments as the reference):
``HessianLinearOperator``, ``GGNLinearOperator``, ``EFLinearOperator``, ``GGNDiagonalLinearOperator``,
``KFACLinearOperator``,
``EKFACLinearOperator``, the Jacobian operators, the structured operators they are assembled from, and the
consumers of their products (CG / Neumann / LSMR inverses, Lanczos, randomised trace / diagonal estimators).  All arithmetic runs in hand-written sm_100a CUDA kernels
behind the C ABI of ``include/curvb200.h``; there is no CPU fallback.
"""

from .curvature import (CurvatureLinearOperator, EFLinearOperator, GGNLinearOperator,
                        HessianLinearOperator)
from .dense import DiagonalLinearOperator, IdentityLinearOperator, TensorLinearOperator
from .estimators import (hutchinson_diag, hutchinson_squared_fro, hutchinson_trace, hutchpp_trace, xdiag,
                         xtrace)
from .ggn_diagonal import GGNDiagonalComputer, GGNDiagonalLinearOperator
from .inverse import CGInverseLinearOperator, LSMRInverseLinearOperator, NeumannInverseLinearOperator
from .jacobian import JacobianLinearOperator, TransposedJacobianLinearOperator
from .kfac import EKFACLinearOperator, FisherType, KFACLinearOperator, KFACType
from .lanczos import fast_lanczos, lanczos_eigsh
from .linop import PyTorchLinearOperator
from .structured import (BlockDiagonalLinearOperator, EighDecomposedLinearOperator,
                         FromCanonicalLinearOperator, KroneckerProductLinearOperator,
                         ToCanonicalLinearOperator)

__all__ = [
    "PyTorchLinearOperator",
    "CurvatureLinearOperator",
    "GGNLinearOperator",
    "EFLinearOperator",
    "GGNDiagonalLinearOperator",
    "GGNDiagonalComputer",
    "fast_lanczos",
    "lanczos_eigsh",
    "HessianLinearOperator",
    "JacobianLinearOperator",
    "TransposedJacobianLinearOperator",
    "KFACLinearOperator",
    "EKFACLinearOperator",
    "FisherType",
    "KFACType",
    "KroneckerProductLinearOperator",
    "EighDecomposedLinearOperator",
    "BlockDiagonalLinearOperator",
    "ToCanonicalLinearOperator",
    "FromCanonicalLinearOperator",
    "TensorLinearOperator",
    "DiagonalLinearOperator",
    "IdentityLinearOperator",
    "CGInverseLinearOperator",
    "LSMRInverseLinearOperator",
    "NeumannInverseLinearOperator",
    "hutchinson_trace",
    "hutchpp_trace",
    "xtrace",
    "hutchinson_diag",
    "xdiag",
    "hutchinson_squared_fro",
]
