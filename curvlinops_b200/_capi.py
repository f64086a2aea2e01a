"""ctypes binding of ``libcurvb200.so`` (C ABI declared in ``include/curvb200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU
fallback: if the library is missing, :func:`lib` raises, and every compute entry point returns an
error when no CUDA device is present.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcurvb200.so")

# enums (keep in sync with include/curvb200.h)
(OP_INPUT, OP_CONV, OP_AFFINE, OP_RELU, OP_ADD, OP_MAXPOOL, OP_AVGPOOL, OP_SIGMOID, OP_TANH, OP_LAYERNORM, OP_GELU,
 OP_ATTENTION, OP_RESHAPE, OP_CLSCAT, OP_POSADD, OP_TOKSEL) = range(16)
LOSS_CE, LOSS_MSE, LOSS_BCE = range(3)
KIND_GGN, KIND_GGN_MC, KIND_HESSIAN, KIND_JVP, KIND_VJP, KIND_FORWARD = range(6)
ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 1, 2, 3, 4


class ValueDesc(C.Structure):
    _fields_ = [("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("has_tangent", C.c_int)]


class ParamDesc(C.Structure):
    _fields_ = [("numel", C.c_longlong), ("offset", C.c_longlong)]


class NodeDesc(C.Structure):
    _fields_ = [("op", C.c_int), ("in0", C.c_int), ("in1", C.c_int), ("out", C.c_int),
                ("p0", C.c_int), ("p1", C.c_int),
                ("c0", C.c_int), ("c1", C.c_int), ("c2", C.c_int), ("c3", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
                ("ph", C.c_int), ("pw", C.c_int), ("eps", C.c_float)]


EXPORTS = [
    "curv_program_create", "curv_program_destroy", "curv_program_workspace_bytes",
    "curv_program_value_layout", "curv_matmat_batch", "curv_matmat_batch_sync", "curv_kfac_accumulate_batch",
    "curv_kron_apply", "curv_kron_apply_tc", "curv_kron_apply_tc_workspace", "curv_kron_apply_tc_factor_bytes",
    "curv_eigh_apply", "curv_gemm", "curv_gemm_batched", "curv_ekfac_correction_batch",
    "curv_lanczos_reorth", "curv_lanczos_reorth_workspace", "curv_last_error", "curv_abi_version",
    "curv_launch_count", "curv_add_launch_count", "curv_set_tensor_core_mode", "curv_profile_enable", "curv_profile_read",
    "curv_profile_read_class", "curv_launch_config",
]

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises ``RuntimeError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            " curvlinops_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    L.curv_program_create.argtypes = [C.POINTER(ValueDesc), i, C.POINTER(NodeDesc), i,
                                      C.POINTER(ParamDesc), i, i, i, i, C.POINTER(vp)]
    L.curv_program_create.restype = i
    L.curv_program_destroy.argtypes = [vp]
    L.curv_program_destroy.restype = None
    L.curv_program_workspace_bytes.argtypes = [vp]
    L.curv_program_workspace_bytes.restype = C.c_size_t
    L.curv_program_value_layout.argtypes = [vp, i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                            C.POINTER(i)]
    L.curv_program_value_layout.restype = i
    L.curv_matmat_batch.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, i, vp, vp,
                                    i, i, i, f, f, vp, C.c_size_t, vp]
    L.curv_matmat_batch.restype = i
    L.curv_matmat_batch_sync.argtypes = L.curv_matmat_batch.argtypes + [C.POINTER(vp), C.POINTER(vp)]
    L.curv_matmat_batch_sync.restype = i
    L.curv_kfac_accumulate_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(i), i,
                                             C.POINTER(vp), C.POINTER(vp), C.POINTER(i), vp, i, f, f,
                                             vp, C.c_size_t, vp]
    L.curv_kfac_accumulate_batch.restype = i
    L.curv_ekfac_correction_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(i), i, C.POINTER(vp),
                                              C.POINTER(vp), C.POINTER(vp), C.POINTER(i), vp, i, f, vp, C.c_size_t, vp]
    L.curv_ekfac_correction_batch.restype = i
    L.curv_kron_apply.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp]
    L.curv_kron_apply.restype = i
    L.curv_kron_apply_tc.argtypes = [vp, vp, i, i, i, vp, vp, vp, C.c_size_t, i, vp, C.c_size_t, vp]
    L.curv_kron_apply_tc.restype = i
    L.curv_kron_apply_tc_workspace.argtypes = [i, i, i]
    L.curv_kron_apply_tc_workspace.restype = C.c_size_t
    L.curv_kron_apply_tc_factor_bytes.argtypes = [i, i]
    L.curv_kron_apply_tc_factor_bytes.restype = C.c_size_t
    L.curv_eigh_apply.argtypes = [vp, vp, vp, f, i, i, i, i, vp, vp, vp, vp, vp]
    L.curv_eigh_apply.restype = i
    L.curv_gemm.argtypes = [i, i, i, i, i, f, vp, i, vp, i, f, vp, i, vp]
    L.curv_gemm.restype = i
    L.curv_gemm_batched.argtypes = [i, i, i, i, i, f, vp, i, ll, vp, i, ll, f, vp, i, ll, i, vp]
    L.curv_gemm_batched.restype = i
    L.curv_lanczos_reorth_workspace.argtypes = [i, ll]
    L.curv_lanczos_reorth_workspace.restype = ll
    L.curv_lanczos_reorth.argtypes = [vp, ll, i, vp, ll, i, vp, vp, ll, vp]
    L.curv_lanczos_reorth.restype = i
    L.curv_last_error.argtypes = []
    L.curv_last_error.restype = C.c_char_p
    L.curv_abi_version.argtypes = []
    L.curv_abi_version.restype = i
    L.curv_launch_count.argtypes = []
    L.curv_launch_count.restype = ll
    L.curv_add_launch_count.argtypes = [ll]
    L.curv_add_launch_count.restype = None
    L.curv_set_tensor_core_mode.argtypes = [i]
    L.curv_set_tensor_core_mode.restype = i
    L.curv_profile_enable.argtypes = [i]
    L.curv_profile_enable.restype = i
    L.curv_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ll)]
    L.curv_profile_read.restype = i
    L.curv_profile_read_class.argtypes = [i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ll)]
    L.curv_profile_read_class.restype = i
    L.curv_launch_config.argtypes = []
    L.curv_launch_config.restype = i
    _lib = L
    return L


def check(rc: int) -> None:
    """Map a status code to the exception class the reference raises in the same situation."""
    if rc == 0:
        return
    msg = lib().curv_last_error().decode()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ptr_array(ptrs) -> C.Array:
    arr = (C.c_void_p * max(1, len(ptrs)))()
    for j, p in enumerate(ptrs):
        arr[j] = p
    return arr
