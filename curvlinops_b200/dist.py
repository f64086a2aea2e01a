"""Data-parallel sharding of the mini-batch sum (new functionality, the reference is single-device).

``G V = sum_b alpha_b sum_{n in batch b} (...)`` is linear in the data (reference
``curvlinops/_torch_base.py:937-944``; no cross-sample coupling with BatchNorm in eval mode), so each
rank processes a contiguous slice of every mini-batch with the *global* normalisation
(``1/B_global`` inside the loss Hessian, ``B_global/N`` outside) and the flat ``[P, K]`` result (or the
Kronecker factors) is summed with ONE all-reduce -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
Opt-in: call :func:`enable` after ``torch.distributed.init_process_group``.
"""

from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor
from torch.nn import CrossEntropyLoss

_enabled = False
_group = None


def enable(flag: bool = True, group=None) -> None:
    """Turn batch sharding + all-reduce on/off for all engine-backed operators of this process."""
    global _enabled, _group
    _enabled, _group = flag, group


def rank_world() -> tuple[int, int]:
    if _enabled and dist.is_available() and dist.is_initialized():
        return dist.get_rank(_group), dist.get_world_size(_group)
    return 0, 1


def grid(world: int, K: int) -> tuple[int, int]:
    """``(batch groups, column groups)`` of the 2-d sharding of one product over ``world`` ranks.  Rank ``r`` takes
    batch shard ``r // column_groups`` and the columns of group ``r % column_groups``; the one all-reduce of ``[P, K]``
    then sums over the batch groups and fills in the other groups' columns (they are zero on this rank).  Splitting
    the columns as well keeps twice the samples per rank (fuller tiles) and halves the per-product work that does not
    depend on the batch (packing the K tangent weights, split-K finishes); it costs one extra primal forward per column
    group (+3 % FLOPs for K = 8), so it is only used from 4 ranks on and keeps at least 4 columns per rank (the
    multi-slot weight-gradient kernel multiplies 4 slots per MMA)."""
    cg = 2 if (world >= 4 and world % 2 == 0 and K % 2 == 0 and K // 2 >= 4) else 1
    return world // cg, cg


def shard_bounds(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of a mini-batch owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(X: Tensor, y: Tensor, rank: int, world: int, loss_func, engine=None):
    """Local slice of ``(X, y)`` and the loss-Hessian constants ``(exact, mc)`` of the GLOBAL batch."""
    if not isinstance(X, Tensor):
        raise NotImplementedError("Data-parallel sharding needs tensor inputs X.")
    B = X.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    if loss_func is None or loss_func.reduction == "sum":
        scales = (1.0, 1.0)
    elif isinstance(loss_func, CrossEntropyLoss):
        scales = (1.0 / B, 1.0 / B)
    else:
        scales = (1.0 / (B * y.shape[-1]), 1.0 / B)
    if hi == lo:
        return None, None, scales
    Xs = X[lo:hi]
    Xs._curv_shard = (lo, hi, B)  # global sample indices of the slice (Monte-Carlo draws are keyed on them)
    return Xs, y[lo:hi], scales


def all_reduce_sum(t: Tensor) -> None:
    """In-place sum over ranks (one collective per matmat / factor build)."""
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)


def all_gather_rows(full: Tensor, part: Tensor) -> None:
    """``full[r * n:(r + 1) * n] = part`` of rank ``r`` (``part``: ``[n, K]``, ``full``: ``[world * n, K]``)."""
    dist.all_gather_into_tensor(full, part, group=_group)


def reduce_scatter_rows(part: Tensor, full: Tensor) -> None:
    """``part`` of rank ``r`` = sum over ranks of ``full[r * n:(r + 1) * n]``."""
    dist.reduce_scatter_tensor(part, full, op=dist.ReduceOp.SUM, group=_group)


def shared_pinned_tensor(name: str, shape, dtype=torch.float32, device=None) -> Tensor:
    """A host tensor backed by POSIX shared memory (``/dev/shm/<name>``) that every rank of one box can map, registered
    with CUDA as pinned memory: one copy of V / of the result in host memory feeds all GPUs of the box (each rank
    moves its row block, see ``CurvatureLinearOperator.matmat_pinned``).  The first caller creates the file; remove it
    with :func:`release_shared` when done."""
    import math

    n = math.prod(shape)
    t = torch.from_file(f"/dev/shm/{name}", shared=True, size=n, dtype=dtype).view(*shape)
    if torch.cuda.is_available():
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n * t.element_size(), 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed with code {int(rc)}")
    return t


def release_shared(name: str, tensor: Tensor | None = None) -> None:
    import os

    if tensor is not None and torch.cuda.is_available():
        torch.cuda.cudart().cudaHostUnregister(tensor.data_ptr())
    try:
        os.unlink(f"/dev/shm/{name}")
    except FileNotFoundError:
        pass
