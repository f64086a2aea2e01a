"""World-size-2 gloo test (CPU) of the data-parallel plumbing: batch sharding with global normalisation
and the single all-reduce of the [P, K] result.  The CUDA engine is replaced by a stand-in that computes
the exact mini-batch GGN product of a linear model with torch ops (TEST ONLY: it exercises
curvlinops_b200.dist and CurvatureLinearOperator._matmat, not the kernels)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from curvlinops_b200 import dist as cdist


def test_shard_bounds_cover_batch():
    for B in (1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [cdist.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeEngine:
    """GGN of f(x) = W x with MSE loss: G = scale * 2 * sum_n (x_n x_n^T (x) I) -- computed densely."""

    def __init__(self, W):
        self.W = W

    def matmat_batch(self, kind, X, y, V, out, alpha, mc_grad=None, scale=None, cols=None):
        o, i = self.W.shape
        if scale is None:
            scale = 1.0 / (X.shape[0] * o)
        A = X.T @ X  # [i, i]
        c0, cn = (0, V.shape[1]) if cols is None else cols
        Vm = V[:, c0:c0 + cn].reshape(o, i, -1)
        out[:, c0:c0 + cn] += (alpha * 2.0 * scale) * torch.einsum("ij,ojk->oik", A, Vm).reshape(o * i, -1)

    def _check_supported(self):
        pass


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from curvlinops_b200.curvature import GGNLinearOperator

        torch.manual_seed(0)
        W = torch.rand(3, 5)
        data = [(torch.rand(7, 5), torch.rand(7, 3)), (torch.rand(4, 5), torch.rand(4, 3))]
        V = torch.rand(15, 8 if world == 4 else 2)  # world 4, K = 8: 2-d sharding (2 batch groups x 2 column groups)

        def build():
            op = GGNLinearOperator.__new__(GGNLinearOperator)
            op._params = {"W": W}
            op._loss_func = torch.nn.MSELoss()
            op._data = data
            op._progressbar = False
            op._batch_size_fn = lambda X: X.shape[0]
            op._N_data = 11
            op._mc_samples = 0
            op._engine = _FakeEngine(W)
            from curvlinops_b200.linop import PyTorchLinearOperator
            PyTorchLinearOperator.__init__(op, [(3, 5)], [(3, 5)])
            return op

        cdist.enable(False)
        single = build() @ V
        cdist.enable(True)
        sharded = build() @ V
        cdist.enable(False)
        ok = torch.allclose(single, sharded, rtol=1e-5, atol=1e-7)
        flag = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(bool(flag.item()))
    finally:
        dist.destroy_process_group()


def test_grid():
    assert cdist.grid(1, 8) == (1, 1) and cdist.grid(2, 8) == (2, 1)
    assert cdist.grid(4, 8) == (2, 2) and cdist.grid(8, 8) == (4, 2)
    assert cdist.grid(8, 4) == (8, 1) and cdist.grid(8, 1) == (8, 1) and cdist.grid(5, 8) == (5, 1)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_matmat_equals_single_process(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
