"""Curvature linear operators backed by the CUDA engine.

Same constructor signatures, attributes and error behaviour as the reference's
``CurvatureLinearOperator`` (``curvlinops/_torch_base.py:817-1007``) with its empirical-risk mixin
(``curvlinops/_empirical_risk.py:20-439``), ``GGNLinearOperator`` (``curvlinops/ggn.py:171-366``) and
``HessianLinearOperator`` (``curvlinops/hessian.py:72-145``).  What differs is *how* a mini-batch
product is computed: instead of ``vmap(jvp/vjp)`` through autograd, ``_matmat`` hands the whole
``[P, K]`` matrix to ``curv_matmat_batch`` (``include/curvb200.h``), which runs the fused
forward+Jv / backward+J^T sweeps.  With ``torch.distributed`` initialised, each rank processes its
slice of every mini-batch and the ``[P, K]`` result is summed with one NCCL all-reduce
(linearity of the data sum, ``_torch_base.py:937-944``).
"""

from __future__ import annotations

import copy
import os
import weakref
from collections.abc import Callable, Iterable, MutableMapping

import torch
from torch import Tensor
from torch.nn import BCEWithLogitsLoss, CrossEntropyLoss, Module, MSELoss

from . import _capi as capi
from . import dist as cdist
from .engine import Engine
from .linop import PyTorchLinearOperator, report_allclose


class _FunctionalCall:
    """``(params, X) -> module(X)`` as a picklable callable (operators are saved with ``torch.save``)."""

    def __init__(self, module: Module):
        self.module = module

    def __call__(self, params: dict[str, Tensor], *inputs):
        return torch.func.functional_call(self.module, params, inputs)


def make_functional_call(module: Module) -> Callable:
    """``(params, X) -> module(X)`` with ``params`` overriding the module's own tensors
    (role of ``curvlinops/utils.py:267-297``)."""
    return _FunctionalCall(module)


def _leading_dim(X) -> int:
    """Default ``batch_size_fn`` (``_empirical_risk.py:83-85``); a named function so operators stay picklable."""
    return X.shape[0]


class CurvatureLinearOperator(PyTorchLinearOperator):
    """Base class of the engine-backed curvature matrices.

    Attributes:
        FIXED_DATA_ORDER: the data loader must yield identical batches in identical order.
        KIND: which product ``curv_matmat_batch`` computes.
    """

    FIXED_DATA_ORDER: bool = False
    NEEDS_NUM_PER_EXAMPLE_LOSS_TERMS: bool = False
    KIND: int = capi.KIND_GGN
    SELF_ADJOINT: bool = True

    def __init__(
        self,
        model_func: Module | Callable[[dict[str, Tensor], Tensor | MutableMapping], Tensor],
        loss_func: Callable[[Tensor, Tensor], Tensor] | None,
        params: dict[str, Tensor],
        data: Iterable[tuple[Tensor | MutableMapping, Tensor]],
        progressbar: bool = False,
        check_deterministic: bool = True,
        num_data: int | None = None,
        num_per_example_loss_terms: int | None = None,
        batch_size_fn: Callable[[MutableMapping | Tensor], int] | None = None,
    ):
        if isinstance(next(iter(data))[0], MutableMapping) and batch_size_fn is None:
            raise ValueError("When using dict-like custom data, `batch_size_fn` is required.")
        if not isinstance(params, dict):
            raise TypeError(
                f"params must be a dict[str, Tensor], got {type(params).__name__}. "
                "Use dict(model.named_parameters()) instead of list(model.parameters())."
            )
        if isinstance(model_func, Module):
            self._model_func = make_functional_call(model_func)
        elif callable(model_func):
            self._model_func = model_func
        else:
            raise ValueError(
                f"model_func must be an nn.Module or a callable, got {type(model_func).__name__}."
            )
        self._params = params
        self._loss_func = loss_func
        self._data = data
        self._progressbar = progressbar
        self._batch_size_fn = _leading_dim if batch_size_fn is None else batch_size_fn
        self._input_key = None
        first_X = next(iter(data))[0]
        if isinstance(first_X, MutableMapping):
            self._adapt_dict_inputs(first_X)
        self._engine = Engine(self._model_func, loss_func, params)
        self._N_data, self._num_per_example_loss_terms = self._get_data_statistics(
            num_data, num_per_example_loss_terms
        )
        if check_deterministic:
            self._check_deterministic()
        shapes = [tuple(p.shape) for p in params.values()]
        PyTorchLinearOperator.__init__(self, shapes, shapes)
        if check_deterministic:
            self._check_deterministic_matvec()

    # ---- dict-like inputs (reference _empirical_risk.py:37-119, test/cases.py:36-60) ------------------
    def _adapt_dict_inputs(self, first_X: MutableMapping) -> None:
        """The engine differentiates a function of ONE input tensor.  For dict-like mini-batches the entry that
        holds it (the only floating-point tensor of shape ``[B, C]`` / ``[B, C, H, W]``) becomes that tensor and
        the model function is wrapped to put it back into a copy of the mapping; the remaining entries must not
        be tensors and are taken from the first mini-batch (they are constants of the traced program).  The data
        loop then yields the tensor, tagged with the batch size ``batch_size_fn`` reports for the mapping."""
        tensors = {k: v for k, v in first_X.items() if isinstance(v, Tensor)}
        keys = [k for k, v in tensors.items() if v.is_floating_point() and v.ndim in (2, 4)]
        if len(keys) != 1 or len(tensors) != 1:
            raise NotImplementedError(
                "The B200 engine supports dict-like inputs with exactly one tensor entry (floating point, shape "
                f"[B, C] or [B, C, H, W]); got tensor entries { {k: tuple(v.shape) for k, v in tensors.items()} }."
            )
        (key,) = keys
        template, inner, user_bs = first_X, self._model_func, self._batch_size_fn

        def model_on_tensor(params, Xt):
            X = copy.copy(template)
            X[key] = Xt
            return inner(params, X)

        self._input_key = key
        self._model_func = model_on_tensor
        self._batch_size_fn = lambda X: getattr(X, "_curv_batch_size", None) or user_bs(X)

    # ---- properties ----------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        devs = {p.device for p in self._params.values()}
        if len(devs) != 1:
            raise RuntimeError(f"Could not infer device. Parameters live on {devs}.")
        return devs.pop()

    @property
    def dtype(self) -> torch.dtype:
        dts = {p.dtype for p in self._params.values()}
        if len(dts) != 1:
            raise RuntimeError(f"Could not infer data type. Parameters have types {dts}.")
        return dts.pop()

    # ---- data loop (reference _empirical_risk.py:121-177, 311-352) --------------------------------
    def _get_data_statistics(self, num_data, num_per_example_loss_terms):
        need_n = num_data is None
        need_t = (self.NEEDS_NUM_PER_EXAMPLE_LOSS_TERMS and self._loss_func is not None
                  and num_per_example_loss_terms is None)
        if not need_n and not need_t:
            return num_data, num_per_example_loss_terms
        n_acc, t_acc = 0, 0
        for X, y in self._loop_over_data(desc="data_statistics"):
            n_acc += self._batch_size_fn(X)
            t_acc += y.numel() if isinstance(self._loss_func, CrossEntropyLoss) else y.shape[:-1].numel()
        N = n_acc if need_n else num_data
        if need_t:
            if t_acc % N != 0:
                raise ValueError(
                    "The number of loss terms must be divisible by the number of data points; "
                    f"num_loss_terms={t_acc}, N_data={N}."
                )
            num_per_example_loss_terms = t_acc // N
        return N, num_per_example_loss_terms

    def _loop_over_data(self, desc: str | None = None, to_device: bool = True, stage: bool = False):
        """``stage``: uploads land in buffers the operator keeps (see :meth:`_upload`); only the product loops ask for
        it -- two simultaneous iterations (the determinism probes) must not share buffers."""
        it = self._data
        dev = self.device if to_device else None
        if self._progressbar:
            from tqdm import tqdm

            it = tqdm(it, desc=f"{self.__class__.__name__}{'' if desc is None else '.' + desc} (on {dev})")
        for i, (X, y) in enumerate(it):
            if isinstance(X, Tensor):
                X = (self._upload(X, (i, "X")) if stage else X.to(dev)) if to_device else X
            elif self._input_key is not None:
                Xt = X[self._input_key]
                Xt = (Xt.to(dev) if to_device else Xt).detach()  # a new tensor object: the tag stays off the caller's
                Xt._curv_batch_size = self._batch_size_fn(X)
                X = Xt
            yield X, ((self._upload(y, (i, "y")) if stage else y.to(dev)) if to_device else y)

    def _upload(self, t: Tensor, key) -> Tensor:
        """Host-resident mini-batches are uploaded by every product (``_torch_base.py:937-944`` does the same).  The
        first few land in device buffers the operator keeps, so their addresses do not change from product to product
        and the captured CUDA graph of a product stays valid (a fresh allocation per upload made every host-data
        product run eagerly or re-capture); data loaders with many batches fall back to plain uploads."""
        dev = self.device
        if t.device == dev:
            return t
        stage = self.__dict__.setdefault("_upload_stage", {})
        buf = stage.get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            if buf is None and len(stage) >= 8:
                return t.to(dev, non_blocking=True)
            buf = stage[key] = torch.empty(t.shape, dtype=t.dtype, device=dev)
        buf.copy_(t, non_blocking=True)
        return buf

    def _get_normalization_factor(self, X, y) -> float:
        return {"sum": 1.0, "mean": self._batch_size_fn(X) / self._N_data}[self._loss_func.reduction]

    # ---- gradient / loss through the engine (reference _empirical_risk.py:354-439) -----------------
    def _batch_prediction_loss_gradient(self, X, y):
        """prediction, weighted loss and weighted parameter gradient of one mini-batch:
        the gradient is ``J^T (dl/df)`` with ``J^T`` applied by the engine (CURV_KIND_VJP)."""
        pred = self._engine.predict(X)
        if self._loss_func is None:
            return pred, None, None
        w = self._get_normalization_factor(X, y)
        f = pred.detach().requires_grad_(True)
        with torch.enable_grad():
            loss = self._loss_func(f, y) * w
            (gf,) = torch.autograd.grad(loss, f)
        P = sum(p.numel() for p in self._params.values())
        out = torch.zeros(P, 1, device=pred.device, dtype=torch.float32)
        self._engine.matmat_batch(capi.KIND_VJP, X, None, gf.reshape(*gf.shape, 1).contiguous(), out, 1.0)
        grads = [g.reshape(p.shape) for g, p in zip(out[:, 0].split([p.numel() for p in self._params.values()]),
                                                    self._params.values())]
        return pred, loss.detach(), grads

    def _check_deterministic(self, rtol: float = 5e-5, atol: float = 1e-6):
        """Two passes over the data must give the same total loss / gradient (and, with
        ``FIXED_DATA_ORDER``, the same batches): reference ``_empirical_risk.py:179-291``."""
        has_loss = self._loss_func is not None
        tot = [None, None]
        for (X1, y1), (X2, y2) in zip(self._loop_over_data(), self._loop_over_data()):
            r1 = self._batch_prediction_loss_gradient(X1, y1)
            r2 = self._batch_prediction_loss_gradient(X2, y2)
            if self.FIXED_DATA_ORDER:
                if isinstance(X1, Tensor) and not report_allclose(X1, X2, rtol=rtol, atol=atol):
                    raise RuntimeError("Check for deterministic X failed.")
                if not report_allclose(y1, y2, rtol=rtol, atol=atol):
                    raise RuntimeError("Check for deterministic y failed.")
                if not report_allclose(r1[0], r2[0], rtol=rtol, atol=atol):
                    raise RuntimeError("Check for deterministic batch prediction failed.")
                if has_loss:
                    if not report_allclose(r1[1], r2[1], rtol=rtol, atol=atol):
                        raise RuntimeError("Check for deterministic batch loss failed.")
                    if any(not report_allclose(a, b, rtol=rtol, atol=atol) for a, b in zip(r1[2], r2[2])):
                        raise RuntimeError("Check for deterministic batch gradient failed.")
            if has_loss:
                for j, r in enumerate((r1, r2)):
                    if tot[j] is None:
                        tot[j] = [r[1].clone(), [g.clone() for g in r[2]]]
                    else:
                        tot[j][0] += r[1]
                        for a, b in zip(tot[j][1], r[2]):
                            a += b
        if has_loss and tot[0] is not None:
            if not report_allclose(tot[0][0], tot[1][0], rtol=rtol, atol=atol):
                raise RuntimeError("Check for deterministic total loss failed.")
            if any(not report_allclose(a, b, rtol=rtol, atol=atol) for a, b in zip(tot[0][1], tot[1][1])):
                raise RuntimeError("Check for deterministic total gradient failed.")

    def _gradient_and_loss(self) -> tuple[list[Tensor], Tensor]:
        if self._loss_func is None:
            raise ValueError("No loss function specified.")
        total_loss, total_grad = None, None
        for X, y in self._loop_over_data(desc="gradient_and_loss"):
            _, loss, grads = self._batch_prediction_loss_gradient(X, y)
            if total_grad is None:
                total_loss, total_grad = loss.clone(), [g.clone() for g in grads]
            else:
                total_loss += loss
                for a, b in zip(total_grad, grads):
                    a += b
        return total_grad, total_loss

    # ---- the product (reference _torch_base.py:923-944) -------------------------------------------
    def _flat_matrix(self, M: list[Tensor]) -> Tensor:
        K = M[0].shape[-1]
        return torch.cat([m.reshape(-1, K) for m in M]).to(torch.float32).contiguous()

    def _unflatten(self, out: Tensor, like: list[Tensor]) -> list[Tensor]:
        K = out.shape[1]
        parts = out.split([s.numel() for s in self._out_shape])
        return [p.reshape(*s, K).to(m.dtype) for p, s, m in zip(parts, self._out_shape, like)]

    def _batch_call(self, X, y, V: Tensor, out: Tensor, alpha: float):
        self._engine.matmat_batch(self.KIND, X, y, V, out, alpha)

    def _matmat(self, M: list[Tensor]) -> list[Tensor]:
        return self._unflatten(self._product(self._flat_matrix(M)), M)

    def _matmat_flat(self, V: Tensor) -> Tensor:
        """``A @ V`` for a flat ``[P, K]`` matrix without the split into per-parameter views and the two
        concatenations of the list format (``_torch_base.py:289-298,414-422``): the engine works on this layout."""
        return self._product(V.to(torch.float32).contiguous()).to(V.dtype)

    def _local_batches(self, desc: str, shard: tuple[int, int] | None = None):
        """Mini-batches as this rank processes them: ``(X, y, alpha, scales)`` on the operator's device.  With batch
        sharding on, only the rank's contiguous slice of each mini-batch is moved to the device (``X`` is ``None``
        for a rank whose slice is empty); ``alpha`` and the loss-Hessian constants stay those of the GLOBAL batch.
        ``shard = (index, count)`` of the batch shard (default: rank / world)."""
        rank, world = cdist.rank_world() if shard is None else shard
        dev = self.device
        for bi, (X, y) in enumerate(self._loop_over_data(desc=desc, to_device=(world == 1), stage=True)):
            alpha = self._get_normalization_factor(X, y)
            if world == 1:
                yield X, y, alpha, None
                continue
            Xs, ys, scales = cdist.shard_batch(X, y, rank, world, self._loss_func, self._engine)
            if Xs is not None and Xs.device == dev:
                # device-resident data: hand out the SAME slice objects product after product (the engine keys its
                # converted copies on the tensor object)
                cache = self.__dict__.setdefault("_slice_cache", {})
                hit = cache.get(bi)
                if hit is not None and hit[0]() is X and hit[1] == Xs._curv_shard:
                    Xs, ys = hit[2], hit[3]
                else:
                    cache[bi] = (weakref.ref(X), Xs._curv_shard, Xs, ys)
            elif Xs is not None:
                shard_info = Xs._curv_shard
                Xs, ys = self._upload(Xs, (bi, "Xs")), self._upload(ys, (bi, "ys"))
                Xs._curv_shard = shard_info
            else:
                self._on_empty_shard(X, y)
            yield Xs, ys, alpha, scales

    def _on_empty_shard(self, X, y) -> None:
        """This rank holds no sample of the mini-batch (fewer samples than ranks)."""

    def _param_buckets(self, K: int, bucket_bytes: int, first_bytes: int | None = None):
        """Row ranges ``(lo, hi, [param indices])`` of the flat ``[P, K]`` matrices made of whole parameters, about
        ``bucket_bytes`` each (the first one at most ``first_bytes``: it is the piece that cannot overlap)."""
        sizes = [p.numel() for p in self._params.values()]
        buckets, lo, plist, acc = [], 0, [], 0
        for i, n in enumerate(sizes):
            plist.append(i)
            acc += n
            limit = min(bucket_bytes, first_bytes) if (not buckets and first_bytes) else bucket_bytes
            if acc * K * 4 >= limit or i == len(sizes) - 1:
                buckets.append((lo, lo + acc, plist))
                lo, plist, acc = lo + acc, [], 0
        return buckets

    #: with batch sharding on: sum the [P, K] result in parameter buckets on a side stream, each bucket as soon as
    #: the backward sweep has finished its rows (it finishes the LAST parameters first), instead of one blocking
    #: all-reduce after the last kernel
    #: (measured at 8 x B200, round 2: 15.9 ms with the overlap vs 14.8 ms with one blocking all-reduce after a
    #: CUDA-graph replay -- the persistent contraction kernels own every SM with a static tile schedule, so NCCL's
    #: CTAs delay whole kernels, and the streaming entry point cannot be graph-replayed.  Off by default.)
    OVERLAP_ALLREDUCE = os.environ.get("CURV_OVERLAP_ALLREDUCE", "0") != "0"

    def _product(self, V: Tensor) -> Tensor:
        """``[P, K]`` fp32 -> ``[P, K]`` fp32: the loop over mini-batches (``_torch_base.py:937-944``)."""
        from .engine import MAX_COLUMNS_PER_SWEEP

        out = torch.zeros_like(V)
        rank, world = cdist.rank_world()
        K = V.shape[1]
        # 2-d sharding (batch x columns) from 4 ranks on, see dist.grid
        n_bg, n_cg = cdist.grid(world, K)
        cols = None if n_cg == 1 else ((rank % n_cg) * (K // n_cg), K // n_cg)
        overlap = (world > 1 and self.OVERLAP_ALLREDUCE and V.device.type == "cuda" and K <= MAX_COLUMNS_PER_SWEEP
                   and cols is None)
        batches = iter(self._local_batches("_matmat", None if world == 1 else (rank // n_cg, n_bg)))
        cur = next(batches, None)
        reduced = False
        while cur is not None:
            nxt = next(batches, None)
            X, y, alpha, scales = cur
            if world > 1:  # data-parallel shard of this mini-batch, weights stay global
                if overlap and nxt is None:
                    self._last_batch_with_overlapped_all_reduce(X, y, V, out, alpha, scales)
                    reduced = True
                elif X is not None:
                    self._batch_call_sharded(X, y, V, out, alpha, scales, **({} if cols is None else {"cols": cols}))
            else:
                if not isinstance(X, Tensor):
                    raise NotImplementedError("The B200 engine needs tensor inputs X.")
                self._batch_call(X, y, V, out, alpha)
            cur = nxt
        if world > 1 and not reduced:
            cdist.all_reduce_sum(out)
        return out

    def _last_batch_with_overlapped_all_reduce(self, X, y, V, out, alpha, scales, bucket_bytes: int = 64 << 20):
        dev = V.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream(dev)
        comm = self._comm_stream
        # buckets in parameter (= forward) order; the sweep completes them back to front.  Every rank builds the same
        # list, so the sequence of collectives matches even for a rank without samples.
        buckets = self._param_buckets(V.shape[1], bucket_bytes, first_bytes=8 << 20)
        n = len(self._params)
        out_done, events = [None] * n, []
        for blo, bhi, ps in buckets:
            e = torch.cuda.Event()
            e.record(main)  # materialises the handle; the engine re-records it when the bucket's rows are final
            events.append(e)
            for i in ps:
                out_done[i] = e
        if X is not None:
            self._batch_call_sharded(X, y, V, out, alpha, scales, out_done=out_done)
        tail = torch.cuda.Event()
        tail.record(main)
        with torch.cuda.stream(comm):
            for (blo, bhi, _), e in zip(reversed(buckets), reversed(events)):
                comm.wait_event(e if X is not None else tail)
                cdist.all_reduce_sum(out[blo:bhi])
        main.wait_stream(comm)
        out.record_stream(comm)

    @staticmethod
    def _shard_kw(out_done, cols) -> dict:
        kw = {} if out_done is None else {"out_done": out_done}
        if cols is not None:
            kw["cols"] = cols
        return kw

    def _batch_call_sharded(self, X, y, V, out, alpha, scale, out_done=None, cols=None):
        self._engine.matmat_batch(self.KIND, X, y, V, out, alpha, scale=scale[0], **self._shard_kw(out_done, cols))

    # ---- host-resident operands: pipelined upload / download ----------------------------------------
    #: a kind whose mini-batch product is one engine call with (X, y, V, out) only may stream
    _STREAMABLE = True

    def matmat_pinned(self, V: Tensor, out: Tensor | None = None, bucket_bytes: int = 48 << 20) -> Tensor:
        """``A @ V`` for a host-resident flat ``[P, K]`` matrix, returned on the host.

        This is what the reference's SciPy bridge does per product (``_torch_base.py:560-592``: matrix to the
        device, multiply, result back to the host), but pipelined: V is uploaded in parameter buckets on a copy
        stream while the forward sweep already runs on the first layers (the engine waits per parameter for its
        columns), and the result is downloaded bucket by bucket while the backward sweep is still producing the
        earlier parameters (it finishes the last parameters first).  Use pinned host tensors for truly
        asynchronous copies; the returned tensor is complete once the current stream has been synchronised.
        """
        dev = self.device
        sizes = [p.numel() for p in self._params.values()]
        Pn = sum(sizes)
        if V.ndim != 2 or V.shape[0] != Pn:
            raise ValueError(f"Expected a flat [P={Pn}, K] matrix, got {tuple(V.shape)}.")
        K = V.shape[1]
        from .engine import MAX_COLUMNS_PER_SWEEP

        rank, world = cdist.rank_world()
        if (dev.type != "cuda" or not self._STREAMABLE or K > MAX_COLUMNS_PER_SWEEP
                or getattr(self, "_mc_samples", 0) > 0):
            # general path (sharded / sampled / wide products): whole-matrix copies around the resident product,
            # asynchronous when the host tensors are pinned
            res = self @ V.to(dev, non_blocking=True)
            if out is None:
                out = torch.empty(res.shape, dtype=res.dtype, pin_memory=(dev.type == "cuda"))
            out.copy_(res, non_blocking=True)
            return out
        if world > 1:
            return self._matmat_pinned_sharded(V, out, rank, world)
        # bf16 operators keep V / the result in bf16 on the host (half the PCIe bytes); the engine's fp32 matrices
        # are filled / drained by conversions on the copy streams
        lowp = V.dtype == torch.bfloat16
        if not lowp:
            V = V.to(torch.float32)
        if out is None:
            out = torch.empty(Pn, K, dtype=V.dtype, pin_memory=True)
        # buckets of whole parameters, about bucket_bytes each; the first one is small: it gates the start of the
        # forward sweep on the way in and is the only download that cannot overlap the backward sweep
        buckets, lo, plist, acc = [], 0, [], 0
        for i, n in enumerate(sizes):
            plist.append(i)
            acc += n
            limit = min(bucket_bytes, 4 << 20) if not buckets else bucket_bytes
            if acc * K * 4 >= limit or i == len(sizes) - 1:
                buckets.append((lo, lo + acc, plist))
                lo, plist, acc = lo + acc, [], 0
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._copy_streams
        Vd = torch.empty(Pn, K, dtype=torch.float32, device=dev)
        outd = torch.zeros(Pn, K, dtype=torch.float32, device=dev)
        Vlow = torch.empty(Pn, K, dtype=V.dtype, device=dev) if lowp else None
        outlow = torch.empty(Pn, K, dtype=out.dtype, device=dev) if out.dtype != torch.float32 else None
        v_ready, out_done, bucket_events = [None] * len(sizes), [None] * len(sizes), []
        # the first mini-batch goes up BEFORE V: host-to-device copies share one DMA queue, and the forward
        # sweep needs X first
        batches = iter(self._loop_over_data(desc="matmat_pinned"))
        cur = next(batches, None)
        s_in.wait_stream(main)
        with torch.cuda.stream(s_in):
            for blo, bhi, ps in buckets:
                if lowp:
                    Vlow[blo:bhi].copy_(V[blo:bhi], non_blocking=True)
                    Vd[blo:bhi].copy_(Vlow[blo:bhi])
                else:
                    Vd[blo:bhi].copy_(V[blo:bhi], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                for i in ps:
                    v_ready[i] = e
        for blo, bhi, ps in buckets:
            e = torch.cuda.Event()
            e.record(main)  # materialises the handle; the engine records it again when the rows are final
            bucket_events.append(e)
            for i in ps:
                out_done[i] = e
        first = True
        while cur is not None:
            nxt = next(batches, None)
            X, y = cur
            if not isinstance(X, Tensor):
                raise NotImplementedError("The B200 engine needs tensor inputs X.")
            alpha = self._get_normalization_factor(X, y)
            if world > 1:  # data-parallel shard (see _matmat); the result is only final after the all-reduce
                Xs, ys, scale = cdist.shard_batch(X, y, rank, world, self._loss_func, self._engine)
                if Xs is not None:
                    self._engine.matmat_batch(self.KIND, Xs, ys, Vd, outd, alpha, scale=scale[0],
                                              v_ready=v_ready if first else None)
                else:
                    main.wait_stream(s_in)
            else:
                self._engine.matmat_batch(self.KIND, X, y, Vd, outd, alpha, v_ready=v_ready if first else None,
                                          out_done=out_done if nxt is None else None)
            first, cur = False, nxt
        if world > 1:
            main.wait_stream(s_in)
            cdist.all_reduce_sum(outd)
            out.copy_(outd if outlow is None else outlow.copy_(outd), non_blocking=True)
        else:
            with torch.cuda.stream(s_out):
                for (blo, bhi, _), e in zip(reversed(buckets), reversed(bucket_events)):
                    s_out.wait_event(e)
                    if outlow is not None:
                        outlow[blo:bhi].copy_(outd[blo:bhi])
                        out[blo:bhi].copy_(outlow[blo:bhi], non_blocking=True)
                    else:
                        out[blo:bhi].copy_(outd[blo:bhi], non_blocking=True)
            main.wait_stream(s_out)
        main.wait_stream(s_in)
        Vd.record_stream(s_in)
        outd.record_stream(s_out)
        if Vlow is not None:
            Vlow.record_stream(s_in)
        if outlow is not None:
            outlow.record_stream(s_out)
        return out

    def _matmat_pinned_sharded(self, V: Tensor, out: Tensor | None, rank: int, world: int) -> Tensor:
        """Host-operand product on several ranks of one box: the host matrices are moved ONCE in total, not once per
        rank.  Rank r uploads rows ``row_block(r)`` of V (1/world of the bytes) and NCCL all-gathers the blocks over
        NVLink; after the sweeps a reduce-scatter leaves rank r with rows ``row_block(r)`` of the summed result,
        which it downloads into ``out[row_block(r)]``.  Hand every rank a view of the SAME host buffers
        (:func:`curvlinops_b200.dist.shared_pinned_tensor`) and ``out`` holds the whole result once all ranks have
        synchronised; with private buffers each rank's ``out`` holds its row block only (the other rows are left
        untouched)."""
        dev = self.device
        Pn, K = V.shape
        chunk = -(-Pn // world)
        lo, hi = min(Pn, rank * chunk), min(Pn, (rank + 1) * chunk)
        if out is None:
            out = torch.zeros(Pn, K, dtype=V.dtype, pin_memory=True)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, _ = self._copy_streams
        Vd = torch.empty(world * chunk, K, dtype=torch.float32, device=dev)
        outd = torch.zeros(world * chunk, K, dtype=torch.float32, device=dev)
        part = torch.zeros(chunk, K, dtype=torch.float32, device=dev)
        s_in.wait_stream(main)
        with torch.cuda.stream(s_in):  # upload + all-gather on the side stream: X goes up on the main stream meanwhile
            if hi > lo:
                stage = torch.empty(hi - lo, K, dtype=V.dtype, device=dev)
                stage.copy_(V[lo:hi], non_blocking=True)
                part[: hi - lo].copy_(stage)
            cdist.all_gather_rows(Vd, part)
        n_bg, n_cg = cdist.grid(world, K)
        cols = None if n_cg == 1 else ((rank % n_cg) * (K // n_cg), K // n_cg)
        batches = list(self._local_batches("matmat_pinned", (rank // n_cg, n_bg)))
        main.wait_stream(s_in)
        for X, y, alpha, scales in batches:
            if X is not None:
                self._batch_call_sharded(X, y, Vd[:Pn], outd[:Pn], alpha, scales, cols=cols)
        res = torch.empty(chunk, K, dtype=torch.float32, device=dev)
        cdist.reduce_scatter_rows(res, outd)
        if hi > lo:
            out[lo:hi].copy_(res[: hi - lo] if out.dtype == torch.float32 else res[: hi - lo].to(out.dtype),
                             non_blocking=True)
        for t in (Vd, part):
            t.record_stream(s_in)
        return out

    def __getstate__(self):
        # compiled programs / workspaces are per-process handles: rebuild lazily after unpickling
        st = self.__dict__.copy()
        st["_engine"] = None
        st["_copy_streams"] = None
        st.pop("_mc_cache", None)
        st.pop("_upload_stage", None)
        st.pop("_slice_cache", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._engine = Engine(self._model_func, self._loss_func, self._params)


class GGNLinearOperator(CurvatureLinearOperator):
    r"""Generalized Gauss-Newton matrix :math:`c\sum_n J_n^\top \nabla^2_{f_n}\ell\, J_n` of an empirical
    risk, or its Monte-Carlo (Fisher) approximation when ``mc_samples > 0``
    (reference ``curvlinops/ggn.py:171-366``)."""

    SELF_ADJOINT = True
    KIND = capi.KIND_GGN
    MC_SUPPORTED_LOSSES = (MSELoss, CrossEntropyLoss, BCEWithLogitsLoss)

    def __init__(self, model_func, loss_func, params, data, progressbar=False, check_deterministic=True,
                 num_data=None, batch_size_fn=None, mc_samples: int = 0, seed: int = 2147483647):
        self._mc_samples = mc_samples
        self._mc_grad_override = None  # tests: list of [B, M, C] tensors, one per mini-batch
        if mc_samples > 0:
            if not isinstance(loss_func, self.MC_SUPPORTED_LOSSES):
                raise NotImplementedError(
                    f"MC-GGN requires loss in {self.MC_SUPPORTED_LOSSES}. Got: {loss_func}."
                )
            self.FIXED_DATA_ORDER = True
            self._seed = seed
        super().__init__(model_func, loss_func, params, data, progressbar=progressbar,
                         check_deterministic=check_deterministic, num_data=num_data,
                         batch_size_fn=batch_size_fn)

    def _product(self, V):
        if self._mc_samples > 0:
            self._batch_index = 0
            dev = self.device
            with torch.random.fork_rng(devices=[dev] if dev.type == "cuda" else []):
                torch.manual_seed(self._seed)  # same stream as the reference (ggn.py:337-341)
                return super()._product(V)
        return super()._product(V)

    def _on_empty_shard(self, X, y) -> None:
        if self._mc_samples > 0:  # consume the variates of this mini-batch: the stream stays aligned across ranks
            B, M, dev = X.shape[0], self._mc_samples, self.device
            n = B * M if isinstance(self._loss_func, CrossEntropyLoss) else B * M * y.shape[-1]
            (torch.randn if isinstance(self._loss_func, MSELoss) else torch.rand)(n, device=dev)

    def _mc_scale(self, batch: int) -> float:
        return 1.0 / batch if self._loss_func.reduction == "mean" else 1.0

    def _mc_grad(self, X, shard=None):
        """Would-be gradients of one mini-batch.  Every product re-seeds the generator (same stream as the reference,
        ``ggn.py:337-341``), so as long as the parameters and the mini-batch are unchanged the draws of batch i are the
        same tensor every time: they are kept per (batch position, data, parameter versions) and a hit only consumes
        the variates the draw would have used, which keeps the stream aligned for the batches that follow.  That saves
        the primal forward pass + softmax + sampling per product (about 14 % of a ViT-B/16 MC-GGN product)."""
        key = (self._batch_index, self._seed, self._mc_samples, X.data_ptr(), X._version, tuple(X.shape), shard,
               tuple((p.data_ptr(), p._version) for p in self._params.values()))
        cache = self.__dict__.setdefault("_mc_cache", {})
        hit = cache.get(self._batch_index)
        if hit is not None and hit[0] == key:
            lo, hi, Bg = (0, X.shape[0], X.shape[0]) if shard is None else shard
            lf, M, C = self._loss_func, self._mc_samples, hit[1].shape[-1]
            n = Bg * M if isinstance(lf, CrossEntropyLoss) else Bg * M * C
            (torch.randn if isinstance(lf, MSELoss) else torch.rand)(n, device=X.device, dtype=hit[1].dtype)
            return hit[1]
        g = self._engine.mc_grad_outputs(X, self._mc_samples, shard=shard)
        if hit is not None and hit[1].shape == g.shape and hit[1].dtype == g.dtype:
            hit[1].copy_(g)  # same address as before: the product's CUDA graph (keyed on it) stays valid
            g = hit[1]
        if hit is not None or len(cache) < 64:
            cache[self._batch_index] = (key, g)
        return g

    def _batch_call(self, X, y, V, out, alpha):
        if self._mc_samples == 0:
            return super()._batch_call(X, y, V, out, alpha)
        if self._mc_grad_override is not None:
            g = self._mc_grad_override[self._batch_index].to(X.device)
        else:
            g = self._mc_grad(X)
        self._batch_index += 1
        self._engine.matmat_batch(capi.KIND_GGN_MC, X, y, V, out, alpha, mc_grad=g,
                                  scale=self._mc_scale(X.shape[0]))

    def _batch_call_sharded(self, X, y, V, out, alpha, scale, out_done=None, cols=None):
        if self._mc_samples == 0:
            return super()._batch_call_sharded(X, y, V, out, alpha, scale, out_done=out_done, cols=cols)
        g = self._mc_grad(X, shard=getattr(X, "_curv_shard", None))
        self._batch_index += 1
        self._engine.matmat_batch(capi.KIND_GGN_MC, X, y, V, out, alpha, mc_grad=g, scale=scale[1],
                                  **self._shard_kw(out_done, cols))


class EFLinearOperator(CurvatureLinearOperator):
    r"""Uncentered gradient covariance ('empirical Fisher') :math:`c\sum_n \nabla_\theta\ell_n
    \nabla_\theta\ell_n^\top` (reference ``curvlinops/gradient_moments.py:90-151``).

    The reference evaluates it as the GGN of the pseudo-loss :math:`\frac{1}{2c}\sum_n\langle f_n, g_n\rangle^2`
    with :math:`g_n = \partial\ell_n/\partial f_n` (``gradient_moments.py:48-86``); here that is the engine's two
    sweeps with the rank-one per-sample loss term :math:`g_n g_n^\top / c` (the Monte-Carlo kind with one
    deterministic 'sample'): one primal forward for :math:`g_n`, then forward + Jv and backward + J^T."""

    SELF_ADJOINT = True
    KIND = capi.KIND_GGN_MC
    _STREAMABLE = False  # needs the per-sample loss gradients as an extra operand
    SUPPORTED_LOSSES = (MSELoss, CrossEntropyLoss, BCEWithLogitsLoss)

    def __init__(self, model_func, loss_func, params, data, progressbar=False, check_deterministic=True,
                 num_data=None, batch_size_fn=None):
        if not isinstance(loss_func, self.SUPPORTED_LOSSES):
            raise NotImplementedError(f"Loss must be one of {self.SUPPORTED_LOSSES}. Got: {loss_func}.")
        super().__init__(model_func, loss_func, params, data, progressbar=progressbar,
                         check_deterministic=check_deterministic, num_data=num_data,
                         batch_size_fn=batch_size_fn)

    def _grad_outputs(self, X, y) -> Tensor:
        """``[B, 1, C]`` per-sample gradients of the unreduced loss w.r.t. the prediction."""
        f = self._engine.predict(X)
        lf = self._loss_func
        if isinstance(lf, CrossEntropyLoss):
            g = torch.softmax(f, dim=1) - torch.nn.functional.one_hot(y, f.shape[1]).to(f.dtype)
        elif isinstance(lf, MSELoss):
            g = 2.0 * (f - y.to(f.dtype))
        else:
            g = torch.sigmoid(f) - y.to(f.dtype)
        return g.unsqueeze(1)

    def _batch_call(self, X, y, V, out, alpha):
        from .engine import loss_scale

        g = self._grad_outputs(X, y)
        self._engine.matmat_batch(capi.KIND_GGN_MC, X, y, V, out, alpha, mc_grad=g,
                                  scale=loss_scale(self._loss_func, X.shape[0], g.shape[-1]))

    def _batch_call_sharded(self, X, y, V, out, alpha, scale, out_done=None, cols=None):
        self._engine.matmat_batch(capi.KIND_GGN_MC, X, y, V, out, alpha, mc_grad=self._grad_outputs(X, y),
                                  scale=scale[0], **self._shard_kw(out_done, cols))


class HessianLinearOperator(CurvatureLinearOperator):
    r"""Hessian :math:`c\sum_n \nabla^2_\theta \ell(f_\theta(x_n), y_n)` of an empirical risk
    (reference ``curvlinops/hessian.py:72-145``), applied with a hand-written R-op
    (forward + tangent sweep, plain backward, R-backward) instead of forward-over-reverse autograd."""

    SELF_ADJOINT = True
    KIND = capi.KIND_HESSIAN
