"""Host-side driver of the CUDA engine: program cache, workspace, per-mini-batch calls.

This is the piece that replaces ``self._mp(X, y, M)`` -- the vmapped ``torch.func`` product of the
reference (``curvlinops/_torch_base.py:946-989``) -- by calls into ``libcurvb200.so``.
"""

from __future__ import annotations

import ctypes as C
import itertools
import math
import os
import weakref

import torch
from torch import Tensor
from torch.nn import BCEWithLogitsLoss, CrossEntropyLoss, MSELoss

from . import _capi as capi
from .capture import LayerProgram, capture

#: number of columns processed per sweep (tangent slots held in the workspace)
MAX_COLUMNS_PER_SWEEP = 8

#: Replay the launches of a mini-batch product as one CUDA graph.  A product is a fixed sequence of ~450 kernel
#: launches (plus tensor-map encodes) whose arguments depend only on pointers and shapes, so the second call
#: with the same (program, data, parameter, column-count) key is captured and later calls replay it: the
#: launch gaps (a few % of a ResNet-18 step, most of an MLP step) disappear.
#: Set to False to launch eagerly (per-launch profiling does so automatically).
CUDA_GRAPHS = os.environ.get("CURV_CUDA_GRAPHS", "1") != "0"
_MAX_GRAPHS = 8


def _loss_code(loss_func) -> int:
    if isinstance(loss_func, CrossEntropyLoss):
        if loss_func.weight is not None or loss_func.label_smoothing != 0.0 or loss_func.ignore_index >= 0:
            raise NotImplementedError("CrossEntropyLoss options (weight/label_smoothing/ignore_index).")
        return capi.LOSS_CE
    if isinstance(loss_func, MSELoss):
        return capi.LOSS_MSE
    if isinstance(loss_func, BCEWithLogitsLoss):
        if loss_func.weight is not None or loss_func.pos_weight is not None:
            raise NotImplementedError("BCEWithLogitsLoss weights.")
        return capi.LOSS_BCE
    raise NotImplementedError(
        f"Loss {loss_func} is not supported by the B200 engine (MSELoss, CrossEntropyLoss, BCEWithLogitsLoss)."
    )


def loss_scale(loss_func, batch: int, classes: int) -> float:
    """Reduction constant of the mini-batch loss Hessian (``ggn_utils.py:116-124`` times the batch mean)."""
    if loss_func.reduction == "sum":
        return 1.0
    if loss_func.reduction != "mean":
        raise ValueError(f"Unsupported reduction {loss_func.reduction!r}.")
    return 1.0 / batch if isinstance(loss_func, CrossEntropyLoss) else 1.0 / (batch * classes)


_program_serial = itertools.count()

#: Layer programs are a function of (model, parameter selection, input shape / dtype) only: operators built
#: repeatedly on the same module (a KFAC operator per optimisation step, the three operators of a comparison) share
#: one trace.  Keyed on the identity AND the version counters of every tensor the module owns, so a module whose
#: buffers / parameters were modified in place is traced again (constants are baked into the program).
_CAPTURE_CACHE: dict = {}
_CAPTURE_CACHE_MAX = 16


def _cached_capture(model_func, params: dict[str, Tensor], X: Tensor, fuse_relu: bool) -> LayerProgram:
    module = getattr(model_func, "module", None)
    if not isinstance(module, torch.nn.Module):
        return capture(model_func, params, X, fuse_relu=fuse_relu)
    state = tuple((id(t), t._version) for t in itertools.chain(module.parameters(), module.buffers()))
    key = (id(module), module.training, tuple(params), tuple(id(p) for p in params.values()), tuple(X.shape), X.dtype,
           str(X.device), fuse_relu, state)
    hit = _CAPTURE_CACHE.get(key)
    if hit is not None and hit[0]() is module:
        return hit[1]
    lp = capture(model_func, params, X, fuse_relu=fuse_relu)
    if len(_CAPTURE_CACHE) >= _CAPTURE_CACHE_MAX:
        _CAPTURE_CACHE.pop(next(iter(_CAPTURE_CACHE)))
    _CAPTURE_CACHE[key] = (weakref.ref(module), lp)
    return lp


class CompiledProgram:
    """A layer program planned for one (input shape, kmax, hessian) combination."""

    def __init__(self, model_func, params: dict[str, Tensor], X: Tensor, kmax: int, hessian: bool | int):
        # `hessian` is the flag word of curv_program_create: 1 = R-op storage, 2 = KFAC scratch, 4 = bf16 arithmetic
        # traced in the parameters' dtype (the engine itself is handed fp32 copies of bf16 data)
        pdt = next(iter(params.values())).dtype if params else X.dtype
        self.lp: LayerProgram = _cached_capture(model_func, params, X if X.dtype == pdt else X.to(pdt),
                                                fuse_relu=not (int(hessian) & 1))
        self.kmax = kmax
        self.serial = next(_program_serial)  # CUDA-graph cache key (an id() could be reused after a rebuild)
        self.batch = X.shape[0]
        self.device = X.device
        lp = self.lp
        vals = (capi.ValueDesc * len(lp.values))(*[capi.ValueDesc(c, h, w, int(t)) for c, h, w, t in lp.values])
        nodes = (capi.NodeDesc * len(lp.nodes))(*[capi.NodeDesc(**n) for n in lp.nodes])
        offs, o = [], 0
        for p in params.values():
            offs.append(o)
            o += p.numel()
        self.P = o
        pds = (capi.ParamDesc * max(1, len(offs)))(*[capi.ParamDesc(p.numel(), off)
                                                     for p, off in zip(params.values(), offs)])
        handle = C.c_void_p()
        capi.check(capi.lib().curv_program_create(vals, len(lp.values), nodes, len(lp.nodes), pds, len(offs),
                                                  self.batch, kmax, int(hessian), C.byref(handle)))
        self.handle = handle
        self.ws_bytes = capi.lib().curv_program_workspace_bytes(handle)
        self.consts = [t.to(self.device) for t in lp.consts]
        self.const_ptrs = capi.ptr_array([t.data_ptr() for t in self.consts])
        self.out_value = lp.nodes[-1]["out"]
        self.classes = lp.out_features

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                capi.lib().curv_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def engine_input(self, X: Tensor) -> Tensor:
        """The network input as the library reads it: NCHW / [B, C]; a token sequence [B, T, D] is handed over as
        [B, D, 1, T] (one transposed copy per call: the library's input kernel expects channels-first)."""
        return X.transpose(1, 2).contiguous() if self.lp.tokens_input else X

    def value_view(self, ws: Tensor, value_id: int, nslots: int) -> Tensor:
        """``[nslots, B, H, W, Cp]`` view of a value's storage inside the workspace (tests / MC draw)."""
        off, sb, cp = C.c_size_t(), C.c_size_t(), C.c_int()
        capi.check(capi.lib().curv_program_value_layout(self.handle, value_id, C.byref(off), C.byref(sb),
                                                        C.byref(cp)))
        c, h, w, _ = self.lp.values[value_id]
        n = sb.value // 4
        flat = ws.view(torch.float32)[off.value // 4: off.value // 4 + n * nslots]
        return flat.view(nslots, self.batch, h, w, cp.value)


class Engine:
    """Runs curvature matmats of one operator through the CUDA library."""

    def __init__(self, model_func, loss_func, params: dict[str, Tensor]):
        self.model_func = model_func
        self.loss_func = loss_func
        self.params = params
        self._programs: dict = {}
        self._ws: Tensor | None = None
        self._graphs: dict = {}    # baked-in pointers except V / out -> {"first", "exact", "staged"} (see _matmat_batch)
        self._x32: list = []       # bf16 operators: [(data_ptr, version, shape, fp32 copy of the mini-batch input)]
        self._p32: dict = {}       # bf16 operators: name -> (data_ptr, version, fp32 copy of the parameter)
        capi.lib()  # fail loudly if the CUDA library has not been built

    # -- plumbing ------------------------------------------------------------------------------
    def _check_supported(self):
        dtypes = {p.dtype for p in self.params.values()}
        if len(dtypes) > 1:
            raise RuntimeError(f"Could not infer data type. Parameters have types {dtypes}.")
        for n, p in self.params.items():
            if p.dtype not in (torch.float32, torch.bfloat16):
                raise NotImplementedError(
                    f"The B200 engine computes in float32 or bfloat16; parameter {n!r} has dtype {p.dtype}."
                )
            if p.device.type != "cuda":
                raise RuntimeError(
                    "curvlinops_b200 runs on CUDA devices only (no CPU fallback); parameter "
                    f"{n!r} lives on {p.device}."
                )

    @property
    def bf16(self) -> bool:
        """bf16 operator (``_torch_base.py:586-589``: bf16 in, bf16 out): the contractions run as ONE bf16
        ``tcgen05.mma`` per product with fp32 accumulation (program flag 4) instead of the three fp16 hi/lo MMAs of
        the fp32-grade path; everything between the contractions stays fp32."""
        return next(iter(self.params.values())).dtype == torch.bfloat16

    def program(self, X: Tensor, kmax: int, hessian: bool | int) -> CompiledProgram:
        hessian = int(hessian) | (4 if self.bf16 else 0)
        key = (tuple(X.shape), hessian)
        prog = self._programs.get(key)
        if prog is None or prog.kmax < kmax:
            if prog is not None:  # graphs captured for the replaced program bake in its plan: drop them
                self._graphs = {k: v for k, v in self._graphs.items() if k[0] != prog.serial}
            prog = CompiledProgram(self.model_func, self.params, X, kmax, hessian)
            self._programs[key] = prog
        return prog

    def workspace(self, nbytes: int, device) -> Tensor:
        if self._ws is None or self._ws.numel() * 4 < nbytes or self._ws.device != device:
            self._ws = None
            self._graphs.clear()  # captured graphs bake in the old workspace address
            self._ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        return self._ws

    def _param_ptrs(self):
        if self.bf16:  # the library reads fp32 parameters: exact upcasts, refreshed when a parameter changes
            ps = []
            for n, p in self.params.items():
                hit = self._p32.get(n)
                if hit is None or hit[0] != p.data_ptr() or hit[1] != p._version:
                    hit = (p.data_ptr(), p._version, p.detach().to(torch.float32).contiguous())
                    self._p32[n] = hit
                ps.append(hit[2])
        else:
            ps = [p if p.is_contiguous() else p.contiguous() for p in self.params.values()]
        return ps, capi.ptr_array([p.data_ptr() for p in ps])

    def _input32(self, X: Tensor) -> Tensor:
        """fp32 contiguous view of a mini-batch input.  Low-precision / strided inputs are converted once per
        (storage, version) and kept (4 most recent): the copy keeps its address from product to product, which the
        CUDA-graph key needs, and an eigensolver that revisits the same batches does not pay for it again."""
        if X.dtype == torch.float32 and X.is_contiguous():
            return X
        sig = (tuple(X.shape), X.dtype, X.device)
        for i, hit in enumerate(self._x32):
            # identity of the tensor OBJECT, not its address: a data loader's fresh batch may land on the address of a
            # freed one with the same version counter
            if hit[0]() is X and hit[2] == sig:
                if hit[1] != X._version:  # same buffer, new contents (a host mini-batch uploaded again): convert in
                    hit[3].copy_(X)       # place, the fp32 copy keeps its address
                    self._x32[i] = (hit[0], X._version, sig, hit[3])
                return hit[3]
        X32 = X.to(torch.float32).contiguous()
        if X.device.type == "cuda":
            self._x32 = [h for h in self._x32 if h[0]() is not None][-3:] + [(weakref.ref(X), X._version, sig, X32)]
        return X32

    # -- the hot call ----------------------------------------------------------------------------
    def matmat_batch(self, kind: int, X: Tensor, y: Tensor | None, V: Tensor, out: Tensor, alpha: float,
                     mc_grad: Tensor | None = None, scale: float | None = None,
                     v_ready: list | None = None, out_done: list | None = None,
                     cols: tuple[int, int] | None = None) -> None:
        """``out += alpha * (mini-batch matrix) @ V`` with ``V``/``out`` flat ``[P, K]`` fp32 on device.

        ``v_ready`` / ``out_done``: optional per-parameter ``torch.cuda.Event`` lists (entries may be ``None``) for
        the streaming entry point ``curv_matmat_batch_sync`` (see ``include/curvb200.h``); needs ``K`` columns
        that fit one sweep.  ``cols = (first, count)``: only these columns of ``V`` / ``out`` are processed (2-d
        sharding over ranks: batch x columns)."""
        self._check_supported()
        with torch.cuda.device(X.device if X.device.type == "cuda" else V.device):
            self._matmat_batch(kind, X, y, V, out, alpha, mc_grad, scale, v_ready, out_done, cols)

    def _matmat_batch(self, kind, X, y, V, out, alpha, mc_grad, scale, v_ready, out_done, cols=None) -> None:
        K = V.shape[-1]
        c0, cn = (0, K) if cols is None else cols
        kc = min(cn, MAX_COLUMNS_PER_SWEEP)
        X = self._input32(X)
        prog = self.program(X, kc, kind == capi.KIND_HESSIAN)
        X = prog.engine_input(X)
        ws = self.workspace(prog.ws_bytes, X.device)
        keep, pptrs = self._param_ptrs()
        loss = 0
        if kind in (capi.KIND_GGN, capi.KIND_GGN_MC, capi.KIND_HESSIAN):
            loss = _loss_code(self.loss_func)
            if scale is None:
                scale = loss_scale(self.loss_func, X.shape[0], prog.classes)
            if loss == capi.LOSS_CE:
                if y.dtype != torch.int64 or y.ndim != 1:
                    raise NotImplementedError("CrossEntropyLoss needs int64 class labels of shape [batch].")
            else:
                y = y.to(torch.float32)
            y = y.contiguous()
        stream = torch.cuda.current_stream(X.device).cuda_stream
        M = 0
        if mc_grad is not None:
            mc_grad = mc_grad.to(torch.float32).contiguous()
            M = mc_grad.shape[1]
        if v_ready is not None or out_done is not None:
            if K > kc:
                raise ValueError(f"Streaming products take at most {kc} columns per call.")
            n = len(keep)
            ev_in = capi.ptr_array([e.cuda_event if e is not None else None for e in (v_ready or [None] * n)])
            ev_out = capi.ptr_array([e.cuda_event if e is not None else None for e in (out_done or [None] * n)])
            capi.check(capi.lib().curv_matmat_batch_sync(
                prog.handle, kind, loss, pptrs, prog.const_ptrs, X.data_ptr(),
                0 if y is None else y.data_ptr(), 0 if mc_grad is None else mc_grad.data_ptr(), M,
                V.data_ptr(), out.data_ptr(), K, K, 0, float(scale or 1.0), float(alpha),
                ws.data_ptr(), ws.numel() * 4, stream,
                ev_in if v_ready is not None else None, ev_out if out_done is not None else None))
            del keep
            return

        def launch(Vt: Tensor, outt: Tensor, strm: int) -> None:
            for k0 in range(c0, c0 + cn, kc):
                kk = min(kc, c0 + cn - k0)
                capi.check(capi.lib().curv_matmat_batch(
                    prog.handle, kind, loss, pptrs, prog.const_ptrs, X.data_ptr(),
                    0 if y is None else y.data_ptr(), 0 if mc_grad is None else mc_grad.data_ptr(), M,
                    Vt.data_ptr(), outt.data_ptr(), kk, K, k0, float(scale or 1.0), float(alpha),
                    ws.data_ptr(), ws.numel() * 4, strm))

        cfg = capi.lib().curv_launch_config()
        if not CUDA_GRAPHS or (cfg >> 16) or torch.cuda.is_current_stream_capturing():
            launch(V, out, stream)
            del keep
            return
        # Two kinds of captured graph per product (`base` = everything baked in except the V / out addresses).
        # EXACT: in steady state the caching allocator hands a caller that builds V / out per product the same blocks
        # again; when the second call repeats the first call's addresses the graph is captured on them and replays
        # need no copies.  STAGED: a caller whose V / out addresses move (an eigensolver allocating vectors of its own
        # between products) gets ONE graph on engine-owned V / out buffers, replayed between two device copies
        # ([P, K] floats each way) - never a capture per address pair.
        base = (prog.serial, kind, loss, K, c0, cn, float(scale or 1.0), float(alpha), X.data_ptr(),
                0 if y is None else y.data_ptr(), 0 if mc_grad is None else mc_grad.data_ptr(), M,
                tuple(p.data_ptr() for p in keep), ws.data_ptr(), cfg)
        addr = (V.data_ptr(), out.data_ptr())

        def captured(Vt, outt):
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(X.device)
            n0 = capi.lib().curv_launch_count()
            with torch.cuda.graph(g):
                launch(Vt, outt, torch.cuda.current_stream(X.device).cuda_stream)
            return [g, capi.lib().curv_launch_count() - n0]  # the capture counted the first replay

        entry = self._graphs.get(base)
        if entry is None:  # first sighting: eager (also warms up one-time initialisation inside the library)
            if len(self._graphs) >= _MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[base] = {"first": addr, "exact": None, "staged": None}
            launch(V, out, stream)
            del keep
            return
        if entry["exact"] is None and entry["staged"] is None and entry["first"] == addr:
            entry["exact"] = captured(V, out)  # records the launches, does not run them
            entry["exact"][0].replay()
        elif entry["exact"] is not None and entry["first"] == addr:
            capi.lib().curv_add_launch_count(entry["exact"][1])
            entry["exact"][0].replay()
        else:
            st = entry["staged"]
            if st is None:
                Vs, outs = torch.empty_like(V), torch.empty_like(out)
                Vs.copy_(V)
                outs.copy_(out)
                st = entry["staged"] = captured(Vs, outs) + [Vs, outs]
            else:
                st[2].copy_(V)
                st[3].copy_(out)
                capi.lib().curv_add_launch_count(st[1])
            st[0].replay()
            out.copy_(st[3])
        del keep

    def predict(self, X: Tensor) -> Tensor:
        """Primal forward through the engine; returns the prediction ``[B, C]`` (a copy)."""
        self._check_supported()
        with torch.cuda.device(X.device):
            return self._predict(X)

    def _predict(self, X: Tensor) -> Tensor:
        X = X.to(torch.float32).contiguous()
        prog = self.program(X, 1, False)
        X = prog.engine_input(X)
        ws = self.workspace(prog.ws_bytes, X.device)
        keep, pptrs = self._param_ptrs()
        stream = torch.cuda.current_stream(X.device).cuda_stream
        capi.check(capi.lib().curv_matmat_batch(
            prog.handle, capi.KIND_FORWARD, 0, pptrs, prog.const_ptrs, X.data_ptr(), 0, 0, 0, 0, 0, 0, 1, 0,
            1.0, 1.0, ws.data_ptr(), ws.numel() * 4, stream))
        del keep
        return prog.value_view(ws, prog.out_value, 1)[0, :, 0, 0, : prog.classes].clone()

    def mc_grad_outputs(self, X: Tensor, mc_samples: int, shard: tuple[int, int, int] | None = None) -> Tensor:
        """Would-be gradients ``[B, M, C] / sqrt(M)`` of the samples in ``X`` (``ggn_utils.py:174-271,369-372``), drawn
        from the global RNG of the operator's device.

        The draw is keyed on the GLOBAL sample index: uniform / normal variates are generated for every sample of the
        whole mini-batch (``shard = (lo, hi, B_global)``: ``X`` holds samples ``lo..hi`` of it) and the labels follow
        by inverse-CDF sampling, so a rank that processes a slice of the mini-batch draws exactly what a single
        process draws for those samples: Monte-Carlo products do not depend on the number of ranks.
        (The reference's ``multinomial`` consumes the stream differently; parity with ITS samples is tested by handing
        the engine the reference's draws.)"""
        f = self.predict(X)
        B, Cc = f.shape
        lo, hi, Bg = (0, B, B) if shard is None else shard
        lf = self.loss_func
        if isinstance(lf, CrossEntropyLoss):
            p = torch.softmax(f, dim=1)
            u = torch.rand(Bg, mc_samples, dtype=f.dtype, device=f.device)[lo:hi]
            cdf = p.cumsum(1)
            yhat = torch.searchsorted(cdf, u * cdf[:, -1:]).clamp_(max=Cc - 1)
            g = p.unsqueeze(1) - torch.nn.functional.one_hot(yhat, num_classes=Cc).to(f.dtype)
        elif isinstance(lf, MSELoss):
            c = 1.0 / Cc if lf.reduction == "mean" else 1.0
            g = torch.randn(Bg, mc_samples, Cc, dtype=f.dtype, device=f.device)[lo:hi] * math.sqrt(2 * c)
        elif isinstance(lf, BCEWithLogitsLoss):
            c = 1.0 / Cc if lf.reduction == "mean" else 1.0
            s = torch.sigmoid(f).unsqueeze(1).expand(B, mc_samples, Cc)
            u = torch.rand(Bg, mc_samples, Cc, dtype=f.dtype, device=f.device)[lo:hi]
            g = math.sqrt(c) * (s - (u < s).to(f.dtype))
        else:
            raise NotImplementedError(f"MC sampling for {lf}.")
        return g / math.sqrt(mc_samples)
