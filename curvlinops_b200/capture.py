"""Model -> static layer program.

The reference differentiates an opaque ``(params, X) -> prediction`` function with ``torch.func``
(``curvlinops/ggn.py:61-71``, ``curvlinops/hessian.py:66``).  This engine instead traces that function
once per input shape into ATen ops (the same ``make_fx(functionalize(f))`` trace the reference's own
KFAC IO collector uses, ``curvlinops/computers/io_collector/collector.py:107``) and lowers the graph to
the fixed op set the CUDA kernels implement.  Anything outside that set raises
``NotImplementedError`` -- there is no autograd / CPU fallback.

Supported ATen ops: convolution (groups=1, dilation=1), addmm / mm (+t) i.e. ``nn.Linear`` on 2-d inputs and on
token sequences ``[B, T, D]`` (a 1x1 convolution over the tokens), native_batch_norm in eval mode, native_layer_norm
over the last dimension of ``[B, D]`` / ``[B, T, D]`` tensors, relu / sigmoid / tanh / gelu (exact), add.Tensor
(residual), max_pool2d, mean over (H, W) or over the tokens / adaptive_avg_pool2d(1), view / flatten / reshape
between 4-d and 2-d and between ``[B, T, D]`` and ``[B*T, D]``, detach.

Activations are stored channels-last (``[B, H, W, C]``); a token sequence ``[B, T, D]`` is the map H = 1, W = T, C = D.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch
from torch import Tensor
from torch.func import functionalize
from torch.fx.experimental.proxy_tensor import make_fx

from . import _capi as capi

aten = torch.ops.aten


@dataclass
class LayerProgram:
    """Host-side description handed to ``curv_program_create``."""

    values: list = field(default_factory=list)  # (C, H, W, has_tangent)
    nodes: list = field(default_factory=list)   # dicts with NodeDesc fields
    consts: list = field(default_factory=list)  # constant tensors referenced by c0..c3
    conv_nodes: dict = field(default_factory=dict)  # weight param name -> node index
    bias_nodes: dict = field(default_factory=dict)  # bias param name -> node index
    out_features: int = 0
    tied: set = field(default_factory=set)  # parameters used by more than one layer
    conv_usages: dict = field(default_factory=dict)  # weight param name -> [node indices] (several: weight tying)
    bias_usages: dict = field(default_factory=dict)  # bias param name -> [node indices]
    tokens_input: bool = False  # the network input is a token sequence [B, T, D] (handed to the engine as [B, D, 1, T])

    def add_value(self, C, H, W, tan):
        self.values.append([int(C), int(H), int(W), bool(tan)])
        return len(self.values) - 1

    def add_const(self, t: Tensor) -> int:
        self.consts.append(t.detach().to(torch.float32).contiguous())
        return len(self.consts) - 1

    def add_node(self, op, in0=-1, in1=-1, out=-1, p0=-1, p1=-1, c0=-1, c1=-1, c2=-1, c3=-1,
                 kh=1, kw=1, sh=1, sw=1, ph=0, pw=0, eps=0.0):
        self.nodes.append(dict(op=op, in0=in0, in1=in1, out=out, p0=p0, p1=p1, c0=c0, c1=c1, c2=c2,
                               c3=c3, kh=kh, kw=kw, sh=sh, sw=sw, ph=ph, pw=pw, eps=eps))
        return len(self.nodes) - 1


class _Ref:
    """What an FX node evaluates to during lowering."""

    def __init__(self, kind, **kw):
        self.kind = kind  # "act" | "param" | "const" | "paramT" | "constT" | "tuple"
        self.__dict__.update(kw)


def _pair(x):
    return (int(x[0]), int(x[1])) if len(x) == 2 else (int(x[0]), int(x[0]))


def _simple_view(node, ref, B, prog) -> bool:
    """Reshapes the plain table handles itself: [B, T, C] <-> [B*T, C] of a token activation; flatten of a feature map
    to [B, C*H*W] and back to [B, C, H, W]."""
    if ref.kind != "act" or node.target not in (aten.view.default, aten._unsafe_view.default, aten.reshape.default):
        return False
    C, H, W, _ = prog.values[ref.value]
    shape = tuple(int(d) for d in node.meta["val"].shape) if "val" in node.meta else None
    tk = getattr(ref, "tok", None)
    if tk is None:
        return shape in ((B, C * H * W), (B, C, H, W))
    return (tk == 3 and shape == (B * W, C)) or (tk == 2 and shape == (B, W, C))


def capture(model_func, params: dict[str, Tensor], X: Tensor, fuse_relu: bool = True) -> LayerProgram:
    """Trace ``model_func(params, X)`` and lower it to a :class:`LayerProgram`.

    ``fuse_relu``: a ReLU whose only producer is a BatchNorm or a residual add (and which is that value's only
    consumer) is folded into the producing node (``kh = 2``): one streaming pass instead of two.

    ``params`` are the *differentiated* tensors (the operator's ``params`` dict, in order); every other
    tensor the function touches becomes a constant, exactly like ``functional_call`` falls back to the
    module's own parameters and buffers (reference ``curvlinops/utils.py:267-297``).
    """
    if not isinstance(X, Tensor) or X.ndim not in (2, 3, 4):
        raise NotImplementedError(
            f"The B200 engine supports tensor inputs of shape [B, C], [B, T, C] or [B, C, H, W]; got {type(X).__name__}"
            + (f" {tuple(X.shape)}" if isinstance(X, Tensor) else "") + "."
        )
    names = list(params.keys())

    def f(p, x):
        return model_func(p, x)

    with torch.no_grad():
        gm = make_fx(functionalize(f), tracing_mode="fake", _allow_non_fake_inputs=True)(params, X)

    prog = LayerProgram()
    env: dict = {}
    producer: dict = {}  # value id -> index of the node that writes it
    placeholders = [n for n in gm.graph.nodes if n.op == "placeholder"]
    if len(placeholders) != len(names) + 1:
        raise NotImplementedError("Unexpected trace signature (non-tensor leaves in params?).")
    for i, n in enumerate(placeholders[:-1]):
        env[n] = _Ref("param", index=i, shape=tuple(params[names[i]].shape))
    xin = placeholders[-1]
    if X.ndim == 4:
        v = prog.add_value(X.shape[1], X.shape[2], X.shape[3], False)
    elif X.ndim == 3:  # tokens [B, T, D]: already channels-last, the engine reads it as [B, D, 1, T] NCHW ... see below
        v = prog.add_value(X.shape[2], 1, X.shape[1], False)
    else:
        v = prog.add_value(X.shape[1], 1, 1, False)
    prog.add_node(capi.OP_INPUT, out=v)
    prog.tokens_input = X.ndim == 3
    env[xin] = _Ref("act", value=v, flat=False, tok=3 if X.ndim == 3 else None)
    Bsz = int(X.shape[0])

    def tok_of(ref):
        """None: [B, C, H, W] / [B, C] tensor; 3: token sequence [B, T, D]; 2: its [B*T, D] view."""
        return getattr(ref, "tok", None)

    def like(xr, value):
        return _Ref("act", value=value, flat=xr.flat, tok=tok_of(xr))

    # ---- layout bookkeeping of token tensors ---------------------------------------------------------------------
    # The engine stores a token value as [B, T, C] whatever the traced program does with it.  nn.MultiheadAttention
    # works sequence-first and splits heads through a chain of view / transpose / select nodes; instead of pattern
    # matching that chain, a "lazy" reference carries an explicit index map: an integer tensor shaped like the traced
    # tensor whose entries are the flat positions in the engine value it reads.  View-like nodes are applied to the
    # map; as soon as the map equals one of the layouts a layer can consume, the reference becomes an ordinary
    # activation again with the tag tok = 3 ([B, T, C]), 2 ([B*T, C]), "t3" ([T, B, C]) or "t2" ([T*B, C]).
    _maps: dict = {}

    def token_maps(value):
        C, H, W, _ = prog.values[value]
        if (W, C) not in _maps:
            base = torch.arange(Bsz * W * C, dtype=torch.int32).view(Bsz, W, C)
            _maps[(W, C)] = {3: base, 2: base.reshape(Bsz * W, C), "t3": base.transpose(0, 1),
                             "t2": base.transpose(0, 1).reshape(W * Bsz, C)}
        return _maps[(W, C)]

    def to_lazy(ref):
        if ref.kind == "lazy":
            return ref
        if ref.kind != "act":
            return None
        C, H, W, _ = shape_of(ref)
        tk = tok_of(ref)
        if tk is not None:
            return _Ref("lazy", value=ref.value, idx=token_maps(ref.value)[tk])
        # feature map: the traced tensor is [B, C, H, W] (or its flattening), the engine stores [B, H, W, C]
        idx = torch.arange(Bsz * H * W * C, dtype=torch.int32).view(Bsz, H, W, C).permute(0, 3, 1, 2)
        return _Ref("lazy", value=ref.value, idx=idx.reshape(Bsz, C * H * W) if ref.flat else idx)

    def resolve(lz):
        """lazy -> tagged activation if its map is one of the consumable layouts, else the lazy reference itself."""
        C, H, W, tan = prog.values[lz.value]
        shape = tuple(lz.idx.shape)
        if shape == (Bsz, H * W, C) and torch.equal(lz.idx, torch.arange(Bsz * H * W * C, dtype=torch.int32).view(shape)):
            if H == 1:
                return _Ref("act", value=lz.value, flat=False, tok=3)
            # [B, C, H, W] patch map read as a [B, H*W, C] token sequence: same elements of the channels-last storage
            ov = prog.add_value(C, 1, H * W, tan)
            prog.add_node(capi.OP_RESHAPE, in0=lz.value, out=ov)
            return _Ref("act", value=ov, flat=False, tok=3)
        if H != 1:
            return lz
        maps = token_maps(lz.value)
        for tag, m in maps.items():
            if tuple(m.shape) == shape and torch.equal(m, lz.idx):
                return _Ref("act", value=lz.value, flat=False, tok=tag)
        if shape == (Bsz, C):  # one token of every sequence (the class-token read-out)
            t0 = int(lz.idx[0, 0]) // C
            if 0 <= t0 < W and torch.equal(lz.idx, maps[3][:, t0, :]):
                ov = prog.add_value(C, 1, 1, tan)
                prog.add_node(capi.OP_TOKSEL, in0=lz.value, out=ov, kw=t0)
                return _Ref("act", value=ov, flat=True)
        return lz

    VIEW_LIKE = {aten.view.default, aten._unsafe_view.default, aten.reshape.default, aten.unsqueeze.default,
                 aten.squeeze.dim, aten.transpose.int, aten.permute.default, aten.select.int}

    def apply_view(node, lz):
        t, args = node.target, node.args
        if t in (aten.view.default, aten._unsafe_view.default, aten.reshape.default):
            idx = lz.idx.reshape(tuple(int(d) for d in node.meta["val"].shape))
        elif t == aten.unsqueeze.default:
            idx = lz.idx.unsqueeze(int(args[1]))
        elif t == aten.squeeze.dim:
            idx = lz.idx.squeeze(int(args[1]))
        elif t == aten.transpose.int:
            idx = lz.idx.transpose(int(args[1]), int(args[2]))
        elif t == aten.permute.default:
            idx = lz.idx.permute(*[int(d) for d in args[1]])
        else:
            idx = lz.idx.select(int(args[1]), int(args[2]))
        return resolve(_Ref("lazy", value=lz.value, idx=idx))

    # liveness: only lower nodes the output depends on
    out_node = [n for n in gm.graph.nodes if n.op == "output"][0]
    live = set()
    stack = list(out_node.all_input_nodes)
    while stack:
        n = stack.pop()
        if not isinstance(n, torch.fx.Node) or n in live:
            continue
        live.add(n)
        stack.extend(a for a in n.all_input_nodes)

    # Consumers per ENGINE value: detach / alias / clone / contiguous / view / getitem nodes resolve to the value of
    # their input, so the readers of a value are the non-alias users of every FX node of its alias class.
    alias_targets = (aten.detach.default, aten.alias.default, aten.clone.default, aten.contiguous.default,
                     aten.view.default, aten._unsafe_view.default, aten.reshape.default, aten.flatten.using_ints)

    def is_alias(n):
        return n.op == "call_function" and (n.target in alias_targets or "getitem" in str(n.target))

    def alias_root(n):
        while is_alias(n) and isinstance(n.args[0], torch.fx.Node):
            n = n.args[0]
        return n

    readers: dict = {}
    for n in gm.graph.nodes:
        if n.op == "placeholder" or is_alias(n) or (n.op != "output" and n not in live):
            continue
        for r in {alias_root(i) for i in n.all_input_nodes}:
            readers[r] = readers.get(r, 0) + 1

    def shape_of(ref):
        return prog.values[ref.value]

    def tan_of(*refs):
        return any(prog.values[r.value][3] for r in refs if r is not None and r.kind == "act")

    def unsupported(node, why=""):
        raise NotImplementedError(
            f"Operation {node.target} is not supported by the B200 curvature engine"
            f"{': ' + why if why else ''}. Supported layers: Linear, Conv2d, BatchNorm2d (eval), LayerNorm, ReLU,"
            " Sigmoid, Tanh, GELU, MaxPool2d, global average pooling, residual adds, flatten, token sequences with"
            " multi-head self-attention, class token and position embedding."
        )

    def weight_slots(ref):
        """(param index or -1, const index or -1) for a weight-like reference."""
        if ref is None:
            return -1, -1
        if ref.kind == "param":
            return ref.index, -1
        if ref.kind == "const":
            return -1, prog.add_const(ref.tensor)
        raise NotImplementedError("Parameters must be used directly by a supported layer.")

    def emit_conv(node, xr, wref, bref, kh, kw, sh, sw, ph, pw, cout, ho, wo):
        p0, c0 = weight_slots(wref)
        p1, c1 = weight_slots(bref)
        tan = tan_of(xr) or p0 >= 0 or p1 >= 0
        ov = prog.add_value(cout, ho, wo, tan)
        ni = prog.add_node(capi.OP_CONV, in0=xr.value, out=ov, p0=p0, p1=p1, c0=c0, c1=c1, kh=kh, kw=kw,
                           sh=sh, sw=sw, ph=ph, pw=pw)
        if p0 >= 0:
            if names[p0] in prog.conv_nodes:
                prog.tied.add(names[p0])
            prog.conv_nodes[names[p0]] = ni
            prog.conv_usages.setdefault(names[p0], []).append(ni)
        if p1 >= 0:
            prog.bias_nodes[names[p1]] = ni
            prog.bias_usages.setdefault(names[p1], []).append(ni)
        return _Ref("act", value=ov, flat=False)

    def emit_linear(node, xr, wref, bref):
        """x [B, in] @ W^T + b, W stored [out, in] (nn.Linear)."""
        if wref.kind not in ("paramT", "constT"):
            unsupported(node, "matmul operand must be a transposed 2-d weight")
        w = wref.base
        C, H, W, _ = shape_of(xr)
        wshape = w.shape if w.kind == "param" else tuple(w.tensor.shape)
        if tok_of(xr) in (2, "t2"):  # Linear applied to every token: a 1x1 convolution over the [1, T] token map
            if len(wshape) != 2 or wshape[1] != C:
                unsupported(node, f"weight shape {wshape} does not match token features {C}")
            out = emit_conv(node, xr, w, bref, 1, 1, 1, 1, 0, 0, wshape[0], H, W)
            out.tok = tok_of(xr)  # rows in the same order as the input's
            return out
        if tok_of(xr) in (3, "t3"):
            unsupported(node, "matmul on a [B, T, D] tensor that was not flattened to [B*T, D]")
        if len(wshape) != 2 or wshape[1] != C * H * W:
            unsupported(node, f"weight shape {wshape} does not match input features {C * H * W}")
        if (H, W) != (1, 1) and not xr.flat:
            unsupported(node, "Linear on a non-flattened >2-d tensor")
        # Linear after flatten of a [C, H, W] map == 'valid' HxW convolution (NCHW flatten order
        # (c, h, w) is exactly the conv weight layout [out, C, H, W]).
        return emit_conv(node, xr, w, bref, H, W, 1, 1, 0, 0, wshape[0], 1, 1)

    def emit_attention(node, a):
        """scaled_dot_product_attention on the head-split views of one packed projection -> OP_ATTENTION."""
        schema = [arg.name for arg in node.target._schema.arguments]
        vals = dict(zip(schema, node.args))  # arguments left at their defaults are absent
        vals.update(node.kwargs)
        if vals.get("dropout_p") not in (None, 0, 0.0) or vals.get("is_causal") or vals.get("scale") is not None \
                or vals.get("attn_mask") is not None or vals.get("attn_bias") is not None:
            unsupported(node, "attention with dropout, masks, causal masking or a custom scale")
        q, k, v = (to_lazy(r) if isinstance(r, _Ref) else None for r in a[:3])
        if q is None or k is None or v is None or not (q.value == k.value == v.value) or q.idx.ndim != 4:
            unsupported(node, "attention whose query / key / value are not head-split views of one packed projection")
        C3, H, T, tan = prog.values[q.value]
        heads, E = int(q.idx.shape[1]), C3 // 3
        if H != 1 or C3 % 3 or E % heads or E % 8:
            unsupported(node, "attention needs a packed [B, T, 3E] projection with E a multiple of 8 and of the head count")
        grid = torch.arange(Bsz * T * C3, dtype=torch.int32).view(Bsz, T, 3, heads, E // heads)
        for j, r in enumerate((q, k, v)):
            if not torch.equal(r.idx, grid[:, :, j].permute(0, 2, 1, 3)):
                unsupported(node, "attention operands are not the q / k / v thirds of the packed projection, split into heads")
        ov = prog.add_value(E, 1, T, tan)
        prog.add_node(capi.OP_ATTENTION, in0=q.value, out=ov, kh=heads)
        out_idx = torch.arange(Bsz * T * E, dtype=torch.int32).view(Bsz, T, heads, E // heads).permute(0, 2, 1, 3)
        return _Ref("tuple", items=[_Ref("lazy", value=ov, idx=out_idx)] + [None] * 9)

    for node in gm.graph.nodes:
        if node.op in ("placeholder", "output") or node not in live:
            continue
        if node.op == "get_attr":
            env[node] = _Ref("const", tensor=getattr(gm, node.target))
            continue
        for ni_, nd_ in enumerate(prog.nodes[len(producer):], start=len(producer)):
            producer[nd_["out"]] = ni_
        t = node.target
        a = [env.get(x, x) if isinstance(x, torch.fx.Node) else x for x in node.args]
        if t in (aten.detach.default, aten.alias.default, aten.clone.default, aten.contiguous.default):
            env[node] = a[0]
        elif (t in VIEW_LIKE and isinstance(a[0], _Ref) and a[0].kind in ("lazy", "act")
              and not _simple_view(node, a[0], Bsz, prog)):
            env[node] = apply_view(node, to_lazy(a[0]))
        elif t == aten.unbind.int and isinstance(a[0], _Ref) and a[0].kind in ("lazy", "act"):
            lz = to_lazy(a[0])  # q, k, v = qkv.unbind(0): one select per item
            dim = int(a[1]) if len(a) > 1 else 0
            env[node] = _Ref("tuple", items=[resolve(_Ref("lazy", value=lz.value, idx=lz.idx.select(dim, i)))
                                             for i in range(lz.idx.shape[dim])])
        elif t == aten.expand.default and isinstance(a[0], _Ref) and a[0].kind in ("param", "const"):
            env[node] = _Ref("expand", base=a[0], shape=tuple(int(d) for d in node.meta["val"].shape))
        elif t == aten.cat.default:
            items = [env.get(x) for x in node.args[0]]
            dim = node.args[1] if len(node.args) > 1 else 0
            if (len(items) != 2 or items[0] is None or items[1] is None or items[0].kind != "expand"
                    or items[1].kind != "act" or tok_of(items[1]) != 3 or dim != 1):
                unsupported(node, "only cat([class_token.expand(B, -1, -1), tokens], dim=1)")
            C, H, W, tan = shape_of(items[1])
            base = items[0].base
            bshape = base.shape if base.kind == "param" else tuple(base.tensor.shape)
            if tuple(bshape) != (1, 1, C) or items[0].shape != (Bsz, 1, C) or C % 8:
                unsupported(node, "class token must be [1, 1, C] with C a multiple of 8")
            p0, c0 = weight_slots(base)
            ov = prog.add_value(C, 1, W + 1, tan or p0 >= 0)
            prog.add_node(capi.OP_CLSCAT, in0=items[1].value, out=ov, p0=p0, c0=c0)
            env[node] = _Ref("act", value=ov, flat=False, tok=3)
        elif "_scaled_dot_product" in str(t):
            env[node] = emit_attention(node, a)
        elif t == aten.t.default or (t == aten.transpose.int and a[0].kind in ("param", "const")):
            if a[0].kind not in ("param", "const"):
                unsupported(node, "transpose of an activation")
            env[node] = _Ref(a[0].kind + "T", base=a[0])
        elif t == aten.convolution.default:
            xr, wref, bref, stride, padding, dilation, transposed, _outpad, groups = a
            if xr.kind != "act" or transposed or groups != 1 or _pair(dilation) != (1, 1) or tok_of(xr):
                unsupported(node, "only plain 2-d convolutions (groups=1, dilation=1) of [B, C, H, W] tensors")
            wshape = wref.shape if wref.kind == "param" else tuple(wref.tensor.shape)
            if len(wshape) != 4:
                unsupported(node, "only 2-d convolutions")
            (sh, sw), (ph, pw) = _pair(stride), _pair(padding)
            C, H, W, _ = shape_of(xr)
            ho = (H + 2 * ph - wshape[2]) // sh + 1
            wo = (W + 2 * pw - wshape[3]) // sw + 1
            env[node] = emit_conv(node, xr, wref, bref, wshape[2], wshape[3], sh, sw, ph, pw, wshape[0],
                                  ho, wo)
        elif t == aten.addmm.default:
            bref, xr, wref = a[0], a[1], a[2]
            if xr.kind != "act" or bref.kind not in ("param", "const"):
                unsupported(node)
            env[node] = emit_linear(node, xr, wref, bref)
        elif t == aten.mm.default:
            xr, wref = a[0], a[1]
            if xr.kind != "act":
                unsupported(node)
            env[node] = emit_linear(node, xr, wref, None)
        elif t in (aten.native_batch_norm.default, aten._native_batch_norm_legit_no_training.default,
                   aten._native_batch_norm_legit.default, aten.cudnn_batch_norm.default):
            if t == aten._native_batch_norm_legit_no_training.default:
                xr, wref, bref, rm, rv, _mom, eps = a
                training = False
            elif t == aten.cudnn_batch_norm.default:
                xr, wref, bref, rm, rv, training, _mom, eps = a
            else:
                xr, wref, bref, rm, rv, training, _mom, eps = a
            if tok_of(xr):
                unsupported(node, "BatchNorm on a token sequence")
            if training or rm is None or rv is None:
                raise NotImplementedError(
                    "BatchNorm in training mode couples the samples of a mini-batch; put the model in"
                    " eval() mode (the reference's determinism check rejects it as well)."
                )
            p0, c0 = weight_slots(wref)
            p1, c1 = weight_slots(bref)
            C, H, W, _ = shape_of(xr)
            tan = tan_of(xr) or p0 >= 0 or p1 >= 0
            ov = prog.add_value(C, H, W, tan)
            prog.add_node(capi.OP_AFFINE, in0=xr.value, out=ov, p0=p0, p1=p1, c0=c0, c1=c1,
                          c2=prog.add_const(rm.tensor), c3=prog.add_const(rv.tensor), eps=float(eps))
            env[node] = _Ref("tuple", items=[_Ref("act", value=ov, flat=False), None, None])
        elif t.__name__ == "getitem" or str(t) == "<built-in function getitem>":
            src, idx = a
            if src.kind != "tuple":
                unsupported(node)
            env[node] = src.items[idx]
        elif t == aten.native_layer_norm.default:
            xr, nshape, wref, bref, eps = a
            C, H, W, _ = shape_of(xr)
            if xr.kind != "act" or [int(d) for d in nshape] != [C] or (tok_of(xr) is None and (H, W) != (1, 1)):
                unsupported(node, "LayerNorm over anything but the last dimension of a [B, D] / [B, T, D] tensor")
            p0, c0 = weight_slots(wref)
            p1, c1 = weight_slots(bref)
            ov = prog.add_value(C, H, W, tan_of(xr) or p0 >= 0 or p1 >= 0)
            prog.add_node(capi.OP_LAYERNORM, in0=xr.value, out=ov, p0=p0, p1=p1, c0=c0, c1=c1, eps=float(eps))
            env[node] = _Ref("tuple", items=[like(xr, ov), None, None])
        elif t == aten.gelu.default:
            xr = a[0]
            if node.kwargs.get("approximate", "none") != "none" or (len(a) > 1 and a[1] != "none"):
                unsupported(node, "only the exact (erf) GELU")
            C, H, W, tan = shape_of(xr)
            ov = prog.add_value(C, H, W, tan)
            prog.add_node(capi.OP_GELU, in0=xr.value, out=ov)
            env[node] = like(xr, ov)
        elif t in (aten.relu.default, aten.sigmoid.default, aten.tanh.default):
            xr = a[0]
            src = node.args[0]
            prod = producer.get(xr.value) if xr.kind == "act" else None
            if (fuse_relu and t == aten.relu.default and prod is not None and readers.get(alias_root(src), 0) == 1
                    and prog.nodes[prod]["op"] in (capi.OP_AFFINE, capi.OP_ADD) and prog.nodes[prod]["kh"] != 2):
                prog.nodes[prod]["kh"] = 2  # fused: the producing node now writes relu(...)
                env[node] = like(xr, xr.value)
                continue
            op = {aten.relu.default: capi.OP_RELU, aten.sigmoid.default: capi.OP_SIGMOID,
                  aten.tanh.default: capi.OP_TANH}[t]
            C, H, W, tan = shape_of(xr)
            ov = prog.add_value(C, H, W, tan)
            prog.add_node(op, in0=xr.value, out=ov)
            env[node] = like(xr, ov)
        elif t == aten.add.Tensor and isinstance(a[0], _Ref) and isinstance(a[1], _Ref) and a[0].kind == "act" \
                and a[1].kind in ("param", "const") and node.kwargs.get("alpha", 1) == 1:
            # x @ W^T traced as mm followed by "+ bias" (nn.MultiheadAttention's packed projection): fold the bias
            # into the producing layer
            xr, bref = a[0], a[1]
            prod = producer.get(xr.value)
            bshape = bref.shape if bref.kind == "param" else tuple(bref.tensor.shape)
            if tok_of(xr) == 3 and tuple(bshape) == (1, shape_of(xr)[2], shape_of(xr)[0]) and shape_of(xr)[0] % 8 == 0:
                C, H, W, tan = shape_of(xr)  # position embedding, broadcast over the batch
                p0, c0 = weight_slots(bref)
                ov = prog.add_value(C, 1, W, tan or p0 >= 0)
                prog.add_node(capi.OP_POSADD, in0=xr.value, out=ov, p0=p0, c0=c0)
                env[node] = _Ref("act", value=ov, flat=False, tok=3)
                continue
            if (prod is None or prog.nodes[prod]["op"] != capi.OP_CONV or prog.nodes[prod]["p1"] >= 0
                    or prog.nodes[prod]["c1"] >= 0 or tuple(bshape) != (shape_of(xr)[0],)
                    or readers.get(alias_root(node.args[0]), 0) != 1):
                unsupported(node, "adding a parameter to an activation (other than the bias of the layer that produced it)")
            p1, c1 = weight_slots(bref)
            prog.nodes[prod]["p1"], prog.nodes[prod]["c1"] = p1, c1
            if p1 >= 0:
                prog.bias_nodes[names[p1]] = prod
                prog.bias_usages.setdefault(names[p1], []).append(prod)
                prog.values[xr.value][3] = True
            env[node] = xr
        elif t == aten.add.Tensor:
            xr, yr = a[0], a[1]
            alpha = node.kwargs.get("alpha", 1)
            if not (isinstance(xr, _Ref) and isinstance(yr, _Ref) and xr.kind == yr.kind == "act") \
                    or alpha != 1 or shape_of(xr)[:3] != shape_of(yr)[:3] or tok_of(xr) != tok_of(yr):
                unsupported(node, "only additions of two activations of equal shape")
            C, H, W, _ = shape_of(xr)
            ov = prog.add_value(C, H, W, tan_of(xr, yr))
            prog.add_node(capi.OP_ADD, in0=xr.value, in1=yr.value, out=ov)
            env[node] = like(xr, ov)
        elif t == aten.max_pool2d_with_indices.default:
            xr = a[0]
            if tok_of(xr):
                unsupported(node, "pooling windows on a token sequence")
            ks = _pair(a[1])
            st = _pair(a[2]) if len(a) > 2 and a[2] else ks
            pd = _pair(a[3]) if len(a) > 3 else (0, 0)
            if len(a) > 4 and _pair(a[4]) != (1, 1) or (len(a) > 5 and a[5]):
                unsupported(node, "dilated / ceil_mode pooling")
            C, H, W, tan = shape_of(xr)
            ho = (H + 2 * pd[0] - ks[0]) // st[0] + 1
            wo = (W + 2 * pd[1] - ks[1]) // st[1] + 1
            ov = prog.add_value(C, ho, wo, tan)
            prog.add_node(capi.OP_MAXPOOL, in0=xr.value, out=ov, kh=ks[0], kw=ks[1], sh=st[0], sw=st[1],
                          ph=pd[0], pw=pd[1])
            env[node] = _Ref("tuple", items=[_Ref("act", value=ov, flat=False), None])
        elif t in (aten.mean.dim, aten._adaptive_avg_pool2d.default, aten.adaptive_avg_pool2d.default):
            xr = a[0]
            C, H, W, tan = shape_of(xr)
            if t == aten.mean.dim and tok_of(xr) == 3:
                if sorted(d % 3 for d in a[1]) != [1] or (len(a) > 2 and a[2]):
                    unsupported(node, "mean of a [B, T, D] tensor over anything but the tokens")
                keep = False
            elif tok_of(xr):
                unsupported(node, "pooling of a flattened token sequence")
            elif t == aten.mean.dim:
                dims = sorted(d % 4 for d in a[1])
                if dims != [2, 3]:
                    unsupported(node, "mean over dims other than (H, W)")
                keep = bool(a[2]) if len(a) > 2 else False
            else:
                if _pair(a[1]) != (1, 1):
                    unsupported(node, "adaptive pooling to sizes other than 1x1")
                keep = True
            ov = prog.add_value(C, 1, 1, tan)
            prog.add_node(capi.OP_AVGPOOL, in0=xr.value, out=ov)
            env[node] = _Ref("act", value=ov, flat=not keep)
        elif t in (aten.view.default, aten._unsafe_view.default, aten.reshape.default,
                   aten.flatten.using_ints):
            xr = a[0]
            if xr.kind != "act":
                unsupported(node, "view of a parameter")
            C, H, W, _ = shape_of(xr)
            oshape = tuple(node.meta["val"].shape) if "val" in node.meta else None
            if oshape is None:
                unsupported(node, "missing shape metadata")
            if tok_of(xr):  # token sequences: [B, T, D] <-> [B*T, D]
                if len(oshape) == 2 and tuple(oshape) == (Bsz * W, C):
                    env[node] = _Ref("act", value=xr.value, flat=False, tok=2)
                elif len(oshape) == 3 and tuple(oshape) == (Bsz, W, C):
                    env[node] = _Ref("act", value=xr.value, flat=False, tok=3)
                else:
                    unsupported(node, f"reshape of a token sequence to {tuple(oshape)}")
            elif len(oshape) == 2 and oshape[1] == C * H * W:
                env[node] = _Ref("act", value=xr.value, flat=True)
            elif len(oshape) == 4 and tuple(oshape[1:]) == (C, H, W):
                env[node] = _Ref("act", value=xr.value, flat=False)
            else:
                unsupported(node, f"reshape to {oshape}")
        else:
            unsupported(node)

    res = out_node.args[0]
    if isinstance(res, (list, tuple)):
        if len(res) != 1:
            raise NotImplementedError("The model must return a single tensor.")
        res = res[0]
    ref = env[res]
    if ref.kind != "act":
        raise NotImplementedError("The model output must be an activation tensor.")
    C, H, W, _ = shape_of(ref)
    if (H, W) != (1, 1):
        raise NotImplementedError("The B200 engine needs a 2-d prediction [batch, C].")
    if prog.nodes[-1]["out"] != ref.value:
        raise NotImplementedError("The model output must be produced by the last layer of the trace.")
    prog.out_features = C
    return prog
