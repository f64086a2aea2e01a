"""CPU-only tests: C-ABI exports, capture/planning (host side, no kernels), operator format handling."""
import ctypes as C
import os
import re

import pytest
import torch
from torch import nn

from curvlinops_b200 import _capi as capi
from curvlinops_b200.capture import capture
from curvlinops_b200.curvature import make_functional_call
from curvlinops_b200.engine import CompiledProgram, loss_scale
from curvlinops_b200.linop import PyTorchLinearOperator
from oracle.models import ConvNetBias, MiniResNet, mlp_c1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "curvb200.h")).read()
    declared = set(re.findall(r"\b(curv_[a-z_0-9]+)\s*\(", header))
    L = capi.lib()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in curvb200.h but not exported"
    assert set(capi.EXPORTS) <= declared
    assert L.curv_abi_version() == 1


def test_compute_call_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    model = mlp_c1()
    params = dict(model.named_parameters())
    prog = CompiledProgram(make_functional_call(model), params, torch.randn(4, 64), 2, False)
    ws = torch.empty(prog.ws_bytes // 4 + 1)
    pp = capi.ptr_array([p.data_ptr() for p in params.values()])
    rc = capi.lib().curv_matmat_batch(prog.handle, capi.KIND_GGN, 0, pp, prog.const_ptrs, 0, 0, 0, 0, 0, 0, 1,
                                      1, 0, 1.0, 1.0, ws.data_ptr(), ws.numel() * 4, 0)
    assert rc == capi.ERR_CUDA
    assert b"no CPU fallback" in capi.lib().curv_last_error()


def test_capture_mlp():
    model = mlp_c1()
    params = dict(model.named_parameters())
    lp = capture(make_functional_call(model), params, torch.randn(5, 64))
    ops = [n["op"] for n in lp.nodes]
    assert ops == [capi.OP_INPUT, capi.OP_CONV, capi.OP_RELU, capi.OP_CONV, capi.OP_RELU, capi.OP_CONV,
                   capi.OP_RELU, capi.OP_CONV]
    assert lp.out_features == 10
    assert [n["p0"] for n in lp.nodes if n["op"] == capi.OP_CONV] == [0, 2, 4, 6]
    assert [n["p1"] for n in lp.nodes if n["op"] == capi.OP_CONV] == [1, 3, 5, 7]


def test_capture_resnet_and_param_subset():
    model = MiniResNet().eval()
    allp = dict(model.named_parameters())
    lp = capture(make_functional_call(model), allp, torch.rand(2, 3, 32, 32))
    kinds = [n["op"] for n in lp.nodes]
    assert kinds.count(capi.OP_CONV) == 7 and kinds.count(capi.OP_AFFINE) == 6
    assert kinds.count(capi.OP_ADD) == 2 and kinds.count(capi.OP_MAXPOOL) == 1
    # subset + reordering: excluded parameters become constants, early values carry no tangent
    sub = {k: allp[k] for k in ["fc.bias", "layer2.0.conv1.weight"]}
    lp2 = capture(make_functional_call(model), sub, torch.rand(2, 3, 32, 32))
    convs = [n for n in lp2.nodes if n["op"] == capi.OP_CONV]
    assert sum(n["p0"] >= 0 for n in convs) == 1 and sum(n["p1"] >= 0 for n in convs) == 1
    assert not lp2.values[convs[0]["out"]][3]  # stem output does not depend on the selected params
    assert lp2.values[lp2.nodes[-1]["out"]][3]


def test_capture_flatten_linear_becomes_valid_conv():
    model = ConvNetBias().eval()
    lp = capture(make_functional_call(model), dict(model.named_parameters()), torch.rand(2, 3, 8, 8))
    last = lp.nodes[-1]
    assert last["op"] == capi.OP_CONV and (last["kh"], last["kw"]) == (3, 3)


def test_capture_rejects_unsupported():
    model = nn.Sequential(nn.Linear(4, 4), nn.Softplus(), nn.Linear(4, 2))
    with pytest.raises(NotImplementedError, match="not supported by the B200"):
        capture(make_functional_call(model), dict(model.named_parameters()), torch.rand(2, 4))
    tanh_gelu = nn.Sequential(nn.Linear(4, 4), nn.GELU(approximate="tanh"), nn.Linear(4, 2))
    with pytest.raises(NotImplementedError, match="exact"):
        capture(make_functional_call(tanh_gelu), dict(tanh_gelu.named_parameters()), torch.rand(2, 4))
    bn = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4)).train()
    with pytest.raises(NotImplementedError, match="eval"):
        capture(make_functional_call(bn), dict(bn.named_parameters()), torch.rand(2, 3, 8, 8))


def test_program_plan_host_only():
    model = MiniResNet().eval()
    params = dict(model.named_parameters())
    prog = CompiledProgram(make_functional_call(model), params, torch.rand(3, 3, 32, 32), 4, False)
    progh = CompiledProgram(make_functional_call(model), params, torch.rand(3, 3, 32, 32), 4, True)
    assert 0 < prog.ws_bytes < progh.ws_bytes
    assert prog.P == sum(p.numel() for p in params.values())


def test_loss_scale():
    assert loss_scale(nn.CrossEntropyLoss(), 8, 10) == 1 / 8
    assert loss_scale(nn.MSELoss(), 8, 10) == 1 / 80
    assert loss_scale(nn.MSELoss(reduction="sum"), 8, 10) == 1.0


class _Dense(PyTorchLinearOperator):
    """Mock operator: block matrix given densely (pattern of reference test__torch_base.py:17-128)."""

    def __init__(self, A, in_shape, out_shape):
        super().__init__(in_shape, out_shape)
        self.A = A

    device = property(lambda self: self.A.device)
    dtype = property(lambda self: self.A.dtype)

    def _matmat(self, X):
        K = X[0].shape[-1]
        flat = torch.cat([x.reshape(-1, K) for x in X])
        Y = self.A @ flat
        return [y.reshape(*s, K) for y, s in zip(Y.split(self._out_shape_flat), self._out_shape)]

    def _adjoint(self):
        return _Dense(self.A.T, self._out_shape, self._in_shape)


def test_linop_formats_and_errors():
    torch.manual_seed(0)
    in_shape, out_shape = [(2, 3), (4,)], [(5,), (1, 2)]
    A = torch.rand(7, 10, dtype=torch.float64)
    op = _Dense(A, in_shape, out_shape)
    assert op.shape == (7, 10)
    x = torch.rand(10, dtype=torch.float64)
    Xm = torch.rand(10, 3, dtype=torch.float64)
    torch.testing.assert_close(op @ x, A @ x)
    torch.testing.assert_close(op @ Xm, A @ Xm)
    xl = [x[:6].reshape(2, 3), x[6:]]
    yl = op @ xl
    assert [tuple(y.shape) for y in yl] == [(5,), (1, 2)]
    torch.testing.assert_close(torch.cat([y.flatten() for y in yl]), A @ x)
    Xl = [Xm[:6].reshape(2, 3, 3), Xm[6:].reshape(4, 3)]
    Yl = op @ Xl
    torch.testing.assert_close(torch.cat([y.reshape(-1, 3) for y in Yl]), A @ Xm)
    # left multiplication (leading K)
    z = torch.rand(7, dtype=torch.float64)
    Z = torch.rand(2, 7, dtype=torch.float64)
    torch.testing.assert_close(z @ op, z @ A)
    torch.testing.assert_close(Z @ op, Z @ A)
    # scipy export
    S = op.to_scipy()
    assert S.dtype == "float64"
    torch.testing.assert_close(torch.from_numpy(S @ Xm.numpy()), A @ Xm)
    torch.testing.assert_close(torch.from_numpy(S.rmatvec(z.numpy())), A.T @ z)
    # algebra
    torch.testing.assert_close((2 * op + op / 2 - op) @ x, 1.5 * (A @ x))
    chain = op.adjoint() @ op
    torch.testing.assert_close(chain @ x, A.T @ (A @ x))
    assert len(chain) == 2
    # errors
    with pytest.raises(ValueError, match="must be non-empty."):
        _Dense(A, [], out_shape)
    with pytest.raises(ValueError, match="Input must be tensor or list of tensors."):
        op @ x.numpy()
    with pytest.raises(ValueError, match="Input list must contain tensors with shapes"):
        op @ [torch.rand(2, 3), torch.rand(5)]
    with pytest.raises(ValueError, match="Input tensor must have shape"):
        op @ torch.rand(11)
    with pytest.raises(ValueError, match="Input list must have 2 tensors. Got 1."):
        op @ [torch.rand(2, 3)]
    with pytest.raises(ValueError, match="Shape mismatch"):
        op @ op
    with pytest.raises(ValueError, match="Dtype mismatch"):
        op + _Dense(A.float(), in_shape, out_shape)


def test_dict_like_inputs_are_adapted_to_the_tensor_entry():
    """Dict-like mini-batches (reference test/cases.py:36-60, 148-168): the single tensor entry becomes the engine's
    input, ``batch_size_fn`` still sees the mapping, other entries are constants of the traced model function."""
    from collections import UserDict

    import pytest
    from torch import nn

    from curvlinops_b200 import GGNLinearOperator
    from curvlinops_b200.capture import capture

    class DictModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.net = nn.Sequential(nn.Linear(10, 5), nn.ReLU(), nn.Linear(5, 3))

        def forward(self, data):
            assert data["tag"] == "a"
            return self.net(data["x"].to(next(self.parameters()).device))

    torch.manual_seed(0)
    m = DictModel().eval()
    data = [(UserDict({"x": torch.rand(3, 10), "tag": "a"}), torch.randint(0, 3, (3,))),
            (UserDict({"x": torch.rand(4, 10), "tag": "a"}), torch.randint(0, 3, (4,)))]
    params = dict(m.named_parameters())
    with pytest.raises(ValueError, match="`batch_size_fn` is required"):
        GGNLinearOperator(m, nn.CrossEntropyLoss(), params, data, check_deterministic=False)
    seen = []

    def batch_size_fn(X):
        seen.append(type(X).__name__)
        return X["x"].shape[0]

    G = GGNLinearOperator(m, nn.CrossEntropyLoss(), params, data, check_deterministic=False,
                          batch_size_fn=batch_size_fn)
    assert G._N_data == 7 and set(seen) == {"UserDict"}
    batches = list(G._loop_over_data())
    assert [tuple(X.shape) for X, _ in batches] == [(3, 10), (4, 10)]
    assert [G._batch_size_fn(X) for X, _ in batches] == [3, 4]
    assert abs(G._get_normalization_factor(*batches[1]) - 4 / 7) < 1e-12
    assert not hasattr(data[0][0]["x"], "_curv_batch_size")  # the caller's tensors stay untouched
    torch.testing.assert_close(G._model_func(params, batches[1][0]), m(data[1][0]))
    lp = capture(G._model_func, params, batches[0][0])
    assert [n["op"] for n in lp.nodes] == [capi.OP_INPUT, capi.OP_CONV, capi.OP_RELU, capi.OP_CONV]

    from curvlinops_b200 import JacobianLinearOperator

    with pytest.raises(NotImplementedError, match="need tensor inputs"):
        JacobianLinearOperator(m, params, data, check_deterministic=False, batch_size_fn=batch_size_fn)

    two = [(UserDict({"x": torch.rand(3, 10), "mask": torch.ones(3, 10)}), torch.randint(0, 3, (3,)))]
    with pytest.raises(NotImplementedError, match="exactly one tensor entry"):
        GGNLinearOperator(m, nn.CrossEntropyLoss(), params, two, check_deterministic=False,
                          batch_size_fn=batch_size_fn)


class _AliasedReLU(nn.Module):
    """relu on a clone of a BN output that is ALSO read un-activated: the ReLU must not be folded into the BN."""

    def __init__(self):
        super().__init__()
        self.conv, self.bn, self.fc = nn.Conv2d(3, 4, 3, padding=1), nn.BatchNorm2d(4), nn.Linear(4, 2)

    def forward(self, x):
        y = self.bn(self.conv(x))
        z = torch.relu(y.clone()) + y.detach().clone()
        return self.fc(z.mean((2, 3)))


def test_capture_relu_fusion_counts_readers_of_alias_nodes():
    model = _AliasedReLU().eval()
    lp = capture(make_functional_call(model), dict(model.named_parameters()), torch.rand(2, 3, 8, 8))
    affine = [n for n in lp.nodes if n["op"] == capi.OP_AFFINE]
    assert len(affine) == 1 and affine[0]["kh"] != 2  # not fused: the un-activated value has a second reader
    assert [n["op"] for n in lp.nodes].count(capi.OP_RELU) == 1
    add = [n for n in lp.nodes if n["op"] == capi.OP_ADD][0]
    assert add["in0"] != add["in1"]
    # the plain chain still fuses
    seq = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                        nn.Linear(4, 2)).eval()
    lp2 = capture(make_functional_call(seq), dict(seq.named_parameters()), torch.rand(2, 3, 8, 8))
    assert [n for n in lp2.nodes if n["op"] == capi.OP_AFFINE][0]["kh"] == 2


def test_functional_call_and_default_batch_size_fn_are_picklable():
    import pickle

    from curvlinops_b200.curvature import _leading_dim

    f = pickle.loads(pickle.dumps(make_functional_call(nn.Linear(3, 2))))
    assert f({}, torch.ones(1, 3)).shape == (1, 2)
    assert pickle.loads(pickle.dumps(_leading_dim))(torch.ones(5, 1)) == 5


def test_capture_layernorm_gelu_tokens():
    from oracle.models import TokenMLP, mlp_ln_gelu

    m = TokenMLP().eval()
    lp = capture(make_functional_call(m), dict(m.named_parameters()), torch.rand(3, 7, 12))
    ops = [n["op"] for n in lp.nodes]
    assert lp.tokens_input and ops.count(capi.OP_LAYERNORM) == 2 and ops.count(capi.OP_GELU) == 1
    assert ops.count(capi.OP_CONV) == 3 and ops.count(capi.OP_ADD) == 1 and ops.count(capi.OP_AVGPOOL) == 1
    assert lp.values[0][:3] == [12, 1, 7] and lp.values[lp.nodes[-1]["out"]][:3] == [5, 1, 1]  # tokens = the W axis
    ln = [n for n in lp.nodes if n["op"] == capi.OP_LAYERNORM][0]
    assert ln["p0"] >= 0 and ln["p1"] >= 0 and abs(ln["eps"] - 1e-5) < 1e-12
    m2 = mlp_ln_gelu().eval()
    lp2 = capture(make_functional_call(m2), dict(m2.named_parameters()), torch.rand(4, 16))
    assert [n["op"] for n in lp2.nodes] == [capi.OP_INPUT, capi.OP_CONV, capi.OP_LAYERNORM, capi.OP_GELU, capi.OP_CONV,
                                           capi.OP_LAYERNORM, capi.OP_RELU, capi.OP_CONV]
    # LayerNorm over the width of an image tensor is not the channel axis: rejected
    bad = nn.Sequential(nn.Conv2d(3, 4, 3), nn.LayerNorm(6), nn.Flatten(), nn.Linear(4 * 6 * 6, 2))
    with pytest.raises(NotImplementedError, match="LayerNorm"):
        capture(make_functional_call(bad), dict(bad.named_parameters()), torch.rand(2, 3, 8, 8))


def test_capture_attention_and_vision_transformer():
    """nn.MultiheadAttention's traced view chain is recognised through index maps; torchvision's VisionTransformer lowers
    to conv -> reshape (alias) -> class-token concat -> position add -> encoder blocks -> token read-out -> head; the plans
    build on the host; R-op programs and unsupported attention variants are rejected."""
    from torchvision.models.vision_transformer import VisionTransformer

    from oracle.models import TransformerBlock

    m = TransformerBlock(dim=16, heads=2, hidden=32, layers=2).eval()
    params = dict(m.named_parameters())
    lp = capture(make_functional_call(m), params, torch.rand(3, 7, 16))
    ops = [n["op"] for n in lp.nodes]
    assert ops.count(capi.OP_ATTENTION) == 2 and ops.count(capi.OP_CONV) == 2 * 4 + 1 and ops.count(capi.OP_ADD) == 4
    att = [n for n in lp.nodes if n["op"] == capi.OP_ATTENTION][0]
    assert att["kh"] == 2 and lp.values[att["in0"]][:3] == [48, 1, 7] and lp.values[att["out"]][:3] == [16, 1, 7]
    inproj = lp.nodes[lp.conv_nodes["blocks.0.attn.in_proj_weight"]]
    assert list(params)[inproj["p1"]] == "blocks.0.attn.in_proj_bias"  # "mm + bias" folded into the layer
    prog = CompiledProgram(make_functional_call(m), params, torch.rand(3, 7, 16), 4, False)
    assert prog.ws_bytes > 0
    with pytest.raises(NotImplementedError, match="R-op"):
        CompiledProgram(make_functional_call(m), params, torch.rand(3, 7, 16), 4, True)

    vit = VisionTransformer(image_size=32, patch_size=8, num_layers=1, num_heads=4, hidden_dim=32, mlp_dim=64,
                            num_classes=5).eval()
    vp = dict(vit.named_parameters())
    lp = capture(make_functional_call(vit), vp, torch.rand(2, 3, 32, 32))
    ops = [n["op"] for n in lp.nodes]
    assert ops[:5] == [capi.OP_INPUT, capi.OP_CONV, capi.OP_RESHAPE, capi.OP_CLSCAT, capi.OP_POSADD]
    assert ops[-2:] == [capi.OP_TOKSEL, capi.OP_CONV] and ops.count(capi.OP_ATTENTION) == 1
    names = list(vp)
    assert names[lp.nodes[3]["p0"]] == "class_token" and names[lp.nodes[4]["p0"]] == "encoder.pos_embedding"
    assert lp.values[lp.nodes[3]["out"]][:3] == [32, 1, 17] and lp.nodes[-2]["kw"] == 0
    assert CompiledProgram(make_functional_call(vit), vp, torch.rand(2, 3, 32, 32), 2, False).ws_bytes > 0
    # parameters left out of `params` become constants
    sub = {n: p for n, p in vp.items() if "mlp" in n}
    lp = capture(make_functional_call(vit), sub, torch.rand(2, 3, 32, 32))
    assert lp.nodes[3]["p0"] == -1 and lp.nodes[3]["c0"] >= 0 and not lp.values[lp.nodes[4]["out"]][3]

    class OwnAttention(nn.Module):  # the timm-style spelling: reshape / permute / unbind around F.scaled_dot_product_attention
        def __init__(self, causal):
            super().__init__()
            self.qkv, self.head, self.causal = nn.Linear(16, 48), nn.Linear(16, 3), causal

        def forward(self, x):
            q, k, v = self.qkv(x).view(x.shape[0], x.shape[1], 3, 2, 8).permute(2, 0, 3, 1, 4)
            o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=self.causal)
            return self.head(o.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], 16).mean(1))

    own = OwnAttention(False).eval()
    lp = capture(make_functional_call(own), dict(own.named_parameters()), torch.rand(2, 5, 16))
    assert [n["op"] for n in lp.nodes] == [capi.OP_INPUT, capi.OP_CONV, capi.OP_ATTENTION, capi.OP_AVGPOOL, capi.OP_CONV]
    bad = OwnAttention(True).eval()
    with pytest.raises(NotImplementedError, match="causal"):
        capture(make_functional_call(bad), dict(bad.named_parameters()), torch.rand(2, 5, 16))
