"""Diagnostic (run on the GPU box): per-value forward check of the engine against torch hooks, and
per-parameter GGN errors, so that one gpurun call localises a broken kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from curvlinops_b200 import GGNLinearOperator, _capi as capi
from curvlinops_b200.engine import Engine
from curvlinops_b200.curvature import make_functional_call
from tests.golden_utils import load_case

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "miniresnet_ce_mean"
model, loss, data, fx = load_case(name, dtype=torch.float32, device="cuda")
params = dict(model.named_parameters())
X, y = data[0]
eng = Engine(make_functional_call(model), loss, params)
pred = eng.predict(X)
ref = model(X)
print("prediction max err", (pred - ref).abs().max().item(), "scale", ref.abs().max().item())
prog = eng.program(X, 1, False)
# compare each engine value with an eager re-execution of the lowered program in torch
import torch.nn.functional as F
lp = prog.lp
vals = {}
ws = eng._ws
names = list(params.keys())
plist = list(params.values())
for n in lp.nodes:
    op = n["op"]
    if op == capi.OP_INPUT:
        vals[n["out"]] = X
    elif op == capi.OP_CONV:
        w = plist[n["p0"]] if n["p0"] >= 0 else prog.consts[n["c0"]]
        b = plist[n["p1"]] if n["p1"] >= 0 else (prog.consts[n["c1"]] if n["c1"] >= 0 else None)
        xin = vals[n["in0"]]
        w4 = w.reshape(w.shape[0], xin.shape[1], n["kh"], n["kw"])
        vals[n["out"]] = F.conv2d(xin, w4, b, (n["sh"], n["sw"]), (n["ph"], n["pw"]))
    elif op == capi.OP_AFFINE:
        g = plist[n["p0"]] if n["p0"] >= 0 else (prog.consts[n["c0"]] if n["c0"] >= 0 else None)
        b = plist[n["p1"]] if n["p1"] >= 0 else (prog.consts[n["c1"]] if n["c1"] >= 0 else None)
        vals[n["out"]] = F.batch_norm(vals[n["in0"]], prog.consts[n["c2"]], prog.consts[n["c3"]], g, b, False, 0.0, n["eps"])
    elif op == capi.OP_RELU:
        vals[n["out"]] = vals[n["in0"]].relu()
    elif op == capi.OP_ADD:
        vals[n["out"]] = vals[n["in0"]] + vals[n["in1"]]
    elif op == capi.OP_MAXPOOL:
        vals[n["out"]] = F.max_pool2d(vals[n["in0"]], (n["kh"], n["kw"]), (n["sh"], n["sw"]), (n["ph"], n["pw"]))
    elif op == capi.OP_AVGPOOL:
        vals[n["out"]] = vals[n["in0"]].mean((2, 3), keepdim=True)
    v = vals[n["out"]]
    if v.ndim == 2:
        v = v[:, :, None, None]
        vals[n["out"]] = v
    got = prog.value_view(ws, n["out"], 1)[0][..., : v.shape[1]].permute(0, 3, 1, 2)
    print(f"node op={op} out={n['out']} shape={tuple(v.shape)} max err {(got - v).abs().max().item():.3e} scale {v.abs().max().item():.3e}")

G = GGNLinearOperator(model, loss, params, data, check_deterministic=False)
V = fx["V"].float().cuda()
got = (G @ V).double().cpu()
refg = fx["ggn"]
o = 0
for nme, p in params.items():
    g, r = got[o:o + p.numel()], refg[o:o + p.numel()]
    print(f"GGN {nme:30s} max|ref|={r.abs().max():.3e} max|err|={(g - r).abs().max():.3e}")
    o += p.numel()
