"""TEST INFRASTRUCTURE (manual script, not collected by pytest; it is the only checker that uses the oracle at full
size, hence it lives under tests/).  Run on the GPU box:  python tests/fullsize_parity_gpu.py [batch]

Full-size parity: ResNet-18, B x 3 x 224 x 224, GGN @ 1-2 columns.
Reference = the oracle's two-sweep restatement evaluated in float64 on the GPU (same algorithm as
oracle/curvature_oracle.py, device-agnostic torch ops).  Reports, per engine mode, the max error relative
to max|ref| and whether allclose(rtol=1e-4, atol=1e-5*max|ref|) holds; per-parameter worst offenders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator, _capi as capi
from oracle import curvature_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Kc = 2
torch.manual_seed(0)
dev = torch.device("cuda")
model = torchvision.models.resnet18().eval().to(dev)
X = torch.rand(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
P = sum(p.numel() for p in params.values())
V = torch.rand(P, Kc, device=dev)
loss = torch.nn.CrossEntropyLoss()
# fp64 reference on the GPU
m64 = torchvision.models.resnet18().eval().to(dev).double()
m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
p64 = dict(m64.named_parameters())
Vl = [v.reshape(*p.shape, Kc).double() for v, p in zip(V.split([p.numel() for p in p64.values()]), p64.values())]
ref = orc.ggn_matmat(m64, loss, p64, [(X.double(), y)], Vl)
ref = torch.cat([r.reshape(-1, Kc) for r in ref])
scale = ref.abs().max().item()
G = GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False)
for name, mode in [("simt fp32", 0), ("half-split tcgen05 (default)", 1),
                   ("half-split, split passes (no planes modes)", 0x401),
                   ("half-split, no N-stacked stem kernel", 0x801), ("3xTF32 tcgen05", 0x201)]:
    capi.lib().curv_set_tensor_core_mode(mode)
    got = (G @ V).double()
    err = (got - ref).abs()
    ok = torch.allclose(got, ref, rtol=1e-4, atol=1e-5 * scale)
    frac_bad = (~torch.isclose(got, ref, rtol=1e-4, atol=1e-5 * scale)).double().mean().item()
    print(f"[{name:44s}] max|err|/max|ref| = {err.max().item() / scale:.3e}  allclose(1e-4) = {ok}  "
          f"violations = {frac_bad:.2e}")
    o = 0
    worst = []
    for n, p in params.items():
        e = err[o:o + p.numel()].max().item(); r = ref[o:o + p.numel()].abs().max().item()
        worst.append((e / max(r, 1e-30), n, e, r)); o += p.numel()
    for w in sorted(worst, reverse=True)[:4]:
        print(f"      {w[1]:32s} max|err|={w[2]:.3e} max|ref|={w[3]:.3e} ratio={w[0]:.2e}")
capi.lib().curv_set_tensor_core_mode(1)
