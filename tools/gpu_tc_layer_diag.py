"""Per-layer comparison of the tcgen05 gather GEMM against the fp32 SIMT kernels (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator, _capi as capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
dev = torch.device("cuda")
model = torchvision.models.resnet18().eval().to(dev)
X = torch.rand(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
V = torch.rand(sum(p.numel() for p in params.values()), 1, device=dev)
G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False)
eng = G._engine
snaps = {}
for mode in (0, 1 | (2 << 4)):
    capi.lib().curv_set_tensor_core_mode(mode)
    eng.predict(X)  # primal forward only
    prog = eng.program(X, 1, False)
    ws = eng._ws
    snaps[mode] = [prog.value_view(ws, n["out"], 1)[0].clone() for n in prog.lp.nodes]
capi.lib().curv_set_tensor_core_mode(1)
a, b = snaps[0], snaps[1 | (2 << 4)]
for n, x, z in zip(prog.lp.nodes, a, b):
    sc = x.abs().max().item()
    print(f"op={n['op']} out={n['out']:3d} shape={tuple(x.shape)} max|diff|/max = {(x - z).abs().max().item() / max(sc, 1e-30):.3e}")
