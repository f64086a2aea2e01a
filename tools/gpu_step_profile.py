"""Per-kernel time breakdown of ONE C2 step (GPU box) via CUPTI activity records (torch.profiler): every kernel
on the device is recorded, including the library's own launches.  usage: python tools/gpu_step_profile.py [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from torch.profiler import profile, ProfilerActivity
from curvlinops_b200 import GGNLinearOperator, HessianLinearOperator, _capi as capi

mode = int(sys.argv[1], 0) if len(sys.argv) > 1 else 1
B, K = int(os.environ.get("CURV_B", 128)), int(os.environ.get("CURV_K", 8))
torch.manual_seed(0)
dev = torch.device("cuda")
dt = torch.bfloat16 if os.environ.get("CURV_DTYPE") == "bf16" else torch.float32
model = getattr(torchvision.models, os.environ.get("CURV_MODEL", "resnet18"))().eval().to(dev).to(dt)
X, y = torch.rand(B, 3, 224, 224, device=dev).to(dt), torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
P = sum(p.numel() for p in params.values())
V = torch.rand(P, K, device=dev).to(dt)
Op = HessianLinearOperator if os.environ.get('CURV_OP') == 'hessian' else GGNLinearOperator
G = Op(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False, num_data=B)
capi.lib().curv_set_tensor_core_mode(mode)
for _ in range(2):
    out = G @ V
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = G @ V
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# one step ({os.environ.get('CURV_MODEL', 'resnet18')}, B={B}, K={K}, {dt}), mode={mode:#x}: sum of kernel times {tot:.2f} ms over {sum(r[1] for r in rows)} launches")
for k, n, ms in rows[:40]:
    print(f"{100*ms/tot:6.2f}% {ms:9.3f} ms  n={n:4d}  {k[:110]}")
# per-launch durations of the contraction kernels, in launch order (layer attribution by position)
evs = [e for e in prof.events() if e.device_time_total > 0 and ("gemm" in e.name)]
evs.sort(key=lambda e: e.time_range.start)
print("# contraction launches in order: name, ms")
for e in evs:
    nm = "gather128" if "hs<128>" in e.name else "gather64" if "hs<64>" in e.name else "wgrad_hs" if "wgrad_gemm_hs" in e.name else e.name[:40]
    print(f"{nm:12s} {e.device_time_total / 1e3:7.3f}")
