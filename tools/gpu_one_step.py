"""One C2 step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off ...` captures (GPU box):
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/gpu_one_step.py
  ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_hs" -c 12 -o R python tools/gpu_one_step.py
Launches eagerly (no CUDA graph) so that every kernel is a separate profiled launch."""
import os, sys
os.environ["CURV_CUDA_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator

B, K = int(os.environ.get("CURV_B", 128)), 8
torch.manual_seed(0)
dev = torch.device("cuda")
dt = torch.bfloat16 if os.environ.get("CURV_DTYPE") == "bf16" else torch.float32  # CURV_DTYPE=bf16: the c2-bf16 step
model = torchvision.models.resnet18().eval().to(dev).to(dt)
X, y = torch.rand(B, 3, 224, 224, device=dev).to(dt), torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
V = torch.rand(sum(p.numel() for p in params.values()), K, device=dev).to(dt)
G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False, num_data=B)
out = G @ V  # warm-up (one-time initialisation)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = G @ V
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("checksum", float(out.float().abs().sum()))
