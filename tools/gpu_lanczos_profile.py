"""Where a Lanczos step on the bf16 ResNet-50 GGN spends its time (GPU box): kernel time by name and host-side op time."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from torch.profiler import profile, ProfilerActivity
from curvlinops_b200 import GGNLinearOperator
from curvlinops_b200.lanczos import lanczos_eigsh

warnings.simplefilter("ignore")
dev = torch.device("cuda")
torch.manual_seed(0)
B = int(os.environ.get("CURV_B", 64))
model = torchvision.models.resnet50().eval().to(torch.bfloat16).to(dev)
X, y = torch.rand(B, 3, 224, 224, device=dev).to(torch.bfloat16), torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False, num_data=B)
run = lambda: lanczos_eigsh(G, k=10, ncv=30, maxiter=30, tol=0.0, return_info=True)
run()
torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print(f"wall {1e3 * (time.perf_counter() - t0) / 30:.2f} ms per step")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(((e.device_time_total / 1e3, e.count, e.key) for e in ka if e.device_time_total > 0), reverse=True)
tot = sum(e.self_device_time_total for e in ka) / 1e3
print(f"# device: sum of self kernel time {tot / 30:.2f} ms per step")
for ms, n, k in rows[:25]:
    print(f"{ms / 30:8.3f} ms/step  n={n:5d}  {k[:100]}")
rows = sorted(((e.self_cpu_time_total / 1e3, e.count, e.key) for e in ka), reverse=True)
print("# host (self CPU time)")
for ms, n, k in rows[:15]:
    print(f"{ms / 30:8.3f} ms/step  n={n:5d}  {k[:100]}")
