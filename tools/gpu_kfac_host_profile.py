"""Host-side cost of one KFAC factor build at C3 shapes (GPU box): wall time vs kernel time, CUDA API calls by time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda")
model, X, y, params = bench.kfac_problem(torch, 128, torch.bfloat16, dev)
from curvlinops_b200 import KFACLinearOperator
kw = dict(fisher_type="mc", mc_samples=1, separate_weight_and_bias=False, check_deterministic=False, num_data=128)
build = lambda: KFACLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], **kw)
for _ in range(3):
    build()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    build()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per build {1e3 * (t1 - t0) / 5:.2f} ms; incl. final sync {1e3 * (t2 - t0) / 5:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    build()
    torch.cuda.synchronize()
ka = prof.key_averages()
dev_tot = sum(e.self_device_time_total for e in ka) / 1e3
print(f"sum of kernel time {dev_tot:.2f} ms")
rows = sorted(((e.self_cpu_time_total / 1e3, e.count, e.key) for e in ka), reverse=True)
for ms, n, k in rows[:25]:
    print(f"{ms:8.3f} ms  n={n:5d}  {k[:90]}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); build(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
