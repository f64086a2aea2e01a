"""C3 (BASELINE.json configs[2]) on the GPU box: KFAC factor build, damped inverse, inverse apply on ResNet-18,
128 x 3 x 224 x 224, Conv2d / Linear parameters, joint bias, MC Fisher (1 sample) -- wall times (CUDA events) and a
per-kernel breakdown of one factor build (CUPTI via torch.profiler).  usage: python tools/gpu_kfac_profile.py [bf16|f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from torch.profiler import profile, ProfilerActivity
from curvlinops_b200 import KFACLinearOperator

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
B = int(os.environ.get("CURV_B", 128))
torch.manual_seed(0)
dev = torch.device("cuda")
model = torchvision.models.resnet18().eval().to(dev).to(dt)
X, y = torch.rand(B, 3, 224, 224, device=dev).to(dt), torch.randint(0, 1000, (B,), device=dev)
mods = dict(model.named_modules())
names = [n for n, m in mods.items() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
params = {f"{n}.{pn}": p for n in names for pn, p in mods[n].named_parameters(recurse=False)}
P = sum(p.numel() for p in params.values())
loss = torch.nn.CrossEntropyLoss()


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), r


build = lambda: KFACLinearOperator(model, loss, params, [(X, y)], fisher_type="mc", mc_samples=1,
                                   separate_weight_and_bias=False, check_deterministic=False, num_data=B)
build()
t_build, Kop = timed(build)
t_inv, Kinv = timed(lambda: Kop.inverse(damping=1e-3))
v = torch.rand(P, device=dev).to(dt)
Kinv @ v
t_apply, _ = timed(lambda: Kinv @ v, reps=10)
fl = 0.0
for n in names:
    m = mods[n]
    if isinstance(m, torch.nn.Conv2d):
        ho = (224 if n == "conv1" else None)
fl_total = 3.81e12 * B / 128
print(f"# C3 {dt} B={B}: factor build {t_build:.2f} ms ({fl_total / t_build / 1e9:.0f} TFLOP/s of the 3.81 TFLOP Gram work), "
      f"damped inverse (cuSOLVER) {t_inv:.2f} ms, inverse apply K=1 {t_apply:.3f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    build()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# one factor build: sum of kernel times {tot:.2f} ms over {sum(r[1] for r in rows)} launches")
for k, n, ms in rows[:25]:
    print(f"{100*ms/tot:6.2f}% {ms:9.3f} ms  n={n:4d}  {k[:100]}")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    Kinv @ v
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# one inverse apply: sum of kernel times {tot:.3f} ms over {sum(r[1] for r in rows)} launches")
for k, n, ms in rows[:12]:
    print(f"{100*ms/tot:6.2f}% {ms:9.3f} ms  n={n:4d}  {k[:100]}")

# ---- EKFAC: eigendecomposition (torch.linalg.eigh, library) + eigenvalue correction (csrc/ekfac.cuh)
if os.environ.get("CURV_EKFAC", "1") != "0":
    from curvlinops_b200 import EKFACLinearOperator
    from curvlinops_b200.kfac import EKFACComputer

    comp = EKFACComputer(model, loss, params, [(X, y)], fisher_type="mc", mc_samples=1, separate_weight_and_bias=False,
                         check_deterministic=False, num_data=B)
    A, G, mapping = super(EKFACComputer, comp).compute()
    f32 = lambda v_: getattr(v_, "_curv_fp32", v_).float()
    t_eigh, (QA, QG) = timed(lambda: ({k: torch.linalg.eigh(f32(v_)).eigenvectors for k, v_ in A.items()},
                                      {k: torch.linalg.eigh(f32(v_)).eigenvectors for k, v_ in G.items()}), reps=1)
    comp._eigenvalue_correction(QA, QG, mapping)
    t_corr, lam = timed(lambda: comp._eigenvalue_correction(QA, QG, mapping), reps=2)
    print(f"# EKFAC {dt} B={B}: eigh of the 42 factors (cuSOLVER) {t_eigh:.1f} ms, eigenvalue correction {t_corr:.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        comp._eigenvalue_correction(QA, QG, mapping)
        torch.cuda.synchronize()
    rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print(f"# one eigenvalue correction: sum of kernel times {tot:.2f} ms over {sum(r[1] for r in rows)} launches")
    for k, n, ms in rows[:14]:
        print(f"{100*ms/tot:6.2f}% {ms:9.3f} ms  n={n:4d}  {k[:100]}")
