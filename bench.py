#!/usr/bin/env python
"""Benchmark of the hot path: GGN matrix-matrix product on ResNet-18 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one ``G @ V`` over one synthetic mini-batch (random-init torchvision ResNet-18 in eval mode,
X = rand(128, 3, 224, 224), y = randint(1000), V = rand(P, 8), fp32, CrossEntropyLoss-mean).
``value`` = P*K / t_step (param-dim x vectors per second), inputs resident in HBM.  With N > 1 the
mini-batch is sharded over the ranks (strong scaling: total work fixed) and the [P, K] result is summed
with one NCCL all-reduce.  ``e2e`` is the same product through the public operator API with HOST
buffers (``matmat_pinned``: pinned X, y and V copied to the device and the result copied back inside the timed
region, V / result pipelined in parameter buckets against the sweeps).
``--impl reference`` times the UNMODIFIED reference (``oracle/_ref``, copied from /root/reference by
``oracle/build_ref.py``; git-ignored, travels with the snapshot) on the host cores at full size.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ggn_matvec_throughput"  # ("hessian_matvec_throughput" for --config c2-hessian)
UNIT = "param*vec/s"
B, K = 128, 8
#: --config: c2 = BASELINE.json configs[1] (the configuration the metric is quoted on; default);
#: c2-bf16 = the same workload as a bf16 operator (bf16 parameters / data / vectors, the dtype of configs[2:])
WORKLOADS = {
    "c2": ("ResNet-18 random-init, synthetic 128x3x224x224, GGNLinearOperator @ 8 vectors, fp32", "f32"),
    "c2-bf16": ("ResNet-18 random-init, synthetic 128x3x224x224, GGNLinearOperator @ 8 vectors, bf16", "bf16"),
    # the C2 workload with the full Hessian (hand-written R-op, 3xTF32 tensor-core kernels) instead of the GGN
    "c2-hessian": ("ResNet-18 random-init, synthetic 128x3x224x224, HessianLinearOperator @ 8 vectors, fp32", "f32"),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case; pure latency (6.6 MFLOP per product)
    "c1": ("3-layer MLP (D=64, 4 Linear+ReLU, CE loss), batch=32, HessianLinearOperator @ 1 vector", "f32"),
    # BASELINE.json configs[2]: its own metric (a step = one factor build), printed as an extra line
    "c3": ("ResNet-18 random-init, synthetic 128x3x224x224, KFACLinearOperator factor build (Conv2d/Linear params, "
           "joint bias, MC Fisher 1 sample) + damped inverse 1e-3 + inverse apply @ 1 vector, bf16", "bf16"),
    # BASELINE.json configs[3]: ViT-B/16, MC-GGN (1 sample), 4 vectors, bf16, 32 samples per GPU (256 on 8)
    "c4": ("ViT-B/16 random-init, synthetic (32 per GPU)x3x224x224, GGNLinearOperator(mc_samples=1) @ 4 vectors, bf16",
           "bf16"),
    # BASELINE.json configs[4]: its own metric (matvecs/s of the GGN under a Lanczos eigensolver), printed as an extra line
    "c5": ("ResNet-50 random-init, synthetic (64 per GPU)x3x224x224, GGNLinearOperator matvecs under Lanczos "
           "eigsh(k=10), bf16", "bf16"),
}
CONFIG = "c2"
WORKLOAD = WORKLOADS[CONFIG][0]


def build_problem(torch, batch):
    import torchvision

    torch.manual_seed(0)
    model = torchvision.models.resnet18().eval()
    X = torch.rand(batch, 3, 224, 224)
    y = torch.randint(0, 1000, (batch,))
    if WORKLOADS[CONFIG][1] == "bf16":
        model, X = model.to(torch.bfloat16), X.to(torch.bfloat16)
    return model, X, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic_per_launch():
    """Mean dram__bytes_read + dram__bytes_write per launch of the gather GEMM kernels, from the committed
    `ncu --set full` capture of the first contraction launches of one step of this configuration
    (profiles/r2_hs_kernels_ncu_full_{fp32,bf16}.txt, made by tools/r2_gpu_job.sh with tools/gpu_one_step.py in the
    same commit as the kernels; a profiler cannot run inside the timed bench)."""
    import re

    path = os.path.join(ROOT, "profiles", "r2_hs_kernels_ncu_full_%s.txt" % ("bf16" if WORKLOADS[CONFIG][1] == "bf16"
                                                                              else "fp32"))
    try:
        blocks = open(path).read().split("---")
    except OSError:
        return None
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot, n = 0.0, 0
    for b in blocks:
        if "gather_gemm_hs<" not in b:
            continue
        vals = [re.search(name + r"\s+([\d.]+)\s+(\w+)", b) for name in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
        if all(vals):
            tot += sum(float(v.group(1)) * unit.get(v.group(2), 1.0) for v in vals)
            n += 1
    return tot / n if n else None


def config_dict(n_gpus):
    """The workload both arms are measured on (identical dict in the product and the reference line)."""
    return {"workload": WORKLOAD, "params": 11689512, "columns": K, "batch": B,
            "parallelism": f"dp{n_gpus} (mini-batch sharded over the ranks, one all-reduce of [P,K])",
            "l2": "working set (GBs of activations) far exceeds the 126 MB L2; no flush needed"}


def host_threads():
    """Threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1,
    which would otherwise pin the reference to one thread)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def reference_ggn(torch, device, batch, k, repeats, warmup=0):
    """The UNMODIFIED reference (`oracle/_ref`, built by oracle/build_ref.py) through its own public API:
    ``GGNLinearOperator(model, CrossEntropyLoss(), params, [(X, y)]) @ V`` on `device`, same seeds / inputs as the
    GPU arm (benchmark protocol of docs/examples/basic_usage/benchmark_execute.py:287-302: synchronize on both sides,
    minimum over the repeats).  Returns (seconds per product, P)."""
    from oracle.build_ref import import_reference

    ref = import_reference()
    model, X, y = build_problem(torch, batch)
    model = model.to(device)
    X, y = X.to(device), y.to(device)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    torch.manual_seed(1)
    V = torch.rand(P, k).to(device).to(next(iter(params.values())).dtype)
    RefOp = ref.HessianLinearOperator if CONFIG == "c2-hessian" else ref.GGNLinearOperator
    G = RefOp(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False, num_data=batch)
    sync = torch.cuda.synchronize if device.type == "cuda" else (lambda: None)
    times = []
    for i in range(warmup + repeats):
        sync()
        t0 = time.perf_counter()
        G @ V
        sync()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return min(times), P


def cpu_reference_run(torch, repeats, sample_batch=B, sample_k=K):
    """Reference on the host cores.  Full size by default (`--impl reference`); the GPU arm's `cpu_baseline` leg uses a
    bounded sample of the mini-batch (samples are independent, so the cost is linear in the batch) with all K
    columns."""
    torch.set_num_threads(host_threads())
    global CONFIG
    cfg, CONFIG = CONFIG, ("c2" if CONFIG == "c2-bf16" else CONFIG)  # bf16 is timed in fp32 on the CPU (CPU bf16
    try:                        # convolutions are not representative; the reference's inverse cannot run in bf16)
        t, P = reference_ggn(torch, torch.device("cpu"), sample_batch, sample_k, repeats)
    finally:
        CONFIG = cfg
    t_full = t * (B / sample_batch) * (K / sample_k)
    full = sample_batch == B and sample_k == K
    return P * K / t_full, t_full, {
        "kind": "reference", "cores": torch.get_num_threads(),
        "sample": (f"full step (B={B}, K={K}), best of {repeats}" if full else
                   f"{sample_batch} of {B} samples, {sample_k} of {K} columns, best of {repeats}, "
                   "extrapolated linearly to the full step") + (", fp32 on the CPU" if cfg == "c2-bf16" else "")}


def gpu_library_baseline(torch, dev, repeats=3):
    """Second bar: the unmodified reference on THIS B200 through torch CUDA (cuDNN / cuBLAS,
    eager autograd), full size, strict fp32 and with TF32 allowed."""
    out = {"what": "unmodified reference GGNLinearOperator @ V on the same GPU via torch CUDA (eager), "
                   f"B={B}, K={K}, min of {repeats} after 1 warm-up", "unit": UNIT}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, tf32 in ((("bf16", True),) if WORKLOADS[CONFIG][1] == "bf16" else (("fp32", False), ("tf32", True))):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            t, P = reference_ggn(torch, dev, B, K, repeats, warmup=1)
            out[name] = {"ms_per_step": t * 1e3, "value": P * K / t}
            torch.cuda.empty_cache()
    except Exception as e:  # reported, never fatal: the product arm has already been measured
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def kfac_problem(torch, batch, dtype, device):
    model, X, y = build_problem(torch, batch)
    model, X, y = model.to(device).to(dtype), X.to(device).to(dtype), y.to(device)
    mods = dict(model.named_modules())
    names = [n for n, m in mods.items() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    params = {f"{n}.{pn}": p for n in names for pn, p in mods[n].named_parameters(recurse=False)}
    return model, X, y, params


def run_c1(torch, args):
    """C1: Hessian-vector product of the 13 130-parameter MLP, one vector: latency (CUDA-graph replay of the R-op
    sweeps) next to the reference on the host cores (its own protocol: minimum of 10, benchmark_execute.py:287-302)."""
    from curvlinops_b200 import HessianLinearOperator, _capi as capi
    from oracle.models import mlp_c1

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = mlp_c1()
    X, y = torch.randn(32, 64), torch.randint(0, 10, (32,))
    loss = torch.nn.CrossEntropyLoss()
    md = mlp_c1().to(dev)
    md.load_state_dict(model.state_dict())
    params = dict(md.named_parameters())
    P = sum(p.numel() for p in params.values())
    H = HessianLinearOperator(md, loss, params, [(X.to(dev), y.to(dev))], check_deterministic=False)
    torch.manual_seed(1)
    v_host = torch.rand(P).pin_memory()
    v = v_host.to(dev)
    for _ in range(max(3, args.warmup) + 2):
        H @ v
    L0 = capi.lib().curv_launch_count()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(args.steps, 50)
    e0.record()
    for _ in range(steps):
        H @ v
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = capi.lib().curv_launch_count() - L0
    t0 = time.perf_counter()
    for _ in range(steps):
        out = (H @ v_host.to(dev, non_blocking=True)).cpu()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    line = {
        "metric": "hessian_matvec_latency", "value": ms, "unit": "ms", "n_gpus": 1, "steps": steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "params": P, "columns": 1, "batch": 32,
                   "l2": "latency-bound: the whole problem (13 130 parameters, 32 x 64 inputs) sits in L2; the product is "
                         "one CUDA-graph replay of ~60 small launches"},
        "gpu_launches": int(launches), "param_vec_per_s": P / (ms / 1e3),
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": P * 4, "d2h_bytes_per_step": P * 4,
                "what": "host vector in, host result out (wall clock, includes the synchronising download)"},
        "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "6.6 MFLOP and ~0.2 MB per product: launch / latency territory, no roofline applies "
                             "(SURVEY 8d)"},
    }
    if not args.no_cpu_baseline:
        from oracle.build_ref import import_reference

        ref = import_reference()
        torch.set_num_threads(host_threads())
        Hr = ref.HessianLinearOperator(model, loss, dict(model.named_parameters()), [(X, y)], check_deterministic=False)
        vr = torch.rand(P)
        Hr @ vr
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            Hr @ vr
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": min(ts) * 1e3, "unit": "ms", "cores": torch.get_num_threads(),
                                "kind": "reference", "sample": "the full C1 product, minimum of 10 after 1 warm-up"}
    print(json.dumps(line))


def run_c3(torch, args):
    """C3: KFAC factor build (value), damped inverse, inverse apply; reference times beside (BASELINE.md section 4:
    19.1 s / 4.43 s / 0.252 s on 8 CPU cores; the reference cannot invert in bf16, so its arms run in fp32)."""
    import ctypes as C

    from curvlinops_b200 import KFACLinearOperator, _capi as capi

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model, X, y, params = kfac_problem(torch, B, torch.bfloat16, dev)
    P = sum(p.numel() for p in params.values())
    loss = torch.nn.CrossEntropyLoss()
    kw = dict(fisher_type="mc", mc_samples=1, separate_weight_and_bias=False, check_deterministic=False, num_data=B)
    build = lambda: KFACLinearOperator(model, loss, params, [(X, y)], **kw)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, r

    for _ in range(max(3, args.warmup)):
        build()
    sampler = ClockSampler(0)
    sampler.start()
    L0 = capi.lib().curv_launch_count()
    ms_build, Kop = timed(build, args.steps)
    launches = capi.lib().curv_launch_count() - L0
    clocks = sampler.stop()
    Kop.inverse(damping=1e-3)  # first call initialises cuSOLVER
    ms_inv, Kinv = timed(lambda: Kop.inverse(damping=1e-3), 3)
    torch.manual_seed(1)
    v_host = torch.rand(P).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(P, dtype=torch.bfloat16).pin_memory()
    v = v_host.to(dev)
    Kinv @ v
    ms_apply, _ = timed(lambda: Kinv @ v, 10)
    # per-launch timing of the Gram contractions (class 1 = wgrad-type kernel) of one build
    capi.lib().curv_profile_enable(1)
    build()
    torch.cuda.synchronize()
    ms, fl, cnt = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
    capi.lib().curv_profile_read(ms, fl, cnt)
    capi.lib().curv_profile_enable(0)
    X_host, y_host = X.cpu().pin_memory(), y.cpu().pin_memory()

    def e2e():  # host data in, preconditioned host vector out
        Xd, yd = X_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)
        Kd = KFACLinearOperator(model, loss, params, [(Xd, yd)], **kw).inverse(damping=1e-3)
        out_host.copy_(Kd @ v_host.to(dev, non_blocking=True), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e()
    ms_e2e, _ = timed(e2e, 2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    gram_tf = (fl[1] / 1e12) / (ms[1] / 1e3) if ms[1] > 0 else 0.0
    out = {
        "metric": "kfac_factor_build_time", "value": ms_build, "unit": "ms", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_build, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "params": P, "batch": B, "factors": 42,
                   "l2": "activations and patch planes (GBs) far exceed the 126 MB L2; no flush needed"},
        "clocks": clocks, "gpu_launches": int(launches),
        "phases_ms": {"factor_build": ms_build, "damped_inverse_cusolver": ms_inv, "inverse_apply_1_vector": ms_apply,
                      "note": "factor build and apply are this repo's kernels; the Cholesky factorisation / inversion "
                              "of the 42 factors is torch.linalg (cuSOLVER), a stated library call"},
        "e2e": {"value": ms_e2e, "unit": "ms", "what": "host X, y, v -> factor build -> damped inverse -> apply -> host",
                "h2d_bytes_per_step": int(X_host.numel() * 2 + y_host.numel() * 8 + P * 2),
                "d2h_bytes_per_step": int(P * 2)},
        "roofline": {"bound": "tensor", "kernel": "wgrad_gemm_hs<1> as Gram kernel (A and G factors)",
                     "achieved": gram_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gram_tf / peak_tf, "traffic": None,
                     "launches_timed": int(cnt[1]), "share_of_step": ms[1] / ms_build,
                     "note": "algorithmic Gram FLOPs 2 * rows * width^2 per factor (both triangles are computed: "
                             "3.81 TFLOP per build, SURVEY 8d) / CUDA-event time of the launches",
                     "step": {"algorithmic_tflop": 3.81, "tflops": 3.81 / (ms_build / 1e3)}},
    }
    if not args.no_cpu_baseline:
        from oracle.build_ref import import_reference

        ref = import_reference()
        rkw = dict(fisher_type=ref.FisherType.MC, mc_samples=1, separate_weight_and_bias=False,
                   check_deterministic=False, num_data=B)
        try:  # second bar: the unmodified reference on this GPU (fp32: it cannot invert bf16 factors)
            m32, X32, y32, p32 = kfac_problem(torch, B, torch.float32, dev)
            rb = lambda: ref.KFACLinearOperator(m32, loss, p32, [(X32, y32)], **rkw)
            rb()
            t_b, Kr = timed(rb, 2)
            t_i, Kri = timed(lambda: Kr.inverse(damping=1e-3), 2)
            vr = torch.rand(P, device=dev)
            Kri @ vr
            t_a, _ = timed(lambda: Kri @ vr, 5)
            out["gpu_library_baseline"] = {"what": "unmodified reference KFACLinearOperator on the same GPU (torch CUDA, "
                                                   "fp32, hooks backend)", "factor_build_ms": t_b,
                                           "damped_inverse_ms": t_i, "inverse_apply_ms": t_a}
            del Kr, Kri, m32, X32
            torch.cuda.empty_cache()
        except Exception as e:
            out["gpu_library_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.set_num_threads(host_threads())
        sb = 16  # bounded CPU sample: the factor build is linear in the batch, inverse / apply do not depend on it
        mc, Xc, yc, pc = kfac_problem(torch, sb, torch.float32, torch.device("cpu"))
        t0 = time.perf_counter()
        Kc = ref.KFACLinearOperator(mc, loss, pc, [(Xc, yc)], **{**rkw, "num_data": sb})
        t1 = time.perf_counter()
        Kci = Kc.inverse(damping=1e-3)
        t2 = time.perf_counter()
        vc = torch.rand(P)
        Kci @ vc
        t3 = time.perf_counter()
        Kci @ vc
        t4 = time.perf_counter()
        out["cpu_baseline"] = {"value": (t1 - t0) * 1e3 * B / sb, "unit": "ms", "cores": torch.get_num_threads(),
                               "kind": "reference",
                               "sample": f"factor build on {sb} of {B} samples (fp32 on the CPU), extrapolated linearly; "
                                         "inverse and apply at full size",
                               "damped_inverse_ms": (t2 - t1) * 1e3, "inverse_apply_ms": (t4 - t3) * 1e3}
    print(json.dumps(out))


def run_c4(torch, args):
    """C4: MC-GGN (one sample) of a bf16 ViT-B/16 applied to 4 vectors; weak scaling, 32 examples per GPU (256 on 8: the
    configuration's global batch); mini-batch sharded over the ranks, one all-reduce of [P, 4] per product."""
    import torch.distributed as dist
    import torchvision

    from curvlinops_b200 import GGNLinearOperator, _capi as capi
    from curvlinops_b200 import dist as cdist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cdist.enable(True)
    PB, KV = int(os.environ.get("CURV_C4_PER_GPU", 32)), 4
    GB = PB * world
    torch.manual_seed(0)
    model = torchvision.models.vit_b_16().eval()
    with torch.no_grad():  # torchvision zero-initialises the classification head: the GGN would vanish
        model.heads.head.weight.normal_(0.0, 0.02)
    model = model.to(torch.bfloat16).to(dev)
    X = torch.rand(GB, 3, 224, 224).to(torch.bfloat16)
    y = torch.randint(0, 1000, (GB,))
    X_host, y_host = X.pin_memory(), y.pin_memory()
    Xd, yd = X.to(dev), y.to(dev)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    loss = torch.nn.CrossEntropyLoss()
    kw = dict(check_deterministic=False, num_data=GB, mc_samples=1)
    G = GGNLinearOperator(model, loss, params, [(Xd, yd)], **kw)
    G_host = GGNLinearOperator(model, loss, params, [(X_host, y_host)], **kw)
    G_host._engine = G._engine
    torch.manual_seed(1)
    shm = None
    if world > 1:  # one copy of V / of the result in host memory for all ranks (shared, registered as pinned)
        shm = f"curvbench_c4_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            V_host = cdist.shared_pinned_tensor(shm + "_V", (P, KV), torch.bfloat16)
            V_host.copy_(torch.rand(P, KV).to(torch.bfloat16))
            out_host = cdist.shared_pinned_tensor(shm + "_out", (P, KV), torch.bfloat16)
        dist.barrier()
        if rank != 0:
            V_host = cdist.shared_pinned_tensor(shm + "_V", (P, KV), torch.bfloat16)
            out_host = cdist.shared_pinned_tensor(shm + "_out", (P, KV), torch.bfloat16)
    else:
        V_host = torch.rand(P, KV).to(torch.bfloat16).pin_memory()
        out_host = torch.empty(P, KV, dtype=torch.bfloat16).pin_memory()
    Vd = V_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def step_e2e():  # host X / y (uploaded inside), host V in, host result out: the public host-operand API
        G_host.matmat_pinned(V_host, out_host)
        torch.cuda.current_stream().synchronize()

    t0 = time.perf_counter()
    G @ Vd
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    for _ in range(max(3, args.warmup)):
        G @ Vd
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    L0 = capi.lib().curv_launch_count()
    ms_step = timed(lambda: G @ Vd, args.steps)
    launches = capi.lib().curv_launch_count() - L0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(3):  # first call eager, second captures its CUDA graph, third replays
        step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 3)))
    # repeatability (same seed -> same draws) and positive semi-definiteness on the timed operator
    a, b = G @ Vd[:, :1], G @ Vd[:, :1]
    same = bool(torch.equal(a, b))
    quad = float((Vd[:, 0].float() * a[:, 0].float()).sum())
    barrier()
    ref_out = (G @ Vd).float()
    e2e_check = float((out_host.to(dev).float() - ref_out).abs().max() / ref_out.abs().max())
    barrier()
    if world > 1 and rank == 0:
        cdist.release_shared(shm + "_V")
        cdist.release_shared(shm + "_out")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "ggn_matvec_throughput", "value": P * KV / (ms_step / 1e3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "params": P, "columns": KV, "global_batch": GB, "mc_samples": 1,
                   "parallelism": f"dp{world} (mini-batch sharded over the ranks, one all-reduce of [P,4] per product)",
                   "l2": "activations (GBs) far exceed the 126 MB L2; no flush needed"},
        "clocks": clocks, "gpu_launches": int(launches),
        "first_product_s": t_first,
        "e2e": {"value": P * KV / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(X_host.numel() * 2 + y_host.numel() * 8 + V_host.numel() * 2),
                "d2h_bytes_per_step": int(P * KV * 2), "vs_resident_max_rel_diff": e2e_check,
                "note": "matmat_pinned: V / result pipelined against the sweeps; with N > 1 they live once in shared "
                        "pinned host memory and every rank moves its 1/N row block (byte counts are per job)"},
        "self_check": {"repeatable": same, "vT_G_v": quad},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.build_ref import import_reference

        ref = import_reference()
        try:  # the unmodified reference on this GPU (needs train() with p = 0 and the math attention path: SURVEY 8c)
            from torch.nn.attention import SDPBackend, sdpa_kernel

            model.train()
            with sdpa_kernel(SDPBackend.MATH):
                Gr = ref.GGNLinearOperator(model, loss, params, [(Xd, yd)], check_deterministic=False, num_data=GB,
                                           mc_samples=1)
                Gr @ Vd
                t_r = timed(lambda: Gr @ Vd, 2)
            out["gpu_library_baseline"] = {"what": "unmodified reference GGNLinearOperator(mc_samples=1) @ V on the same GPU "
                                                   "(torch CUDA, bf16, math attention path)", "ms": t_r,
                                           "value": P * KV / (t_r / 1e3), "unit": UNIT}
            del Gr
        except Exception as e:
            out["gpu_library_baseline"] = {"bf16_error": f"{type(e).__name__}: {e}"[:300]}
            try:  # the reference does not run this configuration in bf16 here: time it in fp32 with TF32 allowed
                from torch.nn.attention import SDPBackend, sdpa_kernel

                torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
                m32 = torchvision.models.vit_b_16().train().to(dev)
                p32 = dict(m32.named_parameters())
                with sdpa_kernel(SDPBackend.MATH):
                    Gr = ref.GGNLinearOperator(m32, loss, p32, [(Xd.float(), yd)], check_deterministic=False,
                                               num_data=GB, mc_samples=1)
                    V32 = Vd.float()
                    Gr @ V32
                    t_r = timed(lambda: Gr @ V32, 2)
                out["gpu_library_baseline"].update({
                    "what": "unmodified reference GGNLinearOperator(mc_samples=1) @ V on the same GPU (torch CUDA, fp32 "
                            "parameters with TF32 matmuls allowed, math attention path)", "ms": t_r,
                    "value": P * KV / (t_r / 1e3), "unit": UNIT})
                del Gr, m32
            except Exception as e2:
                out["gpu_library_baseline"]["fp32_error"] = f"{type(e2).__name__}: {e2}"[:300]
        model.eval()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_c5(torch, args):
    """C5: GGN of a bf16 ResNet-50 feeding a Lanczos eigensolver (k = 10); metric matvecs/s.  Weak scaling: 64 examples
    per GPU (512 on 8, the configuration's global batch), the mini-batch sharded over the ranks, one all-reduce of the
    [P] result per product.  `value` = device-resident products per second inside `lanczos_eigsh` (products + the
    fp32 Lanczos vector algebra); `e2e` = the reference's own usage, `scipy.sparse.linalg.eigsh(G.to_scipy(), k=10)`:
    every product takes a host NumPy vector and returns one (host<->device copies inside)."""
    import numpy
    import torch.distributed as dist
    import torchvision
    from scipy.sparse.linalg import ArpackNoConvergence, eigsh

    from curvlinops_b200 import GGNLinearOperator, _capi as capi
    from curvlinops_b200 import dist as cdist
    from curvlinops_b200.lanczos import lanczos_eigsh

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cdist.enable(True)
    PB = int(os.environ.get("CURV_C5_PER_GPU", 64))
    GB = PB * world
    torch.manual_seed(0)
    model = torchvision.models.resnet50().eval().to(torch.bfloat16).to(dev)
    X = torch.rand(GB, 3, 224, 224).to(torch.bfloat16)
    y = torch.randint(0, 1000, (GB,))
    X_host, y_host = X.pin_memory(), y.pin_memory()
    Xd, yd = X.to(dev), y.to(dev)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    loss = torch.nn.CrossEntropyLoss()
    G = GGNLinearOperator(model, loss, params, [(Xd, yd)], check_deterministic=False, num_data=GB)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, r

    torch.manual_seed(1)
    v = torch.rand(P, device=dev).to(torch.bfloat16)
    if world > 1:
        dist.broadcast(v, 0)
    for _ in range(max(3, args.warmup)):
        G @ v
    ms_mv, _ = timed(lambda: G @ v, max(args.steps, 5))
    # the eigensolver: a fixed number of Lanczos steps (tol=0 never stops early, so every rank does the same work)
    m_steps = int(os.environ.get("CURV_C5_LANCZOS_STEPS", 30))
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    L0 = capi.lib().curv_launch_count()
    import warnings

    solve = lambda: lanczos_eigsh(G, k=10, ncv=m_steps, maxiter=m_steps, tol=0.0, return_info=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        solve()  # warm-up solve: allocator blocks of the Lanczos basis, LAPACK initialisation, graph capture
        ms_run, (evals, _, m_done) = timed(solve, 1)
    launches = capi.lib().curv_launch_count() - L0
    clocks = sampler.stop() if rank == 0 else None

    # e2e: ARPACK on the host drives the operator through the SciPy bridge (rank 0's vectors broadcast to the others)
    count = [0]
    G_host = GGNLinearOperator(model, loss, params, [(X_host, y_host)], check_deterministic=False, num_data=GB)
    G_host._engine = G._engine  # host-resident data (uploaded by every product), shared compiled program
    bridge = G_host.to_scipy(dtype=numpy.float32)
    stop = torch.zeros(1, device=dev)

    def mv(x):  # rank 0: called by ARPACK; the other ranks serve products until rank 0 says stop
        count[0] += 1
        if world > 1:
            dist.broadcast(stop, 0)
            xt = torch.from_numpy(numpy.ascontiguousarray(x, dtype=numpy.float32)).to(dev)
            dist.broadcast(xt, 0)
        return bridge.matvec(x)

    def e2e_run():  # host mini-batch uploaded by every product; Ritz pairs copied back to the host at the end
        ev, vecs, mm = lanczos_eigsh(G_host, k=10, ncv=m_steps, maxiter=m_steps, tol=0.0, return_info=True)
        return ev.cpu(), vecs.cpu(), mm

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e2e_run()  # warm-up (graph capture for the uploaded copies' addresses)
        ms_e2e, (_, vecs_host, m_e2e) = timed(e2e_run, 1)
    t_e2e = None
    if rank == 0:
        from scipy.sparse.linalg import LinearOperator as SL

        Aop = SL((P, P), matvec=mv, dtype=numpy.float32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            ev_sp = eigsh(Aop, k=10, ncv=m_steps, maxiter=1, tol=1e-2, which="LA",
                          v0=numpy.ones(P, dtype=numpy.float32), return_eigenvectors=False)
        except ArpackNoConvergence as e:
            ev_sp = e.eigenvalues
        t_e2e = time.perf_counter() - t0
        if world > 1:
            stop.fill_(1.0)
            dist.broadcast(stop, 0)
    else:
        while True:
            dist.broadcast(stop, 0)
            if stop.item() > 0:
                break
            xt = torch.empty(P, device=dev)
            dist.broadcast(xt, 0)
            G_host @ xt.to(torch.bfloat16)
    if world > 1:
        barrier()
    if rank != 0:
        dist.destroy_process_group()
        return
    lam = [float(e) for e in evals.float().cpu()]
    out = {
        "metric": "ggn_lanczos_matvecs_per_s", "value": m_done / (ms_run / 1e3), "unit": "matvec/s", "n_gpus": world,
        "steps": 1, "warmup": max(3, args.warmup), "ms_per_step": ms_run / m_done, "higher_is_better": True,
        "scaling": "weak", "warmup_note": "3 products + one full 30-step solve before the timed solve", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "params": P, "global_batch": GB, "lanczos_steps": m_done, "k": 10,
                   "parallelism": f"dp{world} (mini-batch sharded over the ranks, one all-reduce of [P] per product)",
                   "l2": "activations (GBs) far exceed the 126 MB L2; no flush needed"},
        "clocks": clocks, "gpu_launches": int(launches),
        "matvec_only": {"ms": ms_mv, "matvec_per_s": 1e3 / ms_mv,
                        "note": "G @ v alone, device-resident; the rest of a Lanczos step is the fp32 three-term "
                                "recurrence and the full re-orthogonalisation (torch GEMVs on [m, P])"},
        "largest_ritz_values": lam[-3:],
        "e2e": {"value": m_e2e / (ms_e2e / 1e3), "unit": "matvec/s", "matvecs": m_e2e,
                "what": "lanczos_eigsh(G, k=10) with the mini-batch in pinned host memory (uploaded by every product), "
                        "the 10 Ritz pairs copied to the host at the end; a step = one product",
                "h2d_bytes_per_step": int(X_host.numel() * 2 + y_host.numel() * 8),
                "d2h_bytes_per_step": int(vecs_host.numel() * 2 // m_e2e)},
        "scipy_eigsh_bridge": {"value": count[0] / t_e2e, "unit": "matvec/s", "matvecs": count[0],
                               "what": "the reference's own usage, scipy.sparse.linalg.eigsh(G.to_scipy(), k=10, ncv=%d, "
                                       "one restart cycle): host NumPy vector in / out per product, ARPACK's "
                                       "single-threaded vector algebra on the host included" % m_steps,
                               "h2d_bytes_per_step": int(P * 4 + X_host.numel() * 2 + y_host.numel() * 8),
                               "d2h_bytes_per_step": int(P * 4)},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.build_ref import import_reference

        ref = import_reference()
        try:  # the unmodified reference on this GPU, same model / data / dtype
            Gr = ref.GGNLinearOperator(model, loss, params, [(Xd, yd)], check_deterministic=False, num_data=GB)
            Gr @ v
            t_r, _ = timed(lambda: Gr @ v, 3)
            out["gpu_library_baseline"] = {"what": "unmodified reference GGNLinearOperator @ v on the same GPU "
                                                   "(torch CUDA, bf16)", "ms": t_r, "matvec_per_s": 1e3 / t_r}
            del Gr
        except Exception as e:
            out["gpu_library_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.set_num_threads(host_threads())
        sb = 8  # bounded CPU sample: the product is linear in the batch
        mc = torchvision.models.resnet50().eval()
        Xc, yc = torch.rand(sb, 3, 224, 224), torch.randint(0, 1000, (sb,))
        pc = dict(mc.named_parameters())
        Gc = ref.GGNLinearOperator(mc, loss, pc, [(Xc, yc)], check_deterministic=False, num_data=sb)
        vc = torch.rand(P)
        t0 = time.perf_counter()
        Gc @ vc
        t1 = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / (t1 * GB / sb), "unit": "matvec/s", "cores": torch.get_num_threads(),
                               "kind": "reference",
                               "sample": f"one product on {sb} of {GB} samples (fp32 on the CPU), extrapolated linearly"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global CONFIG, WORKLOAD
    CONFIG, WORKLOAD = args.config, WORKLOADS[args.config][0]
    DTYPE = WORKLOADS[CONFIG][1]
    global METRIC
    if CONFIG == "c2-hessian":
        METRIC = "hessian_matvec_throughput"
    if args.impl == "reference":  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core it may
        for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ.pop(v, None)
    import torch

    if CONFIG == "c5" and args.impl != "reference":
        return run_c5(torch, args)
    if CONFIG == "c4" and args.impl != "reference":
        return run_c4(torch, args)
    if CONFIG in ("c1", "c3") and args.impl != "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            (run_c3 if CONFIG == "c3" else run_c1)(torch, args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and not CONFIG.startswith("c2"):
        if rank == 0:  # the reference arm times the C2 workloads; the other configurations carry the reference's own
            # numbers (same GPU and host cores) in the gpu_library_baseline / cpu_baseline objects of their line
            print(json.dumps({"impl": "reference", "unavailable": f"--impl reference covers --config c2 / c2-bf16 / "
                              f"c2-hessian; run `bench.py --config {CONFIG}` for the reference numbers of this configuration"}))
        return
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's own CPU path at FULL size (one product takes tens of seconds on the host cores): best of
        # at most 2 products, no separate warm-up (stated in the line), so the run ends within a few minutes
        steps = max(1, min(args.steps, 2))
        # CURV_BENCH_REF_SAMPLE=<samples>: contract test of this arm on a build container (the line then says so)
        value, t_full, cb = cpu_reference_run(torch, steps, sample_batch=int(os.environ.get("CURV_BENCH_REF_SAMPLE", B)))
        cb["value"] = value
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "ms_per_step": t_full * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(args.gpus),
            "note": "unmodified reference (oracle/_ref) on the host cores, torch CPU kernels, full size; "
                    "ms_per_step is the best of `steps` products; no GPU is used by this arm",
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch.distributed as dist

    from curvlinops_b200 import GGNLinearOperator, HessianLinearOperator, _capi as capi
    from curvlinops_b200 import dist as cdist

    Operator = HessianLinearOperator if CONFIG == "c2-hessian" else GGNLinearOperator

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cdist.enable(True)
    model, X, y = build_problem(torch, B)
    model = model.to(dev)
    params = dict(model.named_parameters())
    P = sum(p.numel() for p in params.values())
    torch.manual_seed(1)
    pdt = next(iter(params.values())).dtype
    shm = None
    if world > 1:
        # one copy of V / of the result in host memory for all ranks of the box (POSIX shared memory registered as
        # pinned): each rank moves 1/world of the bytes over PCIe, NVLink does the rest
        shm = f"curvbench_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            V_host = cdist.shared_pinned_tensor(shm + "_V", (P, K), pdt)
            V_host.copy_(torch.rand(P, K).to(pdt))
            out_host = cdist.shared_pinned_tensor(shm + "_out", (P, K), pdt)
        dist.barrier()
        if rank != 0:
            V_host = cdist.shared_pinned_tensor(shm + "_V", (P, K), pdt)
            out_host = cdist.shared_pinned_tensor(shm + "_out", (P, K), pdt)
    else:
        V_host = torch.rand(P, K).to(pdt).pin_memory()
        out_host = torch.empty(P, K, dtype=pdt).pin_memory()
    X_host, y_host = X.pin_memory(), y.pin_memory()
    Xd, yd, Vd = X.to(dev), y.to(dev), V_host.to(dev)
    loss = torch.nn.CrossEntropyLoss()
    G = Operator(model, loss, params, [(Xd, yd)], check_deterministic=False, num_data=B)
    G_host = Operator(model, loss, params, [(X_host, y_host)], check_deterministic=False, num_data=B)
    G_host._engine = G._engine  # share compiled program + workspace

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    def step_device():
        return G @ Vd

    def step_e2e():
        # public host-operand API: V (pinned) in, result (pinned) out, X / y uploaded from pinned memory inside;
        # the upload of V and the download of the result are pipelined against the sweeps in parameter buckets
        G_host.matmat_pinned(V_host, out_host)
        torch.cuda.current_stream().synchronize()  # the result is on the host when the step ends
        return out_host

    for _ in range(max(3, args.warmup)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L0 = capi.lib().curv_launch_count()
    ms_step = timed(step_device, args.steps)
    launches = capi.lib().curv_launch_count() - L0
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel timing of the contraction kernels (CUDA events on the launching stream)
    import ctypes as C

    capi.lib().curv_profile_enable(1)
    barrier()
    for _ in range(min(2, args.steps)):
        step_device()
    barrier()
    ms, fl, cnt = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
    capi.lib().curv_profile_read(ms, fl, cnt)
    ms2, fl2, cnt2 = C.c_double(), C.c_double(), C.c_longlong()
    capi.lib().curv_profile_read_class(2, C.byref(ms2), C.byref(fl2), C.byref(cnt2))
    capi.lib().curv_profile_enable(0)
    nprof = min(2, args.steps)

    for _ in range(3):  # first call eager, second captures its CUDA graph, third replays
        step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 3)))

    # self-check outside the timed region: default tcgen05 path vs the exact-fp32 SIMT kernels, full size
    got = G @ Vd[:, :2]
    old_mode = capi.lib().curv_set_tensor_core_mode(0)
    ref = G @ Vd[:, :2]
    capi.lib().curv_set_tensor_core_mode(old_mode)
    self_check = float((got.float() - ref.float()).abs().max() / ref.float().abs().max())

    if world > 1:  # the shared result must equal the device-resident product
        barrier()
        e2e_check = float((out_host.to(dev).float() - (G @ Vd).float()).abs().max() / out_host.float().abs().max())
        barrier()
        if rank == 0:
            cdist.release_shared(shm + "_V")
            cdist.release_shared(shm + "_out")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
    dom = 0 if ms[0] >= ms[1] else 1
    ach = (fl[dom] / 1e12) / (ms[dom] / 1e3) if ms[dom] > 0 else 0.0
    roof = {
        "bound": "tensor", "kernel": ["gather_gemm (forward + dgrad)", "wgrad_gemm"][dom],
        "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
        "traffic": ncu_traffic_per_launch(),
        "peak_source": peak_src, "launches_timed": int(cnt[dom]),
        "share_of_step": (ms[dom] / nprof) / ms_step,
        "note": ("bf16 operator: one bf16 tcgen05 MMA per product, fp32 accumulation in tensor memory"
                 if DTYPE == "bf16" else
                 "fp32-grade result: every product is 3 fp16 tcgen05 MMAs on hi/lo split operands (half-split), "
                 "so the tensor-pipe ceiling for algorithmic FLOPs is peak/3; peak is the measured dense bf16/fp16 rate"),
        "frac_of_split_ceiling": None if DTYPE == "bf16" else ach / (peak_tf / 3.0),
        "traffic_note": "bytes of DRAM traffic per gather-GEMM launch (ncu --set full capture of forward launches of one "
                        "fp32 step, profiles/; equals the algorithmic bytes of those launches: fp16 hi/lo planes in, "
                        "fp32 result out, nothing re-read)",
        # whole step against both roofs (15.33 TFLOP and 50.4 GB algorithmic per C2 product)
        "step": {"algorithmic_tflop": 15.33, "tflops": 15.33 / (ms_step / 1e3),
                 "frac_of_tensor_peak": 15.33 / (ms_step / 1e3) / peak_tf,
                 "algorithmic_gb": 50.4 if DTYPE != "bf16" else 25.6,
                 "gbs": (50.4 if DTYPE != "bf16" else 25.6) / (ms_step / 1e3),
                 "frac_of_hbm_peak": (50.4 if DTYPE != "bf16" else 25.6) / (ms_step / 1e3) / peaks.get("hbm_gbs", 6550.0)},
        "other": {"kernel": ["gather_gemm", "wgrad_gemm"][1 - dom],
                  "tflops": (fl[1 - dom] / 1e12) / (ms[1 - dom] / 1e3) if ms[1 - dom] > 0 else 0.0,
                  "share_of_step": (ms[1 - dom] / nprof) / ms_step},
        "split_passes": {"kernel": "hs_absmax + hs_split (fp32 -> fp16 hi/lo planes)", "launches": int(cnt2.value),
                         "share_of_step": (ms2.value / nprof) / ms_step},
    }
    out = {
        "metric": METRIC, "value": P * K / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": config_dict(world),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": P * K / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(X_host.numel() * X_host.element_size() + y_host.numel() * 8
                                          + V_host.numel() * V_host.element_size()),
                "d2h_bytes_per_step": int(P * K * out_host.element_size()),
                **({"vs_resident_max_rel_diff": e2e_check,
                    "note": "V and the result live once in shared pinned host memory; every rank moves its 1/N row "
                            "block (H2D + NCCL all-gather, reduce-scatter + D2H); byte counts are per job"}
                   if world > 1 else {})},
        "roofline": roof,
        "self_check": {"tcgen05_vs_fp32_simt_max_rel_err": self_check, "columns": 2,
                       "note": "default tcgen05 path vs the exact-fp32 SIMT kernels on the same inputs, full size "
                               "(tolerance of the config: 1e-4 fp32, 1e-2 bf16)"},
    }
    if not args.no_cpu_baseline and world == 1:
        out["gpu_library_baseline"] = gpu_library_baseline(torch, dev)
        value, t_full, cb = cpu_reference_run(torch, 1, sample_batch=32)
        cb["value"], cb["unit"] = value, UNIT
        out["cpu_baseline"] = cb
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
