"""Time the C2 step under several engine modes in one process (GPU box).  usage: python tools/gpu_variant_bench.py mode[,mode...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator, _capi as capi

modes = [int(m, 0) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1"])]
B, K = int(os.environ.get("CURV_B", 128)), 8  # CURV_B = 16 emulates one rank of an 8-GPU run
torch.manual_seed(0)
dev = torch.device("cuda")
model = torchvision.models.resnet18().eval().to(dev)
X, y = torch.rand(B, 3, 224, 224, device=dev), torch.randint(0, 1000, (B,), device=dev)
params = dict(model.named_parameters())
P = sum(p.numel() for p in params.values())
V = torch.rand(P, K, device=dev)
G = GGNLinearOperator(model, torch.nn.CrossEntropyLoss(), params, [(X, y)], check_deterministic=False, num_data=B)
L = capi.lib()
ref = None
for mode in modes:
    L.curv_set_tensor_core_mode(mode)
    for _ in range(5):  # eager, then CUDA-graph capture for both alternating (V, out) address sets
        out = G @ V
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        out = G @ V
    e1.record(); torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / 4
    L.curv_profile_enable(1)
    out = G @ V
    torch.cuda.synchronize()
    ms, fl, cnt = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_longlong * 2)()
    L.curv_profile_read(ms, fl, cnt); L.curv_profile_enable(0)
    if ref is None:
        ref = out.clone()
    dev_ = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"mode={mode:#x}: {ms_step:7.2f} ms/step  gather {ms[0]:6.1f} ms ({fl[0]/ms[0]/1e9:6.1f} TF/s)  "
          f"wgrad {ms[1]:6.1f} ms ({fl[1]/ms[1]/1e9:6.1f} TF/s)  other {ms_step-ms[0]-ms[1]:5.1f} ms  "
          f"dev-vs-first {dev_:.2e}", flush=True)
L.curv_set_tensor_core_mode(1)
