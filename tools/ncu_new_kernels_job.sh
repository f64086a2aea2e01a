# ncu --set full of the kernels added late in round 2 (attention products, Lanczos re-orthogonalisation, split-K pre-sum)
set -x
mkdir -p gpurun_out
cat > /tmp/one_vit_step.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator
dev = torch.device("cuda")
torch.manual_seed(0)
m = torchvision.models.vit_b_16().eval().to(torch.bfloat16).to(dev)
X, y = torch.rand(32, 3, 224, 224, device=dev).to(torch.bfloat16), torch.randint(0, 1000, (32,), device=dev)
p = dict(m.named_parameters())
V = torch.rand(sum(q.numel() for q in p.values()), 4, device=dev).to(torch.bfloat16)
os.environ["CURV_CUDA_GRAPHS"] = "0"
G = GGNLinearOperator(m, torch.nn.CrossEntropyLoss(), p, [(X, y)], check_deterministic=False, num_data=32)
G @ V
torch.cuda.synchronize()
PY
CURV_CUDA_GRAPHS=0 timeout 900 ncu --set full --clock-control none -k regex:attn_sgemm -c 10 --csv --page raw --log-file gpurun_out/ncu_attn_raw.csv python /tmp/one_vit_step.py > gpurun_out/ncu_attn.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/ncu_attn_raw.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_tensor.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [(w, names.index(w)) for w in want if w in names]
with open("gpurun_out/r2_attn_sgemm_ncu_full.txt", "w") as f:
    for r in rows[hdr + 2:]:
        if len(r) < len(names): continue
        f.write("---\n")
        for w, i in idx:
            f.write(f"{w:70s} {r[i]} {units[i]}\n")
print(open("gpurun_out/r2_attn_sgemm_ncu_full.txt").read()[:3000])
PY
