"""ResNet-50 GGN accuracy diagnostic (GPU box): the engine and torch autograd in the SAME dtype against the float64
oracle on the same inputs; logit scale; per-parameter error of the engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from curvlinops_b200 import GGNLinearOperator
from oracle import curvature_oracle as orc

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
B, K = int(os.environ.get("CURV_B", 8)), 2
name = os.environ.get("CURV_MODEL", "resnet50")
loss = torch.nn.CrossEntropyLoss()
for dt in (torch.float32, torch.bfloat16):
    torch.manual_seed(0)
    model = getattr(torchvision.models, name)().eval().to(dev).to(dt)
    X = torch.rand(B, 3, 224, 224, device=dev).to(dt)
    y = torch.randint(0, 1000, (B,), device=dev)
    params = dict(model.named_parameters())
    sizes = [p.numel() for p in params.values()]
    P = sum(sizes)
    V = torch.rand(P, K, device=dev).to(dt)
    m64 = getattr(torchvision.models, name)().eval().to(dev).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    p64 = dict(m64.named_parameters())
    split = lambda V, ps, d: [v.reshape(*p.shape, K).to(d) for v, p in zip(V.split(sizes), ps.values())]
    ref = torch.cat([r.reshape(-1, K) for r in orc.ggn_matmat(m64, loss, p64, [(X.double(), y)], split(V, p64, torch.float64))])
    with torch.no_grad():
        lg = m64(X.double())
        pmax = lg.softmax(-1).max(-1).values
    print(f"== {name} {dt}: logits absmax {lg.abs().max():.3g}, std {lg.std():.3g}; softmax max prob per sample "
          f"min {pmax.min():.4f} mean {pmax.mean():.4f}")
    tor = torch.cat([r.reshape(-1, K) for r in orc.ggn_matmat(model, loss, params, [(X, y)], split(V, params, dt))]).double()
    got = (GGNLinearOperator(model, loss, params, [(X, y)], check_deterministic=False) @ V).double()
    sc = ref.abs().max()
    print(f"engine vs f64: {((got - ref).abs().max() / sc):.3e}   torch autograd ({dt}) vs f64: {((tor - ref).abs().max() / sc):.3e}"
          f"   engine vs torch: {((got - tor).abs().max() / sc):.3e}")
    rows, o = [], 0
    for n, sz in zip(params, sizes):
        r = ref[o:o + sz]
        rows.append(((got[o:o + sz] - r).abs().max().item() / sc.item(), (tor[o:o + sz] - r).abs().max().item() / sc.item(),
                     r.abs().max().item() / sc.item(), n))
        o += sz
    rows.sort(reverse=True)
    for e, t, m, n in rows[:6]:
        print(f"   {n:34s} engine {e:.2e}  torch {t:.2e}  (block max/global max {m:.2e})")
