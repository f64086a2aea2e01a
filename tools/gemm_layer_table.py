"""Per-layer algorithmic TFLOP/s of the contraction launches of one C2 step (ResNet-18) from a step profile.

    python tools/gemm_layer_table.py [profiles/r1_s2_step_profile_latest.txt] [--batch 128] [--columns 8]

The profile (tools/gpu_step_profile.py) lists the contraction launches in execution order: the forward sweep visits
the 20 convolutions in network order (the fc layer runs on the SIMT kernels), the backward sweep visits them in
reverse order with one wgrad and - except for the stem - one dgrad launch each.  FLOPs per launch:
forward  2*M*N*Kd * (1 + K*(has_tangent_input + 1)),  wgrad 2*M*N*Kd*K,  dgrad 2*M*N*Kd*K  (algorithmic, un-padded;
three fp16 MMAs per product are NOT counted).  Runs anywhere (text in, text out).
"""
import argparse
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def resnet18_convs(B: int):
    """(name, M, N, Kd, input carries tangents) in forward order."""
    L = [("conv1 7x7/2", B * 112 * 112, 64, 7 * 7 * 3, False)]
    cin, hw = 64, 56
    for li, cout in enumerate([64, 128, 256, 512], start=1):
        for blk in range(2):
            stride = 2 if (li > 1 and blk == 0) else 1
            ho = hw // stride
            L.append((f"layer{li}.{blk}.conv1 3x3/{stride}", B * ho * ho, cout, 9 * cin, True))
            L.append((f"layer{li}.{blk}.conv2 3x3", B * ho * ho, cout, 9 * cout, True))
            if stride == 2:
                L.append((f"layer{li}.{blk}.downsample 1x1/2", B * ho * ho, cout, cin, True))
            cin, hw = cout, ho
    return L


def parse_launches(path: str):
    rows, on = [], False
    for line in open(path):
        if line.startswith("# contraction launches"):
            on = True
            continue
        if on:
            m = re.match(r"^(.*?)\s+([\d.]+)\s*$", line.rstrip())
            if m:
                rows.append((m.group(1).strip(), float(m.group(2))))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("profile", nargs="?", default=os.path.join(ROOT, "profiles", "r1_s2_step_profile_latest.txt"))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--columns", type=int, default=8)
    args = ap.parse_args()
    K = args.columns
    convs = resnet18_convs(args.batch)
    rows = [r for r in parse_launches(args.profile) if "simt" not in r[0]]
    fwd, bwd = rows[:len(convs)], rows[len(convs):]
    print(f"# {args.profile}: B={args.batch}, K={K}; algorithmic TFLOP/s per launch")
    print(f"{'layer':34s} {'kernel':10s} {'GFLOP':>8s} {'ms':>7s} {'TFLOP/s':>8s}")
    tot_f = tot_t = 0.0

    def line(name, kern, fl, ms):
        nonlocal tot_f, tot_t
        tot_f += fl
        tot_t += ms
        print(f"{name:34s} {kern[:10]:10s} {fl / 1e9:8.1f} {ms:7.3f} {fl / 1e12 / (ms / 1e3):8.1f}")

    for (name, M, N, Kd, tan), (kern, ms) in zip(convs, fwd):
        line("fwd  " + name, kern, 2.0 * M * N * Kd * (1 + K * ((1 if tan else 0) + 1)), ms)
    it = iter(bwd)
    for name, M, N, Kd, tan in reversed(convs):
        kern, ms = next(it)
        assert "wgrad" in kern, (name, kern)
        line("wgrad " + name, kern, 2.0 * M * N * Kd * K, ms)
        if tan:
            kern, ms = next(it)
            assert "gather" in kern, (name, kern)
            line("dgrad " + name, kern, 2.0 * M * N * Kd * K, ms)
    print(f"{'total':34s} {'':10s} {tot_f / 1e9:8.1f} {tot_t:7.3f} {tot_f / 1e12 / (tot_t / 1e3):8.1f}")


if __name__ == "__main__":
    main()
