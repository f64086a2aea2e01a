"""HBM roofline of the streaming (elementwise) kernels of one C2 step from a per-kernel step profile.

    python tools/streaming_roofline.py [profiles/r1_s2_step_profile_latest.txt] [--batch 128] [--columns 8]

Reads the per-kernel times written by tools/gpu_step_profile.py, attaches the ALGORITHMIC bytes of each kernel class
on ResNet-18 (every fp32 slot / fp16 plane pair it has to read or write once) and prints time at the measured copy
bandwidth (MEASURED_PEAKS.json, else 6553 GB/s) vs measured time.  Runs anywhere (no GPU): it only parses text.
"""
import argparse
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def resnet18_elements(B: int) -> dict:
    """Activation elements (NHWC, channels already multiples of 8) of the maps the streaming kernels touch."""
    e = {"stem": B * 112 * 112 * 64, "l1": B * 56 * 56 * 64, "l2": B * 28 * 28 * 128, "l3": B * 14 * 14 * 256,
         "l4": B * 7 * 7 * 512}
    e["blocks"] = 2 * (e["l1"] + e["l2"] + e["l3"] + e["l4"])  # one map per BasicBlock (bn1, bn2, join)
    e["downsample"] = e["l2"] + e["l3"] + e["l4"]
    return e


def algorithmic_bytes(B: int, K: int) -> dict:
    e, S = resnet18_elements(B), 1 + K
    return {
        # forward: S slots (primal + K tangents); backward: K cotangent slots
        "affine_fwd_kernel<true>": e["blocks"] * (S * 4 + S * 4 + 4),        # fp32 in; plane pairs + fp32 primal out
        "affine_fwd_kernel<false>": (e["stem"] + e["blocks"] + e["downsample"]) * (2 * S * 4),
        "affine_bwd_kernel": (e["stem"] + 2 * e["blocks"] + e["downsample"]) * (K * 4 + 4 + 4 + K * 4),
        "add_relu_fwd_kernel": e["blocks"] * (3 * S * 4),
        "add_relu_bwd_kernel": e["blocks"] * ((K + 1) * 4 + 2 * K * 4),
        "maxpool_fwd_kernel": e["stem"] * S * 4 + e["l1"] * S * 4 + e["l1"],
        "maxpool_bwd_kernel": e["stem"] * K * 4 + e["l1"] * K * 4 + e["l1"],
    }


def parse_profile(path: str) -> dict:
    """kernel name (without the curv:: prefix and the argument list) -> (ms, launches)."""
    out = {}
    pat = re.compile(r"^\s*[\d.]+%\s+([\d.]+) ms\s+n=\s*(\d+)\s+(?:void\s+)?(?:curv::)?([\w:<>, ]+?)\(")
    for line in open(path):
        m = pat.match(line)
        if m:
            out[m.group(3).strip()] = (float(m.group(1)), int(m.group(2)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("profile", nargs="?", default=os.path.join(ROOT, "profiles", "r1_s2_step_profile_latest.txt"))
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--columns", type=int, default=8)
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6553.0
    times = parse_profile(args.profile)
    tot_meas = tot_peak = 0.0
    print(f"# {args.profile}: B={args.batch}, K={args.columns}, peak {peak:.0f} GB/s")
    print(f"{'kernel':28s} {'n':>3s} {'GB':>7s} {'at peak':>9s} {'measured':>9s} {'frac':>6s}")
    for name, nbytes in algorithmic_bytes(args.batch, args.columns).items():
        if name not in times:
            print(f"{name:28s}   - not in the profile")
            continue
        ms, n = times[name]
        at_peak = nbytes / (peak * 1e9) * 1e3
        tot_meas += ms
        tot_peak += at_peak
        print(f"{name:28s} {n:3d} {nbytes / 1e9:7.2f} {at_peak:7.2f}ms {ms:7.2f}ms {100 * at_peak / ms:5.0f}%")
    if tot_meas:
        print(f"{'total':28s}     {'':7s} {tot_peak:7.2f}ms {tot_meas:7.2f}ms {100 * tot_peak / tot_meas:5.0f}%")


if __name__ == "__main__":
    main()
