"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/r2_gpu_job.sh).
usage: python tools/ncu_launch_summary.py launches.csv "title" > summary.txt"""
import csv, io, re, sys

lines = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = {}
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print("# `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over `python tools/gpu_one_step.py`")
print("# (ONE product, launched eagerly).  Per-launch times are serialised / cold-cache: compare SHARES with bench.py's live")
print("# CUDA-event shares (roofline.share_of_step) and the CUPTI step profile, not absolute times.")
print(f"total kernel time {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ms / tot:6.2f}% {ms:9.3f} ms  n={n:4d}  {name}")
