"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and
mean device time, share of the listed time.  Usage: summarize_launches.py launches.csv > summary.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
reader = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in reader:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"<.*", "", name) if not name.startswith(("kpreg", "void kpreg")) else name
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    agg[name][0] += 1
    agg[name][1] += us
    total += us
print(f"| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"| `{name[:110]}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% |")
print(f"\nlisted launches: {sum(v[0] for v in agg.values())}, listed device time: {total / 1000:.2f} ms")
ours = sum(us for name, (n, us) in agg.items() if "kpreg" in name or "cub::" in name)
print(f"kpreg_b200 kernels (incl. CUB sort/scan): {100 * ours / total:.1f}% of listed device time")
