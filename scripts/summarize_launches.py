"""Per-kernel breakdown of ONE training step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): the launches between the last two
adam_kernel launches (one full forward + backward + update).  usage: summarize_launches.py launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    rows.append((r["Kernel Name"], ns))
adam = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k]
assert len(adam) >= 2, "need at least two steps in the capture"
step = rows[adam[-2] + 1:adam[-1] + 1]
agg = OrderedDict()
for k, ns in step:
    k = re.sub(r"\(.*", "", k)
    a = agg.setdefault(k, [0.0, 0])
    a[0] += ns
    a[1] += 1
total = sum(a[0] for a in agg.values())
print(f"one training step: {len(step)} launches, serialised sum {total / 1e6:.3f} ms")
print()
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * ns / total:6.2f}%  {ns / 1e6:9.3f} ms  x{n:4d}  {k}")
