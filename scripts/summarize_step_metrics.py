"""Per-kernel roofline table of ONE training step from an ncu metrics log (scripts/profile_step_metrics.sh): for every kernel
name of the last full step (between the last two adam launches): launches, total time, share, DRAM bytes and GB/s
(read + written, per launch average and the best launch), DRAM throughput %, tensor-pipe % (time-weighted).
usage: summarize_step_metrics.py step_metrics.csv"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
launches = OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    d = launches.setdefault(key, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", "") or 0)
    unit = r.get("Metric Unit", "")
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
    if name.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[name] = v
rows = list(launches.values())
adam = [i for i, d in enumerate(rows) if "adam_kernel" in d["name"]]
assert len(adam) >= 2, "need at least two steps in the capture"
step = rows[adam[-2] + 1:adam[-1] + 1]
agg = OrderedDict()
for d in step:
    k = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
    a = agg.setdefault(k, {"n": 0, "ns": 0.0, "bytes": 0.0, "tensor_w": 0.0, "dram_w": 0.0, "best": 0.0})
    ns = d.get("gpu__time_duration.sum", 0.0)
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["n"] += 1; a["ns"] += ns; a["bytes"] += by
    a["tensor_w"] += ns * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["dram_w"] += ns * d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    if ns > 0:
        a["best"] = max(a["best"], by / ns)
total = sum(a["ns"] for a in agg.values())
print(f"one training step: {len(step)} launches, serialised sum {total / 1e6:.3f} ms (ncu, cold cache per launch); "
      f"measured HBM peak {PEAK:.0f} GB/s")
print(f"{'share':>7} {'time ms':>9} {'n':>4} {'DRAM MB/launch':>15} {'GB/s avg':>9} {'of peak':>8} {'GB/s best':>10} {'dram %':>7} {'tensor %':>9}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    gbs = a["bytes"] / a["ns"] if a["ns"] else 0.0
    print(f"{100 * a['ns'] / total:6.2f}% {a['ns'] / 1e6:9.3f} {a['n']:4d} {a['bytes'] / a['n'] / 1e6:15.1f} {gbs:9.0f} {gbs / PEAK:8.2f} "
          f"{a['best']:10.0f} {a['dram_w'] / max(a['ns'], 1):7.1f} {a['tensor_w'] / max(a['ns'], 1):9.1f}  {k}")
