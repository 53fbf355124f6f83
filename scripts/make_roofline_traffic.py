"""profiles/roofline_traffic.json from an `ncu --set full` capture of the dominant kernel (scripts/capture_ncu.sh fwd_h, run
alone at the benchmark size: 4704 ROIs): DRAM bytes read + written per launch and per ROI.  bench.py multiplies the per-ROI
figure by the ROI count of the configuration it runs.
usage: python scripts/make_roofline_traffic.py gpurun_out/ncu_fwd_h.raw.csv 4704 [gpurun_out/ncu_fwd.raw.csv]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        u, v = d[k]
        tot += float(v.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return tot, d["Kernel Name"][1], d["gpu__time_duration.sum"]


n_roi = int(sys.argv[2])
out = {"source": f"scripts/capture_ncu.sh + scripts/make_roofline_traffic.py: dram__bytes_read.sum + dram__bytes_write.sum of one "
                 f"`ncu --set full --clock-control none` capture per kernel, run alone on {n_roi} ROIs (benchmark size)",
       "n_roi": n_roi}
b, name, dur = dram_bytes(sys.argv[1])
out.update(mask_conv_fwd_h16_dram_bytes_per_launch=b, mask_conv_fwd_h16_dram_bytes_per_roi=b / n_roi,
           mask_conv_fwd_h16_kernel=name[:100], mask_conv_fwd_h16_duration=f"{dur[1]} {dur[0]}",
           mask_conv_fwd_h16_algorithmic_bytes_per_roi=2 * 225 * 256 * 2)
if len(sys.argv) > 3:
    b, name, dur = dram_bytes(sys.argv[3])
    out.update(mask_conv_fwd_dram_bytes_per_launch=b, mask_conv_fwd_dram_bytes_per_roi=b / n_roi, mask_conv_fwd_kernel=name[:100],
               mask_conv_fwd_algorithmic_bytes_per_roi=2 * 225 * 256 * 4)
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out))
