"""Micro-benchmark of the HBM-bound kernel families at the benchmark shapes (N(0,1) activations in the exact
layer shapes of 8a at B=32, S=224; >= 3 warm-ups, CUDA events on the launching stream, median of 20, 256 MB written
between iterations so that the 126 MB L2 is cold): depthwise 3x3 forward / backward-data / backward-filter for the 14
blocks, BatchNorm statistics / apply / backward on the same tensors, ROIAlign forward / backward.  Prints one line per
kernel and layer with algorithmic bytes, time, GB/s and the fraction of the measured HBM peak (MEASURED_PEAKS.json), plus
the per-family totals; writes gpurun_out/hbm_kernels.json.  These are the kernels north_star wants at >= 60 % of 8 TB/s.

    gpurun --timeout 300 -- 'python scripts/bench_hbm_kernels.py'
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mask-yolo_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from myolo import _cabi as C

B, S = 32, 224
# (block, Cin of the depthwise = channels, stride, input H)   myolo/model.py:68-77, 256-268
DW = [(1, 32, 1, 112), (2, 64, 2, 112), (3, 64, 1, 56), (4, 128, 2, 56), (5, 256, 1, 28), (6, 256, 1, 28), (7, 512, 2, 28),
      (8, 512, 1, 14), (9, 512, 1, 14), (10, 512, 1, 14), (11, 512, 1, 14), (12, 512, 1, 14), (13, 512, 2, 14), (14, 1024, 1, 7)]
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = lambda: torch.cuda.current_stream().cuda_stream        # noqa: E731


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


rows = []


def report(family, layer, nbytes, ms):
    gbs = nbytes / ms / 1e6
    rows.append(dict(family=family, layer=layer, bytes=int(nbytes), ms=ms, GBps=gbs, frac_of_measured_peak=gbs / PEAK, frac_of_8TBps=gbs / 8000.0))
    print("%-18s %-10s %8.1f MB %8.4f ms %8.0f GB/s  %.2f of measured peak" % (family, layer, nbytes / 1e6, ms, gbs, gbs / PEAK))


for k, Cc, s, H in DW:
    Ho = (H + 2 - 3) // s + 1
    x = torch.randn(B, H, H, Cc, device="cuda")
    w = torch.randn(3, 3, Cc, device="cuda")
    y = torch.empty(B, Ho, Ho, Cc, device="cuda")
    dy = torch.randn(B, Ho, Ho, Cc, device="cuda")
    dx = torch.empty_like(x)
    dw = torch.empty(3, 3, Cc, device="cuda")
    xv = C.view(x, B, H, H, Cc)
    io = 4.0 * (x.numel() + y.numel()) + 36 * Cc
    report("dw_fwd", "dw%d" % k, io, timed(lambda: C.call("myolo_dwconv3x3_fwd", xv, w, y, s, st())))
    report("dw_bwd_data", "dw%d" % k, io, timed(lambda: C.call("myolo_dwconv3x3_bwd_data", dy, w, dx, B, H, H, Cc, s, st())))
    report("dw_bwd_filter", "dw%d" % k, io, timed(lambda: C.call("myolo_dwconv3x3_bwd_filter", xv, dy, dw, s, st())))
    # the BatchNorm that follows the depthwise conv, on its output tensor
    mean, var = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    g, b = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    yv, ya = C.view(y, B, Ho, Ho, Cc), torch.empty_like(y)
    y.normal_()
    report("bn_stats", "dw%d_bn" % k, 4.0 * y.numel(), timed(lambda: C.call("myolo_bn_stats", yv, mean, var, ws, st())))
    report("bn_apply", "dw%d_bn" % k, 8.0 * y.numel(),
           timed(lambda: C.call("myolo_bn_apply", yv, C.view(ya, B, Ho, Ho, Cc), mean, var, g, b, 1e-3, C.ACT_RELU6, st())))
    dg, db, dya = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda"), torch.empty_like(y)
    report("bn_bwd", "dw%d_bn" % k, 16.0 * y.numel(),        # two passes: reductions (x, dy) then dx (x, dy -> dx)
           timed(lambda: C.call("myolo_bn_bwd", yv, C.view(dy, B, Ho, Ho, Cc), C.view(dya, B, Ho, Ho, Cc), mean, var, g, b, 1e-3,
                                C.ACT_RELU6, 1, dg, db, ws, st())))
    del x, y, dy, dx, ya, dya

# ROIAlign at the benchmark size: 147 ROIs per image on the 28x28x256 feature map, 14x14 samples each
R, P, F, Cf = 147, 14, 28, 256
feat = torch.randn(B, F, F, Cf, device="cuda")
c = torch.rand(B, R, 2, device="cuda") * 0.8 + 0.1
wh = torch.rand(B, R, 2, device="cuda") * 0.4 + 0.05
rois = torch.cat([c - wh / 2, c + wh / 2], -1).contiguous()
out = torch.empty(B * R, P, P, Cf, device="cuda")
fv, ov = C.view(feat, B, F, F, Cf), C.view(out, B * R, P, P, Cf)
report("roialign_fwd", "fp32 out", 4.0 * (out.numel() + feat.numel()) + 16 * B * R,
       timed(lambda: C.call("myolo_roialign_fwd", fv, rois, B * R, R, P, ov, 0, st())))
dfeat = torch.zeros_like(feat)
dv = C.view(dfeat, B, F, F, Cf)
report("roialign_bwd", "fp32 dout", 4.0 * (out.numel() + 2 * feat.numel()),
       timed(lambda: C.call("myolo_roialign_bwd", ov, rois, B * R, R, P, dv, st())))

fam = {}
for r in rows:
    f = fam.setdefault(r["family"], [0.0, 0.0])
    f[0] += r["bytes"]
    f[1] += r["ms"]
print()
for k, (nb, ms) in fam.items():
    print("%-18s total %8.1f MB %8.4f ms  %8.0f GB/s  %.2f of measured peak, %.2f of 8 TB/s" % (k, nb / 1e6, ms, nb / ms / 1e6, nb / ms / 1e6 / PEAK, nb / ms / 1e6 / 8000))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(B=B, S=S, hbm_peak_GBps=PEAK, rows=rows, families={k: dict(bytes=v[0], ms=v[1], GBps=v[0] / v[1] / 1e6) for k, v in fam.items()}),
          open(os.path.join(ROOT, "gpurun_out", "hbm_kernels.json"), "w"), indent=1)
