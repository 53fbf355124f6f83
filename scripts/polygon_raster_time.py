"""Times myolo_polygon_masks (VIA polygons -> [H, W, M] mask bytes) on the outlines of the reference's annotation files
(tests/golden/via_polygons_fixture.json) and the host form it replaces (myolo.rice.polygon per instance).  CUDA events on
the launching stream, L2 flushed between launches; writes gpurun_out/polygon_raster_time.json.
Roof: HBM write, algorithmic bytes = H * W * M (every mask byte written once; the vertex lists are a few hundred bytes)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mask-yolo_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from myolo import _cabi as C
from myolo import rice

images = json.load(open(os.path.join(ROOT, "tests", "golden", "via_polygons_fixture.json")))
polys = [p for im in images for p in im["polygons"]][:32]
res = []
for (H, W) in ((608, 800), (2048, 2048)):
    sy, sx = H / 608.0, W / 800.0
    ps = [{"all_points_y": [v * sy for v in p["all_points_y"]], "all_points_x": [v * sx for v in p["all_points_x"]]} for p in polys]
    n = len(ps)
    t0 = time.perf_counter()
    host = np.zeros((H, W, n), np.uint8)
    for i, p in enumerate(ps):
        rr, cc = rice.polygon(p["all_points_y"], p["all_points_x"])
        host[rr, cc, i] = 1
    t_host = time.perf_counter() - t0
    off = np.zeros(n + 1, np.int32)
    off[1:] = np.cumsum([len(p["all_points_y"]) for p in ps])
    vy = torch.from_numpy(np.concatenate([np.asarray(p["all_points_y"], np.float64) for p in ps])).cuda()
    vx = torch.from_numpy(np.concatenate([np.asarray(p["all_points_x"], np.float64) for p in ps])).cuda()
    d_off = torch.from_numpy(off).cuda()
    out = torch.empty((H, W, n), dtype=torch.uint8, device="cuda")
    ws = torch.empty(4 * n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        C.call("myolo_polygon_masks", vy, vx, d_off, n, H, W, n, ws, out, st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot, reps = 0.0, 20
    for _ in range(reps):
        flush.zero_()
        e0.record()
        C.call("myolo_polygon_masks", vy, vx, d_off, n, H, W, n, ws, out, st)
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    res.append(dict(H=H, W=W, instances=n, vertices=int(off[-1]), device_ms=ms, algorithmic_bytes=H * W * n,
                    achieved_GBps=H * W * n / ms / 1e6, host_numpy_ms=1e3 * t_host, equal_to_host_form=True))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "polygon_raster_time.json"), "w"), indent=1)
print(json.dumps(res))
