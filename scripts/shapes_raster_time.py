"""Times the device Shapes generator (myolo_shapes_raster + myolo_encode_yolo_targets) at the benchmark batch and the host
chain it replaces (ShapesDataset + load_image_gt + BatchGenerator) on the same images.  CUDA events on the launching stream;
writes gpurun_out/shapes_raster_time.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mask-yolo_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from myolo import myolo_utils as mutils
from myolo.shapes import DeviceShapes, ShapesConfig, ShapesDataset, spec_table


class Cfg(ShapesConfig):
    BATCH_SIZE = 32


cfg = Cfg()
B, S, M = 32, 224, cfg.MAX_GT_INSTANCES
ds = ShapesDataset(1234)
ds.load_shapes(8 * B, S, S)
ds.prepare()
t0 = time.perf_counter()
tab = spec_table(ds)
t_spec = (time.perf_counter() - t0) / 8
t0 = time.perf_counter()
info = [list(mutils.load_image_gt(ds, cfg, i, use_mini_mask=False)) for i in range(2 * B)]
gen = mutils.BatchGenerator(info, cfg, mode="training", shuffle=False, norm=True)
host = [gen[i][0] for i in range(2)]
t_host = (time.perf_counter() - t0) / 2
feeder = DeviceShapes(cfg)
for k in range(3):
    feeder.batch(tab[k * B:(k + 1) * B])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 40
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for k in range(n):
    flush.zero_()                                   # write 256 MB: the previous batch leaves the 126 MB L2
    e0.record()
    feeder.batch(tab[(k % 8) * B:(k % 8 + 1) * B])
    e1.record()
    e1.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / n
out_bytes = B * S * S * (12 + M) + B * cfg.TRUE_BOX_BUFFER * 36
res = dict(workload="Shapes 224x224 batch 32, M=%d" % M, device_ms_per_batch=ms, algorithmic_bytes=out_bytes,
           achieved_GBps=out_bytes / ms / 1e6, host_chain_ms_per_batch=1e3 * t_host, spec_table_ms_per_batch=1e3 * t_spec,
           h2d_bytes_per_batch=int(tab[:B].nbytes), host_fed_h2d_bytes_per_batch=int(sum(np.asarray(x).nbytes for x in host[0])),
           note="device time includes the pinned spec upload, both raster kernels and the target-encoding launch")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "shapes_raster_time.json"), "w"), indent=1)
print(json.dumps(res))
