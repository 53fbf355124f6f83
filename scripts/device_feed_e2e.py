"""End-to-end images/s of the training step when the Shapes batches are generated on the device (myolo.shapes.DeviceShapes)
instead of being uploaded: per step the host draws nothing but copies a 4.6 KB spec table; the step reads back its two
losses.  Same model / batch / precision as bench.py; writes gpurun_out/device_feed_e2e.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mask-yolo_b200")):
    sys.path.insert(0, p)
import torch
from bench import bench_config
from myolo.engine import init_params
from myolo.model import MaskYOLO
from myolo.shapes import DeviceShapes, ShapesDataset, spec_table

B, S, K, W = 32, 224, 20, 5
cfg = bench_config(B, S)
model = MaskYOLO("training", cfg, precision="h16")
c = model.engine.cfg
model.engine.load_params(init_params(c["NB"], c["NC"], 0, "trained_like"))
ds = ShapesDataset(1234)
ds.load_shapes(8 * B, S, S)
ds.prepare()
tab = spec_table(ds)
feeder = DeviceShapes(cfg)
for i in range(W):
    vals = model.keras_model.train_on_batch(feeder.batch(tab[(i % 8) * B:(i % 8 + 1) * B]))
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(K):
    vals = model.keras_model.train_on_batch(feeder.batch(tab[(i % 8) * B:(i % 8 + 1) * B]))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res = dict(workload="Shapes 224x224 batch 32, h16, inputs generated on the device from spec tables", steps=K,
           images_per_sec=B * K / dt, ms_per_step=1e3 * dt / K, h2d_bytes_per_step=int(tab[:B].nbytes),
           d2h_bytes_per_step=int(model.last_d2h_bytes), loss=[float(v) for v in vals])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "device_feed_e2e.json"), "w"), indent=1)
print(json.dumps(res))
