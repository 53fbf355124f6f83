"""Host-side cost of one training step: time to ENQUEUE a step (no synchronisation) vs the device time."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import torch
from myolo.model import MaskYOLO
from myolo.shapes import ShapesConfig, make_batches

class Cfg(ShapesConfig):
    BATCH_SIZE = 32
cfg = Cfg()
model = MaskYOLO("training", cfg)
hb = make_batches(cfg, 1, seed=3)[0]
dev = [t.clone() for t in model._stage(hb)]
eng = model.engine
eng.inputs_ready = None
for _ in range(3):
    eng.train_step(dev, 1e-3)
torch.cuda.synchronize()
# enqueue cost: two steps at a time (well below the launch-queue depth, so the host never waits for the device)
best = 1e9
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        eng.train_step(dev, 1e-3)
    best = min(best, (time.perf_counter() - t0) / 2)
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    eng.train_step(dev, 1e-3)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * best:.2f} ms/step, device-bound step {1e3 * (t2 - t0) / n:.2f} ms  (replay={'on' if eng._replay_on else 'off'})")
