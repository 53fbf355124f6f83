"""Convergence sanity: trains Mask-YOLO on synthetic Shapes for a few hundred steps and prints the losses."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import numpy as np, torch
from myolo.model import MaskYOLO
from myolo.shapes import ShapesConfig, make_batches

S = int(sys.argv[1]) if len(sys.argv) > 1 else 224
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
prec = sys.argv[3] if len(sys.argv) > 3 else "h16"

class Cfg(ShapesConfig):
    BATCH_SIZE = 16
    IMAGE_SHAPE = [S, S, 3]
    IMAGE_MIN_DIM = IMAGE_MAX_DIM = S
    GRID_H = GRID_W = S // 32

cfg = Cfg()
np.random.seed(0)
model = MaskYOLO("training", cfg, precision=prec)
batches = make_batches(cfg, 8, seed=11)
t0 = time.time()
for i in range(steps):
    v = model.keras_model.train_on_batch(batches[i % len(batches)])
    if i % 20 == 0 or i == steps - 1:
        npos = int(model.engine.n_pos.sum().item())
        gs = f"  loss scale 2^{int(np.log2(model.engine.gs[0].item()))}" if prec == "h16" else ""
        print(f"step {i:4d}  loss {v[0]:9.4f}  yolo {v[1]:9.4f}  mask {v[2]:7.4f}  positives {npos}{gs}", flush=True)
print(f"{steps} steps in {time.time() - t0:.1f}s")
