"""How stable is the 12-epoch training run of tests/test_api_gpu.py::test_train_loop_checkpoint_and_reload from run to run?
(fp32 atomics order + chaotic early training): prints first / last-three losses of N runs."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_api_gpu import _cfg
from myolo.model import MaskYOLO
from myolo.shapes import ShapesDataset
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for run in range(n):
    cfg = _cfg()
    tr, va = ShapesDataset(seed=1), ShapesDataset(seed=2)
    tr.load_shapes(12, 128, 128); tr.prepare()
    va.load_shapes(4, 128, 128); va.prepare()
    np.random.seed(0)
    with tempfile.TemporaryDirectory() as d:
        model = MaskYOLO(mode="training", config=cfg, model_dir=d)
        hist = model.train(tr, va, learning_rate=cfg.LEARNING_RATE, epochs=12, layers="all", verbose=0)
    L = hist["loss"]
    print("run %d: first %.3f  last3 %s  ratio %.3f  val %.3f" % (run, L[0], ["%.3f" % v for v in L[-3:]], min(L[-3:]) / L[0], hist["val_loss"][-1]), flush=True)
