"""Diagnostic: per-layer deviation of the tf32 engine from the fp32 engine on the same batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import torch
from myolo.engine import Engine, init_params
from tests import helpers as Hh

S, B = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "tf32"
c = Hh.engine_cfg(S=S)
P = init_params(c["NB"], c["NC"], 100, "trained_like")
img = torch.rand(B, S, S, 3, generator=torch.Generator().manual_seed(1)).cuda()
snap = {}
for prec in ("fp32", mode):
    eng = Engine(c, B, "inference", prec, params=P)
    eng.forward(img, training=True)
    masks = eng.mask_head(eng.A["proposals"], training=True)
    torch.cuda.synchronize()
    d = {k: (v[0] + v[1] if (k.startswith("ad") and v.dim() == 5) else v.clone()) for k, v in eng.A.items()}
    d["c4"] = eng.c4.dense(); d["feat"] = eng.feat.dense(); d["x0"] = eng.x0.dense()
    for i in (1, 2, 3, 4):
        d[f"ma{i}"] = eng.ma[i].dense()
    snap[prec] = d
    del eng
for k in snap["fp32"]:
    if k not in snap[mode] or snap["fp32"][k].shape != snap[mode][k].shape:
        continue
    a, b = snap["fp32"][k], snap[mode][k]
    rel = ((a - b).norm() / a.norm().clamp_min(1e-30)).item()
    print(f"{k:12s} shape {tuple(a.shape)!s:24s} relL2 {rel:.3e}  maxabs {(a - b).abs().max().item():.3e}  ref rms {a.pow(2).mean().sqrt().item():.3e}")
