"""Runs the dominant kernel alone at benchmark size (mask-head 3x3 conv, 4704 ROIs of 14x14x256)
for ncu captures and quick timing:  python scripts/profile_conv.py [n_roi] [iters] [which]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import torch
from myolo import _cabi as C
from myolo.pf import PF, conv3x3_shifts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4704
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
C.device_check(0); C.set_precision(C.PREC_TF32)
st = torch.cuda.current_stream().cuda_stream
x, y = PF(n, 14, 14, 256), PF(n, 14, 14, 256)
x.valid().normal_(); y.valid().normal_()
w = torch.randn(9, 256, 256, device="cuda") / 48
bias = torch.zeros(256, device="cuda")
dw = torch.zeros(9, 256, 256, device="cuda")
flops = 2.0 * n * 196 * 2304 * 256
NC = 4
y4 = PF(n, 14, 14, 1024) if which == "deconv" else None
kd = torch.randn(1024, 256, device="cuda") / 16 if which == "deconv" else None
w1 = torch.randn(256, NC, device="cuda") / 16
b1 = torch.zeros(NC, device="cuda")
masks = torch.empty(n, 28, 28, NC, device="cuda") if which == "deconv" else None
ids = torch.zeros(n, dtype=torch.int32, device="cuda")
if which == "deconv":
    flops = 2.0 * n * 196 * 256 * 1024
def run():
    if which == "deconv":
        C.call("myolo_deconv_mask_fwd", x.rows, kd, bias, w1, b1, masks, ids, y4.rows, n, 14, 14, 256, NC, st)
    elif which == "fwd":
        C.call("myolo_conv3x3_fwd", x.rows, w, y.rows, n, 14, 14, 256, 256, bias, None, None, 0, st)
    elif which == "wgrad":
        C.call("myolo_conv3x3_wgrad", x.rows, y.rows, dw, n, 14, 14, 256, 256, st)
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{which}: n_roi {n}  {ms:.3f} ms/launch  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic, valid pixels only)")
