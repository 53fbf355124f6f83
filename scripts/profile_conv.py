"""Runs one mask-head kernel alone at benchmark size (4704 ROIs of 14x14x256) for ncu captures and quick timing:
    python scripts/profile_conv.py [n_roi] [iters] [which]
which: fwd | wgrad | deconv (tf32 operands)  |  fwd_h | dgrad_h | wgrad_h | deconv_h (IEEE-half operands, kind::f16)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import torch
from myolo import _cabi as C
from myolo.pf import PF, conv3x3_shifts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4704
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "fwd_h"
half = which.endswith("_h")
dt = torch.float16 if half else torch.float32
C.device_check(0); C.set_precision(C.PREC_TF32)
st = torch.cuda.current_stream().cuda_stream
x, y = PF(n, 14, 14, 256, dtype=dt), PF(n, 14, 14, 256, dtype=dt)
x.valid().normal_(); y.valid().normal_()
a_out = PF(n, 14, 14, 256, dtype=dt)
a_out.valid().normal_().clamp_(min=0)
w = (torch.randn(9, 256, 256, device="cuda") / 48).to(dt)
bias = torch.zeros(256, device="cuda")
ones = torch.ones(256, device="cuda")
dw = torch.zeros(9, 256, 256, device="cuda")
ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
flops = 2.0 * n * 196 * 2304 * 256
NC = 4
deconv = which.startswith("deconv")
y4 = PF(n, 14, 14, 1024) if deconv else None
kd = (torch.randn(1024, 256, device="cuda") / 16).to(dt) if deconv else None
w1 = torch.randn(256, NC, device="cuda") / 16
b1 = torch.zeros(NC, device="cuda")
masks = torch.empty(n, 28, 28, NC, device="cuda") if deconv else None
ids = torch.zeros(n, dtype=torch.int32, device="cuda")
if deconv:
    flops = 2.0 * n * 196 * 256 * 1024
sh = C.int_array(conv3x3_shifts(14))
shn = C.int_array(conv3x3_shifts(14, negate=True))
M = x.M
def run():
    if which == "deconv":
        C.call("myolo_deconv_mask_fwd", x.rows, kd, bias, w1, b1, masks, ids, y4.rows, n, 14, 14, 256, NC, st)
    elif which == "deconv_h":
        C.call("myolo_deconv_mask_fwd_h", x.rows, kd, bias, w1, b1, masks, ids, y4.rows, n, 14, 14, 256, NC, st)
    elif which == "fwd":
        C.call("myolo_conv3x3_fwd", x.rows, w, y.rows, n, 14, 14, 256, 256, bias, None, None, 0, st)
    elif which == "wgrad":
        C.call("myolo_conv3x3_wgrad", x.rows, y.rows, dw, n, 14, 14, 256, 256, st)
    elif which == "fwd_h":
        C.call("myolo_gemm_taps_h", x.rows, 256, w, None, 0, y.rows, 256, M, 256, 256, 9, sh, bias, ones, bias, C.ACT_RELU, 15, 225, None, st)
    elif which == "dgrad_h":
        C.call("myolo_gemm_taps_bnbwd_h", x.rows, 256, w, None, y.rows, 256, M, 256, 256, 9, shn, 15, 225, a_out.rows, ones, bias, ones,
               1e-3, C.ACT_RELU, dw[0, 0], dw[0, 1], dw[0, 2], ws, None, st)
    elif which == "wgrad_h":
        C.call("myolo_gemm_taps_wgrad_h", x.rows, 256, y.rows, 256, dw, M, 256, 256, 9, sh, 0, None, st)
    else:
        raise SystemExit(f"unknown kernel {which}")
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{which}: n_roi {n}  {ms:.3f} ms/launch  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic, valid pixels only)")
