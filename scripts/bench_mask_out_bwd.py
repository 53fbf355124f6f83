import sys, os
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mask-yolo_b200"))
import torch
from myolo import _cabi as C
from myolo.pf import PF
C.device_check(0)
st = torch.cuda.current_stream().cuda_stream
n, H, W, Cm, NC = 4704, 14, 14, 256, 4
y4 = PF(n, H, W, 4 * Cm); y4.valid().normal_()
bd, w1 = torch.randn(Cm, device="cuda") * 0.1, torch.randn(Cm, NC, device="cuda") / 16
dlogit = torch.zeros(n, 2 * H, 2 * W, NC, device="cuda")
ids = torch.zeros(n, dtype=torch.int32, device="cuda")
pos = [5, 700, 1500, 2222, 3000, 3999, 4700]
for r in pos:
    dlogit[r, :, :, 1] = torch.randn(2 * H, 2 * W, device="cuda") * 3e-5
    ids[r] = 1
gs = torch.tensor([1.0, 1.0, 0, 0], device="cuda")
C.call("myolo_grad_scale", dlogit, dlogit.numel(), gs, st)
out = PF(n, H, W, 4 * Cm, dtype=torch.float16)
prev = torch.zeros(n, dtype=torch.int32, device="cuda")
g = [torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")]
def t(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms")
t(lambda: C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, out.rows, g[0], g[1], g[2], n, H, W, Cm, NC, gs, None, None, st), "no ids")
t(lambda: C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, out.rows, g[0], g[1], g[2], n, H, W, Cm, NC, gs, ids, None, st), "ids")
t(lambda: C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, out.rows, g[0], g[1], g[2], n, H, W, Cm, NC, gs, ids, prev, st), "ids+prev")
