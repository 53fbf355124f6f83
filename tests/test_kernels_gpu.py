"""Per-kernel parity: every C-ABI entry point of libmyolo_sm100.so against the CPU oracle
(oracle/myolo_oracle.py) on the same seeded inputs.  Index / mask-target / ROI-selection work is
compared bit-exactly; fp32 arithmetic within the tolerance written next to each check; the
tcgen05 (tf32-operand) kernels within 2e-3 of the output scale (tf32 has a 10-bit mantissa)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import myolo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from myolo import _cabi
    _cabi.device_check(0)
    return _cabi


def stream():
    return torch.cuda.current_stream().cuda_stream


def cuda(t):
    return t.contiguous().cuda()


def close(a, b, tol, what=""):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item() / scale
    assert err <= tol, f"{what}: rel-to-max err {err:.3e} > {tol}"


# ----------------------------------------------------------------------------- K2 depthwise
@pytest.mark.parametrize("B,H,W,Cc,s", [(2, 16, 16, 32, 1), (2, 16, 16, 64, 2), (1, 13, 11, 32, 1), (3, 14, 14, 96, 2),
                                        (2, 7, 7, 128, 1)])
def test_dwconv(C, B, H, W, Cc, s):
    torch.manual_seed(0)
    x = torch.randn(B, H, W, Cc)
    w = torch.randn(3, 3, Cc, 1)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y = O.depthwise3x3_nhwc(xr, wr, s)
    dy = torch.randn_like(y)
    y.backward(dy)
    xd, wd, dyd = cuda(x), cuda(w.reshape(3, 3, Cc)), cuda(dy)
    yd = torch.empty(y.shape, device="cuda")
    C.call("myolo_dwconv3x3_fwd", C.view(xd, B, H, W, Cc), wd, yd, s, stream())
    close(yd, y, 1e-6, "dw fwd")
    dxd = torch.empty_like(xd)
    C.call("myolo_dwconv3x3_bwd_data", dyd, wd, dxd, B, H, W, Cc, s, stream())
    close(dxd, xr.grad, 1e-6, "dw bwd data")
    dwd = torch.empty(3, 3, Cc, device="cuda")
    C.call("myolo_dwconv3x3_bwd_filter", C.view(xd, B, H, W, Cc), dyd, dwd, s, stream())
    close(dwd, wr.grad.reshape(3, 3, Cc), 2e-6, "dw bwd filter")


@pytest.mark.parametrize("B,H,W,Cc,s", [(2, 16, 16, 32, 1), (3, 14, 14, 96, 2), (1, 13, 11, 64, 1), (4, 28, 28, 256, 2)])
def test_dwconv_fused_bn(C, B, H, W, Cc, s):
    """Depthwise forward with the producer's BN + ReLU6 applied on load and the batch statistics of its own output taken in
    the epilogue; filter gradient with the same BN on load -- against the oracle's separate steps."""
    torch.manual_seed(40)
    xpre = torch.randn(B, H, W, Cc) * 2 + 0.5
    w = torch.randn(3, 3, Cc, 1)
    mean, var = torch.randn(Cc) * 0.3, torch.rand(Cc) + 0.5
    gamma, beta = torch.rand(Cc) + 0.5, torch.randn(Cc) * 0.5
    a = O.relu6((xpre - mean) / torch.sqrt(var + 1e-3) * gamma + beta)
    wr = w.clone().requires_grad_(True)
    y = O.depthwise3x3_nhwc(a, wr, s)
    dy = torch.randn_like(y)
    y.backward(dy)
    ws = torch.zeros(8192, dtype=torch.float64, device="cuda")
    xd, wd = cuda(xpre), cuda(w.reshape(3, 3, Cc))
    md, vd, gd, bd = cuda(mean), cuda(var), cuda(gamma), cuda(beta)
    yd = torch.empty(y.shape, device="cuda")
    om, ov = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    for rep in range(2):                       # twice: the workspace must come back zeroed
        C.call("myolo_dwconv3x3_fwd_bn", C.view(xd, B, H, W, Cc), wd, yd, s, md, vd, gd, bd, 1e-3, C.ACT_RELU6, om, ov, ws, stream())
        close(yd, y, 5e-6, "fused dw fwd")
        yf = y.detach().reshape(-1, Cc).double()
        close(om, yf.mean(0).float(), 2e-5, "epilogue mean")
        close(ov, yf.var(0, unbiased=False).float(), 2e-5, "epilogue variance")
        assert ws.abs().max().item() == 0
    # statistics only / BN on load only
    a_d = cuda(a)
    C.call("myolo_dwconv3x3_fwd_bn", C.view(a_d, B, H, W, Cc), wd, yd, s, None, None, None, None, 0.0, 0, om, ov, ws, stream())
    close(yd, y, 2e-6, "dw fwd + stats"); close(om, yf.mean(0).float(), 2e-5, "mean")
    C.call("myolo_dwconv3x3_fwd_bn", C.view(xd, B, H, W, Cc), wd, yd, s, md, vd, gd, bd, 1e-3, C.ACT_RELU6, None, None, None, stream())
    close(yd, y, 5e-6, "dw fwd, BN on load")
    dwd = torch.empty(3, 3, Cc, device="cuda")
    C.call("myolo_dwconv3x3_bwd_filter_bn", C.view(xd, B, H, W, Cc), cuda(dy), dwd, s, md, vd, gd, bd, 1e-3, C.ACT_RELU6, stream())
    close(dwd, wr.grad.reshape(3, 3, Cc), 6e-6, "dw bwd filter, BN on load")


@pytest.mark.parametrize("M,K,N,x3", [(1000, 64, 64, 1), (777, 128, 256, 1), (4096 + 37, 32, 128, 0), (300, 512, 1024, 1), (130, 64, 32, 0)])
def test_gemm_epilogue_statistics(C, M, K, N, x3):
    """myolo_gemm_taps_tc_stats: the pointwise GEMM (3xTF32 operand triple or single pass) with per-channel batch mean /
    biased variance of its result reduced in the epilogue, against the same GEMM + a separate statistics pass."""
    torch.manual_seed(41)
    A = torch.randn(M, K) * 1.5 + 0.2
    W = torch.randn(K, N) / K ** 0.5
    Wd = cuda(W)
    if x3:
        hi = A.cuda().clone()
        lo = torch.empty_like(hi)
        C.call("myolo_split_tf32", C.view(hi, 1, 1, M, K), C.view(hi, 1, 1, M, K), C.view(lo, 1, 1, M, K), stream())
        Ad = torch.cat([hi, lo], 0).contiguous()
        Wt = torch.empty(3, N, K, device="cuda")              # [B_hi | B_hi | B_lo], each transposed to [N][K]
        C.call("myolo_prep_weights", Wd, Wt, 1, K, N, 1, 2, stream())
        sh, nt = C.int_array([0, M, 0]), 3
    else:
        Ad, Wt, sh, nt = cuda(A), _prep(C, Wd, 1, K, N, 1, 1), None, 1
    ref = torch.empty(M, N, device="cuda")
    C.call("myolo_gemm_taps_tc", Ad, K, Wt, ref, N, M, N, K, nt, sh, None, None, None, 0, 0, 0, 0, stream())
    out = torch.full((M, N), 7.0, device="cuda")
    mean, var = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    ws = torch.zeros(8192, dtype=torch.float64, device="cuda")
    for rep in range(2):
        C.call("myolo_gemm_taps_tc_stats", Ad, K, Wt, out, N, M, N, K, nt, sh, 0, 0, mean, var, ws, M, stream())
        assert torch.equal(out, ref)
        close(mean, ref.double().mean(0).float(), 2e-5, "epilogue mean")
        close(var, ref.double().var(0, unbiased=False).float(), 2e-5, "epilogue variance")
        assert ws.abs().max().item() == 0
    close(ref, (A.double() @ W.double()).float(), 2e-3 if not x3 else 3e-5, "gemm")


@pytest.mark.parametrize("M,K,N,stats", [(4096 + 37, 64, 128, 1), (6272, 512, 512, 1), (25088, 128, 256, 1), (8200, 32, 1024, 0)])
def test_gemm_k_segments_on_the_window_kernel(C, M, K, N, stats):
    """myolo_gemm_segs_win: the 3xTF32 pointwise GEMM as three k-segments of the persistent window kernel (hi / lo operand
    planes lo_off rows apart in one matrix), with the batch statistics of the result in the epilogue, against the
    one-tile kernel on the same operands."""
    torch.manual_seed(42)
    A = torch.randn(M, K) * 1.5 + 0.2
    W = torch.randn(K, N) / K ** 0.5
    Wd = cuda(W)
    hi = A.cuda().clone()
    lo = torch.empty_like(hi)
    C.call("myolo_split_tf32", C.view(hi, 1, 1, M, K), C.view(hi, 1, 1, M, K), C.view(lo, 1, 1, M, K), stream())
    Ad = torch.cat([hi, lo], 0).contiguous()
    Wt = torch.empty(3, N, K, device="cuda")
    C.call("myolo_prep_weights", Wd, Wt, 1, K, N, 1, 2, stream())
    sh = C.int_array([0, M, 0])
    ref = torch.empty(M, N, device="cuda")
    C.call("myolo_gemm_taps_tc", Ad, K, Wt, ref, N, M, N, K, 3, sh, None, None, None, 0, 0, 0, 0, stream())
    assert C.lib().myolo_gemm_segs_win_supported(K, N, M, N, K, 3) == 1
    out = torch.full((M, N), 7.0, device="cuda")
    mean, var = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    ws = torch.zeros(8192, dtype=torch.float64, device="cuda")
    for rep in range(2):
        C.call("myolo_gemm_segs_win", Ad, K, Wt, out, N, M, N, K, 3, sh, mean if stats else None, var if stats else None,
               ws if stats else None, stream())
        close(out, ref, 2e-6, "k-segment GEMM vs tap GEMM (same products, different accumulation order)")
        if stats:
            close(mean, ref.double().mean(0).float(), 2e-5, "epilogue mean")
            close(var, ref.double().var(0, unbiased=False).float(), 2e-5, "epilogue variance")
            assert ws.abs().max().item() == 0
    close(out, (A.double() @ W.double()).float(), 3e-5, "3xTF32 product")


# ----------------------------------------------------------------------------- K1 stem conv
def test_conv1(C):
    torch.manual_seed(1)
    B, S = 2, 32
    x = torch.rand(B, S, S, 3)
    w = (torch.randn(3, 3, 3, 32) * 0.2).requires_grad_(True)
    y = O.conv2d_nhwc(x, w, stride=2, pad=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    xd, wd = cuda(x), cuda(w.detach())
    yd = torch.empty(y.shape, device="cuda")
    C.call("myolo_conv1_fwd", xd, wd, yd, B, S, 32, stream())
    close(yd, y, 1e-6, "conv1 fwd")
    dwd = torch.empty(3, 3, 3, 32, device="cuda")
    C.call("myolo_conv1_wgrad", xd, cuda(dy), dwd, B, S, 32, stream())
    close(dwd, w.grad, 2e-6, "conv1 wgrad")


# ----------------------------------------------------------------------------- K4 batch norm
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("train", [1, 0])
def test_bn(C, act, train):
    torch.manual_seed(2)
    B, H, W, Cc = 3, 9, 7, 64
    x = (torch.randn(B, H, W, Cc) * 2 + 0.5).requires_grad_(True)
    g = (torch.rand(Cc) + 0.5).requires_grad_(True)
    b = (torch.randn(Cc) * 0.3).requires_grad_(True)
    mm, mv = torch.randn(Cc) * 0.1, torch.rand(Cc) + 0.5
    if train:
        y, mean, var, n = O.bn_train(x, g, b)
    else:
        y, mean, var = O.bn_infer(x, g, b, mm, mv), mm, mv
    fa = {0: lambda t: t, 1: torch.relu, 2: O.relu6}[act]
    ya = fa(y)
    dy = torch.randn_like(ya)
    ya.backward(dy)
    xd, gd, bd = cuda(x.detach()), cuda(g.detach()), cuda(b.detach())
    meand, vard = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    xv = C.view(xd, B, H, W, Cc)
    if train:
        C.call("myolo_bn_stats", xv, meand, vard, ws, stream())
        close(meand, mean, 1e-6, "bn mean")
        close(vard, var, 1e-5, "bn var")
    else:
        meand, vard = cuda(mm), cuda(mv)
    yd = torch.empty_like(xd)
    C.call("myolo_bn_apply", xv, C.view(yd, B, H, W, Cc), meand, vard, gd, bd, 1e-3, act, stream())
    close(yd, ya, 2e-6, "bn apply")
    dxd = torch.empty_like(xd)
    dg, db = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    dyd = cuda(dy)
    C.call("myolo_bn_bwd", xv, C.view(dyd, B, H, W, Cc), C.view(dxd, B, H, W, Cc), meand, vard, gd, bd, 1e-3, act, train,
           dg, db, ws, stream())
    close(dxd, x.grad, 2e-5, "bn dx")
    close(dg, g.grad, 2e-5, "bn dgamma")
    close(db, b.grad, 2e-5, "bn dbeta")


def test_bn_moving_update_and_colsum(C):
    torch.manual_seed(3)
    Cc, n = 32, 5 * 7 * 7
    val = torch.rand(Cc) + 0.1
    biased = torch.zeros(Cc)
    bd, md = cuda(biased), torch.empty(Cc, device="cuda")
    exp_b = biased.clone()
    for step in (1, 2, 3):
        C.call("myolo_bn_moving_update", cuda(val), bd, md, Cc, 0.99, step, 1, float(n), 1e-3, stream())
        v = O.moving_update_value(val, n)
        exp_b = exp_b - (exp_b - v) * (1 - 0.99)
        close(md, exp_b / (1 - 0.99 ** step), 2e-5, "moving var")
    x = torch.randn(4, 5, 6, 64)
    out = torch.empty(64, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    xd = cuda(x)
    C.call("myolo_colsum", C.view(xd, 4, 5, 6, 64), out, ws, stream())
    close(out, x.sum((0, 1, 2)), 1e-5, "colsum")
    x5 = torch.randn(4, 5, 6, 20)
    out5 = torch.empty(20, device="cuda")
    x5d = cuda(x5)
    C.call("myolo_colsum", C.view(x5d, 4, 5, 6, 20), out5, ws, stream())
    close(out5, x5.sum((0, 1, 2)), 1e-5, "colsum generic")


# ----------------------------------------------------------------------------- K7 ROIAlign
def _boxes(n, seed):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g)
    wh = torch.rand(n, 2, generator=g) * 0.6
    b = torch.cat([c - wh / 2, c + wh / 2], 1)        # some fall outside [0,1] -> extrapolation
    b[0] = torch.tensor([0.0, 0.0, 0.0, 0.0])          # zero-padded roi
    b[1] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    b[2] = torch.tensor([0.2, 0.3, 0.2, 0.9])          # degenerate
    b[3] = torch.tensor([-0.5, 0.1, 1.5, 0.8])
    b[4] = torch.tensor([0.7, 0.8, 0.3, 0.2])          # reversed box: sample positions run right to left
    b[5] = torch.tensor([0.1, 0.45, 0.9, 0.55])        # samples several feature columns apart
    b[6] = torch.tensor([0.0, 0.0, 12.0 / 11.0 * 0.5, 0.5])
    return b


@pytest.mark.parametrize("Cc", [64, 128, 256])      # 128 / 256: the run-length backward kernel; 64: one reduction per sample
def test_roialign(C, Cc):
    torch.manual_seed(4)
    B, Fh, R, P = 2, 12, 9, 14
    feat = torch.randn(B, Fh, Fh, Cc).requires_grad_(True)
    boxes = _boxes(B * R, 5)
    idx = torch.arange(B).repeat_interleave(R)
    out = O.crop_and_resize(feat, boxes, idx, P, P)
    dout = torch.randn_like(out)
    out.backward(dout)
    fd, bd = cuda(feat.detach()), cuda(boxes)
    od = torch.empty(B * R, P, P, Cc, device="cuda")
    C.call("myolo_roialign_fwd", C.view(fd, B, Fh, Fh, Cc), bd, B * R, R, P, C.view(od, B * R, P, P, Cc), 0, stream())
    assert torch.equal(od.cpu(), out.detach()), "ROIAlign forward must be bit-exact vs the oracle"
    dfd = torch.zeros_like(fd)
    dod = cuda(dout)
    C.call("myolo_roialign_bwd", C.view(dod, B * R, P, P, Cc), bd, B * R, R, P, C.view(dfd, B, Fh, Fh, Cc), stream())
    close(dfd, feat.grad, 1e-5, "roialign bwd")
    if Cc >= 128:     # the same from a half, loss-scaled gradient (h16 mode): the scale is removed while reading
        from myolo.pf import PF
        dh = PF(B * R, P, P, Cc, dtype=torch.float16)
        dh.valid().copy_(dod * 64.0)
        ref = torch.zeros_like(fd)
        C.call("myolo_roialign_bwd", C.view(dh.valid().float().contiguous() / 64.0, B * R, P, P, Cc), bd, B * R, R, P,
               C.view(ref, B, Fh, Fh, Cc), stream())
        dfh = torch.zeros_like(fd)
        us = torch.tensor([1.0 / 64.0], device="cuda")
        C.call("myolo_roialign_bwd_h", dh.view(), bd, B * R, R, P, C.view(dfh, B, Fh, Fh, Cc), us, stream())
        close(dfh, ref, 5e-6, "roialign bwd from the half gradient (fp32 atomics order)")
        with pytest.raises(C.MyoloError):
            C.call("myolo_roialign_bwd_h", dh.view(), bd, B * R, R, P, C.view(torch.zeros(B, Fh, Fh, 64, device="cuda"), B, Fh, Fh, 64),
                   us, stream())


# ----------------------------------------------------------------------------- tap-GEMM family
def _prep(C, w, ntaps, rows, cols, transpose, rnd=0):
    out = torch.empty(ntaps, cols if transpose else rows, rows if transpose else cols, device="cuda")
    C.call("myolo_prep_weights", w, out, ntaps, rows, cols, transpose, rnd, stream())
    return out


@pytest.mark.parametrize("impl,tol", [("ffma", 2e-6), ("tc", 2e-3)])
@pytest.mark.parametrize("M,K,N", [(300, 64, 64), (1000, 32, 128), (257, 256, 512), (64, 1024, 32), (4100, 128, 256)])
def test_gemm_pointwise(C, impl, tol, M, K, N):
    torch.manual_seed(6)
    A = torch.randn(M, K)
    W = torch.randn(K, N) / K ** 0.5
    bias = torch.randn(N)
    ref = torch.relu(A.double() @ W.double() + bias.double()).float()
    Ad, Wd = cuda(A), cuda(W)
    Wt = _prep(C, Wd, 1, K, N, 1)
    assert torch.equal(Wt[0].cpu(), W.t().contiguous())
    out = torch.full((M, N), 7.0, device="cuda")
    C.call(f"myolo_gemm_taps_{impl}", Ad, K, Wt, out, N, M, N, K, 1, None, cuda(bias), None, None, C.ACT_RELU, 0, 0, 0,
           stream())
    close(out, ref, tol, f"pointwise {impl}")
    # dgrad form + accumulate
    dY = torch.randn(M, N)
    base = torch.randn(M, K)
    dX = cuda(base)
    C.call(f"myolo_gemm_taps_{impl}", cuda(dY), N, Wd, dX, K, M, K, N, 1, None, None, None, None, 0, 0, 0, 1, stream())
    close(dX, (base.double() + dY.double() @ W.double().t()).float(), tol, f"dgrad {impl}")


@pytest.mark.parametrize("impl,tol", [("ffma", 3e-6), ("tc", 2e-3)])
@pytest.mark.parametrize("n,H,W,Ci,Co", [(3, 14, 14, 64, 64), (2, 7, 9, 32, 128), (5, 14, 14, 256, 256)])
def test_conv3x3_pf(C, impl, tol, n, H, W, Ci, Co):
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(7)
    x = torch.randn(n, H, W, Ci)
    w = (torch.randn(3, 3, Ci, Co) / (9 * Ci) ** 0.5).requires_grad_(True)
    bias = torch.randn(Co)
    xr = x.clone().requires_grad_(True)
    y = O.conv2d_nhwc(xr, w, 1, 1) + bias
    dy = torch.randn_like(y)
    y.backward(dy)
    px = PF(n, H, W, Ci).load_dense(cuda(x))
    py = PF(n, H, W, Co)
    wd = cuda(w.detach().reshape(9, Ci, Co))
    wt = _prep(C, wd, 9, Ci, Co, 1)
    sh = C.int_array(conv3x3_shifts(W))
    M = px.M
    C.call(f"myolo_gemm_taps_{impl}", px.rows, Ci, wt, py.rows, Co, M, Co, Ci, 9, sh, cuda(bias), None, None, 0, W + 1,
           (H + 1) * (W + 1), 0, stream())
    close(py.dense(), y, tol, f"conv3x3 fwd {impl}")
    full = py.rows.view(n, H + 1, W + 1, Co)
    assert full[:, 0].abs().max().item() == 0 and full[:, :, 0].abs().max().item() == 0, "pad rows must stay zero"
    # dgrad
    pdy = PF(n, H, W, Co).load_dense(cuda(dy))
    pdx = PF(n, H, W, Ci)
    shn = C.int_array(conv3x3_shifts(W, negate=True))
    C.call(f"myolo_gemm_taps_{impl}", pdy.rows, Co, wd, pdx.rows, Ci, M, Ci, Co, 9, shn, None, None, None, 0, W + 1,
           (H + 1) * (W + 1), 0, stream())
    close(pdx.dense(), xr.grad, tol, f"conv3x3 dgrad {impl}")
    # wgrad
    dw = torch.zeros(9, Ci, Co, device="cuda")
    C.call(f"myolo_gemm_taps_wgrad_{impl}",      # tcgen05 path: Ci only needs whole 32-channel boxes
           px.rows, Ci, pdy.rows, Co, dw, M, Co, Ci, 9, sh, 0, stream())
    close(dw, w.grad.reshape(9, Ci, Co), tol, f"conv3x3 wgrad {impl}")
    dwt = torch.zeros(9, Co, Ci, device="cuda")
    C.call(f"myolo_gemm_taps_wgrad_{impl}",      # tcgen05 path: Ci only needs whole 32-channel boxes
           px.rows, Ci, pdy.rows, Co, dwt, M, Co, Ci, 9, sh, 1, stream())
    close(dwt, w.grad.reshape(9, Ci, Co).transpose(1, 2), tol, f"conv3x3 wgrad^T {impl}")


def test_named_conv_wrappers(C):
    """myolo_pwconv_* / myolo_conv3x3_* in both precision modes."""
    from myolo.pf import PF
    torch.manual_seed(8)
    M, Ci, Co = 777, 128, 64
    x, w, dy = torch.randn(M, Ci), torch.randn(Ci, Co) / Ci ** 0.5, torch.randn(M, Co)
    for mode, tol in ((C.PREC_FP32, 3e-6), (C.PREC_TF32, 2e-3)):
        C.set_precision(mode)
        try:
            wd = cuda(w)
            wt = _prep(C, wd, 1, Ci, Co, 1)
            y = torch.empty(M, Co, device="cuda")
            C.call("myolo_pwconv_fwd", cuda(x), wt, y, M, Ci, Co, None, stream())
            close(y, x @ w, tol, "pwconv fwd")
            dx = torch.empty(M, Ci, device="cuda")
            C.call("myolo_pwconv_dgrad", cuda(dy), wd, dx, M, Ci, Co, stream())
            close(dx, dy @ w.t(), tol, "pwconv dgrad")
            dw = torch.empty(Ci, Co, device="cuda")
            C.call("myolo_pwconv_wgrad", cuda(x), cuda(dy), dw, M, Ci, Co, stream())
            close(dw, x.t() @ dy, tol, "pwconv wgrad")
            n, H, W = 4, 14, 14
            xi = torch.randn(n, H, W, Ci)
            k = torch.randn(3, 3, Ci, Co) / (9 * Ci) ** 0.5
            px, py = PF(n, H, W, Ci).load_dense(cuda(xi)), PF(n, H, W, Co)
            kt = _prep(C, cuda(k.reshape(9, Ci, Co)), 9, Ci, Co, 1)
            C.call("myolo_conv3x3_fwd", px.rows, kt, py.rows, n, H, W, Ci, Co, None, None, None, 0, stream())
            close(py.dense(), O.conv2d_nhwc(xi, k, 1, 1), tol, "conv3x3 fwd wrapper")
        finally:
            C.set_precision(C.PREC_FP32)


# ----------------------------------------------------------------------------- K10/K11 mask tail
def test_mask_out(C):
    from myolo.pf import PF
    torch.manual_seed(9)
    n, H, W, Cm, NC = 3, 14, 14, 256, 4
    a4 = torch.randn(n, H, W, Cm)
    kd = (torch.randn(2, 2, Cm, Cm) / Cm ** 0.5).requires_grad_(True)       # [a,b,co,ci]
    bd = (torch.randn(Cm) * 0.1).requires_grad_(True)
    w1 = (torch.randn(Cm, NC) / Cm ** 0.5).requires_grad_(True)
    b1 = (torch.randn(NC) * 0.1).requires_grad_(True)
    a4r = a4.clone().requires_grad_(True)
    y = F.conv_transpose2d(a4r.permute(0, 3, 1, 2), kd.permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    h = torch.relu(y + bd)
    logit = h @ w1 + b1
    masks = torch.sigmoid(logit)
    dlogit = torch.zeros_like(logit)
    dlogit[0, :, :, 2] = torch.randn(2 * H, 2 * W)
    dlogit[2, :, :, 1] = torch.randn(2 * H, 2 * W)
    logit.backward(dlogit)
    pa = PF(n, H, W, Cm).load_dense(cuda(a4))
    y4 = PF(n, H, W, 4 * Cm)
    kdd = cuda(kd.detach().reshape(4 * Cm, Cm))
    C.call("myolo_gemm_taps_ffma", pa.rows, Cm, kdd, y4.rows, 4 * Cm, pa.M, 4 * Cm, Cm, 1, None, None, None, None, 0,
           W + 1, (H + 1) * (W + 1), 0, stream())
    md = torch.empty(n, 2 * H, 2 * W, NC, device="cuda")
    C.call("myolo_mask_out_fwd", y4.rows, cuda(bd.detach()), cuda(w1.detach()), cuda(b1.detach()), md, n, H, W, Cm, NC, stream())
    close(md, masks, 5e-6, "mask_out fwd")
    dy4 = PF(n, H, W, 4 * Cm)
    dw1, db1, dbd = torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")
    C.call("myolo_mask_out_bwd", y4.rows, cuda(bd.detach()), cuda(w1.detach()), cuda(dlogit), dy4.rows, dw1, db1, dbd,
           n, H, W, Cm, NC, stream())
    close(dw1, w1.grad, 2e-5, "dw1")
    close(db1, b1.grad, 2e-5, "db1")
    close(dbd, bd.grad, 2e-5, "dbd")
    # dy4 -> da4 (dgrad GEMM, Bt = Kd^T per (a,b,co) -> [ci][(a,b,co)]) and dKd (wgrad, transposed output)
    kdt = _prep(C, kdd, 1, 4 * Cm, Cm, 1)                                   # [Cm][4Cm]
    da = PF(n, H, W, Cm)
    C.call("myolo_gemm_taps_ffma", dy4.rows, 4 * Cm, kdt, da.rows, Cm, pa.M, Cm, 4 * Cm, 1, None, None, None, None, 0,
           W + 1, (H + 1) * (W + 1), 0, stream())
    close(da.dense(), a4r.grad, 2e-5, "deconv dgrad")
    dkd = torch.zeros(4 * Cm, Cm, device="cuda")
    C.call("myolo_gemm_taps_wgrad_ffma", pa.rows, Cm, dy4.rows, 4 * Cm, dkd, pa.M, 4 * Cm, Cm, 1, None, 1, stream())
    close(dkd, kd.grad.reshape(4 * Cm, Cm), 2e-5, "deconv wgrad")


# ----------------------------------------------------------------------------- K12 decode
CFG = dict(GRID_H=7, GRID_W=7, N_BOX=3, NUM_CLASSES=4, ANCHORS=[0.6, 0.9, 1.5, 1.4, 2.5, 2.8], TRAIN_ROIS_PER_IMAGE=147,
           MASK_SHAPE=[28, 28], MASK_POOL_SIZE=14, COORD_SCALE=1.0, NO_OBJECT_SCALE=1.0, OBJECT_SCALE=5.0,
           CLASS_SCALE=1.0, CLASS_WEIGHTS=np.ones(4, dtype="float32"), WARM_UP_BATCHES=0, TRUE_BOX_BUFFER=15)


def test_yolo_decode(C):
    torch.manual_seed(10)
    B = 3
    yp = torch.randn(B, 7, 7, 3, 9)
    boxes = O.decode_yolo(yp, CFG)
    det = O.detections_layer(yp, CFG)
    bd, dd = torch.empty(B, 147, 4, device="cuda"), torch.empty(B, 147, 6, device="cuda")
    C.call("myolo_yolo_decode", cuda(yp), cuda(torch.tensor(CFG["ANCHORS"])), bd, dd, B, 7, 7, 3, 4, stream())
    close(bd, boxes, 2e-6, "decode boxes")
    close(dd[..., :5], det[..., :5], 2e-6, "detections")
    assert torch.equal(dd[..., 5].cpu(), det[..., 5])


# ----------------------------------------------------------------------------- K13 targets
def _shapes_batch(B, S, M, seed):
    """Synthetic GT: axis-aligned rectangles / discs with pixel boxes (x1,y1,x2,y2), x2/y2 exclusive."""
    rng = np.random.RandomState(seed)
    ids = np.zeros((B, M), np.int32)
    boxes = np.zeros((B, M, 4), np.float32)
    masks = np.zeros((B, S, S, M), np.uint8)
    yy, xx = np.mgrid[0:S, 0:S]
    for b in range(B):
        for m in range(rng.randint(1, 4)):
            cx, cy = rng.randint(20, S - 20, 2)
            r = rng.randint(8, S // 4)
            if rng.rand() < 0.5:
                mk = (np.abs(xx - cx) <= r) & (np.abs(yy - cy) <= r)
            else:
                mk = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
            masks[b, :, :, m] = mk
            ys, xs = np.where(mk)
            boxes[b, m] = [xs.min(), ys.min(), xs.max() + 1, ys.max() + 1]
            ids[b, m] = rng.randint(1, 4)
    return ids, boxes, masks


def _proposals_near(boxes_px, S, R, seed):
    rng = np.random.RandomState(seed)
    B = boxes_px.shape[0]
    props = rng.rand(B, R, 4).astype(np.float32)
    props[..., 2:] = props[..., :2] + rng.rand(B, R, 2).astype(np.float32) * 0.5
    for b in range(B):
        for j in range(0, R, 3):                       # every third proposal is a jittered GT box
            g = boxes_px[b, rng.randint(0, 3)]
            if g[2] <= g[0]:
                continue
            props[b, j] = g / S + rng.randn(4).astype(np.float32) * 0.02
    props[0, 5] = np.nan                               # NaN proposal -> neither pos nor neg
    return props


def test_detect_mask_targets_bit_exact(C):
    B, S, M, R = 4, 128, 10, 48
    ids, boxes, masks = _shapes_batch(B, S, M, 11)
    props = _proposals_near(boxes, S, R, 12)
    cfg = dict(CFG, TRAIN_ROIS_PER_IMAGE=R)
    gtb = O.norm_boxes_graph(torch.tensor(boxes), S, S)
    rois, tids, tm = O.detect_mask_targets(torch.tensor(props), torch.tensor(ids), gtb, torch.tensor(masks).bool(), cfg)
    rd = torch.empty(B, R, 4, device="cuda")
    td = torch.empty(B, R, dtype=torch.int32, device="cuda")
    md = torch.empty(B, R, 28, 28, device="cuda")
    npos = torch.empty(B, dtype=torch.int32, device="cuda")
    src = torch.empty(B, R, dtype=torch.int32, device="cuda")
    rgt = torch.empty(B, R, dtype=torch.int32, device="cuda")
    C.call("myolo_detect_mask_targets", cuda(torch.tensor(props)), cuda(torch.tensor(ids)), cuda(torch.tensor(boxes)),
           cuda(torch.tensor(masks)), B, R, M, M, S, 28, 28, rd, td, md, npos, src, rgt, stream())
    assert (tids > 0).sum() > 10, "test should exercise positives"
    assert torch.equal(td.cpu(), tids), "target class ids must be bit-exact"
    assert torch.equal(rd.cpu().view(torch.int32), rois.view(torch.int32)), "roi selection/order must be bit-exact"
    assert torch.equal(md.cpu(), tm), "mask targets must be bit-exact"
    assert torch.equal(npos.cpu().long(), (tids > 0).sum(1))


# ----------------------------------------------------------------------------- K15 / K16 losses
def _yolo_inputs(B, seed):
    rng = np.random.RandomState(seed)
    G, NB, NC, TB = 7, 3, 4, 15
    yt = np.zeros((B, G, G, NB, 5 + NC), np.float32)
    tb = np.zeros((B, 1, 1, 1, TB, 4), np.float32)
    for b in range(B):
        for k in range(rng.randint(1, 4)):
            cx, cy = rng.rand(2) * G
            w, h = rng.rand(2) * 3 + 0.3
            a = rng.randint(NB)
            yt[b, int(cy), int(cx), a, :5] = [cx, cy, w, h, 1]
            yt[b, int(cy), int(cx), a, 5 + rng.randint(1, NC)] = 1
            tb[b, 0, 0, 0, k] = [cx, cy, w, h]
    yp = rng.randn(B, G, G, NB, 5 + NC).astype(np.float32)
    return torch.tensor(yt), torch.tensor(yp), torch.tensor(tb)


@pytest.mark.parametrize("warm", [0, 1])
def test_yolo_loss(C, warm):
    B = 4
    yt, yp, tb = _yolo_inputs(B, 13)
    cfg = dict(CFG, WARM_UP_BATCHES=10 if warm else 0)
    ypr = yp.clone().requires_grad_(True)
    loss = O.yolo_custom_loss(yt, ypr, tb, cfg, seen=1.0)
    loss.backward()
    lo = torch.empty(5, device="cuda")
    dyp = torch.empty_like(yp, device="cuda")
    ws = torch.zeros(8, dtype=torch.float64, device="cuda")
    sc = C.float_array([cfg["OBJECT_SCALE"], cfg["NO_OBJECT_SCALE"], cfg["COORD_SCALE"], cfg["CLASS_SCALE"]])
    C.call("myolo_yolo_loss", cuda(yt), cuda(yp), cuda(tb.reshape(B, 15, 4)), cuda(torch.tensor(cfg["ANCHORS"])),
           cuda(torch.tensor(cfg["CLASS_WEIGHTS"])), B, 7, 7, 3, 4, 15, sc, warm, 1.0, lo, dyp, ws, stream())
    assert abs(lo[0].item() - loss.item()) <= 1e-5 * max(1.0, abs(loss.item())), (lo[0].item(), loss.item())
    close(dyp, ypr.grad, 2e-5, "yolo loss grad")


def test_mask_loss(C):
    torch.manual_seed(14)
    n, NC = 12, 4
    logits = torch.randn(n, 28, 28, NC) * 3
    logits[0, 0, 0, :] = 40.0                          # saturated -> clip region
    lr = logits.clone().requires_grad_(True)
    masks = torch.sigmoid(lr)
    tm = (torch.rand(n, 28, 28) > 0.5).float()
    ids = torch.tensor([2, 0, 1, 0, 3, 0, 0, 1, 0, 0, 2, 0], dtype=torch.int32)
    loss = O.myolo_mask_loss_graph(tm[None], ids[None], masks[None])
    loss.backward()
    lo = torch.empty(1, device="cuda")
    dl = torch.empty(n, 28, 28, NC, device="cuda")
    ws = torch.zeros(2, dtype=torch.float64, device="cuda")
    C.call("myolo_mask_loss", cuda(masks.detach()), cuda(tm), cuda(ids), n, 28, 28, NC, 1.0, lo, dl, ws, stream())
    assert abs(lo.item() - loss.item()) <= 2e-6 * max(1.0, loss.item())
    close(dl, lr.grad, 5e-5, "mask loss dlogit")
    ids0 = torch.zeros(n, dtype=torch.int32)
    C.call("myolo_mask_loss", cuda(masks.detach()), cuda(tm), cuda(ids0), n, 28, 28, NC, 1.0, lo, dl, ws, stream())
    assert lo.item() == 0.0 and dl.abs().max().item() == 0.0


# ----------------------------------------------------------------------------- K17 Adam
def test_adam(C):
    torch.manual_seed(15)
    n = 1003
    p, g = torch.randn(n), torch.randn(n)
    m, v = torch.zeros(n), torch.zeros(n)
    pd, md, vd = cuda(p), cuda(m), cuda(v)
    for t in (1, 2, 3):
        lr_t = 1e-3 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        p = p - lr_t * m / (v.sqrt() + 1e-8)
        C.call("myolo_adam_step", pd, cuda(g), md, vd, n, lr_t, 0.9, 0.999, 1e-8, 1.0, stream())
    close(pd, p, 1e-6, "adam")


def test_error_paths(C):
    """Reference-style error behaviour: bad arguments raise, the message names the failed check."""
    x = torch.zeros(1, 4, 4, 30, device="cuda")
    with pytest.raises(C.MyoloError, match="argument check failed"):
        C.call("myolo_dwconv3x3_fwd", C.view(x, 1, 4, 4, 30), x, x, 1, stream())
    with pytest.raises(C.MyoloError):
        C.call("myolo_conv1_fwd", torch.zeros(4), x, x, 1, 4, 32, stream())   # CPU tensor: no CPU path


@pytest.mark.parametrize("n,H,W,Ci", [(40, 14, 14, 256), (3, 14, 14, 64), (300, 14, 14, 256), (17, 7, 9, 128)])
def test_conv3x3_window_kernel(C, n, H, W, Ci):
    """Persistent windowed tcgen05 conv (conv_win_tcgen05.cu) vs the exact CUDA-core kernel: forward
    with the fused bias/BN/ReLU epilogue, and the dgrad form (negated shifts)."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(20)
    Co = 256
    px = PF(n, H, W, Ci)
    px.valid().normal_()
    w = torch.randn(9, Ci, Co, device="cuda") / (9 * Ci) ** 0.5
    wt = _prep(C, w, 9, Ci, Co, 1)
    bias, scale, shift = (torch.randn(Co, device="cuda") * 0.1, torch.rand(Co, device="cuda") + 0.5,
                          torch.randn(Co, device="cuda") * 0.1)
    sh = C.int_array(conv3x3_shifts(W))
    args = (Co, Ci, 9, sh, bias, scale, shift, C.ACT_RELU, W + 1, (H + 1) * (W + 1), 0, stream())
    ref, out = PF(n, H, W, Co), PF(n, H, W, Co)
    C.call("myolo_gemm_taps_ffma", px.rows, Ci, wt, ref.rows, Co, px.M, *args)
    assert C.lib().myolo_gemm_taps_win_supported(Ci, Co, px.M, Co, Ci, 9, ctypes.addressof(sh), 0) == 1
    C.call("myolo_gemm_taps_win", px.rows, Ci, wt, out.rows, Co, px.M, *args)
    close(out.rows, ref.rows, 2e-3, "windowed conv fwd")
    assert out.storage[:px.C].abs().max().item() == 0
    # dgrad form: A = dy [M, 256], Bt = w[t] as [N=Ci? no: N must be 256] -> use square case only
    if Ci == 256:
        shn = C.int_array(conv3x3_shifts(W, negate=True))
        a2 = (Co, Ci, 9, shn, None, None, None, 0, W + 1, (H + 1) * (W + 1), 0, stream())
        r2, o2 = PF(n, H, W, Ci), PF(n, H, W, Ci)
        C.call("myolo_gemm_taps_ffma", ref.rows, Co, w, r2.rows, Ci, px.M, *a2)
        C.call("myolo_gemm_taps_win", ref.rows, Co, w, o2.rows, Ci, px.M, *a2)
        close(o2.rows, r2.rows, 2e-3, "windowed conv dgrad")


@pytest.mark.parametrize("n,NC", [(37, 4), (150, 7), (150, 81)])
def test_deconv_mask_fused(C, n, NC):
    """tcgen05 (tf32) deconv GEMM with the mask tail in its epilogue vs the unfused exact path: exact-fp32 FMA chains in
    registers up to seven classes, a second (half-operand) tcgen05 GEMM beyond."""
    from myolo.pf import PF
    torch.manual_seed(21)
    H, W, Cm = 14, 14, 256
    pa = PF(n, H, W, Cm)
    pa.valid().normal_()
    kd = torch.randn(4 * Cm, Cm, device="cuda") / Cm ** 0.5
    bd, w1, b1 = torch.randn(Cm, device="cuda") * 0.1, torch.randn(Cm, NC, device="cuda") / Cm ** 0.5, torch.randn(NC, device="cuda") * 0.1
    ids = torch.zeros(n, dtype=torch.int32, device="cuda")
    ids[[0, 5, n - 1]] = torch.tensor([1, 3, 2], dtype=torch.int32, device="cuda")
    y_ref = PF(n, H, W, 4 * Cm)
    C.call("myolo_gemm_taps_ffma", pa.rows, Cm, kd, y_ref.rows, 4 * Cm, pa.M, 4 * Cm, Cm, 1, None, None, None, None, 0,
           W + 1, (H + 1) * (W + 1), 0, stream())
    m_ref = torch.empty(n, 2 * H, 2 * W, NC, device="cuda")
    C.call("myolo_mask_out_fwd", y_ref.rows, bd, w1, b1, m_ref, n, H, W, Cm, NC, stream())
    y4 = PF(n, H, W, 4 * Cm)
    m = torch.full((n, 2 * H, 2 * W, NC), -1.0, device="cuda")
    assert C.lib().myolo_deconv_mask_fwd_supported(Cm, NC) == 1
    C.call("myolo_deconv_mask_fwd", pa.rows, kd, bd, w1, b1, m, ids, y4.rows, n, H, W, Cm, NC, stream())
    torch.cuda.synchronize()
    assert m.min().item() >= 0.0, "every mask element was written"
    close(m, m_ref, 1e-3, "fused masks")
    yv, rv = y4.valid(), y_ref.valid()
    for r in range(n):
        if ids[r] > 0:
            close(yv[r], rv[r], 2e-3, "y4 of a positive roi")
        else:
            assert yv[r].abs().max().item() == 0, "y4 rows of non-positive rois are not written"


def test_detect_postprocess_matches_numpy_pipeline(C):
    """Device zero-area drop / top-k / threshold / NMB / mask paste against the host functions: myolo_utils.NMB (88-113) and
    myolo_utils.unmold_mask (883-912: int() truncation, clamp, resize into the CLIPPED box), both pinned to the reference's
    own source (tests/test_reference_golden.py, tests/test_reference_graph_golden.py)."""
    from myolo import myolo_utils as mu
    for seed, (B, R, NC, S, K), thr, nms, spread in ((30, (3, 147, 4, 224, 10), 0.9, 0.5, 0.8), (31, (2, 60, 4, 96, 10), 0.0, 2.0, 1.0)):
        rng = np.random.RandomState(seed)
        det = np.zeros((B, R, 6), np.float32)
        c = rng.rand(B, R, 2) * spread + (1 - spread) / 2          # spread 1.0: many boxes leave the image
        wh = rng.rand(B, R, 2) * 0.4 + 0.05
        det[..., 0:2], det[..., 2:4] = c - wh / 2, c + wh / 2
        det[:, :20, :4] = det[:, 20:40, :4] + rng.randn(B, 20, 4).astype(np.float32) * 0.01     # near-duplicates -> suppression
        det[..., 4] = rng.permutation(B * R).reshape(B, R) / float(B * R)                        # distinct scores
        det[..., 5] = rng.randint(0, NC, (B, R))
        top = np.argsort(det[0, :, 4])[::-1]
        det[0, top[1], 2] = det[0, top[1], 0]                                                    # a high-scoring box without area
        masks = rng.rand(B, R, 28, 28, NC).astype(np.float32)
        dd, md = cuda(torch.tensor(det)), cuda(torch.tensor(masks))
        i32 = lambda *sh: torch.empty(sh, dtype=torch.int32, device="cuda")      # noqa: E731
        idx, boxes, cls, cnt = i32(B, K), i32(B, K, 4), i32(B, K), i32(B)
        score = torch.empty(B, K, device="cuda")
        pm = torch.empty(B, K, S, S, dtype=torch.uint8, device="cuda")
        C.call("myolo_detect_postprocess", dd, md, B, R, NC, S, 28, 28, K, thr, nms, idx, boxes, cls, score, cnt, pm, stream())
        total_px = mism = 0
        for b in range(B):
            area = (det[b, :, 2] - det[b, :, 0]) * (det[b, :, 3] - det[b, :, 1])
            live = np.where(area > 0)[0]                              # decode_masks: np.delete of the zero-area rows
            order = live[np.argsort(det[b, live, 4])[::-1][:K]]
            order = [i for i in order if det[b, i, 4] >= thr]
            keep = list(mu.NMB(det[b, order, :4], det[b, order, 5], np.asarray(order), (S, S, 3), nms_threshold=nms)) if order else []
            n = int(cnt[b].item())
            assert idx[b, :n].cpu().tolist() == [int(k) for k in keep], (b, idx[b].cpu().tolist(), keep)
            assert (idx[b, n:] == -1).all()
            if b == 0:
                assert int(top[1]) not in idx[b].cpu().tolist()
            px = (det[b, keep, :4] * np.float32(S)).astype(np.int32)  # int() truncation
            exp_boxes = np.stack([np.clip(px[:, 0], 0, S), np.clip(px[:, 1], 0, S), np.clip(px[:, 2], 1, S), np.clip(px[:, 3], 1, S)], 1) \
                if keep else np.zeros((0, 4), np.int32)
            assert np.array_equal(boxes[b, :n].cpu().numpy(), exp_boxes)
            assert cls[b, :n].cpu().tolist() == [int(det[b, k, 5]) for k in keep]
            for j, k in enumerate(keep):
                ref = mu.unmold_mask(masks[b, k, :, :, int(det[b, k, 5])], det[b, k, :4], (S, S, 3))
                got = pm[b, j].bool().cpu().numpy()
                total_px += ref.sum()
                mism += (ref != got).sum()
            assert pm[b, n:].sum().item() == 0
        assert total_px > 1000 and mism <= 2e-3 * total_px, (mism, total_px)     # cv2 vs device rounding at exactly 0.5


def test_dgrad_with_fused_bn_backward(C):
    """myolo_gemm_taps_bnbwd (dgrad GEMM + BN/ReLU backward in the epilogue) against the exact two-step path:
    CUDA-core dgrad followed by myolo_bn_act_bwd_from_output."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(22)
    n, H, W, Cc = 40, 14, 14, 256
    g_in = PF(n, H, W, Cc)
    g_in.valid().normal_()
    a_out = PF(n, H, W, Cc)
    a_out.valid().copy_(torch.relu(torch.randn(n, H, W, Cc, device="cuda")))
    w = torch.randn(9, Cc, Cc, device="cuda") / (9 * Cc) ** 0.5
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    var = torch.rand(Cc, device="cuda") + 0.5
    shn = C.int_array(conv3x3_shifts(W, negate=True))
    pfw, pfb, M = W + 1, (H + 1) * (W + 1), g_in.M
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    ref = PF(n, H, W, Cc)
    C.call("myolo_gemm_taps_ffma", g_in.rows, Cc, w, ref.rows, Cc, M, Cc, Cc, 9, shn, None, None, None, 0, pfw, pfb, 0, stream())
    dg_r, db_r, dbias_r = (torch.empty(Cc, device="cuda") for _ in range(3))
    C.call("myolo_bn_act_bwd_from_output", a_out.view(), ref.view(), ref.view(), gamma, beta, var, 1e-3, C.ACT_RELU, dg_r, db_r,
           dbias_r, ws, stream())
    out = PF(n, H, W, Cc)
    dg, db, dbias = (torch.empty(Cc, device="cuda") for _ in range(3))
    assert C.lib().myolo_gemm_taps_bnbwd_supported(Cc, Cc, M, Cc, Cc, 9, ctypes.addressof(shn)) == 1
    C.call("myolo_gemm_taps_bnbwd", g_in.rows, Cc, w, out.rows, Cc, M, Cc, Cc, 9, shn, pfw, pfb, a_out.rows, gamma, beta, var,
           1e-3, C.ACT_RELU, dg, db, dbias, ws, stream())
    close(out.rows, ref.rows, 2e-3, "fused d(pre-BN)")
    close(db, db_r, 2e-3, "fused dbeta")
    close(dg, dg_r, 2e-3, "fused dgamma")
    close(dbias, dbias_r, 2e-3, "fused dbias")
    assert ws.abs().max().item() == 0, "BN workspace must be left zero"
    assert out.storage[:Cc].abs().max().item() == 0 and out.rows.view(n, H + 1, W + 1, Cc)[:, 0].abs().max().item() == 0


def test_target_encoding_on_device_matches_batch_generator(C):
    """myolo_extract_bboxes / myolo_encode_yolo_targets against the host BatchGenerator (bit-exact)."""
    from myolo import myolo_utils as mu
    from myolo.shapes import ShapesConfig, make_batches

    class Cfg(ShapesConfig):
        BATCH_SIZE = 8
    cfg = Cfg()
    images, tb, yt, ids, boxes, masks = make_batches(cfg, 1, seed=77)[0]
    assert (ids > 0).sum() >= 8
    tb_d, yt_d, boxes_d = mu.encode_targets_device(cfg, torch.tensor(ids).cuda(), None, torch.tensor(masks).cuda())
    assert np.array_equal(boxes_d.cpu().numpy(), boxes), "extract_bboxes"
    assert np.array_equal(yt_d.cpu().numpy(), yt.astype(np.float32)), "yolo_target"
    assert np.array_equal(tb_d.cpu().numpy(), tb.astype(np.float32)), "true_boxes"


def test_staging_helpers_split_fold_copy_and_batched_moving_update(C):
    """The small layout / precision staging entry points: tf32 hi/lo split (3xTF32 operands), BN folding,
    strided view copy (dense <-> padded-flat, accumulate, tf32 rounding) and the batched moving-average update."""
    import struct
    from myolo.pf import PF
    torch.manual_seed(23)
    n, H, W, Cc = 3, 6, 5, 32
    x = torch.randn(n, H, W, Cc, device="cuda")
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    xv = C.view(x, n, H, W, Cc)
    C.call("myolo_split_tf32", xv, C.view(hi, n, H, W, Cc), C.view(lo, n, H, W, Cc), stream())
    assert (hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0 and (lo.view(torch.int32) & 0x1FFF).abs().max().item() == 0
    assert ((hi + lo) - x).abs().max().item() <= 2 ** -21 * x.abs().max().item()        # hi+lo recovers ~22 mantissa bits
    # split BN apply == split of the plain BN apply
    g, b = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    mean, var = torch.randn(Cc, device="cuda") * 0.1, torch.rand(Cc, device="cuda") + 0.5
    y = torch.empty_like(x)
    C.call("myolo_bn_apply", xv, C.view(y, n, H, W, Cc), mean, var, g, b, 1e-3, C.ACT_RELU6, stream())
    yh, yl = torch.empty_like(x), torch.empty_like(x)
    C.call("myolo_bn_apply_split", xv, C.view(yh, n, H, W, Cc), C.view(yl, n, H, W, Cc), mean, var, g, b, 1e-3, C.ACT_RELU6, stream())
    assert ((yh + yl) - y).abs().max().item() <= 2 ** -20 * max(y.abs().max().item(), 1.0)
    # fold: scale/shift reproduce the BN affine map
    sc, sh = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    C.call("myolo_bn_fold", g, b, mean, var, 1e-3, sc, sh, Cc, stream())
    close(x * sc + sh, (x - mean) * torch.rsqrt(var + 1e-3) * g + b, 2e-6, "bn fold")
    # view copy: dense -> padded-flat, then accumulate back
    pf = PF(n, H, W, Cc)
    C.call("myolo_view_copy", xv, pf.view(), 0, stream())
    assert torch.equal(pf.dense(), x) and pf.rows.view(n, H + 1, W + 1, Cc)[:, 0].abs().max().item() == 0
    acc = x.clone()
    C.call("myolo_view_copy", pf.view(), C.view(acc, n, H, W, Cc), 1, stream())
    assert torch.equal(acc, 2 * x)
    # batched moving update == per-layer update
    val = [torch.rand(Cc, device="cuda") + 0.1 for _ in range(2)]
    bi = [torch.zeros(Cc, device="cuda") for _ in range(4)]
    mo = [torch.zeros(Cc, device="cuda") for _ in range(4)]
    npix = 77.0
    corr = (npix / (npix - 1)) * (npix / (npix - (1 + 1e-3)))
    rec = struct.pack("<QQQif", val[0].data_ptr(), bi[0].data_ptr(), mo[0].data_ptr(), Cc, 1.0) + \
        struct.pack("<QQQif", val[1].data_ptr(), bi[1].data_ptr(), mo[1].data_ptr(), Cc, corr)
    table = torch.frombuffer(bytearray(rec), dtype=torch.uint8).cuda()
    for step in (1, 2):
        C.call("myolo_bn_moving_update_batch", table, 2, 0.99, step, stream())
        C.call("myolo_bn_moving_update", val[0], bi[2], mo[2], Cc, 0.99, step, 0, npix, 1e-3, stream())
        C.call("myolo_bn_moving_update", val[1], bi[3], mo[3], Cc, 0.99, step, 1, npix, 1e-3, stream())
        close(mo[0], mo[2], 1e-6, "batched moving mean")
        close(mo[1], mo[3], 1e-6, "batched moving var")


def test_prep_weights_batch_equals_single_jobs(C):
    """one-launch staging (device job table) against the per-job entry points, all four modes"""
    import struct
    torch.manual_seed(50)
    specs = [(9, 64, 96, 1, 1), (1, 100, 40, 0, 0), (3, 32, 256, 1, 2), (9, 256, 256, 0, 3), (1, 1024, 256, 1, 3), (2, 33, 65, 1, 3)]
    rec, tiles, outs, refs = b"", 0, [], []
    keep = []
    for ntaps, rows, cols, tr, mode in specs:
        w = torch.randn(ntaps, rows, cols, device="cuda")
        n = ntaps * rows * cols
        if mode == 3:
            out, ref = torch.zeros(n, dtype=torch.float16, device="cuda"), torch.zeros(n, dtype=torch.float16, device="cuda")
            C.call("myolo_prep_weights_h", w, ref, ntaps, rows, cols, tr, stream())
        else:
            out, ref = torch.zeros(n * (3 if mode == 2 else 1), device="cuda"), torch.zeros(n * (3 if mode == 2 else 1), device="cuda")
            C.call("myolo_prep_weights", w, ref, ntaps, rows, cols, tr, mode, stream())
        rec += struct.pack("<QQiiiiiiii", w.data_ptr(), out.data_ptr(), ntaps, rows, cols, tr, mode, tiles, 0, 0)
        tiles += ntaps * ((rows + 31) // 32) * ((cols + 31) // 32)
        outs.append(out); refs.append(ref); keep.append(w)
    # zero-padded staging (conv_23: 27 output channels -> 32): transposed 3xTF32 triple [3][32][K] and plain [K][32]
    K, N, NP = 96, 27, 32
    w23 = torch.randn(K, N, device="cuda")
    pad_t = torch.zeros(3, NP, K, device="cuda")
    rec += struct.pack("<QQiiiiiiii", w23.data_ptr(), pad_t.data_ptr(), 1, K, N, 1, 2, tiles, 0, NP * K)
    tiles += ((K + 31) // 32) * ((N + 31) // 32)
    pad_n = torch.zeros(K, NP, device="cuda")
    rec += struct.pack("<QQiiiiiiii", w23.data_ptr(), pad_n.data_ptr(), 1, K, N, 0, 1, tiles, NP, 0)
    tiles += ((K + 31) // 32) * ((N + 31) // 32)
    table = torch.frombuffer(bytearray(rec), dtype=torch.uint8).cuda()
    C.call("myolo_prep_weights_batch", table, len(specs) + 2, tiles, stream())
    torch.cuda.synchronize()
    for o, r, sp in zip(outs, refs, specs):
        assert torch.equal(o, r), sp
    ref3 = torch.zeros(3, N, K, device="cuda")
    C.call("myolo_prep_weights", w23, ref3, 1, K, N, 1, 2, stream())
    assert torch.equal(pad_t[:, :N], ref3) and pad_t[:, N:].abs().max().item() == 0
    ref1 = torch.zeros(K, N, device="cuda")
    C.call("myolo_prep_weights", w23, ref1, 1, K, N, 0, 1, stream())
    assert torch.equal(pad_n[:, :N], ref1) and pad_n[:, N:].abs().max().item() == 0
    # myolo_copy_cols: cut the dense columns out of a padded matrix and back
    src = torch.randn(50, NP, device="cuda")
    dense = torch.empty(50, N, device="cuda")
    C.call("myolo_copy_cols", src, NP, dense, N, 50, N, stream())
    assert torch.equal(dense, src[:, :N])
    back = torch.zeros(50, NP, device="cuda")
    C.call("myolo_copy_cols", dense, N, back, NP, 50, N, stream())
    assert torch.equal(back[:, :N], dense) and back[:, N:].abs().max().item() == 0


def test_detect_postprocess_nmb_matches_reference_golden(C):
    """The device NMB against the outputs of the REFERENCE's own NMB (tests/golden/reference_utils_fixture.npz):
    same-class candidates only, and a candidate that was dropped still suppresses later ones."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_utils_fixture.npz"))
    S, K = 224, 10
    for case in range(4):
        bx, cls, ind = gold[f"nmb{case}_boxes"], gold[f"nmb{case}_class_ids"], gold[f"nmb{case}_indices"]
        n = len(ind)
        det = np.zeros((1, n, 6), np.float32)
        det[0, :, :4] = bx
        det[0, :, 4] = 1.0 - 0.05 * np.arange(n)            # candidate order = descending score
        det[0, :, 5] = cls
        dd = cuda(torch.tensor(det))
        i32 = lambda *sh: torch.empty(sh, dtype=torch.int32, device="cuda")
        idx, boxes, cl, cnt = i32(1, K), i32(1, K, 4), i32(1, K), i32(1)
        score = torch.empty(1, K, device="cuda")
        C.call("myolo_detect_postprocess", dd, None, 1, n, 4, S, 28, 28, K, 0.0, float(0.3 + 0.2 * case), idx, boxes, cl, score,
               cnt, None, stream())
        kept_pos = [list(ind).index(v) for v in gold[f"nmb{case}_kept"]]
        assert idx[0, :int(cnt[0].item())].cpu().tolist() == kept_pos, case


def test_target_encoding_on_device_matches_reference_golden(C):
    """myolo_extract_bboxes / myolo_encode_yolo_targets against what the REFERENCE's BatchGenerator.__getitem__ and
    extract_bboxes produced for the same ground truth (tests/golden/reference_utils_fixture.npz): bit-exact."""
    import os
    from myolo import myolo_utils as mu
    from tests.test_reference_golden import _Cfg
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_utils_fixture.npz"))
    cfg = _Cfg()
    for b in range(int(gold["bg_n_batches"])):
        ids, boxes, masks = gold[f"bg_batch{b}_gt_class_ids"], gold[f"bg_batch{b}_gt_boxes"], gold[f"bg_batch{b}_gt_masks"]
        tb_d, yt_d, _ = mu.encode_targets_device(cfg, torch.tensor(ids).cuda(), torch.tensor(boxes).cuda())
        assert np.array_equal(yt_d.cpu().numpy(), gold[f"bg_batch{b}_yolo_target"].astype(np.float32)), "yolo_target"
        assert np.array_equal(tb_d.cpu().numpy(), gold[f"bg_batch{b}_true_boxes"].astype(np.float32)), "true_boxes"
        gm = torch.tensor(masks).to(torch.uint8).cuda().contiguous()
        B, S, M = gm.shape[0], gm.shape[1], gm.shape[3]
        bx = torch.empty(B, M, 4, dtype=torch.int32, device="cuda")
        C.call("myolo_extract_bboxes", gm, B, S, M, bx, stream())
        assert np.array_equal(bx.cpu().numpy(), boxes[:, :M]), "extract_bboxes"
