"""Per-kernel parity of the half-operand (tcgen05 kind::f16) mask-head kernels.

Inputs are drawn so that every operand is exactly representable in IEEE half; the exact CUDA-core kernels
(already checked against the oracle in test_kernels_gpu.py) then compute the same products in fp32, so the
fp32 outputs of the half-operand kernels must agree to accumulation-order rounding and the half outputs to
one half ulp (2^-11 relative) of it.  ROIAlign's and BN's half outputs are compared bit-exactly with the
round-to-nearest-even conversion of the oracle-checked fp32 result."""
import ctypes

import pytest
import torch

from oracle import myolo_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from myolo import _cabi
    _cabi.device_check(0)
    return _cabi


def stream():
    return torch.cuda.current_stream().cuda_stream


def hq(t):
    """values exactly representable in half, kept in fp32"""
    return t.half().float()


def close(a, b, tol, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item() / scale
    assert err <= tol, f"{what}: rel-to-max err {err:.3e} > {tol}"


def half_pf(pf32):
    """half twin of a fp32 padded-flat tensor (same geometry, converted storage)"""
    from myolo.pf import PF
    h = PF(pf32.n, pf32.H, pf32.W, pf32.C, dtype=torch.float16)
    h.rows.copy_(pf32.rows)
    return h


def test_prep_weights_h(C):
    torch.manual_seed(1)
    w = torch.randn(9, 64, 96, device="cuda")
    out = torch.empty(9, 96, 64, dtype=torch.float16, device="cuda")
    C.call("myolo_prep_weights_h", w, out, 9, 64, 96, 1, stream())
    assert torch.equal(out, w.transpose(1, 2).contiguous().half())
    out2 = torch.empty(9, 64, 96, dtype=torch.float16, device="cuda")
    C.call("myolo_prep_weights_h", w * 1e5, out2, 9, 64, 96, 0, stream())
    assert torch.equal(out2, (w * 1e5).clamp(-65504, 65504).half()), "conversion saturates instead of producing inf"


@pytest.mark.parametrize("n,H,W", [(40, 14, 14), (300, 14, 14), (19, 7, 9)])
def test_conv3x3_half_operands(C, n, H, W):
    """forward conv with the folded bias/BN/ReLU epilogue, fp32 + half outputs; then the dgrad form with a
    device-side accumulator scale."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(30)
    Ci = Co = 256
    px = PF(n, H, W, Ci)
    px.valid().copy_(hq(torch.randn(n, H, W, Ci, device="cuda")))
    w = hq(torch.randn(9, Ci, Co, device="cuda") / (9 * Ci) ** 0.5)
    wt = torch.empty(9 * Ci * Co, device="cuda")
    C.call("myolo_prep_weights", w, wt, 9, Ci, Co, 1, 0, stream())
    wth = torch.empty(9 * Ci * Co, dtype=torch.float16, device="cuda")
    C.call("myolo_prep_weights_h", w, wth, 9, Ci, Co, 1, stream())
    bias, scale, shift = (torch.randn(Co, device="cuda") * 0.1, torch.rand(Co, device="cuda") + 0.5,
                          torch.randn(Co, device="cuda") * 0.1)
    sh = C.int_array(conv3x3_shifts(W))
    pfw, pfb, M = W + 1, (H + 1) * (W + 1), px.M
    ref = PF(n, H, W, Co)
    C.call("myolo_gemm_taps_ffma", px.rows, Ci, wt, ref.rows, Co, M, Co, Ci, 9, sh, bias, scale, shift, C.ACT_RELU, pfw, pfb, 0, stream())
    assert C.lib().myolo_gemm_taps_h_supported(Ci, M, Co, Ci, 9, ctypes.addressof(sh)) == 1
    xh = half_pf(px)
    out, outh = PF(n, H, W, Co), PF(n, H, W, Co, dtype=torch.float16)
    # one output per launch: fp32 (unrounded accumulator + epilogue) ...
    C.call("myolo_gemm_taps_h", xh.rows, Ci, wth, out.rows, Co, None, 0, M, Co, Ci, 9, sh, bias, scale, shift, C.ACT_RELU,
           pfw, pfb, None, stream())
    # ... or half
    C.call("myolo_gemm_taps_h", xh.rows, Ci, wth, None, 0, outh.rows, Co, M, Co, Ci, 9, sh, bias, scale, shift, C.ACT_RELU,
           pfw, pfb, None, stream())
    torch.cuda.synchronize()
    close(out.rows, ref.rows, 2e-5, "half-operand conv, fp32 output")
    assert torch.equal(outh.rows, out.rows.half()), "the half output is the round-to-nearest-even of the fp32 one"
    assert outh.storage[:Co].abs().max().item() == 0 and outh.rows.view(n, H + 1, W + 1, Co)[:, 0].abs().max().item() == 0
    assert outh.rows.view(n, H + 1, W + 1, Co)[:, :, 0].abs().max().item() == 0, "pad pixels stay zero"
    with pytest.raises(C.MyoloError):
        C.call("myolo_gemm_taps_h", xh.rows, Ci, wth, out.rows, Co, outh.rows, Co, M, Co, Ci, 9, sh, bias, scale, shift, C.ACT_RELU,
               pfw, pfb, None, stream())
    # fp32 output + the batch statistics of the result over the valid pixels in the epilogue (myolo_mask_conv1 -> bn1)
    ws = torch.zeros(8192, dtype=torch.float64, device="cuda")
    o2, ref2 = PF(n, H, W, Co), PF(n, H, W, Co)
    C.call("myolo_gemm_taps_h", xh.rows, Ci, wth, ref2.rows, Co, None, 0, M, Co, Ci, 9, sh, bias, None, None, C.ACT_NONE, pfw, pfb,
           None, stream())
    mean, var = torch.empty(Co, device="cuda"), torch.empty(Co, device="cuda")
    for rep in range(2):                      # twice: the workspace comes back zeroed
        pivot = None if rep == 0 else bias * 0.5 + 0.3
        C.call("myolo_gemm_taps_h_stats", xh.rows, Ci, wth, o2.rows, Co, M, Co, Ci, 9, sh, bias, pfw, pfb, pivot, mean, var, ws,
               n * H * W, stream())
        assert torch.equal(o2.rows, ref2.rows)
        flat = ref2.valid().reshape(-1, Co).double()
        close(mean, flat.mean(0).float(), 2e-5, "epilogue mean")
        close(var, flat.var(0, unbiased=False).float(), 2e-5, "epilogue variance")
        assert ws.abs().max().item() == 0
    # dgrad form (negated shifts, un-transposed weights), fp32 output scaled by a device scalar
    shn = C.int_array(conv3x3_shifts(W, negate=True))
    wh = w.half()
    r2 = PF(n, H, W, Ci)
    C.call("myolo_gemm_taps_ffma", ref.rows, Co, w, r2.rows, Ci, M, Ci, Co, 9, shn, None, None, None, 0, pfw, pfb, 0, stream())
    gh = half_pf(ref)
    r2h, g32 = PF(n, H, W, Ci), PF(n, H, W, Co)
    g32.rows.copy_(gh.rows)
    C.call("myolo_gemm_taps_ffma", g32.rows, Co, w, r2h.rows, Ci, M, Ci, Co, 9, shn, None, None, None, 0, pfw, pfb, 0, stream())
    sc = torch.tensor([0.25], device="cuda")
    o2 = PF(n, H, W, Ci)
    C.call("myolo_gemm_taps_h", gh.rows, Co, wh, o2.rows, Ci, None, 0, M, Ci, Co, 9, shn, None, None, None, 0, pfw, pfb, sc, stream())
    close(o2.rows * 4.0, r2h.rows, 2e-5, "half-operand dgrad with accumulator scale")


@pytest.mark.parametrize("n,NC", [(37, 4), (150, 2), (37, 16), (37, 17), (150, 81), (20, 128)])
def test_deconv_mask_tail_half_operands(C, n, NC):
    """Half-operand deconv GEMM with the mask tail as a SECOND tcgen05 GEMM in its epilogue (h = relu(acc + bias) as half
    x the 1x1 kernel as half, logits in TMEM) against the exact two-kernel path.  The 1x1 kernel is half-representable
    here, so the only extra rounding is h -> half (2^-11 relative per element).  n = 150 gives every CTA pair several
    work items (the tail GEMM of item i is issued inside the main loop of item i+1)."""
    from myolo.pf import PF
    torch.manual_seed(31)
    H, W, Cm = 14, 14, 256
    pa = PF(n, H, W, Cm)
    pa.valid().copy_(hq(torch.randn(n, H, W, Cm, device="cuda")))
    kd = hq(torch.randn(4 * Cm, Cm, device="cuda") / Cm ** 0.5)
    bd, w1, b1 = torch.randn(Cm, device="cuda") * 0.1, hq(torch.randn(Cm, NC, device="cuda") / Cm ** 0.5), torch.randn(NC, device="cuda") * 0.1
    ids = torch.zeros(n, dtype=torch.int32, device="cuda")
    ids[[0, 5, n - 1]] = torch.tensor([1, min(3, NC - 1), 1], dtype=torch.int32, device="cuda")
    y_ref = PF(n, H, W, 4 * Cm)
    C.call("myolo_gemm_taps_ffma", pa.rows, Cm, kd, y_ref.rows, 4 * Cm, pa.M, 4 * Cm, Cm, 1, None, None, None, None, 0,
           W + 1, (H + 1) * (W + 1), 0, stream())
    m_ref = torch.empty(n, 2 * H, 2 * W, NC, device="cuda")
    C.call("myolo_mask_out_fwd", y_ref.rows, bd, w1, b1, m_ref, n, H, W, Cm, NC, stream())
    y4 = PF(n, H, W, 4 * Cm)
    m = torch.full((n, 2 * H, 2 * W, NC), -1.0, device="cuda")
    assert C.lib().myolo_deconv_mask_fwd_supported(Cm, NC) == 1
    C.call("myolo_deconv_mask_fwd_h", half_pf(pa).rows, kd.half(), bd, w1, b1, m, ids, y4.rows, n, H, W, Cm, NC, stream())
    torch.cuda.synchronize()
    assert m.min().item() >= 0.0, "every mask element was written"
    close(m, m_ref, 3e-4, "masks of the half-operand deconv + tensor-core tail")
    yv, rv = y4.valid(), y_ref.valid()
    for r in range(n):
        if ids[r] > 0:
            close(yv[r], rv[r], 2e-5, "y4 of a positive roi")
        else:
            assert yv[r].abs().max().item() == 0


def test_dgrad_with_fused_bn_backward_half(C):
    """myolo_gemm_taps_bnbwd_h against the exact two-step path on the same (half-representable) tensors; the
    incoming gradient carries a loss scale of 64 that the reduced dgamma / dbeta / dbias must not."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(32)
    n, H, W, Cc = 40, 14, 14, 256
    S = 64.0
    g_true = PF(n, H, W, Cc)
    g_true.valid().copy_(hq(torch.randn(n, H, W, Cc, device="cuda")))
    a_out = PF(n, H, W, Cc)
    a_out.valid().copy_(hq(torch.relu(torch.randn(n, H, W, Cc, device="cuda"))))
    w = hq(torch.randn(9, Cc, Cc, device="cuda") / (9 * Cc) ** 0.5)
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    var = torch.rand(Cc, device="cuda") + 0.5
    shn = C.int_array(conv3x3_shifts(W, negate=True))
    pfw, pfb, M = W + 1, (H + 1) * (W + 1), g_true.M
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    ref = PF(n, H, W, Cc)
    C.call("myolo_gemm_taps_ffma", g_true.rows, Cc, w, ref.rows, Cc, M, Cc, Cc, 9, shn, None, None, None, 0, pfw, pfb, 0, stream())
    dg_r, db_r, dbias_r = (torch.empty(Cc, device="cuda") for _ in range(3))
    C.call("myolo_bn_act_bwd_from_output", a_out.view(), ref.view(), ref.view(), gamma, beta, var, 1e-3, C.ACT_RELU, dg_r, db_r,
           dbias_r, ws, stream())
    g_scaled = PF(n, H, W, Cc, dtype=torch.float16)
    g_scaled.rows.copy_(g_true.rows * S)
    out32, outh = PF(n, H, W, Cc), PF(n, H, W, Cc, dtype=torch.float16)
    dg, db, dbias = (torch.empty(Cc, device="cuda") for _ in range(3))
    unscale = torch.tensor([1.0 / S], device="cuda")
    C.call("myolo_gemm_taps_bnbwd_h", g_scaled.rows, Cc, w.half(), out32.rows, None, Cc, M, Cc, Cc, 9, shn, pfw, pfb,
           half_pf(a_out).rows, gamma, beta, var, 1e-3, C.ACT_RELU, dg, db, dbias, ws, unscale, stream())
    C.call("myolo_gemm_taps_bnbwd_h", g_scaled.rows, Cc, w.half(), None, outh.rows, Cc, M, Cc, Cc, 9, shn, pfw, pfb,
           half_pf(a_out).rows, gamma, beta, var, 1e-3, C.ACT_RELU, dg, db, dbias, ws, unscale, stream())
    torch.cuda.synchronize()
    close(out32.rows / S, ref.rows, 2e-5, "fused d(pre-BN), fp32 output")
    assert torch.equal(outh.rows, out32.rows.half())
    close(db, db_r, 1e-4, "fused dbeta")
    close(dg, dg_r, 1e-4, "fused dgamma")
    close(dbias, dbias_r, 1e-4, "fused dbias")
    assert ws.abs().max().item() == 0, "BN workspace must be left zero"
    assert outh.storage[:Cc].abs().max().item() == 0 and outh.rows.view(n, H + 1, W + 1, Cc)[:, 0].abs().max().item() == 0


def test_roialign_and_bn_half_outputs(C):
    from myolo.pf import PF
    torch.manual_seed(33)
    B, Fh, Cc, R, P = 2, 12, 64, 9, 14
    feat = torch.randn(B, Fh, Fh, Cc)
    g = torch.Generator().manual_seed(5)
    c = torch.rand(B * R, 2, generator=g)
    wh = torch.rand(B * R, 2, generator=g) * 0.6
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    boxes[0] = 0.0
    boxes[1] = torch.tensor([-0.5, 0.1, 1.5, 0.8])
    idx = torch.arange(B).repeat_interleave(R)
    ref = O.crop_and_resize(feat, boxes, idx, P, P)
    fd, bd = feat.cuda(), boxes.cuda()
    o32, oh = PF(B * R, P, P, Cc), PF(B * R, P, P, Cc, dtype=torch.float16)
    C.call("myolo_roialign_fwd_h", C.view(fd, B, Fh, Fh, Cc), bd, B * R, R, P, o32.view(), oh.view(), stream())
    assert torch.equal(o32.valid().cpu(), ref), "fp32 ROIAlign output stays bit-exact vs the oracle"
    assert torch.equal(oh.valid().cpu(), ref.half()), "half output = round-to-nearest-even of it"
    oh2 = PF(B * R, P, P, Cc, dtype=torch.float16)
    C.call("myolo_roialign_fwd_h", C.view(fd, B, Fh, Fh, Cc), bd, B * R, R, P, None, oh2.view(), stream())
    assert torch.equal(oh2.rows, oh.rows)
    # BN + ReLU with half / half-rounded fp32 outputs vs myolo_bn_apply
    x = PF(5, P, P, Cc)
    x.valid().normal_()
    mean, var = torch.randn(Cc, device="cuda") * 0.1, torch.rand(Cc, device="cuda") + 0.5
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    y_ref, y32, yh = PF(5, P, P, Cc), PF(5, P, P, Cc), PF(5, P, P, Cc, dtype=torch.float16)
    C.call("myolo_bn_apply", x.view(), y_ref.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, stream())
    C.call("myolo_bn_apply_h", x.view(), y32.view(), yh.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, stream())
    assert torch.equal(yh.rows, y_ref.rows.half())
    assert torch.equal(y32.rows, y_ref.rows.half().float())


@pytest.mark.parametrize("n,H,W,K,N,taps", [(40, 14, 14, 256, 256, 9), (300, 14, 14, 256, 256, 9), (23, 7, 9, 64, 128, 9),
                                            (37, 14, 14, 256, 1024, 1), (11, 14, 14, 128, 64, 9)])
def test_wgrad_half_operands(C, n, H, W, K, N, taps):
    """filter gradient with both operands MN-major half (kind::f16) vs the exact CUDA-core kernel on the same
    half-representable tensors; the gradient operand carries a loss scale that out_scale removes."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(40)
    S = 32.0
    a = PF(n, H, W, K)
    a.valid().copy_(hq(torch.randn(n, H, W, K, device="cuda")))
    d = PF(n, H, W, N)
    d.valid().copy_(hq(torch.randn(n, H, W, N, device="cuda") * 0.1))
    sh = C.int_array(conv3x3_shifts(W)) if taps == 9 else None
    M = a.M
    for tr in ((0, 1) if taps == 1 else (0,)):
        ref = torch.zeros(taps * K * N, device="cuda")
        C.call("myolo_gemm_taps_wgrad_ffma", a.rows, K, d.rows, N, ref, M, N, K, taps, sh, tr, stream())
        ah = half_pf(a)
        dh = PF(n, H, W, N, dtype=torch.float16)
        dh.rows.copy_(d.rows * S)
        out = torch.zeros(taps * K * N, device="cuda")
        assert C.lib().myolo_gemm_taps_wgrad_h_supported(K, N, M, N, K, taps) == 1
        unscale = torch.tensor([1.0 / S], device="cuda")
        C.call("myolo_gemm_taps_wgrad_h", ah.rows, K, dh.rows, N, out, M, N, K, taps, sh, tr, unscale, stream())
        torch.cuda.synchronize()
        close(out, ref, 3e-5, f"half-operand wgrad (transpose_out={tr})")


def test_mask_out_bwd_half_and_grad_scale(C):
    from myolo.pf import PF
    torch.manual_seed(41)
    n, H, W, Cm, NC = 21, 14, 14, 256, 4
    y4 = PF(n, H, W, 4 * Cm)
    y4.valid().normal_()
    bd, w1 = torch.randn(Cm, device="cuda") * 0.1, torch.randn(Cm, NC, device="cuda") / Cm ** 0.5
    dlogit = torch.zeros(n, 2 * H, 2 * W, NC, device="cuda")
    for r, k in ((0, 1), (7, 3), (20, 2)):
        dlogit[r, :, :, k] = torch.randn(2 * H, 2 * W, device="cuda") * 3e-5
    gs = torch.tensor([1.0, 1.0, 0.0, 0.0], device="cuda")
    C.call("myolo_grad_scale", dlogit, dlogit.numel(), gs, stream())
    m = dlogit.abs().max().item()
    S = gs[0].item()
    assert 8.0 <= m * S < 16.0 and gs[1].item() == 1.0 / S and gs[2].item() == 0.0, (m, S)
    import math
    assert math.log2(S) == int(math.log2(S)), "the loss scale is a power of two"
    ref = PF(n, H, W, 4 * Cm)
    gr = [torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")]
    C.call("myolo_mask_out_bwd", y4.rows, bd, w1, dlogit, ref.rows, gr[0], gr[1], gr[2], n, H, W, Cm, NC, stream())
    outh = PF(n, H, W, 4 * Cm, dtype=torch.float16)
    gh = [torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")]
    C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, outh.rows, gh[0], gh[1], gh[2], n, H, W, Cm, NC, gs, None, None, stream())
    assert torch.equal(outh.rows, (ref.rows * S).half()), "dy4 = half(S * exact)"
    for a, b in zip(gh, gr):
        close(a, b, 1e-5, "unscaled parameter gradients of the mask tail")
    # with the target ids: non-positive rois are zero-filled without reading dlogit
    ids = torch.zeros(n, dtype=torch.int32, device="cuda")
    ids[[0, 7, 20]] = torch.tensor([1, 3, 2], dtype=torch.int32, device="cuda")
    outh2 = PF(n, H, W, 4 * Cm, dtype=torch.float16)
    outh2.valid().fill_(7.0)
    gh2 = [torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")]
    C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, outh2.rows, gh2[0], gh2[1], gh2[2], n, H, W, Cm, NC, gs, ids, None, stream())
    assert torch.equal(outh2.rows, outh.rows)
    for a, b in zip(gh2, gh):
        close(a, b, 1e-5, "same parameter gradients with the id fast path")
    # persistent buffer + prev_ids: only rows whose roi is or was positive are touched; a second step with another
    # positive set must leave exactly the new set non-zero
    buf = PF(n, H, W, 4 * Cm, dtype=torch.float16)
    prev = torch.zeros(n, dtype=torch.int32, device="cuda")
    g3 = [torch.zeros(Cm, NC, device="cuda"), torch.zeros(NC, device="cuda"), torch.zeros(Cm, device="cuda")]
    C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit, buf.rows, g3[0], g3[1], g3[2], n, H, W, Cm, NC, gs, ids, prev, stream())
    assert torch.equal(buf.rows, outh.rows) and torch.equal(prev, ids)
    dlogit2 = torch.zeros_like(dlogit)
    dlogit2[3, :, :, 1] = torch.randn(2 * H, 2 * W, device="cuda") * 3e-5
    dlogit2[7, :, :, 2] = torch.randn(2 * H, 2 * W, device="cuda") * 3e-5
    ids2 = torch.zeros(n, dtype=torch.int32, device="cuda")
    ids2[[3, 7]] = torch.tensor([1, 2], dtype=torch.int32, device="cuda")
    ref2 = PF(n, H, W, 4 * Cm, dtype=torch.float16)
    C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit2, ref2.rows, g3[0], g3[1], g3[2], n, H, W, Cm, NC, gs, None, None, stream())
    C.call("myolo_mask_out_bwd_h", y4.rows, bd, w1, dlogit2, buf.rows, g3[0], g3[1], g3[2], n, H, W, Cm, NC, gs, ids2, prev, stream())
    assert torch.equal(buf.rows, ref2.rows) and torch.equal(prev, ids2)
    # all-zero gradient -> S = 1
    z = torch.zeros(1024, device="cuda")
    C.call("myolo_grad_scale", z, 1024, gs, stream())
    assert gs[0].item() == 1.0 and gs[1].item() == 1.0


def test_bn_backward_half_output(C):
    from myolo.pf import PF
    torch.manual_seed(42)
    n, P, Cc = 9, 14, 64
    x, dy = PF(n, P, P, Cc), PF(n, P, P, Cc)
    x.valid().normal_()
    dy.valid().normal_()
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    mean, var = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    C.call("myolo_bn_stats", x.view(), mean, var, ws, stream())
    ref = PF(n, P, P, Cc)
    dg_r, db_r, dg, db = (torch.empty(Cc, device="cuda") for _ in range(4))
    C.call("myolo_bn_bwd", x.view(), dy.view(), ref.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, 1, dg_r, db_r, ws, stream())
    outh = PF(n, P, P, Cc, dtype=torch.float16)
    sc = torch.tensor([128.0], device="cuda")
    C.call("myolo_bn_bwd_h", x.view(), dy.view(), outh.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, 1, dg, db, ws, sc, stream())
    assert torch.equal(outh.rows, (ref.rows * 128.0).half())
    assert torch.equal(dg, dg_r) and torch.equal(db, db_r)


def test_bn_backward_half_in_half_out(C):
    """batch-statistics BN backward on a loss-scaled half gradient, in place: dx keeps the scale, dgamma / dbeta do not"""
    from myolo.pf import PF
    torch.manual_seed(43)
    n, P, Cc, S = 9, 14, 64, 256.0
    x, dy = PF(n, P, P, Cc), PF(n, P, P, Cc)
    x.valid().normal_()
    dy.valid().copy_(hq(torch.randn(n, P, P, Cc, device="cuda") * 0.01))
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    mean, var = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    C.call("myolo_bn_stats", x.view(), mean, var, ws, stream())
    ref = PF(n, P, P, Cc)
    dg_r, db_r, dg, db = (torch.empty(Cc, device="cuda") for _ in range(4))
    C.call("myolo_bn_bwd", x.view(), dy.view(), ref.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, 1, dg_r, db_r, ws, stream())
    gh = PF(n, P, P, Cc, dtype=torch.float16)
    gh.rows.copy_(dy.rows * S)
    us = torch.tensor([1.0 / S], device="cuda")
    C.call("myolo_bn_bwd_hh", x.view(), gh.view(), gh.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, 1, dg, db, ws, us, stream())
    close(gh.rows.float() / S, ref.rows, 6e-4, "dx (half, scaled)")
    close(dg, dg_r, 1e-5, "dgamma")
    close(db, db_r, 1e-5, "dbeta")
    assert ws[:2064].abs().max().item() == 0, "tickets and sums are left zero (the coefficient table behind them is scratch)"


def test_bn1_half_path_forward(C):
    """myolo_mask_conv1 -> myolo_mask_bn1 on a HALF pre-BN tensor: the conv stores half and takes the batch statistics of the
    half-rounded values in its epilogue (myolo_gemm_taps_hh_stats), BN + ReLU reads half and writes half (myolo_bn_apply_hh)."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(51)
    n, H, W, Ci, Co = 300, 14, 14, 256, 256
    x = PF(n, H, W, Ci)
    x.valid().copy_(hq(torch.randn(n, H, W, Ci, device="cuda")))
    w = hq(torch.randn(9, Ci, Co, device="cuda") / (9 * Ci) ** 0.5)
    wth = w.transpose(1, 2).contiguous().half()
    bias = torch.randn(Co, device="cuda") * 0.5
    sh = C.int_array(conv3x3_shifts(W))
    pfw, pfb, M = W + 1, (H + 1) * (W + 1), x.M
    xh = half_pf(x)
    ref32 = PF(n, H, W, Co)
    C.call("myolo_gemm_taps_h", xh.rows, Ci, wth, ref32.rows, Co, None, 0, M, Co, Ci, 9, sh, bias, None, None, C.ACT_NONE, pfw, pfb,
           None, stream())
    ws = torch.zeros(8192, dtype=torch.float64, device="cuda")
    mean, var = torch.empty(Co, device="cuda"), torch.empty(Co, device="cuda")
    zh = PF(n, H, W, Co, dtype=torch.float16)
    for rep in range(2):
        pivot = None if rep == 0 else bias + 0.1
        C.call("myolo_gemm_taps_hh_stats", xh.rows, Ci, wth, zh.rows, Co, M, Co, Ci, 9, sh, bias, pfw, pfb, pivot, mean, var, ws,
               n * H * W, stream())
        assert torch.equal(zh.rows, ref32.rows.half()), "half result = round-to-nearest-even of the fp32 one"
        flat = zh.valid().reshape(-1, Co).double()
        close(mean, flat.mean(0).float(), 2e-5, "mean of the half-rounded result")
        close(var, flat.var(0, unbiased=False).float(), 2e-5, "variance of the half-rounded result")
        assert ws.abs().max().item() == 0
    assert zh.storage[:Co].abs().max().item() == 0 and zh.rows.view(n, H + 1, W + 1, Co)[:, 0].abs().max().item() == 0
    gamma, beta = torch.rand(Co, device="cuda") + 0.5, torch.randn(Co, device="cuda") * 0.1
    z32 = PF(n, H, W, Co)
    z32.rows.copy_(zh.rows)
    y_ref, yh = PF(n, H, W, Co), PF(n, H, W, Co, dtype=torch.float16)
    C.call("myolo_bn_apply", z32.view(), y_ref.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, stream())
    C.call("myolo_bn_apply_hh", zh.view(), yh.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, stream())
    assert torch.equal(yh.rows, y_ref.rows.half())


@pytest.mark.parametrize("n", [40, 300])
def test_bn1_half_path_backward(C, n):
    """Backward of a batch-statistics BN + ReLU split over the epilogue of the GEMM that produces d(a)
    (myolo_gemm_taps_bnbwd_sums_h: relu'(a), gamma * rs, the two column sums) and one in-place correction pass
    (myolo_bn_bwd_batch_fix_hh), against the exact path: CUDA-core data-gradient GEMM, then myolo_bn_bwd on fp32."""
    from myolo.pf import PF, conv3x3_shifts
    torch.manual_seed(52)
    H, W, Cc, S = 14, 14, 256, 64.0
    z = PF(n, H, W, Cc)
    z.valid().copy_(hq(torch.randn(n, H, W, Cc, device="cuda") * 1.3 + 0.2))
    gamma, beta = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda") * 0.1
    mean, var = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    ws = torch.zeros(4112, dtype=torch.float64, device="cuda")
    C.call("myolo_bn_stats", z.view(), mean, var, ws, stream())
    zh = half_pf(z)
    ah = PF(n, H, W, Cc, dtype=torch.float16)
    C.call("myolo_bn_apply_hh", zh.view(), ah.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, stream())
    g_true = PF(n, H, W, Cc)
    g_true.valid().copy_(hq(torch.randn(n, H, W, Cc, device="cuda")))
    w = hq(torch.randn(9, Cc, Cc, device="cuda") / (9 * Cc) ** 0.5)
    shn = C.int_array(conv3x3_shifts(W, negate=True))
    pfw, pfb, M = W + 1, (H + 1) * (W + 1), z.M
    # exact: d(a) on CUDA cores in fp32, then the three-pass BN backward on the fp32 pre-BN tensor
    da = PF(n, H, W, Cc)
    C.call("myolo_gemm_taps_ffma", g_true.rows, Cc, w, da.rows, Cc, M, Cc, Cc, 9, shn, None, None, None, 0, pfw, pfb, 0, stream())
    ref = PF(n, H, W, Cc)
    dg_r, db_r, dg, db = (torch.empty(Cc, device="cuda") for _ in range(4))
    C.call("myolo_bn_bwd", z.view(), da.view(), ref.view(), mean, var, gamma, beta, 1e-3, C.ACT_RELU, 1, dg_r, db_r, ws, stream())
    # fused: loss-scaled half gradient in, half d(pre-BN) out
    g_scaled = PF(n, H, W, Cc, dtype=torch.float16)
    g_scaled.rows.copy_(g_true.rows * S)
    g1 = PF(n, H, W, Cc, dtype=torch.float16)
    us = torch.tensor([1.0 / S], device="cuda")
    for rep in range(2):
        C.call("myolo_gemm_taps_bnbwd_sums_h", g_scaled.rows, Cc, w.half(), g1.rows, Cc, M, Cc, Cc, 9, shn, pfw, pfb, ah.rows,
               gamma, beta, var, 1e-3, C.ACT_RELU, ws, stream())
        assert ws[16:16 + 2 * Cc].abs().max().item() > 0, "the column sums wait in the workspace"
        C.call("myolo_bn_bwd_batch_fix_hh", zh.view(), g1.view(), mean, var, gamma, 1e-3, dg, db, ws, us, stream())
        torch.cuda.synchronize()
        # two half roundings of values up to ~S * max|d(a)| * gamma * rs
        close(g1.rows.float() / S, ref.rows, 1.5e-3, "d(pre-BN), half, scaled")
        close(dg, dg_r, 2e-3, "dgamma (xhat recovered from the half activation)")
        close(db, db_r, 1e-4, "dbeta")
        assert ws[:2064].abs().max().item() == 0, "tickets and sums are left zero"
        v = g1.rows.view(n, H + 1, W + 1, Cc)
        assert v[:, 0].abs().max().item() == 0 and v[:, :, 0].abs().max().item() == 0, "pad pixels stay zero"
    # rows that were zero before the correction pass (the sparse backward's other rois) receive the mean terms only
    g1.rows.zero_()
    ws[16:16 + Cc] = 3.0
    ws[16 + Cc:16 + 2 * Cc] = -2.0
    C.call("myolo_bn_bwd_batch_fix_hh", zh.view(), g1.view(), mean, var, gamma, 1e-3, dg, db, ws, None, stream())
    rs = 1.0 / torch.sqrt(var + 1e-3)
    npx = n * H * W
    want = -(gamma * rs) * (3.0 / npx + (zh.valid().float() - mean) * rs * (-2.0 / npx))
    close(g1.valid().float(), want, 2e-3, "mean terms on zero rows")
    assert torch.equal(dg, torch.full_like(dg, -2.0)) and torch.equal(db, torch.full_like(db, 3.0))
