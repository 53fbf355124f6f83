"""CPU tests of the oracle (oracle/myolo_oracle.py): structure pinned against the GraphDef the
reference ships (tests/golden/graph_fixture.json), hand-computed micro-cases for the third-party
semantics it restates (tf.image.crop_and_resize, Keras BCE, YOLO loss, Keras Adam / BN update) and
fp64-vs-fp32 self-consistency.  The reference has no golden vectors of its own; the vectors
made by executing its source are checked in test_reference_graph_golden.py / test_reference_golden.py."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import myolo_oracle as O
from tests import helpers as Hh

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "graph_fixture.json")))


def test_variable_names_and_shapes_match_graph_fixture():
    from myolo.engine import param_specs
    specs = {n: list(s) for n, s, _ in param_specs(3, 4)}          # fixture: N_BOX 3, NUM_CLASSES 4
    fx = {k: v for k, v in FIX["variables"].items()
          if not k.startswith(("training", "Adam", "yolo_sum_loss")) and "/biased" not in k and "/local_step" not in k}
    assert set(specs) == set(fx), (set(specs) ^ set(fx))
    for k, shp in fx.items():
        assert specs[k] == shp, (k, specs[k], shp)
    n_all = sum(int(np.prod(s)) for s in specs.values())
    n_tr = sum(int(np.prod(s)) for n, s, t in param_specs(3, 4) if t)
    # SURVEY 10.3 totals (7 321 763 / 7 296 035) also count the loss-state scalars `seen` / `total_recall`
    # (yolo_sum_loss/Variable*, tf.Variables that are trainable by default; the Lambda is traced twice)
    n_state = sum(1 for k in FIX["variables"] if k.startswith("yolo_sum_loss/Variable"))
    assert (n_all + n_state, n_tr + n_state) == (7321763, 7296035)


def test_op_attributes_match_graph_fixture():
    oa, co = FIX["op_attrs"], FIX["consts"]
    assert oa["conv1/convolution"]["strides"] == [1, 2, 2, 1] and oa["conv1/convolution"]["padding"] == "VALID"
    strides = {k: s for k, _, _, s in O.BACKBONE_BLOCKS + O.YOLO_BLOCKS}
    for k, s in strides.items():
        key = f"conv_dw_{k}/depthwise" if k <= 6 else f"yolo_model/conv_dw_{k}/depthwise"
        a = oa.get(key) or oa[f"conv_dw_{k}/depthwise"]
        assert a["padding"] == "VALID" and a["strides"] == [1, s, s, 1], (k, a)
        pad = co.get(f"conv_pad_{k}/Pad/paddings") or co[f"yolo_model/conv_pad_{k}/Pad/paddings"]
        assert pad == [[0, 0], [1, 1], [1, 1], [0, 0]]
    assert abs(oa["conv1_bn/FusedBatchNorm"]["epsilon"] - O.BN_EPS) < 1e-9
    assert co["roi_align_mask/CropAndResize/crop_size"] == [14, 14]
    assert co["detect_mask_targets/CropAndResize/crop_size"] == [28, 28]
    assert co["detect_mask_targets/GreaterEqual/y"] == 0.5 and co["detect_mask_targets/Less/y"] == 0.5
    assert abs(co["yolo_sum_loss/Less/y"] - 0.6) < 1e-6
    assert abs(co["Adam/lr/initial_value"] - 1e-3) < 1e-9 and abs(co["Adam/beta_1/initial_value"] - 0.9) < 1e-6
    assert FIX["placeholders"]["input_yolo_target"] == [-1, 7, 7, 3, 9]
    assert FIX["placeholders"]["output_rois/Placeholder"] == [-1, 147, 4]


def test_crop_and_resize_hand_cases():
    # 3x3 ramp image, value = 10*y + x
    img = torch.tensor([[0., 1, 2], [10, 11, 12], [20, 21, 22]]).reshape(1, 3, 3, 1)
    idx = torch.zeros(1, dtype=torch.long)
    full = O.crop_and_resize(img, torch.tensor([[0., 0., 1., 1.]]), idx, 3, 3)[0, :, :, 0]
    assert torch.equal(full, img[0, :, :, 0])                                    # identity box
    half = O.crop_and_resize(img, torch.tensor([[0., 0., 1., 1.]]), idx, 2, 2)[0, :, :, 0]
    assert torch.equal(half, torch.tensor([[0., 2.], [20., 22.]]))               # corners only
    mid = O.crop_and_resize(img, torch.tensor([[0.25, 0.25, 0.75, 0.75]]), idx, 2, 2)[0, :, :, 0]
    assert torch.allclose(mid, torch.tensor([[5.5, 6.5], [15.5, 16.5]]))         # bilinear at (0.5,0.5)...(1.5,1.5)
    out = O.crop_and_resize(img, torch.tensor([[-0.5, 0., 1.5, 1.]]), idx, 3, 3)[0, :, :, 0]
    assert torch.equal(out[0], torch.zeros(3)) and torch.equal(out[2], torch.zeros(3))   # rows outside -> 0
    assert torch.equal(out[1], torch.tensor([10., 11., 12.]))
    zero = O.crop_and_resize(img + 7, torch.tensor([[0., 0., 0., 0.]]), idx, 14, 14)
    assert torch.equal(zero, torch.full((1, 14, 14, 1), 7.0))                     # padded roi samples pixel (0,0)
    one = O.crop_and_resize(img, torch.tensor([[0., 0., 1., 1.]]), idx, 1, 1)
    assert one.item() == 11.0                                                      # crop 1 -> box centre
    nan = O.crop_and_resize(img, torch.tensor([[float("nan"), 0., 1., 1.]]), idx, 2, 2)
    assert torch.equal(nan, torch.zeros(1, 2, 2, 1))


def test_crop_and_resize_gradient_is_bilinear_scatter():
    img = torch.zeros(1, 4, 4, 1, requires_grad=True)
    out = O.crop_and_resize(img, torch.tensor([[0.1, 0.2, 0.7, 0.9]]), torch.zeros(1, dtype=torch.long), 3, 3)
    out.sum().backward()
    assert abs(img.grad.sum().item() - 9.0) < 1e-5                                # weights of each sample sum to 1


def test_roi_align_swaps_axes_like_the_reference():
    """PyramidROIAlign passes (x1,y1,x2,y2) to an op reading (y1,x1,y2,x2): SURVEY Q2."""
    feat = torch.arange(16.).reshape(1, 4, 4, 1)                                  # value = 4*row + col
    rois = torch.tensor([[[0.0, 1.0, 0.0, 1.0]]])                                  # x fixed at 0, y spans 0..1
    out = O.pyramid_roi_align(rois, feat, 2)[0, 0, :, :, 0]
    # read as y1=0,x1=1,y2=0,x2=1 -> row 0 everywhere, column 3 everywhere
    assert torch.equal(out, torch.full((2, 2), 3.0))


def test_mask_loss_matches_plain_bce():
    torch.manual_seed(0)
    p = torch.rand(1, 3, 28, 28, 4) * 0.98 + 0.01
    t = (torch.rand(1, 3, 28, 28) > 0.5).float()
    ids = torch.tensor([[2, 0, 1]], dtype=torch.int32)
    got = O.myolo_mask_loss_graph(t, ids, p)
    sel = torch.stack([p[0, 0, :, :, 2], p[0, 2, :, :, 1]])
    tt = torch.stack([t[0, 0], t[0, 2]])
    ref = -(tt * sel.log() + (1 - tt) * (1 - sel).log()).mean()
    assert abs(got.item() - ref.item()) < 1e-5
    assert O.myolo_mask_loss_graph(t, torch.zeros(1, 3, dtype=torch.int32), p).item() == 0.0


def test_yolo_loss_single_object_by_hand():
    cfg = dict(GRID_H=2, GRID_W=2, N_BOX=1, NUM_CLASSES=2, ANCHORS=[1.0, 1.0], COORD_SCALE=1.0, NO_OBJECT_SCALE=1.0,
               OBJECT_SCALE=5.0, CLASS_SCALE=1.0, CLASS_WEIGHTS=np.ones(2, "float32"), WARM_UP_BATCHES=0)
    yt = torch.zeros(1, 2, 2, 1, 7)
    yt[0, 0, 1, 0] = torch.tensor([1.5, 0.5, 1.0, 1.0, 1.0, 0.0, 1.0])          # cell row 0, col 1
    tb = torch.zeros(1, 1, 1, 1, 3, 4)
    tb[0, 0, 0, 0, 0] = torch.tensor([1.5, 0.5, 1.0, 1.0])
    yp = torch.zeros(1, 2, 2, 1, 7)                                              # sigmoid(0)=.5, exp(0)=1: perfect box in that cell
    loss = O.yolo_custom_loss(yt, yp, tb, cfg)
    # object cell: iou 1 -> (1-0.5)^2*5 ; three empty cells: best iou < 0.6 -> (0-0.5)^2*1 each; nb_conf = 4
    conf = (0.25 * 5 + 3 * 0.25) / (4 + 1e-6) / 2
    cls = math.log(2.0) / (1 + 1e-6)
    assert abs(loss.item() - (conf + cls)) < 1e-5, (loss.item(), conf + cls)


def test_decode_matches_closed_form():
    cfg = dict(GRID_H=2, GRID_W=2, N_BOX=1, NUM_CLASSES=1, ANCHORS=[2.0, 1.0])
    yp = torch.zeros(1, 2, 2, 1, 6)
    b = O.decode_yolo(yp, cfg)                                                    # order (row, col, box)
    # cell (row 1, col 0): centre ((0.5+0)/2, (0.5+1)/2), w = 2/2, h = 1/2
    assert torch.allclose(b[0, 2], torch.tensor([0.25 - 0.5, 0.75 - 0.25, 0.25 + 0.5, 0.75 + 0.25]))
    d = O.detections_layer(yp, cfg)
    assert d.shape == (1, 4, 6) and torch.allclose(d[..., 4], torch.full((1, 4), 0.5))


def test_targets_partition_and_padding():
    cfg = dict(TRAIN_ROIS_PER_IMAGE=5, MASK_SHAPE=[28, 28])
    props = torch.tensor([[0.5, 0.5, 0.9, 0.9], [0.1, 0.1, 0.4, 0.4], [float("nan")] * 4, [0.1, 0.1, 0.45, 0.4], [0.6, 0.6, 0.7, 0.7]])
    gtb = O.norm_boxes_graph(torch.tensor([[6.0, 6.0, 27.0, 27.0], [0, 0, 0, 0]]), 64, 64)
    masks = torch.zeros(64, 64, 2, dtype=torch.bool)
    masks[6:27, 6:27, 0] = True
    rois, ids, tm = O.detect_mask_target_graph(props, torch.tensor([3, 0]), gtb, masks, cfg)
    assert ids.tolist() == [3, 3, 0, 0, 0]                                        # positives first, original order
    assert torch.equal(rois[0], props[1]) and torch.equal(rois[1], props[3]) and torch.equal(rois[2], props[0])
    assert torch.equal(rois[4], torch.zeros(4))                                   # NaN proposal dropped -> zero padding
    assert tm[0].min() == 1.0 and tm[2:].abs().sum() == 0
    assert gtb[1, :2].tolist() == [0.0, 0.0] and torch.allclose(gtb[1, 2:], torch.tensor([-1 / 63, -1 / 63]))


def test_adam_and_moving_average_first_step():
    c = Hh.engine_cfg(S=64)
    oc = Hh.oracle_cfg(c)
    from myolo.engine import init_params
    P = init_params(c["NB"], c["NC"], 0, "keras")
    img = torch.rand(2, 64, 64, 3, generator=torch.Generator().manual_seed(0))
    inputs = Hh.batch_from_boxes(c, 2, img, Hh.random_boxes(2, 2, 1), 2)
    P0 = {k: v.clone() for k, v in P.items()}
    opt = {}
    out, g = O.train_step(P, opt, Hh.to_oracle_inputs(inputs), oc, lr=1e-3)
    k = "conv1/kernel"
    moved = (P[k] - P0[k]).abs()
    nz = g[k].abs() > 1e-6
    assert torch.allclose(moved[nz], torch.full_like(moved[nz], 1e-3), atol=2e-5)   # Keras Adam step 1: lr * g/(|g|+eps)
    name, mean, varv = out["bn_record"].items[0]
    assert torch.allclose(P[name + "/moving_mean"], mean, atol=1e-6)               # zero-debias: step-1 average == value
    assert opt["t"] == 1 and opt["seen"] == 1.0


def test_fp64_fp32_self_consistency():
    c = Hh.engine_cfg(S=64)
    oc = Hh.oracle_cfg(c)
    from myolo.engine import init_params
    P = init_params(c["NB"], c["NC"], 3, "trained_like")
    img = torch.rand(2, 64, 64, 3, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a = O.forward_inference(P, img, oc)
        b = O.forward_inference({k: v.double() for k, v in P.items()}, img.double(), oc)
    assert (a["detections"][..., :5] - b["detections"][..., :5].float()).abs().max() < 2e-4
    assert (a["yolo_output"] - b["yolo_output"].float()).abs().max() < 5e-4


def test_crop_and_resize_equals_align_corners_bilinear_sampling_inside_the_image():
    """Independent cross-check of the restated TF primitive: for samples that fall inside the image,
    tf.image.crop_and_resize is plain bilinear sampling on the corner-aligned grid, which PyTorch implements separately as
    grid_sample(align_corners=True).  (Outside the image the two differ by design: TF writes the extrapolation value for
    the whole sample, grid_sample blends with zero padding -- those samples are covered by the hand-computed cases.)"""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    B, H, W, C, N, ch, cw = 2, 9, 13, 3, 12, 7, 5
    img = torch.randn(B, H, W, C, generator=g, dtype=torch.float64)
    y1x1 = torch.rand(N, 2, generator=g, dtype=torch.float64) * 0.5
    hw = torch.rand(N, 2, generator=g, dtype=torch.float64) * 0.5          # boxes stay inside [0, 1]
    boxes = torch.cat([y1x1, y1x1 + hw], 1)                                # (y1, x1, y2, x2)
    idx = torch.arange(N) % B
    got = O.crop_and_resize(img, boxes, idx, ch, cw)
    ys = boxes[:, 0:1] + (boxes[:, 2:3] - boxes[:, 0:1]) * torch.linspace(0, 1, ch, dtype=torch.float64)[None]     # [N, ch] in [0,1]
    xs = boxes[:, 1:2] + (boxes[:, 3:4] - boxes[:, 1:2]) * torch.linspace(0, 1, cw, dtype=torch.float64)[None]
    grid = torch.stack([(2 * xs - 1)[:, None, :].expand(N, ch, cw), (2 * ys - 1)[:, :, None].expand(N, ch, cw)], -1)
    ref = F.grid_sample(img[idx].permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert torch.allclose(got, ref.permute(0, 2, 3, 1), atol=1e-12)
