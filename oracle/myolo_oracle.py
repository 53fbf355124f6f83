"""CPU oracle for the Mask-YOLO hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``mask-yolo_b200/``) never imports it and has no CPU fallback.

PARITY STATUS.  The reference (jianing-sun/Mask-YOLO @ 402dbd9) ships no golden vectors and no
known-answer tests, and its third-party stack (TensorFlow 1.x / Keras 2.x / keras_applications /
mrcnn) cannot be installed here (Python 3.12).  This file is a *restatement* of the reference's
algorithm in torch-CPU tensor ops (fp32 by default, fp64 when the inputs are fp64), function by
function.  It is pinned as follows:
  * PINNED to the reference's own Python source, executed here: every function of myolo/model.py on
    the path -- yolo_custom_loss (both branches), DecodeYOLOLayer, DetectionsLayer, norm_boxes_graph,
    trim_zeros_graph, overlaps_graph, DetectMaskTargetLayer / detect_mask_target_graph,
    PyramidROIAlign, myolo_mask_loss_graph, conv_block, mobilenet_graph, yolo_branch_graph,
    build_mask_graph and the whole MaskYOLO(mode, config).build for 'training' and 'inference' -- runs
    UNMODIFIED from /root/reference over eager numpy/torch stand-ins for its TensorFlow / Keras
    primitives (tests/golden/tf1_numpy_shim.py, keras2_layers_shim.py); the outputs are committed as
    tests/golden/reference_graph_fixture.npz (tests/golden/make_reference_graph_fixtures.py) and
    tests/test_reference_graph_golden.py holds this oracle to them: selections, orderings, class ids
    and mask targets exact, fp64 network outputs to 1e-9, losses to 1e-5.  The host-side numpy
    functions of myolo/myolo_utils.py and example/shapes/dataset_shapes.py run as they are
    (tests/golden/make_reference_fixtures.py, tests/test_reference_golden.py).
  * PARITY UNPINNED for the third-party PRIMITIVES only: the TF kernels (``crop_and_resize``,
    ``FusedBatchNorm``, conv / SAME padding, reductions), Keras ``BatchNormalization`` moving update /
    ``Adam`` / ``binary_crossentropy`` and ``keras_applications.mobilenet._depthwise_conv_block`` are
    restated from their published behaviour (here and, independently, in the stand-ins); their
    structure (variable names/shapes, paddings, strides, eps, crop sizes, thresholds, Adam constants,
    BN train flags) is pinned against the GraphDef the reference ships (tests/golden/graph_fixture.json),
    plus hand-computed micro-cases (tests/test_oracle.py) and fp64-vs-fp32 self-consistency.

All ``file:line`` citations are relative to /root/reference/.
Layout everywhere: activations NHWC, conv kernels HWIO, depthwise [3,3,C,1],
transposed-conv kernel [2,2,Cout,Cin] (Keras), parameters keyed by the Keras variable names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # Keras BatchNormalization default; graph fixture: epsilon 0.001
BN_MOMENTUM = 0.99     # Keras default

# (block id, Cin, Cout, stride)  myolo/model.py:68-77 (backbone) and 256-268 (yolo branch)
BACKBONE_BLOCKS = [(1, 32, 64, 1), (2, 64, 64, 2), (3, 64, 128, 1), (4, 128, 256, 2),
                   (5, 256, 256, 1), (6, 256, 512, 1)]
YOLO_BLOCKS = [(7, 512, 512, 2), (8, 512, 512, 1), (9, 512, 512, 1), (10, 512, 512, 1),
               (11, 512, 512, 1), (12, 512, 512, 1), (13, 512, 1024, 2), (14, 1024, 1024, 1)]


# --------------------------------------------------------------------------------------
# primitive ops (third-party semantics restated)
# --------------------------------------------------------------------------------------
def conv2d_nhwc(x, w_hwio, stride=1, pad=0):
    """tf.nn.conv2d NHWC/HWIO with explicit symmetric zero padding `pad`."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def depthwise3x3_nhwc(x, w_33c1, stride):
    """ZeroPadding2D((1,1)) + DepthwiseConv2D 3x3 VALID, stride s  (graph fixture:
    conv_pad_k paddings [[0,0],[1,1],[1,1],[0,0]], conv_dw_k/depthwise padding VALID)."""
    c = x.shape[-1]
    w = w_33c1.permute(2, 3, 0, 1)  # [C,1,3,3]
    y = F.conv2d(x.permute(0, 3, 1, 2), w, stride=stride, padding=1, groups=c)
    return y.permute(0, 2, 3, 1).contiguous()


def relu6(x):  # myolo/model.py:38-39
    return torch.clamp(x, 0.0, 6.0)


def bn_train(x, gamma, beta, eps=BN_EPS):
    """TF FusedBatchNorm is_training=True: normalise with biased batch variance.
    Returns y, batch mean, biased batch variance, element count per channel."""
    c = x.shape[-1]
    xf = x.reshape(-1, c)
    n = xf.shape[0]
    mean = xf.mean(0)
    var = ((xf - mean) ** 2).mean(0)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return y, mean, var, n


def bn_infer(x, gamma, beta, mmean, mvar, eps=BN_EPS):
    """Keras inference-mode BN (myolo_mask_bn2..4 always, everything when learning phase 0)."""
    return (x - mmean) * torch.rsqrt(mvar + eps) * gamma + beta


def moving_update_value(var_biased, n, eps=BN_EPS):
    """Value Keras feeds to the moving variance: TF FusedBatchNorm's `batch_variance` output is
    Bessel-corrected (n/(n-1)), and Keras 2.1.6-2.2.x multiplies by n/(n-(1+eps)) again."""
    bessel = n / max(n - 1, 1)
    return var_biased * bessel * (n / (n - (1.0 + eps)))


def crop_and_resize(image, boxes, box_ind, crop_h, crop_w):
    """tf.image.crop_and_resize (TF 1.x CPU kernel, bilinear, extrapolation_value 0).
    image [B,H,W,C]; boxes [N,4] = (y1,x1,y2,x2) normalised; box_ind [N] int64.
    Differentiable w.r.t. image (== CropAndResizeGradImage), not w.r.t. boxes.  SURVEY Q3."""
    boxes = boxes.detach()
    H, W = image.shape[1], image.shape[2]
    N = boxes.shape[0]
    dt = image.dtype
    y1, x1, y2, x2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    if crop_h > 1:
        hs = (y2 - y1) * (H - 1) / (crop_h - 1)
        in_y = (y1 * (H - 1))[:, None] + torch.arange(crop_h, dtype=dt)[None, :] * hs[:, None]
    else:
        in_y = (0.5 * (y1 + y2) * (H - 1))[:, None]
    if crop_w > 1:
        ws = (x2 - x1) * (W - 1) / (crop_w - 1)
        in_x = (x1 * (W - 1))[:, None] + torch.arange(crop_w, dtype=dt)[None, :] * ws[:, None]
    else:
        in_x = (0.5 * (x1 + x2) * (W - 1))[:, None]
    vy = (in_y >= 0) & (in_y <= H - 1)            # NaN -> False -> extrapolation (0)
    vx = (in_x >= 0) & (in_x <= W - 1)
    in_y = torch.where(vy, in_y, torch.zeros_like(in_y))
    in_x = torch.where(vx, in_x, torch.zeros_like(in_x))
    top = torch.floor(in_y)
    bot = torch.ceil(in_y)
    ly = (in_y - top)[:, :, None, None]
    left = torch.floor(in_x)
    right = torch.ceil(in_x)
    lx = (in_x - left)[:, None, :, None]
    ti, bi, li, ri = top.long(), bot.long(), left.long(), right.long()
    b = box_ind.long()[:, None, None]

    def g(yy, xx):
        return image[b, yy[:, :, None], xx[:, None, :]]          # [N,ch,cw,C]

    tl, tr, bl, br = g(ti, li), g(ti, ri), g(bi, li), g(bi, ri)
    t = tl + (tr - tl) * lx
    bt = bl + (br - bl) * lx
    out = t + (bt - t) * ly
    valid = (vy[:, :, None] & vx[:, None, :])[..., None]
    return torch.where(valid, out, torch.zeros_like(out))


# --------------------------------------------------------------------------------------
# network  (myolo/model.py)
# --------------------------------------------------------------------------------------
class _BNRec:
    """Collects (name, batch mean, value for moving variance) for the Keras moving update."""

    def __init__(self):
        self.items: List[Tuple[str, torch.Tensor, torch.Tensor]] = []


def _bn(x, P, name, training, rec):
    g, b = P[name + "/gamma"], P[name + "/beta"]
    if training:
        y, mean, var, n = bn_train(x, g, b)
        if rec is not None:
            rec.items.append((name, mean.detach(), moving_update_value(var.detach(), n)))
        return y
    return bn_infer(x, g, b, P[name + "/moving_mean"], P[name + "/moving_variance"])


def conv_block(x, P, training, rec):
    """myolo/model.py:42-52  ZeroPad(1,1) -> 3x3 s2 VALID no bias -> BN -> ReLU6."""
    x = conv2d_nhwc(x, P["conv1/kernel"], stride=2, pad=1)
    return relu6(_bn(x, P, "conv1_bn", training, rec))


def depthwise_conv_block(x, P, k, stride, training, rec):
    """keras_applications.mobilenet._depthwise_conv_block (older, symmetric-pad variant pinned by
    the graph fixture); call sites myolo/model.py:68-77, 256-268."""
    x = depthwise3x3_nhwc(x, P[f"conv_dw_{k}/depthwise_kernel"], stride)
    x = relu6(_bn(x, P, f"conv_dw_{k}_bn", training, rec))
    x = conv2d_nhwc(x, P[f"conv_pw_{k}/kernel"], 1, 0)
    return relu6(_bn(x, P, f"conv_pw_{k}_bn", training, rec))


def mobilenet_graph(x, P, training, rec=None):
    """myolo/model.py:55-79 -> C4 [B,S/8,S/8,512]."""
    x = conv_block(x, P, training, rec)
    for k, _, _, s in BACKBONE_BLOCKS:
        x = depthwise_conv_block(x, P, k, s, training, rec)
    return x


def yolo_branch_graph(c4, P, cfg, training, rec=None):
    """myolo/model.py:249-278 -> [B,G,G,N_BOX,5+NC]."""
    x = c4
    for k, _, _, s in YOLO_BLOCKS:
        x = depthwise_conv_block(x, P, k, s, training, rec)
    x = conv2d_nhwc(x, P["conv_23/kernel"], 1, 0) + P["conv_23/bias"]
    B = x.shape[0]
    return x.reshape(B, cfg["GRID_H"], cfg["GRID_W"], cfg["N_BOX"], 5 + cfg["NUM_CLASSES"])


def _cell_grid(cfg, dt):
    """myolo/model.py:89-93 / 1445-1449: cell[...,0] = column index, cell[...,1] = its transpose
    (row index when GRID_H == GRID_W)."""
    gh, gw = cfg["GRID_H"], cfg["GRID_W"]
    cell_x = torch.arange(gw, dtype=dt).repeat(gh).reshape(1, gh, gw, 1, 1)
    cell_y = cell_x.permute(0, 2, 1, 3, 4)
    return torch.cat([cell_x, cell_y], -1)                      # [1,gh,gw,1,2] broadcast over B, N_BOX


def _anchors(cfg, dt):
    return torch.tensor(np.reshape(np.asarray(cfg["ANCHORS"], dtype=np.float64), [1, 1, 1, cfg["N_BOX"], 2]), dtype=dt)


def decode_yolo(y_pred, cfg):
    """DecodeYOLOLayer.call myolo/model.py:1442-1473 -> [B,R,4] (x1,y1,x2,y2) normalised."""
    dt = y_pred.dtype
    gw = float(cfg["GRID_W"])
    xy = (torch.sigmoid(y_pred[..., :2]) + _cell_grid(cfg, dt)) / gw
    wh = torch.exp(y_pred[..., 2:4]) * _anchors(cfg, dt) / gw
    half = wh / 2.0
    out = torch.cat([xy - half, xy + half], -1)
    return out.reshape(y_pred.shape[0], -1, 4)


def detections_layer(y_pred, cfg):
    """DetectionsLayer.call myolo/model.py:1493-1538 -> [B,R,6]."""
    boxes = decode_yolo(y_pred, cfg)
    B = y_pred.shape[0]
    conf = torch.sigmoid(y_pred[..., 4]).reshape(B, -1, 1)
    cls = torch.argmax(y_pred[..., 5:], -1).to(y_pred.dtype).reshape(B, -1, 1)
    return torch.cat([boxes, conf, cls], -1)


def norm_boxes_graph(boxes, h, w):
    """myolo/model.py:1394-1408.  `shape` = image.shape[1:3]; the reference unpacks it as (w, h)
    -> scale = [s0, s1, s0, s1] - 1, shift [0,0,1,1]."""
    dt = boxes.dtype
    scale = torch.tensor([h, w, h, w], dtype=dt) - 1.0
    shift = torch.tensor([0.0, 0.0, 1.0, 1.0], dtype=dt)
    return (boxes - shift) / scale


def overlaps_graph(b1, b2):
    """myolo/model.py:420-454, boxes (x1,y1,x2,y2)."""
    a = b1[:, None, :]
    b = b2[None, :, :]
    x1 = torch.maximum(a[..., 0], b[..., 0])
    y1 = torch.maximum(a[..., 1], b[..., 1])
    x2 = torch.minimum(a[..., 2], b[..., 2])
    y2 = torch.minimum(a[..., 3], b[..., 3])
    inter = torch.clamp(x2 - x1, min=0) * torch.clamp(y2 - y1, min=0)
    a1 = (a[..., 3] - a[..., 1]) * (a[..., 2] - a[..., 0])
    a2 = (b[..., 3] - b[..., 1]) * (b[..., 2] - b[..., 0])
    union = a1 + a2 - inter
    return inter / union


def detect_mask_target_graph(proposals, gt_class_ids, gt_boxes, gt_masks, cfg):
    """myolo/model.py:457-602 for ONE image.  proposals [R,4] normalised; gt_class_ids [M] int;
    gt_boxes [M,4] normalised (already through norm_boxes_graph); gt_masks [S,S,M] bool."""
    proposals = proposals.detach()
    dt = proposals.dtype
    R = cfg["TRAIN_ROIS_PER_IMAGE"]
    mh, mw = cfg["MASK_SHAPE"]
    # trim_zeros_graph 1411-1420 (never trims padded rows in practice)
    nz = gt_boxes.abs().sum(1) != 0
    gt_boxes = gt_boxes[nz]
    gt_class_ids = gt_class_ids[nz]
    gt_masks = gt_masks[:, :, nz]
    ov = overlaps_graph(proposals, gt_boxes)                     # [R,M']
    if ov.shape[1] > 0:
        # tf.reduce_max: NaN compares false -> keep TF semantics via explicit nan handling
        iou_max = torch.amax(ov, 1)
    else:
        iou_max = torch.full((proposals.shape[0],), -float("inf"), dtype=dt)
    pos = torch.nonzero(iou_max >= 0.5)[:, 0]
    neg = torch.nonzero(iou_max < 0.5)[:, 0]
    pos_rois = proposals[pos]
    neg_rois = proposals[neg]
    if ov.shape[1] > 0 and pos.numel() > 0:
        assign = torch.argmax(ov[pos], 1)
    else:
        assign = torch.zeros((0,), dtype=torch.long)
    ids = gt_class_ids[assign].to(torch.int32)
    if pos.numel() > 0:
        roi_masks = gt_masks.permute(2, 0, 1)[assign].to(dt)[..., None]   # [P,S,S,1]
        boxes = torch.stack([pos_rois[:, 1], pos_rois[:, 0], pos_rois[:, 3], pos_rois[:, 2]], 1)
        masks = crop_and_resize(roi_masks, boxes, torch.arange(pos.numel()), mh, mw)[..., 0]
        masks = torch.round(masks)                               # tf.round = half-to-even
    else:
        masks = torch.zeros((0, mh, mw), dtype=dt)
    rois = torch.cat([pos_rois, neg_rois], 0)
    npad = max(R - rois.shape[0], 0)
    rois = torch.cat([rois, torch.zeros((npad, 4), dtype=dt)], 0)
    nneg = neg_rois.shape[0]
    ids = torch.cat([ids, torch.zeros((nneg + npad,), dtype=torch.int32)], 0)
    masks = torch.cat([masks, torch.zeros((nneg + npad, mh, mw), dtype=dt)], 0)
    return rois, ids, masks


def detect_mask_targets(proposals, gt_class_ids, gt_boxes_norm, gt_masks, cfg):
    """DetectMaskTargetLayer.call myolo/model.py:635-649 (batch_slice: myolo_utils.py:929-963)."""
    outs = [detect_mask_target_graph(proposals[i], gt_class_ids[i], gt_boxes_norm[i], gt_masks[i], cfg)
            for i in range(proposals.shape[0])]
    return tuple(torch.stack(o, 0) for o in zip(*outs))


def pyramid_roi_align(rois, feat, pool):
    """PyramidROIAlign.call myolo/model.py:327-410.  Level is clamped to 0 for every box; boxes
    are handed to crop_and_resize AS (x1,y1,x2,y2) although the op reads (y1,x1,y2,x2)
    (385-387) -> x/y swapped sampling, reproduced as-is.  The final top-k reorder
    (402-405) is the identity.  Returns [B,R,pool,pool,C]."""
    B, R = rois.shape[0], rois.shape[1]
    boxes = rois.reshape(-1, 4)
    idx = torch.arange(B).repeat_interleave(R)
    out = crop_and_resize(feat, boxes, idx, pool, pool)
    return out.reshape(B, R, pool, pool, feat.shape[-1])


def build_mask_graph(rois, feat, P, cfg, training, rec=None):
    """myolo/model.py:668-715 -> [B,R,28,28,NC].  bn1 follows the learning phase (690), bn2..4 are
    called with training=train_bn=False -> moving statistics always (695-708)."""
    pool = cfg["MASK_POOL_SIZE"]
    x = pyramid_roi_align(rois, feat, pool)
    B, R = x.shape[0], x.shape[1]
    x = x.reshape(B * R, pool, pool, -1)
    x = conv2d_nhwc(x, P["myolo_mask_conv1/kernel"], 1, 1) + P["myolo_mask_conv1/bias"]
    x = torch.relu(_bn(x, P, "myolo_mask_bn1", training, rec))
    for i in (2, 3, 4):
        x = conv2d_nhwc(x, P[f"myolo_mask_conv{i}/kernel"], 1, 1) + P[f"myolo_mask_conv{i}/bias"]
        x = torch.relu(_bn(x, P, f"myolo_mask_bn{i}", False, None))
    # Conv2DTranspose 2x2 s2 VALID, kernel [2,2,Cout,Cin]: out[2i+a,2j+b,co] = sum_ci in[i,j,ci] K[a,b,co,ci]
    wd = P["myolo_mask_deconv/kernel"]
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), wd.permute(3, 2, 0, 1), stride=2)
    x = torch.relu(y.permute(0, 2, 3, 1) + P["myolo_mask_deconv/bias"])
    x = conv2d_nhwc(x.contiguous(), P["myolo_mask/kernel"], 1, 0) + P["myolo_mask/bias"]
    x = torch.sigmoid(x)
    return x.reshape(B, R, x.shape[1], x.shape[2], x.shape[3])


def myolo_mask_loss_graph(target_masks, target_class_ids, pred_masks):
    """myolo/model.py:718-754 with Keras K.binary_crossentropy (probabilities -> clipped logits)."""
    dt = pred_masks.dtype
    ids = target_class_ids.reshape(-1)
    tm = target_masks.reshape(-1, target_masks.shape[2], target_masks.shape[3])
    pm = pred_masks.reshape(-1, pred_masks.shape[2], pred_masks.shape[3], pred_masks.shape[4])
    pos = torch.nonzero(ids > 0)[:, 0]
    if pos.numel() == 0:
        return torch.zeros((), dtype=dt)
    y_true = tm[pos]
    y_pred = pm[pos, :, :, ids[pos].long()]
    eps = 1e-7
    p = torch.clamp(y_pred, eps, 1 - eps)
    z = torch.log(p / (1 - p))
    loss = torch.clamp(z, min=0) - z * y_true + torch.log1p(torch.exp(-torch.abs(z)))
    return loss.mean()


def yolo_custom_loss(y_true, y_pred, true_boxes, cfg, seen=1.0):
    """myolo/model.py:86-242 (config read from the instance).  `seen` is the value of
    the counter AFTER this batch's assign_add (197)."""
    dt = y_pred.dtype
    cell = _cell_grid(cfg, dt)
    anc = _anchors(cfg, dt)
    pxy = torch.sigmoid(y_pred[..., :2]) + cell
    pwh = torch.exp(y_pred[..., 2:4]) * anc
    pconf = torch.sigmoid(y_pred[..., 4])
    pcls = y_pred[..., 5:]
    txy = y_true[..., 0:2]
    twh = y_true[..., 2:4]
    tmin, tmax = txy - twh / 2.0, txy + twh / 2.0
    pmin, pmax = pxy - pwh / 2.0, pxy + pwh / 2.0
    iwh = torch.clamp(torch.minimum(pmax, tmax) - torch.maximum(pmin, tmin), min=0.0)
    inter = iwh[..., 0] * iwh[..., 1]
    union = pwh[..., 0] * pwh[..., 1] + twh[..., 0] * twh[..., 1] - inter
    iou = inter / union
    tconf = iou * y_true[..., 4]
    tcls = torch.argmax(y_true[..., 5:], -1)
    coord_mask = y_true[..., 4:5] * cfg["COORD_SCALE"]
    # best IoU against the true-box buffer (159-184)
    tb = true_boxes.to(dt)                                       # [B,1,1,1,TB,4]
    bxy, bwh = tb[..., 0:2], tb[..., 2:4]
    bmin, bmax = bxy - bwh / 2.0, bxy + bwh / 2.0
    qxy, qwh = pxy.unsqueeze(4), pwh.unsqueeze(4)
    qmin, qmax = qxy - qwh / 2.0, qxy + qwh / 2.0
    iwh2 = torch.clamp(torch.minimum(qmax, bmax) - torch.maximum(qmin, bmin), min=0.0)
    inter2 = iwh2[..., 0] * iwh2[..., 1]
    union2 = qwh[..., 0] * qwh[..., 1] + bwh[..., 0] * bwh[..., 1] - inter2
    best = torch.amax(inter2 / union2, 4)
    obj = y_true[..., 4]
    conf_mask = (best < 0.6).to(dt) * (1 - obj) * cfg["NO_OBJECT_SCALE"] + obj * cfg["OBJECT_SCALE"]
    cw = torch.as_tensor(np.asarray(cfg["CLASS_WEIGHTS"], dtype=np.float64), dtype=dt)
    class_mask = obj * cw[tcls] * cfg["CLASS_SCALE"]
    if seen < cfg["WARM_UP_BATCHES"]:                            # 196-207 (dead at WARM_UP_BATCHES=0)
        nobox = (coord_mask < cfg["COORD_SCALE"] / 2.0).to(dt)
        txy = txy + (0.5 + cell) * nobox
        twh = twh + torch.ones_like(twh) * anc * nobox
        coord_mask = torch.ones_like(coord_mask)
    nb_coord = (coord_mask > 0).to(dt).sum()
    nb_conf = (conf_mask > 0).to(dt).sum()
    nb_class = (class_mask > 0).to(dt).sum()
    loss_xy = ((txy - pxy) ** 2 * coord_mask).sum() / (nb_coord + 1e-6) / 2.0
    loss_wh = ((twh - pwh) ** 2 * coord_mask).sum() / (nb_coord + 1e-6) / 2.0
    loss_conf = ((tconf - pconf) ** 2 * conf_mask).sum() / (nb_conf + 1e-6) / 2.0
    ce = F.cross_entropy(pcls.reshape(-1, pcls.shape[-1]), tcls.reshape(-1), reduction="none").reshape(tcls.shape)
    loss_class = (ce * class_mask).sum() / (nb_class + 1e-6)
    return loss_xy + loss_wh + loss_conf + loss_class


# --------------------------------------------------------------------------------------
# whole-model passes  (MaskYOLO.build, myolo/model.py:787-941)
# --------------------------------------------------------------------------------------
def forward_training(P, inputs, cfg, seen=1.0):
    """mode='training' graph (844-901).  inputs = [image, true_boxes, yolo_target, gt_class_ids,
    gt_boxes(px), gt_masks(bool)].  Returns dict of the 6 model outputs + total loss + BN record."""
    image, true_boxes, yolo_target, gt_ids, gt_boxes, gt_masks = inputs
    rec = _BNRec()
    c4 = mobilenet_graph(image, P, True, rec)
    feat = conv2d_nhwc(c4, P["feature_map/kernel"], 1, 1) + P["feature_map/bias"]      # 848
    yolo_out = yolo_branch_graph(c4, P, cfg, True, rec)                                # 851-852
    proposals = decode_yolo(yolo_out, cfg)                                             # 874
    gtb = norm_boxes_graph(gt_boxes.to(image.dtype), image.shape[1], image.shape[2])   # 819-820
    rois, tids, tmasks = detect_mask_targets(proposals, gt_ids, gtb, gt_masks, cfg)    # 876-878
    masks = build_mask_graph(rois, feat, P, cfg, True, rec)                            # 880-882
    yl = yolo_custom_loss(yolo_target.to(image.dtype), yolo_out, true_boxes, cfg, seen)  # 888-889
    ml = myolo_mask_loss_graph(tmasks, tids, masks)                                    # 892-893
    lw = cfg.get("LOSS_WEIGHTS", {})
    total = yl * lw.get("yolo_sum_loss", 1.0) + ml * lw.get("myolo_mask_loss", 1.0)    # compile 1087-1094
    return dict(yolo_output=yolo_out, yolo_proposals=proposals, output_rois=rois, myolo_mask=masks,
                yolo_sum_loss=yl, mask_loss=ml, loss=total, target_class_ids=tids, target_mask=tmasks,
                feature_map=feat, c4=c4, bn_record=rec)


def forward_inference(P, image, cfg):
    """mode='inference' graph (922-936): learning phase 0 -> every BN uses moving statistics."""
    c4 = mobilenet_graph(image, P, False)
    feat = conv2d_nhwc(c4, P["feature_map/kernel"], 1, 1) + P["feature_map/bias"]
    yolo_out = yolo_branch_graph(c4, P, cfg, False)
    det = detections_layer(yolo_out, cfg)
    masks = build_mask_graph(det[..., :4], feat, P, cfg, False)
    return dict(yolo_output=yolo_out, detections=det, myolo_mask=masks)


def trainable_names(P):
    return sorted(k for k in P if k.rsplit("/", 1)[1] in ("kernel", "bias", "gamma", "beta", "depthwise_kernel"))


def train_step(P, opt_state, inputs, cfg, lr=1e-3):
    """One Keras fit step: forward (training phase), backward, Adam (compile 1071-1075, Keras
    form), BN moving-average update (zero-debiased assign_moving_average, momentum 0.99).
    P is updated IN PLACE; returns the forward outputs and the gradient dict."""
    names = trainable_names(P)
    leaves = {k: P[k].detach().clone().requires_grad_(True) for k in names}
    Q = dict(P)
    Q.update(leaves)
    opt_state["seen"] = opt_state.get("seen", 0.0) + 1.0
    out = forward_training(Q, inputs, cfg, seen=opt_state["seen"])
    grads = torch.autograd.grad(out["loss"], [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, grads)}
    t = opt_state.get("t", 0) + 1
    opt_state["t"] = t
    b1, b2, eps = 0.9, 0.999, 1e-8
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    with torch.no_grad():
        for k in names:
            g = grads[k]
            m = opt_state.setdefault("m/" + k, torch.zeros_like(g))
            v = opt_state.setdefault("v/" + k, torch.zeros_like(g))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            P[k] = P[k] - lr_t * m / (torch.sqrt(v) + eps)
        for name, mean, varv in out["bn_record"].items:
            for suffix, val in (("moving_mean", mean), ("moving_variance", varv)):
                key = f"{name}/{suffix}"
                biased = opt_state.setdefault("biased/" + key, torch.zeros_like(val))
                step = opt_state.get("step/" + key, 0) + 1
                opt_state["step/" + key] = step
                biased.sub_((biased - val) * (1 - BN_MOMENTUM))
                P[key] = biased / (1 - BN_MOMENTUM ** step)
    out = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return out, grads
