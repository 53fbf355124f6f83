"""Seeded inputs of the reference-graph golden cases, shared by the generator (make_reference_graph_fixtures.py, runs
where /root/reference exists) and the tests (which regenerate the same inputs instead of storing them).  Only
np.random.RandomState is used: its streams are frozen across numpy versions."""
import numpy as np

CASES = {
    # the Shapes configuration of the shipped graph: 3 anchors, 4 classes, 15-wide box buffer
    "shapes": dict(B=2, G=7, NB=3, NC=4, TB=15, M=15, S=96, C=4, seed=101,
                   ANCHORS=[1.27273, 1.277385, 2.47446, 2.56253, 4.03843, 4.07434], CLASS_WEIGHTS=[1.0, 1.0, 1.0, 1.0]),
    # base Config values at HEAD: 5 anchors, 2 classes, 10-wide buffer; non-trivial class weights and scales
    "base": dict(B=1, G=4, NB=5, NC=2, TB=10, M=10, S=64, C=8, seed=202,
                 ANCHORS=[1.27, 1.31, 1.95, 1.85, 2.40, 2.72, 3.20, 3.32, 5.06, 5.05], CLASS_WEIGHTS=[0.5, 2.0],
                 OBJECT_SCALE=5.0, NO_OBJECT_SCALE=0.7, COORD_SCALE=1.5, CLASS_SCALE=1.2),
    # the second image has NO ground truth at all (every padded row is zero): trim_zeros leaves nothing, the IoU matrix
    # is [R, 0], every ROI is negative (model.py:487-545 on empty tensors)
    "nogt": dict(B=2, G=4, NB=3, NC=4, TB=6, M=6, S=64, C=4, seed=505, empty_images=[1],
                 ANCHORS=[0.6, 0.6, 1.2, 1.3, 2.0, 2.1], CLASS_WEIGHTS=[1.0, 1.0, 1.0, 1.0]),
}


def fuzz_case(i):
    """Random configuration number i for the live differential test (tests/test_reference_live_fuzz.py)."""
    rs = np.random.RandomState(9000 + i)
    NB, NC = int(rs.randint(1, 6)), int(rs.randint(2, 7))
    TB = int(rs.randint(5, 13))
    return dict(B=int(rs.randint(1, 4)), G=int(rs.randint(2, 8)), NB=NB, NC=NC, TB=TB, M=TB, S=int(rs.choice([48, 64, 96])), C=4,
                seed=7000 + i, ANCHORS=[float(v) for v in rs.uniform(0.4, 4.5, size=2 * NB)],
                CLASS_WEIGHTS=[float(v) for v in rs.uniform(0.5, 2.0, size=NC)], OBJECT_SCALE=float(rs.uniform(1, 6)),
                NO_OBJECT_SCALE=float(rs.uniform(0.3, 1.5)), COORD_SCALE=float(rs.uniform(0.5, 2.0)), CLASS_SCALE=float(rs.uniform(0.5, 2.0)),
                empty_images=[0] if i % 7 == 3 else [])


def build(name):
    """-> dict of numpy inputs for one case (a name in CASES, or a dict such as fuzz_case(i) returns).  Ground truth: `n` axis-aligned ellipses per image; y_pred: N(0,1) logits,
    except that the predictor responsible for each instance decodes to (roughly) the instance's box, so that positive
    ROIs, class assignments and mask targets are exercised."""
    c = dict(CASES[name]) if isinstance(name, str) else dict(name)
    rng = np.random.RandomState(c["seed"])
    B, G, NB, NC, TB, M, S = c["B"], c["G"], c["NB"], c["NC"], c["TB"], c["M"], c["S"]
    anchors = np.asarray(c["ANCHORS"], np.float64).reshape(NB, 2)
    y_pred = rng.standard_normal((B, G, G, NB, 5 + NC)).astype(np.float32)
    y_true = np.zeros((B, G, G, NB, 5 + NC), np.float32)
    true_boxes = np.zeros((B, 1, 1, 1, TB, 4), np.float32)
    ids = np.zeros((B, TB), np.int32)
    boxes_px = np.zeros((B, TB, 4), np.float32)
    masks = np.zeros((B, S, S, M), bool)
    yy, xx = np.mgrid[0:S, 0:S]
    for b in range(B):
        n = int(rng.randint(2, 5))
        if b in c.get("empty_images", ()):
            n = 0
        for k in range(n):
            w, h = rng.uniform(0.18, 0.45, size=2) * S
            cx, cy = rng.uniform(0.25, 0.75, size=2) * S
            x1, y1, x2, y2 = int(cx - w / 2), int(cy - h / 2), int(cx + w / 2), int(cy + h / 2)
            masks[b, :, :, k] = ((xx - (x1 + x2 - 1) / 2.0) / ((x2 - x1) / 2.0)) ** 2 + \
                                ((yy - (y1 + y2 - 1) / 2.0) / ((y2 - y1) / 2.0)) ** 2 <= 1.0
            ids[b, k] = int(rng.randint(1, NC))
            boxes_px[b, k] = (x1, y1, x2, y2)
            # BatchGenerator-style encoding in grid units
            gcx, gcy = 0.5 * (x1 + x2) / (S / G), 0.5 * (y1 + y2) / (S / G)
            gw, gh = (x2 - x1) / (S / G), (y2 - y1) / (S / G)
            gx, gy = int(gcx), int(gcy)
            a = int(np.argmin(np.abs(anchors[:, 0] - gw) + np.abs(anchors[:, 1] - gh)))
            y_true[b, gy, gx, a, 0:4] = (gcx, gcy, gw, gh)
            y_true[b, gy, gx, a, 4] = 1.0
            y_true[b, gy, gx, a, 5:] = 0.0
            y_true[b, gy, gx, a, 5 + ids[b, k]] = 1.0
            true_boxes[b, 0, 0, 0, k] = (gcx, gcy, gw, gh)
            # responsible predictor decodes close to the instance (with a little noise)
            fx, fy = np.clip(gcx - gx, 0.05, 0.95), np.clip(gcy - gy, 0.05, 0.95)
            y_pred[b, gy, gx, a, 0] = np.log(fx / (1 - fx)) + 0.1 * rng.standard_normal()
            y_pred[b, gy, gx, a, 1] = np.log(fy / (1 - fy)) + 0.1 * rng.standard_normal()
            y_pred[b, gy, gx, a, 2] = np.log(gw / anchors[a, 0]) + 0.05 * rng.standard_normal()
            y_pred[b, gy, gx, a, 3] = np.log(gh / anchors[a, 1]) + 0.05 * rng.standard_normal()
    feat = rng.standard_normal((B, S // 8, S // 8, c["C"])).astype(np.float32)
    R = G * G * NB
    pred_masks = rng.uniform(0.0, 1.0, size=(B, R, 28, 28, NC)).astype(np.float32)
    pred_masks[0, 0, 0, :4, :] = (0.0, 1.0, 1e-9, 1.0 - 1e-9)[:NC] if NC <= 4 else 0.5   # exercise the 1e-7 clip
    c.update(y_pred=y_pred, y_true=y_true, true_boxes=true_boxes, gt_class_ids=ids, gt_boxes_px=boxes_px, gt_masks=masks,
             feat=feat, pred_masks=pred_masks, R=R)
    return c


# ---- network cases: the reference's graph BUILDERS (conv_block / mobilenet_graph / yolo_branch_graph / build_mask_graph)
NET = dict(B=2, S=64, NB=3, NC=4, R=10, seed=303)


def weights(nb, nc, seed):
    """Every variable of the model (names and shapes from the product's param_specs, which a CPU test pins to the
    reference's GraphDef) filled from a frozen RandomState stream: He-scaled kernels, non-trivial BN statistics."""
    from myolo.engine import param_specs
    rs = np.random.RandomState(seed)
    P = {}
    for name, shape, _ in param_specs(nb, nc):
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("kernel", "depthwise_kernel"):
            fan_in = shape[0] * shape[1] * (shape[2] if leaf == "kernel" else 1)
            if name.startswith("myolo_mask_deconv"):
                fan_in = shape[3]
            P[name] = (rs.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        elif leaf in ("gamma", "moving_variance"):
            P[name] = rs.uniform(0.6, 1.4, size=shape).astype(np.float32)
        elif leaf in ("beta", "moving_mean"):
            P[name] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
        else:
            P[name] = (0.05 * rs.standard_normal(shape)).astype(np.float32)
    return P


def net_fuzz_case(i):
    rs = np.random.RandomState(9500 + i)
    return dict(B=int(rs.randint(1, 3)), S=int(rs.choice([64, 96, 128])), NB=int(rs.choice([1, 2, 3, 5])), NC=int(rs.choice([2, 4, 5])),
                R=int(rs.randint(3, 9)), seed=9600 + i)


def net_inputs(case=None):
    c = dict(NET if case is None else case)
    rs = np.random.RandomState(c["seed"] + 1)
    c["image"] = rs.uniform(0.0, 1.0, size=(c["B"], c["S"], c["S"], 3)).astype(np.float32)
    F = c["S"] // 8
    c["feat"] = rs.standard_normal((c["B"], F, F, 256)).astype(np.float32)
    x1y1 = rs.uniform(-0.1, 0.6, size=(c["B"], c["R"], 2))
    wh = rs.uniform(0.1, 0.6, size=(c["B"], c["R"], 2))
    rois = np.concatenate([x1y1, x1y1 + wh], -1).astype(np.float32)
    rois[:, -1] = 0.0                                              # a zero-padded ROI row
    c["rois"] = rois
    return c


# ---- whole-model case: MaskYOLO.build (787-941), modes 'training' and 'inference'
BUILD = dict(B=2, S=64, G=2, NB=3, NC=4, TB=15, M=15, R=12, seed=404, ANCHORS=[0.5, 0.6, 0.8, 0.7, 1.1, 1.2])


def build_image():
    c = dict(BUILD)
    rs = np.random.RandomState(c["seed"] + 1)
    c["image"] = rs.uniform(0.0, 1.0, size=(c["B"], c["S"], c["S"], 3)).astype(np.float32)
    return c


def gt_from_boxes(c, gt_class_ids, gt_boxes_px):
    """Masks (ellipse inscribed in each pixel box), YOLO target and true-box buffer for given padded gt arrays."""
    B, S, G, NB, NC, TB, M = c["B"], c["S"], c["G"], c["NB"], c["NC"], c["TB"], c["M"]
    masks = np.zeros((B, S, S, M), bool)
    y_true = np.zeros((B, G, G, NB, 5 + NC), np.float32)
    true_boxes = np.zeros((B, 1, 1, 1, TB, 4), np.float32)
    yy, xx = np.mgrid[0:S, 0:S]
    for b in range(B):
        for k in range(TB):
            if gt_class_ids[b, k] == 0:
                continue
            x1, y1, x2, y2 = [float(v) for v in gt_boxes_px[b, k]]
            masks[b, :, :, k] = ((xx - (x1 + x2 - 1) / 2.0) / ((x2 - x1) / 2.0)) ** 2 + \
                                ((yy - (y1 + y2 - 1) / 2.0) / ((y2 - y1) / 2.0)) ** 2 <= 1.0
            gcx, gcy, gw, gh = 0.5 * (x1 + x2) / (S / G), 0.5 * (y1 + y2) / (S / G), (x2 - x1) / (S / G), (y2 - y1) / (S / G)
            gx, gy = min(int(gcx), G - 1), min(int(gcy), G - 1)
            y_true[b, gy, gx, k % NB, 0:5] = (gcx, gcy, gw, gh, 1.0)
            y_true[b, gy, gx, k % NB, 5:] = 0.0
            y_true[b, gy, gx, k % NB, 5 + gt_class_ids[b, k]] = 1.0
            true_boxes[b, 0, 0, 0, k] = (gcx, gcy, gw, gh)
    return masks, y_true, true_boxes


# ---- decode_masks / unmold_mask case (model.py:1330-1391, myolo_utils.py:883-912)
def decode_masks_inputs(S=96, N=14, NC=4, seed=606):
    """detections [1,N,6] (normalised boxes incl. boxes leaving the image, a zero-area and an inverted one; score; class)
    and myolo_mask [1,N,28,28,NC]."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(0.1, 0.9, size=(N, 2))
    wh = rs.uniform(0.08, 0.6, size=(N, 2))
    det = np.zeros((1, N, 6), np.float32)
    det[0, :, 0:2], det[0, :, 2:4] = c - wh / 2, c + wh / 2            # some corners fall below 0 or above 1
    det[0, 3, 2] = det[0, 3, 0]                                          # zero area -> filtered out
    det[0, 5, [0, 2]] = det[0, 5, [2, 0]]                                # inverted -> negative area -> filtered out
    det[0, 7, :4] = (0.97, 0.2, 1.4, 0.9)                                # hugs the right border
    det[0, :, 4] = rs.uniform(0, 1, size=N)
    det[0, :, 5] = rs.randint(0, NC, size=N)
    masks = rs.uniform(0, 1, size=(1, N, 28, 28, NC)).astype(np.float32)
    yy, xx = np.mgrid[0:28, 0:28]
    masks[0, :, :, :, :] *= 0.2
    for i in range(N):                                                   # a blob per detection so that masks have structure
        cy, cx, r = rs.uniform(8, 20), rs.uniform(8, 20), rs.uniform(4, 12)
        masks[0, i, :, :, int(det[0, i, 5])] += 0.75 * (((yy - cy) ** 2 + (xx - cx) ** 2) <= r * r)
    return dict(S=S, N=N, NC=NC, detections=det, myolo_mask=masks)
