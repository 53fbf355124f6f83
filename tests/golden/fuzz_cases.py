"""Seeded random cases for the live differential test of the host functions (tests/test_reference_live_fuzz.py): `run(U, n)`
drives ANY implementation `U` of the myolo_utils interface -- the reference's module or the package's -- through the same
n cases and returns comparable plain arrays."""
import numpy as np


class Cfg(object):
    BATCH_SIZE = 3
    IMAGE_SHAPE = [224, 224, 3]                 # the reference's BatchGenerator hard-codes 224 x 224 buffers (737, 744)
    GRID_H = GRID_W = 7
    N_BOX = 3
    NUM_CLASSES = 4
    ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434]
    TRUE_BOX_BUFFER = 6
    MAX_GT_INSTANCES = 6


def run(U, n):
    rs = np.random.RandomState(4242)
    out = {}
    for k in range(n):
        # decode_one_yolo_output (myolo_utils.py:36-85)
        G, NB, NC = int(rs.randint(2, 6)), 3, int(rs.randint(2, 6))
        netout = rs.normal(0, 1.5, size=(G, G, NB, 5 + NC))
        netout[..., 4] += rs.uniform(-1, 1)
        ot, nt = float(rs.uniform(0.15, 0.6)), float(rs.uniform(0.2, 0.7))
        got = U.decode_one_yolo_output(netout.copy(), Cfg.ANCHORS, NC, obj_threshold=ot, nms_threshold=nt)
        out["dec%d" % k] = np.array([[b.xmin, b.ymin, b.xmax, b.ymax, b.c, b.get_label(), b.get_score()] for b in got]).reshape(-1, 7)
        # NMB (88-113) + bbox_iou_2 (201-228) + bbox_iou (187-198)
        m = int(rs.randint(1, 12))
        c = rs.uniform(0.2, 0.8, size=(m, 2))
        wh = rs.uniform(0.05, 0.5, size=(m, 2))
        bx = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
        if m > 2:
            bx[1] = bx[0] + rs.uniform(-0.02, 0.02, size=4)
        cls = rs.randint(1, 3, size=m)
        idx = rs.permutation(100)[:m]
        out["nmb%d" % k] = np.asarray(U.NMB(bx, cls, idx.copy(), [224, 224, 3], nms_threshold=float(rs.uniform(0.2, 0.8))))
        out["iou2_%d" % k] = np.array([U.bbox_iou_2(bx[i], bx[(i + 1) % m], [224, 224, 3]) for i in range(m)])
        out["iou_%d" % k] = np.array([U.bbox_iou(U.BoundBox(*bx[i]), U.BoundBox(*bx[(i + 1) % m])) for i in range(m)])
        # extract_bboxes (247-271) on random blobs, one channel left empty
        S = 40
        yy, xx = np.mgrid[0:S, 0:S]
        masks = np.zeros((S, S, 4), bool)
        for ch in range(3):
            cy, cx, r = rs.uniform(0, S, size=3)
            masks[:, :, ch] = (yy - cy) ** 2 + (xx - cx) ** 2 <= (0.4 * r) ** 2
        out["eb%d" % k] = U.extract_bboxes(masks)
    # BatchGenerator.__getitem__ (727-860) on random integer boxes: border cells, shared cells, 0..8 instances (> buffer)
    S = Cfg.IMAGE_SHAPE[0]
    info = []
    for i in range(max(3, n // 4)):
        k = int(rs.randint(0, 9))                # up to 8 instances: more than the 6-wide buffer -> np.random.choice
        x1 = rs.randint(0, S - 8, size=k)
        y1 = rs.randint(0, S - 8, size=k)
        boxes = np.stack([x1, y1, np.minimum(x1 + rs.randint(4, 120, size=k), S), np.minimum(y1 + rs.randint(4, 120, size=k), S)], 1).astype(np.int32) \
            if k else np.zeros((0, 4), np.int32)
        ids = rs.randint(1, Cfg.NUM_CLASSES, size=k).astype(np.int32)
        gm = np.zeros((S, S, k), bool)
        for j in range(k):
            gm[boxes[j, 1]:boxes[j, 3], boxes[j, 0]:boxes[j, 2], j] = True
        image = rs.randint(0, 256, size=(S, S, 3)).astype(np.uint8)
        info.append([image, ids, boxes, gm])
    gen = U.BatchGenerator(info, Cfg(), mode="training", shuffle=False, norm=True)
    for b in range(len(gen)):
        np.random.seed(1000 + b)                    # the sub-sampling of over-full images draws from the global generator
        inputs, _ = gen[b]
        for name, arr in zip(("images", "true_boxes", "yolo_target", "gt_class_ids", "gt_boxes", "gt_masks"), inputs):
            out["bg%d_%s" % (b, name)] = np.asarray(arr) if name != "images" else np.asarray(arr).astype(np.float64).sum(axis=(1, 2))
    return out
