"""Host data path of Mask-YOLO (reference: myolo/myolo_utils.py): ground-truth loading, the YOLO
target encoding done by `BatchGenerator` / `data_generator`, box helpers and the numpy
post-processing used by `MaskYOLO.detect`.  Same names, arguments and conventions as the reference
(boxes are (x1, y1, x2, y2) with x2/y2 exclusive, myolo_utils.py:247-271); everything here is
numpy on the host -- the device side starts where the arrays returned by `__getitem__` are copied
into the engine's pinned staging buffers (myolo/model.py).
"""
from __future__ import annotations

import logging
import math
import random

import numpy as np

try:                                    # optional: only the resize paths need them
    import cv2
except Exception:                       # pragma: no cover
    cv2 = None
try:
    import scipy.ndimage
except Exception:                       # pragma: no cover
    scipy = None


# --------------------------------------------------------------------------- boxes
class BoundBox:
    """Axis-aligned box with optional objectness `c` and class scores (myolo_utils.py:161-184)."""

    def __init__(self, xmin, ymin, xmax, ymax, c=None, classes=None):
        self.xmin, self.ymin, self.xmax, self.ymax = xmin, ymin, xmax, ymax
        self.c = c
        self.classes = classes
        self.label = -1
        self.score = -1

    def get_label(self):
        if self.label == -1:
            self.label = int(np.argmax(self.classes))
        return self.label

    def get_score(self):
        if self.score == -1:
            self.score = self.classes[self.get_label()]
        return self.score


def _interval_overlap(interval_a, interval_b):
    """Length of the overlap of two 1-D intervals (myolo_utils.py:231-244; may be called with
    a2 < b1, in which case the result is 0)."""
    a1, a2 = interval_a
    b1, b2 = interval_b
    if b1 < a1:
        return 0 if b2 < a1 else min(a2, b2) - a1
    return 0 if a2 < b1 else min(a2, b2) - b1


def bbox_iou(box1, box2):
    """IoU of two BoundBox objects (myolo_utils.py:187-198)."""
    iw = _interval_overlap([box1.xmin, box1.xmax], [box2.xmin, box2.xmax])
    ih = _interval_overlap([box1.ymin, box1.ymax], [box2.ymin, box2.ymax])
    inter = iw * ih
    union = (box1.xmax - box1.xmin) * (box1.ymax - box1.ymin) + (box2.xmax - box2.xmin) * (box2.ymax - box2.ymin) - inter
    return float(inter) / union


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x, axis=-1, t=-100.0):
    """The reference's softmax (myolo_utils.py:21-33): subtracts the GLOBAL max and rescales when the
    minimum falls below t."""
    x = x - np.max(x)
    if np.min(x) < t:
        x = x / np.min(x) * t
    e = np.exp(x)
    return e / e.sum(axis, keepdims=True)


def extract_bboxes(mask):
    """[H, W, N] masks -> int32 [N, (x1, y1, x2, y2)], x2/y2 one past the last set pixel; all-zero
    masks give a zero box (myolo_utils.py:247-271)."""
    n = mask.shape[-1]
    boxes = np.zeros([n, 4], dtype=np.int32)
    for i in range(n):
        m = mask[:, :, i]
        cols = np.flatnonzero(m.any(axis=0))
        rows = np.flatnonzero(m.any(axis=1))
        if cols.size:
            boxes[i] = (cols[0], rows[0], cols[-1] + 1, rows[-1] + 1)
    return boxes


# --------------------------------------------------------------------------- image / mask loading
def resize(image, output_shape, order=1, mode='constant', cval=0, clip=True, preserve_range=False,
           anti_aliasing=False, anti_aliasing_sigma=None):
    """Bilinear resize standing in for the reference's skimage wrapper (myolo_utils.py:433-454);
    skimage is not a dependency here, cv2 provides the interpolation."""
    if cv2 is None:
        raise ImportError("resizing needs cv2")
    out = cv2.resize(np.asarray(image, dtype=np.float32), (int(output_shape[1]), int(output_shape[0])),
                     interpolation=cv2.INTER_LINEAR if order == 1 else cv2.INTER_NEAREST)
    return out


def resize_image(image, net_image_shape):
    """Stretch to the network shape (aspect ratio is NOT kept, myolo_utils.py:369-390); a no-op when the
    dataset already has the network size.  Returns (image, [scale_h, scale_w])."""
    h, w = image.shape[:2]
    scale = [net_image_shape[0] / h, net_image_shape[1] / w]
    if scale != [1, 1]:
        image = resize(image, (round(h * scale[0]), round(w * scale[1])), preserve_range=True).astype(image.dtype)
    return image, scale


def resize_mask(mask, scale):
    """Nearest-neighbour zoom of [H, W, N] masks (scipy.ndimage.zoom order 0, myolo_utils.py:393-410)."""
    if list(scale) == [1, 1]:
        return mask
    return scipy.ndimage.zoom(mask, zoom=[scale[0], scale[1], 1], order=0)


def minimize_mask(bbox, mask, mini_shape):
    """Crop each instance to its box and resize to mini_shape (myolo_utils.py:413-430)."""
    mini = np.zeros(tuple(mini_shape) + (mask.shape[-1],), dtype=bool)
    for i in range(mask.shape[-1]):
        x1, y1, x2, y2 = bbox[i][:4]
        m = mask[y1:y2, x1:x2, i].astype(np.float32)
        if m.size == 0:
            raise Exception("Invalid bounding box with area of zero")
        mini[:, :, i] = np.around(resize(m, mini_shape)).astype(bool)
    return mini


def load_image_gt(dataset, config, image_id, augment=False, augmentation=None, use_mini_mask=False):
    """(image, class_ids, bbox (x1,y1,x2,y2), mask [H,W,N]) for one dataset image
    (myolo_utils.py:274-366).  Instances whose mask is empty after resizing are dropped."""
    image = dataset.load_image(image_id)
    mask, class_ids = dataset.load_mask(image_id)
    image, scale = resize_image(image, config.IMAGE_SHAPE)
    mask = resize_mask(mask, scale)
    if augment:
        logging.warning("'augment' is deprecated. Use 'augmentation' instead.")
        if random.randint(0, 1):
            image, mask = np.fliplr(image), np.fliplr(mask)
    if augmentation:
        raise NotImplementedError("imgaug augmentation is not available in this build")
    keep = mask.sum(axis=(0, 1)) > 0
    mask = mask[:, :, keep]
    class_ids = class_ids[keep]
    bbox = extract_bboxes(mask)
    if use_mini_mask:
        mask = minimize_mask(bbox, mask, config.MINI_MASK_SHAPE)
    return image, class_ids, bbox, mask


# --------------------------------------------------------------------------- YOLO target encoding
def _n_box(config):
    return len(config.ANCHORS) // 2


def _grid(config):
    s = int(config.IMAGE_SHAPE[0])
    g = s // 32
    return (g, g) if (config.GRID_H * 32 != s or config.GRID_W * 32 != s) else (config.GRID_H, config.GRID_W)


def encode_instance(config, anchors, gt_class_ids, gt_boxes, yolo_target, true_boxes):
    """Write one image's YOLO target rows (myolo_utils.py:769-820): centre and size in grid units, the
    cell that holds the centre, the anchor with the best IoU against (0,0,w,h) (first maximum wins),
    [cx, cy, w, h, 1, one-hot class] at [gy, gx, anchor] and the box into the true-box ring buffer."""
    gh, gw = yolo_target.shape[0], yolo_target.shape[1]
    cell_w = float(config.IMAGE_SHAPE[0]) / gw
    cell_h = float(config.IMAGE_SHAPE[1]) / gh
    slot = 0
    for i in range(gt_boxes.shape[0]):
        xmin, ymin, xmax, ymax = (gt_boxes[i][j] for j in range(4))
        cx = .5 * (xmin + xmax) / cell_w
        cy = .5 * (ymin + ymax) / cell_h
        gx, gy = int(np.floor(cx)), int(np.floor(cy))
        if gx < gw and gy < gh:
            w = (xmax - xmin) / cell_w
            h = (ymax - ymin) / cell_h
            probe = BoundBox(0, 0, w, h)
            best, best_iou = -1, -1
            for j, a in enumerate(anchors):
                iou = bbox_iou(probe, a)
                if best_iou < iou:
                    best, best_iou = j, iou
            yolo_target[gy, gx, best, 0:4] = (cx, cy, w, h)
            yolo_target[gy, gx, best, 4] = 1.
            yolo_target[gy, gx, best, 5 + gt_class_ids[i]] = 1
            true_boxes[0, 0, 0, slot] = (cx, cy, w, h)
            slot = (slot + 1) % true_boxes.shape[3]


class BatchGenerator(object):
    """Sequence of training batches (myolo_utils.py:689-860, a keras.utils.Sequence in the reference).

    all_info: list of [image, gt_class_ids, gt_boxes, gt_masks] as load_image_gt returns them.
    __getitem__(idx) -> (inputs, []) with inputs =
      mode 'training': [images f32, true_boxes [B,1,1,1,TB,4], yolo_target [B,GH,GW,NB,5+NC],
                        gt_class_ids i32 [B,TB], gt_boxes i32 [B,TB,4], gt_masks bool [B,H,W,MAX_GT]]
      mode 'yolo'    : the first three.
    The last batch is filled up from the preceding images (l_bound is moved back)."""

    def __init__(self, all_info, config, mode, shuffle=True, jitter=False, norm=False):
        assert mode in ['yolo', 'training']
        self.config, self.mode, self.all_info = config, mode, all_info
        self.shuffle, self.jitter, self.norm = shuffle, jitter, norm
        self.anchors = [BoundBox(0, 0, config.ANCHORS[2 * i], config.ANCHORS[2 * i + 1]) for i in range(_n_box(config))]
        if shuffle:
            np.random.shuffle(self.all_info)

    def __len__(self):
        return int(np.ceil(float(len(self.all_info)) / self.config.BATCH_SIZE))

    def num_classes(self):
        return self.config.NUM_CLASSES

    def size(self):
        return len(self.all_info)

    def load_image(self, i):
        return self.all_info[i][0]

    def __getitem__(self, idx):
        cfg = self.config
        bs = cfg.BATCH_SIZE
        lo, hi = idx * bs, (idx + 1) * bs
        if hi > len(self.all_info):
            hi = len(self.all_info)
            lo = max(0, hi - bs)
        n = hi - lo
        H, W = int(cfg.IMAGE_SHAPE[0]), int(cfg.IMAGE_SHAPE[1])
        gh, gw = _grid(cfg)
        nb, nc, tb = _n_box(cfg), cfg.NUM_CLASSES, cfg.TRUE_BOX_BUFFER
        images = np.zeros((n, H, W, 3), dtype=np.float32)
        yolo_target = np.zeros((n, gh, gw, nb, 4 + 1 + nc))
        true_boxes = np.zeros((n, 1, 1, 1, tb, 4))
        ids = np.zeros((n, tb), dtype=np.int32)
        boxes = np.zeros((n, tb, 4), dtype=np.int32)
        masks = np.zeros((n, H, W, cfg.MAX_GT_INSTANCES), dtype=bool)
        for k, (image, gt_class_ids, gt_boxes, gt_masks) in enumerate(self.all_info[lo:hi]):
            if gt_boxes.shape[0] > tb:
                print('find instances more than ' + str(tb) + ' in an image')
                pick = np.random.choice(np.arange(gt_boxes.shape[0]), tb, replace=False)
                gt_class_ids, gt_boxes, gt_masks = gt_class_ids[pick], gt_boxes[pick], gt_masks[:, :, pick]
            encode_instance(cfg, self.anchors, gt_class_ids, gt_boxes, yolo_target[k], true_boxes[k])
            images[k] = image / 255. if self.norm else image[:, :, ::-1]
            ids[k, :gt_class_ids.shape[0]] = gt_class_ids
            boxes[k, :gt_boxes.shape[0]] = gt_boxes
            m = min(gt_masks.shape[-1], cfg.MAX_GT_INSTANCES)
            masks[k, :, :, :m] = gt_masks[:, :, :m]
        if self.mode == 'yolo':
            return [images, true_boxes, yolo_target], []
        return [images, true_boxes, yolo_target, ids, boxes, masks], []

    def on_epoch_end(self):
        if self.shuffle:
            np.random.shuffle(self.all_info)


def data_generator(dataset, config, shuffle=True, augment=False, augmentation=None, batch_size=1,
                   no_augmentation_sources=None, norm=False):
    """The older python-generator variant (myolo_utils.py:457-686): loads images on the fly and yields
    ([images, true_boxes, yolo_target], []) -- the three YOLO inputs only.  A failing image is logged
    and skipped; the fifth failure is re-raised.  `no_augmentation_sources` (dataset sources that are never
    augmented) is accepted for signature parity; imgaug augmentation is not available in this build.  As in
    the reference, norm=False hands the uint8 image values over unscaled."""
    anchors = [BoundBox(0, 0, config.ANCHORS[2 * i], config.ANCHORS[2 * i + 1]) for i in range(_n_box(config))]
    image_ids = np.copy(dataset.image_ids)
    H, W = int(config.IMAGE_SHAPE[0]), int(config.IMAGE_SHAPE[1])
    gh, gw = _grid(config)
    nb, nc, tb = _n_box(config), config.NUM_CLASSES, config.TRUE_BOX_BUFFER
    index, b, errors = -1, 0, 0
    while True:
        try:
            index = (index + 1) % len(image_ids)
            if shuffle and index == 0:
                np.random.shuffle(image_ids)
            image, gt_class_ids, gt_boxes, gt_masks = load_image_gt(dataset, config, image_ids[index], augment=augment,
                                                                   augmentation=augmentation,
                                                                   use_mini_mask=config.USE_MINI_MASK)
            if not np.any(gt_class_ids > 0):
                continue
            if b == 0:
                images = np.zeros((batch_size, H, W, 3), dtype=np.float32)
                yolo_target = np.zeros((batch_size, gh, gw, nb, 4 + 1 + nc))
                true_boxes = np.zeros((batch_size, 1, 1, 1, tb, 4))
            if gt_boxes.shape[0] > tb:
                pick = np.random.choice(np.arange(gt_boxes.shape[0]), tb, replace=False)
                gt_class_ids, gt_boxes = gt_class_ids[pick], gt_boxes[pick]
            encode_instance(config, anchors, gt_class_ids, gt_boxes, yolo_target[b], true_boxes[b])
            images[b] = image / 255. if norm else image
            b += 1
            if b >= batch_size:
                yield [images, true_boxes, yolo_target], []
                b = 0
        except (GeneratorExit, KeyboardInterrupt):
            raise
        except Exception:
            logging.exception("Error processing image {}".format(dataset.image_info[image_ids[index]] if hasattr(dataset, "image_info") else index))
            errors += 1
            if errors > 5:
                raise


def encode_targets_device(config, gt_class_ids, gt_boxes, gt_masks=None):
    """Device-side version of the per-batch target construction of BatchGenerator.__getitem__ (769-820):
    padded int32 gt arrays [B,TB] / [B,TB,4] on the GPU -> (true_boxes [B,1,1,1,TB,4], yolo_target
    [B,GH,GW,NB,5+NC]) fp32 on the GPU.  With gt_boxes=None the boxes are first extracted from gt_masks
    [B,S,S,M] (extract_bboxes, 247-271)."""
    import torch
    from . import _cabi as C
    from .config import resolve
    c = resolve(config)
    st = torch.cuda.current_stream().cuda_stream
    ids = gt_class_ids.int().contiguous()
    B, M = ids.shape
    if gt_boxes is None:
        gm = gt_masks.to(torch.uint8).contiguous()
        assert gm.shape[3] == M
        gt_boxes = torch.empty(B, M, 4, dtype=torch.int32, device=ids.device)
        C.call("myolo_extract_bboxes", gm, B, c["S"], M, gt_boxes, st)
    boxes = gt_boxes.int().contiguous()
    yt = torch.empty(B, c["G"], c["G"], c["NB"], 5 + c["NC"], device=ids.device)
    tb = torch.empty(B, 1, 1, 1, c["TB"], 4, device=ids.device)
    anchors = torch.tensor(c["ANCHORS"], dtype=torch.float32, device=ids.device)
    C.call("myolo_encode_yolo_targets", ids, boxes, B, M, c["S"], c["G"], c["NB"], c["NC"], c["TB"], anchors, yt, tb, st)
    return tb, yt, boxes


def batch_slice(inputs, graph_fn, batch_size, names=None):
    """Apply graph_fn to each batch slice and stack the results (myolo_utils.py:929-963).  Kept for API
    compatibility with torch tensors / numpy arrays; the engine's target kernel is batched natively."""
    import torch
    if not isinstance(inputs, list):
        inputs = [inputs]
    outs = []
    for i in range(batch_size):
        o = graph_fn(*[x[i] for x in inputs])
        outs.append(list(o) if isinstance(o, (tuple, list)) else [o])
    res = [torch.stack([torch.as_tensor(v) for v in col], 0) for col in zip(*outs)]
    return res[0] if len(res) == 1 else res


# --------------------------------------------------------------------------- inference post-processing
def decode_one_yolo_output(netout, anchors, nb_class=None, obj_threshold=0.3, nms_threshold=0.3):
    """numpy YOLO decode + per-class NMS of ONE image's [GH,GW,NB,5+NC] output, argument order and arithmetic of
    myolo_utils.py:36-85 (pinned by tests/test_reference_golden.py).  Returns BoundBox list with normalised corner
    coordinates.  Unlike the reference the caller's array is not modified."""
    gh, gw, nb = netout.shape[:3]
    netout = np.array(netout, copy=True)
    nb_class = netout.shape[-1] - 5 if nb_class is None else nb_class
    netout[..., 4] = _sigmoid(netout[..., 4])
    netout[..., 5:] = netout[..., 4][..., np.newaxis] * _softmax(netout[..., 5:])
    netout[..., 5:] *= netout[..., 5:] > obj_threshold
    boxes = []
    for row in range(gh):
        for col in range(gw):
            for b in range(nb):
                classes = netout[row, col, b, 5:]
                if np.sum(classes) > 0:
                    x, y, w, h = netout[row, col, b, :4]
                    x = (col + _sigmoid(x)) / gw
                    y = (row + _sigmoid(y)) / gh
                    w = anchors[2 * b + 0] * np.exp(w) / gw
                    h = anchors[2 * b + 1] * np.exp(h) / gh
                    boxes.append(BoundBox(x - w / 2, y - h / 2, x + w / 2, y + h / 2, netout[row, col, b, 4], classes))
    for c in range(nb_class):
        order = list(reversed(np.argsort([bx.classes[c] for bx in boxes])))
        for i, bi in enumerate(order):
            if boxes[bi].classes[c] == 0:
                continue
            for bj in order[i + 1:]:
                if bbox_iou(boxes[bi], boxes[bj]) >= nms_threshold:
                    boxes[bj].classes[c] = 0
    return [bx for bx in boxes if bx.get_score() > obj_threshold]


def bbox_iou_2(box1, box2, image_shape=None):
    """IoU of two normalised (x1, y1, x2, y2) sequences, evaluated in pixels of `image_shape` exactly like
    myolo_utils.py:201-228 (image_shape None: coordinates taken as they are)."""
    w, h = (image_shape[0], image_shape[1]) if image_shape is not None else (1, 1)
    b1 = (box1[0] * w, box1[1] * h, box1[2] * w, box1[3] * h)
    b2 = (box2[0] * w, box2[1] * h, box2[2] * w, box2[3] * h)
    iw = _interval_overlap([b1[0], b1[2]], [b2[0], b2[2]])
    ih = _interval_overlap([b1[1], b1[3]], [b2[1], b2[3]])
    inter = iw * ih
    w1, h1 = b1[2] - b1[0], b1[3] - b1[1]
    w2, h2 = b2[2] - b2[0], b2[3] - b2[1]
    union = w1 * h1 + w2 * h2 - inter
    return float(inter) / union


def NMB(boxes, class_ids, indices, image_shape, nms_threshold=0.3):
    """"Suppress non-maximal boxes" exactly as myolo_utils.py:88-113: `boxes` / `class_ids` are the candidates in
    descending score order, `indices` their detection indices; candidate j is dropped when ANY earlier candidate i
    of the same class has IoU >= nms_threshold with it -- including an i that was dropped itself (this is not
    greedy NMS; pinned by tests/test_reference_golden.py).  Returns the surviving entries of `indices`."""
    remove = []
    for i in range(len(indices)):
        for j in range(i + 1, len(indices)):
            if bbox_iou_2(boxes[i], boxes[j], image_shape) >= nms_threshold and class_ids[i] == class_ids[j]:
                remove.append(j)
    return np.delete(indices, remove)


def unmold_mask(mask, bbox, image_shape):
    """A small soft mask -> full-size boolean mask (myolo_utils.py:883-912).  `bbox` = (x1, y1, x2, y2) NORMALISED, as
    DetectionsLayer emits it; the reference's integer rules are kept: corners truncated with int(), x1/y1 clamped to
    [0, size], x2/y2 to [1, size], the mask resized (bilinear) to the CLIPPED box and thresholded at 0.5.  (The reference
    resizes with scikit-image, which is not available here: cv2's bilinear resize stands in for it.  A box that is empty
    after truncation gives an empty mask; the reference fails with a broadcasting error there.)"""
    threshold = 0.5
    w, h = image_shape[0], image_shape[1]
    x1, y1, x2, y2 = bbox
    x1 = min(max(0, int(x1 * w)), w)
    x2 = min(max(1, int(x2 * w)), w)
    y1 = min(max(0, int(y1 * h)), h)
    y2 = min(max(1, int(y2 * h)), h)
    full_mask = np.zeros(image_shape[:2], dtype=bool)
    if x2 <= x1 or y2 <= y1:
        return full_mask
    m = resize(mask, (max(1, y2 - y1), max(1, x2 - x1)))
    full_mask[y1:y2, x1:x2] = np.where(m >= threshold, 1, 0).astype(bool)
    return full_mask


def box_refinement_graph(box, gt_box):
    """Refinement (dy, dx, log dh, log dw) that maps `box` onto `gt_box` (myolo_utils.py:116-139); the reference's
    coordinate naming is kept: columns 0/2 span the "width", 1/3 the "height".  Works on torch tensors [N, 4]."""
    import torch
    box, gt_box = box.to(torch.float32), gt_box.to(torch.float32)
    width, height = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1]
    center_x, center_y = box[:, 0] + 0.5 * width, box[:, 1] + 0.5 * height
    gt_width, gt_height = gt_box[:, 2] - gt_box[:, 0], gt_box[:, 3] - gt_box[:, 1]
    gt_center_x, gt_center_y = gt_box[:, 0] + 0.5 * gt_width, gt_box[:, 1] + 0.5 * gt_height
    return torch.stack([(gt_center_y - center_y) / height, (gt_center_x - center_x) / width,
                        torch.log(gt_height / height), torch.log(gt_width / width)], dim=1)


def compute_backbone_shapes(config, image_shape):
    """[height, width] of the backbone's feature map (myolo_utils.py:142-150).  The reference divides by the LIST
    config.BACKBONE_STRIDES (a TypeError at HEAD); the single stride it holds is used here."""
    assert config.BACKBONE in ["mobilenet"]
    stride = config.BACKBONE_STRIDES[0] if isinstance(config.BACKBONE_STRIDES, (list, tuple)) else config.BACKBONE_STRIDES
    return np.array([int(math.ceil(image_shape[0] / stride)), int(math.ceil(image_shape[1] / stride))])


def mold_image(images, config):
    """RGB image(s) -> float32 minus config.MEAN_PIXEL (myolo_utils.py:153-158; the base Config defines no MEAN_PIXEL,
    so -- exactly as in the reference -- the caller's config has to)."""
    return images.astype(np.float32) - config.MEAN_PIXEL


def draw_boxes(image, boxes, labels):
    """Draw BoundBox rectangles and 'label score' captions into `image` in place (myolo_utils.py:863-880)."""
    if cv2 is None:
        raise ImportError("draw_boxes needs cv2")
    image_h, image_w, _ = image.shape
    for box in boxes:
        xmin, ymin = int(box.xmin * image_w), int(box.ymin * image_h)
        xmax, ymax = int(box.xmax * image_w), int(box.ymax * image_h)
        cv2.rectangle(image, (xmin, ymin), (xmax, ymax), (0, 255, 0), 1)
        cv2.putText(image, labels[box.get_label()] + ' ' + str(box.get_score()), (xmin, ymax - 13),
                    cv2.FONT_HERSHEY_SIMPLEX, 1.5e-3 * image_h, (0, 255, 0), 1)
    return image


def resize_one_image(image, gt_box, gt_mask, new_shape):
    """myolo_utils.py:915-925, kept as it is at HEAD: rescales and clips `gt_box` IN PLACE (index 2 is computed from
    gt_box[1], as in the reference) and returns None; the resized image is discarded."""
    if cv2 is None:
        raise ImportError("resize_one_image needs cv2")
    original_w, original_h = image.shape[0], image.shape[1]
    new_w, new_h = new_shape[0], new_shape[1]
    cv2.resize(image, (new_w, new_h))
    gt_box[0], gt_box[2] = int(gt_box[0] * float(new_w) / original_w), int(gt_box[1] * float(new_w) / original_w)
    gt_box[1], gt_box[3] = int(gt_box[1] * float(new_h) / original_h), int(gt_box[3] * float(new_h) / original_h)
    gt_box[0], gt_box[2] = max(min(gt_box[0], new_w), 0), max(min(gt_box[2], new_w), 0)
    gt_box[1], gt_box[3] = max(min(gt_box[1], new_h), 0), max(min(gt_box[3], new_h), 0)
