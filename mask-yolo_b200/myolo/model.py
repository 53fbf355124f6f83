"""Mask-YOLO model API on the B200 engine.

Same public surface as the reference's myolo/model.py -- `MaskYOLO(mode, config, model_dir,
yolo_pretrain_dir, yolo_trainable)` with `.keras_model`, `.train`, `.compile`, `.set_trainable`,
`.load_weights`, `.infer_yolo`, `.detect`, `.decode_masks`, and the module-level layer/graph names
(`PyramidROIAlign`, `DetectMaskTargetLayer`, `DecodeYOLOLayer`, `DetectionsLayer`,
`yolo_custom_loss`, `myolo_mask_loss_graph`, ...) -- but nothing here builds a Keras graph: every
name is an operator over device tensors that calls the hand-written sm_100a kernels through the C
ABI (include/myolo_b200.h), and `MaskYOLO` drives `myolo.engine.Engine`, which owns all HBM buffers.
There is no CPU path: constructing a model without an sm_100 GPU raises.
"""
from __future__ import annotations

import datetime
import math
import os
import re
import time
from typing import List, Optional

import numpy as np
import torch

from . import _cabi as C
from . import checkpoint
from . import myolo_utils as mutils
from .config import Config, resolve
from .engine import Engine, init_params, param_specs, MASK_C


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


# ------------------------------------------------------------------------------------------------
# operators (reference: the Keras layers / graph functions of myolo/model.py)
# ------------------------------------------------------------------------------------------------
def relu6(x):
    """model.py:38-39.  Elementwise utility (the engine fuses ReLU6 into its BN-apply kernel)."""
    return torch.clamp(x, 0.0, 6.0)


def norm_boxes_graph(boxes, shape):
    """Pixel (x1,y1,x2,y2) -> normalised: (box - [0,0,1,1]) / ([s0,s1,s0,s1] - 1)  (model.py:1394-1408)."""
    s0, s1 = float(shape[0]), float(shape[1])
    scale = torch.tensor([s0, s1, s0, s1], dtype=boxes.dtype, device=boxes.device) - 1.0
    shift = torch.tensor([0.0, 0.0, 1.0, 1.0], dtype=boxes.dtype, device=boxes.device)
    return (boxes - shift) / scale


def trim_zeros_graph(boxes, name='trim_zeros'):
    """Drop all-zero rows; returns (boxes, keep mask)  (model.py:1411-1420)."""
    keep = boxes.abs().sum(1) != 0
    return boxes[keep], keep


def overlaps_graph(boxes1, boxes2):
    """IoU matrix [N1, N2] of (x1,y1,x2,y2) boxes  (model.py:420-454)."""
    a, b = boxes1[:, None, :], boxes2[None, :, :]
    iw = (torch.minimum(a[..., 2], b[..., 2]) - torch.maximum(a[..., 0], b[..., 0])).clamp(min=0)
    ih = (torch.minimum(a[..., 3], b[..., 3]) - torch.maximum(a[..., 1], b[..., 1])).clamp(min=0)
    inter = iw * ih
    union = (a[..., 3] - a[..., 1]) * (a[..., 2] - a[..., 0]) + (b[..., 3] - b[..., 1]) * (b[..., 2] - b[..., 0]) - inter
    return inter / union


class _Layer(object):
    def __call__(self, inputs):
        return self.call(inputs)


class DecodeYOLOLayer(_Layer):
    """yolo_output [B,G,G,NB,5+NC] -> proposals [B, G*G*NB, (x1,y1,x2,y2)] normalised  (model.py:1429-1476)."""

    def __init__(self, config, **kwargs):
        self.name, self.config, self.cfg = kwargs.get("name"), config, resolve(config)

    def call(self, inputs):
        y = _f32(inputs[0] if isinstance(inputs, (list, tuple)) else inputs)
        B, G, NB, NC = y.shape[0], self.cfg["G"], self.cfg["NB"], self.cfg["NC"]
        out = torch.empty(B, G * G * NB, 4, device=y.device)
        anchors = torch.tensor(self.cfg["ANCHORS"], dtype=torch.float32, device=y.device)
        C.call("myolo_yolo_decode", y, anchors, out, None, B, G, G, NB, NC, _stream())
        return out

    def compute_output_shape(self, input_shape):
        return (None, self.cfg["R"], 4)


class DetectionsLayer(_Layer):
    """yolo_output -> [B, R, (x1,y1,x2,y2, sigmoid(conf), argmax class)]  (model.py:1479-1541)."""

    def __init__(self, config, **kwargs):
        self.name, self.config, self.cfg = kwargs.get("name"), config, resolve(config)

    def call(self, inputs):
        y = _f32(inputs[0] if isinstance(inputs, (list, tuple)) else inputs)
        B, G, NB, NC = y.shape[0], self.cfg["G"], self.cfg["NB"], self.cfg["NC"]
        det = torch.empty(B, G * G * NB, 6, device=y.device)
        anchors = torch.tensor(self.cfg["ANCHORS"], dtype=torch.float32, device=y.device)
        C.call("myolo_yolo_decode", y, anchors, None, det, B, G, G, NB, NC, _stream())
        return det

    def compute_output_shape(self, input_shape):
        return (None, self.cfg["R"], 6)


class PyramidROIAlign(_Layer):
    """Single-level ROIAlign = crop_and_resize(feature_map, boxes, box->image index, pool_shape)
    (model.py:299-413).  boxes [B,R,4] are handed to the sampler as they are, i.e. (x1,y1,x2,y2) read
    as (y1,x1,y2,x2) -- the reference's behaviour.  Returns [B,R,P,P,C]."""

    def __init__(self, pool_shape, **kwargs):
        self.name = kwargs.get("name")
        self.pool_shape = tuple(pool_shape)
        assert self.pool_shape[0] == self.pool_shape[1]

    def call(self, inputs):
        boxes, feat = _f32(inputs[0]), _f32(inputs[1])
        B, R = boxes.shape[0], boxes.shape[1]
        Fh, Fw, Cc = feat.shape[1], feat.shape[2], feat.shape[3]
        P = self.pool_shape[0]
        out = torch.empty(B * R, P, P, Cc, device=feat.device)
        C.call("myolo_roialign_fwd", C.view(feat, B, Fh, Fw, Cc), boxes, B * R, R, P, C.view(out, B * R, P, P, Cc), 0,
               _stream())
        return out.view(B, R, P, P, Cc)

    def compute_output_shape(self, input_shape):
        return input_shape[0][:2] + self.pool_shape + (input_shape[1][-1],)


class DetectMaskTargetLayer(_Layer):
    """proposals + ground truth -> (rois, target_class_ids, None, target_mask)  (model.py:605-661,
    detect_mask_target_graph 457-602).  gt_boxes are PIXEL (x1,y1,x2,y2) as the model input carries
    them: norm_boxes_graph (model.py:819-820) is fused into the kernel."""

    def __init__(self, config, **kwargs):
        self.name, self.config, self.cfg = kwargs.get("name"), config, resolve(config)

    def compute_output_shape(self, input_shape):
        return [(None, input_shape[0][1], 4), (None, input_shape[0][1]), (None, None),
                (None, input_shape[0][1], self.cfg["MASK_SHAPE"][0], self.cfg["MASK_SHAPE"][1])]

    def compute_mask(self, inputs, mask=None):
        return [None, None, None, None]

    def call(self, inputs):
        props, ids, boxes, masks = inputs
        props, boxes = _f32(props), _f32(boxes)
        ids = ids.int().contiguous()
        masks = masks.to(torch.uint8).contiguous()
        B, R = props.shape[0], props.shape[1]
        mh, mw = self.cfg["MASK_SHAPE"]
        dev = props.device
        rois = torch.empty(B, R, 4, device=dev)
        tids = torch.empty(B, R, dtype=torch.int32, device=dev)
        tmask = torch.empty(B, R, mh, mw, device=dev)
        scratch = [torch.empty(B, dtype=torch.int32, device=dev)] + [torch.empty(B, R, dtype=torch.int32, device=dev) for _ in range(2)]
        C.call("myolo_detect_mask_targets", props, ids, boxes, masks, B, R, ids.shape[1], masks.shape[3], self.cfg["S"],
               mh, mw, rois, tids, tmask, scratch[0], scratch[1], scratch[2], _stream())
        return [rois, tids, None, tmask]


def log2_graph(x):
    """model.py:299-301"""
    return torch.log(x) / math.log(2.0)


def detect_mask_target_graph(yolo_proposals, gt_class_ids, gt_boxes, gt_masks, config):
    """One image (model.py:457-602): thin wrapper over the batched kernel."""
    r = DetectMaskTargetLayer(config).call([yolo_proposals[None], gt_class_ids[None], gt_boxes[None], gt_masks[None]])
    return r[0][0], r[1][0], None, r[3][0]


def yolo_custom_loss(y_true, y_pred, true_boxes, config=None):
    """YOLOv2 loss of model.py:86-242 (xy + wh + conf + class), scalar device tensor."""
    cfg = resolve(config if config is not None else Config())
    y_true, y_pred, tb = _f32(y_true), _f32(y_pred), _f32(true_boxes)
    B, G, NB, NC, TB = y_pred.shape[0], cfg["G"], cfg["NB"], cfg["NC"], cfg["TB"]
    dev = y_pred.device
    out = torch.empty(5, device=dev)
    ws = torch.zeros(8, dtype=torch.float64, device=dev)
    sc = C.float_array([cfg["OBJECT_SCALE"], cfg["NO_OBJECT_SCALE"], cfg["COORD_SCALE"], cfg["CLASS_SCALE"]])
    C.call("myolo_yolo_loss", y_true, y_pred, tb, torch.tensor(cfg["ANCHORS"], dtype=torch.float32, device=dev),
           torch.tensor(cfg["CLASS_WEIGHTS"], device=dev), B, G, G, NB, NC, TB, sc, 0, 1.0, out, None, ws, _stream())
    return out[0]


def myolo_mask_loss_graph(target_masks, target_class_ids, pred_masks):
    """Binary cross-entropy over the positive ROIs' class-specific masks (model.py:718-754)."""
    pm = _f32(pred_masks)
    n = pm.shape[0] * pm.shape[1]
    mh, mw, nc = pm.shape[2], pm.shape[3], pm.shape[4]
    out = torch.empty(1, device=pm.device)
    ws = torch.zeros(2, dtype=torch.float64, device=pm.device)
    C.call("myolo_mask_loss", pm, _f32(target_masks), target_class_ids.int().contiguous(), n, mh, mw, nc, 1.0, out, None,
           ws, _stream())
    return out[0]


# The reference's graph builders add layers to Keras' implicit default graph.  Here the weights live in an Engine, and the
# "implicit graph" is the engine of the most recently built MaskYOLO (or the one passed as `engine=`): the builders keep the
# reference's positional signatures and run the corresponding part of that engine on device tensors.
_CURRENT = {"engine": None}


def _engine(engine=None):
    e = engine if engine is not None else _CURRENT["engine"]
    if e is None:
        raise RuntimeError("no engine: build a MaskYOLO first (its weights are what these graph functions run on) or pass engine=")
    return e


def conv_block(inputs, filters, alpha=1.0, kernel=(3, 3), strides=(1, 1), *, engine=None):
    """model.py:42-52: ZeroPad + 3x3 conv + BN + ReLU6 (inference-mode BN).  The engine holds exactly the stem the
    reference builds with it (32 filters, stride 2, model.py:66); other arguments are refused."""
    if int(filters * alpha) != 32 or tuple(kernel) != (3, 3) or tuple(strides) != (2, 2):
        raise NotImplementedError("the engine's stem is conv_block(x, 32, strides=(2, 2)) as in mobilenet_graph (model.py:66)")
    e = _engine(engine)
    e.forward(inputs, training=False)
    return e.A["a0"]


def mobilenet_graph(input_image, architecture, stage5=False, alpha=1.0, depth_multiplier=1, *, engine=None, training=False):
    """Truncated MobileNet-v1 backbone -> C4 [B,S/8,S/8,512]  (model.py:55-79)."""
    assert architecture == 'mobilenet'
    e = _engine(engine)
    e.forward(input_image, training=training)
    return e.c4.dense() if e.with_mask and not e.x3 else e.A["ap6"]


def yolo_branch_graph(x, config=None, alpha=1.0, depth_multiplier=1, *, engine=None, training=False):
    """YOLO branch -> [B,G,G,NB,5+NC]  (model.py:249-278).  `x` is either the image (the whole backbone + branch pass is
    run) or the C4 tensor mobilenet_graph just returned for this engine (the branch output of that same pass is returned:
    the engine evaluates backbone and branch together)."""
    e = _engine(engine)
    if x.shape[-1] == 3:
        return e.forward(x, training=training)
    if x.shape[-1] != 512 or "yolo" not in e.A:
        raise ValueError("x must be the image or the C4 feature map of the engine's last mobilenet_graph pass")
    return e.A["yolo"].view(e.B, e.cfg["G"], e.cfg["G"], e.NB, 5 + e.NC)


def build_yolo_model(config, depth=None, batch=None):
    """The nested 'yolo_model' of the reference (model.py:281-292) as an engine in mode 'yolo'."""
    return Engine(resolve(config), batch or config.BATCH_SIZE, "yolo")


def build_mask_graph(rois, feature_maps=None, pool_size=None, num_classes=None, train_bn=False, *, engine=None, training=False):
    """Mask head on [B,R,4] rois -> [B,R,28,28,NC] sigmoid masks  (model.py:668-715).  The feature map is the one the
    engine's preceding backbone pass left behind (`feature_maps` is accepted for signature parity); bn2-4 always use
    their moving statistics, as in the reference, whatever `train_bn` says at HEAD (it is passed as training=False)."""
    e = _engine(engine)
    if pool_size is not None and int(pool_size) != int(e.cfg["POOL"]):
        raise ValueError("pool_size %r differs from the engine's MASK_POOL_SIZE %r" % (pool_size, e.cfg["POOL"]))
    if num_classes is not None and int(num_classes) != int(e.NC):
        raise ValueError("num_classes %r differs from the engine's %r" % (num_classes, e.NC))
    return e.mask_head(_f32(rois), training)


class Prefetcher(object):
    """Builds `sequence[i]` for i in `indices` on background threads and yields the results IN ORDER, never more than
    `depth` of them ahead of the consumer -- what Keras' OrderedEnqueuer does for fit_generator(max_queue_size=depth).
    numpy releases the GIL in the large copies that dominate BatchGenerator.__getitem__, so the batch for step k+1 is
    assembled while the GPU runs step k.  An exception in a worker is re-raised at the position of its item."""

    def __init__(self, sequence, indices, depth=3, workers=1):
        self.sequence, self.indices = sequence, list(indices)
        self.depth, self.workers = max(1, int(depth)), max(1, int(workers))

    def __iter__(self):
        import collections
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=min(self.workers, self.depth), thread_name_prefix="myolo-prefetch")
        pending = collections.deque()
        todo = iter(self.indices)
        try:
            for i in todo:
                pending.append(pool.submit(self.sequence.__getitem__, i))
                if len(pending) >= self.depth:
                    break
            while pending:
                item = pending.popleft().result()
                for i in todo:                        # keep the window full: one new item per consumed item
                    pending.append(pool.submit(self.sequence.__getitem__, i))
                    break
                yield item
        finally:
            for f in pending:
                f.cancel()
            pool.shutdown(wait=True)


class DetectResults(list):
    """What MaskYOLO.detect returns.  The reference's code returns `[{"bboxes", "class_ids", "confidence_scores",
    "full_masks"}]` (model.py:1316-1328) while its docstring documents the keys "rois", "class_ids", "scores", "masks":
    `results[0]["bboxes"]` and `results["rois"]` both work here."""
    _ALIASES = {"rois": "bboxes", "scores": "confidence_scores", "masks": "full_masks"}

    def __init__(self, boxes, class_ids, scores, masks):
        list.__init__(self, [{"bboxes": boxes, "class_ids": class_ids, "confidence_scores": scores, "full_masks": masks}])

    def __getitem__(self, key):
        if isinstance(key, str):
            return list.__getitem__(self, 0)[self._ALIASES.get(key, key)]
        return list.__getitem__(self, key)

    def keys(self):
        return list(list.__getitem__(self, 0)) + list(self._ALIASES)


# ------------------------------------------------------------------------------------------------
# keras_model-like handle
# ------------------------------------------------------------------------------------------------
class _ModelHandle(object):
    """What `MaskYOLO.keras_model` exposes: predict / train_on_batch / test_on_batch / summary /
    get_weights-by-name, over the engine."""

    def __init__(self, owner: "MaskYOLO"):
        self._o = owner
        self.metrics_names = ["loss", "yolo_sum_loss", "myolo_mask_loss"] if owner.mode == "training" else ["loss", "yolo_sum_loss"]
        self.name = {"training": "mask+yolo", "yolo": "only_yolo", "inference": "mask_yolo_inference"}[owner.mode]

    def predict(self, inputs, batch_size=None, verbose=0):
        return self._o._predict(inputs)

    def train_on_batch(self, inputs, targets=None):
        return self._o._train_on_batch(inputs)

    def test_on_batch(self, inputs, targets=None):
        return self._o._train_on_batch(inputs, update=False)

    def summary(self):
        lines = ["Model: %s" % self.name]
        total = trainable = 0
        for name, shape, tr in self._o.engine.specs:
            n = int(np.prod(shape))
            total += n
            trainable += n if tr else 0
            lines.append("  %-40s %-20s %10d%s" % (name, tuple(shape), n, "" if tr else "  (non-trainable)"))
        lines.append("Total params: %d  Trainable: %d  Non-trainable: %d" % (total, trainable, total - trainable))
        return "\n".join(lines)

    def get_weights(self):
        return self._o.engine.state_dict()

    def save_weights(self, path):
        checkpoint.write_checkpoint(path, self._o.engine.state_dict())

    def load_weights(self, path, by_name=False):
        self._o.load_weights(path, by_name=by_name)


# ------------------------------------------------------------------------------------------------
# MaskYOLO
# ------------------------------------------------------------------------------------------------
class MaskYOLO:
    """MobileNet backbone + YOLOv2 head + Mask R-CNN-style mask branch (model.py:761-1391)."""

    def __init__(self, mode, config, model_dir=None, yolo_pretrain_dir=None, yolo_trainable=True,
                 precision: Optional[str] = None, device: int = 0, seed: int = 0):
        assert mode in ['training', 'inference', 'yolo']
        self.mode = mode
        self.config = config
        self.model_dir = model_dir
        self.yolo_pretrain_dir = yolo_pretrain_dir
        self.yolo_trainable = yolo_trainable
        self.precision = precision or os.environ.get("MYOLO_PRECISION", "h16")
        self.device, self.seed = device, seed
        self.learning_rate = getattr(config, "LEARNING_RATE", 1e-3)
        self.allreduce = None                    # set by myolo.ddp.attach() for data-parallel training
        self.keras_model = self.build(mode=mode, config=config)
        self.epoch = 0

    # ---- construction
    def build(self, mode, config):
        assert mode in ['training', 'inference', 'yolo']
        w, h = config.IMAGE_SHAPE[:2]
        if w % 32 != 0 or h % 32 != 0:
            raise Exception("Image size must be dividable by 32 to adapt with YOLO framework. "
                            "For example, use 224, 256, 288, 320, 356, ... etc. ")
        self.cfg = resolve(config)
        batch = int(config.BATCH_SIZE) if mode != "inference" else int(getattr(config, "BATCH_SIZE", 1))
        self.engine = Engine(self.cfg, batch, mode, self.precision, self.device, seed=self.seed)
        _CURRENT["engine"] = self.engine                  # what the module-level graph functions run on by default
        self._stage_bufs, self._stage_shapes = None, None
        self._slot_free = [None] * self._STAGE_SLOTS
        self._last_slot = None
        if self.yolo_pretrain_dir is not None:
            self.load_weights(self.yolo_pretrain_dir, by_name=True)
            if not self.yolo_trainable:
                # model.py:854-868 sets trainable=False on every layer of the 'whole_yolo_branch' model: the backbone
                # layers one by one, and the nested 'yolo_model' (conv_dw_7..14, conv_pw_7..14, conv_23) as a CONTAINER.
                # set_trainable() later re-opens layers by regex but never touches the container's own flag, so the
                # nested model stays frozen for good (base=True), the backbone only until the next set_trainable().
                nested = re.compile(r"(conv_(dw|pw)_(7|8|9|1[0-4])(_bn)?|conv_23)/.*")
                self.engine.set_trainable(lambda n: not nested.fullmatch(n), base=True)
                self.engine.set_trainable(lambda n: n.startswith(("feature_map", "myolo_mask")))
        return _ModelHandle(self)

    # ---- host <-> device staging (pinned buffers, one async copy per input)
    _WANT = [torch.float32, torch.float32, torch.float32, torch.int32, torch.float32, torch.uint8]
    _STAGE_SLOTS = 2

    def _stage(self, inputs: List[np.ndarray]):
        """numpy batch (what BatchGenerator yields) -> device tensors.  Each input is converted into a page-locked staging
        buffer (numpy copy: float64 -> float32 for the two YOLO tensors, bool -> bytes for the masks) and uploaded with one
        asynchronous copy on a side stream.  Two slots of staging + device buffers alternate, so the upload of step k+1 can
        run while step k still reads its inputs (fit_batches); the recorded launch plan of the engine is re-pointed at the
        slot's addresses on replay."""
        want = self._WANT
        if all(isinstance(x, torch.Tensor) and x.is_cuda for x in inputs):
            # batch already resident in HBM (myolo.shapes.DeviceShapes): nothing to stage
            for x, dt in zip(inputs, want):
                if x.dtype != dt or not x.is_contiguous():
                    raise TypeError("device inputs must be contiguous %s tensors, got %s" % (dt, x.dtype))
            self.engine.inputs_ready = None
            self.last_h2d_bytes = 0
            return list(inputs)
        shapes = [tuple(np.shape(x)) for x in inputs]
        if self._stage_bufs is None or self._stage_shapes != shapes:
            self._stage_bufs, self._stage_shapes, self._stage_k = [], shapes, 0
            for _ in range(self._STAGE_SLOTS):
                slot = []
                for shp, dt in zip(shapes, want):
                    host = torch.empty(shp, dtype=dt).pin_memory()
                    slot.append((host, host.numpy(), torch.empty_like(host, device=self.engine.dev)))
                self._stage_bufs.append([slot, None])           # [buffers, event: last upload from this slot has completed]
        # copies run on a side stream in consumption order: the step starts as soon as the IMAGE has
        # landed; the ground-truth tensors (two thirds of the bytes) arrive under the backbone forward and
        # are awaited right before the first kernel that reads them (Engine.inputs_ready)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.engine.dev)
        entry = self._stage_bufs[self._stage_k % self._STAGE_SLOTS]
        self._stage_k += 1
        slot = entry[0]
        if entry[1] is not None:
            entry[1].synchronize()                        # the upload that last read this pinned slot has left it
        main = torch.cuda.current_stream(self.engine.dev)
        if self._slot_free[(self._stage_k - 1) % self._STAGE_SLOTS] is not None:
            # the step that last consumed this slot's device tensors has finished reading them
            self._copy_stream.wait_event(self._slot_free[(self._stage_k - 1) % self._STAGE_SLOTS])
        out, nbytes = [], 0
        with torch.cuda.stream(self._copy_stream):
            for k, ((host, host_np, devt), x) in enumerate(zip(slot, inputs)):
                if isinstance(x, torch.Tensor) and x.is_pinned() and x.dtype == host.dtype and x.is_contiguous():
                    host = x                              # already pinned and typed (pin_inputs): no host-side copy
                elif isinstance(x, torch.Tensor):
                    host.copy_(x)
                else:
                    x = np.asarray(x)
                    np.copyto(host_np, x.view(np.uint8) if x.dtype == np.bool_ else x, casting="unsafe")
                devt.copy_(host, non_blocking=True)
                if k == 0:
                    ev_img = torch.cuda.Event()
                    ev_img.record(self._copy_stream)
                nbytes += host.numel() * host.element_size()
                out.append(devt)
            ev_all = torch.cuda.Event()
            ev_all.record(self._copy_stream)
        entry[1] = ev_all
        main.wait_event(ev_img)
        self.engine.inputs_ready = ev_all
        self.last_h2d_bytes = nbytes
        self._last_slot = (self._stage_k - 1) % self._STAGE_SLOTS
        return out

    def _mark_slot_consumed(self):
        """Record, on the compute stream, that the step just enqueued is the last reader of its staging slot."""
        k = getattr(self, "_last_slot", None)
        if k is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.engine.dev))
            self._slot_free[k] = ev
            self._last_slot = None

    @staticmethod
    def pin_inputs(inputs):
        """BatchGenerator output (numpy) -> page-locked torch tensors in the dtypes the engine consumes; a data
        loader that produces these lets train_on_batch skip its staging copy."""
        want = [torch.float32, torch.float32, torch.float32, torch.int32, torch.float32, torch.uint8]
        out = []
        for x, dt in zip(inputs, want):
            t = torch.empty(tuple(np.shape(x)), dtype=dt).pin_memory()
            t.copy_(torch.from_numpy(np.ascontiguousarray(x)))
            out.append(t)
        return out

    def _enqueue_step(self, inputs, update=True, lr=None):
        """Stage one batch and enqueue its step; returns the device-side outputs (nothing is read back yet)."""
        dev_inputs = self._stage(inputs)
        eng = self.engine
        if update:
            out = eng.train_step(dev_inputs, lr if lr is not None else self.learning_rate, self.allreduce)
        else:
            out = eng.forward_training(dev_inputs, learning_phase=False)     # Keras validates with learning phase 0
        self._mark_slot_consumed()
        res = [out["yolo_sum_loss"]] + ([out["mask_loss"]] if "mask_loss" in out else [])
        if getattr(self, "_loss_ring", None) is None or self._loss_ring[0].numel() != len(res):
            self._loss_ring = [torch.empty(len(res), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._loss_k = 0
        host = self._loss_ring[self._loss_k % len(self._loss_ring)]
        self._loss_k += 1
        host.copy_(torch.stack(res), non_blocking=True)      # device -> host read of the step's losses (pinned, async)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(eng.dev))
        self.last_d2h_bytes = 4 * len(res)
        self.last_outputs = out
        return host, done

    def _read_step(self, pending):
        host, done = pending
        done.synchronize()
        vals = host.tolist()
        lw = self.cfg["LOSS_WEIGHTS"]
        total = vals[0] * lw.get("yolo_sum_loss", 1.0) + (vals[1] * lw.get("myolo_mask_loss", 1.0) if len(vals) > 1 else 0.0)
        return [total] + vals

    def _train_on_batch(self, inputs, update=True, lr=None):
        return self._read_step(self._enqueue_step(inputs, update, lr))

    def fit_batches(self, batches, update=True, lr=None):
        """The fit loop's inner pipeline (what Keras' fit_generator does around train_on_batch, model.py:1047-1059):
        consumes an iterable of BatchGenerator outputs and yields [loss, yolo_sum_loss, myolo_mask_loss] per batch, in
        order.  The losses of step k are read after step k+1 has been staged and enqueued, so the host-side conversion
        and upload of the next batch overlap the device's work on the current one; every batch is still copied from
        host memory and every step's losses are still read back."""
        pending = None
        for inputs in batches:
            nxt = self._enqueue_step(inputs, update, lr)
            if pending is not None:
                yield self._read_step(pending)
            pending = nxt
        if pending is not None:
            yield self._read_step(pending)

    def _predict(self, inputs):
        image = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        img = torch.as_tensor(np.ascontiguousarray(image), dtype=torch.float32).pin_memory().to(self.engine.dev, non_blocking=True)
        eng = self.engine
        if self.mode == "inference":
            yolo, det, masks = eng.forward_inference(img)
            return [yolo.cpu().numpy(), det.cpu().numpy(), masks.cpu().numpy()]
        yolo = eng.forward(img, training=False)
        return [yolo.cpu().numpy()]

    # ---- training API
    def compile(self, learning_rate, momentum=None):
        """Adam(lr, beta_1 0.9, beta_2 0.999, epsilon 1e-8) over LOSS_WEIGHTS-weighted losses
        (model.py:1062-1118).  `momentum` is accepted and unused, as in the reference."""
        self.learning_rate = float(learning_rate)
        self.engine.reset_optimizer()             # a new keras.optimizers.Adam per compile() (model.py:1071-1075)

    def set_trainable(self, layer_regex, keras_model=None, indent=0, verbose=1):
        """Train only the layers whose Keras name fully matches layer_regex (model.py:1120-1155)."""
        rx = re.compile(layer_regex)
        self.engine.set_trainable(lambda name: bool(rx.fullmatch(name.split("/")[0])))

    def load_weights(self, filepath, by_name=False, exclude=None):
        """Load a checkpoint keyed by the Keras variable names: a Keras 2.x .h5 weight file (myolo.h5lite), a torch.save dict
        (what train() writes), .npz or .safetensors (myolo.checkpoint).  by_name tolerates missing variables; `exclude` drops
        layers by name, the nested 'yolo_model' included (model.py:1157-1196 semantics)."""
        sd = checkpoint.read_checkpoint(filepath, exclude)
        self.engine.load_params(sd, strict=not (by_name or exclude))

    def train(self, train_dataset, val_dataset, learning_rate, epochs, layers, augmentation=None, custom_callbacks=None,
              no_augmentation_sources=None, max_cached=(50, 6), verbose=1, max_queue_size=3, workers=2):
        """fit loop of model.py:943-1060: caches the first 50 / 6 images of the datasets, builds the
        BatchGenerators, trains `layers` with Adam and writes './saved_model_<Mon DD-HH-MM>.pt' after
        every epoch.  Batches are built ahead of the GPU by background threads, in order, at most `max_queue_size` of
        them (the reference's fit_generator(max_queue_size=3) enqueuer, model.py:1047-1058).  Returns the per-epoch
        history dict."""
        layer_regex = {"all": ".*"}
        layers = layer_regex.get(layers, layers)
        cfg = self.config
        n_tr = min(max_cached[0], len(train_dataset.image_ids))
        n_va = min(max_cached[1], len(val_dataset.image_ids)) if val_dataset is not None else 0
        train_info = [list(mutils.load_image_gt(train_dataset, cfg, i, use_mini_mask=cfg.USE_MINI_MASK)) for i in range(n_tr)]
        val_info = [list(mutils.load_image_gt(val_dataset, cfg, i, use_mini_mask=cfg.USE_MINI_MASK)) for i in range(n_va)]
        gen_mode = "yolo" if self.mode == "yolo" else "training"
        train_gen = mutils.BatchGenerator(train_info, cfg, mode=gen_mode, shuffle=True, jitter=False, norm=True)
        val_gen = mutils.BatchGenerator(val_info, cfg, mode=gen_mode, shuffle=True, jitter=False, norm=True) if n_va else None
        self.set_trainable(layers)
        self.compile(learning_rate, getattr(cfg, "LEARNING_MOMENTUM", 0.9))
        stamp = datetime.datetime.now().strftime("%b %d-%H-%M")
        ckpt = os.path.join(self.model_dir or ".", "saved_model_" + stamp + ".pt")
        history = {"loss": [], "yolo_sum_loss": [], "myolo_mask_loss": [], "val_loss": []}
        B = self.engine.B
        for ep in range(self.epoch, epochs):
            t0, acc, nb = time.time(), np.zeros(3), 0
            full = (inputs for inputs, _ in Prefetcher(train_gen, range(len(train_gen)), max_queue_size, workers)
                    if inputs[0].shape[0] == B)
            for vals in self.fit_batches(full):
                acc[:len(vals)] += vals
                nb += 1
            acc /= max(nb, 1)
            vloss = float("nan")
            if val_gen is not None:
                vfull = (inputs for inputs, _ in Prefetcher(val_gen, range(len(val_gen)), max_queue_size, workers)
                         if inputs[0].shape[0] == B)
                vs = [v[0] for v in self.fit_batches(vfull, update=False)]
                vloss = float(np.mean(vs)) if vs else float("nan")
            for k, v in zip(("loss", "yolo_sum_loss", "myolo_mask_loss"), acc):
                history[k].append(float(v))
            history["val_loss"].append(vloss)
            if verbose:
                print("Epoch %d/%d - %.1fs - loss: %.4f - yolo_sum_loss: %.4f - myolo_mask_loss: %.4f - val_loss: %.4f"
                      % (ep + 1, epochs, time.time() - t0, acc[0], acc[1], acc[2], vloss))
            torch.save(self.engine.state_dict(), ckpt)
            train_gen.on_epoch_end()
        self.epoch = max(self.epoch, epochs)
        return history

    # ---- inference API
    def infer_yolo(self, image, weights_dir=None, save_path=None, display=False):
        """YOLO-only inference on one uint8 image (model.py:1198-1236): returns the decoded BoundBox
        list after the numpy NMS of decode_one_yolo_output."""
        assert image.dtype == np.uint8 and list(image.shape) == list(self.config.IMAGE_SHAPE)
        if weights_dir is not None:
            self.load_weights(weights_dir)
        x = (image / 255.)[None].astype(np.float32)
        netout = self._predict_b1(x)[0][0]
        return mutils.decode_one_yolo_output(netout, anchors=self.cfg["ANCHORS"], nms_threshold=0.3, obj_threshold=0.35,
                                             nb_class=self.cfg["NC"])       # thresholds of model.py:1227-1231

    def _engine_b1(self):
        """Batch-1 engine sharing this model's weights (detect / infer_yolo run one image at a time)."""
        if self.engine.B == 1:
            return self.engine
        src = self.engine
        if getattr(self, "_eng1", None) is None:
            mode = "inference" if self.mode != "yolo" else "yolo"
            self._eng1 = Engine(self.cfg, 1, mode, self.precision, self.device, params=src.state_dict())
            self._eng1_version = src.version
        elif self._eng1_version != src.version:           # weights changed since the last call: device-to-device copy
            self._eng1.params.copy_(src.params)
            for k, v in src.stats.items():
                self._eng1.stats[k].copy_(v)
            self._eng1.refresh_weights()
            self._eng1_version = src.version
        return self._eng1

    def _predict_b1(self, x):
        eng = self._engine_b1()
        img = torch.from_numpy(x).to(eng.dev)
        if eng.with_mask:
            yolo, det, masks = eng.forward_inference(img)
            return [yolo.cpu().numpy(), det.cpu().numpy(), masks.cpu().numpy()]
        return [eng.forward(img, training=False).cpu().numpy()]

    def detect(self, image, weights_dir=None, save_path=None, cs_threshold=0.35, display=False, top_k=10):
        """Full inference on one uint8 image (model.py:1238-1328), end to end on the GPU: network, top-10 by
        confidence, confidence threshold, NMB and mask paste (myolo_detect_postprocess).  Returns DetectResults: a
        one-element list holding the dict the reference's code builds ('bboxes', 'class_ids', 'confidence_scores',
        'full_masks', model.py:1316-1321) that also answers to the keys its docstring promises ('rois' (x1,y1,x2,y2
        pixels), 'class_ids', 'scores', boolean 'masks' [H,W,N]).  The reference's debugging overrides (hard-coded
        indices 1306, fixed 224 scale 1307) are not reproduced."""
        assert self.mode == "inference", "Create model in inference mode."
        assert image.dtype == np.uint8 and list(image.shape) == list(self.config.IMAGE_SHAPE)
        if weights_dir is not None:
            self.load_weights(weights_dir)
        eng = self._engine_b1()
        x = torch.from_numpy((image / 255.)[None].astype(np.float32)).to(eng.dev)
        eng.forward_inference(x)
        idx, boxes, cls, score, cnt, pm = eng.postprocess(top_k=top_k, cs_threshold=cs_threshold, nms_threshold=0.7)   # model.py:1304
        n = int(cnt[0].item())
        masks = pm[0, :n].permute(1, 2, 0).bool().cpu().numpy() if n else np.zeros(tuple(image.shape[:2]) + (0,), bool)
        return DetectResults(boxes[0, :n].cpu().numpy(), cls[0, :n].cpu().numpy(), score[0, :n].cpu().numpy(), masks)

    def detect_for_one(self, images, verbose=0):
        """The call the reference's example scripts make (example/shapes/infer_shapes.py:52, example/rice/rice_dataset.py:232)
        although myolo/model.py defines no such method at HEAD: detect() on a one-element list of images, same
        result list (`results[0]['rois' | 'masks' | 'class_ids' | 'scores']`)."""
        assert len(images) == 1, "only detect for one image per time"
        return self.detect(images[0])

    def decode_masks(self, detections, myolo_mask, image_shape):
        """Network outputs of ONE image -> (boxes [N,4] normalised (x1,y1,x2,y2), class_ids [N], scores [N], full-size
        boolean masks [H,W,N])  (model.py:1330-1391): class-specific 28x28 masks, zero-area boxes dropped, each mask
        resized into its box by unmold_mask."""
        assert len(detections) == 1            # only detect for one image per time
        assert len(myolo_mask) == 1
        assert list(image_shape) == list(self.config.IMAGE_SHAPE)
        detection = np.asarray(detections[0])
        masks_all = np.asarray(myolo_mask[0])
        N = len(detection)
        boxes = detection[:N, :4]
        scores = detection[:N, 4]
        class_ids = detection[:N, 5].astype(np.int32)
        masks = masks_all[np.arange(N), :, :, class_ids]
        exclude_ix = np.where((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) <= 0)[0]
        if exclude_ix.shape[0] > 0:
            boxes = np.delete(boxes, exclude_ix, axis=0)
            class_ids = np.delete(class_ids, exclude_ix, axis=0)
            scores = np.delete(scores, exclude_ix, axis=0)
            masks = np.delete(masks, exclude_ix, axis=0)
            N = class_ids.shape[0]
        full_masks = [mutils.unmold_mask(masks[i], boxes[i], image_shape) for i in range(N)]
        full_masks = np.stack(full_masks, axis=-1) if full_masks else np.empty(tuple(image_shape[:2]) + (0,))
        return boxes, class_ids, scores, full_masks
