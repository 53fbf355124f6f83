"""Synthetic Shapes data (squares, circles, triangles on a random background), the workload of
example/shapes/dataset_shapes.py:53-204 with a seeded `random.Random` instead of the global generator.
Shapes are rasterised with the same cv2 calls as the reference (120-135) -- specs, images and masks are then
identical to the reference's for the same seed (tests/test_reference_golden.py, golden vectors from the real file) --
and with a close numpy restatement when cv2 is not installed.

ShapesConfig mirrors dataset_shapes.py:14-50.  ShapesDataset follows the mrcnn Dataset protocol
(load_image / load_mask / image_ids), so load_image_gt and BatchGenerator consume it unchanged."""
import math
import random

import numpy as np

from mrcnn import utils
from .config import Config

try:
    import cv2
except Exception:       # pragma: no cover - the numpy restatement below takes over
    cv2 = None


class ShapesConfig(Config):
    NAME = "shapes"
    LABELS = ['background', 'square', 'circle', 'triangle']
    GPU_COUNT = 0
    IMAGES_PER_GPU = 8
    BATCH_SIZE = 16
    NUM_CLASSES = 1 + 3
    IMAGE_MIN_DIM = 224
    IMAGE_MAX_DIM = 224
    ANCHORS = [1.27273, 1.277385, 2.47446, 2.56253, 4.03843, 4.07434]
    N_BOX = 3                                   # = len(ANCHORS)//2: what the shipped TF graph was built with
    TRUE_BOX_BUFFER = 15
    MAX_GT_INSTANCES = 15                       # gt ids/boxes (TRUE_BOX_BUFFER wide) and gt masks must agree (model.py:487-493)
    CLASS_WEIGHTS = np.ones(NUM_CLASSES, dtype='float32')
    TRAIN_ROIS_PER_IMAGE = Config.GRID_H * Config.GRID_W * N_BOX
    USE_MINI_MASK = False


class ShapesDataset(utils.Dataset):
    def __init__(self, seed=None):
        super().__init__()
        self.rng = random.Random(seed)

    def load_shapes(self, count, height, width):
        self.add_class("shapes", 1, "square")
        self.add_class("shapes", 2, "circle")
        self.add_class("shapes", 3, "triangle")
        for i in range(count):
            bg_color, shapes = self.random_image(height, width)
            self.add_image("shapes", image_id=i, path=None, width=width, height=height, bg_color=bg_color, shapes=shapes)

    def image_reference(self, image_id):
        info = self.image_info[image_id]
        return info["shapes"] if info["source"] == "shapes" else super().image_reference(image_id)

    @staticmethod
    def _raster(shape, dims, h, w):
        x, y, s = dims
        yy, xx = np.mgrid[0:h, 0:w]
        if shape == "square":
            return (np.abs(xx - x) <= s) & (np.abs(yy - y) <= s)
        if shape == "circle":
            return (xx - x) ** 2 + (yy - y) ** 2 <= s * s
        # triangle with apex (x, y-s) and base corners (x -+ s/sin60, y+s)
        half = s / math.sin(math.radians(60))
        t = (yy - (y - s)) / (2.0 * s)
        return (t >= 0) & (t <= 1) & (np.abs(xx - x) <= half * t)

    @classmethod
    def draw_shape(cls, image, shape, dims, color):
        """dataset_shapes.py:120-135 (cv2.rectangle / circle / fillPoly, filled); `image` is [h, w, c] uint8."""
        x, y, s = dims
        if cv2 is None:
            image[cls._raster(shape, dims, image.shape[0], image.shape[1])] = color
            return image
        image = np.ascontiguousarray(image)
        if shape == "square":
            cv2.rectangle(image, (x - s, y - s), (x + s, y + s), color, -1)
        elif shape == "circle":
            cv2.circle(image, (x, y), s, color, -1)
        elif shape == "triangle":
            points = np.array([[(x, y - s), (x - s / math.sin(math.radians(60)), y + s),
                                (x + s / math.sin(math.radians(60)), y + s)]], dtype=np.int32)
            cv2.fillPoly(image, points, color)
        return image

    def load_image(self, image_id):
        info = self.image_info[image_id]
        img = np.ones([info["height"], info["width"], 3], dtype=np.uint8) * np.array(info["bg_color"], dtype=np.uint8).reshape(1, 1, 3)
        for shape, color, dims in info["shapes"]:
            img = self.draw_shape(img, shape, dims, color)
        return img

    def load_mask(self, image_id):
        info = self.image_info[image_id]
        shapes = info["shapes"]
        h, w = info["height"], info["width"]
        mask = np.zeros([h, w, len(shapes)], dtype=np.uint8)
        for i, (shape, _, dims) in enumerate(shapes):
            mask[:, :, i:i + 1] = self.draw_shape(mask[:, :, i:i + 1].copy(), shape, dims, 1)
        occlusion = np.logical_not(mask[:, :, -1]).astype(np.uint8) if shapes else None
        for i in range(len(shapes) - 2, -1, -1):      # later shapes occlude earlier ones
            mask[:, :, i] = mask[:, :, i] * occlusion
            occlusion = np.logical_and(occlusion, np.logical_not(mask[:, :, i]))
        class_ids = np.array([self.class_names.index(s[0]) for s in shapes], dtype=np.int32)
        return mask.astype(bool), class_ids

    def random_shape(self, height, width):
        shape = self.rng.choice(["square", "circle", "triangle"])
        color = tuple(self.rng.randint(0, 255) for _ in range(3))
        buffer = 20
        y = self.rng.randint(buffer, height - buffer - 1)
        x = self.rng.randint(buffer, width - buffer - 1)
        s = self.rng.randint(buffer, max(buffer, height // 4))
        return shape, color, (x, y, s)

    def random_image(self, height, width):
        bg_color = [self.rng.randint(0, 255) for _ in range(3)]
        shapes, boxes = [], []
        for _ in range(self.rng.randint(1, 4)):
            shape, color, dims = self.random_shape(height, width)
            shapes.append((shape, color, dims))
            x, y, s = dims
            boxes.append([y - s, x - s, y + s, x + s])
        keep = utils.non_max_suppression(np.array(boxes), np.arange(len(shapes)), 0.3)
        return bg_color, [s for i, s in enumerate(shapes) if i in keep]


def make_batches(config, n_batches, seed=1234, mode="training"):
    """`n_batches` BatchGenerator outputs (lists of numpy arrays) of synthetic Shapes at the config's
    IMAGE_SHAPE / BATCH_SIZE -- the benchmark's and smoke test's input source."""
    from . import myolo_utils as mutils
    S = int(config.IMAGE_SHAPE[0])
    ds = ShapesDataset(seed)
    ds.load_shapes(n_batches * config.BATCH_SIZE, S, S)
    ds.prepare()
    info = [list(mutils.load_image_gt(ds, config, i, use_mini_mask=False)) for i in ds.image_ids]
    state = np.random.get_state()
    np.random.seed(seed)
    gen = mutils.BatchGenerator(info, config, mode=mode, shuffle=False, norm=True)
    out = [gen[i][0] for i in range(len(gen))]
    np.random.set_state(state)
    return out


# --------------------------------------------------------------------------- device-side generation
_TYPE_ID = {"square": 1, "circle": 2, "triangle": 3}          # = class ids (load_shapes above)


def spec_table(dataset, image_ids=None, max_shapes=4):
    """The integers the device rasteriser needs, one row per image: [bg r, g, b, n] + max_shapes x
    [type, r, g, b, x, y, s, 0] (int32).  This is all that crosses the host -> device link per image."""
    ids = dataset.image_ids if image_ids is None else image_ids
    tab = np.zeros((len(ids), 4 + 8 * max_shapes), dtype=np.int32)
    for row, i in zip(tab, ids):
        info = dataset.image_info[i]
        shapes = info["shapes"]
        if len(shapes) > max_shapes:
            raise ValueError("image %d has %d shapes, max_shapes is %d" % (i, len(shapes), max_shapes))
        row[0:3] = info["bg_color"]
        row[3] = len(shapes)
        for k, (shape, color, (x, y, s)) in enumerate(shapes):
            row[4 + 8 * k: 4 + 8 * k + 7] = (_TYPE_ID[shape],) + tuple(color) + (x, y, s)
    return tab


class DeviceShapes(object):
    """Builds a training batch of the Shapes workload ON the GPU from its spec table: images, instance masks with
    occlusion, class ids, boxes (myolo_shapes_raster) and the YOLO target / true-box tensors (myolo_encode_yolo_targets) --
    the six inputs BatchGenerator(mode='training', norm=True) yields, bit for bit (tests/test_shapes_raster.py), without the
    host rasterisation and the 43 MB per-step upload.  Buffers are allocated once; batch() overwrites them, so the returned
    tensors are valid until the next call (the engine's recorded step keeps reading the same addresses)."""

    def __init__(self, config, device=0, max_shapes=4):
        import torch
        from .config import resolve
        self.torch = torch
        self.c = c = resolve(config)
        self.B, self.MS = int(config.BATCH_SIZE), int(max_shapes)
        self.M = int(config.MAX_GT_INSTANCES)
        B, S, G, TB = self.B, c["S"], c["G"], c["TB"]
        dev = self.dev = torch.device("cuda", device)
        # ring of pinned upload buffers: the async copy of call k may still be queued when call k+1 fills the next one
        self.spec_ring = [torch.empty(B, 4 + 8 * self.MS, dtype=torch.int32).pin_memory() for _ in range(4)]
        self.spec_done = [None] * len(self.spec_ring)
        self.calls = 0
        self.spec_dev = torch.empty_like(self.spec_ring[0], device=dev)
        self.ws = torch.empty(B * self.MS * (2 * S + 1), dtype=torch.int32, device=dev)
        self.images = torch.empty(B, S, S, 3, dtype=torch.float32, device=dev)
        self.masks = torch.empty(B, S, S, self.M, dtype=torch.uint8, device=dev)
        self.ids = torch.empty(B, TB, dtype=torch.int32, device=dev)
        self.boxes = torch.empty(B, TB, 4, dtype=torch.int32, device=dev)
        self.boxes_f = torch.empty(B, TB, 4, dtype=torch.float32, device=dev)
        self.yolo_target = torch.empty(B, G, G, c["NB"], 5 + c["NC"], dtype=torch.float32, device=dev)
        self.true_boxes = torch.empty(B, 1, 1, 1, TB, 4, dtype=torch.float32, device=dev)
        self.anchors = torch.tensor(c["ANCHORS"], dtype=torch.float32, device=dev)

    def batch(self, specs, image_u8=None):
        """specs: int32 [B, 4+8*max_shapes] (spec_table).  Returns [images, true_boxes, yolo_target, gt_class_ids,
        gt_boxes (float), gt_masks (bytes)] as device tensors; `image_u8` (optional [B,S,S,3] byte tensor) also receives
        the un-normalised image."""
        from . import _cabi as C
        torch, c = self.torch, self.c
        slot = self.calls % len(self.spec_ring)
        self.calls += 1
        host = self.spec_ring[slot]
        assert tuple(specs.shape) == tuple(host.shape), (specs.shape, host.shape)
        if self.spec_done[slot] is not None:
            self.spec_done[slot].synchronize()            # the upload that last used this pinned buffer has left it
        host.copy_(torch.from_numpy(np.ascontiguousarray(specs, dtype=np.int32)))
        self.spec_dev.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.spec_done[slot] = ev
        st = torch.cuda.current_stream(self.dev).cuda_stream
        C.call("myolo_shapes_raster", self.spec_dev, self.B, c["S"], self.MS, self.M, c["TB"], self.ws, self.images,
               image_u8, self.masks, self.ids, self.boxes, self.boxes_f, st)
        C.call("myolo_encode_yolo_targets", self.ids, self.boxes, self.B, c["TB"], c["S"], c["G"], c["NB"], c["NC"],
               c["TB"], self.anchors, self.yolo_target, self.true_boxes, st)
        return [self.images, self.true_boxes, self.yolo_target, self.ids, self.boxes_f, self.masks]
