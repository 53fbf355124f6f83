"""the Shapes workload generated on the device (myolo_shapes_raster, myolo.shapes.DeviceShapes).

CPU part: the row-extent functions the kernels are built from (csrc/shapes_extents.h, compiled here with g++) against
cv2.rectangle / cv2.circle / cv2.fillPoly pixel for pixel, and the kernels' composition rule ("the last shape covering a
pixel owns it") against the host chain load_image / load_mask / load_image_gt / extract_bboxes.
GPU part: DeviceShapes.batch() against BatchGenerator on the same images -- all six model inputs bit-exact -- and a
training step fed from device tensors against the same step fed from host arrays."""
import ctypes
import math
import os
import random
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from myolo import myolo_utils as mutils
from myolo.shapes import ShapesConfig, ShapesDataset, make_batches, spec_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIN60 = math.sin(math.radians(60))


@pytest.fixture(scope="module")
def ext(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shapes") / "shapes_extents_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "mask-yolo_b200", "csrc"),
                           os.path.join(ROOT, "tests", "shapes_extents_harness.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.shape_rows_host.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    lib.shape_rows_host.restype = None

    def rows(t, x, y, s, W, H):
        lo, hi = np.empty(H, np.int32), np.empty(H, np.int32)
        lib.shape_rows_host(t, x, y, s, W, H, lo.ctypes.data, hi.ctypes.data)
        return lo, hi
    return rows


def _mask_from_rows(lo, hi, W):
    c = np.arange(W)[None, :]
    return ((c >= lo[:, None]) & (c <= hi[:, None])).astype(np.uint8)


def _cv(t, x, y, s, W, H):
    m = np.zeros((H, W), np.uint8)
    if t == 1:
        cv2.rectangle(m, (x - s, y - s), (x + s, y + s), 1, -1)
    elif t == 2:
        cv2.circle(m, (x, y), s, 1, -1)
    else:
        cv2.fillPoly(m, np.array([[(x, y - s), (x - s / SIN60, y + s), (x + s / SIN60, y + s)]], dtype=np.int32), 1)
    return m


def test_row_extents_equal_cv2_on_the_generator_domain(ext):
    """dataset_shapes.py:137-158 draws centres in [20, S-21] and sizes in [20, S/4]: random samples plus the corners of
    that domain, three shape kinds, five image sizes."""
    rng = random.Random(11)
    n = 0
    for S in (64, 128, 224, 416, 640):
        cases = [(t, x, y, s) for t in (1, 2, 3) for x in (20, S - 21) for y in (20, S - 21) for s in (20, max(20, S // 4))]
        cases += [(rng.randint(1, 3), rng.randint(20, S - 21), rng.randint(20, S - 21), rng.randint(20, max(20, S // 4)))
                  for _ in range(2500 if S <= 224 else 500)]
        for t, x, y, s in cases:
            lo, hi = ext(t, x, y, s, S, S)
            assert np.array_equal(_mask_from_rows(lo, hi, S), _cv(t, x, y, s, S, S)), (S, t, x, y, s)
            n += 1
    assert n > 8000


def test_all_three_shapes_equal_cv2_anywhere(ext):
    """Rectangles, circles and triangles are exact for any centre / size / image shape: clipped on every side, degenerate
    (s = 0), apex outside the image, a side that touches the image in a single pixel, fully outside."""
    rng = random.Random(12)
    per_kind = {1: 0, 2: 0, 3: 0}
    for _ in range(9000):
        W, H, t = rng.randint(8, 80), rng.randint(8, 80), rng.randint(1, 3)
        x, y, s = rng.randint(-40, 120), rng.randint(-40, 120), rng.randint(0, 70)
        lo, hi = ext(t, x, y, s, W, H)
        assert np.array_equal(_mask_from_rows(lo, hi, W), _cv(t, x, y, s, W, H)), (W, H, t, x, y, s)
        per_kind[t] += 1
    assert min(per_kind.values()) > 2500
    # the cases that pinned the clipped-edge rule of fillPoly (a side reduced to one border pixel by clipLine)
    for W, H, x, y, s in ((37, 49, 39, 91, 47), (41, 27, 69, 10, 33), (67, 13, 93, -10, 24), (63, 64, -31, -5, 27)):
        lo, hi = ext(3, x, y, s, W, H)
        assert np.array_equal(_mask_from_rows(lo, hi, W), _cv(3, x, y, s, W, H)), (W, H, x, y, s)


def _compose(rows_fn, info, S, M, TB):
    """What the two kernels of shapes.cu compute, in numpy: owner = last shape covering the pixel; image colour and mask
    channel follow from the owner; instances without a visible pixel are dropped and the rest compacted."""
    shapes = info["shapes"]
    owner = np.full((S, S), -1, np.int32)
    for i, (name, _, (x, y, s)) in enumerate(shapes):
        lo, hi = rows_fn({"square": 1, "circle": 2, "triangle": 3}[name], x, y, s, S, S)
        owner[_mask_from_rows(lo, hi, S).astype(bool)] = i
    image = np.empty((S, S, 3), np.uint8)
    image[:] = np.array(info["bg_color"], np.uint8)
    masks = np.zeros((S, S, M), bool)
    ids, boxes, k = np.zeros(TB, np.int32), np.zeros((TB, 4), np.int32), 0
    for i, (name, color, _) in enumerate(shapes):
        vis = owner == i
        image[vis] = color
        if vis.any():
            masks[:, :, k] = vis
            ids[k] = {"square": 1, "circle": 2, "triangle": 3}[name]
            r, c = np.flatnonzero(vis.any(1)), np.flatnonzero(vis.any(0))
            boxes[k] = (c[0], r[0], c[-1] + 1, r[-1] + 1)
            k += 1
    return image, masks, ids, boxes


@pytest.mark.parametrize("S,count,seed", [(128, 40, 3), (224, 60, 1234)])
def test_owner_composition_equals_host_chain(ext, S, count, seed):
    class Cfg(ShapesConfig):
        IMAGE_SHAPE = [S, S, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = S
    cfg = Cfg()
    ds = ShapesDataset(seed)
    ds.load_shapes(count, S, S)
    ds.prepare()
    dropped = 0
    for i in ds.image_ids:
        image, class_ids, bbox, mask = mutils.load_image_gt(ds, cfg, i, use_mini_mask=False)
        im2, m2, ids2, bx2 = _compose(ext, ds.image_info[i], S, cfg.MAX_GT_INSTANCES, cfg.TRUE_BOX_BUFFER)
        n = class_ids.shape[0]
        dropped += len(ds.image_info[i]["shapes"]) - n
        assert np.array_equal(image, im2)
        assert np.array_equal(mask, m2[:, :, :n]) and not m2[:, :, n:].any()
        assert np.array_equal(class_ids, ids2[:n]) and not ids2[n:].any()
        assert np.array_equal(bbox, bx2[:n]) and not bx2[n:].any()


def test_spec_table_layout():
    ds = ShapesDataset(5)
    ds.load_shapes(6, 128, 128)
    ds.prepare()
    tab = spec_table(ds, max_shapes=4)
    assert tab.shape == (6, 36) and tab.dtype == np.int32
    for row, i in zip(tab, ds.image_ids):
        info = ds.image_info[i]
        assert row[:3].tolist() == list(info["bg_color"]) and row[3] == len(info["shapes"])
        for k, (name, color, dims) in enumerate(info["shapes"]):
            assert row[4 + 8 * k: 12 + 8 * k].tolist() == [["square", "circle", "triangle"].index(name) + 1, *color, *dims, 0]
        assert not row[4 + 8 * len(info["shapes"]):].any()
    with pytest.raises(ValueError):
        spec_table(ds, max_shapes=0)


# ------------------------------------------------------------------------------------------------ GPU
def _cfg(S, B, anchors=None):
    class Cfg(ShapesConfig):
        BATCH_SIZE = B
        IMAGE_SHAPE = [S, S, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = S
        GRID_H = GRID_W = S // 32
    if anchors is not None:
        Cfg.ANCHORS = anchors
    return Cfg()


@pytest.mark.gpu
@pytest.mark.parametrize("S,B,nb", [(224, 32, 2), (128, 4, 3), (416, 8, 1)])
def test_device_batches_equal_batchgenerator(S, B, nb):
    import torch
    from myolo.shapes import DeviceShapes
    cfg = _cfg(S, B)
    host = make_batches(cfg, nb, seed=77)
    ds = ShapesDataset(77)
    ds.load_shapes(nb * B, S, S)
    ds.prepare()
    tab = spec_table(ds)
    feeder = DeviceShapes(cfg)
    u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device="cuda")
    for k in range(nb):
        dev = feeder.batch(tab[k * B:(k + 1) * B], image_u8=u8)
        torch.cuda.synchronize()
        images, true_boxes, yolo_target, ids, boxes, masks = host[k]
        assert np.array_equal(dev[0].cpu().numpy(), images)                                    # float32(uint8 / 255.)
        assert np.array_equal(u8.cpu().numpy(), np.stack([ds.load_image(i) for i in range(k * B, (k + 1) * B)]))
        assert np.array_equal(dev[1].cpu().numpy(), true_boxes.astype(np.float32))
        assert np.array_equal(dev[2].cpu().numpy(), yolo_target.astype(np.float32))
        assert dev[3].dtype == torch.int32 and np.array_equal(dev[3].cpu().numpy(), ids)
        assert np.array_equal(dev[4].cpu().numpy(), boxes.astype(np.float32))
        assert np.array_equal(feeder.boxes.cpu().numpy(), boxes)
        assert np.array_equal(dev[5].cpu().numpy().astype(bool), masks)
        # extract_bboxes kernel on the rasterised masks agrees with the boxes the rasteriser tracked itself
        from myolo import _cabi as C
        bx = torch.empty(B, masks.shape[3], 4, dtype=torch.int32, device="cuda")
        C.call("myolo_extract_bboxes", dev[5], B, S, masks.shape[3], bx, torch.cuda.current_stream().cuda_stream)
        assert torch.equal(bx[:, :boxes.shape[1]], feeder.boxes[:, :bx.shape[1]])


@pytest.mark.gpu
def test_training_step_from_device_batches_equals_host_fed_step():
    import torch
    from myolo.model import MaskYOLO
    from myolo.shapes import DeviceShapes
    cfg = _cfg(128, 4, anchors=[0.6, 0.6, 1.2, 1.3, 2.0, 2.1])
    host = make_batches(cfg, 2, seed=9)
    ds = ShapesDataset(9)
    ds.load_shapes(8, 128, 128)
    ds.prepare()
    tab = spec_table(ds)
    a, b = MaskYOLO("training", cfg, seed=3), MaskYOLO("training", cfg, seed=3)
    feeder = DeviceShapes(cfg)
    for k in range(2):
        va = a.keras_model.train_on_batch(host[k])
        vb = b.keras_model.train_on_batch(feeder.batch(tab[4 * k:4 * k + 4]))
        assert b.last_h2d_bytes == 0
        assert np.allclose(va, vb, rtol=1e-4, atol=1e-6), (k, va, vb)
    with pytest.raises(TypeError):
        b.keras_model.train_on_batch([t.double() for t in feeder.batch(tab[:4])])


@pytest.mark.gpu
def test_shapes_raster_rejects_bad_arguments():
    import torch
    from myolo import _cabi as C
    z = torch.zeros(64, dtype=torch.int32, device="cuda")
    m = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(C.MyoloError):          # S not a multiple of 16
        C.call("myolo_shapes_raster", z, 1, 40, 4, 10, 10, z, None, None, m, z, z, None, 0)
    with pytest.raises(C.MyoloError):          # more shapes than mask channels
        C.call("myolo_shapes_raster", z, 1, 64, 4, 2, 10, z, None, None, m, z, z, None, 0)


def test_owner_composition_drops_fully_occluded_instances_like_load_image_gt(ext):
    """A hand-made image the generator's NMS would never produce: shape 0 is completely covered by shape 2, shape 1 partly.
    load_image_gt drops the invisible instance and keeps the order of the others; so does the owner rule."""
    S = 96

    class Cfg(ShapesConfig):
        IMAGE_SHAPE = [S, S, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = S

    cfg = Cfg()
    ds = ShapesDataset(0)
    ds.add_class("shapes", 1, "square")
    ds.add_class("shapes", 2, "circle")
    ds.add_class("shapes", 3, "triangle")
    shapes = [("circle", (10, 20, 30), (40, 40, 8)), ("triangle", (50, 60, 70), (60, 50, 20)), ("square", (90, 100, 110), (42, 42, 14))]
    ds.add_image("shapes", image_id=0, path=None, width=S, height=S, bg_color=[1, 2, 3], shapes=shapes)
    ds.add_image("shapes", image_id=1, path=None, width=S, height=S, bg_color=[7, 8, 9], shapes=[])
    ds.prepare()
    image, class_ids, bbox, mask = mutils.load_image_gt(ds, cfg, 0, use_mini_mask=False)
    assert class_ids.tolist() == [3, 1] and mask.shape == (S, S, 2)               # the circle is gone
    im2, m2, ids2, bx2 = _compose(ext, ds.image_info[0], S, cfg.MAX_GT_INSTANCES, cfg.TRUE_BOX_BUFFER)
    assert np.array_equal(image, im2) and np.array_equal(mask, m2[:, :, :2]) and not m2[:, :, 2:].any()
    assert ids2[:3].tolist() == [3, 1, 0] and np.array_equal(bbox, bx2[:2]) and not bx2[2:].any()
    tab = spec_table(ds)
    assert tab[0, 3] == 3 and tab[1, 3] == 0 and tab[1, :3].tolist() == [7, 8, 9] and not tab[1, 4:].any()
    im3, m3, ids3, bx3 = _compose(ext, ds.image_info[1], S, cfg.MAX_GT_INSTANCES, cfg.TRUE_BOX_BUFFER)
    assert (im3 == np.array([7, 8, 9], np.uint8)).all() and not m3.any() and not ids3.any() and not bx3.any()


@pytest.mark.gpu
def test_device_raster_drops_fully_occluded_instances_and_handles_empty_images():
    import torch
    from myolo import _cabi as C
    S, MS, M, TB = 96, 4, 10, 10
    specs = np.zeros((2, 4 + 8 * MS), np.int32)
    specs[0, :4] = (1, 2, 3, 3)
    specs[0, 4:11] = (2, 10, 20, 30, 40, 40, 8)            # circle, later covered completely by the square
    specs[0, 12:19] = (3, 50, 60, 70, 60, 50, 20)          # triangle, partly covered
    specs[0, 20:27] = (1, 90, 100, 110, 42, 42, 14)        # square
    specs[1, :4] = (7, 8, 9, 0)                            # no shapes at all
    dev = torch.device("cuda")
    sp = torch.from_numpy(specs).to(dev)
    ws = torch.empty(2 * MS * (2 * S + 1), dtype=torch.int32, device=dev)
    img = torch.empty(2, S, S, 3, dtype=torch.uint8, device=dev)
    masks = torch.empty(2, S, S, M, dtype=torch.uint8, device=dev)
    ids = torch.empty(2, TB, dtype=torch.int32, device=dev)
    boxes = torch.empty(2, TB, 4, dtype=torch.int32, device=dev)
    C.call("myolo_shapes_raster", sp, 2, S, MS, M, TB, ws, None, img, masks, ids, boxes, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()

    class Cfg(ShapesConfig):
        IMAGE_SHAPE = [S, S, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = S
        MAX_GT_INSTANCES = M
        TRUE_BOX_BUFFER = TB

    ds = ShapesDataset(0)
    for k, name in enumerate(("square", "circle", "triangle")):
        ds.add_class("shapes", k + 1, name)
    ds.add_image("shapes", image_id=0, path=None, width=S, height=S, bg_color=[1, 2, 3],
                 shapes=[("circle", (10, 20, 30), (40, 40, 8)), ("triangle", (50, 60, 70), (60, 50, 20)), ("square", (90, 100, 110), (42, 42, 14))])
    ds.prepare()
    image, class_ids, bbox, mask = mutils.load_image_gt(ds, Cfg(), 0, use_mini_mask=False)
    assert np.array_equal(img[0].cpu().numpy(), image)
    assert ids[0].cpu().tolist() == [3, 1] + [0] * (TB - 2) and np.array_equal(boxes[0, :2].cpu().numpy(), bbox)
    assert np.array_equal(masks[0, :, :, :2].cpu().numpy().astype(bool), mask) and not masks[0, :, :, 2:].any()
    assert (img[1].cpu().numpy() == np.array([7, 8, 9], np.uint8)).all() and not masks[1].any() and not ids[1].any() and not boxes.cpu()[1].any()
