"""SURVEY 8f row 4, second half: VIA polygon annotations -> instance masks (RiceDataset.load_mask,
example/rice/rice_dataset.py:135-159 = skimage.draw.polygon per instance).

scikit-image is not installable here and the reference ships no rasterised masks, so the rule is pinned as follows
(PARITY UNPINNED against the library itself, as the oracle's header says):
  * the oracle (oracle/via_polygon_oracle.py, plain Python loops over the published algorithm) against hand-computed cases;
  * the package's host form (myolo.rice.polygon, vectorised numpy) and the device kernel's inclusion test compiled for the
    host (csrc/polygon_pip.h via tests/polygon_pip_harness.cpp) against the oracle, bit for bit, on the outlines of the
    reference's own annotation files (tests/golden/via_polygons_fixture.json) and on random float / concave /
    self-intersecting / out-of-range outlines;
  * GPU: myolo_polygon_masks against the oracle, byte for byte, on the same inputs; RiceDataset.load_mask_device against
    load_mask; the masks feed extract_bboxes / load_image_gt unchanged."""
import ctypes
import json
import math
import os
import subprocess

import numpy as np
import pytest

from myolo import rice
from oracle import via_polygon_oracle as VO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "via_polygons_fixture.json")


def _fixture_images():
    return json.load(open(FIXTURE))


def _size_for(polygons, margin=3):
    h = int(math.ceil(max(max(p["all_points_y"]) for p in polygons))) + margin
    w = int(math.ceil(max(max(p["all_points_x"]) for p in polygons))) + margin
    return h, w


def _random_outlines(seed, n, size):
    """Float vertices, some outside the image on the low side (clipped at 0 by the rule), star-shaped (concave) and
    random-order (self-intersecting) outlines, horizontal edges and repeated vertices."""
    rng = np.random.RandomState(seed)
    out = []
    for k in range(n):
        nv = rng.randint(3, 12)
        kind = k % 4
        if kind == 0:        # star around a centre: concave
            cy, cx = rng.uniform(5, size - 6, 2)
            ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
            rad = rng.uniform(1.0, min(cy, cx, size - 1 - cy, size - 1 - cx), nv)
            ys, xs = cy + rad * np.sin(ang), cx + rad * np.cos(ang)
        elif kind == 1:      # random order: self-intersecting, float
            ys, xs = rng.uniform(0, size - 1, nv), rng.uniform(0, size - 1, nv)
        elif kind == 2:      # integer vertices with horizontal / vertical edges and a repeated vertex
            ys, xs = rng.randint(0, size, nv).astype(float), rng.randint(0, size, nv).astype(float)
            ys[1], xs[2] = ys[0], xs[1]
            ys[-1], xs[-1] = ys[0], xs[0]
        else:                # partly above / left of the image
            ys, xs = rng.uniform(-8, size - 1, nv), rng.uniform(-8, size - 1, nv)
        out.append({"all_points_y": [float(v) for v in ys], "all_points_x": [float(v) for v in xs]})
    return out


def _oracle_mask(polygons, h, w):
    return VO.load_mask(polygons, h, w)[0]


# ------------------------------------------------------------------------------------------------ the oracle itself
def test_oracle_hand_computed_cases():
    # axis-aligned square (2,2)-(6,6): rows and columns follow the half-open rule [2, 6) of the crossing test
    rr, cc = VO.polygon([2, 2, 6, 6], [2, 6, 6, 2])
    m = np.zeros((9, 9), int)
    m[rr, cc] = 1
    want = np.zeros((9, 9), int)
    want[2:6, 2:6] = 1
    assert (m == want).all()
    # vertex order does not matter for a simple outline
    rr2, cc2 = VO.polygon([6, 6, 2, 2], [2, 6, 6, 2])
    assert sorted(zip(rr, cc)) == sorted(zip(rr2, cc2))
    # right triangle (0,0), (0,4), (4,0) in (r, c): pixel (r, c) is inside iff c < 4 - r, r in [0, 4)
    rr, cc = VO.polygon([0, 0, 4], [0, 4, 0])
    assert sorted(zip(rr.tolist(), cc.tolist())) == [(r, c) for r in range(4) for c in range(4 - r)]
    # half-integer square: pixel centres 1..3 are strictly inside
    rr, cc = VO.polygon([0.5, 0.5, 3.5, 3.5], [0.5, 3.5, 3.5, 0.5])
    assert sorted(zip(rr.tolist(), cc.tolist())) == [(r, c) for r in (1, 2, 3) for c in (1, 2, 3)]
    # concave "U" (r, c): outer 0..6 x 0..6 with the slot rows 0..3, columns 2..4 cut out
    ys = [0, 0, 4, 4, 0, 0, 6, 6]
    xs = [0, 2, 2, 4, 4, 6, 6, 0]
    rr, cc = VO.polygon(ys, xs)
    m = np.zeros((8, 8), int)
    m[rr, cc] = 1
    want = np.zeros((8, 8), int)
    want[0:6, 0:6] = 1
    want[0:4, 2:4] = 0
    assert (m == want).all()
    # bow-tie (self-intersecting; vertical edges at c = 0 and c = 8): even-odd parity leaves the left and right triangles
    rr, cc = VO.polygon([0, 8, 0, 8], [0, 8, 8, 0])
    m = np.zeros((9, 9), int)
    m[rr, cc] = 1
    assert m[4, 1] == 1 and m[4, 6] == 1 and m[1, 4] == 0 and m[7, 4] == 0
    assert all(m[r, c] == (c < min(r, 8 - r) or c >= max(r, 8 - r)) for r in range(9) for c in range(8))
    # negative coordinates are clipped at 0 by the bounding box; `shape` clips the upper side
    rr, cc = VO.polygon([-3, -3, 2, 2], [-3, 2, 2, -3])
    assert sorted(zip(rr.tolist(), cc.tolist())) == [(r, c) for r in (0, 1) for c in (0, 1)]
    rr, cc = VO.polygon([0, 0, 10, 10], [0, 10, 10, 0], shape=(4, 5))
    assert rr.max() == 3 and cc.max() == 4 and len(rr) == 20
    # output order: row-major, like the library's double loop
    rr, cc = VO.polygon([2, 2, 6, 6], [2, 6, 6, 2])
    assert list(zip(rr, cc)) == sorted(zip(rr, cc))


def test_oracle_load_mask_contract():
    polys = [{"all_points_y": [1, 1, 5, 5], "all_points_x": [1, 5, 5, 1]},
             {"all_points_y": [3, 3, 7, 7], "all_points_x": [3, 7, 7, 3]}]
    mask, ids = VO.load_mask(polys, 9, 10)
    assert mask.dtype == bool and mask.shape == (9, 10, 2) and ids.dtype == np.int32 and ids.tolist() == [1, 1]
    assert mask[:, :, 0].sum() == 16 and mask[:, :, 1].sum() == 16 and (mask[:, :, 0] & mask[:, :, 1]).sum() == 4   # overlaps stay
    with pytest.raises(IndexError):                      # the reference indexes unclipped coordinates
        VO.load_mask([{"all_points_y": [0, 0, 12, 12], "all_points_x": [0, 4, 4, 0]}], 9, 10)


# ------------------------------------------------------------------------------------------------ host form + device rule on the CPU
@pytest.fixture(scope="module")
def pip_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pip") / "polygon_pip_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "mask-yolo_b200", "csrc"),
                           os.path.join(ROOT, "tests", "polygon_pip_harness.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.polygon_mask_host.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.polygon_mask_host.restype = None

    def run(p, h, w):
        ys, xs = np.asarray(p["all_points_y"], np.float64), np.asarray(p["all_points_x"], np.float64)
        m = np.empty((h, w), np.uint8)
        lib.polygon_mask_host(len(ys), ys.ctypes.data, xs.ctypes.data, h, w, m.ctypes.data)
        return m
    return run


def _cases():
    cases = [(im["polygons"],) + _size_for(im["polygons"]) for im in _fixture_images()[:8]]
    cases += [(_random_outlines(s, 6, 40), 40, 40) for s in (1, 2, 3)]
    cases += [(_random_outlines(9, 4, 33), 33, 61)]
    return cases


def test_package_host_form_equals_oracle():
    for polys, h, w in _cases():
        for p in polys:
            rr, cc = rice.polygon(p["all_points_y"], p["all_points_x"])
            orr, occ = VO.polygon(p["all_points_y"], p["all_points_x"])
            assert rr.dtype == orr.dtype and np.array_equal(rr, orr) and np.array_equal(cc, occ)       # same pixels, same order
            rr, cc = rice.polygon(p["all_points_y"], p["all_points_x"], shape=(h // 2, w // 2))
            orr, occ = VO.polygon(p["all_points_y"], p["all_points_x"], shape=(h // 2, w // 2))
            assert np.array_equal(rr, orr) and np.array_equal(cc, occ)


def test_device_rule_compiled_for_the_host_equals_oracle(pip_host):
    for polys, h, w in _cases():
        want = _oracle_mask(polys, h, w)
        for i, p in enumerate(polys):
            assert np.array_equal(pip_host(p, h, w).astype(bool), want[:, :, i])


def _write_via_dataset(root, images, v1=False, word="food"):
    """A VIA project on disk: <root>/train/*.png + via_<word>_annotation.json (2.x list regions, or 1.x dict regions)."""
    cv2 = pytest.importorskip("cv2")
    d = os.path.join(root, "train")
    os.makedirs(d, exist_ok=True)
    ann = {}
    for k, (polys, h, w) in enumerate(images):
        name = "%d.png" % k
        cv2.imwrite(os.path.join(d, name), np.full((h, w, 3), 40 + k, np.uint8))
        regs = [{"shape_attributes": dict(name="polygon", **p), "region_attributes": {}} for p in polys]
        ann[name + "123"] = {"filename": name, "size": 123, "file_attributes": {},
                             "regions": {str(i): r for i, r in enumerate(regs)} if v1 else regs}
    ann["empty.png9"] = {"filename": "empty.png", "size": 9, "regions": [] if not v1 else {}, "file_attributes": {}}
    json.dump(ann, open(os.path.join(d, "via_%s_annotation.json" % word), "w"))
    return root


@pytest.mark.parametrize("v1", [False, True])
def test_rice_dataset_host_chain(tmp_path, v1):
    """load_rice (both VIA region encodings, unannotated images skipped) -> load_mask -> extract_bboxes / load_image_gt."""
    from myolo import myolo_utils as mutils
    images = [(im["polygons"],) + _size_for(im["polygons"], margin=5) for im in _fixture_images()[:3]]
    ds = rice.RiceDataset()
    ds.load_rice(_write_via_dataset(str(tmp_path), images, v1), "train")
    ds.prepare()
    assert len(ds.image_ids) == 3 and ds.class_names == ["BG", "food"]
    for k, (polys, h, w) in enumerate(images):
        info = ds.image_info[k]
        assert (info["height"], info["width"]) == (h, w) and ds.image_reference(k) == info["path"]
        mask, ids = ds.load_mask(k)
        want, wids = VO.load_mask(polys, h, w)
        assert mask.dtype == bool and np.array_equal(mask, want) and np.array_equal(ids, wids) and ids.dtype == np.int32
        boxes = mutils.extract_bboxes(mask)
        for i, p in enumerate(polys):           # the box of a simple outline hugs its vertices (half-open on the high side)
            ys, xs = np.nonzero(want[:, :, i].any(1))[0], np.nonzero(want[:, :, i].any(0))[0]
            assert boxes[i].tolist() == [xs[0], ys[0], xs[-1] + 1, ys[-1] + 1]
    cfg = rice.RiceConfig()
    image, cls, boxes, masks = mutils.load_image_gt(ds, cfg, 0, use_mini_mask=False)
    assert image.shape[:2] == masks.shape[:2] and masks.shape[-1] == len(cls) == len(boxes) and (cls == 1).all()


def test_rice_config_matches_reference_values():
    c = rice.RiceConfig()
    assert (c.NAME, c.IMAGES_PER_GPU, c.GPU_COUNT, c.NUM_CLASSES) == ("food", 2, 0, 2)       # rice_dataset.py:60-82


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_device_masks_equal_oracle_byte_for_byte():
    import torch
    dp = rice.DevicePolygons(0)
    for polys, h, w in _cases():
        got = dp.masks(polys, h, w)
        torch.cuda.synchronize()
        assert got.dtype == torch.uint8 and tuple(got.shape) == (h, w, len(polys))
        assert np.array_equal(got.cpu().numpy().astype(bool), _oracle_mask(polys, h, w))
    # every fixture image at a size whose pixel count is not a multiple of the CTA's 256 pixels, many instances per launch
    allp = [p for im in _fixture_images() for p in im["polygons"]]
    h, w = _size_for(allp, margin=2)
    got = dp.masks(allp, h, w).cpu().numpy().astype(bool)
    assert np.array_equal(got, _oracle_mask(allp, h, w))


@pytest.mark.gpu
def test_device_entry_point_contract():
    import torch
    from myolo import _cabi as C
    dp = rice.DevicePolygons(0)
    # channels beyond n_inst are zero-filled (M > n_inst), the output is fully overwritten
    p = {"all_points_y": [1.0, 1.0, 5.0, 5.0], "all_points_x": [1.0, 5.0, 5.0, 1.0]}
    vy, vx = torch.tensor(p["all_points_y"], dtype=torch.float64).cuda(), torch.tensor(p["all_points_x"], dtype=torch.float64).cuda()
    off = torch.tensor([0, 4], dtype=torch.int32).cuda()
    out = torch.full((7, 9, 4), 255, dtype=torch.uint8, device="cuda")
    ws = torch.empty(16, dtype=torch.int32, device="cuda")
    C.call("myolo_polygon_masks", vy, vx, off, 1, 7, 9, 4, ws, out, None)
    o = out.cpu().numpy()
    assert o[:, :, 1:].sum() == 0 and np.array_equal(o[:, :, 0].astype(bool), _oracle_mask([p], 7, 9)[:, :, 0])
    with pytest.raises(C.MyoloError, match="argument check failed"):
        C.call("myolo_polygon_masks", vy, vx, off, 5, 7, 9, 4, ws, out, None)          # more instances than channels
    with pytest.raises(IndexError):                                                # the reference's error, before the launch
        dp.masks([{"all_points_y": [0, 0, 12, 12], "all_points_x": [0, 4, 4, 0]}], 9, 10)
    assert tuple(dp.masks([], 5, 6).shape) == (5, 6, 0)
    # a vertex ON the last row / column yields no pixel there (half-open rule): no error, like the reference
    ok = dp.masks([{"all_points_y": [0, 0, 8, 8], "all_points_x": [0, 9, 9, 0]}], 9, 10).cpu().numpy()
    assert ok[:, :, 0].sum() == 8 * 9


@pytest.mark.gpu
def test_rice_dataset_device_masks_equal_host_masks(tmp_path):
    import torch
    from myolo import _cabi as C
    images = [(im["polygons"],) + _size_for(im["polygons"], margin=4) for im in _fixture_images()[8:12]]
    ds = rice.RiceDataset()
    ds.load_rice(_write_via_dataset(str(tmp_path), images), "train")
    ds.prepare()
    dp = rice.DevicePolygons(0)
    for k in ds.image_ids:
        host, ids = ds.load_mask(k)
        dev, dids = ds.load_mask_device(k, dp)
        assert np.array_equal(dev.cpu().numpy().astype(bool), host) and np.array_equal(ids, dids)
    # square image: the device masks feed the device box extraction (myolo_extract_bboxes) like the host masks feed extract_bboxes
    from myolo import myolo_utils as mutils
    polys = images[0][0]
    s = max(_size_for(polys, margin=4))
    s += (-s) % 16
    m = dp.masks(polys, s, s)
    boxes = torch.zeros((1, m.shape[-1], 4), dtype=torch.int32, device="cuda")
    C.call("myolo_extract_bboxes", m.unsqueeze(0).contiguous(), 1, s, m.shape[-1], boxes, None)
    assert np.array_equal(boxes[0].cpu().numpy(), mutils.extract_bboxes(m.cpu().numpy().astype(bool)))


# ------------------------------------------------------------------------------------------------ live: the reference's own class
@pytest.mark.parametrize("example,word", [("rice", "food"), ("food", "rice")])
def test_rice_dataset_equals_the_references_own_class_live(tmp_path, example, word):
    """The reference's example/rice/rice_dataset.py, UNMODIFIED, imported here over this package (myolo.config, myolo.model,
    mrcnn.utils) with a stand-in for the two scikit-image calls it makes (`skimage.io.imread` -> cv2, `skimage.draw.polygon`
    -> the oracle): its RiceConfig / RiceDataset.load_rice / load_mask / image_reference against myolo.rice on the same VIA
    project, both region encodings.  Pins everything but the polygon primitive (container-only: needs /root/reference)."""
    import sys
    import types
    ex = "/root/reference/example/" + example        # example/rice says "food" everywhere, example/food says "rice"
    if not os.path.isdir(ex):
        pytest.skip("reference checkout not present on this box")
    cv2 = pytest.importorskip("cv2")
    sk, sk_draw, sk_io, sk_color = (types.ModuleType(n) for n in ("skimage", "skimage.draw", "skimage.io", "skimage.color"))
    sk_draw.polygon = VO.polygon
    sk_io.imread = lambda path: cv2.imread(path)[:, :, ::-1]
    sk.draw, sk.io, sk.color = sk_draw, sk_io, sk_color
    saved = {k: sys.modules.get(k) for k in ("skimage", "skimage.draw", "skimage.io", "skimage.color", "rice_dataset")}
    sys.modules.update({"skimage": sk, "skimage.draw": sk_draw, "skimage.io": sk_io, "skimage.color": sk_color})
    sys.path.insert(0, ex)
    try:
        import rice_dataset as ref                      # the reference's own file
        rc, mc = ref.RiceConfig(), (rice.RiceConfig() if word == "food" else rice.FoodExampleRiceConfig())
        for k in ("NAME", "IMAGES_PER_GPU", "GPU_COUNT", "NUM_CLASSES", "BATCH_SIZE"):
            assert getattr(rc, k) == getattr(mc, k), k
        images = [(im["polygons"],) + _size_for(im["polygons"], margin=5) for im in _fixture_images()[12:16]]
        for v1 in (False, True):
            root = _write_via_dataset(str(tmp_path / ("v1" if v1 else "v2")), images, v1, word)
            a, b = ref.RiceDataset(), rice.RiceDataset(source=word)
            a.load_rice(root, "train")
            b.load_rice(root, "train")
            a.prepare()
            b.prepare()
            assert a.class_info == b.class_info and len(a.image_info) == len(b.image_info) == len(images)
            for ia, ib in zip(a.image_info, b.image_info):
                assert ia == ib                          # id, source, path, width, height, polygons
            for k in a.image_ids:
                (ma, ca), (mb, cb) = a.load_mask(k), b.load_mask(k)
                assert ma.dtype == mb.dtype and np.array_equal(ma, mb) and ca.dtype == cb.dtype and np.array_equal(ca, cb)
                assert a.image_reference(k) == b.image_reference(k)
    finally:
        sys.path.remove(ex)
        sys.modules.pop("rice_dataset", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
