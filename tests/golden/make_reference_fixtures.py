"""Golden vectors produced by the REFERENCE's own code.

Run HERE (the container that has /root/reference); the .npz it writes is committed so tests never read
/root/reference at run time:

    python tests/golden/make_reference_fixtures.py

What can run: the host-side, pure-numpy functions of /root/reference/myolo/myolo_utils.py -- the target encoding of
BatchGenerator.__getitem__ (727-860), extract_bboxes (247-271), bbox_iou / bbox_iou_2 / _interval_overlap (187-244),
NMB (88-113), decode_one_yolo_output (36-85), _sigmoid / _softmax (21-33).  The module itself imports tensorflow,
keras, mrcnn, skimage, imgaug, matplotlib and distutils at the top; none of those exist in this image, so they are
replaced by EMPTY stub modules before the file is executed -- the functions listed above never touch them.  (The
Keras/TensorFlow graph of myolo/model.py cannot be run this way)
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/myolo/myolo_utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_utils_fixture.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_utils():
    class _Any(object):
        def __init__(self, *a, **k):
            pass

    _stub("tensorflow")
    _stub("mrcnn", utils=_stub("mrcnn.utils", Dataset=_Any))
    _stub("distutils", version=_stub("distutils.version", LooseVersion=_Any))
    sk = _stub("skimage", __version__="0.0")
    for sub in ("color", "io", "transform"):
        setattr(sk, sub, _stub("skimage." + sub))
    _stub("keras", utils=_stub("keras.utils", Sequence=object))
    _stub("imgaug", augmenters=_stub("imgaug.augmenters"))
    _stub("matplotlib", pyplot=_stub("matplotlib.pyplot"), patches=_stub("matplotlib.patches", Rectangle=_Any))
    if not hasattr(np, "bool"):          # the reference uses the numpy<1.24 alias
        np.bool = bool
    spec = importlib.util.spec_from_file_location("ref_myolo_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_shapes():
    """The REAL example/shapes/dataset_shapes.py (random specs 137-180, cv2 rasterisation 120-135, occlusion handling
    98-118).  Its `mrcnn` base class comes from this repository's shim (bookkeeping + matterport's greedy NMS restated);
    `myolo.model` (TensorFlow) is stubbed, `myolo.config` and `myolo.myolo_utils` are the reference's own files."""
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    spec = importlib.util.spec_from_file_location("mrcnn.utils", os.path.join(repo, "mask-yolo_b200", "mrcnn", "utils.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    _stub("mrcnn", utils=shim, visualize=_stub("mrcnn.visualize"))
    sys.modules["mrcnn.utils"] = shim
    for k in [k for k in sys.modules if k == "myolo" or k.startswith("myolo.")]:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    try:
        import myolo                                             # the reference's package
        assert all(p.startswith("/root/reference") for p in myolo.__path__), list(myolo.__path__)   # a namespace package
        sys.modules["myolo.model"] = types.ModuleType("myolo.model")
        myolo.model = sys.modules["myolo.model"]
        spec = importlib.util.spec_from_file_location("ref_dataset_shapes", "/root/reference/example/shapes/dataset_shapes.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove("/root/reference")
    return mod


class RefConfig(object):
    """The attributes BatchGenerator reads, with the Shapes values of example/shapes (3 anchors as in the shipped
    graph, 4 classes incl. background)."""
    BATCH_SIZE = 4
    IMAGE_SHAPE = [224, 224, 3]
    GRID_H = GRID_W = 7
    N_BOX = 3
    NUM_CLASSES = 4
    ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434]
    TRUE_BOX_BUFFER = 15
    MAX_GT_INSTANCES = 10


def make_all_info(rng, n_images, S=224):
    """BatchGenerator's `all_info` records [image, gt_class_ids, gt_boxes (x1,y1,x2,y2 px, x2/y2 exclusive), gt_masks]
    from random axis-aligned ellipses; instance counts 0..6, some boxes touching the border (grid index == G is
    dropped by the reference), two instances that fall into the same cell."""
    yy, xx = np.mgrid[0:S, 0:S]
    info = []
    for i in range(n_images):
        n = int(rng.randint(0, 7))
        masks = np.zeros((S, S, n), dtype=bool)
        ids = rng.randint(1, 4, size=n).astype(np.int32)
        for k in range(n):
            cx, cy = rng.uniform(10, S - 10, size=2)
            rx, ry = rng.uniform(6, 60, size=2)
            if i == 1 and k == 0:
                cx, cy, rx, ry = S - 3.0, S - 3.0, 8.0, 8.0          # hugs the bottom-right corner
            if i == 2 and k == 1 and n > 1:
                cx, cy = 100.0, 100.0
            if i == 2 and k == 0 and n > 1:
                cx, cy = 101.0, 99.0                                  # same cell as instance 1
            masks[:, :, k] = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
        image = ((xx[..., None] * (i + 1) + yy[..., None] * 3 + np.arange(3) * 40) % 256).astype(np.uint8)   # compressible
        info.append([image, ids, None, masks])
    return info


def main():
    ref = load_reference_utils()
    rng = np.random.RandomState(20260101)
    out = {}

    # ---- extract_bboxes (247-271)
    m = np.zeros((40, 50, 5), dtype=bool)
    m[3:9, 10:31, 0] = True
    m[39, 49, 1] = True
    m[0, 0, 2] = True
    m[5:30, 7, 3] = True                   # channel 4 stays empty
    out["eb_mask"] = m
    out["eb_boxes"] = ref.extract_bboxes(m)

    # ---- BatchGenerator.__getitem__ (727-860), mode 'training', norm=True, no shuffle
    cfg = RefConfig()
    info = make_all_info(rng, 6)
    for rec in info:
        keep = rec[3].sum(axis=(0, 1)) > 0                 # load_image_gt drops empty instances (347-349)
        rec[3] = rec[3][:, :, keep]
        rec[1] = rec[1][keep]
        rec[2] = ref.extract_bboxes(rec[3])
    gen = ref.BatchGenerator(info, cfg, mode="training", shuffle=False, norm=True)
    out["bg_n_images"] = np.int64(len(info))
    for i, rec in enumerate(info):
        out[f"bg_image_{i}"], out[f"bg_ids_{i}"], out[f"bg_boxes_{i}"], out[f"bg_masks_{i}"] = rec
    for b in range(len(gen)):
        inputs, outputs = gen[b]
        assert outputs == []
        for name, arr in zip(("images", "true_boxes", "yolo_target", "gt_class_ids", "gt_boxes", "gt_masks"), inputs):
            if name == "images":         # 2.4 MB per batch: keep dtype/shape, a checksum and one image row
                out[f"bg_batch{b}_images_meta"] = np.array(arr.shape + (arr.dtype.itemsize,), dtype=np.int64)
                out[f"bg_batch{b}_images_sum"] = arr.astype(np.float64).sum(axis=(1, 2, 3))
                out[f"bg_batch{b}_images_row"] = arr[:, 17]
            else:
                out[f"bg_batch{b}_{name}"] = arr
    out["bg_n_batches"] = np.int64(len(gen))

    # ---- bbox_iou (187-198) on BoundBox pairs, bbox_iou_2 (201-228) on normalised corner arrays
    a = rng.uniform(0, 1, size=(40, 2))
    boxes = np.concatenate([a, a + rng.uniform(0.01, 0.6, size=(40, 2))], axis=1)       # x1,y1,x2,y2
    boxes[5] = boxes[4]                                                                  # identical pair
    out["iou_boxes"] = boxes
    out["iou_pairs"] = np.array([[i, (i * 7 + 3) % 40] for i in range(40)], dtype=np.int64)
    out["iou_values"] = np.array([ref.bbox_iou(ref.BoundBox(*boxes[i]), ref.BoundBox(*boxes[j])) for i, j in out["iou_pairs"]])
    out["iou2_values"] = np.array([ref.bbox_iou_2(boxes[i], boxes[j], [224, 224, 3]) for i, j in out["iou_pairs"]])

    # ---- NMB (88-113): candidates already in score order, as detect() passes them (model.py:1291-1304)
    for case in range(4):
        n = 10
        c = rng.uniform(0.2, 0.8, size=(n, 2))
        wh = rng.uniform(0.1, 0.5, size=(n, 2))
        bx = np.concatenate([c - wh / 2, c + wh / 2], axis=1)
        bx[1] = bx[0] + 0.01                        # near-duplicate of the best box
        bx[3] = bx[1] + 0.02                        # overlaps box 1 (suppressed itself) more than box 0
        cls = rng.randint(1, 3, size=n)
        cls[1] = cls[0]
        cls[3] = cls[1]
        if case == 0:                               # chain A > B > C: A-B and B-C overlap 0.39, A-C only 0.14
            bx[0] = [0.10, 0.10, 0.50, 0.50]
            bx[1] = [0.20, 0.20, 0.60, 0.60]
            bx[2] = [0.30, 0.30, 0.70, 0.70]
            cls[:3] = 1
            bx[3:] += 2.0                           # everything else far away
        idx = rng.permutation(200)[:n]
        out[f"nmb{case}_boxes"], out[f"nmb{case}_class_ids"], out[f"nmb{case}_indices"] = bx, cls, idx
        out[f"nmb{case}_kept"] = np.asarray(ref.NMB(bx, cls, idx.copy(), [224, 224, 3], nms_threshold=0.3 + 0.2 * case))

    # ---- decode_one_yolo_output (36-85)
    for case in range(3):
        netout = rng.normal(0, 1.5, size=(4, 4, 3, 9))
        netout[..., 4] += 0.5
        out[f"dec{case}_netout"] = netout.copy()
        got = ref.decode_one_yolo_output(netout.copy(), cfg.ANCHORS, 4, obj_threshold=0.3, nms_threshold=0.3)
        out[f"dec{case}_boxes"] = np.array([[b.xmin, b.ymin, b.xmax, b.ymax, b.c] for b in got]).reshape(-1, 5)
        out[f"dec{case}_classes"] = np.array([b.classes for b in got]).reshape(-1, 4)
        out[f"dec{case}_label_score"] = np.array([[b.get_label(), b.get_score()] for b in got]).reshape(-1, 2)

    # ---- _sigmoid / _softmax (21-33)
    x = rng.normal(0, 30, size=(5, 7))
    out["sm_x"] = x
    out["sm_sigmoid"] = ref._sigmoid(x)
    out["sm_softmax"] = ref._softmax(x.copy())
    out["sm_softmax_small"] = ref._softmax(x.copy() / 30.0)

    # ---- ShapesDataset (example/shapes/dataset_shapes.py:53-180): specs, images, masks for seeded `random`
    import random
    shp = load_reference_shapes()
    for tag, seed, count, size in (("a", 1234, 10, 224), ("b", 7, 4, 128)):
        random.seed(seed)
        ds = shp.ShapesDataset()
        ds.load_shapes(count, size, size)
        ds.prepare()
        out[f"shp{tag}_meta"] = np.array([seed, count, size], dtype=np.int64)
        for i in ds.image_ids:
            info = ds.image_info[i]
            img = ds.load_image(i)
            mask, ids = ds.load_mask(i)
            out[f"shp{tag}_{i}_bg"] = np.asarray(info["bg_color"], dtype=np.int64)
            out[f"shp{tag}_{i}_specs"] = np.array([[["square", "circle", "triangle"].index(s[0])] + list(s[1]) + list(s[2])
                                                    for s in info["shapes"]], dtype=np.int64)
            out[f"shp{tag}_{i}_ids"] = ids
            out[f"shp{tag}_{i}_mask_bits"] = np.packbits(mask.astype(np.uint8))
            out[f"shp{tag}_{i}_mask_shape"] = np.array(mask.shape, dtype=np.int64)
            out[f"shp{tag}_{i}_image_rowsum"] = img.astype(np.int64).sum(axis=(1, 2))
            out[f"shp{tag}_{i}_image_colsum"] = img.astype(np.int64).sum(axis=(0, 2))
            out[f"shp{tag}_{i}_boxes"] = ref.extract_bboxes(mask)

    # ---- load_image_gt (274-366) and data_generator (457-686) of the reference's myolo_utils on its own ShapesDataset
    refu = sys.modules["myolo.myolo_utils"]
    assert refu.__file__.startswith("/root/reference")

    class GenConfig(shp.ShapesConfig):
        IMAGE_SHAPE = [128, 128, 3]
        GRID_H = GRID_W = 4
        N_BOX = 3
        TRUE_BOX_BUFFER = 10
        MAX_GT_INSTANCES = 10

    gcfg = GenConfig()
    random.seed(99)
    ds = shp.ShapesDataset()
    ds.load_shapes(5, 128, 128)
    ds.prepare()
    for i in ds.image_ids:
        image, class_ids, bbox, mask = refu.load_image_gt(ds, gcfg, i, use_mini_mask=False)
        out[f"lig_{i}_image_rowsum"] = image.astype(np.int64).sum(axis=(1, 2))
        out[f"lig_{i}_class_ids"], out[f"lig_{i}_bbox"] = class_ids, bbox
        out[f"lig_{i}_mask_bits"], out[f"lig_{i}_mask_shape"] = np.packbits(mask.astype(np.uint8)), np.array(mask.shape, dtype=np.int64)
    gen = refu.data_generator(ds, gcfg, shuffle=False, batch_size=2, norm=True)
    for b in range(3):                                   # 3 batches of 2 over 5 images: the third wraps around
        (images, true_boxes, yolo_target), outputs = next(gen)
        assert outputs == []
        out[f"dg_batch{b}_images_sum"] = images.astype(np.float64).sum(axis=(1, 2, 3))
        out[f"dg_batch{b}_images_row"] = images[:, 31]
        out[f"dg_batch{b}_images_meta"] = np.array(images.shape + (images.dtype.itemsize,), dtype=np.int64)
        out[f"dg_batch{b}_true_boxes"], out[f"dg_batch{b}_yolo_target"] = true_boxes, yolo_target
    gen.close()

    # ---- Config (myolo/config.py) and ShapesConfig (dataset_shapes.py:14-50): every public class attribute and the
    # attributes __init__ derives
    import json

    def dump(cls):
        def norm(v):
            return v.tolist() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, tuple) else v)
        d = {k: norm(getattr(cls, k)) for k in dir(cls) if not k.startswith("_") and not callable(getattr(cls, k))}
        d["__instance__"] = {k: norm(v) for k, v in vars(cls()).items()}
        return json.dumps(d, sort_keys=True)

    import myolo.config as refcfg                # still the reference's module (loaded by load_reference_shapes)
    assert refcfg.__file__.startswith("/root/reference")
    out["config_json"] = np.array(dump(refcfg.Config))
    out["shapes_config_json"] = np.array(dump(shp.ShapesConfig))

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KB")


def fuzz(path, n):
    """Reference side of tests/test_reference_live_fuzz.py: the reference's own myolo_utils through tests/golden/fuzz_cases."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fuzz_cases
    ref = load_reference_utils()
    assert ref.__file__ if hasattr(ref, "__file__") else True
    np.savez_compressed(path, **fuzz_cases.run(ref, n))


SHAPES_FUZZ = ((11, 128, 40), (12, 224, 40), (13, 320, 20), (14, 96, 40))        # (seed, image size, images)


def fuzz_shapes(path):
    """Reference side of the live Shapes test: its own ShapesDataset (seeded `random`) and load_image_gt."""
    import random
    load_reference_utils()
    shp = load_reference_shapes()
    refu = sys.modules["myolo.myolo_utils"]
    assert refu.__file__.startswith("/root/reference")
    out = {}
    for seed, size, count in SHAPES_FUZZ:
        class Cfg(shp.ShapesConfig):
            IMAGE_SHAPE = [size, size, 3]
        random.seed(seed)
        ds = shp.ShapesDataset()
        ds.load_shapes(count, size, size)
        ds.prepare()
        for i in ds.image_ids:
            info = ds.image_info[i]
            image, class_ids, bbox, mask = refu.load_image_gt(ds, Cfg(), i, use_mini_mask=False)
            tag = "%d_%d" % (seed, i)
            out[tag + "_specs"] = np.array([[["square", "circle", "triangle"].index(s[0])] + list(s[1]) + list(s[2]) for s in info["shapes"]],
                                           dtype=np.int64).reshape(-1, 7)
            out[tag + "_bg"] = np.asarray(info["bg_color"], dtype=np.int64)
            out[tag + "_image"] = np.array([image.astype(np.int64).sum(), (image.astype(np.int64) * np.arange(1, size + 1)[:, None, None]).sum()])
            out[tag + "_ids"], out[tag + "_bbox"] = class_ids, bbox
            out[tag + "_mask"] = np.packbits(mask.astype(np.uint8))
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--fuzz-shapes":
        fuzz_shapes(sys.argv[2])
    elif len(sys.argv) == 4 and sys.argv[1] == "--fuzz":
        fuzz(sys.argv[2], int(sys.argv[3]))
    else:
        main()
