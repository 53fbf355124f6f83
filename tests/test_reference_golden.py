"""Host-side functions of the package against golden vectors produced by the REFERENCE's own code
(tests/golden/reference_utils_fixture.npz, written by tests/golden/make_reference_fixtures.py from
/root/reference/myolo/myolo_utils.py with its third-party imports stubbed out).  These pin SURVEY 8a row a14 and the
numpy side of 8f rows 1-2 to the reference itself; the device kernels for the same rows are compared with the same
vectors in tests/test_kernels_gpu.py.  Integer / index / boolean outputs must be identical, float64 outputs equal to
the last bit (same operations in the same order)."""
import os

import numpy as np
import pytest

from myolo import myolo_utils as mutils
from myolo.shapes import ShapesConfig

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_utils_fixture.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(FIX)


class _Cfg(ShapesConfig):
    BATCH_SIZE = 4
    IMAGE_SHAPE = [224, 224, 3]
    GRID_H = GRID_W = 7
    N_BOX = 3
    NUM_CLASSES = 4
    ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434]
    TRUE_BOX_BUFFER = 15
    MAX_GT_INSTANCES = 10


def test_extract_bboxes(gold):
    got = mutils.extract_bboxes(gold["eb_mask"])
    assert got.dtype == gold["eb_boxes"].dtype and np.array_equal(got, gold["eb_boxes"])


def _all_info(gold):
    n = int(gold["bg_n_images"])
    return [[gold[f"bg_image_{i}"], gold[f"bg_ids_{i}"], gold[f"bg_boxes_{i}"], gold[f"bg_masks_{i}"]] for i in range(n)]


def test_batch_generator_target_encoding(gold):
    """BatchGenerator.__getitem__ (myolo_utils.py:727-860): images / true boxes / YOLO target / padded GT arrays"""
    info = _all_info(gold)
    for rec in info:      # the boxes the reference's extract_bboxes produced for these masks
        assert np.array_equal(mutils.extract_bboxes(rec[3]), rec[2])
    gen = mutils.BatchGenerator(info, _Cfg(), mode="training", shuffle=False, norm=True)
    assert len(gen) == int(gold["bg_n_batches"])
    for b in range(len(gen)):
        (images, tb, yt, ids, boxes, masks), outputs = gen[b]
        assert outputs == []
        meta = gold[f"bg_batch{b}_images_meta"]
        assert tuple(images.shape) == tuple(meta[:4]) and images.dtype.itemsize == meta[4]
        assert np.array_equal(images.astype(np.float64).sum(axis=(1, 2, 3)), gold[f"bg_batch{b}_images_sum"])
        assert np.array_equal(images[:, 17], gold[f"bg_batch{b}_images_row"])
        for name, got in (("true_boxes", tb), ("yolo_target", yt), ("gt_class_ids", ids), ("gt_boxes", boxes), ("gt_masks", masks)):
            ref = gold[f"bg_batch{b}_{name}"]
            assert got.shape == ref.shape and got.dtype == ref.dtype, (name, got.dtype, ref.dtype)
            assert np.array_equal(got, ref), (b, name)
    # the fixture exercises the interesting branches
    yt_all = np.concatenate([gold[f"bg_batch{b}_yolo_target"] for b in range(len(gen))])
    assert yt_all[..., 4].sum() >= 8


def test_iou_helpers(gold):
    boxes = gold["iou_boxes"]
    for (i, j), v, v2 in zip(gold["iou_pairs"], gold["iou_values"], gold["iou2_values"]):
        assert mutils.bbox_iou(mutils.BoundBox(*boxes[i]), mutils.BoundBox(*boxes[j])) == v
        assert mutils.bbox_iou_2(boxes[i], boxes[j], [224, 224, 3]) == v2


def test_nmb(gold):
    """NMB (myolo_utils.py:88-113): a candidate is dropped when ANY earlier candidate of the same class overlaps it by
    at least the threshold -- also an earlier candidate that was dropped itself (this is not greedy NMS)."""
    differs_from_greedy = False
    for case in range(4):
        bx, cls, idx = gold[f"nmb{case}_boxes"], gold[f"nmb{case}_class_ids"], gold[f"nmb{case}_indices"]
        thr = 0.3 + 0.2 * case
        kept = mutils.NMB(bx, cls, idx.copy(), [224, 224, 3], nms_threshold=thr)
        assert np.array_equal(np.asarray(kept), gold[f"nmb{case}_kept"]), case
        greedy = []
        for a in range(len(idx)):
            if all(not (cls[a] == cls[k] and mutils.bbox_iou_2(bx[k], bx[a], [224, 224, 3]) >= thr) for k in greedy):
                greedy.append(a)
        differs_from_greedy |= list(idx[greedy]) != list(gold[f"nmb{case}_kept"])
    assert differs_from_greedy, "the fixture must contain a chain A>B>C that separates the reference rule from greedy NMS"


def test_decode_one_yolo_output(gold):
    anchors = _Cfg.ANCHORS
    for case in range(3):
        got = mutils.decode_one_yolo_output(gold[f"dec{case}_netout"].copy(), anchors, 4, obj_threshold=0.3, nms_threshold=0.3)
        ref = gold[f"dec{case}_boxes"]
        assert len(got) == len(ref) and len(ref) > 0
        arr = np.array([[b.xmin, b.ymin, b.xmax, b.ymax, b.c] for b in got])
        assert np.array_equal(arr, ref)
        assert np.array_equal(np.array([b.classes for b in got]), gold[f"dec{case}_classes"])
        assert np.array_equal(np.array([[b.get_label(), b.get_score()] for b in got]), gold[f"dec{case}_label_score"])


def test_sigmoid_softmax(gold):
    x = gold["sm_x"]
    assert np.array_equal(mutils._sigmoid(x), gold["sm_sigmoid"])
    assert np.array_equal(mutils._softmax(x.copy()), gold["sm_softmax"])
    assert np.array_equal(mutils._softmax(x.copy() / 30.0), gold["sm_softmax_small"])


def test_shapes_dataset_matches_reference(gold):
    """ShapesDataset (example/shapes/dataset_shapes.py:53-180): for the same seed the random specs, the rasterised
    images and the occlusion-resolved masks are those of the reference's own file (cv2 rasterisation)."""
    pytest.importorskip("cv2")
    from myolo.shapes import ShapesDataset
    for tag in ("a", "b"):
        seed, count, size = (int(v) for v in gold[f"shp{tag}_meta"])
        ds = ShapesDataset(seed=seed)
        ds.load_shapes(count, size, size)
        ds.prepare()
        assert len(ds.image_ids) == count
        for i in ds.image_ids:
            info = ds.image_info[i]
            assert list(info["bg_color"]) == gold[f"shp{tag}_{i}_bg"].tolist()
            specs = [[["square", "circle", "triangle"].index(s[0])] + list(s[1]) + list(s[2]) for s in info["shapes"]]
            assert specs == gold[f"shp{tag}_{i}_specs"].tolist(), (tag, i)
            mask, ids = ds.load_mask(i)
            assert np.array_equal(ids, gold[f"shp{tag}_{i}_ids"]) and ids.dtype == gold[f"shp{tag}_{i}_ids"].dtype
            assert list(mask.shape) == gold[f"shp{tag}_{i}_mask_shape"].tolist() and mask.dtype == bool
            assert np.array_equal(np.packbits(mask.astype(np.uint8)), gold[f"shp{tag}_{i}_mask_bits"]), (tag, i)
            img = ds.load_image(i)
            assert img.dtype == np.uint8
            assert np.array_equal(img.astype(np.int64).sum(axis=(1, 2)), gold[f"shp{tag}_{i}_image_rowsum"])
            assert np.array_equal(img.astype(np.int64).sum(axis=(0, 2)), gold[f"shp{tag}_{i}_image_colsum"])
            assert np.array_equal(mutils.extract_bboxes(mask), gold[f"shp{tag}_{i}_boxes"])


def test_config_attributes_match_reference(gold):
    """Config (myolo/config.py) attribute for attribute; ShapesConfig (dataset_shapes.py:14-50) except the fields
    the reference leaves inconsistent with its own shipped graph (N_BOX inherited as 5 with 3 anchors)."""
    import json
    from myolo.config import Config

    def dump(cls):
        def norm(v):
            return v.tolist() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, tuple) else v)
        d = {k: norm(getattr(cls, k)) for k in dir(cls) if not k.startswith("_") and not callable(getattr(cls, k))}
        d["__instance__"] = {k: norm(v) for k, v in vars(cls()).items()}
        return d

    ref = json.loads(str(gold["config_json"]))
    assert dump(Config) == ref
    ref_s, mine_s = json.loads(str(gold["shapes_config_json"])), dump(ShapesConfig)
    # inherited from the base class although ShapesConfig changes what they depend on: N_BOX 5 with 3 anchors,
    # CLASS_WEIGHTS of length 2 with 4 classes (tf.gather out of range); TRUE_BOX_BUFFER / MAX_GT_INSTANCES are 10 at
    # HEAD but the graph the reference ships was built 15 wide (graph_fixture.json: input_true_boxes [-1,1,1,1,15,4])
    repaired = {"N_BOX", "TRUE_BOX_BUFFER", "MAX_GT_INSTANCES", "TRAIN_ROIS_PER_IMAGE", "CLASS_WEIGHTS"}
    for k in set(ref_s) | set(mine_s):
        if k in repaired or k == "__instance__":
            continue
        assert mine_s.get(k) == ref_s.get(k), k
    assert ref_s["N_BOX"] == 5 and len(ref_s["ANCHORS"]) == 6 and mine_s["N_BOX"] == 3
    assert ref_s["TRUE_BOX_BUFFER"] == 10 and mine_s["TRUE_BOX_BUFFER"] == mine_s["MAX_GT_INSTANCES"] == 15
    assert len(ref_s["CLASS_WEIGHTS"]) == 2 and ref_s["NUM_CLASSES"] == 4 and len(mine_s["CLASS_WEIGHTS"]) == 4
    for k in set(ref_s["__instance__"]) | set(mine_s["__instance__"]):
        if k not in repaired:
            assert mine_s["__instance__"].get(k) == ref_s["__instance__"].get(k), k


def test_load_image_gt_and_data_generator_match_reference(gold):
    """load_image_gt (myolo_utils.py:274-366) and the older python generator data_generator (457-686) on the Shapes
    dataset, against the reference's own functions run on its own ShapesDataset with the same seed."""
    from myolo.shapes import ShapesDataset

    class GenConfig(ShapesConfig):
        IMAGE_SHAPE = [128, 128, 3]
        GRID_H = GRID_W = 4
        N_BOX = 3
        TRUE_BOX_BUFFER = 10
        MAX_GT_INSTANCES = 10

    cfg = GenConfig()
    ds = ShapesDataset(seed=99)
    ds.load_shapes(5, 128, 128)
    ds.prepare()
    for i in ds.image_ids:
        image, class_ids, bbox, mask = mutils.load_image_gt(ds, cfg, i, use_mini_mask=False)
        assert image.dtype == np.uint8 and np.array_equal(image.astype(np.int64).sum(axis=(1, 2)), gold[f"lig_{i}_image_rowsum"])
        assert np.array_equal(class_ids, gold[f"lig_{i}_class_ids"]) and class_ids.dtype == gold[f"lig_{i}_class_ids"].dtype
        assert np.array_equal(bbox, gold[f"lig_{i}_bbox"]) and bbox.dtype == gold[f"lig_{i}_bbox"].dtype
        assert list(mask.shape) == gold[f"lig_{i}_mask_shape"].tolist()
        assert np.array_equal(np.packbits(mask.astype(np.uint8)), gold[f"lig_{i}_mask_bits"])
    gen = mutils.data_generator(ds, cfg, shuffle=False, batch_size=2, norm=True)
    for b in range(3):
        (images, true_boxes, yolo_target), outputs = next(gen)
        assert outputs == []
        assert list(images.shape) + [images.dtype.itemsize] == gold[f"dg_batch{b}_images_meta"].tolist()
        assert np.array_equal(images.astype(np.float64).sum(axis=(1, 2, 3)), gold[f"dg_batch{b}_images_sum"])
        assert np.array_equal(images[:, 31], gold[f"dg_batch{b}_images_row"])
        for name, arr in (("true_boxes", true_boxes), ("yolo_target", yolo_target)):
            ref = gold[f"dg_batch{b}_{name}"]
            assert arr.shape == ref.shape and arr.dtype == ref.dtype and np.array_equal(arr, ref), (b, name)
    gen.close()
