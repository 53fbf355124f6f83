"""Live differential test of the host functions: where the reference checkout is present (this container; never the GPU
box, and nothing GPU-marked depends on it) the reference's OWN myolo_utils -- executed in a subprocess with its
third-party imports stubbed, exactly like the golden-vector generator does -- and the package run through the same 60
seeded random cases (tests/golden/fuzz_cases.py): decode_one_yolo_output, NMB, bbox_iou, bbox_iou_2, extract_bboxes and
the BatchGenerator encoding must agree to the last bit.  Skipped when /root/reference does not exist; the committed golden
vectors (tests/test_reference_golden.py) are what travels."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/myolo/myolo_utils.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is only present in the build container")
def test_host_functions_agree_with_the_reference_on_random_cases(tmp_path):
    n = 60
    out = str(tmp_path / "ref_fuzz.npz")
    env = dict(os.environ, PYTHONPATH="")                    # the subprocess must not see this package's `myolo`
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_fixtures.py"), "--fuzz", out, str(n)],
                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(out)
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import fuzz_cases
    from myolo import myolo_utils as U
    assert U.__file__.startswith(os.path.dirname(HERE))       # the package, not the reference
    mine = fuzz_cases.run(U, n)
    assert set(mine) == set(ref.files) and len(mine) > 5 * n
    for k in sorted(mine):
        a, b = np.asarray(mine[k]), ref[k]
        assert a.shape == b.shape and a.dtype == b.dtype, (k, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), k
    # the cases are not vacuous
    assert sum(mine["dec%d" % k].shape[0] for k in range(n)) > n
    assert sum(len(mine["nmb%d" % k]) for k in range(n)) < sum(len(mine["iou_%d" % k]) for k in range(n))     # something was suppressed
    assert any((mine["eb%d" % k][3] == 0).all() for k in range(n))                                               # the empty channel


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is only present in the build container")
def test_oracle_agrees_with_the_reference_graph_functions_on_random_configurations(tmp_path):
    """The ORACLE against the reference's own graph functions (yolo_custom_loss incl. warm-up, DecodeYOLOLayer,
    DetectionsLayer, norm_boxes_graph, DetectMaskTargetLayer, PyramidROIAlign, myolo_mask_loss_graph) run live over the
    numpy stand-in for TensorFlow, on 24 random configurations: grid 2..7, 1..5 anchors, 2..6 classes, buffer 5..12, random
    loss scales / class weights / anchors, warm-up on and off, images without ground truth."""
    import torch
    n = 24
    out = str(tmp_path / "ref_graph_fuzz.npz")
    env = dict(os.environ, PYTHONPATH="")
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_graph_fixtures.py"), "--fuzz", out, str(n)],
                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(out)
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import graph_case_inputs as GI
    from oracle import myolo_oracle as O
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x))                     # noqa: E731
    positives = 0
    for i in range(n):
        c = GI.build(GI.fuzz_case(i))
        g = lambda k: ref["%d/%s" % (i, k)]                                     # noqa: E731
        warm = 3 if i % 3 == 1 else 0
        cfg = dict(GRID_H=c["G"], GRID_W=c["G"], N_BOX=c["NB"], NUM_CLASSES=c["NC"], ANCHORS=c["ANCHORS"],
                   TRAIN_ROIS_PER_IMAGE=c["R"], MASK_SHAPE=[28, 28], MASK_POOL_SIZE=14, COORD_SCALE=c["COORD_SCALE"],
                   NO_OBJECT_SCALE=c["NO_OBJECT_SCALE"], OBJECT_SCALE=c["OBJECT_SCALE"], CLASS_SCALE=c["CLASS_SCALE"],
                   CLASS_WEIGHTS=np.asarray(c["CLASS_WEIGHTS"], np.float32), WARM_UP_BATCHES=warm, TRUE_BOX_BUFFER=c["TB"])
        y_pred = t(c["y_pred"])
        loss = O.yolo_custom_loss(t(c["y_true"]), y_pred, t(c["true_boxes"]), cfg, seen=1.0).item()
        assert np.isclose(loss, g("yolo_loss"), rtol=2e-5), (i, loss, g("yolo_loss"))
        scale = max(1.0, float(np.abs(g("proposals")).max()))
        assert np.allclose(O.decode_yolo(y_pred, cfg).numpy(), g("proposals"), rtol=0, atol=2e-6 * scale), i
        det = O.detections_layer(y_pred, cfg).numpy()
        assert np.allclose(det[..., :5], g("detections")[..., :5], rtol=0, atol=2e-6 * scale) and np.array_equal(det[..., 5], g("detections")[..., 5]), i
        gt_norm = O.norm_boxes_graph(t(c["gt_boxes_px"]), c["S"], c["S"])
        assert np.allclose(gt_norm.numpy(), g("gt_boxes_norm"), rtol=0, atol=1e-7), i
        rois, tids, tmasks = O.detect_mask_targets(t(g("proposals")), t(c["gt_class_ids"]), t(g("gt_boxes_norm")), t(c["gt_masks"]), cfg)
        assert np.array_equal(tids.numpy(), g("target_class_ids")) and np.array_equal(rois.numpy(), g("rois")), i
        assert np.array_equal(np.packbits(tmasks.numpy().astype(np.uint8)), g("target_masks_bits")), i
        positives += int((tids.numpy() > 0).sum())
        pooled = O.pyramid_roi_align(t(g("rois")), t(c["feat"]), 14).numpy()[:, ::7]
        assert np.allclose(pooled, g("pooled_every7"), rtol=0, atol=5e-6), i
        ml = O.myolo_mask_loss_graph(tmasks, tids, t(c["pred_masks"])).item()
        assert np.isclose(ml, g("mask_loss"), rtol=2e-5, atol=1e-7), (i, ml, g("mask_loss"))
    assert positives > n                                                        # the cases exercise the positive-ROI path


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is only present in the build container")
def test_oracle_agrees_with_the_reference_graph_builders_on_random_configurations(tmp_path):
    """conv_block + mobilenet_graph, yolo_branch_graph and build_mask_graph of the reference run live (Keras-layer
    stand-ins, fp64) against the fp64 oracle for 5 random (batch, image size, anchors, classes) configurations, each with
    its own weights, in both learning phases."""
    import torch
    n = 5
    out = str(tmp_path / "ref_net_fuzz.npz")
    env = dict(os.environ, PYTHONPATH="")
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_graph_fixtures.py"), "--fuzz-net", out, str(n)],
                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(out)
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import graph_case_inputs as GI
    from oracle import myolo_oracle as O
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).double()            # noqa: E731

    def close(a, b, tol, what):
        err, scale = float(np.abs(a - b).max()), max(1.0, float(np.abs(b).max()))
        assert err <= tol * scale, (what, err, scale)

    for i in range(n):
        c = GI.net_inputs(GI.net_fuzz_case(i))
        P = {k: t(v) for k, v in GI.weights(c["NB"], c["NC"], c["seed"]).items()}
        cfg = dict(GRID_H=c["S"] // 32, GRID_W=c["S"] // 32, N_BOX=c["NB"], NUM_CLASSES=c["NC"], MASK_POOL_SIZE=14)
        for phase in (1, 0):
            c3 = O.mobilenet_graph(t(c["image"]), P, bool(phase))
            close(c3.numpy()[..., ::16], ref["%d/%d/c3" % (i, phase)], 1e-9, (i, phase, "c3"))
            close(O.yolo_branch_graph(c3, P, cfg, bool(phase)).numpy(), ref["%d/%d/yolo" % (i, phase)], 1e-9, (i, phase, "yolo"))
            masks = O.build_mask_graph(t(c["rois"]), t(c["feat"]), P, cfg, bool(phase))
            close(masks.numpy()[:, ::2, ::3, ::3], ref["%d/%d/masks" % (i, phase)], 5e-6, (i, phase, "masks"))


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is only present in the build container")
def test_shapes_dataset_and_device_raster_rule_agree_with_the_reference_dataset(tmp_path):
    """140 images of the reference's own ShapesDataset (seeded `random`, four image sizes) through its own load_image_gt,
    against (1) the package's ShapesDataset + load_image_gt and (2) the owner-per-pixel rule of the DEVICE rasteriser
    evaluated in numpy over the g++ build of csrc/shapes_extents.h -- specs, images, class ids, boxes and every mask bit."""
    out = str(tmp_path / "ref_shapes_fuzz.npz")
    env = dict(os.environ, PYTHONPATH="")
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_fixtures.py"), "--fuzz-shapes", out],
                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(out)
    import ctypes
    from myolo import myolo_utils as U
    from myolo.shapes import ShapesConfig, ShapesDataset
    so = str(tmp_path / "shapes_extents_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(os.path.dirname(HERE), "mask-yolo_b200", "csrc"),
                           os.path.join(HERE, "shapes_extents_harness.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.shape_rows_host.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    lib.shape_rows_host.restype = None
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_reference_fixtures import SHAPES_FUZZ
    n_img = n_tri = 0
    for seed, size, count in SHAPES_FUZZ:
        class Cfg(ShapesConfig):
            IMAGE_SHAPE = [size, size, 3]
            IMAGE_MIN_DIM = IMAGE_MAX_DIM = size
        ds = ShapesDataset(seed)
        ds.load_shapes(count, size, size)
        ds.prepare()
        weights = np.arange(1, size + 1)[:, None, None]
        for i in ds.image_ids:
            tag, info = "%d_%d" % (seed, i), ds.image_info[i]
            specs = np.array([[["square", "circle", "triangle"].index(s[0])] + list(s[1]) + list(s[2]) for s in info["shapes"]],
                             dtype=np.int64).reshape(-1, 7)
            assert np.array_equal(specs, ref[tag + "_specs"]) and list(info["bg_color"]) == ref[tag + "_bg"].tolist(), tag
            image, class_ids, bbox, mask = U.load_image_gt(ds, Cfg(), i, use_mini_mask=False)
            chk = lambda im: [im.astype(np.int64).sum(), (im.astype(np.int64) * weights).sum()]        # noqa: E731
            assert chk(image) == ref[tag + "_image"].tolist() and np.array_equal(class_ids, ref[tag + "_ids"]), tag
            assert np.array_equal(bbox, ref[tag + "_bbox"]) and np.array_equal(np.packbits(mask.astype(np.uint8)), ref[tag + "_mask"]), tag
            # the device rule: owner = last shape covering the pixel
            owner = np.full((size, size), -1, np.int32)
            cols = np.arange(size)[None, :]
            for k, (t, _, _, _, x, y, s) in enumerate(specs.tolist()):
                lo, hi = np.empty(size, np.int32), np.empty(size, np.int32)
                lib.shape_rows_host(t + 1, x, y, s, size, size, lo.ctypes.data, hi.ctypes.data)
                owner[(cols >= lo[:, None]) & (cols <= hi[:, None])] = k
                n_tri += t == 2
            img2 = np.empty((size, size, 3), np.uint8)
            img2[:] = np.asarray(info["bg_color"], np.uint8)
            keep = []
            for k, row in enumerate(specs.tolist()):
                img2[owner == k] = row[1:4]
                if (owner == k).any():
                    keep.append(k)
            assert chk(img2) == ref[tag + "_image"].tolist(), tag
            m2 = np.stack([owner == k for k in keep], -1) if keep else np.zeros((size, size, 0), bool)
            assert np.array_equal(np.packbits(m2.astype(np.uint8)), ref[tag + "_mask"]), tag
            assert [specs[k][0] + 1 for k in keep] == ref[tag + "_ids"].tolist(), tag
            n_img += 1
    assert n_img == 140 and n_tri > 40
