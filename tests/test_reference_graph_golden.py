"""The oracle (and, on the GPU, the device kernels) against golden vectors produced by the REFERENCE's own
myolo/model.py source, executed eagerly over a numpy stand-in for its TensorFlow/Keras primitives
(tests/golden/make_reference_graph_fixtures.py + tf1_numpy_shim.py; the vectors are committed, inputs are regenerated
from seeds by tests/golden/graph_case_inputs.py).  This pins the reference's formulas -- YOLO loss incl. the warm-up
branch, box decoding, detections, box normalisation, IoU, positive/negative ROI selection and ordering, class / mask
targets, the x/y-swapped ROIAlign call, the mask loss -- to its code rather than to a restatement.

Tolerances: float32 arithmetic in a different association order -> 2e-6 absolute on O(1) coordinates, 1e-5 relative on
the reduced losses; everything integer / boolean is exact."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import graph_case_inputs as GI          # noqa: E402

from oracle import myolo_oracle as O    # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_graph_fixture.npz"))


def _ocfg(c, warmup=0):
    return dict(GRID_H=c["G"], GRID_W=c["G"], N_BOX=c["NB"], NUM_CLASSES=c["NC"], ANCHORS=c["ANCHORS"],
                TRAIN_ROIS_PER_IMAGE=c["R"], MASK_SHAPE=[28, 28], MASK_POOL_SIZE=14, COORD_SCALE=c.get("COORD_SCALE", 1.0),
                NO_OBJECT_SCALE=c.get("NO_OBJECT_SCALE", 1.0), OBJECT_SCALE=c.get("OBJECT_SCALE", 5.0),
                CLASS_SCALE=c.get("CLASS_SCALE", 1.0), CLASS_WEIGHTS=np.asarray(c["CLASS_WEIGHTS"], np.float32),
                WARM_UP_BATCHES=warmup, TRUE_BOX_BUFFER=c["TB"], LOSS_WEIGHTS={"yolo_sum_loss": 1.0, "myolo_mask_loss": 1.0})


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("name", list(GI.CASES))
def test_oracle_equals_reference_source(gold, name):
    c = GI.build(name)
    g = lambda k: gold[name + "/" + k]                                             # noqa: E731
    cfg = _ocfg(c)
    y_pred, y_true, tb = _t(c["y_pred"]), _t(c["y_true"]), _t(c["true_boxes"])
    # a11 yolo_custom_loss, normal and warm-up branch
    assert np.isclose(O.yolo_custom_loss(y_true, y_pred, tb, cfg).item(), g("yolo_loss"), rtol=1e-5)
    assert np.isclose(O.yolo_custom_loss(y_true, y_pred, tb, _ocfg(c, warmup=3), seen=1.0).item(), g("yolo_loss_warmup"), rtol=1e-5)
    # a5 / a12 decode + detections
    props = O.decode_yolo(y_pred, cfg)
    assert np.allclose(props.numpy(), g("proposals"), rtol=0, atol=2e-6 * max(1.0, np.abs(g("proposals")).max()))
    det = O.detections_layer(y_pred, cfg).numpy()
    assert np.allclose(det[..., :5], g("detections")[..., :5], rtol=0, atol=2e-6 * max(1.0, np.abs(g("proposals")).max()))
    assert np.array_equal(det[..., 5], g("detections")[..., 5])
    # a6 norm_boxes_graph, trim_zeros_graph, overlaps_graph
    S = c["S"]
    gt_norm = O.norm_boxes_graph(_t(c["gt_boxes_px"]), S, S)
    assert np.allclose(gt_norm.numpy(), g("gt_boxes_norm"), rtol=0, atol=1e-7)
    nz = g("trim_nonzero_0")
    assert np.array_equal((gt_norm[0].abs().sum(1) != 0).numpy(), nz)
    ov = O.overlaps_graph(_t(g("proposals")[0]), _t(g("gt_boxes_norm")[0][nz]))
    assert np.allclose(ov.numpy(), g("overlaps_0"), rtol=0, atol=2e-6, equal_nan=True)
    # a7 DetectMaskTargetLayer on the reference's own proposals / normalised boxes (identical inputs -> exact selection)
    rois, tids, tmasks = O.detect_mask_targets(_t(g("proposals")), _t(c["gt_class_ids"]), _t(g("gt_boxes_norm")),
                                               _t(c["gt_masks"]), cfg)
    assert np.array_equal(tids.numpy(), g("target_class_ids")) and (tids.numpy() > 0).sum() >= 2
    assert np.array_equal(rois.numpy(), g("rois"))                                # same rows in the same order, bit for bit
    ref_masks = np.unpackbits(g("target_masks_bits"))[:int(np.prod(g("target_masks_shape")))].reshape(g("target_masks_shape"))
    assert np.array_equal(tmasks.numpy().astype(np.uint8), ref_masks)
    # a8 PyramidROIAlign (boxes handed over as x1,y1,x2,y2: the reference's swapped sampling)
    pooled = O.pyramid_roi_align(_t(g("rois")), _t(c["feat"]), 14).numpy()[:, ::5]
    assert pooled.shape == g("pooled_every5").shape
    assert np.allclose(pooled, g("pooled_every5"), rtol=0, atol=5e-6)
    # a10 myolo_mask_loss_graph
    tm = _t(ref_masks.astype(np.float32))
    ml = O.myolo_mask_loss_graph(tm, _t(g("target_class_ids")), _t(c["pred_masks"]))
    assert np.isclose(ml.item(), g("mask_loss"), rtol=1e-5)
    assert O.myolo_mask_loss_graph(tm, torch.zeros_like(_t(g("target_class_ids"))), _t(c["pred_masks"])).item() == g("mask_loss_no_positives") == 0.0


@pytest.mark.parametrize("phase", [1, 0])
def test_oracle_networks_equal_reference_builders(gold, phase):
    """a1-a4 and a9: conv_block + mobilenet_graph, yolo_branch_graph and build_mask_graph as the reference's own source
    wires them (Keras layers restated by tests/golden/keras2_layers_shim.py), learning phase 1 and 0.  Also pins which
    BatchNormalization layers follow the learning phase: all 29 of the backbone / YOLO branch and, in the mask head,
    ONLY myolo_mask_bn1 (bn2-4 are called with training=False, model.py:695-708)."""
    c = GI.net_inputs()
    P = {k: _t(v).double() for k, v in GI.weights(c["NB"], c["NC"], c["seed"]).items()}     # fp64 oracle vs fp64 layers
    tag = "net/phase%d/" % phase
    training = bool(phase)
    cfg = dict(GRID_H=c["S"] // 32, GRID_W=c["S"] // 32, N_BOX=c["NB"], NUM_CLASSES=c["NC"], MASK_POOL_SIZE=14)
    rec = O._BNRec()
    c3 = O.mobilenet_graph(_t(c["image"]).double(), P, training, rec)
    yolo = O.yolo_branch_graph(c3, P, cfg, training, rec)
    masks = O.build_mask_graph(_t(c["rois"]).double(), _t(c["feat"]).double(), P, cfg, training, rec)

    def close(a, ref, tol, what):
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(a - ref).max())
        assert err <= tol * scale, (what, err, scale)

    close(c3.numpy()[..., ::8], gold[tag + "c3_every8"], 1e-9, "backbone feature map")
    close(yolo.numpy(), gold[tag + "yolo"], 1e-9, "yolo branch output")
    # the ROIAlign in front of the mask head runs in float32 in the reference's graph (tf.image.crop_and_resize)
    close(masks.numpy()[:, ::3, ::2, ::2], gold[tag + "masks_sub"], 5e-6, "mask head output")
    names, batch = [str(n) for n in gold[tag + "bn_names"]], gold[tag + "bn_batch_stats"]
    assert len(names) == 33 and names[:2] == ["conv1_bn", "conv_dw_1_bn"] and names[-4:] == ["myolo_mask_bn%d" % i for i in (1, 2, 3, 4)]
    assert [n for n, b in zip(names, batch) if b] == [n for n, _, _ in rec.items]        # same layers, same order
    if training:
        assert batch[:29].all() and batch[29:].tolist() == [True, False, False, False]
    else:
        assert not batch.any()


def test_oracle_whole_model_equals_reference_build(gold):
    """MaskYOLO(mode, config).build (761-941) from the reference's source, training graph (six outputs, learning phase 1)
    and inference graph (three outputs, phase 0), against oracle.forward_training / forward_inference in fp64: same
    weights, same image, ground truth taken from the fixture."""
    c = GI.build_image()
    P = {k: _t(v).double() for k, v in GI.weights(GI.NET["NB"], GI.NET["NC"], GI.NET["seed"]).items()}
    ids, boxes = gold["build/gt_class_ids"], gold["build/gt_boxes_px"]
    masks, y_true, true_boxes = GI.gt_from_boxes(c, ids, boxes)
    cfg = _ocfg(dict(c, CLASS_WEIGHTS=[1.0] * c["NC"]))
    image = _t(c["image"]).double()
    out = O.forward_training(P, [image, _t(true_boxes).double(), _t(y_true).double(), _t(ids), _t(boxes).double(), _t(masks)], cfg)

    def close(a, ref, tol, what):
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(np.asarray(a) - ref).max())
        assert err <= tol * scale, (what, err, scale)

    g = lambda k: gold["build/training/" + k]                                     # noqa: E731
    close(out["yolo_output"].numpy(), g("yolo_output"), 1e-9, "yolo_output")
    close(out["yolo_proposals"].numpy(), g("yolo_proposals"), 1e-6, "yolo_proposals")       # float32 cell grid / anchors in the graph
    close(out["output_rois"].numpy(), g("output_rois"), 1e-6, "output_rois")
    assert np.array_equal(np.abs(out["output_rois"].numpy()).sum(-1) > 0, np.abs(g("output_rois")).sum(-1) > 0)
    assert (out["target_class_ids"].numpy() > 0).sum() >= 2
    close(out["myolo_mask"].numpy()[:, ::2, ::3, ::3], g("myolo_mask"), 5e-6, "myolo_mask")
    assert np.isclose(out["yolo_sum_loss"].item(), g("yolo_sum_loss"), rtol=1e-6)
    assert np.isclose(out["mask_loss"].item(), g("mask_loss"), rtol=1e-5)
    # mode 'yolo' (906-920): backbone + YOLO branch + YOLO loss only, same learning phase
    close(out["yolo_output"].numpy(), gold["build/yolo/yolo_output"], 1e-9, "yolo-mode yolo_output")
    assert np.isclose(out["yolo_sum_loss"].item(), gold["build/yolo/yolo_sum_loss"], rtol=1e-6)
    inf = O.forward_inference(P, image, cfg)
    g = lambda k: gold["build/inference/" + k]                                    # noqa: E731
    close(inf["yolo_output"].numpy(), g("yolo_output"), 1e-9, "inference yolo_output")
    close(inf["detections"].numpy()[..., :5], g("detections")[..., :5], 1e-6, "detections")
    assert np.array_equal(inf["detections"].numpy()[..., 5], g("detections")[..., 5])
    close(inf["myolo_mask"].numpy()[:, ::2, ::3, ::3], g("myolo_mask"), 5e-6, "inference myolo_mask")


@pytest.mark.parametrize("name", list(GI.CASES))
def test_package_tensor_helpers_equal_reference_source(gold, name):
    """The package's own norm_boxes_graph / trim_zeros_graph / overlaps_graph (plain tensor functions of myolo.model,
    usable on any device) against the reference-source vectors."""
    from myolo import model as M
    c = GI.build(name)
    g = lambda k: gold[name + "/" + k]                                             # noqa: E731
    S = c["S"]
    gt_norm = M.norm_boxes_graph(_t(c["gt_boxes_px"]), (S, S))
    assert np.allclose(gt_norm.numpy(), g("gt_boxes_norm"), rtol=0, atol=1e-7)
    boxes, keep = M.trim_zeros_graph(gt_norm[0])
    assert np.array_equal(keep.numpy(), g("trim_nonzero_0")) and boxes.shape[0] == int(g("trim_nonzero_0").sum())
    ov = M.overlaps_graph(_t(g("proposals")[0]), _t(g("gt_boxes_norm")[0][g("trim_nonzero_0")]))
    assert np.allclose(ov.numpy(), g("overlaps_0"), rtol=0, atol=2e-6, equal_nan=True)


def test_decode_masks_and_unmold_mask_equal_reference_source(gold):
    """MaskYOLO.decode_masks (model.py:1330-1391) / unmold_mask (myolo_utils.py:883-912) of the package against the
    reference's own code: kept detections, order, class ids, scores and every pixel of the pasted full-size masks (boxes
    that leave the image, a zero-area and an inverted box included).  Both sides resize with the same cv2 call."""
    from myolo.model import MaskYOLO
    from myolo.shapes import ShapesConfig
    c = GI.decode_masks_inputs()

    class Cfg(ShapesConfig):
        IMAGE_SHAPE = [c["S"], c["S"], 3]

    m = MaskYOLO.__new__(MaskYOLO)
    m.config = Cfg()
    boxes, class_ids, scores, full = m.decode_masks(c["detections"], c["myolo_mask"], (c["S"], c["S"], 3))
    assert np.array_equal(boxes, gold["decode_masks/boxes"]) and np.array_equal(scores, gold["decode_masks/scores"])
    assert np.array_equal(class_ids, gold["decode_masks/class_ids"]) and class_ids.dtype == gold["decode_masks/class_ids"].dtype
    assert list(full.shape) == gold["decode_masks/full_shape"].tolist() and full.dtype == bool
    assert np.array_equal(np.packbits(full.astype(np.uint8)), gold["decode_masks/full_bits"])
    with pytest.raises(AssertionError):
        m.decode_masks(np.concatenate([c["detections"]] * 2), c["myolo_mask"], (c["S"], c["S"], 3))


def test_shim_crop_and_resize_micro_cases():
    """The one non-trivial primitive the stand-in supplies, against hand-computed values (tf.image.crop_and_resize:
    corners map to [0, size-1], samples outside take the extrapolation value 0, crop size 1 samples the box centre)."""
    import tf1_numpy_shim as tfs
    img = np.arange(12, dtype=np.float32).reshape(1, 3, 4, 1)                      # value = 4*y + x
    full = tfs.image.crop_and_resize(img, np.array([[0, 0, 1, 1]], np.float32), np.array([0]), [3, 4]).a
    assert np.array_equal(full, img)
    mid = tfs.image.crop_and_resize(img, np.array([[0, 0, 1, 1]], np.float32), np.array([0]), [2, 2]).a[0, :, :, 0]
    assert np.array_equal(mid, [[0, 3], [8, 11]])
    one = tfs.image.crop_and_resize(img, np.array([[0.25, 0.5, 0.75, 1.0]], np.float32), np.array([0]), [1, 1]).a
    assert np.isclose(one.item(), 4 * 1.0 + 2.25)                                  # centre: y = 0.5*2 = 1, x = 0.75*3
    out = tfs.image.crop_and_resize(img, np.array([[-0.5, 0, 0.5, 1.5]], np.float32), np.array([0]), [3, 4]).a[0, :, :, 0]
    assert np.array_equal(out[0], [0, 0, 0, 0]) and out[1, 0] == 0.0 and np.array_equal(out[1:, 3], [0, 0])   # y=-1 row, x=4.5 col
    assert np.isclose(out[2, 1], 4 * 1.0 + 1.5)


# ------------------------------------------------------------------------------------------------ GPU
_GPU_CASES = list(GI.CASES)


@pytest.mark.gpu
@pytest.mark.parametrize("name", _GPU_CASES)
def test_device_kernels_equal_reference_source(gold, name):
    """The C-ABI entry points of rows a5/a7/a8/a10/a11/a12 directly against the reference-source vectors."""
    from myolo import _cabi as C
    c = GI.build(name)
    g = lambda k: gold[name + "/" + k]                                             # noqa: E731
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    B, G, NB, NC, TB, M, S, R = c["B"], c["G"], c["NB"], c["NC"], c["TB"], c["M"], c["S"], c["R"]
    y_pred = _t(c["y_pred"]).to(dev)
    anchors = torch.tensor(c["ANCHORS"], dtype=torch.float32, device=dev)
    box_tol = 5e-6 * max(1.0, np.abs(g("proposals")).max())
    # a5 / a12
    props = torch.empty(B, R, 4, device=dev)
    det = torch.empty(B, R, 6, device=dev)
    C.call("myolo_yolo_decode", y_pred, anchors, props, det, B, G, G, NB, NC, st)
    assert np.allclose(props.cpu().numpy(), g("proposals"), rtol=0, atol=box_tol)
    assert np.allclose(det.cpu().numpy()[..., :5], g("detections")[..., :5], rtol=0, atol=box_tol)
    assert np.array_equal(det.cpu().numpy()[..., 5], g("detections")[..., 5])
    # a11, normal and warm-up branch
    cw = torch.tensor(c["CLASS_WEIGHTS"], dtype=torch.float32, device=dev)
    sc = C.float_array([c.get("OBJECT_SCALE", 5.0), c.get("NO_OBJECT_SCALE", 1.0), c.get("COORD_SCALE", 1.0), c.get("CLASS_SCALE", 1.0)])
    lo = torch.empty(5, device=dev)
    ws = torch.zeros(8, dtype=torch.float64, device=dev)
    for warm, key in ((0, "yolo_loss"), (1, "yolo_loss_warmup")):
        C.call("myolo_yolo_loss", _t(c["y_true"]).to(dev), y_pred, _t(c["true_boxes"]).reshape(B, TB, 4).to(dev), anchors, cw,
               B, G, G, NB, NC, TB, sc, warm, 1.0, lo, None, ws, st)
        assert np.isclose(lo[0].item(), g(key), rtol=3e-5), (key, lo[0].item(), g(key))
    # a6 + a7 from the reference's proposals: selection, order, class ids and 28x28 targets exact
    rois = torch.empty(B, R, 4, device=dev)
    tids = torch.empty(B, R, dtype=torch.int32, device=dev)
    tmask = torch.empty(B, R, 28, 28, device=dev)
    npos = torch.empty(B, dtype=torch.int32, device=dev)
    src = torch.empty(B, R, dtype=torch.int32, device=dev)
    rgt = torch.empty(B, R, dtype=torch.int32, device=dev)
    C.call("myolo_detect_mask_targets", _t(g("proposals")).to(dev), _t(c["gt_class_ids"]).to(dev), _t(c["gt_boxes_px"]).to(dev),
           _t(c["gt_masks"].astype(np.uint8)).to(dev), B, R, TB, M, S, 28, 28, rois, tids, tmask, npos, src, rgt, st)
    assert np.array_equal(tids.cpu().numpy(), g("target_class_ids"))
    assert np.array_equal(rois.cpu().numpy(), g("rois"))
    assert np.array_equal(npos.cpu().numpy(), (g("target_class_ids") > 0).sum(1))
    ref_masks = np.unpackbits(g("target_masks_bits"))[:int(np.prod(g("target_masks_shape")))].reshape(g("target_masks_shape"))
    assert np.array_equal(tmask.cpu().numpy().astype(np.uint8), ref_masks)
    # a8: crop_and_resize over the reference's rois (x/y swapped, as the reference calls it)
    feat = _t(c["feat"]).to(dev)
    F = feat.shape[1]
    pooled = torch.empty(B * R, 14, 14, c["C"], device=dev)
    C.call("myolo_roialign_fwd", C.view(feat, B, F, F, c["C"]), rois, B * R, R, 14, C.view(pooled, B * R, 14, 14, c["C"]), 0, st)
    assert np.allclose(pooled.reshape(B, R, 14, 14, c["C"])[:, ::5].cpu().numpy(), g("pooled_every5"), rtol=0, atol=5e-6)
    # a10
    lm = torch.empty(1, device=dev)
    C.call("myolo_mask_loss", _t(c["pred_masks"]).reshape(B * R, 28, 28, NC).to(dev), tmask, tids, B * R, 28, 28, NC, 1.0, lm,
           None, torch.zeros(2, dtype=torch.float64, device=dev), st)
    assert np.isclose(lm.item(), g("mask_loss"), rtol=2e-5), (lm.item(), g("mask_loss"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GI.CASES))
def test_module_level_layer_operators_equal_reference_source(gold, name):
    """DecodeYOLOLayer / DetectionsLayer / DetectMaskTargetLayer / PyramidROIAlign / yolo_custom_loss /
    myolo_mask_loss_graph of the package, called the way the reference's build() calls them, on device tensors."""
    from myolo import model as M
    from myolo.config import Config
    c = GI.build(name)
    g = lambda k: gold[name + "/" + k]                                             # noqa: E731

    class Cfg(Config):
        BATCH_SIZE, GRID_H, GRID_W, N_BOX, NUM_CLASSES = c["B"], c["G"], c["G"], c["NB"], c["NC"]
        ANCHORS, TRUE_BOX_BUFFER, MAX_GT_INSTANCES = list(c["ANCHORS"]), c["TB"], c["M"]
        IMAGE_SHAPE = [c["S"], c["S"], 3]
        CLASS_WEIGHTS = np.asarray(c["CLASS_WEIGHTS"], dtype="float32")
        OBJECT_SCALE, NO_OBJECT_SCALE = c.get("OBJECT_SCALE", 5.0), c.get("NO_OBJECT_SCALE", 1.0)
        COORD_SCALE, CLASS_SCALE = c.get("COORD_SCALE", 1.0), c.get("CLASS_SCALE", 1.0)
        TRAIN_ROIS_PER_IMAGE = c["R"]

    cfg = Cfg()
    # the decode kernels take the grid from the tensor; resolve() derives G from IMAGE_SHAPE (S // 32), which these
    # synthetic cases do not follow, so the layers get a config whose IMAGE_SHAPE matches the grid
    class GridCfg(Cfg):
        IMAGE_SHAPE = [32 * c["G"], 32 * c["G"], 3]
    gcfg = GridCfg()
    dev = torch.device("cuda")
    y_pred = _t(c["y_pred"]).to(dev)
    tol = 5e-6 * max(1.0, np.abs(g("proposals")).max())
    props = M.DecodeYOLOLayer(name='decode_yolo_layer', config=gcfg)([y_pred])
    assert np.allclose(props.cpu().numpy(), g("proposals"), rtol=0, atol=tol)
    det = M.DetectionsLayer(name="decode_yolo_layer", config=gcfg)([y_pred])
    assert np.allclose(det.cpu().numpy()[..., :5], g("detections")[..., :5], rtol=0, atol=tol)
    assert np.array_equal(det.cpu().numpy()[..., 5], g("detections")[..., 5])
    loss = M.yolo_custom_loss(_t(c["y_true"]).to(dev), y_pred, _t(c["true_boxes"]).to(dev), gcfg)
    assert np.isclose(loss.item(), g("yolo_loss"), rtol=3e-5)
    rois, tids, _, tmask = M.DetectMaskTargetLayer(cfg, name='detect_mask_targets')(
        [_t(g("proposals")).to(dev), _t(c["gt_class_ids"]).to(dev), _t(c["gt_boxes_px"]).to(dev), _t(c["gt_masks"]).to(dev)])
    assert np.array_equal(tids.cpu().numpy(), g("target_class_ids")) and np.array_equal(rois.cpu().numpy(), g("rois"))
    pooled = M.PyramidROIAlign([14, 14], name="roi_align_mask")([rois, _t(c["feat"]).to(dev)])
    assert np.allclose(pooled[:, ::5].cpu().numpy(), g("pooled_every5"), rtol=0, atol=5e-6)
    ml = M.myolo_mask_loss_graph(tmask, tids, _t(c["pred_masks"]).to(dev))
    assert np.isclose(ml.item(), g("mask_loss"), rtol=2e-5)
