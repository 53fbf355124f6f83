"""Golden vectors from the reference's OWN graph-building source (myolo/model.py), executed eagerly.

Run HERE (the container that has /root/reference); the .npz it writes is committed, tests never read /root/reference:

    python tests/golden/make_reference_graph_fixtures.py

TensorFlow 1.x / Keras 2.x are not installable, so `tensorflow`, `keras.backend` and the other third-party imports of
myolo/model.py are replaced by tests/golden/tf1_numpy_shim.py (primitive ops over numpy float32, restated from their
documented behaviour) and empty stubs.  The functions below then run UNMODIFIED from /root/reference/myolo/model.py:

    yolo_custom_loss 86-242 (both the normal and the warm-up branch)      DecodeYOLOLayer.call 1442-1473
    DetectionsLayer.call 1493-1538        norm_boxes_graph 1394-1408      trim_zeros_graph 1411-1420
    overlaps_graph 420-454                detect_mask_target_graph 457-602 through DetectMaskTargetLayer.call 635-649
    PyramidROIAlign.call 327-410          myolo_mask_loss_graph 718-754

Two things are patched, both documented reference defects: (1) model.py reads its configuration from the module
global `config` = the base Config CLASS (model.py:25), so the case's values are set as class attributes; (2)
DetectMaskTargetLayer.call uses the undefined name `utils` (644) -- it is bound to the reference's own
myolo.myolo_utils, where batch_slice (929-963) lives."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import graph_case_inputs as GI          # noqa: E402
import keras2_layers_shim as kls        # noqa: E402
import tf1_numpy_shim as tfs            # noqa: E402

PRODUCT = os.path.join(os.path.dirname(os.path.dirname(HERE)), "mask-yolo_b200")

OUT = os.path.join(HERE, "reference_graph_fixture.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LooseVersion(object):
    def __init__(self, s):
        self.v = tuple(int(p) for p in str(s).split(".")[:3] if p.isdigit())

    def __ge__(self, o):
        return self.v >= o.v


class _Any(object):
    def __init__(self, *a, **k):
        pass


def load_reference_model():
    sys.modules["tensorflow"] = tfs
    kb = _stub("keras.backend", shape=tfs.shape, reshape=tfs.k_reshape, switch=tfs.k_switch, mean=tfs.k_mean,
               binary_crossentropy=tfs.k_binary_crossentropy)
    layers = _stub("keras.layers", **{k: getattr(kls, k) for k in ("ZeroPadding2D", "Conv2D", "DepthwiseConv2D", "Conv2DTranspose",
                                                                   "BatchNormalization", "Activation", "Reshape", "TimeDistributed",
                                                                   "Input", "Lambda")})
    _stub("keras", __version__="2.2.4", backend=kb, engine=_stub("keras.engine", Layer=kls.Layer),
          layers=layers, models=_stub("keras.models", Model=kls.Model), utils=_stub("keras.utils", Sequence=object))
    _stub("keras_applications", get_keras_submodule=lambda name: {"backend": kls.backend}[name],
          mobilenet=_stub("keras_applications.mobilenet", _depthwise_conv_block=kls.depthwise_conv_block),
          mobilenet_v2=_stub("keras_applications.mobilenet_v2", MobileNetV2=None))
    _stub("pytz", timezone=lambda name: None)
    _stub("distutils", version=_stub("distutils.version", LooseVersion=_LooseVersion))
    _stub("matplotlib", pyplot=_stub("matplotlib.pyplot"), patches=_stub("matplotlib.patches", Rectangle=_Any))
    _stub("mrcnn", utils=_stub("mrcnn.utils", Dataset=_Any))
    sk = _stub("skimage", __version__="0.0")
    for sub in ("color", "io", "transform"):
        setattr(sk, sub, _stub("skimage." + sub))
    _stub("imgaug", augmenters=_stub("imgaug.augmenters"))
    if not hasattr(np, "bool"):
        np.bool = bool
    for k in [k for k in sys.modules if k == "myolo" or k.startswith("myolo.")]:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    try:
        import myolo                                              # the reference's (namespace) package
        assert all(p.startswith("/root/reference") for p in myolo.__path__), list(myolo.__path__)
        sys.modules["myolo.visualize"] = types.ModuleType("myolo.visualize")    # matplotlib / IPython plotting only
        myolo.visualize = sys.modules["myolo.visualize"]
        import myolo.model as ref_model
        import myolo.myolo_utils as ref_utils
    finally:
        sys.path.remove("/root/reference")
    assert ref_model.__file__.startswith("/root/reference/")
    ref_model.utils = ref_utils                                   # model.py:644 uses the undefined name `utils`
    return ref_model


def set_config(ref_model, c, warmup=0):
    cfg = ref_model.config                                        # the Config CLASS (model.py:25)
    cfg.BATCH_SIZE, cfg.GRID_H, cfg.GRID_W, cfg.N_BOX = c["B"], c["G"], c["G"], c["NB"]
    cfg.NUM_CLASSES, cfg.ANCHORS, cfg.TRUE_BOX_BUFFER = c["NC"], list(c["ANCHORS"]), c["TB"]
    cfg.CLASS_WEIGHTS = np.asarray(c["CLASS_WEIGHTS"], dtype="float32")
    for k, d in (("OBJECT_SCALE", 5.0), ("NO_OBJECT_SCALE", 1.0), ("COORD_SCALE", 1.0), ("CLASS_SCALE", 1.0)):
        setattr(cfg, k, c.get(k, d))
    cfg.WARM_UP_BATCHES = warmup
    cfg.TRAIN_ROIS_PER_IMAGE = c["R"]
    cfg.MASK_SHAPE, cfg.USE_MINI_MASK, cfg.MAX_GT_INSTANCES = [28, 28], False, c["M"]
    return cfg()


def product_weights():
    """Variable names / shapes come from the product's param_specs; it lives in a package that is also called `myolo`, so
    the weights are drawn BEFORE the reference's package is imported and the product's modules are dropped again."""
    sys.path.insert(0, PRODUCT)
    try:
        c = GI.NET
        return GI.weights(c["NB"], c["NC"], c["seed"])
    finally:
        sys.path.remove(PRODUCT)
        for k in [k for k in sys.modules if k == "myolo" or k.startswith("myolo.") or k == "mrcnn" or k.startswith("mrcnn.")]:
            del sys.modules[k]


def net_cases(M, out, W):
    """conv_block + mobilenet_graph (42-79), yolo_branch_graph (249-278), build_mask_graph (668-715) from the reference's
    source, in both learning phases.  Stored: strided samples of the activations (the tests regenerate inputs and weights
    from seeds) and, per phase, which statistics every BatchNormalization used."""
    c = GI.net_inputs()
    kls.WEIGHTS.clear()
    kls.WEIGHTS.update(W)

    class Cfg(object):
        N_BOX, NUM_CLASSES, GRID_H, GRID_W = c["NB"], c["NC"], c["S"] // 32, c["S"] // 32

    for phase in (1, 0):
        kls.STATE["learning_phase"] = phase
        del kls.USED[:]
        c3 = M.mobilenet_graph(tfs.T(c["image"]), "mobilenet")
        yolo = M.yolo_branch_graph(c3, Cfg())
        n_backbone_bn = len(kls.USED)
        masks = M.build_mask_graph(tfs.T(c["rois"]), [tfs.T(c["feat"])], 14, c["NC"], train_bn=False)
        tag = "net/phase%d/" % phase
        assert c3.a.shape == (c["B"], c["S"] // 8, c["S"] // 8, 512) and yolo.a.shape == (c["B"], 2, 2, c["NB"], 5 + c["NC"])
        assert masks.a.shape == (c["B"], c["R"], 28, 28, c["NC"])
        out[tag + "c3_every8"] = c3.a[..., ::8]
        out[tag + "yolo"] = yolo.a
        out[tag + "masks_sub"] = masks.a[:, ::3, ::2, ::2]
        out[tag + "bn_names"] = np.asarray([n for n, _ in kls.USED])
        out[tag + "bn_batch_stats"] = np.asarray([k == "batch" for _, k in kls.USED])
        print(tag, "c3 |max|", np.abs(c3.a).max(), "yolo |max|", np.abs(yolo.a).max(), "masks mean", masks.a.mean(),
              "BN layers", n_backbone_bn, "+", len(kls.USED) - n_backbone_bn,
              "mask BNs on batch statistics:", [n for n, k in kls.USED[n_backbone_bn:] if k == "batch"])


def build_cases(M, out, W):
    """MaskYOLO(mode, config) -> MaskYOLO.build (761-941) from the reference's source: the training graph (six outputs) in
    learning phase 1 and the inference graph (three outputs) in phase 0.  Ground-truth boxes are taken from the model's
    own proposals so that positive ROIs exist and the mask loss is exercised; they are stored (the test regenerates the
    image, the weights and everything derived from the boxes)."""
    c = GI.build_image()
    kls.WEIGHTS.clear()
    kls.WEIGHTS.update(W)
    base = dict(c, CLASS_WEIGHTS=[1.0] * c["NC"])
    cfg = set_config(M, base)
    cfg.IMAGE_SHAPE, cfg.BACKBONE, cfg.TOP_FEATURE_MAP_DEPTH, cfg.SECOND_PHASE_YOLO_DEPTH = [c["S"], c["S"], 3], "mobilenet", 256, 512
    cfg.MASK_POOL_SIZE = 14
    B, S, TB = c["B"], c["S"], c["TB"]
    # proposals of the untrained network -> ground truth that overlaps some of them
    kls.STATE["learning_phase"] = 1
    c4 = M.mobilenet_graph(tfs.T(c["image"]), "mobilenet")
    props = M.DecodeYOLOLayer(config=cfg).call([M.yolo_branch_graph(c4, cfg)]).a
    ids = np.zeros((B, TB), np.int32)
    boxes = np.zeros((B, TB, 4), np.float32)
    rs = np.random.RandomState(c["seed"] + 2)
    for b in range(B):
        k = 0
        for r in rs.permutation(props.shape[1]):
            px = np.round(props[b, r] * (S - 1) + np.array([0, 0, 1, 1])).clip(0, S)
            if px[2] - px[0] >= 8 and px[3] - px[1] >= 8 and k < 3:
                ids[b, k], boxes[b, k] = int(rs.randint(1, c["NC"])), px
                k += 1
        assert k >= 2, "untrained proposals too degenerate for this seed"
    masks, y_true, true_boxes = GI.gt_from_boxes(c, ids, boxes)
    out["build/gt_class_ids"], out["build/gt_boxes_px"] = ids, boxes
    kls.FEEDS.clear()
    kls.FEEDS.update(input_image=c["image"], input_true_boxes=true_boxes, input_yolo_target=y_true, input_gt_class_ids=ids,
                     input_gt_boxes=boxes, input_gt_masks=masks, input_yolo_feature_map=c4.a)
    model = M.MaskYOLO(mode="training", config=cfg).keras_model
    names = ["yolo_output", "yolo_proposals", "output_rois", "myolo_mask", "yolo_sum_loss", "mask_loss"]
    assert model.name == "mask+yolo" and len(model.outputs) == 6
    for n, t in zip(names, model.outputs):
        out["build/training/" + n] = np.asarray(t.a)[:, ::2, ::3, ::3] if n == "myolo_mask" else np.asarray(t.a)
    assert (np.abs(out["build/training/output_rois"]).sum(-1) > 0).any() and out["build/training/mask_loss"] > 0
    print("build/training: yolo_sum_loss", out["build/training/yolo_sum_loss"], "mask_loss", out["build/training/mask_loss"])
    model = M.MaskYOLO(mode="yolo", config=cfg).keras_model          # same feeds, learning phase still 1
    assert model.name == "only_yolo" and len(model.outputs) == 2
    out["build/yolo/yolo_output"], out["build/yolo/yolo_sum_loss"] = np.asarray(model.outputs[0].a), np.asarray(model.outputs[1].a)
    kls.STATE["learning_phase"] = 0
    c4 = M.mobilenet_graph(tfs.T(c["image"]), "mobilenet")
    kls.FEEDS["input_yolo_feature_map"] = c4.a
    model = M.MaskYOLO(mode="inference", config=cfg).keras_model
    assert model.name == "mask_yolo_inference" and len(model.outputs) == 3
    for n, t in zip(["yolo_output", "detections", "myolo_mask"], model.outputs):
        out["build/inference/" + n] = np.asarray(t.a)[:, ::2, ::3, ::3] if n == "myolo_mask" else np.asarray(t.a)
    print("build/inference: detections", out["build/inference/detections"].shape, "masks", out["build/inference/myolo_mask"].shape)


def decode_masks_case(M, out):
    """MaskYOLO.decode_masks (1330-1391) and myolo_utils.unmold_mask (883-912) from the reference's source.  The one
    primitive underneath, the reference's `resize` wrapper around scikit-image (absent), is replaced by the cv2 bilinear
    resize the package uses -- what is pinned is everything around it: class-specific mask selection, the zero-area
    filter, int() truncation and clamping of the normalised corners, resizing into the CLIPPED box, threshold, paste."""
    import cv2
    c = GI.decode_masks_inputs()
    refu = M.mutils
    assert refu.__file__.startswith("/root/reference/")
    refu.resize = lambda image, output_shape, **kw: cv2.resize(np.asarray(image, dtype=np.float32),
                                                               (int(output_shape[1]), int(output_shape[0])),
                                                               interpolation=cv2.INTER_LINEAR)
    M.config.IMAGE_SHAPE = [c["S"], c["S"], 3]
    obj = M.MaskYOLO.__new__(M.MaskYOLO)
    boxes, class_ids, scores, full = M.MaskYOLO.decode_masks(obj, c["detections"], c["myolo_mask"], (c["S"], c["S"], 3))
    assert boxes.shape[0] == c["N"] - 2 and full.shape == (c["S"], c["S"], c["N"] - 2) and full.any()
    out["decode_masks/boxes"], out["decode_masks/class_ids"], out["decode_masks/scores"] = boxes, class_ids, scores
    out["decode_masks/full_bits"], out["decode_masks/full_shape"] = np.packbits(full.astype(np.uint8)), np.asarray(full.shape)
    print("decode_masks:", boxes.shape[0], "kept of", c["N"], "mask pixels", int(full.sum()))


def main():
    W = product_weights()
    M = load_reference_model()
    T = tfs.T
    out = {}
    for name in GI.CASES:
        c = GI.build(name)
        cfg = set_config(M, c)
        y_pred, y_true, tb = T(c["y_pred"]), T(c["y_true"]), T(c["true_boxes"])
        out[name + "/yolo_loss"] = M.yolo_custom_loss(y_true, y_pred, tb).a
        set_config(M, c, warmup=3)                                # seen = 1 after assign_add < 3: warm-up branch (196-207)
        out[name + "/yolo_loss_warmup"] = M.yolo_custom_loss(y_true, y_pred, tb).a
        cfg = set_config(M, c)
        props = M.DecodeYOLOLayer(config=cfg).call([y_pred])
        out[name + "/proposals"] = props.a
        out[name + "/detections"] = M.DetectionsLayer(config=cfg).call([y_pred]).a
        S = c["S"]
        gt_norm = M.norm_boxes_graph(T(c["gt_boxes_px"]), T(np.asarray([S, S], np.int32)))
        out[name + "/gt_boxes_norm"] = gt_norm.a
        trimmed, nz = M.trim_zeros_graph(gt_norm[0])
        out[name + "/trim_nonzero_0"] = nz.a
        out[name + "/overlaps_0"] = M.overlaps_graph(props[0], trimmed).a
        rois, tids, _, tmasks = M.DetectMaskTargetLayer(cfg).call([props, T(c["gt_class_ids"]), gt_norm, T(c["gt_masks"])])
        out[name + "/rois"] = rois.a
        out[name + "/target_class_ids"] = tids.a
        assert set(np.unique(tmasks.a)) <= {0.0, 1.0}
        out[name + "/target_masks_bits"] = np.packbits(tmasks.a.astype(np.uint8))
        out[name + "/target_masks_shape"] = np.asarray(tmasks.a.shape)
        assert (tids.a > 0).sum() >= 2, "case has too few positive ROIs to be useful"
        pooled = M.PyramidROIAlign([14, 14]).call([rois, T(c["feat"])])
        assert pooled.a.shape == (c["B"], c["R"], 14, 14, c["C"])
        out[name + "/pooled_every5"] = pooled.a[:, ::5]
        out[name + "/mask_loss"] = M.myolo_mask_loss_graph(tmasks, tids, T(c["pred_masks"])).a
        out[name + "/mask_loss_no_positives"] = M.myolo_mask_loss_graph(tmasks, T(np.zeros_like(tids.a)), T(c["pred_masks"])).a
        print(name, "yolo_loss", out[name + "/yolo_loss"], "warm-up", out[name + "/yolo_loss_warmup"], "positives",
              (tids.a > 0).sum(1), "mask_loss", out[name + "/mask_loss"])
    net_cases(M, out, W)
    build_cases(M, out, W)
    decode_masks_case(M, out)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


def fuzz(path, n):
    """Reference side of the live differential test of the graph functions: n random configurations (grid, anchors,
    classes, buffer width, loss scales, class weights, warm-up on/off, some images without ground truth)."""
    M = load_reference_model()
    T = tfs.T
    out = {}
    for i in range(n):
        c = GI.build(GI.fuzz_case(i))
        warm = 3 if i % 3 == 1 else 0
        cfg = set_config(M, c, warmup=warm)
        y_pred, y_true, tb = T(c["y_pred"]), T(c["y_true"]), T(c["true_boxes"])
        out["%d/yolo_loss" % i] = M.yolo_custom_loss(y_true, y_pred, tb).a
        props = M.DecodeYOLOLayer(config=cfg).call([y_pred])
        out["%d/proposals" % i] = props.a
        out["%d/detections" % i] = M.DetectionsLayer(config=cfg).call([y_pred]).a
        gt_norm = M.norm_boxes_graph(T(c["gt_boxes_px"]), T(np.asarray([c["S"], c["S"]], np.int32)))
        out["%d/gt_boxes_norm" % i] = gt_norm.a
        rois, tids, _, tmasks = M.DetectMaskTargetLayer(cfg).call([props, T(c["gt_class_ids"]), gt_norm, T(c["gt_masks"])])
        out["%d/rois" % i], out["%d/target_class_ids" % i] = rois.a, tids.a
        out["%d/target_masks_bits" % i] = np.packbits(tmasks.a.astype(np.uint8))
        out["%d/pooled_every7" % i] = M.PyramidROIAlign([14, 14]).call([rois, T(c["feat"])]).a[:, ::7]
        out["%d/mask_loss" % i] = M.myolo_mask_loss_graph(tmasks, tids, T(c["pred_masks"])).a
    np.savez_compressed(path, **out)


def fuzz_net(path, n):
    """Reference side of the live differential test of the graph BUILDERS: n random (batch, image size, anchors, classes)
    configurations with their own weights, both learning phases."""
    cases = [GI.net_fuzz_case(i) for i in range(n)]
    sys.path.insert(0, PRODUCT)
    try:
        weights = [GI.weights(c["NB"], c["NC"], c["seed"]) for c in cases]      # before the reference's `myolo` is imported
    finally:
        sys.path.remove(PRODUCT)
        for k in [k for k in sys.modules if k == "myolo" or k.startswith("myolo.") or k == "mrcnn" or k.startswith("mrcnn.")]:
            del sys.modules[k]
    M = load_reference_model()
    out = {}
    for i, (case, W) in enumerate(zip(cases, weights)):
        c = GI.net_inputs(case)
        kls.WEIGHTS.clear()
        kls.WEIGHTS.update(W)

        class Cfg(object):
            N_BOX, NUM_CLASSES, GRID_H, GRID_W = c["NB"], c["NC"], c["S"] // 32, c["S"] // 32

        for phase in (1, 0):
            kls.STATE["learning_phase"] = phase
            c3 = M.mobilenet_graph(tfs.T(c["image"]), "mobilenet")
            out["%d/%d/c3" % (i, phase)] = c3.a[..., ::16]
            out["%d/%d/yolo" % (i, phase)] = M.yolo_branch_graph(c3, Cfg()).a
            out["%d/%d/masks" % (i, phase)] = M.build_mask_graph(tfs.T(c["rois"]), [tfs.T(c["feat"])], 14, c["NC"], train_bn=False).a[:, ::2, ::3, ::3]
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--fuzz-net":
        fuzz_net(sys.argv[2], int(sys.argv[3]))
    elif len(sys.argv) == 4 and sys.argv[1] == "--fuzz":
        fuzz(sys.argv[2], int(sys.argv[3]))
    else:
        main()
