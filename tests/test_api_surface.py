"""Drop-in boundary: every public name of the reference's myolo.model / myolo.myolo_utils / myolo.config -- read
from its source by tests/golden/make_api_surface.py, committed as tests/golden/reference_api_surface.json -- exists in
this package, and every callable takes the reference's positional parameters, in the reference's order, with the
reference's defaults; extra parameters may only FOLLOW them (precision=, device=, engine=, ...)."""
import ast
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_api_surface.json")))

# defaults that deliberately differ from the reference's, with the reason (a parameter that is REQUIRED in the reference
# may be optional here without being listed: every reference call still binds)
DEFAULT_DIFFS = {
    ("myolo.model", "MaskYOLO.detect", "save_path"): "nothing is written unless a path is given (reference: './img_results/')",
    ("myolo.model", "MaskYOLO.detect", "display"): "no plotting libraries in this build (reference: True)",
    ("myolo.model", "MaskYOLO.infer_yolo", "save_path"): "nothing is written unless a path is given (reference: './img_results/')",
    ("myolo.model", "MaskYOLO.infer_yolo", "display"): "no plotting libraries in this build (reference: True)",
}


def _positional(fn):
    ps = [p for p in inspect.signature(fn).parameters.values()
          if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    return ps


def _check(where, ref_sig, fn):
    mine = _positional(fn)
    names = [p.name for p in mine]
    assert names[:len(ref_sig["args"])] == ref_sig["args"], (where, "reference", ref_sig["args"], "package", names)
    n_def = len(ref_sig["defaults"])
    ref_defaults = dict(zip(ref_sig["args"][len(ref_sig["args"]) - n_def:], ref_sig["defaults"]))
    for p in mine[:len(ref_sig["args"])]:
        key = tuple(where.split(":")) + (p.name,)
        if p.name in ref_defaults:
            assert p.default is not inspect.Parameter.empty, (where, p.name, "must stay optional")
            try:
                want = ast.literal_eval(ref_defaults[p.name])
            except Exception:
                continue                                    # a non-literal default (e.g. a name): presence is enough
            if key in DEFAULT_DIFFS:
                continue
            assert p.default == want, (where, p.name, "reference default", want, "package default", p.default)
    if ref_sig["kwarg"]:
        assert any(p.kind == p.VAR_KEYWORD for p in inspect.signature(fn).parameters.values()), (where, "**kwargs")


@pytest.mark.parametrize("mod", sorted(REF))
def test_public_names_and_signatures_match_reference(mod):
    pkg = importlib.import_module(mod)
    for name, sig in REF[mod]["functions"].items():
        assert hasattr(pkg, name), (mod, "function missing", name, "reference line", sig["line"])
        _check("%s:%s" % (mod, name), sig, getattr(pkg, name))
    for cname, c in REF[mod]["classes"].items():
        assert hasattr(pkg, cname), (mod, "class missing", cname)
        cls = getattr(pkg, cname)
        for mname, sig in c["methods"].items():
            assert hasattr(cls, mname), (mod, "method missing", cname + "." + mname, "reference line", sig["line"])
            _check("%s:%s.%s" % (mod, cname, mname), sig, getattr(cls, mname))


def test_reference_call_patterns_bind():
    """The call sites of the reference's own build() (model.py:844-936) bind against this package's signatures."""
    from myolo import model as M
    from myolo.shapes import ShapesConfig
    cfg = ShapesConfig()
    inspect.signature(M.DecodeYOLOLayer.__init__).bind(None, name='decode_yolo_layer', config=cfg)
    inspect.signature(M.DetectionsLayer.__init__).bind(None, name="decode_yolo_layer", config=cfg)
    inspect.signature(M.DetectMaskTargetLayer.__init__).bind(None, cfg, name='detect_mask_targets')
    inspect.signature(M.PyramidROIAlign.__init__).bind(None, [14, 14], name="roi_align_mask")
    inspect.signature(M.build_mask_graph).bind("rois", ["feature_map"], cfg.MASK_POOL_SIZE, cfg.NUM_CLASSES)
    inspect.signature(M.mobilenet_graph).bind("image", cfg.BACKBONE, stage5=False)
    inspect.signature(M.conv_block).bind("image", 32, strides=(2, 2))
    inspect.signature(M.yolo_branch_graph).bind("c4", cfg)
    inspect.signature(M.yolo_custom_loss).bind("y_true", "y_pred", "true_boxes")
    inspect.signature(M.MaskYOLO.__init__).bind(None, mode="training", config=cfg, model_dir=None)
    inspect.signature(M.MaskYOLO.detect).bind(None, "image", "weights.h5", None, 0.35, False)
    # host-side layer objects can be built without a device
    d = M.DecodeYOLOLayer(name='decode_yolo_layer', config=cfg)
    assert d.name == 'decode_yolo_layer' and d.compute_output_shape((None, 7, 7, 3, 9)) == (None, 147, 4)
    t = M.DetectMaskTargetLayer(cfg, name='detect_mask_targets')
    assert t.compute_mask(None) == [None] * 4
    assert t.compute_output_shape([(None, 147, 4)]) == [(None, 147, 4), (None, 147), (None, None), (None, 147, 28, 28)]
    with pytest.raises(RuntimeError, match="no engine"):
        M._CURRENT["engine"] = None
        M.mobilenet_graph("image", "mobilenet")


def test_small_exported_helpers():
    """box_refinement_graph / compute_backbone_shapes / mold_image / resize_one_image / log2_graph (off the hot path; the
    reference exports them)."""
    import numpy as np
    import torch
    from myolo import model as M
    from myolo import myolo_utils as U
    from myolo.config import Config
    box = torch.tensor([[0.0, 0.0, 2.0, 4.0]])
    gt = torch.tensor([[1.0, 1.0, 5.0, 3.0]])
    r = U.box_refinement_graph(box, gt)
    assert torch.allclose(r, torch.tensor([[(2.0 - 2.0) / 4.0, (3.0 - 1.0) / 2.0, np.log(2.0 / 4.0), np.log(4.0 / 2.0)]], dtype=torch.float32))
    assert U.compute_backbone_shapes(Config(), [224, 224, 3]).tolist() == [28, 28]

    class WithMean(Config):
        MEAN_PIXEL = np.array([1.0, 2.0, 3.0])
    assert U.mold_image(np.ones((2, 2, 3), np.uint8), WithMean()).tolist() == [[[0.0, -1.0, -2.0]] * 2] * 2
    with pytest.raises(AttributeError):
        U.mold_image(np.ones((2, 2, 3), np.uint8), Config())
    gt_box = [10, 20, 30, 40]
    assert U.resize_one_image(np.zeros((100, 100, 3), np.uint8), gt_box, None, (50, 50)) is None
    assert gt_box == [5, 10, 10, 20]                         # index 2 from gt_box[1], as at HEAD
    assert torch.allclose(M.log2_graph(torch.tensor([8.0])), torch.tensor([3.0]))


def test_detect_results_answer_to_both_key_sets():
    """detect() results: the list-of-one-dict the reference's code returns and the keys its docstring documents."""
    import numpy as np
    from myolo.model import DetectResults
    r = DetectResults(np.zeros((2, 4), np.int32), np.array([1, 2]), np.array([0.9, 0.8]), np.zeros((8, 8, 2), bool))
    assert len(r) == 1 and set(r[0]) == {"bboxes", "class_ids", "confidence_scores", "full_masks"}
    assert r["rois"] is r[0]["bboxes"] and r["scores"] is r[0]["confidence_scores"] and r["masks"] is r[0]["full_masks"]
    assert r["class_ids"].tolist() == [1, 2] and [d["full_masks"].shape for d in r] == [(8, 8, 2)]
