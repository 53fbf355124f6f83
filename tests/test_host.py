"""Host-side mirror of the reference interface: Config resolution, the BatchGenerator /
data_generator target encoding, box helpers, the Shapes dataset and the mrcnn shim."""
import os
import sys

import numpy as np
import pytest

from myolo import myolo_utils as mutils
from myolo.config import Config, resolve
from myolo.shapes import ShapesConfig, ShapesDataset, make_batches


def test_resolve_repairs_inconsistent_inherited_fields():
    class Broken(Config):                       # what example/shapes/dataset_shapes.py does at HEAD
        NUM_CLASSES = 4
        ANCHORS = [1.27273, 1.277385, 2.47446, 2.56253, 4.03843, 4.07434]
        BATCH_SIZE = 16
    c = resolve(Broken())
    assert (c["NB"], c["NC"], c["R"], c["G"]) == (3, 4, 147, 7)
    assert c["CLASS_WEIGHTS"].shape == (4,) and len(c["ANCHORS"]) == 6

    class Big(Config):
        IMAGE_SHAPE = [416, 416, 3]
    assert resolve(Big())["G"] == 13 and resolve(Big())["R"] == 13 * 13 * 5
    assert Config().TRAIN_ROIS_PER_IMAGE == 245 and Config.MASK_SHAPE == [28, 28]


def test_extract_bboxes_and_iou_helpers():
    m = np.zeros((10, 12, 3), bool)
    m[2:5, 3:9, 0] = True
    m[7, 11, 1] = True
    b = mutils.extract_bboxes(m)
    assert b.tolist() == [[3, 2, 9, 5], [11, 7, 12, 8], [0, 0, 0, 0]] and b.dtype == np.int32
    a, c = mutils.BoundBox(0, 0, 2, 2), mutils.BoundBox(1, 1, 3, 3)
    assert abs(mutils.bbox_iou(a, c) - 1 / 7) < 1e-12
    assert mutils.bbox_iou(a, mutils.BoundBox(5, 5, 6, 6)) == 0
    assert mutils._interval_overlap([0, 2], [1, 5]) == 1 and mutils._interval_overlap([3, 4], [0, 1]) == 0
    bx = np.array([[0, 0, 10, 10], [1, 1, 10, 10], [20, 20, 30, 30]]) / 32.0
    assert mutils.NMB(bx, np.array([1, 1, 1]), np.array([7, 3, 5]), [32, 32, 3], nms_threshold=0.5).tolist() == [7, 5]
    assert mutils.NMB(bx, np.array([1, 2, 1]), np.array([7, 3, 5]), [32, 32, 3], nms_threshold=0.5).tolist() == [7, 3, 5]


def test_batch_generator_encoding():
    cfg = ShapesConfig()
    cfg.BATCH_SIZE = 2
    S = 224
    img = np.full((S, S, 3), 128, np.uint8)
    mask = np.zeros((S, S, 1), bool)
    mask[64:128, 32:96, 0] = True                      # box (32,64,96,128): centre (64,96) px = (2.0, 3.0) cells
    info = [[img, np.array([2], np.int32), mutils.extract_bboxes(mask), mask] for _ in range(3)]
    gen = mutils.BatchGenerator(info, cfg, mode="training", shuffle=False, norm=True)
    assert len(gen) == 2 and gen.size() == 3
    (images, tb, yt, ids, boxes, masks), outs = gen[1]           # last batch is filled up from the previous image
    assert outs == [] and images.shape == (2, S, S, 3) and images.dtype == np.float32
    assert abs(images.max() - 128 / 255.) < 1e-6
    assert tb.shape == (2, 1, 1, 1, 15, 4) and yt.shape == (2, 7, 7, 3, 9)
    assert ids.shape == (2, 15) and boxes.shape == (2, 15, 4) and boxes.dtype == np.int32 and masks.dtype == bool
    cell = yt[0, 3, 2]                                   # [gy=3, gx=2]
    a = int(np.argmax(cell[:, 4]))
    assert cell[a, :5].tolist() == [2.0, 3.0, 2.0, 2.0, 1.0] and cell[a, 5 + 2] == 1 and cell[:, 4].sum() == 1
    assert a == 1                                        # 2x2-cell box best matches anchor 2.47x2.56
    assert tb[0, 0, 0, 0, 0].tolist() == [2.0, 3.0, 2.0, 2.0] and ids[0, 0] == 2
    y3 = mutils.BatchGenerator(info, cfg, mode="yolo", shuffle=False, norm=True)[0][0]
    assert len(y3) == 3


def test_shapes_dataset_and_data_generator():
    ds = ShapesDataset(seed=3)
    ds.load_shapes(6, 128, 128)
    ds.prepare()
    assert ds.num_classes == 4 and ds.class_names == ["BG", "square", "circle", "triangle"]
    img = ds.load_image(0)
    mask, cls = ds.load_mask(0)
    assert img.shape == (128, 128, 3) and img.dtype == np.uint8 and mask.shape[:2] == (128, 128) and mask.shape[2] == len(cls)
    assert ((mask.sum(-1)) <= 1).all()                  # occlusion handling: instance masks never overlap

    class C128(ShapesConfig):
        IMAGE_SHAPE = [128, 128, 3]
        GRID_H = GRID_W = 4
    g = mutils.data_generator(ds, C128(), shuffle=False, batch_size=2, norm=True)
    (images, tb, yt), _ = next(g)
    assert images.shape == (2, 128, 128, 3) and yt.shape == (2, 4, 4, 3, 9) and yt[..., 4].sum() >= 2
    b = make_batches(C128(), 1, seed=5)
    assert len(b) == 1 and b[0][5].shape == (16, 128, 128, 15)
    b2 = make_batches(C128(), 1, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(b[0], b2[0]))       # seeded -> reproducible


def test_decode_one_yolo_output_and_unmold():
    netout = np.full((2, 2, 1, 7), -8.0)
    netout[1, 0, 0] = [0, 0, 0, 0, 8.0, 9.0, -9.0]
    boxes = mutils.decode_one_yolo_output(netout, [1.0, 1.0], obj_threshold=0.3, nb_class=2)
    assert len(boxes) == 1 and boxes[0].get_label() == 0
    assert abs(boxes[0].xmin - (0.25 - 0.25)) < 1e-9 and abs(boxes[0].ymax - (0.75 + 0.25)) < 1e-9
    # unmold_mask keeps the reference's rules: normalised box, int() truncation, clamp, resize to the clipped box
    full = mutils.unmold_mask(np.ones((28, 28), np.float32), [10.9 / 64, 20.2 / 64, 30.99 / 64, 1.5], (64, 64, 3))
    assert full.sum() == 20 * 44 and full[20:64, 10:30].all()
    half = np.zeros((28, 28), np.float32)
    half[:, 14:] = 1.0
    full = mutils.unmold_mask(half, [0.0, 0.0, 2.0, 1.0], (64, 64, 3))      # x2 clamped to 64: the mask is squeezed into the clipped box
    assert full[:, 32:].all() and not full[:, :31].any()
    assert not mutils.unmold_mask(half, [1.0, 0.0, 1.0, 1.0], (64, 64, 3)).any()


def test_mrcnn_shim_and_reference_example_imports():
    from mrcnn import utils
    keep = utils.non_max_suppression(np.array([[0, 0, 10, 10], [0, 0, 9, 10], [20, 20, 30, 30]]), np.array([0.5, 0.9, 0.1]), 0.3)
    assert keep.tolist() == [1, 2]
    ex = "/root/reference/example/shapes"
    if not os.path.isdir(ex):
        pytest.skip("reference checkout not present on this box")
    sys.path.insert(0, ex)
    try:
        import dataset_shapes                       # the reference's own file, unmodified
        cfg = dataset_shapes.ShapesConfig()
        ds = dataset_shapes.ShapesDataset()
        ds.load_shapes(2, 224, 224)
        ds.prepare()
        image, cls, bbox, mask = mutils.load_image_gt(ds, cfg, 0)
        assert image.shape == (224, 224, 3) and bbox.shape[1] == 4 and mask.shape[-1] == len(cls)
        assert resolve(cfg)["NB"] == 3
    finally:
        sys.path.remove(ex)
        sys.modules.pop("dataset_shapes", None)


def test_checkpoint_containers_round_trip_by_variable_name(tmp_path):
    """.pt / .npz / .safetensors hold the same {Keras variable name: array} mapping; HDF5-style names
    (':0' suffix, nested-model prefix) are normalised; `exclude` drops whole layers; .h5 is refused with instructions."""
    import torch
    from myolo import checkpoint as ck
    from myolo.engine import param_specs
    specs = param_specs(3, 4)
    rs = np.random.RandomState(0)
    sd = {name: torch.from_numpy(rs.standard_normal(shape).astype(np.float32)) for name, shape, _ in specs[:12]}
    for ext in (".pt", ".npz", ".safetensors"):
        path = str(tmp_path / ("w" + ext))
        ck.write_checkpoint(path, sd)
        assert os.path.exists(path) and not os.path.exists(path + ".npz")
        back = ck.read_checkpoint(path)
        assert list(back) == list(sd) or set(back) == set(sd)
        assert all(torch.equal(back[k], sd[k]) and back[k].dtype == torch.float32 for k in sd)
    keras_style = {("yolo_model/" if i % 2 else "") + k + ":0": v.double() for i, (k, v) in enumerate(sd.items())}
    path = str(tmp_path / "keras_names.npz")
    with open(path, "wb") as f:
        np.savez(f, **{k: v.numpy() for k, v in keras_style.items()})
    back = ck.read_checkpoint(path)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert ck.normalise_key("yolo_model/conv_dw_7/depthwise_kernel:0") == "conv_dw_7/depthwise_kernel"
    assert ck.normalise_key("conv1/kernel") == "conv1/kernel"
    kept = ck.select(sd, exclude=["conv1_bn"])
    assert not any(k.startswith("conv1_bn/") for k in kept) and len(kept) == len(sd) - 4 and "conv1/kernel" in kept
    # Keras HDF5 weight files go through myolo.h5lite (tests/test_h5lite.py): same variables, same values
    for name in ("weights.h5", "weights.hdf5"):
        ck.write_checkpoint(str(tmp_path / name), sd)
        back = ck.read_checkpoint(str(tmp_path / name))
        assert set(back) == set(sd) and all(torch.equal(back[k], torch.as_tensor(sd[k]).float()) for k in sd)


def test_iou_parity_helpers():
    import torch
    from tests import helpers as Hh
    a = torch.tensor([[0.0, 0.0, 2.0, 2.0], [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 3.0, 3.0]])
    b = torch.tensor([[1.0, 0.0, 3.0, 2.0], [0.0, 0.0, 0.0, 0.0], [1.0, 1.0, 3.0, 3.0]])
    assert torch.allclose(Hh.box_iou_pairs(a, b), torch.tensor([1.0 / 3.0, 1.0, 1.0], dtype=torch.float64))
    m = torch.zeros(2, 4, 4, 3)
    m[0, :2, :, 0] = 0.9                                   # 8 pixels on
    n = m.clone()
    n[0, 2, :, 0] = 0.6                                    # 12 pixels on -> IoU 8/12
    n[1, 0, 0, 2] = 0.49                                   # below the threshold: still empty
    iou = Hh.mask_iou(m, n)
    assert iou.shape == (2, 3) and abs(iou[0, 0].item() - 8.0 / 12.0) < 1e-12 and iou[1, 2].item() == 1.0 and iou[0, 1].item() == 1.0


def test_prefetcher_is_ordered_bounded_and_propagates_errors():
    import threading
    import time
    from myolo.model import Prefetcher

    class Seq(object):
        def __init__(self, fail_at=None):
            self.lock, self.inflight, self.peak, self.built, self.fail_at = threading.Lock(), 0, 0, [], fail_at

        def __getitem__(self, i):
            with self.lock:
                self.inflight += 1
                self.peak = max(self.peak, self.inflight)
                self.built.append(i)
            time.sleep(0.004 * (4 - i % 4))                 # later items of a window finish first
            with self.lock:
                self.inflight -= 1
            if i == self.fail_at:
                raise ValueError("item %d" % i)
            return i * i

    s = Seq()
    consumed = []
    for v in Prefetcher(s, range(11), depth=3, workers=2):
        consumed.append(v)
        assert len(s.built) <= len(consumed) + 3            # never more than `depth` items ahead of the consumer
    assert consumed == [i * i for i in range(11)] and s.peak <= 2
    s = Seq(fail_at=5)
    got = []
    with pytest.raises(ValueError, match="item 5"):
        for v in Prefetcher(s, range(11), depth=3, workers=2):
            got.append(v)
    assert got == [0, 1, 4, 9, 16] and max(s.built) <= 8    # raised at its position; the producer stopped
    assert list(Prefetcher(Seq(), [], 3, 2)) == [] and list(Prefetcher(Seq(), [3], 1, 1)) == [9]


def test_train_loop_on_stub_engine(tmp_path):
    """MaskYOLO.train (model.py:943-1060) with the device step stubbed out: caching, BatchGenerator construction, the
    prefetched epoch loop, validation with update=False, history and the per-epoch checkpoint."""
    import torch
    from myolo.model import MaskYOLO

    class C128(ShapesConfig):
        BATCH_SIZE = 4
        IMAGE_SHAPE = [128, 128, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = 128
        GRID_H = GRID_W = 4

    class StubEngine(object):
        B = 4
        trainable = None

        def state_dict(self):
            return {"conv1/kernel": torch.zeros(3, 3, 3, 32)}

        def set_trainable(self, pred, base=False):
            self.trainable = [n for n in ("conv1/kernel", "myolo_mask_conv1/kernel") if pred(n)]

        def reset_optimizer(self):
            self.resets = getattr(self, "resets", 0) + 1

    cfg = C128()
    tr, va = ShapesDataset(seed=1), ShapesDataset(seed=2)
    tr.load_shapes(10, 128, 128); tr.prepare()
    va.load_shapes(4, 128, 128); va.prepare()
    m = MaskYOLO.__new__(MaskYOLO)
    m.mode, m.config, m.model_dir, m.epoch, m.engine, m.learning_rate = "training", cfg, str(tmp_path), 0, StubEngine(), None
    calls = []

    def fake_step(inputs, update=True, lr=None):
        assert len(inputs) == 6 and inputs[0].shape == (4, 128, 128, 3) and inputs[0].dtype == np.float32
        assert inputs[5].shape == (4, 128, 128, cfg.MAX_GT_INSTANCES) and inputs[3].shape == (4, cfg.TRUE_BOX_BUFFER)
        calls.append(update)
        return [3.0, 1.0, 2.0]

    # the fit loop stages + enqueues step k+1 before it reads the losses of step k (MaskYOLO.fit_batches)
    order = []
    m._enqueue_step = lambda inputs, update=True, lr=None: (order.append("enqueue"), fake_step(inputs, update, lr))[1]
    m._read_step = lambda pending: (order.append("read"), pending)[1]
    np.random.seed(0)
    hist = m.train(tr, va, learning_rate=0.01, epochs=2, layers=r"(myolo_mask.*)", verbose=0)
    # 10 images / batch 4 -> 3 batches (the last one refilled from the preceding images), 4 val images -> 1 batch
    assert calls == [True, True, True, False] * 2
    assert hist == {"loss": [3.0, 3.0], "yolo_sum_loss": [1.0, 1.0], "myolo_mask_loss": [2.0, 2.0], "val_loss": [3.0, 3.0]}
    assert m.engine.trainable == ["myolo_mask_conv1/kernel"] and m.learning_rate == 0.01 and m.epoch == 2
    assert m.engine.resets == 1                                   # compile() starts a fresh Adam (model.py:1071-1075)
    assert order[:7] == ["enqueue", "enqueue", "read", "enqueue", "read", "read", "enqueue"]     # 3 train batches, then validation
    saved = [f for f in os.listdir(str(tmp_path)) if f.startswith("saved_model_") and f.endswith(".pt")]
    assert len(saved) == 1 and "conv1/kernel" in torch.load(os.path.join(str(tmp_path), saved[0]))


def test_padded_flat_layout_turns_a_3x3_same_conv_into_nine_shifted_gemms():
    """DESIGN section 3, on the CPU: in the padded-flat layout a 3x3 SAME convolution is sum_t A[row + shift_t] @ W_t with
    shift(dy,dx) = (dy-1)(W+1) + (dx-1), tiles cannot bleed into each other, and the transposed (negated) shifts give the
    data gradient; guard rows and pad rows stay zero."""
    import torch
    import torch.nn.functional as F
    from myolo import pf
    n, H, W, C, Co = 3, 5, 4, 6, 7
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, H, W, C, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, C, Co, generator=g, dtype=torch.float64)          # HWIO
    A = pf.PF(n, H, W, C, device="cpu", dtype=torch.float64).load_dense(x)
    assert A.M == n * (H + 1) * (W + 1) and A.rows.shape == (A.M, C)
    assert torch.equal(A.dense(), x)
    full = A.storage.view(-1, C)
    gr = pf.guard_rows(W)
    assert gr >= W + 2 and not full[:gr].any() and not full[gr + A.M:].any()     # zero guards around the matrix
    grid = A.rows.view(n, H + 1, W + 1, C)
    assert not grid[:, 0].any() and not grid[:, :, 0].any()                      # one zero row-block / one zero pixel per line
    shifts = pf.conv3x3_shifts(W)
    assert shifts == [(dy - 1) * (W + 1) + (dx - 1) for dy in range(3) for dx in range(3)] and max(map(abs, shifts)) == W + 2

    def tap_gemm(storage_rows, M, weights, sh):
        out = torch.zeros(M, weights.shape[-1], dtype=torch.float64)
        for t, s in enumerate(sh):
            out += storage_rows[gr + s: gr + s + M] @ weights[t]
        return out

    y = tap_gemm(full, A.M, w.reshape(9, C, Co), shifts).view(n, H + 1, W + 1, Co)[:, 1:, 1:]
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-12)
    # data gradient: dy in PF layout (pads zero), taps with negated shifts and transposed weights
    dy = torch.randn(n, H, W, Co, generator=g, dtype=torch.float64)
    D = pf.PF(n, H, W, Co, device="cpu", dtype=torch.float64).load_dense(dy)
    dx = tap_gemm(D.storage.view(-1, Co), D.M, w.reshape(9, C, Co).transpose(1, 2), pf.conv3x3_shifts(W, negate=True))
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1).backward(dy)
    assert torch.allclose(dx.view(n, H + 1, W + 1, C)[:, 1:, 1:], xr.grad, atol=1e-12)
    v = A.view()
    assert (v.n, v.h, v.w, v.c, v.sh, v.sn) == (n, H, W, C, (W + 1) * C, (H + 1) * (W + 1) * C)
    assert v.p == A.rows.data_ptr() + 8 * ((W + 1) + 1) * C                     # first valid pixel, float64 here


def test_detect_for_one_is_detect_on_a_one_element_list():
    """the example scripts call model.detect_for_one([image], verbose=1); the reference defines no such method."""
    from myolo.model import MaskYOLO
    m = MaskYOLO.__new__(MaskYOLO)              # no engine: only the delegation is checked here
    seen = []
    m.detect = lambda image: (seen.append(image), ["result"])[1]
    img = np.zeros((4, 4, 3), np.uint8)
    assert m.detect_for_one([img], verbose=1) == ["result"] and seen[0] is img
    with pytest.raises(AssertionError):
        m.detect_for_one([img, img])
