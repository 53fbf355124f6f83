"""The reference-facing Python API on the GPU: MaskYOLO.train / keras_model.train_on_batch / predict /
detect / load_weights / set_trainable, driven the way example/shapes drives the reference."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg():
    from myolo.shapes import ShapesConfig

    class C128(ShapesConfig):
        BATCH_SIZE = 4
        IMAGE_SHAPE = [128, 128, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = 128
        GRID_H = GRID_W = 4
        ANCHORS = [0.6, 0.6, 1.2, 1.3, 2.0, 2.1]
    return C128()


def test_train_loop_checkpoint_and_reload(tmp_path):
    from myolo.model import MaskYOLO
    from myolo.shapes import ShapesDataset
    cfg = _cfg()
    tr, va = ShapesDataset(seed=1), ShapesDataset(seed=2)
    tr.load_shapes(12, 128, 128); tr.prepare()
    va.load_shapes(4, 128, 128); va.prepare()
    np.random.seed(0)                          # BatchGenerator shuffles with the global numpy RNG
    model = MaskYOLO(mode="training", config=cfg, model_dir=str(tmp_path))
    assert model.keras_model.metrics_names == ["loss", "yolo_sum_loss", "myolo_mask_loss"]
    assert "Total params" in model.keras_model.summary()
    hist = model.train(tr, va, learning_rate=cfg.LEARNING_RATE, epochs=12, layers="all", verbose=0)
    assert len(hist["loss"]) == 12 and all(np.isfinite(hist["loss"])) and np.isfinite(hist["val_loss"][-1])
    assert min(hist["loss"][-3:]) < 0.5 * hist["loss"][0], hist["loss"]     # 36 Adam steps: the loss comes down
    ckpts = glob.glob(os.path.join(str(tmp_path), "saved_model_*.pt"))
    assert len(ckpts) == 1
    sd = model.engine.state_dict()
    other = MaskYOLO(mode="inference", config=cfg, seed=99)
    other.load_weights(ckpts[0])
    sd2 = other.engine.state_dict()
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    # inference API on one uint8 image
    img = tr.load_image(0)
    res = other.detect(img, cs_threshold=0.0)
    n = res["rois"].shape[0]
    assert res["masks"].shape == (128, 128, n) and res["class_ids"].shape == (n,) and res["scores"].shape == (n,)
    outs = other.keras_model.predict([np.repeat((img / 255.).astype(np.float32)[None], cfg.BATCH_SIZE, 0), None])
    assert outs[0].shape == (4, 4, 4, 3, 9) and outs[1].shape == (4, 48, 6) and outs[2].shape == (4, 48, 28, 28, 4)
    with pytest.raises(AssertionError):
        other.detect(img.astype(np.float32))


def test_set_trainable_freezes_layers_and_bad_image_size_raises():
    from myolo.model import MaskYOLO
    from myolo.shapes import make_batches
    cfg = _cfg()
    model = MaskYOLO(mode="training", config=cfg)
    batch = make_batches(cfg, 1, seed=3)[0]
    before = model.engine.state_dict()
    model.set_trainable(r"(conv_(pw|dw)_1[0-4].*)|(conv_23)")
    vals = model.keras_model.train_on_batch(batch)
    assert len(vals) == 3 and all(np.isfinite(vals)) and abs(vals[0] - (vals[1] + vals[2])) < 1e-4
    after = model.engine.state_dict()
    assert torch.equal(before["conv_pw_3/kernel"], after["conv_pw_3/kernel"])           # frozen
    assert torch.equal(before["conv1/kernel"], after["conv1/kernel"])
    assert torch.equal(before["myolo_mask_conv2/kernel"], after["myolo_mask_conv2/kernel"])
    assert not torch.equal(before["conv_23/kernel"], after["conv_23/kernel"])           # trainable
    assert not torch.equal(before["conv_pw_12/kernel"], after["conv_pw_12/kernel"])
    assert not torch.equal(before["conv_pw_3_bn/moving_mean"], after["conv_pw_3_bn/moving_mean"])   # BN still in training phase
    v2 = model.keras_model.test_on_batch(batch)
    assert all(np.isfinite(v2))

    class Bad(type(cfg)):
        IMAGE_SHAPE = [100, 100, 3]
    with pytest.raises(Exception, match="dividable by 32"):
        MaskYOLO(mode="training", config=Bad())
