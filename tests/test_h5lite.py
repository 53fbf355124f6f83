"""myolo.h5lite: the HDF5 subset Keras 2.x weight files use, restated in pure Python, and
MaskYOLO.load_weights(filepath, by_name, exclude) semantics over such files (myolo/model.py:1157-1196) -- no h5py."""
import struct

import numpy as np
import pytest
import torch

from myolo import checkpoint, h5lite
from myolo.engine import init_params, param_specs


def _state(nb=3, nc=4, seed=3):
    return init_params(nb, nc, seed, "trained_like")


def test_keras_weight_file_roundtrip_with_nested_yolo_model(tmp_path):
    sd = _state()
    path = str(tmp_path / "saved_model.h5")
    checkpoint.write_checkpoint(path, sd)
    f = h5lite.File(path)
    layer_names = [n.decode() for n in f.attrs["layer_names"]]
    assert "yolo_model" in layer_names and "conv1" in layer_names and "conv_dw_7" not in layer_names
    assert len(f.keys()) == len(layer_names) > 16                      # several SNODs below the root B-tree (8 links each)
    assert f.attrs["backend"] == b"tensorflow"
    # Keras' layout: /conv1/conv1/kernel:0 and /yolo_model/conv_dw_7/depthwise_kernel:0
    assert np.array_equal(f["conv1/conv1/kernel:0"].read(), sd["conv1/kernel"].numpy())
    assert np.array_equal(f["yolo_model"]["conv_dw_7/depthwise_kernel:0"].read(), sd["conv_dw_7/depthwise_kernel"].numpy())
    wn = [w.decode() for w in f["yolo_model"].attrs["weight_names"]]
    assert "conv_23/bias:0" in wn and "conv_pw_14_bn/moving_variance:0" in wn and len(wn) == 8 * 10 + 2
    with pytest.raises(KeyError):
        f["conv1/nothing"]
    back = checkpoint.read_checkpoint(path)
    assert list(back) == [n for n, _, _ in param_specs(3, 4)] or set(back) == set(sd)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_load_weights_semantics_by_name_and_exclude(tmp_path):
    sd = _state()
    path = str(tmp_path / "w.h5")
    checkpoint.write_checkpoint(path, sd)
    # exclude a top-level layer, and the nested model as a whole (model.py:1170-1180 filters layers by name)
    part = checkpoint.read_checkpoint(path, exclude=["myolo_mask_conv1", "yolo_model"])
    assert "myolo_mask_conv1/kernel" not in part and "myolo_mask_conv2/kernel" in part
    assert not any(k.startswith(("conv_dw_7", "conv_pw_14", "conv_23")) for k in part) and "conv_pw_6/kernel" in part
    assert "conv_dw_9_bn/gamma" not in checkpoint.read_checkpoint(path, exclude=["conv_dw_9_bn"])

    class Eng:                                       # load_params' contract, without a GPU
        def __init__(self):
            self.specs = param_specs(3, 4)
            self.got = {}

        def load_params(self, P, strict=True):
            for name, shape, _ in self.specs:
                if name not in P:
                    if strict:
                        raise KeyError(name)
                    continue
                self.got[name] = torch.as_tensor(P[name]).reshape(shape)

    from myolo.model import MaskYOLO
    m = MaskYOLO.__new__(MaskYOLO)
    m.engine = Eng()
    m.load_weights(path)                             # strict: every variable present
    assert len(m.engine.got) == len(sd)
    m.engine = Eng()
    with pytest.raises(KeyError):
        sub = {k: v for k, v in sd.items() if not k.startswith("myolo_mask")}
        checkpoint.write_checkpoint(str(tmp_path / "yolo_only.h5"), sub)
        m.load_weights(str(tmp_path / "yolo_only.h5"))
    m.engine = Eng()
    m.load_weights(str(tmp_path / "yolo_only.h5"), by_name=True)        # by_name tolerates the missing mask head
    assert "conv_23/kernel" in m.engine.got and "myolo_mask/kernel" not in m.engine.got
    m.engine = Eng()
    m.load_weights(path, by_name=True, exclude=["yolo_model"])
    assert "conv_pw_6/kernel" in m.engine.got and "conv_23/kernel" not in m.engine.got


def test_writer_emits_the_structures_of_the_format_specification(tmp_path):
    path = str(tmp_path / "t.h5")
    h5lite.write_tree(path, {"g": ({"d": np.arange(6, dtype=np.float32).reshape(2, 3)}, {"a": np.array([b"xy", b"z"])}),
                             "s": np.float64(2.5)}, {"n": np.int32(7)})
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                      # superblock version 0
    assert raw[13] == 8 and raw[14] == 8 and struct.unpack_from("<HH", raw, 16) == (4, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert (base, free, drv, eof) == (0, h5lite.UNDEF, h5lite.UNDEF, len(raw))
    root_hdr, cache_type = struct.unpack_from("<Q", raw, 64)[0], struct.unpack_from("<I", raw, 72)[0]
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert cache_type == 1 and raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    assert raw[root_hdr] == 1 and struct.unpack_from("<H", raw, root_hdr + 16)[0] == 0x0011   # v1 header, symbol table first
    f = h5lite.File(path)
    assert sorted(f.keys()) == ["g", "s"] and f.attrs["n"] == 7
    assert f["s"].read() == 2.5 and f["s"].shape == ()
    assert list(f["g"].attrs["a"]) == [b"xy", b"z"]
    d = f["g/d"]
    assert d.shape == (2, 3) and np.array_equal(np.asarray(d), np.arange(6, dtype=np.float32).reshape(2, 3))


def test_reader_on_a_hand_assembled_file_with_other_legal_encodings(tmp_path):
    """A file put together byte by byte from the format specification, using encodings the writer above never produces:
    superblock version 1, a version-2 dataspace, a compact dataset, big-endian data, an attribute in a header
    continuation block."""
    buf = bytearray(2048)

    def msg(typ, data):
        data = bytes(data) + b"\0" * ((-len(data)) % 8)
        return struct.pack("<HHB3x", typ, len(data), 0) + data

    # dataset object header at 512: dataspace v2 (rank 1, dim 3), big-endian float64, compact layout, continuation -> 1024
    values = np.array([1.5, -2.0, 3.25], dtype=">f8")
    dspace = struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 3)
    dtype = struct.pack("<BBBBI", 0x11, 0x21, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    layout = struct.pack("<BBH", 3, 0, values.nbytes) + values.tobytes()
    first = msg(0x0001, dspace) + msg(0x0003, dtype) + msg(0x0008, layout) + msg(0x0010, struct.pack("<QQ", 1024, 64))
    buf[512:512 + 16] = struct.pack("<BxHII4x", 1, 5, 1, len(first))
    buf[528:528 + len(first)] = first
    name = b"unit\0"
    adt = struct.pack("<BBBBI", 0x13, 0x01, 0, 0, 2)
    ads = struct.pack("<BBB5x", 1, 0, 0)
    att = struct.pack("<BxHHH", 1, len(name), len(adt), len(ads)) + name + b"\0" * 3 + adt + ads + b"mm"
    cont = msg(0x000C, att)
    assert len(cont) <= 64
    buf[1024:1024 + len(cont)] = cont
    # root group at 256: symbol table -> B-tree 320 (one SNOD at 400), local heap 384 with data segment at 1200
    heap_data = b"\0" * 8 + b"temps\0\0\0"
    buf[1200:1200 + len(heap_data)] = heap_data
    buf[384:416] = struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap_data), h5lite.UNDEF, 1200)
    buf[400 + 32:400 + 32 + 8 + 40] = struct.pack("<4sBxH", b"SNOD", 1, 1) + struct.pack("<QQII16x", 8, 512, 0, 0)
    buf[320:320 + 24 + 24] = struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1, h5lite.UNDEF, h5lite.UNDEF) + struct.pack("<QQQ", 0, 432, 8)
    root = msg(0x0011, struct.pack("<QQ", 320, 384))
    buf[256:272] = struct.pack("<BxHII4x", 1, 1, 1, len(root))
    buf[272:272 + len(root)] = root
    # superblock version 1: two extra fields (indexed storage K, reserved) before the addresses
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 1, 0, 0, 0, 8, 8, 4, 16, 0) + struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", 0, h5lite.UNDEF, len(buf), h5lite.UNDEF) + struct.pack("<QQII", 0, 256, 1, 0) + struct.pack("<QQ", 320, 384)
    buf[:len(sb)] = sb
    path = str(tmp_path / "hand.h5")
    open(path, "wb").write(bytes(buf))
    f = h5lite.File(path)
    assert f.keys() == ["temps"]
    d = f["temps"]
    assert d.shape == (3,) and d.read().tolist() == [1.5, -2.0, 3.25] and d.attrs["unit"] == b"mm"


def test_unsupported_features_are_named(tmp_path):
    path = str(tmp_path / "t.h5")
    h5lite.write_tree(path, {"d": np.zeros(3, np.float32)})
    raw = bytearray(open(path, "rb").read())
    raw[8] = 2
    open(path, "wb").write(bytes(raw))
    with pytest.raises(h5lite.H5Unsupported, match="superblock version 2"):
        h5lite.File(path)
    open(path, "wb").write(b"not hdf5 at all")
    with pytest.raises(ValueError):
        h5lite.File(path)
