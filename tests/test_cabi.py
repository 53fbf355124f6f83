"""The C-ABI boundary without a GPU: the in-tree shared library loads and exports every symbol
include/myolo_b200.h declares; argument validation and the no-CPU-fallback rule hold."""
import ctypes
import os

import pytest
import torch


def test_library_exports_every_declared_symbol():
    from myolo import _cabi
    protos = _cabi.parse_header()
    assert len(protos) >= 40
    for must in ("myolo_dwconv3x3_fwd", "myolo_pwconv_fwd", "myolo_conv3x3_wgrad", "myolo_roialign_bwd", "myolo_gemm_taps_tc",
                 "myolo_detect_mask_targets", "myolo_yolo_loss", "myolo_mask_out_bwd", "myolo_adam_step", "myolo_bn_apply_split"):
        assert must in protos
    lib = _cabi.lib()                      # binds restype/argtypes for each prototype: AttributeError if one is missing
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.myolo_version() >= 100
    assert lib.myolo_last_error() is not None


def test_precision_switch_and_argument_errors_need_no_device():
    from myolo import _cabi as C
    C.set_precision(C.PREC_TF32)
    assert C.get_precision() == C.PREC_TF32
    C.set_precision(C.PREC_FP32)
    with pytest.raises(C.MyoloError, match="argument check failed"):
        C.call("myolo_set_precision", 7)
    assert C.lib().myolo_gemm_taps_tc_supported(256, 256, 1000, 256, 256, 9, 0) == 1
    assert C.lib().myolo_gemm_taps_tc_supported(256, 45, 1000, 45, 256, 1, 0) == 0      # N=45 -> CUDA-core kernel


def test_no_cpu_fallback():
    from myolo import _cabi as C
    with pytest.raises(C.MyoloError, match="CPU tensor"):
        C.call("myolo_conv1_fwd", torch.zeros(4), torch.zeros(4), torch.zeros(4), 1, 4, 32, None)
    if not torch.cuda.is_available():
        from myolo.engine import Engine
        from tests import helpers as Hh
        with pytest.raises(C.MyoloError, match="no CPU fallback"):
            Engine(Hh.engine_cfg(S=64), 1)


def test_product_code_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mask-yolo_b200")
    import re
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|.*importlib.*oracle|.*oracle[/.]myolo_oracle)", re.M)
    n = 0
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                n += 1
                assert not pat.search(open(os.path.join(dp, f), errors="ignore").read()), os.path.join(dp, f)
    assert n >= 8


def test_allreduce_entry_points_bind_nccl_at_run_time():
    """myolo_allreduce_* resolve NCCL with dlsym from the library PyTorch ships; ncclGetUniqueId needs no GPU, so the
    binding itself is checked here (communicator creation and the all-reduce are -m gpu)."""
    import ctypes
    from myolo import _cabi as C
    from myolo import ddp
    path = ddp.load_nccl_global()
    assert "nccl" in path
    a, b = ctypes.create_string_buffer(128), ctypes.create_string_buffer(128)
    assert C.call("myolo_allreduce_unique_id", a) == 0 and C.call("myolo_allreduce_unique_id", b) == 0
    assert any(a.raw) and a.raw != b.raw                    # 128 opaque bytes, different per call
    with pytest.raises(C.MyoloError):
        C.call("myolo_allreduce_run", None, None, 0, 0)     # argument checks come before any NCCL call
    with pytest.raises(C.MyoloError):
        C.call("myolo_allreduce_init", a, 3, 2, ctypes.addressof(ctypes.c_void_p()))
