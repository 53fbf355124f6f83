"""ctypes binding of libmyolo_sm100.so (the C ABI declared in include/myolo_b200.h).

The prototypes are parsed from the header itself, so the header is the single source of truth for
the boundary.  There is no CPU or library fallback: if the shared library is missing, or the
device is not an sm_100-class GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)                      # mask-yolo_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
HEADER = os.path.join(REPO_ROOT, "include", "myolo_b200.h")
LIB_PATH = os.environ.get("MYOLO_LIB") or os.path.join(PKG_ROOT, "lib", "libmyolo_sm100.so")   # MYOLO_LIB: A/B runs of two builds

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
ROUND_TF32 = 0x100
PREC_FP32, PREC_TF32 = 0, 1


class View(ctypes.Structure):
    """myolo_view: strided NHWC view (innermost C contiguous, pixel stride = C)."""
    _fields_ = [("p", ctypes.c_void_p), ("sn", ctypes.c_longlong), ("sh", ctypes.c_longlong),
                ("n", ctypes.c_int), ("h", ctypes.c_int), ("w", ctypes.c_int), ("c", ctypes.c_int)]


_CTYPE = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
    "myolo_stream": ctypes.c_void_p,
}


def parse_header(path: str = HEADER):
    """Returns {name: (restype, [(ctype, argname), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(myolo_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    typ, an = a.rsplit(" ", 1)
                    argl.append((_CTYPE[typ.replace("const ", "").strip()], an))
        protos[name] = (restype, argl)
    return protos


class MyoloError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None
_protos = None


def lib():
    """Loads the library once.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MyoloError(f"{LIB_PATH} not found: build it with `make -C {PKG_ROOT}/csrc` "
                             "(there is no CPU fallback for the Mask-YOLO hot path)")
        l = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (restype, argl) in _protos.items():
            fn = getattr(l, name)          # AttributeError if the header declares a missing symbol
            fn.restype = restype
            fn.argtypes = [t for t, _ in argl]
        _lib = l
    return _lib


def _conv(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):             # torch tensor
        if not x.is_cuda:
            raise MyoloError("device pointer argument is a CPU tensor (no CPU path exists)")
        if not x.is_contiguous():
            raise MyoloError("tensor arguments must be contiguous")
        return x.data_ptr()
    if isinstance(x, View):
        return ctypes.addressof(x)
    if isinstance(x, ctypes.Array):
        return ctypes.addressof(x)
    return x


# device kernels launched per entry point (anything not listed launches exactly one); bench.py reports
# the sum over the timed region as `gpu_launches`
KERNELS_PER_CALL = {"myolo_bn_stats": 1, "myolo_bn_bwd": 2, "myolo_bn_bwd_h": 2, "myolo_bn_bwd_hh": 2, "myolo_grad_scale": 2,
                    "myolo_gemm_taps_bnbwd": 2, "myolo_gemm_taps_bnbwd_h": 2, "myolo_gemm_taps_h_stats": 2, "myolo_gemm_taps_hh_stats": 2, "myolo_bn_bwd_batch_fix_hh": 2, "myolo_gemm_segs_win": 2, "myolo_gemm_segs_win_supported": 0, "myolo_gemm_taps_h_supported": 0,
                    "myolo_gemm_taps_wgrad_h_supported": 0, "myolo_colsum": 1, "myolo_detect_mask_targets": 2,
                    "myolo_mask_loss": 3, "myolo_yolo_loss": 3, "myolo_shapes_raster": 2, "myolo_polygon_masks": 2, "myolo_version": 0, "myolo_last_error": 0,
                    "myolo_device_check": 0, "myolo_set_precision": 0, "myolo_get_precision": 0, "myolo_set_wgrad_sms": 0,
                    "myolo_gemm_taps_tc_supported": 0, "myolo_gemm_taps_wgrad_tc_supported": 0}
launch_count = 0


# ---- launch recording (Engine.train_step): the sequence of C-ABI calls of a training step and their converted
# arguments is the same every step, so the engine records it once and replays the raw ctypes calls afterwards
# (the Python around ~290 launches cost 15 ms per step, against 19 ms of device time).
_rec = None


def start_recording():
    global _rec
    _rec = []


def stop_recording():
    global _rec
    r, _rec = _rec, None
    return r


class no_record(object):
    """Suspend the launch recording: for a host action that is itself recorded with record_py and issues C-ABI calls whose
    arguments change from step to step (it runs again at every replay)."""

    def __enter__(self):
        global _rec
        self.saved, _rec = _rec, None

    def __exit__(self, *a):
        global _rec
        _rec = self.saved
        return False


def record_py(fn):
    """Run a host-side action (event record / wait, torch fill, hook) and, while recording, note it in sequence."""
    fn()
    if _rec is not None:
        _rec.append([None, fn, None, 0, "py"])


# profiling only (scripts/whatif_critical_path.sh): entry points named in MYOLO_WHATIF_SKIP are NOT launched, so that the
# step time without them shows what each costs on the critical path.  The step's results are garbage; bench.py marks
# such a line invalid.
WHATIF_SKIP = frozenset(x for x in os.environ.get("MYOLO_WHATIF_SKIP", "").split(",") if x)


def call(name: str, *args):
    """Invoke a C-ABI entry point; tensors -> device pointers; nonzero status -> MyoloError."""
    global launch_count
    if WHATIF_SKIP and name in WHATIF_SKIP:
        return 0
    l = lib()
    fn = getattr(l, name)
    n = KERNELS_PER_CALL.get(name, 1)
    launch_count += n
    conv = [_conv(a) for a in args]         # `args` keeps Views / arrays alive across the call
    if _rec is not None:
        _rec.append([fn, conv, args, n, name])
    rc = fn(*conv)
    if rc != 0:
        raise MyoloError(f"{name} failed ({rc}): {l.myolo_last_error().decode()}")
    return rc


def replay(entries):
    """Re-issue a recorded sequence: raw ctypes calls with the stored arguments, host actions in between."""
    global launch_count
    for e in entries:
        fn = e[0]
        if fn is None:
            e[1]()
        else:
            rc = fn(*e[1])
            if rc != 0:
                raise MyoloError(f"{e[4]} failed ({rc}): {lib().myolo_last_error().decode()}")
            launch_count += e[3]


def view(t, n, h, w, c, sn=None, sh=None, offset=0):
    """View over tensor `t` starting `offset` elements in; dense unless strides are given."""
    sh = w * c if sh is None else sh
    sn = h * sh if sn is None else sn
    return View(t.data_ptr() + 4 * offset, sn, sh, n, h, w, c)


def device_check(dev: int = 0):
    call("myolo_device_check", dev)


def set_precision(mode: int):
    call("myolo_set_precision", mode)


def get_precision() -> int:
    return lib().myolo_get_precision()


def int_array(vals):
    return (ctypes.c_int * len(vals))(*vals)


def float_array(vals):
    return (ctypes.c_float * len(vals))(*vals)
