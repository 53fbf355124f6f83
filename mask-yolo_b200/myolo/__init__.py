"""B200-native Mask-YOLO hot path behind the reference's Python API (myolo.config / myolo.model /
myolo.myolo_utils).  The arithmetic lives in libmyolo_sm100.so (hand-written sm_100a CUDA, C ABI in
include/myolo_b200.h); this package is the host side."""
