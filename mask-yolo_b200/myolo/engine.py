"""Executor of the Mask-YOLO hot path on one B200: forward, backward, Adam and the Keras BN
moving-average update, entirely through the C ABI of libmyolo_sm100.so (include/myolo_b200.h).

This replaces what the reference delegates to one `tf.Session.run(train_op)` per batch
(myolo/model.py:1047-1059 -> the graph wired in MaskYOLO.build, 787-941).  PyTorch is used for
device memory and streams only; no torch operator computes anything on the hot path and there is
no CPU fallback (the C-ABI loader raises when the library or an sm_100 device is missing).

Data layout in HBM: activations fp32 NHWC; every tensor that feeds a 3x3
tensor-core convolution lives in the padded-flat (PF) layout of myolo/pf.py; parameters, gradients
and the two Adam moments are four flat fp32 buffers with identical offsets (one fused Adam launch,
one NCCL all-reduce), keyed by the Keras variable names of the reference graph.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi as C
from .pf import PF, conv3x3_shifts

BN_EPS = 1e-3        # Keras BatchNormalization default (graph fixture: epsilon 0.001)
BN_MOMENTUM = 0.99

# (block id, Cin, Cout, stride): myolo/model.py:68-77 (backbone) and 256-268 (yolo branch)
BACKBONE_BLOCKS = [(1, 32, 64, 1), (2, 64, 64, 2), (3, 64, 128, 1), (4, 128, 256, 2), (5, 256, 256, 1), (6, 256, 512, 1)]
YOLO_BLOCKS = [(7, 512, 512, 2), (8, 512, 512, 1), (9, 512, 512, 1), (10, 512, 512, 1), (11, 512, 512, 1),
               (12, 512, 512, 1), (13, 512, 1024, 2), (14, 1024, 1024, 1)]
MASK_C = 256          # TOP_FEATURE_MAP_DEPTH (config.py:91) == mask head width (model.py:688-711)


def param_specs(nb: int, nc: int):
    """[(keras variable name, shape, trainable)] in flat-buffer order: backbone, yolo branch, then
    feature_map + mask head LAST (their gradients are complete first in the backward pass, so the
    tail of the flat gradient buffer can be all-reduced while the backbone is still running)."""
    sp = [("conv1/kernel", (3, 3, 3, 32), True)]

    def bn(name, c):
        return [(f"{name}/gamma", (c,), True), (f"{name}/beta", (c,), True),
                (f"{name}/moving_mean", (c,), False), (f"{name}/moving_variance", (c,), False)]

    sp += bn("conv1_bn", 32)
    for k, ci, co, _ in BACKBONE_BLOCKS + YOLO_BLOCKS:
        sp.append((f"conv_dw_{k}/depthwise_kernel", (3, 3, ci, 1), True))
        sp += bn(f"conv_dw_{k}_bn", ci)
        sp.append((f"conv_pw_{k}/kernel", (1, 1, ci, co), True))
        sp += bn(f"conv_pw_{k}_bn", co)
    sp += [("conv_23/kernel", (1, 1, 1024, nb * (5 + nc)), True), ("conv_23/bias", (nb * (5 + nc),), True)]
    sp += [("feature_map/kernel", (3, 3, 512, MASK_C), True), ("feature_map/bias", (MASK_C,), True)]
    for i in (1, 2, 3, 4):
        sp += [(f"myolo_mask_conv{i}/kernel", (3, 3, MASK_C, MASK_C), True), (f"myolo_mask_conv{i}/bias", (MASK_C,), True)]
        sp += bn(f"myolo_mask_bn{i}", MASK_C)
    sp += [("myolo_mask_deconv/kernel", (2, 2, MASK_C, MASK_C), True), ("myolo_mask_deconv/bias", (MASK_C,), True)]
    sp += [("myolo_mask/kernel", (1, 1, MASK_C, nc), True), ("myolo_mask/bias", (nc,), True)]
    return sp


def init_params(nb: int, nc: int, seed: int = 0, kind: str = "keras") -> Dict[str, torch.Tensor]:
    """CPU fp32 parameter dict keyed by Keras names.  kind='keras': the Keras defaults the reference
    relies on (glorot_uniform kernels, zero biases, BN gamma 1 / beta 0 / mean 0 / var 1).
    kind='trained_like': same kernels with perturbed BN statistics and affine terms so that moving
    statistics, biases and the inference-mode mask BNs are numerically exercised."""
    g = torch.Generator().manual_seed(seed)
    P = OrderedDict()
    for name, shape, _ in param_specs(nb, nc):
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("kernel", "depthwise_kernel"):
            rf = shape[0] * shape[1]
            fan_in, fan_out = rf * shape[2], rf * shape[3]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            P[name] = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif leaf in ("gamma", "moving_variance"):
            P[name] = torch.ones(shape)
        else:
            P[name] = torch.zeros(shape)
        if kind == "trained_like":
            if leaf == "gamma":
                P[name] = 0.8 + 0.4 * torch.rand(shape, generator=g)
            elif leaf == "beta":
                P[name] = 0.2 * torch.randn(shape, generator=g)
            elif leaf == "moving_mean":
                P[name] = 0.1 * torch.randn(shape, generator=g)
            elif leaf == "moving_variance":
                P[name] = 0.6 + 0.8 * torch.rand(shape, generator=g)
            elif leaf == "bias":
                P[name] = 0.05 * torch.randn(shape, generator=g)
    return P


class _BN:
    """One BatchNormalization layer: views into the flat buffers + batch/moving statistics."""
    __slots__ = ("name", "c", "gamma", "beta", "dgamma", "dbeta", "mmean", "mvar", "bmean", "bvar", "mean", "var", "step")


class Engine:
    """Owns every device buffer of one replica and runs the step.  `cfg` is a plain dict (built by
    myolo.model from a Config instance, SURVEY Q1 rules applied): S, G, NB, NC, TB, MAXGT, R,
    ANCHORS, POOL, MASK_SHAPE, OBJECT/NO_OBJECT/COORD/CLASS scales, CLASS_WEIGHTS, WARM_UP_BATCHES,
    LOSS_WEIGHTS."""

    def __init__(self, cfg: dict, batch: int, mode: str = "training", precision: str = "tf32", device: int = 0,
                 params: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0, sparse_backward: Optional[bool] = None):
        assert mode in ("training", "inference", "yolo")
        if not torch.cuda.is_available():
            raise C.MyoloError("the Mask-YOLO hot path needs an sm_100 (B200) GPU; there is no CPU fallback")
        C.device_check(device)
        self.cfg, self.B, self.mode = cfg, batch, mode
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.set_precision(precision)
        S, G = cfg["S"], cfg["G"]
        if S % 32 != 0 or G != S // 32:
            raise Exception("Image size must be dividable by 32 to adapt with YOLO framework. "
                            "For example, use 224, 256, 288, 320, 356, ... etc. ")   # model.py:791-794
        self.NB, self.NC, self.TB, self.R = cfg["NB"], cfg["NC"], cfg["TB"], cfg["G"] * cfg["G"] * cfg["NB"]
        self.with_mask = mode != "yolo"
        self._frozen = False
        self._frozen_names = set()
        self._moving_key = None
        self._neg_shifts = C.int_array(conv3x3_shifts(cfg["POOL"], negate=True))
        self._shift_cache = {}
        self.inputs_ready = None       # optional CUDA event: inputs[1:] of forward_training are complete
        self.kernel_events = None      # bench.py: list collecting (start, end) CUDA events of the dominant kernel
        self._phases_on = os.environ.get("MYOLO_PHASES", "0") != "0"
        self._phase_evs, self._phase_order = {}, []
        # backbone backward: the two filter-gradient kernels of a block (pointwise wgrad, depthwise bwd_filter) run on a
        # side stream next to the data-gradient chain (they are 20-70 us kernels that fill a fraction of the SMs)
        self._side = torch.cuda.Stream(device=self.dev) if os.environ.get("MYOLO_BWD_STREAMS", "1") != "0" else None
        # third stream: the YOLO branch's backward next to the mask head's (Engine.backward)
        self._ystream = torch.cuda.Stream(device=self.dev) if os.environ.get("MYOLO_Y_OVERLAP", "1") != "0" else None
        # ... started already behind the yolo loss, i.e. next to the mask head's forward as well (A/B switch)
        self._y_early = os.environ.get("MYOLO_Y_EARLY", "1") != "0"
        self._yside = torch.cuda.Stream(device=self.dev) if (self._ystream is not None and os.environ.get("MYOLO_Y_SIDE", "1") != "0") else None
        # fourth stream: the mask head's filter gradients, off the data-gradient chain (Engine._backward_mask_h16)
        self._wstream = torch.cuda.Stream(device=self.dev) if (os.environ.get("MYOLO_W_OVERLAP", "1") != "0" and precision == "h16") else None
        self._w_used = False
        # ... only the last MYOLO_W_DEFER of the five (issue order deconv, conv4, conv3, conv2, conv1), sized for MYOLO_W_SMS SMs
        self._w_defer = max(0, min(5, int(os.environ.get("MYOLO_W_DEFER", "3"))))
        self._w_sms = max(32, min(148, int(os.environ.get("MYOLO_W_SMS", "110"))))
        # exact sparse backward of the mask head (h16 mode; OFF by default, never the headline number): above
        # myolo_mask_bn1 only the rois with a target class carry gradient, see _sparse_mask_middle
        if sparse_backward is None:
            sparse_backward = os.environ.get("MYOLO_SPARSE_BWD", "0") != "0"
        self.sparse_backward = bool(sparse_backward) and precision == "h16" and mode == "training"
        self.sparse_stats = {"steps": 0, "sparse": 0, "dense_fallback": 0, "no_positives": 0, "rois": 0}
        self._evs = {}
        self._replay_on = os.environ.get("MYOLO_REPLAY", "1") != "0"
        # BatchNormalization fusions of the backbone (bit mask; A/B switch): 1 = batch statistics of a depthwise output in
        # the depthwise kernel's epilogue, 2 = BN + ReLU6 after conv1 / a pointwise conv applied by the next depthwise
        # kernel (forward and filter gradient) while it loads, 4 = statistics of a pointwise output in the GEMM epilogue,
        # 8 = statistics of myolo_mask_conv1's output (myolo_mask_bn1) in the conv kernel's epilogue (h16 mode),
        # 16 = myolo_mask_bn1 on a HALF pre-BN tensor (statistics of the half-rounded values, taken in the epilogue) and its
        # backward's reduction pass in the epilogue of myolo_mask_conv2's data-gradient GEMM (h16 mode; implies 8)
        self._fuse_bn = int(os.environ.get("MYOLO_FUSE_BN", "29"))
        self._bn1_half = bool(self._fuse_bn & 16) and precision == "h16"
        # h16: d(x0), the data gradient of myolo_mask_conv1, stays a loss-scaled half tensor like every other gradient of the
        # mask head; ROIAlign's backward removes the scale while it reads (A/B switch)
        self._dx0_half = os.environ.get("MYOLO_DX0_HALF", "1") != "0" and precision == "h16"
        self._pw_win = os.environ.get("MYOLO_PW_WIN", "0") != "0"      # pointwise forward on the persistent window kernel (A/B)      # measured best: profiles/r02_fuse_bn_ab.txt
        self._deferred = {}
        self._plan = None
        self.t = 0                     # Adam iteration
        self.version = 0               # bumped whenever the weights change (load_params, apply_updates)
        self.seen = 0                  # yolo_custom_loss `seen` counter (model.py:95, 197)
        self._alloc_params(params if params is not None else init_params(self.NB, self.NC, seed))
        self._alloc_acts()
        self.refresh_weights()

    # ------------------------------------------------------------------ parameters
    def set_precision(self, precision: str):
        """fp32      : exact-fp32 CUDA-core GEMMs everywhere (parity configuration);
        tf32      : tcgen05 kind::tf32 everywhere, single pass (fastest, ~1e-3 relative per layer);
        tf32x3    : tcgen05 everywhere; the FORWARD GEMMs of backbone / yolo branch / feature_map run
                    as 3xTF32 (operands split hi+lo, A*B ~= Ah*Bh + Al*Bh + Ah*Bl: fp32-grade
                    outputs), mask head and every backward GEMM single-pass tf32;
        tf32x3_all: as tf32x3, with the mask-head forward convolutions in 3xTF32 as well;
        h16       : as tf32x3, with the mask head (98.8 % of the FLOPs) on tcgen05 kind::f16: activations and
                    staged weights stored as IEEE half (the same 10 explicit mantissa bits tf32 keeps, at
                    twice the tensor rate and half the bytes), fp32 accumulation and fp32 epilogues."""
        assert precision in ("fp32", "tf32", "tf32x3", "tf32x3_all", "h16")
        self.precision = precision
        self.tc = precision != "fp32"
        self.x3 = precision in ("tf32x3", "tf32x3_all", "h16")
        self.x3m = precision == "tf32x3_all"
        self.h16 = precision == "h16"
        self.rnd = C.ROUND_TF32 if self.tc else 0
        C.set_precision(C.PREC_TF32 if self.tc else C.PREC_FP32)

    def _alloc_params(self, P):
        specs = param_specs(self.NB, self.NC)
        self.specs = specs
        off, self.offs = 0, OrderedDict()
        for name, shape, tr in specs:
            if tr:
                n = int(np.prod(shape))
                self.offs[name] = (off, n, shape)
                off += (n + 63) // 64 * 64                    # 256-byte aligned slices: kernels read weights in
                                                              # place through TMA (dgrad B operand); a slice that
                                                              # straddles 128-byte lines costs 2x L2 requests
        self.n_flat = off
        z = lambda: torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.params, self.grads, self.adam_m, self.adam_v = z(), z(), z(), z()
        self.p, self.g = {}, {}
        for name, (o, n, shape) in self.offs.items():
            self.p[name] = self.params[o:o + n].view(shape)
            self.g[name] = self.grads[o:o + n].view(shape)
        self.trainable_mask = torch.ones(off, dtype=torch.float32, device=self.dev)
        self.stats = {name: torch.zeros(shape, dtype=torch.float32, device=self.dev) for name, shape, tr in specs if not tr}
        self.bn: Dict[str, _BN] = {}
        for name, shape, tr in specs:
            if name.endswith("/gamma"):
                b = _BN()
                b.name, b.c = name[:-6], shape[0]
                b.gamma, b.beta = self.p[b.name + "/gamma"], self.p[b.name + "/beta"]
                b.dgamma, b.dbeta = self.g[b.name + "/gamma"], self.g[b.name + "/beta"]
                b.mmean, b.mvar = self.stats[b.name + "/moving_mean"], self.stats[b.name + "/moving_variance"]
                b.bmean = torch.zeros(b.c, device=self.dev)     # zero-debiased accumulators ("biased")
                b.bvar = torch.zeros(b.c, device=self.dev)
                b.mean = torch.zeros(b.c, device=self.dev)      # this batch's statistics
                b.var = torch.zeros(b.c, device=self.dev)
                b.step = 0
                self.bn[b.name] = b
        # boundary of the "late" bucket (feature_map + mask head) inside the flat buffers
        self.tail_off = self.offs["feature_map/kernel"][0]
        self.load_params(P)
        # GEMM-side weight copies: per-tap transposed ([t][Cout][Cin], K-major B operand of the forward)
        self.wt = {}
        for name, (o, n, shape) in self.offs.items():
            if name.endswith("/kernel") and name != "conv1/kernel":
                self.wt[name] = torch.zeros(n * (3 if self._is_x3(name) else 1), dtype=torch.float32, device=self.dev)
        # conv_23 (model.py:271) on tcgen05: its N_BOX*(5+NC) output channels (27 / 35 / 430) are padded to a multiple of 32
        # with zero weights -- forward Bt [Npad][1024] (3xTF32 triple in the x3 modes), data-gradient Bt [1024][Npad]
        self.ny = self.NB * (5 + self.NC)
        self.ny_pad = (self.ny + 31) // 32 * 32
        if self.tc:
            self.wt["conv_23/kernel"] = torch.zeros(self.ny_pad * 1024 * (3 if self.x3 else 1), dtype=torch.float32, device=self.dev)
            self.w23_d = torch.zeros(1024 * self.ny_pad, dtype=torch.float32, device=self.dev)
            self.b23_pad = torch.zeros(self.ny_pad, dtype=torch.float32, device=self.dev)
        # "h16": IEEE-half staging of the mask-head weights.  fwd = per-tap transposed [t][Cout][Cin] (deconv: the
        # Keras layout is already [4*Cout][Cin]).
        # dgrad = the HWIO kernel as it is ([t][Cin][Cout] = Bt of the data-gradient GEMM; deconv: transposed [Cin][4*Cout]).
        self.wth, self.wth_d = {}, {}
        if self.h16:
            for name, (o, n, shape) in self.offs.items():
                if name.startswith("myolo_mask_conv") and name.endswith("/kernel") or name == "myolo_mask_deconv/kernel":
                    self.wth[name] = torch.zeros(n, dtype=torch.float16, device=self.dev)
                    if self.mode == "training":
                        self.wth_d[name] = torch.zeros(n, dtype=torch.float16, device=self.dev)
            self.gs = torch.tensor([1.0, 1.0, 0.0, 0.0], dtype=torch.float32, device=self.dev)   # loss scale {S, 1/S, scratch}
        self.ws = torch.zeros(8192, dtype=torch.float64, device=self.dev)       # BN family: zero between calls
        self.ws_y = torch.zeros(8192, dtype=torch.float64, device=self.dev)     # the same for the YOLO-branch backward stream
        self.ws_loss = torch.zeros(16, dtype=torch.float64, device=self.dev)
        self.anchors = torch.tensor(self.cfg["ANCHORS"], dtype=torch.float32, device=self.dev)
        self.class_w = torch.tensor(np.asarray(self.cfg["CLASS_WEIGHTS"], dtype=np.float32), device=self.dev)
        self.scales = C.float_array([self.cfg["OBJECT_SCALE"], self.cfg["NO_OBJECT_SCALE"], self.cfg["COORD_SCALE"],
                                     self.cfg["CLASS_SCALE"]])

    def _is_x3(self, name: str) -> bool:
        """Does the forward GEMM of this kernel run as 3xTF32?"""
        if name.startswith("myolo_mask_conv"):
            return self.x3m
        if name.startswith("myolo_mask"):
            return False            # deconv: single pass; the 1x1 mask conv runs inside the deconv kernel's epilogue
        return self.x3

    def load_params(self, P: Dict[str, torch.Tensor], strict: bool = True):
        for name, shape, tr in self.specs:
            if name not in P:
                if strict:
                    raise KeyError(name)
                continue
            src = torch.as_tensor(P[name], dtype=torch.float32).reshape(shape)
            (self.p[name] if tr else self.stats[name]).copy_(src)
        self.version = getattr(self, "version", 0) + 1
        if hasattr(self, "wt"):
            self.refresh_weights()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = OrderedDict()
        for name, shape, tr in self.specs:
            out[name] = (self.p[name] if tr else self.stats[name]).detach().cpu().clone()
        return out

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return OrderedDict((n, self.g[n].detach().cpu().clone()) for n in self.offs)

    def set_trainable(self, predicate, base: bool = False):
        """predicate(keras variable name) -> bool.  A frozen variable is left out of the optimizer altogether (its value
        and its Adam moments stay bit-for-bit; a frozen BatchNormalization layer also keeps its moving statistics), which
        is what Keras does with `layer.trainable = False` (model.py:1120-1155).
        base=True installs a PERSISTENT freeze that later calls can only narrow, never lift: the reference's
        set_trainable recurses into the nested 'yolo_model' but never touches the Model object itself, so
        `yolo_model.trainable = False` (yolo_trainable=False, model.py:854-868) survives train(layers='all')."""
        if base:
            self._base_trainable = {name for name in self.offs if predicate(name)}
            predicate = lambda name: True                                           # noqa: E731
        allowed = getattr(self, "_base_trainable", None)
        mask = torch.zeros(self.n_flat, dtype=torch.float32)
        self._frozen = False
        self._frozen_names = set()
        for name, (o, n, _) in self.offs.items():
            if predicate(name) and (allowed is None or name in allowed):
                mask[o:o + n] = 1.0
            else:
                self._frozen = True
                self._frozen_names.add(name)
        self.trainable_mask.copy_(mask)
        self._moving_key = None            # the set of BN layers whose moving statistics advance may have changed

    def reset_optimizer(self):
        """Fresh Adam state: the reference builds a new keras.optimizers.Adam in every compile() (model.py:1071-1075),
        i.e. iteration count and both moment estimates start from zero."""
        self.t = 0
        self.adam_m.zero_()
        self.adam_v.zero_()

    def _prep_jobs(self):
        """(src tensor, dst tensor, ntaps, rows, cols, transpose, mode, out_ld, out_total) for every GEMM-side weight copy;
        fields as in myolo_prep_weights_batch."""
        jobs = []
        for name, buf in self.wt.items():
            mode = 2 if self._is_x3(name) else (1 if self.tc else 0)
            shape = self.offs[name][2]
            if name == "conv_23/kernel":
                if self.tc:     # zero-padded staging: [Npad][1024] per copy
                    jobs.append((self.p[name], buf, 1, 1024, self.ny, 1, mode, 0, self.ny_pad * 1024))
                    jobs.append((self.p[name], self.w23_d, 1, 1024, self.ny, 0, 1, self.ny_pad, 0))
                    jobs.append((self.p["conv_23/bias"], self.b23_pad, 1, 1, self.ny, 0, 0, 0, 0))
                else:           # exact CUDA-core kernel: plain transposed copy
                    jobs.append((self.p[name], buf, 1, 1024, self.ny, 1, 0, 0, 0))
            elif name == "myolo_mask_deconv/kernel":
                # Keras [2,2,Cout,Cin] is already the forward Bt ([N=4*Cout][K=Cin]); stage its
                # transpose [Cin][4*Cout] for the dgrad GEMM.
                jobs.append((self.p[name], buf, 1, 4 * shape[2], shape[3], 1, mode, 0, 0))
            else:
                jobs.append((self.p[name], buf, shape[0] * shape[1], shape[2], shape[3], 1, mode, 0, 0))
        for name, buf in self.wth.items():
            shape = self.offs[name][2]
            if name == "myolo_mask_deconv/kernel":
                jobs.append((self.p[name], buf, 1, 4 * shape[2], shape[3], 0, 3, 0, 0))
            else:
                jobs.append((self.p[name], buf, shape[0] * shape[1], shape[2], shape[3], 1, 3, 0, 0))
        for name, buf in self.wth_d.items():
            shape = self.offs[name][2]
            if name == "myolo_mask_deconv/kernel":
                jobs.append((self.p[name], buf, 1, 4 * shape[2], shape[3], 1, 3, 0, 0))
            else:
                jobs.append((self.p[name], buf, shape[0] * shape[1], shape[2], shape[3], 0, 3, 0, 0))
        if self.h16:    # the tf32 staging of the mask-head weights is never read in h16 mode
            jobs = [j for j in jobs if not (j[6] != 3 and any(j[1] is self.wt[n] for n in self.wth))]
        return jobs

    def refresh_weights(self):
        """Re-stage the transposed / rounded / split / half GEMM weight operands after an update: one launch over a
        device-side job table (built once; the buffers never move)."""
        if getattr(self, "_prep_table", None) is None:
            import struct
            rec, tiles = b"", 0
            jobs = self._prep_jobs()
            for src, dst, ntaps, rows, cols, tr, mode, out_ld, out_total in jobs:
                rec += struct.pack("<QQiiiiiiii", src.data_ptr(), dst.data_ptr(), ntaps, rows, cols, tr, mode, tiles, out_ld, out_total)
                tiles += ntaps * ((rows + 31) // 32) * ((cols + 31) // 32)
            self._prep_table = torch.frombuffer(bytearray(rec), dtype=torch.uint8).to(self.dev)
            self._prep_n, self._prep_tiles = len(jobs), tiles
        C.call("myolo_prep_weights_batch", self._prep_table, self._prep_n, self._prep_tiles, self._st())

    # ------------------------------------------------------------------ activations
    def _alloc_acts(self):
        B, S, dev = self.B, self.cfg["S"], self.dev
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        A = {}
        H = S // 2
        A["y0"], A["a0"] = f(B, H, H, 32), f(B, H, H, 32)
        self.geo = {}
        maxel = B * H * H * 32
        for k, ci, co, s in BACKBONE_BLOCKS + YOLO_BLOCKS:
            Ho = (H - 1) // s + 1
            self.geo[k] = (H, Ho, ci, co, s)
            A[f"yd{k}"] = f(B, Ho, Ho, ci)
            # pointwise-GEMM operand; in 3xTF32 mode [0] = tf32 high part, [1] = low part
            A[f"ad{k}"] = f(2, B, Ho, Ho, ci) if self.x3 else f(B, Ho, Ho, ci)
            A[f"yp{k}"] = f(B, Ho, Ho, co)
            if k == 6 and self.with_mask:
                self.c4 = PF(B, Ho, Ho, co, device=dev, split=self.x3)   # C4 feeds the 3x3 feature_map conv
                if self.x3:
                    A["ap6"] = f(B, Ho, Ho, co)                   # full-precision C4 for the depthwise consumer
            elif k == 14 and self.x3:
                A[f"ap{k}"] = f(2, B, Ho, Ho, co)                 # hi / lo operand pair of the 3xTF32 conv_23 GEMM
            else:
                A[f"ap{k}"] = f(B, Ho, Ho, co)
            maxel = max(maxel, B * H * H * ci, B * Ho * Ho * co)
            H = Ho
        self.F = self.geo[6][1]
        G, NB, NC, R = self.cfg["G"], self.NB, self.NC, self.R
        assert H == G
        A["yolo"] = f(B, G, G, NB * (5 + NC))
        if self.tc:     # tensor-core conv_23: result / gradient with the channel count padded to a multiple of 32
            A["yolo_pad"] = torch.zeros(B * G * G, self.ny_pad, dtype=torch.float32, device=dev)
            if self.mode != "inference":
                A["dyolo_pad"] = torch.zeros(B * G * G, self.ny_pad, dtype=torch.float32, device=dev)
        A["proposals"] = f(B, R, 4)
        A["detections"] = f(B, R, 6)
        self.loss_yolo = torch.zeros(5, device=dev)
        self.loss_mask = torch.zeros(1, device=dev)
        if self.mode != "inference":
            A["dyolo"] = f(B, G, G, NB * (5 + NC))
            self.gx, self.gy = f(maxel), f(maxel)                 # backbone gradient ping-pong
        if self.with_mask:
            F_, P_ = self.F, self.cfg["POOL"]
            n = B * R
            self.n_roi = n
            self.feat = PF(B, F_, F_, MASK_C, device=dev)
            # pre-BN conv outputs: only conv1 (batch-statistics BN) needs one; conv2..4 fold their
            # fixed-statistics BN + ReLU into the GEMM epilogue (unless the 3xTF32 mask mode splits them)
            self.my = [None] + [PF(n, P_, P_, MASK_C, device=dev) if ((i == 1 and not self._bn1_half) or self.x3m) else None
                                for i in range(1, 5)]
            self.z1h = PF(n, P_, P_, MASK_C, device=dev, dtype=torch.float16) if self._bn1_half else None
            self.bn_scale = torch.zeros(4, MASK_C, device=dev)
            self.bn_shift = torch.zeros(4, MASK_C, device=dev)
            if self.h16:     # the conv operands x0, a1..a4 exist as IEEE half only
                self.mah = [PF(n, P_, P_, MASK_C, device=dev, dtype=torch.float16) for _ in range(5)]
                self.x0, self.ma = None, [None] * 5
            else:
                self.x0 = PF(n, P_, P_, MASK_C, device=dev, split=self.x3m)
                self.ma = [self.x0] + [PF(n, P_, P_, MASK_C, device=dev, split=self.x3m and i < 4) for i in range(1, 5)]
            self.y4d = PF(n, P_, P_, 4 * MASK_C, device=dev)
            mh, mw = self.cfg["MASK_SHAPE"]
            assert (mh, mw) == (2 * P_, 2 * P_)
            A["masks"] = f(n, mh, mw, NC)
            if self.mode == "training":
                A["rois"] = f(B, R, 4)
                self.target_ids = torch.zeros(B, R, dtype=torch.int32, device=dev)
                A["target_masks"] = f(B, R, mh, mw)
                self.n_pos = torch.zeros(B, dtype=torch.int32, device=dev)
                self.roi_src = torch.zeros(B, R, dtype=torch.int32, device=dev)
                self.roi_gt = torch.zeros(B, R, dtype=torch.int32, device=dev)
                A["dlogit"] = f(n, mh, mw, NC)
                if self.h16:     # loss-scaled half gradients; fp32 only for d(a1) (batch-statistics BN) and d(x0) (ROIAlign)
                    self.dy4h = PF(n, P_, P_, 4 * MASK_C, device=dev, dtype=torch.float16)
                    self.dy4h_ids = torch.zeros(n, dtype=torch.int32, device=dev)   # which rois' rows of dy4h are non-zero
                    self.mgh = [PF(n, P_, P_, MASK_C, device=dev, dtype=torch.float16) for _ in range(4 if self._wstream is not None else 2)]
                else:
                    self.dy4d = PF(n, P_, P_, 4 * MASK_C, device=dev)
                if self.h16 and self.sparse_backward:
                    # compact padded-flat tensors for up to pcap positive rois (more than that: the step runs dense)
                    self.pcap = max(64, n // 8)
                    hp = lambda c: PF(self.pcap, P_, P_, c, device=dev, dtype=torch.float16)      # noqa: E731
                    self.sp_a = [None] + [hp(MASK_C) for _ in range(4)]        # a1..a4 of the positive rois
                    self.sp_dy4 = hp(4 * MASK_C)
                    self.sp_g = [hp(MASK_C) for _ in range(2)]
                    self.sp_list = torch.zeros(self.pcap, dtype=torch.int32, device=dev)
                    self.sp_list_host = torch.zeros(self.pcap, dtype=torch.int32).pin_memory()
                if self.h16 and self._dx0_half:
                    self.mg = [None, PF(n, P_, P_, MASK_C, device=dev, dtype=torch.float16)]
                else:
                    self.mg = [PF(n, P_, P_, MASK_C, device=dev) for _ in range(1 if self.h16 else 2)]
                    if self.h16:
                        self.mg = [None, self.mg[0]]
                self.dfeat = PF(B, F_, F_, MASK_C, device=dev)
                self.dc4 = PF(B, F_, F_, 512, device=dev)
        self.A = A

    # ------------------------------------------------------------------ small helpers
    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _v(t):
        n, h, w, c = t.shape
        return C.view(t, n, h, w, c)

    def _bn_fwd(self, name, xv, yv, act, training, n_pix, yv_lo=None, stats=True, apply=True):
        """stats=False: the producing kernel has already left the batch statistics in b.mean / b.var (epilogue fusion);
        apply=False: the consumer applies the normalisation + activation while it loads (no post-BN tensor)."""
        b, st = self.bn[name], self._st()
        if training:
            if stats:
                C.call("myolo_bn_stats", xv, b.mean, b.var, self.ws, st)
            self._bn_touched.append((b, n_pix))
            mean, var = b.mean, b.var
        else:
            mean, var = b.mmean, b.mvar
        if not apply:
            return
        if yv_lo is None:
            C.call("myolo_bn_apply", xv, yv, mean, var, b.gamma, b.beta, BN_EPS, act, st)
        else:
            C.call("myolo_bn_apply_split", xv, yv, yv_lo, mean, var, b.gamma, b.beta, BN_EPS, act & 0xff, st)

    def _gemm_fwd(self, a_rows, lo_off, name, out_rows, M, N, K, shifts, bias, pf_w1, pf_blk, scale=None, shift=None,
                  act=0, stream=None):
        """Forward conv GEMM through the tap-GEMM entry point; 3xTF32 = the tap list tripled over the
        (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo) operand pairs (lo_off = row distance hi -> lo)."""
        base = list(shifts) if shifts is not None else [0]
        if self._is_x3(name):
            sh = base + [v + lo_off for v in base] + base
        else:
            sh = base
        key = (name, tuple(sh))
        arr = self._shift_cache.get(key)
        if arr is None:
            arr = self._shift_cache[key] = C.int_array(sh)
        timed = name.startswith("myolo_mask_conv")
        if timed:
            C.record_py(self._ke_begin)
        C.call("myolo_gemm_taps", a_rows, K, self.wt[name], out_rows, N, M, N, K, len(sh), arr, bias, scale, shift,
               act, pf_w1, pf_blk, 0, self._st() if stream is None else stream)
        if timed:
            C.record_py(self._ke_end)

    def _gemm_fwd_stats(self, a_rows, lo_off, name, out_rows, M, N, K, b):
        """Pointwise forward GEMM with the batch statistics of its result (BN layer `b`) reduced in the epilogue.  3xTF32
        layers with M >= 4096 and N % 128 == 0 run on the persistent window kernel as three k-segments (TMA-store epilogue,
        CTA pairs); the rest on the one-tile kernel with the operand triple as taps, as in _gemm_fwd."""
        x3 = self._is_x3(name)
        sh = [0, lo_off, 0] if x3 else [0]
        key = (name, tuple(sh))
        arr = self._shift_cache.get(key)
        if arr is None:
            arr = self._shift_cache[key] = C.int_array(sh)
        if x3 and self._pw_win and C.lib().myolo_gemm_segs_win_supported(K, N, M, N, K, 3):
            C.call("myolo_gemm_segs_win", a_rows, K, self.wt[name], out_rows, N, M, N, K, 3, arr, b.mean, b.var, self.ws, self._st())
            return
        C.call("myolo_gemm_taps_tc_stats", a_rows, K, self.wt[name], out_rows, N, M, N, K, len(sh), arr, 0, 0, b.mean, b.var,
               self.ws, M, self._st())

    def _phase(self, name):
        """Diagnostic (MYOLO_PHASES=1, scripts/phase_timeline.sh): a timing event on the step's main stream at a phase
        boundary; recorded as a host action, so it is re-recorded at every replay.  phase_times() reads the last step's."""
        if not self._phases_on:
            return
        e = self._phase_evs.get(name)
        if e is None:
            e = self._phase_evs[name] = torch.cuda.Event(enable_timing=True)
            self._phase_order.append(name)
        main = torch.cuda.current_stream()
        C.record_py(lambda: e.record(main))

    def phase_times(self):
        """[(phase, ms since the previous boundary)] of the last completed step (synchronises)."""
        torch.cuda.synchronize()
        ev = [(n, self._phase_evs[n]) for n in self._phase_order]
        return [(n1, e0.elapsed_time(e1)) for (_, e0), (n1, e1) in zip(ev[:-1], ev[1:])]

    def _ke_begin(self):
        """bench.py hook: CUDA events around the dominant kernel (only while kernel_events is a list)"""
        if self.kernel_events is not None:
            self._ke0 = torch.cuda.Event(enable_timing=True)
            self._ke0.record()

    def _ke_end(self):
        if self.kernel_events is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.kernel_events.append((self._ke0, e1))

    def _bn_bwd(self, name, xv, dyv, act, training, st=None, ws=None):
        b = self.bn[name]
        mean, var = (b.mean, b.var) if training else (b.mmean, b.mvar)
        C.call("myolo_bn_bwd", xv, dyv, dyv, mean, var, b.gamma, b.beta, BN_EPS, act, 1 if training else 0,
               b.dgamma, b.dbeta, self.ws if ws is None else ws, self._st() if st is None else st)

    # ------------------------------------------------------------------ forward
    def forward(self, image: torch.Tensor, training: Optional[bool] = None):
        """image [B,S,S,3] fp32 on the device, values in 0..1 (BatchGenerator divides by 255).
        Backbone + feature_map + yolo branch + decode (model.py:844-874, 922-926)."""
        training = (self.mode != "inference") if training is None else training
        A, B, st = self.A, self.B, self._st()
        S = self.cfg["S"]
        assert tuple(image.shape) == (B, S, S, 3) and image.is_cuda and image.dtype == torch.float32
        self._bn_touched: List = []
        self._image = image
        relu6 = C.ACT_RELU6
        # Fusions of the backbone's BatchNormalization layers (training phase only):
        #   * batch statistics of a depthwise / pointwise output are reduced in the epilogue of the kernel that produces it;
        #   * BN + ReLU6 after conv1 and after a pointwise conv is applied by the NEXT depthwise kernel while it stages its
        #     input (and again by that layer's filter-gradient kernel in the backward pass): the post-BN activation is
        #     never written, except where a GEMM reads it through TMA (block 6 -> feature_map, block 14 -> conv_23).
        fm = self._fuse_bn if training else 0
        fuse, dw_stats = bool(fm & 2), bool(fm & 1)
        self._deferred = {}          # block id (0 = conv1) -> BN layer whose apply was left to the consumer
        # conv_block (model.py:42-52)
        C.call("myolo_conv1_fwd", image, self.p["conv1/kernel"], A["y0"], B, S, 32, st)
        self._bn_fwd("conv1_bn", self._v(A["y0"]), self._v(A["a0"]), relu6, training, B * (S // 2) ** 2, apply=not fuse)
        xin_view = self._v(A["y0"] if fuse else A["a0"])
        in_bn = self.bn["conv1_bn"] if fuse else None
        if fuse:
            self._deferred[0] = in_bn
        for k, ci, co, s in BACKBONE_BLOCKS + YOLO_BLOCKS:
            Hi, Ho, _, _, _ = self.geo[k]
            npix = B * Ho * Ho
            # _depthwise_conv_block: ZeroPad(1,1) + depthwise 3x3 VALID stride s -> BN -> ReLU6
            bd = self.bn[f"conv_dw_{k}_bn"]
            if fuse or dw_stats:
                ib = in_bn
                C.call("myolo_dwconv3x3_fwd_bn", xin_view, self.p[f"conv_dw_{k}/depthwise_kernel"], A[f"yd{k}"], s,
                       ib.mean if ib else None, ib.var if ib else None, ib.gamma if ib else None, ib.beta if ib else None,
                       BN_EPS, relu6, bd.mean if dw_stats else None, bd.var if dw_stats else None,
                       self.ws if dw_stats else None, st)
            else:
                C.call("myolo_dwconv3x3_fwd", xin_view, self.p[f"conv_dw_{k}/depthwise_kernel"], A[f"yd{k}"], s, st)
            ad = A[f"ad{k}"]
            if self.x3:
                self._bn_fwd(f"conv_dw_{k}_bn", self._v(A[f"yd{k}"]), self._v(ad[0]), relu6, training, npix, self._v(ad[1]),
                             stats=not dw_stats)
            else:
                self._bn_fwd(f"conv_dw_{k}_bn", self._v(A[f"yd{k}"]), self._v(ad), relu6 | self.rnd, training, npix, stats=not dw_stats)
            # pointwise 1x1 -> BN -> ReLU6
            pw_stats = bool(fm & 4) and self.tc
            if pw_stats:
                self._gemm_fwd_stats(ad, npix, f"conv_pw_{k}/kernel", A[f"yp{k}"], npix, co, ci, self.bn[f"conv_pw_{k}_bn"])
            else:
                self._gemm_fwd(ad, npix, f"conv_pw_{k}/kernel", A[f"yp{k}"], npix, co, ci, None, None, 0, 0)
            defer = fuse and k not in (6, 14)
            in_bn = None
            if k == 6 and self.with_mask:
                if self.x3:
                    out_view = self._v(A["ap6"])
                    self._bn_fwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), out_view, relu6, training, npix, stats=not pw_stats)
                    C.call("myolo_split_tf32", out_view, self.c4.view(), self.c4.view(lo=True), st)
                else:
                    out_view = self.c4.view()
                    self._bn_fwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), out_view, relu6 | self.rnd, training, npix,
                                 stats=not pw_stats)
            elif defer:
                self._bn_fwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), None, relu6, training, npix, stats=not pw_stats, apply=False)
                out_view = self._v(A[f"yp{k}"])
                in_bn = self.bn[f"conv_pw_{k}_bn"]
                self._deferred[k] = in_bn
            elif k == 14 and self.x3:
                out_view = self._v(A["ap14"][0])
                self._bn_fwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), out_view, relu6, training, npix, self._v(A["ap14"][1]),
                             stats=not pw_stats)
            else:
                out_view = self._v(A[f"ap{k}"])
                self._bn_fwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), out_view, relu6 | (self.rnd if k == 14 else 0), training,
                             npix, stats=not pw_stats)
            xin_view = out_view
            if k == 6 and self.with_mask:
                # myolo_feature_maps = Conv2D(256, 3x3, SAME)(C4) + bias   (model.py:848)
                # (on the side stream: 225 one-tile CTAs = 1.5 waves, next to the small kernels of the yolo branch;
                # mask_head() joins before ROIAlign reads the feature map)
                F_ = self.F
                fm_stream = None
                if self._side is not None and os.environ.get("MYOLO_FM_OVERLAP", "1") != "0":
                    e = self._evs.setdefault("fm_fork", torch.cuda.Event())
                    cur, side_ = torch.cuda.current_stream(), self._side
                    C.record_py(lambda: (e.record(cur), side_.wait_event(e)))
                    fm_stream = self._side.cuda_stream
                self._gemm_fwd(self.c4.rows, self.c4.lo_off, "feature_map/kernel", self.feat.rows, self.c4.M, MASK_C, 512,
                               conv3x3_shifts(F_), self.p["feature_map/bias"], F_ + 1, (F_ + 1) * (F_ + 1), stream=fm_stream)
                if fm_stream is not None:
                    e2, side_ = self._evs.setdefault("fm_done", torch.cuda.Event()), self._side
                    C.record_py(lambda: e2.record(side_))
                    self._fm_pending = True
        G, NB, NC = self.cfg["G"], self.NB, self.NC
        # conv_23 (model.py:271) + reshape [B,G,G,NB,5+NC] (273): a pure view of the NHWC result
        ny, m23 = self.ny, B * G * G
        if self.tc:
            # tcgen05 tiles (3xTF32 operand triple in the x3 modes) on the zero-padded channel count, bias in the epilogue;
            # the dense [.., ny] tensor the decode / loss kernels read is then cut out of the padded result
            sh = [0, m23, 0] if self.x3 else [0]
            arr = self._shift_cache.get(("c23", m23, self.x3))
            if arr is None:
                arr = self._shift_cache[("c23", m23, self.x3)] = C.int_array(sh)
            C.call("myolo_gemm_taps_tc", A["ap14"], 1024, self.wt["conv_23/kernel"], A["yolo_pad"], self.ny_pad, m23, self.ny_pad,
                   1024, len(sh), arr, self.b23_pad, None, None, C.ACT_NONE, 0, 0, 0, st)
            C.call("myolo_copy_cols", A["yolo_pad"], self.ny_pad, A["yolo"], ny, m23, ny, st)
        else:
            C.call("myolo_gemm_taps_ffma", A["ap14"], 1024, self.wt["conv_23/kernel"], A["yolo"], ny, m23, ny, 1024, 1, None,
                   self.p["conv_23/bias"], None, None, C.ACT_NONE, 0, 0, 0, st)
        # DecodeYOLOLayer / DetectionsLayer (model.py:1442-1473, 1493-1538)
        C.call("myolo_yolo_decode", A["yolo"], self.anchors, A["proposals"], A["detections"], B, G, G, NB, NC, st)
        return A["yolo"].view(B, G, G, NB, 5 + NC)

    def mask_head(self, rois: torch.Tensor, training: bool):
        """build_mask_graph (model.py:668-715) on rois [B,R,4] (x1,y1,x2,y2, passed to ROIAlign as-is:
        SURVEY Q2).  bn1 follows the learning phase, bn2..4 always use moving statistics."""
        A, st, n = self.A, self._st(), self.n_roi
        P_ = self.cfg["POOL"]
        npix = n * P_ * P_
        if getattr(self, "_fm_pending", False):      # the feature_map conv ran on the side stream
            cur, e2 = torch.cuda.current_stream(), self._evs["fm_done"]
            C.record_py(lambda: cur.wait_event(e2))
            self._fm_pending = False
        if self.h16:
            return self._mask_head_h16(rois, training)
        C.call("myolo_roialign_fwd", self.feat.view(), rois, n, self.R, P_, self.x0.view(),
               1 if (self.rnd and not self.x3m) else 0, st)
        if self.x3m:
            C.call("myolo_split_tf32", self.x0.view(), self.x0.view(), self.x0.view(lo=True), st)
        sh3 = conv3x3_shifts(P_)
        self._mask_fused = [False] * 5
        for i in (1, 2, 3, 4):
            a_in = self.ma[i - 1]
            batch_stats = training and i == 1
            if batch_stats or self.x3m:
                self._gemm_fwd(a_in.rows, a_in.lo_off, f"myolo_mask_conv{i}/kernel", self.my[i].rows, a_in.M, MASK_C, MASK_C,
                               sh3, self.p[f"myolo_mask_conv{i}/bias"], P_ + 1, (P_ + 1) * (P_ + 1))
                lo = self.ma[i].view(lo=True) if self.ma[i].rows_lo is not None else None
                self._bn_fwd(f"myolo_mask_bn{i}", self.my[i].view(), self.ma[i].view(), C.ACT_RELU | self.rnd,
                             batch_stats, npix, lo)
            else:
                # fixed-statistics BN + ReLU folded into the conv epilogue: a_i = relu((conv+bias)*scale + shift)
                b = self.bn[f"myolo_mask_bn{i}"]
                C.call("myolo_bn_fold", b.gamma, b.beta, b.mmean, b.mvar, BN_EPS, self.bn_scale[i - 1], self.bn_shift[i - 1],
                       MASK_C, st)
                self._gemm_fwd(a_in.rows, a_in.lo_off, f"myolo_mask_conv{i}/kernel", self.ma[i].rows, a_in.M, MASK_C, MASK_C,
                               sh3, self.p[f"myolo_mask_conv{i}/bias"], P_ + 1, (P_ + 1) * (P_ + 1),
                               self.bn_scale[i - 1], self.bn_shift[i - 1], C.ACT_RELU | self.rnd)
                self._mask_fused[i] = True
        # Conv2DTranspose 2x2 s2 as one GEMM [rows,256] x [256, 4*256]; bias/ReLU/1x1/sigmoid in mask_out
        a4 = self.ma[4]
        if self.tc and C.lib().myolo_deconv_mask_fwd_supported(MASK_C, self.NC):
            # one tcgen05 kernel: deconv GEMM + bias + ReLU + 1x1 conv + sigmoid; the deconv activation is
            # kept only for positive rois (all the backward pass reads)
            ids = self.target_ids if self.mode == "training" else None
            C.call("myolo_deconv_mask_fwd", a4.rows, self.p["myolo_mask_deconv/kernel"], self.p["myolo_mask_deconv/bias"],
                   self.p["myolo_mask/kernel"], self.p["myolo_mask/bias"], A["masks"], ids, self.y4d.rows, n, P_, P_,
                   MASK_C, self.NC, st)
        else:
            C.call("myolo_gemm_taps", a4.rows, MASK_C, self.p["myolo_mask_deconv/kernel"], self.y4d.rows, 4 * MASK_C, a4.M,
                   4 * MASK_C, MASK_C, 1, None, None, None, None, C.ACT_NONE, P_ + 1, (P_ + 1) * (P_ + 1), 0, st)
            C.call("myolo_mask_out_fwd", self.y4d.rows, self.p["myolo_mask_deconv/bias"], self.p["myolo_mask/kernel"],
                   self.p["myolo_mask/bias"], A["masks"], n, P_, P_, MASK_C, self.NC, st)
        mh, mw = self.cfg["MASK_SHAPE"]
        return A["masks"].view(self.B, self.R, mh, mw, self.NC)

    def _mask_head_h16(self, rois: torch.Tensor, training: bool):
        """mask_head on tcgen05 kind::f16: every conv reads IEEE-half activations / weights, accumulates in fp32 and
        stores its result as half (what the next conv and the whole backward pass read)."""
        A, st, n = self.A, self._st(), self.n_roi
        P_ = self.cfg["POOL"]
        npix = n * P_ * P_
        pfw, pfb = P_ + 1, (P_ + 1) * (P_ + 1)
        C.call("myolo_roialign_fwd_h", self.feat.view(), rois, n, self.R, P_, None, self.mah[0].view(), st)
        key = ("h16", P_)
        sh3 = self._shift_cache.get(key)
        if sh3 is None:
            sh3 = self._shift_cache[key] = C.int_array(conv3x3_shifts(P_))
        self._mask_fused = [False] * 5
        M = self.mah[0].M
        for i in (1, 2, 3, 4):
            name = f"myolo_mask_conv{i}/kernel"
            a_in = self.mah[i - 1]
            C.record_py(self._ke_begin)
            if training and i == 1:
                # batch-statistics BN: the pre-BN tensor is kept in fp32 (statistics, backward); its batch statistics are
                # reduced in the conv's epilogue (bit 8 of MYOLO_FUSE_BN)
                if self._bn1_half:
                    b1 = self.bn["myolo_mask_bn1"]
                    C.call("myolo_gemm_taps_hh_stats", a_in.rows, MASK_C, self.wth[name], self.z1h.rows, MASK_C, M, MASK_C,
                           MASK_C, 9, sh3, self.p[f"myolo_mask_conv{i}/bias"], pfw, pfb, b1.mmean, b1.mean, b1.var, self.ws, npix, st)
                elif self._fuse_bn & 8:
                    b1 = self.bn["myolo_mask_bn1"]
                    C.call("myolo_gemm_taps_h_stats", a_in.rows, MASK_C, self.wth[name], self.my[1].rows, MASK_C, M, MASK_C,
                           MASK_C, 9, sh3, self.p[f"myolo_mask_conv{i}/bias"], pfw, pfb, b1.mmean, b1.mean, b1.var, self.ws, npix, st)
                else:
                    C.call("myolo_gemm_taps_h", a_in.rows, MASK_C, self.wth[name], self.my[1].rows, MASK_C, None, 0, M, MASK_C,
                           MASK_C, 9, sh3, self.p[f"myolo_mask_conv{i}/bias"], None, None, C.ACT_NONE, pfw, pfb, None, st)
            else:
                b = self.bn[f"myolo_mask_bn{i}"]
                C.call("myolo_bn_fold", b.gamma, b.beta, b.mmean, b.mvar, BN_EPS, self.bn_scale[i - 1], self.bn_shift[i - 1],
                       MASK_C, st)
                C.call("myolo_gemm_taps_h", a_in.rows, MASK_C, self.wth[name], None, 0, self.mah[i].rows, MASK_C,
                       M, MASK_C, MASK_C, 9, sh3, self.p[f"myolo_mask_conv{i}/bias"], self.bn_scale[i - 1],
                       self.bn_shift[i - 1], C.ACT_RELU, pfw, pfb, None, st)
                self._mask_fused[i] = True
            C.record_py(self._ke_end)
            if training and i == 1:
                b = self.bn["myolo_mask_bn1"]
                if not (self._fuse_bn & 8) and not self._bn1_half:
                    C.call("myolo_bn_stats", self.my[1].view(), b.mean, b.var, self.ws, st)
                self._bn_touched.append((b, npix))
                if self._bn1_half:
                    C.call("myolo_bn_apply_hh", self.z1h.view(), self.mah[1].view(), b.mean, b.var, b.gamma, b.beta, BN_EPS,
                           C.ACT_RELU, st)
                else:
                    C.call("myolo_bn_apply_h", self.my[1].view(), None, self.mah[1].view(), b.mean, b.var, b.gamma,
                           b.beta, BN_EPS, C.ACT_RELU, st)
        if C.lib().myolo_deconv_mask_fwd_supported(MASK_C, self.NC):
            ids = self.target_ids if self.mode == "training" else None
            C.call("myolo_deconv_mask_fwd_h", self.mah[4].rows, self.wth["myolo_mask_deconv/kernel"], self.p["myolo_mask_deconv/bias"],
                   self.p["myolo_mask/kernel"], self.p["myolo_mask/bias"], A["masks"], ids, self.y4d.rows, n, P_, P_,
                   MASK_C, self.NC, st)
        else:       # many classes (NC > 7): deconv GEMM to y4, then the mask tail as its own kernel
            C.call("myolo_gemm_taps_h", self.mah[4].rows, MASK_C, self.wth["myolo_mask_deconv/kernel"], self.y4d.rows, 4 * MASK_C,
                   None, 0, M, 4 * MASK_C, MASK_C, 1, None, None, None, None, C.ACT_NONE, pfw, pfb, None, st)
            C.call("myolo_mask_out_fwd", self.y4d.rows, self.p["myolo_mask_deconv/bias"], self.p["myolo_mask/kernel"],
                   self.p["myolo_mask/bias"], A["masks"], n, P_, P_, MASK_C, self.NC, st)
        mh, mw = self.cfg["MASK_SHAPE"]
        return A["masks"].view(self.B, self.R, mh, mw, self.NC)

    def forward_inference(self, image):
        """mode='inference' graph (model.py:922-936) -> [yolo_output, detections, myolo_mask]."""
        yolo = self.forward(image, training=False)
        det = self.A["detections"]
        # detection_boxes = detections[..., :4] (model.py:927) == the decoded proposals buffer
        masks = self.mask_head(self.A["proposals"], training=False) if self.with_mask else None
        return yolo, det, masks

    def postprocess(self, top_k: int = 10, cs_threshold: float = 0.35, nms_threshold: float = 0.5):
        """Device-side tail of MaskYOLO.detect (model.py:1290-1304, 1330-1391) on the detections / masks left by
        forward_inference: top-k by confidence, threshold, NMB, mask paste.  Returns device tensors
        (index [B,K], boxes [B,K,4] int32 px, class [B,K], score [B,K], count [B], masks [B,K,S,S] uint8)."""
        B, R, S, dev = self.B, self.R, self.cfg["S"], self.dev
        mh, mw = self.cfg["MASK_SHAPE"]
        i32 = lambda *sh: torch.empty(sh, dtype=torch.int32, device=dev)
        idx, boxes, cls, cnt = i32(B, top_k), i32(B, top_k, 4), i32(B, top_k), i32(B)
        score = torch.empty(B, top_k, dtype=torch.float32, device=dev)
        pm = torch.empty(B, top_k, S, S, dtype=torch.uint8, device=dev) if self.with_mask else None
        C.call("myolo_detect_postprocess", self.A["detections"], self.A["masks"] if self.with_mask else None, B, R, self.NC, S,
               mh, mw, top_k, float(cs_threshold), float(nms_threshold), idx, boxes, cls, score, cnt, pm, self._st())
        return idx, boxes, cls, score, cnt, pm

    def _wait_inputs(self):
        if self.inputs_ready is not None:
            torch.cuda.current_stream().wait_event(self.inputs_ready)
            self.inputs_ready = None

    def forward_training(self, inputs, learning_phase: bool = True, start_backward: bool = False):
        """mode='training' graph (model.py:844-901); learning_phase=False evaluates the same graph the way Keras
        validates (every BN on its moving statistics).  inputs as BatchGenerator yields them:
        [image, true_boxes [B,1,1,1,TB,4], yolo_target [B,G,G,NB,5+NC], gt_class_ids [B,M] i32,
        gt_boxes [B,M,4] px (x1,y1,x2,y2), gt_masks [B,S,S,M] bool] -- device tensors."""
        image, true_boxes, yolo_target = inputs[0], inputs[1], inputs[2]
        A, B, st, cfg = self.A, self.B, self._st(), self.cfg
        G, NB, NC, TB, R = cfg["G"], self.NB, self.NC, self.TB, self.R
        self._phase("start")
        yolo = self.forward(image, training=learning_phase)
        self._phase("backbone + yolo branch forward, decode")
        C.record_py(self._wait_inputs)         # ground-truth tensors still in flight on the caller's copy stream
        self.seen += 1
        warm = 1 if self.seen < cfg.get("WARM_UP_BATCHES", 0) else 0
        lw = cfg.get("LOSS_WEIGHTS", {})
        assert true_boxes.dtype == torch.float32 and yolo_target.dtype == torch.float32
        C.call("myolo_yolo_loss", yolo_target, A["yolo"], true_boxes, self.anchors, self.class_w, B, G, G, NB, NC, TB,
               self.scales, warm, float(lw.get("yolo_sum_loss", 1.0)), self.loss_yolo, A["dyolo"], self.ws_loss, st)
        out = dict(yolo_output=yolo, yolo_proposals=A["proposals"], yolo_sum_loss=self.loss_yolo[0])
        self._y_started = False
        if start_backward and learning_phase and self.with_mask and self._ystream is not None and self._y_early:
            # train_step: the YOLO branch's backward (conv_23, blocks 14..7) needs nothing but the yolo loss, so it starts
            # HERE on its own stream, next to the mask head's FORWARD and backward (12 ms of persistent tensor-core kernels)
            # instead of next to the backward alone; Engine.backward picks the chain up again at block 6
            self._start_yolo_branch_backward()
        if self.with_mask:
            gt_ids, gt_boxes, gt_masks = inputs[3], inputs[4], inputs[5]
            M = gt_ids.shape[1]
            mh, mw = cfg["MASK_SHAPE"]
            assert gt_ids.dtype == torch.int32 and gt_boxes.dtype == torch.float32 and gt_masks.dtype in (torch.uint8, torch.bool)
            C.call("myolo_detect_mask_targets", A["proposals"], gt_ids, gt_boxes, gt_masks, B, R, M, gt_masks.shape[3],
                   cfg["S"], mh, mw,
                   A["rois"], self.target_ids, A["target_masks"], self.n_pos, self.roi_src, self.roi_gt, st)
            self._phase("yolo loss, mask targets (+ yolo-branch backward issued on its stream)")
            masks = self.mask_head(A["rois"], training=learning_phase)
            C.call("myolo_mask_loss", A["masks"], A["target_masks"], self.target_ids, self.n_roi, mh, mw, NC,
                   float(lw.get("myolo_mask_loss", 1.0)), self.loss_mask, A["dlogit"], self.ws_loss, st)
            self._phase("ROIAlign + mask head forward + mask loss")
            out.update(output_rois=A["rois"], myolo_mask=masks, mask_loss=self.loss_mask[0],
                       target_class_ids=self.target_ids, target_mask=A["target_masks"])
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, on_tail_ready=None):
        """Gradients of LOSS_WEIGHTS-weighted (yolo_sum_loss + mask_loss) w.r.t. every trainable
        variable, into the flat gradient buffer.  `on_tail_ready()` is invoked once the
        feature_map + mask-head slice [tail_off:] is final (hook for the overlapped all-reduce).

        Streams: the backward of the YOLO branch (conv_23, blocks 14..7) depends on the yolo loss only, so it runs on its
        own stream NEXT TO the mask-head backward (eight milliseconds of persistent tensor-core kernels, one CTA per SM,
        which leave the issue slots of every SM mostly idle): ~40 latency-bound launches leave the critical path.  The two
        chains meet at block 6, where d(C4) of the feature_map branch is added.  The filter-gradient kernels of blocks
        6..1 run on the side stream next to the data-gradient chain, as before."""
        A, B = self.A, self.B
        main, side, ys = torch.cuda.current_stream(), self._side, self._ystream
        overlap = self.with_mask and ys is not None
        blocks = list(reversed(BACKBONE_BLOCKS + YOLO_BLOCKS))
        n_y = len(YOLO_BLOCKS)
        if overlap:
            if not getattr(self, "_y_started", False):
                self._start_yolo_branch_backward()
            self._y_started = False
            e1 = self._ev("y_done")
            self._backward_mask()
            if on_tail_ready is not None and not self._w_used:
                C.record_py(on_tail_ready)
            C.record_py(lambda: main.wait_event(e1))
            self._phase("join of the yolo-branch backward")
            self._backward_blocks(blocks[n_y:], main, side, self.ws)
        else:
            C.record_py(self.grads.zero_)
            if self.with_mask:
                self._backward_mask()
            if on_tail_ready is not None and not self._w_used:
                C.record_py(on_tail_ready)
            self._backward_conv23(main, side, self.ws)
            self._backward_blocks(blocks, main, side, self.ws)
        S = self.cfg["S"]
        H0 = S // 2
        st = main.cuda_stream
        self._bn_bwd("conv1_bn", self._v(A["y0"]), C.view(self.gx, B, H0, H0, 32), C.ACT_RELU6, True)
        C.call("myolo_conv1_wgrad", self._image, self.gx, self.g["conv1/kernel"], B, S, 32, st)
        self._phase("backbone blocks 6..1 + conv1 backward (data-gradient chain)")
        if side is not None and "f_done" in self._evs:
            e = self._evs["f_done"]
            C.record_py(lambda: main.wait_event(e))     # every gradient is in the flat buffer once main passes this point
        if self._w_used:                                # ... and the mask head's filter gradients
            W, ew = self._wstream, self._ev("w_done")
            C.record_py(lambda: (ew.record(W), main.wait_event(ew)))
            if on_tail_ready is not None:
                C.record_py(on_tail_ready)
        self._phase("join of the filter-gradient streams")

    def _start_yolo_branch_backward(self):
        """Zeroes the flat gradient buffer and issues conv_23's and blocks 14..7's backward on the Y stream, behind
        everything issued on the current stream so far (the yolo loss gradient)."""
        main, ys = torch.cuda.current_stream(), self._ystream
        blocks = list(reversed(BACKBONE_BLOCKS + YOLO_BLOCKS))
        C.record_py(self.grads.zero_)
        e0, e1 = self._ev("y_fork"), self._ev("y_done")
        C.record_py(lambda: (e0.record(main), ys.wait_event(e0)))          # after grads.zero_()
        # The chain only advances in the gaps between the persistent mask-head kernels (its GEMMs need the shared memory those
        # hold): with its filter gradients on a stream of their own (MYOLO_Y_SIDE=1) a gap starts two of its kernels instead
        # of one, and the chain needs half as many gaps.
        yside = self._yside
        self._backward_conv23(ys, yside, self.ws_y)
        self._backward_blocks(blocks[:len(YOLO_BLOCKS)], ys, yside, self.ws_y)
        if yside is not None:
            ef = self._evs["f_done"]                      # the chain's last filter-gradient kernel
            C.record_py(lambda: ys.wait_event(ef))
        C.record_py(lambda: e1.record(ys))
        self._y_started = True

    def _ev(self, name):        # events are created once and re-recorded every step
        e = self._evs.get(name)
        if e is None:
            e = self._evs[name] = torch.cuda.Event()
        return e

    def _backward_conv23(self, main, side, ws):
        """conv_23 (model.py:271): bias / kernel gradients and d(ap14) into gx, on stream `main` (kernel gradient on `side`)."""
        A, B = self.A, self.B
        G, ny = self.cfg["G"], self.ny
        st = main.cuda_stream
        sst = side.cuda_stream if side is not None else st
        dy = A["dyolo"]
        C.call("myolo_colsum", self._v(dy), self.g["conv_23/bias"], ws, st)
        if side is not None:
            e = self._ev("f23")
            C.record_py(lambda: (e.record(main), side.wait_event(e)))
        ap14 = A["ap14"][0] if self.x3 else A["ap14"]
        C.call("myolo_pwconv_wgrad", ap14, dy, self.g["conv_23/kernel"], B * G * G, 1024, ny, sst)
        if self.tc:     # data gradient on tcgen05: dy padded to ny_pad channels against the zero-padded [1024][ny_pad] kernel
            C.call("myolo_copy_cols", dy, ny, A["dyolo_pad"], self.ny_pad, B * G * G, ny, st)
            C.call("myolo_gemm_taps", A["dyolo_pad"], self.ny_pad, self.w23_d, self.gx, 1024, B * G * G, 1024, self.ny_pad, 1,
                   None, None, None, None, C.ACT_NONE, 0, 0, 0, st)
        else:
            C.call("myolo_pwconv_dgrad", dy, self.p["conv_23/kernel"], self.gx, B * G * G, 1024, ny, st)

    def _backward_blocks(self, blocks, main, side, ws):
        """Backward of depthwise-separable blocks (given last to first); d(block output) arrives in gx and d(block input)
        leaves in gx.  Data-gradient chain on `main`; with a `side` stream the two filter-gradient kernels of a block run
        there (they are 20-70 us kernels that fill a fraction of the SMs)."""
        A, B = self.A, self.B
        relu6 = C.ACT_RELU6
        gx, gy = self.gx, self.gy
        st = main.cuda_stream
        sst = side.cuda_stream if side is not None else st

        def fork(name):     # the side stream may start once everything issued on main so far is done
            if side is not None:
                e = self._ev(name)
                C.record_py(lambda: (e.record(main), side.wait_event(e)))

        def mark(name):     # remember the side stream's position
            if side is not None:
                e = self._ev(name)
                C.record_py(lambda: e.record(side))

        def join(name):     # main waits for that position
            if side is not None and name in self._evs:
                e = self._evs[name]
                C.record_py(lambda: main.wait_event(e))

        first = True
        for k, ci, co, s in blocks:
            Hi, Ho, _, _, _ = self.geo[k]
            npix = B * Ho * Ho
            if k == 6 and self.with_mask:
                # join the feature_map branch: d(C4) += dgrad of the 3x3 conv (padded-flat -> dense)
                C.call("myolo_view_copy", self.dc4.view(), C.view(gx, B, Ho, Ho, co), 1, st)
            d_ap = C.view(gx, B, Ho, Ho, co)
            self._bn_bwd(f"conv_pw_{k}_bn", self._v(A[f"yp{k}"]), d_ap, relu6, True, st, ws)
            ad_hi = A[f"ad{k}"][0] if self.x3 else A[f"ad{k}"]
            fork("fw")
            C.call("myolo_pwconv_wgrad", ad_hi, gx, self.g[f"conv_pw_{k}/kernel"], npix, ci, co, sst)      # reads gx
            mark("w_done")
            if not first:
                join("f_done")          # the previous block's bwd_filter has finished reading gy
            first = False
            C.call("myolo_pwconv_dgrad", gx, self.p[f"conv_pw_{k}/kernel"], gy, npix, ci, co, st)
            d_ad = C.view(gy, B, Ho, Ho, ci)
            self._bn_bwd(f"conv_dw_{k}_bn", self._v(A[f"yd{k}"]), d_ad, relu6, True, st, ws)
            ib = self._deferred.get(k - 1)
            if ib is not None:          # the block's input exists only as the producer's pre-BN output: BN + ReLU6 on load
                xin = self._v(A["y0"] if k == 1 else A[f"yp{k - 1}"])
            elif k == 1:
                xin = self._v(A["a0"])
            elif k == 7 and self.with_mask and not self.x3:
                xin = self.c4.view()
            else:
                xin = self._v(A[f"ap{k - 1}"])
            fork("ff")
            if ib is not None:
                C.call("myolo_dwconv3x3_bwd_filter_bn", xin, gy, self.g[f"conv_dw_{k}/depthwise_kernel"], s, ib.mean, ib.var,
                       ib.gamma, ib.beta, BN_EPS, relu6, sst)                                           # reads gy
            else:
                C.call("myolo_dwconv3x3_bwd_filter", xin, gy, self.g[f"conv_dw_{k}/depthwise_kernel"], s, sst)  # reads gy
            mark("f_done")
            join("w_done")              # the pointwise wgrad has finished reading gx
            C.call("myolo_dwconv3x3_bwd_data", gy, self.p[f"conv_dw_{k}/depthwise_kernel"], gx, B, Hi, Hi, ci, s, st)

    def _backward_mask_h16(self):
        """Backward of the mask head on tcgen05 kind::f16.  The gradient tensors are half, multiplied by the
        power-of-two loss scale gs[0] (device scalar, from max|dlogit|); every parameter gradient and the two fp32
        hand-overs -- d(a1) into the batch-statistics BN backward, d(x0) into the ROIAlign backward -- are un-scaled
        by gs[1] where they are produced."""
        A, st, n = self.A, self._st(), self.n_roi
        P_ = self.cfg["POOL"]
        pfw, pfb = P_ + 1, (P_ + 1) * (P_ + 1)
        M = self.mah[0].M
        if self._mask_fused[1]:
            raise C.MyoloError("backward needs the forward pass of the same step in the learning phase (batch-statistics bn1)")
        gs, ugs = self.gs, self.gs[1:]
        sh3 = self._shift_cache[("h16", P_)]
        shn = self._neg_shifts
        # Filter gradients feed nothing but Adam: with a W stream they are issued there BEHIND the mask head's data-gradient
        # chain (one event after its last GEMM), so that they run next to ROIAlign's backward, the feature_map backward and
        # blocks 6..1 of the backbone instead of in front of them; Engine.backward joins the stream before the optimizer.
        # Every layer then keeps its own gradient tensor (mgh[0..3]) instead of two ping-pong buffers.
        main, W = torch.cuda.current_stream(), self._wstream
        wst = W.cuda_stream if W is not None else st
        self._w_used = W is not None

        deferred = []
        n_seen = [0]

        def wgrad(*args):       # args end with the stream handle; issue order: deconv, conv4, conv3, conv2, conv1
            k = n_seen[0]
            n_seen[0] += 1
            if W is not None and k >= 5 - self._w_defer:
                deferred.append(args)
            elif W is not None:     # not deferred: inline on the main stream, full width
                C.call("myolo_gemm_taps_wgrad_h", *(args[:-1] + (st,)))
            else:
                C.call("myolo_gemm_taps_wgrad_h", *args)

        C.call("myolo_grad_scale", A["dlogit"], A["dlogit"].numel(), gs, st)
        C.call("myolo_mask_out_bwd_h", self.y4d.rows, self.p["myolo_mask_deconv/bias"], self.p["myolo_mask/kernel"],
               A["dlogit"], self.dy4h.rows, self.g["myolo_mask/kernel"], self.g["myolo_mask/bias"],
               self.g["myolo_mask_deconv/bias"], n, P_, P_, MASK_C, self.NC, gs, self.target_ids, self.dy4h_ids, st)
        if not self.sparse_backward:
            wgrad(self.mah[4].rows, MASK_C, self.dy4h.rows, 4 * MASK_C, self.g["myolo_mask_deconv/kernel"],
                  M, 4 * MASK_C, MASK_C, 1, None, 1, ugs, wst)
        G = self.mgh if W is not None else [self.mgh[0], self.mgh[1], self.mgh[0], self.mgh[1]]     # d(pre-BN) of conv4..conv1
        if self.sparse_backward:
            # everything between the loss and d(a1) on the positive rois only; a host action (it reads the number of
            # positives) that is recorded as such and issues its own launches at every replay
            g1 = self.mgh[1]
            C.record_py(self._sparse_mask_middle)
        else:
            g1 = self._dense_mask_middle(G, wgrad, M, pfw, pfb, sh3, shn, ugs, st)
        b = self.bn["myolo_mask_bn1"]
        if self._bn1_half:      # the reduction pass ran in the epilogue of conv2's data-gradient GEMM (_conv2_dgrad)
            C.call("myolo_bn_bwd_batch_fix_hh", self.z1h.view(), g1.view(), b.mean, b.var, b.gamma, BN_EPS, b.dgamma, b.dbeta,
                   self.ws, ugs, st)
        else:
            C.call("myolo_bn_bwd_hh", self.my[1].view(), g1.view(), g1.view(), b.mean, b.var, b.gamma, b.beta, BN_EPS,
                   C.ACT_RELU, 1, b.dgamma, b.dbeta, self.ws, ugs, st)
        name = "myolo_mask_conv1/kernel"
        wgrad(self.mah[0].rows, MASK_C, g1.rows, MASK_C, self.g[name], M, MASK_C, MASK_C, 9, sh3, 0, ugs, wst)
        if self._dx0_half:
            C.call("myolo_gemm_taps_h", g1.rows, MASK_C, self.wth_d[name], None, 0, self.mg[1].rows, MASK_C, M, MASK_C, MASK_C, 9,
                   shn, None, None, None, C.ACT_NONE, pfw, pfb, None, st)
        else:
            C.call("myolo_gemm_taps_h", g1.rows, MASK_C, self.wth_d[name], self.mg[1].rows, MASK_C, None, 0, M, MASK_C, MASK_C, 9,
                   shn, None, None, None, C.ACT_NONE, pfw, pfb, ugs, st)
        if W is not None:       # the data-gradient chain of the mask head is issued: the filter gradients start behind it
            e = self._ev("w_start")
            C.record_py(lambda: (e.record(main), W.wait_event(e)))
            if self._w_sms < 148:
                C.call("myolo_set_wgrad_sms", self._w_sms)
            for args in deferred:
                C.call("myolo_gemm_taps_wgrad_h", *args)
            if self._w_sms < 148:
                C.call("myolo_set_wgrad_sms", 148)
        return self.mg[1]

    def _dense_mask_middle(self, G, wgrad, M, pfw, pfb, sh3, shn, ugs, st):
        """deconv data gradient + BN4 backward, conv4..conv2 filter / data gradients with the fused BN backward, conv2's data
        gradient: from dy4h to d(a1) (scaled half, returned) over ALL rois."""
        def dgrad_bn(src_rows, lda, name, dst_rows, K, ntaps, shifts, layer):
            b = self.bn[f"myolo_mask_bn{layer}"]
            C.call("myolo_gemm_taps_bnbwd_h", src_rows, lda, self.wth_d[name], None, dst_rows, MASK_C, M, MASK_C, K, ntaps, shifts,
                   pfw, pfb, self.mah[layer].rows, b.gamma, b.beta, b.mvar, BN_EPS, C.ACT_RELU, b.dgamma, b.dbeta,
                   self.g[f"myolo_mask_conv{layer}/bias"], self.ws, ugs, st)

        dgrad_bn(self.dy4h.rows, 4 * MASK_C, "myolo_mask_deconv/kernel", G[0].rows, 4 * MASK_C, 1, None, 4)
        for i in (4, 3, 2):
            name = f"myolo_mask_conv{i}/kernel"
            gi = G[4 - i]
            wgrad(self.mah[i - 1].rows, MASK_C, gi.rows, MASK_C, self.g[name], M, MASK_C, MASK_C, 9, sh3, 0, ugs,
                  self._wstream.cuda_stream if self._w_used else st)
            if i > 2:
                dgrad_bn(gi.rows, MASK_C, name, G[4 - i + 1].rows, MASK_C, 9, shn, i - 1)
        # d(a1) as scaled half
        g2, g1 = G[2], G[3]
        self._conv2_dgrad(g2.rows, g1.rows, self.mah[1].rows, M, pfw, pfb, shn, st)
        return g1

    def _conv2_dgrad(self, g2_rows, g1_rows, a1_rows, M, pfw, pfb, shn, st):
        """Data gradient of myolo_mask_conv2 = d(a1).  With the half bn1 path its epilogue already applies relu'(a1) and
        gamma * rs and leaves bn1's two column sums in the BN workspace for myolo_bn_bwd_batch_fix_hh."""
        if self._bn1_half:
            b = self.bn["myolo_mask_bn1"]
            C.call("myolo_gemm_taps_bnbwd_sums_h", g2_rows, MASK_C, self.wth_d["myolo_mask_conv2/kernel"], g1_rows, MASK_C, M,
                   MASK_C, MASK_C, 9, shn, pfw, pfb, a1_rows, b.gamma, b.beta, b.var, BN_EPS, C.ACT_RELU, self.ws, st)
        else:
            C.call("myolo_gemm_taps_h", g2_rows, MASK_C, self.wth_d["myolo_mask_conv2/kernel"], None, 0, g1_rows, MASK_C, M,
                   MASK_C, MASK_C, 9, shn, None, None, None, C.ACT_NONE, pfw, pfb, None, st)

    def _sparse_mask_middle(self):
        """Exact sparse form of _dense_mask_middle.  The mask loss touches the rois with a target class only
        (myolo_mask_loss_graph gathers them, model.py:718-754), bn2..bn4 use fixed statistics and every roi is its own
        "image" for the convolutions, so d(a4)..d(a1) and every summand of the filter / BN gradients of conv2..conv4 and of
        the deconvolution are identically zero for all other rois.  Their padded-flat tiles are gathered into compact
        tensors, the SAME kernels run on P tiles instead of n_roi, and d(a1) is scattered back into a zeroed full tensor
        for the batch-statistics backward of bn1, which couples all rois.  Runs as a host action inside the recorded step
        (it needs the number of positives on the host: one 4*B-byte read).  More positives than the compact tensors hold:
        the dense form runs for that step."""
        with C.no_record():
            st = self._st()
            B, R, n = self.B, self.R, self.n_roi
            P_ = self.cfg["POOL"]
            pfw, pfb = P_ + 1, (P_ + 1) * (P_ + 1)
            M = self.mah[0].M
            ugs = self.gs[1:]
            sh3, shn = self._shift_cache[("h16", P_)], self._neg_shifts
            npos = self.n_pos.cpu().tolist()             # the positives are the first n_pos[b] rois of image b (a7)
            idx = [b * R + j for b in range(B) for j in range(npos[b])]
            P = len(idx)
            self.sparse_stats["steps"] += 1
            self.sparse_stats["rois"] += P
            g1 = self.mgh[1]
            if P > self.pcap:
                self.sparse_stats["dense_fallback"] += 1
                G = [self.mgh[0], self.mgh[1], self.mgh[0], self.mgh[1]]
                call = lambda *a: C.call("myolo_gemm_taps_wgrad_h", *a)                       # noqa: E731
                call(self.mah[4].rows, MASK_C, self.dy4h.rows, 4 * MASK_C, self.g["myolo_mask_deconv/kernel"], M, 4 * MASK_C,
                     MASK_C, 1, None, 1, ugs, st)
                w_used, self._w_used = self._w_used, False
                self._dense_mask_middle(G, call, M, pfw, pfb, sh3, shn, ugs, st)
                self._w_used = w_used
                return
            C.record_py(g1.storage.zero_)                 # (not recording here: simply runs) d(a1) of the other rois
            if P == 0:
                self.sparse_stats["no_positives"] += 1
                return
            self.sparse_stats["sparse"] += 1
            self.sp_list_host[:P] = torch.tensor(idx, dtype=torch.int32)
            self.sp_list[:P].copy_(self.sp_list_host[:P], non_blocking=True)
            tile = pfb * MASK_C * 2                        # bytes of one roi's padded-flat tile, half
            for i in (1, 2, 3, 4):
                C.call("myolo_copy_tiles", self.mah[i].rows, self.sp_a[i].rows, self.sp_list, P, tile, 0, st)
            C.call("myolo_copy_tiles", self.dy4h.rows, self.sp_dy4.rows, self.sp_list, P, 4 * tile, 0, st)
            Mc = P * pfb
            a, dy4, (ga, gb) = self.sp_a, self.sp_dy4, self.sp_g
            C.call("myolo_gemm_taps_wgrad_h", a[4].rows, MASK_C, dy4.rows, 4 * MASK_C, self.g["myolo_mask_deconv/kernel"], Mc,
                   4 * MASK_C, MASK_C, 1, None, 1, ugs, st)

            def dgrad_bn(src_rows, lda, name, dst_rows, K, ntaps, shifts, layer):
                b = self.bn[f"myolo_mask_bn{layer}"]
                C.call("myolo_gemm_taps_bnbwd_h", src_rows, lda, self.wth_d[name], None, dst_rows, MASK_C, Mc, MASK_C, K, ntaps,
                       shifts, pfw, pfb, a[layer].rows, b.gamma, b.beta, b.mvar, BN_EPS, C.ACT_RELU, b.dgamma, b.dbeta,
                       self.g[f"myolo_mask_conv{layer}/bias"], self.ws, ugs, st)

            dgrad_bn(dy4.rows, 4 * MASK_C, "myolo_mask_deconv/kernel", ga.rows, 4 * MASK_C, 1, None, 4)
            for i in (4, 3, 2):
                name = f"myolo_mask_conv{i}/kernel"
                C.call("myolo_gemm_taps_wgrad_h", a[i - 1].rows, MASK_C, ga.rows, MASK_C, self.g[name], Mc, MASK_C, MASK_C, 9,
                       sh3, 0, ugs, st)
                if i > 2:
                    dgrad_bn(ga.rows, MASK_C, name, gb.rows, MASK_C, 9, shn, i - 1)
                    ga, gb = gb, ga
            self._conv2_dgrad(ga.rows, gb.rows, a[1].rows, Mc, pfw, pfb, shn, st)
            C.call("myolo_copy_tiles", gb.rows, g1.rows, self.sp_list, P, tile, 1, st)

    def _backward_mask(self):
        A, st, n = self.A, self._st(), self.n_roi
        P_, B, F_ = self.cfg["POOL"], self.B, self.F
        pfw, pfb = P_ + 1, (P_ + 1) * (P_ + 1)
        if self.h16:
            g0 = self._backward_mask_h16()
            self._phase("mask head backward (loss .. d(x0))")
            self._backward_feature_map(g0)
            self._phase("ROIAlign backward + feature_map backward")
            return
        a4 = self.ma[4]
        C.call("myolo_mask_out_bwd", self.y4d.rows, self.p["myolo_mask_deconv/bias"], self.p["myolo_mask/kernel"],
               A["dlogit"], self.dy4d.rows, self.g["myolo_mask/kernel"], self.g["myolo_mask/bias"],
               self.g["myolo_mask_deconv/bias"], n, P_, P_, MASK_C, self.NC, st)
        # deconv: dKd[(a,b,co)][ci] = sum_p dy4[p][(a,b,co)] a4[p][ci]  (transposed wgrad output = Keras layout)
        C.call("myolo_gemm_taps_wgrad", a4.rows, MASK_C, self.dy4d.rows, 4 * MASK_C, self.g["myolo_mask_deconv/kernel"],
               a4.M, 4 * MASK_C, MASK_C, 1, None, 1, st)
        g0, g1 = self.mg
        M = a4.M
        shn = self._neg_shifts

        def fusable(i):      # BN_i backward can ride in the epilogue of the GEMM that produces d(a_i)
            return i >= 2 and self._mask_fused[i] and self.tc and M >= 4096 and MASK_C == 256

        def dgrad(src_rows, lda, bt, dst_rows, K, ntaps, shifts, bn_layer):
            """d(a_bn_layer) = GEMM(src); with bn_layer set, its fixed-statistics BN + ReLU backward is fused into
            the epilogue (dst = d(pre-BN), dgamma / dbeta / conv-bias gradients reduced on the fly)."""
            if bn_layer:
                b = self.bn[f"myolo_mask_bn{bn_layer}"]
                C.call("myolo_gemm_taps_bnbwd", src_rows, lda, bt, dst_rows, MASK_C, M, MASK_C, K, ntaps, shifts, pfw, pfb,
                       self.ma[bn_layer].rows, b.gamma, b.beta, b.mvar, BN_EPS, C.ACT_RELU | self.rnd, b.dgamma, b.dbeta,
                       self.g[f"myolo_mask_conv{bn_layer}/bias"], self.ws, st)
            else:
                C.call("myolo_gemm_taps", src_rows, lda, bt, dst_rows, MASK_C, M, MASK_C, K, ntaps, shifts, None, None, None,
                       C.ACT_NONE, pfw, pfb, 0, st)

        bn_done = fusable(4)
        dgrad(self.dy4d.rows, 4 * MASK_C, self.wt["myolo_mask_deconv/kernel"], g0.rows, 4 * MASK_C, 1, None, 4 if bn_done else 0)
        for i in (4, 3, 2, 1):
            if not bn_done:
                if self._mask_fused[i]:
                    b = self.bn[f"myolo_mask_bn{i}"]
                    C.call("myolo_bn_act_bwd_from_output", self.ma[i].view(), g0.view(), g0.view(), b.gamma, b.beta, b.mvar,
                           BN_EPS, C.ACT_RELU | self.rnd, b.dgamma, b.dbeta, self.g[f"myolo_mask_conv{i}/bias"], self.ws, st)
                else:
                    self._bn_bwd(f"myolo_mask_bn{i}", self.my[i].view(), g0.view(), C.ACT_RELU | self.rnd, i == 1)
                    if i != 1:      # conv1's bias feeds a batch-statistics BN: its gradient is identically zero
                        C.call("myolo_colsum", g0.view(), self.g[f"myolo_mask_conv{i}/bias"], self.ws, st)
            C.call("myolo_conv3x3_wgrad", self.ma[i - 1].rows, g0.rows, self.g[f"myolo_mask_conv{i}/kernel"], n, P_, P_,
                   MASK_C, MASK_C, st)
            bn_done = i >= 2 and fusable(i - 1)
            dgrad(g0.rows, MASK_C, self.p[f"myolo_mask_conv{i}/kernel"], g1.rows, MASK_C, 9, shn, (i - 1) if bn_done else 0)
            g0, g1 = g1, g0
        self._backward_feature_map(g0)

    def _backward_feature_map(self, g0):
        """d(x0) -> CropAndResizeGradImage -> feature_map conv backward (model.py:848, 385-387)."""
        A, st, n = self.A, self._st(), self.n_roi
        P_, B, F_ = self.cfg["POOL"], self.B, self.F
        # CropAndResizeGradImage into the feature-map gradient
        C.record_py(self.dfeat.storage.zero_)
        if g0.rows.dtype == torch.float16:
            C.call("myolo_roialign_bwd_h", g0.view(), A["rois"], n, self.R, P_, self.dfeat.view(), self.gs[1:], st)
        else:
            C.call("myolo_roialign_bwd", g0.view(), A["rois"], n, self.R, P_, self.dfeat.view(), st)
        C.call("myolo_colsum", self.dfeat.view(), self.g["feature_map/bias"], self.ws, st)
        W = self._wstream if getattr(self, "_w_used", False) else None
        if W is not None:
            main, e = torch.cuda.current_stream(), self._ev("w_fm")
            C.record_py(lambda: (e.record(main), W.wait_event(e)))
        C.call("myolo_conv3x3_wgrad", self.c4.rows, self.dfeat.rows, self.g["feature_map/kernel"], B, F_, F_, 512, MASK_C,
               W.cuda_stream if W is not None else st)
        C.call("myolo_conv3x3_dgrad", self.dfeat.rows, self.p["feature_map/kernel"], self.dc4.rows, B, F_, F_, 512, MASK_C, st)

    # ------------------------------------------------------------------ optimizer
    def apply_updates(self, lr: float = 1e-3, grad_scale: float = 1.0):
        """Keras Adam (model.py:1071-1075: lr, beta_1 0.9, beta_2 0.999, epsilon 1e-8, no decay) over
        the flat buffers, then the Keras BN moving-average update (momentum 0.99, TF zero-debias)."""
        st = self._st()
        self.t += 1
        b1, b2 = 0.9, 0.999
        lr_t = lr * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        if self._frozen:                                # set_trainable(): frozen variables are not in the optimizer
            C.call("myolo_adam_step_masked", self.params, self.grads, self.adam_m, self.adam_v, self.trainable_mask,
                   self.n_flat, lr_t, b1, b2, 1e-8, grad_scale, st)
            # Keras drops the moving-average updates of a BatchNormalization layer whose `trainable` is False
            fz = self._frozen_names
            self._bn_touched = [(b, n) for b, n in self._bn_touched if (b.name + "/gamma") not in fz]
        else:
            C.call("myolo_adam_step", self.params, self.grads, self.adam_m, self.adam_v, self.n_flat, lr_t, b1, b2, 1e-8,
                   grad_scale, st)
        if self._bn_touched:
            # one launch for every (layer, statistic) that shares a step count (the zero-debias factor 1 - momentum^step is
            # a kernel argument); layers that were frozen for a while lag behind and form their own group.  The record
            # tables are rebuilt only when the set of batch-statistics layers changes (it never does within a run).
            base = min(b.step for b, _ in self._bn_touched)
            key = tuple((id(b), b.step - base) for b, _ in self._bn_touched)     # membership and grouping
            if self._moving_key != key:
                import struct
                groups = {}
                for b, npix in self._bn_touched:
                    groups.setdefault(b.step, []).append((b, npix))
                self._moving_tables = []
                for _, members in sorted(groups.items()):
                    rec = b""
                    for b, npix in members:
                        n = float(npix)
                        corr = (n / max(n - 1.0, 1.0)) * (n / (n - (1.0 + BN_EPS)))
                        rec += struct.pack("<QQQif", b.mean.data_ptr(), b.bmean.data_ptr(), b.mmean.data_ptr(), b.c, 1.0)
                        rec += struct.pack("<QQQif", b.var.data_ptr(), b.bvar.data_ptr(), b.mvar.data_ptr(), b.c, corr)
                    self._moving_tables.append((torch.frombuffer(bytearray(rec), dtype=torch.uint8).to(self.dev),
                                                2 * len(members), members[0][0]))
                self._moving_key = key
            for b, _ in self._bn_touched:
                b.step += 1
            for table, n, b0 in self._moving_tables:
                C.call("myolo_bn_moving_update_batch", table, n, BN_MOMENTUM, b0.step, st)
        self._bn_touched = []
        self.version += 1
        self.refresh_weights()
        self._phase("Adam + BN moving update + weight staging")

    def train_step(self, inputs, lr: float = 1e-3, allreduce=None):
        """One fit step.  `allreduce(flat_grads, lo, hi)` (optional) sums gradient slices across
        replicas; it is called for the tail bucket as soon as it is ready and for the head at the end.

        The forward + backward part is a fixed sequence of ~290 C-ABI launches with fixed arguments (only the input
        pointers change), so it is recorded on the first step and replayed afterwards as raw ctypes calls
        (MYOLO_REPLAY=0 disables this); the optimizer part carries per-step scalars and stays dynamic."""
        C.set_precision(C.PREC_TF32 if self.tc else C.PREC_FP32)    # process-global dispatch switch: another engine may have moved it
        if self._replay_on:
            warm = 1 if (self.seen + 1) < self.cfg.get("WARM_UP_BATCHES", 0) else 0
            key = (tuple((tuple(t.shape), t.dtype) for t in inputs), warm, id(allreduce),
                   torch.cuda.current_stream().cuda_stream)
            if self._plan is not None and self._plan["key"] == key:
                out = self._replay_step(inputs)
            else:
                out = self._record_step(inputs, allreduce, key)
        else:
            out = self.forward_training(inputs, start_backward=True)
            if allreduce is None:
                self.backward()
            else:
                self.backward(on_tail_ready=lambda: allreduce(self.grads, self.tail_off, self.n_flat))
        scale = allreduce(self.grads, 0, self.tail_off) if allreduce is not None else 1.0
        self.apply_updates(lr, scale if scale is not None else 1.0)
        return out

    def _record_step(self, inputs, allreduce, key):
        self._plan = None
        C.start_recording()
        try:
            out = self.forward_training(inputs, start_backward=True)
            if allreduce is None:
                self.backward()
            else:
                self.backward(on_tail_ready=lambda: allreduce(self.grads, self.tail_off, self.n_flat))
        finally:
            entries = C.stop_recording()
        # replay patches: (converted-argument list, argument position, input index) for every argument that WAS one of the
        # input tensors when the call was issued (identity of the Python object at that position, not its integer value)
        ids = {id(t): k for k, t in enumerate(inputs)}
        assert len(ids) == len(inputs), "input tensors must not alias each other"
        patches = [(e[1], ai, ids[id(a)]) for e in entries if e[0] is not None
                   for ai, a in enumerate(e[2]) if id(a) in ids]
        self._plan = dict(key=key, entries=entries, patches=patches, out=out, bn_touched=list(self._bn_touched))
        return out

    def _replay_step(self, inputs):
        plan = self._plan
        for conv, ai, k in plan["patches"]:
            conv[ai] = inputs[k].data_ptr()
        self._image = inputs[0]
        self._bn_touched = list(plan["bn_touched"])
        self.seen += 1
        C.replay(plan["entries"])
        return plan["out"]
