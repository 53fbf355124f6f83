#!/usr/bin/env python
"""Mask-YOLO benchmark: images/sec of one full training step (forward, both losses, backward, Adam,
BN moving update).  Prints ONE JSON line (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (sm_100a engine), BASELINE.json configs[1]:
                                                                 Shapes 224x224, per-GPU batch 32 (weak scaling, configs[3])
  python bench.py --config c3|c5 ...                             configs[2]: rice-like 416x416 batch 16, NC=2, R=845 dense ROIs
                                                                 configs[4]: COCO-shape 640x640 batch 8/GPU, NC=81, R=2000
  python bench.py --impl reference ...                           the reference's algorithm on the host CPU, full batch per step
                                                                 (oracle port: the Keras/TF code cannot run here)

Our arm times the headline precision (`h16`) and, in the same invocation, the fp32-class mode (`tf32x3`) -> `fp32_class`.
`value` = device-timed steps on batches resident in HBM; `e2e` = the public fit loop (MaskYOLO.fit_batches, the body of
MaskYOLO.train) fed numpy batches as BatchGenerator yields them: conversion into pinned memory, host->device copy and the
device->host read of the losses of EVERY step inside the timed region; `e2e_train_on_batch` = the same through the strictly
synchronous keras_model.train_on_batch(numpy batch) call.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "mask-yolo_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

METRIC = "images/sec fwd+bwd @224x224 Shapes"


def metric_name(args):
    return METRIC if args.config == "c2" and args.size == 224 else f"images/sec fwd+bwd @{args.size}x{args.size} ({args.config})"
WORKLOADS = {
    # name: (image side, per-GPU batch, description)
    "c2": (224, 32, "Shapes 224x224 batch 32/GPU, MobileNet+YOLO+ROIAlign+mask fwd+bwd+Adam"),
    "c3": (416, 16, "synthetic rice-like 416x416 batch 16/GPU (2 classes, 5 anchors, 845 dense ROIs/image), fwd+bwd+Adam"),
    "c5": (640, 8, "synthetic COCO-shape 640x640 batch 8/GPU (81 classes, 5 anchors, 2000 ROIs/image), fwd+bwd+Adam"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--precision", default=os.environ.get("MYOLO_PRECISION", "h16"))
    ap.add_argument("--no-fp32-class", action="store_true", help="skip the second (tf32x3) timing")
    ap.add_argument("--no-sparse", action="store_true", help="skip the secondary exact-sparse-backward timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    S, B, _ = WORKLOADS[a.config]
    a.size = a.size or S
    a.batch = a.batch or B
    return a


# ------------------------------------------------------------------------------------------------ workloads
def bench_config(batch, size, name="c2"):
    """Config instance of a workload (the reference's Config schema)."""
    from myolo.config import Config
    from myolo.shapes import ShapesConfig
    if name == "c2":
        class BenchConfig(ShapesConfig):
            BATCH_SIZE = batch
            IMAGE_SHAPE = [size, size, 3]
            IMAGE_MIN_DIM = IMAGE_MAX_DIM = size
            GRID_H = GRID_W = size // 32
        return BenchConfig()
    nc = 2 if name == "c3" else 81              # example/rice/rice_dataset.py:60-82 (background + rice); COCO 80 + background

    class BigConfig(Config):                     # base Config: 5 anchors, TRUE_BOX_BUFFER 10, MAX_GT_INSTANCES 10
        NAME = "rice_like" if name == "c3" else "coco_shape"
        BATCH_SIZE = batch
        NUM_CLASSES = nc
        IMAGE_SHAPE = [size, size, 3]
        IMAGE_MIN_DIM = IMAGE_MAX_DIM = size
        GRID_H = GRID_W = size // 32
        N_BOX = 5
        TRAIN_ROIS_PER_IMAGE = (size // 32) ** 2 * 5
        CLASS_WEIGHTS = np.ones(nc, dtype="float32")
    return BigConfig()


def init_weights(c, name):
    """Random 'trained-like' weights (no checkpoints offline).  For the dense-ROI workloads the detection head is damped
    so that the decoded proposals sit near the anchor priors, which is what the ground truth below is laid on."""
    from myolo.engine import init_params
    P = init_params(c["NB"], c["NC"], 0, "trained_like")
    if name != "c2":
        P["conv_23/kernel"] = P["conv_23/kernel"] * 0.05
        P["conv_23/bias"] = torch.zeros_like(P["conv_23/bias"])
    return P


def make_host_batches(cfg, name, n_batches, seed):
    """`n_batches` training batches as lists of numpy arrays in BatchGenerator's format and dtypes."""
    from myolo.config import resolve
    if name == "c2":
        from myolo.shapes import make_batches
        return make_batches(cfg, n_batches, seed=seed)
    from tests import helpers as Hh
    c = dict(resolve(cfg))
    B, S, G, NB = int(cfg.BATCH_SIZE), c["S"], c["G"], c["NB"]
    anc = np.asarray(c["ANCHORS"], np.float64).reshape(NB, 2)
    out = []
    for k in range(n_batches):
        rng = np.random.RandomState(seed + 17 * k)
        image = torch.from_numpy(rng.rand(B, S, S, 3).astype(np.float32))
        boxes = []
        for b in range(B):
            bl = []
            for m in range(c["MAXGT"] - (b % 3 if name == "c3" else 0)):          # c3: 8-10 instances per image
                if name == "c3":
                    # elongated grains on the anchor priors of a random cell (the two largest anchors): every prior of the
                    # neighbourhood overlaps them -> dense positives
                    a = anc[NB - 1 - (m % 2)] * (1.0 + 0.1 * (rng.rand(2) - 0.5)) * np.array([1.0, 0.8 + 0.4 * rng.rand()])
                    cx, cy = (rng.randint(2, G - 2, 2) + 0.5) / G
                    w, h = a / G
                else:
                    cx, cy = rng.rand(2) * 0.7 + 0.15
                    w, h = rng.rand(2) * 0.35 + 0.05
                bl.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
            boxes.append(bl)
        t = Hh.batch_from_boxes(c, B, image, boxes, seed + 17 * k + 1)
        t[1], t[2] = t[1].double(), t[2].double()                                   # BatchGenerator yields float64 here
        out.append([x.numpy() if x.dtype != torch.uint8 else x.numpy().astype(bool) for x in t])
    return out


def workload_config(args, c, world):
    """The `config` object of the JSON line: identical for both arms (same workload, metric and unit)."""
    return {"workload": WORKLOADS[args.config][2] if (args.size, args.batch) == WORKLOADS[args.config][:2]
            else f"{args.config} at {args.size}x{args.size} batch {args.batch}/GPU",
            "name": args.config, "image_size": args.size, "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "N_BOX": c["NB"], "NUM_CLASSES": c["NC"], "rois_per_image": c["R"], "parallelism": f"dp{world}",
            "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush",
            "weights": "random trained-like init (no checkpoints offline)"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.rows, self.proc = dev, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def oracle_inputs(batch_np, nimg):
    return [torch.from_numpy(np.ascontiguousarray(batch_np[0][:nimg])).float(), torch.from_numpy(np.asarray(batch_np[1][:nimg])).float(),
            torch.from_numpy(np.asarray(batch_np[2][:nimg])).float(), torch.from_numpy(np.asarray(batch_np[3][:nimg])),
            torch.from_numpy(np.asarray(batch_np[4][:nimg])).float(), torch.from_numpy(np.asarray(batch_np[5][:nimg])).bool()]


def oracle_step_fn(cfg, batch_np, nimg, name="c2"):
    """Closure running oracle.train_step (fwd+bwd+Adam+BN update) on the first `nimg` images; every call starts from the
    same initial weights.  Returns (closure, initial weights)."""
    from myolo.config import resolve
    from oracle import myolo_oracle as O
    from tests import helpers as Hh
    c = resolve(cfg)
    oc = Hh.oracle_cfg(c)
    P0 = init_weights(c, name)
    x = oracle_inputs(batch_np, nimg)

    def step():
        P = {k: v.clone() for k, v in P0.items()}
        return O.train_step(P, {}, x, oc, lr=1e-3)
    return step, P0


def flops_per_image(c):
    n_roi = c["R"]
    mask = 4 * 2 * 196 * 2304 * 256 * n_roi + 2 * 196 * 256 * 1024 * n_roi + 2 * 784 * 256 * c["NC"] * n_roi
    back = (1.263e9 + 1.85e9) * (c["S"] / 224.0) ** 2
    return 3.0 * (mask + back)


def run_reference(args):
    """Reference arm: the reference's algorithm (oracle port; Keras/TF 1.x is not installable here) on the host cores
    with every thread torch can use, on the FULL per-GPU batch of the workload each step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from myolo.config import resolve
    cfg = bench_config(args.batch, args.size, args.config)
    c = resolve(cfg)
    batch = make_host_batches(cfg, args.config, 1, seed=1234)[0]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nimg = args.batch
    step, _ = oracle_step_fn(cfg, batch, nimg, args.config)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = nimg * args.steps / dt
    sample = f"{args.steps} oracle.train_step calls (fwd+bwd+Adam, torch CPU fp32, {cores} threads) on the full batch of {nimg} " \
             f"images of the workload (NB={c['NB']}, NC={c['NC']}, R={c['R']})"
    print(json.dumps({"impl": "reference", "metric": metric_name(args), "value": v, "unit": "images/sec", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": workload_config(args, c, args.gpus),
                      "sample_images_per_step": nimg,
                      "note": "rank 0 only: one host runs the reference's CPU path on one per-GPU batch per step",
                      "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ our arm
def time_precision(args, precision, cfg, host_batches, world, rank, local, with_kernel_events, sparse=False):
    """Build the model in `precision`, time K device-resident steps (CUDA events) and the end-to-end loops."""
    import torch.distributed as dist
    from myolo import _cabi as C
    from myolo import ddp
    from myolo.model import MaskYOLO
    os.environ["MYOLO_SPARSE_BWD"] = "1" if sparse else "0"
    model = MaskYOLO("training", cfg, precision=precision, device=local)
    os.environ["MYOLO_SPARSE_BWD"] = "0"
    eng = model.engine
    c = eng.cfg
    eng.load_params(init_weights(c, args.config))
    if world > 1:
        ddp.attach(model)
    pool = len(host_batches)
    dev_batches = []
    for hb in host_batches:
        staged = model._stage(hb)
        torch.cuda.synchronize()
        dev_batches.append([t.clone() for t in staged])
    eng.inputs_ready = None
    torch.cuda.synchronize()
    B, K, W = args.batch, args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.kernel_events = []
    for i in range(W):
        eng.train_step(dev_batches[i % pool], 1e-3, model.allreduce)
    eng.kernel_events = [] if with_kernel_events else None
    barrier()
    launches0 = C.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        out = eng.train_step(dev_batches[i % pool], 1e-3, model.allreduce)
    ev1.record()
    barrier()
    launches = C.launch_count - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    kev = list(eng.kernel_events or [])
    eng.kernel_events = None
    npos = int(eng.n_pos.sum().item())
    res = {"precision": precision, "launches": launches, "kev": kev, "npos": npos,
           "loss": [float(out["yolo_sum_loss"]), float(out["mask_loss"])]}
    # ---- end to end: numpy batches (BatchGenerator's dtypes) -> pinned staging -> H2D -> step -> D2H losses, every step
    if not args.no_e2e:
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

        def timed(run):
            run(2)
            barrier()
            t0 = time.perf_counter()
            run(K)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return dt.item()

        def run_fit(n):
            for _ in model.fit_batches(host_batches[i % pool] for i in range(n)):
                pass

        def run_sync(n):
            for i in range(n):
                model.keras_model.train_on_batch(host_batches[i % pool])

        dt = timed(run_fit)
        res["e2e"] = {"value": B * world * K / dt, "unit": "images/sec", "h2d_bytes_per_step": int(model.last_h2d_bytes),
                      "d2h_bytes_per_step": int(model.last_d2h_bytes), "ms_per_step": 1e3 * dt / K,
                      "path": "MaskYOLO.fit_batches (the fit loop of MaskYOLO.train) over numpy batches: host conversion into pinned "
                              "staging + H2D + step + D2H of the losses, every step; losses of step k read after step k+1 is enqueued"}
        dt = timed(run_sync)
        res["e2e_train_on_batch"] = {"value": B * world * K / dt, "unit": "images/sec", "ms_per_step": 1e3 * dt / K,
                                     "path": "keras_model.train_on_batch(numpy batch), synchronous per step"}
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res["ms_total"] = ms.item()
    res["ms_per_step"] = ms.item() / K
    res["value"] = B * world * K / (ms.item() / 1e3)
    res["n_roi"] = eng.n_roi
    res["cfg"] = c
    res["sparse_stats"] = dict(eng.sparse_stats)
    res["phases"] = eng.phase_times() if eng._phases_on else None      # MYOLO_PHASES=1: main-stream phase boundaries of the last step
    del model, eng, dev_batches, out
    import gc
    gc.collect()                      # MaskYOLO <-> keras_model handle is a reference cycle: without this the engine's
    torch.cuda.empty_cache()          # buffers (80 GB in tf32x3 at config 5) outlive the function
    return res


def parity_leg(args, cfg, batch_np, nimg, precisions, want_cpu):
    """One oracle.train_step on the first `nimg` images of the workload's first batch -- timed (cpu_baseline) and used
    as the checker for one step of the engine from the same weights (parity): outside every GPU-timed region."""
    from myolo.config import resolve
    from myolo.model import MaskYOLO
    from tests import helpers as Hh
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, P0 = oracle_step_fn(cfg, batch_np, nimg, args.config)
    t0 = time.perf_counter()
    oout, _ = step()
    dt, n = time.perf_counter() - t0, 1
    if dt < 8.0:                                         # small sample: repeat for a steadier rate
        t0 = time.perf_counter()
        while n < 4 and time.perf_counter() - t0 < 16:
            step(); n += 1
        dt = (time.perf_counter() - t0) / max(n - 1, 1) if n > 1 else dt
    c = resolve(cfg)
    cpu = {"value": nimg / dt, "unit": "images/sec", "cores": cores, "kind": "port",
           "sample": f"oracle.train_step (fwd+bwd+Adam, torch CPU fp32) on {nimg} image(s) of the workload's first batch "
                     f"({args.size}x{args.size}, NB={c['NB']}, NC={c['NC']}, R={c['R']}), {n} call(s)"} if want_cpu else None
    parity = {}
    if args.no_parity:
        return cpu, None
    sub_cfg = bench_config(nimg, args.size, args.config)
    sub = [np.ascontiguousarray(x[:nimg]) for x in batch_np]
    for prec in precisions:
        try:
            model = MaskYOLO("training", sub_cfg, precision=prec)
            model.engine.load_params(P0)
            vals = model.keras_model.train_on_batch(sub)
            torch.cuda.synchronize()
            dev = model.last_outputs
            m = Hh.step_parity(dev, oout, args.size // 8)
            parity[prec] = {k: v for k, v in m.items() if not k.startswith("_")}
            parity[prec].update(loss=vals[0], oracle_loss=oout["loss"].item())
            del model, dev
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        except Exception as e:                      # never lose the throughput line to the side check
            parity[prec] = {"error": repr(e)[:200]}
    parity["against"] = f"oracle.train_step (CPU restatement of the reference path) on {nimg} image(s) of the benchmark workload, one step"
    return cpu, parity


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = bench_config(args.batch, args.size, args.config)
    host_batches = make_host_batches(cfg, args.config, 3, seed=1234 + rank)
    K, W = args.steps, args.warmup
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    main_res = time_precision(args, args.precision, cfg, host_batches, world, rank, local, True)
    clocks = sampler.stop()
    fp32 = None
    if not args.no_fp32_class and args.precision != "tf32x3":
        r2 = time_precision(args, "tf32x3", cfg, host_batches, world, rank, local, False)
        fp32 = {"precision": "tf32x3", "dtype": "3xTF32 forward GEMMs (fp32-grade), tf32 backward GEMMs, f32 accumulate/epilogues",
                "value": r2["value"], "unit": "images/sec", "ms_per_step": r2["ms_per_step"],
                "e2e": r2.get("e2e"), "e2e_train_on_batch": r2.get("e2e_train_on_batch"), "gpu_launches": r2["launches"],
                "loss": r2["loss"]}
    sparse = None
    if not args.no_sparse and args.precision == "h16":
        r3 = time_precision(args, "h16", cfg, host_batches, world, rank, local, False, sparse=True)
        sparse = {"what": "SECONDARY figure, never the headline: the same step with the exact sparse backward of the mask head "
                          "(Engine(sparse_backward=True)): above myolo_mask_bn1 only the rois with a target class carry gradient, "
                          "so conv2..conv4 / deconv backward run on their tiles alone; gradients equal the dense step's up to "
                          "fp32 summation order (tests/test_model_gpu.py::test_sparse_mask_backward_equals_dense). `value` above "
                          "is the DENSE step.",
                  "value": r3["value"], "unit": "images/sec", "ms_per_step": r3["ms_per_step"], "e2e": r3.get("e2e"),
                  "positive_rois_per_step": r3["sparse_stats"]["rois"] / max(r3["sparse_stats"]["steps"], 1),
                  "steps_sparse_dense_nopos": [r3["sparse_stats"]["sparse"], r3["sparse_stats"]["dense_fallback"],
                                               r3["sparse_stats"]["no_positives"]], "loss": r3["loss"]}
    c = main_res["cfg"]
    # ---- roofline of the dominant kernel
    pk, pk_src = peaks()
    roof = None
    kev = main_res["kev"]
    if kev:
        durs = [a.elapsed_time(b) for a, b in kev]
        avg_ms = float(np.mean(durs))
        flops = 2.0 * main_res["n_roi"] * 196 * 2304 * 256
        ach = flops / (avg_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        h16 = args.precision == "h16"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            key = "mask_conv_fwd_h16" if h16 else "mask_conv_fwd"
            per_roi = tj.get(key + "_dram_bytes_per_roi")
            if per_roi is not None:
                traffic, traffic_src = per_roi * main_res["n_roi"], tj.get("source")
            elif args.config == "c2":
                traffic, traffic_src = tj.get(key + "_dram_bytes_per_launch"), tj.get("source_h16" if h16 else "source")
        except Exception:
            pass
        roof = {"kernel": "tc_conv_win_kernel<256> (mask-head 3x3 conv forward, persistent 9-tap tcgen05 kind::%s, CTA pairs)" % ("f16" if h16 else "tf32"),
                "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": avg_ms, "launches_timed": len(durs),
                "algorithmic_flops_per_launch": flops,
                "peak_source": pk_src + (" bf16 sustained (kind::f16 runs at the bf16 rate)" if h16 else
                                         " bf16 sustained (tf32 runs at half the bf16 tensor rate: nominal 1.1 vs 2.25 PFLOP/s)"),
                "step_share": sum(durs) / main_res["ms_total"]}
    cpu = parity = None
    if rank == 0 and (not args.no_cpu_baseline or not args.no_parity):
        nimg = args.batch if args.config == "c2" else 2
        precs = [args.precision] + (["tf32x3"] if fp32 is not None else [])
        cpu, parity = parity_leg(args, cfg, host_batches[0], nimg, precs, world == 1 and not args.no_cpu_baseline)
    if rank == 0:
        fl = flops_per_image(c)
        value = main_res["value"]
        line = {"metric": metric_name(args), "value": value, "unit": "images/sec", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "h16": "f16 operands (mask head) / 3xtf32 (backbone), f32 accumulate"}.get(args.precision, "tf32"),
                "data": "synthetic", "config": workload_config(args, c, world), "precision": args.precision,
                "positive_rois_last_step": main_res["npos"],
                "e2e": main_res.get("e2e"), "e2e_train_on_batch": main_res.get("e2e_train_on_batch"),
                "gpu_launches": main_res["launches"], "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "fp32_class": fp32, "sparse_backward": sparse, "parity": parity,
                "model_tflops_per_s": value * fl / 1e12, "frac_of_conv_roofline": value * fl / 1e12 / pk["bf16_tflops_sustained"],
                "loss": main_res["loss"]}
        if main_res.get("phases"):       # diagnostic (MYOLO_PHASES=1): the extra events cost launch-chain overlap, not a benchmark line
            line["phases_ms"] = main_res["phases"]
            line["invalid"] = "MYOLO_PHASES=1 (diagnostic timing events inside the step)"
        from myolo import _cabi
        if _cabi.WHATIF_SKIP:       # profiling run with entry points switched off: timing experiment, not a benchmark
            line["invalid"] = "MYOLO_WHATIF_SKIP=" + ",".join(sorted(_cabi.WHATIF_SKIP))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
