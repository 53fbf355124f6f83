"""Shared builders for the model-level parity tests, smoke() and bench.py's cpu_baseline leg:
synthetic Shapes-like batches in the exact format BatchGenerator yields, and the
oracle-side config dict."""
import numpy as np
import torch


def engine_cfg(S=64, NB=3, NC=4, TB=5, anchors=None, maxgt=4):
    G = S // 32
    anchors = anchors if anchors is not None else [0.5, 0.6, 0.9, 0.8, 1.2, 1.3, 1.6, 1.5, 0.4, 1.0][:2 * NB]
    return dict(S=S, G=G, NB=NB, NC=NC, TB=TB, MAXGT=maxgt, R=G * G * NB, ANCHORS=list(anchors), POOL=14,
                MASK_SHAPE=[28, 28], OBJECT_SCALE=5.0, NO_OBJECT_SCALE=1.0, COORD_SCALE=1.0, CLASS_SCALE=1.0,
                CLASS_WEIGHTS=np.ones(NC, dtype="float32"), WARM_UP_BATCHES=0,
                LOSS_WEIGHTS={"yolo_sum_loss": 1.0, "myolo_mask_loss": 1.0})


def oracle_cfg(c):
    return dict(GRID_H=c["G"], GRID_W=c["G"], N_BOX=c["NB"], NUM_CLASSES=c["NC"], ANCHORS=c["ANCHORS"],
                TRAIN_ROIS_PER_IMAGE=c["R"], MASK_SHAPE=c["MASK_SHAPE"], MASK_POOL_SIZE=c["POOL"],
                COORD_SCALE=c["COORD_SCALE"], NO_OBJECT_SCALE=c["NO_OBJECT_SCALE"], OBJECT_SCALE=c["OBJECT_SCALE"],
                CLASS_SCALE=c["CLASS_SCALE"], CLASS_WEIGHTS=c["CLASS_WEIGHTS"], WARM_UP_BATCHES=c["WARM_UP_BATCHES"],
                TRUE_BOX_BUFFER=c["TB"], LOSS_WEIGHTS=c["LOSS_WEIGHTS"])


def batch_from_boxes(c, B, image, boxes_norm, seed):
    """Training inputs whose GT instances are ellipses inscribed in the given normalised boxes
    (list per image of (x1,y1,x2,y2)); the YOLO target / true-box buffer are encoded like
    BatchGenerator.__getitem__ does (myolo_utils.py:769-820)."""
    rng = np.random.RandomState(seed)
    S, G, NB, NC, TB, M = c["S"], c["G"], c["NB"], c["NC"], c["TB"], c["MAXGT"]
    ids = np.zeros((B, M), np.int32)
    gtb = np.zeros((B, M, 4), np.float32)
    masks = np.zeros((B, S, S, M), np.uint8)
    yt = np.zeros((B, G, G, NB, 5 + NC), np.float32)
    tb = np.zeros((B, 1, 1, 1, TB, 4), np.float32)
    yy, xx = np.mgrid[0:S, 0:S]
    anc = np.asarray(c["ANCHORS"], np.float32).reshape(NB, 2)
    for b in range(B):
        for m, bx in enumerate(boxes_norm[b][:M]):
            x1, y1, x2, y2 = [float(np.clip(v, 0.0, 1.0)) * (S - 1) for v in bx]
            x1, y1, x2, y2 = int(round(x1)), int(round(y1)), int(round(x2)) + 1, int(round(y2)) + 1
            if x2 - x1 < 3 or y2 - y1 < 3:
                continue
            cx, cy, rx, ry = (x1 + x2 - 1) / 2.0, (y1 + y2 - 1) / 2.0, (x2 - x1) / 2.0, (y2 - y1) / 2.0
            mk = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0
            mk[y1, x1:x2] = True; mk[y2 - 1, x1:x2] = True; mk[y1:y2, x1] = True; mk[y1:y2, x2 - 1] = True
            masks[b, :, :, m] = mk
            gtb[b, m] = [x1, y1, x2, y2]
            cls = rng.randint(1, NC)
            ids[b, m] = cls
            gcx, gcy = 0.5 * (x1 + x2) / (S / G), 0.5 * (y1 + y2) / (S / G)
            gw, gh = (x2 - x1) / (S / G), (y2 - y1) / (S / G)
            gx_, gy_ = int(gcx), int(gcy)
            if gx_ < G and gy_ < G:
                inter = np.minimum(anc[:, 0], gw) * np.minimum(anc[:, 1], gh)
                iou = inter / (anc[:, 0] * anc[:, 1] + gw * gh - inter)
                a = int(np.argmax(iou))
                yt[b, gy_, gx_, a, :5] = [gcx, gcy, gw, gh, 1.0]
                yt[b, gy_, gx_, a, 5:] = 0
                yt[b, gy_, gx_, a, 5 + cls] = 1
                tb[b, 0, 0, 0, m % TB] = [gcx, gcy, gw, gh]
    return [torch.as_tensor(image), torch.tensor(tb), torch.tensor(yt), torch.tensor(ids), torch.tensor(gtb),
            torch.tensor(masks)]


def random_boxes(B, n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for b in range(B):
        c = rng.rand(n, 2) * 0.6 + 0.2
        wh = rng.rand(n, 2) * 0.3 + 0.15
        out.append(np.concatenate([c - wh / 2, c + wh / 2], 1).tolist())
    return out


def to_oracle_inputs(inputs):
    image, tb, yt, ids, gtb, masks = inputs
    return [image, tb, yt, ids, gtb, masks.bool()]


def to_device(inputs, dev="cuda"):
    return [t.contiguous().to(dev) for t in inputs]


def box_iou_pairs(a, b):
    """Element-wise IoU of two box tensors [..., 4] (x1, y1, x2, y2); two empty boxes count as identical (1)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    iw = (torch.minimum(a[..., 2], b[..., 2]) - torch.maximum(a[..., 0], b[..., 0])).clamp(min=0)
    ih = (torch.minimum(a[..., 3], b[..., 3]) - torch.maximum(a[..., 1], b[..., 1])).clamp(min=0)
    inter = iw * ih
    area = lambda t: (t[..., 2] - t[..., 0]).clamp(min=0) * (t[..., 3] - t[..., 1]).clamp(min=0)   # noqa: E731
    union = area(a) + area(b) - inter
    return torch.where(union > 0, inter / union.clamp(min=1e-300), torch.ones_like(union))


def mask_iou(a, b, thr=0.5, return_union=False):
    """IoU of the masks binarised at `thr`, per (roi, class): inputs [..., H, W, NC] -> [..., NC]; two empty masks
    count as identical (1).  With return_union also the pixel count of the union (a mask of a few pixels turns one
    probability on either side of `thr` into an IoU of 0 or 1/2)."""
    a, b = a.detach().cpu() >= thr, b.detach().cpu() >= thr
    inter = (a & b).sum(dim=(-3, -2)).double()
    union = (a | b).sum(dim=(-3, -2)).double()
    iou = torch.where(union > 0, inter / union.clamp(min=1), torch.ones_like(union))
    return (iou, union) if return_union else iou


def sample_validity(rois, F_, pool=14):
    """Sample-validity pattern of crop_and_resize for [..., 4] boxes on an F x F map, in the oracle's float32 arithmetic
    (oracle.crop_and_resize): [n, 2, pool] booleans.  Two sets of ROI coordinates that differ in this pattern make
    crop_and_resize zero different samples -- a discontinuity of the reference's own function, in any implementation."""
    r = rois.detach().float().cpu().reshape(-1, 4)
    ar = torch.arange(pool, dtype=torch.float32)[None, :]
    pats = []
    for lo, hi in ((0, 2), (1, 3)):
        step = (r[:, hi] - r[:, lo]) * (F_ - 1) / (pool - 1)
        pos = (r[:, lo] * (F_ - 1))[:, None] + ar * step[:, None]
        pats.append((pos >= 0) & (pos <= F_ - 1))
    return torch.stack(pats, 1)


def step_parity(dev, ref, F_):
    """Outputs of one fit step of the engine (`dev`) against oracle.train_step's (`ref`): the figures BASELINE.json's metric
    asks for (box / mask IoU vs the reference path) and the 1e-3 checks of north_star.  Boxes and class scores: max abs error
    and the same relative to the output scale (decoded boxes reach |coordinate| ~ 20 for large anchors).  Masks: over the
    ROIs whose sample-validity pattern agrees under both sets of ROI coordinates, in images with identical ROI selection."""
    B, R = ref["target_class_ids"].shape
    NC = ref["myolo_mask"].shape[-1]
    out = {}
    for k, name in (("yolo_proposals", "box"), ("yolo_output", "class_score")):
        r = ref[k].float()
        e = (dev[k].detach().cpu().reshape(r.shape) - r).abs().max().item()
        out[f"max_abs_{name}_err"] = e
        out[f"{name}_scale"] = r.abs().max().item()
        out[f"max_{name}_err_rel_to_scale"] = e / max(1.0, r.abs().max().item())
    biou = box_iou_pairs(dev["yolo_proposals"], ref["yolo_proposals"])
    same_img = (dev["target_class_ids"].cpu().int() == ref["target_class_ids"].int()).all(dim=1)
    flips = (sample_validity(dev["output_rois"], F_) != sample_validity(ref["output_rois"], F_)).flatten(1).any(dim=1)
    keep = same_img[:, None].expand(B, R).reshape(-1) & ~flips
    dm = (dev["myolo_mask"].detach().cpu().reshape(B * R, -1, NC) - ref["myolo_mask"].float().reshape(B * R, -1, NC)).abs()
    mh = int(round(dm.shape[1] ** 0.5))
    miou, munion = mask_iou(dev["myolo_mask"].detach().cpu().reshape(B * R, mh, mh, NC)[keep],
                            ref["myolo_mask"].reshape(B * R, mh, mh, NC)[keep], return_union=True)
    big = munion >= 16          # binarised masks of fewer pixels: one probability within the error of 0.5 decides the IoU
    out.update(box_iou_mean=biou.mean().item(), box_iou_min=biou.min().item(),
               mask_iou_mean=miou.mean().item(), mask_iou_min=miou.min().item(),
               mask_iou_min_union_ge_16px=miou[big].min().item() if bool(big.any()) else None,
               masks_with_union_lt_16px=int((~big & (munion > 0)).sum()),
               max_abs_mask_err=dm[keep].max().item() if bool(keep.any()) else None,
               rois_compared=int(keep.sum()), rois_total=B * R, rois_with_flipped_border_sample=int(flips.sum()),
               images_with_identical_roi_selection=int(same_img.sum()), images=B,
               roi_selection_identical=bool(same_img.all()),
               positive_rois=int((ref["target_class_ids"] > 0).sum().item()))
    out["_same_img"], out["_keep"], out["_flips"] = same_img, keep, flips
    return out
