"""VIA polygon datasets (the reference's rice / food examples), host and device.

RiceConfig mirrors example/rice/rice_dataset.py:60-82; RiceDataset mirrors 89-168 (load_rice for VIA 1.x dict and 2.x
list `regions`, load_mask, image_reference) over the mrcnn Dataset protocol, so load_image_gt and BatchGenerator consume
it unchanged.  Two things the reference takes from scikit-image, which is not a dependency of this package:
  * `skimage.io.imread` (only the image size is used in load_rice): cv2 reads the file instead;
  * `skimage.draw.polygon`: `polygon()` below restates its published algorithm (float64 crossing-number test over the
    polygon's clipped bounding box, see csrc/polygon_pip.h for the rule and its provenance) in vectorised numpy, with the
    library's output order (row-major).
`DevicePolygons` is the device form of load_mask: all instances of an image in one launch of myolo_polygon_masks
(csrc/polygon.cu), same bytes as the host form.  It needs the CUDA library; there is no fallback from it to the host form.
"""
import json
import math
import os

import numpy as np

from mrcnn import utils
from .config import Config


class RiceConfig(Config):
    """example/rice/rice_dataset.py:60-82 (that file names its configuration, source and class "food")."""
    NAME = "food"
    IMAGES_PER_GPU = 2
    GPU_COUNT = 0
    NUM_CLASSES = 1 + 1


class FoodExampleRiceConfig(RiceConfig):
    """The RiceConfig of example/food/rice_dataset.py:60-82: the same file with the two words swapped (NAME "rice")."""
    NAME = "rice"


def polygon(r, c, shape=None):
    """skimage.draw.polygon(r, c, shape=None) -> (rr, cc): the pixels whose centre passes the crossing-number test, rows
    first.  Every operation is the library's, in its order, on float64 arrays (numpy rounds each one individually)."""
    r = np.atleast_1d(np.asarray(r, dtype=np.float64))
    c = np.atleast_1d(np.asarray(c, dtype=np.float64))
    minr, maxr = int(max(0, r.min())), int(math.ceil(r.max()))
    minc, maxc = int(max(0, c.min())), int(math.ceil(c.max()))
    if shape is not None:
        maxr, maxc = min(shape[0] - 1, maxr), min(shape[1] - 1, maxc)
    if maxr < minr or maxc < minc:
        return np.zeros(0, np.intp), np.zeros(0, np.intp)
    y = np.arange(minr, maxr + 1, dtype=np.float64)[:, None]
    x = np.arange(minc, maxc + 1, dtype=np.float64)[None, :]
    inside = np.zeros((y.shape[0], x.shape[1]), dtype=bool)
    n = r.shape[0]
    j = n - 1
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            span = ((r[i] <= y) & (y < r[j])) | ((r[j] <= y) & (y < r[i]))          # [rows, 1]
            if span.any():
                t = (c[j] - c[i]) * (y - r[i]) / (r[j] - r[i]) + c[i]               # [rows, 1]
                inside ^= span & (x < t)
            j = i
    rr, cc = np.nonzero(inside)
    return (rr + minr).astype(np.intp), (cc + minc).astype(np.intp)


def _check_inside(polygons, height, width):
    """The reference indexes `mask[rr, cc, i]` with unclipped coordinates: an outline that yields a pixel beyond the
    image raises IndexError there.  The same error is raised here BEFORE a device launch (the kernel would clip).  Only an
    outline with a vertex beyond the last row / column can yield such a pixel; those (rare) ones are rasterised on the
    host to decide."""
    for k, p in enumerate(polygons):
        ys, xs = p['all_points_y'], p['all_points_x']
        if len(ys) == 0 or len(ys) != len(xs):
            raise ValueError("polygon %d: empty or ragged vertex lists" % k)
        if max(ys) > height - 1 or max(xs) > width - 1:
            rr, cc = polygon(ys, xs)
            if len(rr) and (rr.max() >= height or cc.max() >= width):
                raise IndexError("polygon %d reaches pixel (%d, %d) of a %d x %d image" % (k, rr.max(), cc.max(), height, width))


class DevicePolygons(object):
    """load_mask on the device.  masks(polygons, height, width) -> uint8 cuda tensor [height, width, n] (the reference's
    layout; `.bool()` gives load_mask's dtype).  The vertex lists are packed on the host (a few hundred bytes) and copied
    with the launch; the mask tensor is written once by the kernel and never touches the host."""

    def __init__(self, device=0):
        import torch
        from . import _cabi as C
        C.lib()                                   # raises if the library is missing: no CPU path behind this class
        self.C, self.torch = C, torch
        self.dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device

    def masks(self, polygons, height, width):
        torch, C = self.torch, self.C
        n = len(polygons)
        if n == 0:
            return torch.zeros((height, width, 0), dtype=torch.uint8, device=self.dev)
        if n > 128:
            raise ValueError("at most 128 instances per image")
        _check_inside(polygons, height, width)
        off = np.zeros(n + 1, np.int32)
        off[1:] = np.cumsum([len(p['all_points_y']) for p in polygons])
        vy = np.concatenate([np.asarray(p['all_points_y'], np.float64) for p in polygons])
        vx = np.concatenate([np.asarray(p['all_points_x'], np.float64) for p in polygons])
        d_vy, d_vx = torch.from_numpy(vy).to(self.dev), torch.from_numpy(vx).to(self.dev)
        d_off = torch.from_numpy(off).to(self.dev)
        out = torch.empty((height, width, n), dtype=torch.uint8, device=self.dev)
        ws = torch.empty(4 * n, dtype=torch.int32, device=self.dev)
        C.call("myolo_polygon_masks", d_vy, d_vx, d_off, n, height, width, n, ws, out,
               torch.cuda.current_stream(self.dev).cuda_stream)
        return out


class RiceDataset(utils.Dataset):
    """example/rice/rice_dataset.py:89-168.  `source` is the word that file uses for its dataset source, its one class and
    its annotation file (`via_<source>_annotation.json`): "food" in example/rice (the default), "rice" in the otherwise
    identical example/food/rice_dataset.py."""

    def __init__(self, source="food", class_map=None):
        super().__init__(class_map)
        self.source = source

    def load_rice(self, dataset_dir, subset, annotation_file=None):
        """VIA json -> one image record per annotated file, with its polygons (`shape_attributes` dicts).  The reference
        hard-codes the file name (rice_dataset.py:102); it is the default here."""
        src = self.source
        annotation_file = annotation_file or "via_%s_annotation.json" % src
        self.add_class(src, 1, src)
        assert subset in ["train", "val"]
        dataset_dir = os.path.join(dataset_dir, subset)
        with open(os.path.join(dataset_dir, annotation_file)) as fh:
            annotations = list(json.load(fh).values())
        annotations = [a for a in annotations if a['regions']]
        for a in annotations:
            if type(a['regions']) is dict:                                   # VIA 1.x
                polygons = [r['shape_attributes'] for r in a['regions'].values()]
            else:                                                            # VIA 2.x
                polygons = [r['shape_attributes'] for r in a['regions']]
            image_path = os.path.join(dataset_dir, a['filename'])
            height, width = self._image_size(image_path)
            self.add_image(src, image_id=a['filename'], path=image_path, width=width, height=height, polygons=polygons)

    @staticmethod
    def _image_size(path):
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError("cannot read " + str(path))
        return img.shape[:2]

    def load_mask(self, image_id):
        """(bool [height, width, instances], int32 ones [instances]); rice_dataset.py:135-159."""
        info = self.image_info[image_id]
        if info["source"] != self.source:
            return super().load_mask(image_id)
        mask = np.zeros([info["height"], info["width"], len(info["polygons"])], dtype=np.uint8)
        for i, p in enumerate(info["polygons"]):
            rr, cc = polygon(p['all_points_y'], p['all_points_x'])
            mask[rr, cc, i] = 1
        return mask.astype(bool), np.ones([mask.shape[-1]], dtype=np.int32)

    def load_mask_device(self, image_id, rasteriser):
        """load_mask through a DevicePolygons instance: (uint8 cuda tensor [height, width, instances], int32 ones)."""
        info = self.image_info[image_id]
        m = rasteriser.masks(info["polygons"], info["height"], info["width"])
        return m, np.ones([m.shape[-1]], dtype=np.int32)

    def image_reference(self, image_id):
        info = self.image_info[image_id]
        if info["source"] == self.source:
            return info["path"]
        return super().image_reference(image_id)
