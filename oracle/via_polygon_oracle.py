"""CPU oracle for the VIA-polygon rasteriser  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` may import this module; the product package never does (tests/test_cabi.py enforces it).

What it restates: ``RiceDataset.load_mask`` (example/rice/rice_dataset.py:135-159; example/food/rice_dataset.py is the
same file), whose arithmetic lives in ``skimage.draw.polygon`` -- a third-party, un-pinned dependency (no requirements
file; scikit-image 0.13 / 0.14 are the releases contemporary with the reference, 2018) that is NOT installable here.
Its published algorithm is restated below in plain Python loops over Python floats (= IEEE float64), one reference
statement per line:

    skimage/draw/_draw.pyx  _polygon(r, c, shape):
        minr = int(max(0, r.min()));  maxr = int(ceil(r.max()))
        minc = int(max(0, c.min()));  maxc = int(ceil(c.max()))
        if shape is not None:  maxr = min(shape[0] - 1, maxr);  maxc = min(shape[1] - 1, maxc)
        for r in range(minr, maxr + 1):
            for c in range(minc, maxc + 1):
                if point_in_polygon(nr_verts, cptr, rptr, c, r):  rr.append(r); cc.append(c)
    skimage/_shared/geometry.pxd  point_in_polygon(nr_verts, xp, yp, x, y):
        c = 0;  j = nr_verts - 1
        for i in range(nr_verts):
            if ((((yp[i] <= y) and (y < yp[j])) or ((yp[j] <= y) and (y < yp[i])))
                    and (x < (xp[j] - xp[i]) * (y - yp[i]) / (yp[j] - yp[i]) + xp[i])):
                c = not c
            j = i

PARITY UNPINNED against scikit-image itself (the reference holds no golden masks either: its datasets directory ships
the VIA json files without the images).  Pinned by hand-computed cases in tests/test_via_polygons.py (axis-aligned
squares: the half-open [min, max) rule on both axes; a right triangle; a concave and a self-intersecting outline).
"""
import math

import numpy as np


def point_in_polygon(xp, yp, x, y):
    """skimage/_shared/geometry.pxd point_in_polygon: crossing parity of the ray from (x, y) towards +x."""
    c = False
    n = len(xp)
    j = n - 1
    for i in range(n):
        if ((yp[i] <= y < yp[j]) or (yp[j] <= y < yp[i])) and \
                (x < (xp[j] - xp[i]) * (y - yp[i]) / (yp[j] - yp[i]) + xp[i]):
            c = not c
        j = i
    return c


def polygon(r, c, shape=None):
    """skimage.draw.polygon(r, c, shape=None) -> (rr, cc), row-major order like the library's double loop."""
    r = [float(v) for v in np.atleast_1d(r)]
    c = [float(v) for v in np.atleast_1d(c)]
    minr = int(max(0, min(r)))
    maxr = int(math.ceil(max(r)))
    minc = int(max(0, min(c)))
    maxc = int(math.ceil(max(c)))
    if shape is not None:
        maxr = min(shape[0] - 1, maxr)
        maxc = min(shape[1] - 1, maxc)
    rr, cc = [], []
    for ri in range(minr, maxr + 1):
        for ci in range(minc, maxc + 1):
            if point_in_polygon(c, r, float(ci), float(ri)):
                rr.append(ri)
                cc.append(ci)
    return np.array(rr, dtype=np.intp), np.array(cc, dtype=np.intp)


def load_mask(polygons, height, width):
    """RiceDataset.load_mask (rice_dataset.py:135-159) on a list of VIA `shape_attributes` dicts:
    (bool [height, width, n], int32 ones [n]).  Like the reference, a polygon reaching outside the image raises IndexError."""
    mask = np.zeros([height, width, len(polygons)], dtype=np.uint8)
    for i, p in enumerate(polygons):
        rr, cc = polygon(p['all_points_y'], p['all_points_x'])
        mask[rr, cc, i] = 1
    return mask.astype(bool), np.ones([mask.shape[-1]], dtype=np.int32)
