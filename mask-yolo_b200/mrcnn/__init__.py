"""Minimal stand-in for matterport's `mrcnn` package, which the reference examples import
(example/shapes/dataset_shapes.py:6-7, myolo_utils.py:4) but which is not vendored anywhere.
Only what those call sites use: utils.Dataset, utils.non_max_suppression / compute_iou, and a
visualize module whose plotting entry points are no-ops when matplotlib is absent.  Importing it
also restores the numpy aliases (np.bool, np.float, np.int) the 2018-era examples rely on."""
import numpy as _np

for _alias, _typ in (("bool", bool), ("float", float), ("int", int)):
    if _alias not in _np.__dict__:
        setattr(_np, _alias, _typ)
