"""example/shapes/train_shapes.py:8 imports mrcnn.model yet calls modellib.MaskYOLO."""
from myolo.model import MaskYOLO  # noqa: F401
