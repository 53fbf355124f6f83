"""Mask-YOLO base configuration: the class-attribute Config surface of the reference
(myolo/config.py:15-257) -- subclass it and override attributes, exactly as
example/shapes/dataset_shapes.py:14-50 does -- plus `resolve()`, which turns an instance into the
plain dict the sm_100a engine is built from.

The reference reads several fields from the *base class* instead of the instance it is given
(model.py:25 aliases the class as a module global), which makes ShapesConfig
inconsistent at HEAD.  Here everything is read from the instance, and `resolve()` repairs
inherited values that contradict the overridden ones (N_BOX vs ANCHORS, CLASS_WEIGHTS vs
NUM_CLASSES, GRID vs IMAGE_SHAPE, TRAIN_ROIS_PER_IMAGE).
"""
import numpy as np


class Config(object):
    """Base configuration class.  Create a sub-class and override what needs to change."""
    # ---- YOLO head (config.py:22-39)
    NUM_CLASSES = 1 + 1                  # background + classes
    LABELS = ['background', 'food']
    ANCHORS = [1.27, 1.31, 1.95, 1.85, 2.40, 2.72, 3.20, 3.32, 5.06, 5.05]
    N_BOX = 5
    GRID_H, GRID_W = 7, 7
    TRUE_BOX_BUFFER = 10
    BATCH_SIZE = 1
    OBJECT_SCALE = 5.0
    COORD_SCALE = 1.0
    CLASS_SCALE = 1.0
    NO_OBJECT_SCALE = 1.0
    WARM_UP_BATCHES = 0
    CLASS_WEIGHTS = np.ones(NUM_CLASSES, dtype='float32')

    NAME = None                          # override in sub-classes
    GPU_COUNT = 0
    IMAGES_PER_GPU = (BATCH_SIZE / GPU_COUNT) if GPU_COUNT != 0 else 0
    STEPS_PER_EPOCH = 1000
    VALIDATION_STEPS = 5

    # ---- backbone / heads (config.py:74-108)
    BACKBONE = "mobilenet"
    COMPUTE_BACKBONE_SHAPE = None
    BACKBONE_STRIDES = [8]
    FPN_CLASSIF_FC_LAYERS_SIZE = 1024
    TOP_FEATURE_MAP_DEPTH = 256
    SECOND_PHASE_YOLO_DEPTH = 512
    RPN_ANCHOR_SCALES = (32, 64, 128, 256, 512)
    RPN_ANCHOR_RATIOS = [0.5, 1, 2]
    RPN_ANCHOR_STRIDE = 1
    RPN_NMS_THRESHOLD = 0.7

    # ---- image / mask geometry (config.py:122-180)
    USE_MINI_MASK = False
    MINI_MASK_SHAPE = (56, 56)
    IMAGE_RESIZE_MODE = "square"
    IMAGE_MIN_DIM = 224
    IMAGE_MAX_DIM = 224
    IMAGE_MIN_SCALE = 0
    IMAGE_CHANNEL_COUNT = 3
    TRAIN_ROIS_PER_IMAGE = GRID_H * GRID_W * N_BOX
    POOL_SIZE = 7
    MASK_POOL_SIZE = 14
    MASK_SHAPE = [28, 28]
    MAX_GT_INSTANCES = 10

    # ---- optimisation (config.py:200-230)
    LEARNING_RATE = 0.001
    LEARNING_MOMENTUM = 0.9
    WEIGHT_DECAY = 0.0001
    LOSS_WEIGHTS = {
        "yolo_sum_loss": 1.,
        "myolo_mask_loss": 1.,
    }
    TRAIN_BN = False
    GRADIENT_CLIP_NORM = 5.0
    IMAGE_SHAPE = [224, 224, 3]

    def display(self):
        """Display Configuration values."""
        print("\nConfigurations:")
        for a in dir(self):
            if not a.startswith("__") and not callable(getattr(self, a)):
                print("{:30} {}".format(a, getattr(self, a)))
        print("\n")


def resolve(config) -> dict:
    """Engine configuration derived from a Config instance (SURVEY Q1 build rule)."""
    shape = list(config.IMAGE_SHAPE)
    S = int(shape[0])
    if int(shape[1]) != S:
        raise Exception("Mask-YOLO builds its cell grid as a transpose (model.py:1447): IMAGE_SHAPE must be square")
    anchors = [float(a) for a in config.ANCHORS]
    nb = len(anchors) // 2                       # the anchors, not the inherited N_BOX, are authoritative
    nc = int(config.NUM_CLASSES)
    g = S // 32                                   # the only grid the backbone strides can produce
    cw = np.asarray(config.CLASS_WEIGHTS, dtype=np.float32)
    if cw.shape[0] != nc:
        cw = np.ones(nc, dtype=np.float32)
    return dict(S=S, G=g, NB=nb, NC=nc, TB=int(config.TRUE_BOX_BUFFER), MAXGT=int(config.MAX_GT_INSTANCES),
                R=g * g * nb, ANCHORS=anchors[:2 * nb], POOL=int(config.MASK_POOL_SIZE),
                MASK_SHAPE=[int(v) for v in config.MASK_SHAPE], OBJECT_SCALE=float(config.OBJECT_SCALE),
                NO_OBJECT_SCALE=float(config.NO_OBJECT_SCALE), COORD_SCALE=float(config.COORD_SCALE),
                CLASS_SCALE=float(config.CLASS_SCALE), CLASS_WEIGHTS=cw, WARM_UP_BATCHES=int(config.WARM_UP_BATCHES),
                LOSS_WEIGHTS=dict(config.LOSS_WEIGHTS))
