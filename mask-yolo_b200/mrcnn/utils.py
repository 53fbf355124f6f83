"""mrcnn.utils shim: Dataset base class and greedy NMS (semantics restated in SURVEY.md Q8)."""
import numpy as np


class Dataset(object):
    """Base dataset: subclasses add classes/images and override load_image / load_mask."""

    def __init__(self, class_map=None):
        self._image_ids = []
        self.image_info = []
        self.class_info = [{"source": "", "id": 0, "name": "BG"}]
        self.source_class_ids = {}

    def add_class(self, source, class_id, class_name):
        assert "." not in source, "Source name cannot contain a dot"
        for info in self.class_info:
            if info["source"] == source and info["id"] == class_id:
                return
        self.class_info.append({"source": source, "id": class_id, "name": class_name})

    def add_image(self, source, image_id, path, **kwargs):
        info = {"id": image_id, "source": source, "path": path}
        info.update(kwargs)
        self.image_info.append(info)

    def image_reference(self, image_id):
        return ""

    def prepare(self, class_map=None):
        self.num_classes = len(self.class_info)
        self.class_ids = np.arange(self.num_classes)
        self.class_names = [",".join(c["name"].split(",")[:1]) for c in self.class_info]
        self.num_images = len(self.image_info)
        self._image_ids = np.arange(self.num_images)
        self.class_from_source_map = {"{}.{}".format(i["source"], i["id"]): k for k, i in enumerate(self.class_info)}
        self.image_from_source_map = {"{}.{}".format(i["source"], i["id"]): k for k, i in enumerate(self.image_info)}
        self.sources = sorted(set(i["source"] for i in self.class_info))
        self.source_class_ids = {}
        for source in self.sources:
            self.source_class_ids[source] = [k for k, i in enumerate(self.class_info) if k == 0 or i["source"] == source]

    def map_source_class_id(self, source_class_id):
        return self.class_from_source_map[source_class_id]

    @property
    def image_ids(self):
        return self._image_ids

    def source_image_link(self, image_id):
        return self.image_info[image_id]["path"]

    def load_image(self, image_id):
        import cv2
        img = cv2.imread(self.image_info[image_id]["path"], cv2.IMREAD_COLOR)
        if img is None:
            raise IOError("cannot read " + str(self.image_info[image_id]["path"]))
        return img[:, :, ::-1].copy()

    def load_mask(self, image_id):
        return np.empty([0, 0, 0]), np.empty([0], np.int32)


def compute_iou(box, boxes, box_area, boxes_area):
    """IoU of one (y1,x1,y2,x2) box against many."""
    y1 = np.maximum(box[0], boxes[:, 0])
    y2 = np.minimum(box[2], boxes[:, 2])
    x1 = np.maximum(box[1], boxes[:, 1])
    x2 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
    return inter / (box_area + boxes_area[:] - inter)


def non_max_suppression(boxes, scores, threshold):
    """Greedy NMS over (y1,x1,y2,x2) boxes; drops boxes with IoU > threshold against a kept one."""
    assert boxes.shape[0] > 0
    boxes = boxes.astype(np.float32)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    order = scores.argsort()[::-1]
    pick = []
    while len(order) > 0:
        i = order[0]
        pick.append(i)
        iou = compute_iou(boxes[i], boxes[order[1:]], area[i], area[order[1:]])
        order = order[1:][iou <= threshold]
    return np.array(pick, dtype=np.int32)
