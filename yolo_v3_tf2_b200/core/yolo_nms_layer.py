"""Drop-in for reference core/yolo_nms_layer.py."""
from .yolo_nms import yolo_nms


class YoloNmsLayer:
    """reference core/yolo_nms_layer.py:16-29: stores the three parameters, ``call`` forwards to ``yolo_nms``."""

    def __init__(self, yolo_max_boxes, nms_iou_threshold, nms_score_threshold, **kwargs):
        self.yolo_max_boxes = yolo_max_boxes
        self.nms_iou_threshold = nms_iou_threshold
        self.nms_score_threshold = nms_score_threshold

    def call(self, decoded_outputs, **kwargs):
        return yolo_nms(decoded_outputs, self.yolo_max_boxes, self.nms_iou_threshold, self.nms_score_threshold)

    __call__ = call
