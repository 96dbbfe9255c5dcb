"""B200-native YOLOv3 inference hot path (backbone + neck + heads -> decode -> NMS).

Drop-in for the reference's ``core.parse_model.ParseModel`` / ``core.yolo_decode_layer.yolo_decode`` /
``core.yolo_nms_layer.YoloNmsLayer`` surface (ronen-halevy/yolo-v3-tf2); every tensor op runs in hand-written sm_100a
CUDA behind the C ABI in include/y3b200.h.  See DESIGN.md.
"""
from . import _lib  # noqa: F401
from .core.parse_model import ParseModel  # noqa: F401
from .core.yolo_decode_layer import yolo_decode  # noqa: F401
from .core.yolo_nms import yolo_nms  # noqa: F401
from .core.yolo_nms_layer import YoloNmsLayer  # noqa: F401
from .core.utils import get_anchors, resize_image, preprocess_images  # noqa: F401
from .inference import Inference, Detector  # noqa: F401

__version__ = "0.1.0"
