"""Drop-in for reference core/yolo_nms.py (class-agnostic padded NMS)."""
import numpy as np
import torch

from .. import _lib


def _dev(t, device, dtype):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    return t.contiguous().to(dtype)


def nms_padded(bboxes, scores, yolo_max_boxes, nms_iou_threshold, nms_score_threshold):
    """tf.image.non_max_suppression_padded(pad_to_max_output_size=True) as used by reference core/yolo_nms.py:26-33.
    bboxes [B,N,4], scores [B,N] (CUDA float32) -> (selected_indices_padded [B,max] int32, num_valid [B] int32)."""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    bboxes = _dev(bboxes, dev, torch.float32)
    scores = _dev(scores, dev, torch.float32)
    B, N = scores.shape
    sel = torch.empty((B, int(yolo_max_boxes)), dtype=torch.int32, device=dev)
    nvalid = torch.empty((B,), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().y3_nms(ctx.handle, _lib.ptr(bboxes), _lib.ptr(scores), B, N, int(yolo_max_boxes),
                                 float(nms_iou_threshold), float(nms_score_threshold), _lib.ptr(sel), _lib.ptr(nvalid),
                                 _lib.ptr(status), _lib.stream_ptr()))
    return sel, nvalid, status


def yolo_nms(outputs, yolo_max_boxes, nms_iou_threshold, nms_score_threshold, check_status=True):
    """reference core/yolo_nms.py:15-34.  outputs = (bboxes [B,N,4], confidence [B,N,1], class_probs [B,N,C]).
    Returns the same 5-tuple: (bboxes [B,N,4] f32, class_indices [B,N] i64, scores [B,N] f32,
    selected_indices_padded [B,max] i32, num_valid_detections [B] i32)."""
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    bboxes, confidence, class_probs = outputs
    bboxes = _dev(bboxes, dev, torch.float32)
    confidence = _dev(confidence, dev, torch.float32)
    class_probs = _dev(class_probs, dev, torch.float32)
    B = bboxes.shape[0]
    bboxes = bboxes.reshape(B, -1, 4)
    N = bboxes.shape[1]
    Cn = class_probs.shape[-1]
    scores = torch.empty((B, N), dtype=torch.float32, device=dev)
    class_indices = torch.empty((B, N), dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().y3_class_reduce(ctx.handle, _lib.ptr(class_probs), _lib.ptr(confidence), B, N, int(Cn),
                                          _lib.ptr(scores), _lib.ptr(class_indices), _lib.stream_ptr()))
    sel, nvalid, status = nms_padded(bboxes, scores, yolo_max_boxes, nms_iou_threshold, nms_score_threshold)
    if check_status and int(status.max().item()) != 0:
        raise _lib.Y3Unsupported("NMS kept-list overflow: more than 1024 surviving boxes without a positive coordinate")
    return bboxes, class_indices, scores, sel, nvalid
