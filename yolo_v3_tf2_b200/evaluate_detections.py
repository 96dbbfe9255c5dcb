"""Mirror of reference evaluate_detections.py (``EvaluateDetections``): per-class preds / gts / tp / fp / fn counters of
detections against ground truth, accumulated on the GPU for whole batches (one CTA per image)."""
import torch

from . import _lib


class EvaluateDetections:
    """reference evaluate_detections.py:16-35, 136-164.  ``evaluate`` takes one image like the reference;
    ``evaluate_batch`` takes the padded batch tensors the detector produces.  ``counters`` is a dict of torch tensors
    with the reference's keys."""

    def __init__(self, nclasses, iou_thresh):
        self.nclasses = int(nclasses)
        self.iou_thresh = float(iou_thresh)
        self._buf = None

    def _buffer(self, device):
        if self._buf is None:
            self._buf = torch.zeros(5 * self.nclasses + 2, dtype=torch.int32, device=device)
        return self._buf

    @property
    def counters(self):
        c = self.nclasses
        b = self._buf if self._buf is not None else torch.zeros(5 * c + 2, dtype=torch.int32)
        return {"preds": b[0:c], "gts": b[c:2 * c], "tp": b[2 * c:3 * c], "fp": b[3 * c:4 * c], "fn": b[4 * c:5 * c],
                "examples": b[5 * c], "errors": b[5 * c + 1]}

    def evaluate_batch(self, det_boxes, det_classes, num_det, gt_boxes, gt_classes, num_gt):
        """det_boxes [B,max_det,4] f32, det_classes [B,max_det] i64, num_det [B] i32 (``Detector.detections`` output),
        gt_boxes [B,max_gt,4] f32, gt_classes [B,max_gt] i32, num_gt [B] i32 -- CUDA tensors."""
        ctx = _lib.context(det_boxes.device.index)
        buf = self._buffer(det_boxes.device)
        B, max_det = det_classes.shape
        max_gt = gt_classes.shape[1]
        args = [det_boxes.contiguous().float(), det_classes.contiguous().long(), num_det.contiguous().int(),
                gt_boxes.contiguous().float(), gt_classes.contiguous().int(), num_gt.contiguous().int()]
        _lib.check(_lib.lib().y3_evaluate(ctx.handle, _lib.ptr(args[0]), _lib.ptr(args[1]), _lib.ptr(args[2]), int(max_det),
                                          _lib.ptr(args[3]), _lib.ptr(args[4]), _lib.ptr(args[5]), int(max_gt), int(B),
                                          self.nclasses, self.iou_thresh, _lib.ptr(buf), _lib.stream_ptr()))
        for t in args:
            t.record_stream(torch.cuda.current_stream())
        return self.counters

    def evaluate(self, pred_bboxes, pred_classes, gt_bboxes, gt_classes):
        """One image, the reference's signature (evaluate_detections.py:136)."""
        dev = torch.device("cuda", _lib.context().device)
        pb = torch.as_tensor(pred_bboxes, dtype=torch.float32, device=dev).reshape(1, -1, 4)
        pc = torch.as_tensor(pred_classes, device=dev).long().reshape(1, -1)
        gb = torch.as_tensor(gt_bboxes, dtype=torch.float32, device=dev).reshape(1, -1, 4)
        gc = torch.as_tensor(gt_classes, device=dev).int().reshape(1, -1)
        nd = torch.tensor([pc.shape[1]], dtype=torch.int32, device=dev)
        ng = torch.tensor([gc.shape[1]], dtype=torch.int32, device=dev)
        if pc.shape[1] == 0:
            pb, pc = torch.zeros((1, 1, 4), device=dev), torch.zeros((1, 1), dtype=torch.long, device=dev)
        if gc.shape[1] == 0:
            gb, gc = torch.zeros((1, 1, 4), device=dev), torch.zeros((1, 1), dtype=torch.int32, device=dev)
        return self.evaluate_batch(pb, pc, nd, gb, gc, ng)
