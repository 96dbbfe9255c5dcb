"""The assembly of reference inference.py:83-117 (model -> yolo_decode -> YoloNmsLayer) and its
``gather_valid_detections_results`` (inference.py:21-28), without the file / plotting loop (host I/O, out of scope)."""
import numpy as np
import torch
import yaml

from . import _lib
from .core.parse_model import ParseModel, Y3Model
from .core.utils import get_anchors
from .core.yolo_decode_layer import yolo_decode
from .core.yolo_nms import nms_padded
from .core.yolo_nms_layer import YoloNmsLayer


class Inference:
    @staticmethod
    def gather_valid_detections_results(bboxes_padded, class_indices_padded, scores_padded,
                                        selected_indices_padded, num_valid_detections):
        """reference inference.py:21-28, for ONE image: rows ``selected[:num_valid]`` of bboxes [N,4], classes [N],
        scores [N]."""
        n = int(num_valid_detections)
        idx = torch.as_tensor(selected_indices_padded)[:n].long()
        bboxes = torch.as_tensor(bboxes_padded)[idx.to(torch.as_tensor(bboxes_padded).device)]
        classes = torch.as_tensor(class_indices_padded)[idx.to(torch.as_tensor(class_indices_padded).device)]
        scores = torch.as_tensor(scores_padded)[idx.to(torch.as_tensor(scores_padded).device)]
        return bboxes, classes, scores


def gather_detections_batched(bboxes, class_indices, scores, selected, num_valid, packed=False):
    """Batched, zero-padded device version of ``gather_valid_detections_results``:
    -> (boxes [B,max,4] f32, classes [B,max] i64, scores [B,max] f32); rows >= num_valid[b] are zero.
    ``packed=True`` appends the float32 record tensor [B, max*6 + 1] of ``distributed.pack_detections``, written by the
    same kernel."""
    ctx = _lib.context()
    B, N = scores.shape
    mx = selected.shape[1]
    dev = scores.device
    ob = torch.empty((B, mx, 4), dtype=torch.float32, device=dev)
    oc = torch.empty((B, mx), dtype=torch.int64, device=dev)
    os_ = torch.empty((B, mx), dtype=torch.float32, device=dev)
    rec = torch.empty((B, mx * 6 + 1), dtype=torch.float32, device=dev) if packed else None
    _lib.check(_lib.lib().y3_gather_detections_packed(ctx.handle, _lib.ptr(bboxes), _lib.ptr(class_indices),
                                                      _lib.ptr(scores), _lib.ptr(selected), _lib.ptr(num_valid), B, N, mx,
                                                      _lib.ptr(ob), _lib.ptr(oc), _lib.ptr(os_), _lib.ptr(rec),
                                                      _lib.stream_ptr()))
    if packed:
        return ob, oc, os_, rec
    return ob, oc, os_


class Detector:
    """model -> decode -> NMS on one GPU, the composition of reference inference.py:109-117.

    ``detect(x)`` returns the reference's 5-tuple (bboxes [B,N,4], class_indices [B,N] int64, scores [B,N],
    selected_indices_padded [B,max] int32, num_valid_detections [B] int32).  ``fused=True`` (default) lets the decode
    kernel emit scores / class ids itself; ``fused=False`` runs the reference's exact op sequence
    (yolo_decode -> YoloNmsLayer).  Results are identical either way."""

    def __init__(self, model: Y3Model, anchors_table, nclasses, yolo_max_boxes=100, nms_iou_threshold=0.5,
                 nms_score_threshold=0.1, fused=True, check_status=True):
        self.model = model
        self.anchors = np.asarray(anchors_table, dtype=np.float32)
        self.nclasses = int(nclasses)
        self.max_boxes = int(yolo_max_boxes)
        self.iou_thr = float(nms_iou_threshold)
        self.score_thr = float(nms_score_threshold)
        self.fused = fused
        # NMS reports a per-image status (1 = its kept list overflowed: only possible with a score threshold <= 0, where
        # suppressed all-zero boxes stay candidates).  Eager calls raise on it (one device->host read); inside a CUDA
        # graph capture nothing can be read back, so the status tensor is left in ``last_status`` for the caller.
        self.check_status = check_status
        self.last_status = None
        self.nms_layer = YoloNmsLayer(yolo_max_boxes, nms_iou_threshold, nms_score_threshold)

    @classmethod
    def from_config(cls, detect_config_file, search_dirs=()):
        """Build from a reference ``detect_config*.yaml`` (keys: model_config_file, classes_name_file, anchors_file,
        yolo_max_boxes, nms_iou_threshold, nms_score_threshold, input_weights_path; inference.py:52-71)."""
        with open(detect_config_file, "r") as f:
            cfg = yaml.safe_load(f)
        from . import graph as graph_mod
        anchors = get_anchors(graph_mod._resolve(cfg["anchors_file"], search_dirs)).astype(np.float32)
        names = [c.strip() for c in open(graph_mod._resolve(cfg["classes_name_file"], search_dirs)).readlines()]
        nclasses = len(names)   # inference.py:84-85
        with open(graph_mod._resolve(cfg["model_config_file"], search_dirs), "r") as f:
            model_config = yaml.safe_load(f)
        model = ParseModel().build_model(None, model_config["sub_models_configs"], model_config["output_stage"],
                                         nclasses=nclasses, search_dirs=search_dirs)
        det = cls(model, anchors, nclasses, cfg["yolo_max_boxes"], cfg["nms_iou_threshold"], cfg["nms_score_threshold"])
        det.class_names = names
        det.config = cfg
        return det

    def detect(self, x):
        if self.fused:
            # pitched head outputs (pixel pitch rounded up to 4 floats): TMA-store epilogue in the head convs
            grids = self.model(x, padded=True)
            # compact decode: NMS reads boxes and scores only, so objectness / class probabilities are never written
            bboxes, _, _, scores, cls = yolo_decode(grids, self.anchors, self.nclasses, compact=True)
            sel, nvalid, status = nms_padded(bboxes, scores, self.max_boxes, self.iou_thr, self.score_thr)
            self.last_status = status
            if self.check_status and not torch.cuda.is_current_stream_capturing() and int(status.max().item()) != 0:
                raise _lib.Y3Unsupported("NMS kept-list overflow: more than 1024 surviving boxes without a positive "
                                         "coordinate (selected indices / num_valid are truncated)")
            return bboxes, cls, scores, sel, nvalid
        grids = self.model(x)
        decoded = yolo_decode(grids, self.anchors, self.nclasses)
        return self.nms_layer(decoded)

    __call__ = detect
    predict = detect

    def detections(self, x, packed=False):
        """detect + batched gather: (boxes [B,max,4], classes [B,max], scores [B,max], num_valid [B]); ``packed=True``
        appends the float32 record tensor [B, max*6 + 1] the multi-GPU gather sends (same kernel)."""
        bboxes, cls, scores, sel, nvalid = self.detect(x)
        out = gather_detections_batched(bboxes, cls, scores, sel, nvalid, packed=packed)
        return out[:3] + (nvalid,) + out[3:]

    def detections_graphed(self, x, packed=False, static_input=False):
        """``detections`` replayed from a CUDA graph: the ~80 launches of a step (75 convs with their programmatic
        dependencies, decode, NMS, gather) are captured once per input shape and then cost one launch, which takes the
        host out of the loop.  ``x`` is copied into the graph's static input; the returned tensors are the graph's
        static outputs (overwritten by the next call).  ``packed=True`` appends the packed record tensor
        ``[B, max*6 + 1]`` of ``distributed.pack_detections`` (built inside the graph), which is what the multi-GPU gather
        sends.  ``static_input=True``: the caller promises to reuse this very tensor (same memory) for later batches --
        a serving loop that rotates a few device input buffers -- so the graph reads it in place and the copy into a
        static input (133 MB per 64-image batch) is skipped; one graph is kept per such buffer."""
        if static_input and not (x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.is_contiguous()):
            raise ValueError("static_input needs a contiguous float32 or uint8 CUDA tensor")
        key = (tuple(x.shape), x.device.index, bool(packed), x.dtype, x.data_ptr() if static_input else None)
        graphs = self.__dict__.setdefault("_graphs", {})
        ent = graphs.get(key)
        if ent is None:
            static_x = x if static_input else torch.empty_like(
                x, dtype=torch.uint8 if x.dtype == torch.uint8 else torch.float32).contiguous()
            if not static_input:
                static_x.copy_(x)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):            # warm-up outside the capture: lazy function attributes, arenas
                for _ in range(2):
                    self.detections(static_x)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = self.detections(static_x, packed=packed)
            ent = graphs[key] = (g, static_x, outs)
        g, static_x, outs = ent
        if not static_input:
            static_x.copy_(x, non_blocking=True)
        g.replay()
        return outs
