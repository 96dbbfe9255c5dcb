#!/usr/bin/env python
"""NMS / decode micro-benchmark on the bench workload (Keras-default init: every box passes the score threshold) and on
sparse synthetic scores.  GPU only; profiling aid."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import numpy as np, torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import configs
from yolo_v3_tf2_b200.core.yolo_nms import nms_padded

B = 64
model = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
x = torch.rand((B, 416, 416, 3), device="cuda")
grids = model(x)
anchors = configs.coco_anchors()
bboxes, conf, probs, scores, cls = y3.yolo_decode(grids, anchors, 80, with_scores=True)
torch.cuda.synchronize()
print("scores: min %.3f max %.3f" % (scores.min().item(), scores.max().item()))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("decode (fused scores): %.1f us" % timeit(lambda: y3.yolo_decode(grids, anchors, 80, with_scores=True)))
for thr in (0.1, 0.28, 0.30, 0.32, 0.34):
    npass = (scores > thr).sum(1).float().mean().item()
    sel, nv, st = nms_padded(bboxes, scores, 100, 0.5, thr)
    print("nms thr %.2f: %.1f us   (mean passing %.0f, mean valid %.1f)" % (
        thr, timeit(lambda: nms_padded(bboxes, scores, 100, 0.5, thr)), npass, nv.float().mean().item()))
rng = np.random.default_rng(0)
s2 = torch.from_numpy((rng.random((B, 10647)) ** 8).astype(np.float32)).cuda()
for thr in (0.004, 0.1, 0.5):
    npass = (s2 > thr).sum(1).float().mean().item()
    print("sparse scores thr %.3f: %.1f us (mean passing %.0f)" % (thr, timeit(lambda: nms_padded(bboxes, s2, 100, 0.5, thr)), npass))
