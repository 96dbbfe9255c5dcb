#!/usr/bin/env python
"""Measurements for the SURVEY section 8 'next' rows on a B200 (GPU only): GPU pre-processing (f-2), YOLOv3-tiny incl. its
maxpool kernel (f-3), evaluation counters (f-4).  Prints a markdown table (committed under profiles/)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import json
import numpy as np
import torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import configs, _lib
from yolo_v3_tf2_b200.evaluate_detections import EvaluateDetections

try:
    HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    HBM = 6545.3


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
B = 64
# ---- f-2: pre-processing.  uint8 frames -> 416x416 float32 / 255 (core/load_tfrecords.py:46) and resize_image (core/utils.py:17-28)
for (h, w) in [(480, 640), (1080, 1920)]:
    frames = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(B)]
    out = torch.empty((B, 416, 416, 3), device="cuda")
    ms = timeit(lambda: y3.preprocess_images(frames, 416, 416, divide_by_255=True, out=out))
    # algorithmic bytes: output written once; each output pixel reads at most 4 source pixels (12 bytes), capped by the source size
    byt = B * (416 * 416 * 3 * 4 + min(416 * 416 * 4 * 3, h * w * 3))
    rows.append((f"pre-processing {B} x {h}x{w} uint8 -> 416x416 f32 (/255)", ms * 1e3, byt / ms / 1e6, f"{B / ms * 1e3:.0f} img/s"))
    stacked = torch.stack(frames)   # one [B, h, w, 3] tensor: the vectorised-descriptor path of preprocess_images
    ms = timeit(lambda: y3.preprocess_images(stacked, 416, 416, divide_by_255=True, out=out))
    rows.append((f"  ... same, frames as ONE [B,h,w,3] tensor", ms * 1e3, byt / ms / 1e6, f"{B / ms * 1e3:.0f} img/s"))
    ms = timeit(lambda: y3.preprocess_images(frames, 416, 416, preserve_aspect_ratio=True, out=out))
    rows.append((f"resize_image (aspect + pad) {B} x {h}x{w} -> 416x416", ms * 1e3, byt / ms / 1e6, f"{B / ms * 1e3:.0f} img/s"))
    # the kernel alone (descriptors prebuilt on the device): what the HBM roofline applies to
    sy, sx = int((np.float32(h) / np.float32(416)).view(np.int32)), int((np.float32(w) / np.float32(416)).view(np.int32))
    desc = torch.tensor([[f.data_ptr(), h, w, 0, 416, 416, 0, 0, sy, sx] for f in frames], dtype=torch.int64).cuda()
    ctx = _lib.context()
    ms = timeit(lambda: _lib.check(_lib.lib().y3_preprocess(ctx.handle, _lib.ptr(desc), B, 416, 416, 1, _lib.ptr(out), _lib.stream_ptr())))
    rows.append((f"  ... preprocess_kernel alone", ms * 1e3, byt / ms / 1e6, ""))
# ---- f-3: yolov3-tiny
tiny = y3.ParseModel.builtin_yolov3_tiny(80).init_weights("keras", seed=0)
x = torch.rand((B, 416, 416, 3), device="cuda")
outs = tiny(x, padded=True)
ms = timeit(lambda: tiny(x, padded=True, outs=outs))
fl = 0
p = tiny.plan(416, 416, 1)
ci = 0
from yolo_v3_tf2_b200 import _lib
for l, pl in zip(tiny.graph.layers, p["layers"]):
    if l.op == _lib.OP_CONV:
        k, cin, cout, _ = tiny.conv_shapes[ci]
        ci += 1
        fl += 2 * pl["H"] * pl["W"] * cout * k * k * cin
rows.append((f"YOLOv3-tiny forward, batch {B}, 416x416 ({fl / 1e9:.2f} GFLOP/img, 13 convs + 6 maxpools)", ms * 1e3,
             None, f"{B / ms * 1e3:.0f} img/s, {fl * B / ms / 1e9:.0f} TFLOP/s"))
det = y3.Detector(tiny, configs.tiny_anchors(), 80)
det.detections_graphed(x)
ms = timeit(lambda: det.detections_graphed(x))
rows.append((f"YOLOv3-tiny forward + decode + NMS + gather (graph replay), batch {B}", ms * 1e3, None, f"{B / ms * 1e3:.0f} img/s"))
per = [(l, t) for l, t in tiny.profile_layers(x)]
mp = [(l, t) for l, t in per if tiny.graph.layers[l].op == _lib.OP_MAXPOOL]
for l, t in mp[:2]:
    pl = p["layers"][l]
    s = tiny.graph.layers[l].stride
    cp = max(32, pl["C"])
    byt = B * (pl["H"] * s * pl["W"] * s + pl["H"] * pl["W"]) * cp * 2
    rows.append((f"maxpool 2x2/{s} -> {pl['H']}x{pl['W']}x{pl['C']} (layer {l})", t * 1e3, byt / t / 1e6, ""))
# ---- f-4: evaluation counters
rng = np.random.default_rng(0)
ob, oc, os_, nv = y3.Detector(y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0), configs.coco_anchors(), 80).detections(x)
gt = torch.rand((B, 20, 2), device="cuda")
gtb = torch.cat([gt, gt + 0.1], -1)
gtc = torch.randint(0, 80, (B, 20), dtype=torch.int32, device="cuda")
gtn = torch.full((B,), 20, dtype=torch.int32, device="cuda")
ev = EvaluateDetections(80, 0.5)
ms = timeit(lambda: ev.evaluate_batch(ob, oc, nv, gtb, gtc, gtn))
rows.append((f"evaluation counters, {B} images x 100 detections x 20 ground-truth boxes", ms * 1e3, None, ""))

print("| what | us | GB/s (algorithmic) | of measured HBM peak | note |\n|---|---|---|---|---|")
for name, us, gbs, note in rows:
    print(f"| {name} | {us:.1f} | {'' if gbs is None else f'{gbs:.0f}'} | {'' if gbs is None else f'{gbs / HBM:.2f}'} | {note} |")
