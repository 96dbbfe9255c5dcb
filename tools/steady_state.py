#!/usr/bin/env python
"""How does the step time evolve over a few seconds of continuous load (power capping)?  Prints the mean step time and
the SM clock per window of 50 steps.  GPU only; profiling aid.  usage: python tools/steady_state.py [steps] [sync_every]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import subprocess, sys, time
import torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
sync_every = int(sys.argv[2]) if len(sys.argv) > 2 else 0
model = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
det = y3.Detector(model, configs.coco_anchors(), 80)
x = torch.rand((64, 416, 416, 3), device="cuda")
for _ in range(5):
    det.detections(x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                       stdout=subprocess.PIPE, text=True)
t0 = time.perf_counter()
ev[0].record()
mode = sys.argv[3] if len(sys.argv) > 3 else "detections"
from yolo_v3_tf2_b200.core.yolo_nms import nms_padded
from yolo_v3_tf2_b200.inference import gather_detections_batched
xs = [torch.rand((64, 416, 416, 3), device="cuda") for _ in range(3)]
fe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for i in range(n):
    if mode == "detections":
        det.detections(x)
    elif mode == "graph":
        det.detections_graphed(xs[i % 3])
    else:
        xx = xs[i % 3] if "rot" in mode else x
        if "ev" in mode:
            fe[i][0].record()
        grids = model(xx, padded=True)
        if "ev" in mode:
            fe[i][1].record()
        bboxes, conf, probs, scores, cls = y3.yolo_decode(grids, configs.coco_anchors(), 80, with_scores=True)
        sel, nv, status = nms_padded(bboxes, scores, 100, 0.5, 0.1)
        gather_detections_batched(bboxes, cls, scores, sel, nv)
    ev[i + 1].record()
    if sync_every and (i + 1) % sync_every == 0:
        torch.cuda.synchronize()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
smi.terminate()
clk = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
print(f"{n} steps in {wall:.2f} s wall; sync_every={sync_every} mode={mode}")
for w in range(0, n, 50):
    ms = ev[w].elapsed_time(ev[min(w + 50, n)]) / (min(w + 50, n) - w)
    print(f"  steps {w:4d}-{min(w + 50, n):4d}: {ms:.3f} ms/step")
print("nvidia-smi (clocks.sm MHz, power W) every 200 ms:", " | ".join(clk[:40]))
