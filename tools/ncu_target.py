#!/usr/bin/env python
"""Tiny target for ncu: two forward passes (+ decode + NMS) of the bench workload; profile the second one.
usage: python tools/ncu_target.py [batch] [size]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import sys
import torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import configs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
m = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
det = y3.Detector(m, configs.coco_anchors(), 80)
x = torch.rand((B, S, S, 3), device="cuda")
for _ in range(2):
    out = det.detections(x)
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in out])
