#!/usr/bin/env python
"""Per-layer device times and achieved TFLOP/s of one forward pass (profiling aid; GPU only).
usage: python tools/profile_layers.py [batch] [size]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import sys
import numpy as np
import torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
m = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
x = torch.rand((B, S, S, 3), device="cuda")
for _ in range(3):
    m(x)
torch.cuda.synchronize()
runs = [m.profile_layers(x) for _ in range(5)]
plan = m.plan(S, S, B)["layers"]
g = m.graph
conv_of_layer = {li: ci for ci, li in enumerate(g.conv_layers)}
tot_ms = tot_fl = 0.0
print(f"{'layer':>5} {'k':>2} {'s':>2} {'cin':>5} {'cout':>5} {'HxW':>9} {'BN':>4} {'ms':>8} {'TFLOP/s':>8} {'GB/s':>7}")
for i, (layer, _) in enumerate(runs[0]):
    ms = float(np.median([r[i][1] for r in runs]))
    l = g.layers[layer]
    pl = plan[layer]
    if l.op == _lib.OP_CONV:
        k, cin, cout, _ = m.conv_shapes[conv_of_layer[layer]]
        fl = 2.0 * B * pl["H"] * pl["W"] * cout * k * k * cin
        hin = pl["H"] * l.stride
        byt = B * (hin * hin * cin * (4 if cin == 3 else 2) + pl["H"] * pl["W"] * cout * 2 * (2 if pl["fused_add"] >= 0 else 1))
        print(f"{layer:5d} {k:2d} {l.stride:2d} {cin:5d} {cout:5d} {pl['H']:4d}x{pl['W']:<4d} {pl['block_n']:4d} {ms:8.4f} "
              f"{fl / ms / 1e9:8.1f} {byt / ms / 1e6:7.0f}")
        tot_fl += fl
    tot_ms += ms
print(f"total {tot_ms:.3f} ms  {tot_fl / tot_ms / 1e9:.1f} TFLOP/s  {B / tot_ms * 1e3:.0f} img/s (forward only)")

# whole forward without per-layer events (what PDL / launch overlap actually buys)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
outs = m(x)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    m(x, outs=outs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"whole forward, back-to-back launches: {ms:.3f} ms  {tot_fl / ms / 1e9:.1f} TFLOP/s  {B / ms * 1e3:.0f} img/s")
