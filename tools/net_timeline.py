#!/usr/bin/env python
"""Timeline of a whole forward pass from per-CTA %globaltimer stamps (y3_dbg_timestamps_net; profiling library only):
for every launch, when its CTAs enter, get their first operands, issue their last MMA and exit, relative to the first
entry of the pass.  Shows how much of a layer's ramp and tail overlaps its neighbours (layer chaining on / off).
usage: Y3_PROF_LIB=1 [Y3_CHAIN=0] python tools/net_timeline.py [batch] [size] [first_step] [last_step]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import sys
import numpy as np
import torch
import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import _lib

assert _os.environ.get("Y3_PROF_LIB") == "1", "needs the profiling library (Y3_PROF_LIB=1)"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
S0 = int(sys.argv[3]) if len(sys.argv) > 3 else 9
S1 = int(sys.argv[4]) if len(sys.argv) > 4 else 30
m = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=1)
x = (torch.rand((B, S, S, 3), device="cuda") * 255).to(torch.uint8)
outs = m(x, padded=True)
for _ in range(5):
    m(x, outs=outs, padded=True)
torch.cuda.synchronize()
plan = m.plan(S, S, B)
nsteps = len(plan["steps"])
STRIDE = 32 * 160
ts = torch.zeros(nsteps * STRIDE, dtype=torch.int64, device="cuda")
lib = _lib.lib()
lib.y3_dbg_timestamps_net(_lib.ptr(ts))
m(x, outs=outs, padded=True)
torch.cuda.synchronize()
lib.y3_dbg_timestamps_net(None)
t = ts.cpu().numpy().reshape(nsteps, 160, 32).astype(np.float64)
t0 = t[:, :, 0][t[:, :, 0] > 0].min()
print(f"{'step':>4} {'layer':>5} {'ch':>2} {'rounds':>6} | {'entry min':>9} {'med':>8} {'max':>8} | {'1st operands min':>16} {'med':>8} {'max':>8} |"
      f" {'last MMA max':>12} | {'exit min':>8} {'med':>8} {'max':>8} | {'span':>6}")
prev_exit = None
for k in range(S0, min(S1 + 1, nsteps)):
    st = plan["steps"][k]
    e = t[k, :, 0]
    used = e > 0
    if not used.any():
        print(f"{k:4d} {st['layer']:5d}   (no stamps: not the CTA-pair kernel)")
        continue
    def col(c):
        v = t[k, used, c]
        v = v[v > 0]
        return (v - t0) / 1e3 if len(v) else np.array([np.nan])
    en, op, mm, ex = col(0), col(4), col(5), col(11)
    print(f"{k:4d} {st['layer']:5d} {st['chained']:2d} {st['tiles'] / max(1, st['ctas']):6.2f} | {en.min():9.2f} {np.median(en):8.2f} {en.max():8.2f} |"
          f" {op.min():16.2f} {np.median(op):8.2f} {op.max():8.2f} | {mm.max():12.2f} | {ex.min():8.2f} {np.median(ex):8.2f} {ex.max():8.2f} |"
          f" {ex.max() - en.min():6.2f}")
