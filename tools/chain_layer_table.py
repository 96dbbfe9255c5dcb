#!/usr/bin/env python
"""Per-layer times INSIDE the persistent multi-layer launches, from the per-CTA %globaltimer stamps of the profiling
library (same stamps as tools/net_timeline.py): a markdown table of every conv step of one forward pass with

  span    = last CTA's exit  - first CTA's entry of that layer (overlaps its neighbours inside a run)
  advance = how much later this layer's last exit is than the latest exit of all earlier layers: the time the layer
            ADDS to the pass (sums to the length of the pass), the per-layer cost once ramp / tail overlap is counted
  TFLOP/s = the layer's algorithmic FLOPs / advance

usage: Y3_PROF_LIB=1 python tools/chain_layer_table.py [batch] [size] [out.md]"""
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
OUT = sys.argv[3] if len(sys.argv) > 3 else None
m = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=1)
x = (torch.rand((B, S, S, 3), device="cuda") * 255).to(torch.uint8)
outs = m(x, padded=True)
for _ in range(5):
    m(x, outs=outs, padded=True)
torch.cuda.synchronize()
plan = m.plan(S, S, B)
steps, lplan, layers = plan["steps"], plan["layers"], m.graph.layers
nsteps = len(steps)
STRIDE = 32 * 160
lib = _lib.lib()
REPS = 5
spans, advs = np.zeros((REPS, nsteps)), np.zeros((REPS, nsteps))
total = np.zeros(REPS)
for rep in range(REPS):
    ts = torch.zeros(nsteps * STRIDE, dtype=torch.int64, device="cuda")
    lib.y3_dbg_timestamps_net(_lib.ptr(ts))
    m(x, outs=outs, padded=True)
    torch.cuda.synchronize()
    lib.y3_dbg_timestamps_net(None)
    t = ts.cpu().numpy().reshape(nsteps, 160, 32).astype(np.float64)
    latest = None
    t0 = None
    for k in range(nsteps):
        e = t[k, :, 0]
        used = e > 0
        if not used.any():
            spans[rep, k] = advs[rep, k] = np.nan
            continue
        ex = t[k, used, :][:, 7:12]        # epilogue-group / kernel exit stamps
        ex = ex[ex > 0]
        en = e[used].min()
        if t0 is None:
            t0 = en
        last = ex.max() if len(ex) else np.nan
        spans[rep, k] = (last - en) / 1e3
        advs[rep, k] = (last - (latest if latest is not None else en)) / 1e3
        latest = last if latest is None else max(latest, last)
    total[rep] = (latest - t0) / 1e3 if latest is not None else np.nan
span, adv = np.nanmedian(spans, axis=0), np.nanmedian(advs, axis=0)

rows = ["| step | layer | conv | output | run | tiles / CTA pairs | GFLOP | span us | advance us | TFLOP/s over advance |",
        "|---|---|---|---|---|---|---|---|---|---|"]
for k in range(nsteps):
    st = steps[k]
    L = layers[st["layer"]]
    lp = lplan[st["layer"]]
    cin = 3 if L.src0 <= 0 else lplan[L.src0 - 1]["C"]
    gf = 2.0 * B * lp["H"] * lp["W"] * L.filters * L.ksize * L.ksize * cin / 1e9
    run = f"{st['run_first']}+{st['run_len']}" if st["run_len"] > 0 else "single"
    if np.isnan(span[k]):
        rows.append(f"| {k} | {st['layer']} | {L.ksize}x{L.ksize}/{L.stride} {cin}->{L.filters} | {lp['H']}x{lp['W']} | {run} | - | {gf:.1f} | (no stamps: not a CTA-pair kernel) | | |")
        continue
    a = max(adv[k], 1e-3)
    rows.append(f"| {k} | {st['layer']} | {L.ksize}x{L.ksize}/{L.stride} {cin}->{L.filters}{' +Add' if lp['fused_add'] > 0 else ''} | {lp['H']}x{lp['W']} | {run} |"
                f" {st['tiles']} / {st['ctas']} | {gf:.1f} | {span[k]:.1f} | {adv[k]:.1f} | {gf / a * 1e3:.0f} |")
txt = (f"Per-layer times of one forward pass, batch {B}, {S}x{S}, uint8 input, eager launches of the profiling library with "
       f"%globaltimer stamps (median of {REPS} passes; the stamps cost ~3 % themselves).  Stamped span of the pass: "
       f"{np.nanmedian(total):.0f} us.\n\n" + "\n".join(rows) + "\n")
print(txt)
if OUT:
    open(OUT, "w").write(txt)
