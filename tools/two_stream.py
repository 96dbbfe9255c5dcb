#!/usr/bin/env python
"""Experiment: one batch of 64 as a single forward pass vs two half batches on two CUDA streams (their layer kernels
overlap at the tails).  GPU only; profiling aid."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import sys
import torch
import yolo_v3_tf2_b200 as y3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = 416
m = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
m2 = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
x = torch.rand((B, S, S, 3), device="cuda")
for parts in (1, 2, 4):
    models = [m, m2, y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0),
              y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)][:parts]
    xs = list(x.chunk(parts))
    streams = [torch.cuda.Stream() for _ in range(parts)]
    outs = [None] * parts
    for _ in range(3):
        for i in range(parts):
            with torch.cuda.stream(streams[i]):
                outs[i] = models[i](xs[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for _ in range(n):
        for i in range(parts):
            with torch.cuda.stream(streams[i]):
                models[i](xs[i], outs=outs[i])
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{parts} stream(s) x batch {B // parts}: {ms:.3f} ms per {B} images  ({B / ms * 1e3:.0f} img/s)")
