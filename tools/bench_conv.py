#!/usr/bin/env python
"""Micro-benchmark of single conv layers through the unit-test entry (GPU only).
usage: [Y3_DBG=n] python tools/bench_conv.py"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import numpy as np, torch
from yolo_v3_tf2_b200 import _lib
ctx = _lib.context()
lib = _lib.lib()
B = 64
cases = [("1x1 256->128 @52", 52, 256, 128, 1, 1, False), ("3x3 128->256 @52", 52, 128, 256, 3, 1, False),
         ("3x3 128->256 @52 +res", 52, 128, 256, 3, 1, True), ("1x1 512->256 @26", 26, 512, 256, 1, 1, False),
         ("3x3 64->128 @104 +res", 104, 64, 128, 3, 1, True), ("1x1 128->64 @104", 104, 128, 64, 1, 1, False),
         ("3x3 256->512 @26 +res", 26, 256, 512, 3, 1, True), ("3x3 512->1024 @13 +res", 13, 512, 1024, 3, 1, True)]
if _os.environ.get("Y3_BC_CASES"):
    cases = [c for i, c in enumerate(cases) if str(i) in _os.environ["Y3_BC_CASES"].split(",")]
for name, g, cin, cout, k, stride, res in cases:
    x = torch.randn((B, g, g, cin), device="cuda").to(torch.bfloat16)
    bn = lib.y3_conv_block_n(cin, cout)
    cp = ((cout + bn - 1) // bn) * bn
    w = (torch.randn((cp, k, k, cin), device="cuda") / (k * k * cin) ** 0.5).to(torch.bfloat16)
    b = torch.zeros(cp, device="cuda")
    r = torch.randn((B, g, g, cout), device="cuda").to(torch.bfloat16) if res else None
    o = torch.empty((B, g, g, cout), device="cuda", dtype=torch.bfloat16)
    def run():
        _lib.check(lib.y3_conv2d_bf16(ctx.handle, _lib.ptr(x), B, g, g, cin, cin, _lib.ptr(w), _lib.ptr(b), k, stride, cout, 1,
                                      _lib.ptr(r), cout, _lib.ptr(o), cout, 0, 0, _lib.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(10):
        flush.zero_()
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    fl = 2.0 * B * g * g * cout * k * k * cin
    print(f"{name:24s} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
    if k == 3 and stride == 1:
        # the same layer through the flat-patch kernel (haloed input)
        bk = 64 if cin % 64 == 0 else 32
        bnf = 64 if cout <= 64 else (128 if cout <= 128 else 256)
        cpf = ((cout + bnf - 1) // bnf) * bnf
        wf = (torch.randn((cpf, cin // bk, 3, 3, bk), device="cuda") / (9 * cin) ** 0.5).to(torch.bfloat16)
        bf = torch.zeros(cpf, device="cuda")
        xp = torch.zeros((B, g + 1, g + 1, cin), device="cuda", dtype=torch.bfloat16)
        xp[:, :g, :g] = x
        def runf():
            _lib.check(lib.y3_conv2d_flat_bf16(ctx.handle, _lib.ptr(xp), B, g, g, cin, _lib.ptr(wf), _lib.ptr(bf), cout, 1,
                                               _lib.ptr(r), cout, _lib.ptr(o), cout, _lib.stream_ptr()))
        for _ in range(3):
            runf()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0.record(); runf(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"{'   flat-patch kernel':24s} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
