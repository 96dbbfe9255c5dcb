#!/usr/bin/env python
"""Does UMMA accept row-shifted descriptors into a TMA-swizzled tile? (GPU only)"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import numpy as np, torch
from yolo_v3_tf2_b200 import _lib
ctx = _lib.context(); lib = _lib.lib()
for swz in (128, 64):
    bk = swz // 2
    rows = 300
    x = torch.randn((rows, bk), device="cuda").to(torch.bfloat16)
    w = torch.randn((64, bk), device="cuda").to(torch.bfloat16)
    for mode in (0, 1):
        res = []
        for shift in (0, 1, 2, 3, 5, 7, 8, 9, 13, 16, 27, 55, 100, 171):
            out = torch.zeros((128, 64), device="cuda")
            _lib.check(lib.y3_dbg_umma_shift(ctx.handle, _lib.ptr(x), rows, _lib.ptr(w), swz, shift, mode, _lib.ptr(out), _lib.stream_ptr()))
            torch.cuda.synchronize()
            ref = x[shift:shift + 128].float() @ w.float().T
            err = (out - ref).abs().max().item()
            res.append((shift, round(err, 4)))
        print(f"swz={swz} base_off_mode={mode}:", res)
