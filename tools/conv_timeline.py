#!/usr/bin/env python
"""In-kernel timeline of single conv layers (CTA-pair kernel) from per-CTA %globaltimer stamps (y3_dbg_timestamps).
Prints, per layer, the time from the earliest CTA entry to each milestone (min / median / max over CTAs), so the fixed
cost of a launch (prologue, first operand latency, tail, store drain, teardown) can be read off directly.
usage: python tools/conv_timeline.py        (GPU only; profiling aid)"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import numpy as np, torch
from yolo_v3_tf2_b200 import _lib
ctx = _lib.context()
lib = _lib.lib()
B = 64
NAMES = ["entry", "prologue done", "pdl wait done", "last load issued", "first operands", "last MMA issued",
         "first accum done", "epi g0 handoff", "epi g1 handoff", "epi g0 stores done", "epi g1 stores done", "exit"]
CLK = ["chunk1 start", "out of TMEM", "residual landed", "math+STS issued", "fenced", "store issued", "res prefetch issued",
       "chunk2 out of TMEM"]
import os
cases = [("1x1 512->256 @13", 13, 512, 256, 1, 1, False), ("1x1 256->128 @52", 52, 256, 128, 1, 1, False),
         ("1x1 512->256 @26", 26, 512, 256, 1, 1, False), ("1x1 1024->512 @13", 13, 1024, 512, 1, 1, False),
         ("3x3 128->256 @52 +res", 52, 128, 256, 3, 1, True), ("3x3 64->128 @104 +res", 104, 64, 128, 3, 1, True),
         ("3x3 512->1024 @13", 13, 512, 1024, 3, 1, False)]
if os.environ.get("Y3_TL_CASES"):
    cases = [c for i, c in enumerate(cases) if str(i) in os.environ["Y3_TL_CASES"].split(",")]
ts = torch.zeros(32 * 148, dtype=torch.int64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
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
    for mode in ("cold (L2 flushed)", "warm (same launch repeated)"):
        if mode.startswith("cold"):
            flush.zero_()
        else:
            run()
        torch.cuda.synchronize()
        ts.zero_()
        lib.y3_dbg_timestamps(_lib.ptr(ts))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        lib.y3_dbg_timestamps(None)
        t = ts.cpu().numpy().reshape(148, 32).astype(np.float64)
        import os
        os.makedirs("gpurun_out", exist_ok=True)
        np.save(f"gpurun_out/tl_{name.replace(' ', '_').replace('>', '')}_{mode.split()[0]}.npy", t)
        used = t[:, 0] > 0
        t = t[used]
        t0 = t[:, 0].min()
        print(f"== {name}  {mode}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, {used.sum()} CTAs")
        for kk, nm in enumerate(NAMES):
            col = t[:, kk]
            col = col[col > 0]
            if len(col) == 0:
                continue
            d = (col - t0) / 1e3
            print(f"   {nm:20s} min {d.min():7.2f}  med {np.median(d):7.2f}  max {d.max():7.2f} us   (n={len(col)})")
        lead = t[:, 27] > 0
        if lead.any():
            tl = t[lead]
            print(f"   MMA warp (leader CTAs, median cycles): loop {np.median(tl[:, 27]):.0f}, waiting for operands "
                  f"{np.median(tl[:, 26]):.0f}, waiting for a free accumulator {np.median(tl[:, 25]):.0f}; "
                  f"producer waiting for a free stage {np.median(t[:, 24]):.0f}")
        ep = t[:, 28] > 0
        if ep.any():
            te = t[ep]
            print(f"   epilogue warp 4 (median cycles): loop {np.median(te[:, 28]):.0f}, waiting for an accumulator "
                  f"{np.median(te[:, 29]):.0f}, for a staging slot {np.median(te[:, 30]):.0f}, for the residual {np.median(te[:, 31]):.0f}")
        ok = t[:, 16] > 0
        if ok.any():
            c = t[ok][:, 16:24]
            c = c - c[:, :1]
            print("   cycles since chunk1 start (median over CTAs): " +
                  ", ".join(f"{nm} {np.median(c[:, i][c[:, i] >= 0]) if (t[ok][:, 16 + i] > 0).any() else -1:.0f}" for i, nm in enumerate(CLK)))
