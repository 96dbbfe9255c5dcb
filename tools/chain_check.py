#!/usr/bin/env python
"""Layer chaining check (GPU only): the forward pass must give bit-identical logits on every run -- a consumer that read a
tile before its producer finished it would show up as a run-to-run difference -- and the same bits with the flags off.
Prints one JSON line: output hashes, mismatching runs, eager and CUDA-graph forward times.

usage: python tools/chain_check.py [batch] [size] [reps] [classes]
       Y3_PROF_LIB=1 Y3_CHAIN=0 python tools/chain_check.py ...      (same sources, every layer waits for its predecessor)
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import hashlib
import json
import sys
import torch
import yolo_v3_tf2_b200 as y3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 30
NC = int(sys.argv[4]) if len(sys.argv) > 4 else 80

m = y3.ParseModel.builtin_yolov3(NC).init_weights("variance", seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((B, S, S, 3), device="cuda", generator=g)
x8 = (x * 255).to(torch.uint8)


def digest(ts):
    h = hashlib.sha256()
    for t in ts:
        h.update(t.cpu().numpy().tobytes())
    return h.hexdigest()[:16]


res = {"batch": B, "size": S, "classes": NC, "reps": REPS,
       "lib": "prof" if _os.environ.get("Y3_PROF_LIB") == "1" else "release",
       "env": {k: v for k, v in _os.environ.items() if k.startswith("Y3_")}}
for name, inp, kw in (("f32", x, {}), ("u8_padded", x8, {"padded": True})):
    ref = [o.clone() for o in m(inp, **kw)]
    torch.cuda.synchronize()
    bad = 0
    outs = None
    for r in range(REPS):
        outs = m(inp, outs=outs, **kw)
        torch.cuda.synchronize()
        if not all(torch.equal(a, b) for a, b in zip(outs, ref)):
            bad += 1
        if r % 3 == 0:   # something else in between: another batch size through the same net (same arena, other tiling)
            m(inp[: max(1, B // 3)], **kw)
    res[name] = {"hash": digest(ref), "mismatching_runs": bad, "finite": all(bool(torch.isfinite(o).all()) for o in ref)}

# timing: eager back-to-back launches, and the same pass replayed from a CUDA graph
outs = m(x8, padded=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(10):
    m(x8, outs=outs, padded=True)
e0.record()
for _ in range(50):
    m(x8, outs=outs, padded=True)
e1.record()
torch.cuda.synchronize()
res["eager_ms"] = e0.elapsed_time(e1) / 50
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    m(x8, outs=outs, padded=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        m(x8, outs=outs, padded=True)
    for _ in range(200):
        graph.replay()
    e0.record(st)
    for _ in range(100):
        graph.replay()
    e1.record(st)
torch.cuda.synchronize()
res["graph_ms"] = e0.elapsed_time(e1) / 100
res["graph_hash"] = digest(outs)
res["graph_matches_eager"] = res["graph_hash"] == res["u8_padded"]["hash"]

# A/B inside this process: the same pass captured with every layer as its own launch, replayed alternately
from yolo_v3_tf2_b200 import _lib
_lib.lib().y3_dbg_set_chain_runs(0)
with torch.cuda.stream(st):
    m(x8, outs=outs, padded=True)
    graph_b = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph_b, stream=st):
        m(x8, outs=outs, padded=True)
    graph_b.replay()
torch.cuda.synchronize()
res["per_layer_graph_hash_matches"] = digest(outs) == res["u8_padded"]["hash"]
_lib.lib().y3_dbg_set_chain_runs(1)
ta, tb = [], []
with torch.cuda.stream(st):
    for rnd in range(12):
        for gr, acc in ((graph, ta), (graph_b, tb)):
            gr.replay()
            e0.record(st)
            for _ in range(25):
                gr.replay()
            e1.record(st)
            e1.synchronize()
            acc.append(e0.elapsed_time(e1) / 25)
import statistics
res["ab_runs_ms"] = statistics.median(ta)
res["ab_per_layer_ms"] = statistics.median(tb)
print(json.dumps(res))
