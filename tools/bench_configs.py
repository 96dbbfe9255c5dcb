#!/usr/bin/env python
"""Timings of BASELINE.json configs[2..4] on ONE B200 (config 3 = 608x608 batch 256,
config 4 = decode + NMS only at batch 512 with the score / IoU threshold sweep, config 5 = custom-class heads at
batch 128).  CUDA events on the launching stream, warm-up first, inputs larger than L2.  Parity of the same cases is in
tests/ (test_decode_nms_gpu.py, test_net_gpu.py); this file only measures.  Writes a markdown table to stdout."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root
import argparse
import numpy as np
import torch

import yolo_v3_tf2_b200 as y3
from yolo_v3_tf2_b200 import configs
from yolo_v3_tf2_b200.core.yolo_nms import nms_padded, yolo_nms

HBM = 6545.3   # MEASURED_PEAKS.json hbm_gbs


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3     # us


def gen_grids(B, sizes, C, obj_mean, seed):
    """SURVEY 8d config-4 distribution, generated on the device (timing only; the parity tests use the numpy twin
    tests/y3_test_util.synth_grids)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = []
    for s in sizes:
        t = torch.empty((B, s, s, 3, 5 + C), device="cuda")
        t[..., 0:2] = torch.randn((B, s, s, 3, 2), device="cuda", generator=g)
        t[..., 2:4] = torch.randn((B, s, s, 3, 2), device="cuda", generator=g).clamp_(-4, 4)
        t[..., 4] = torch.randn((B, s, s, 3), device="cuda", generator=g) * 2 + obj_mean
        t[..., 5:] = torch.randn((B, s, s, 3, C), device="cuda", generator=g) * 2 - 2
        out.append(t)
    return out


def config4(B=512, C=80):
    sizes = (13, 26, 52)
    N = 3 * sum(s * s for s in sizes)
    F = 5 + C
    anchors = configs.coco_anchors()
    print("\n## Config 4: decode + NMS only, N = %d x C = %d, batch %d (one GPU)\n" % (N, C, B))
    for name, mean in (("dense (obj ~ N(0,2))", 0.0), ("sparse (obj ~ N(-6,2))", -6.0)):
        print("| stage | us / batch | algorithmic GB/s | of measured HBM peak (%.0f GB/s) | images/s |" % HBM)
        print("|---|---|---|---|---|")
        grids = gen_grids(B, sizes, C, mean, 0)
        t = timeit(lambda: y3.yolo_decode(grids, anchors, C))
        by = 2.0 * N * F * 4 * B
        print("| `yolo_decode`, %s | %.1f | %.0f | %.2f | %.0f |" % (name, t, by / t / 1e3, by / t / 1e3 / HBM, B / t * 1e6))
        t = timeit(lambda: y3.yolo_decode(grids, anchors, C, with_scores=True))
        by2 = by + 12.0 * N * B
        print("| `yolo_decode` + fused scores / class ids, %s | %.1f | %.0f | %.2f | %.0f |" % (
            name, t, by2 / t / 1e3, by2 / t / 1e3 / HBM, B / t * 1e6))
        t = timeit(lambda: y3.yolo_decode(grids, anchors, C, compact=True))
        by4 = (N * F * 4 + 28.0 * N) * B
        print("| compact `yolo_decode` (boxes + scores + class ids only: what the fused pipeline runs), %s | %.1f | %.0f | %.2f | %.0f |" % (
            name, t, by4 / t / 1e3, by4 / t / 1e3 / HBM, B / t * 1e6))
        dec = y3.yolo_decode(grids, anchors, C, with_scores=True)
        bboxes, conf, probs, scores, cls = dec
        t = timeit(lambda: yolo_nms((bboxes, conf, probs), 100, 0.5, 0.1, check_status=False))
        t_n = timeit(lambda: nms_padded(bboxes, scores, 100, 0.5, 0.1))
        by3 = (N * C * 4 + N * 4 + N * 12.0) * B
        print("| class reduce (`yolo_nms` minus the suppression stage), %s | %.1f | %.0f | %.2f | %.0f |" % (
            name, t - t_n, by3 / (t - t_n) / 1e3, by3 / (t - t_n) / 1e3 / HBM, B / (t - t_n) * 1e6))
        del grids
        print("")
        print("NMS sweep, %s: us per %d-image batch (images/s) [mean boxes passing the score threshold / mean kept]\n" % (name, B))
        print("| score_thr | iou 0.3 | iou 0.5 | iou 0.7 |")
        print("|---|---|---|---|")
        for sthr in (0.004, 0.1, 0.2, 0.5, 0.9):
            npass = (scores > sthr).sum(1).float().mean().item()
            cells = []
            for iou in (0.3, 0.5, 0.7):
                sel, nv, st = nms_padded(bboxes, scores, 100, iou, sthr)
                assert int(st.max().item()) == 0
                t = timeit(lambda: nms_padded(bboxes, scores, 100, iou, sthr), n=10, warm=2)
                cells.append("%.0f (%.2f M img/s) [%.0f / %.1f]" % (t, B / t, npass, nv.float().mean().item()))
            print("| %.3f | %s |" % (sthr, " | ".join(cells)))
        print("")
        del dec, bboxes, conf, probs, scores, cls
        torch.cuda.empty_cache()


def full_path(title, C, H, B, steps=30):
    model = y3.ParseModel.builtin_yolov3(C).init_weights("keras", seed=0)
    anchors = configs.coco_anchors()
    det = y3.Detector(model, anchors, C)
    xs = [torch.rand((B, H, H, 3), device="cuda") for _ in range(2)]
    k = [0]

    def step():
        k[0] ^= 1
        det.detections_graphed(xs[k[0]], static_input=True)

    def fwd():
        k[0] ^= 1
        model(xs[k[0]], padded=True)
    for _ in range(10):
        step()
    t = timeit(step, n=steps, warm=10)
    tf = timeit(fwd, n=steps, warm=5)
    from bench import conv_flops
    flops = conv_flops(model, H, H)[0]
    ob, oc, os_, nv = det.detections_graphed(xs[0], static_input=True)
    torch.cuda.synchronize()
    print("| %s | %d | %.3f | %.0f | %.3f | %s | %.1f |" % (
        title, B, t / 1e3, B / t * 1e6, tf / 1e3,
        ("%.0f" % (flops * B / tf / 1e6)) if flops else "", nv.float().mean().item()))
    del det, model, xs
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    print("# BASELINE.json configs[2..4] on one B200 (`python tools/bench_configs.py`, CUDA events)")
    if a.only in ("", "full"):
        print("\n## Configs 2, 3, 5: whole path (forward + decode + NMS + gather, CUDA-graph replay, device-resident inputs)\n")
        print("| config | images / step | ms / step | images/s | forward only, ms (eager launches) | forward TFLOP/s | mean detections kept |")
        print("|---|---|---|---|---|---|---|")
        full_path("2: 416x416, C=80", 80, 416, 64)
        full_path("3: 608x608, C=80 (grids 19/38/76, N = 22 743), the whole 256-image global batch on one GPU", 80, 608, 256, steps=10)
        full_path("3: 608x608, C=80, the 8-GPU shard (32 images)", 80, 608, 32)
        full_path("5: 416x416, C=37 (126-channel heads)", 37, 416, 128, steps=20)
        full_path("5: 416x416, C=38 (129-channel heads, `len(pets_breed.names)`)", 38, 416, 128, steps=20)
    if a.only in ("", "c4"):
        config4()
