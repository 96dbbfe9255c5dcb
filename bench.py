#!/usr/bin/env python
"""Benchmark of the YOLOv3 hot path (backbone+neck+heads -> decode -> NMS) on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (TF is absent: the oracle port)

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic images.
Workload at N=1: BASELINE.json configs[1] = YOLOv3 Darknet-53, 80 classes, 416x416, batch 64, bf16 activations,
random-init weights (Keras defaults, the reference's own fresh-model state).  N > 1: weak scaling, 64 images per rank.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCLASSES = 80
GFLOP_PER_IMAGE_416 = 65.864          # SURVEY.md section 8d (2*MAC over the 75 convs, C=80)


def workload_name(size, batch):
    return (f"YOLOv3 Darknet-53 C=80 {size}x{size} batch {batch}/GPU, random-init (Keras defaults), "
            f"forward + decode + NMS(max 100, iou 0.5, score 0.1) + gather")


def conv_flops(model, H, W, batch=1):
    """(conv FLOPs per image, kernel launches of one forward pass at this batch size).  Consecutive CTA-pair layers run
    as ONE persistent launch (csrc/conv_chain.cuh), so the launch count is taken from the planner's step list."""
    from yolo_v3_tf2_b200 import _lib
    p = model.plan(H, W, batch)
    fl, ci = 0, 0
    for l, pl in zip(model.graph.layers, p["layers"]):
        if l.op == _lib.OP_CONV:
            k, cin, cout, _ = model.conv_shapes[ci]
            ci += 1
            fl += 2 * pl["H"] * pl["W"] * cout * k * k * cin
    launches = sum(1 for i, st in enumerate(p["steps"]) if st["run_first"] < 0 or st["run_first"] == i)
    return fl, launches


# DRAM traffic of the conv launches of one forward pass: ONE named ncu capture, tied to the commit it was taken at
# (tools/gpu_profile.sh + tools/summarize_profiles.py write it); not "whichever file is newest".
TRAFFIC_CAPTURE = os.path.join("profiles", "r2_traffic.json")


def conv_traffic(size, batch):
    """(bytes, provenance) from TRAFFIC_CAPTURE when it was taken on this workload, else (None, reason)."""
    path = os.path.join(ROOT, TRAFFIC_CAPTURE)
    if not os.path.exists(path):
        return None, f"{TRAFFIC_CAPTURE} not present"
    d = json.load(open(path))
    if d.get("size", 416) != size or d.get("batch", 64) != batch:
        return None, f"{TRAFFIC_CAPTURE} was captured at {d.get('size', 416)}^2 batch {d.get('batch', 64)}"
    return d["dram_bytes_read"] + d["dram_bytes_write"], f"{TRAFFIC_CAPTURE} (ncu capture at commit {d.get('commit', '?')})"


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled by an `nvidia-smi -lms` subprocess while the timed regions
    run (an in-process NVML polling thread was measured to slow kernel launches down 3x, so the sampler lives in its
    own process)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=250):
        import subprocess
        import tempfile
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def start(self):
        return self

    def stop(self):
        import statistics
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        clocks, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.out.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clocks.append(int(float(parts[0])))
                mx = int(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        return {"sm_mhz": (int(statistics.median(clocks)) if clocks else None), "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(clocks)}


def cpu_reference_step(model, anchors, x_np, torch_threads):
    """The reference's CPU path restated (oracle): forward + decode + class reduce + NMS for a small batch."""
    import torch
    from oracle import net_oracle, decode_oracle, c_oracle
    torch.set_num_threads(torch_threads)
    grids = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x_np)
    b, c, p = decode_oracle.yolo_decode(grids, anchors, NCLASSES)
    cls, sc = decode_oracle.class_reduce(c, p)
    sel, nv = c_oracle.nms(b, sc, 100, 0.5, 0.1)
    return sel, nv


def run_reference(args, rank, world):
    """--impl reference: TensorFlow 2.8.1 is not installable here (no network, Python 3.12), so the arm times the
    CPU restatement of the reference (oracle) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import numpy as np
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    cores = os.cpu_count() or 1
    model = y3.ParseModel.builtin_yolov3(NCLASSES).init_weights("keras", seed=0)
    anchors = configs.coco_anchors()
    sample_b = args.ref_batch
    x = np.random.default_rng(0).random((sample_b, args.size, args.size, 3), dtype=np.float32)
    for _ in range(args.warmup):
        cpu_reference_step(model, anchors, x, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(model, anchors, x, cores)
    dt = time.perf_counter() - t0
    v = sample_b * args.steps / dt
    line = {"impl": "reference", "metric": f"images/sec ({args.size}^2, backbone+decode+NMS)", "value": v, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.size, args.batch), "images_per_gpu": args.batch,
                       "images_per_step": sample_b,
                       "note": f"reference arm: CPU restatement of the reference (TensorFlow 2.8.1 is not installable here), "
                               f"each step is a bounded sample of {sample_b} images of that workload"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{sample_b} images of {args.size}x{args.size} per step x {args.steps} steps"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: this many images per step in total, split evenly over the ranks "
                         "(BASELINE config 3: --size 608 --global-batch 256)")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--input", default="u8", choices=["u8", "f32"],
                    help="u8: uint8 images, x/255 inside the stem conv (the serving input, inference.py:157-158); "
                         "f32: float32 images in [0,1] (the Keras model's own input)")
    ap.add_argument("--settle-seconds", type=float, default=2.0,
                    help="untimed steady load after the W warm-up steps so the SM clock has settled (reported as settle_steps)")
    ap.add_argument("--ref-batch", type=int, default=8,
                    help="images per step of the CPU reference arm (8 keeps all host cores busy; a step is ~0.5 s on 16 cores)")
    ap.add_argument("--cpu-baseline-images", type=int, default=96,
                    help="images of the bounded CPU-baseline sample (about 10 s of host work, processed 8 at a time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling mode for runs under ncu: eager launches, no CUDA graphs, only W warm-up steps "
                         "(numbers printed in this mode are not benchmark values)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs, distributed as y3dist
    from yolo_v3_tf2_b200.core.yolo_nms import nms_padded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    S = args.size
    strong = args.global_batch > 0
    if strong:
        lo, hi = y3dist.shard_range(args.global_batch, rank, world)
        B = hi - lo
    else:
        B = args.batch
    K = args.steps
    u8 = args.input == "u8"
    model = y3.ParseModel.builtin_yolov3(NCLASSES).init_weights("keras", seed=0)
    anchors = configs.coco_anchors()
    det = y3.Detector(model, anchors, NCLASSES, yolo_max_boxes=100, nms_iou_threshold=0.5, nms_score_threshold=0.1)
    flops_img, launches_fwd = conv_flops(model, S, S, B)
    mx = det.max_boxes

    # synthetic inputs: per-rank seed; a few rotating device buffers + pinned host copies
    nbuf = 3
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    if u8:
        host = [torch.randint(0, 256, (B, S, S, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    else:
        host = [torch.rand((B, S, S, 3), generator=gen, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    xs = [h.to(dev) for h in host]
    in_bytes = host[0].numel() * host[0].element_size()

    no_gather = os.environ.get("Y3_BENCH_NO_GATHER") == "1"   # diagnosis only: N > 1 without the NCCL gather

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed steps call the public serving entry point Detector.detections_graphed(): the launches of a step are
    # replayed from a CUDA graph, so the number does not depend on how fast this box's host can issue launches.
    def gstep(x):
        if args.ncu:
            local = det.detections(x, packed=(world > 1))
            if world > 1 and not no_gather:
                y3dist.gather_packed(local[4])
            return local[:4]
        # static_input: the loop rotates over a fixed set of device input buffers, one graph per buffer reads it in place
        if world > 1 and not no_gather:
            local = det.detections_graphed(x, packed=True, static_input=True)   # records packed inside the graph
            y3dist.gather_packed(local[4])                      # the only launch outside it: one NCCL all-gather
            return local[:4]
        return det.detections_graphed(x, static_input=True)

    # ---------------- warm-up: exactly the W steps asked for, then a separately reported settle phase ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    W = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps
    for i in range(W):
        gstep(xs[i % nbuf])
    sync_all()
    # settle: a fresh process needs ~2 s of steady load before the SM clock / power state stops moving (the first timed
    # loop measured 10-30 % low otherwise).  Not counted as warm-up steps; reported as settle_steps.
    settle = 0
    if not args.ncu and args.settle_seconds > 0:
        t0 = time.perf_counter()
        for i in range(10):
            gstep(xs[i % nbuf])
        sync_all()
        dt10 = max(time.perf_counter() - t0, 1e-4)
        settle = 10 + min(4000, int(args.settle_seconds / dt10 * 10))
        if world > 1:   # every rank must issue the same number of collectives
            tw = torch.tensor([settle], dtype=torch.int64, device=dev)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            settle = int(tw.item())
        for i in range(settle - 10):
            gstep(xs[i % nbuf])
        sync_all()

    # ---------------- device-resident timing (value) ----------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        gstep(xs[i % nbuf])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max / 1e3)

    # ---------------- roofline: forward pass / decode / NMS, timed INTERLEAVED with the step ----------------
    # Every iteration replays the whole step and then, between their own event pairs, the forward pass alone, the
    # decode alone and the NMS alone (each its own small CUDA graph on this step's data).  All four are measured under the
    # same clocks and power state, so forward_ms <= step_ms holds by construction and nothing is subtracted.
    roof = {}
    if not args.ncu:
        fx = xs[0].clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fouts = model(fx, padded=True)
            dense = model(fx)
            dec_c = y3.yolo_decode(fouts, anchors, NCLASSES, compact=True)
            dec_f = y3.yolo_decode(dense, anchors, NCLASSES)
            nms_padded(dec_c[0], dec_c[3], mx, 0.5, 0.1)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g_fwd, g_decc, g_decf, g_nms = (torch.cuda.CUDAGraph() for _ in range(4))
        with torch.cuda.graph(g_fwd):
            model(fx, padded=True, outs=fouts)
        with torch.cuda.graph(g_decc):
            dec_c = y3.yolo_decode(fouts, anchors, NCLASSES, compact=True)
        with torch.cuda.graph(g_decf):
            dec_f = y3.yolo_decode(dense, anchors, NCLASSES)
        with torch.cuda.graph(g_nms):
            nms_out = nms_padded(dec_c[0], dec_c[3], mx, 0.5, 0.1)
        parts = {"step": None, "forward": g_fwd, "decode_fused": g_decc, "decode_full": g_decf, "nms": g_nms}
        evs = {k: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
               for k in parts}
        for i in range(3):
            for g in (g_fwd, g_decc, g_decf, g_nms):
                g.replay()
        torch.cuda.synchronize()
        for i in range(K):
            for name, g in parts.items():
                a, b = evs[name][i]
                a.record()
                if g is None:
                    det.detections_graphed(xs[i % nbuf], packed=(world > 1 and not no_gather), static_input=True)
                else:
                    g.replay()
                b.record()
        torch.cuda.synchronize()
        roof = {k: sum(a.elapsed_time(b) for a, b in v) / K for k, v in evs.items()}
    fwd_ms = roof.get("forward", float("nan"))
    achieved_tflops = flops_img * B / (fwd_ms / 1e3) / 1e12 if roof else None

    # ---------------- end to end through the public API with host buffers (e2e) ----------------
    out_host = [torch.empty((B, mx * 6 + 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    xin = [torch.empty_like(xs[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    def e2e_loop(nsteps):
        # double-buffered serving loop: the H2D copy of step i+1 overlaps the compute of step i
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for b in range(2):
            free[b].record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[0])
            xin[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(nsteps):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[nxt])
                    xin[nxt].copy_(host[(i + 1) % nbuf], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_stream.wait_event(ready[cur])
            # the public serving call: the whole step replayed from a CUDA graph (one launch), then the NCCL gather
            if args.ncu:
                rec = det.detections(xin[cur], packed=True)[4]
            else:
                rec = det.detections_graphed(xin[cur], packed=True, static_input=True)[4]
            if world > 1 and not no_gather:
                y3dist.gather_packed(rec)
            free[cur].record(main_stream)
            out_host[cur].copy_(rec, non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(3)
    sync_all()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_loop(K)
    s1.record()
    sync_all()
    e2e_ms = s0.elapsed_time(s1)
    clocks = sampler.stop()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t.item()) / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm = float(peaks.get("hbm_gbs", 6500.0))
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.5 TB/s"
        traffic, traffic_src = conv_traffic(S, B)
        N = 3 * sum((S // s) ** 2 for s in (32, 16, 8))
        F = 5 + NCLASSES
        line = {
            "metric": f"images/sec ({S}^2, backbone+decode+NMS)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": W, "settle_steps": settle, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(S, B), "images_per_gpu": B, "images_per_step": B * world,
                       "input": ("uint8 [B,S,S,3] images, x/255 (core/load_tfrecords.py:46) inside the stem conv, bit-identical "
                                 "to feeding float32" if u8 else "float32 [B,S,S,3] in [0,1]"),
                       "l2": "inputs rotate over 3 device buffers; each step streams a ~1 GB activation arena "
                             "(>> 126 MB L2) so no step starts with a warm L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": B * (mx * 6 + 1) * 4,
                    "note": "Detector.detections_graphed() (the step replayed from a CUDA graph) on pinned host batches, "
                            "double-buffered H2D on a copy stream, detections copied back to pinned host memory every step"},
            # kernels of this library per timed step: the conv launches + decode + NMS + gather (the memset of the
            # layer-chaining counters is a driver memset node and is not counted)
            "gpu_launches": K * (launches_fwd + 3),
        }
        if roof:
            line["roofline"] = {
                "bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved_tflops / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": f"tcgen05 implicit-GEMM conv kernels (the whole forward pass: 75 conv layers in {launches_fwd} launches), "
                          "replayed from its own CUDA graph between CUDA events right after each timed-style step "
                          "(interleaved, same clocks); achieved = algorithmic conv FLOPs of one batch / that time; "
                          "algorithmic bytes = 190 MB/img",
                "forward_ms": fwd_ms, "step_ms_interleaved": roof["step"], "flops_per_step": flops_img * B}
            dec_bytes = 2 * N * F * 4 * B                      # read N*F*4, write N*(4+1+C)*4 per image
            decc_bytes = (N * F * 4 + N * (16 + 4 + 8)) * B    # fused form: boxes + scores + int64 class ids out
            nms_bytes = (N * 20 + (mx + 1) * 4) * B
            line["roofline_decode"] = {
                "bound": "hbm", "achieved": dec_bytes / (roof["decode_full"] / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": dec_bytes / (roof["decode_full"] / 1e3) / 1e9 / hbm, "ms": roof["decode_full"],
                "algorithmic_bytes": dec_bytes, "peak_source": hbm_src,
                "kernel": "decode_kernel as the reference's yolo_decode (boxes + confidence + class_probs written)"}
            line["roofline_decode_fused"] = {
                "bound": "hbm", "achieved": decc_bytes / (roof["decode_fused"] / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": decc_bytes / (roof["decode_fused"] / 1e3) / 1e9 / hbm, "ms": roof["decode_fused"],
                "algorithmic_bytes": decc_bytes,
                "kernel": "decode_kernel as run inside the step (compact: boxes + scores + class ids; probabilities are "
                          "never written because NMS does not read them)"}
            line["roofline_nms"] = {
                "bound": "latency", "achieved": nms_bytes / (roof["nms"] / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": nms_bytes / (roof["nms"] / 1e3) / 1e9 / hbm, "ms": roof["nms"], "algorithmic_bytes": nms_bytes,
                "images_per_s": B / (roof["nms"] / 1e3),
                "kernel": "nms_kernel (one CTA per image; sort + IoU bitmask chunks: latency-bound, 212 940 B/img)"}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            nimg = args.cpu_baseline_images
            xc = torch.cat([host[i % nbuf] for i in range((nimg + B - 1) // B)])[:nimg]
            xc = (xc.float() / 255.0).numpy() if u8 else xc.numpy()
            cpu_reference_step(model, anchors, xc[:1], cores)   # warm-up
            t0 = time.perf_counter()
            for i0 in range(0, nimg, 8):                       # 8 images at a time bounds the oracle's memory
                cpu_reference_step(model, anchors, xc[i0:i0 + 8], cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nimg / dt, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{nimg} images of {S}x{S}, torch-CPU fp32 oracle forward + numpy decode "
                                              f"+ C NMS (TensorFlow absent: CPU restatement of the reference)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
