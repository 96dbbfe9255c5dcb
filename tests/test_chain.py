"""Layer chaining / persistent multi-layer launches (csrc/conv_chain.cuh, ChainArgs in csrc/conv_tc.cuh).

CPU part: the launch plan (which steps form a run, who waits on whom, tile orders) from a planning-only context.
GPU part: the persistent launches give bit-identical logits to one launch per layer, run after run, at sizes from one
tile per layer to the full 416^2 x 64 batch -- a consumer that read a tile before its producer had finished it, or a
buffer recycled under a layer still reading it, would show up as a difference."""
import numpy as np
import pytest

import yolo_v3_tf2_b200 as y3


def _plans(model, sizes):
    return {k: model.plan(k[0], k[0], k[1]) for k in sizes}


@pytest.mark.parametrize("builder", ["yolov3", "tiny", "thin"])
def test_chain_plan_is_consistent(built, builder):
    model = {"yolov3": lambda: y3.ParseModel.builtin_yolov3(80),
             "tiny": lambda: y3.ParseModel.builtin_yolov3_tiny(80),
             "thin": lambda: y3.ParseModel.builtin_yolov3(80, thin_heads=True)}[builder]()
    for (size, batch), plan in _plans(model, [(416, 64), (416, 1), (64, 2), (608, 32), (96, 3)]).items():
        steps = plan["steps"]
        layers = plan["layers"]
        in_run = 0
        for i, s in enumerate(steps):
            if s["run_len"] == 0:
                assert s["run_first"] == -1 and s["chained"] == 0
                continue
            in_run += 1
            first, n = s["run_first"], s["run_len"]
            assert n >= 2 and first <= i < first + n
            # a run is contiguous and all its members agree on it
            for k in range(first, first + n):
                assert steps[k]["run_first"] == first and steps[k]["run_len"] == n and steps[k]["ctas"] == s["ctas"]
            # producers it waits on are earlier members of the SAME run (older ones are complete when the launch starts)
            for dep in (s["dep_step"], s["res_step"]):
                assert dep == -1 or first <= dep < i
            assert s["posts"] == 1 and s["chained"] == 1
            assert 0 <= s["rot"] < max(1, s["tiles"]) and s["rev"] in (0, 1)
            assert 0 <= s["vshift"] < s["ctas"] <= 74
            assert s["rows_per_group"] == 256 and s["tiles"] % s["tiles_n"] == 0
            # only bf16 tensor-core conv layers with a dense output join a run (kernel 1, no fused upsample)
            assert layers[s["layer"]]["kernel"] == 1 and layers[s["layer"]]["fused_upsample"] == 0
        if builder == "yolov3" and size >= 96:
            assert in_run >= 55, f"{in_run} of {len(steps)} launches in runs at {size}/{batch}"


def test_chain_plan_yolov3_runs(built):
    """Darknet-53 at the bench size: the backbone from the first CTA-pair layer to head0's 3x3 is ONE launch."""
    plan = y3.ParseModel.builtin_yolov3(80).plan(416, 416, 64)
    runs = sorted({(s["run_first"], s["run_len"]) for s in plan["steps"] if s["run_len"]})
    assert len(runs) == 3 and runs[0][1] >= 49 and runs[1][1] == 6 and runs[2][1] == 6
    # the two route layers (inputs written by two launches) start a run, never sit inside one
    for first, _ in runs[1:]:
        assert plan["steps"][first]["dep_step"] == -1
    # buffers stay alive two launches past their last reader: the arena grows, but not by much
    assert 1.0e9 < plan["arena_bytes"] < 2.2e9


@pytest.mark.gpu
@pytest.mark.parametrize("size,B,C", [(64, 1, 80), (96, 3, 80), (224, 5, 38), (416, 8, 80), (608, 4, 37), (416, 64, 80)])
def test_runs_equal_per_layer_launches(cuda, size, B, C):
    import torch
    from yolo_v3_tf2_b200 import _lib
    lib = _lib.lib()
    model = y3.ParseModel.builtin_yolov3(C).init_weights("variance", seed=3)
    g = torch.Generator(device="cuda").manual_seed(size + B)
    x = (torch.rand((B, size, size, 3), device="cuda", generator=g) * 255).to(torch.uint8)
    try:
        lib.y3_dbg_set_chain_runs(0)
        ref = [o.clone() for o in model(x)]
        ref_p = [o.clone() for o in model(x, padded=True)]
        lib.y3_dbg_set_chain_runs(1)
        reps = 12 if size * size * B < 416 * 416 * 16 else 5
        for r in range(reps):
            outs = model(x)
            assert all(torch.equal(a, b) for a, b in zip(outs, ref)), f"run {r}: logits differ from per-layer launches"
            if r % 2 == 0:
                model(x[: max(1, B // 2)])   # another batch size through the same arena and flags in between
            outs_p = model(x, padded=True)
            assert all(torch.equal(a, b) for a, b in zip(outs_p, ref_p)), f"run {r}: pitched logits differ"
        assert all(bool(torch.isfinite(o).all()) for o in ref)
    finally:
        lib.y3_dbg_set_chain_runs(1)
    torch.cuda.synchronize()
    assert _lib.context().watchdog_code() == 0


@pytest.mark.gpu
def test_runs_inside_a_cuda_graph(cuda):
    """Detector.detections_graphed replays the persistent launches from a CUDA graph (flags reset by a memset node)."""
    import torch
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=5)
    det = y3.Detector(model, configs.coco_anchors(), 80, nms_score_threshold=1e-4)
    x = (torch.rand((6, 160, 160, 3), device="cuda") * 255).to(torch.uint8)
    eager = [t.clone() for t in det.detections(x)]
    for _ in range(4):
        outs = det.detections_graphed(x)
        torch.cuda.synchronize()
        for a, b in zip(outs, eager):
            assert torch.equal(a, b)
