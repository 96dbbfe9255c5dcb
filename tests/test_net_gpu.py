"""Whole-network parity: ParseModel-built YOLOv3 on the GPU vs the torch-CPU oracle of the same graph and weights.

Stated bf16 tolerance (activations are stored in bf16 between the 75 convs, accumulation is fp32): per head,
relative L2 error of the logits <= 2e-2 and max-abs error <= 6e-2 * max|ref| (+0.02); decoded boxes of the GPU logits within
2e-2 (image-fraction units) of the oracle's for boxes whose w/h logits are moderate.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(grids, ref, rel_tol=2e-2):
    for k, (g, r) in enumerate(zip(grids, ref)):
        g = g.cpu().numpy()
        assert g.shape == r.shape
        rel = np.linalg.norm(g - r) / np.linalg.norm(r)
        mx = np.abs(g - r).max()
        assert rel <= rel_tol, f"head {k}: relative L2 error {rel}"
        assert mx <= 6e-2 * np.abs(r).max() + 0.02, f"head {k}: max abs error {mx} (max |ref| {np.abs(r).max()})"


@pytest.mark.parametrize("init,size,B,C", [("variance", 64, 2, 80), ("keras", 96, 1, 80), ("variance", 128, 3, 38),
                                           ("variance", 416, 1, 80), ("variance", 96, 2, 37),
                                           ("variance", 608, 1, 80), ("keras", 416, 1, 37)])
def test_forward_vs_oracle(cuda, init, size, B, C):
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle
    model = y3.ParseModel.builtin_yolov3(C).init_weights(init, seed=3)
    rng = np.random.default_rng(0)
    x = rng.random((B, size, size, 3), dtype=np.float32)
    grids = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    from yolo_v3_tf2_b200 import _lib
    assert _lib.context().watchdog_code() == 0
    ref = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x)
    assert [tuple(g.shape) for g in grids] == [(B, size // s, size // s, 3, 5 + C) for s in (32, 16, 8)]
    _compare(grids, ref)


@pytest.mark.parametrize("init,size,B,C", [("variance", 64, 2, 80), ("keras", 96, 3, 80), ("variance", 416, 2, 80),
                                           ("variance", 160, 1, 37)])
def test_tiny_forward_vs_oracle(cuda, init, size, B, C):
    """YOLOv3-tiny (config/models/yolov3_tiny): maxpool kernel incl. the stride-1 'same' pool, the 16-filter stem stored
    as 32 channels, 2 heads."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle
    model = y3.ParseModel.builtin_yolov3_tiny(C).init_weights(init, seed=11)
    x = np.random.default_rng(1).random((B, size, size, 3), dtype=np.float32)
    grids = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    ref = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x)
    assert [tuple(g.shape) for g in grids] == [(B, size // s, size // s, 3, 5 + C) for s in (32, 16)]
    _compare(grids, ref)


def test_maxpool_variants_vs_oracle(cuda):
    """'valid' pooling, 3x3 stride-2 'same' pooling and odd spatial sizes through a small graph."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle

    def conv(f, size=3, bn=True, act="leaky"):
        d = {"type": "convolutional", "filters": f, "size": size, "stride": 1, "pad": 1, "activation": act}
        if bn:
            d["batch_normalize"] = 1
        return d

    def pool(size, stride, padding):
        return {"type": "maxpool", "size_xy": [size, size], "stride_xy": [stride, stride], "padding": padding}

    layers = [{"type": "route", "source": {"inputs": [0]}}, conv(32), pool(3, 2, "same"), conv(64), pool(2, 2, "valid"),
              conv(64), pool(2, 1, "same"), pool(3, 1, "valid"), conv(64), conv("3*(2+2+1+nclasses)", 1, bn=False, act="linear"),
              {"type": "yolo", "grid_size": 0}]
    subs = [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}]
    model = y3.ParseModel().build_model(None, subs, "head", nclasses=3, layer_lists={"a": layers})
    model.init_weights("variance", seed=2)
    x = np.random.default_rng(4).random((2, 96, 96, 3), dtype=np.float32)   # 96 -> 48 -> 24 -> 24 -> 22
    grids = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    ref = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x)
    assert tuple(grids[0].shape) == (2, 22, 22, 3, 8)
    _compare(grids, ref)


@pytest.mark.parametrize("size,B,C", [(96, 2, 80), (416, 1, 80)])
def test_thin_heads_forward_vs_oracle(cuda, size, B, C):
    """The wiring of the reference's model_thin_heads.yaml (multi-output necks, negative entry_index,
    backbone tapped before two shortcut adds, whose Adds then run as separate kernels) against the oracle of the same
    graph."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle
    model = y3.ParseModel.builtin_yolov3(C, thin_heads=True).init_weights("variance", seed=9)
    x = np.random.default_rng(1).random((B, size, size, 3), dtype=np.float32)
    grids = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    ref = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x)
    assert [tuple(g.shape) for g in grids] == [(B, size // s, size // s, 3, 5 + C) for s in (32, 16, 8)]
    _compare(grids, ref)


def test_forward_batch_invariance_and_rebatch(cuda):
    """Images are independent: a batch of 5 gives the same rows as 5 single-image calls (bit-exact), and growing the
    batch re-plans the arena."""
    import torch
    import yolo_v3_tf2_b200 as y3
    model = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=5)
    x = torch.rand((5, 96, 96, 3), device="cuda")
    one = [model(x[i:i + 1]) for i in range(5)]
    allb = model(x)
    for k in range(3):
        assert torch.equal(allb[k], torch.cat([o[k] for o in one], 0))


@pytest.mark.parametrize("size,B,C", [(416, 64, 80), (608, 32, 80), (416, 128, 37)])
def test_full_size_batches_are_image_independent(cuda, size, B, C):
    """BASELINE.json configs 2 / 3 (the 8-GPU shard) / 5 at their full per-GPU sizes, where the CPU oracle would take
    minutes: the size-independent property is that images are independent -- every probed row of the full batch is
    bit-identical to the same image run alone (which test_forward_vs_oracle ties to the oracle at B = 1), through the
    forward pass and through decode + NMS + gather."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(C).init_weights("variance", seed=17)
    x = torch.rand((B, size, size, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(size + B))
    det = y3.Detector(model, configs.coco_anchors(), C)
    grids = [g.clone() for g in model(x)]
    full = [t.clone() for t in det.detections(x)]
    assert [tuple(g.shape) for g in grids] == [(B, size // s, size // s, 3, 5 + C) for s in (32, 16, 8)]
    for i in (0, B // 2 - 1, B - 1):
        one = model(x[i:i + 1].contiguous())
        for k in range(3):
            assert torch.equal(one[k][0], grids[k][i]), (i, k)
        d1 = det.detections(x[i:i + 1].contiguous())
        for a, b in zip(d1, full):
            assert torch.equal(a[0], b[i]), i
    from yolo_v3_tf2_b200 import _lib
    assert _lib.context().watchdog_code() == 0


@pytest.mark.parametrize("C", [80, 38, 3])
def test_pitched_head_outputs_equal_dense(cuda, C):
    """``model(x, padded=True)`` (fp32 TMA-store epilogue in the head convs, pixel pitch rounded up to 4 floats) holds
    exactly the numbers of ``model(x)``, and ``yolo_decode`` of the pitched grids equals the decode of the dense ones."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(C).init_weights("variance", seed=13)
    x = torch.rand((3, 128, 128, 3), device="cuda")
    dense = model(x)
    pitched = model(x, padded=True)
    F = 5 + C
    for d, p in zip(dense, pitched):
        assert p.shape[:3] == d.shape[:3] and p.shape[3] == (3 * F + 3) // 4 * 4
        assert torch.equal(p[..., :3 * F].reshape(d.shape), d)
    a = y3.yolo_decode(dense, configs.coco_anchors(), C, with_scores=True)
    b = y3.yolo_decode(pitched, configs.coco_anchors(), C, with_scores=True)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_detector_end_to_end(cuda):
    """model -> decode -> NMS (inference.py:109-117): fused and reference-sequence paths agree bit for bit, and NMS on
    the GPU's own decoded tensors equals the oracle NMS on those same tensors."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    from oracle import c_oracle
    model = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=7)
    x = torch.rand((3, 160, 160, 3), device="cuda")
    anchors = configs.coco_anchors()
    fused = y3.Detector(model, anchors, 80, fused=True).detect(x)
    plain = y3.Detector(model, anchors, 80, fused=False).detect(x)
    for a, b in zip(fused, plain):
        assert torch.equal(a, b)
    rsel, rnv = c_oracle.nms(fused[0].cpu().numpy(), fused[2].cpu().numpy(), 100, 0.5, 0.1)
    assert np.array_equal(fused[3].cpu().numpy(), rsel) and np.array_equal(fused[4].cpu().numpy(), rnv)
    ob, oc, os_, nv = y3.Detector(model, anchors, 80).detections(x)
    for b in range(3):
        n = int(nv[b])
        assert torch.equal(ob[b, :n], fused[0][b][fused[3][b, :n].long()])
        assert (ob[b, n:] == 0).all() and (os_[b, n:] == 0).all()


def test_graphed_detections_equal_eager(cuda):
    """The CUDA-graph replay of a step (75 PDL-chained cluster launches + decode + NMS + gather) returns exactly what
    the eager call returns, for successive different inputs."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=21)
    det = y3.Detector(model, configs.coco_anchors(), 80)
    for seed in range(3):
        x = torch.rand((2, 96, 96, 3), device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed))
        eager = [t.clone() for t in det.detections(x)]
        graphed = det.detections_graphed(x)
        torch.cuda.synchronize()
        for a, b in zip(eager, graphed):
            assert torch.equal(a, b)
    # static_input: the graph reads the caller's buffer in place; refilling that buffer changes the next replay
    buf = torch.rand((2, 96, 96, 3), device="cuda")
    first = [t.clone() for t in det.detections_graphed(buf, static_input=True)]
    for a, b in zip(first, det.detections(buf.clone())):
        assert torch.equal(a, b)
    buf.copy_(torch.rand((2, 96, 96, 3), device="cuda"))
    second = det.detections_graphed(buf, static_input=True)
    for a, b in zip(second, det.detections(buf.clone())):
        assert torch.equal(a, b)


def test_reference_yaml_and_darknet_weights_roundtrip(cuda, tmp_path):
    """Weights in the reference's layouts: Darknet .weights (convert.py:36-74) and Keras set_weights order."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import weights as wm
    m1 = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=9)
    p = tmp_path / "yolov3.weights"
    wm.write_darknet_weights(str(p), m1._params)
    m2 = y3.ParseModel.builtin_yolov3(80)
    m2.load_weights(str(p)).expect_partial()
    m3 = y3.ParseModel.builtin_yolov3(80)
    m3.set_weights(m1.get_weights())
    x = torch.rand((1, 64, 64, 3), device="cuda")
    a, b, c = m1(x), m2(x), m3(x)
    for k in range(3):
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k])


@pytest.mark.parametrize("tiny", [False, True])
def test_uint8_input_equals_float_input(cuda, tiny):
    """uint8 serving input (reference inference.py:157-158 / core/load_tfrecords.py:46: resize(...) / 255): the stem
    conv divides by 255 itself through a 256-entry table of (bf16 hi, bf16 lo) pairs; the logits are bit-identical to
    feeding float32(x) / 255 -- for YOLOv3 (stride-1 tensor-core stem) and YOLOv3-tiny (16-filter stem stored as 32
    channels, same kernel), dense and pitched outputs, and through the whole detector."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    from oracle import preprocess_oracle
    model = (y3.ParseModel.builtin_yolov3_tiny(80) if tiny else y3.ParseModel.builtin_yolov3(80)).init_weights("variance", seed=5)
    g = torch.Generator(device="cpu").manual_seed(7)
    xu = torch.randint(0, 256, (3, 96, 128, 3), dtype=torch.uint8, generator=g)
    xu[0, :4] = 0
    xu[1, -3:] = 255
    # the float tensor the reference's pipeline would feed: numpy float32 division, as in the pre-processing oracle
    xf = torch.from_numpy(xu.numpy().astype(np.float32) / np.float32(255))
    a = model(xu.cuda())
    b = model(xf.cuda())
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    ap = model(xu.cuda(), padded=True)
    bp = model(xf.cuda(), padded=True)
    for u, v in zip(ap, bp):
        assert torch.equal(u, v)
    # numpy uint8 goes the same way, and predict() keeps the dtype
    c = model(xu.numpy())
    for u, v in zip(a, c):
        assert torch.equal(u, v)
    anchors = configs.coco_anchors() if not tiny else configs.coco_anchors()[:2]
    det = y3.Detector(model, anchors, 80, nms_score_threshold=0.05)
    du = det.detections(xu.cuda())
    df = det.detections(xf.cuda())
    for u, v in zip(du, df):
        assert torch.equal(u, v)
    gu = det.detections_graphed(xu.cuda())
    torch.cuda.synchronize()
    for u, v in zip(du, gu):
        assert torch.equal(u, v)


def test_net_rebuild_keeps_captured_graphs_valid(cuda):
    """growing the batch re-plans the net; the old net (arena, weights, tensor maps) must stay alive because a
    CUDA graph captured on it still replays."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=2)
    det = y3.Detector(model, configs.coco_anchors(), 80, nms_score_threshold=0.05)
    x2 = torch.rand((2, 96, 96, 3), device="cuda")
    want2 = [t.clone() for t in det.detections(x2)]
    g2 = [t.clone() for t in det.detections_graphed(x2)]
    x5 = torch.rand((5, 96, 96, 3), device="cuda")
    want5 = [t.clone() for t in det.detections(x5)]           # batch 5 > 2: the net is rebuilt
    junk = [torch.full((64 << 20,), 7, dtype=torch.uint8, device="cuda") for _ in range(4)]   # reuse freed memory, if any
    again2 = det.detections_graphed(x2)                        # replays the graph captured on the first net
    torch.cuda.synchronize()
    for a, b, c in zip(want2, g2, again2):
        assert torch.equal(a, b) and torch.equal(a, c)
    for a, b in zip(want5, det.detections_graphed(x5)):
        assert torch.equal(a, b)
    del junk
    # new weights reach every net, including the retired one the batch-2 graph replays on
    model.init_weights("variance", seed=3)
    fresh = [t.clone() for t in det.detections(x2)]
    assert not torch.equal(fresh[0], want2[0])
    for a, b in zip(fresh, det.detections_graphed(x2)):
        assert torch.equal(a, b)
