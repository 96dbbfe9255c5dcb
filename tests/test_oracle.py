"""The oracle against its golden vectors and against itself (two NMS formulations + the C port; fp32 vs fp64 convs)."""
import os

import numpy as np
import pytest

from y3_test_util import cluster_boxes, synth_grids

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_decode_golden():
    from oracle import decode_oracle, c_oracle
    z = np.load(os.path.join(GOLD, "decode_small.npz"))
    grids = [z["g0"], z["g1"], z["g2"]]
    b, c, p = decode_oracle.yolo_decode(grids, z["anchors"], 6)
    assert np.array_equal(b, z["bboxes"]) and np.array_equal(c, z["conf"]) and np.array_equal(p, z["probs"])
    cls, sc = decode_oracle.class_reduce(c, p)
    assert np.array_equal(cls, z["cls"]) and np.array_equal(sc, z["scores"])
    # C port: same formulas, libm expf instead of numpy's -> a couple of ulp
    cb, cc, cp = c_oracle.decode(grids, z["anchors"], 6)
    np.testing.assert_allclose(cb, b, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(cp, p, rtol=2e-6, atol=2e-6)
    ccls, csc = c_oracle.class_reduce(c, p)
    assert np.array_equal(ccls, cls) and np.array_equal(csc, sc)


def test_decode_layout_properties():
    """flat index n = off_s + (i*g + j)*3 + a ; box centre inside its cell ; scales concatenated 13 -> 26 -> 52."""
    from oracle import decode_oracle
    from yolo_v3_tf2_b200 import configs
    grids = [np.zeros((1, g, g, 3, 7), np.float32) for g in (13, 26, 52)]
    b, c, p = decode_oracle.yolo_decode(grids, configs.coco_anchors(), 2)
    assert b.shape == (1, 10647, 4) and c.shape == (1, 10647, 1) and p.shape == (1, 10647, 2)
    assert np.all(c == 0.5) and np.all(p == 0.5)
    n = 507 + (5 * 26 + 7) * 3 + 1          # scale 1, row 5, col 7, anchor 1
    cx, cy = (b[0, n, 0] + b[0, n, 2]) / 2, (b[0, n, 1] + b[0, n, 3]) / 2
    assert abs(cx - 7.5 / 26) < 1e-6 and abs(cy - 5.5 / 26) < 1e-6
    w = b[0, n, 2] - b[0, n, 0]
    assert abs(w - configs.coco_anchors()[1, 1, 0]) < 1e-6


def test_nms_golden_all_formulations():
    from oracle import nms_oracle, c_oracle
    z = np.load(os.path.join(GOLD, "nms_cases.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_boxes")})
    assert len(names) >= 6
    for n in names:
        b, s = z[f"{n}_boxes"], z[f"{n}_scores"]
        mx, iou, sthr = z[f"{n}_params"]
        for f in (nms_oracle.nms_padded_tiled, nms_oracle.nms_padded_greedy):
            sel, nv = f(b, s, int(mx), float(iou), float(sthr))
            assert np.array_equal(sel, z[f"{n}_sel"]) and nv == z[f"{n}_nv"], n
        sel, nv = c_oracle.nms(b[None], s[None], int(mx), float(iou), float(sthr))
        assert np.array_equal(sel[0], z[f"{n}_sel"]) and nv[0] == z[f"{n}_nv"], n


@pytest.mark.parametrize("seed", range(4))
def test_nms_formulations_agree_random(seed):
    from oracle import nms_oracle, c_oracle
    rng = np.random.default_rng(seed)
    b, s = cluster_boxes(int(rng.integers(600, 2500)), int(rng.integers(3, 60)), seed + 100)
    s[:: int(rng.integers(3, 9))] = s[1]
    for iou in (0.3, 0.5, 0.7):
        for sthr in (0.004, 0.5):
            t = nms_oracle.nms_padded_tiled(b, s, 100, iou, sthr)
            g = nms_oracle.nms_padded_greedy(b, s, 100, iou, sthr)
            c = c_oracle.nms(b[None], s[None], 100, iou, sthr)
            assert np.array_equal(t[0], g[0]) and t[1] == g[1]
            assert np.array_equal(c[0][0], t[0]) and c[1][0] == t[1]


@pytest.mark.parametrize("seed", range(6))
def test_nms_oracle_vs_torchvision(seed):
    """Third opinion on the greedy core of the restatement: torchvision.ops.nms (an implementation written by other
    people, 'iou > thr', no epsilon, no TF-specific rules) must select the same boxes in the same order wherever the
    TF-specific rules cannot matter: distinct scores (no tie-break), every box with a positive coordinate, and no
    pair of boxes whose IoU is within 3e-5 of the threshold ('>=' vs '>', the 1e-8 in the denominator)."""
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    from oracle import nms_oracle, c_oracle
    rng = np.random.default_rng(seed)
    N = 500
    b, _ = cluster_boxes(N, 25, seed + 7)
    b = np.abs(b) + np.float32(1e-3)
    lo, hi = np.minimum(b[:, :2], b[:, 2:]), np.maximum(b[:, :2], b[:, 2:])
    hi = np.maximum(hi, lo + np.float32(0.03))                       # sides >= 0.03: the 1e-8 then shifts an IoU by < 2e-5
    b = np.concatenate([lo, hi], 1).astype(np.float32)
    s = ((rng.permutation(N) + 0.5) / N).astype(np.float32)          # distinct scores
    iou_all = tv.ops.box_iou(torch.from_numpy(b), torch.from_numpy(b)).numpy()
    assert np.abs(nms_oracle.bbox_overlap(b, b) - iou_all).max() < 2e-5   # same IoU up to rounding and the 1e-8
    checked = 0
    for thr in (0.3, 0.5, 0.7):
        for sthr in (0.1, 0.6):
            keep = s > sthr
            idx = np.nonzero(keep)[0]
            if np.any(np.abs(iou_all[np.ix_(idx, idx)] - thr) < 3e-5):
                continue                                             # a borderline pair: implementations may differ
            checked += 1
            ref = idx[tv.ops.nms(torch.from_numpy(b[idx]), torch.from_numpy(s[idx]), thr).numpy()][:100]
            for sel, nv in (nms_oracle.nms_padded_greedy(b, s, 100, thr, sthr),
                            nms_oracle.nms_padded_tiled(b, s, 100, thr, sthr),
                            tuple(x[0] for x in c_oracle.nms(b[None], s[None], 100, thr, sthr))):
                assert nv == len(ref)
                assert np.array_equal(sel[:nv], ref)
    assert checked >= 1


def test_nms_semantics_by_hand():
    """Tiny hand-checkable cases for each rule in SURVEY.md row a14."""
    from oracle import nms_oracle
    b = np.array([[0, 0, .4, .4], [0, 0, .4, .41], [.5, .5, .9, .9], [0, 0, .4, .4]], np.float32)
    s = np.array([.9, .8, .3, .9], np.float32)
    sel, nv = nms_oracle.nms_padded_greedy(b, s, 5, 0.5, 0.1)
    assert nv == 2 and sel.tolist() == [0, 2, 0, 0, 0]        # tie 0/3 -> lower index first; 1 and 3 suppressed
    sel, nv = nms_oracle.nms_padded_greedy(b, s, 5, 0.5, 0.3)  # strict '>' score filter drops score == thr
    assert nv == 1 and sel.tolist() == [0, 0, 0, 0, 0]
    sel, nv = nms_oracle.nms_padded_greedy(b, s, 1, 0.5, 0.1)  # max_output_size caps
    assert nv == 1
    # all-coords <= 0 box survives (it suppresses) but is never selected
    b2 = np.array([[-.5, -.5, -.1, -.1], [-.5, -.5, -.1, -.1001], [.1, .1, .2, .2]], np.float32)
    sel, nv = nms_oracle.nms_padded_greedy(b2, np.array([.9, .8, .7], np.float32), 3, 0.5, 0.1)
    assert nv == 1 and sel.tolist() == [2, 0, 0]
    t = nms_oracle.nms_padded_tiled(b2, np.array([.9, .8, .7], np.float32), 3, 0.5, 0.1)
    assert t[1] == 1 and t[0].tolist() == [2, 0, 0]


@pytest.mark.parametrize("init", ["variance", "keras"])
def test_net_golden(init):
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle
    z = np.load(os.path.join(GOLD, f"net64_{init}.npz"))
    m = y3.ParseModel.builtin_yolov3(80).init_weights(init, seed=int(z["seed"]))
    assert abs(m._params[0].kernel.astype(np.float64).sum() - float(z["w0_sum"])) < 1e-9      # generator is deterministic
    assert abs(m._params[74].kernel.astype(np.float64).sum() - float(z["w74_sum"])) < 1e-9
    outs = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, z["x"])
    for o, k in zip(outs, ("g0", "g1", "g2")):
        # conv summation order may differ between CPU kernels/thread counts: compare with a float32 tolerance
        np.testing.assert_allclose(o, z[k], rtol=2e-4, atol=2e-4)


def test_net_oracle_fp32_vs_fp64():
    """Bounds the oracle's own float32 noise, so the bf16 tolerance of the GPU tests is stated against truth."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import net_oracle
    m = y3.ParseModel.builtin_yolov3(80).init_weights("variance", seed=3)
    x = np.random.default_rng(1).random((1, 64, 64, 3), dtype=np.float32)
    a = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, x)
    b = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, x, dtype=torch.float64)
    for u, v in zip(a, b):
        assert np.linalg.norm(u - v) / np.linalg.norm(v) < 1e-5


def test_net_oracle_conv_semantics():
    """Asymmetric stride-2 padding (top/left only), concat order and nearest upsample, by hand."""
    from oracle import net_oracle
    x = np.zeros((1, 4, 4, 1), np.float32)
    x[0, 0, 0, 0] = 1.0
    k = np.zeros((3, 3, 1, 1), np.float32)
    k[0, 0, 0, 0] = 5.0     # top-left tap sees the padded zero row/col for output (0,0)
    k[1, 1, 0, 0] = 1.0     # centre tap: output(0,0) reads input(0,0) when pad_top=pad_left=1
    y = net_oracle.conv_layer(x, k, np.zeros(1, np.float32), 3, 2, False)
    assert y.shape == (1, 2, 2, 1) and y[0, 0, 0, 0] == 1.0 and y.sum() == 1.0
    k2 = np.zeros((3, 3, 1, 1), np.float32)
    k2[2, 2, 0, 0] = 1.0    # bottom-right tap: output(1,1) reads input(3,3) -> no bottom/right padding needed
    x2 = np.zeros((1, 4, 4, 1), np.float32)
    x2[0, 3, 3, 0] = 2.0
    y2 = net_oracle.conv_layer(x2, k2, np.zeros(1, np.float32), 3, 2, False)
    assert y2[0, 1, 1, 0] == 2.0
    up = net_oracle.conv_layer(np.arange(4, dtype=np.float32).reshape(1, 2, 2, 1), np.ones((1, 1, 1, 1), np.float32),
                               np.zeros(1, np.float32), 1, 1, False, upsample=True)
    assert up[0, :, :, 0].tolist() == [[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]]


def test_maxpool_oracle_tf_same_rule():
    """Keras MaxPooling2D 'same' (parse_model.py:78-99): ceil(H/stride) outputs, padding after the data when the total is
    odd, and the padding never wins; 'valid' drops the ragged edge."""
    import torch
    from oracle import net_oracle
    a = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5) - 30.0      # all negative: zero padding would win
    s1 = net_oracle.maxpool_tf(a, 2, 1, True)                                   # yolov3-tiny's last pool
    assert s1.shape == (1, 1, 5, 5)
    exp = a.clone()
    exp[..., :4, :4] = a[..., 1:, 1:]
    exp[..., :4, 4] = a[..., 1:, 4]
    exp[..., 4, :4] = a[..., 4, 1:]
    assert torch.equal(s1, exp)
    s2 = net_oracle.maxpool_tf(a, 2, 2, True)
    assert s2.shape == (1, 1, 3, 3) and s2[0, 0, 2, 2] == a[0, 0, 4, 4] and s2[0, 0, 0, 0] == a[0, 0, 1, 1]
    v = net_oracle.maxpool_tf(a, 2, 2, False)
    assert v.shape == (1, 1, 2, 2) and v[0, 0, 1, 1] == a[0, 0, 3, 3]
    s3 = net_oracle.maxpool_tf(a, 3, 2, True)      # total padding 2 -> one before, one after
    assert s3.shape == (1, 1, 3, 3) and s3[0, 0, 0, 0] == a[0, 0, 1, 1] and s3[0, 0, 2, 2] == a[0, 0, 4, 4]


def test_preprocess_oracle_vs_torch_bilinear():
    """The numpy restatement of TF2's ResizeBilinear(half_pixel_centers=True) against torch's independent implementation
    of the same formula (align_corners=False, no antialias): equal up to float32 rounding, up- and down-sampling."""
    import torch
    import torch.nn.functional as F
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(0)
    for (h, w, oh, ow) in [(37, 53, 64, 64), (480, 640, 416, 416), (100, 80, 33, 57), (5, 5, 5, 5)]:
        img = rng.random((h, w, 3), dtype=np.float32)
        ours = po.resize_bilinear(img, oh, ow)
        ref = F.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None], size=(oh, ow), mode="bilinear",
                            align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
        # the two differ only in how the source coordinate (up to ~640 here, float32 ulp 6e-5) is rounded before the
        # fractional weight is taken: |diff| <= ulp(coord) * |pixel difference| ~ 3e-5
        np.testing.assert_allclose(ours, ref, rtol=0, atol=5e-5)
    # identity resize returns the image; uint8 sources are converted, not rescaled
    u8 = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    assert np.array_equal(po.resize_bilinear(u8, 9, 7), u8.astype(np.float32))
    # aspect-preserving resize + centred zero padding (core/utils.py:17-28)
    out = po.resize_image(rng.random((50, 100, 3), dtype=np.float32), 64, 64)
    assert out.shape == (64, 64, 3) and po.aspect_size(50, 100, 64, 64) == (32, 64)
    assert (out[:16] == 0).all() and (out[48:] == 0).all() and (out[16:48] != 0).any()


def test_preprocess_oracle_vs_opencv_bilinear():
    """Third opinion on the resize restatement: OpenCV's INTER_LINEAR on float32 images uses the same half-pixel-centre,
    edge-clamped formula as TF2's tf.image.resize(bilinear, antialias=False)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(1)
    for (h, w, oh, ow) in [(37, 53, 64, 64), (480, 640, 416, 416), (100, 80, 33, 57), (416, 416, 608, 608)]:
        img = rng.random((h, w, 3), dtype=np.float32)
        ours = po.resize_bilinear(img, oh, ow)
        ref = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
        np.testing.assert_allclose(ours, ref, rtol=0, atol=5e-5)


def test_next_rows_golden():
    """The oracles of the 'next' rows against their committed vectors (tests/golden/make_golden.py: golden_next_rows)."""
    import yolo_v3_tf2_b200 as y3
    from oracle import evaluate_oracle, net_oracle, preprocess_oracle
    z = np.load(os.path.join(GOLD, "tiny64_variance.npz"))
    m = y3.ParseModel.builtin_yolov3_tiny(80).init_weights("variance", seed=int(z["seed"]))
    outs = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, z["x"])
    for o, k in zip(outs, ("g0", "g1")):
        np.testing.assert_allclose(o, z[k], rtol=2e-4, atol=2e-4)
    z = np.load(os.path.join(GOLD, "preprocess_small.npz"))
    assert np.array_equal(preprocess_oracle.resize(z["u8"], 32, divide_by_255=True), z["u8_resized_div255"])
    assert np.array_equal(preprocess_oracle.resize(z["f32"], 40), z["f32_resized"])
    assert np.array_equal(preprocess_oracle.resize_image(z["u8"], 32, 48), z["u8_aspect"])
    assert np.array_equal(preprocess_oracle.resize_image(z["f32"], 64, 64), z["f32_aspect"])
    z = np.load(os.path.join(GOLD, "evaluate_small.npz"))
    n = int(z["nclasses"])
    c = evaluate_oracle.new_counters(n)
    for b in range(len(z["dn"])):
        evaluate_oracle.evaluate(c, n, 0.5, z["det"][b, :z["dn"][b]], z["dcls"][b, :z["dn"][b]], z["gt"][b, :z["gn"][b]],
                                 z["gcls"][b, :z["gn"][b]])
    for k in ("preds", "gts", "tp", "fp", "fn"):
        assert np.array_equal(c[k], z[k]), k
    assert c["examples"] == int(z["examples"])


def test_division_free_div255_is_the_ieee_quotient(tmp_path):
    """preprocess.cuh replaces `resize(...) / 255` (core/load_tfrecords.py:46) by a multiply and two FMAs; the C checker
    compares it with the IEEE division on every 13th float of [0, 256] here (all of them: 34 s, run once, 0 mismatches)."""
    import os
    import subprocess
    src = os.path.join(os.path.dirname(__file__), "div255_check.c")
    exe = str(tmp_path / "div255_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
    out = subprocess.run([exe, "13"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout
