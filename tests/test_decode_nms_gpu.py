"""GPU parity of decode / class-reduce / NMS (through the Python mirror of the reference API -> C ABI) vs the oracle.

Decode tolerance: sigmoid/exp come from different math libraries (CUDA expf <= 2 ulp, numpy float32 exp <= 1 ulp), so
probabilities/objectness must match within 2e-6 absolute and boxes within 2e-6 + 2e-6*|ref|.
NMS: bit-exact selected indices and counts on identical (oracle-decoded) inputs.
"""
import numpy as np
import pytest

from y3_test_util import synth_grids, cluster_boxes

pytestmark = pytest.mark.gpu


def _anchors():
    from yolo_v3_tf2_b200 import configs
    return configs.coco_anchors()


@pytest.mark.parametrize("B,sizes,C", [
    (2, (13, 26, 52), 80), (3, (13, 26, 52), 38), (1, (13, 26, 52), 37), (2, (19, 38, 76), 80), (5, (2, 4, 8), 3),
    (1, ((13, 26), (26, 52), (52, 104)), 80),   # non-square grids: the reference's (gh, gw) divisor quirk
])
def test_decode_vs_oracle(cuda, B, sizes, C):
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import decode_oracle
    grids = synth_grids(B, sizes, C, seed=B + C)
    ref = decode_oracle.yolo_decode(grids, _anchors(), C)
    got = y3.yolo_decode([torch.from_numpy(g).cuda() for g in grids], _anchors(), C, with_scores=True)
    torch.cuda.synchronize()
    for name, g, r in zip(("bboxes", "conf", "probs"), got[:3], ref):
        g = g.cpu().numpy()
        assert g.shape == r.shape, name
        np.testing.assert_allclose(g, r, rtol=2e-6, atol=2e-6, err_msg=name)
    # fused scores / classes must be bit-identical to a class-reduce of the kernel's OWN probabilities
    cls_ref, sc_ref = decode_oracle.class_reduce(got[1].cpu().numpy(), got[2].cpu().numpy())
    assert np.array_equal(got[4].cpu().numpy(), cls_ref)
    assert np.array_equal(got[3].cpu().numpy(), sc_ref)
    # and the un-fused API returns the same three tensors bit for bit
    got2 = y3.yolo_decode([torch.from_numpy(g).cuda() for g in grids], _anchors(), C)
    for a, b in zip(got[:3], got2):
        assert torch.equal(a, b)


def test_decode_accepts_numpy_and_checks_shapes(cuda):
    import yolo_v3_tf2_b200 as y3
    grids = synth_grids(1, (2, 4, 8), 5, seed=0)
    b, c, p = y3.yolo_decode(grids, _anchors(), 5)
    assert b.shape == (1, 3 * (4 + 16 + 64), 4) and c.shape[-1] == 1 and p.shape[-1] == 5
    with pytest.raises(ValueError):
        y3.yolo_decode(grids, _anchors(), 6)


@pytest.mark.parametrize("C,N,misalign", [(80, 1000, 0), (38, 1000, 0), (5, 1001, 0), (81, 777, 0), (600, 131, 0),
                                          (80, 333, 1), (20000, 7, 0)])
def test_class_reduce_vs_oracle(cuda, C, N, misalign):
    """staged kernel (C % 4 == 0 vector scan, even / odd scalar scans, partial last chunk with a ragged 16-byte tail,
    wide class vectors) and the warp-per-record fallback (class_probs not 16-byte aligned, C too wide to stage)"""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import decode_oracle
    rng = np.random.default_rng(C)
    B = 3
    probs = rng.random((B, N, C)).astype(np.float32)
    probs[0, :50, :] = 0.25            # all-equal rows: first index must win
    probs[1, 5, C - 1] = probs[1, 5, 3] = 2.0   # duplicated max: lower index wins
    probs[2, 6, 1:] = probs[2, 6, 0]   # maximum at index 0 tied with every later class
    conf = rng.random((B, N, 1)).astype(np.float32)
    boxes = np.zeros((B, N, 4), np.float32)
    pd = torch.from_numpy(probs).cuda()
    if misalign:
        flat = torch.empty(pd.numel() + 4, dtype=torch.float32, device="cuda")
        pd = flat[misalign:misalign + pd.numel()].copy_(pd.reshape(-1)).view(B, N, C)
        assert pd.data_ptr() % 16 != 0
    out = y3.yolo_nms((torch.from_numpy(boxes).cuda(), torch.from_numpy(conf).cuda(), pd), 10, 0.5, 0.1)
    cls_ref, sc_ref = decode_oracle.class_reduce(conf, probs)
    assert out[1].dtype == torch.int64
    assert np.array_equal(out[1].cpu().numpy(), cls_ref)
    assert np.array_equal(out[2].cpu().numpy(), sc_ref)


def _check_nms(boxes, scores, max_boxes, iou, sthr, impl="c"):
    import torch
    from yolo_v3_tf2_b200.core.yolo_nms import nms_padded
    from oracle import nms_oracle, c_oracle
    sel, nv, status = nms_padded(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), max_boxes, iou, sthr)
    torch.cuda.synchronize()
    assert int(status.max().item()) == 0
    if impl == "c":
        rsel, rnv = c_oracle.nms(boxes, scores, max_boxes, iou, sthr)
    else:
        rsel, rnv = nms_oracle.nms_batch(boxes, scores, max_boxes, iou, sthr, impl)
    assert np.array_equal(nv.cpu().numpy(), rnv), (nv.cpu().numpy(), rnv)
    assert np.array_equal(sel.cpu().numpy(), rsel)
    return rnv


@pytest.mark.parametrize("iou", [0.3, 0.5, 0.7])
@pytest.mark.parametrize("sthr", [0.004, 0.1, 0.5, 0.9])
def test_nms_clustered_vs_oracles(cuda, iou, sthr):
    cases = [cluster_boxes(2000, 15, 0), cluster_boxes(3000, 40, 1), cluster_boxes(5000, 150, 2), cluster_boxes(700, 3, 3)]
    for b, s in cases:
        s = s.copy()
        s[::7] = s[3]          # exact score ties
        b = b.copy()
        b[10] = b[11]          # duplicate boxes (iou == 1)
        for impl in ("c", "tiled"):
            _check_nms(b[None], s[None], 100, iou, sthr, impl)


def test_nms_decoded_dense_and_sparse(cuda):
    """config 4 style: oracle-decoded random logits, dense (obj ~ N(0,2)) and sparse (obj ~ N(-6,2))."""
    from oracle import decode_oracle
    for obj_mean in (0.0, -6.0):
        grids = synth_grids(4, (13, 26, 52), 80, seed=11, obj_mean=obj_mean)
        bboxes, conf, probs = decode_oracle.yolo_decode(grids, _anchors(), 80)
        cls, scores = decode_oracle.class_reduce(conf, probs)
        for iou in (0.3, 0.5, 0.7):
            for sthr in (0.004, 0.1, 0.5, 0.9):
                _check_nms(bboxes, scores, 100, iou, sthr)


def test_nms_edge_cases(cuda):
    rng = np.random.default_rng(3)
    # (a) nothing passes the threshold -> num_valid 0, all-zero indices
    b, s = cluster_boxes(600, 10, 5)
    nv = _check_nms(b[None], (s * 0.05)[None], 100, 0.5, 0.1)
    assert nv[0] == 0
    # (b) fewer survivors than max_boxes, several images with different counts
    bs = np.stack([cluster_boxes(900, k, 10 + k)[0] for k in (2, 5, 9)])
    ss = np.stack([cluster_boxes(900, k, 10 + k)[1] for k in (2, 5, 9)])
    _check_nms(bs, ss, 100, 0.5, 0.1)
    # (c) exact iou == threshold: two unit-overlap boxes with iou exactly 1/3, and thr = float32(1/3) -> suppressed (>=)
    b3 = np.array([[0.0, 0.0, 0.5, 1.0], [0.25, 0.0, 0.75, 1.0], [0.9, 0.9, 1.0, 1.0]], np.float32)
    s3 = np.array([0.9, 0.8, 0.7], np.float32)
    from oracle import nms_oracle
    thr = float(nms_oracle.bbox_overlap(b3[:1], b3[1:2])[0, 0])
    _check_nms(b3[None], s3[None], 10, thr, 0.1)
    _check_nms(b3[None], s3[None], 10, float(np.nextafter(np.float32(thr), np.float32(1))), 0.1)
    # (d) boxes with no positive coordinate survive but are never selected ("invisible" boxes)
    b4 = np.array([[-0.5, -0.5, -0.1, -0.1], [-0.45, -0.45, -0.1, -0.1], [0.1, 0.1, 0.4, 0.4], [-0.3, -0.2, 0.0, 0.0]], np.float32)
    s4 = np.array([0.9, 0.8, 0.7, 0.6], np.float32)
    _check_nms(b4[None], s4[None], 10, 0.5, 0.1)
    # (e) max_boxes larger than N, and max_boxes == 1
    b5, s5 = cluster_boxes(40, 40, 7, jitter=0.2)
    _check_nms(b5[None], s5[None], 100, 0.5, 0.1)
    _check_nms(b5[None], s5[None], 1, 0.5, 0.1)
    # (f) negative score threshold / scores (general, un-compacted path)
    b6, s6 = cluster_boxes(500, 20, 8)
    _check_nms(b6[None], (s6 - 0.5)[None], 50, 0.5, -0.25)
    # (g) N not a multiple of anything, N just above a power of two
    b7, s7 = cluster_boxes(1025, 30, 9)
    _check_nms(b7[None], s7[None], 100, 0.5, 0.1)


def test_nms_top_candidate_prefilter_paths(cuda):
    """More than 2048 passing candidates: the kernel sorts only the best 1024 (radix select).  Exercised here:
    (a) 608x608 size, N = 22 743 dense (largest shared-memory footprint), (b) the best 1024 collapse to a handful of
    survivors so the image must be redone with all candidates, (c) > 4096 candidates tied at the selection key."""
    from oracle import decode_oracle
    grids = synth_grids(2, (19, 38, 76), 80, seed=5, obj_mean=1.0)
    bboxes, conf, probs = decode_oracle.yolo_decode(grids, _anchors(), 80)
    cls, scores = decode_oracle.class_reduce(conf, probs)
    assert bboxes.shape[1] == 22743 and (scores > 0.05).sum(1).min() > 8000
    _check_nms(bboxes, scores, 100, 0.5, 0.05)
    # (b) 1500 near-identical high-scoring boxes (one survivor) + 4000 scattered low-scoring ones (the other 99)
    rng = np.random.default_rng(17)
    hot = np.array([0.4, 0.4, 0.6, 0.6], np.float32) + rng.normal(0, 1e-3, (1500, 4)).astype(np.float32)
    c = rng.random((4000, 2)).astype(np.float32)
    cold = np.concatenate([c - 0.01, c + 0.01], 1).astype(np.float32)
    b = np.concatenate([hot, cold])
    s = np.concatenate([0.8 + 0.1 * rng.random(1500), 0.2 + 0.1 * rng.random(4000)]).astype(np.float32)
    perm = rng.permutation(len(s))
    nv = _check_nms(b[perm][None], s[perm][None], 100, 0.5, 0.1)
    assert nv[0] == 100
    # (c) 5000 candidates with the same score: ties are broken by the lower index
    s_tie = np.full(5500, 0.5, np.float32)
    s_tie[:500] = 0.9
    _check_nms(b[None], s_tie[None], 100, 0.5, 0.1)
    _check_nms(b[perm][None], s_tie[None], 100, 0.5, 0.1, "tiled")


def test_nms_full_size_batch_properties(cuda):
    """BASELINE config 4 size (N=10647, B=64 here): checked against the C oracle on every image, plus
    size-independent properties: sorted-by-score output, zero padding, survivors mutually below the IoU threshold."""
    import torch
    from yolo_v3_tf2_b200.core.yolo_nms import nms_padded
    from oracle import decode_oracle, nms_oracle
    grids = synth_grids(64, (13, 26, 52), 80, seed=21, obj_mean=-3.0)
    bboxes, conf, probs = decode_oracle.yolo_decode(grids, _anchors(), 80)
    cls, scores = decode_oracle.class_reduce(conf, probs)
    _check_nms(bboxes, scores, 100, 0.5, 0.1)
    sel, nv, _ = nms_padded(torch.from_numpy(bboxes).cuda(), torch.from_numpy(scores).cuda(), 100, 0.5, 0.1)
    sel, nv = sel.cpu().numpy(), nv.cpu().numpy()
    for b in range(0, 64, 7):
        n = nv[b]
        assert (sel[b, n:] == 0).all()
        sc = scores[b, sel[b, :n]]
        assert (np.diff(sc) <= 0).all() and (sc > 0.1).all()
        iou = nms_oracle.bbox_overlap(bboxes[b, sel[b, :n]], bboxes[b, sel[b, :n]])
        assert (iou[np.triu_indices(n, 1)] < 0.5).all()


def test_yolo_nms_layer_matches_reference_api(cuda):
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import decode_oracle, nms_oracle
    grids = synth_grids(2, (13, 26, 52), 80, seed=4, obj_mean=-2.0)
    dec = decode_oracle.yolo_decode(grids, _anchors(), 80)
    layer = y3.YoloNmsLayer(100, 0.5, 0.1)
    out = layer(tuple(torch.from_numpy(d).cuda() for d in dec))
    ref = nms_oracle.yolo_nms(dec, 100, 0.5, 0.1)
    assert len(out) == 5
    assert out[0].shape == ref[0].shape and out[1].dtype == torch.int64 and out[3].dtype == torch.int32
    for o, r in zip(out, ref):
        assert np.array_equal(o.cpu().numpy(), r)
    # gather_valid_detections_results (inference.py:21-28)
    bb, cc, ss = y3.Inference.gather_valid_detections_results(out[0][0], out[1][0], out[2][0], out[3][0], out[4][0])
    n = int(ref[4][0])
    assert np.array_equal(bb.cpu().numpy(), ref[0][0][ref[3][0][:n]])
    assert np.array_equal(cc.cpu().numpy(), ref[1][0][ref[3][0][:n]])
    assert np.array_equal(ss.cpu().numpy(), ref[2][0][ref[3][0][:n]])


@pytest.mark.parametrize("C", [80, 37])
def test_compact_decode_and_packed_gather(cuda, C):
    """The fused pipeline's decode writes only boxes / scores / class ids (conf and probs are not needed by yolo_nms,
    core/yolo_nms.py:18-33) and its gather also emits the packed records of distributed.pack_detections: both must
    equal the full-output calls bit for bit."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    from yolo_v3_tf2_b200.core.yolo_nms import nms_padded
    from yolo_v3_tf2_b200.inference import gather_detections_batched
    from yolo_v3_tf2_b200.distributed import pack_detections
    rng = np.random.default_rng(C)
    B = 3
    grids = [torch.from_numpy(rng.normal(0, 1.5, (B, g, g, 3, 5 + C)).astype(np.float32)).cuda() for g in (13, 26, 52)]
    anchors = configs.coco_anchors()
    full = y3.yolo_decode(grids, anchors, C, with_scores=True)
    comp = y3.yolo_decode(grids, anchors, C, compact=True)
    assert comp[1] is None and comp[2] is None
    for k in (0, 3, 4):
        assert torch.equal(full[k], comp[k])
    sel, nv, st = nms_padded(comp[0], comp[3], 100, 0.5, 0.1)
    ob, oc, os_ = gather_detections_batched(comp[0], comp[4], comp[3], sel, nv)
    pb, pc, ps, rec = gather_detections_batched(comp[0], comp[4], comp[3], sel, nv, packed=True)
    assert torch.equal(ob, pb) and torch.equal(oc, pc) and torch.equal(os_, ps)
    assert torch.equal(rec, pack_detections(ob, oc, os_, nv))
    assert int(nv.min()) > 0


@pytest.mark.parametrize("C", [80, 37, 3])
def test_compact_decode_class_ties_and_saturation(cuda, C):
    """The compact decode finds the class from the raw logits and takes one sigmoid; it must still return exactly the
    first arg-max of the float32 probabilities (core/yolo_nms.py:18-24 applies tf.argmax to sigmoid outputs): saturated
    logits (sigmoid == 1.0f for several classes), exact and near ties, all-zero probabilities (logits < -104), denormal
    probabilities, NaN logits, +-inf."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    rng = np.random.default_rng(100 + C)
    B = 2
    grids = []
    for g in (13, 26, 52):
        t = rng.normal(0, 1.5, (B, g, g, 3, 5 + C)).astype(np.float32)
        cls = t[..., 5:].reshape(-1, C)
        n = cls.shape[0]
        kind = rng.integers(0, 12, n)
        for k in range(n):
            row = cls[k]
            a, b = rng.choice(C, 2, replace=False) if C > 1 else (0, 0)
            if kind[k] == 0:      # several saturated classes: sigmoid == 1.0f for all of them
                row[rng.choice(C, min(C, 3), replace=False)] = rng.uniform(17.5, 60.0, min(C, 3))
            elif kind[k] == 1:    # exact tie of the two largest logits
                row[a] = row[b] = np.float32(row.max() + 1.0)
            elif kind[k] == 2:    # near tie: 1 ulp apart, at moderate and at large logits
                base = np.float32(rng.choice([0.3, 2.0, 5.5, 7.0, 9.0, 12.0, 15.0]))
                row[a] = base
                row[b] = np.nextafter(base, np.float32(100.0))
            elif kind[k] == 3:    # everything underflows to probability 0 (class 0 must win)
                row[:] = rng.uniform(-200.0, -105.0, C)
            elif kind[k] == 4:    # denormal probabilities
                row[:] = rng.uniform(-103.0, -88.0, C)
            elif kind[k] == 5:    # the flat part of the sigmoid: logits within a few 1e-3 of each other around 6..16
                row[:] = np.float32(rng.uniform(5.0, 16.0)) + rng.uniform(-4e-3, 4e-3, C).astype(np.float32)
            elif kind[k] == 6:
                row[a] = np.float32("inf")
            elif kind[k] == 7:
                row[:] = np.float32("-inf")
            elif kind[k] == 8 and k % 7 == 0:
                row[a] = np.float32("nan")
        grids.append(torch.from_numpy(t).cuda())
    anchors = configs.coco_anchors()
    full = y3.yolo_decode(grids, anchors, C, with_scores=True)
    comp = y3.yolo_decode(grids, anchors, C, compact=True)
    probs = full[2].cpu().numpy()
    # the non-compact kernel's own scores / classes are the class reduce of the probabilities it wrote ...
    finite = ~np.isnan(probs).any(axis=2)
    assert np.array_equal(full[4].cpu().numpy()[finite], probs.argmax(axis=2)[finite])
    # ... and the compact path agrees with it bit for bit (NaN scores compare equal as bit patterns)
    assert torch.equal(full[0], comp[0])
    assert np.array_equal(full[4].cpu().numpy(), comp[4].cpu().numpy())
    assert np.array_equal(full[3].cpu().numpy().view(np.uint32), comp[3].cpu().numpy().view(np.uint32))


def test_nms_rejects_unsupported_parameters(cuda):
    """parameters the kernel cannot honour are refused instead of returning truncated results."""
    import torch
    from yolo_v3_tf2_b200 import _lib
    from yolo_v3_tf2_b200.core.yolo_nms import nms_padded
    b = torch.rand((1, 64, 4), device="cuda")
    s = torch.rand((1, 64), device="cuda")
    with pytest.raises(_lib.Y3Unsupported, match="yolo_max_boxes"):
        nms_padded(b, s, 769, 0.5, 0.1)
    with pytest.raises(_lib.Y3Unsupported, match="nms_iou_threshold"):
        nms_padded(b, s, 100, 0.0, 0.1)
    nms_padded(b, s, 768, 0.5, 0.1)
