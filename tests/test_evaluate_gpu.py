"""Evaluation counters (reference evaluate_detections.py) on the GPU vs the numpy oracle:
bit-exact integer counters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(rng, B, max_det, max_gt, nclasses):
    gt_n = rng.integers(0, max_gt + 1, B)
    det_n = rng.integers(0, max_det + 1, B)
    c = rng.random((B, max_gt, 2)).astype(np.float32)
    wh = (rng.random((B, max_gt, 2)) * 0.2 + 0.05).astype(np.float32)
    gt = np.concatenate([c - wh / 2, c + wh / 2], -1).astype(np.float32)
    gcls = rng.integers(0, nclasses, (B, max_gt)).astype(np.int32)
    # detections: jittered copies of ground-truth boxes (some with the wrong class) plus random boxes
    det = np.zeros((B, max_det, 4), np.float32)
    dcls = np.zeros((B, max_det), np.int64)
    for b in range(B):
        for p in range(max_det):
            if gt_n[b] and rng.random() < 0.7:
                g = rng.integers(0, gt_n[b])
                det[b, p] = gt[b, g] + rng.normal(0, 0.02, 4).astype(np.float32)
                dcls[b, p] = gcls[b, g] if rng.random() < 0.8 else rng.integers(0, nclasses)
            else:
                cc = rng.random(2).astype(np.float32)
                det[b, p] = np.concatenate([cc - 0.05, cc + 0.05])
                dcls[b, p] = rng.integers(0, nclasses)
    return det, dcls, det_n.astype(np.int32), gt, gcls, gt_n.astype(np.int32)


def test_counters_match_oracle(cuda):
    import torch
    from yolo_v3_tf2_b200.evaluate_detections import EvaluateDetections
    from oracle import evaluate_oracle as eo
    rng = np.random.default_rng(0)
    nclasses = 7
    ev = EvaluateDetections(nclasses, 0.5)
    ref = eo.new_counters(nclasses)
    for B, md, mg in [(16, 100, 20), (5, 3, 1), (9, 40, 60)]:
        det, dcls, dn, gt, gcls, gn = _case(rng, B, md, mg, nclasses)
        ev.evaluate_batch(*(torch.from_numpy(a).cuda() for a in (det, dcls, dn, gt, gcls, gn)))
        for b in range(B):
            eo.evaluate(ref, nclasses, 0.5, det[b, :dn[b]], dcls[b, :dn[b]], gt[b, :gn[b]], gcls[b, :gn[b]])
    torch.cuda.synchronize()
    got = ev.counters
    for k in ("preds", "gts", "tp", "fp", "fn"):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    assert int(got["examples"]) == ref["examples"] == 30 and int(got["errors"]) == 0
    assert ref["tp"].sum() > 50 and ref["fp"].sum() > 50 and ref["fn"].sum() > 10      # the case has teeth


def test_single_image_api_and_bad_class(cuda):
    """The reference's per-image signature; a class id outside [0, nclasses) skips the sample and counts an error
    (update_counters' except branch, evaluate_detections.py:66-72)."""
    from yolo_v3_tf2_b200.evaluate_detections import EvaluateDetections
    from oracle import evaluate_oracle as eo
    ev = EvaluateDetections(3, 0.5)
    ref = eo.new_counters(3)
    pb = np.array([[0.1, 0.1, 0.4, 0.4], [0.12, 0.1, 0.4, 0.42], [0.6, 0.6, 0.9, 0.9]], np.float32)
    pc = np.array([1, 1, 2])
    gb = np.array([[0.1, 0.1, 0.4, 0.4], [0.5, 0.5, 0.7, 0.7]], np.float32)
    gc = np.array([1, 0])
    ev.evaluate(pb, pc, gb, gc)
    eo.evaluate(ref, 3, 0.5, pb, pc, gb, gc)
    # both predictions of class 1 claim the same ground-truth box and both count (vectorised decision in the reference)
    assert ref["tp"].tolist() == [0, 2, 0] and ref["fp"].tolist() == [0, 0, 1] and ref["fn"].tolist() == [1, 0, 0]
    ev.evaluate(pb, pc, gb, np.array([1, -1]))
    eo.evaluate(ref, 3, 0.5, pb, pc, gb, np.array([1, -1]))
    ev.evaluate(np.zeros((0, 4), np.float32), np.zeros(0, np.int64), gb, gc)        # no detections: all ground truth missed
    eo.evaluate(ref, 3, 0.5, np.zeros((0, 4), np.float32), np.zeros(0, np.int64), gb, gc)
    got = ev.counters
    for k in ("preds", "gts", "tp", "fp", "fn"):
        assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
    assert int(got["examples"]) == 2 and int(got["errors"]) == 1
