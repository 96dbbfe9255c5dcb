"""Oracle for the evaluation counters: numpy restatement of reference evaluate_detections.py:39-135.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned (TensorFlow unavailable)."""
import numpy as np


def iou_alg(box_1, box_2):
    """evaluate_detections.py:39-48 for box_1 [P,4] against box_2 [G,4] -> [P,G] (float32, separate roundings)."""
    b1 = box_1[:, None, :].astype(np.float32)
    b2 = box_2[None, :, :].astype(np.float32)
    ow = np.maximum(np.minimum(b1[..., 2], b2[..., 2]) - np.maximum(b1[..., 0], b2[..., 0]), np.float32(0))
    oh = np.maximum(np.minimum(b1[..., 3], b2[..., 3]) - np.maximum(b1[..., 1], b2[..., 1]), np.float32(0))
    ov = ow * oh
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        return ov / (a1 + a2 - ov)


def evaluate(counters, nclasses, iou_thresh, pred_bboxes, pred_classes, gt_bboxes, gt_classes):
    """One image: calc_iou (:122-133), process_decisions (:83-118), update_counters (:57-80).  ``counters`` is a dict of
    int arrays preds/gts/tp/fp/fn [nclasses] plus ints examples/errors, updated in place."""
    pc = np.asarray(pred_classes).astype(np.int64)
    gc = np.asarray(gt_classes).astype(np.int64)
    if ((gc < 0) | (gc >= nclasses)).any() or ((pc < 0) | (pc >= nclasses)).any():
        counters["errors"] += 1
        return counters
    P, G = len(pc), len(gc)
    assigned = np.zeros(G, bool)
    decisions = np.zeros(P, bool)
    if P and G:
        iou = iou_alg(np.asarray(pred_bboxes, np.float32).reshape(P, 4), np.asarray(gt_bboxes, np.float32).reshape(G, 4))
        arg = iou.argmax(axis=-1)
        mx = iou[np.arange(P), arg]
        decisions = (mx > np.float32(iou_thresh)) & (gc[arg] == pc) & ~assigned[arg]   # assigned is all False here
        np.logical_or.at(assigned, arg[decisions], True)
    np.add.at(counters["tp"], pc, decisions.astype(np.int64))
    np.add.at(counters["fp"], pc, (~decisions).astype(np.int64))
    np.add.at(counters["fn"], gc, (~assigned).astype(np.int64))
    np.add.at(counters["gts"], gc, 1)
    np.add.at(counters["preds"], pc, 1)
    counters["examples"] += 1
    return counters


def new_counters(nclasses):
    c = {k: np.zeros(nclasses, np.int64) for k in ("preds", "gts", "tp", "fp", "fn")}
    c["examples"] = 0
    c["errors"] = 0
    return c
