"""Oracle for the NMS stage: restatement of tf.image.non_max_suppression_padded (TensorFlow 2.8.1,
tensorflow/python/ops/image_ops_impl.py: non_max_suppression_padded_v2, _suppression_loop_body, _cross_suppression,
_self_suppression, _bbox_overlap) as called by reference core/yolo_nms.py:26-33 with pad_to_max_output_size=True.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: TensorFlow is not a vendored dependency of
/root/reference and is not installable here; the algorithm below is restated from the published TF 2.8 source.  Two
independent formulations live here and must agree on every test vector:
  * ``nms_padded_tiled``  -- TF's tile_size=512 algorithm, step for step, in numpy
  * ``nms_padded_greedy`` -- plain greedy NMS with the same comparison rules
plus a C port of the greedy form (oracle/nms_oracle.c) used for large batches.

Documented deviation: TF's in-tile fixed-point loop stops when the float32 sum of the IoU matrix changes by <= the
threshold between iterations, which can (for a removed row whose IoU sum is within rounding of the threshold) end one
iteration early; the restatement iterates to convergence, i.e. the greedy result.
"""
import numpy as np

EPS = np.float32(1e-8)


def bbox_overlap(a, b):
    """_bbox_overlap: a [n,4], b [m,4] -> iou [n,m], float32, every op separately rounded."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    a0, a1, a2, a3 = (a[:, i:i + 1] for i in range(4))       # [n,1]
    b0, b1, b2, b3 = (b[:, i][None, :] for i in range(4))    # [1,m]
    i_1min = np.maximum(a1, b1)
    i_1max = np.minimum(a3, b3)
    i_0min = np.maximum(a0, b0)
    i_0max = np.minimum(a2, b2)
    i_area = np.maximum(i_1max - i_1min, np.float32(0)) * np.maximum(i_0max - i_0min, np.float32(0))
    a_area = (a2 - a0) * (a3 - a1)
    b_area = (b2 - b0) * (b3 - b1)
    u_area = a_area + b_area - i_area + EPS
    with np.errstate(divide="ignore", invalid="ignore"):
        return (i_area / u_area).astype(np.float32)


def _prepare(boxes, scores, score_threshold):
    """score filter + stable descending sort (argsort DESCENDING is top_k: ties -> lower index first)."""
    boxes = np.asarray(boxes, np.float32).copy()
    scores = np.asarray(scores, np.float32).copy()
    if score_threshold != float("-inf"):
        mask = (scores > np.float32(score_threshold)).astype(np.float32)
        scores = scores * mask
        boxes = boxes * mask[:, None]
    # canonicalize_coordinates: decided on the first box only; decoded boxes are already (min, min, max, max)
    if not (boxes[0, 0] <= boxes[0, 2]):
        boxes[:, [0, 2]] = boxes[:, [2, 0]]
    if not (boxes[0, 1] <= boxes[0, 3]):
        boxes[:, [1, 3]] = boxes[:, [3, 1]]
    order = np.argsort(-scores, kind="stable")
    # -0.0 and 0.0 compare equal in top_k; np.argsort on the negated array keeps index order for equal keys
    return boxes[order], scores[order], order.astype(np.int32)


def nms_padded_tiled(boxes, scores, max_output_size, iou_threshold, score_threshold, tile_size=512):
    """One image.  boxes [N,4], scores [N] -> (selected_indices_padded [max] int32, num_valid int32)."""
    N = boxes.shape[0]
    thr = np.float32(iou_threshold)
    sboxes, _, order = _prepare(boxes, scores, score_threshold)
    padded = int(np.ceil(max(N, max_output_size) / tile_size)) * tile_size
    sb = np.zeros((padded, 4), np.float32)
    sb[:N] = sboxes
    num_tiles = padded // tile_size
    output_size = 0
    idx = 0
    while output_size < max_output_size and idx < num_tiles:
        sl = slice(idx * tile_size, (idx + 1) * tile_size)
        box_slice = sb[sl].copy()
        for inner in range(idx):                                             # _cross_suppression
            new_slice = sb[inner * tile_size:(inner + 1) * tile_size]
            iou = bbox_overlap(new_slice, box_slice)
            keep = np.all(iou < thr, axis=0)
            box_slice = box_slice * keep[:, None].astype(np.float32)
        iou = bbox_overlap(box_slice, box_slice)                             # _suppression_loop_body
        r = np.arange(tile_size)
        upper = r[None, :] > r[:, None]
        iou = iou * (upper & (iou >= thr)).astype(np.float32)
        while True:                                                          # _self_suppression to convergence
            can_suppress_others = (np.max(iou, axis=0) < thr).astype(np.float32)[:, None]
            row_keep = (np.max(can_suppress_others * iou, axis=0) < thr).astype(np.float32)[:, None]
            new_iou = row_keep * iou
            changed = not np.array_equal(new_iou, iou)
            iou = new_iou
            if not changed:
                break
        suppressed = np.sum(iou, axis=0) > 0
        box_slice = box_slice * (1.0 - suppressed.astype(np.float32))[:, None]
        sb[sl] = box_slice
        output_size += int(np.sum(np.any(box_slice > 0, axis=1)))
        idx += 1
    num_valid = min(output_size, max_output_size)
    valid_pos = np.nonzero(np.any(sb > 0, axis=1))[0][:max_output_size]
    out = np.zeros(max_output_size, np.int32)
    pos = np.minimum(valid_pos[:num_valid], N - 1)
    out[:len(pos)] = order[pos]
    return out, np.int32(num_valid)


def nms_padded_greedy(boxes, scores, max_output_size, iou_threshold, score_threshold):
    """Independent formulation: walk the sorted boxes, keep a box unless an earlier kept box has iou >= thr."""
    N = boxes.shape[0]
    thr = np.float32(iou_threshold)
    sboxes, _, order = _prepare(boxes, scores, score_threshold)
    kept = np.zeros((0, 4), np.float32)
    sel = []
    for i in range(N):
        b = sboxes[i:i + 1]
        if kept.shape[0]:
            iou = bbox_overlap(kept, b)[:, 0]
            if np.any(iou >= thr):
                continue
        kept = np.concatenate([kept, b], 0)
        if np.any(b > 0):
            sel.append(order[i])
            if len(sel) >= max_output_size:
                break
    out = np.zeros(max_output_size, np.int32)
    out[:len(sel)] = sel
    return out, np.int32(len(sel))


def nms_batch(boxes, scores, max_output_size, iou_threshold, score_threshold, impl="greedy"):
    """[B,N,4], [B,N] -> ([B,max] int32, [B] int32).  Note TF decides coordinate canonicalisation from image 0's
    first box for the whole batch; decoded boxes never trigger it, and the per-image rule here matches for them."""
    f = {"greedy": nms_padded_greedy, "tiled": nms_padded_tiled}[impl]
    sels, nv = [], []
    for b in range(boxes.shape[0]):
        s, n = f(boxes[b], scores[b], max_output_size, iou_threshold, score_threshold)
        sels.append(s)
        nv.append(n)
    return np.stack(sels), np.array(nv, np.int32)


def yolo_nms(outputs, yolo_max_boxes, nms_iou_threshold, nms_score_threshold, impl="greedy"):
    """reference core/yolo_nms.py:15-34 -> (bboxes, class_indices int64, scores, selected_indices_padded, num_valid)."""
    from .decode_oracle import class_reduce
    bboxes, confidence, class_probs = outputs
    class_indices, scores = class_reduce(confidence, class_probs)
    bboxes = np.asarray(bboxes, np.float32).reshape(bboxes.shape[0], -1, 4)
    sel, nv = nms_batch(bboxes, scores, yolo_max_boxes, nms_iou_threshold, nms_score_threshold, impl)
    return bboxes, class_indices, scores, sel, nv
