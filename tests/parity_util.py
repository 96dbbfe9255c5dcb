"""Shared parity measurements: GPU path vs the CPU oracle on the same seeded inputs and weights.

Used by tests/test_parity_full_gpu.py (asserts the stated tolerances) and tools/parity_table.py (writes the achieved
errors to profiles/r2_parity.md).  Follows reference inference.py:109-116 (model -> yolo_decode -> YoloNmsLayer) and
core/yolo_decode_layer.py:4-12 for what is compared.
"""
import numpy as np


def oracle_forward_chunked(model, x_np, chunk=8):
    """torch-CPU fp32 oracle of the whole graph, a few images at a time (bounds the oracle's activation memory)."""
    from oracle import net_oracle
    outs = None
    for i0 in range(0, x_np.shape[0], chunk):
        o = net_oracle.forward(model.graph.layers, model.graph.outputs, model._params, x_np[i0:i0 + chunk])
        if outs is None:
            outs = [[] for _ in o]
        for k, g in enumerate(o):
            outs[k].append(g)
    return [np.concatenate(c, 0) for c in outs]


def logits_errors(gpu_grids, ref_grids):
    """per head: relative L2 error, max-abs error, max |ref|, and max-abs relative to max |ref|"""
    rows = []
    for g, r in zip(gpu_grids, ref_grids):
        g = np.asarray(g, np.float32)
        d = (g.astype(np.float64) - r.astype(np.float64))
        rel = float(np.linalg.norm(d) / max(np.linalg.norm(r.astype(np.float64)), 1e-30))
        mx = float(np.abs(d).max())
        mr = float(np.abs(r).max())
        rows.append({"rel_l2": rel, "max_abs": mx, "max_ref": mr, "max_abs_over_max_ref": mx / max(mr, 1e-30)})
    return rows


def box_errors(gpu_grids, ref_grids, anchors, nclasses, twh_limit=2.0):
    """Decoded boxes of the GPU's own logits vs the oracle's decode of the oracle's logits
    (core/yolo_decode_layer.py:4-12), over the boxes whose reference w/h logits are moderate (|t_wh| <= twh_limit:
    exp() amplifies a logit error e into a relative size error exp(e) - 1, so the absolute error of a huge box says
    nothing about the conv stack).  Image-fraction units.
      centre_abs : max |(xy centre)_gpu - (xy centre)_ref|
      wh_rel     : max |wh_gpu / wh_ref - 1|
      box_abs    : max |corner_gpu - corner_ref| over the same boxes"""
    from oracle import decode_oracle
    bg, cg, pg = decode_oracle.yolo_decode(gpu_grids, anchors, nclasses)
    br, cr, pr = decode_oracle.yolo_decode(ref_grids, anchors, nclasses)
    twh = np.concatenate([np.abs(np.asarray(r)[..., 2:4]).max(-1).reshape(r.shape[0], -1) for r in ref_grids], 1)
    ok = twh <= twh_limit
    cen_g, cen_r = (bg[..., :2] + bg[..., 2:]) / 2, (br[..., :2] + br[..., 2:]) / 2
    wh_g, wh_r = bg[..., 2:] - bg[..., :2], br[..., 2:] - br[..., :2]
    out = {
        "boxes_compared": int(ok.sum()), "boxes_total": int(ok.size),
        "centre_abs": float(np.abs(cen_g - cen_r)[ok].max()),
        "wh_rel": float(np.abs(wh_g / wh_r - 1.0)[ok].max()),
        "box_abs": float(np.abs(bg - br)[ok].max()),
        "conf_abs": float(np.abs(cg - cr).max()),
        "prob_abs": float(np.abs(pg - pr).max()),
    }
    return out, (bg, cg, pg), (br, cr, pr)


def _iou_matrix(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-8)


def nms_set_overlap(dec_gpu, dec_ref, max_boxes, iou_thr, score_thr, match_iou=0.9):
    """Informational: NMS of the GPU's own decoded tensors vs NMS of the oracle's, both through
    the oracle NMS, compared as sets: the fraction of reference detections that have a GPU detection of the same class
    with IoU >= match_iou, and the fraction of images whose detection COUNT agrees."""
    from oracle import decode_oracle, c_oracle
    res = []
    for (b, c, p) in (dec_gpu, dec_ref):
        cls, sc = decode_oracle.class_reduce(c, p)
        sel, nv = c_oracle.nms(b, sc, max_boxes, iou_thr, score_thr)
        res.append((b, cls, sel, nv))
    (bg, clg, selg, nvg), (br, clr, selr, nvr) = res
    matched = total = same_count = 0
    for i in range(bg.shape[0]):
        ig, ir = selg[i, :nvg[i]], selr[i, :nvr[i]]
        same_count += int(nvg[i] == nvr[i])
        total += len(ir)
        if len(ir) == 0 or len(ig) == 0:
            continue
        iou = _iou_matrix(br[i][ir], bg[i][ig])
        same_cls = clr[i][ir][:, None] == clg[i][ig][None, :]
        matched += int(((iou >= match_iou) & same_cls).any(1).sum())
    return {"ref_detections": int(total), "matched": int(matched),
            "matched_frac": (matched / total) if total else 1.0,
            "images_same_count_frac": same_count / bg.shape[0],
            "gpu_detections": int(nvg.sum())}


def measure(init, size, B, C, seed=3, with_nms=True, score_thr=0.1, chunk=8, nms_images=None):
    """GPU forward (+decode) vs oracle for one configuration; returns a dict of achieved errors.  ``nms_images``: the
    informational NMS set overlap (the C oracle NMS on dense init-V boxes takes seconds per image) looks at the first
    nms_images images only; logits and boxes are always compared for every image."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    model = y3.ParseModel.builtin_yolov3(C).init_weights(init, seed=seed)
    rng = np.random.default_rng(size * 1000 + B)
    x = rng.random((B, size, size, 3), dtype=np.float32)
    grids = [g.cpu().numpy() for g in model(torch.from_numpy(x).cuda())]
    torch.cuda.synchronize()
    ref = oracle_forward_chunked(model, x, chunk)
    anchors = configs.coco_anchors()
    out = {"init": init, "size": size, "B": B, "C": C, "heads": logits_errors(grids, ref)}
    bx, dg, dr = box_errors(grids, ref, anchors, C)
    out["boxes"] = bx
    # the GPU decode kernel on the GPU's own logits agrees with the oracle decode of those same logits (decode parity)
    dk = y3.yolo_decode([torch.from_numpy(g).cuda() for g in grids], anchors, C)
    out["decode_kernel_vs_oracle_abs"] = float(max(np.abs(a.cpu().numpy() - b).max() for a, b in zip(dk, dg)))
    if with_nms:
        k = B if nms_images is None else min(B, nms_images)
        out["nms"] = nms_set_overlap(tuple(a[:k] for a in dg), tuple(a[:k] for a in dr), 100, 0.5, score_thr)
        out["nms"]["images"] = k
    model.close()
    return out
