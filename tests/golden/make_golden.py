"""Generates tests/golden/*.npz -- small seeded input/output vectors produced by the oracle.

The reference ships no golden vectors for this path (its fixtures are 0-byte files) and TensorFlow cannot run here, so
these vectors pin OUR restatement (oracle/) rather than the reference binary: they guard the oracle against silent
regressions and give the GPU tests fixed known-answer inputs.  Where the two independent NMS formulations (TF's tiled
algorithm and the greedy loop) are both applicable, the script asserts they agree before writing.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import decode_oracle, net_oracle, nms_oracle  # noqa: E402
from y3_test_util import cluster_boxes, synth_grids  # noqa: E402
from yolo_v3_tf2_b200 import configs  # noqa: E402
import yolo_v3_tf2_b200 as y3  # noqa: E402


def golden_decode():
    anchors = configs.coco_anchors()
    grids = synth_grids(2, (2, 4, 8), 6, seed=123)
    b, c, p = decode_oracle.yolo_decode(grids, anchors, 6)
    cls, sc = decode_oracle.class_reduce(c, p)
    np.savez_compressed(os.path.join(HERE, "decode_small.npz"), g0=grids[0], g1=grids[1], g2=grids[2], anchors=anchors,
                        bboxes=b, conf=c, probs=p, cls=cls, scores=sc)


def golden_nms():
    out = {}
    cases = {
        "cluster15": cluster_boxes(1500, 15, 0) + (100, 0.5, 0.1),
        "cluster40_iou03": cluster_boxes(2000, 40, 1) + (100, 0.3, 0.1),
        "cluster150_iou07": cluster_boxes(1200, 150, 2) + (100, 0.7, 0.5),
        "few": cluster_boxes(300, 4, 3) + (100, 0.5, 0.1),
        "none_pass": (cluster_boxes(200, 4, 4)[0], cluster_boxes(200, 4, 4)[1] * 0.05, 50, 0.5, 0.1),
    }
    b, s = cluster_boxes(900, 20, 5)
    s[::5] = s[2]
    b[7] = b[8]
    cases["ties_dups"] = (b, s, 100, 0.5, 0.1)
    for name, (b, s, mx, iou, sthr) in cases.items():
        t = nms_oracle.nms_padded_tiled(b, s, mx, iou, sthr)
        g = nms_oracle.nms_padded_greedy(b, s, mx, iou, sthr)
        assert np.array_equal(t[0], g[0]) and t[1] == g[1], name
        out[f"{name}_boxes"], out[f"{name}_scores"] = b, s
        out[f"{name}_params"] = np.array([mx, iou, sthr], np.float64)
        out[f"{name}_sel"], out[f"{name}_nv"] = t[0], np.int32(t[1])
    np.savez_compressed(os.path.join(HERE, "nms_cases.npz"), **out)


def golden_net():
    """YOLOv3 (80 classes) on one 64x64 image: weights are regenerated from the seed, only x and the expected grids
    are stored; plus a handful of intermediate activations' statistics."""
    for init, seed in (("variance", 3), ("keras", 0)):
        m = y3.ParseModel.builtin_yolov3(80).init_weights(init, seed=seed)
        x = np.random.default_rng(42).random((1, 64, 64, 3), dtype=np.float32)
        outs = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, x)
        np.savez_compressed(os.path.join(HERE, f"net64_{init}.npz"), x=x, seed=np.int64(seed), g0=outs[0], g1=outs[1], g2=outs[2],
                            w0_sum=np.float64(m._params[0].kernel.astype(np.float64).sum()),
                            w74_sum=np.float64(m._params[74].kernel.astype(np.float64).sum()))


def golden_next_rows():
    """Small vectors for the 'next' rows: YOLOv3-tiny forward (maxpool), pre-processing, evaluation counters."""
    from oracle import evaluate_oracle, preprocess_oracle
    m = y3.ParseModel.builtin_yolov3_tiny(80).init_weights("variance", seed=11)
    x = np.random.default_rng(7).random((1, 64, 64, 3), dtype=np.float32)
    outs = net_oracle.forward(m.graph.layers, m.graph.outputs, m._params, x)
    np.savez_compressed(os.path.join(HERE, "tiny64_variance.npz"), x=x, seed=np.int64(11), g0=outs[0], g1=outs[1])
    rng = np.random.default_rng(9)
    u8 = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    f32 = rng.random((48, 30, 3), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "preprocess_small.npz"), u8=u8, f32=f32,
                        u8_resized_div255=preprocess_oracle.resize(u8, 32, divide_by_255=True),
                        f32_resized=preprocess_oracle.resize(f32, 40),
                        u8_aspect=preprocess_oracle.resize_image(u8, 32, 48),
                        f32_aspect=preprocess_oracle.resize_image(f32, 64, 64))
    nclasses = 5
    det = rng.random((6, 12, 2)).astype(np.float32)
    det = np.concatenate([det, det + rng.random((6, 12, 2)).astype(np.float32) * 0.3 + 0.05], -1)
    gt = det[:, :8] + rng.normal(0, 0.02, (6, 8, 4)).astype(np.float32)
    dcls = rng.integers(0, nclasses, (6, 12)).astype(np.int64)
    gcls = np.where(rng.random((6, 8)) < 0.7, dcls[:, :8], rng.integers(0, nclasses, (6, 8))).astype(np.int32)
    dn = rng.integers(0, 13, 6).astype(np.int32)
    gn = rng.integers(0, 9, 6).astype(np.int32)
    c = evaluate_oracle.new_counters(nclasses)
    for b in range(6):
        evaluate_oracle.evaluate(c, nclasses, 0.5, det[b, :dn[b]], dcls[b, :dn[b]], gt[b, :gn[b]], gcls[b, :gn[b]])
    np.savez_compressed(os.path.join(HERE, "evaluate_small.npz"), det=det, dcls=dcls, dn=dn, gt=gt, gcls=gcls, gn=gn,
                        nclasses=np.int64(nclasses), preds=c["preds"], gts=c["gts"], tp=c["tp"], fp=c["fp"], fn=c["fn"],
                        examples=np.int64(c["examples"]))


if __name__ == "__main__":
    golden_decode()
    golden_nms()
    golden_net()
    golden_next_rows()
    print("golden vectors written to", HERE)
