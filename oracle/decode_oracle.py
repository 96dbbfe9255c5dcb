"""Oracle for yolo_decode: numpy line-for-line restatement of reference core/yolo_decode_layer.py:4-36.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned (TensorFlow unavailable): sigmoid/exp are numpy's
float32 routines, which can differ from Eigen's by a few ulp -- hence the decode tolerance in tests/.
"""
import numpy as np


def _sigmoid(x):
    x = np.asarray(x, np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def _arrange_bbox(xy, wh):
    # yolo_decode_layer.py:4-12 ; grid[i, j] = (j, i) -> (x offset = column, y offset = row)
    gh, gw = xy.shape[1:3]
    gx, gy = np.meshgrid(np.arange(gw), np.arange(gh))
    grid = np.stack([gx, gy], axis=-1)[:, :, None, :].astype(np.float32)          # [gh, gw, 1, 2]
    # tf.cast(grid_size) is (gh, gw) and divides (x, y) elementwise exactly like the reference does
    xy = (xy + grid) / np.array([gh, gw], dtype=np.float32)
    xy_min = xy - wh / np.float32(2)
    xy_max = xy + wh / np.float32(2)
    return np.concatenate([xy_min, xy_max], axis=-1).astype(np.float32)


def yolo_decode(model_output_grids, anchors_table, nclasses):
    """-> (bboxes [B,N,4], confidence [B,N,1], class_probs [B,N,C]) float32."""
    anchors_table = np.asarray(anchors_table, np.float32)
    boxes, confs, probs = [], [], []
    for g, anchors in zip(model_output_grids, anchors_table):
        g = np.asarray(g, np.float32)
        B = g.shape[0]
        xy, wh, obj, cls = g[..., 0:2], g[..., 2:4], g[..., 4:5], g[..., 5:5 + nclasses]   # :16-17
        xy = _sigmoid(xy)                                                                  # :19
        obj = _sigmoid(obj)                                                                # :20
        cls = _sigmoid(cls)                                                                # :21
        box = _arrange_bbox(xy, (np.exp(wh) * anchors.reshape(1, 1, 1, 3, 2)).astype(np.float32))   # :23
        boxes.append(box.reshape(B, -1, 4))                                                # :26-28
        confs.append(obj.reshape(B, -1, 1))                                                # :30-32
        probs.append(cls.reshape(B, -1, nclasses))                                         # :34
    return np.concatenate(boxes, 1), np.concatenate(confs, 1), np.concatenate(probs, 1)


def class_reduce(confidence, class_probs):
    """reference core/yolo_nms.py:18-24 -> (class_indices int64 [B,N], scores float32 [B,N])."""
    class_indices = np.argmax(class_probs, axis=-1).astype(np.int64)
    max_p = np.max(class_probs, axis=-1, keepdims=True)
    scores = (np.asarray(confidence, np.float32) * max_p.astype(np.float32)).squeeze(-1)
    return class_indices, scores.astype(np.float32)
