"""ctypes loader for oracle/nms_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liby3oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def nms(boxes, scores, max_boxes, iou_thr, score_thr):
    boxes = np.ascontiguousarray(boxes, np.float32)
    scores = np.ascontiguousarray(scores, np.float32)
    B, N = scores.shape
    sel = np.zeros((B, max_boxes), np.int32)
    nv = np.zeros((B,), np.int32)
    rc = lib().y3o_nms(boxes.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p), B, N, int(max_boxes),
                       C.c_float(iou_thr), C.c_float(score_thr), sel.ctypes.data_as(C.c_void_p),
                       nv.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return sel, nv


def decode(grids, anchors, nclasses):
    B = grids[0].shape[0]
    N = sum(g.shape[1] * g.shape[2] * 3 for g in grids)
    bboxes = np.zeros((B, N, 4), np.float32)
    conf = np.zeros((B, N, 1), np.float32)
    probs = np.zeros((B, N, nclasses), np.float32)
    off = 0
    anchors = np.ascontiguousarray(anchors, np.float32)
    for s, g in enumerate(grids):
        g = np.ascontiguousarray(g, np.float32)
        lib().y3o_decode_scale(g.ctypes.data_as(C.c_void_p), B, g.shape[1], g.shape[2], int(nclasses),
                               anchors[s].ctypes.data_as(C.c_void_p), N, off, bboxes.ctypes.data_as(C.c_void_p),
                               conf.ctypes.data_as(C.c_void_p), probs.ctypes.data_as(C.c_void_p))
        off += g.shape[1] * g.shape[2] * 3
    return bboxes, conf, probs


def class_reduce(conf, probs):
    probs = np.ascontiguousarray(probs, np.float32)
    conf = np.ascontiguousarray(conf, np.float32).reshape(probs.shape[:-1])
    nrec = int(np.prod(probs.shape[:-1]))
    scores = np.zeros(probs.shape[:-1], np.float32)
    cls = np.zeros(probs.shape[:-1], np.int64)
    lib().y3o_class_reduce(probs.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p), C.c_longlong(nrec),
                           int(probs.shape[-1]), scores.ctypes.data_as(C.c_void_p), cls.ctypes.data_as(C.c_void_p))
    return cls, scores
