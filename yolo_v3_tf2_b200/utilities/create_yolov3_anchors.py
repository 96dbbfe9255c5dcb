"""Mirror of reference utilities/create_yolov3_anchors.py: k-means over the ground-truth box
(width, height) pairs -> anchors sorted by area, written as 'w, h' lines with '%10.5f'.

Host utility, not on the hot path (a few thousand boxes; scikit-learn's KMeans exactly as the reference uses it, so the
same ``random_state`` gives the same anchors as the reference would).  Dataset readers (tfrecords / COCO json) are out of
scope: the functions take the label arrays directly.

Reference quirk kept as is: the file is written in ASCENDING area order (create_yolov3_anchors.py:37-39, :115) while
``get_anchors`` hands row block 0 to the coarsest grid, for which the shipped COCO anchors file lists the LARGEST anchors
first; ``descending=True`` writes the order the decode expects.
"""
import os
import pathlib

import numpy as np


def sort_anchors(anchors):
    """create_yolov3_anchors.py:37-39."""
    return anchors[(anchors[:, 0] * anchors[:, 1]).argsort()]


def arrange_wh_array(labels):
    """create_yolov3_anchors.py:44-51: labels [..., >=4] = (xmin, ymin, xmax, ymax, ...) -> [n, 2] (w, h) of the boxes
    that are not all-zero padding (the reference keeps a pair only when BOTH w and h differ from 0)."""
    labels = np.asarray(labels, dtype=np.float32).reshape(-1, np.shape(labels)[-1])
    w_h = np.stack([labels[:, 2] - labels[:, 0], labels[:, 3] - labels[:, 1]], axis=-1)
    return w_h[np.all(w_h != 0.0, axis=-1)]


def creat_yolo_anchors(labels, n_clusters, random_state=None):
    """create_yolov3_anchors.py:54-65 (name as in the reference)."""
    from sklearn.cluster import KMeans
    w_h = arrange_wh_array(labels)
    if len(w_h) < n_clusters:
        raise ValueError(f"{len(w_h)} boxes cannot form {n_clusters} anchors")
    kmeans = KMeans(n_clusters=n_clusters, random_state=random_state, n_init=10)
    kmeans.fit(w_h)
    return sort_anchors(kmeans.cluster_centers_).astype(np.float32)


def save_anchors(anchors_out_file, anchors, descending=False):
    """create_yolov3_anchors.py:111-115: ``np.savetxt(file, anchors, delimiter=',', fmt='%10.5f')``."""
    head, _ = os.path.split(anchors_out_file)
    if head:
        pathlib.Path(head).mkdir(parents=True, exist_ok=True)
    a = np.asarray(anchors, dtype=np.float32)
    if descending:
        a = a[::-1]
    np.savetxt(anchors_out_file, a, delimiter=',', fmt='%10.5f')
