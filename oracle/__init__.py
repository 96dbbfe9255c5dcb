"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's hot path (ronen-halevy/yolo-v3-tf2).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package, and only as the checker or the reported CPU
baseline -- never as part of the product path (yolo_v3_tf2_b200/ must not import it).

PARITY UNPINNED: the reference's arithmetic lives in TensorFlow 2.8.1 / Keras 2.8.0 (requirements.txt:16,41), which
is not installed here and cannot be (no network, Python 3.12), and the reference ships no golden vectors for this
path (its fixtures are 0-byte files).  The restatement is therefore cross-checked only against
itself: two independent NMS formulations (TF's tiled algorithm and a plain greedy loop, plus a C port) must agree on
every vector, and the conv stack is checked in float64 against float32.
"""
