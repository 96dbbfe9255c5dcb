"""Weight layouts of the reference (host side): Darknet file order (convert.py:36-74, 93-137) and Keras order."""
import numpy as np
import pytest


def test_darknet_roundtrip_and_layout(tmp_path):
    from yolo_v3_tf2_b200 import weights as wm
    shapes = [(3, 3, 8, True), (1, 8, 4, True), (1, 4, 6, False)]
    params = wm.init_variance_preserving(shapes, seed=1)
    path = str(tmp_path / "t.weights")
    wm.write_darknet_weights(path, params)
    raw = np.fromfile(path, dtype=np.float32)
    hdr = 5
    # first conv: [beta, gamma, mean, var] then kernel as (Cout, Cin, kh, kw)
    np.testing.assert_array_equal(raw[hdr:hdr + 8], params[0].beta)
    np.testing.assert_array_equal(raw[hdr + 8:hdr + 16], params[0].gamma)
    np.testing.assert_array_equal(raw[hdr + 32:hdr + 32 + 8 * 3 * 9].reshape(8, 3, 3, 3).transpose(2, 3, 1, 0), params[0].kernel)
    back = wm.read_darknet_weights(path, shapes)
    for a, b in zip(params, back):
        for x, y in zip(a.as_list(), b.as_list()):
            np.testing.assert_array_equal(x, y)
    with pytest.raises(ValueError, match="truncated"):
        wm.read_darknet_weights(path, shapes + [(1, 6, 6, True)])


def test_keras_order_roundtrip():
    import yolo_v3_tf2_b200 as y3
    m = y3.ParseModel.builtin_yolov3(80).init_weights("keras", seed=0)
    w = m.get_weights()
    assert len(w) == 72 * 5 + 3 * 2
    assert w[0].shape == (3, 3, 3, 32) and w[1].shape == (32,)           # conv2d kernel, then its BN gamma
    m2 = y3.ParseModel.builtin_yolov3(80)
    m2.set_weights(w)
    for a, b in zip(m.get_weights(), m2.get_weights()):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        m2.set_weights(w[:-1])
    # Keras defaults: BN identity statistics, zero head bias
    assert np.all(m._params[0].gamma == 1) and np.all(m._params[0].var == 1) and np.all(m._params[58].bias == 0)


def test_model_without_weights_or_gpu_fails_loudly():
    import torch
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    m = y3.ParseModel.builtin_yolov3(80)
    with pytest.raises(_lib.Y3Error, match="no weights"):
        m(np.zeros((1, 64, 64, 3), np.float32))
    if not torch.cuda.is_available():
        m.init_weights("keras")
        with pytest.raises(_lib.Y3Error, match="no CPU fallback|no CUDA"):
            m(np.zeros((1, 64, 64, 3), np.float32))
