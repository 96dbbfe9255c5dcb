"""Weight layouts of the reference (host side): Darknet file order (convert.py:36-74, 93-137) and Keras order."""
import os

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


def test_tf_checkpoint_roundtrip_object_based(tmp_path):
    """TensorFlow checkpoint (tensor bundle) written with the reference model's object-based variable names and read
    back without TensorFlow (inference.py:102 ``model.load_weights(prefix).expect_partial()``)."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import tf_checkpoint as tc
    m = y3.ParseModel.builtin_yolov3_tiny(80).init_weights("variance", seed=4)
    prefix = str(tmp_path / "yolov3_train_8.tf")
    m.save_weights(prefix)
    idx = tc.read_index(prefix + ".index")
    assert idx[""]["num_shards"] == 1 and idx[""]["endianness"] == 0
    # Keras numbers Model.layers by decreasing depth: backbone(7 convs)=0, neck0(1)=1, neck1(1)=2, head0(2)=3, head1(2)=4;
    # head0's bias-only conv is the third layer with weights of sub-model 3
    assert "layer_with_weights-0/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE" in idx
    assert "layer_with_weights-0/layer_with_weights-1/moving_variance/.ATTRIBUTES/VARIABLE_VALUE" in idx
    assert "layer_with_weights-3/layer_with_weights-2/bias/.ATTRIBUTES/VARIABLE_VALUE" in idx
    assert idx["layer_with_weights-3/layer_with_weights-2/bias/.ATTRIBUTES/VARIABLE_VALUE"]["shape"] == [255]
    assert idx["layer_with_weights-2/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"]["shape"] == [1, 1, 256, 128]
    e = idx["layer_with_weights-0/layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE"]
    assert e["shape"] == [3, 3, 3, 16] and e["dtype"] == 1 and e["size"] == 3 * 3 * 3 * 16 * 4
    for path in (prefix, prefix + ".index"):
        m2 = y3.ParseModel.builtin_yolov3_tiny(80)
        m2.load_weights(path).expect_partial()
        for a, b in zip(m.get_weights(), m2.get_weights()):
            np.testing.assert_array_equal(a, b)
    # a model with a different head does not silently accept the file
    with pytest.raises(ValueError, match="kernel shape"):
        y3.ParseModel.builtin_yolov3_tiny(3).load_weights(prefix)
    with pytest.raises(FileNotFoundError):
        y3.ParseModel.builtin_yolov3_tiny(80).load_weights(str(tmp_path / "missing.tf"))


def test_keras_depth_order_of_sub_models():
    """``layer_with_weights-<i>`` follows Model.layers, which Keras sorts by decreasing depth -- yolov3:
    backbone, neck0, neck1, neck2, head0, head1, head2 (NOT the sub_models_configs order); inside a sub-model conv and
    batch-normalization layers alternate."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import graph
    for thin in (False, True):
        g = y3.ParseModel.builtin_yolov3(80, thin_heads=thin).graph
        slots = graph.keras_weight_slots(g)
        order = {}
        for ci, (i, jc, jb) in enumerate(slots):
            order.setdefault(i, g.layers[g.conv_layers[ci]].sub_model)
        assert [order[i] for i in sorted(order)] == ["backbone", "neck0", "neck1", "neck2", "head0", "head1", "head2"]
        # the backbone: 52 convs each followed by its BN -> j = 0..103; heads: conv, bn, biased conv
        assert slots[0] == (0, 0, 1) and slots[51] == (0, 102, 103)
        assert slots[57] == (4, 0, 1) and slots[58] == (4, 2, None) and slots[74] == (6, 2, None)
        assert slots[59] == (2, 0, 1)          # neck1's first conv (creation index 59) lives in sub-model 2
    gt = y3.ParseModel.builtin_yolov3_tiny(80).graph
    st = graph.keras_weight_slots(gt)
    assert [s[0] for s in st] == [0] * 7 + [1, 3, 3, 2, 4, 4]


def test_tf_checkpoint_handmade_fixture():
    """tests/golden/tf_bundle_handmade.* was assembled byte by byte by tests/golden/make_tf_bundle_fixture.py, an
    encoder that shares no code with yolo_v3_tf2_b200.tf_checkpoint (prefix-compressed keys with restart interval 4,
    snappy-compressed blocks, a 9-block index with shortest-separator keys, a DT_STRING object-graph entry, Keras
    depth-ordered variable names written out by hand).  The reader must return exactly the generator's arrays."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_tf_bundle_fixture as fx
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import tf_checkpoint as tc
    idx = tc.read_index(fx.PREFIX + ".index")
    assert len(idx) == 57 and idx[""]["num_shards"] == 1
    assert idx["_CHECKPOINTABLE_OBJECT_GRAPH"]["dtype"] == 7
    back = tc.read_checkpoint(fx.PREFIX)
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" not in back and int(back["save_counter" + fx.SUFFIX]) == 8
    m = y3.ParseModel().build_model(None, fx.SMALL_MODEL["sub_models_configs"], "head", nclasses=fx.NCLASSES,
                                    layer_lists=fx.SMALL_LAYERS)
    m.load_weights(fx.PREFIX).expect_partial()
    for p, (k, b, g, be, mu, v) in zip(m._params, fx.expected_params()):
        np.testing.assert_array_equal(p.kernel, k)
        if b is not None:
            np.testing.assert_array_equal(p.bias, b)
        else:
            for got, want in ((p.gamma, g), (p.beta, be), (p.mean, mu), (p.var, v)):
                np.testing.assert_array_equal(got, want)
    # a corrupted tensor payload is caught by the per-tensor crc32c of the BundleEntryProto
    import shutil, tempfile
    d = tempfile.mkdtemp()
    for suf in (".index", ".data-00000-of-00001"):
        shutil.copy(fx.PREFIX + suf, os.path.join(d, "c" + suf))
    raw = bytearray(open(os.path.join(d, "c.data-00000-of-00001"), "rb").read())
    raw[100] ^= 0x40
    open(os.path.join(d, "c.data-00000-of-00001"), "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="crc32c"):
        tc.read_checkpoint(os.path.join(d, "c"))


def test_tf_checkpoint_format_details(tmp_path):
    """Many small tensors (several data blocks, prefix-compressed keys), name-based keys, other dtypes, a snappy
    compressed block and corruption checks."""
    from yolo_v3_tf2_b200 import tf_checkpoint as tc
    rng = np.random.default_rng(0)
    tensors = {f"scope/var_{i:04d}/weights": rng.standard_normal((i % 5 + 1, 3)).astype(np.float32) for i in range(300)}
    tensors["global_step"] = np.array(12345, np.int64)
    tensors["flags"] = np.array([True, False, True])
    prefix = str(tmp_path / "many")
    tc.write_checkpoint(prefix, tensors, block_entries=17)
    back = tc.read_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        np.testing.assert_array_equal(back[k], tensors[k])
        assert back[k].dtype == tensors[k].dtype
    # name-based (TF1-style) checkpoint of a conv + BN + biased conv
    shapes = [(3, 3, 8, True), (1, 8, 6, False)]
    named = {"conv2d/kernel": rng.standard_normal((3, 3, 3, 8)).astype(np.float32),
             "batch_normalization/gamma": np.ones(8, np.float32), "batch_normalization/beta": np.zeros(8, np.float32),
             "batch_normalization/moving_mean": np.full(8, 0.5, np.float32),
             "batch_normalization/moving_variance": np.full(8, 2.0, np.float32),
             "conv2d_1/kernel": rng.standard_normal((1, 1, 8, 6)).astype(np.float32), "conv2d_1/bias": np.arange(6, dtype=np.float32)}
    tc.write_checkpoint(str(tmp_path / "named"), named)
    ps = tc.params_from_checkpoint(str(tmp_path / "named"), shapes)
    np.testing.assert_array_equal(ps[0].kernel, named["conv2d/kernel"])
    np.testing.assert_array_equal(ps[0].var, named["batch_normalization/moving_variance"])
    np.testing.assert_array_equal(ps[1].bias, named["conv2d_1/bias"])
    # snappy raw format: literal + overlapping copy
    comp = bytes([11]) + bytes([(3 - 1) << 2]) + b"abc" + bytes([((8 - 4) << 2) | 1, 3])
    assert tc._snappy_decompress(comp) == b"abcabcabcab"
    # crc32c known answer (RFC 3720 test vector: 32 bytes of zeros)
    assert tc._crc32c(bytes(32)) == 0x8A9136AA
    # corruption: flipped byte inside the index block, bad magic
    raw = bytearray(open(prefix + ".index", "rb").read())
    bad = bytearray(raw)
    bad[-60] ^= 0xFF
    open(str(tmp_path / "bad.index"), "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tc.read_index(str(tmp_path / "bad.index"))
    bad = bytearray(raw)
    bad[-1] ^= 0xFF
    open(str(tmp_path / "bad2.index"), "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        tc.read_index(str(tmp_path / "bad2.index"))


def test_anchor_generation_roundtrip(tmp_path):
    """utilities/create_yolov3_anchors.py mirror: k-means over (w, h), ascending area order, '%10.5f' file format that
    get_anchors (core/utils.py:31-37) reads back as [n_scales, 3, 2]."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200.utilities import create_yolov3_anchors as ca
    rng = np.random.default_rng(0)
    centers = np.array([[0.05, 0.06], [0.1, 0.2], [0.2, 0.12], [0.3, 0.35], [0.5, 0.4], [0.7, 0.8]], np.float32)
    wh = np.concatenate([c + rng.normal(0, 0.004, (200, 2)).astype(np.float32) for c in centers])
    xy = rng.random((len(wh), 2)).astype(np.float32) * 0.1
    labels = np.concatenate([xy, xy + wh, np.ones((len(wh), 1), np.float32)], axis=1)
    labels = np.concatenate([labels, np.zeros((50, 5), np.float32)])          # zero padding rows are ignored
    anchors = ca.creat_yolo_anchors(labels.reshape(-1, 10, 5), 6, random_state=0)
    assert anchors.shape == (6, 2) and anchors.dtype == np.float32
    areas = anchors[:, 0] * anchors[:, 1]
    assert (np.diff(areas) > 0).all()
    np.testing.assert_allclose(anchors, centers[np.argsort(centers[:, 0] * centers[:, 1])], atol=0.01)
    path = str(tmp_path / "anchors" / "a.txt")
    ca.save_anchors(path, anchors)
    assert all(len(line.split(",")) == 2 for line in open(path).read().strip().splitlines())
    back = y3.get_anchors(path)
    assert back.shape == (2, 3, 2)
    np.testing.assert_allclose(back.reshape(6, 2), anchors, atol=1e-5)
    ca.save_anchors(path, anchors, descending=True)
    np.testing.assert_allclose(y3.get_anchors(path).reshape(6, 2), anchors[::-1], atol=1e-5)


def test_crc32c_against_tensorflow_written_bytes():
    """The checksum of the tensor-bundle reader (crc32c::Mask(crc32c::Value), reference inference.py:102 /
    train.py:93-104 via model.load_weights) checked on bytes TensorFlow itself wrote: the framing of the reference's own
    TFRecord files uses the same masked crc32c (tests/golden/make_tf_written_fixture.py copies the bytes verbatim)."""
    import ctypes as C
    import struct
    from yolo_v3_tf2_b200 import _lib
    from yolo_v3_tf2_b200 import tf_checkpoint as tc
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "tf_written_crc.npz"))
    headers = fx["headers"]
    assert headers.shape == (100, 12)
    lib = _lib.lib()
    lib.y3_crc32c.restype = C.c_uint32
    lib.y3_crc32c.argtypes = [C.c_uint32, C.c_void_p, C.c_int64]

    def native(b):
        a = np.frombuffer(b, np.uint8)
        return int(lib.y3_crc32c(0, a.ctypes.data_as(C.c_void_p), a.nbytes))

    for h in headers:
        h = h.tobytes()
        want, = struct.unpack("<I", h[8:12])
        assert tc._mask_crc(tc._crc32c(h[:8])) == want
        assert tc._mask_crc(native(h[:8])) == want
    rec = fx["record"].tobytes()
    n, = struct.unpack_from("<Q", rec, 0)
    assert len(rec) == 16 + n
    want, = struct.unpack_from("<I", rec, 12 + n)
    data = rec[12:12 + n]
    assert tc._mask_crc(tc._crc32c(data)) == want
    assert tc._mask_crc(native(data)) == want
    # incremental form used by the reader on split buffers
    assert tc._crc32c(data[1000:], tc._crc32c(data[:1000])) == tc._crc32c(data)
    # the reader's protobuf field parser (BundleEntryProto / BundleHeaderProto) on a protobuf-library-written message:
    # the record is a tf.train.Example {features {feature: map<string, Feature>}} (reference core/load_tfrecords.py:20-31)
    (fn, wt, feats), = list(tc._pb_fields(data))
    assert (fn, wt) == (1, 2)
    keys = {}
    for fn, wt, kv in tc._pb_fields(feats):
        assert (fn, wt) == (1, 2)
        (kf, _, key), (vf, _, val) = list(tc._pb_fields(kv))
        assert (kf, vf) == (1, 2)
        keys[key.decode()] = list(tc._pb_fields(val))
    assert set(keys) == {"image/encoded", "image/object/bbox/xmin", "image/object/bbox/ymin", "image/object/bbox/xmax",
                         "image/object/bbox/ymax", "image/object/class/label"}
    (_, _, jpeg), = list(tc._pb_fields(keys["image/encoded"][0][2]))            # BytesList.value
    assert jpeg[:3] == b"\xff\xd8\xff"                                            # a JPEG stream, intact
    (_, _, packed), = list(tc._pb_fields(keys["image/object/bbox/xmin"][0][2]))  # FloatList.value, packed
    xmin = np.frombuffer(packed, "<f4")
    assert xmin.size == 3 and np.all((xmin >= 0) & (xmin <= 1))
