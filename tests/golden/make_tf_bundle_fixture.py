#!/usr/bin/env python
"""Hand-assembled TensorFlow tensor-bundle fixture (tests/golden/tf_bundle_handmade.*), written WITHOUT
yolo_v3_tf2_b200.tf_checkpoint: an independent encoder of the published formats, so the reader is not validated only by
its own writer.

Formats restated (TensorFlow 2.8 sources):
  * tensorflow/core/lib/io/table_format.txt, block_builder.cc, table_builder.cc, format.cc  (LevelDB-derived table:
    prefix-compressed keys with restart points, per-block trailer = 1-byte compression type + masked crc32c of
    contents+type, index block keyed by shortest separators, 48-byte footer with magic 0xdb4775248b80fb57)
  * snappy format_description.txt (raw format: varint length, literal / copy tags)
  * tensorflow/core/protobuf/tensor_bundle.proto (BundleHeaderProto, BundleEntryProto), tensor_shape.proto
  * tensorflow/core/util/tensor_bundle/tensor_bundle.cc: keys sorted bytewise, "" holds the header, tensor bytes in
    key order inside <prefix>.data-00000-of-00001, entry crc32c = masked crc32c of the tensor bytes
  * Keras object-based names: layer_with_weights-<i>/layer_with_weights-<j>/<var>/.ATTRIBUTES/VARIABLE_VALUE with
    <i>, <j> numbering Model.layers (sorted by decreasing depth, keras/engine/functional.py), plus the string tensor
    _CHECKPOINTABLE_OBJECT_GRAPH that every TF2 checkpoint carries

What the fixture exercises that the package's own writer does not: restart interval 4 (most keys share a prefix with
their predecessor), snappy-compressed data blocks next to uncompressed ones, 7-entry data blocks (multi-block index),
shortest-separator index keys (not equal to any real key), a DT_STRING entry, non-zero metaindex handle.

The model is a small 5-sub-model graph (SMALL_MODEL below; sub_models_configs order backbone, neck0, head0, neck1, head1)
whose Keras depth order is backbone, neck0, neck1, head0, head1 -- the (i, j) table KERAS_SLOTS is written out BY HAND.
usage: python tests/golden/make_tf_bundle_fixture.py
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PREFIX = os.path.join(HERE, "tf_bundle_handmade")
NCLASSES = 1


def conv(f, size, stride=1, bn=True, act="leaky"):
    d = {"type": "convolutional", "filters": f, "size": size, "stride": stride, "pad": 1, "activation": act}
    if bn:
        d["batch_normalize"] = 1
    return d


HEAD_F = "3*(5+nclasses)"
SMALL_MODEL = {
    "sub_models_configs": [
        {"name": "backbone", "layers_config_file": "backbone", "outputs_layers": [-2, -1]},
        {"name": "neck0", "layers_config_file": "neck0", "outputs_layers": [-1],
         "inputs": {"source": [{"name": "backbone", "entry_index": 1}]}},
        {"name": "head0", "layers_config_file": "head0", "outputs_layers": [-1],
         "inputs": {"source": [{"name": "neck0", "entry_index": 0}]}},
        {"name": "neck1", "layers_config_file": "neck1", "outputs_layers": [-1],
         "inputs": {"source": [{"name": "neck0", "entry_index": 0}, {"name": "backbone", "entry_index": 0}]}},
        {"name": "head1", "layers_config_file": "head1", "outputs_layers": [-1],
         "inputs": {"source": [{"name": "neck1", "entry_index": 0}]}},
    ],
    "output_stage": "head",
}
SMALL_LAYERS = {
    "backbone": [{"type": "route", "source": {"inputs": [0]}}, conv(8, 3), conv(8, 3, 2), conv(8, 1), conv(8, 3),
                 {"type": "shortcut", "from": -3, "activation": "linear"}, conv(16, 3, 2)],
    "neck0": [{"type": "route", "source": {"inputs": [0]}}, conv(8, 1)],
    "head0": [{"type": "route", "source": {"inputs": [0]}}, conv(16, 3), conv(HEAD_F, 1, bn=False, act="linear"),
              {"type": "yolo", "grid_size": 0}],
    "neck1": [{"type": "route", "source": {"inputs": [0]}}, conv(8, 1), {"type": "upsample", "stride": 2},
              {"type": "route", "source": {"layers": [-1], "inputs": [1]}}, conv(8, 1)],
    "head1": [{"type": "route", "source": {"inputs": [0]}}, conv(8, 3), conv(HEAD_F, 1, bn=False, act="linear"),
              {"type": "yolo", "grid_size": 0}],
}
# conv creation order -> (k, cin, cout, bn) and, BY HAND, its Keras slot (i, j_conv, j_bn):
#   Model.layers by decreasing depth: backbone(4) neck0(3) neck1(2) head0(1)... heads are the outputs (depth 0), neck1
#   feeds head1 only (depth 1), neck0 feeds head0 (1) and neck1 (2) -> depth 2, backbone -> 3.  Ties (head0, head1) keep the
#   order of the depth-first walk from the outputs [head0, head1].
CONVS = [
    # backbone = layer_with_weights-0
    (3, 3, 8, True, (0, 0, 1)), (3, 8, 8, True, (0, 2, 3)), (1, 8, 8, True, (0, 4, 5)), (3, 8, 8, True, (0, 6, 7)),
    (3, 8, 16, True, (0, 8, 9)),
    # neck0 = layer_with_weights-1
    (1, 16, 8, True, (1, 0, 1)),
    # head0 = layer_with_weights-3 (after neck1!)
    (3, 8, 16, True, (3, 0, 1)), (1, 16, 18, False, (3, 2, None)),
    # neck1 = layer_with_weights-2
    (1, 8, 8, True, (2, 0, 1)), (1, 16, 8, True, (2, 2, 3)),
    # head1 = layer_with_weights-4
    (3, 8, 8, True, (4, 0, 1)), (1, 8, 18, False, (4, 2, None)),
]
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def expected_params(seed=20261018):
    """[(kernel, bias | None, gamma, beta, mean, var | None)] in conv creation order, from one seeded generator."""
    rng = np.random.default_rng(seed)
    out = []
    for k, cin, cout, bn, _ in CONVS:
        kern = rng.standard_normal((k, k, cin, cout)).astype(np.float32) * np.float32(0.1)
        if bn:
            out.append((kern, None, rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.normal(0, 0.1, cout).astype(np.float32),
                        rng.normal(0, 0.1, cout).astype(np.float32), rng.uniform(0.5, 1.5, cout).astype(np.float32)))
        else:
            out.append((kern, rng.normal(0, 0.1, cout).astype(np.float32), None, None, None, None))
    return out


# ------------------------------------------------------------------ independent encoders
def varint(v):
    b = bytearray()
    while v >= 0x80:
        b.append((v & 0x7F) | 0x80)
        v >>= 7
    b.append(v)
    return bytes(b)


def crc32c_bitwise(data):
    """Castagnoli CRC, reflected polynomial 0x82F63B78, one bit at a time (deliberately not table driven)."""
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def crc32c_fast(data):
    """Same CRC via numpy-free slicing-by-1 on a lazily built table -- used for the (larger) tensor payloads; checked
    against the bitwise form below."""
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
        tbl.append(c)
    crc = 0xFFFFFFFF
    for byte in data:
        crc = tbl[(crc ^ byte) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def snappy_compress(data):
    """Greedy raw-snappy encoder: 4-byte hash matches -> copy with 2-byte offset (tag 10), everything else literals."""
    out = bytearray(varint(len(data)))
    table = {}
    i, lit_start, n = 0, 0, len(data)

    def flush_literal(end):
        nonlocal lit_start
        pos = lit_start
        while pos < end:
            ln = min(end - pos, 65536)
            if ln <= 60:
                out.append((ln - 1) << 2)
            elif ln <= 256:
                out.append(60 << 2)
                out.append(ln - 1)
            else:
                out.append(61 << 2)
                out.extend(struct.pack("<H", ln - 1))
            out.extend(data[pos:pos + ln])
            pos += ln
        lit_start = end

    while i + 4 <= n:
        key = data[i:i + 4]
        cand = table.get(key)
        table[key] = i
        if cand is not None and i - cand <= 0xFFFF:
            ln = 4
            while i + ln < n and ln < 64 and data[cand + ln] == data[i + ln]:
                ln += 1
            flush_literal(i)
            out.append(((ln - 1) << 2) | 2)
            out += struct.pack("<H", i - cand)
            i += ln
            lit_start = i
        else:
            i += 1
    flush_literal(n)
    return bytes(out)


def build_block(items, restart_interval):
    body, restarts, prev = bytearray(), [], b""
    for idx, (key, val) in enumerate(items):
        shared = 0
        if idx % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        body += varint(shared) + varint(len(key) - shared) + varint(len(val)) + key[shared:] + val
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def shortest_separator(a, b):
    """leveldb BytewiseComparator::FindShortestSeparator: a string s with a <= s < b, as short as possible."""
    i = 0
    while i < min(len(a), len(b)) and a[i] == b[i]:
        i += 1
    if i < min(len(a), len(b)) and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def pb_varint(fn, v):
    return varint((fn << 3) | 0) + varint(v)


def pb_bytes(fn, b):
    return varint((fn << 3) | 2) + varint(len(b)) + b


def pb_fixed32(fn, v):
    return varint((fn << 3) | 5) + struct.pack("<I", v)


def entry_proto(dtype, shape, offset, size, crc):
    shape_pb = b"".join(pb_bytes(2, pb_varint(1, d)) for d in shape)
    e = pb_varint(1, dtype) + pb_bytes(2, shape_pb)
    if offset:
        e += pb_varint(4, offset)          # proto3: zero-valued scalars are omitted (shard_id 0, offset 0)
    e += pb_varint(5, size) + pb_fixed32(6, crc)
    return e


def main():
    params = expected_params()
    tensors = {}
    for (k, cin, cout, bn, (i, jc, jb)), (kern, bias, gamma, beta, mean, var) in zip(CONVS, params):
        base = f"layer_with_weights-{i}/layer_with_weights-{jc}"
        tensors[base + "/kernel" + SUFFIX] = kern
        if bn:
            bb = f"layer_with_weights-{i}/layer_with_weights-{jb}"
            for nm, a in (("gamma", gamma), ("beta", beta), ("moving_mean", mean), ("moving_variance", var)):
                tensors[bb + "/" + nm + SUFFIX] = a
        else:
            tensors[base + "/bias" + SUFFIX] = bias
    # every TF2 checkpoint also stores the serialized object graph as a scalar DT_STRING tensor and Keras adds
    # save_counter; the reader must skip / tolerate them
    tensors["save_counter" + SUFFIX] = np.array(8, np.int64)
    keys = sorted(k.encode() for k in list(tensors) + ["_CHECKPOINTABLE_OBJECT_GRAPH"])
    data = bytearray()
    items = [(b"", pb_varint(1, 1) + pb_bytes(3, pb_varint(1, 1)))]   # BundleHeaderProto: num_shards 1, version{producer 1}
    for key in keys:
        name = key.decode()
        if name == "_CHECKPOINTABLE_OBJECT_GRAPH":
            payload = varint(5) + b"graph"             # string tensors: varint lengths, then the bytes
            off = len(data)
            data += payload
            items.append((key, entry_proto(7, [], off, len(payload), masked(crc32c_bitwise(payload)))))
            continue
        a = np.ascontiguousarray(tensors[name])
        raw = a.tobytes()
        off = len(data)
        data += raw
        dt = {np.dtype(np.float32): 1, np.dtype(np.int64): 9}[a.dtype]
        items.append((key, entry_proto(dt, list(a.shape), off, len(raw), masked(crc32c_fast(raw)))))
    assert crc32c_fast(b"123456789") == crc32c_bitwise(b"123456789") == 0xE3069283   # the CRC-32C check value

    out = bytearray()

    def put(body, compress):
        ctype = 0
        if compress:
            c = snappy_compress(body)
            if len(c) < len(body):
                body, ctype = c, 1
        off = len(out)
        out.extend(body)
        trailer = bytes([ctype])
        out.extend(trailer + struct.pack("<I", masked(crc32c_bitwise(body + trailer))))
        return varint(off) + varint(len(body)), ctype

    index_items, used_snappy = [], 0
    per_block = 7
    for bi, i0 in enumerate(range(0, len(items), per_block)):
        blk = items[i0:i0 + per_block]
        handle, ctype = put(build_block(blk, restart_interval=4), compress=(bi % 2 == 0))
        used_snappy += ctype
        nxt = items[i0 + per_block][0] if i0 + per_block < len(items) else None
        sep = shortest_separator(blk[-1][0], nxt) if nxt is not None else blk[-1][0] + b"\x00"
        index_items.append((sep, handle))
    assert used_snappy >= 2 and len(index_items) >= 5
    meta_handle, _ = put(build_block([], 1), compress=False)
    index_handle, _ = put(build_block(index_items, restart_interval=1), compress=False)
    footer = meta_handle + index_handle
    footer += bytes(40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    out += footer
    with open(PREFIX + ".index", "wb") as f:
        f.write(bytes(out))
    with open(PREFIX + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    print("wrote", PREFIX + ".index", len(out), "bytes;", len(data), "data bytes;", len(items), "entries,",
          len(index_items), "data blocks,", used_snappy, "snappy")


if __name__ == "__main__":
    main()
