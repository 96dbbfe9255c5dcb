"""TensorFlow checkpoint ("tensor bundle") reader -- and a writer for round-trip tests -- without TensorFlow.

The reference saves and restores weights with ``model.save_weights(prefix)`` / ``model.load_weights(prefix)``
(train.py:76-78, 93-104; inference.py:102), i.e. TF2 object-based checkpoints ``<prefix>.index`` +
``<prefix>.data-00000-of-00001``.  TensorFlow is not available here, so the format is read directly.  It is restated
from the published TensorFlow sources (tensorflow/core/util/tensor_bundle/tensor_bundle.cc,
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/io/table_format.txt -- the LevelDB table format):

  .index  = sorted string table:  [data block]* [metaindex block] [index block] [48-byte footer]
            footer      = metaindex BlockHandle, index BlockHandle (varint64 offset, varint64 size), zero padding to
                          40 bytes, magic 0xdb4775248b80fb57 (little endian)
            block       = entries, uint32 restart offsets[], uint32 num_restarts; followed in the file by a 1-byte
                          compression type (0 none, 1 snappy) and a 4-byte masked crc32c
            entry       = varint32 shared_key_bytes, varint32 unshared_key_bytes, varint32 value_bytes, key delta, value
            index block = (separator key >= last key of a data block) -> BlockHandle of that data block
            key ""      -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
            key <name>  -> BundleEntryProto  {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size,
                                              6: crc32c (fixed32)}
  .data-SSSSS-of-NNNNN = raw little-endian tensor bytes at [offset, offset + size)

Keras object-based variable names: ``layer_with_weights-<i>/layer_with_weights-<j>/<var>/.ATTRIBUTES/VARIABLE_VALUE``
where i enumerates the sub-models with weights in ``model.layers`` order and j the layers with weights inside the
sub-model, also in ``layers`` order; <var> is kernel / bias / gamma / beta / moving_mean / moving_variance.  Keras sorts
``Model.layers`` by decreasing depth (longest path to an output), NOT by creation order, so the sub-models of yolov3 are
numbered backbone, neck0, neck1, neck2, head0, head1, head2 (yolov3-tiny: backbone, neck0, neck1, head0, head1); the
mapping conv -> (i, j) is computed from the graph by ``graph.keras_weight_slots`` and passed in as ``slots``.
Name-based (TF1-style) keys ``conv2d_<n>/kernel``, ``batch_normalization_<n>/gamma`` are accepted as well.

NOT validated against a file written by TensorFlow (none can be produced here and the reference ships only 0-byte
checkpoint stubs): tests round-trip through the writer below, which follows the same specification.
"""
import os
import re
import struct

import numpy as np

from .weights import ConvParams

_MAGIC = 0xdb4775248b80fb57
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
       19: np.float16}
_DT_INV = {np.dtype(v): k for k, v in _DT.items()}
_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_CRC_VERIFY_MAX = 1 << 16


# ---------------------------------------------------------------- varints / protobuf wire format
def _varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf):
    """Yields (field number, wire type, value) of one protobuf message (varint, fixed32/64 and length-delimited)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for fn, wt, v in _pb_fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:          # repeated Dim
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    e["shape"].append(size)
        elif fn == 3:
            e["shard_id"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc32c"] = v
        elif fn == 7:
            e["sliced"] = True
    return e


# ---------------------------------------------------------------- crc32c (Castagnoli), only used on small buffers
_CRC_TABLE = None


def _tensor_crc(arr):
    """masked crc32c of a tensor's bytes as stored in BundleEntryProto.crc32c; None if it cannot be computed cheaply"""
    a = np.ascontiguousarray(arr)
    try:
        from . import _lib
        import ctypes as C
        return _mask_crc(int(_lib.lib().y3_crc32c(0, a.ctypes.data_as(C.c_void_p), a.nbytes)))
    except Exception:
        if a.nbytes <= _CRC_VERIFY_MAX:
            return _mask_crc(_crc32c(a.tobytes()))
        return None


def _crc32c(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------- snappy (raw format) decoder for compressed blocks
def _snappy_decompress(buf):
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                       # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                 # may overlap
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


# ---------------------------------------------------------------- table reader
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype, crc = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and _mask_crc(_crc32c(raw[:size + 1])) != crc:
        raise ValueError("table block checksum mismatch")
    if ctype == 1:
        body = _snappy_decompress(body)
    elif ctype != 0:
        raise ValueError(f"unknown block compression type {ctype}")
    return body


def _block_entries(body):
    nrest = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * nrest
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(body, pos)
        unshared, pos = _varint(body, pos)
        vlen, pos = _varint(body, pos)
        key = key[:shared] + body[pos:pos + unshared]
        pos += unshared
        yield key, body[pos:pos + vlen]
        pos += vlen


def read_index(index_path):
    """{tensor name: entry dict} plus the header dict under the key ''."""
    out = {}
    with open(index_path, "rb") as f:
        f.seek(0, os.SEEK_END)
        fsize = f.tell()
        if fsize < 48:
            raise ValueError(f"{index_path}: not a tensor-bundle index (too small)")
        f.seek(fsize - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
            raise ValueError(f"{index_path}: not a tensor-bundle index (bad magic)")
        pos = 0
        _, pos = _varint(footer, pos)           # metaindex offset
        _, pos = _varint(footer, pos)           # metaindex size
        ioff, pos = _varint(footer, pos)
        isize, pos = _varint(footer, pos)
        for _, handle in _block_entries(_read_block(f, ioff, isize)):
            boff, p2 = _varint(handle, 0)
            bsize, _ = _varint(handle, p2)
            for key, val in _block_entries(_read_block(f, boff, bsize)):
                name = key.decode("utf-8", "replace")
                if name == "":
                    hdr = {"num_shards": 1, "endianness": 0}
                    for fn, _, v in _pb_fields(val):
                        if fn == 1:
                            hdr["num_shards"] = v
                        elif fn == 2:
                            hdr["endianness"] = v
                    out[""] = hdr
                else:
                    out[name] = _parse_entry(val)
    return out


def read_checkpoint(prefix, names=None):
    """{name: numpy array} of every numeric tensor (or of ``names``) in the bundle ``<prefix>.index`` /
    ``<prefix>.data-*``.  String tensors (the object graph) are skipped."""
    prefix = _strip_prefix(prefix)
    idx = read_index(prefix + ".index")
    hdr = idx.pop("", {"num_shards": 1, "endianness": 0})
    if hdr.get("endianness", 0) != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    nshards = hdr.get("num_shards", 1)
    files = {}
    out = {}
    try:
        for name, e in idx.items():
            if names is not None and name not in names:
                continue
            if e["dtype"] not in _DT or e["sliced"]:
                continue
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open(f"{prefix}.data-{sid:05d}-of-{nshards:05d}", "rb")
            f = files[sid]
            f.seek(e["offset"])
            a = np.fromfile(f, dtype=_DT[e["dtype"]], count=int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1)
            if a.nbytes != e["size"]:
                raise ValueError(f"{name}: truncated tensor data")
            # per-tensor checksum (BundleEntryProto.crc32c, masked), computed by the shared library's y3_crc32c (pure
            # Python only for small tensors when the library is not built); table blocks are always verified
            if e["crc32c"] is not None:
                got = _tensor_crc(a)
                if got is not None and got != e["crc32c"]:
                    raise ValueError(f"{name}: tensor data crc32c mismatch")
            out[name] = a.reshape(e["shape"])
    finally:
        for f in files.values():
            f.close()
    return out


def _strip_prefix(path):
    path = str(path)
    for suf in (".index",):
        if path.endswith(suf):
            return path[:-len(suf)]
    m = re.match(r"^(.*)\.data-\d{5}-of-\d{5}$", path)
    return m.group(1) if m else path


def is_checkpoint(path):
    return os.path.exists(_strip_prefix(path) + ".index")


# ---------------------------------------------------------------- variables -> conv parameters
_VARS = ("kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance")


def _object_slots(tensors):
    """{(i, j, ...): {var: array}} for the object-based keys ``layer_with_weights-<i>/layer_with_weights-<j>/<var>``."""
    obj = {}
    for name, arr in tensors.items():
        if not name.endswith(_SUFFIX):
            continue
        parts = name[:-len(_SUFFIX)].split("/")
        if parts[-1] not in _VARS or len(parts) < 2:
            continue
        path = []
        for comp in parts[:-1]:
            m = re.fullmatch(r"layer_with_weights-(\d+)", comp)
            if not m:
                path = None
                break
            path.append(int(m.group(1)))
        if path:
            obj.setdefault(tuple(path), {})[parts[-1]] = arr
    return obj


def _layer_slots(tensors):
    """Ordered list of {var: array} per layer with weights, from object-based or name-based keys (creation order is
    assumed for object-based keys: only right when the caller has no graph to derive the Keras order from)."""
    obj = {}
    for name, arr in tensors.items():
        if not name.endswith(_SUFFIX):
            continue
        parts = name[:-len(_SUFFIX)].split("/")
        if parts[-1] not in _VARS or len(parts) < 2:
            continue
        path = []
        ok = True
        for comp in parts[:-1]:
            m = re.fullmatch(r"layer_with_weights-(\d+)", comp)
            if not m:
                ok = False
                break
            path.append(int(m.group(1)))
        if ok:
            obj.setdefault(tuple(path), {})[parts[-1]] = arr
    if obj:
        return [obj[k] for k in sorted(obj)]
    named = {}
    for name, arr in tensors.items():
        parts = name.split("/")
        if len(parts) < 2 or parts[-1] not in _VARS:
            continue
        m = re.fullmatch(r"(conv2d|batch_normalization)(?:_(\d+))?", parts[-2])
        if not m:
            continue
        # conv2d_<n> precedes its batch_normalization_<m>: order by (kind-independent creation index is unknown), so
        # convs and BNs are ordered separately and zipped by the consumer
        named.setdefault((m.group(1), int(m.group(2) or 0)), {})[parts[-1]] = arr
    convs = [named[k] for k in sorted(k for k in named if k[0] == "conv2d")]
    bns = [named[k] for k in sorted(k for k in named if k[0] == "batch_normalization")]
    out, bi = [], 0
    for c in convs:
        out.append(c)
        if "bias" not in c and bi < len(bns):
            out.append(bns[bi])
            bi += 1
    return out


def params_from_checkpoint(prefix, conv_shapes, slots=None):
    """ConvParams list in conv creation order from a Keras checkpoint of the reference's model.  ``slots`` =
    ``graph.keras_weight_slots(g)``: per conv (i, j_conv, j_bn) of its ``layer_with_weights-<i>/layer_with_weights-<j>``
    keys (Keras depth order); without it the keys are taken in sorted order (name-based checkpoints)."""
    tensors = read_checkpoint(prefix)
    obj = _object_slots(tensors) if slots is not None else {}
    if obj:
        out = []
        for ci, ((k, cin, cout, bn), (i, jc, jb)) in enumerate(zip(conv_shapes, slots)):
            conv = obj.get((i, jc))
            if conv is None or "kernel" not in conv:
                raise ValueError(f"checkpoint has no kernel for conv {ci} (layer_with_weights-{i}/layer_with_weights-{jc})")
            kern = np.asarray(conv["kernel"], np.float32)
            if kern.shape != (k, k, cin, cout):
                raise ValueError(f"conv {ci}: checkpoint kernel shape {kern.shape}, model expects {(k, k, cin, cout)}")
            if bn:
                b = obj.get((i, jb))
                if b is None or "gamma" not in b:
                    raise ValueError(f"checkpoint has no batch-normalization variables for conv {ci}")
                vecs = [np.asarray(b[v], np.float32) for v in ("gamma", "beta", "moving_mean", "moving_variance")]
                if any(v.shape != (cout,) for v in vecs):
                    raise ValueError(f"conv {ci}: batch-normalization vectors have the wrong shape")
                out.append(ConvParams(kern, gamma=vecs[0], beta=vecs[1], mean=vecs[2], var=vecs[3]))
            else:
                if "bias" not in conv:
                    raise ValueError(f"conv {ci}: checkpoint has no bias")
                out.append(ConvParams(kern, bias=np.asarray(conv["bias"], np.float32)))
        return out
    slots = _layer_slots(tensors)
    out, si = [], 0
    for ci, (k, cin, cout, bn) in enumerate(conv_shapes):
        if si >= len(slots) or "kernel" not in slots[si]:
            raise ValueError(f"checkpoint has no kernel for conv {ci}")
        conv = slots[si]
        si += 1
        kern = np.asarray(conv["kernel"], np.float32)
        if kern.shape != (k, k, cin, cout):
            raise ValueError(f"conv {ci}: checkpoint kernel shape {kern.shape}, model expects {(k, k, cin, cout)}")
        if bn:
            if si >= len(slots) or "gamma" not in slots[si]:
                raise ValueError(f"checkpoint has no batch-normalization variables for conv {ci}")
            b = slots[si]
            si += 1
            vecs = [np.asarray(b[v], np.float32) for v in ("gamma", "beta", "moving_mean", "moving_variance")]
            if any(v.shape != (cout,) for v in vecs):
                raise ValueError(f"conv {ci}: batch-normalization vectors have the wrong shape")
            out.append(ConvParams(kern, gamma=vecs[0], beta=vecs[1], mean=vecs[2], var=vecs[3]))
        else:
            if "bias" not in conv:
                raise ValueError(f"conv {ci}: checkpoint has no bias")
            out.append(ConvParams(kern, bias=np.asarray(conv["bias"], np.float32)))
    return out


# ---------------------------------------------------------------- writer (tests / exporting to the reference)
def _pb_varint_field(fn, v):
    return _put_varint(fn << 3) + _put_varint(v)


def _pb_bytes_field(fn, b):
    return _put_varint((fn << 3) | 2) + _put_varint(len(b)) + b


def _build_block(items, restart_interval=16):
    body, restarts, last = bytearray(), [], b""
    for i, (key, val) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_checkpoint(prefix, tensors, block_entries=64):
    """Write {name: numpy array} as a single-shard tensor bundle (uncompressed blocks, like TensorFlow's BundleWriter)."""
    prefix = _strip_prefix(prefix)
    names = sorted(tensors, key=lambda s: s.encode())
    entries = []
    with open(prefix + ".data-00000-of-00001", "wb") as df:
        for name in names:
            a = np.ascontiguousarray(tensors[name])
            if a.dtype not in _DT_INV:
                raise ValueError(f"{name}: unsupported dtype {a.dtype}")
            off = df.tell()
            raw = a.tobytes()
            df.write(raw)
            shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in a.shape)
            crc = _tensor_crc(a)
            if crc is None:
                raise RuntimeError("writing a checkpoint needs liby3b200.so (y3_crc32c): TensorFlow verifies every tensor's crc32c")
            ent = (_pb_varint_field(1, _DT_INV[a.dtype]) + _pb_bytes_field(2, shape) + _pb_varint_field(4, off) +
                   _pb_varint_field(5, len(raw)) + _put_varint((6 << 3) | 5) + struct.pack("<I", crc))
            entries.append((name.encode(), ent))
    header = _pb_varint_field(1, 1) + _pb_varint_field(2, 0) + _pb_bytes_field(3, _pb_varint_field(1, 1))
    items = [(b"", header)] + entries
    with open(prefix + ".index", "wb") as f:
        index_items = []

        def put_block(body):
            off = f.tell()
            f.write(body)
            f.write(b"\x00" + struct.pack("<I", _mask_crc(_crc32c(body + b"\x00"))))
            return _put_varint(off) + _put_varint(len(body))

        for i in range(0, len(items), block_entries):
            chunk = items[i:i + block_entries]
            index_items.append((chunk[-1][0], put_block(_build_block(chunk))))
        meta = put_block(_build_block([]))
        index = put_block(_build_block(index_items, restart_interval=1))
        footer = meta + index
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
        f.write(footer)


def keras_variable_names(slots, conv_shapes):
    """Object-based names of the reference model's variables, per conv in creation order: [kernel, bias] or
    [kernel, gamma, beta, moving_mean, moving_variance].  ``slots`` = ``graph.keras_weight_slots(g)``."""
    names = []
    for (i, jc, jb), (_, _, _, bn) in zip(slots, conv_shapes):
        base = f"layer_with_weights-{i}/layer_with_weights-{jc}"
        if bn:
            bbase = f"layer_with_weights-{i}/layer_with_weights-{jb}"
            names.append([base + "/kernel" + _SUFFIX] + [bbase + "/" + v + _SUFFIX
                                                          for v in ("gamma", "beta", "moving_mean", "moving_variance")])
        else:
            names.append([base + "/kernel" + _SUFFIX, base + "/bias" + _SUFFIX])
    return names
