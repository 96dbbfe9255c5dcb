"""Extracts TensorFlow-WRITTEN known answers for the masked crc32c that the TF checkpoint reader (f-1) depends on.

The reference's checkpoints are 0-byte placeholders, but its TFRecord files
(/root/reference/datasets/shapes/three_mixed_shapes/*/input/tfrecords/*/*.tfrec) were written by TensorFlow's
RecordWriter, whose framing uses the same checksum as the tensor bundle (lib/hash/crc32c.h: crc32c::Mask(crc32c::Value)):

    uint64 length | uint32 masked_crc32c(length bytes) | data[length] | uint32 masked_crc32c(data)

This script copies, byte for byte and without recomputing anything,
  * the 12-byte header (length + its masked crc) of every record of one file     -> `headers` [R, 12] uint8
  * the smallest complete record (header + data + data crc, ~4.5 KB)              -> `record`  uint8
into tests/golden/tf_written_crc.npz.  tests/test_weights.py checks tf_checkpoint._crc32c / _mask_crc and the C-ABI's
y3_crc32c against them: these are bytes TensorFlow produced, not bytes this repository's writer produced.

Run in the build container (the reference is not present on the GPU box):  python tests/golden/make_tf_written_fixture.py
"""
import os
import struct

import numpy as np

SRC = "/root/reference/datasets/shapes/three_mixed_shapes/white/input/tfrecords/test/file_00_100.tfrec"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    buf = open(SRC, "rb").read()
    pos, spans = 0, []
    while pos < len(buf):
        n, = struct.unpack_from("<Q", buf, pos)
        spans.append((pos, n))
        pos += 16 + n
    assert pos == len(buf)
    headers = np.stack([np.frombuffer(buf[p:p + 12], np.uint8) for p, _ in spans])
    p, n = min(spans, key=lambda s: s[1])
    record = np.frombuffer(buf[p:p + 16 + n], np.uint8)
    np.savez_compressed(os.path.join(HERE, "tf_written_crc.npz"), headers=headers, record=record,
                        source=np.array(SRC.replace("/root/reference/", "")))
    print(len(spans), "records; smallest", n, "bytes")


if __name__ == "__main__":
    main()
