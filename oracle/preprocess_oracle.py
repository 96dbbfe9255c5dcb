"""Oracle for the input pre-processing: numpy restatement of the TensorFlow ops the reference calls.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned: TensorFlow is unavailable; the bilinear kernel is
cross-checked against torch's independent implementation of the same half-pixel formula in tests/test_oracle.py.

  * tf.image.resize(img, size)  (inference.py:157-158, core/load_tfrecords.py:46): method bilinear, antialias False
    -> ResizeBilinear(half_pixel_centers=True).  TensorFlow 2.8 kernels/image/resize_bilinear_op.cc +
    image_resizer_state.h: scale = in / float(out); in_f = (i + 0.5) * scale - 0.5; lower = max(floor(in_f), 0);
    upper = min(ceil(in_f), in - 1); lerp = in_f - floor(in_f);
    top = tl + (tr - tl) * x_lerp; bottom = bl + (br - bl) * x_lerp; out = top + (bottom - top) * y_lerp   (float32)
  * preserve_aspect_ratio=True (core/utils.py:17-22): scale = min(th / h, tw / w); size = round(h * scale), round(w * scale)
  * tf.image.pad_to_bounding_box(img, (th - sh) // 2, (tw - sw) // 2, th, tw) (core/utils.py:25-27): zero padding
"""
import numpy as np


def _weights(out_size, in_size):
    scale = np.float32(in_size) / np.float32(out_size)
    i = np.arange(out_size, dtype=np.float32)
    f = (i + np.float32(0.5)) * scale - np.float32(0.5)
    fl = np.floor(f)
    lo = np.maximum(fl.astype(np.int64), 0)
    hi = np.minimum(np.ceil(f).astype(np.int64), in_size - 1)
    return lo, hi, (f - fl).astype(np.float32)


def resize_bilinear(img, out_h, out_w):
    """img: [H, W, C] uint8 or float32 -> [out_h, out_w, C] float32."""
    a = np.asarray(img).astype(np.float32)
    y0, y1, ly = _weights(out_h, a.shape[0])
    x0, x1, lx = _weights(out_w, a.shape[1])
    lx = lx[None, :, None]
    ly = ly[:, None, None]
    tl, tr = a[y0][:, x0], a[y0][:, x1]
    bl, br = a[y1][:, x0], a[y1][:, x1]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return (top + (bot - top) * ly).astype(np.float32)


def aspect_size(h, w, th, tw):
    sc = np.minimum(np.float32(th) / np.float32(h), np.float32(tw) / np.float32(w))
    return int(np.round(np.float32(h) * sc)), int(np.round(np.float32(w) * sc))


def resize_image(img, target_height, target_width):
    """core/utils.py:17-28."""
    oh, ow = aspect_size(img.shape[0], img.shape[1], target_height, target_width)
    r = resize_bilinear(img, oh, ow)
    out = np.zeros((target_height, target_width, img.shape[2]), np.float32)
    oy, ox = (target_height - oh) // 2, (target_width - ow) // 2
    out[oy:oy + oh, ox:ox + ow] = r
    return out


def resize(img, size, divide_by_255=False):
    """tf.image.resize(img, (size, size)) [/ 255]."""
    r = resize_bilinear(img, size, size)
    return (r / np.float32(255.0)).astype(np.float32) if divide_by_255 else r
