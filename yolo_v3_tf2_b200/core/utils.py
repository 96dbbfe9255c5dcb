"""Mirror of the hot-path part of reference core/utils.py (anchors, class count) and of the image pre-processing that
feeds the model (``tf.image.resize`` call sites + ``resize_image``), which runs on the GPU here."""
import numpy as np


def get_anchors(anchors_file):
    """reference core/utils.py:31-37: ``loadtxt(delimiter=',')`` then ``reshape(-1, 3, 2)``; row block 0 belongs to the
    coarsest grid.  Values are fractions of the image side."""
    nanchors_per_scale = 3
    anchor_entry_size = 2
    anchors_table = np.loadtxt(anchors_file, dtype=float, delimiter=',')
    anchors_table = anchors_table.reshape(-1, nanchors_per_scale, anchor_entry_size)
    return anchors_table


def count_file_lines(filename):
    """reference core/utils.py:40-43 (``nclasses = len(lines)`` -- 38 for datasets/pets_breed.names)."""
    with open(filename, 'r') as fp:
        nlines = len(fp.readlines())
    return nlines


def _aspect_size(h, w, target_height, target_width):
    """tf.image.resize(preserve_aspect_ratio=True): scale = min(th / h, tw / w), new size = round(h * scale), round(w * scale)
    (float32 arithmetic, round half to even like tf.round)."""
    sh = np.float32(target_height) / np.float32(h)
    sw = np.float32(target_width) / np.float32(w)
    sc = np.minimum(sh, sw)
    return int(np.round(np.float32(h) * sc)), int(np.round(np.float32(w) * sc))


_DESC_RING = {}


def _descriptor_upload(rows, dev, slots=8):
    """Image descriptors -> device without a blocking pageable copy: a small ring of pinned host buffers per device, each
    guarded by an event so that a slot is not rewritten while its previous upload is still in flight."""
    import torch
    ring = _DESC_RING.setdefault(dev.index, {"i": 0, "bufs": [None] * slots, "events": [None] * slots})
    i = ring["i"] = (ring["i"] + 1) % slots
    n = len(rows)
    if ring["bufs"][i] is None or ring["bufs"][i].shape[0] < n:
        ring["bufs"][i] = torch.empty((max(n, 64), 10), dtype=torch.int64).pin_memory()
        ring["events"][i] = torch.cuda.Event()
    else:
        ring["events"][i].synchronize()
    host = ring["bufs"][i][:n]
    host.numpy()[:] = np.asarray(rows, dtype=np.int64)
    desc = host.to(dev, non_blocking=True)
    ring["events"][i].record(torch.cuda.current_stream())
    return desc


def preprocess_images(images, target_height, target_width, preserve_aspect_ratio=False, divide_by_255=False, out=None):
    """Resize a list of [H, W, 3] images (torch CUDA tensors, uint8 or float32; numpy arrays are copied to the GPU) to
    one float32 batch [B, target_height, target_width, 3] with ONE kernel launch.
    ``preserve_aspect_ratio=False``: ``tf.image.resize(img, (th, tw))`` (inference.py:157-158; with
    ``divide_by_255`` the tfrecord path core/load_tfrecords.py:46).  ``True``: ``resize_image`` (core/utils.py:17-28):
    aspect-preserving resize, then centred zero padding."""
    import torch
    from .. import _lib
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    keep, rows = [], []
    if (isinstance(images, np.ndarray) or torch.is_tensor(images)) and images.ndim == 4:
        # one [B, H, W, 3] batch of equally sized images: a single device copy and vectorised descriptors (the per-image
        # loop below costs ~15 us of host time per image, which bounded the call at 0.07 of the copy bandwidth)
        batch = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
        if batch.shape[3] != 3:
            raise ValueError(f"batch shape {tuple(batch.shape)} is not [B, H, W, 3]")
        if batch.dtype not in (torch.uint8, torch.float32):
            batch = batch.float()
        batch = batch.to(dev).contiguous()
        keep.append(batch)
        nb, h, w = int(batch.shape[0]), int(batch.shape[1]), int(batch.shape[2])
        if preserve_aspect_ratio:
            oh, ow = _aspect_size(h, w, target_height, target_width)
            oy, ox = (target_height - oh) // 2, (target_width - ow) // 2
        else:
            oh, ow, oy, ox = target_height, target_width, 0, 0
        rows = np.empty((nb, 10), dtype=np.int64)
        rows[:, 0] = batch.data_ptr() + np.arange(nb, dtype=np.int64) * (h * w * 3 * batch.element_size())
        rows[:, 1:] = [h, w, 0 if batch.dtype == torch.uint8 else 1, oh, ow, oy, ox, 0, 0]
        images = ()
    for img in images:
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img))
        if img.dim() != 3 or img.shape[2] != 3:
            raise ValueError(f"image shape {tuple(img.shape)} is not [H, W, 3]")
        if img.dtype not in (torch.uint8, torch.float32):
            img = img.float()
        img = img.to(dev).contiguous()
        keep.append(img)
        h, w = int(img.shape[0]), int(img.shape[1])
        if preserve_aspect_ratio:
            oh, ow = _aspect_size(h, w, target_height, target_width)
            oy, ox = (target_height - oh) // 2, (target_width - ow) // 2
        else:
            oh, ow, oy, ox = target_height, target_width, 0, 0
        rows.append([img.data_ptr(), h, w, 0 if img.dtype == torch.uint8 else 1, oh, ow, oy, ox, 0, 0])
    rows = np.asarray(rows, dtype=np.int64)
    B = int(rows.shape[0])
    # the scale factors of ResizeBilinear, float32 in / float32 out, passed as bit patterns
    rows[:, 8] = (rows[:, 1].astype(np.float32) / rows[:, 4].astype(np.float32)).view(np.int32)
    rows[:, 9] = (rows[:, 2].astype(np.float32) / rows[:, 5].astype(np.float32)).view(np.int32)
    desc = _descriptor_upload(rows, dev)
    if out is None:
        out = torch.empty((B, target_height, target_width, 3), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().y3_preprocess(ctx.handle, _lib.ptr(desc), B, int(target_height), int(target_width),
                                        1 if divide_by_255 else 0, _lib.ptr(out), _lib.stream_ptr()))
    # the kernel reads `keep` and `desc` asynchronously: tie their lifetime to the stream
    for t in keep + [desc]:
        t.record_stream(torch.cuda.current_stream())
    return out


def resize_image(img, target_height, target_width):
    """reference core/utils.py:17-28 for one image [H, W, 3] or a batch [B, H, W, 3] (all images of a batch share H, W)."""
    if img.ndim == 4:
        return preprocess_images(img, target_height, target_width, preserve_aspect_ratio=True)
    return preprocess_images([img], target_height, target_width, preserve_aspect_ratio=True)[0]
