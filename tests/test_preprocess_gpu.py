"""GPU pre-processing vs the numpy oracle: bit-exact (every float32 operation of the
TensorFlow formula is rounded separately on both sides)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["uint8", "float32"])
def test_resize_matches_oracle_bit_exact(cuda, dtype):
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(3)
    shapes = [(480, 640), (333, 500), (416, 416), (37, 53), (1080, 1920)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if dtype == "uint8" else rng.random((h, w, 3), dtype=np.float32)
            for h, w in shapes]
    div = dtype == "uint8"
    out = y3.preprocess_images([torch.from_numpy(i).cuda() for i in imgs], 416, 416, divide_by_255=div)
    torch.cuda.synchronize()
    assert out.shape == (len(shapes), 416, 416, 3) and out.dtype == torch.float32
    for k, img in enumerate(imgs):
        assert np.array_equal(out[k].cpu().numpy(), po.resize(img, 416, divide_by_255=div)), f"image {k}"


def test_resize_image_aspect_preserving(cuda):
    """reference core/utils.py:17-28: resize with preserve_aspect_ratio, then pad_to_bounding_box (centred, zeros)."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(5)
    for (h, w, th, tw) in [(480, 640, 416, 416), (640, 480, 416, 416), (100, 300, 608, 608), (50, 50, 64, 96)]:
        img = rng.random((h, w, 3), dtype=np.float32)
        got = y3.resize_image(torch.from_numpy(img).cuda(), th, tw)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), po.resize_image(img, th, tw)), (h, w, th, tw)
    batch = rng.integers(0, 256, (3, 120, 200, 3), dtype=np.uint8)
    got = y3.resize_image(torch.from_numpy(batch).cuda(), 416, 416)
    for k in range(3):
        assert np.array_equal(got[k].cpu().numpy(), po.resize_image(batch[k], 416, 416))


def test_preprocess_feeds_the_model(cuda):
    """uint8 frames -> GPU resize / 255 -> model: same grids as feeding the oracle-resized float image."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(7)
    frames = [rng.integers(0, 256, (90, 130, 3), dtype=np.uint8) for _ in range(2)]
    model = y3.ParseModel.builtin_yolov3_tiny(80).init_weights("variance", seed=1)
    x = y3.preprocess_images(frames, 96, 96, divide_by_255=True)
    a = model(x)
    b = model(np.stack([po.resize(f, 96, divide_by_255=True) for f in frames]))
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_batch_tensor_equals_list_of_images(cuda):
    """preprocess_images on ONE [B, H, W, 3] tensor (vectorised descriptors, a single device copy) gives exactly what the
    per-image list path gives (tf.image.resize of inference.py:157-158 and resize_image of core/utils.py:17-28)."""
    import torch
    import yolo_v3_tf2_b200 as y3
    rng = np.random.default_rng(11)
    for dtype in (np.uint8, np.float32):
        batch = (rng.integers(0, 256, (5, 90, 130, 3)).astype(dtype) if dtype == np.uint8
                 else rng.random((5, 90, 130, 3), dtype=np.float32))
        for kw in ({"divide_by_255": True}, {"preserve_aspect_ratio": True}):
            a = y3.preprocess_images(batch, 64, 96, **kw)                      # numpy batch
            b = y3.preprocess_images(torch.from_numpy(batch).cuda(), 64, 96, **kw)   # CUDA batch
            c = y3.preprocess_images([torch.from_numpy(f).cuda() for f in batch], 64, 96, **kw)
            torch.cuda.synchronize()
            assert torch.equal(a, c) and torch.equal(b, c)


@pytest.mark.parametrize("shape,out_hw", [((7, 91, 131, 3), (64, 96)), ((3, 5, 3, 3), (32, 32)), ((2, 417, 419, 3), (416, 416)),
                                          ((4, 2, 2, 3), (8, 8))])
def test_uint8_batch_unaligned_images_bit_exact(cuda, shape, out_hw):
    """uint8 frames inside one [B, h, w, 3] tensor start at arbitrary byte alignments (h * w * 3 odd): the kernel's aligned
    8-byte word loads, their byte-load fall-back at both ends of every image buffer, tiny images and both up- and
    down-scaling must give the oracle's bits (tf.image.resize(...) / 255, inference.py:157-158, core/load_tfrecords.py:46)."""
    import torch
    import yolo_v3_tf2_b200 as y3
    from oracle import preprocess_oracle as po
    rng = np.random.default_rng(sum(shape))
    batch = rng.integers(0, 256, shape, dtype=np.uint8)
    # an odd byte offset for the whole batch as well: a view that starts 1 byte into its allocation
    raw = torch.zeros(batch.size + 1, dtype=torch.uint8, device="cuda")
    raw[1:] = torch.from_numpy(batch.reshape(-1)).cuda()
    view = raw[1:].view(*shape)
    out = y3.preprocess_images(view, out_hw[0], out_hw[1], divide_by_255=True)
    torch.cuda.synchronize()
    for k in range(shape[0]):
        ref = (po.resize_bilinear(batch[k], out_hw[0], out_hw[1]) / np.float32(255.0)).astype(np.float32)
        assert np.array_equal(out[k].cpu().numpy(), ref), f"image {k}"
