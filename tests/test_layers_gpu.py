"""Layer-level parity of the first Darknet-53 layers inside the whole network (reference core/parse_model.py:13-56,
143-160; config/models/yolov3/backbone.yaml): every checked layer's GPU output (read back with y3_net_read_layer) is
compared with the oracle's conv applied to the GPU's OWN input of that layer, so the tolerance is that of one layer --
bf16 rounding of the output, |err| <= 2^-8 |ref| + 2e-3 -- not of the accumulated network.  This is what pins the
pixel-pair kernels of the Cin = 32 layers (ConvArgs::ksize_w: image borders, odd / even output columns, the fused
residual), the stem and the first CTA-pair layers at a tolerance that a single wrong filter tap cannot hide in.
"""
import numpy as np
import pytest

from y3_test_util import bf16_round

pytestmark = pytest.mark.gpu


def _folded(p):
    """BN folded the way y3_net_load_conv does it: scale in double -> float32, kernel * scale in float32 -> bf16."""
    k = np.asarray(p.kernel, np.float32)
    if p.gamma is None:
        return bf16_round(k), np.asarray(p.bias, np.float32)
    sc = np.asarray(p.gamma, np.float64) / np.sqrt(np.asarray(p.var, np.float64) + 1e-3)
    shift = (np.asarray(p.beta, np.float64) - np.asarray(p.mean, np.float64) * sc).astype(np.float32)
    return bf16_round(k * sc.astype(np.float32)), shift


def _truncated(n_entries, tiny=False, seed=5):
    """The first n_entries entries of the backbone's layer list (reference config/models/yolov3/backbone.yaml or
    yolov3_tiny) followed by a linear 1x1 head conv + yolo layer, so that the last backbone tensor is still live when the
    forward pass ends (arena buffers are recycled three launches after their last reader)."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    if tiny:
        _, files = configs.yolov3_tiny_config()
        base = [v for k, v in files.items() if "backbone" in k][0]
    else:
        base = configs.backbone_layers()
    layers = list(base[:n_entries]) + [
        {"type": "convolutional", "filters": configs.FILTER_EXPR, "size": 1, "stride": 1, "pad": 1, "activation": "linear"},
        {"type": "yolo", "grid_size": 0}]
    subs = [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}]
    model = y3.ParseModel().build_model(None, subs, "head", nclasses=3, layer_lists={"a": layers})
    return model.init_weights("variance", seed=seed)


def _check_last_backbone_layer(model, x, u8=False):
    """Compare the last batch-normalised conv of `model` (with the shortcut fused into it, if one follows) with the
    oracle conv applied to the GPU's own input (and residual) of that layer."""
    import torch
    from oracle import net_oracle
    model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    L = model.graph.layers
    convs = [i for i, l in enumerate(L) if l.op == net_oracle.OP_CONV]
    i = convs[-2]                      # the head conv is the last one
    l = L[i]
    x_f = (x.astype(np.float32) / np.float32(255.0)) if u8 else x

    def tensor(t):   # tensor id -> float32 NHWC as the GPU holds it (tensor t is the output of layer t - 1)
        return x_f if t == 0 else model.read_layer(t - 1, x.shape)

    w, b = _folded(model._params[len(convs) - 2])
    a = tensor(l.src0)
    if l.src0 == 0:
        a = a.astype(np.float32)       # the stem reads the fp32 image itself (hi / lo bf16 split: no rounding)
    res = None
    out_layer = i
    if L[i + 1].op == net_oracle.OP_SHORTCUT:
        s = L[i + 1]
        assert i + 1 in (s.src0, s.src1)
        res = tensor(s.src1 if s.src0 == i + 1 else s.src0)
        out_layer = i + 1
    ref = net_oracle.conv_layer(a, w, b, l.ksize, l.stride, l.activation == 1, residual=res, dtype=torch.float64)
    got = model.read_layer(out_layer, x.shape)
    assert got.shape == ref.shape, (i, got.shape, ref.shape)
    err = np.abs(got - ref)
    tol = 2.0 ** -8 * np.abs(ref) + 2e-3
    k = int(np.argmax(err - tol))
    assert (err <= tol).all(), (f"layer {i} (k={l.ksize} s={l.stride} f={l.filters}): {int((err > tol).sum())} of {err.size} "
                                f"values out of tolerance, worst at {np.unravel_index(k, err.shape)}: got {got.flat[k]} "
                                f"ref {ref.flat[k]}")
    assert np.abs(ref).max() > 0.05    # the comparison is not vacuous


# entries of backbone_layers(): 0 route, 1 stem, 2 3x3/2 32->64, 3 1x1 64->32, 4 3x3 32->64, 5 shortcut, 6 3x3/2 64->128,
# 7 1x1, 8 3x3 128, 9 shortcut
@pytest.mark.parametrize("n_entries", [2, 3, 4, 6, 7, 10])
@pytest.mark.parametrize("size,B", [(64, 2), (96, 3), (416, 1), (160, 5)])
def test_first_layers_vs_oracle(cuda, n_entries, size, B):
    model = _truncated(n_entries)
    x = np.random.default_rng(size + B).random((B, size, size, 3), dtype=np.float32)
    _check_last_backbone_layer(model, x)


@pytest.mark.parametrize("n_entries", [2, 3, 6])
def test_first_layers_uint8_input(cuda, n_entries):
    model = _truncated(n_entries, seed=6)
    x = np.random.default_rng(9).integers(0, 256, (2, 128, 128, 3), dtype=np.uint8)
    _check_last_backbone_layer(model, x, u8=True)


@pytest.mark.parametrize("n_entries", [2, 4])
def test_tiny_first_layers_vs_oracle(cuda, n_entries):
    """YOLOv3-tiny: 16-filter stem stored as 32 channels -> maxpool -> 3x3 16->32 (pixel-pair view with zero weights on
    the padding channels)."""
    model = _truncated(n_entries, tiny=True, seed=7)
    x = np.random.default_rng(3).random((2, 96, 96, 3), dtype=np.float32)
    _check_last_backbone_layer(model, x)
