"""Oracle for the conv stack: torch-CPU restatement of the Keras graph built by reference core/parse_model.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned: TensorFlow is unavailable.

Semantics restated, with the reference line each follows:
  * stride > 1  -> ZeroPadding2D(((1,0),(1,0))) then 'valid'           parse_model.py:34-35, 31
  * stride == 1 and pad == 1 -> 'same', else 'valid'                    parse_model.py:31
  * Conv2D(use_bias = not batch_normalize, linear)                      parse_model.py:37-43
  * BatchNormalization() inference: gamma*(x-mean)/sqrt(var+1e-3)+beta  parse_model.py:45-46 (Keras default epsilon)
  * LeakyReLU(alpha=0.1)                                                parse_model.py:51-52
  * Add()([from, x])                                                    parse_model.py:155-156
  * UpSampling2D(size=2) nearest                                        parse_model.py:71-72
  * Concatenate(axis=3)([layers..., inputs...])                         parse_model.py:116-134
  * Reshape((g, g, 3, 5+C)) -- pure view, grid taken from the tensor    parse_model.py:209-210
  * MaxPooling2D(size, strides, padding): 'same' = ceil(H/stride) outputs, total padding
    max((Ho-1)*stride + size - H, 0) with the smaller half first, padding never wins the max (-inf)  parse_model.py:78-99
"""
import numpy as np
import torch
import torch.nn.functional as F

OP_CONV, OP_SHORTCUT, OP_UPSAMPLE, OP_CONCAT, OP_YOLO, OP_MAXPOOL = range(6)
BN_EPS = 1e-3


def forward(layers, outputs, params, x_nhwc, dtype=torch.float32, keep=None):
    """layers: list of records with fields op, src0, src1, ksize, stride, filters, pad, batch_normalize, activation
    (yolo_v3_tf2_b200.graph.Layer); params: per-conv objects with kernel (HWIO) and bias or gamma/beta/mean/var;
    x_nhwc: [B,H,W,3] float array.  Returns the list of output grids [B,g,g,3,5+C] (numpy, float32/64).
    ``keep``: optional set of tensor ids whose NHWC activations are also returned (dict id -> array)."""
    x = torch.as_tensor(np.asarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2).contiguous()
    t = {0: x}
    ci = 0
    with torch.no_grad():
        for i, l in enumerate(layers):
            a = t[l.src0]
            if l.op == OP_CONV:
                p = params[ci]
                ci += 1
                w = torch.as_tensor(p.kernel).to(dtype).permute(3, 2, 0, 1).contiguous()   # HWIO -> OIHW
                if l.stride > 1:
                    a = F.pad(a, (1, 0, 1, 0))
                    pad = 0
                else:
                    pad = (l.ksize - 1) // 2 if l.pad == 1 else 0
                bias = None if l.batch_normalize else torch.as_tensor(p.bias).to(dtype)
                y = F.conv2d(a, w, bias, stride=l.stride, padding=pad)
                if l.batch_normalize:
                    g, b, m, v = (torch.as_tensor(q).to(dtype).view(1, -1, 1, 1) for q in (p.gamma, p.beta, p.mean, p.var))
                    y = g * (y - m) / torch.sqrt(v + BN_EPS) + b
                if l.activation == 1:
                    y = F.leaky_relu(y, 0.1)
            elif l.op == OP_SHORTCUT:
                y = t[l.src1] + a
            elif l.op == OP_UPSAMPLE:
                y = F.interpolate(a, scale_factor=l.stride, mode="nearest")
            elif l.op == OP_CONCAT:
                y = torch.cat([a, t[l.src1]], dim=1)
            elif l.op == OP_YOLO:
                y = a
            elif l.op == OP_MAXPOOL:
                y = maxpool_tf(a, l.ksize, l.stride, l.pad == 1)
            else:
                raise ValueError(f"oracle: unsupported op {l.op}")
            t[i + 1] = y
    outs = []
    for o in outputs:
        y = t[o].permute(0, 2, 3, 1).contiguous()
        B, gh, gw, ch = y.shape
        outs.append(y.reshape(B, gh, gw, 3, ch // 3).numpy())
    if keep is not None:
        return outs, {k: t[k].permute(0, 2, 3, 1).contiguous().numpy() for k in keep}
    return outs


def maxpool_tf(a_nchw, size, stride, same):
    """Keras/TF MaxPooling2D on an NCHW torch tensor with TF's 'same' padding rule (asymmetric, -inf padding)."""
    if same:
        H, W = a_nchw.shape[2], a_nchw.shape[3]
        Ho, Wo = -(-H // stride), -(-W // stride)
        ph = max((Ho - 1) * stride + size - H, 0)
        pw = max((Wo - 1) * stride + size - W, 0)
        a_nchw = F.pad(a_nchw, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
    return F.max_pool2d(a_nchw, size, stride)


def conv_layer(x_nhwc, kernel_hwio, bias, ksize, stride, leaky, residual=None, upsample=False, dtype=torch.float32):
    """One fused conv step (conv + bias + leaky + residual + upsample) for kernel unit tests."""
    a = torch.as_tensor(np.asarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
    w = torch.as_tensor(np.asarray(kernel_hwio)).to(dtype).permute(3, 2, 0, 1).contiguous()
    if stride > 1:
        a = F.pad(a, (1, 0, 1, 0))
        pad = 0
    else:
        pad = (ksize - 1) // 2
    y = F.conv2d(a, w, torch.as_tensor(np.asarray(bias)).to(dtype), stride=stride, padding=pad)
    if leaky:
        y = F.leaky_relu(y, 0.1)
    if residual is not None:
        y = y + torch.as_tensor(np.asarray(residual)).to(dtype).permute(0, 3, 1, 2)
    if upsample:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    return y.permute(0, 2, 3, 1).contiguous().numpy()
