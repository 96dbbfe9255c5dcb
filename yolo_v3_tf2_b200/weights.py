"""Weights in the reference's layout (host side, numpy).

Per conv (creation order = Keras ``conv2d``, ``conv2d_1`` ... = Darknet file order, reference convert.py:93-137):
  kernel  (kh, kw, Cin, Cout) float32  -- Keras HWIO
  bias    (Cout,)                      -- only for convs without batch_normalize (the three 1x1 head convs)
  bn      gamma, beta, moving_mean, moving_variance (Cout,) each -- Keras BatchNormalization variable order,
          inference formula gamma*(x-mean)/sqrt(var+1e-3)+beta (Keras default epsilon, parse_model.py:45-46)
"""
import numpy as np

BN_EPS = 1e-3   # tf.keras.layers.BatchNormalization default


class ConvParams:
    __slots__ = ("kernel", "bias", "gamma", "beta", "mean", "var")

    def __init__(self, kernel, bias=None, gamma=None, beta=None, mean=None, var=None):
        self.kernel, self.bias, self.gamma, self.beta, self.mean, self.var = kernel, bias, gamma, beta, mean, var

    @property
    def has_bn(self):
        return self.gamma is not None

    def as_list(self):
        """Keras variable order of the conv layer followed by its BN layer."""
        if self.has_bn:
            return [self.kernel, self.gamma, self.beta, self.mean, self.var]
        return [self.kernel, self.bias]


def init_keras_default(conv_shapes, seed=0):
    """init-K: what a freshly built reference model holds (convert.py:160-168 sanity forward): glorot-uniform kernels,
    zero bias, BN gamma=1 beta=0 mean=0 var=1."""
    rng = np.random.default_rng(seed)
    out = []
    for k, cin, cout, bn in conv_shapes:
        fan_in, fan_out = k * k * cin, k * k * cout
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        kern = rng.uniform(-lim, lim, size=(k, k, cin, cout)).astype(np.float32)
        if bn:
            out.append(ConvParams(kern, gamma=np.ones(cout, np.float32), beta=np.zeros(cout, np.float32),
                                  mean=np.zeros(cout, np.float32), var=np.ones(cout, np.float32)))
        else:
            out.append(ConvParams(kern, bias=np.zeros(cout, np.float32)))
    return out


def init_variance_preserving(conv_shapes, seed=0, obj_bias=-4.0, nclasses=None, residual_convs=()):
    """init-V: activations stay O(1) through all 75 layers so tolerance tests have teeth.
    He-style kernels for LeakyReLU(0.1); BN statistics are non-trivial but gamma/sqrt(var+eps) has unit mean square;
    convs that feed a shortcut (``residual_convs``: conv indices) are damped so the 23 residual adds do not blow the
    variance up; head bias pushes objectness down like a trained net."""
    rng = np.random.default_rng(seed)
    residual_convs = set(residual_convs)
    out = []
    for i, (k, cin, cout, bn) in enumerate(conv_shapes):
        fan_in = k * k * cin
        kern = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (1.01 * fan_in))).astype(np.float32)
        if bn:
            var = rng.uniform(0.5, 1.5, cout)
            gain = rng.uniform(0.75, 1.15, cout) * (np.sqrt(0.05) if i in residual_convs else 1.0)
            out.append(ConvParams(kern,
                                  gamma=(gain * np.sqrt(var + BN_EPS)).astype(np.float32),
                                  beta=(rng.standard_normal(cout) * 0.1).astype(np.float32),
                                  mean=(rng.standard_normal(cout) * 0.1).astype(np.float32),
                                  var=var.astype(np.float32)))
        else:
            b = np.zeros(cout, np.float32)
            if nclasses is not None and cout == 3 * (5 + nclasses):
                b.reshape(3, 5 + nclasses)[:, 4] = obj_bias
            out.append(ConvParams(kern * np.float32(0.5), bias=b))
    return out


def read_darknet_weights(path, conv_shapes):
    """Darknet ``.weights`` reader following reference convert.py:36-74, 93-95: int32[5] header, then per conv
    [beta, gamma, mean, var] (re-ordered to Keras [gamma, beta, mean, var]) or bias, then the kernel stored
    (Cout, Cin, kh, kw) and transposed to (kh, kw, Cin, Cout)."""
    out = []
    with open(path, "rb") as f:
        header = np.fromfile(f, dtype=np.int32, count=5)
        if header.size != 5:
            raise ValueError(f"{path}: truncated header")
        for k, cin, cout, bn in conv_shapes:
            if bn:
                v = np.fromfile(f, dtype=np.float32, count=4 * cout)
                if v.size != 4 * cout:
                    raise ValueError(f"{path}: truncated BN block")
                beta, gamma, mean, var = v.reshape(4, cout)
                bias = None
            else:
                bias = np.fromfile(f, dtype=np.float32, count=cout)
                if bias.size != cout:
                    raise ValueError(f"{path}: truncated bias block")
            n = cout * cin * k * k
            w = np.fromfile(f, dtype=np.float32, count=n)
            if w.size != n:
                raise ValueError(f"{path}: truncated kernel block")
            kern = np.ascontiguousarray(w.reshape(cout, cin, k, k).transpose(2, 3, 1, 0))
            if bn:
                out.append(ConvParams(kern, gamma=gamma.copy(), beta=beta.copy(), mean=mean.copy(), var=var.copy()))
            else:
                out.append(ConvParams(kern, bias=bias))
    return out


def write_darknet_weights(path, params):
    """Inverse of read_darknet_weights (used by tests to round-trip the format)."""
    with open(path, "wb") as f:
        np.array([0, 2, 0, 0, 0], dtype=np.int32).tofile(f)
        for p in params:
            if p.has_bn:
                np.stack([p.beta, p.gamma, p.mean, p.var]).astype(np.float32).tofile(f)
            else:
                p.bias.astype(np.float32).tofile(f)
            np.ascontiguousarray(p.kernel.transpose(3, 2, 0, 1)).astype(np.float32).tofile(f)


def params_from_list(arrays, conv_shapes):
    """Keras ``set_weights`` order: per conv layer [kernel, (bias)] followed by its BN [gamma, beta, mean, var]."""
    out, i = [], 0
    for k, cin, cout, bn in conv_shapes:
        kern = np.asarray(arrays[i], np.float32)
        if kern.shape != (k, k, cin, cout):
            raise ValueError(f"weight {i}: expected kernel shape {(k, k, cin, cout)}, got {kern.shape}")
        i += 1
        n = 4 if bn else 1
        vecs = [np.asarray(a, np.float32) for a in arrays[i:i + n]]
        if len(vecs) != n or any(v.shape != (cout,) for v in vecs):
            raise ValueError(f"weight {i}: expected {n} vectors of shape ({cout},)")
        i += n
        out.append(ConvParams(kern, gamma=vecs[0], beta=vecs[1], mean=vecs[2], var=vecs[3]) if bn
                   else ConvParams(kern, bias=vecs[0]))
    if i != len(arrays):
        raise ValueError(f"expected {i} weight arrays, got {len(arrays)}")
    return out
