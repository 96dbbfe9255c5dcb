"""Shared helpers for the parity tests (inputs generators mirror SURVEY.md section 8d)."""
import numpy as np


def bf16_round(a):
    """Round float32 -> bfloat16 -> float32 (round to nearest even), numpy only."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def pack_weights(kernel_hwio, cout_pad):
    """[kh,kw,Cin,Cout] -> [Cout_pad][kh][kw][Cin] float32 (caller converts to bf16)."""
    kh, kw, cin, cout = kernel_hwio.shape
    w = np.zeros((cout_pad, kh, kw, cin), np.float32)
    w[:cout] = kernel_hwio.transpose(3, 0, 1, 2)
    return w


def synth_grids(B, sizes, nclasses, seed=0, obj_mean=0.0):
    """config-4 generator: t_xy, t_wh ~ N(0,1) (wh clipped to [-4,4]), t_cls ~ N(-2,2), t_obj ~ N(obj_mean, 2)."""
    rng = np.random.default_rng(seed)
    out = []
    for g in sizes:
        gh, gw = (g, g) if np.isscalar(g) else g
        t = np.empty((B, gh, gw, 3, 5 + nclasses), np.float32)
        t[..., 0:2] = rng.standard_normal((B, gh, gw, 3, 2))
        t[..., 2:4] = np.clip(rng.standard_normal((B, gh, gw, 3, 2)), -4, 4)
        t[..., 4] = rng.standard_normal((B, gh, gw, 3)) * 2 + obj_mean
        t[..., 5:] = rng.standard_normal((B, gh, gw, 3, nclasses)) * 2 - 2
        out.append(t)
    return out


def cluster_boxes(N, K, seed, jitter=0.02):
    """heavily overlapping boxes around K objects -> exercises multi-chunk suppression"""
    rng = np.random.default_rng(seed)
    cc = rng.random((K, 2)).astype(np.float32)
    cw = (rng.random((K, 2)) * 0.3 + 0.05).astype(np.float32)
    k = rng.integers(0, K, N)
    c = cc[k] + rng.normal(0, jitter, (N, 2)).astype(np.float32)
    wh = cw[k] * (1 + rng.normal(0, 0.1, (N, 2))).astype(np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    s = rng.random(N).astype(np.float32)
    return b, s
