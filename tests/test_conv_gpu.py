"""GPU parity of the tcgen05 implicit-GEMM conv (through the C ABI) against the torch-CPU oracle.

Tolerance (bf16 path): inputs/weights are rounded to bf16 on both sides and the oracle accumulates in fp32, so the only
differences are summation order and the final bf16 rounding of the output: |err| <= 2^-8 * |ref| + 2e-3 * sqrt(K)/32.
"""
import ctypes as C

import numpy as np
import pytest

from y3_test_util import bf16_round, pack_weights

pytestmark = pytest.mark.gpu


def _unswizzle(raw, swz):
    """raw smem image [128 rows][swz bytes] written by TMA with SWIZZLE_<swz>B -> logical [128][swz/2] bf16 bits.
    16-byte chunk c of row r is stored at chunk c ^ ((r >> s) & m): 128B: r & 7; 64B: (r >> 1) & 3."""
    rows = raw.reshape(128, swz // 16, 16)
    out = np.empty_like(rows)
    for r in range(128):
        x = (r & 7) if swz == 128 else ((r >> 1) & 3)
        for c in range(swz // 16):
            out[r, c] = rows[r, c ^ x]
    return out.reshape(128, swz).view(np.uint16)


def _ref_tile(x, ksize, stride, tap_r, tap_s, c0, m0, nch):
    B, H, W, Cin = x.shape
    pad_lo = 1 if stride > 1 else (ksize - 1) // 2
    pad_hi = 0 if stride > 1 else (ksize - 1) - pad_lo
    Ho = (H + pad_lo + pad_hi - ksize) // stride + 1
    Wo = (W + pad_lo + pad_hi - ksize) // stride + 1
    out = np.zeros((128, nch), np.float32)
    for i in range(128):
        m = m0 + i
        n, rem = divmod(m, Ho * Wo)
        p, q = divmod(rem, Wo)
        if n >= B:
            continue
        y, xx = p * stride - pad_lo + tap_r, q * stride - pad_lo + tap_s
        if 0 <= y < H and 0 <= xx < W:
            out[i] = x[n, y, xx, c0:c0 + nch]
    return out


@pytest.mark.parametrize("ksize,stride,swz,H,W,Cin", [
    (1, 1, 128, 8, 8, 64), (3, 1, 128, 13, 13, 64), (3, 2, 128, 16, 16, 128), (3, 1, 64, 12, 12, 32),
    (3, 2, 64, 16, 16, 32), (3, 1, 128, 26, 26, 128),
])
def test_tma_tile(cuda, ksize, stride, swz, H, W, Cin):
    import torch
    from yolo_v3_tf2_b200 import _lib
    ctx = _lib.context()
    rng = np.random.default_rng(1)
    B = 3
    x = bf16_round(rng.standard_normal((B, H, W, Cin)).astype(np.float32))
    xd = torch.from_numpy(x).cuda().to(torch.bfloat16)
    nch = swz // 2
    Ho = H if stride == 1 else H // 2
    M = B * Ho * Ho
    taps = [(0, 0)] if ksize == 1 else [(0, 0), (1, 1), (2, 2), (0, 2), (2, 0)]
    for m0 in sorted({0, 128, ((M - 1) // 128) * 128}):
        if m0 >= M:
            continue
        for (r, s) in taps:
            for c0 in sorted({0, Cin - nch}):
                out = torch.zeros(128 * swz, dtype=torch.uint8, device="cuda")
                _lib.check(_lib.lib().y3_dbg_tma_tile(ctx.handle, _lib.ptr(xd), B, H, W, Cin, Cin, ksize, stride, swz,
                                                     r, s, c0, m0, _lib.ptr(out), _lib.stream_ptr()))
                torch.cuda.synchronize()
                got = _unswizzle(out.cpu().numpy(), swz)
                got = (got.astype(np.uint32) << 16).view(np.float32)
                want = _ref_tile(x, ksize, stride, r, s, c0, m0, nch)
                np.testing.assert_array_equal(got, want, err_msg=f"tile m0={m0} tap=({r},{s}) c0={c0}")


CONV_CASES = [
    # (B, H, W, Cin, Cout, k, stride, leaky, residual, upsample, out_fp32)
    (2, 8, 8, 64, 64, 1, 1, 1, False, False, False),        # smallest 1x1: 2-D tiled A map
    (2, 13, 13, 128, 256, 1, 1, 1, False, False, False),    # M=338 (tail tile), N=256
    (3, 13, 13, 64, 128, 3, 1, 1, False, False, False),     # im2col, 'same' padding
    (3, 13, 13, 64, 128, 3, 1, 1, True, False, False),      # + residual
    (2, 16, 16, 64, 128, 3, 2, 1, False, False, False),     # stride 2, asymmetric padding
    (2, 32, 32, 32, 64, 3, 2, 1, False, False, False),      # Cin=32 -> 64-byte swizzle
    (2, 16, 16, 32, 64, 3, 1, 1, True, False, False),       # Cin=32, stride 1, residual
    (2, 13, 13, 512, 256, 1, 1, 1, False, True, False),     # 1x1 + fused 2x upsample
    (2, 13, 13, 256, 255, 1, 1, 0, False, False, True),     # head: linear, bias, fp32 out, N=255
    (4, 13, 13, 512, 1024, 3, 1, 1, True, False, False),    # deep K=4608, 4 N tiles
    (2, 26, 26, 768, 256, 1, 1, 1, False, False, False),    # concat-width input
    (5, 19, 19, 256, 512, 3, 1, 1, False, False, False),    # 608-style odd grid, many tiles
    (1, 52, 52, 128, 256, 3, 1, 1, True, False, False),
    (2, 8, 8, 64, 32, 1, 1, 1, False, False, False),        # N=32
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv2d_vs_oracle(cuda, case):
    import torch
    from yolo_v3_tf2_b200 import _lib
    from oracle import net_oracle
    B, H, W, Cin, Cout, k, stride, leaky, use_res, up, fp32 = case
    ctx = _lib.context()
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = bf16_round(rng.standard_normal((B, H, W, Cin)).astype(np.float32))
    kern = bf16_round((rng.standard_normal((k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    bias = rng.standard_normal(Cout).astype(np.float32) * 0.1
    Ho = H if stride == 1 else H // 2
    Wo = W if stride == 1 else W // 2
    res = bf16_round(rng.standard_normal((B, Ho, Wo, Cout)).astype(np.float32)) if use_res else None
    ref = net_oracle.conv_layer(x, kern, bias, k, stride, leaky, res, up)

    bn = _lib.lib().y3_conv_block_n(Cin, Cout)
    cout_pad = ((Cout + bn - 1) // bn) * bn
    wd = torch.from_numpy(pack_weights(kern, cout_pad)).cuda().to(torch.bfloat16).contiguous()
    bd = torch.zeros(cout_pad, dtype=torch.float32, device="cuda")
    bd[:Cout] = torch.from_numpy(bias).cuda()
    xd = torch.from_numpy(x).cuda().to(torch.bfloat16).contiguous()
    rd = torch.from_numpy(res).cuda().to(torch.bfloat16).contiguous() if use_res else None
    oshape = (B, Ho * (2 if up else 1), Wo * (2 if up else 1), Cout)
    od = torch.full(oshape, float("nan"), dtype=torch.float32 if fp32 else torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().y3_conv2d_bf16(ctx.handle, _lib.ptr(xd), B, H, W, Cin, Cin, _lib.ptr(wd), _lib.ptr(bd), k, stride,
                                         Cout, leaky, _lib.ptr(rd), Cout, _lib.ptr(od), Cout, int(fp32), int(up),
                                         _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert ctx.watchdog_code() == 0
    got = od.float().cpu().numpy()
    assert np.isfinite(got).all(), "unwritten / non-finite outputs"
    K = k * k * Cin
    tol = 2.0 ** -8 * np.abs(ref) + 2e-3 * np.sqrt(K) / 32 + 1e-3
    err = np.abs(got - ref)
    assert (err <= tol).all(), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}; ref there {ref.flat[err.argmax()]}"


def test_conv2d_concat_slice_views(cuda):
    """Input read from, and output written into, channel slices of wider buffers (zero-copy concat)."""
    import torch
    from yolo_v3_tf2_b200 import _lib
    from oracle import net_oracle
    ctx = _lib.context()
    rng = np.random.default_rng(5)
    B, H, W, Cin, Cout = 2, 13, 13, 128, 128
    wide_in = bf16_round(rng.standard_normal((B, H, W, 384)).astype(np.float32))
    x = wide_in[..., 128:256]
    kern = bf16_round((rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32))
    bias = np.zeros(Cout, np.float32)
    ref = net_oracle.conv_layer(x, kern, bias, 3, 1, 1)
    wd = torch.from_numpy(pack_weights(kern, Cout)).cuda().to(torch.bfloat16).contiguous()
    bd = torch.zeros(Cout, dtype=torch.float32, device="cuda")
    ind = torch.from_numpy(wide_in).cuda().to(torch.bfloat16).contiguous()
    outd = torch.zeros((B, H, W, 256), dtype=torch.bfloat16, device="cuda")
    xin = C.c_void_p(ind.data_ptr() + 128 * 2)
    oout = C.c_void_p(outd.data_ptr() + 64 * 2)
    _lib.check(_lib.lib().y3_conv2d_bf16(ctx.handle, xin, B, H, W, Cin, 384, _lib.ptr(wd), _lib.ptr(bd), 3, 1, Cout, 1, None, 0,
                                         oout, 256, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = outd.float().cpu().numpy()
    assert (got[..., :64] == 0).all() and (got[..., 192:] == 0).all(), "wrote outside the channel slice"
    err = np.abs(got[..., 64:192] - ref)
    assert (err <= 2.0 ** -8 * np.abs(ref) + 5e-3).all(), err.max()


@pytest.mark.parametrize("B,H,W,stride", [(2, 32, 32, 1), (1, 416, 416, 1), (3, 35, 29, 1), (2, 64, 64, 2)])
def test_conv_stem_vs_oracle(cuda, B, H, W, stride):
    """3-channel stem on the tensor cores: fp32 image split into bf16 hi + lo halves, so only the weights are rounded."""
    import torch
    from yolo_v3_tf2_b200 import _lib
    from oracle import net_oracle
    if stride == 2 and (H % 2 or W % 2):
        pytest.skip("even sizes only")
    ctx = _lib.context()
    rng = np.random.default_rng(B * 1000 + H)
    x = rng.random((B, H, W, 3), dtype=np.float32)
    kern = bf16_round((rng.standard_normal((3, 3, 3, 32)) / np.sqrt(27)).astype(np.float32))
    bias = (rng.standard_normal(32) * 0.1).astype(np.float32)
    ref = net_oracle.conv_layer(x, kern, bias, 3, stride, 1)
    wp = np.zeros((32, 64), np.float32)
    flat = kern.reshape(27, 32).T          # [o][k], k ordered (r, s, c)
    wp[:, :27] = flat
    wp[:, 32:59] = flat
    wd = torch.from_numpy(wp).cuda().to(torch.bfloat16).contiguous()
    bd = torch.from_numpy(bias).cuda()
    xd = torch.from_numpy(x).cuda()
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    od = torch.full((B, Ho, Wo, 32), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().y3_conv2d_stem_f32(ctx.handle, _lib.ptr(xd), B, H, W, _lib.ptr(wd), _lib.ptr(bd), stride, 1,
                                             _lib.ptr(od), 32, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert ctx.watchdog_code() == 0
    got = od.float().cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - ref)
    # input precision ~2^-16, output rounded to bf16
    assert (err <= 2.0 ** -8 * np.abs(ref) + 1e-3).all(), err.max()


FLAT_CASES = [
    # (B, H, W, Cin, Cout, leaky, residual)
    (2, 13, 13, 64, 128, 1, False), (3, 13, 13, 128, 256, 1, True), (2, 26, 26, 256, 512, 1, True),
    (1, 52, 52, 128, 256, 1, True), (2, 104, 104, 64, 128, 1, True), (1, 208, 208, 32, 64, 1, True),
    (5, 19, 19, 128, 256, 0, False), (2, 13, 13, 512, 1024, 1, True), (1, 9, 17, 64, 64, 1, False),
]


@pytest.mark.parametrize("case", FLAT_CASES, ids=[str(c) for c in FLAT_CASES])
def test_conv_flat_vs_oracle(cuda, case):
    """3x3 stride-1 conv on the zero-haloed flat layout [B, H+1, W+1, C]: one staged patch per 64-channel block, the
    nine taps are row-shifted UMMA descriptors.  Same tolerance as the im2col-fed kernel."""
    import torch
    from yolo_v3_tf2_b200 import _lib
    from oracle import net_oracle
    B, H, W, Cin, Cout, leaky, use_res = case
    ctx = _lib.context()
    rng = np.random.default_rng(hash(case) & 0xffff)
    x = bf16_round(rng.standard_normal((B, H, W, Cin)).astype(np.float32))
    kern = bf16_round((rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32))
    bias = rng.standard_normal(Cout).astype(np.float32) * 0.1
    res = bf16_round(rng.standard_normal((B, H, W, Cout)).astype(np.float32)) if use_res else None
    ref = net_oracle.conv_layer(x, kern, bias, 3, 1, leaky, res)
    bk = 64 if Cin % 64 == 0 else 32
    bn = 64 if Cout <= 64 else (128 if Cout <= 128 else 256)
    cout_pad = ((Cout + bn - 1) // bn) * bn
    wp = np.zeros((cout_pad, Cin // bk, 3, 3, bk), np.float32)
    wp[:Cout] = kern.reshape(3, 3, Cin // bk, bk, Cout).transpose(4, 2, 0, 1, 3)
    wd = torch.from_numpy(wp).cuda().to(torch.bfloat16).contiguous()
    bd = torch.zeros(cout_pad, dtype=torch.float32, device="cuda")
    bd[:Cout] = torch.from_numpy(bias).cuda()
    xpad = np.zeros((B, H + 1, W + 1, Cin), np.float32)
    xpad[:, :H, :W] = x
    xd = torch.from_numpy(xpad).cuda().to(torch.bfloat16).contiguous()
    rd = torch.from_numpy(res).cuda().to(torch.bfloat16).contiguous() if use_res else None
    od = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().y3_conv2d_flat_bf16(ctx.handle, _lib.ptr(xd), B, H, W, Cin, _lib.ptr(wd), _lib.ptr(bd), Cout, leaky,
                                              _lib.ptr(rd), Cout, _lib.ptr(od), Cout, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert ctx.watchdog_code() == 0
    got = od.float().cpu().numpy()
    assert np.isfinite(got).all(), "unwritten / non-finite outputs"
    K = 9 * Cin
    tol = 2.0 ** -8 * np.abs(ref) + 2e-3 * np.sqrt(K) / 32 + 1e-3
    err = np.abs(got - ref)
    assert (err <= tol).all(), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"
