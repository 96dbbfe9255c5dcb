// Input pre-processing of the reference, on the GPU:
//   tf.image.resize(img, (S, S))                       inference.py:157-158, core/load_tfrecords.py:46 (then / 255)
//   resize_image(img, th, tw): tf.image.resize(preserve_aspect_ratio=True) + pad_to_bounding_box   core/utils.py:17-28
// tf.image.resize's default is ResizeBilinear with half_pixel_centers=True and no antialiasing:
//   scale = in / out (float);  in_f = (i + 0.5) * scale - 0.5;  lo = max(floor(in_f), 0);  hi = min(ceil(in_f), in - 1);
//   lerp = in_f - floor(in_f);  top = tl + (tr - tl) * xl;  bot = bl + (br - bl) * xl;  out = top + (bot - top) * yl
// every step a separately rounded float32 operation (the intrinsics below keep the compiler from fusing them), so the
// result is bit-identical to the numpy restatement in oracle/preprocess_oracle.py.
// HBM-bound: one thread per output pixel (3 channels = 12 bytes written, <= 4 source pixels read through L1/L2).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace y3 {

struct ImageDesc {          // one per image, device memory (10 x int64)
    long long src;          // device pointer: uint8 or float32 [H, W, 3]
    long long H, W;
    long long dtype;        // 0 uint8, 1 float32
    long long out_h, out_w; // resized size (== canvas size without aspect preservation)
    long long off_y, off_x; // top-left corner inside the canvas (pad_to_bounding_box)
    long long scale_y, scale_x;   // float32 bit patterns of H / out_h and W / out_w (IEEE division done once on the host)
};

struct PreprocessArgs {
    const ImageDesc* desc;
    float* out;             // [B, dst_h, dst_w, 3] float32
    int B, dst_h, dst_w;
    float mul;              // 1/255 for uint8 sources of the tfrecord path, else 1
    int use_mul;
    unsigned long long div_magic;   // ceil(2^40 / dst_w) when idx / dst_w == (idx * magic) >> 40 for every pixel index, else 0
};

// The kernel was bound by the XU pipe (ncu: 66 % busy -- int <-> float conversions, floor / ceil and the three IEEE
// divisions are 16-lane operations): 12 byte -> float conversions, 8 rounding / conversion steps and 3 reciprocals per
// output pixel.  The helpers below do the same arithmetic on the ALU / FMA pipes, bit for bit.
// exact float of an integer 0 <= i < 2^23 (no I2F): 2^23 + i has i in its mantissa
__device__ __forceinline__ float small_uint_to_float(unsigned i) { return __fsub_rn(__uint_as_float(0x4B000000u | i), 8388608.0f); }
// exact int of an integral float |f| < 2^22 (no F2I): 1.5 * 2^23 + f has f in its mantissa (two's complement)
__device__ __forceinline__ int integral_float_to_int(float f) { return __float_as_int(__fadd_rn(f, 12582912.0f)) - 0x4B400000; }
// v / 255 correctly rounded, for 0 <= v <= 256, without the reciprocal: q0 = v * RN(1/255), one residual correction.
// Equal to the IEEE quotient for ALL 1 132 462 081 floats of that range (exhaustive check: tests/div255_check.c).
__device__ __forceinline__ float div255(float v) {
    const float r = 0.003921568859368563f;   // RN(1 / 255) = 0x3B808081
    const float q0 = __fmul_rn(v, r);
    const float e = __fmaf_rn(-q0, 255.0f, v);
    return __fmaf_rn(e, r, q0);
}

__device__ __forceinline__ float load_px(const ImageDesc& d, long long idx) {
    if (d.dtype == 0) return (float)reinterpret_cast<const uint8_t*>(d.src)[idx];
    return reinterpret_cast<const float*>(d.src)[idx];
}

// grid = (ceil(dst_h * dst_w / 256), B): blockIdx.y is the image, so no 64-bit division per pixel; the image's
// descriptor is read once per block into shared memory.
__global__ void __launch_bounds__(256) preprocess_kernel(const PreprocessArgs a) {
    __shared__ ImageDesc sd;
    const int b = blockIdx.y;
    if (threadIdx.x < (int)(sizeof(ImageDesc) / 8))
        reinterpret_cast<long long*>(&sd)[threadIdx.x] = reinterpret_cast<const long long*>(a.desc + b)[threadIdx.x];
    __syncthreads();
    const int per = a.dst_h * a.dst_w;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= per) return;
    // (the compiler's 32-bit division goes through I2F / MUFU.RCP / F2I: three more XU operations per pixel)
    const int y = a.div_magic ? (int)(((unsigned long long)(unsigned)idx * a.div_magic) >> 40) : idx / a.dst_w;
    const int x = idx - y * a.dst_w;
    const int H = (int)sd.H, W = (int)sd.W, out_h = (int)sd.out_h, out_w = (int)sd.out_w;
    float r[3] = {0.f, 0.f, 0.f};
    const int yy = y - (int)sd.off_y, xx = x - (int)sd.off_x;
    if (yy >= 0 && yy < out_h && xx >= 0 && xx < out_w) {
        const float sy = __int_as_float((int)sd.scale_y), sx = __int_as_float((int)sd.scale_x);
        const float fy = __fsub_rn(__fmul_rn(__fadd_rn(small_uint_to_float((unsigned)yy), 0.5f), sy), 0.5f);
        const float fx = __fsub_rn(__fmul_rn(__fadd_rn(small_uint_to_float((unsigned)xx), 0.5f), sx), 0.5f);
        const float fly = floorf(fy), flx = floorf(fx);
        // ceil(f) = floor(f) + (f > floor(f)); the conversion without F2I needs |f| < 2^22 (any real image)
        const bool small = (H | W) < (1 << 22);
        const int yfl = small ? integral_float_to_int(fly) : (int)fly, xfl = small ? integral_float_to_int(flx) : (int)flx;
        const int y0 = max(yfl, 0), y1 = min(yfl + (fy > fly ? 1 : 0), H - 1);
        const int x0 = max(xfl, 0), x1 = min(xfl + (fx > flx ? 1 : 0), W - 1);
        const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
        const long long r0 = (long long)y0 * W, r1 = (long long)y1 * W;
        float tl[3], tr[3], bl[3], br[3];
        if (sd.dtype == 0) {
            const uint8_t* p = reinterpret_cast<const uint8_t*>(sd.src);
            const uint8_t *ptl = p + (r0 + x0) * 3, *ptr_ = p + (r0 + x1) * 3, *pbl = p + (r1 + x0) * 3, *pbr = p + (r1 + x1) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                tl[c] = small_uint_to_float(__ldg(ptl + c)); tr[c] = small_uint_to_float(__ldg(ptr_ + c));
                bl[c] = small_uint_to_float(__ldg(pbl + c)); br[c] = small_uint_to_float(__ldg(pbr + c));
            }
        } else {
            const float* p = reinterpret_cast<const float*>(sd.src);
            const float *ptl = p + (r0 + x0) * 3, *ptr_ = p + (r0 + x1) * 3, *pbl = p + (r1 + x0) * 3, *pbr = p + (r1 + x1) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                tl[c] = __ldg(ptl + c); tr[c] = __ldg(ptr_ + c);
                bl[c] = __ldg(pbl + c); br[c] = __ldg(pbr + c);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float top = __fadd_rn(tl[c], __fmul_rn(__fsub_rn(tr[c], tl[c]), lx));
            const float bot = __fadd_rn(bl[c], __fmul_rn(__fsub_rn(br[c], bl[c]), lx));
            float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            // the reference divides: resize(...) / 255.  uint8 sources give 0 <= v <= 255 (div255 is the IEEE quotient
            // there); float sources may hold anything and take the division itself
            if (a.use_mul) v = (sd.dtype == 0) ? div255(v) : __fdiv_rn(v, 255.0f);
            r[c] = v;
        }
    }
    float* o = a.out + ((long long)b * per + idx) * 3;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}

// uint8 -> float32 with the reference's `/ 255` (core/load_tfrecords.py:46): only used when the stem conv has no uint8
// variant of its own (stride-2 or CUDA-core stems)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                        long long n, float div) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        out[i] = __fdiv_rn((float)in[i], div);
}

}  // namespace y3
