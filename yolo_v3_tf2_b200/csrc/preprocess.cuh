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

struct ImageDesc {          // one per image, device memory (8 x int64)
    long long src;          // device pointer: uint8 or float32 [H, W, 3]
    long long H, W;
    long long dtype;        // 0 uint8, 1 float32
    long long out_h, out_w; // resized size (== canvas size without aspect preservation)
    long long off_y, off_x; // top-left corner inside the canvas (pad_to_bounding_box)
};

struct PreprocessArgs {
    const ImageDesc* desc;
    float* out;             // [B, dst_h, dst_w, 3] float32
    int B, dst_h, dst_w;
    float mul;              // 1/255 for uint8 sources of the tfrecord path, else 1
    int use_mul;
};

__device__ __forceinline__ float load_px(const ImageDesc& d, long long idx) {
    if (d.dtype == 0) return (float)reinterpret_cast<const uint8_t*>(d.src)[idx];
    return reinterpret_cast<const float*>(d.src)[idx];
}

__global__ void preprocess_kernel(const PreprocessArgs a) {
    const long long per = (long long)a.dst_h * a.dst_w;
    const long long total = per * a.B;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / per);
        const int rem = (int)(e - (long long)b * per);
        const int y = rem / a.dst_w, x = rem - y * a.dst_w;
        const ImageDesc d = a.desc[b];
        float r[3] = {0.f, 0.f, 0.f};
        const int yy = y - (int)d.off_y, xx = x - (int)d.off_x;
        if (yy >= 0 && yy < d.out_h && xx >= 0 && xx < d.out_w) {
            const float sy = __fdiv_rn((float)d.H, (float)d.out_h), sx = __fdiv_rn((float)d.W, (float)d.out_w);
            const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)yy, 0.5f), sy), 0.5f);
            const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)xx, 0.5f), sx), 0.5f);
            const float fly = floorf(fy), flx = floorf(fx);
            const long long y0 = max((long long)fly, 0LL), y1 = min((long long)ceilf(fy), d.H - 1);
            const long long x0 = max((long long)flx, 0LL), x1 = min((long long)ceilf(fx), d.W - 1);
            const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tl = load_px(d, (y0 * d.W + x0) * 3 + c), tr = load_px(d, (y0 * d.W + x1) * 3 + c);
                const float bl = load_px(d, (y1 * d.W + x0) * 3 + c), br = load_px(d, (y1 * d.W + x1) * 3 + c);
                const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
                const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
                float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
                if (a.use_mul) v = __fdiv_rn(v, 255.0f);   // the reference divides: resize(...) / 255
                r[c] = v;
            }
        }
        float* o = a.out + e * 3;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    }
}

}  // namespace y3
