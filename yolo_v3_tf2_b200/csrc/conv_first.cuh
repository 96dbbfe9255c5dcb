// Direct convolution for the 3-channel stem conv (reference config/models/yolov3/backbone.yaml first entry:
// 3x3, stride 1, 32 filters, BN + leaky; built by core/parse_model.py:13-56).  K = 27 is too small for a tensor-core
// tile, and the layer is bound by its 416x416x32 output, so it runs on the CUDA cores: fp32 NHWC image in
// (inference.py:157-158 feeds float32 in [0,1]), BN-folded fp32 weights in shared memory, bf16 NHWC out.
// Also used (slow path) for any other conv whose Cin is not a multiple of 32.
#pragma once
#include "ptx.cuh"

namespace y3 {

struct ConvFirstArgs {
    const float* x;        // [B, H, W, CIN] fp32
    const float* w;        // [k*k*CIN][COUT] fp32, BN folded, K ordered (r, s, c)
    const float* bias;     // [COUT]
    __nv_bfloat16* out;    // view: pixel stride out_stride elements
    long long out_stride;
    int B, H, W, Ho, Wo;
    int ksize, stride, pad_lo;   // pad_lo = padding before (top/left)
    int leaky;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128) conv_first_kernel(const ConvFirstArgs a) {
    extern __shared__ __align__(16) float wsm[];   // [k*k*CIN][COUT] + [COUT]
    const int KK = a.ksize * a.ksize * CIN;
    for (int i = threadIdx.x; i < KK * COUT; i += blockDim.x) wsm[i] = a.w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) wsm[KK * COUT + i] = a.bias[i];
    __syncthreads();

    const long long M = (long long)a.B * a.Ho * a.Wo;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(m / (a.Ho * a.Wo));
        const int rem = (int)(m - (long long)n * a.Ho * a.Wo);
        const int po = rem / a.Wo, qo = rem - po * a.Wo;
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = wsm[KK * COUT + c];
        for (int r = 0; r < a.ksize; ++r) {
            const int y = po * a.stride - a.pad_lo + r;
            if (y < 0 || y >= a.H) continue;
            for (int s = 0; s < a.ksize; ++s) {
                const int xx = qo * a.stride - a.pad_lo + s;
                if (xx < 0 || xx >= a.W) continue;
                const float* px = a.x + (((long long)n * a.H + y) * a.W + xx) * CIN;
                const float* wt = wsm + ((r * a.ksize + s) * CIN) * COUT;
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float v = __ldg(px + ci);
                    const float4* w4 = reinterpret_cast<const float4*>(wt + ci * COUT);
#pragma unroll
                    for (int j = 0; j < COUT / 4; ++j) {
                        const float4 ww = w4[j];
                        acc[4 * j + 0] = fmaf(v, ww.x, acc[4 * j + 0]);
                        acc[4 * j + 1] = fmaf(v, ww.y, acc[4 * j + 1]);
                        acc[4 * j + 2] = fmaf(v, ww.z, acc[4 * j + 2]);
                        acc[4 * j + 3] = fmaf(v, ww.w, acc[4 * j + 3]);
                    }
                }
            }
        }
        uint4* op = reinterpret_cast<uint4*>(a.out + m * a.out_stride);
#pragma unroll
        for (int j = 0; j < COUT / 8; ++j) {
            __nv_bfloat162 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v0 = acc[8 * j + 2 * e], v1 = acc[8 * j + 2 * e + 1];
                if (a.leaky) {
                    v0 = v0 > 0.f ? v0 : 0.1f * v0;
                    v1 = v1 > 0.f ? v1 : 0.1f * v1;
                }
                h2[e] = __floats2bfloat162_rn(v0, v1);
            }
            op[j] = *reinterpret_cast<uint4*>(h2);
        }
    }
}

// ---------------- small stand-alone ops for graphs whose add / upsample / concat cannot be fused ----------------
// All operate on bf16 NHWC views (pixel stride in elements, channels a multiple of 8).
struct ViewArgs {
    const __nv_bfloat16* a;  long long a_stride;
    const __nv_bfloat16* b;  long long b_stride;
    __nv_bfloat16* out;      long long out_stride;
    long long npix;          // output pixels
    int C;                   // channels copied / added
    int Ho, Wo;              // output spatial dims (upsample only)
};

__global__ void add_views_kernel(const ViewArgs v) {   // reference parse_model.py:155-156  Add()([from, x])
    const int c8 = v.C >> 3;
    const long long total = v.npix * c8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pix = e / c8;
        const int c = (int)(e - pix * c8) << 3;
        const uint4 ua = *reinterpret_cast<const uint4*>(v.a + pix * v.a_stride + c);
        const uint4 ub = *reinterpret_cast<const uint4*>(v.b + pix * v.b_stride + c);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
        __nv_bfloat162 ho[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]);
            ho[i] = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
        }
        *reinterpret_cast<uint4*>(v.out + pix * v.out_stride + c) = *reinterpret_cast<uint4*>(ho);
    }
}

__global__ void copy_view_kernel(const ViewArgs v) {   // concat operand that could not be produced in place
    const int c8 = v.C >> 3;
    const long long total = v.npix * c8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pix = e / c8;
        const int c = (int)(e - pix * c8) << 3;
        *reinterpret_cast<uint4*>(v.out + pix * v.out_stride + c) =
            *reinterpret_cast<const uint4*>(v.a + pix * v.a_stride + c);
    }
}

__global__ void upsample2_view_kernel(const ViewArgs v) {   // reference parse_model.py:71-72 UpSampling2D(2), nearest
    const int c8 = v.C >> 3;
    const long long total = v.npix * c8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pix = e / c8;
        const int c = (int)(e - pix * c8) << 3;
        const int hw = v.Ho * v.Wo;
        const long long n = pix / hw;
        const int rem = (int)(pix - n * hw);
        const int y = rem / v.Wo, x = rem - y * v.Wo;
        const long long src = (n * (v.Ho >> 1) + (y >> 1)) * (v.Wo >> 1) + (x >> 1);
        *reinterpret_cast<uint4*>(v.out + pix * v.out_stride + c) =
            *reinterpret_cast<const uint4*>(v.a + src * v.a_stride + c);
    }
}

// reference parse_model.py:78-99  MaxPooling2D(pool_size, strides, padding): windows are clipped at the border, i.e. the
// 'same' padding never wins the max (yolov3-tiny: 2x2 stride 2, and 2x2 stride 1 'same' = window {x, x+1} clipped)
struct PoolArgs {
    const __nv_bfloat16* a;  long long a_stride;
    __nv_bfloat16* out;      long long out_stride;
    long long npix;          // output pixels
    int C;                   // channels (multiple of 8)
    int H, W, Ho, Wo;
    int size, stride, pad_lo;
};

__global__ void maxpool_view_kernel(const PoolArgs v) {
    const int c8 = v.C >> 3;
    const long long total = v.npix * c8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pix = e / c8;
        const int c = (int)(e - pix * c8) << 3;
        const int hw = v.Ho * v.Wo;
        const long long n = pix / hw;
        const int rem = (int)(pix - n * hw);
        const int yo = rem / v.Wo, xo = rem - yo * v.Wo;
        const int y0 = yo * v.stride - v.pad_lo, x0 = xo * v.stride - v.pad_lo;
        __nv_bfloat162 m[4];
        bool first = true;
        for (int dy = 0; dy < v.size; ++dy) {
            const int y = y0 + dy;
            if (y < 0 || y >= v.H) continue;
            for (int dx = 0; dx < v.size; ++dx) {
                const int x = x0 + dx;
                if (x < 0 || x >= v.W) continue;
                const uint4 u = *reinterpret_cast<const uint4*>(v.a + ((n * v.H + y) * v.W + x) * v.a_stride + c);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) m[i] = first ? h[i] : __hmax2(m[i], h[i]);
                first = false;
            }
        }
        if (first) {
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = __floats2bfloat162_rn(0.f, 0.f);
        }
        *reinterpret_cast<uint4*>(v.out + pix * v.out_stride + c) = *reinterpret_cast<uint4*>(m);
    }
}

}  // namespace y3
