// YOLO decode: the elementwise stage of reference core/yolo_decode_layer.py:4-36 as ONE memory-bound kernel.
//
// Input : 3 grids  t[s] = [B, gh_s, gw_s, 3, 5+C] fp32  (model outputs, reference core/parse_model.py:209-210)
// Output: bboxes [B, N, 4] (xmin, ymin, xmax, ymax; image fractions, unclipped), confidence [B, N, 1],
//         class_probs [B, N, C];  N = 3 * sum_s gh_s*gw_s, scales concatenated in model-output order,
//         flat index n = off_s + (i*gw + j)*3 + a.
// Optional fused tail (reference core/yolo_nms.py:18-24): scores[B,N] = conf * max_c prob, class_idx[B,N] = argmax_c
// (first max wins, int64) so the NMS stage does not have to re-read class_probs.
//
// Each CTA stages a contiguous run of records in shared memory with one 1-D bulk-TMA copy, then
//   (1) one thread per record does the box / objectness math (and the class max if requested),
//   (2) all threads stream sigmoid(class logits) back out with 16-byte stores (C % 4 == 0) or scalar stores.
#pragma once
#include "ptx.cuh"

namespace y3 {

constexpr int kDecodeThreads = 256;
constexpr int kDecodeRecs = 64;   // records per CTA (multiple of 4 keeps every chunk start 16-byte aligned)

struct DecodeArgs {
    const float* in[3];
    int gh[3], gw[3];
    int rec_off[3];        // first record of scale s inside one image
    int chunk_begin[4];    // prefix sum of per-scale chunk counts
    int pitch[3];          // floats between consecutive pixels (3*(5+C) dense, or the padded pitch of the head convs)
    int recs[3];           // records per chunk (kDecodeRecs; a multiple of 3 = whole pixels when the pitch is padded)
    float anchors[18];     // [3 scales][3 anchors][w, h]
    int B, C, N;
    float* bboxes;
    float* conf;           // conf and probs may both be null when scores / cls are requested ("compact" decode: the fused
    float* probs;          // pipeline's NMS only reads boxes and scores, so 225 of the 425 MB per 64 images are not written)
    float* scores;         // optional
    long long* cls;        // optional
    int stage_bytes;       // size of the record staging buffer (the record-index array follows it)
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Which class logits can reach the float32 maximum of sigmoidf_acc over a record whose largest logit is m (p_m =
// sigmoidf_acc(m))?  sigmoidf_acc(x) = sigmoid(x) (1 + e) with |e| <= 4e-7 (expf: 2 ulp on a term that enters with weight
// <= 1, one rounding each for the add and the divide), so sigmoidf_acc(x) < sigmoidf_acc(m) is certain once
// sigmoid(m) / sigmoid(x) > 1 + 8e-7.  d/dx ln sigmoid = 1 - sigmoid is decreasing, hence
// ln sigmoid(m) - ln sigmoid(x) >= (m - x) (1 - sigmoid(m)), and x < m - 2e-6 / (1 - p_m) is out of reach with a 2.5 x
// margin.  Where the sigmoid saturates (p_m == 1) the bound is -inf: every class stays a candidate.  Callers treat
// m < -80 (denormal / zero probabilities, where the relative error model does not hold) separately.
__device__ __forceinline__ float class_tie_threshold(float m, float p_m) {
    return (p_m < 1.0f) ? m - 2e-6f / (1.0f - p_m) : -INFINITY;   // (m = +inf would give inf - inf)
}

__global__ void __launch_bounds__(kDecodeThreads) decode_kernel(const DecodeArgs a) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int F = 5 + a.C;
    float* rec = reinterpret_cast<float*>(dsm);                         // staged records (pixel pitch as in memory)
    float4* rec_aux = reinterpret_cast<float4*>(dsm + a.stage_bytes);   // [kDecodeRecs] (cell column, cell row, anchor w, anchor h)
    int* out_rec = reinterpret_cast<int*>(rec_aux + kDecodeRecs);       // [kDecodeRecs] global record index
    int* rec_base = out_rec + kDecodeRecs;                               // [kDecodeRecs] float offset of the record in `rec`
    __shared__ __align__(8) uint64_t bar;

    int s = 0;
    if ((int)blockIdx.x >= a.chunk_begin[1]) s = 1;
    if ((int)blockIdx.x >= a.chunk_begin[2]) s = 2;
    // B * N <= 2^31 - 1 (host-checked), so records are counted in 32 bits; only byte offsets need 64
    const unsigned chunk = blockIdx.x - a.chunk_begin[s];
    const unsigned per_img = (unsigned)(a.gh[s] * a.gw[s] * 3);
    const unsigned total = (unsigned)a.B * per_img;
    const int recs = a.recs[s];
    const int pitch = a.pitch[s];
    const bool padded = pitch != 3 * F;       // then recs % 3 == 0 and a chunk starts at a pixel boundary
    const unsigned r0 = chunk * (unsigned)recs;
    const int nrec = (int)min((unsigned)recs, total - r0);
    const float* src = padded ? a.in[s] + (size_t)(r0 / 3u) * (size_t)pitch : a.in[s] + (size_t)r0 * (size_t)F;
    const int nfl = padded ? (nrec / 3) * pitch : nrec * F;
    const uint32_t bulk_bytes = ((uint32_t)nfl * 4u) & ~15u;
    // A pixel pitch that is a multiple of 32 floats would put record (pixel, anchor) of every pixel in the same three
    // banks (11-way conflicts in the per-record phases: measured 135 vs 92 us); such pixels are staged 4 floats further apart.
    const int spitch = (padded && (pitch & 31) == 0) ? pitch + 4 : pitch;
    auto rec_at = [&](int t) { return rec + rec_base[t]; };

    const uint32_t bar_s = smem_u32(&bar);
    if (threadIdx.x == 0) {
        mbar_init(bar_s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (spitch != pitch) {
        // one copy per pixel (pitch * 4 bytes, a multiple of 16), each issued by its own thread (one thread issuing all
        // 21 serialised ~1 us of the CTA's life).  A copy may complete before thread 0's arrive.expect_tx: the
        // transaction count then goes negative for a moment, and the phase cannot complete before that one arrival.
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bar_s, bulk_bytes);
        if ((int)threadIdx.x < nrec / 3)
            bulk_load_1d(smem_u32(rec + threadIdx.x * spitch), src + (long long)threadIdx.x * pitch, (uint32_t)pitch * 4u, bar_s);
    } else if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar_s, bulk_bytes);
        if (bulk_bytes) bulk_load_1d(smem_u32(rec), src, bulk_bytes, bar_s);
    }
    // Per-record tables, written by warps 2-3 while the copies are in flight: where record t of the chunk sits in the
    // staging buffer, where it goes, its cell and its anchor (the per-record phases are bound by their instruction count:
    // the divisions are done once per record here, not by every thread that touches the record).
    if ((int)threadIdx.x >= kDecodeRecs && (int)threadIdx.x < 2 * kDecodeRecs) {
        const int t = (int)threadIdx.x - kDecodeRecs;
        rec_base[t] = padded ? (t / 3) * spitch + (t - (t / 3) * 3) * F : t * F;
        const unsigned fr = r0 + (unsigned)min(t, max(nrec - 1, 0));
        const unsigned b = fr / per_img;
        const unsigned local = fr - b * per_img;
        const unsigned cell = local / 3u;
        const unsigned anc = local - cell * 3u;
        const unsigned gi = cell / (unsigned)a.gw[s];
        const unsigned gj = cell - gi * (unsigned)a.gw[s];
        out_rec[t] = (int)(b * (unsigned)a.N + (unsigned)a.rec_off[s] + local);
        rec_aux[t] = make_float4((float)gj, (float)gi, a.anchors[(s * 3 + (int)anc) * 2 + 0], a.anchors[(s * 3 + (int)anc) * 2 + 1]);
    }
    // the (at most 3) floats past the last 16-byte multiple
    for (int i = (int)(bulk_bytes >> 2) + threadIdx.x; i < nfl; i += kDecodeThreads) rec[i] = src[i];
    mbar_wait(bar_s, 0, 0x600);
    __syncthreads();

    const int C = a.C;
    if (a.probs == nullptr && a.conf == nullptr && a.scores != nullptr) {
        // ---- compact decode (the fused pipeline: NMS reads only boxes and scores), four threads per record ----
        // The kernel is bound by instructions issued per record (ncu: 82 % SM throughput at 0.40 of the HBM rate), so this
        // path is written for instruction count:
        //  * record -> (image, cell, anchor) index arithmetic is done ONCE per record by the two set-up warps above, while
        //    the bulk copies are in flight (rec_aux / out_rec), not by all four threads of the record;
        //  * Box: ONE exp + ONE division sequence per warp gives sigmoid t_x, sigmoid t_y, exp t_w, exp t_h in the four
        //    lanes of a record (part 0..3), ONE more division gives c_x, c_y, w/2, h/2 (x / 2 == x * 0.5 exactly); each
        //    lane then forms ONE box coordinate, so a warp stores 128 contiguous bytes with one instruction;
        //  * Class: what is needed is max_c sigmoid(t_c) and the FIRST class that attains it, not the C probabilities.
        //    sigmoid is increasing, so only classes whose logit is close to the largest one, m, can attain the float32
        //    maximum (class_tie_threshold).  One pass finds m, its first index and the runner-up; objectness and
        //    sigmoid(m) share ONE sigmoid sequence (lanes 0 / 1).  Only if the runner-up is within the threshold
        //    (near ties, saturation, tiny probabilities, NaN) are the candidates' sigmoids evaluated and reduced as
        //    (probability, lowest class).  Same operations on the same operands as the per-record path below: the result
        //    is bit-identical to the class reduce of the probabilities this kernel writes in its non-compact mode.
        static_assert(kDecodeThreads == 4 * kDecodeRecs, "four threads per record");
        const int t = (int)threadIdx.x >> 2, part = (int)threadIdx.x & 3;
        const bool live = t < nrec;
        const int tl = live ? t : 0;
        const float* r = rec + rec_base[tl];
        const int orec = out_rec[tl];
        const int base = (int)(threadIdx.x & 31u) & ~3;
        {
            const float aux = reinterpret_cast<const float*>(rec_aux)[tl * 4 + part];   // column, row, anchor w, anchor h
            const float x = r[part];
            const float ex = expf(part < 2 ? -x : x);                      // sigmoidf_acc(x) = 1 / (1 + expf(-x))
            const float sg = 1.0f / (1.0f + ex);
            // the reference divides (x, y) by (rows, cols) of the grid, see the per-record path
            const float num = part < 2 ? __fadd_rn(sg, aux) : __fmul_rn(ex, aux);
            const float den = part < 2 ? (float)(part == 0 ? a.gh[s] : a.gw[s]) : 2.0f;
            const float q = __fdiv_rn(num, den);                           // c_x, c_y, w/2, h/2
            const float c = __shfl_sync(0xffffffffu, q, base + (part & 1));
            const float h = __shfl_sync(0xffffffffu, q, base + 2 + (part & 1));
            const float coord = part < 2 ? __fsub_rn(c, h) : __fadd_rn(c, h);   // xmin, ymin, xmax, ymax
            if (live) a.bboxes[(size_t)orec * 4 + part] = coord;
        }
        // one pass: the largest logit m1 (first class attaining it: i1) and the runner-up value m2, per thread and then
        // merged over the record's four threads.  An exact tie sets m2 = m1, so it can never pass for "unique" below.
        // m2 = max(m2, min(m1, v)) is the runner-up update for v > m1 and for v <= m1 alike.  zsum = sum of 0 * v is NaN
        // exactly when some v is NaN or infinite: such records take the all-candidates path (where a NaN also makes
        // m2 = m1 here, it does not matter).
        float m1 = -INFINITY, m2 = -INFINITY, zsum = 0.0f;
        int i1 = 0x7fffffff;
        if (live) {
            const float* rc = r + 5 + part;
            const int n = (C - part + 3) >> 2;
#pragma unroll 10
            for (int k = 0; k < n; ++k) {
                const float v = rc[4 * k];
                zsum = fmaf(v, 0.0f, zsum);
                i1 = (v > m1) ? k : i1;
                m2 = fmaxf(m2, fminf(m1, v));
                m1 = fmaxf(m1, v);
            }
            i1 = (i1 == 0x7fffffff) ? i1 : 4 * i1 + part;
        }
        const bool nan = zsum != zsum;
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const float o1 = __shfl_xor_sync(0xffffffffu, m1, o);
            const float o2 = __shfl_xor_sync(0xffffffffu, m2, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i1, o);
            if (o1 > m1) { m2 = fmaxf(m1, o2); m1 = o1; i1 = oi; }
            else if (o1 == m1) { m2 = m1; i1 = min(i1, oi); }
            else m2 = fmaxf(m2, o1);
        }
        const bool any_nan = ((__ballot_sync(0xffffffffu, nan) >> base) & 0xFu) != 0u;
        // objectness (lane 0) and sigmoid(m1) (lane 1) in one sequence
        const float sg = sigmoidf_acc(part == 0 ? r[4] : m1);
        const float obj = __shfl_sync(0xffffffffu, sg, base);
        const float p_m = __shfl_sync(0xffffffffu, sg, base + 1);
        const float thr = (any_nan || !(m1 >= -80.0f)) ? -INFINITY : class_tie_threshold(m1, p_m);
        // runner-up below the threshold: the arg-max of the logits is the strict arg-max of the float32 probabilities and
        // its probability is the one just computed -- the common case costs one sigmoid per record
        const bool unique = m2 < thr;
        float best = p_m;
        int bi = i1;
        if (__any_sync(0xffffffffu, live && !unique)) {
            // some record of this warp has several candidates (near ties, saturation, NaN, tiny probabilities): its four
            // threads evaluate the sigmoid of every candidate and reduce (probability, lowest class)
            float sb = -INFINITY;
            int si = 0x7fffffff;
            if (live && !unique)
                for (int c = part; c < C; c += 4) {
                    const float v = r[5 + c];
                    if (v >= thr) {
                        const float pc = sigmoidf_acc(v);
                        if (pc > sb) { sb = pc; si = c; }      // classes ascend: the first maximum of this thread is kept
                    }
                }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, sb, o);
                const int oi = __shfl_xor_sync(0xffffffffu, si, o);
                if (ob > sb || (ob == sb && oi < si)) { sb = ob; si = oi; }
            }
            if (!unique) {
                best = sb; bi = si;
                // the sequential scan this replaces starts from class 0 and replaces it only by a strictly larger value:
                // a NaN in class 0 is never replaced, a NaN anywhere else never wins
                const float p0 = r[5];
                if (p0 != p0 || bi == 0x7fffffff) { best = sigmoidf_acc(p0); bi = 0; }
            }
        }
        if (live && part == 0) {
            a.scores[orec] = __fmul_rn(obj, best);
            a.cls[orec] = (long long)bi;
        }
        return;
    }

    // ---- (1) per-record box / objectness (+ class max) ----
    if ((int)threadIdx.x < nrec) {
        const int t = threadIdx.x;
        const long long orec = out_rec[t];
        const float4 aux = rec_aux[t];              // cell column, cell row, anchor w, anchor h
        float* r = rec_at(t);
        const float sx = sigmoidf_acc(r[0]);
        const float sy = sigmoidf_acc(r[1]);
        const float w = __fmul_rn(expf(r[2]), aux.z);
        const float h = __fmul_rn(expf(r[3]), aux.w);
        const float obj = sigmoidf_acc(r[4]);
        // The reference divides the (x, y) pair elementwise by tf.shape(xy)[1:3] = (gh, gw)
        // (yolo_decode_layer.py:5,8): x by the row count, y by the column count.  Identical for square grids; mirrored
        // literally so non-square inputs give the reference's numbers.
        const float cx = __fdiv_rn(__fadd_rn(sx, aux.x), (float)a.gh[s]);
        const float cy = __fdiv_rn(__fadd_rn(sy, aux.y), (float)a.gw[s]);
        const float hw = w * 0.5f, hh = h * 0.5f;
        float4 box;
        box.x = __fsub_rn(cx, hw);
        box.y = __fsub_rn(cy, hh);
        box.z = __fadd_rn(cx, hw);
        box.w = __fadd_rn(cy, hh);
        reinterpret_cast<float4*>(a.bboxes)[orec] = box;
        if (a.conf != nullptr) a.conf[orec] = obj;
        r[4] = obj;   // kept for the fused score in phase (3)
    }
    __syncthreads();

    // ---- (2) class probabilities ----
    if ((C & 3) == 0) {
        const int c4 = C >> 2;
        const int n4 = nrec * c4;
        for (int e = threadIdx.x; e < n4; e += kDecodeThreads) {
            const int ri = e / c4;
            const int c = (e - ri * c4) << 2;
            float* r = rec_at(ri) + 5 + c;
            float4 o;
            o.x = sigmoidf_acc(r[0]);
            o.y = sigmoidf_acc(r[1]);
            o.z = sigmoidf_acc(r[2]);
            o.w = sigmoidf_acc(r[3]);
            if (a.probs != nullptr) *reinterpret_cast<float4*>(a.probs + (long long)out_rec[ri] * C + c) = o;
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = o.w;
        }
    } else {
        const int n1 = nrec * C;
        for (int e = threadIdx.x; e < n1; e += kDecodeThreads) {
            const int ri = e / C;
            const int c = e - ri * C;
            float* r = rec_at(ri) + 5 + c;
            const float pc = sigmoidf_acc(*r);
            if (a.probs != nullptr) a.probs[(long long)out_rec[ri] * C + c] = pc;
            *r = pc;
        }
    }

    // ---- (3) fused class max: score = conf * max_c prob, class = first arg-max (reference core/yolo_nms.py:18-24) ----
    if (a.scores != nullptr) {
        __syncthreads();
        if ((int)threadIdx.x < nrec) {
            const float* r = rec_at((int)threadIdx.x);
            float best = r[5];
            int bi = 0;
            for (int c = 1; c < C; ++c) {
                const float pc = r[5 + c];
                if (pc > best) { best = pc; bi = c; }
            }
            const long long orec = out_rec[threadIdx.x];
            a.scores[orec] = __fmul_rn(r[4], best);
            a.cls[orec] = (long long)bi;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Class reduce for the stand-alone yolo_nms entry point (reference core/yolo_nms.py:18-24):
// class_idx = argmax_c probs (first max wins), score = conf * max_c probs.
//
// HBM-bound: N*C*4 bytes in per image, 12 bytes per record out.  Each CTA brings a contiguous run of `recs` records
// into shared memory with ONE 1-D bulk copy (no register staging, several CTAs per SM keep ~200 KB in flight), then
// one thread per record scans its C values out of shared memory.  The scan starts at a per-thread skewed class so the
// 32 lanes of a warp (record stride C floats) hit different banks; ties keep the lowest class index whatever the order.
// (The first version, one warp per record with a shuffle reduction, was bound by instruction issue: 0.28 of the copy
// bandwidth; it is kept below for unaligned pointers and very large C.)
// ------------------------------------------------------------------------------------------------
constexpr int kReduceThreads = 128;

__global__ void __launch_bounds__(kReduceThreads) class_reduce_kernel(const float* __restrict__ probs,
                                                                      const float* __restrict__ conf, long long nrec,
                                                                      int C, int recs, float* __restrict__ scores,
                                                                      long long* __restrict__ cls) {
    extern __shared__ __align__(16) uint8_t dsm[];
    float* rec = reinterpret_cast<float*>(dsm);
    __shared__ __align__(8) uint64_t bar;
    const long long r0 = (long long)blockIdx.x * recs;
    const int n = (int)min((long long)recs, nrec - r0);
    const float* src = probs + r0 * C;          // 16-byte aligned: recs % 4 == 0 and the base is (host-checked)
    const int nfl = n * C;
    const uint32_t bulk_bytes = ((uint32_t)nfl * 4u) & ~15u;
    const uint32_t bar_s = smem_u32(&bar);
    if (threadIdx.x == 0) {
        mbar_init(bar_s, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar_s, bulk_bytes);
        if (bulk_bytes) bulk_load_1d(smem_u32(rec), src, bulk_bytes, bar_s);
    }
    for (int i = (int)(bulk_bytes >> 2) + threadIdx.x; i < nfl; i += kReduceThreads) rec[i] = src[i];
    __syncthreads();                            // barrier init visible to the waiters, tail floats written
    mbar_wait(bar_s, 0, 0x680);
    for (int t = threadIdx.x; t < n; t += kReduceThreads) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        auto upd = [&](float v, int c) {
            if (v > best || (v == best && c < bi)) { best = v; bi = c; }
        };
        if ((C & 3) == 0) {
            const float4* r4 = reinterpret_cast<const float4*>(rec + (size_t)t * C);
            const int n4 = C >> 2;
            int j = t % n4;
            for (int i = 0; i < n4; ++i) {
                const float4 v = r4[j];
                upd(v.x, 4 * j); upd(v.y, 4 * j + 1); upd(v.z, 4 * j + 2); upd(v.w, 4 * j + 3);
                if (++j == n4) j = 0;
            }
        } else {
            const float* r = rec + (size_t)t * C;
            int j = (C & 1) ? 0 : t % C;        // odd C: the record stride is already conflict-free
            for (int i = 0; i < C; ++i) {
                upd(r[j], j);
                if (++j == C) j = 0;
            }
        }
        scores[r0 + t] = __fmul_rn(__ldg(conf + r0 + t), best);
        cls[r0 + t] = (long long)(bi == 0x7fffffff ? 0 : bi);
    }
}

// fallback: one warp per record, coalesced scalar loads, shuffle reduction
__global__ void __launch_bounds__(256) class_reduce_warp_kernel(const float* __restrict__ probs,
                                                           const float* __restrict__ conf, long long nrec, int C,
                                                           float* __restrict__ scores, long long* __restrict__ cls) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < nrec; r += nwarps) {
        const float* p = probs + r * C;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float v = __ldg(p + c);
            if (v > best) { best = v; bi = c; }   // strictly greater: earlier index wins inside a lane
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            scores[r] = __fmul_rn(__ldg(conf + r), best);
            cls[r] = (long long)(bi == 0x7fffffff ? 0 : bi);
        }
    }
}

}  // namespace y3
