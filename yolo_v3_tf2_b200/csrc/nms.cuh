// Class-agnostic padded NMS with the semantics of tf.image.non_max_suppression_padded(pad_to_max_output_size=True)
// as called by reference core/yolo_nms.py:26-33 (restated in oracle/nms_oracle.py):
//   1. masked score = score > score_thr ? score : 0 ; filtered boxes become all-zero boxes
//   2. stable descending sort by masked score (ties: lower index first)
//   3. greedy suppression in sorted order, a box is suppressed by an earlier surviving box when
//      iou >= iou_thr, iou = inter / (area_a + area_b - inter + 1e-8) in fp32 with separately rounded operations
//   4. selected = first max_boxes surviving boxes that have ANY coordinate > 0, mapped back to input indices,
//      zero padded; num_valid = their count.
//
// One CTA (1024 threads) per image:
//   select  : when more than kNmsSortAll candidates pass the score threshold, a 3-pass radix select (12 + 12 + 8 bits,
//             shared-memory histograms) finds the key of the kNmsTopT-th best candidate and only the candidates at or
//             above it are sorted and visited: the greedy loop stops at max_boxes survivors long before it runs out of
//             them (dense worst case: 10 647 candidates, ~600 visited).  If it does run out, the image is redone with
//             all candidates -- same result, just slower.
//   sort    : (score key, index) pairs in shared memory, bitonic network with a (key desc, index asc) comparator
//   suppress: candidates are consumed in sorted order in chunks of 256:
//             A  every candidate of the chunk is tested against the boxes kept so far (all threads),
//             B  256x256 upper-triangular IoU bitmask of the chunk built in shared memory (all threads),
//             C  one warp resolves the chunk sequentially with word-wide bit operations, visiting only survivors,
//             D  survivors are appended to the kept list / output (ballot + popc compaction).
//   The loop stops as soon as max_boxes valid boxes are selected.
#pragma once
#include "ptx.cuh"

namespace y3 {

constexpr int kNmsThreads = 1024;
constexpr int kNmsChunk = 256;
constexpr int kNmsKeptCap = 1024;
constexpr int kNmsMaxN = 32768;
constexpr int kNmsSortAll = 2048;   // up to this many candidates are simply sorted
constexpr int kNmsTopT = 1024;      // otherwise: the best kNmsTopT (plus ties with the last of them)
constexpr int kNmsTopCap = 4096;    // more than this many selected (mass ties): sort everything instead
constexpr int kNmsBins = 4096;

struct NmsArgs {
    const float* boxes;      // [B, N, 4]
    const float* scores;     // [B, N]
    int B, N, NP;            // NP = power of two >= N
    int max_boxes;
    float iou_thr, score_thr;
    int* selected;           // [B, max_boxes] int32, zero padded
    int* num_valid;          // [B]
    int* status;             // [B] 0 ok, 1 kept-list overflow (pathological input)
};

__device__ __forceinline__ uint32_t score_key(float s) {
    uint32_t u = __float_as_uint(s);
    if (u == 0x80000000u) u = 0u;   // -0 == +0 for ordering
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// IoU exactly as TF's _bbox_overlap: every operation separately rounded (no FMA contraction), IEEE division.
__device__ __forceinline__ float iou_tf(const float4 a, const float4 b) {
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fsub_rn(ix2, ix1), 0.0f);
    const float ih = fmaxf(__fsub_rn(iy2, iy1), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    const float area_a = __fmul_rn(__fsub_rn(a.w, a.y), __fsub_rn(a.z, a.x));
    const float area_b = __fmul_rn(__fsub_rn(b.w, b.y), __fsub_rn(b.z, b.x));
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-8f);
    return __fdiv_rn(inter, uni);
}

// counts[w] = items of warp w (32 warps): woff = items of the warps before `wid`, tot = all of them.  Every warp scans the
// 32 counts with shuffles (the first version had every thread add them up in a 32-iteration loop, per 1024-wide tile).
__device__ __forceinline__ void warp_counts_scan(const int* counts, int wid, int lane, int& woff, int& tot) {
    static_assert(kNmsThreads == 1024, "one count per lane");
    const int c = counts[lane];
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    tot = __shfl_sync(0xffffffffu, inc, 31);
    woff = __shfl_sync(0xffffffffu, inc - c, wid);
}

// iou_tf(a, b) >= thr for thr > 0, without the division when the boxes do not intersect: inter == 0 gives
// iou = 0 / uni = +-0 (or NaN when uni == 0), never >= a positive threshold.  Most pairs of a chunk do not overlap.
__device__ __forceinline__ bool iou_ge_pos(const float4 a, const float4 b, float thr) {
    const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
    const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
    const float iw = fmaxf(__fsub_rn(ix2, ix1), 0.0f);
    const float ih = fmaxf(__fsub_rn(iy2, iy1), 0.0f);
    const float inter = __fmul_rn(iw, ih);
    if (!(inter > 0.0f)) return false;
    const float area_a = __fmul_rn(__fsub_rn(a.w, a.y), __fsub_rn(a.z, a.x));
    const float area_b = __fmul_rn(__fsub_rn(b.w, b.y), __fsub_rn(b.z, b.x));
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-8f);
    return __fdiv_rn(inter, uni) >= thr;
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const NmsArgs a) {
    extern __shared__ __align__(16) uint8_t nsm[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(nsm);                       // [NP]
    uint16_t* idxs = reinterpret_cast<uint16_t*>(nsm + (size_t)a.NP * 4);    // [NP]
    uint8_t* tail = nsm + (size_t)a.NP * 6;
    float4* kept = reinterpret_cast<float4*>(tail);                          // [kNmsKeptCap]
    float4* cbox = kept + kNmsKeptCap;                                       // [kNmsChunk]
    uint32_t* mask = reinterpret_cast<uint32_t*>(cbox + kNmsChunk);          // [kNmsChunk][8]
    uint32_t* dead = mask + kNmsChunk * 8;                                   // [8] chunk-level dead bits
    uint32_t* validw = dead + 8;                                             // [8] chunk boxes with a coordinate > 0
    // radix-select histogram [kNmsBins]: aliases the kept list, which is only used after the sort (16 KB each; at
    // N = 22 743 (608x608) keys + indices alone take 192 KB of the 227 KB)
    uint32_t* hist = reinterpret_cast<uint32_t*>(kept);
    static_assert(kNmsBins * 4 <= kNmsKeptCap * 16, "histogram must fit in the kept list");
    __shared__ int s_nkept, s_nsel, s_overflow;
    __shared__ int s_warp_cnt[kNmsThreads / 32];
    __shared__ int s_scan[kNmsThreads / 32];
    __shared__ uint32_t s_sel_bin, s_sel_above;

    const int img = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const float* sc = a.scores + (long long)img * a.N;
    const float4* bx = reinterpret_cast<const float4*>(a.boxes) + (long long)img * a.N;
    int* sel = a.selected + (long long)img * a.max_boxes;

    // With score_thr >= 0 and iou_thr > 0 the filtered (all-zero) boxes sort after every passing box, never suppress
    // and are never selected, so only passing boxes need to be sorted/visited.  Otherwise keep all N candidates.
    const bool compact = (a.score_thr >= 0.0f) && (a.iou_thr > 0.0f);
    const bool pos_thr = a.iou_thr > 0.0f;

    if (tid == 0) { s_nkept = 0; s_nsel = 0; s_overflow = 0; }
    for (int i = tid; i < a.max_boxes; i += kNmsThreads) sel[i] = 0;
    __syncthreads();

    // attempt 0 may visit only the best candidates; attempt 1 (rare) redoes the image with all of them
    for (int attempt = 0; attempt < 2; ++attempt) {
    if (attempt == 1) {
        if (tid == 0) { s_nkept = 0; s_nsel = 0; s_overflow = 0; }
        for (int i = tid; i < a.max_boxes; i += kNmsThreads) sel[i] = 0;
        __syncthreads();
    }
    // ---------------- candidate list (index order) ----------------
    int ncand;
    if (compact) {
        // ordered compaction, one pass of 1024-wide tiles: ballot inside warps, running base across tiles
        int base = 0;
        for (int t0 = 0; t0 < a.N; t0 += kNmsThreads) {
            const int i = t0 + tid;
            const float s = (i < a.N) ? sc[i] : 0.0f;
            const bool pass = (i < a.N) && (s > a.score_thr);
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_warp_cnt[wid] = __popc(bal);
            __syncthreads();
            int woff, tot;
            warp_counts_scan(s_warp_cnt, wid, lane, woff, tot);
            if (pass) {
                const int pos = base + woff + __popc(bal & ((1u << lane) - 1u));
                keys[pos] = score_key(s);
                idxs[pos] = (uint16_t)i;
            }
            base += tot;
            __syncthreads();
        }
        ncand = base;
    } else {
        for (int i = tid; i < a.N; i += kNmsThreads) {
            const float s = sc[i];
            const float ms = (s > a.score_thr) ? s : 0.0f;
            keys[i] = score_key(ms);
            idxs[i] = (uint16_t)i;
        }
        ncand = a.N;
    }
    // ---------------- top-T selection (radix select on the 32-bit keys) ----------------
    bool partial = false;          // true: keys[0 .. ncand) hold only the best candidates of a longer list
    if (attempt == 0 && ncand > kNmsSortAll) {
        uint32_t prefix = 0u, pmask = 0u;
        int need = kNmsTopT;       // rank (1-based, from the top) of the key we are looking for
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = (pass == 0) ? 20 : (pass == 1 ? 8 : 0);
            const uint32_t bmask = (pass == 2) ? 0xFFu : 0xFFFu;
            for (int i = tid; i < kNmsBins; i += kNmsThreads) hist[i] = 0u;
            __syncthreads();
            // Detection scores share their exponent and leading mantissa bits, so in the first pass most of the 10 647
            // keys of a dense image fall into a handful of bins and same-address shared-memory atomics serialise -- across
            // the lanes of a warp AND across the 32 warps.  Hence two levels of aggregation: the lanes that hit the same
            // bin send one update (match_any), and while a whole warp keeps hitting ONE bin it only counts in a register
            // and flushes when the bin changes (one atomic per warp per run instead of one per 32 keys).
            uint32_t run_bin = 0xFFFFFFFFu, run_cnt = 0u;      // warp-uniform
            for (int t0 = 0; t0 < ncand; t0 += kNmsThreads) {
                const int i = t0 + tid;
                const uint32_t k = (i < ncand) ? keys[i] : 0u;
                const bool act = (i < ncand) && ((k & pmask) == prefix);
                const uint32_t bin = act ? ((k >> shift) & bmask) : 0xFFFFFFFFu;
                const uint32_t am = __ballot_sync(0xffffffffu, act);
                if (am == 0u) continue;
                const uint32_t bin0 = __shfl_sync(0xffffffffu, bin, __ffs(am) - 1);
                if (__all_sync(0xffffffffu, !act || bin == bin0)) {       // every active lane in the same bin
                    if (bin0 != run_bin) {
                        if (run_cnt != 0u && lane == 0) atomicAdd(&hist[run_bin], run_cnt);
                        run_bin = bin0;
                        run_cnt = 0u;
                    }
                    run_cnt += (uint32_t)__popc(am);
                } else {
                    const uint32_t grp = __match_any_sync(0xffffffffu, bin);
                    if (act && lane == __ffs(grp) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(grp));
                }
            }
            if (run_cnt != 0u && lane == 0) atomicAdd(&hist[run_bin], run_cnt);
            __syncthreads();
            // suffix sums from the top bin: thread t owns bins [4t, 4t+4), highest bins = highest t
            const uint32_t h0 = hist[4 * tid], h1 = hist[4 * tid + 1], h2 = hist[4 * tid + 2], h3 = hist[4 * tid + 3];
            const int mine = (int)(h0 + h1 + h2 + h3);
            // inclusive suffix scan over threads: warp level (lanes above), then warps above
            int suf = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += v;
            }
            if (lane == 0) s_scan[wid] = suf;
            __syncthreads();
            int above_warps = 0;
            for (int w = wid + 1; w < kNmsThreads / 32; ++w) above_warps += s_scan[w];
            const int incl = suf + above_warps;        // keys in bins >= 4t
            const int excl = incl - mine;              // keys in bins >= 4t + 4
            if (excl < need && incl >= need) {
                // the crossing is in one of my four bins (walk down from the highest)
                int acc = excl;
                const uint32_t hb[4] = {h0, h1, h2, h3};
                for (int b = 3; b >= 0; --b) {
                    if (acc + (int)hb[b] >= need) { s_sel_bin = (uint32_t)(4 * tid + b); s_sel_above = (uint32_t)acc; break; }
                    acc += (int)hb[b];
                }
            }
            __syncthreads();
            need -= (int)s_sel_above;
            prefix |= s_sel_bin << shift;
            pmask |= bmask << shift;
            __syncthreads();
        }
        // prefix is now the key of the kNmsTopT-th best candidate: keep every candidate with key >= prefix (ordered
        // in-place compaction, a tile's reads complete before its writes and writes never pass the reads)
        int base = 0;
        for (int t0 = 0; t0 < ncand; t0 += kNmsThreads) {
            const int i = t0 + tid;
            const uint32_t k = (i < ncand) ? keys[i] : 0u;
            const uint16_t x = (i < ncand) ? idxs[i] : (uint16_t)0;
            const bool pass = (i < ncand) && (k >= prefix);
            const uint32_t bal = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_warp_cnt[wid] = __popc(bal);
            __syncthreads();
            int woff, tot;
            warp_counts_scan(s_warp_cnt, wid, lane, woff, tot);
            if (pass && base + tot <= kNmsTopCap) {
                const int pos = base + woff + __popc(bal & ((1u << lane) - 1u));
                keys[pos] = k;
                idxs[pos] = x;
            }
            base += tot;
            __syncthreads();
        }
        if (base > kNmsTopCap) continue;   // mass ties: the list is damaged, redo with everything (attempt 1)
        partial = base < ncand;
        ncand = base;
    }
    // pad to a power of two with entries that sort last
    int np = 32;
    while (np < ncand) np <<= 1;
    for (int i = ncand + tid; i < np; i += kNmsThreads) { keys[i] = 0u; idxs[i] = 0xFFFFu; }
    __syncthreads();

    // ---------------- bitonic sort: descending key, ascending index on ties ----------------
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (np >> 1); t += kNmsThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                const int l = i | j;
                const uint32_t ki = keys[i], kl = keys[l];
                const uint16_t xi = idxs[i], xl = idxs[l];
                const bool i_first = (ki > kl) || (ki == kl && xi < xl);   // "i sorts before l"
                const bool want_first = ((i & k) == 0);                    // this sub-sequence is in final order
                if (i_first != want_first) {
                    keys[i] = kl; keys[l] = ki;
                    idxs[i] = xl; idxs[l] = xi;
                }
            }
            __syncthreads();
        }
    }

    // ---------------- greedy suppression over sorted candidates ----------------
    for (int c0 = 0; c0 < ncand; c0 += kNmsChunk) {
        const int nc = min(kNmsChunk, ncand - c0);
        const int nkept = s_nkept;
        // load chunk boxes (masked), clear dead bits
        if (tid < kNmsChunk) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < nc) {
                const int oi = idxs[c0 + tid];
                const bool pass = sc[oi] > a.score_thr;
                if (pass) b = bx[oi];
            }
            cbox[tid] = b;
            const uint32_t vb = __ballot_sync(0xffffffffu, b.x > 0.f || b.y > 0.f || b.z > 0.f || b.w > 0.f);
            if (lane == 0) validw[wid] = vb;
        }
        if (tid < 8) dead[tid] = 0u;
        __syncthreads();

        // A: chunk candidates vs kept boxes.  thread -> (candidate = tid % 256, kept subset = tid / 256)
        {
            const int ci = tid & (kNmsChunk - 1);
            if (ci < nc) {
                const float4 b = cbox[ci];
                bool d = false;
                if (pos_thr) {
                    for (int k = tid >> 8; k < nkept && !d; k += kNmsThreads / kNmsChunk) d = iou_ge_pos(kept[k], b, a.iou_thr);
                } else {
                    for (int k = tid >> 8; k < nkept && !d; k += kNmsThreads / kNmsChunk) d = iou_tf(kept[k], b) >= a.iou_thr;
                }
                if (d) atomicOr(&dead[ci >> 5], 1u << (ci & 31));
            }
        }
        __syncthreads();

        // B: intra-chunk bitmask. thread -> (row i = tid / 4, two 32-column words w = (tid % 4) * 2 + {0,1})
        {
            const int i = tid >> 2;
            const bool row_alive = (i < nc) && !((dead[i >> 5] >> (i & 31)) & 1u);
            const float4 bi = cbox[i];
#pragma unroll
            for (int ww = 0; ww < 2; ++ww) {
                const int w = ((tid & 3) << 1) + ww;
                uint32_t bits = 0u;
                if (row_alive && (w * 32 + 31) > i) {
                    const int jlo = max(w * 32, i + 1);
                    const int jhi = min(w * 32 + 32, nc);
                    if (pos_thr) {
                        for (int j = jlo; j < jhi; ++j)
                            if (iou_ge_pos(bi, cbox[j], a.iou_thr)) bits |= 1u << (j & 31);
                    } else {
                        for (int j = jlo; j < jhi; ++j)
                            if (iou_tf(bi, cbox[j]) >= a.iou_thr) bits |= 1u << (j & 31);
                    }
                }
                mask[i * 8 + w] = bits;
            }
        }
        __syncthreads();

        // C: sequential resolution by warp 0; lane w (< 8) owns dead word w.  The chunk is resolved in 8 blocks of 32 rows:
        // inside a block every lane runs the same register-only loop over the still-alive rows (the row's 32-bit
        // intra-block mask word comes from the lane that holds it, one shuffle per surviving row), then the survivors'
        // mask words are OR-ed into the dead words of the later blocks, one lane per word.  The loop stops at the
        // max_boxes-th valid survivor: nothing after it can be selected.  (The first version walked all 256 rows with a
        // ballot + two find-first-set + a shuffle + a shared-memory load per surviving row while the other 31 warps
        // waited at the barrier below: 40 % of the kernel.)
        if (wid == 0) {
            uint32_t dw = (lane < 8) ? dead[lane] : 0xffffffffu;
            if (lane < 8) {   // bits beyond nc are dead
                const int lo = lane * 32;
                if (nc <= lo) dw = 0xffffffffu;
                else if (nc < lo + 32) dw |= ~((1u << (nc - lo)) - 1u);
            }
            const uint32_t vw = (lane < 8) ? validw[lane] : 0u;
            int need = a.max_boxes - s_nsel;          // >= 1: the chunk loop stops once max_boxes are selected
            bool done = false;
            for (int sb = 0; sb < kNmsChunk / 32 && sb * 32 < nc; ++sb) {
                if (done) {                            // everything after the last needed survivor is irrelevant
                    if (lane == sb) dw = 0xffffffffu;
                    continue;
                }
                const uint32_t mrow = mask[(sb * 32 + lane) * 8 + sb];     // row (sb, lane): columns of the same block
                uint32_t d = __shfl_sync(0xffffffffu, dw, sb);
                const uint32_t vm = __shfl_sync(0xffffffffu, vw, sb);
                uint32_t todo = ~d, surv = 0u;
                while (todo != 0u) {
                    const int r = __ffs(todo) - 1;
                    surv |= 1u << r;
                    if (((vm >> r) & 1u) && --need == 0) { done = true; break; }
                    d |= __shfl_sync(0xffffffffu, mrow, r);
                    todo = ~d & ~((2u << r) - 1u);     // alive rows after r (r == 31: the mask is all ones)
                }
                if (lane == sb) dw = ~surv;
                if (lane > sb && lane < 8 && !done) {
                    uint32_t sv = surv;
                    while (sv != 0u) {
                        const int r = __ffs(sv) - 1;
                        sv &= sv - 1u;
                        dw |= mask[(sb * 32 + r) * 8 + lane];
                    }
                }
            }
            if (lane < 8) dead[lane] = dw;
        }
        __syncthreads();

        // D: append survivors in order (first kNmsChunk threads = 8 warps), count the valid ones
        if (tid < kNmsChunk) {
            const bool alive = !((dead[tid >> 5] >> (tid & 31)) & 1u);
            const float4 b = cbox[tid];
            const bool is_zero = (b.x == 0.f && b.y == 0.f && b.z == 0.f && b.w == 0.f);
            // all-zero boxes can only matter as suppressors when iou_thr <= 0
            const bool keep = alive && !(compact && is_zero);
            const bool valid = alive && (b.x > 0.f || b.y > 0.f || b.z > 0.f || b.w > 0.f);
            const uint32_t kb = __ballot_sync(0xffffffffu, keep);
            const uint32_t vb = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) { s_warp_cnt[wid] = __popc(kb); s_warp_cnt[8 + wid] = __popc(vb); }
            // named barrier over the 256 participating threads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            int koff = s_nkept, voff = s_nsel;
            for (int w = 0; w < wid; ++w) { koff += s_warp_cnt[w]; voff += s_warp_cnt[8 + w]; }
            const uint32_t lm = (1u << lane) - 1u;
            if (keep) {
                const int kp = koff + __popc(kb & lm);
                if (kp < kNmsKeptCap) kept[kp] = b; else s_overflow = 1;
            }
            if (valid) {
                const int vp = voff + __popc(vb & lm);
                if (vp < a.max_boxes) sel[vp] = (int)idxs[c0 + tid];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                int kt = 0, vt = 0;
                for (int w = 0; w < 8; ++w) { kt += s_warp_cnt[w]; vt += s_warp_cnt[8 + w]; }
                s_nkept = min(s_nkept + kt, kNmsKeptCap);
                s_nsel = s_nsel + vt;
            }
        }
        __syncthreads();
        if (s_nsel >= a.max_boxes || s_overflow) break;
    }
    // the best-candidates list ran out before max_boxes boxes were selected: lower-scored candidates may still qualify
    if (partial && s_nsel < a.max_boxes && !s_overflow) continue;
    break;
    }   // attempt
    if (tid == 0) {
        a.num_valid[img] = min(s_nsel, a.max_boxes);
        a.status[img] = s_overflow;
    }
}

// gather_valid_detections_results of reference inference.py:21-28, batched and zero padded to max_boxes rows.
__global__ void gather_detections_kernel(const float* __restrict__ boxes, const long long* __restrict__ cls,
                                         const float* __restrict__ scores, const int* __restrict__ selected,
                                         const int* __restrict__ num_valid, int B, int N, int max_boxes,
                                         float* __restrict__ out_boxes, long long* __restrict__ out_cls,
                                         float* __restrict__ out_scores, float* __restrict__ packed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * max_boxes) return;
    const int b = t / max_boxes, k = t - b * max_boxes;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    long long c = 0;
    float s = 0.f;
    if (k < num_valid[b]) {
        const long long src = (long long)b * N + selected[t];
        bb = reinterpret_cast<const float4*>(boxes)[src];
        c = cls[src];
        s = scores[src];
    }
    reinterpret_cast<float4*>(out_boxes)[t] = bb;
    out_cls[t] = c;
    out_scores[t] = s;
    if (packed != nullptr) {
        // the record the multi-GPU gather sends (distributed.pack_detections): per image max_boxes x (x1, y1, x2, y2,
        // score, class) then num_valid, all float32 (class ids < 2^24 are exact)
        float* rec = packed + (long long)b * (max_boxes * 6 + 1);
        float* r = rec + k * 6;
        r[0] = bb.x; r[1] = bb.y; r[2] = bb.z; r[3] = bb.w; r[4] = s; r[5] = (float)c;
        if (k == 0) rec[max_boxes * 6] = (float)num_valid[b];
    }
}

}  // namespace y3
