// 3x3 stride-1 'same' conv with 32 input channels on a BAND-RESIDENT input (reference core/parse_model.py:13-56, 143-160;
// backbone.yaml: the 3x3 32 -> 64 conv of the first residual block; yolov3-tiny: 16 -> 32 and 32 -> 64).
//
// Why: fed by im2col TMA these layers read every input byte 9 - 12 times through L2 (profiles/r2_first_layers.md: 768 B
// per output pixel, ~9 TB/s, 0.25 ms against an HBM floor of 0.14 ms).  Here a CTA keeps a band of R + 2 input rows in
// shared memory, fetched ONCE with one tiled 4-D TMA box (32 channels x P pixels starting at x = -1 x (R + 2) rows
// starting at y0 - 1: the image border is zero filled by the TMA unit, no haloed layout in global memory), and the nine
// taps of a tile are nine UMMA descriptors into that band:
//   flat pixel v = i * P + x of the band (row pitch P), A row of output v for tap (r, s) = band pixel
//   v + r * P + s  ->  a tile of 128 consecutive flat pixels is a descriptor at byte (t * 128 + r * P + s) * 64
//   (the hardware swizzles on absolute shared-memory address bits, so a row-shifted window into the TMA-written band
//   is exact -- the same property conv_flat.cuh and conv_stem.cuh rely on).
// The row pitch P is W + 2 rounded up to a multiple of 32, so that a 32-pixel chunk of flat pixels (one epilogue warp's
// TMEM lanes) never straddles two image rows: flat pixels with x >= W are junk (computed, never stored -- the 3-D output
// map clips them), and so are the rows past the band in its last tile.  The epilogue is the TMA one: the residual (Add) of
// a chunk is TMA-loaded into the staging slot three chunks ahead, added in place, and the slot is TMA-stored.
// (A first version let a thread write its pixel with 256-bit global stores and load the residual the same way: correct,
// but every warp instruction then touches 32 different 128-byte lines and the layer was bound by the load/store unit --
// 0.231 ms with 8 epilogue warps, 0.260 ms with 16, against 0.106 ms for the main loop alone.)
// Weights (9 taps x BN rows x 64 B) are resident; accumulators are 8 TMEM stages of BN columns.
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-15 epilogue (three groups of four, tile j -> group j % 3).
#pragma once
#include "conv_stem.cuh"

namespace y3 {

struct BandArgs {
    int B, H, W;
    int R;                 // output rows per band; H % R == 0
    int P;                 // flat row pitch: W + 2 rounded up to a multiple of 32
    int T;                 // 128-pixel tiles per band = ceil(R * P / 128)
    int cout;              // stored output channels (= BN)
    const float* bias;     // [BN]
    int leaky;
    const __nv_bfloat16* residual;   // optional, dense pixel indexing
    long long res_stride;            // elements between consecutive pixels of the residual view
    int dbg;
};

constexpr int kBandEpiGroups = 3;         // epilogue groups of four warps; tile j of a CTA is drained by group j % 3
constexpr int kBandEpiWarps = 4 * kBandEpiGroups;
constexpr int kBandThreads = 32 * (4 + kBandEpiWarps);
constexpr int kBandAccStages = 8;
constexpr int kBandRing = 3;              // staging slots per epilogue warp (32 pixels x 32 channels = 2 KB each)
constexpr int kBandStgBytes = kBandRing * 2048;

__host__ __device__ inline int band_in_bytes(int R, int P) { return (R + 2) * P * 64; }                 // one TMA box
__host__ __device__ inline int band_pitch_bytes(int R, int P) { return (band_in_bytes(R, P) + 1023) & ~1023; }
template <int BN>
__host__ __device__ inline int band_smem_bytes(int R, int P) {
    return 1024 + 2 * band_pitch_bytes(R, P) + 9 * BN * 64 + kBandEpiWarps * kBandStgBytes + 512;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_tile_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kBandThreads, 1)
conv_band_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const BandArgs p) {
    static_assert(BN == 32 || BN == 64, "N tile = stored output channels");
    constexpr uint32_t TMEM_COLS = kBandAccStages * BN;          // 256 / 512
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int R = p.R, P = p.P, T = p.T;
    const uint32_t in_bytes = (uint32_t)band_in_bytes(R, P);
    const uint32_t in_pitch = (uint32_t)band_pitch_bytes(R, P);
    // bands first, then weights, staging, barriers: the rows a band's last tile reads past its end lie in the next region
    // -- junk rows of junk outputs, but inside the allocation
    const uint32_t smem_in = smem_base;
    const uint32_t smem_w = smem_in + 2u * in_pitch;
    const uint32_t smem_stg = smem_w + 9u * BN * 64u;
    const uint32_t bar_base = smem_stg + (uint32_t)(kBandEpiWarps * kBandStgBytes);
    auto in_full = [&](int b) { return bar_base + 8u * b; };
    auto in_empty = [&](int b) { return bar_base + 8u * (2 + b); };
    auto tfull = [&](int s) { return bar_base + 8u * (4 + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (4 + kBandAccStages + s); };
    const uint32_t wfull = bar_base + 8u * (4 + 2 * kBandAccStages);
    const uint32_t tmem_ptr_smem = bar_base + 8u * (5 + 2 * kBandAccStages);
    auto res_bar = [&](int w, int slot) { return bar_base + 8u * (6 + 2 * kBandAccStages) + 8u * (uint32_t)(w * kBandRing + slot); };
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bands_per_img = p.H / R;
    const int num_bands = p.B * bands_per_img;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        if (p.residual) tma_prefetch_desc(&tmR);
    }
    if (warp == 1 && lane == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(in_full(b), 1); mbar_init(in_empty(b), 1); }
        for (int s = 0; s < kBandAccStages; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 128); }
        mbar_init(wfull, 1);
        for (int w = 0; w < kBandEpiWarps * kBandRing; ++w) mbar_init(res_bar(0, 0) + 8u * w, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if (warp == 0 && elect_one()) {
        // the weights do not depend on the previous layer
        mbar_arrive_expect_tx(wfull, 9u * BN * 64u);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem_w + (uint32_t)tap * BN * 64u, &tmB, wfull, tap * 32, 0);
    }
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer: one box per band =====================
        const bool leader = elect_one();
        int k = 0;
        for (int band = blockIdx.x; band < num_bands; band += gridDim.x, ++k) {
            const int b = k & 1;
            const int n = band / bands_per_img;
            const int y0 = (band - n * bands_per_img) * R;
            mbar_wait(in_empty(b), (uint32_t)(((k >> 1) & 1) ^ 1), 0x100 + b);
            if (leader) {
                mbar_arrive_expect_tx(in_full(b), (p.dbg & 4) ? 0u : in_bytes);
                if (!(p.dbg & 4)) tma_load_tile_4d(smem_in + (uint32_t)b * in_pitch, &tmA, in_full(b), 0, -1, y0 - 1, n);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(128, BN);
        mbar_wait(wfull, 0, 0x700);
        tc_fence_after();
        const uint64_t bdesc0 = make_smem_desc<64>(smem_w);
        int j = 0, k = 0;
        for (int band = blockIdx.x; band < num_bands; band += gridDim.x, ++k) {
            const int b = k & 1;
            mbar_wait(in_full(b), (uint32_t)((k >> 1) & 1), 0x300 + b);
            tc_fence_after();
            const uint32_t in_b = smem_in + (uint32_t)b * in_pitch;
            for (int t = 0; t < T; ++t, ++j) {
                const int s = j & (kBandAccStages - 1);
                mbar_wait(tempty(s), (uint32_t)(((j / kBandAccStages) & 1) ^ 1), 0x200 + s);
                tc_fence_after();
                if (leader && !(p.dbg & 8)) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(s * BN);
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            const uint64_t adesc = make_smem_desc<64>(in_b + (uint32_t)(t * 128 + r * P + sx) * 64u);
                            const uint64_t bdesc = bdesc0 + (uint64_t)(((r * 3 + sx) * BN * 64) >> 4);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
                                umma_bf16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                          (uint32_t)((r | sx | kk) != 0));
                        }
                }
                if (leader) umma_commit(tfull(s));
            }
            if (leader) umma_commit(in_empty(b));
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue: three groups of four warps, tile j of the CTA goes to group j % 3 =====================
        // (an accumulator stage is drained by whichever group owns the tile: always exactly four warps = 128 arrivals)
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        constexpr int NH = BN / 32;
        const uint32_t stg = smem_stg + (uint32_t)(warp - 4) * kBandStgBytes;
        const float slope = p.leaky ? 0.1f : 1.0f;
        const bool has_res = p.residual != nullptr && !(p.dbg & 1);
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);          // SWIZZLE_64B: 16-byte piece index ^ ((row >> 1) & 3)
        const uint32_t row_off = (uint32_t)lane * 64u;
        // work items of a warp: (tile j of this CTA with j % 3 == eg, half h), in order; an item is VALID when its 32
        // flat pixels lie on an image row of the band and start left of W.  The cursor advances incrementally: a first
        // version recomputed band / tile / row with three integer divisions per item and per prefetch, 27 % of the
        // instructions the kernel executed (ncu source view).
        const int chunks_per_row = P >> 5;
        const uint32_t inv_cpr = (65536u + (uint32_t)chunks_per_row - 1u) / (uint32_t)chunks_per_row;   // c / cpr for c < 256
        struct Item {
            int j, h, t, band, ybase;     // tile counter of this CTA, half, tile inside the band, band, n * H + y0
            int x0, yrow;
            bool any, valid;
        };
        auto place = [&](Item& it) {      // (t, band, ybase) -> chunk position
            it.any = it.band < num_bands;
            const int c = it.t * 4 + q;                               // 32-pixel chunk index inside the band
            const int i = (int)(((uint32_t)c * inv_cpr) >> 16);       // image row of the band (P % 32 == 0)
            it.x0 = (c - i * chunks_per_row) << 5;
            it.yrow = it.ybase + i;
            it.valid = it.any && i < R && it.x0 < p.W;
        };
        auto set_band = [&](Item& it) {
            const int n = it.band / bands_per_img;
            it.ybase = n * p.H + (it.band - n * bands_per_img) * R;
        };
        auto next_item = [&](Item it) {
            if (it.h + 1 < NH) { ++it.h; return it; }
            it.h = 0;
            it.j += kBandEpiGroups;
            it.t += kBandEpiGroups;
            if (it.t >= T) {
                do { it.t -= T; it.band += (int)gridDim.x; } while (it.t >= T);
                set_band(it);
            }
            place(it);
            return it;
        };
        auto first_item = [&]() {
            Item it;
            it.j = eg; it.h = 0; it.t = eg; it.band = (int)blockIdx.x;
            while (it.t >= T) { it.t -= T; it.band += (int)gridDim.x; }
            set_band(it);
            place(it);
            return it;
        };
        Item cur = first_item(), pf = cur;
        uint32_t gv = 0, gp = 0;       // valid items processed / valid items whose residual load has been issued
        auto issue_res = [&]() {       // lane 0: residual of the next valid item of the prefetch cursor, if any
            while (pf.any && !pf.valid) pf = next_item(pf);
            if (!pf.any) return;
            const uint32_t slot = gp % kBandRing;
            const uint32_t bar = res_bar(warp - 4, (int)slot);
            mbar_arrive_expect_tx(bar, 2048u);
            tma_load_3d(stg + slot * 2048u, &tmR, bar, pf.h * 32, pf.x0, pf.yrow);
            pf = next_item(pf);
            ++gp;
        };
        if (has_res && lane == 0)
            for (int i = 0; i < kBandRing - 1; ++i) issue_res();
        while (cur.any) {
            const int s = cur.j & (kBandAccStages - 1);
            if (cur.h == 0) {
                mbar_wait(tfull(s), (uint32_t)((cur.j / kBandAccStages) & 1), 0x400 + s);
                tc_fence_after();
            }
            if (!cur.valid) {                               // junk pixels: nothing to read
                if (cur.h == NH - 1) {
                    tc_fence_before();
                    mbar_arrive(tempty(s));
                }
                cur = next_item(cur);
                continue;
            }
            const uint32_t slot = gv % kBandRing;
            const uint32_t buf = stg + slot * 2048u;
            if (!has_res) {
                if (lane == 0) tma_store_wait_read<kBandRing - 1>();   // the store that last used this slot has read it
                __syncwarp();
            }
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * BN + cur.h * 32), acc);
            const float4* bp = reinterpret_cast<const float4*>(p.bias + cur.h * 32);
            float4 bz[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bz[c] = __ldg(bp + c);
            tmem_ld_wait();
            if (cur.h == NH - 1) {                          // every TMEM read of this accumulator has completed
                tc_fence_before();
                mbar_arrive(tempty(s));
            }
            if (!(p.dbg & 1)) {
                if (has_res) mbar_wait(res_bar(warp - 4, (int)slot), (gv / kBandRing) & 1u, 0x600 + slot);
                const float* bzf = reinterpret_cast<const float*>(bz);
                const float2 s2 = make_float2(slope, slope);
#pragma unroll
                for (int c = 0; c < 4; ++c) {               // 16-byte piece c: channels 8c .. 8c+7 of this half
                    const uint32_t addr = buf + row_off + ((((uint32_t)c) ^ sw) << 4);
                    uint4 rr = make_uint4(0u, 0u, 0u, 0u);
                    if (has_res) rr = ld_shared_v4_relaxed(addr);
                    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
                    __nv_bfloat162 o2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = 8 * c + 2 * e;
                        const float2 y = __fadd2_rn(make_float2(__uint_as_float(acc[ch]), __uint_as_float(acc[ch + 1])),
                                                    make_float2(bzf[ch], bzf[ch + 1]));
                        const float2 z = __fmul2_rn(y, s2);
                        float2 o = make_float2(fmaxf(y.x, z.x), fmaxf(y.y, z.y));
                        if (has_res) o = __fadd2_rn(o, __bfloat1622float2(r2[e]));
                        o2[e] = __floats2bfloat162_rn(o.x, o.y);
                    }
                    st_shared_v4_relaxed(addr, *reinterpret_cast<uint4*>(o2));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (!(p.dbg & 2)) tma_store_3d(&tmO, buf, cur.h * 32, cur.x0, cur.yrow);
                    tma_store_commit();
                    if (has_res) {
                        // every store but the one just issued has read its slot: the slot of item gv - 1 is free again,
                        // and it is the slot of item gv + kBandRing - 1
                        tma_store_wait_read<1>();
                        issue_res();
                    }
                }
            }
            ++gv;
            cur = next_item(cur);
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace y3
