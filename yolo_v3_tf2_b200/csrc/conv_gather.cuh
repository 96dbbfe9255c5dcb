// Implicit-GEMM conv whose A operand is built by software im2col ("gather") instead of TMA, with the whole weight
// matrix resident in shared memory.  Used where TMA's per-row issue rate, not bytes, is the limit:
//   * the 3-channel stem conv (reference backbone.yaml first conv, parse_model.py:13-56): K = 27 is padded to one
//     64-wide K block holding bf16(x) in columns 0..26 and bf16(x - bf16(x)) in columns 32..58 (hi/lo split, so the
//     fp32 image loses no precision); the weights are duplicated for both halves.
//   * 3x3 convs with Cin == 32 (64-byte rows): one K block per filter tap.
// Same tcgen05/TMEM pipeline and the same epilogue as conv_tc_kernel; only the producer differs:
//   the producer warps (128 threads per group) build the swizzled K-major tile the UMMA descriptor expects -- the stem
//   with 27 scalar loads per output pixel, the Cin == 32 layers with 16-byte cp.async (4 lanes per 64-byte row) --
//   fence the async proxy and arrive on the stage's mbarrier.  Warp 0 loads all K blocks of the weights once with TMA.
// The Cin == 32 path is opt-in (Y3_GATHER_CIN32=1): measured 0.45 ms (one producer group) / 0.38 ms (two) per layer
// against 0.27 ms for the TMA im2col kernel, so the planner keeps TMA for those layers.
#pragma once
#include <type_traits>

#include "conv_tc.cuh"

namespace y3 {

// warp layout: 0 weights loader, 1 MMA issuer, 2 TMEM allocator, 3 idle, then 4*NEPI epilogue warps, then 4*NPROD
// producer warps.  Tiles are dealt round-robin to the epilogue groups (group g owns accumulator stage g) and to the
// producer groups, so two tiles are gathered / drained concurrently: these layers are latency bound, not byte bound.
constexpr int kGatherEpiGroups = 2;

template <int NPROD>
constexpr int gather_threads() { return 32 * (4 + 4 * kGatherEpiGroups + 4 * NPROD); }

template <int BLOCK_N, int SWZ, int STAGES>
struct GatherSmem {
    static constexpr int A_BYTES = kBlockM * SWZ;
    static constexpr int B_BYTES = BLOCK_N * SWZ;
    static constexpr int XPOSE_BYTES = kGatherEpiGroups * 4 * kEpiWarpBytes;
    static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16 + 8 * kEpiMaxBufs * 4 * kGatherEpiGroups;
    static constexpr int total(int num_k_blocks) {
        return 1024 + STAGES * A_BYTES + num_k_blocks * B_BYTES + XPOSE_BYTES + BAR_BYTES;
    }
};

// byte offset of 16-byte chunk j of row r inside a K-major tile written with SWIZZLE_<SWZ>B
template <int SWZ>
__device__ __forceinline__ uint32_t swz_off(int r, int j) {
    if (SWZ == 128) return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
    return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
}

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills (padding halo / rows past M)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// COL (stride-1 stem only): the column-sharing producer described at the producer role below.
// U8 (COL only): the image is uint8 [B,H,W,3] and the network input is x / in_div (reference inference.py:157-158,
// core/load_tfrecords.py:46: `/ 255`).  A byte has 256 possible values, so the float division, the bf16 rounding and the
// bf16 remainder are looked up in a 256-entry shared-memory table (hi | lo << 16) built once per CTA with exactly the
// arithmetic of the float path -- results are bit-identical to feeding float32(x) / in_div, with a quarter of the input
// bytes and fewer producer instructions (9 LDS + 9 PRMT instead of ~40 conversion instructions per pixel column).
template <int BLOCK_N, int SWZ, int STAGES, bool STEM, int NPROD_ = (STEM ? 2 : 1), bool COL = false, bool U8 = false>
__global__ void __launch_bounds__(gather_threads<NPROD_>(), 1)
conv_gather_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmR, const ConvArgs p) {
    using S = GatherSmem<BLOCK_N, SWZ, STAGES>;
    static_assert(!COL || (STEM && SWZ == 128), "column producer is a stem variant");
    static_assert(!U8 || COL, "uint8 input is a variant of the column producer");
    __shared__ uint32_t u8_lut[U8 ? 256 : 1];
    constexpr int NEPI = kGatherEpiGroups;
    constexpr int NPROD = NPROD_;
    constexpr int BLOCK_K = SWZ / 2;
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(!STEM || SWZ == 128, "stem uses one 64-wide K block");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int nkb = p.num_k_blocks;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + STAGES * S::A_BYTES;
    const uint32_t tiles_bytes = STAGES * S::A_BYTES + nkb * S::B_BYTES;
    const uint32_t bar_base = smem_base + tiles_bytes + S::XPOSE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t bfull_bar = bar_base + 8u * (2 * STAGES + 4);
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 5);
    auto res_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 5) + 16u + 8u * kEpiMaxBufs * w; };   // ring of epilogue warp w
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + tiles_bytes + S::XPOSE_BYTES + 8 * (2 * STAGES + 5));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m;     // tiles_n == 1 (host-enforced)
    // number of tiles this CTA processes: tile(j) = blockIdx.x + j * gridDim.x
    const int my_tiles = (num_tiles > (int)blockIdx.x) ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        mbar_init(bfull_bar, 1);
        for (int w = 0; w < kEpiMaxBufs * 4 * NEPI; ++w) mbar_init(res_bar(0) + 8u * w, 1);
        fence_mbar_init();
    }
    if (warp == 3 && lane == 0 && p.tma_out) {
        tma_prefetch_desc(&tmO);
        if (p.residual) tma_prefetch_desc(&tmR);
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
    }
    if constexpr (U8) {
        if (threadIdx.x < 256) {
            const float f = __fdiv_rn((float)threadIdx.x, p.in_div);
            const __nv_bfloat16 hi = __float2bfloat16_rn(f);
            const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(f, __bfloat162float(hi)));
            u8_lut[threadIdx.x] = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
        }
    }
    if constexpr (COL) {
        // columns 54..63 of every row are read by the MMAs but never written by the producers: clear the stages once
        // (stale shared memory could hold NaN patterns, and NaN x 0-weight is NaN)
        for (uint32_t o = threadIdx.x * 16u; o < (uint32_t)(STAGES * S::A_BYTES); o += blockDim.x * 16u)
            st_shared_v4(smem_a + o, make_uint4(0u, 0u, 0u, 0u));
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous layer's tail;
    // from here on we touch activations it wrote (and buffers it may still be reading), so wait for it to finish.
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== weights: all K blocks, once =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(bfull_bar, (uint32_t)(nkb * S::B_BYTES));
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem_b + kb * S::B_BYTES, &tmB, bfull_bar, kb * BLOCK_K, 0);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
        const uint64_t adesc0 = make_smem_desc<SWZ>(smem_a);
        const uint64_t bdesc0 = make_smem_desc<SWZ>(smem_b);
        int stage = 0;
        uint32_t phase = 0;
        mbar_wait(bfull_bar, 0, 0x700);
        for (int j = 0; j < my_tiles; ++j) {
            const int acc = j & 1;
            mbar_wait(tempty_bar(acc), (uint32_t)(((j >> 1) & 1) ^ 1), 0x200 + acc);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(stage), phase, 0x300 + stage);
                fence_proxy_async_smem();   // producer writes came through the generic proxy (st.shared / cp.async)
                tc_fence_after();
                if (leader) {
                    const uint64_t adesc = adesc0 + (uint64_t)(stage * (S::A_BYTES >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(kb * (S::B_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(empty_bar(stage));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + 4 * NEPI) {
        // ===================== epilogue groups =====================
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        float* xp = reinterpret_cast<float*>(smem_gen + tiles_bytes + (warp - 4) * kEpiWarpBytes);
        if (p.tma_out) {
            const uint32_t stg = smem_base + tiles_bytes + (uint32_t)((warp - 4) * kEpiWarpBytes);
            const EpiTiles et{(int)blockIdx.x + eg * (int)gridDim.x, NEPI * (int)gridDim.x, num_tiles, 1, 1, 0};
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * BLOCK_N);
            if (BLOCK_N >= 64 && p.tma_out == 64)
                epilogue_role_tma<64, 2>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty_bar(eg), false, nullptr, 7 + eg);
            else
                epilogue_role_tma<32, 4>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty_bar(eg), false, nullptr, 7 + eg);
        } else {
            for (int j = eg; j < my_tiles; j += NEPI) {
                const int acc = j & 1;
                const int tile = tile_id(p, blockIdx.x + j * gridDim.x, num_tiles);
                mbar_wait(tfull_bar(acc), (uint32_t)((j >> 1) & 1), 0x400 + acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                epilogue_tile(p, BLOCK_N, tile * kBlockM, 0, t_row, q, lane, xp);
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                epilogue_tile_post(p, tile * kBlockM + q * 32, lane);
            }
        }
    } else if (warp >= 4 + 4 * NEPI) {
        // ===================== A producers: one output pixel (tile row) per thread =====================
        const int pg = (warp - 4 - 4 * NEPI) >> 2;
        const int row = (int)threadIdx.x - 32 * (4 + 4 * NEPI) - 128 * pg;
        const int hw = p.Ho * p.Wo;
        if constexpr (COL) {
            // Column-sharing stem producer (3x3, stride 1, pad 1).  The one-thread-per-output-pixel producer below loads
            // and converts 27 values per pixel although horizontally adjacent pixels share 18 of them, and the kernel
            // is bound by instructions issued per tile (finding 15).  Here a thread loads and converts only
            // the 3 rows x 3 channels of ITS pixel column (9 values -> 9 hi + 9 lo bf16 = two 16-byte pieces + 4 bytes)
            // and stores them three times: as filter column s = 1 of its own tile row, s = 2 of the row of the pixel to
            // its left and s = 0 of the row of the pixel to its right.  K layout of a row (weights packed to match), with
            // i = r*3 + c, hi = bf16(x), lo = bf16(x - hi): columns s*16 + [0,9) = hi_i, s*16 + 9 + [0,7) = lo_0..lo_6,
            // columns 48 + 2s, 49 + 2s = lo_7, lo_8; columns 54..63 are zero.  Every (row, s) slot has exactly one
            // writer: the neighbour in the same image row and tile, else the row's own thread (zeros at the image border,
            // or the halo column at a tile edge).
            using In = typename std::conditional<U8, uint8_t, float>::type;      // element of the image
            using Val = typename std::conditional<U8, uint32_t, float>::type;    // what a thread keeps per value
            const In* src = reinterpret_cast<const In*>(p.src);
            const int W = p.W;
            auto loadcol = [&](long long pix, int y, bool valid, Val (&v)[9]) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int yy = y - 1 + r;
                    const bool ok = valid && yy >= 0 && yy < p.H;
                    const In* px = src + (pix + (long long)(r - 1) * W) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[r * 3 + c] = ok ? (Val)__ldg(px + c) : (Val)0;   // byte 0 -> table entry (0, 0)
                }
            };
            auto convert = [&](const Val (&vv)[9], uint4& p0, uint4& p1, uint32_t& p2) {
                if constexpr (U8) {
                    uint32_t w[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) w[i] = u8_lut[vv[i]];
                    // w = hi | lo << 16: (hi_a, hi_b) = prmt 0x5410, (lo_a, lo_b) = prmt 0x7632, (hi_a, lo_b) = prmt 0x7610
                    p0 = make_uint4(__byte_perm(w[0], w[1], 0x5410), __byte_perm(w[2], w[3], 0x5410),
                                    __byte_perm(w[4], w[5], 0x5410), __byte_perm(w[6], w[7], 0x5410));
                    p1 = make_uint4(__byte_perm(w[8], w[0], 0x7610), __byte_perm(w[1], w[2], 0x7632),
                                    __byte_perm(w[3], w[4], 0x7632), __byte_perm(w[5], w[6], 0x7632));
                    p2 = __byte_perm(w[7], w[8], 0x7632);
                } else {
                float v[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) v[i] = (float)vv[i];
                __nv_bfloat162 h[4], l[4];
                float f[9];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                    const float2 t = __bfloat1622float2(h[i]);
                    f[2 * i] = t.x; f[2 * i + 1] = t.y;
                }
                const __nv_bfloat16 h8 = __float2bfloat16_rn(v[8]);
                f[8] = __bfloat162float(h8);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v[0] - f[0]);
#pragma unroll
                for (int i = 0; i < 4; ++i) l[i] = __floats2bfloat162_rn(v[2 * i + 1] - f[2 * i + 1], v[2 * i + 2] - f[2 * i + 2]);
                const __nv_bfloat162 mid = __halves2bfloat162(h8, l0);
                p0 = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                                *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
                p1 = make_uint4(*reinterpret_cast<const uint32_t*>(&mid), *reinterpret_cast<uint32_t*>(&l[0]),
                                *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]));
                p2 = *reinterpret_cast<uint32_t*>(&l[3]);
                }
            };
            // pixel of this thread in tile j, its column (and the halo column the tile-edge threads need)
            struct Pix { int m, y, x; bool valid, halo; };
            auto locate = [&](int j) {
                Pix q;
                q.m = tile_id(p, blockIdx.x + j * gridDim.x, num_tiles) * kBlockM + row;   // < 2^31 (host-checked)
                q.valid = (j < my_tiles) && (q.m < p.M) && !(Y3_DBG_BITS(p) & 4);
                const int n = q.m / hw;
                const int rem = q.m - n * hw;
                q.y = rem / W;
                q.x = rem - q.y * W;
                q.halo = (row == 0 && q.x > 0) || (row == kBlockM - 1 && q.x < W - 1);
                return q;
            };
            Pix qn = locate(pg);
            Val vn[9], hn[9];
            loadcol(qn.m, qn.y, qn.valid, vn);
            if (qn.halo) loadcol(qn.m + (row == 0 ? -1 : 1), qn.y, qn.valid, hn);
            for (int j = pg; j < my_tiles; j += NPROD) {
                const Pix q = qn;
                uint4 c0, c1, e0, e1;
                uint32_t c2, e2;
                convert(vn, c0, c1, c2);
                if (q.halo) convert(hn, e0, e1, e2);
                qn = locate(j + NPROD);
                loadcol(qn.m, qn.y, qn.valid, vn);
                if (qn.halo) loadcol(qn.m + (row == 0 ? -1 : 1), qn.y, qn.valid, hn);
                const int stage = j % STAGES;                      // one K block per tile
                const uint32_t phase = (uint32_t)((j / STAGES) & 1);
                mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                const uint32_t a_s = smem_a + stage * S::A_BYTES;
                auto put = [&](int drow, int sx, const uint4& p0, const uint4& p1, uint32_t p2) {
                    st_shared_v4(a_s + swz_off<128>(drow, 2 * sx), p0);
                    st_shared_v4(a_s + swz_off<128>(drow, 2 * sx + 1), p1);
                    st_shared_u32(a_s + swz_off<128>(drow, 6) + 4u * (uint32_t)sx, p2);
                };
                if (!(Y3_DBG_BITS(p) & 32)) {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    put(row, 1, c0, c1, c2);
                    if (row > 0 && q.x > 0) put(row - 1, 2, c0, c1, c2);
                    if (row < kBlockM - 1 && q.x < W - 1) put(row + 1, 0, c0, c1, c2);
                    if (q.x == 0) put(row, 0, z, z, 0u);
                    if (q.x == W - 1) put(row, 2, z, z, 0u);
                    if (q.halo) put(row, row == 0 ? 0 : 2, e0, e1, e2);
                }
                fence_proxy_async_smem();
                mbar_arrive(full_bar(stage));
            }
        } else if constexpr (STEM) {
            // 27 fp32 inputs -> hi/lo bf16 halves of one 64-wide K block; the loads of the group's next tile are in
            // flight while the current one is converted and stored
            const float* src = reinterpret_cast<const float*>(p.src);
            auto load27 = [&](int j, float (&v)[27]) {
                const int m = tile_id(p, blockIdx.x + j * gridDim.x, num_tiles) * kBlockM + row;
                const bool valid = (j < my_tiles) && (m < p.M) && !(Y3_DBG_BITS(p) & 4);   // Y3_DBG=4: no image loads (profiling)
                const int n = m / hw;
                const int rem = m - n * hw;
                const int po = rem / p.Wo;
                const int qo = rem - po * p.Wo;
                const int y0 = po * p.stride + p.lower;
                const int x0 = qo * p.stride + p.lower;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int y = y0 + r;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const int x = x0 + s;
                        const bool ok = valid && y >= 0 && y < p.H && x >= 0 && x < p.W;
                        const float* px = src + (((long long)n * p.H + y) * p.W + x) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = ok ? __ldg(px + c) : 0.0f;
                    }
                }
            };
            // Measured: 3 / 4 producer groups (without the register prefetch, 80 / 72 registers) take 0.51 / 0.65 ms against
            // 0.35 ms for 2 groups; with image loads, producer stores and epilogue all disabled (Y3_DBG=37) the kernel
            // still takes 0.24 ms -- it is bound by the instructions issued per 128 x 32 tile, not by latency or bytes.
            float vn[27];
            load27(pg, vn);
            for (int j = pg; j < my_tiles; j += NPROD) {
                float v[27];
#pragma unroll
                for (int i = 0; i < 27; ++i) v[i] = vn[i];
                load27(j + NPROD, vn);
                uint32_t w[32];   // 64 bf16: [0,27) hi, [32,59) lo, rest zero
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = 0u;
                // two values per step with the packed conversions (cvt.rn.bf16x2.f32): hi = bf16(x), lo = bf16(x - hi);
                // about a third of the instructions of the per-element shift-and-or packing it replaces
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    const float a0 = v[2 * i], a1 = (2 * i + 1 < 27) ? v[2 * i + 1] : 0.0f;
                    const __nv_bfloat162 hi2 = __floats2bfloat162_rn(a0, a1);
                    const float2 hf = __bfloat1622float2(hi2);
                    const __nv_bfloat162 lo2 = __floats2bfloat162_rn(a0 - hf.x, a1 - hf.y);
                    w[i] = *reinterpret_cast<const uint32_t*>(&hi2);
                    w[16 + i] = *reinterpret_cast<const uint32_t*>(&lo2);
                }
                const int stage = j % STAGES;                      // one K block per tile
                const uint32_t phase = (uint32_t)((j / STAGES) & 1);
                mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                const uint32_t a_s = smem_a + stage * S::A_BYTES;
                if (!(Y3_DBG_BITS(p) & 32)) {                               // Y3_DBG=32: no shared-memory stores (profiling)
#pragma unroll
                    for (int jc = 0; jc < 8; ++jc)
                        st_shared_v4(a_s + swz_off<128>(row, jc), make_uint4(w[4 * jc], w[4 * jc + 1], w[4 * jc + 2], w[4 * jc + 3]));
                }
                fence_proxy_async_smem();
                mbar_arrive(full_bar(stage));
            }
        } else {
            // bf16 input, Cin == BLOCK_K == 32 (64-byte rows): one K block per filter tap, copied with cp.async straight
            // into the swizzled stage (no register staging, up to STAGES taps in flight).
            // Lane mapping: 4 lanes x 16 bytes cover ONE 64-byte row, a warp instruction covers 8 consecutive rows, and
            // a thread handles rows sr, sr + 8, sr + 16, sr + 24 of its warp's 32-row block.  (The first version gave
            // every thread one whole row: a warp instruction then touched 32 different rows -- half-used 32-byte
            // sectors on the global side and 4-way bank conflicts on the swizzled shared-memory side; adding producer
            // warps made it slower, 0.47 -> 0.63 ms.)
            // With NPROD producer groups the K blocks (taps) are dealt round-robin: group pg fills every NPROD-th stage.
            static_assert(STEM || SWZ == 64, "lane mapping assumes 4 chunks per row");
            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.src);
            const int wrow = (row & ~31) + (lane >> 2);      // first of this thread's 4 rows
            const int jc = lane & 3;
            int stage = 0, turn = 0;
            uint32_t phase = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int m0 = tile_id(p, blockIdx.x + j * gridDim.x, num_tiles) * kBlockM + wrow;
                const __nv_bfloat16* base[4];
                int y0[4], x0[4];
                uint32_t dst[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = m0 + 8 * i;
                    const int n = m / hw;
                    const int rem = m - n * hw;
                    const int po = rem / p.Wo;
                    const int qo = rem - po * p.Wo;
                    // rows past M: push y0 out of range so that every tap is zero filled
                    y0[i] = (m < p.M) ? po * p.stride + p.lower : -4;
                    x0[i] = qo * p.stride + p.lower;
                    base[i] = src + (((long long)n * p.H + y0[i]) * p.W + x0[i]) * p.src_stride + jc * 8;
                    dst[i] = swz_off<64>(wrow + 8 * i, jc);
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    if (turn == pg) {
                        const int r = kb / p.ksize, sx = kb - r * p.ksize;
                        const long long tap = ((long long)r * p.W + sx) * p.src_stride;
                        mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                        const uint32_t a_s = smem_a + stage * S::A_BYTES;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = (unsigned)(y0[i] + r) < (unsigned)p.H && (unsigned)(x0[i] + sx) < (unsigned)p.W;
                            cp_async_16(a_s + dst[i], ok ? (const void*)(base[i] + tap) : (const void*)src, ok ? 16u : 0u);
                        }
                        cp_async_arrive_noinc(full_bar(stage));
                    }
                    if (++turn == NPROD) turn = 0;
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace y3
