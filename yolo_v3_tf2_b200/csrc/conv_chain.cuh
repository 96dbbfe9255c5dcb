// Persistent multi-layer conv kernel: ONE launch walks a run of consecutive conv layers (CTA-pair tcgen05 tiles, as
// conv_tc2.cuh), each CTA pair moving from its tiles of layer k straight to its tiles of layer k+1.
//
// Why: per launch a layer pays, on every SM, a ramp (prologue, TMEM allocation, cluster sync, first operands from a cold
// pipeline: ~4 us from the predecessor's last exit to the first MMA) and a tail (last MMA -> last tile's epilogue ->
// store drain -> teardown: ~6 us), plus the idle SMs of its last, partial round of tiles -- measured with per-CTA
// %globaltimer stamps (tools/net_timeline.py): ~10 of the ~80 us of a 3x3 128->256 @52 layer and ~10 of the ~34 us of a
// 1x1 256->128 @52 layer.  One CTA per SM (TMEM and shared memory are both full) means a successor launch cannot hide
// it: its CTA only starts when this one has exited.  Inside one kernel nothing has to be torn down between layers:
//   * the producer warp runs ahead into layer k+1 (operand ring, barriers and phases simply continue) as soon as the
//     128-row blocks of layer k's output that its next tile reads have been posted (ChainArgs flags, conv_tc.cuh),
//   * the MMA warp issues layer k+1's first tile while the epilogue warps are still draining layer k's last one,
//   * a pair without a tile in layer k's last partial round starts layer k+1 a tile earlier, and the host rotates which
//     pairs get the extra tile from layer to layer (ChainLayer::vshift) so the partial rounds even out over the run
//     instead of costing every layer a whole round.
// Layers of a run may differ in tile width (BLOCK_N 128 / 256, run-time here), filter size, stride, residual; what they
// share is the CTA-pair tiling, the 128-byte swizzle (Cin % 64 == 0) and the bf16 TMA-store epilogue in 64-column
// chunks.  Everything else (stem, Cin = 32 layers, fp32 heads, fused upsample) stays a launch of its own.
//
// Deadlock freedom: the grid is at most one CTA per SM, so every CTA is resident; a CTA only ever waits for tiles of
// EARLIER layers, which their owners reach in order without waiting for anything later.
// Buffer reuse: before a pair starts layer k it waits until layer k-2 is complete (ChainArgs::gate_done), so at most
// three consecutive layers are in flight and the arena planner keeps a buffer alive two launches past its last reader.
#pragma once
#include "conv_tc2.cuh"

namespace y3 {

struct alignas(128) ChainLayer {
    CUtensorMap tmA, tmB, tmO, tmR;   // A operand (2-D tiled or im2col), weights (box = block_n / 2 rows), output, residual
    ConvArgs p;                       // as for conv_tc2_kernel; p.ch wired to the run's flags (null = input older than the run)
    int block_n;                      // 128 or 256
    int vshift;                       // pair c walks the tile sequence of virtual pair (c + vshift) mod pairs
};

constexpr int kChainStages = 5;
struct ChainSmem {
    static constexpr int A_BYTES = kBlockM * 128;
    static constexpr int B_BYTES = 128 * 128;            // half of a 256-row weight tile; 128-wide tiles use the first half
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TILE_BYTES = kChainStages * STAGE_BYTES;
    static constexpr int XPOSE_BYTES = kConvEpiGroups * 4 * kEpiWarpBytes;
    static constexpr int BAR_BYTES = (2 * kChainStages + 4) * 8 + 16 + 8 * 2 * 4 * kConvEpiGroups;   // [+8]: tmem ptr, [+12]: dep counter
    static constexpr int TOTAL = 1024 + TILE_BYTES + XPOSE_BYTES + BAR_BYTES;
};
static_assert(ChainSmem::TOTAL <= 232448, "shared memory budget");

// what an epilogue warp carries from layer to layer
struct ChainEpiState {
    uint32_t g;        // chunks stored so far: chunk n lives in staging slot n % 2
    uint32_t jj;       // tiles of this group so far (accumulator phase)
    uint32_t rphase;   // bit b: parity of the residual barrier of staging slot b
};

// One layer's share of the TMA-store epilogue (epilogue_role_tma<64, 2> of conv_tc.cuh, with the ring, the accumulator
// phase and the residual barrier phases continuing from the previous layer).
__device__ __forceinline__ void chain_epilogue_layer(const ConvArgs& p, const CUtensorMap* tmO, const CUtensorMap* tmR,
                                                     int block_n, const EpiTiles et, uint32_t t_acc, int q, int lane,
                                                     uint32_t stg, uint32_t res_bar0, uint32_t tfull_bar,
                                                     uint32_t tempty_addr, bool tempty_remote, ChainEpiState& st) {
    constexpr int CW = 64, NBUF = 2;
    constexpr uint32_t ROW_BYTES = CW * 2;
    constexpr uint32_t BUF_BYTES = 32 * ROW_BYTES;
    constexpr int kPostLag = 2;   // chunks: ~1.4 us, about the completion latency of a bulk store
    const bool has_res = p.residual != nullptr;
    const float slope = p.leaky ? 0.1f : 1.0f;
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t row_off = (uint32_t)lane * ROW_BYTES;

    EpiCursor<CW> pr, pf, pp;   // chunk being processed / chunk whose residual is fetched next / tile posted next
    pr.tile = et.first;
    pr.load(p, et, block_n, q);
    pf = pr;
    pp = pr;
    uint32_t g = st.g, gp = st.g;
    const uint32_t g0 = st.g;
    uint32_t pp_gend = g0 + (pp.valid ? (uint32_t)pp.nch : 0u);   // value of g + 1 once pp's tile has been issued completely
    uint32_t posted = 0;
    bool res_all = p.ch.res_flags == nullptr;
    auto issue_res = [&]() {   // lane 0 only
        if (!res_all && pf.c == 0 && pf.row < p.M) {
            if (ld_acquire_gpu(p.ch.res_done) >= p.ch.res_total) res_all = true;
            else flag_wait_ge(p.ch.res_flags + (pf.row >> 7), p.ch.res_need, 0x920);
            fence_proxy_async_all();
        }
        const uint32_t b = gp % NBUF;
        mbar_arrive_expect_tx(res_bar0 + 8u * b, BUF_BYTES);
        tma_load_2d(stg + b * BUF_BYTES, tmR, res_bar0 + 8u * b, pf.ncol, pf.row);
    };
    if (has_res && pf.valid) {
        // the residual lands in the staging slot of its chunk: the previous layer's stores must have read it first
        if (lane == 0) {
            tma_store_wait_read<0>();
            issue_res();
        }
        pf.next(p, et, block_n, q);
        ++gp;
    }

#pragma unroll 1
    while (pr.valid) {
        if (pr.c == 0) {
            mbar_wait(tfull_bar, st.jj & 1u, 0x400);
            tc_fence_after();
        }
        const uint32_t b = g % NBUF;
        const uint32_t buf = stg + b * BUF_BYTES;
        if (!has_res) {
            if (lane == 0) tma_store_wait_read<NBUF - 1>();
            __syncwarp();
        }
        uint32_t v[CW];
        tmem_ld_32x32(t_acc + (uint32_t)(pr.c * CW), *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32(t_acc + (uint32_t)(pr.c * CW + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[CW - 32]));
        const float4* bp = reinterpret_cast<const float4*>(p.bias + pr.ncol);
        float4 bz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bz[j] = __ldg(bp + j);
        tmem_ld_wait();
        if (pr.c == pr.nch - 1) {
            tc_fence_before();
            if (tempty_remote) mbar_arrive_cluster(tempty_addr);
            else mbar_arrive(tempty_addr);
            ++st.jj;
        }
        if (has_res) {
            mbar_wait(res_bar0 + 8u * b, (st.rphase >> b) & 1u, 0x600 + b);
            st.rphase ^= 1u << b;
        }
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
            const float4 b0 = (j < 4) ? bz[2 * j] : __ldg(bp + 2 * j), b1 = (j < 4) ? bz[2 * j + 1] : __ldg(bp + 2 * j + 1);
            float f[8];
            {
                const float2 s2 = make_float2(slope, slope);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 y = __fadd2_rn(make_float2(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1])),
                                                make_float2(bb[2 * e], bb[2 * e + 1]));
                    const float2 z = __fmul2_rn(y, s2);
                    f[2 * e] = fmaxf(y.x, z.x);
                    f[2 * e + 1] = fmaxf(y.y, z.y);
                }
            }
            const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
            if (has_res) {
                const uint4 r = ld_shared_v4_relaxed(addr);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 o = __fadd2_rn(make_float2(f[2 * e], f[2 * e + 1]), __bfloat1622float2(h2[e]));
                    f[2 * e] = o.x;
                    f[2 * e + 1] = o.y;
                }
            }
            __nv_bfloat162 o2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            st_shared_v4_relaxed(addr, *reinterpret_cast<uint4*>(o2));
        }
        fence_proxy_async_smem();
        __syncwarp();
        const bool post_now = pp.valid && g + 1 >= pp_gend + kPostLag;
        if (lane == 0) {
            tma_store_2d(tmO, buf, pr.ncol, pr.row);
            tma_store_commit();
            if (post_now) {
                tma_store_wait<kPostLag>();   // everything but the kPostLag newest groups has completed
                chain_post(p, pp.row);
                ++posted;
            }
            if (has_res && pf.valid) {
                tma_store_wait_read<1>();
                issue_res();
            }
        }
        if (has_res && pf.valid) {
            pf.next(p, et, block_n, q);
            ++gp;
        }
        if (post_now) {
            pp.tile += et.step;
            pp.load(p, et, block_n, q);
            if (pp.valid) pp_gend += (uint32_t)pp.nch;
        }
        pr.next(p, et, block_n, q);
        ++g;
    }
    // end of the layer for this warp: everything it stored becomes visible, the rest of its tiles are posted
    if (lane == 0 && g != g0) {
        tma_store_wait<0>();
        while (pp.valid) {
            chain_post(p, pp.row);
            ++posted;
            pp.tile += et.step;
            pp.load(p, et, block_n, q);
        }
        chain_post_done(p, posted);
    }
    __syncwarp();
    st.g = g;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_chain_kernel(const ChainLayer* __restrict__ layers, int n_layers) {
    using S = ChainSmem;
    constexpr int STAGES = kChainStages;
    constexpr int BLOCK_K = 64;
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t ACC_COLS = 256;   // accumulator stage a sits at column a * 256 whatever the layer's tile width

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + (uint32_t)(STAGES * S::A_BYTES);
    const uint32_t bar_base = smem_base + S::TILE_BYTES + S::XPOSE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 4);
    auto res_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 4) + 16u + 8u * 2 * w; };   // 2 slots per epilogue warp
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + S::TILE_BYTES + S::XPOSE_BYTES + 8 * (2 * STAGES + 4));

    const uint32_t dep_cnt_smem = tmem_ptr_smem + 4u;   // tiles whose inputs the dependency warp has seen complete
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = (int)cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cid = (int)blockIdx.x / 2;
    const int pairs = (int)gridDim.x / 2;

    if (warp == 3 && lane == 0) st_shared_u32(dep_cnt_smem, 0u);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 256);
        }
        for (int w = 0; w < 2 * 4 * kConvEpiGroups; ++w) mbar_init(res_bar(0) + 8u * w, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc2(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    // the run as a whole waits for the launch before it; between its own layers only the tile flags order things
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        const bool leader_lane = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t tiles_seen = 0;   // tiles of this CTA so far, over all layers
#pragma unroll 1
        for (int li = 0; li < n_layers; ++li) {
            const ChainLayer* L = layers + li;
            const ConvArgs p = L->p;
            const CUtensorMap* tmA = &L->tmA;
            const CUtensorMap* tmB = &L->tmB;
            const int bn = L->block_n;
            const uint32_t stage_tx = 2u * (uint32_t)(S::A_BYTES + (bn / 2) * 128);
            const int num_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
            int vc = cid + L->vshift;
            if (vc >= pairs) vc -= pairs;
            const int hw = p.Ho * p.Wo;
            if (lane == 0) ts_mark(p.ts, 0);   // producer reaches the layer
#pragma unroll 1
            for (int tile = vc; tile < num_tiles; tile += pairs) {
                const int tid_ = tile_id(p, tile, num_tiles);
                const int tmg = tid_ / p.tiles_n, tn = tid_ - tmg * p.tiles_n;
                const int tm = tmg * 2 + cta_rank;
                const int m0 = tm * kBlockM;
                const int nb = tn * bn + cta_rank * (bn / 2);
                int kcoord = 0;
                // inputs of this tile complete?  (the dependency warp polls the global flags ahead of us)
                ++tiles_seen;
                if (ld_shared_acquire_u32(dep_cnt_smem) < tiles_seen) {
                    const long long t0 = clock64();
                    while (ld_shared_acquire_u32(dep_cnt_smem) < tiles_seen) {
                        if (clock64() - t0 > 8000000000LL) {
                            atomicExch(&g_watchdog_flag, 0x930u);
                            __threadfence_system();
                            __trap();
                        }
                    }
                }
                fence_proxy_async_all();   // the observation above (generic proxy) orders the TMA loads below (async proxy)
                if (p.a_im2col) {
                    const int cn = m0 / hw;
                    const int rem = m0 - cn * hw;
                    const int po = rem / p.Wo;
                    const int qo = rem - po * p.Wo;
                    const int cw = qo * p.stride + p.lower;
                    const int ch = po * p.stride + p.lower;
                    for (int r = 0; r < p.ksize; ++r) {
                        for (int sx = 0; sx < p.ksize; ++sx) {
                            for (int c0 = 0; c0 < p.kblocks_per_tap * BLOCK_K; c0 += BLOCK_K) {
                                mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                                if (leader_lane) {
                                    const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
                                    if (is_leader) mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                                    tma2_load_im2col_4d(smem_a + stage * S::A_BYTES, tmA, lead_full, c0, cw, ch, cn,
                                                        (uint16_t)sx, (uint16_t)r);
                                    tma2_load_2d(smem_b + stage * S::B_BYTES, tmB, lead_full, kcoord, nb);
                                }
                                kcoord += BLOCK_K;
                                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                            }
                        }
                    }
                } else {
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                        if (leader_lane) {
                            const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
                            if (is_leader) mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                            tma2_load_2d(smem_a + stage * S::A_BYTES, tmA, lead_full, kcoord, m0);
                            tma2_load_2d(smem_b + stage * S::B_BYTES, tmB, lead_full, kcoord, nb);
                        }
                        kcoord += BLOCK_K;
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader) {
            const bool leader_lane = elect_one();
            const uint64_t adesc0 = make_smem_desc<128>(smem_a);
            const uint64_t bdesc0 = make_smem_desc<128>(smem_b);
            int stage = 0;
            uint32_t phase = 0;
            int j = 0;   // tiles of this pair so far, over all layers: accumulator stage j & 1
#pragma unroll 1
            for (int li = 0; li < n_layers; ++li) {
                const ChainLayer* L = layers + li;
                const int bn = L->block_n;
                const int tiles_m = L->p.tiles_m, tiles_n = L->p.tiles_n, nkb = L->p.num_k_blocks;
                const uint32_t idesc = make_idesc_bf16(2 * kBlockM, bn);
                const int num_tiles = ((tiles_m + 1) / 2) * tiles_n;
                int vc = cid + L->vshift;
                if (vc >= pairs) vc -= pairs;
                unsigned long long* ts = L->p.ts;
                bool first = true;
#pragma unroll 1
                for (int tile = vc; tile < num_tiles; tile += pairs, ++j) {
                    const int acc = j & 1;
                    mbar_wait(tempty_bar(acc), (uint32_t)(((j >> 1) & 1) ^ 1), 0x200 + acc);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(full_bar(stage), phase, 0x300 + stage);
                        tc_fence_after();
                        if (first && lane == 0) ts_mark(ts, 4);   // first operands of the layer landed
                        first = false;
                        if (leader_lane) {
                            const uint64_t adesc = adesc0 + (uint64_t)(stage * (S::A_BYTES >> 4));
                            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (S::B_BYTES >> 4));
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                           (uint32_t)((kb | k) != 0));
                            umma2_commit_mc(empty_bar(stage), 3);
                            if (kb == nkb - 1) umma2_commit_mc(tfull_bar(acc), 3);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                if (lane == 0) ts_mark(ts, 5);   // last MMA of the layer issued
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== dependency warp (both CTAs) =====================
        // Walks the same tile sequence as the producer, ahead of it: for every tile it waits (global flags, ~1 us per
        // poll) until the layer two back is complete (gate, once per layer) and the rows the tile reads are written,
        // then bumps a shared-memory counter.  The producer only reads that counter, so the polling latency stays out of
        // its issue path (a poll per tile inside the producer cost more than the launch boundaries the run removes).
        uint32_t cleared = 0;
#pragma unroll 1
        for (int li = 0; li < n_layers; ++li) {
            const ChainLayer* L = layers + li;
            const ConvArgs p = L->p;
            if (p.ch.gate_done != nullptr) {   // layer li - 2 complete: its buffers may be recycled by this layer's output
                if (lane == 0) flag_wait_ge(p.ch.gate_done, p.ch.gate_total, 0x900);
                __syncwarp();
            }
            const int num_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
            int vc = cid + L->vshift;
            if (vc >= pairs) vc -= pairs;
            bool dep_all = p.ch.dep_flags == nullptr;
#pragma unroll 1
            for (int tile = vc; tile < num_tiles; tile += pairs) {
                if (!dep_all) {
                    const int tid_ = tile_id(p, tile, num_tiles);
                    const int tm = (tid_ / p.tiles_n) * 2 + cta_rank;
                    dep_all = chain_wait_a(p, tm * kBlockM, lane);
                }
                ++cleared;
                if (lane == 0) st_shared_release_u32(dep_cnt_smem, cleared);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue groups (both CTAs, each drains its own 128 TMEM lanes) =====================
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        const uint32_t stg = smem_base + S::TILE_BYTES + (uint32_t)((warp - 4) * kEpiWarpBytes);
        const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)eg * ACC_COLS;
        const uint32_t tempty = is_leader ? tempty_bar(eg) : mapa_shared(tempty_bar(eg), 0);
        ChainEpiState st{0u, 0u, 0u};
        int j0 = 0;   // tiles of this pair in the layers before
#pragma unroll 1
        for (int li = 0; li < n_layers; ++li) {
            const ChainLayer* L = layers + li;
            const ConvArgs p = L->p;
            const int num_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
            int vc = cid + L->vshift;
            if (vc >= pairs) vc -= pairs;
            const int n_my = vc < num_tiles ? (num_tiles - vc + pairs - 1) / pairs : 0;
            const int i0 = (eg - j0) & 1;   // this group takes the pair's tiles whose running number is eg mod 2
            const EpiTiles et{vc + i0 * pairs, kConvEpiGroups * pairs, num_tiles, p.tiles_n, 2, cta_rank};
            chain_epilogue_layer(p, &L->tmO, &L->tmR, L->block_n, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                 tempty, !is_leader, st);
            if (lane == 0 && q == 0) ts_mark(p.ts, eg == 0 ? 9 : 11);   // this group's share of the layer stored and posted
            j0 += n_my;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

}  // namespace y3
