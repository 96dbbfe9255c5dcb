// CTA-pair (cta_group::2) variant of the implicit-GEMM conv: one tcgen05.mma spans the tensor cores of two SMs.
//
// Why: with 128 x 256 tiles on one SM, every K block writes 48 KB into shared memory (TMA) and reads 48 KB out of it
// (UMMA operands) per 512 tensor cycles = 188 B/clk against the 128 B/clk an SM's shared memory can move, so the single-CTA
// kernel tops out near 55 % tensor utilisation.  A CTA pair computes a 256 x BLOCK_N tile: each CTA stages its own 128
// pixel rows of A and only HALF of the weight rows; the hardware feeds both halves to both tensor cores.  Per SM and
// K block: 32 KB in + 32 KB out = 128 B/clk for the same 512 tensor cycles, and the smaller stage leaves room for 6
// pipeline stages instead of 4.
//
// Protocol (rank 0 = leader):
//   full[s]    leader's barrier: count 1 (leader producer's arrive.expect_tx of BOTH CTAs' bytes); both CTAs' TMA
//              loads signal it (cta_group::2 loads may complete on the peer CTA's barrier)
//   empty[s]   one per CTA, count 1: tcgen05.commit of the leader's MMA warp, multicast to both CTAs
//   tfull[a]   one per CTA, count 1: commit multicast after the last K block -> each CTA's epilogue drains ITS 128 lanes
//   tempty[a]  leader's barrier, count 256: the 128 epilogue threads of each CTA arrive (the peer remotely)
#pragma once
#include "conv_tc.cuh"

namespace y3 {

template <int BLOCK_N, int SWZ, int STAGES>
struct Conv2Smem {
    static constexpr int A_BYTES = kBlockM * SWZ;
    static constexpr int B_BYTES = (BLOCK_N / 2) * SWZ;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int XPOSE_BYTES = kConvEpiGroups * 4 * kEpiWarpBytes;
    static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16 + 8 * kEpiMaxBufs * 4 * kConvEpiGroups;
    static constexpr int TOTAL = 1024 + TILE_BYTES + XPOSE_BYTES + BAR_BYTES;
    // weights-resident variant: `stages` A tiles + all K blocks of this CTA's half of the weight tile
    static constexpr int total_resident(int stages, int num_k_blocks) {
        return 1024 + stages * A_BYTES + num_k_blocks * B_BYTES + XPOSE_BYTES + BAR_BYTES;
    }
};

// BRES (weights resident): when this CTA's half of the [BLOCK_N x K] weight tile fits in shared memory next to the A
// pipeline, it is loaded ONCE per launch -- before griddepcontrol.wait, i.e. while the previous layer is still
// finishing -- instead of once per output tile, and the stage ring carries only the A operand.  The layers it applies
// to (1x1 with K <= 768, 3x3 64->128) are bound by what the SM can pull from L2 (finding 7c/7d):
// the re-read weight tile was 25-33 % of that traffic.  Every cluster must then stay on one N tile: the host only
// picks this variant when the number of clusters is a multiple of tiles_n (tile % tiles_n is then constant).
// STAGES is the barrier-array size (maximum depth); the depth actually used is p.stages.
template <int BLOCK_N, int SWZ, int STAGES, bool BRES = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvArgs p) {
    using S = Conv2Smem<BLOCK_N, SWZ, STAGES>;
    constexpr int BLOCK_K = SWZ / 2;
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "UMMA N (cta_group::2: multiple of 16)");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int nst = BRES ? p.stages : STAGES;                       // pipeline depth in use
    const uint32_t tile_bytes = BRES ? (uint32_t)(nst * S::A_BYTES + p.num_k_blocks * S::B_BYTES) : (uint32_t)S::TILE_BYTES;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + (uint32_t)(nst * S::A_BYTES);   // BRES: the resident weights, else the B stages
    const uint32_t bar_base = smem_base + tile_bytes + S::XPOSE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t bfull_bar = bar_base + 8u * (2 * STAGES + 4);   // resident weights landed (leader's, both CTAs signal it)
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 5);
    auto res_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 5) + 16u + 8u * kEpiMaxBufs * w; };   // ring of epilogue warp w
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + tile_bytes + S::XPOSE_BYTES + 8 * (2 * STAGES + 5));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = (int)cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int num_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;   // (pairs of M tiles) x N tiles
    const int first_tile = (int)blockIdx.x / 2;
    const int tile_step = (int)gridDim.x / 2;

    if (threadIdx.x == 0) ts_mark(p.ts, 0);   // kernel entry
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 256);
        }
        for (int w = 0; w < kEpiMaxBufs * 4 * kConvEpiGroups; ++w) mbar_init(res_bar(0) + 8u * w, 1);
        mbar_init(bfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc2(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish2();
    }
    if (warp == 3 && lane == 0) {
        if (p.tma_out) {
            tma_prefetch_desc(&tmO);
            if (p.residual) tma_prefetch_desc(&tmR);
        }
        chain_gate(p);   // chained layers: the layer two launches back is complete before any thread lets the successor start
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if constexpr (BRES) {
        // the weights do not depend on the previous layer: fetch this CTA's half of the tile before waiting for it
        if (warp == 0 && elect_one()) {
            const int tn = tile_id(p, first_tile, num_tiles) % p.tiles_n;   // constant for this cluster (host-enforced)
            const int nb = tn * BLOCK_N + cta_rank * (BLOCK_N / 2);
            const uint32_t lead_bfull = mapa_shared(bfull_bar, 0);
            if (is_leader) mbar_arrive_expect_tx(bfull_bar, (uint32_t)(2 * p.num_k_blocks * S::B_BYTES));
            for (int kb = 0; kb < p.num_k_blocks; ++kb)
                tma2_load_2d(smem_b + kb * S::B_BYTES, &tmB, lead_bfull, kb * BLOCK_K, nb);
        }
    }
    if (threadIdx.x == 0) ts_mark(p.ts, 1);   // prologue done
    // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous layer's tail;
    // from here on we touch activations it wrote (and buffers it may still be reading), so wait for it to finish.
    // chained layers (ChainArgs) wait tile by tile in the producer / epilogue warps instead
    const bool chained = chain_enabled(p);
    pdl_launch_dependents();   // chained: the gate (chain_gate) was passed before the barrier above
    if (!chained) pdl_wait();
    if (threadIdx.x == 0) ts_mark(p.ts, 2);   // predecessor grid complete

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        const bool leader_lane = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        bool dep_all = false;   // chained: the whole input has been seen complete
        const int hw = p.Ho * p.Wo;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int tid_ = tile_id(p, tile, num_tiles);
            const int tmg = tid_ / p.tiles_n, tn = tid_ - tmg * p.tiles_n;
            const int tm = tmg * 2 + cta_rank;
            const int m0 = tm * kBlockM;
            const int nb = tn * BLOCK_N + cta_rank * (BLOCK_N / 2);   // this CTA's half of the weight rows
            int kcoord = 0;
            if (chained && !dep_all) dep_all = chain_wait_a(p, m0, lane);
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
                                if constexpr (BRES) {
                                    if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * S::A_BYTES);
                                    tma2_load_im2col_4d(smem_a + stage * S::A_BYTES, &tmA, lead_full, c0, cw, ch, cn,
                                                        (uint16_t)sx, (uint16_t)r);
                                } else {
                                    if (is_leader)
                                        mbar_arrive_expect_tx(full_bar(stage), (Y3_DBG_BITS(p) & 4) ? 2 * S::B_BYTES : 2 * S::STAGE_BYTES);
                                    if (!(Y3_DBG_BITS(p) & 4))
                                        tma2_load_im2col_4d(smem_a + stage * S::A_BYTES, &tmA, lead_full, c0, cw, ch, cn,
                                                            (uint16_t)sx, (uint16_t)r);
                                    tma2_load_2d(smem_b + stage * S::B_BYTES, &tmB, lead_full, kcoord, nb);
                                }
                            }
                            kcoord += BLOCK_K;
                            if (++stage == nst) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            } else {
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                    if (leader_lane) {
                        const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
                        if constexpr (BRES) {
                            if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * S::A_BYTES);
                            tma2_load_2d(smem_a + stage * S::A_BYTES, &tmA, lead_full, kcoord, m0);
                        } else {
                            if (is_leader)
                                mbar_arrive_expect_tx(full_bar(stage), (Y3_DBG_BITS(p) & 4) ? 2 * S::B_BYTES : 2 * S::STAGE_BYTES);
                            if (!(Y3_DBG_BITS(p) & 4)) tma2_load_2d(smem_a + stage * S::A_BYTES, &tmA, lead_full, kcoord, m0);
                            tma2_load_2d(smem_b + stage * S::B_BYTES, &tmB, lead_full, kcoord, nb);
                        }
                    }
                    kcoord += BLOCK_K;
                    if (++stage == nst) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (lane == 0) ts_mark(p.ts, 3);   // producer issued its last load
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader) {
            const bool leader_lane = elect_one();
            constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N);
            const uint64_t adesc0 = make_smem_desc<SWZ>(smem_a);
            const uint64_t bdesc0 = make_smem_desc<SWZ>(smem_b);
            int stage = 0;
            uint32_t phase = 0;
            int j = 0;
            if constexpr (BRES) {
                mbar_wait(bfull_bar, 0, 0x700);   // both halves of the weight tile are resident
                tc_fence_after();
            }
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++j) {
                const int acc = j & 1;
                mbar_wait(tempty_bar(acc), (uint32_t)(((j >> 1) & 1) ^ 1), 0x200 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x300 + stage);
                    tc_fence_after();
                    if (j == 0 && kb == 0 && lane == 0) ts_mark(p.ts, 4);   // first operands landed
                    if (leader_lane) {
                        const uint64_t adesc = adesc0 + (uint64_t)(stage * (S::A_BYTES >> 4));
                        const uint64_t bdesc = bdesc0 + (uint64_t)((BRES ? kb : stage) * (S::B_BYTES >> 4));
                        if (!(Y3_DBG_BITS(p) & 8)) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                           (uint32_t)((kb | k) != 0));
                        }
                        umma2_commit_mc(empty_bar(stage), 3);               // both producers may refill the stage
                        if (kb == p.num_k_blocks - 1) umma2_commit_mc(tfull_bar(acc), 3);   // both epilogues may drain
                    }
                    if (++stage == nst) { stage = 0; phase ^= 1u; }
                }
            }
            if (lane == 0) ts_mark(p.ts, 5);   // last MMA issued
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue groups (both CTAs, each drains its own 128 TMEM lanes) =====================
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        float* xp = reinterpret_cast<float*>(smem_gen + tile_bytes + (warp - 4) * kEpiWarpBytes);
        if (p.tma_out) {
            const uint32_t stg = smem_base + tile_bytes + (uint32_t)((warp - 4) * kEpiWarpBytes);
            const EpiTiles et{first_tile + eg * tile_step, kConvEpiGroups * tile_step, num_tiles, p.tiles_n, 2, cta_rank};
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * BLOCK_N);
            const uint32_t tempty = is_leader ? tempty_bar(eg) : mapa_shared(tempty_bar(eg), 0);
            if (p.tma_out == 64)
                epilogue_role_tma<64, 2>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty, !is_leader, p.ts, 7 + eg);
            else if (p.out_fp32)
                epilogue_role_tma<32, 2, true>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                               tempty, !is_leader, p.ts, 7 + eg);
            else
                epilogue_role_tma<32, 4>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty, !is_leader, p.ts, 7 + eg);
        } else {
            int j = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++j) {
                if ((j % kConvEpiGroups) != eg) continue;
                const int acc = j & 1;
                const int tid_ = tile_id(p, tile, num_tiles);
                const int tmg = tid_ / p.tiles_n, tn = tid_ - tmg * p.tiles_n;
                const int tm = tmg * 2 + cta_rank;
                mbar_wait(tfull_bar(acc), (uint32_t)((j >> 1) & 1), 0x400 + acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                epilogue_tile(p, BLOCK_N, tm * kBlockM, tn, t_row, q, lane, xp);
                tc_fence_before();
                if (is_leader) mbar_arrive(tempty_bar(acc));
                else mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
                epilogue_tile_post(p, tm * kBlockM + q * 32, lane);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
    if (threadIdx.x == 0) ts_mark(p.ts, 11);   // exit
}

}  // namespace y3
