// 3x3 stride-1 'same' convolution on a zero-haloed flat activation layout -- the kernel for 93 % of YOLOv3's FLOPs.
//
// Problem it solves: an im2col-fed implicit GEMM fetches every input pixel 9 times (once per filter tap) and every
// weight once per 128-pixel tile; for a 256 x 256 tile with K = 1152 that is 1.18 MB through the L2 -> SM fabric for
// 151 MFLOP.  Ablations showed the layer still takes 39 of 93 us with the MMAs, the A loads and
// the epilogue math all removed: it is bound by that fabric traffic, not by the tensor pipe.
//
// Layout trick: the producer 1x1 conv stores its output as [B, H+1, W+1, C] with a zero last row / last column.  In
// that FLAT pixel order a 3x3 tap is a constant row offset:  in(p; r, s) = X[p + (r-1)*(W+1) + (s-1)], and every
// out-of-image neighbour lands on a zero halo pixel (or outside the tensor, which TMA zero-fills).  So for a tile of
// 128 consecutive flat pixels ONE patch of 130 + 2(W+1) pixels x 64 channels is staged, and the nine taps are nine
// UMMA shared-memory descriptors that differ only in their start row (the hardware swizzles on absolute shared-memory
// address bits, so a row-shifted window into a TMA-swizzled patch is exact -- tools_test_shift.py).  A traffic drops
// ~4-7x; K is ordered (channel block, r, s, c) so one patch serves 9 consecutive K blocks.
//
// CTA pair (cta_group::2): 256 flat pixels x block_n outputs per pair; each CTA stages its own patch and half of the
// weight rows.  Outputs on halo positions are discarded by the epilogue's row map (output and residual are dense).
#pragma once
#include "conv_tc2.cuh"

namespace y3 {

template <int SWZ>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvArgs p) {
    constexpr int BLOCK_K = SWZ / 2;
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr int ACC_STRIDE = 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int PST = p.pst, BST = p.bst;
    const uint32_t box_bytes = (uint32_t)p.box_rows * SWZ;
    const uint32_t patch_tx = box_bytes * (uint32_t)p.patch_boxes;           // bytes one CTA's patch brings in
    const uint32_t patch_bytes = (patch_tx + 1023u) & ~1023u;                // stage pitch (keeps 1024-byte alignment)
    const uint32_t b_bytes = (uint32_t)(p.block_n / 2) * SWZ;
    const uint32_t smem_p = smem_base;
    const uint32_t smem_b = smem_p + PST * patch_bytes;
    const uint32_t xpose_off = PST * patch_bytes + BST * b_bytes;
    float* xpose = reinterpret_cast<float*>(smem_gen + xpose_off);
    const uint32_t bar_base = smem_base + xpose_off + kConvEpiGroups * 4 * kXposeWarpFloats * 4;
    auto pfull_bar = [&](int s) { return bar_base + 8u * s; };
    auto pempty_bar = [&](int s) { return bar_base + 8u * (PST + s); };
    auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * PST + s); };
    auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * PST + BST + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PST + 2 * BST + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PST + 2 * BST + 2 + a); };
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * PST + 2 * BST + 4);
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(
        smem_gen + xpose_off + kConvEpiGroups * 4 * kXposeWarpFloats * 4 + 8 * (2 * PST + 2 * BST + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = (int)cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int num_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
    const int first_tile = (int)blockIdx.x / 2;
    const int tile_step = (int)gridDim.x / 2;
    const int wp = p.it_w;                      // flat row pitch in pixels (W + 1)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PST; ++s) { mbar_init(pfull_bar(s), 1); mbar_init(pempty_bar(s), 1); }
        for (int s = 0; s < BST; ++s) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
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
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        const bool leader_lane = elect_one();
        int ps = 0, bs = 0;
        uint32_t pphase = 0, bphase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int tmg = tile / p.tiles_n, tn = tile - tmg * p.tiles_n;
            const int p0 = (tmg * 2 + cta_rank) * kBlockM;
            const int row0 = p0 - wp - 1;                                   // first flat pixel of the patch (may be < 0)
            const int nb = tn * p.block_n + cta_rank * (p.block_n / 2);     // this CTA's half of the weight rows
            int kcoord = 0;
            for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(pempty_bar(ps), pphase ^ 1u, 0x100 + ps);
                if (leader_lane) {
                    const uint32_t lead = mapa_shared(pfull_bar(ps), 0);
                    if (is_leader) mbar_arrive_expect_tx(pfull_bar(ps), (Y3_DBG_BITS(p) & 4) ? 0u : 2 * patch_tx);
                    if (!(Y3_DBG_BITS(p) & 4))
                        for (int b = 0; b < p.patch_boxes; ++b)
                            tma2_load_2d(smem_p + ps * patch_bytes + b * box_bytes, &tmA, lead, cb * BLOCK_K,
                                         row0 + b * p.box_rows);
                }
                if (++ps == PST) { ps = 0; pphase ^= 1u; }
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(bempty_bar(bs), bphase ^ 1u, 0x180 + bs);
                    if (leader_lane) {
                        const uint32_t lead = mapa_shared(bfull_bar(bs), 0);
                        if (is_leader) mbar_arrive_expect_tx(bfull_bar(bs), (Y3_DBG_BITS(p) & 16) ? 0u : 2 * b_bytes);
                        if (!(Y3_DBG_BITS(p) & 16)) tma2_load_2d(smem_b + bs * b_bytes, &tmB, lead, kcoord, nb);
                    }
                    kcoord += BLOCK_K;
                    if (++bs == BST) { bs = 0; bphase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader) {
            const bool leader_lane = elect_one();
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, p.block_n);
            const uint64_t pdesc0 = make_smem_desc<SWZ>(smem_p);
            const uint64_t bdesc0 = make_smem_desc<SWZ>(smem_b);
            int ps = 0, bs = 0;
            uint32_t pphase = 0, bphase = 0;
            int j = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++j) {
                const int acc = j & 1;
                mbar_wait(tempty_bar(acc), (uint32_t)(((j >> 1) & 1) ^ 1), 0x200 + acc);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    mbar_wait(pfull_bar(ps), pphase, 0x300 + ps);
                    const uint64_t pdesc = pdesc0 + (uint64_t)(ps * (patch_bytes >> 4));
                    int tap_off = 0;                                 // (r * wp + s) * SWZ >> 4
                    for (int r = 0; r < 3; ++r) {
                        for (int sx = 0; sx < 3; ++sx) {
                            mbar_wait(bfull_bar(bs), bphase, 0x380 + bs);
                            tc_fence_after();
                            if (leader_lane) {
                                const uint64_t adesc = pdesc + (uint64_t)(((r * wp + sx) * SWZ) >> 4);
                                const uint64_t bdesc = bdesc0 + (uint64_t)(bs * (b_bytes >> 4));
                                if (!(Y3_DBG_BITS(p) & 8)) {
#pragma unroll
                                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                        umma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                                   (uint32_t)((cb | r | sx | k) != 0));
                                }
                                umma2_commit_mc(bempty_bar(bs), 3);
                            }
                            if (++bs == BST) { bs = 0; bphase ^= 1u; }
                        }
                    }
                    (void)tap_off;
                    if (leader_lane) {
                        umma2_commit_mc(pempty_bar(ps), 3);                                  // patch may be refilled
                        if (cb == p.cblocks - 1) umma2_commit_mc(tfull_bar(acc), 3);         // accumulator complete
                    }
                    if (++ps == PST) { ps = 0; pphase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue groups (both CTAs) =====================
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        float* xp = xpose + (warp - 4) * kXposeWarpFloats;
        int j = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++j) {
            if ((j % kConvEpiGroups) != eg) continue;
            const int acc = j & 1;
            const int tmg = tile / p.tiles_n, tn = tile - tmg * p.tiles_n;
            const int p0 = (tmg * 2 + cta_rank) * kBlockM;
            mbar_wait(tfull_bar(acc), (uint32_t)((j >> 1) & 1), 0x400 + acc);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
            epilogue_tile(p, p.block_n, p0, tn, t_row, q, lane, xp);
            tc_fence_before();
            if (is_leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
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
