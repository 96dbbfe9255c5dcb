// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces the Keras call chain ZeroPadding2D -> Conv2D -> BatchNormalization -> LeakyReLU (-> Add) (-> UpSampling2D)
// of reference core/parse_model.py:13-56 (conv), :143-160 (shortcut), :59-75 (upsample) for every conv whose input
// channel count is a multiple of 32 (all of Darknet-53 + neck + heads except the very first 3-channel conv).
//
//   GEMM view:  D[M, N] = A[M, K] * W[N, K]^T       M = B*Ho*Wo output pixels, N = Cout, K = k*k*Cin ordered (r, s, c)
//   A tile   :  128 pixels x BLOCK_K channels of ONE filter tap, fetched by TMA straight from the NHWC activation:
//               - 1x1 convs: plain 2-D tiled tensor map over [pixels, channels]
//               - 3x3 convs: im2col-mode tensor map (C, W, H, N); the hardware walks 128 consecutive output pixels,
//                 applies the traversal stride (1 or 2), adds the tap offset and zero-fills the padding halo.  The
//                 reference's asymmetric stride-2 padding ((1,0),(1,0)) then VALID (parse_model.py:34-35) is expressed
//                 by the bounding-box corners lower = -1, upper = 0 - (k-1).
//   W tile   :  BLOCK_N x BLOCK_K slice of the BN-folded bf16 weight matrix [Cout_pad, K] (2-D tiled tensor map)
//   D        :  fp32 accumulator in TMEM, 128 lanes x BLOCK_N columns, double buffered so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//
// Warp roles (384 threads, one persistent CTA per SM):
//   warp 0  TMA producer (one elected lane)         warp 1  MMA issuer (one lane issues tcgen05.mma)
//   warp 2  TMEM allocator                          warp 3  idle
//   warps 4-7 / 8-11  two epilogue groups, one per accumulator stage: tcgen05.ld -> +bias -> LeakyReLU(0.1) -> +residual -> bf16 (or fp32 for heads) -> global,
//              optionally replicated 2x2 (nearest upsample) into a channel slice of a wider buffer (concat).
#pragma once
#include "ptx.cuh"

namespace y3 {

// Cross-layer tile flags ("layer chaining").  A conv layer normally starts its main loop with
// griddepcontrol.wait, i.e. after its predecessor has finished completely: per launch that costs the predecessor's tail
// (last tile's epilogue, store drain, teardown), this layer's ramp (first operands from a cold pipeline) and the idle
// SMs of the predecessor's last, partial round of tiles.  With the flags a layer instead waits, tile by tile, for
// exactly the 128-row M tiles of its input that the tile reads:
//   producer  every epilogue warp, once the bulk stores of its 32 rows x BLOCK_N block have completed, adds 1 to
//             post_flags[row / 128] and to post_done (red.release.gpu)
//   consumer  the TMA producer warp polls dep_flags[t] >= dep_need for the M tiles t its next A tile reads (3x3: the
//             rows above and below included); the epilogue warps do the same for the residual rows; once
//             dep_done == dep_total the whole input is known to be complete and the polls stop
//   gate      before a CTA executes griddepcontrol.launch_dependents it waits until the layer TWO launches back has
//             completed (gate_done == gate_total).  Layer K+1 can therefore only be resident once layer K-2 is
//             finished: at most three layers are in flight, which is what the arena planner's live ranges assume
//             (a buffer is not reused until two launches after its last reader).
// All pointers are null when the feature is off (single-layer entry points, layers fed by a concat or a non-conv
// kernel): the kernel then uses griddepcontrol.wait as before.  Counters are zeroed at the start of every forward pass.
struct ChainArgs {
    uint32_t* post_flags;        // [CL * ceil(tiles_m / CL) + 1] arrivals per 128-row M tile of THIS layer's output
    uint32_t* post_done;         // arrivals over all tiles
    const uint32_t* dep_flags;   // producer of the A operand
    const uint32_t* dep_done;
    uint32_t dep_need, dep_total;
    int dep_h, dep_w;            // spatial size of the A operand (rows / columns per image)
    const uint32_t* res_flags;   // producer of the residual (same pixel indexing as this layer's output)
    const uint32_t* res_done;
    uint32_t res_need, res_total;
    const uint32_t* gate_done;   // the layer two launches back
    uint32_t gate_total;
    uint32_t mode;               // profiling build only (Y3_CHAIN_MODE): 1 = post with red.release instead of red.relaxed
};

struct ConvArgs {
    int M;                 // B*Ho*Wo
    int Ho, Wo;            // output spatial size
    int stride;            // 1 or 2
    int lower;             // im2col lower corner (= -pad_before) for w and h
    int a_im2col;          // 0: 2-D tiled A map (1x1), 1: im2col A map
    int ksize;             // 1 or 3
    // Pixel-pair view of a Cin = 32 input (conv_tc_kernel only; equal to ksize / stride / lower otherwise).  Two
    // horizontally adjacent pixels are ONE 64-channel "pixel" of 128 bytes, so a 3x3 conv becomes 3 rows x 2 pair columns
    // = 6 TMA rows of 128 B per output instead of 9 rows of 64 B (these layers are bound by the TMA unit's row rate).
    //   stride 2: output x reads pairs x-1 (odd half) and x            -> taps (r, S), S in {0,1}, w traversal stride 1
    //   stride 1: the GEMM row is an output PAIR and the N tile the parity j of the pixel inside it (tiles_n = 2, the
    //             output row = 2 x Cout channels): output 2X+j reads pairs X-1+j, X+j -> the tap offset is S + j
    //   (a_shift_n = 1).  Weights: K index (r, S, i, c) holds w[r][s = 2S + i + j - 1][c], zero when s is outside 0..2.
    int ksize_w, stride_w, lower_w;
    int a_shift_n;         // added to the im2col w offset per N tile index
    int kblocks_per_tap;   // Cin / BLOCK_K
    int num_k_blocks;      // ksize*ksize*kblocks_per_tap
    int tiles_m, tiles_n;
    int cout;              // valid output channels (<= tiles_n*BLOCK_N)
    const float* bias;     // [tiles_n*BLOCK_N] fp32 (folded BN shift or conv bias, zero padded)
    int leaky;             // 1: LeakyReLU(0.1)
    const __nv_bfloat16* residual;  // optional, same pixel indexing as the output (dense pixel index m)
    long long res_stride;  // elements between consecutive pixels of the residual view
    void* out;             // bf16 (or fp32 if out_fp32) view base
    long long out_stride;  // elements between consecutive pixels of the output view
    int out_fp32;
    int upsample;          // 1: write every output pixel to the 2x2 block (2p+dy, 2q+dx) of a (2Ho, 2Wo) view
    // Iteration space of the GEMM M index: per image an it_h x it_w grid.  Normally (Ho, Wo); (Ho+1, Wo+1) when the
    // conv walks the zero-haloed flat layout [B, Ho+1, Wo+1, C] of its input (rows with y == Ho or x == Wo are halo
    // positions and produce no output).  out_padded: the output itself is stored in that haloed layout (and the
    // epilogue writes the zero halo next to the last column / row).
    int it_h, it_w;
    int out_padded;
    // flat-patch 3x3 kernel only (conv_flat.cuh)
    int block_n;           // N tile (64 / 128 / 256), runtime there
    int cblocks;           // Cin / BLOCK_K
    int patch_boxes, box_rows;   // one patch = patch_boxes TMA boxes of box_rows haloed-flat pixels
    int pst, bst;          // pipeline depth: patches / weight blocks
    // gather kernel only (software im2col): the input view
    const void* src;       // bf16 NHWC view, or the fp32 [B,H,W,3] image for the stem
    long long src_stride;  // elements between consecutive input pixels
    int H, W;              // input spatial size
    int rot;               // tile sequence rotation: position seq of the sequence is tile (seq + rot) mod num_tiles (before
                           //    rev is applied); the planner picks it so that a chained layer starts on tiles whose input
                           //    its predecessor completed a round ago instead of on the ones still in flight
    ChainArgs ch;          // cross-layer tile flags, all null = off
    int rev;               // 1: walk the output tiles from the last to the first.  The planner alternates the direction from
                           //    layer to layer so that a layer starts on the pixels its predecessor wrote LAST, which are
                           //    still in L2 (each 52x52 tensor is 88 MB of a 126 MB L2)
    int stem_col;          // stem only: column-sharing producer (conv_gather.cuh), weights in its K order
    float in_div;          // stem only, uint8 image: the network input is (float)byte / in_div (255 in the reference)
    int stages;            // weights-resident kernels only: A-operand pipeline depth (what fits next to the weights)
    int tma_out;           // 0: register-transpose epilogue; 32 / 64: bf16 dense output written with TMA stores in chunks
                           //    of that many columns (epilogue_role_tma); tmO (and tmR when a residual is fused) are
                           //    tensor maps with a 32-row x tma_out-column box
    unsigned long long* ts;   // profiling: when non-null every CTA records %globaltimer at 12 points into ts[32*blockIdx.x + k]
    int dbg;               // profiling knobs (env Y3_DBG): 1 epilogue drains TMEM only, 2 no global stores,
                           // 4 producer skips the A loads, 8 no MMAs are issued (results are garbage)
};

constexpr int kConvEpiGroups = 2;   // epilogue warp groups; group g drains accumulator stage g (tiles j % 2 == g)
constexpr int kConvThreads = 32 * (4 + 4 * kConvEpiGroups);
constexpr int kBlockM = 128;

// position in a CTA's tile sequence -> tile id (see ConvArgs::rev)
__device__ __forceinline__ int tile_id(const ConvArgs& p, int seq, int num_tiles) {
    seq += p.rot;
    if (seq >= num_tiles) seq -= num_tiles;
    return p.rev ? num_tiles - 1 - seq : seq;
}

// ---- layer chaining (ChainArgs) ----
// The gate is executed by one thread of the CTA during the prologue, ahead of the CTA-wide barrier that precedes
// griddepcontrol.launch_dependents (issued by every thread: one thread's trigger does not release the successor).
__device__ __forceinline__ bool chain_enabled(const ConvArgs& p) { return p.ch.dep_flags != nullptr; }
__device__ __forceinline__ void chain_gate(const ConvArgs& p) {
    if (p.ch.gate_done != nullptr) flag_wait_ge(p.ch.gate_done, p.ch.gate_total, 0x900);
}
// Whole warp: wait until every input pixel read by the A tile of output rows [m0, m0 + 128) has been written.  Returns
// true once the whole producer layer is known to be complete (the caller then stops calling).
__device__ __forceinline__ bool chain_wait_a(const ConvArgs& p, int m0, int lane) {
    const ChainArgs& c = p.ch;
    uint32_t d = 0;
    if (lane == 0) d = ld_acquire_gpu(c.dep_done);
    d = __shfl_sync(0xffffffffu, d, 0);
    if (d >= c.dep_total) {
        fence_proxy_async_all();
        return true;
    }
    if (m0 >= p.M) return false;   // phantom tile of an odd pair: reads nothing that matters
    const int m1 = min(m0 + kBlockM, p.M) - 1;
    int lo = m0, hi = m1;
    if (p.a_im2col) {
        const int hw = p.Ho * p.Wo;
        const int n0 = m0 / hw, y0 = (m0 - n0 * hw) / p.Wo;
        const int n1 = m1 / hw, y1 = (m1 - n1 * hw) / p.Wo;
        const int yi0 = max(0, y0 * p.stride + p.lower);
        const int yi1 = min(c.dep_h - 1, y1 * p.stride + p.lower + p.ksize - 1);
        lo = (n0 * c.dep_h + yi0) * c.dep_w;
        hi = (n1 * c.dep_h + yi1) * c.dep_w + c.dep_w - 1;
    }
    for (int t = (lo >> 7) + lane; t <= (hi >> 7); t += 32) flag_wait_ge(c.dep_flags + t, c.dep_need, 0x910);
    __syncwarp();
    fence_proxy_async_all();   // the acquires above (generic proxy) order the TMA loads that follow (async proxy)
    return false;
}
// One lane, after the global writes of output rows [row, row + 32) of one tile have completed and are visible to it.
// The layer-wide counter is posted once per warp, when it is done (chain_post_done): nobody needs it earlier.
// The increment itself is relaxed: the caller has already waited for the COMPLETION of the bulk stores it announces
// (cp.async.bulk.wait_group, not .read -- the data is in L2 before the reduction is even issued), or has fenced its
// ordinary stores (epilogue_tile_post).  A red.release here compiles to MEMBAR.ALL.GPU + RED and cost 0.3 ms per
// forward pass (measured, ChainArgs::mode bit 0 of the profiling build selects it for comparison).
__device__ __forceinline__ void chain_post(const ConvArgs& p, int row) {
#ifdef Y3_PROFILING
    if (p.ch.mode & 1u) {
        red_release_gpu_add(p.ch.post_flags + (row >> 7), 1u);
        return;
    }
#endif
    red_relaxed_gpu_add(p.ch.post_flags + (row >> 7), 1u);
}
__device__ __forceinline__ void chain_post_done(const ConvArgs& p, uint32_t tiles) {
    if (tiles) red_relaxed_gpu_add(p.ch.post_done, tiles);
}
// Whole warp, after epilogue_tile (ordinary st.global by every lane): make all lanes' stores visible, then post once.
__device__ __forceinline__ void epilogue_tile_post(const ConvArgs& p, int row, int lane) {
    if (p.ch.post_flags == nullptr) return;
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        chain_post(p, row);
        chain_post_done(p, 1u);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BLOCK_N accumulator tile, executed by one of the four epilogue warps (TMEM lane quarter q).
//   TMEM -> registers (thread = one pixel row, 32 fp32 columns per chunk) -> +bias -> LeakyReLU(0.1)
//   -> per-warp 32x33 fp32 transpose tile in shared memory (conflict free both ways)
//   -> re-read so that 4 lanes cover 32 consecutive channels of one pixel: residual loads and output stores are
//      full 32-byte sectors (8 pixels x 64 B per instruction) instead of 32 scattered 16-byte pieces
//   -> +residual (fp32) -> bf16 -> global   (optionally replicated to the 2x2 nearest-upsample positions)
// fp32 outputs (the 3*(5+C)-channel heads, pixel stride not a multiple of 4 floats) store 32 consecutive floats of
// one pixel per instruction.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kXposePitch = 36;                       // floats per transpose-tile row: 144 B keeps 16-byte alignment
constexpr int kXposeWarpFloats = 32 * kXposePitch;    // and makes both the 128-bit writes and reads conflict free

// Where does GEMM row m (iteration index) go?  Identity for dense->dense layers; otherwise decode (n, y, x) on the
// iteration grid, drop halo rows, and re-encode for the output / residual layouts.
struct RowMap {
    bool valid;
    bool last_x, last_y;    // pixel sits in the last column / row of its image (needed to zero the halo it borders)
    long long out;          // output pixel index (top-left pixel of the 2x2 block when upsampling)
    long long res;          // residual pixel index (dense layout)
};

__device__ __forceinline__ RowMap map_row(const ConvArgs& p, int m) {
    RowMap r;
    if (p.it_h == p.Ho && p.it_w == p.Wo && !p.out_padded && !p.upsample) {
        r.valid = m < p.M;
        r.last_x = r.last_y = false;
        r.out = r.res = m;
        return r;
    }
    const int per = p.it_h * p.it_w;
    const int n = m / per;
    const int rem = m - n * per;
    const int y = rem / p.it_w;
    const int x = rem - y * p.it_w;
    r.valid = (m < p.M) && (y < p.Ho) && (x < p.Wo);
    r.last_x = (x == p.Wo - 1);
    r.last_y = (y == p.Ho - 1);
    const long long dense = ((long long)n * p.Ho + y) * p.Wo + x;
    r.res = dense;
    if (p.upsample) r.out = ((long long)n * 2 * p.Ho + 2 * y) * (2LL * p.Wo) + 2 * x;
    else if (p.out_padded) r.out = ((long long)n * (p.Ho + 1) + y) * (p.Wo + 1) + x;
    else r.out = dense;
    return r;
}

// Epilogue of one 128 x block_n accumulator tile whose first GEMM row is m_base, for TMEM lane quarter q.
__device__ __forceinline__ void epilogue_tile(const ConvArgs& p, int block_n, int m_base, int tn, uint32_t t_row, int q,
                                              int lane, float* xp) {
    const int n_base = tn * block_n;
    const int m_w = m_base + q * 32;           // first GEMM row of this warp's slab
    const int sub = lane >> 2;                 // 0..7 : pixel row inside an 8-row group
    const int seg = lane & 3;                  // 0..3 : 8-channel (16 B) segment inside the 32-channel chunk
    const bool has_res = p.residual != nullptr;
    const int nchunks = min(block_n / 32, (p.cout - n_base + 31) / 32);
    const float slope = p.leaky ? 0.1f : 1.0f;   // LeakyReLU(0.1)(x) = max(x, 0.1x); slope 1 makes it the identity
    float4* xrow = reinterpret_cast<float4*>(xp + lane * kXposePitch);

    // the four rows this lane stores (transposed mapping): where they go, where their residual comes from
    RowMap rm[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) rm[it] = map_row(p, m_w + it * 8 + sub);

    // residual tile rows of this lane (4 rows x 16 B per chunk), prefetched one chunk ahead
    uint4 rcur[4], rnext[4];
    auto load_res = [&](int c, uint4 (&r4)[4]) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            r4[it] = make_uint4(0u, 0u, 0u, 0u);
            if (has_res && rm[it].valid)
                r4[it] = __ldg(reinterpret_cast<const uint4*>(p.residual + rm[it].res * p.res_stride + n_base + c * 32 +
                                                              seg * 8));
        }
    };
    if (has_res && nchunks > 0) load_res(0, rnext);

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int ncol = n_base + c * 32;   // first output channel of this chunk
        uint32_t v[32];
        tmem_ld_32x32(t_row + (uint32_t)(c * 32), v);
#pragma unroll
        for (int it = 0; it < 4; ++it) rcur[it] = rnext[it];
        if (has_res && c + 1 < nchunks) load_res(c + 1, rnext);
        float4 bias4[8];   // issued before the TMEM wait so their latency is hidden behind it
        {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + ncol);
#pragma unroll
            for (int j = 0; j < 8; ++j) bias4[j] = __ldg(bp + j);
        }
        tmem_ld_wait();
        if (Y3_DBG_BITS(p) & 1) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 f;
            f.x = __uint_as_float(v[4 * j + 0]) + bias4[j].x;
            f.y = __uint_as_float(v[4 * j + 1]) + bias4[j].y;
            f.z = __uint_as_float(v[4 * j + 2]) + bias4[j].z;
            f.w = __uint_as_float(v[4 * j + 3]) + bias4[j].w;
            f.x = fmaxf(f.x, slope * f.x);
            f.y = fmaxf(f.y, slope * f.y);
            f.z = fmaxf(f.z, slope * f.z);
            f.w = fmaxf(f.w, slope * f.w);
            xrow[j] = f;
        }
        __syncwarp();
        if (!p.out_fp32) {
            __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + sub;
                if (rm[it].valid) {
                    const float4* src = reinterpret_cast<const float4*>(xp + r * kXposePitch + seg * 8);
                    const float4 lo = src[0], hi = src[1];
                    float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    if (has_res) {
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rcur[it]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 r2 = __bfloat1622float2(h2[e]);
                            f[2 * e] += r2.x;
                            f[2 * e + 1] += r2.y;
                        }
                    }
                    __nv_bfloat162 o2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                    const uint4 o = *reinterpret_cast<uint4*>(o2);
                    __nv_bfloat16* dst = ob + rm[it].out * p.out_stride + ncol + seg * 8;
                    if (Y3_DBG_BITS(p) & 2) {
                        if (o.x == 0x12345678u && o.y == 0x9abcdef0u) ob[0] = __float2bfloat16(0.f);   // keep the math alive
                    } else if (p.upsample) {
                        const long long up_row = 2LL * p.Wo * p.out_stride;
                        *reinterpret_cast<uint4*>(dst) = o;
                        *reinterpret_cast<uint4*>(dst + p.out_stride) = o;
                        *reinterpret_cast<uint4*>(dst + up_row) = o;
                        *reinterpret_cast<uint4*>(dst + up_row + p.out_stride) = o;
                    } else {
                        *reinterpret_cast<uint4*>(dst) = o;
                        if (p.out_padded && (rm[it].last_x || rm[it].last_y)) {
                            // keep the zero halo of the [B, Ho+1, Wo+1, C] layout intact (arena buffers are recycled)
                            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                            const long long row = (long long)(p.Wo + 1) * p.out_stride;
                            if (rm[it].last_x) *reinterpret_cast<uint4*>(dst + p.out_stride) = z;
                            if (rm[it].last_y) *reinterpret_cast<uint4*>(dst + row) = z;
                            if (rm[it].last_x && rm[it].last_y) *reinterpret_cast<uint4*>(dst + row + p.out_stride) = z;
                        }
                    }
                }
            }
        } else {
            // fp32 heads: dense -> dense only (host-enforced), 32 consecutive floats of one pixel per instruction
            float* of = reinterpret_cast<float*>(p.out);
            const bool col_ok = (ncol + lane) < p.cout;
            const int rows = min(32, p.M - m_w);
            float* dst = of + (long long)m_w * p.out_stride + ncol + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                if (r < rows && col_ok) dst[(long long)r * p.out_stride] = xp[r * kXposePitch + lane];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-store epilogue for the common case (bf16 output, dense pixel indexing, no upsample).  One epilogue warp owns the
// 32 TMEM lanes (= 32 output pixels) q*32.. of every tile of its group and walks them in chunks of CW columns through a
// ring of NBUF staging buffers (32 rows x CW bf16, written with the tensor map's swizzle):
//   [residual]  one lane TMA-loads the 32 x CW residual block of a chunk NBUF-1 chunks ahead of the one being
//               processed (also across tile boundaries, and before the accumulator is even complete)
//   TMEM -> registers (thread = pixel row) -> accumulator handed back to the MMA warp right after the tile's last load
//   -> +bias -> LeakyReLU(0.1) [-> + residual read back from the staging buffer] -> bf16
//   -> 16-byte st.shared into the staging buffer (bank-conflict free) -> fence.proxy.async
//   -> one lane issues a 2-D TMA store of the 32 x CW block; the buffer is reused NBUF chunks later, so the store's
//      read of shared memory is never waited for on the spot.
// No per-row address arithmetic, no transposing re-read, no scattered global stores: ~1/3 of the instructions of
// epilogue_tile.  Rows past M and columns past Cout are clipped by the tensor map.  What motivated it: with loads and
// MMAs disabled the 3x3 128->256 @52 layer still spent 70 of its 96 us in the old epilogue; the first (single-buffer)
// version of this one took 1.3 us per chunk because every chunk waited for the previous store to drain its buffer.
//   CW = 64 -> SWIZZLE_128B rows of 128 B, NBUF = 2;   CW = 32 -> SWIZZLE_64B rows of 64 B, NBUF = 4
//   (residual layers use CW = 32 so the residual prefetch runs 3 chunks ahead).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kEpiWarpBytes = 8192;    // staging ring per epilogue warp; a multiple of the 1024 B swizzle period
constexpr int kEpiMaxBufs = 4;         // residual mbarriers per epilogue warp

// the tile sequence of one epilogue group of one CTA: tile ids first, first + step, ... < end;
// tile -> (tmg = tile / tiles_n, tn = tile % tiles_n), M tile = tmg * cl + cta_rank
struct EpiTiles {
    int first, step, end;
    int tiles_n, cl, cta_rank;
};

template <int CW>
struct EpiCursor {     // walks (tile, chunk) in processing order
    int tile, c, nch, ncol, row;
    bool valid;
    __device__ __forceinline__ void load(const ConvArgs& p, const EpiTiles& et, int block_n, int q) {
        valid = tile < et.end;
        if (!valid) return;
        const int tid_ = tile_id(p, tile, et.end);
        const int tmg = tid_ / et.tiles_n, tn = tid_ - tmg * et.tiles_n;
        const int tm = tmg * et.cl + et.cta_rank;
        const int n_base = tn * block_n;
        nch = min(block_n / CW, (p.cout - n_base + CW - 1) / CW);
        ncol = n_base;
        row = tm * kBlockM + q * 32;
        c = 0;
    }
    __device__ __forceinline__ void next(const ConvArgs& p, const EpiTiles& et, int block_n, int q) {
        if (++c < nch) {
            ncol += CW;
        } else {
            tile += et.step;
            load(p, et, block_n, q);
        }
    }
};

// F32 = true: fp32 output (the head convs when the caller provides a 16-byte aligned pixel pitch): CW = 32 floats per
// 128-byte staging row, no residual.
template <int CW, int NBUF, bool F32 = false>
__device__ __forceinline__ void epilogue_role_tma(const ConvArgs& p, const CUtensorMap* tmO, const CUtensorMap* tmR,
                                                  int block_n, const EpiTiles et, uint32_t t_acc, int q, int lane,
                                                  uint32_t stg, uint32_t res_bar0, uint32_t tfull_bar,
                                                  uint32_t tempty_addr, bool tempty_remote, unsigned long long* ts,
                                                  int ts_slot) {
    static_assert((CW == 64 || CW == 32) && NBUF >= 2 && NBUF <= kEpiMaxBufs &&
                  NBUF * 32 * CW * (F32 ? 4 : 2) <= kEpiWarpBytes && (!F32 || CW == 32), "ring");
    constexpr uint32_t ROW_BYTES = CW * (F32 ? 4 : 2);
    constexpr uint32_t BUF_BYTES = 32 * ROW_BYTES;
    const bool drain_only = (Y3_DBG_BITS(p) & 1) != 0;
    const bool has_res = !F32 && (p.residual != nullptr) && !drain_only;
    const float slope = p.leaky ? 0.1f : 1.0f;   // LeakyReLU(0.1)(x) = max(x, 0.1x); slope 1 makes it the identity
    const uint32_t sw = (ROW_BYTES == 128) ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
    const uint32_t row_off = (uint32_t)lane * ROW_BYTES;

    EpiCursor<CW> pr, pf;    // chunk being processed / chunk whose residual is fetched next
    pr.tile = et.first;
    pr.load(p, et, block_n, q);
    pf = pr;
    uint32_t g = 0, gp = 0;  // their running chunk numbers; chunk n lives in ring slot n % NBUF
    bool res_all = p.ch.res_flags == nullptr;   // chained layer: the residual's producer may still be running
    // posting (chained layers): a cursor that trails `pr` by kPostLag bulk-store groups.  A tile is posted once that
    // many newer groups have been committed, so the wait for its stores' completion never actually stalls.
    constexpr int kPostLag = 4;
    const bool posting = p.ch.post_flags != nullptr;
    EpiCursor<CW> pp = pr;
    uint32_t pp_gend = pp.valid ? (uint32_t)pp.nch : 0u;   // groups committed once pp's tile has been issued completely
    uint32_t posted = 0;
    auto issue_res = [&]() {   // lane 0 only
        if (!res_all && pf.c == 0 && pf.row < p.M) {
            // first chunk of a tile: its 32 residual rows must have been written (later chunks are the same rows)
            if (ld_acquire_gpu(p.ch.res_done) >= p.ch.res_total) res_all = true;
            else flag_wait_ge(p.ch.res_flags + (pf.row >> 7), p.ch.res_need, 0x920);
            fence_proxy_async_all();
        }
        const uint32_t b = gp % NBUF;
        mbar_arrive_expect_tx(res_bar0 + 8u * b, BUF_BYTES);
        tma_load_2d(stg + b * BUF_BYTES, tmR, res_bar0 + 8u * b, pf.ncol, pf.row);
    };
    if (has_res) {
        // all ring slots are free: fetch the residual of the first NBUF-1 chunks before the first accumulator is ready
#pragma unroll 1
        for (int i = 0; i < NBUF - 1 && pf.valid; ++i) {
            if (lane == 0) issue_res();
            pf.next(p, et, block_n, q);
            ++gp;
        }
    }

    uint32_t jj = 0;         // tiles of this group seen so far (accumulator phase)
#pragma unroll 1
    while (pr.valid) {
        if (pr.c == 0) {
            mbar_wait(tfull_bar, jj & 1u, 0x400);
            tc_fence_after();
            if (jj == 0 && lane == 0 && q == 0 && ts_slot == 7) ts_mark(ts, 6);   // first accumulator complete
        }
        const uint32_t b = g % NBUF;
        const uint32_t buf = stg + b * BUF_BYTES;
        if (ts != nullptr && ts_slot == 7 && q == 0 && lane == 0 && g == 1) ts_clock(ts, 16);   // chunk 1 starts
        if (!has_res) {
            // slot b was last used by chunk g - NBUF: its store must have finished reading shared memory
            if (lane == 0) tma_store_wait_read<NBUF - 1>();
            __syncwarp();
        }
        uint32_t v[CW];
        tmem_ld_32x32(t_acc + (uint32_t)(pr.c * CW), *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        if constexpr (CW == 64) tmem_ld_32x32(t_acc + (uint32_t)(pr.c * CW + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[CW - 32]));
        // bias of the first 32 columns: issued before the TMEM wait so the (L1-resident after the first tile) loads
        // overlap it; the shared-memory accesses below carry no memory clobber, so the second half's loads hoist too
        const float4* bp = reinterpret_cast<const float4*>(p.bias + pr.ncol);
        float4 bz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bz[j] = __ldg(bp + j);
        tmem_ld_wait();
        const bool stamp = ts != nullptr && ts_slot == 7 && q == 0 && lane == 0 && g >= 1 && g <= 2;   // chunks 1, 2 of warp 4
        if (stamp) ts_clock(ts, g == 1 ? 17 : 23);   // out of TMEM
        if (pr.c == pr.nch - 1) {
            // every TMEM read of this accumulator has completed -> hand it back to the MMA warp now
            tc_fence_before();
            if (tempty_remote) mbar_arrive_cluster(tempty_addr);
            else mbar_arrive(tempty_addr);
            ++jj;
        }
        if (!drain_only) {
            if (has_res) mbar_wait(res_bar0 + 8u * b, (g / NBUF) & 1u, 0x600 + b);
            if (stamp && g == 1) ts_clock(ts, 18);   // residual landed
#pragma unroll
            for (int j = 0; j < CW / 8; ++j) {      // one 16-byte (8-channel) piece of this thread's row at a time
                const float4 b0 = (j < 4) ? bz[2 * j] : __ldg(bp + 2 * j), b1 = (j < 4) ? bz[2 * j + 1] : __ldg(bp + 2 * j + 1);
                // packed fp32x2 add / multiply (sm_100 FADD2 / FMUL2, IEEE round-to-nearest like the scalar forms):
                // (x + bias), slope * (x + bias), then max -- 5 instructions per value pair instead of 6
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
                if constexpr (F32) {
                    // 8 floats = two 16-byte pieces of the 128-byte row
                    const uint32_t a0 = buf + row_off + ((((uint32_t)(2 * j)) ^ sw) << 4);
                    const uint32_t a1 = buf + row_off + ((((uint32_t)(2 * j + 1)) ^ sw) << 4);
                    st_shared_v4_relaxed(a0, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
                    st_shared_v4_relaxed(a1, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
                    continue;
                }
                const uint32_t addr = buf + row_off + ((((uint32_t)j) ^ sw) << 4);
                if (has_res) {
                    const uint4 r = ld_shared_v4_relaxed(addr);   // ordered after the residual barrier wait (both volatile)
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
                st_shared_v4_relaxed(addr, *reinterpret_cast<uint4*>(o2));   // same address as the load it depends on
            }
            if (stamp && g == 1) ts_clock(ts, 19);   // math + st.shared issued
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA (async proxy); also a compiler barrier
            if (stamp && g == 1) ts_clock(ts, 20);   // fenced
            __syncwarp();
            if (lane == 0) {
                if (!(Y3_DBG_BITS(p) & 2)) tma_store_2d(tmO, buf, pr.ncol, pr.row);
                tma_store_commit();
                if (stamp && g == 1) ts_clock(ts, 21);   // store issued
                if (posting && pp.valid && g + 1 >= pp_gend + kPostLag) {
                    tma_store_wait<kPostLag>();   // everything but the kPostLag newest groups has completed
                    chain_post(p, pp.row);
                    ++posted;
                }
                if (has_res && pf.valid) {
                    // all stores but the one just issued have read their buffers: the slot of chunk g - 1 is free,
                    // and it is the slot of chunk gp = g + NBUF - 1
                    tma_store_wait_read<1>();
                    issue_res();
                }
                if (stamp && g == 1) ts_clock(ts, 22);   // residual prefetch issued
            }
            if (has_res && pf.valid) {
                pf.next(p, et, block_n, q);
                ++gp;
            }
            if (posting && pp.valid && g + 1 >= pp_gend + kPostLag) {   // lane 0 has just posted pp's tile
                pp.tile += et.step;
                pp.load(p, et, block_n, q);
                if (pp.valid) pp_gend += (uint32_t)pp.nch;
            }
        }
        pr.next(p, et, block_n, q);
        ++g;
    }
    if (lane == 0) {
        if (q == 0) ts_mark(ts, ts_slot);         // last chunk handed to the TMA
        tma_store_wait<0>();                      // every bulk store has completed (not just been read) before exit
        if (posting) {
            while (pp.valid) {
                chain_post(p, pp.row);
                ++posted;
                pp.tile += et.step;
                pp.load(p, et, block_n, q);
            }
            chain_post_done(p, posted);
        }
        if (q == 0) ts_mark(ts, ts_slot + 2);
    }
}

template <int BLOCK_N, int SWZ, int STAGES>
struct ConvSmem {
    static constexpr int A_BYTES = kBlockM * SWZ;
    static constexpr int B_BYTES = BLOCK_N * SWZ;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    // epilogue region: per-warp staging ring of the TMA-store epilogue; the register-transpose epilogue (fp32 heads,
    // fused upsample) uses the first 32 x 36 fp32 of each warp's share instead
    static constexpr int XPOSE_BYTES = kConvEpiGroups * 4 * kEpiWarpBytes;
    static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16 + 8 * kEpiMaxBufs * 4 * kConvEpiGroups; // full/empty + tmem full/empty + weights + tmem ptr + residual ring
    static constexpr int TOTAL = 1024 /*align slack*/ + TILE_BYTES + XPOSE_BYTES + BAR_BYTES;
    // weights-resident variant (BRES, see conv_tc2.cuh): `stages` A tiles + all K blocks of the weight tile
    static constexpr int total_resident(int stages, int num_k_blocks) {
        return 1024 + stages * A_BYTES + num_k_blocks * B_BYTES + XPOSE_BYTES + BAR_BYTES;
    }
    static_assert(kEpiWarpBytes >= kXposeWarpFloats * 4 && TILE_BYTES % 1024 == 0, "staging rings share the transpose region");
};

// CLUSTER == 2: two CTAs of a cluster work on two adjacent M tiles of the same N tile.  Each loads its own A tile and
// HALF of the weight tile, multicast into both CTAs' shared memory, so every SM issues 128 + BLOCK_N/2 TMA rows per
// K block instead of 128 + BLOCK_N (the per-SM TMA issue rate is what bounds the 3x3 layers).  A stage may be refilled
// only after BOTH CTAs' MMAs have read it, so tcgen05.commit arrives on the empty barrier of both CTAs (count 2).
// BRES (CLUSTER == 1 only): the whole [BLOCK_N x K] weight tile is loaded once per launch, before
// griddepcontrol.wait, and the stage ring carries only the A operand.  For the Cin = 32 layers (64-byte rows, 9 taps)
// the TMA unit's row rate is the limit and the weight rows were a third of the rows fetched per output tile.
template <int BLOCK_N, int SWZ, int STAGES, int CLUSTER, bool BRES = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvArgs p) {
    using S = ConvSmem<BLOCK_N, SWZ, STAGES>;
    static_assert(!BRES || CLUSTER == 1, "weights-resident variant is single-CTA");
    constexpr int BLOCK_K = SWZ / 2;   // bf16 elements per swizzle row
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N");

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
    const uint32_t bfull_bar = bar_base + 8u * (2 * STAGES + 4);    // resident weights landed
    const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 5);
    auto res_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 5) + 16u + 8u * kEpiMaxBufs * w; };   // ring of epilogue warp w
    volatile uint32_t* tmem_ptr_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + tile_bytes + S::XPOSE_BYTES + 8 * (2 * STAGES + 5));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // work items: (group of CLUSTER adjacent M tiles) x (N tile); every CTA of a cluster walks the same sequence
    const int cta_rank = (CLUSTER > 1) ? (int)cluster_ctarank() : 0;
    const int num_tiles = ((p.tiles_m + CLUSTER - 1) / CLUSTER) * p.tiles_n;
    const int first_tile = (int)blockIdx.x / CLUSTER;
    const int tile_step = (int)gridDim.x / CLUSTER;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CLUSTER);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        for (int w = 0; w < kEpiMaxBufs * 4 * kConvEpiGroups; ++w) mbar_init(res_bar(0) + 8u * w, 1);
        mbar_init(bfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
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
    if (CLUSTER > 1) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    if constexpr (BRES) {
        // the weights do not depend on the previous layer: fetch the tile before waiting for it
        if (warp == 0 && elect_one()) {
            const int n0 = (tile_id(p, first_tile, num_tiles) % p.tiles_n) * BLOCK_N;   // constant for this CTA (host-enforced)
            mbar_arrive_expect_tx(bfull_bar, (uint32_t)(p.num_k_blocks * S::B_BYTES));
            for (int kb = 0; kb < p.num_k_blocks; ++kb)
                tma_load_2d(smem_b + kb * S::B_BYTES, &tmB, bfull_bar, kb * BLOCK_K, n0);
        }
    }
    // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous layer's tail;
    // from here on we touch activations it wrote (and buffers it may still be reading), so wait for it to finish.
    // chained layers (ChainArgs) wait tile by tile in the producer / epilogue warps instead
    const bool chained = chain_enabled(p);
    pdl_launch_dependents();   // chained: the gate (chain_gate) was passed before the barrier above
    if (!chained) pdl_wait();

    // The two single-thread roles run their loops with the WHOLE warp (uniform control flow, so the compiler keeps
    // addresses / descriptors / barrier phases in uniform registers) and predicate only the issuing instructions on
    // one elected lane.  An earlier version wrapped the loops in `if (lane == 0)`: ncu showed ~190 SASS instructions
    // per K block in each role (runtime divisions + ELECT/BRA.U.ANY uniformisation loops), i.e. the kernel was bound by
    // scalar issue latency, not by the tensor pipe.
    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        bool dep_all = false;   // chained: the whole input has been seen complete
        const int hw = p.Ho * p.Wo;
        auto load_b = [&](int st, int kcoord, int n0) {
            if (CLUSTER == 1) {
                tma_load_2d(smem_b + st * S::B_BYTES, &tmB, full_bar(st), kcoord, n0);
            } else {
                constexpr int HALF = BLOCK_N / CLUSTER;
                tma_load_2d_mc(smem_b + st * S::B_BYTES + cta_rank * (HALF * SWZ), &tmB, full_bar(st), kcoord,
                               n0 + cta_rank * HALF, (uint16_t)((1u << CLUSTER) - 1u));
            }
        };
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int tid_ = tile_id(p, tile, num_tiles);
            const int tmg = tid_ / p.tiles_n, tn = tid_ - tmg * p.tiles_n;
            const int tm = tmg * CLUSTER + cta_rank;
            const int m0 = tm * kBlockM;
            const int n0 = tn * BLOCK_N;
            int kcoord = 0;   // K coordinate of the weight tile
            if (chained && !dep_all) dep_all = chain_wait_a(p, m0, lane);
            if (p.a_im2col) {
                const int cn = m0 / hw;
                const int rem = m0 - cn * hw;
                const int po = rem / p.Wo;
                const int qo = rem - po * p.Wo;
                const int cw = qo * p.stride_w + p.lower_w;
                const int ch = po * p.stride + p.lower;
                const int off_w0 = tn * p.a_shift_n;
                for (int r = 0; r < p.ksize; ++r) {
                    for (int sx = 0; sx < p.ksize_w; ++sx) {
                        for (int c0 = 0; c0 < p.kblocks_per_tap * BLOCK_K; c0 += BLOCK_K) {
                            mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                            if (leader) {
                                mbar_arrive_expect_tx(full_bar(stage), BRES ? S::A_BYTES : S::STAGE_BYTES);
                                tma_load_im2col_4d(smem_a + stage * S::A_BYTES, &tmA, full_bar(stage), c0, cw, ch, cn,
                                                   (uint16_t)(sx + off_w0), (uint16_t)r);
                                if constexpr (!BRES) load_b(stage, kcoord, n0);
                            }
                            kcoord += BLOCK_K;
                            if (++stage == nst) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            } else {
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
                    if (leader) {
                        mbar_arrive_expect_tx(full_bar(stage), BRES ? S::A_BYTES : S::STAGE_BYTES);
                        tma_load_2d(smem_a + stage * S::A_BYTES, &tmA, full_bar(stage), kcoord, m0);
                        if constexpr (!BRES) load_b(stage, kcoord, n0);
                    }
                    kcoord += BLOCK_K;
                    if (++stage == nst) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N);
        const uint64_t adesc0 = make_smem_desc<SWZ>(smem_a);
        const uint64_t bdesc0 = make_smem_desc<SWZ>(smem_b);
        int stage = 0;
        uint32_t phase = 0;
        int j = 0;
        if constexpr (BRES) {
            mbar_wait(bfull_bar, 0, 0x700);   // the weight tile is resident
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
                if (leader) {
                    // stage s sits s*A_BYTES (s*B_BYTES) further: +bytes>>4 in the descriptor's address field
                    const uint64_t adesc = adesc0 + (uint64_t)(stage * (S::A_BYTES >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)((BRES ? kb : stage) * (S::B_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((kb | k) != 0));
                    }
                    // smem slot reusable once these MMAs have read it (in a cluster: tell both producers)
                    if (CLUSTER == 1) umma_commit(empty_bar(stage));
                    else umma_commit_mc(empty_bar(stage), (uint16_t)((1u << CLUSTER) - 1u));
                    if (kb == p.num_k_blocks - 1) umma_commit(tfull_bar(acc));
                }
                if (++stage == nst) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue groups =====================
        const int eg = (warp - 4) >> 2;         // group: owns accumulator stage eg
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        float* xp = reinterpret_cast<float*>(smem_gen + tile_bytes + (warp - 4) * kEpiWarpBytes);
        if (p.tma_out) {
            const uint32_t stg = smem_base + tile_bytes + (uint32_t)((warp - 4) * kEpiWarpBytes);
            const EpiTiles et{first_tile + eg * tile_step, kConvEpiGroups * tile_step, num_tiles, p.tiles_n, CLUSTER, cta_rank};
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * BLOCK_N);
            if (BLOCK_N >= 64 && p.tma_out == 64)
                epilogue_role_tma<64, 2>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty_bar(eg), false, nullptr, 7 + eg);
            else if (p.out_fp32)
                epilogue_role_tma<32, 2, true>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                               tempty_bar(eg), false, nullptr, 7 + eg);
            else
                epilogue_role_tma<32, 4>(p, &tmO, &tmR, BLOCK_N, et, t_acc, q, lane, stg, res_bar(warp - 4), tfull_bar(eg),
                                         tempty_bar(eg), false, nullptr, 7 + eg);
        } else {
            int j = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++j) {
                if ((j % kConvEpiGroups) != eg) continue;
                const int acc = j & 1;
                const int tid_ = tile_id(p, tile, num_tiles);
                const int tmg = tid_ / p.tiles_n, tn = tid_ - tmg * p.tiles_n;
                const int tm = tmg * CLUSTER + cta_rank;
                mbar_wait(tfull_bar(acc), (uint32_t)((j >> 1) & 1), 0x400 + acc);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
                epilogue_tile(p, BLOCK_N, tm * kBlockM, tn, t_row, q, lane, xp);
                // all TMEM reads of this accumulator are complete (wait::ld) -> hand it back to the MMA warp
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
                epilogue_tile_post(p, tm * kBlockM + q * 32, lane);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it / arrive on it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// Debug helper: does a UMMA smem descriptor accept a start address that is a whole number of swizzle rows (but not a
// multiple of the 8-row swizzle period) into a TMA-written SWIZZLE_<SWZ>B tile?  Loads X[rows][SWZ/2] and W[64][SWZ/2],
// issues D[128,64] = X[shift .. shift+128) * W^T with the A descriptor advanced by `shift` rows, optionally with the
// descriptor's base-offset field set to the row phase, and writes D (fp32) out.
template <int SWZ>
__global__ void umma_shift_test_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                       int rows, int shift, int base_off_mode, float* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int BLOCK_K = SWZ / 2;
    const uint32_t smem_x = smem_base;
    const uint32_t smem_w = smem_base + (uint32_t)((rows + 127) / 128) * 128u * SWZ;
    const uint32_t bar = smem_w + 64 * SWZ;
    const uint32_t bar2 = bar + 8;
    const uint32_t tptr = bar + 16;
    uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(gen + (tptr - smem_base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tptr, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tptr_gen;
    if (threadIdx.x == 0) {
        // X arrives as whole 128-row boxes (rows past the tensor are zero filled but still counted), then W
        const int nbox = (rows + 127) / 128;
        mbar_arrive_expect_tx(bar, (uint32_t)nbox * 128u * SWZ + 64u * SWZ);
        for (int b = 0; b < nbox; ++b) tma_load_2d(smem_x + b * 128 * SWZ, &tmX, bar, 0, b * 128);
        tma_load_2d(smem_w, &tmW, bar, 0, 0);
    }
    mbar_wait(bar, 0, 0x800);
    tc_fence_after();
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64);
        const uint32_t a_addr = smem_x + (uint32_t)shift * SWZ;
        uint64_t adesc = make_smem_desc<SWZ>(a_addr);
        if (base_off_mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t bdesc = make_smem_desc<SWZ>(smem_w);
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0, 0x801);
    tc_fence_after();
    if (warp < 4) {
        uint32_t v[32];
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// Debug helper: fetch ONE A tile (128 pixels x BLOCK_K channels) through the same TMA path as the conv kernel and copy the
// raw (still swizzled) shared-memory image to global memory, so tests can check the im2col/padding semantics in
// isolation from the MMA.
template <int SWZ>
__global__ void tma_tile_dump_kernel(const __grid_constant__ CUtensorMap tmA, int a_im2col, int c0, int cw, int ch,
                                     int cn, int off_w, int off_h, int m0, uint8_t* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr int BYTES = kBlockM * SWZ;
    const uint32_t bar = smem_base + BYTES;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, BYTES);
        if (a_im2col)
            tma_load_im2col_4d(smem_base, &tmA, bar, c0, cw, ch, cn, (uint16_t)off_w, (uint16_t)off_h);
        else
            tma_load_2d(smem_base, &tmA, bar, c0, m0);
    }
    mbar_wait(bar, 0, 0x500);
    for (int i = threadIdx.x; i < BYTES; i += blockDim.x) out[i] = smem_gen[i];
}

}  // namespace y3
