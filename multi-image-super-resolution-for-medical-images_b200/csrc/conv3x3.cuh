// Persistent implicit-GEMM kernel for Conv2d 3x3 (padding 1) forward and data-gradient on tcgen05/TMEM + TMA.
// Reference ops: nn.Conv2d(k=3, p=1) inside UNetBlock, /root/reference/src/unet_model.py:27,30 (and its dgrad).
//
//   D[pixel, n] = sum_{dh,dw,c} A[pixel + (dh-1, dw-1), c] * Wp[n, (dh*3+dw)*C + c]
//
// Design (this kernel replaced a first-generation one-tile-per-CTA kernel, retired in round 2):
//   * any H, W: edge tiles of a shape that is not a multiple of the 16 x 8 pixel tile reach past the image — TMA zero-fills
//     those loads (the same zeros as the padding), clips the stores, and the epilogue keeps the out-of-image pixels out of
//     the BatchNorm statistics (Conv3Args::ragged).
//   * persistent CTAs (one per SM) walk a static tile schedule; the accumulator is double-buffered in TMEM so the
//     epilogue of tile i overlaps the main loop of tile i+1; barrier set-up / TMEM allocation happen once per CTA.
//   * an output tile is 16 rows x 8 columns of pixels. For a fixed horizontal tap dw, ONE haloed TMA box
//     {64 ch, 8 w, 18 h} serves the three vertical taps: in the K-major 128B-swizzled layout one pixel row of the
//     tile is exactly one 1024-byte swizzle atom, so the operand for tap dh starts dh*1024 bytes into the box.
//     Activation traffic from L2 drops from 9 to 3.4 tile loads per 64-channel chunk.
//   * activations (A ring, 18 KB slots) and weights (B ring, <=16 KB slots) have independent rings; with
//     BLOCK_N = 256 the two 128-column halves of the weight tile reuse the same activation box.
//   * BatchNorm statistics are accumulated in registers across all tiles of the CTA (the schedule keeps the CTA on
//     one column block) and leave through one vector of atomics per CTA.
#pragma once
#include "ptx.cuh"

namespace b200sr {

struct Conv3Args {
    int H, W;
    int tiles_w;       // ceil(W / 8)
    int tiles_hw;      // ceil(H / 16) * ceil(W / 8)
    int ragged;        // H % 16 != 0 or W % 8 != 0: edge tiles reach past the image. TMA zero-fills their loads and clips
                       // their stores; the epilogue only has to keep the out-of-image pixels out of the statistics (and
                       // out of the ReLU-mask reads)
    int n_tiles;       // N / BLOCK_N
    int num_tiles;     // B * tiles_hw * n_tiles
    int cin_chunks;    // C / 64
    int C;             // channels of A (GEMM-K per tap)
    int n_total;       // GEMM-N
    int relu;
    int out_pix_stride;
    int out_c_off;
    int stats_replicas;
    int cout_t;        // MODE 1: channels per (i,j) sub-pixel (n_total = 4 * cout_t); modulus of the affine vectors
    int stats_sum_cols;  // > 0: only the column SUMS of the first stats_sum_cols output columns are needed (ConvTranspose2d
                         // bias gradient from a dgrad): no squares, no work for the other columns
    int stats_slots;   // > 0: deterministic statistics — CTA (blockIdx.x / n_tiles) STORES its partial sums into its own
                       // slot of stats[stats_replicas][2][n_total] (no atomics, no pre-zeroing; unused slots are zeroed
                       // here); 0: legacy mode, partial sums are atomically ADDED into slot blockIdx.x % stats_replicas
    int b_resident;    // the CTA's whole weight block fits the B ring: load it once, keep it for all tiles
    int sa;            // activation ring depth (3 .. C3_SA_MAX): ring bytes the resident weights do not need go to A
                       // stages — at 256^2 the kernel is bound by TMA latency x bytes in flight, not by the MMAs
    __nv_bfloat16* out;
    const float* col_scale;  // nullable, [n_total]
    const float* col_shift;  // nullable, [n_total]
    float* stats;            // nullable, [replicas][2][n_total]
    // nullable ReLU mask of a data-gradient: out = acc * [mask > 0], mask a (B,H,W,n_total) channel slot (the activation
    // of the layer the gradient flows into; Conv+ReLU stacks without BatchNorm: Fast-DDPM DoubleConv, VGG features)
    const __nv_bfloat16* mask;
    int mask_pix_stride;
    int mask_c_off;
    // SPLIT epilogue (fp32-accuracy eval mode, see below): channel distance between the [hi | lo | hi] parts of the output
    int split_stride;
    // Fused train-mode BatchNorm finalize (slot-mode statistics only): the last CTA of a column block to finish (ticket
    // counter per column block) sums that block's statistic slots in slot order and writes scale / shift / mean / invstd
    // and the running statistics — b200sr_bn_finalize without its launch. bn_scale == nullptr: off.
    const float* bn_gamma;
    const float* bn_beta;
    const float* bn_conv_bias;  // nullable
    float* bn_scale;
    float* bn_shift;
    float* bn_mean;
    float* bn_invstd;
    float* bn_rmean;            // nullable (track_running_stats off)
    float* bn_rvar;
    long long* bn_nbt;          // nullable
    unsigned* bn_counters;      // [n_tiles], zero-initialised once, self-resetting
    float bn_count, bn_eps, bn_momentum;
};

// Fixed-order finalize of BLOCK_N channels from `used` statistic slots by all 256 threads of the CTA that drew the last
// ticket. Layout: BLOCK_N/2 float4 column tasks (sum | sum of squares) x 512/BLOCK_N slot lanes; double accumulation like
// bn_finalize_kernel. `scratch` >= 8 KB of shared memory.
template <int BLOCK_N>
__device__ __forceinline__ void c3_bn_finalize(const Conv3Args& args, int n0, int used, uint8_t* scratch) {
    constexpr int TASKS = BLOCK_N / 2;     // float4 columns: [0, BLOCK_N/4) = sums, [BLOCK_N/4, BLOCK_N/2) = squares
    constexpr int LANES = 256 / TASKS;     // 8 / 4 / 2
    const int tid = threadIdx.x;
    double* s_acc = reinterpret_cast<double*>(scratch);  // [LANES][2][BLOCK_N]
    if (tid < 256) {  // (the CTA may have 384 threads: two epilogue teams)
        const int task = tid % TASKS, lane = tid / TASKS;
        const int which = task / (BLOCK_N / 4), c4 = task % (BLOCK_N / 4);
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const float* base = args.stats + static_cast<size_t>(which) * args.n_total + n0 + 4 * c4;
#pragma unroll 8  // this CTA is the last of its column block: every round trip here is on the critical path
        for (int sl = lane; sl < used; sl += LANES) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(sl) * 2 * args.n_total));
            acc[0] += v.x;
            acc[1] += v.y;
            acc[2] += v.z;
            acc[3] += v.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_acc[(lane * 2 + which) * BLOCK_N + 4 * c4 + k] = acc[k];
    }
    __syncthreads();
    if (tid < BLOCK_N) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int l = 0; l < LANES; ++l) {
            s += s_acc[(l * 2 + 0) * BLOCK_N + tid];
            q += s_acc[(l * 2 + 1) * BLOCK_N + tid];
        }
        const int c = n0 + tid;
        const double count = static_cast<double>(args.bn_count);
        const double mean = s / count;
        double var = q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float invstd = 1.0f / sqrtf(static_cast<float>(var) + args.bn_eps);
        const float sc = args.bn_gamma[c] * invstd;
        args.bn_scale[c] = sc;
        args.bn_shift[c] = args.bn_beta[c] - static_cast<float>(mean) * sc;
        args.bn_mean[c] = static_cast<float>(mean);
        args.bn_invstd[c] = invstd;
        if (args.bn_rmean != nullptr) {
            const float b = args.bn_conv_bias ? args.bn_conv_bias[c] : 0.f;
            const float unbiased = static_cast<float>(var * (count / (count - 1.0)));
            args.bn_rmean[c] = (1.f - args.bn_momentum) * args.bn_rmean[c] + args.bn_momentum * (static_cast<float>(mean) + b);
            args.bn_rvar[c] = (1.f - args.bn_momentum) * args.bn_rvar[c] + args.bn_momentum * unbiased;
        }
        if (c == 0 && args.bn_nbt != nullptr) *args.bn_nbt += 1;
    }
}

#ifndef C3_HAS_MASK
#define C3_HAS_MASK 1  // A/B switch for the fused ReLU-mask epilogue (build with -DC3_HAS_MASK=0 to compare)
#endif
constexpr int C3_THREADS = 256;
constexpr int C3_SA_MAX = 8;
constexpr int C3_TILE_H = 16;
constexpr int C3_TILE_W = 8;
constexpr int C3_A_SLOT = (C3_TILE_H + 2) * C3_TILE_W * 128;  // 18432 B: 18 pixel rows x 8 pixels x 64 ch bf16
constexpr int C3_OUT_STAGE = C3_TILE_H * C3_TILE_W * 128;     // 16384 B: 128 pixels x 64 ch bf16

template <int BLOCK_N, bool SPLIT = false>
struct C3Cfg {
    static constexpr int BN_SLOT = BLOCK_N < 128 ? BLOCK_N : 128;  // columns per weight slot = UMMA N
    static constexpr int NH = BLOCK_N / BN_SLOT;
    static constexpr int B_SLOT = BN_SLOT * 128;
    static constexpr int SA = 3;
    // 18 x 8 KB holds every weight of a Cout=64 layer (K <= 1152); the SPLIT epilogue needs a second staging buffer instead
    static constexpr int SB = (BLOCK_N == 64 && !SPLIT) ? 18 : 8;
    static constexpr int NBUF = (BLOCK_N == 64 && !SPLIT) ? 1 : 2;  // output staging buffers (128 pixels x 64 ch)
    static constexpr int STAGING = NBUF * C3_OUT_STAGE;
    static constexpr int RING_BYTES = SA * C3_A_SLOT + SB * B_SLOT;
    static constexpr int SMEM_BYTES =
        RING_BYTES + STAGING + 1024 /* barriers */ + 2048 /* affine vectors */ + 1024 /* alignment slack */;
};

// MODE 0: Conv2d 3x3 forward / dgrad (three haloed boxes per channel chunk, three vertical taps per box)
// MODE 1: ConvTranspose2d k2 s2 forward as a 1-tap GEMM with N = 4*Cout and a pixel-shuffle scatter epilogue + bias
//         (unet_model.py:67-76; writes straight into the concat slot, which replaces torch.cat at :101-113)
// MODE 2: ConvTranspose2d k2 s2 dgrad: four taps (i,j), each gathered through the 5-D view (c, j, w, i, b*H+h)
// MODE 3: Conv2d 1x1 forward / dgrad (DeepCNN downsample branch, ModelLoader.py:347-351): one tap, plain NHWC store
//
// SPLIT (MODE 0 / 1 only): fp32-accuracy inference on the bf16 tensor cores (BASELINE configs[0] is an fp32 forward,
// tolerance 1e-4). An fp32 activation v is stored as THREE bf16 channel groups [hi | lo | hi], hi = bf16(v),
// lo = bf16(v - hi); the weights are packed as [w_hi | w_hi | w_lo] along K (b200sr_pack_jobs kinds 11 / 12), so the
// ordinary main loop over K' = 3*Cin accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in fp32 (the dropped lo*lo term is
// 2^-18 relative). Only the epilogue differs: the fp32 result (after the fp32 affine / ReLU) is split again and the
// three parts are stored at channel offsets 0, split_stride, 2*split_stride of the output slot.
// ConvTranspose modes (1, 2) have K = Cin only (one tap per box), so a tile's main loop is short and its epilogue — 64 KB
// of bf16 output per 128 x 256 tile — is the critical path. They run TWO epilogue teams of four warps (warps 4-7 and
// 8-11, each covering the 128 TMEM lanes): team t converts and stores the 64-column groups g with g % 2 == t through its
// own staging buffer and its own TMA-store issuer thread, so two groups are in flight per CTA.
// The conv3x3 forward / dgrad (MODE 0) with BLOCK_N >= 128 runs two teams as well: with the BatchNorm statistics (a
// 32 x 32 transposing shuffle reduction per 32 columns and statistic) the epilogue of one warp per scheduler is longer than
// the main loop of the small-K layers (K = 576: 36 MMAs per tile).
template <int BLOCK_N, int MODE_T, bool SPLIT>
__host__ __device__ constexpr int c3_teams() {
    return ((MODE_T == 0 || MODE_T == 1 || MODE_T == 2) && BLOCK_N >= 128 && !SPLIT) ? 2 : 1;
}
template <int BLOCK_N, int MODE_T, bool SPLIT>
__host__ __device__ constexpr int c3_threads() {
    return 128 + 128 * c3_teams<BLOCK_N, MODE_T, SPLIT>();
}

//
// PAIR (BLOCK_N = 64, MODE 0, resident weights): the kernel runs as clusters of two CTAs (one TPC) issuing
// tcgen05.mma.cta_group::2 with M = 256: CTA r of a pair owns pixel tile (blockIdx.x + k*gridDim.x) — the pair's two tiles
// are neighbours — stages its own haloed activation boxes and HALF (32) of the 64 weight rows; the leader (rank 0) issues
// every MMA, each CTA's accumulator rows land in its own TMEM and are drained by its own epilogue warps. An N = 64 MMA
// reads 6 KB of shared memory per 32 tensor-core cycles in one CTA (192 B/clk against the 128 B/clk an SM delivers); as a
// pair each SM reads 5 KB. Barriers: TMA bytes of both CTAs are credited to the leader's "full" barriers, MMA completion is
// multicast to both CTAs' "empty" / "accumulator full" barriers, the peer's epilogue arrives remotely on the leader's
// "accumulator empty" barrier.
template <int BLOCK_N, int MODE_T, bool SPLIT = false, bool PAIR = false>
__global__ void __launch_bounds__((c3_threads<BLOCK_N, MODE_T, SPLIT>()), 1) conv3x3_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b,
                                                                const __grid_constant__ CUtensorMap map_out,
                                                                const Conv3Args args) {
    using Cfg = C3Cfg<BLOCK_N, SPLIT>;
    static_assert(!SPLIT || MODE_T == 0 || MODE_T == 1, "SPLIT epilogue exists for conv3x3 forward and ConvT forward");
    static_assert(!PAIR || (BLOCK_N == 64 && MODE_T == 0 && !SPLIT), "CTA pairs exist for the N = 64 conv3x3 kernel");
    const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = pair_rank == 0;
    constexpr int SB = Cfg::SB, NH = Cfg::NH, BN_SLOT = Cfg::BN_SLOT;
    constexpr int TEAMS = c3_teams<BLOCK_N, MODE_T, SPLIT>();
    const int SA = args.sa;
    // MODE 4 = MODE 0 (3x3 conv / dgrad) + the ReLU-mask epilogue: a separate instantiation, so the kernels of the UNet
    // hot path carry none of its registers or branches (measured: +2 % per step when it was a runtime branch of MODE 0)
    constexpr bool MASKED = C3_HAS_MASK && MODE_T == 4;
    constexpr int MODE = MODE_T == 4 ? 0 : MODE_T;
    constexpr int NG = MODE == 0 ? 3 : (MODE == 2 ? 4 : 1);  // activation boxes per 64-channel chunk
    constexpr int NT = MODE == 0 ? 3 : 1;                    // taps served by one box
    constexpr int A_BYTES = MODE == 0 ? C3_A_SLOT : C3_TILE_H * C3_TILE_W * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring_a = smem;
    uint8_t* ring_b = smem + SA * C3_A_SLOT;
    uint8_t* out_stage = smem + Cfg::RING_BYTES;  // 1024-byte aligned (all slot sizes are multiples of 1024)
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES + Cfg::STAGING);
    uint64_t* a_empty = a_full + C3_SA_MAX;
    uint64_t* b_full = a_empty + C3_SA_MAX;
    uint64_t* b_empty = b_full + SB;
    uint64_t* acc_full = b_empty + SB;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // per-column affine (eval-mode folded BatchNorm scale/shift, ConvT bias) of this CTA's column block, staged once
    float* s_scale = reinterpret_cast<float*>(smem + Cfg::RING_BYTES + Cfg::STAGING + 1024);
    float* s_shift = s_scale + 256;

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], PAIR ? 256 : 128 * TEAMS);  // pair: both CTAs' epilogue threads arrive on the leader's
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) {
            tmem_alloc_2sm(tmem_ptr_smem, 2 * BLOCK_N);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(tmem_ptr_smem, 2 * BLOCK_N);
            tmem_relinquish();
        }
    }
    // PDL: everything above ran while the previous kernel of the stream was still draining; from here on its results are
    // needed. The trigger comes AFTER the wait so that "this grid has started" implies "its predecessors have completed".
    griddep_wait();
    griddep_launch_dependents();
    if (warp >= 4 && warp < 8 && blockIdx.x < args.num_tiles) {
        // the tile schedule keeps a CTA on one column block, so its affine vectors can be staged once
        const int n0 = (static_cast<int>(blockIdx.x) % args.n_tiles) * BLOCK_N;
        for (int i = threadIdx.x - 128; i < BLOCK_N; i += 128) {
            const int c = MODE == 1 ? (n0 + i) % args.cout_t : n0 + i;
            s_scale[i] = args.col_scale ? __ldg(args.col_scale + c) : 1.f;
            s_shift[i] = args.col_shift ? __ldg(args.col_shift + c) : 0.f;
        }
    }
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();  // both CTAs' barriers are initialised and their TMEM allocated before anything crosses over
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int chunks = args.cin_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
                const int n_tile = tile % args.n_tiles;
                const int m_tile = tile / args.n_tiles;
                const int img = m_tile / args.tiles_hw;
                const int t_in = m_tile - img * args.tiles_hw;
                const int h0 = (t_in / args.tiles_w) * C3_TILE_H;
                const int w0 = (t_in % args.tiles_w) * C3_TILE_W;
                const int n0 = n_tile * BLOCK_N;
                for (int c = 0; c < chunks; ++c) {
                    for (int dw = 0; dw < NG; ++dw) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        if (PAIR) {
                            // both CTAs load their own box; the bytes of both are credited to the leader's barrier
                            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * A_BYTES);
                            tma_load_4d_2sm(&map_a, &a_full[sa], ring_a + sa * C3_A_SLOT, c * 64, w0 + dw - 1, h0 - 1, img);
                        } else {
                        mbar_arrive_expect_tx(&a_full[sa], A_BYTES);
                        if (MODE == 0)
                            tma_load_4d(&map_a, &a_full[sa], ring_a + sa * C3_A_SLOT, c * 64, w0 + dw - 1, h0 - 1, img);
                        else if (MODE == 1 || MODE == 3)
                            tma_load_4d(&map_a, &a_full[sa], ring_a + sa * C3_A_SLOT, c * 64, w0, h0, img);
                        else
                            tma_load_5d(&map_a, &a_full[sa], ring_a + sa * C3_A_SLOT, c * 64, dw & 1, w0, dw >> 1,
                                        img * args.H + h0);
                        }
                        if (++sa == SA) {
                            sa = 0;
                            pa ^= 1;
                        }
                        for (int dh = 0; dh < NT; ++dh) {
                            const int kcol = (MODE == 0 ? (dh * 3 + dw) : dw) * args.C + c * 64;
#pragma unroll
                            for (int nh = 0; nh < NH; ++nh) {
                                if (PAIR) {
                                    // resident weights (host guarantees it): each CTA keeps 32 of the 64 rows
                                    if (tile == static_cast<int>(blockIdx.x)) {
                                        if (leader) mbar_arrive_expect_tx(&b_full[sb], Cfg::B_SLOT);  // 2 x 4 KB
                                        tma_load_2d_2sm(&map_b, &b_full[sb], ring_b + sb * Cfg::B_SLOT, kcol,
                                                        n0 + static_cast<int>(pair_rank) * (BN_SLOT / 2));
                                        ++sb;
                                    }
                                    continue;
                                }
                                if (args.b_resident) {
                                    // slot = position in the (fixed) per-tile consumption order; loaded once
                                    if (tile == static_cast<int>(blockIdx.x)) {
                                        mbar_arrive_expect_tx(&b_full[sb], Cfg::B_SLOT);
                                        tma_load_2d(&map_b, &b_full[sb], ring_b + sb * Cfg::B_SLOT, kcol,
                                                    n0 + nh * BN_SLOT);
                                        ++sb;
                                    }
                                    continue;
                                }
                                mbar_wait(&b_empty[sb], pb ^ 1);
                                mbar_arrive_expect_tx(&b_full[sb], Cfg::B_SLOT);
                                tma_load_2d(&map_b, &b_full[sb], ring_b + sb * Cfg::B_SLOT, kcol, n0 + nh * BN_SLOT);
                                if (++sb == SB) {
                                    sb = 0;
                                    pb ^= 1;
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if ((!PAIR || leader) && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN_SLOT, 0, 0);
            int sa = 0, sb = 0, as = 0;
            uint32_t pa = 0, pb = 0, pacc = 0;
            for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty[as], pacc ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + as * BLOCK_N;
                const bool first_tile = tile == static_cast<int>(blockIdx.x);
                if (args.b_resident) sb = 0;
                for (int c = 0; c < chunks; ++c) {
                    for (int dw = 0; dw < NG; ++dw) {
                        mbar_wait(&a_full[sa], pa);
                        const uint32_t a_addr = smem_u32(ring_a + sa * C3_A_SLOT);
                        for (int dh = 0; dh < NT; ++dh) {
                            const uint64_t da = umma_smem_desc_sw128(a_addr + dh * 1024, 0, 1024);
#pragma unroll
                            for (int nh = 0; nh < NH; ++nh) {
                                if (!args.b_resident || first_tile) mbar_wait(&b_full[sb], pb);
                                tc_fence_after();
                                const uint64_t db = umma_smem_desc_sw128(smem_u32(ring_b + sb * Cfg::B_SLOT), 0, 1024);
                                const uint32_t first = (c | dw | dh) == 0 ? 0u : 1u;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (PAIR)
                                        umma_bf16_2sm(d_base + nh * BN_SLOT, da + 2 * k, db + 2 * k, idesc, first | k);
                                    else
                                        umma_bf16(d_base + nh * BN_SLOT, da + 2 * k, db + 2 * k, idesc, first | k);
                                }
                                if (args.b_resident) {
                                    ++sb;
                                } else {
                                    umma_commit(&b_empty[sb]);
                                    if (++sb == SB) {
                                        sb = 0;
                                        pb ^= 1;
                                    }
                                }
                            }
                        }
                        if (PAIR)
                            umma_commit_2sm(&a_empty[sa]);  // frees the slot in both CTAs
                        else
                            umma_commit(&a_empty[sa]);
                        if (++sa == SA) {
                            sa = 0;
                            pa ^= 1;
                        }
                    }
                }
                if (PAIR)
                    umma_commit_2sm(&acc_full[as]);  // both CTAs' epilogues
                else
                    umma_commit(&acc_full[as]);
                if (++as == 2) {
                    as = 0;
                    pacc ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int team = (warp - 4) >> 2;  // 0, or 1 for the second epilogue team of the ConvTranspose modes
        const int bar_a = 1 + 2 * team, bar_b = 2 + 2 * team;  // named barriers of this team
        const int row = q * 32 + lane;  // pixel inside the tile: row = h_local * 8 + w_local
        const bool do_stats = args.stats != nullptr;
        const bool affine = args.col_scale != nullptr || args.col_shift != nullptr;
        const bool issuer = threadIdx.x == 128 + 128 * team;  // issues this team's TMA stores
        float st_sum[BLOCK_N / 32], st_sq[BLOCK_N / 32];
#pragma unroll
        for (int i = 0; i < BLOCK_N / 32; ++i) st_sum[i] = st_sq[i] = 0.f;
        int as = 0;
        uint32_t pacc = 0;
        int n0_last = 0;
        uint32_t sbuf = 0;  // staging buffer used by the next 64-column group
        // swizzled position of this thread's pixel row inside a staging buffer (128-byte rows, 16-byte chunk j is
        // stored at chunk j ^ (row & 7)): the layout a SWIZZLE_128B TMA store expects, and conflict-free to write
        const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
        const uint32_t row_xor = static_cast<uint32_t>(row & 7);
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % args.n_tiles;
            const int m_tile = tile / args.n_tiles;
            const int img = m_tile / args.tiles_hw;
            const int t_in = m_tile - img * args.tiles_hw;
            const int h0 = (t_in / args.tiles_w) * C3_TILE_H;
            const int w0 = (t_in % args.tiles_w) * C3_TILE_W;
            const int n0 = n_tile * BLOCK_N;
            n0_last = n0;
            const bool valid = !args.ragged || (h0 + (row >> 3) < args.H && w0 + (row & 7) < args.W);
            mbar_wait(&acc_full[as], pacc);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
#pragma unroll
            for (int grp = 0; grp < BLOCK_N / 64; ++grp) {
                if (SPLIT) {
                    // both staging buffers per group: [0] = hi, [1] = lo; the previous group's stores must have been read
                    if (issuer) tma_store_wait_read<0>();
                    named_bar_sync(2, 128);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int chunk = grp * 2 + half;
                        uint32_t raw[32];
                        tmem_ld32(t_addr + chunk * 32, raw);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                        if (affine) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 sc = *reinterpret_cast<const float4*>(s_scale + chunk * 32 + i);
                                const float4 sh = *reinterpret_cast<const float4*>(s_shift + chunk * 32 + i);
                                v[i] = fmaf(v[i], sc.x, sh.x);
                                v[i + 1] = fmaf(v[i + 1], sc.y, sh.y);
                                v[i + 2] = fmaf(v[i + 2], sc.z, sh.z);
                                v[i + 3] = fmaf(v[i + 3], sc.w, sh.w);
                            }
                        }
                        if (args.relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        uint32_t phi[16], plo[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            phi[i] = *reinterpret_cast<const uint32_t*>(&h2);
                            plo[i] = pack_bf16x2(v[2 * i] - __low2float(h2), v[2 * i + 1] - __high2float(h2));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t off = row_off + (((static_cast<uint32_t>(half * 4 + i)) ^ row_xor) << 4);
                            *reinterpret_cast<uint4*>(out_stage + off) =
                                make_uint4(phi[4 * i], phi[4 * i + 1], phi[4 * i + 2], phi[4 * i + 3]);
                            *reinterpret_cast<uint4*>(out_stage + C3_OUT_STAGE + off) =
                                make_uint4(plo[4 * i], plo[4 * i + 1], plo[4 * i + 2], plo[4 * i + 3]);
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (issuer) {
                        const int col0 = n0 + grp * 64;
                        const int S = args.split_stride;
                        if (MODE == 1) {
                            const int ij = col0 / args.cout_t;
                            const int co = col0 - ij * args.cout_t;
                            tma_store_5d(&map_out, out_stage, co, ij & 1, w0, ij >> 1, img * args.H + h0);
                            tma_store_5d(&map_out, out_stage + C3_OUT_STAGE, co + S, ij & 1, w0, ij >> 1, img * args.H + h0);
                            tma_store_5d(&map_out, out_stage, co + 2 * S, ij & 1, w0, ij >> 1, img * args.H + h0);
                        } else {
                            tma_store_4d(&map_out, out_stage, col0, w0, h0, img);
                            tma_store_4d(&map_out, out_stage + C3_OUT_STAGE, col0 + S, w0, h0, img);
                            tma_store_4d(&map_out, out_stage, col0 + 2 * S, w0, h0, img);
                        }
                        tma_store_commit();
                    }
                    continue;
                }
                if (TEAMS == 2 && (grp & 1) != team) continue;  // the other team's group
                uint8_t* stage = out_stage + (TEAMS == 2 ? team : sbuf) * C3_OUT_STAGE;
                if (Cfg::NBUF == 1 || TEAMS == 2) {
                    // single staging buffer (per team): the previous store must have finished reading it before it is rewritten
                    if (issuer) tma_store_wait_read<0>();
                    named_bar_sync(bar_b, 128);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int chunk = grp * 2 + half;
                    uint4 mraw[4];
                    if (MASKED) {
                        // issued before the TMEM load so that the global-load latency overlaps it
                        const size_t pix = (static_cast<size_t>(img) * args.H + h0 + (row >> 3)) * args.W + w0 + (row & 7);
                        const uint4* mp = reinterpret_cast<const uint4*>(args.mask + pix * args.mask_pix_stride +
                                                                         args.mask_c_off + n0 + chunk * 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) mraw[i] = valid ? __ldg(mp + i) : make_uint4(0u, 0u, 0u, 0u);
                    }
                    uint32_t raw[32];
                    tmem_ld32(t_addr + chunk * 32, raw);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                    if (affine) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 sc = *reinterpret_cast<const float4*>(s_scale + chunk * 32 + i);
                            const float4 sh = *reinterpret_cast<const float4*>(s_shift + chunk * 32 + i);
                            v[i] = fmaf(v[i], sc.x, sh.x);
                            v[i + 1] = fmaf(v[i + 1], sc.y, sh.y);
                            v[i + 2] = fmaf(v[i + 2], sc.z, sh.z);
                            v[i + 3] = fmaf(v[i + 3], sc.w, sh.w);
                        }
                    }
                    if (args.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (MASKED) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mraw[i]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (!(__low2float(mh[j]) > 0.f)) v[i * 8 + 2 * j] = 0.f;
                                if (!(__high2float(mh[j]) > 0.f)) v[i * 8 + 2 * j + 1] = 0.f;
                            }
                        }
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) packed[i] = valid ? pack_bf16x2(v[2 * i], v[2 * i + 1]) : 0u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t j = static_cast<uint32_t>(half * 4 + i);
                        *reinterpret_cast<uint4*>(stage + row_off + ((j ^ row_xor) << 4)) =
                            make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
                    }
                    if (do_stats && (args.stats_sum_cols == 0 || n0 + chunk * 32 < args.stats_sum_cols)) {
                        // statistics of the tensor as stored (bf16-rounded), like BatchNorm reading the conv output
                        float s1[32], s2[32];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&packed[i]);
                            const float a = __low2float(hh), b = __high2float(hh);
                            s1[2 * i] = a;
                            s1[2 * i + 1] = b;
                            s2[2 * i] = a * a;
                            s2[2 * i + 1] = b * b;
                        }
                        st_sum[chunk] += warp_transpose_reduce32(s1, lane);
                        if (args.stats_sum_cols == 0) st_sq[chunk] += warp_transpose_reduce32(s2, lane);
                    }
                }
                // make the generic-proxy writes visible to the TMA engine, make sure the OTHER buffer's previous
                // store has finished reading shared memory (it is the next one to be overwritten), then store
                fence_proxy_async_smem();
                if (Cfg::NBUF == 2 && TEAMS == 1 && issuer) tma_store_wait_read<0>();
                named_bar_sync(bar_a, 128);
                if (issuer) {
                    const int col0 = n0 + grp * 64;
                    if (MODE == 1) {
                        const int ij = col0 / args.cout_t;
                        const int co = col0 - ij * args.cout_t;
                        tma_store_5d(&map_out, stage, co, ij & 1, w0, ij >> 1, img * args.H + h0);
                    } else {
                        tma_store_4d(&map_out, stage, col0, w0, h0, img);
                    }
                    tma_store_commit();
                }
                if (Cfg::NBUF == 2 && TEAMS == 1) sbuf ^= 1;
            }
            // all TMEM reads of this thread have completed (tmem_ld_wait above): hand the accumulator back
            tc_fence_before();
            if (PAIR && !leader)
                mbar_arrive_cluster(&acc_empty[as], 0);  // the leader's MMA thread waits for both CTAs' epilogues
            else
                mbar_arrive(&acc_empty[as]);
            if (++as == 2) {
                as = 0;
                pacc ^= 1;
            }
        }
        if (issuer) tma_store_wait<0>();  // global writes complete before the CTA retires
        if (do_stats && blockIdx.x < args.num_tiles) {
            // the tile schedule keeps this CTA on one column block (gridDim.x % n_tiles == 0, or one tile per CTA)
            if (args.stats_slots > 0) {
                // deterministic: this CTA owns slot (blockIdx.x / n_tiles) of its column block — plain stores; the slots no
                // CTA owns are zeroed by the CTA that owns slot (s mod used), so bn_finalize can sum all of them in order
                const int used = static_cast<int>(gridDim.x) / args.n_tiles;
                const int mine = static_cast<int>(blockIdx.x) / args.n_tiles;
                float* s_red = reinterpret_cast<float*>(out_stage);  // staging tiles are dead (stores drained above)
                constexpr int EW = 4 * TEAMS;  // epilogue warps; each holds the sums of ITS pixel rows x ITS column groups
                const int ew = warp - 4;
                named_bar_sync(5, 128 * TEAMS);
#pragma unroll
                for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                    s_red[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = st_sum[chunk];
                    s_red[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = st_sq[chunk];
                }
                named_bar_sync(5, 128 * TEAMS);
                for (int i = threadIdx.x - 128; i < 2 * BLOCK_N; i += 128 * TEAMS) {
                    const int which = i / BLOCK_N, col = i - which * BLOCK_N;
                    float acc = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < EW; ++w8) acc += s_red[(w8 * 2 + which) * BLOCK_N + col];
                    args.stats[(static_cast<size_t>(mine) * 2 + which) * args.n_total + n0_last + col] = acc;
                    for (int s2 = mine + used; s2 < args.stats_replicas; s2 += used)
                        args.stats[(static_cast<size_t>(s2) * 2 + which) * args.n_total + n0_last + col] = 0.f;
                }
                __threadfence();  // this CTA's slot is visible device-wide before its ticket (fused finalize) is drawn
            } else {
                float* dst = args.stats + static_cast<size_t>(blockIdx.x % args.stats_replicas) * 2 * args.n_total + n0_last;
#pragma unroll
                for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                    atomicAdd(dst + chunk * 32 + lane, st_sum[chunk]);
                    atomicAdd(dst + args.n_total + chunk * 32 + lane, st_sq[chunk]);
                }
            }
        }
    }

    tc_fence_before();
    if (PAIR)
        cluster_sync_all();  // neither CTA's TMEM / barriers go away while the other may still touch them
    else
        __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR)
            tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
        else
            tmem_dealloc(tmem_base, 2 * BLOCK_N);
    }
    if (MODE == 0 && !SPLIT && args.bn_scale != nullptr && blockIdx.x < args.num_tiles) {
        // fused BatchNorm finalize: one ticket per CTA of this column block; the last one sums the slots in slot order
        __shared__ int s_bn_last;
        const int n_tile = static_cast<int>(blockIdx.x) % args.n_tiles;
        const int used = static_cast<int>(gridDim.x) / args.n_tiles;
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(args.bn_counters + n_tile, 1u);
            s_bn_last = (t == static_cast<unsigned>(used) - 1u) ? 1 : 0;
            if (s_bn_last) args.bn_counters[n_tile] = 0u;
        }
        __syncthreads();
        if (s_bn_last) {
            __threadfence();
            c3_bn_finalize<BLOCK_N>(args, n_tile * BLOCK_N, used, out_stage);
        }
    }
}

}  // namespace b200sr
