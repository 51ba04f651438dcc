// Conv2d 3x3 weight gradient on tcgen05/TMEM fed by TMA, second generation (sm_100a).
// Reference op: autograd of nn.Conv2d(k=3, p=1) in UNetBlock, /root/reference/src/unet_model.py:27,30.
//
//   G[(dh,dw), ci, co] = sum_pixels X[pixel + (dh-1, dw-1), ci] * dZ[pixel, co]
//
// The reduction runs over pixels, so both operands are MN-major for the tensor core (a TMA box lands as
// [pixels][64 ch] rows of 128 B with the 128-byte swizzle). One pipeline stage covers 4 rows x 16 columns of
// pixels. What differs from the first-generation kernel (wgrad.cuh, still used for ConvTranspose2d):
//   * few, large TMA boxes: a 5-D view (64 ch, w, h, 64-channel chunk, image) lets ONE box carry several channel
//     chunks, so a stage is 2-4 TMA instructions instead of up to 17 (the producer thread was issue-bound);
//   * halo reuse: for a fixed horizontal tap dw one box of 6 rows serves the three vertical taps (a pixel row of
//     the box is 2048 B = two whole swizzle atoms, so tap dh / k-step r start (r + dh) * 2048 B into the box);
//   * table-driven M tiles (two 64-row MN atoms each, described by start offset + leading byte offset):
//       mode A (Cin % 128 == 0): CTA job = (dw, 128 input channels, N_TILE output channels); the 3 M tiles are the
//                                vertical taps, their two atoms the two channel chunks of the same box;
//       mode B (Cin == 64):      CTA job = (all 9 taps, 64 input channels, 64 output channels); three boxes (one
//                                per dw) sit side by side; 5 M tiles pair taps (one half of the last tile is unused).
//   * split-K over pixel ranges sized so the grid fills whole waves of SMs; partial sums are added to the fp32
//     workspace G[tap][ci][co] with vector reductions.
#pragma once
#include "ptx.cuh"

namespace b200sr {

struct WG3Args {
    int H, W;
    int chunks_w;        // W / 16
    int chunks_hw;       // (H / 4) * (W / 16)
    int total_chunks;    // B * chunks_hw
    int chunks_per_cta;  // split-K slice length
    int mode_b;          // 0: mode A, 1: mode B
    int jobs_ci;         // mode A: Cin / 128
    int jobs_co;         // Cout / N_TILE
    int Cin, Cout;
    float* out;          // [9][Cin][Cout] fp32, pre-zeroed, ADDED into (split_stride == 0)
    long long split_stride;  // > 0: deterministic mode — split-K slice s STORES its partial into out + s * split_stride
                             // (floats); a fixed-order second stage (wgrad_reduce_unpack_kernel) sums the slices
};

constexpr int WG3_THREADS = 256;
constexpr int WG3_ROW_BYTES = 16 * 128;          // one pixel row of a box: 16 pixels x 64 ch bf16
constexpr int WG3_XBOX = 6 * WG3_ROW_BYTES;      // 12288: one (dw, channel chunk) halo box of 6 rows
constexpr int WG3_ZBOX = 4 * WG3_ROW_BYTES;      // 8192: one 64-channel chunk of dZ, 4 rows

#ifndef WG3_SMEM_BUDGET_KB
#define WG3_SMEM_BUDGET_KB 225
#endif

template <int N_TILE, int MODE_B>
struct WG3Cfg {
    static constexpr int X_BYTES = (MODE_B ? 3 : 2) * WG3_XBOX;
    static constexpr int Z_BYTES = (N_TILE / 64) * WG3_ZBOX;
    static constexpr int STAGE_BYTES = X_BYTES + Z_BYTES;
    // shared-memory budget. Measured (tools/bench_overlap.py): capping it at 180 KB so that BatchNorm-backward blocks
    // can co-reside with a wgrad CTA costs nothing alone but does not buy stream overlap either — at 256^2 / 128^2 the
    // wgrad kernels themselves stream 2.8-4 TB/s of operands, so they and the BatchNorm passes share the HBM roofline.
    static constexpr int STAGES = (WG3_SMEM_BUDGET_KB * 1024) / STAGE_BYTES > 6 ? 6 : (WG3_SMEM_BUDGET_KB * 1024) / STAGE_BYTES;
    static constexpr int M_TILES = MODE_B ? 5 : 3;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
    static_assert(M_TILES * N_TILE <= 512, "accumulators exceed TMEM");
};

template <int N_TILE, int MODE_B>
__global__ void __launch_bounds__(WG3_THREADS, 1) wgrad3x3_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ CUtensorMap map_z,
                                                                  const WG3Args args) {
    using Cfg = WG3Cfg<N_TILE, MODE_B>;
    constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, M_TILES = Cfg::M_TILES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    // job decode: blockIdx.x = ((dw * jobs_ci) + ci_group) * jobs_co + co_block   (mode B: dw = ci_group = 0)
    const int co_block = blockIdx.x % args.jobs_co;
    const int rest = blockIdx.x / args.jobs_co;
    const int ci_group = MODE_B ? 0 : rest % args.jobs_ci;
    const int dw_job = MODE_B ? 0 : rest / args.jobs_ci;
    const int n0 = co_block * N_TILE;
    const int chunk_begin = blockIdx.y * args.chunks_per_cta;
    const int chunk_end = min(chunk_begin + args.chunks_per_cta, args.total_chunks);
    const int iters = max(chunk_end - chunk_begin, 0);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_z);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < iters; ++it) {
                const int chunk = chunk_begin + it;
                const int img = chunk / args.chunks_hw;
                const int r = chunk - img * args.chunks_hw;
                const int h0 = (r / args.chunks_w) * 4;
                const int w0 = (r % args.chunks_w) * 16;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = ring + stage * STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                if (MODE_B) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        tma_load_5d(&map_x, &full_bar[stage], st + d * WG3_XBOX, 0, w0 + d - 1, h0 - 1, 0, img);
                } else {
                    tma_load_5d(&map_x, &full_bar[stage], st, 0, w0 + dw_job - 1, h0 - 1, ci_group * 2, img);
                }
                tma_load_5d(&map_z, &full_bar[stage], st + Cfg::X_BYTES, 0, w0, h0, co_block * (N_TILE / 64), img);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, N_TILE, 1, 1);  // both operands MN-major
            // M-tile table: start offset of the first 64-row atom and byte distance to the second one
            uint32_t a_off[M_TILES], a_lbo[M_TILES];
            if (MODE_B) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    a_off[t] = t * WG3_ROW_BYTES;  // (dh = t, dw 0 | dw 1)
                    a_lbo[t] = WG3_XBOX;
                }
                a_off[3] = 2 * WG3_XBOX;                  // (dh 0 | dh 1, dw 2)
                a_lbo[3] = WG3_ROW_BYTES;
                a_off[M_TILES - 1] = 2 * WG3_XBOX + WG3_ROW_BYTES;  // (dh 1 [unused] | dh 2, dw 2)
                a_lbo[M_TILES - 1] = WG3_ROW_BYTES;
            } else {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    a_off[t] = t * WG3_ROW_BYTES;  // dh = t, atoms = the two channel chunks of the box
                    a_lbo[t] = WG3_XBOX;
                }
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(ring + stage * STAGE_BYTES);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    // MN-major SW128: LBO = bytes between 64-element MN atoms, SBO = bytes between 8-pixel K groups
                    const uint64_t db = umma_smem_desc_sw128(st + Cfg::X_BYTES + r * WG3_ROW_BYTES, WG3_ZBOX, 1024);
#pragma unroll
                    for (int t = 0; t < M_TILES; ++t) {
                        const uint64_t da = umma_smem_desc_sw128(st + a_off[t] + r * WG3_ROW_BYTES, a_lbo[t], 1024);
                        umma_bf16(tmem_base + t * N_TILE, da, db, idesc, (it | r) != 0);
                    }
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tmem_full_bar);
        }
    } else if (warp >= 4 && iters > 0) {
        // ===================== epilogue: TMEM -> red.add into the fp32 workspace =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int half = row >> 6;  // warp-uniform
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < M_TILES; ++t) {
            int tap, ci;
            if (MODE_B) {
                if (t < 3) {
                    tap = t * 3 + half;
                } else if (t == 3) {
                    tap = half * 3 + 2;
                } else {
                    if (half == 0) continue;
                    tap = 2 * 3 + 2;
                }
                ci = 0;
            } else {
                tap = t * 3 + dw_job;
                ci = ci_group * 128 + half * 64;
            }
            float* dst_row = args.out + static_cast<size_t>(blockIdx.y) * args.split_stride +
                             (static_cast<size_t>(tap) * args.Cin + ci + (row & 63)) * args.Cout + n0;
#pragma unroll 1
            for (int chunk = 0; chunk < N_TILE / 32; ++chunk) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * N_TILE + chunk * 32, raw);
                tmem_ld_wait();
                if (args.split_stride > 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(dst_row + chunk * 32 + 4 * i) =
                            make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                                        __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        red_add_v4_f32(dst_row + chunk * 32 + 4 * i, __uint_as_float(raw[4 * i]),
                                       __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                                       __uint_as_float(raw[4 * i + 3]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200sr
