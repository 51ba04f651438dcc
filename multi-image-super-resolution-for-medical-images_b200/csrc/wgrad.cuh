// Weight-gradient implicit GEMM on tcgen05/TMEM fed by TMA (sm_100a).
//
//   G[(tap, a), n] = sum_pixels  T[pixel + tap_offset, a] * P[pixel, n]
//
// The reduction (GEMM-K) runs over PIXELS, so both operands are "MN-major" for the tensor core: a TMA box
// {64 ch, 16 w, 2 h, 1 img} lands as [32 pixels][128 B] with the 128-byte swizzle, which is the canonical
// MN-major SWIZZLE_128B UMMA layout (K rows of 128 B, 8-row swizzle atoms, SBO = 1024 B) — the very same
// box shape the forward kernel reads K-major. T is the tapped tensor, P the plain one:
//   * Conv2d 3x3 wgrad (unet_model.py:27,30):  T = layer input x (zero-filled halo via TMA OOB), P = dZ
//       -> G[tap][ci][co]
//   * ConvTranspose2d k2 s2 wgrad (:67-76):     T = grad of the upsampled map, gathered per sub-pixel (i,j)
//       through the 5-D view (c, j, w, i, b*h); P = layer input x  -> G[(i,j)][co][ci]
//
// One CTA owns up to 2*M_TILES "atoms" (tap, 64-channel chunk of T) = M_TILES accumulators of 128 x N_TILE
// fp32 in TMEM (all 512 columns), and a slice of the pixel range (split-K). Partial sums are added into the
// fp32 workspace G with vector reductions (red.global.add.v4.f32); a later kernel transposes G into the
// PyTorch parameter layout.
#pragma once
#include "ptx.cuh"

namespace b200sr {

struct WGradArgs {
    int H, W;            // spatial size of the pixel space (per image) the reduction runs over
    int chunks_w;        // ceil(W / 16)
    int chunks_hw;       // ceil(H / rows) * chunks_w, rows = KPIX / 16 pixel rows per pipeline stage
    int total_chunks;    // B * chunks_hw
    int chunks_per_cta;  // split-K slice length
    int t_mode;          // 0: 4-D taps on T (3x3: 9 taps, 1x1: 1 tap), 1: 5-D gather (4 taps)
    int num_taps;
    int tchunks_per_tap;  // channels of T / 64
    int total_atoms;      // num_taps * tchunks_per_tap
    int n_total;          // channels of P
    float* out;           // [total_atoms*64][n_total] fp32, pre-zeroed (split_stride == 0)
    long long split_stride;  // > 0: split-K slice s STORES its partial into out + s * split_stride (deterministic mode)
};

constexpr int WG_THREADS = 256;
// KPIX = pixels per pipeline stage: 32 (2 rows x 16 cols) or 64 (4 rows x 16 cols; the N_TILE = 256 variant, whose stage then
// carries 8 MMAs of 128 x 256 x 16 per barrier round trip instead of 4). One atom = KPIX pixels x 64 ch x 2 B.
template <int KPIX>
__host__ __device__ constexpr int wg_atom_bytes() {
    return KPIX * 128;
}
template <int N_TILE>
__host__ __device__ constexpr int wg_atoms_per_cta() {
    return 2 * (512 / N_TILE);
}
template <int N_TILE, int KPIX>
__host__ __device__ constexpr int wg_stage_bytes() {
    return (wg_atoms_per_cta<N_TILE>() + N_TILE / 64) * wg_atom_bytes<KPIX>();
}
template <int N_TILE, int STAGES, int KPIX>
__host__ __device__ constexpr int wg_smem_bytes() {
    return STAGES * wg_stage_bytes<N_TILE, KPIX>() + 256 + 1024;
}

template <int N_TILE, int STAGES, int KPIX>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ CUtensorMap map_t,
                                                              const __grid_constant__ CUtensorMap map_p,
                                                              const WGradArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int AT = wg_atoms_per_cta<N_TILE>();
    constexpr int STAGE_BYTES = wg_stage_bytes<N_TILE, KPIX>();
    constexpr int WG_ATOM_BYTES = wg_atom_bytes<KPIX>();
    constexpr int ROWS = KPIX / 16;
    constexpr int NB = N_TILE / 64;
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    const int atom0 = blockIdx.x * AT;
    const int n_atoms = min(AT, args.total_atoms - atom0);
    const int m_tiles = (n_atoms + 1) >> 1;
    const int n0 = blockIdx.y * N_TILE;
    const int chunk_begin = blockIdx.z * args.chunks_per_cta;
    const int chunk_end = min(chunk_begin + args.chunks_per_cta, args.total_chunks);
    const int iters = max(chunk_end - chunk_begin, 0);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_t);
        tma_prefetch_desc(&map_p);
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
            const uint32_t tx = static_cast<uint32_t>(n_atoms + NB) * WG_ATOM_BYTES;
            for (int it = 0; it < iters; ++it) {
                const int chunk = chunk_begin + it;
                const int img = chunk / args.chunks_hw;
                const int r = chunk - img * args.chunks_hw;
                const int h0 = (r / args.chunks_w) * ROWS;
                const int w0 = (r % args.chunks_w) * 16;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = ring + stage * STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], tx);
                for (int a = 0; a < n_atoms; ++a) {
                    const int atom = atom0 + a;
                    const int tap = atom / args.tchunks_per_tap;
                    const int c0 = (atom - tap * args.tchunks_per_tap) * 64;
                    if (args.t_mode == 0) {
                        int dh = 0, dw = 0;
                        if (args.num_taps == 9) {
                            dh = tap / 3 - 1;
                            dw = tap % 3 - 1;
                        }
                        tma_load_4d(&map_t, &full_bar[stage], st + a * WG_ATOM_BYTES, c0, w0 + dw, h0 + dh, img);
                    } else {
                        tma_load_5d(&map_t, &full_bar[stage], st + a * WG_ATOM_BYTES, c0, tap & 1, w0, tap >> 1,
                                    img * args.H + h0);
                    }
                }
                for (int b = 0; b < NB; ++b)
                    tma_load_4d(&map_p, &full_bar[stage], st + (AT + b) * WG_ATOM_BYTES, n0 + b * 64, w0, h0, img);
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
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(ring + stage * STAGE_BYTES);
                // MN-major SW128: LBO = bytes between 64-element MN atoms, SBO = bytes between 8-row K groups
                const uint64_t db = umma_smem_desc_sw128(st + AT * WG_ATOM_BYTES, WG_ATOM_BYTES, 1024);
                for (int mt = 0; mt < m_tiles; ++mt) {
                    const uint64_t da = umma_smem_desc_sw128(st + 2 * mt * WG_ATOM_BYTES, WG_ATOM_BYTES, 1024);
#pragma unroll
                    for (int k = 0; k < KPIX / 16; ++k) {
                        // 16 pixels = 16 rows of 128 B = 2048 B (>>4 = 128)
                        umma_bf16(tmem_base + mt * N_TILE, da + 128 * k, db + 128 * k, idesc, (it | k) != 0);
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
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int a = 2 * mt + (row >> 6);  // warp-uniform: a warp never straddles the two atoms
            if (a >= n_atoms) continue;
            float* dst_row = args.out + static_cast<size_t>(blockIdx.z) * args.split_stride +
                             (static_cast<size_t>(atom0 + a) * 64 + (row & 63)) * args.n_total + n0;
#pragma unroll 1
            for (int chunk = 0; chunk < N_TILE / 32; ++chunk) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mt * N_TILE + chunk * 32, raw);
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
