// Implicit-GEMM forward-style kernel on tcgen05/TMEM fed by TMA (sm_100a).
//
//   D[pixel, n] = sum_{tap, c} A[pixel + tap_offset, c] * Wp[n, tap*C + c]
//
// A is an NHWC bf16 activation tensor (or a channel-slot view of a concat buffer); every (tap, 64-channel) K-block
// is ONE TMA box load {64 ch, 16 w, 8 h, 1 img} whose out-of-bounds elements are zero-filled by the TMA unit,
// which is exactly the zero padding of Conv2d(padding=1). The box lands in shared memory as [128 pixels][128 B]
// with the 128-byte swizzle = the canonical K-major UMMA operand layout, so no thread ever touches the operands.
// Wp is the packed K-major weight matrix [N][taps*C] (bf16).
//
// The same kernel serves (reference ops in /root/reference/src/unet_model.py):
//   * Conv2d 3x3 p1 forward           (:27,:30)  9 taps, offsets (kh-1, kw-1)
//   * Conv2d 3x3 dgrad                            9 taps, weights rotated/transposed by the packer
//   * ConvTranspose2d k2 s2 forward   (:67-76)   1 tap, N = 4*Cout, pixel-shuffle scatter epilogue + bias
//   * ConvTranspose2d k2 s2 dgrad                 4 taps gathered through a 5-D tensor map of the 2H x 2W grad
//
// Epilogue (4 warps, one TMEM lane = one output pixel per thread): optional per-column affine (+ReLU) for the
// eval-mode folded BatchNorm / ConvT bias, bf16 NHWC store into an arbitrary channel slot (this is how
// torch.cat disappears), and optional per-column sum / sum-of-squares for train-mode BatchNorm statistics.
#pragma once
#include "ptx.cuh"

namespace b200sr {

struct IGemmArgs {
    int H, W;            // spatial size of the M-space (per image); M = B*H*W
    int tiles_w;         // W / 16
    int tiles_hw;        // (H / 8) * (W / 16)
    int a_mode;          // 0: 4-D NHWC taps, 1: 5-D convT-dgrad gather (c, j, w, i, b*h)
    int num_taps;        // 9, 1 or 4
    int kc_per_tap;      // channels of A per tap / 64
    int n_total;         // total GEMM-N
    int n_tiles;         // n_total / BLOCK_N
    int epi_mode;        // 0: NHWC store, 1: ConvTranspose pixel-shuffle scatter
    int cout_t;          // epi_mode 1: channels per (i,j) sub-pixel; also modulus of the affine vectors
    int relu;
    int out_pix_stride;  // elements between consecutive output pixels (total channels of the buffer)
    int out_c_off;       // first channel of the slot written
    int stats_replicas;
    int stats_slots;     // 1: deterministic — M tile t STORES into slot t (needs stats_replicas >= #M tiles), rest zeroed
    __nv_bfloat16* out;
    const float* col_scale;  // nullable
    const float* col_shift;  // nullable
    float* stats;            // nullable, [replicas][2][n_total]
};

constexpr int IG_BLOCK_M = 128;
constexpr int IG_BLOCK_K = 64;
constexpr int IG_TILE_W = 16;
constexpr int IG_TILE_H = 8;
constexpr int IG_THREADS = 256;
constexpr int IG_A_BYTES = IG_BLOCK_M * IG_BLOCK_K * 2;  // 16 KB

template <int BLOCK_N>
__host__ __device__ constexpr int ig_stage_bytes() {
    return IG_A_BYTES + BLOCK_N * IG_BLOCK_K * 2;
}
template <int BLOCK_N, int STAGES>
__host__ __device__ constexpr int ig_smem_bytes() {
    // operands + barriers/tmem ptr (256 B) + stats staging + 1 KB alignment slack
    return STAGES * ig_stage_bytes<BLOCK_N>() + 256 + 8 * BLOCK_N * 4 + 1024;
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(IG_THREADS) igemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const IGemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle atoms are 1024 B: align the operand ring by hand.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int STAGE_BYTES = ig_stage_bytes<BLOCK_N>();
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* s_stats = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [4 warps][2][BLOCK_N]

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    // tile coordinates: N-tile fastest so CTAs sharing an A tile run together (L2 reuse)
    const int n_tile = blockIdx.x % args.n_tiles;
    const int m_tile = blockIdx.x / args.n_tiles;
    const int img = m_tile / args.tiles_hw;
    const int t_in = m_tile - img * args.tiles_hw;
    const int h0 = (t_in / args.tiles_w) * IG_TILE_H;
    const int w0 = (t_in % args.tiles_w) * IG_TILE_W;
    const int n0 = n_tile * BLOCK_N;
    const int num_kb = args.num_taps * args.kc_per_tap;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
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
        tmem_alloc(tmem_ptr_smem, BLOCK_N);
        tmem_relinquish();
    }
    if (warp >= 4) {
        for (int i = threadIdx.x - 128; i < 8 * BLOCK_N; i += 128) s_stats[i] = 0.f;
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
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = ring + stage * STAGE_BYTES;
                uint8_t* sb = sa + IG_A_BYTES;
                const int tap = kb / args.kc_per_tap;
                const int c0 = (kb - tap * args.kc_per_tap) * IG_BLOCK_K;
                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                if (args.a_mode == 0) {
                    int dh = 0, dw = 0;
                    if (args.num_taps == 9) {
                        dh = tap / 3 - 1;
                        dw = tap % 3 - 1;
                    }
                    tma_load_4d(&map_a, &full_bar[stage], sa, c0, w0 + dw, h0 + dh, img);
                } else {
                    // (c, j, w, i, img*H + h) view of the (2H x 2W) tensor; tap = i*2 + j
                    tma_load_5d(&map_a, &full_bar[stage], sa, c0, tap & 1, w0, tap >> 1, img * args.H + h0);
                }
                tma_load_2d(&map_b, &full_bar[stage], sb, kb * IG_BLOCK_K, n0);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(IG_BLOCK_M, BLOCK_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(ring + stage * STAGE_BYTES);
                const uint32_t sb = sa + IG_A_BYTES;
                const uint64_t da = umma_smem_desc_sw128(sa, 0, 1024);
                const uint64_t db = umma_smem_desc_sw128(sb, 0, 1024);
#pragma unroll
                for (int k = 0; k < IG_BLOCK_K / 16; ++k) {
                    // +32 B per UMMA_K inside the 128 B swizzle span (encoded >> 4)
                    umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tmem_full_bar);  // accumulator complete
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;             // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;      // pixel inside the tile
        const int h = h0 + row / IG_TILE_W;
        const int w = w0 + row % IG_TILE_W;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const bool do_stats = args.stats != nullptr;

#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + chunk * 32, raw);
            tmem_ld_wait();
            const int col0 = n0 + chunk * 32;  // global GEMM column of raw[0]
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);

            if (args.col_scale != nullptr || args.col_shift != nullptr) {
                const int a0 = col0 % args.cout_t;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float sc = args.col_scale ? __ldg(args.col_scale + a0 + i) : 1.f;
                    const float sh = args.col_shift ? __ldg(args.col_shift + a0 + i) : 0.f;
                    v[i] = fmaf(v[i], sc, sh);
                }
            }
            if (args.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }

            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);

            size_t off;
            if (args.epi_mode == 0) {
                off = (static_cast<size_t>(img * args.H + h) * args.W + w) * args.out_pix_stride + args.out_c_off +
                      col0;
            } else {
                const int ij = col0 / args.cout_t;
                const int co = col0 - ij * args.cout_t;
                const int oh = 2 * h + (ij >> 1);
                const int ow = 2 * w + (ij & 1);
                off = (static_cast<size_t>(img * 2 * args.H + oh) * (2 * args.W) + ow) * args.out_pix_stride +
                      args.out_c_off + co;
            }
            uint4* dst = reinterpret_cast<uint4*>(args.out + off);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);

            if (do_stats) {
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
                const float cs = warp_transpose_reduce32(s1, lane);
                const float cq = warp_transpose_reduce32(s2, lane);
                s_stats[(q * 2 + 0) * BLOCK_N + chunk * 32 + lane] = cs;  // own slot per warp: combined in warp order below
                s_stats[(q * 2 + 1) * BLOCK_N + chunk * 32 + lane] = cq;
            }
        }
        if (do_stats) {
            named_bar_sync(1, 128);
            const int m_tiles = static_cast<int>(gridDim.x) / args.n_tiles;
            for (int i = threadIdx.x - 128; i < 2 * BLOCK_N; i += 128) {
                const int which = i / BLOCK_N, col = i - which * BLOCK_N;
                float acc = 0.f;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) acc += s_stats[(w4 * 2 + which) * BLOCK_N + col];
                if (args.stats_slots) {
                    args.stats[(static_cast<size_t>(m_tile) * 2 + which) * args.n_total + n0 + col] = acc;
                    for (int s2 = m_tile + m_tiles; s2 < args.stats_replicas; s2 += m_tiles)
                        args.stats[(static_cast<size_t>(s2) * 2 + which) * args.n_total + n0 + col] = 0.f;
                } else {
                    atomicAdd(args.stats + (static_cast<size_t>(m_tile % args.stats_replicas) * 2 + which) * args.n_total + n0 + col,
                              acc);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BLOCK_N);
    }
}

}  // namespace b200sr
