// First-layer 3x3 convolutions (Cin = 2: UNet enc1.conv.0, unet_model.py:49 -> :27; Cin = 3: the image channels of
// Fast-DDPM's inc.block.0, ModelLoader.py:554) straight from the fp32 NCHW network input, on warp-level tensor-core
// MMAs. K = 9*Cin = 18 / 27 is padded to 32; the layer is HBM-bound (arithmetic intensity 17 flop/B: it writes 64 bf16
// channels per pixel from 2-3 fp32 inputs), so the point of the MMA is only to get the 64-channel FMA work off the
// CUDA cores (the FFMA version sat at 27 % of the FP32 peak and 0.9 TB/s): m16n8k16 bf16 `mma.sync` with im2col
// fragments built in registers from a haloed fp32 tile in shared memory. (tcgen05 would need the im2col tile
// materialised in swizzled shared memory for 32 K-elements per pixel; at K=32 the legacy warp MMA is already far
// from being the limiter.) The forward pass keeps fp32 operand accuracy with the bf16x3 split (x = x_hi + x_lo,
// w = w_hi + w_lo; x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, the 2^-18 lo*lo term dropped): the network input is the one
// tensor of the path that is NOT bf16 already, and rounding it measurably moves the end-to-end parity (Progressive
// chain 2.4e-2 -> 2.6e-2); three MMAs instead of one are free in an HBM-bound kernel. The weight gradient uses single
// bf16 operands like every other wgrad of the path (its other operand, dZ, is bf16 anyway). Accumulation is fp32.
//
//   forward : D[pixel][co]  = sum_k A[pixel][k] * W[co][k]        (M = 16 pixels of one tile row, N = 64, K = 32)
//   wgrad   : G[k][co]     += sum_pixel A[pixel][k] * dZ[pixel][co] (M = 32 (k), N = 64, K = pixels)
#pragma once
#include "elementwise.cuh"

namespace b200sr {

constexpr int FC_TILE = 16;
constexpr int FC_HT = FC_TILE + 2;
constexpr int FC_COUT = 64;
constexpr int FC_PITCH = FC_COUT + 8;  // bf16 per staged pixel row: 144 B (conflict-free 4-byte and ldmatrix access)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = smem_u32(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

struct FirstConvSrc {
    const float* plane0;  // CIN == 3: (B,1,H,W) first channel (x_t, or the clean target when noise != NULL); else unused
    const float* noise;   // nullable: first channel = coef[b].x*plane0 + coef[b].y*noise   (q_sample fused)
    const float2* coef;   // [B]
    const float* planes;  // (B,2,H,W): the remaining two channels
};

template <int CIN>
__device__ __forceinline__ void fc_load_halo(float* s_x, const FirstConvSrc& src, int img, int h0, int w0, int H, int W,
                                             int tid) {
    float ca = 1.f, cb = 0.f;
    if (CIN == 3 && src.noise != nullptr) {
        const float2 c = src.coef[img];
        ca = c.x;
        cb = c.y;
    }
    for (int i = tid; i < CIN * FC_HT * FC_HT; i += 256) {
        const int ci = i / (FC_HT * FC_HT);
        const int r = i % (FC_HT * FC_HT);
        const int hh = h0 + r / FC_HT - 1, ww = w0 + r % FC_HT - 1;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            if (CIN == 3 && ci == 0) {
                const size_t q = (static_cast<size_t>(img) * H + hh) * W + ww;
                v = src.plane0[q];
                if (src.noise != nullptr) v = ca * v + cb * src.noise[q];
            } else {
                v = src.planes[((static_cast<size_t>(img) * 2 + ci - (CIN - 2)) * H + hh) * W + ww];
            }
        }
        s_x[i] = v;
    }
}

// register-staged variant: the global loads of the NEXT tile are issued before the MMAs of the current one and written
// to the other shared-memory buffer afterwards, so their latency is hidden behind a whole tile of work
template <int CIN>
struct FcHaloRegs {
    static constexpr int N = (CIN * FC_HT * FC_HT + 255) / 256;
    float v[N];
};

template <int CIN>
__device__ __forceinline__ void fc_fetch_halo(FcHaloRegs<CIN>& r, const FirstConvSrc& src, int img, int h0, int w0, int H,
                                              int W, int tid) {
    float ca = 1.f, cb = 0.f;
    if (CIN == 3 && src.noise != nullptr) {
        const float2 c = src.coef[img];
        ca = c.x;
        cb = c.y;
    }
#pragma unroll
    for (int n = 0; n < FcHaloRegs<CIN>::N; ++n) {
        const int i = tid + n * 256;
        float v = 0.f;
        if (i < CIN * FC_HT * FC_HT) {
            const int ci = i / (FC_HT * FC_HT);
            const int rr = i % (FC_HT * FC_HT);
            const int hh = h0 + rr / FC_HT - 1, ww = w0 + rr % FC_HT - 1;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                if (CIN == 3 && ci == 0) {
                    const size_t q = (static_cast<size_t>(img) * H + hh) * W + ww;
                    v = src.plane0[q];
                    if (src.noise != nullptr) v = ca * v + cb * src.noise[q];
                } else {
                    v = src.planes[((static_cast<size_t>(img) * 2 + ci - (CIN - 2)) * H + hh) * W + ww];
                }
            }
        }
        r.v[n] = v;
    }
}

template <int CIN>
__device__ __forceinline__ void fc_store_halo(float* s_x, const FcHaloRegs<CIN>& r, int tid) {
#pragma unroll
    for (int n = 0; n < FcHaloRegs<CIN>::N; ++n) {
        const int i = tid + n * 256;
        if (i < CIN * FC_HT * FC_HT) s_x[i] = r.v[n];
    }
}

// offset of im2col column k inside the halo tile (relative to the pixel's own halo position), or -1 beyond K
template <int CIN>
__device__ __forceinline__ int fc_koff(int k) {
    if (k >= CIN * 9) return -1;
    const int ci = k / 9, t = k % 9;
    return ci * FC_HT * FC_HT + (t / 3) * FC_HT + (t % 3);
}

// ------------------------------------------------------------------------------------------------
// forward. 256 threads = 8 warps; a block iterates over 16x16-pixel tiles, warp w owns tile rows 2w and 2w+1.
// Epilogue: + tb[b][border class][co] (Fast-DDPM time-embedding fold, nullable), * col_scale + col_shift (nullable),
// ReLU, bf16 store through a shared-memory staging tile (16-byte coalesced rows), optional BatchNorm statistics of the
// stored values.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256, 2) first_conv_mma_fwd_kernel(const FirstConvSrc src, const float* __restrict__ wgt,
                                                                    int w_stride,  // floats between output channels
                                                                    const float* __restrict__ tb,
                                                                    const float* __restrict__ col_scale,
                                                                    const float* __restrict__ col_shift, int relu,
                                                                    __nv_bfloat16* __restrict__ out,
                                                                    float* __restrict__ stats, int stats_replicas,
                                                                    int stats_slots, int H, int W, int num_tiles) {
    __shared__ float s_x2[2][CIN * FC_HT * FC_HT];
    __shared__ __align__(16) __nv_bfloat16 s_out[FC_TILE * FC_TILE * FC_PITCH];
    __shared__ uint2 s_blo[2 * 8 * 32];  // low halves of the weight fragments, [ks][nb][lane] (each lane reads its own)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_w = W / FC_TILE;
    const int tiles_hw = tiles_w * (H / FC_TILE);

    // B fragments (weights): high halves resident in registers [ks][nb] -> (b0, b1), low halves in shared memory
    uint32_t bfrag[2][8][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float* wr = wgt + static_cast<size_t>(nb * 8 + g) * w_stride;
            const int k0 = ks * 16 + 2 * t;
            float v[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + (j & 1) + (j >> 1) * 8;
                v[j] = k < CIN * 9 ? wr[k] : 0.f;
                lo[j] = v[j] - bf16_round(v[j]);
            }
            bfrag[ks][nb][0] = pack_bf16x2(v[0], v[1]);
            bfrag[ks][nb][1] = pack_bf16x2(v[2], v[3]);
            if (warp == 0) s_blo[(ks * 8 + nb) * 32 + lane] = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
        }
    // im2col offsets of this thread's 8 K columns
    int koff[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) koff[ks][j] = fc_koff<CIN>(ks * 16 + 2 * t + (j & 1) + (j >> 1) * 8);
    // per-thread epilogue constants: this thread's 16 output channels are nb*8 + 2t + {0,1}
    float st1[8], st2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) st1[k] = st2[k] = 0.f;

    FcHaloRegs<CIN> pre;
    int buf = 0;
    if (static_cast<int>(blockIdx.x) < num_tiles) {
        const int img = blockIdx.x / tiles_hw, t_in = blockIdx.x - img * tiles_hw;
        fc_fetch_halo<CIN>(pre, src, img, (t_in / tiles_w) * FC_TILE, (t_in % tiles_w) * FC_TILE, H, W, tid);
        fc_store_halo<CIN>(s_x2[0], pre, tid);
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_hw;
        const int t_in = tile - img * tiles_hw;
        const int h0 = (t_in / tiles_w) * FC_TILE, w0 = (t_in % tiles_w) * FC_TILE;
        __syncthreads();  // s_x2[buf] is published; the previous tile's s_out copy-out is done
        const int next = tile + gridDim.x;
        if (next < num_tiles) {  // global loads of the next tile fly during this tile's MMAs
            const int nimg = next / tiles_hw, nt = next - nimg * tiles_hw;
            fc_fetch_halo<CIN>(pre, src, nimg, (nt / tiles_w) * FC_TILE, (nt % tiles_w) * FC_TILE, H, W, tid);
        }
        const float* s_x = s_x2[buf];

#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int row = 2 * warp + mb;
            const float* base = s_x + row * FC_HT;
            float acc[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[nb][j] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                float v[4][2];  // [j][pixel column g / g+8]
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = koff[ks][j];
                    v[j][0] = o >= 0 ? base[o + g] : 0.f;
                    v[j][1] = o >= 0 ? base[o + g + 8] : 0.f;
                }
                uint32_t a[4], al[4];
                a[0] = pack_bf16x2(v[0][0], v[1][0]);
                a[1] = pack_bf16x2(v[0][1], v[1][1]);
                a[2] = pack_bf16x2(v[2][0], v[3][0]);
                a[3] = pack_bf16x2(v[2][1], v[3][1]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j][0] -= bf16_round(v[j][0]);
                    v[j][1] -= bf16_round(v[j][1]);
                }
                al[0] = pack_bf16x2(v[0][0], v[1][0]);
                al[1] = pack_bf16x2(v[0][1], v[1][1]);
                al[2] = pack_bf16x2(v[2][0], v[3][0]);
                al[3] = pack_bf16x2(v[2][1], v[3][1]);
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    const uint2 bl = s_blo[(ks * 8 + nb) * 32 + lane];
                    mma_bf16_16816(acc[nb], al, bfrag[ks][nb][0], bfrag[ks][nb][1]);
                    mma_bf16_16816(acc[nb], a, bl.x, bl.y);
                    mma_bf16_16816(acc[nb], a, bfrag[ks][nb][0], bfrag[ks][nb][1]);
                }
            }
            // epilogue -> staging tile
            const int h = h0 + row;
            const int rc = h == 0 ? 0 : (h == H - 1 ? 2 : 1);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int col = g + 8 * half;
                const int w = w0 + col;
                const int cls = rc * 3 + (w == 0 ? 0 : (w == W - 1 ? 2 : 1));
                const float* tbp = tb != nullptr ? tb + (static_cast<size_t>(img) * 9 + cls) * FC_COUT : nullptr;
                __nv_bfloat16* srow = s_out + (row * FC_TILE + col) * FC_PITCH;
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    const int c = nb * 8 + 2 * t;
                    float v0 = acc[nb][2 * half], v1 = acc[nb][2 * half + 1];
                    if (tbp != nullptr) {
                        const float2 b2 = *reinterpret_cast<const float2*>(tbp + c);
                        v0 += b2.x;
                        v1 += b2.y;
                    }
                    if (col_scale != nullptr) {
                        v0 = fmaf(v0, col_scale[c], col_shift[c]);
                        v1 = fmaf(v1, col_scale[c + 1], col_shift[c + 1]);
                    } else if (col_shift != nullptr) {
                        v0 += col_shift[c];
                        v1 += col_shift[c + 1];
                    }
                    if (relu) {
                        v0 = fmaxf(v0, 0.f);
                        v1 = fmaxf(v1, 0.f);
                    }
                    *reinterpret_cast<uint32_t*>(srow + c) = pack_bf16x2(v0, v1);
                }
            }
        }
        __syncthreads();
        // copy-out: 256 pixels x 8 chunks of 16 bytes; chunk index = tid + 256*j => this thread's channel chunk is fixed
        const int c8 = tid & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = (tid >> 3) + 32 * j;
            const uint4 u = *reinterpret_cast<const uint4*>(s_out + p * FC_PITCH + c8 * 8);
            *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * H + h0 + p / FC_TILE) * W + w0 + p % FC_TILE) * FC_COUT +
                                      c8 * 8) = u;
            if (stats != nullptr) {
                const F8 f = unpack8(u);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    st1[k] += f.v[k];
                    st2[k] = fmaf(f.v[k], f.v[k], st2[k]);
                }
            }
        }
        if (next < num_tiles) fc_store_halo<CIN>(s_x2[buf ^ 1], pre, tid);
        buf ^= 1;
    }
    if (stats != nullptr) {
        // lanes with equal (lane & 7) share channels: fold the 4 of a warp with shuffles, the 8 warps through shared
        // memory in a fixed order (deterministic), one global flush per block
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            st1[k] += __shfl_xor_sync(0xffffffffu, st1[k], 8);
            st1[k] += __shfl_xor_sync(0xffffffffu, st1[k], 16);
            st2[k] += __shfl_xor_sync(0xffffffffu, st2[k], 8);
            st2[k] += __shfl_xor_sync(0xffffffffu, st2[k], 16);
        }
        __syncthreads();  // the last copy-out has finished reading the staging tile
        float* s_w = reinterpret_cast<float*>(s_out);  // [8 warps][2][64]
        if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s_w[(warp * 2 + 0) * FC_COUT + lane * 8 + k] = st1[k];
                s_w[(warp * 2 + 1) * FC_COUT + lane * 8 + k] = st2[k];
            }
        }
        __syncthreads();
        if (tid < 2 * FC_COUT) {
            const int which = tid / FC_COUT, c = tid % FC_COUT;
            float acc = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) acc += s_w[(w8 * 2 + which) * FC_COUT + c];
            if (stats_slots > 0) {
                // deterministic: block b STORES into its own slot; slots beyond the grid are zeroed (see conv3x3.cuh)
                stats[static_cast<size_t>(blockIdx.x) * 2 * FC_COUT + tid] = acc;
                for (int s2 = blockIdx.x + gridDim.x; s2 < stats_replicas; s2 += gridDim.x)
                    stats[static_cast<size_t>(s2) * 2 * FC_COUT + tid] = 0.f;
            } else {
                atomicAdd(stats + static_cast<size_t>(blockIdx.x % stats_replicas) * 2 * FC_COUT + tid, acc);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dw[co*w_stride + k] += sum_pixels A[pixel][k] * dZ[pixel][co], k < 9*CIN.
// Warp w accumulates the pixels of tile rows 2w, 2w+1 of every tile the block visits into a 32 x 64 fp32 fragment;
// warps are combined through shared memory in a fixed order at the end. partial == nullptr: one global atomic per weight
// and block into dw (legacy); otherwise block b STORES its [K][64] partial into partial[b] and a fixed-order second stage
// (reduce_partials_kernel) produces dw — bit-reproducible.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256, 2) first_conv_mma_wgrad_kernel(const FirstConvSrc src,
                                                                      const __nv_bfloat16* __restrict__ dz,  // [B][H][W][64]
                                                                      float* __restrict__ dw, int w_stride, int H, int W,
                                                                      int num_tiles, float* __restrict__ partial) {
    __shared__ float s_x[CIN * FC_HT * FC_HT];
    __shared__ __align__(16) __nv_bfloat16 s_dz[FC_TILE * FC_TILE * FC_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles_w = W / FC_TILE;
    const int tiles_hw = tiles_w * (H / FC_TILE);
    int koff[2][2];  // [mb][k = mb*16 + g, + 8]
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        koff[mb][0] = fc_koff<CIN>(mb * 16 + g);
        koff[mb][1] = fc_koff<CIN>(mb * 16 + g + 8);
    }
    float acc[2][8][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mb][nb][j] = 0.f;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_hw;
        const int t_in = tile - img * tiles_hw;
        const int h0 = (t_in / tiles_w) * FC_TILE, w0 = (t_in % tiles_w) * FC_TILE;
        __syncthreads();
        fc_load_halo<CIN>(s_x, src, img, h0, w0, H, W, tid);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = (tid >> 3) + 32 * j, c8 = tid & 7;
            *reinterpret_cast<uint4*>(s_dz + p * FC_PITCH + c8 * 8) = *reinterpret_cast<const uint4*>(
                dz + ((static_cast<size_t>(img) * H + h0 + p / FC_TILE) * W + w0 + p % FC_TILE) * FC_COUT + c8 * 8);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 2; ++r) {  // one k16 step = the 16 pixels of tile row 2*warp + r
            const int row = 2 * warp + r;
            const float* base = s_x + row * FC_HT;
            uint32_t a[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const int o0 = koff[mb][0], o1 = koff[mb][1];
                // A[k][pixel]: a0 = (k=g; pixels 2t,2t+1), a1 = (k=g+8; same), a2 = (k=g; pixels 2t+8,2t+9), a3 = (k=g+8; ..)
                a[mb][0] = o0 >= 0 ? pack_bf16x2(base[o0 + 2 * t], base[o0 + 2 * t + 1]) : 0u;
                a[mb][1] = o1 >= 0 ? pack_bf16x2(base[o1 + 2 * t], base[o1 + 2 * t + 1]) : 0u;
                a[mb][2] = o0 >= 0 ? pack_bf16x2(base[o0 + 2 * t + 8], base[o0 + 2 * t + 9]) : 0u;
                a[mb][3] = o1 >= 0 ? pack_bf16x2(base[o1 + 2 * t + 8], base[o1 + 2 * t + 9]) : 0u;
            }
            // B fragments from the [pixel][co] tile with ldmatrix.trans: lane l addresses row (l & 7) of matrix (l >> 3);
            // matrix mi = pixels (mi & 1)*8 .. +7 of this row, channel block 2*q + (mi >> 1)
            const __nv_bfloat16* prow = s_dz + (row * FC_TILE + (lane & 7) + ((lane >> 3) & 1) * 8) * FC_PITCH + (lane >> 4) * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t b[4];
                ldmatrix_x4_trans(b, prow + q * 16);
                mma_bf16_16816(acc[0][2 * q], a[0], b[0], b[1]);
                mma_bf16_16816(acc[1][2 * q], a[1], b[0], b[1]);
                mma_bf16_16816(acc[0][2 * q + 1], a[0], b[2], b[3]);
                mma_bf16_16816(acc[1][2 * q + 1], a[1], b[2], b[3]);
            }
        }
    }
    // combine the 8 warps in warp order: s_acc[k][co] aliases the dz tile
    __syncthreads();
    float* s_acc = reinterpret_cast<float*>(s_dz);
    for (int i = tid; i < 32 * FC_COUT; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    for (int w8 = 0; w8 < 8; ++w8) {
        if (warp == w8) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    const int c = nb * 8 + 2 * t;
                    s_acc[(mb * 16 + g) * FC_COUT + c] += acc[mb][nb][0];
                    s_acc[(mb * 16 + g) * FC_COUT + c + 1] += acc[mb][nb][1];
                    s_acc[(mb * 16 + g + 8) * FC_COUT + c] += acc[mb][nb][2];
                    s_acc[(mb * 16 + g + 8) * FC_COUT + c + 1] += acc[mb][nb][3];
                }
        }
        __syncthreads();
    }
    if (partial != nullptr) {
        float* dst = partial + static_cast<size_t>(blockIdx.x) * (CIN * 9 * FC_COUT);
        for (int i = tid; i < CIN * 9 * FC_COUT; i += 256) dst[i] = s_acc[i];  // [k][co]
        return;
    }
    for (int i = tid; i < CIN * 9 * FC_COUT; i += 256) {
        const int k = i / FC_COUT, co = i % FC_COUT;
        atomicAdd(dw + static_cast<size_t>(co) * w_stride + k, s_acc[k * FC_COUT + co]);
    }
}

}  // namespace b200sr
