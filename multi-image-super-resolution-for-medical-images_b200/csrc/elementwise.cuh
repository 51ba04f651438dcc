// Bandwidth-bound kernels of the UNet hot path: everything that is not a tensor-core GEMM.
// All activations are NHWC bf16; a tensor may live in a channel slot of a wider buffer, described by
// (base pointer, pixel stride in elements, first channel). 16-byte vector accesses (8 bf16) throughout.
#pragma once
#include "ptx.cuh"

namespace b200sr {

// ------------------------------------------------------------------------------------------------
// vector helpers
// ------------------------------------------------------------------------------------------------
struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 ld_bf16x8(const __nv_bfloat16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    F8 r;
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __low2float(h[i]);
        r.v[2 * i + 1] = __high2float(h[i]);
    }
    return r;
}

__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const F8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]);
    u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]);
    u.w = pack_bf16x2(r.v[6], r.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ F8 ld_f32x8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// weight packing / gradient unpacking (table driven: one launch for the whole network)
// ------------------------------------------------------------------------------------------------
enum PackKind : int {
    PACK_CONV_FWD = 0,     // (Cout,Cin,3,3) f32 -> [Cout][tap*Cin + ci] bf16
    PACK_CONV_DGRAD = 1,   // (Cout,Cin,3,3) f32 -> [Cin][tap'*Cout + co] bf16, tap' = rotated tap
    PACK_CONVT_FWD = 2,    // (Cin,Cout,2,2) f32 -> [(i*2+j)*Cout + co][ci] bf16
    PACK_CONVT_DGRAD = 3,  // (Cin,Cout,2,2) f32 -> [ci][(i*2+j)*Cout + co] bf16
    UNPACK_CONV_WGRAD = 4,   // G[tap][ci][co] f32 -> (Cout,Cin,3,3) f32
    UNPACK_CONVT_WGRAD = 5,  // G[(i,j)][co][ci] f32 -> (Cin,Cout,2,2) f32
    PACK_CONV1X1_FWD = 6,    // (Cout,Cin,1,1) f32 -> [Cout][Cin] bf16
    PACK_CONV1X1_DGRAD = 7,  // (Cout,Cin,1,1) f32 -> [Cin][Cout] bf16
    UNPACK_CONV1X1_WGRAD = 8,  // G[ci][co] f32 -> (Cout,Cin,1,1) f32
    PACK_CONV_BOTH = 9,    // kinds 0 and 1 from ONE read of the parameter tile: dst = forward, aux = dgrad packing
    PACK_CONVT_BOTH = 10,  // kinds 2 and 3 likewise
    // fp32-accuracy eval mode (conv3x3.cuh, SPLIT): each weight as [w_hi | w_hi | w_lo] along K, matching [a_hi | a_lo | a_hi]
    PACK_CONV_FWD_SPLIT3 = 11,   // (Cout,Cin,3,3) f32 -> [Cout][tap*3Cin + part*Cin + ci] bf16
    PACK_CONVT_FWD_SPLIT3 = 12,  // (Cin,Cout,2,2) f32 -> [(i*2+j)*Cout + co][part*Cin + ci] bf16
};

__device__ __forceinline__ __nv_bfloat16 split3_part(float w, int part) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    return part == 2 ? __float2bfloat16_rn(w - __bfloat162float(hi)) : hi;
}

struct PackJob {
    const void* src;
    void* dst;
    int kind;
    int cout;
    int cin;
    int pad;
    long long count;  // elements of dst; kinds 9 / 10: the second destination pointer (dgrad packing)
};

// One 32 x 32 x T tile (T = 9 conv taps or 4 sub-pixels) per block iteration, staged through shared memory so that
// BOTH the global reads and the global writes are coalesced runs (the element-wise version read fp32 weights with a
// 36-byte stride). "outer" is the leading parameter dimension (Cout for Conv2d, Cin for ConvTranspose2d).
constexpr int PK_TILE = 32;

__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJob* __restrict__ jobs) {
    __shared__ float tile[PK_TILE][PK_TILE * 9 + 1];
    const PackJob job = jobs[blockIdx.y];
    const int Cout = job.cout, Cin = job.cin;
    const bool one = job.kind >= PACK_CONV1X1_FWD && job.kind <= UNPACK_CONV1X1_WGRAD;
    const bool conv = one || job.kind == PACK_CONV_FWD || job.kind == PACK_CONV_DGRAD || job.kind == UNPACK_CONV_WGRAD ||
                      job.kind == PACK_CONV_BOTH || job.kind == PACK_CONV_FWD_SPLIT3;
    const int T = one ? 1 : (conv ? 9 : 4);
    const int outer_total = conv ? Cout : Cin, inner_total = conv ? Cin : Cout;
    const int tiles_in = inner_total / PK_TILE;
    const int num_tiles = (outer_total / PK_TILE) * tiles_in;
    const int row = PK_TILE * T;  // floats per outer index inside a tile
    const int tid = threadIdx.x;
    for (int tl = blockIdx.x; tl < num_tiles; tl += gridDim.x) {
        const int o0 = (tl / tiles_in) * PK_TILE, i0 = (tl % tiles_in) * PK_TILE;
        __syncthreads();
        // ---- load ----
        if (job.kind <= PACK_CONVT_DGRAD || job.kind == PACK_CONV1X1_FWD || job.kind == PACK_CONV1X1_DGRAD ||
            job.kind >= PACK_CONV_BOTH) {
            const float* src = static_cast<const float*>(job.src);
            for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                const int o = idx / row, r = idx - o * row;
                tile[o][r] = src[(static_cast<long long>(o0 + o) * inner_total + i0) * T + r];
            }
        } else {
            const float* src = static_cast<const float*>(job.src);  // G[t][inner][outer] (conv) / G[ij][inner][outer]
            for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                const int o = idx % PK_TILE, i = (idx / PK_TILE) % PK_TILE, t = idx / (PK_TILE * PK_TILE);
                tile[o][i * T + t] = src[(static_cast<long long>(t) * inner_total + i0 + i) * outer_total + o0 + o];
            }
        }
        __syncthreads();
        // ---- store ----
        switch (job.kind) {
            case PACK_CONV_FWD: {  // dst[co][t*Cin + ci]
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int i = idx % PK_TILE, t = (idx / PK_TILE) % 9, o = idx / (PK_TILE * 9);
                    dst[static_cast<long long>(o0 + o) * (9LL * Cin) + t * Cin + i0 + i] =
                        __float2bfloat16_rn(tile[o][i * 9 + t]);
                }
                break;
            }
            case PACK_CONV_DGRAD: {  // dst[ci][t*Cout + co] = W[co][ci][8 - t]  (taps rotated by 180 degrees)
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx % PK_TILE, t = (idx / PK_TILE) % 9, i = idx / (PK_TILE * 9);
                    dst[static_cast<long long>(i0 + i) * (9LL * Cout) + t * Cout + o0 + o] =
                        __float2bfloat16_rn(tile[o][i * 9 + 8 - t]);
                }
                break;
            }
            case PACK_CONV_BOTH: {  // both packings of a Conv2d weight from the tile already in shared memory
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                __nv_bfloat16* aux = reinterpret_cast<__nv_bfloat16*>(job.count);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int i = idx % PK_TILE, t = (idx / PK_TILE) % 9, o = idx / (PK_TILE * 9);
                    dst[static_cast<long long>(o0 + o) * (9LL * Cin) + t * Cin + i0 + i] =
                        __float2bfloat16_rn(tile[o][i * 9 + t]);
                }
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx % PK_TILE, t = (idx / PK_TILE) % 9, i = idx / (PK_TILE * 9);
                    aux[static_cast<long long>(i0 + i) * (9LL * Cout) + t * Cout + o0 + o] =
                        __float2bfloat16_rn(tile[o][i * 9 + 8 - t]);
                }
                break;
            }
            case PACK_CONVT_BOTH: {  // both packings of a ConvTranspose2d weight (outer = ci, inner = co)
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                __nv_bfloat16* aux = reinterpret_cast<__nv_bfloat16*>(job.count);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx % PK_TILE, t = (idx / PK_TILE) % 4, i = idx / (PK_TILE * 4);
                    dst[(static_cast<long long>(t) * Cout + i0 + i) * Cin + o0 + o] =
                        __float2bfloat16_rn(tile[o][i * 4 + t]);
                }
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int i = idx % PK_TILE, t = (idx / PK_TILE) % 4, o = idx / (PK_TILE * 4);
                    aux[static_cast<long long>(o0 + o) * (4LL * Cout) + t * Cout + i0 + i] =
                        __float2bfloat16_rn(tile[o][i * 4 + t]);
                }
                break;
            }
            case PACK_CONV_FWD_SPLIT3: {  // dst[co][t*3Cin + part*Cin + ci]
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int i = idx % PK_TILE, t = (idx / PK_TILE) % 9, o = idx / (PK_TILE * 9);
                    const float w = tile[o][i * 9 + t];
#pragma unroll
                    for (int part = 0; part < 3; ++part)
                        dst[static_cast<long long>(o0 + o) * (27LL * Cin) + (t * 3 + part) * Cin + i0 + i] =
                            split3_part(w, part);
                }
                break;
            }
            case PACK_CONVT_FWD_SPLIT3: {  // dst[(t*Cout + co)][part*Cin + ci], outer = ci, inner = co
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx % PK_TILE, t = (idx / PK_TILE) % 4, i = idx / (PK_TILE * 4);
                    const float w = tile[o][i * 4 + t];
#pragma unroll
                    for (int part = 0; part < 3; ++part)
                        dst[(static_cast<long long>(t) * Cout + i0 + i) * (3LL * Cin) + part * Cin + o0 + o] =
                            split3_part(w, part);
                }
                break;
            }
            case PACK_CONV1X1_FWD: {  // dst[co][ci]
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * PK_TILE; idx += 256) {
                    const int i = idx % PK_TILE, o = idx / PK_TILE;
                    dst[static_cast<long long>(o0 + o) * Cin + i0 + i] = __float2bfloat16_rn(tile[o][i]);
                }
                break;
            }
            case PACK_CONV1X1_DGRAD: {  // dst[ci][co]
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * PK_TILE; idx += 256) {
                    const int o = idx % PK_TILE, i = idx / PK_TILE;
                    dst[static_cast<long long>(i0 + i) * Cout + o0 + o] = __float2bfloat16_rn(tile[o][i]);
                }
                break;
            }
            case PACK_CONVT_FWD: {  // dst[ij*Cout + co][ci], outer = ci, inner = co
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx % PK_TILE, t = (idx / PK_TILE) % 4, i = idx / (PK_TILE * 4);
                    dst[(static_cast<long long>(t) * Cout + i0 + i) * Cin + o0 + o] =
                        __float2bfloat16_rn(tile[o][i * 4 + t]);
                }
                break;
            }
            case PACK_CONVT_DGRAD: {  // dst[ci][ij*Cout + co]
                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(job.dst);
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int i = idx % PK_TILE, t = (idx / PK_TILE) % 4, o = idx / (PK_TILE * 4);
                    dst[static_cast<long long>(o0 + o) * (4LL * Cout) + t * Cout + i0 + i] =
                        __float2bfloat16_rn(tile[o][i * 4 + t]);
                }
                break;
            }
            default: {  // UNPACK_*: parameter layout [outer][inner][T], fp32
                // job.pad > 0: the workspace covers a sub-range of the parameter's inner dimension (a conv whose
                // input channels were split over two wgrad launches); pad = inner extent of the destination
                float* dst = static_cast<float*>(job.dst);
                const int inner_dst = job.pad > 0 ? job.pad : inner_total;
                for (int idx = tid; idx < PK_TILE * row; idx += 256) {
                    const int o = idx / row, r = idx - o * row;
                    dst[(static_cast<long long>(o0 + o) * inner_dst + i0) * T + r] = tile[o][r];
                }
                break;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Deterministic cross-block reductions. Every block STORES its partial result into its own slot; the block that
// draws the last ticket sums the slots in slot order. The summation order is a function of the launch geometry
// only (never of the scheduling), so two runs give the same bits; the counter is reset by the last block, so the
// same (zero-initialised) counter serves every later launch on the stream and CUDA-graph replays.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_block_ticket(unsigned* counter, unsigned total) {
    __shared__ int s_last_flag;
    __threadfence();  // this thread's partial stores are visible device-wide before the ticket is drawn
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        s_last_flag = (t == total - 1u) ? 1 : 0;
        if (s_last_flag) *counter = 0u;
    }
    __syncthreads();
    const bool last = s_last_flag != 0;
    if (last) __threadfence();
    return last;
}

// out[map(i)] = sum_{p < nparts} partials[p * part_stride + i], p ascending, i < n.
// map(i) = (i % inner) * stride_mod + (i / inner) * stride_div   (identity: inner = n, stride_mod = 1).
// Block = 32 outputs x 8 part-lanes; lane l sums parts l, l+8, ... then the 8 lanes are combined in lane order.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nparts,
                                                              long long part_stride, int n, float* __restrict__ out,
                                                              int inner, long long stride_mod, long long stride_div,
                                                              float scale) {
    __shared__ float s_p[8][33];
    const int col = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + col;
    float acc = 0.f;
    if (i < n) {
#pragma unroll 8
        for (int p = pl; p < nparts; p += 8) acc += __ldcg(partials + static_cast<long long>(p) * part_stride + i);
    }
    s_p[pl][col] = acc;
    __syncthreads();
    if (pl == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) t += s_p[l][col];
        out[(i % inner) * stride_mod + (i / inner) * stride_div] = t * scale;
    }
}

// In-place first stage of a split-K reduction: parts[0][i] = sum_p parts[p][i] (p ascending) for float4 columns i < n4.
// Fully parallel over the elements AND (for many parts) over `lanes` part-lanes (lane l sums parts l, l+lanes, ...; the
// lanes are then combined in lane order): the summation order is fixed by (nparts, lanes) alone. A thread reads every part
// of its own column before it overwrites part 0, so reducing in place is safe.
__global__ void __launch_bounds__(256) reduce_splits_inplace_kernel(float4* __restrict__ parts, int nparts,
                                                                    long long part_stride4, long long n4, int lanes) {
    __shared__ float4 s_p[256];
    const int cols = 256 / lanes;
    const int col = threadIdx.x % cols, pl = threadIdx.x / cols;
    const long long i = static_cast<long long>(blockIdx.x) * cols + col;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
#pragma unroll 4
        for (int p = pl; p < nparts; p += lanes) {
            const float4 v = __ldcs(parts + static_cast<long long>(p) * part_stride4 + i);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
    }
    if (lanes == 1) {
        if (i < n4) parts[i] = acc;
        return;
    }
    s_p[threadIdx.x] = acc;
    __syncthreads();
    if (pl == 0 && i < n4) {
        float4 t = s_p[col];
        for (int l = 1; l < lanes; ++l) {
            const float4 v = s_p[l * cols + col];
            t.x += v.x;
            t.y += v.y;
            t.z += v.z;
            t.w += v.w;
        }
        parts[i] = t;
    }
}

// Second stage of a split-K weight gradient: sums the per-split partials G_s[t][inner][outer] (s ascending) and writes
// the PyTorch parameter layout dst[outer][inner_off + inner][t] (row length inner_dst) — the fixed-order replacement of
// red.global.add + unpack. T = 9 (Conv2d 3x3: outer = Cout, inner = Cin), 4 (ConvTranspose2d: outer = Cin, inner = Cout)
// or 1 (Conv2d 1x1). One 32 x 32 x T tile per block iteration; reads and writes are coalesced runs.
template <int T>
__global__ void __launch_bounds__(256) wgrad_reduce_unpack_kernel(const float* __restrict__ parts, int nsplits,
                                                                  long long split_stride, int outer_total,
                                                                  int inner_total, int inner_dst, int inner_off,
                                                                  float* __restrict__ dst) {
    // T is a template parameter so that the 4 * T loads of a thread are issued back to back: with a run-time T the
    // per-element predicate kept every load behind the add of the previous one (one DRAM round trip each, 20 us per launch)
    __shared__ float tile[PK_TILE][PK_TILE * T + 1];
    constexpr int ROW = PK_TILE * T;
    constexpr int PER_THREAD = (PK_TILE * ROW) / 256;  // 4 * T
    const int tiles_in = inner_total / PK_TILE;
    const int num_tiles = (outer_total / PK_TILE) * tiles_in;
    const int tid = threadIdx.x;
    for (int tl = blockIdx.x; tl < num_tiles; tl += gridDim.x) {
        const int o0 = (tl / tiles_in) * PK_TILE, i0 = (tl % tiles_in) * PK_TILE;
        float acc[PER_THREAD];
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) acc[j] = 0.f;
        for (int sp = 0; sp < nsplits; ++sp) {
            const float* src = parts + static_cast<long long>(sp) * split_stride;
            float v[PER_THREAD];
#pragma unroll
            for (int j = 0; j < PER_THREAD; ++j) {
                const int idx = tid + 256 * j;
                const int o = idx % PK_TILE, i = (idx / PK_TILE) % PK_TILE, t = idx / (PK_TILE * PK_TILE);
                v[j] = __ldcs(src + (static_cast<long long>(t) * inner_total + i0 + i) * outer_total + o0 + o);
            }
#pragma unroll
            for (int j = 0; j < PER_THREAD; ++j) acc[j] += v[j];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const int idx = tid + 256 * j;
            const int o = idx % PK_TILE, i = (idx / PK_TILE) % PK_TILE, t = idx / (PK_TILE * PK_TILE);
            tile[o][i * T + t] = acc[j];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER_THREAD; ++j) {
            const int idx = tid + 256 * j;
            const int o = idx / ROW, r = idx - o * ROW;
            dst[(static_cast<long long>(o0 + o) * inner_dst + inner_off + i0) * T + r] = tile[o][r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// first layer Conv2d(2 -> 64, 3x3, p1): forward and weight gradient live in firstconv.cuh (tensor-core MMA straight
// from the fp32 NCHW input); only the data gradient below stays on CUDA cores.
// ------------------------------------------------------------------------------------------------
constexpr int C1_TILE = 16;
constexpr int C1_COUT = 64;

// Data gradient of the first layer (needed when the network input itself requires a gradient: stages 2A/2B of the
// Progressive UNet feed stage 1's prediction into their first conv, reference src/ModelLoader.py:258-267).
//   dx[b][ci][h][w] = sum_{kh,kw,co} dZ[b][h-(kh-1)][w-(kw-1)][co] * W[co][ci][kh][kw]        (fp32 NCHW output)
// One thread = one pixel; the 18x18 dZ halo tile (bf16) and the weights live in shared memory.
__global__ void __launch_bounds__(256) conv1_direct_dgrad_kernel(const __nv_bfloat16* __restrict__ dz,  // [B][H][W][64]
                                                                 const float* __restrict__ wgt,          // [64][2][3][3]
                                                                 float* __restrict__ dx,                 // [B][2][H][W]
                                                                 int H, int W, int num_tiles) {
    // rows of 66 bf16 = 33 words: the 32 pixels of a warp read the same channel pair from 32 different banks
    __shared__ __align__(16) __nv_bfloat16 s_dz[(C1_TILE + 2) * (C1_TILE + 2)][C1_COUT + 2];
    __shared__ float2 s_w[9][C1_COUT];  // [tap][co] -> (ci 0, ci 1)
    const int tid = threadIdx.x;
    const int tiles_w = W / C1_TILE;
    const int tiles_hw = tiles_w * (H / C1_TILE);
    for (int i = tid; i < 9 * C1_COUT; i += 256) {
        const int co = i % C1_COUT, t = i / C1_COUT;
        s_w[t][co] = make_float2(wgt[co * 18 + t], wgt[co * 18 + 9 + t]);
    }
    const int ph = tid / C1_TILE, pw = tid % C1_TILE;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_hw;
        const int t_in = tile - img * tiles_hw;
        const int h0 = (t_in / tiles_w) * C1_TILE, w0 = (t_in % tiles_w) * C1_TILE;
        __syncthreads();
        for (int i = tid; i < (C1_TILE + 2) * (C1_TILE + 2) * 8; i += 256) {
            const int p = i >> 3, c8 = i & 7;
            const int hh = h0 + p / (C1_TILE + 2) - 1, ww = w0 + p % (C1_TILE + 2) - 1;
            uint4 u = make_uint4(0, 0, 0, 0);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                u = *reinterpret_cast<const uint4*>(dz + ((static_cast<size_t>(img) * H + hh) * W + ww) * C1_COUT + c8 * 8);
            uint32_t* d = reinterpret_cast<uint32_t*>(&s_dz[p][c8 * 8]);  // 4-byte aligned only (132-byte rows)
            d[0] = u.x;
            d[1] = u.y;
            d[2] = u.z;
            d[3] = u.w;
        }
        __syncthreads();
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            // tap (kh,kw) of the forward conv reads dZ at (h - (kh-1), w - (kw-1)): halo index (ph + 2 - kh, pw + 2 - kw)
            const int kh = t / 3, kw = t % 3;
            const __nv_bfloat16* row = s_dz[(ph + 2 - kh) * (C1_TILE + 2) + (pw + 2 - kw)];
#pragma unroll 8
            for (int co = 0; co < C1_COUT; co += 2) {
                const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(row + co);
                const float g0 = __low2float(g2), g1 = __high2float(g2);
                const float2 wa = s_w[t][co], wb = s_w[t][co + 1];
                a0 = fmaf(g0, wa.x, fmaf(g1, wb.x, a0));
                a1 = fmaf(g0, wa.y, fmaf(g1, wb.y, a1));
            }
        }
        const size_t base = (static_cast<size_t>(img) * 2 * H + h0 + ph) * W + w0 + pw;
        dx[base] = a0;
        dx[base + static_cast<size_t>(H) * W] = a1;
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (train): finalize statistics -> per-channel scale/shift, saved mean/invstd, running stats.
// Reference semantics: nn.BatchNorm2d defaults (eps 1e-5, momentum 0.1, biased var to normalise,
// unbiased var into running_var), unet_model.py:28,31. The conv bias is not added to the stored conv
// output (BN cancels it), so it is added back here for running_mean only.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ stats, int replicas, int C,
                                                          float count, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ conv_bias, float eps, float momentum,
                                                          float* __restrict__ scale, float* __restrict__ shift,
                                                          float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                                          float* __restrict__ running_mean,
                                                          float* __restrict__ running_var,
                                                          long long* __restrict__ num_batches_tracked) {
    // block = 32 channels x 8 slot-lanes; lane l sums slots l, l+8, ... (double), lanes are combined in lane order:
    // the result depends on the slot contents only, not on which CTA wrote them when
    __shared__ double s_s[8][33], s_q[8][33];
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + col;
    double s = 0.0, q = 0.0;
    if (c < C) {
#pragma unroll 8
        for (int r = sl; r < replicas; r += 8) {
            s += stats[(static_cast<size_t>(r) * 2) * C + c];
            q += stats[(static_cast<size_t>(r) * 2 + 1) * C + c];
        }
    }
    s_s[sl][col] = s;
    s_q[sl][col] = q;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
    if (sl != 0 || c >= C) return;
    s = q = 0.0;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        s += s_s[l][col];
        q += s_q[l][col];
    }
    const double inv_count = 1.0 / static_cast<double>(count);
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = 1.0f / sqrtf(static_cast<float>(var) + eps);
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - static_cast<float>(mean) * sc;
    mean_out[c] = static_cast<float>(mean);
    invstd_out[c] = invstd;
    if (running_mean != nullptr) {
        const float b = conv_bias ? conv_bias[c] : 0.f;
        const float unbiased = static_cast<float>(var * (count / (count - 1.0)));
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (static_cast<float>(mean) + b);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// Eval-mode fold: y = relu(scale*conv + shift) with scale = gamma/sqrt(rv+eps),
// shift = beta + (bias - rm)*scale. Tiny; one launch per network via the job table.
struct FoldJob {
    const float* gamma;
    const float* beta;
    const float* rmean;
    const float* rvar;
    const float* conv_bias;
    float* scale;
    float* shift;
    int C;
    int pad;
};
__global__ void bn_fold_eval_kernel(const FoldJob* __restrict__ jobs, float eps) {
    const FoldJob j = jobs[blockIdx.y];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.C; c += gridDim.x * blockDim.x) {
        const float sc = j.gamma[c] * rsqrtf(j.rvar[c] + eps);
        j.scale[c] = sc;
        j.shift[c] = j.beta[c] + ((j.conv_bias ? j.conv_bias[c] : 0.f) - j.rmean[c]) * sc;
    }
}

// ------------------------------------------------------------------------------------------------
// BN-apply + ReLU (+ 2x2 max-pool): reads the raw conv output once, writes the activation into its
// (possibly concat-slot) destination and, for encoder blocks, the pooled tensor in the same pass.
// One thread = 8 channels of a 2x2 pixel quad.
// ------------------------------------------------------------------------------------------------
template <bool EDGE>  // EDGE: H or W odd (never with pooling) — the last 2x2 group of a row / column is partial
__global__ void __launch_bounds__(256) bnrelu_apply_kernel(const __nv_bfloat16* __restrict__ z, int C,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           __nv_bfloat16* __restrict__ act, int act_stride,
                                                           int act_coff, __nv_bfloat16* __restrict__ pooled, int H,
                                                           int W, long long total /* B*ceil(H/2)*ceil(W/2)*(C/8) */) {
    griddep_launch_dependents();  // PDL: a conv kernel launched next may start its set-up while this grid drains
    const int c8n = C >> 3;
    const int W2 = (W + 1) >> 1, H2 = (H + 1) >> 1;  // odd sizes (no pooling then): the last group is partial
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        long long r = idx / c8n;
        const int w2 = static_cast<int>(r % W2);
        r /= W2;
        const int h2 = static_cast<int>(r % H2);
        const int img = static_cast<int>(r / H2);
        const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c);
        F8 mx;
#pragma unroll
        for (int k = 0; k < 8; ++k) mx.v[k] = 0.f;  // post-ReLU values are >= 0
        // no branch between the four loads (they must issue back to back): a pixel past an odd edge re-reads the last row /
        // column instead and is simply not stored
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int hh = 2 * h2 + dy, ww = 2 * w2 + dx;
                const bool inside = !EDGE || (hh < H && ww < W);
                const size_t pix = EDGE ? (static_cast<size_t>(img) * H + min(hh, H - 1)) * W + min(ww, W - 1)
                                        : (static_cast<size_t>(img) * H + hh) * W + ww;
                F8 v = ld_bf16x8(z + pix * C + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v.v[k] = fmaxf(fmaf(v.v[k], sc.v[k], sh.v[k]), 0.f);
                    // pool over the values as stored (bf16), like MaxPool2d reading the activation tensor
                    mx.v[k] = fmaxf(mx.v[k], bf16_round(v.v[k]));
                }
                if (inside) st_bf16x8(act + pix * act_stride + act_coff + c, v);
            }
        if (pooled != nullptr) st_bf16x8(pooled + ((static_cast<size_t>(img) * H2 + h2) * W2 + w2) * C + c, mx);
    }
}

// Train-mode BatchNorm finalize fused into the apply pass (saves one tiny launch per layer on the critical path):
// every thread derives scale/shift of its own 8 channels from the statistic replicas (double precision for
// E[x^2]-E[x]^2, exactly like bn_finalize_kernel); block 0 additionally publishes scale/shift/mean/invstd for the
// backward pass and updates the running statistics (unbiased variance, conv bias re-added to the mean).
__global__ void __launch_bounds__(256, 4) bn_train_apply_kernel(
    const __nv_bfloat16* __restrict__ z, int C, const float* __restrict__ stats, int replicas, float count,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps,
    float momentum, float* __restrict__ scale_out, float* __restrict__ shift_out, float* __restrict__ mean_out,
    float* __restrict__ invstd_out, float* __restrict__ running_mean, float* __restrict__ running_var,
    __nv_bfloat16* __restrict__ act, int act_stride, int act_coff, __nv_bfloat16* __restrict__ pooled, int H, int W,
    long long total /* B*(H/2)*(W/2)*(C/8) */) {
    const int c8n = C >> 3;
    const int W2 = W >> 1, H2 = H >> 1;
    const long long first = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    // grid-stride by a multiple of c8n keeps a thread on the same 8 channels (host guarantees stride % c8n == 0)
    const int c = static_cast<int>(first % c8n) * 8;
    F8 sc, sh;
    {
        double s[8], q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.0;
        for (int r = 0; r < replicas; ++r) {
            const F8 a = ld_f32x8(stats + (static_cast<size_t>(r) * 2) * C + c);
            const F8 b = ld_f32x8(stats + (static_cast<size_t>(r) * 2 + 1) * C + c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s[k] += a.v[k];
                q[k] += b.v[k];
            }
        }
        const F8 ga = ld_f32x8(gamma + c), be = ld_f32x8(beta + c);
        const bool publish = first < c8n;  // the first c8n threads of the grid cover every channel once
        const double inv_count = 1.0 / static_cast<double>(count);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // E[x^2] - E[x]^2 in double (cancellation), the rest in fp32: same formula as bn_finalize_kernel
            const double mean = s[k] * inv_count;
            double var = q[k] * inv_count - mean * mean;
            if (var < 0.0) var = 0.0;
            const float invstd = 1.0f / sqrtf(static_cast<float>(var) + eps);
            sc.v[k] = ga.v[k] * invstd;
            sh.v[k] = be.v[k] - static_cast<float>(mean) * sc.v[k];
            if (publish) {
                scale_out[c + k] = sc.v[k];
                shift_out[c + k] = sh.v[k];
                mean_out[c + k] = static_cast<float>(mean);
                invstd_out[c + k] = invstd;
                if (running_mean != nullptr) {
                    const float b = conv_bias ? conv_bias[c + k] : 0.f;
                    const float unbiased = static_cast<float>(var * (count / (count - 1.0)));
                    running_mean[c + k] = (1.f - momentum) * running_mean[c + k] + momentum * (static_cast<float>(mean) + b);
                    running_var[c + k] = (1.f - momentum) * running_var[c + k] + momentum * unbiased;
                }
            }
        }
    }
    for (long long idx = first; idx < total; idx += stride) {
        long long r = idx / c8n;
        const int w2 = static_cast<int>(r % W2);
        r /= W2;
        const int h2 = static_cast<int>(r % H2);
        const int img = static_cast<int>(r / H2);
        F8 mx;
#pragma unroll
        for (int k = 0; k < 8; ++k) mx.v[k] = 0.f;  // post-ReLU values are >= 0
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const size_t pix = (static_cast<size_t>(img) * H + 2 * h2 + dy) * W + 2 * w2 + dx;
                F8 v = ld_bf16x8(z + pix * C + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v.v[k] = fmaxf(fmaf(v.v[k], sc.v[k], sh.v[k]), 0.f);
                    mx.v[k] = fmaxf(mx.v[k], bf16_round(v.v[k]));
                }
                st_bf16x8(act + pix * act_stride + act_coff + c, v);
            }
        if (pooled != nullptr) st_bf16x8(pooled + ((static_cast<size_t>(img) * H2 + h2) * W2 + w2) * C + c, mx);
    }
}

// Plain MaxPool2d(2,2) forward (unet_model.py:52,55,58,61) on an NHWC slot -> dense NHWC.
__global__ void __launch_bounds__(256) maxpool2x2_fwd_kernel(const __nv_bfloat16* __restrict__ in, int in_stride,
                                                             int in_coff, int C, __nv_bfloat16* __restrict__ out,
                                                             int H, int W, long long total) {
    const int c8n = C >> 3;
    const int W2 = W >> 1, H2 = H >> 1;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        long long r = idx / c8n;
        const int w2 = static_cast<int>(r % W2);
        r /= W2;
        const int h2 = static_cast<int>(r % H2);
        const int img = static_cast<int>(r / H2);
        const size_t p00 = (static_cast<size_t>(img) * H + 2 * h2) * W + 2 * w2;
        const F8 a = ld_bf16x8(in + p00 * in_stride + in_coff + c);
        const F8 b = ld_bf16x8(in + (p00 + 1) * in_stride + in_coff + c);
        const F8 d = ld_bf16x8(in + (p00 + W) * in_stride + in_coff + c);
        const F8 e = ld_bf16x8(in + (p00 + W + 1) * in_stride + in_coff + c);
        F8 m;
#pragma unroll
        for (int k = 0; k < 8; ++k) m.v[k] = fmaxf(fmaxf(a.v[k], b.v[k]), fmaxf(d.v[k], e.v[k]));
        st_bf16x8(out + ((static_cast<size_t>(img) * H2 + h2) * W2 + w2) * C + c, m);
    }
}

// MaxPool2d(2,2) backward fused with the skip-connection gradient add:
//   dY[q] = dskip[q] + (q is the FIRST max of its 2x2 window in row-major order ? dpool[window] : 0)
// (ATen max_pool2d_with_indices tie-breaking). `act` is the forward activation the pool read.
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ act, int act_stride,
                                                             int act_coff, const __nv_bfloat16* __restrict__ dpool,
                                                             const __nv_bfloat16* __restrict__ dskip,
                                                             int dskip_stride, int dskip_coff, int C,
                                                             __nv_bfloat16* __restrict__ dy, int H, int W,
                                                             long long total) {
    griddep_launch_dependents();  // PDL: a conv kernel launched next may start its set-up while this grid drains
    const int c8n = C >> 3;
    const int W2 = W >> 1, H2 = H >> 1;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        long long r = idx / c8n;
        const int w2 = static_cast<int>(r % W2);
        r /= W2;
        const int h2 = static_cast<int>(r % H2);
        const int img = static_cast<int>(r / H2);
        const size_t p00 = (static_cast<size_t>(img) * H + 2 * h2) * W + 2 * w2;
        const size_t pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        F8 a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = ld_bf16x8(act + pix[k] * act_stride + act_coff + c);
        const F8 g = ld_bf16x8(dpool + ((static_cast<size_t>(img) * H2 + h2) * W2 + w2) * C + c);
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int best = 0;
            float bv = a[0].v[k];
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (a[j].v[k] > bv) {
                    bv = a[j].v[k];
                    best = j;
                }
            arg[k] = best;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            F8 o;
            if (dskip != nullptr) {
                o = ld_bf16x8(dskip + pix[j] * dskip_stride + dskip_coff + c);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] += (arg[k] == j) ? g.v[k] : 0.f;
            st_bf16x8(dy + pix[j] * C + c, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward, two passes over (dY, z):
//   m = [scale*z + shift > 0];  g = dY*m;  xhat = (z - mean)*invstd
//   pass 1: S1[c] = sum g, S2[c] = sum g*xhat            (-> dbeta, dgamma)
//   pass 2: dZ = gamma*invstd * (g - S1/N - xhat*S2/N)
// Pass 1 layout: a block owns a channel group (8 channels per thread-column) and a slice of pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                                            int dy_coff, const __nv_bfloat16* __restrict__ z, int C,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            float* __restrict__ sums /* [replicas][2][C] */,
                                                            int replicas, long long npix) {
    // thread layout: tx = channel-vector (8 ch) within a 64-channel group, ty = pixel lane
    const int cg = C >> 6;                       // 64-channel groups
    const int group = blockIdx.x % cg;           // which 64-channel group
    const int slice = blockIdx.x / cg;           // which pixel slice
    const int nslices = gridDim.x / cg;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;  // 8 x 32
    const int c = group * 64 + tx * 8;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c), is = ld_f32x8(invstd + c);
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
    for (long long p = static_cast<long long>(slice) * 32 + ty; p < npix; p += static_cast<long long>(nslices) * 32) {
        const F8 g = ld_bf16x8(dy + p * dy_stride + dy_coff + c);
        const F8 zz = ld_bf16x8(z + p * C + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float y = fmaf(zz.v[k], sc.v[k], sh.v[k]);
            const float gm = y > 0.f ? g.v[k] : 0.f;
            s1[k] += gm;
            s2[k] = fmaf(gm, (zz.v[k] - mu.v[k]) * is.v[k], s2[k]);
        }
    }
    __shared__ float red[2][32][65];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        red[0][ty][tx * 8 + k] = s1[k];
        red[1][ty][tx * 8 + k] = s2[k];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
        float acc = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) acc += red[which][r][ch];
        atomicAdd(sums + (static_cast<size_t>(slice % replicas) * 2 + which) * C + group * 64 + ch, acc);
    }
}

// Bandwidth-tuned variants for C in {64,...,2048} with 256 % (C/8) == 0: a thread owns 8 fixed channels (all
// per-channel constants live in registers) and walks pixels with 4 independent 16-byte loads in flight per tensor.
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ F8 unpack8(const uint4& u) {
    F8 r;
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __low2float(h[i]);
        r.v[2 * i + 1] = __high2float(h[i]);
    }
    return r;
}

constexpr int BNB_UNROLL = 4;

template <bool MASKED>
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_fast_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                                                 int dy_coff, const __nv_bfloat16* __restrict__ z,
                                                                 int C, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd,
                                                                 float* __restrict__ sums, int replicas,
                                                                 long long npix,
                                                                 const __nv_bfloat16* __restrict__ mask_src) {
    // mask_src (nullable, dense [npix][C]): ReLU mask taken from a stored activation (residual blocks, where the
    // ReLU follows the skip addition) instead of from scale*z+shift > 0
    const int CV = C >> 3;            // channel vectors per pixel
    const int PB = 256 / CV;          // pixels per block iteration
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int c = cv * 8;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c);
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
    const long long step = static_cast<long long>(gridDim.x) * PB;
    for (long long p0 = static_cast<long long>(blockIdx.x) * PB + pl; p0 < npix; p0 += step * BNB_UNROLL) {
        uint4 g[BNB_UNROLL], zz[BNB_UNROLL], mm[MASKED ? BNB_UNROLL : 1];
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                g[u] = ld_stream(dy + p * dy_stride + dy_coff + c);
                zz[u] = ld_stream(z + p * C + c);
                if (MASKED) mm[MASKED ? u : 0] = ld_stream(mask_src + p * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            if (p0 + u * step < npix) {
                const F8 gf = unpack8(g[u]), zf = unpack8(zz[u]);
                F8 mf;
                if (MASKED) {
                    mf = unpack8(mm[MASKED ? u : 0]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mf.v[k] = fmaf(zf.v[k], sc.v[k], sh.v[k]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gm = mf.v[k] > 0.f ? gf.v[k] : 0.f;
                    s1[k] += gm;
                    s2[k] = fmaf(gm, zf.v[k] - mu.v[k], s2[k]);
                }
            }
        }
    }
    // reduce over the PB pixel lanes of the block that share a channel vector
    __shared__ float red[2][256][9];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        red[0][threadIdx.x][k] = s1[k];
        red[1][threadIdx.x][k] = s2[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        const int which = i / C, ch = i % C;
        float acc = 0.f;
        for (int r = 0; r < PB; ++r) acc += red[which][r * CV + (ch >> 3)][ch & 7];
        if (which == 1) acc *= invstd[ch];
        atomicAdd(sums + (static_cast<size_t>(blockIdx.x % replicas) * 2 + which) * C + ch, acc);
    }
}

// Block- and grid-level finish shared by the deterministic BatchNorm-backward reductions: thread (tx = tid & 7, ty = tid >> 3)
// holds partial sums of channels blockIdx.y*64 + tx*8 .. +7 over its pixels. Pixel lanes are combined in lane order, the
// block's 128 sums go to its own slot, and the last block of the channel group (ticket) adds the slots in slice order.
__device__ __forceinline__ void bn_red_block_finish(const float (&s1)[8], const float (&s2)[8], float* __restrict__ sums,
                                                    float* __restrict__ ws, unsigned* __restrict__ counters,
                                                    const float* __restrict__ invstd, int C) {
    const int slices = gridDim.x, slice = blockIdx.x, group = blockIdx.y;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    __shared__ float red[2][32][65];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        red[0][ty][tx * 8 + k] = s1[k];
        red[1][ty][tx * 8 + k] = s2[k];
    }
    __syncthreads();
    float* slot = ws + (static_cast<size_t>(group) * slices + slice) * 128;
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
        float acc = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) acc += red[which][r][ch];
        slot[threadIdx.x] = acc;
    }
    if (!last_block_ticket(counters + group, static_cast<unsigned>(slices))) return;
    // last block of this channel group: 32 float4 columns x 8 slice-lanes, lanes combined in lane order
    __shared__ float4 s_t[8][32];
    const int col4 = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const float4* base = reinterpret_cast<const float4*>(ws + static_cast<size_t>(group) * slices * 128) + col4;
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = sl; r < slices; r += 8) {
        const float4 v = __ldcg(base + static_cast<size_t>(r) * 32);
        a4.x += v.x;
        a4.y += v.y;
        a4.z += v.z;
        a4.w += v.w;
    }
    s_t[sl][col4] = a4;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) acc += reinterpret_cast<const float*>(&s_t[l][0])[threadIdx.x];
        if (which == 1) acc *= invstd[group * 64 + ch];
        sums[static_cast<size_t>(which) * C + group * 64 + ch] = acc;
    }
}

// Deterministic BatchNorm+ReLU backward reduction (pass 1). grid = (slices, C/64): a block owns 64 channels (8 threads
// x 8 channels per pixel) and every `slices`-th group of 32 pixels; its 128 partial sums go to its own slot of `ws`; the
// last block of a channel group (ticket counter per group) sums that group's slots in slice order and writes the final
//   sums[0][c] = sum g,  sums[1][c] = invstd[c] * sum g*(z - mean)        (g = dY * [scale*z + shift > 0], or the stored mask)
// No atomics on data, no pre-zeroing: bit-reproducible. counters: C/64 zero-initialised unsigned, self-resetting.
template <bool MASKED>
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_det_kernel(
    const __nv_bfloat16* __restrict__ dy, int dy_stride, int dy_coff, const __nv_bfloat16* __restrict__ z, int C,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, float* __restrict__ sums /* [2][C] */, float* __restrict__ ws,
    unsigned* __restrict__ counters, long long npix, const __nv_bfloat16* __restrict__ mask_src) {
    const int slices = gridDim.x, slice = blockIdx.x, group = blockIdx.y;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;  // 8 channel vectors x 32 pixel lanes
    const int c = group * 64 + tx * 8;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c);
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
    const long long step = static_cast<long long>(slices) * 32;
    for (long long p0 = static_cast<long long>(slice) * 32 + ty; p0 < npix; p0 += step * BNB_UNROLL) {
        uint4 g[BNB_UNROLL], zz[BNB_UNROLL], mm[MASKED ? BNB_UNROLL : 1];
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                g[u] = ld_stream(dy + p * dy_stride + dy_coff + c);
                zz[u] = ld_stream(z + p * C + c);
                if (MASKED) mm[MASKED ? u : 0] = ld_stream(mask_src + p * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            if (p0 + u * step < npix) {
                const F8 gf = unpack8(g[u]), zf = unpack8(zz[u]);
                F8 mf;
                if (MASKED) {
                    mf = unpack8(mm[MASKED ? u : 0]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mf.v[k] = fmaf(zf.v[k], sc.v[k], sh.v[k]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gm = mf.v[k] > 0.f ? gf.v[k] : 0.f;
                    s1[k] += gm;
                    s2[k] = fmaf(gm, zf.v[k] - mu.v[k], s2[k]);
                }
            }
        }
    }
    bn_red_block_finish(s1, s2, sums, ws, counters, invstd, C);
}

// MaxPool2d(2,2) backward + skip-gradient add (maxpool2x2_bwd_kernel) that ALSO produces the BatchNorm+ReLU backward
// reduction of the layer the gradient flows into (encoder conv.3 -> BN -> ReLU -> {skip, pool}): the dY it writes is the very
// tensor bn_bwd_reduce_det would read back, so the separate reduction pass (and its re-read of dY) disappears; the sums are
// taken from the values as stored (bf16), exactly like the two-kernel path. grid = (slices, C/64); a block owns 64
// channels and every slices-th group of 32 2x2 quads.
__global__ void __launch_bounds__(256, 2) maxpool2x2_bwd_bnred_kernel(
    const __nv_bfloat16* __restrict__ act, int act_stride, int act_coff, const __nv_bfloat16* __restrict__ dpool,
    const __nv_bfloat16* __restrict__ dskip, int dskip_stride, int dskip_coff, int C, __nv_bfloat16* __restrict__ dy,
    const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ sums, float* __restrict__ ws,
    unsigned* __restrict__ counters, int H, int W, long long nquads) {
    griddep_launch_dependents();  // PDL: a conv kernel launched next may start its set-up while this grid drains
    const int slices = gridDim.x, slice = blockIdx.x, group = blockIdx.y;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int c = group * 64 + tx * 8;
    const int W2 = W >> 1, H2 = H >> 1;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c);
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
    for (long long q = static_cast<long long>(slice) * 32 + ty; q < nquads; q += static_cast<long long>(slices) * 32) {
        const int w2 = static_cast<int>(q % W2);
        long long r = q / W2;
        const int h2 = static_cast<int>(r % H2);
        const long long img = r / H2;
        const long long p00 = (img * H + 2 * h2) * W + 2 * w2;
        const long long pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        uint4 ua[4], ud[4], uz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ua[j] = ld_stream(act + pix[j] * act_stride + act_coff + c);
            ud[j] = ld_stream(dskip + pix[j] * dskip_stride + dskip_coff + c);
            uz[j] = ld_stream(z + pix[j] * C + c);
        }
        const F8 g = unpack8(ld_stream(dpool + ((img * H2 + h2) * W2 + w2) * C + c));
        F8 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = unpack8(ua[j]);
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int best = 0;
            float bv = a[0].v[k];
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (a[j].v[k] > bv) {
                    bv = a[j].v[k];
                    best = j;
                }
            arg[k] = best;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            F8 o = unpack8(ud[j]);
            const F8 zf = unpack8(uz[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] += (arg[k] == j) ? g.v[k] : 0.f;
            st_bf16x8(dy + pix[j] * C + c, o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float gv = bf16_round(o.v[k]);  // the reduction sees dY as stored
                const float gm = fmaf(zf.v[k], sc.v[k], sh.v[k]) > 0.f ? gv : 0.f;
                s1[k] += gm;
                s2[k] = fmaf(gm, zf.v[k] - mu.v[k], s2[k]);
            }
        }
    }
    bn_red_block_finish(s1, s2, sums, ws, counters, invstd, C);
}

__global__ void __launch_bounds__(256, 3) bn_bwd_apply_fast_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                                                int dy_coff, const __nv_bfloat16* __restrict__ z, int C,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const float* __restrict__ c1,
                                                                const float* __restrict__ c2,
                                                                __nv_bfloat16* __restrict__ dz, long long npix) {
    const int CV = C >> 3;
    const int PB = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int c = cv * 8;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c);
    F8 ka, kb;  // dz = gm*scale + z*ka + kb
    {
        const F8 mu = ld_f32x8(mean + c), is = ld_f32x8(invstd + c), k1 = ld_f32x8(c1 + c), k2 = ld_f32x8(c2 + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float t = is.v[k] * k2.v[k] * sc.v[k];
            ka.v[k] = -t;
            kb.v[k] = mu.v[k] * t - sc.v[k] * k1.v[k];
        }
    }
    const long long step = static_cast<long long>(gridDim.x) * PB;
    for (long long p0 = static_cast<long long>(blockIdx.x) * PB + pl; p0 < npix; p0 += step * BNB_UNROLL) {
        uint4 g[BNB_UNROLL], zz[BNB_UNROLL];
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                g[u] = ld_stream(dy + p * dy_stride + dy_coff + c);
                zz[u] = ld_stream(z + p * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                const F8 gf = unpack8(g[u]), zf = unpack8(zz[u]);
                F8 o;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gm = fmaf(zf.v[k], sc.v[k], sh.v[k]) > 0.f ? gf.v[k] : 0.f;
                    o.v[k] = fmaf(gm, sc.v[k], fmaf(zf.v[k], ka.v[k], kb.v[k]));
                }
                st_bf16x8(dz + p * C + c, o);
            }
        }
    }
}

// BatchNorm backward apply with the finalize step folded in: a thread derives c1 = S1/N, c2 = S2/N of its own 8
// channels from the replicas of the reduce pass; the first pixel lane of block 0 publishes dgamma / dbeta.
template <bool MASKED>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_fused_kernel(
    const __nv_bfloat16* __restrict__ dy, int dy_stride, int dy_coff, const __nv_bfloat16* __restrict__ z, int C,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ sums, int replicas, float count,
    float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dz, long long npix,
    const __nv_bfloat16* __restrict__ mask_src) {
    griddep_launch_dependents();  // PDL: a conv kernel launched next may start its set-up while this grid drains
    const int CV = C >> 3;
    const int PB = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int c = cv * 8;
    const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c);
    F8 ka, kb;  // dz = gm*scale + z*ka + kb
    {
        F8 s1, s2;
#pragma unroll
        for (int k = 0; k < 8; ++k) s1.v[k] = s2.v[k] = 0.f;
        for (int r = 0; r < replicas; ++r) {
            const F8 a = ld_f32x8(sums + (static_cast<size_t>(r) * 2) * C + c);
            const F8 b = ld_f32x8(sums + (static_cast<size_t>(r) * 2 + 1) * C + c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s1.v[k] += a.v[k];
                s2.v[k] += b.v[k];
            }
        }
        if (blockIdx.x == 0 && pl == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                dgamma[c + k] = s2.v[k];
                dbeta[c + k] = s1.v[k];
            }
        }
        const F8 mu = ld_f32x8(mean + c), is = ld_f32x8(invstd + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float k1 = s1.v[k] / count, k2 = s2.v[k] / count;
            const float t = is.v[k] * k2 * sc.v[k];
            ka.v[k] = -t;
            kb.v[k] = mu.v[k] * t - sc.v[k] * k1;
        }
    }
    const long long step = static_cast<long long>(gridDim.x) * PB;
    for (long long p0 = static_cast<long long>(blockIdx.x) * PB + pl; p0 < npix; p0 += step * BNB_UNROLL) {
        uint4 g[BNB_UNROLL], zz[BNB_UNROLL], mm[MASKED ? BNB_UNROLL : 1];
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                g[u] = ld_stream(dy + p * dy_stride + dy_coff + c);
                zz[u] = ld_stream(z + p * C + c);
                if (MASKED) mm[MASKED ? u : 0] = ld_stream(mask_src + p * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < BNB_UNROLL; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                const F8 gf = unpack8(g[u]), zf = unpack8(zz[u]);
                F8 mf;
                if (MASKED) {
                    mf = unpack8(mm[MASKED ? u : 0]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mf.v[k] = fmaf(zf.v[k], sc.v[k], sh.v[k]);
                }
                F8 o;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gm = mf.v[k] > 0.f ? gf.v[k] : 0.f;
                    o.v[k] = fmaf(gm, sc.v[k], fmaf(zf.v[k], ka.v[k], kb.v[k]));
                }
                st_bf16x8(dz + p * C + c, o);
            }
        }
    }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ sums, int replicas, int C, float count,
                                       float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < replicas; ++r) {
        s1 += sums[(static_cast<size_t>(r) * 2) * C + c];
        s2 += sums[(static_cast<size_t>(r) * 2 + 1) * C + c];
    }
    c1[c] = s1 / count;
    c2[c] = s2 / count;
    dgamma[c] = s2;
    dbeta[c] = s1;
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                                           int dy_coff, const __nv_bfloat16* __restrict__ z, int C,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ c1, const float* __restrict__ c2,
                                                           __nv_bfloat16* __restrict__ dz, long long total) {
    const int c8n = C >> 3;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        const long long p = idx / c8n;
        const F8 sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c), is = ld_f32x8(invstd + c);
        const F8 k1 = ld_f32x8(c1 + c), k2 = ld_f32x8(c2 + c);
        const F8 g = ld_bf16x8(dy + p * dy_stride + dy_coff + c);
        const F8 zz = ld_bf16x8(z + p * C + c);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float y = fmaf(zz.v[k], sc.v[k], sh.v[k]);
            const float gm = y > 0.f ? g.v[k] : 0.f;
            const float xh = (zz.v[k] - mu.v[k]) * is.v[k];
            o.v[k] = sc.v[k] * (gm - k1.v[k] - xh * k2.v[k]);  // scale = gamma*invstd
        }
        st_bf16x8(dz + p * C + c, o);
    }
}

// ------------------------------------------------------------------------------------------------
// Perceptual (VGG feature) loss helpers — SURVEY §8(f) row 2. Frozen conv+bias+ReLU stacks have no BatchNorm, so
// their backward only needs the ReLU mask of the stored activation:  out = dy * [act > 0]   (bf16, 8 per thread)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                       const __nv_bfloat16* __restrict__ act,
                                                       __nv_bfloat16* __restrict__ out, long long n8) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const F8 g = unpack8(ld_stream(dy + i * 8)), a = unpack8(ld_stream(act + i * 8));
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = a.v[k] > 0.f ? g.v[k] : 0.f;
        st_bf16x8(out + i * 8, o);
    }
}

// Feature-space MSE between two post-ReLU feature maps and its gradient w.r.t. the first one, already multiplied by
// that layer's ReLU mask:  sums[0] += sum (fp - ft)^2 ;  g = gscale * (fp - ft) * [fp > 0]
__global__ void __launch_bounds__(256) feat_mse_grad_kernel(const __nv_bfloat16* __restrict__ fp,
                                                            const __nv_bfloat16* __restrict__ ft,
                                                            __nv_bfloat16* __restrict__ g, double* __restrict__ sums,
                                                            float gscale, long long n8) {
    float part = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const F8 a = unpack8(ld_stream(fp + i * 8)), b = unpack8(ld_stream(ft + i * 8));
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float d = a.v[k] - b.v[k];
            part = fmaf(d, d, part);
            o.v[k] = a.v[k] > 0.f ? gscale * d : 0.f;
        }
        if (g != nullptr) st_bf16x8(g + i * 8, o);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ float s_red[8];
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += s_red[w];
        atomicAdd(sums, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// 1x1 head: Conv2d(64 -> 1) + bias (unet_model.py:80,117). fp32 output in NCHW (C=1 => same as NHW).
// 8 threads per pixel (one uint4 each), shuffle reduce.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ act,  // [P][64]
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float* __restrict__ out, long long npix) {
    const int sub = threadIdx.x & 7;
    const F8 wv = ld_f32x8(w + sub * 8);
    const float bias = b[0];
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npix;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const F8 a = ld_bf16x8(act + p * 64 + sub * 8);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s = fmaf(a.v[k], wv.v[k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (sub == 0) out[p] = s + bias;
    }
}

// head backward: dA[p][c] = dOut[p]*w[c];  dW[c] = sum_p dOut[p]*a[p][c];  db = sum_p dOut[p]
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dout,
                                                       const __nv_bfloat16* __restrict__ act,
                                                       const float* __restrict__ w, __nv_bfloat16* __restrict__ dact,
                                                       float* __restrict__ dw, float* __restrict__ db, long long npix) {
    const int sub = threadIdx.x & 7;
    const F8 wv = ld_f32x8(w + sub * 8);
    float accw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[k] = 0.f;
    float accb = 0.f;
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npix;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const float g = dout[p];
        const F8 a = ld_bf16x8(act + p * 64 + sub * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o.v[k] = g * wv.v[k];
            accw[k] = fmaf(g, a.v[k], accw[k]);
        }
        st_bf16x8(dact + p * 64 + sub * 8, o);
        if (sub == 0) accb += g;
    }
    __shared__ float s_w[64];
    __shared__ float s_b;
    if (threadIdx.x < 64) s_w[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) s_b = 0.f;
    __syncthreads();
    // reduce over the 4 pixels a warp handles concurrently (lanes with equal sub), then smem atomics
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        accw[k] += __shfl_xor_sync(0xffffffffu, accw[k], 8);
        accw[k] += __shfl_xor_sync(0xffffffffu, accw[k], 16);
    }
    accb += __shfl_xor_sync(0xffffffffu, accb, 8);
    accb += __shfl_xor_sync(0xffffffffu, accb, 16);
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_w[sub * 8 + k], accw[k]);
        if (sub == 0) atomicAdd(&s_b, accb);
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(dw + threadIdx.x, s_w[threadIdx.x]);
    if (threadIdx.x == 0) atomicAdd(db, s_b);
}

// Deterministic variant: block b STORES its 65 partials (dw[64], db) into ws[b][65+]; the last block (ticket) sums the
// slots in block order and WRITES dw / db (no pre-zeroing, no atomics on data). counter: one zero-initialised unsigned.
__global__ void __launch_bounds__(256) head_bwd_det_kernel(const float* __restrict__ dout,
                                                           const __nv_bfloat16* __restrict__ act,
                                                           const float* __restrict__ w, __nv_bfloat16* __restrict__ dact,
                                                           float* __restrict__ dw, float* __restrict__ db,
                                                           float* __restrict__ ws, unsigned* __restrict__ counter,
                                                           long long npix) {
    const int sub = threadIdx.x & 7;
    const F8 wv = ld_f32x8(w + sub * 8);
    float accw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[k] = 0.f;
    float accb = 0.f;
    // four pixels per iteration, their loads issued before the first store (one pixel per iteration left every load
    // waiting behind the previous store: 4.5 TB/s)
    const long long stride = (static_cast<long long>(gridDim.x) * blockDim.x) >> 3;
    long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3;
    for (; p + 3 * stride < npix; p += 4 * stride) {
        float g[4];
        F8 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            g[u] = dout[p + u * stride];
            a[u] = ld_bf16x8(act + (p + u * stride) * 64 + sub * 8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            F8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o.v[k] = g[u] * wv.v[k];
                accw[k] = fmaf(g[u], a[u].v[k], accw[k]);
            }
            st_bf16x8(dact + (p + u * stride) * 64 + sub * 8, o);
            if (sub == 0) accb += g[u];
        }
    }
    for (; p < npix; p += stride) {
        const float g = dout[p];
        const F8 a = ld_bf16x8(act + p * 64 + sub * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o.v[k] = g * wv.v[k];
            accw[k] = fmaf(g, a.v[k], accw[k]);
        }
        st_bf16x8(dact + p * 64 + sub * 8, o);
        if (sub == 0) accb += g;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        accw[k] += __shfl_xor_sync(0xffffffffu, accw[k], 8);
        accw[k] += __shfl_xor_sync(0xffffffffu, accw[k], 16);
    }
    accb += __shfl_xor_sync(0xffffffffu, accb, 8);
    accb += __shfl_xor_sync(0xffffffffu, accb, 16);
    __shared__ float s_w[8][66];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_w[warp][sub * 8 + k] = accw[k];
        if (sub == 0) s_w[warp][64] = accb;
    }
    __syncthreads();
    float* slot = ws + static_cast<size_t>(blockIdx.x) * 72;
    if (threadIdx.x < 65) {
        float acc = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) acc += s_w[w8][threadIdx.x];
        slot[threadIdx.x] = acc;
    }
    if (!last_block_ticket(counter, gridDim.x)) return;
    // 65 outputs x 3 slot-lanes (195 threads), combined in lane order
    __shared__ float s_t[3][66];
    const int o = threadIdx.x % 65, sl = threadIdx.x / 65;
    if (sl < 3) {
        float acc = 0.f;
#pragma unroll 16
        for (int r = sl; r < static_cast<int>(gridDim.x); r += 3) acc += __ldcg(ws + static_cast<size_t>(r) * 72 + o);
        s_t[sl][o] = acc;
    }
    __syncthreads();
    if (threadIdx.x < 65) {
        const float t = s_t[0][threadIdx.x] + s_t[1][threadIdx.x] + s_t[2][threadIdx.x];
        if (threadIdx.x < 64) dw[threadIdx.x] = t;
        else *db = t;
    }
}


// head_bwd_det_kernel that ALSO produces the BatchNorm+ReLU backward reduction (pass 1) of the layer feeding the head
// (dec1.conv.3 -> BN -> ReLU -> final 1x1 conv, unet_model.py:76-80,113-117): the dact it writes is the tensor
// bn_bwd_reduce_det would read back, so that pass (and its 2 B/element re-read of dact) disappears. Sums are taken from the
// values as stored (bf16), like the two-kernel path. grid = (slices, 1), 64 channels: thread (tx = tid & 7, ty = tid >> 3)
// owns channels 8 tx .. 8 tx + 7 of every slices-th group of 32 pixels. ws: [slices][72] head partials, then [slices][128]
// BatchNorm partials; counters: [0] BatchNorm ticket (one channel group), [1] head ticket.
__global__ void __launch_bounds__(256, 2) head_bwd_bnred_kernel(
    const float* __restrict__ dout, const __nv_bfloat16* __restrict__ act, const float* __restrict__ w,
    __nv_bfloat16* __restrict__ dact, float* __restrict__ dw, float* __restrict__ db, const __nv_bfloat16* __restrict__ z,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, float* __restrict__ sums, float* __restrict__ ws, unsigned* __restrict__ counters,
    long long npix) {
    griddep_launch_dependents();
    const int slices = gridDim.x, slice = blockIdx.x;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int c = tx * 8;
    const F8 wv = ld_f32x8(w + c), sc = ld_f32x8(scale + c), sh = ld_f32x8(shift + c), mu = ld_f32x8(mean + c);
    float accw[8], s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[k] = s1[k] = s2[k] = 0.f;
    float accb = 0.f;
    constexpr int U = 4;  // four pixels in flight per thread (two spilled registers at three blocks per SM)
    const long long step = static_cast<long long>(slices) * 32;
    for (long long p0 = static_cast<long long>(slice) * 32 + ty; p0 < npix; p0 += step * U) {
        float g[U];
        uint4 ua[U], uz[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                g[u] = dout[p];
                ua[u] = ld_stream(act + p * 64 + c);
                uz[u] = ld_stream(z + p * 64 + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * step;
            if (p < npix) {
                const F8 a = unpack8(ua[u]), zf = unpack8(uz[u]);
                F8 o;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    o.v[k] = bf16_round(g[u] * wv.v[k]);  // the gradient as stored
                    accw[k] = fmaf(g[u], a.v[k], accw[k]);
                    const float gm = fmaf(zf.v[k], sc.v[k], sh.v[k]) > 0.f ? o.v[k] : 0.f;
                    s1[k] += gm;
                    s2[k] = fmaf(gm, zf.v[k] - mu.v[k], s2[k]);
                }
                st_bf16x8(dact + p * 64 + c, o);
                if (tx == 0) accb += g[u];
            }
        }
    }
    // ---- head partials: pixel lanes combined in lane order, slot per block, last block adds the slots in block order ----
    {
        __shared__ float s_h[32][66];
#pragma unroll
        for (int k = 0; k < 8; ++k) s_h[ty][c + k] = accw[k];
        if (tx == 0) s_h[ty][64] = accb;
        __syncthreads();
        float* slot = ws + static_cast<size_t>(slice) * 72;
        if (threadIdx.x < 65) {
            float acc = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) acc += s_h[r][threadIdx.x];
            slot[threadIdx.x] = acc;
        }
        if (last_block_ticket(counters + 1, static_cast<unsigned>(slices))) {
            __shared__ float s_t[3][66];
            const int o = threadIdx.x % 65, sl = threadIdx.x / 65;
            if (sl < 3) {
                float acc = 0.f;
#pragma unroll 16
                for (int r = sl; r < slices; r += 3) acc += __ldcg(ws + static_cast<size_t>(r) * 72 + o);
                s_t[sl][o] = acc;
            }
            __syncthreads();
            if (threadIdx.x < 65) {
                const float t = s_t[0][threadIdx.x] + s_t[1][threadIdx.x] + s_t[2][threadIdx.x];
                if (threadIdx.x < 64) dw[threadIdx.x] = t;
                else *db = t;
            }
        }
    }
    // ---- BatchNorm backward sums (blockIdx.y == 0: one 64-channel group) ----
    bn_red_block_finish(s1, s2, sums, ws + static_cast<size_t>(slices) * 72, counters, invstd, 64);
}

// ------------------------------------------------------------------------------------------------
// fp32-accuracy eval mode: the bandwidth kernels around the SPLIT tensor-core convolutions (conv3x3.cuh). Activations
// are (B,H,W,3C) bf16 [hi | lo | hi] with hi + lo the fp32 value to 2^-17 relative.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_split3(__nv_bfloat16* p, int part_stride, const F8& v) {
    F8 hi, lo;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        hi.v[k] = bf16_round(v.v[k]);
        lo.v[k] = v.v[k] - hi.v[k];
    }
    st_bf16x8(p, hi);
    st_bf16x8(p + part_stride, lo);
    st_bf16x8(p + 2 * part_stride, hi);
}

// First layer Conv2d(2,64,3,p=1) + folded BatchNorm + ReLU in plain fp32 FMAs (0.15 GFLOP per sample: CUDA cores are
// plenty), input fp32 NCHW, output split (B,H,W,192). One thread = one pixel x 8 output channels.
__global__ void __launch_bounds__(256) conv1_split_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int relu,
                                                              __nv_bfloat16* __restrict__ out, int H, int W,
                                                              long long total /* B*H*W*8 */) {
    __shared__ float s_w[18][64];
    for (int i = threadIdx.x; i < 18 * 64; i += 256) s_w[i % 18][i / 18] = wgt[i];  // wgt[co][ci][3][3] -> [k][co]
    __syncthreads();
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(idx & 7);
        long long p = idx >> 3;
        const int w = static_cast<int>(p % W);
        p /= W;
        const int h = static_cast<int>(p % H);
        const long long img = p / H;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
                const float v = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + ((img * 2 + ci) * H + hh) * W + ww) : 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(v, s_w[ci * 9 + t][c8 * 8 + k], acc[k]);
            }
        F8 r;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = fmaf(acc[k], scale[c8 * 8 + k], shift[c8 * 8 + k]);
            r.v[k] = relu ? fmaxf(v, 0.f) : v;
        }
        st_split3(out + ((img * H + h) * W + w) * 192 + c8 * 8, 64, r);
    }
}

// MaxPool2d(2,2) on a split slot: compares hi + lo, copies the winner's pair. in: channel slot of a (.., in_stride)
// buffer whose parts are in_part apart; out: dense (B,H/2,W/2,3C).
__global__ void __launch_bounds__(256) maxpool2x2_split_kernel(const __nv_bfloat16* __restrict__ in, int in_stride,
                                                               int in_coff, int in_part, int C,
                                                               __nv_bfloat16* __restrict__ out, int H, int W,
                                                               long long total) {
    griddep_launch_dependents();  // PDL: a conv kernel launched next may start its set-up while this grid drains
    const int c8n = C >> 3;
    const int W2 = W >> 1, H2 = H >> 1;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        long long r = idx / c8n;
        const int w2 = static_cast<int>(r % W2);
        r /= W2;
        const int h2 = static_cast<int>(r % H2);
        const long long img = r / H2;
        const long long p00 = (img * H + 2 * h2) * W + 2 * w2;
        const long long pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        F8 bh, bl;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16* q = in + pix[j] * in_stride + in_coff + c;
            const F8 hi = ld_bf16x8(q), lo = ld_bf16x8(q + in_part);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (j == 0 || hi.v[k] + lo.v[k] > bh.v[k] + bl.v[k]) {
                    bh.v[k] = hi.v[k];
                    bl.v[k] = lo.v[k];
                }
            }
        }
        __nv_bfloat16* o = out + ((img * H2 + h2) * W2 + w2) * (3 * C) + c;
        st_bf16x8(o, bh);
        st_bf16x8(o + C, bl);
        st_bf16x8(o + 2 * C, bh);
    }
}

// 1x1 head on a split (B,H,W,192) activation: out = b + sum_c (hi + lo)[c] * w[c], fp32.
__global__ void __launch_bounds__(256) head_split_fwd_kernel(const __nv_bfloat16* __restrict__ act,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             float* __restrict__ out, long long npix) {
    const int sub = threadIdx.x & 7;
    const F8 wv = ld_f32x8(w + sub * 8);
    const float bias = b[0];
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 3; p < npix;
         p += (static_cast<long long>(gridDim.x) * blockDim.x) >> 3) {
        const F8 hi = ld_bf16x8(act + p * 192 + sub * 8), lo = ld_bf16x8(act + p * 192 + 64 + sub * 8);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s = fmaf(hi.v[k] + lo.v[k], wv.v[k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (sub == 0) out[p] = s + bias;
    }
}

// ------------------------------------------------------------------------------------------------
// layout casts at the boundary (used by the per-op tests and by users feeding intermediate tensors)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C,
                                             int HW, long long total) {
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = idx % C;
        const long long r = idx / C;
        const int p = r % HW;
        const long long n = r / HW;
        out[idx] = __float2bfloat16_rn(in[(n * C + c) * HW + p]);
    }
}

__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ in, int in_stride, int in_coff,
                                             float* __restrict__ out, int C, int HW, long long total) {
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = idx % HW;
        const long long r = idx / HW;
        const int c = r % C;
        const long long n = r / C;
        out[idx] = __bfloat162float(in[(n * HW + p) * in_stride + in_coff + c]);
    }
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam semantics, unet_model.py:155: lr 1e-4, betas (0.9,0.999), eps 1e-8, wd 0) over the
// flat fp32 parameter / gradient / moment buffers. step-dependent scalars are computed on the host.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        float lr, float beta1, float beta2, float eps,
                                                        float bias_corr1, float bias_corr2_sqrt, float grad_scale,
                                                        const float* __restrict__ bias_corr_dev,
                                                        int* __restrict__ step_dev) {
    // step-dependent scalars may come from device memory so that the launch can live in a CUDA graph
    if (bias_corr_dev != nullptr) {
        bias_corr1 = bias_corr_dev[0];
        bias_corr2_sqrt = bias_corr_dev[1];
    }
    // step_dev = {completed steps, block ticket}: the step count lives on the device, the kernel derives the bias
    // corrections of step t = completed + 1 itself and the last block to finish publishes t — the host never has to
    // hand over per-step scalars, so an unsynchronised host running several steps ahead cannot skew them
    if (step_dev != nullptr) {
        // one thread per block evaluates the two double-precision pow()s (every thread doing so made the kernel FP64-bound:
        // 138 us for 7.76 M parameters against a 30 us HBM floor)
        __shared__ float s_corr[2];
        if (threadIdx.x == 0) {
            const double t = static_cast<double>(step_dev[0] + 1);
            s_corr[0] = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
            s_corr[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)));
        }
        __syncthreads();
        bias_corr1 = s_corr[0];
        bias_corr2_sqrt = s_corr[1];
    }
    for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
        if (i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i);
            float4 vv = *reinterpret_cast<float4*>(v + i);
            float* pa = &pp.x;
            const float* ga = &gg.x;
            float* ma = &mm.x;
            float* va = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gk = ga[k] * grad_scale;
                ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
                va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
                const float denom = sqrtf(va[k]) / bias_corr2_sqrt + eps;
                pa[k] -= (lr / bias_corr1) * (ma[k] / denom);
            }
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long j = i; j < n; ++j) {
                const float gk = g[j] * grad_scale;
                m[j] = beta1 * m[j] + (1.f - beta1) * gk;
                v[j] = beta2 * v[j] + (1.f - beta2) * gk * gk;
                const float denom = sqrtf(v[j]) / bias_corr2_sqrt + eps;
                p[j] -= (lr / bias_corr1) * (m[j] / denom);
            }
        }
    }
    if (step_dev != nullptr) {
        __syncthreads();  // every thread of the block has read step_dev[0]
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(step_dev + 1), 1u);
            if (ticket == gridDim.x - 1) {  // all blocks have read the old count
                step_dev[1] = 0;
                __threadfence();
                step_dev[0] += 1;
            }
        }
    }
}

}  // namespace b200sr
