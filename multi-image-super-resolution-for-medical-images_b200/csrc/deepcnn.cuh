// Kernels that the DeepCNN residual baseline needs on top of the UNet hot-path kernels (BASELINE
// configs[1]; reference /root/reference/src/ModelLoader.py:276-377). The 3x3 convolutions, BatchNorm statistics /
// finalize / apply, 1x1 convolutions and weight gradients reuse the tensor-core and bandwidth kernels of the UNet
// path; new here:
//   * Conv2d(2 -> 64, 7x7, padding 3) forward / wgrad straight from the fp32 NCHW input (K = 98: CUDA cores)
//   * MaxPool2d(3, stride 1, padding 1) forward / backward (first maximum in row-major window order, like ATen)
//   * residual tail  out = relu(bn2(z2) + identity)  with identity = x or bn_d(z_d), and the matching gradient add
//   * 1x1 head for wide inputs (Conv2d(512 -> 1) + bias) forward / backward
#pragma once
#include "elementwise.cuh"

namespace b200sr {

// ------------------------------------------------------------------------------------------------
// stem: Conv2d(2 -> 64, 7x7, p3). One thread = one pixel x 64 channels (two passes of 32), 16x16 tiles, 22x22 halo.
// ------------------------------------------------------------------------------------------------
constexpr int C7_K = 7;
constexpr int C7_HALO = C1_TILE + C7_K - 1;  // 22
constexpr int C7_TAPS = 2 * C7_K * C7_K;     // 98

__global__ void __launch_bounds__(256) conv7_direct_fwd_kernel(const float* __restrict__ x,     // [B][2][H][W]
                                                               const float* __restrict__ wgt,   // [64][2][7][7]
                                                               __nv_bfloat16* __restrict__ out,  // [B][H][W][64]
                                                               float* __restrict__ stats, int stats_replicas, int H,
                                                               int W, int num_tiles) {
    __shared__ float s_x[2][C7_HALO][C7_HALO + 1];
    __shared__ __align__(16) float s_w[C7_TAPS][C1_COUT];  // [ci*49 + kh*7 + kw][co]
    __shared__ float s_stats[2][C1_COUT];
    const int tid = threadIdx.x;
    const int tiles_w = W / C1_TILE;
    const int tiles_hw = tiles_w * (H / C1_TILE);
    const uint32_t lane = tid & 31;
    const int ph = tid / C1_TILE, pw = tid % C1_TILE;
    for (int i = tid; i < C7_TAPS * C1_COUT; i += 256) {
        const int co = i % C1_COUT, k = i / C1_COUT;
        s_w[k][co] = wgt[co * C7_TAPS + k];
    }
    if (tid < 2 * C1_COUT) (&s_stats[0][0])[tid] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_hw;
        const int t_in = tile - img * tiles_hw;
        const int h0 = (t_in / tiles_w) * C1_TILE, w0 = (t_in % tiles_w) * C1_TILE;
        __syncthreads();
        for (int i = tid; i < 2 * C7_HALO * C7_HALO; i += 256) {
            const int ci = i / (C7_HALO * C7_HALO);
            const int r = i % (C7_HALO * C7_HALO);
            const int hh = h0 + r / C7_HALO - 3, ww = w0 + r % C7_HALO - 3;
            float v = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((static_cast<size_t>(img) * 2 + ci) * H + hh) * W + ww];
            s_x[ci][r / C7_HALO][r % C7_HALO] = v;
        }
        __syncthreads();
        __nv_bfloat16* dst = out + ((static_cast<size_t>(img) * H + h0 + ph) * W + w0 + pw) * C1_COUT;
#pragma unroll 1
        for (int cb = 0; cb < C1_COUT; cb += 32) {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll 1
            for (int ci = 0; ci < 2; ++ci)
#pragma unroll 1
                for (int kh = 0; kh < C7_K; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < C7_K; ++kw) {
                        const float xv = s_x[ci][ph + kh][pw + kw];
                        const float* wrow = &s_w[ci * 49 + kh * 7 + kw][cb];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
                            acc[j] = fmaf(xv, w4.x, acc[j]);
                            acc[j + 1] = fmaf(xv, w4.y, acc[j + 1]);
                            acc[j + 2] = fmaf(xv, w4.z, acc[j + 2]);
                            acc[j + 3] = fmaf(xv, w4.w, acc[j + 3]);
                        }
                    }
                }
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(acc[2 * j], acc[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                reinterpret_cast<uint4*>(dst + cb)[j] =
                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            if (stats != nullptr) {
                float s1[32], s2[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&packed[j]);
                    const float a = __low2float(hh), b = __high2float(hh);
                    s1[2 * j] = a; s1[2 * j + 1] = b;
                    s2[2 * j] = a * a; s2[2 * j + 1] = b * b;
                }
                const float cs = warp_transpose_reduce32(s1, lane);
                const float cq = warp_transpose_reduce32(s2, lane);
                atomicAdd(&s_stats[0][cb + lane], cs);
                atomicAdd(&s_stats[1][cb + lane], cq);
            }
        }
    }
    if (stats != nullptr) {
        __syncthreads();
        float* d = stats + static_cast<size_t>(blockIdx.x % stats_replicas) * 2 * C1_COUT;
        if (tid < 2 * C1_COUT) atomicAdd(d + tid, (&s_stats[0][0])[tid]);
    }
}

// wgrad of the stem: dW[co][ci][kh][kw] = sum_q dZ[q][co] * x[q + (kh-3, kw-3)][ci].
// Thread = (group of 4 output channels, one (ci,kh) row of 7 taps): 28 accumulators; 16 x 14 = 224 active threads.
__global__ void __launch_bounds__(256) conv7_direct_wgrad_kernel(const float* __restrict__ x,            // [B][2][H][W]
                                                                 const __nv_bfloat16* __restrict__ dz,  // [B][H][W][64]
                                                                 float* __restrict__ dw,                 // [64][2][7][7]
                                                                 int H, int W, int num_tiles) {
    __shared__ float s_x[2][C7_HALO][C7_HALO + 1];
    __shared__ __align__(16) __nv_bfloat16 s_dz[C1_TILE * C1_TILE][C1_COUT + 8];
    const int tid = threadIdx.x;
    const int cg = tid & 15;        // channels 4*cg .. 4*cg+3
    const int row = tid >> 4;       // (ci, kh) row: ci = row / 7, kh = row % 7; rows 14, 15 idle
    const bool active = row < 14;
    const int ci = row / 7, kh = row % 7;
    const int tiles_w = W / C1_TILE;
    const int tiles_hw = tiles_w * (H / C1_TILE);
    float acc[4][C7_K];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < C7_K; ++k) acc[j][k] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_hw;
        const int t_in = tile - img * tiles_hw;
        const int h0 = (t_in / tiles_w) * C1_TILE, w0 = (t_in % tiles_w) * C1_TILE;
        __syncthreads();
        for (int i = tid; i < 2 * C7_HALO * C7_HALO; i += 256) {
            const int c = i / (C7_HALO * C7_HALO);
            const int r = i % (C7_HALO * C7_HALO);
            const int hh = h0 + r / C7_HALO - 3, ww = w0 + r % C7_HALO - 3;
            float v = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[((static_cast<size_t>(img) * 2 + c) * H + hh) * W + ww];
            s_x[c][r / C7_HALO][r % C7_HALO] = v;
        }
        for (int i = tid; i < C1_TILE * C1_TILE * 8; i += 256) {
            const int p = i >> 3, c8 = i & 7;
            const int hh = h0 + p / C1_TILE, ww = w0 + p % C1_TILE;
            const uint4 u =
                *reinterpret_cast<const uint4*>(dz + ((static_cast<size_t>(img) * H + hh) * W + ww) * C1_COUT + c8 * 8);
            *reinterpret_cast<uint4*>(&s_dz[p][c8 * 8]) = u;
        }
        __syncthreads();
        if (active) {
#pragma unroll 2
            for (int p = 0; p < C1_TILE * C1_TILE; ++p) {
                const uint2 gu = *reinterpret_cast<const uint2*>(&s_dz[p][cg * 4]);
                const __nv_bfloat162 g01 = *reinterpret_cast<const __nv_bfloat162*>(&gu.x);
                const __nv_bfloat162 g23 = *reinterpret_cast<const __nv_bfloat162*>(&gu.y);
                const float g[4] = {__low2float(g01), __high2float(g01), __low2float(g23), __high2float(g23)};
                const float* xr = &s_x[ci][p / C1_TILE + kh][p % C1_TILE];
#pragma unroll
                for (int kw = 0; kw < C7_K; ++kw) {
                    const float xv = xr[kw];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][kw] = fmaf(g[j], xv, acc[j][kw]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int kw = 0; kw < C7_K; ++kw)
                atomicAdd(dw + (cg * 4 + j) * C7_TAPS + ci * 49 + kh * 7 + kw, acc[j][kw]);
    }
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 1, padding 1) on dense NHWC bf16 (reference ModelLoader.py:327). Padding is -inf.
// One thread = 8 channels of one pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool3x3_fwd_kernel(const __nv_bfloat16* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int C, int H, int W,
                                                             long long total /* B*H*W*(C/8) */) {
    const int c8n = C >> 3;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c8n) * 8;
        long long r = idx / c8n;
        const int w = static_cast<int>(r % W);
        r /= W;
        const int h = static_cast<int>(r % H);
        const long long img = r / H;
        F8 m;
#pragma unroll
        for (int k = 0; k < 8; ++k) m.v[k] = -INFINITY;
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
            for (int dw = -1; dw <= 1; ++dw) {
                const int hh = h + dh, ww = w + dw;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                const F8 v = ld_bf16x8(in + ((img * H + hh) * W + ww) * C + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) m.v[k] = fmaxf(m.v[k], v.v[k]);
            }
        st_bf16x8(out + ((img * H + h) * W + w) * C + c, m);
    }
}

// backward as a gather: dIn[p] = sum over the (up to 9) windows q containing p of dOut[q] * [argmax(q) == p], where
// argmax(q) is the FIRST maximum of window q in row-major order (ATen max_pool2d_with_indices).
// One thread = 2 channels of one pixel (the 5x5 neighbourhood lives in registers).
__global__ void __launch_bounds__(256) maxpool3x3_bwd_kernel(const __nv_bfloat16* __restrict__ in,
                                                             const __nv_bfloat16* __restrict__ dout,
                                                             __nv_bfloat16* __restrict__ din, int C, int H, int W,
                                                             long long total /* B*H*W*(C/2) */) {
    const int c2n = C >> 1;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c2n) * 2;
        long long r = idx / c2n;
        const int w = static_cast<int>(r % W);
        r /= W;
        const int h = static_cast<int>(r % H);
        const long long img = r / H;
        float2 nb[5][5];  // 5x5 neighbourhood of the input around p (missing = -inf)
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const int hh = h + a - 2, ww = w + b - 2;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) {
                    nb[a][b] = make_float2(-INFINITY, -INFINITY);
                } else {
                    const __nv_bfloat162 v =
                        *reinterpret_cast<const __nv_bfloat162*>(in + ((img * H + hh) * W + ww) * C + c);
                    nb[a][b] = make_float2(__low2float(v), __high2float(v));
                }
            }
        float2 acc = make_float2(0.f, 0.f);
        const float2 pv = nb[2][2];
        // window centred at q = p + (qa-1, qb-1), qa,qb in 0..2; p sits at window position (2-qa, 2-qb)
#pragma unroll
        for (int qa = 0; qa < 3; ++qa)
#pragma unroll
            for (int qb = 0; qb < 3; ++qb) {
                const int qh = h + qa - 1, qw = w + qb - 1;
                if (qh < 0 || qh >= H || qw < 0 || qw >= W) continue;
                const __nv_bfloat162 gv =
                    *reinterpret_cast<const __nv_bfloat162*>(dout + ((img * H + qh) * W + qw) * C + c);
                const int pa = 2 - qa, pb = 2 - qb;  // position of p inside window q (row, col)
                bool ax = true, ay = true;
#pragma unroll
                for (int u = 0; u < 3; ++u)
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        if (u == pa && v == pb) continue;
                        const float2 o = nb[qa + u][qb + v];
                        const bool before = (u < pa) || (u == pa && v < pb);
                        // p is the first maximum iff every earlier element is < p and every later element is <= p
                        ax = ax && (before ? (o.x < pv.x) : (o.x <= pv.x));
                        ay = ay && (before ? (o.y < pv.y) : (o.y <= pv.y));
                    }
                if (ax) acc.x += __low2float(gv);
                if (ay) acc.y += __high2float(gv);
            }
        *reinterpret_cast<__nv_bfloat162*>(din + ((img * H + h) * W + w) * C + c) = __floats2bfloat162_rn(acc.x, acc.y);
    }
}

// ------------------------------------------------------------------------------------------------
// residual tail (ModelLoader.py:290-307): out = relu(scale2*z2 + shift2 + identity), identity = x (dense bf16) or
// scale_d*z_d + shift_d (the 1x1 downsample branch). One thread = 8 fixed channels, walks pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_add_relu_kernel(const __nv_bfloat16* __restrict__ z2,
                                                          const float* __restrict__ scale2,
                                                          const float* __restrict__ shift2,
                                                          const __nv_bfloat16* __restrict__ idn,
                                                          const float* __restrict__ scale_d,  // nullable: idn is raw x
                                                          const float* __restrict__ shift_d,
                                                          __nv_bfloat16* __restrict__ out, int C, long long npix) {
    const int CV = C >> 3;
    const int PB = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int c = cv * 8;
    const F8 s2 = ld_f32x8(scale2 + c), h2 = ld_f32x8(shift2 + c);
    F8 sd, hd;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sd.v[k] = 1.f;
        hd.v[k] = 0.f;
    }
    if (scale_d != nullptr) {
        sd = ld_f32x8(scale_d + c);
        hd = ld_f32x8(shift_d + c);
    }
    const long long step = static_cast<long long>(gridDim.x) * PB;
    for (long long p = static_cast<long long>(blockIdx.x) * PB + pl; p < npix; p += step) {
        const F8 a = unpack8(ld_stream(z2 + p * C + c)), b = unpack8(ld_stream(idn + p * C + c));
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            o.v[k] = fmaxf(fmaf(a.v[k], s2.v[k], h2.v[k]) + fmaf(b.v[k], sd.v[k], hd.v[k]), 0.f);
        st_bf16x8(out + p * C + c, o);
    }
}

// gradient join of a residual block: out = a + b * [mask > 0] (mask nullable: plain add). n8 = elements / 8.
__global__ void __launch_bounds__(256) add_masked_kernel(const __nv_bfloat16* __restrict__ a,
                                                         const __nv_bfloat16* __restrict__ b,
                                                         const __nv_bfloat16* __restrict__ mask,
                                                         __nv_bfloat16* __restrict__ out, long long n8) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const F8 x = unpack8(ld_stream(a + i * 8)), y = unpack8(ld_stream(b + i * 8));
        F8 o;
        if (mask != nullptr) {
            const F8 m = unpack8(ld_stream(mask + i * 8));
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = x.v[k] + (m.v[k] > 0.f ? y.v[k] : 0.f);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = x.v[k] + y.v[k];
        }
        st_bf16x8(out + i * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------
// wide 1x1 head: Conv2d(C -> 1) + bias, C % 256 == 0 (DeepCNN output_conv, ModelLoader.py:336,375).
// One warp per pixel: lane owns C/32 consecutive channels.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) headw_fwd_kernel(const __nv_bfloat16* __restrict__ act,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ out, long long npix) {
    constexpr int PER = C / 32;  // channels per lane (multiple of 8)
    const int lane = threadIdx.x & 31;
    float wv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) wv[i] = w[lane * PER + i];
    const float bias = b[0];
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; p < npix; p += warps) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; i += 8) {
            const F8 a = unpack8(ld_stream(act + p * C + lane * PER + i));
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(a.v[k], wv[i + k], s);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[p] = s + bias;
    }
}

template <int C>
__global__ void __launch_bounds__(256) headw_bwd_kernel(const float* __restrict__ dout,
                                                        const __nv_bfloat16* __restrict__ act,
                                                        const float* __restrict__ w, __nv_bfloat16* __restrict__ dact,
                                                        float* __restrict__ dw, float* __restrict__ db, long long npix) {
    constexpr int PER = C / 32;
    const int lane = threadIdx.x & 31;
    float wv[PER], accw[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        wv[i] = w[lane * PER + i];
        accw[i] = 0.f;
    }
    float accb = 0.f;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long p = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; p < npix; p += warps) {
        const float g = dout[p];
#pragma unroll
        for (int i = 0; i < PER; i += 8) {
            const F8 a = unpack8(ld_stream(act + p * C + lane * PER + i));
            F8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o.v[k] = g * wv[i + k];
                accw[i + k] = fmaf(g, a.v[k], accw[i + k]);
            }
            st_bf16x8(dact + p * C + lane * PER + i, o);
        }
        accb += g;
    }
    __shared__ float s_w[C];
    __shared__ float s_b;
    for (int i = threadIdx.x; i < C; i += 256) s_w[i] = 0.f;
    if (threadIdx.x == 0) s_b = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PER; ++i) atomicAdd(&s_w[lane * PER + i], accw[i]);
    if (lane == 0) atomicAdd(&s_b, accb);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(dw + i, s_w[i]);
    if (threadIdx.x == 0) atomicAdd(db, s_b);
}

}  // namespace b200sr
