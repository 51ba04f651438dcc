// Fast-DDPM denoiser kernels (reference src/ModelLoader.py:471-636; SURVEY.md §8(f) row 4, BASELINE configs[4]).
//
// UNet2D concatenates the time embedding, TILED over all pixels, to the 3 image channels (:565-570), so its first
// conv (259 -> 64) spends 98.8 % of its FLOPs on an input that is constant over space. A 3x3 zero-padded conv of a
// spatially constant 256-channel input is a per-sample bias that depends only on which taps fall inside the image:
// 9 border classes (first / interior / last row x first / interior / last column). The kernels here compute
//     z[b,h,w,co] = conv3x3(image channels)[b,h,w,co] + tb[b][class(h,w)][co],
//     tb[b][cls][co] = bias[co] + sum_{taps valid in cls} sum_c W[co][3+c][tap] * e[b][c]
// which is the same sum the reference evaluates, re-associated (19.3 of 19.55 GFLOP per sample never issued), and
// the matching backward: the weight gradient of the time channels needs only border sums of dz.
// All 3x3 convs behind the first one run on the tensor-core kernels of the UNet path (conv3x3.cuh / wgrad3x3.cuh).
#pragma once
#include "elementwise.cuh"

namespace b200sr {

constexpr int FD_TDIM = 256;               // time_dim (ModelLoader.py:543)
constexpr int FD_CIMG = 3;                 // x_t, pre, post
constexpr int FD_CIN0 = FD_CIMG + FD_TDIM;  // 259 input channels of inc.block.0
constexpr int FD_K0 = FD_CIMG * 9;         // 27

// ------------------------------------------------------------------------------------------------
// sinusoidal_timestep_embedding (:475-487) + time_mlp = Linear -> ReLU -> Linear (:547-551). One block per sample;
// a warp per output row so the weight reads are coalesced. emb / hid are kept for the backward pass.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fd_time_mlp_fwd_kernel(const long long* __restrict__ t,
                                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              float* __restrict__ emb, float* __restrict__ hid,
                                                              float* __restrict__ e) {
    __shared__ float s_in[FD_TDIM];
    __shared__ float s_h[FD_TDIM];
    const int b = blockIdx.x, tid = threadIdx.x;
    constexpr int half = FD_TDIM / 2;
    {
        const int i = tid % half;
        const float freq = expf(-logf(10000.f) * static_cast<float>(i) / static_cast<float>(half));
        const float arg = static_cast<float>(t[b]) * freq;
        const float v = tid < half ? sinf(arg) : cosf(arg);
        s_in[tid] = v;
        emb[b * FD_TDIM + tid] = v;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < FD_TDIM; j += 8) {
        const float* row = w1 + j * FD_TDIM;
        float acc = 0.f;
#pragma unroll
        for (int k = lane; k < FD_TDIM; k += 32) acc = fmaf(row[k], s_in[k], acc);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float h = fmaxf(acc + b1[j], 0.f);
            s_h[j] = h;
            hid[b * FD_TDIM + j] = h;
        }
    }
    __syncthreads();
    for (int j = warp; j < FD_TDIM; j += 8) {
        const float* row = w2 + j * FD_TDIM;
        float acc = 0.f;
#pragma unroll
        for (int k = lane; k < FD_TDIM; k += 32) acc = fmaf(row[k], s_h[k], acc);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) e[b * FD_TDIM + j] = acc + b2[j];
    }
}

// tb[b][cls][co] (see the header comment). One block per sample, thread = (co, tap) with tap fastest so that a warp
// reads runs of 9 consecutive weights.
__global__ void __launch_bounds__(576) fd_time_bias_kernel(const float* __restrict__ e,     // [B][256]
                                                           const float* __restrict__ wgt,   // [64][259][3][3]
                                                           const float* __restrict__ bias,  // [64]
                                                           float* __restrict__ tb) {        // [B][9][64]
    __shared__ float s_e[FD_TDIM];
    __shared__ float s_p[9][C1_COUT];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < FD_TDIM) s_e[tid] = e[b * FD_TDIM + tid];
    __syncthreads();
    {
        const int tap = tid % 9, co = tid / 9;
        const float* wp = wgt + (static_cast<size_t>(co) * FD_CIN0 + FD_CIMG) * 9 + tap;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < FD_TDIM; ++c) acc = fmaf(wp[c * 9], s_e[c], acc);
        s_p[tap][co] = acc;
    }
    __syncthreads();
    {
        const int cls = tid / C1_COUT, co = tid % C1_COUT;
        const int rc = cls / 3, cc = cls % 3;
        float s = bias[co];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            if ((rc == 0 && kh == 0) || (rc == 2 && kh == 2)) continue;  // input row h + kh - 1 is outside the image
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                if ((cc == 0 && kw == 0) || (cc == 2 && kw == 2)) continue;
                s += s_p[kh * 3 + kw][co];
            }
        }
        tb[(static_cast<size_t>(b) * 9 + cls) * C1_COUT + co] = s;
    }
}

// inc.block.0 itself (3-channel conv + tb + ReLU, with q_sample fused into the input load) and the weight gradient of its
// image channels run on the tensor-core first-layer kernels in firstconv.cuh (first_conv_mma_{fwd,wgrad}<3>).

// ------------------------------------------------------------------------------------------------
// DoubleConv backward (Conv3x3 + bias -> ReLU, no BatchNorm, :521-533): dz = dy * [act > 0] fused with the bias
// gradient. Per-SAMPLE channel sums ps[b][c] += sum_pixels dz: the first conv also needs them per sample.
// grid (blocks per image, B); C/8 must divide 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fd_relu_bwd_bias_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                                               int dy_coff, const __nv_bfloat16* __restrict__ act,
                                                               int act_stride, int act_coff,
                                                               __nv_bfloat16* __restrict__ dz, float* __restrict__ ps,
                                                               int C, int HW) {
    __shared__ float s_c[256];
    const int tid = threadIdx.x;
    const int c8n = C >> 3;
    const int rows = 256 / c8n;
    const int cl = tid % c8n, pr = tid / c8n;
    const int b = blockIdx.y;
    if (tid < C) s_c[tid] = 0.f;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int p = blockIdx.x * rows + pr; p < HW; p += gridDim.x * rows) {
        const size_t q = static_cast<size_t>(b) * HW + p;
        const F8 g = unpack8(ld_stream(dy + q * dy_stride + dy_coff + cl * 8));
        const F8 a = ld_bf16x8(act + q * act_stride + act_coff + cl * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o.v[k] = a.v[k] > 0.f ? g.v[k] : 0.f;
            acc[k] += o.v[k];
        }
        st_bf16x8(dz + q * C + cl * 8, o);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_c[cl * 8 + k], acc[k]);
    __syncthreads();
    if (tid < C) atomicAdd(ps + static_cast<size_t>(b) * C + tid, s_c[tid]);
}

// The same fusion for gradients that are PRODUCED by a bandwidth-bound kernel: the ReLU mask and the bias sums are
// applied where the gradient is formed, so it never makes a round trip through HBM unmasked.
//   (1) nearest-upsample backward (2x2 block sum of a concat-slot gradient) -> dz of the layer below
__global__ void __launch_bounds__(256) fd_upsample2x_bwd_relu_kernel(const __nv_bfloat16* __restrict__ dout, int dout_stride,
                                                                     int dout_coff, const __nv_bfloat16* __restrict__ act,
                                                                     __nv_bfloat16* __restrict__ dz, float* __restrict__ ps,
                                                                     int C, int h, int w) {
    __shared__ float s_c[256];
    const int tid = threadIdx.x;
    const int c8n = C >> 3;
    const int rows = 256 / c8n;
    const int cl = tid % c8n, pr = tid / c8n;
    const int b = blockIdx.y;
    const int HW = h * w, W2 = 2 * w;
    if (tid < C) s_c[tid] = 0.f;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int p = blockIdx.x * rows + pr; p < HW; p += gridDim.x * rows) {
        const int hh = p / w, ww = p - hh * w;
        const __nv_bfloat16* src =
            dout + ((static_cast<size_t>(b) * 2 * h + 2 * hh) * W2 + 2 * ww) * dout_stride + dout_coff + cl * 8;
        const F8 g0 = unpack8(ld_stream(src)), g1 = unpack8(ld_stream(src + dout_stride));
        const F8 g2 = unpack8(ld_stream(src + static_cast<size_t>(W2) * dout_stride));
        const F8 g3 = unpack8(ld_stream(src + static_cast<size_t>(W2 + 1) * dout_stride));
        const size_t q = static_cast<size_t>(b) * HW + p;
        const F8 a = ld_bf16x8(act + q * C + cl * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // the unfused path stores the 2x2 sum as bf16 before masking: round the same way
            const float sum = bf16_round((g0.v[k] + g1.v[k]) + (g2.v[k] + g3.v[k]));
            o.v[k] = a.v[k] > 0.f ? sum : 0.f;
            acc[k] += o.v[k];
        }
        st_bf16x8(dz + q * C + cl * 8, o);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_c[cl * 8 + k], acc[k]);
    __syncthreads();
    if (tid < C) atomicAdd(ps + static_cast<size_t>(b) * C + tid, s_c[tid]);
}

//   (2) 1x1 head backward (outc, :558,585): d u1 = dOut * w masked by u1 > 0; dW, db of the head; bias sums of up1.2
__global__ void __launch_bounds__(256) fd_head_bwd_relu_kernel(const float* __restrict__ dout,
                                                               const __nv_bfloat16* __restrict__ act,  // [B][HW][64]
                                                               const float* __restrict__ w, __nv_bfloat16* __restrict__ dz,
                                                               float* __restrict__ dw, float* __restrict__ db,
                                                               float* __restrict__ ps, int HW) {
    __shared__ float s_w[64], s_p[64], s_b;
    const int tid = threadIdx.x, sub = tid & 7, pl = tid >> 3;
    const int b = blockIdx.y;
    const F8 wv = ld_f32x8(w + sub * 8);
    if (tid < 64) s_w[tid] = s_p[tid] = 0.f;
    if (tid == 0) s_b = 0.f;
    float accw[8], accp[8], accb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) accw[k] = accp[k] = 0.f;
    for (int p = blockIdx.x * 32 + pl; p < HW; p += gridDim.x * 32) {
        const size_t q = static_cast<size_t>(b) * HW + p;
        const float g = dout[q];
        const F8 a = ld_bf16x8(act + q * 64 + sub * 8);
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o.v[k] = a.v[k] > 0.f ? bf16_round(g * wv.v[k]) : 0.f;
            accw[k] = fmaf(g, a.v[k], accw[k]);
            accp[k] += o.v[k];
        }
        st_bf16x8(dz + q * 64 + sub * 8, o);
        if (sub == 0) accb += g;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        atomicAdd(&s_w[sub * 8 + k], accw[k]);
        atomicAdd(&s_p[sub * 8 + k], accp[k]);
    }
    if (sub == 0) atomicAdd(&s_b, accb);
    __syncthreads();
    if (tid < 64) {
        atomicAdd(dw + tid, s_w[tid]);
        atomicAdd(ps + static_cast<size_t>(b) * 64 + tid, s_p[tid]);
    }
    if (tid == 0) atomicAdd(db, s_b);
}

//   (3) MaxPool2d(2,2) backward + skip-connection gradient (F.max_pool2d, :574-575; torch.cat skip, :580,583),
//       masked by the ReLU of the layer that produced the pooled tensor (`act` is both the arg-max source and the mask)
__global__ void __launch_bounds__(256) fd_maxpool2x2_bwd_relu_kernel(const __nv_bfloat16* __restrict__ act, int act_stride,
                                                                     int act_coff, const __nv_bfloat16* __restrict__ dpool,
                                                                     const __nv_bfloat16* __restrict__ dskip,
                                                                     int dskip_stride, int dskip_coff, int C,
                                                                     __nv_bfloat16* __restrict__ dz, float* __restrict__ ps,
                                                                     int H, int W) {
    __shared__ float s_c[256];
    const int tid = threadIdx.x;
    const int c8n = C >> 3;
    const int rows = 256 / c8n;
    const int cl = tid % c8n, pr = tid / c8n;
    const int c = cl * 8;
    const int b = blockIdx.y;
    const int W2 = W >> 1, H2 = H >> 1;
    if (tid < C) s_c[tid] = 0.f;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int win = blockIdx.x * rows + pr; win < H2 * W2; win += gridDim.x * rows) {
        const int h2 = win / W2, w2 = win - h2 * W2;
        const size_t p00 = (static_cast<size_t>(b) * H + 2 * h2) * W + 2 * w2;
        const size_t pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        F8 a[4];
        uint4 skr[4];  // all loads of the window are issued before the first store
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] = ld_bf16x8(act + pix[j] * act_stride + act_coff + c);
            skr[j] = ld_stream(dskip + pix[j] * dskip_stride + dskip_coff + c);
        }
        const F8 g = unpack8(ld_stream(dpool + ((static_cast<size_t>(b) * H2 + h2) * W2 + w2) * C + c));
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
            const F8 sk = unpack8(skr[j]);
            F8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float v = bf16_round(sk.v[k] + ((arg[k] == j) ? g.v[k] : 0.f));
                o.v[k] = a[j].v[k] > 0.f ? v : 0.f;
                acc[k] += o.v[k];
            }
            st_bf16x8(dz + pix[j] * C + c, o);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_c[c + k], acc[k]);
    __syncthreads();
    if (tid < C) atomicAdd(ps + static_cast<size_t>(b) * C + tid, s_c[tid]);
}

struct FdBiasJob {
    const float* ps;  // [rows][stride] partial sums (per sample, or the statistic replicas of a conv epilogue)
    float* dst;       // [C] bias gradient, ADDED into
    int C;
    int rows;         // 0: one row per sample (B)
    int stride;       // floats between rows; 0: C
    int pad;
};
__global__ void fd_bias_finish_kernel(const FdBiasJob* __restrict__ jobs, int B) {
    const FdBiasJob job = jobs[blockIdx.x];
    const int rows = job.rows > 0 ? job.rows : B;
    const int stride = job.stride > 0 ? job.stride : job.C;
    for (int c = threadIdx.x; c < job.C; c += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < rows; ++b) s += job.ps[static_cast<size_t>(b) * stride + c];
        job.dst[c] += s;
    }
}

// Tap sums of dz for the time channels of inc.block.0:  S[b][tap][co] = sum over pixels whose tap lies inside the image
// = T - (excluded border row) - (excluded border column) + (their corner). T comes from ps (fd_relu_bwd_bias).
__global__ void __launch_bounds__(256) fd_tap_sums_kernel(const __nv_bfloat16* __restrict__ dz,  // [B][H][W][64]
                                                          const float* __restrict__ ps,          // [B][64]
                                                          float* __restrict__ S,                 // [B][9][64]
                                                          int H, int W) {
    __shared__ float s_line[4][C1_COUT];  // row 0, row H-1, col 0, col W-1
    const int b = blockIdx.x, tid = threadIdx.x;
    (&s_line[0][0])[tid] = 0.f;
    __syncthreads();
    const int cl = tid & 7, pl = tid >> 3;  // 8 channel lanes x 32 pixel lanes
    const __nv_bfloat16* base = dz + static_cast<size_t>(b) * H * W * C1_COUT;
#pragma unroll
    for (int line = 0; line < 4; ++line) {
        const int len = line < 2 ? W : H;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int i = pl; i < len; i += 32) {
            size_t pix;
            if (line == 0) pix = i;
            else if (line == 1) pix = static_cast<size_t>(H - 1) * W + i;
            else if (line == 2) pix = static_cast<size_t>(i) * W;
            else pix = static_cast<size_t>(i) * W + W - 1;
            const F8 v = ld_bf16x8(base + pix * C1_COUT + cl * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_line[line][cl * 8 + k], acc[k]);
    }
    __syncthreads();
    for (int i = tid; i < 9 * C1_COUT; i += 256) {
        const int tap = i / C1_COUT, co = i % C1_COUT;
        const int kh = tap / 3, kw = tap % 3;
        float s = ps[b * C1_COUT + co];
        // tap kh == 0 reads input row h-1: pixel row 0 does not contribute; kh == 2: row H-1 does not
        const int er = kh == 0 ? 0 : (kh == 2 ? H - 1 : -1);
        const int ec = kw == 0 ? 0 : (kw == 2 ? W - 1 : -1);
        if (er >= 0) s -= s_line[kh == 0 ? 0 : 1][co];
        if (ec >= 0) s -= s_line[kw == 0 ? 2 : 3][co];
        if (er >= 0 && ec >= 0) s += __bfloat162float(base[(static_cast<size_t>(er) * W + ec) * C1_COUT + co]);
        S[(static_cast<size_t>(b) * 9 + tap) * C1_COUT + co] = s;
    }
}

// dW[co][3+c][tap] += sum_b e[b][c] * S[b][tap][co]      grid 64 (co), 256 threads (c)
__global__ void __launch_bounds__(256) fd_time_wgrad_kernel(const float* __restrict__ e, const float* __restrict__ S,
                                                            float* __restrict__ dw, int B) {
    const int co = blockIdx.x, c = threadIdx.x;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int b = 0; b < B; ++b) {
        const float ev = e[b * FD_TDIM + c];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = fmaf(ev, S[(static_cast<size_t>(b) * 9 + t) * C1_COUT + co], acc[t]);
    }
    float* d = dw + (static_cast<size_t>(co) * FD_CIN0 + FD_CIMG + c) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) d[t] += acc[t];
}

// de[b][c] = sum_{co,tap} W[co][3+c][tap] * S[b][tap][co]      grid B, 256 threads (c)
__global__ void __launch_bounds__(256) fd_time_dgrad_kernel(const float* __restrict__ wgt, const float* __restrict__ S,
                                                            float* __restrict__ de) {
    __shared__ float s_S[9][C1_COUT];
    const int b = blockIdx.x, c = threadIdx.x;
    for (int i = c; i < 9 * C1_COUT; i += 256) (&s_S[0][0])[i] = S[static_cast<size_t>(b) * 9 * C1_COUT + i];
    __syncthreads();
    float acc = 0.f;
    for (int co = 0; co < C1_COUT; ++co) {
        const float* wp = wgt + (static_cast<size_t>(co) * FD_CIN0 + FD_CIMG + c) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc = fmaf(wp[t], s_S[t][co], acc);
    }
    de[b * FD_TDIM + c] = acc;
}

// time_mlp backward, step 1: dh[b][k] = [hid > 0] * sum_j de[b][j] * W2[j][k]     grid B, 256 threads (k)
__global__ void __launch_bounds__(256) fd_time_mlp_bwd_dh_kernel(const float* __restrict__ de, const float* __restrict__ hid,
                                                                 const float* __restrict__ w2, float* __restrict__ dh) {
    __shared__ float s_de[FD_TDIM];
    const int b = blockIdx.x, k = threadIdx.x;
    s_de[k] = de[b * FD_TDIM + k];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < FD_TDIM; ++j) acc = fmaf(s_de[j], w2[j * FD_TDIM + k], acc);
    dh[b * FD_TDIM + k] = hid[b * FD_TDIM + k] > 0.f ? acc : 0.f;
}

// step 2: dW[j][k] += sum_b d[b][j] * in[b][k], db[j] += sum_b d[b][j]; blockIdx.y = 0: layer 2 (d = de, in = hid),
// 1: layer 1 (d = dh, in = emb). grid (256 rows, 2), 256 threads (k).
__global__ void __launch_bounds__(256) fd_time_mlp_bwd_w_kernel(const float* __restrict__ de, const float* __restrict__ dh,
                                                                const float* __restrict__ hid, const float* __restrict__ emb,
                                                                float* __restrict__ dw1, float* __restrict__ db1,
                                                                float* __restrict__ dw2, float* __restrict__ db2, int B) {
    const int j = blockIdx.x, k = threadIdx.x;
    const bool l2 = blockIdx.y == 0;
    const float* d = l2 ? de : dh;
    const float* in = l2 ? hid : emb;
    float acc = 0.f, bs = 0.f;
    for (int b = 0; b < B; ++b) {
        const float dv = d[b * FD_TDIM + j];
        acc = fmaf(dv, in[b * FD_TDIM + k], acc);
        bs += dv;
    }
    (l2 ? dw2 : dw1)[j * FD_TDIM + k] += acc;
    if (k == 0) (l2 ? db2 : db1)[j] += bs;
}

// ------------------------------------------------------------------------------------------------
// F.interpolate(scale_factor=2), mode 'nearest' (:579,582): dense NHWC -> channel slot of the decoder's concat
// buffer (replaces torch.cat, :580,583); backward = 2x2 block sum.
// ------------------------------------------------------------------------------------------------
// One thread = one input pixel x 8 channels: one 16-byte load, four 16-byte stores (32-bit index arithmetic).
__global__ void __launch_bounds__(256) fd_upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ in, int C,
                                                                __nv_bfloat16* __restrict__ out, int out_stride,
                                                                int out_coff, int h, int w, unsigned total) {
    const unsigned c8n = C >> 3;
    const unsigned W2 = 2 * w;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned c = (idx % c8n) * 8;
        unsigned r = idx / c8n;
        const unsigned ww = r % w;
        r /= w;
        const unsigned hh = r % h;
        const unsigned img = r / h;
        const uint4 v = *reinterpret_cast<const uint4*>(in + static_cast<size_t>(idx) * 8);
        __nv_bfloat16* o = out + ((static_cast<size_t>(img) * 2 * h + 2 * hh) * W2 + 2 * ww) * out_stride + out_coff + c;
        *reinterpret_cast<uint4*>(o) = v;
        *reinterpret_cast<uint4*>(o + out_stride) = v;
        *reinterpret_cast<uint4*>(o + static_cast<size_t>(W2) * out_stride) = v;
        *reinterpret_cast<uint4*>(o + static_cast<size_t>(W2 + 1) * out_stride) = v;
    }
}

__global__ void __launch_bounds__(256) fd_upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int dout_stride,
                                                                int dout_coff, int C, __nv_bfloat16* __restrict__ din,
                                                                int h, int w, unsigned total) {
    const unsigned c8n = C >> 3;
    const unsigned W2 = 2 * w;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned c = (idx % c8n) * 8;
        unsigned r = idx / c8n;
        const unsigned ww = r % w;
        r /= w;
        const unsigned hh = r % h;
        const unsigned img = r / h;
        const __nv_bfloat16* p = dout + ((static_cast<size_t>(img) * 2 * h + 2 * hh) * W2 + 2 * ww) * dout_stride + dout_coff + c;
        const F8 a = unpack8(ld_stream(p));
        const F8 b = unpack8(ld_stream(p + dout_stride));
        const F8 d = unpack8(ld_stream(p + static_cast<size_t>(W2) * dout_stride));
        const F8 e = unpack8(ld_stream(p + static_cast<size_t>(W2 + 1) * dout_stride));
        F8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = (a.v[k] + b.v[k]) + (d.v[k] + e.v[k]);
        st_bf16x8(din + static_cast<size_t>(idx) * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------
// diffusion elementwise: q_sample (:515-518) and one deterministic DDIM update (:629-633)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fd_q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                          const float2* __restrict__ coef, float* __restrict__ out,
                                                          int HW, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float2 c = coef[i / HW];
        out[i] = c.x * x0[i] + c.y * noise[i];
    }
}

// x0 = (x - sqrt(1-a)*eps)/sqrt(a);  x = sqrt(a_prev)*x0 + sqrt(1-a_prev)*eps;  clamp(-1,1) after the last step (:635)
__global__ void __launch_bounds__(256) fd_ddim_update_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                                             float sqrt_1ma, float sqrt_a, float sqrt_ap, float sqrt_1map,
                                                             int clamp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float ev = eps[i];
        const float x0 = (x[i] - sqrt_1ma * ev) / sqrt_a;
        float xn = sqrt_ap * x0 + sqrt_1map * ev;
        if (clamp) xn = fminf(fmaxf(xn, -1.f), 1.f);
        x[i] = xn;
    }
}

// ------------------------------------------------------------------------------------------------
// global-norm gradient clipping (torch.nn.utils.clip_grad_norm_, FastDDPM_Training_Fixed.ipynb cell 11: max_norm 1.0)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fd_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    float part = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        part = fmaf(g[i], g[i], part);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ float s_red[8];
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += s_red[w];
        atomicAdd(out, acc);
    }
}

// g *= min(1, max_norm / (pre_scale*sqrt(sumsq) + 1e-6))
__global__ void __launch_bounds__(256) fd_clip_scale_kernel(float* __restrict__ g, long long n,
                                                            const double* __restrict__ sumsq, float max_norm,
                                                            float pre_scale) {
    const float norm = static_cast<float>(sqrt(*sumsq)) * pre_scale;
    const float coef = fminf(max_norm / (norm + 1e-6f), 1.f);
    if (coef >= 1.f) return;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        g[i] *= coef;
}

}  // namespace b200sr
