// Fused MSE + windowed-SSIM loss, forward value AND gradient w.r.t. the prediction in one pass (fp32).
//
//   loss = w_mse * mean((x-y)^2) + w_ssim * (1 - mean(SSIM_map(x, y)))
//
// Reference: nn.MSELoss (unet_model.py:156,180); the combined-loss notebook is missing from the reference
// snapshot, so the SSIM definition is the one frozen in SURVEY.md §8(a11):
//   mode G: 11-tap Gaussian (sigma 1.5) separable window, "valid" map, biased covariance
//   mode U: 7-tap uniform window, "valid" map, sample covariance (x K^2/(K^2-1)) == skimage defaults used by
//           the reference's evaluation code (VolumeVisualization.py:256)
// The window taps and the covariance normalisation are kernel arguments, so both modes are the same code.
//
// With mu = w*x, m_xx = w*x^2, m_xy = w*x*y (valid correlations) and S(mu_x, mu_y, m_xx, m_yy, m_xy) the SSIM
// map, the gradient is three "full" correlations of per-map-pixel coefficient maps with the same window:
//   dS/dx(q) = sum_p w(q-p) [ Gm(p) + 2 x(q) Gxx(p) + y(q) Gxy(p) ]
// Each block owns a 32x32 tile of prediction pixels, stages x,y with a 2(K-1) halo in shared memory, and runs
// the separable filters out of shared memory; loss partial sums leave through two double atomics per block.
#pragma once
#include "elementwise.cuh"

namespace b200sr {

constexpr int LS_T = 32;
constexpr int LS_KMAX = 11;
constexpr int LS_IT = LS_T + 2 * (LS_KMAX - 1);  // 52
constexpr int LS_MT = LS_T + (LS_KMAX - 1);      // 42
constexpr int LS_SMEM_FLOATS = 2 * LS_IT * LS_IT + 5 * LS_IT * LS_MT + 3 * LS_MT * LS_MT;
constexpr int LS_SMEM_BYTES = LS_SMEM_FLOATS * 4;

struct LossArgs {
    const float* pred;    // [B][H][W]
    const float* target;  // [B][H][W]
    float* grad;          // [B][H][W] or null
    double* sums;         // [2]: sum (x-y)^2, sum SSIM map
    int H, W, K;
    float win[LS_KMAX];
    float cov_norm, C1, C2;
    float g_mse;   // w_mse * 2 / (B*H*W)
    float g_ssim;  // -w_ssim / (B * (H-K+1) * (W-K+1))
    // deterministic finish (partials != NULL): block b stores its two partial sums into partials[b][2]; the last block
    // (ticket) adds them in block order and writes out = {loss, mse, mean SSIM} — no atomics, no host-side arithmetic
    double* partials;
    unsigned* counter;
    float* out;
    double inv_n_mse, inv_n_ssim;
    float w_mse, w_ssim;
    // evaluation-metrics mode (per_slice != NULL; reference compute_metrics, src/VolumeVisualization.py:237-269): partial
    // slots hold {sum sq err, sum SSIM, sum abs err}; the last block turns them into per-slice {SSIM, PSNR} and
    // out = {ssim_mean, ssim_std, psnr_mean, psnr_std, mae}. blocks_per_image = tiles of one slice (consecutive blocks).
    float* per_slice;       // [B][2]
    int blocks_per_image;
    int nimages;
};

__device__ __forceinline__ void metrics_finish(const LossArgs& a, const float (&s_red)[3][8], int tid) {
    if (tid < 3) {
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += s_red[tid][w];
        a.partials[static_cast<size_t>(blockIdx.x) * 4 + tid] = acc;
    }
    if (!last_block_ticket(a.counter, gridDim.x)) return;
    // one thread per slice: its tiles' partials in tile order (fixed order -> reproducible)
    __shared__ double s_abs[256];
    double abs_acc = 0.0;
    for (int img = tid; img < a.nimages; img += 256) {
        double sq = 0.0, ss = 0.0, ab = 0.0;
        const double* p = a.partials + static_cast<size_t>(img) * a.blocks_per_image * 4;
        for (int b = 0; b < a.blocks_per_image; ++b) {
            sq += __ldcg(p + 4 * b);
            ss += __ldcg(p + 4 * b + 1);
            ab += __ldcg(p + 4 * b + 2);
        }
        const double pix = static_cast<double>(a.H) * a.W;
        a.per_slice[2 * img] = static_cast<float>(ss / (static_cast<double>(a.H - a.K + 1) * (a.W - a.K + 1)));
        a.per_slice[2 * img + 1] = static_cast<float>(10.0 * log10(1.0 / (sq / pix)));  // PSNR, data_range 1.0
        abs_acc += ab;
    }
    s_abs[tid] = abs_acc;
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
        double mae = 0.0;
        for (int l = 0; l < 256; ++l) mae += s_abs[l];
        double m[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
        for (int i = 0; i < a.nimages; ++i) {
            m[0] += a.per_slice[2 * i];
            m[1] += a.per_slice[2 * i + 1];
        }
        m[0] /= a.nimages;
        m[1] /= a.nimages;
        for (int i = 0; i < a.nimages; ++i) {
            v[0] += (a.per_slice[2 * i] - m[0]) * (a.per_slice[2 * i] - m[0]);
            v[1] += (a.per_slice[2 * i + 1] - m[1]) * (a.per_slice[2 * i + 1] - m[1]);
        }
        a.out[0] = static_cast<float>(m[0]);
        a.out[1] = static_cast<float>(sqrt(v[0] / a.nimages));  // np.std: population standard deviation
        a.out[2] = static_cast<float>(m[1]);
        a.out[3] = static_cast<float>(sqrt(v[1] / a.nimages));
        a.out[4] = static_cast<float>(mae / (static_cast<double>(a.nimages) * a.H * a.W));
    }
}

__device__ __forceinline__ void loss_finish(const LossArgs& a, const float (&s_red)[3][8], int tid) {
    if (a.per_slice != nullptr) return metrics_finish(a, s_red, tid);
    if (a.partials == nullptr) {
        if (tid < 2) {
            double acc = 0.0;
            for (int w = 0; w < 8; ++w) acc += s_red[tid][w];
            atomicAdd(a.sums + tid, acc);
        }
        return;
    }
    if (tid < 2) {
        double acc = 0.0;
        for (int w = 0; w < 8; ++w) acc += s_red[tid][w];
        a.partials[static_cast<size_t>(blockIdx.x) * 2 + tid] = acc;
    }
    if (!last_block_ticket(a.counter, gridDim.x)) return;
    __shared__ double s_t[2][128];
    const int which = tid & 1, sl = tid >> 1;  // 2 sums x 128 block-lanes
    double acc = 0.0;
    for (unsigned b = sl; b < gridDim.x; b += 128) acc += __ldcg(a.partials + static_cast<size_t>(b) * 2 + which);
    s_t[which][sl] = acc;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0, q = 0.0;
        for (int l = 0; l < 128; ++l) {
            m += s_t[0][l];
            q += s_t[1][l];
        }
        const double mse = m * a.inv_n_mse, ssim = q * a.inv_n_ssim;
        a.out[0] = static_cast<float>(a.w_mse * mse + a.w_ssim * (1.0 - ssim));
        a.out[1] = static_cast<float>(mse);
        a.out[2] = static_cast<float>(ssim);
        if (a.sums != nullptr) {
            a.sums[0] = m;
            a.sums[1] = q;
        }
    }
}

__global__ void __launch_bounds__(256) mse_ssim_kernel(const LossArgs a) {
    extern __shared__ float ls_smem[];
    const int K = a.K, R = K - 1;
    const int IT = LS_T + 2 * R, MT = LS_T + R;
    float* s_x = ls_smem;                       // [IT][IT]
    float* s_y = s_x + LS_IT * LS_IT;           // [IT][IT]
    float* s_h = s_y + LS_IT * LS_IT;           // 5 x [IT][MT]; later reused as 3 x [MT][T]
    float* s_g = s_h + 5 * LS_IT * LS_MT;       // 3 x [MT][MT]
    __shared__ float s_win[LS_KMAX];
    __shared__ float s_red[3][8];

    const int tid = threadIdx.x;
    const int tiles_w = (a.W + LS_T - 1) / LS_T;
    const int tiles_h = (a.H + LS_T - 1) / LS_T;
    const int img = blockIdx.x / (tiles_w * tiles_h);
    const int t_in = blockIdx.x % (tiles_w * tiles_h);
    const int qh0 = (t_in / tiles_w) * LS_T, qw0 = (t_in % tiles_w) * LS_T;
    const int ih0 = qh0 - R, iw0 = qw0 - R;
    const float* px = a.pred + static_cast<size_t>(img) * a.H * a.W;
    const float* py = a.target + static_cast<size_t>(img) * a.H * a.W;
    const int MH = a.H - K + 1, MW = a.W - K + 1;  // valid map size

    if (tid < K) s_win[tid] = a.win[tid];
    for (int i = tid; i < IT * IT; i += 256) {
        const int r = i / IT, c = i % IT;
        const int hh = ih0 + r, ww = iw0 + c;
        float vx = 0.f, vy = 0.f;
        if (hh >= 0 && hh < a.H && ww >= 0 && ww < a.W) {
            vx = px[hh * a.W + ww];
            vy = py[hh * a.W + ww];
        }
        s_x[r * LS_IT + c] = vx;
        s_y[r * LS_IT + c] = vy;
    }
    __syncthreads();

    // horizontal pass of the 5 moment images
    for (int i = tid; i < IT * MT; i += 256) {
        const int r = i / MT, b = i % MT;
        float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
        for (int v = 0; v < K; ++v) {
            const float wv = s_win[v];
            const float x = s_x[r * LS_IT + b + v], y = s_y[r * LS_IT + b + v];
            hx = fmaf(wv, x, hx);
            hy = fmaf(wv, y, hy);
            hxx = fmaf(wv, x * x, hxx);
            hyy = fmaf(wv, y * y, hyy);
            hxy = fmaf(wv, x * y, hxy);
        }
        float* d = s_h + r * LS_MT + b;
        d[0 * LS_IT * LS_MT] = hx;
        d[1 * LS_IT * LS_MT] = hy;
        d[2 * LS_IT * LS_MT] = hxx;
        d[3 * LS_IT * LS_MT] = hyy;
        d[4 * LS_IT * LS_MT] = hxy;
    }
    __syncthreads();

    // vertical pass -> SSIM map value + gradient coefficient maps
    float ssim_part = 0.f;
    for (int i = tid; i < MT * MT; i += 256) {
        const int ar = i / MT, b = i % MT;
        const int ph = qh0 - R + ar, pw = qw0 - R + b;
        float gm = 0.f, gxx = 0.f, gxy = 0.f;
        if (ph >= 0 && ph < MH && pw >= 0 && pw < MW) {
            float mx = 0.f, my = 0.f, mxx = 0.f, myy = 0.f, mxy = 0.f;
            for (int u = 0; u < K; ++u) {
                const float wu = s_win[u];
                const float* s = s_h + (ar + u) * LS_MT + b;
                mx = fmaf(wu, s[0 * LS_IT * LS_MT], mx);
                my = fmaf(wu, s[1 * LS_IT * LS_MT], my);
                mxx = fmaf(wu, s[2 * LS_IT * LS_MT], mxx);
                myy = fmaf(wu, s[3 * LS_IT * LS_MT], myy);
                mxy = fmaf(wu, s[4 * LS_IT * LS_MT], mxy);
            }
            const float cn = a.cov_norm;
            const float sxx = cn * (mxx - mx * mx), syy = cn * (myy - my * my), sxy = cn * (mxy - mx * my);
            const float A1 = 2.f * mx * my + a.C1, A2 = 2.f * sxy + a.C2;
            const float B1 = mx * mx + my * my + a.C1, B2 = sxx + syy + a.C2;
            const float inv = 1.f / (B1 * B2);
            const float S = A1 * A2 * inv;
            // partial derivatives of S
            const float dS_dsxy = 2.f * A1 * inv;
            const float dS_dsxx = -S / B2;
            gxx = cn * dS_dsxx;
            gxy = cn * dS_dsxy;
            gm = 2.f * my * A2 * inv - 2.f * mx * S / B1 - cn * my * dS_dsxy - 2.f * cn * mx * dS_dsxx;
            if (ar >= R && b >= R) ssim_part += S;  // map pixels owned by this block
        }
        s_g[0 * LS_MT * LS_MT + ar * LS_MT + b] = gm;
        s_g[1 * LS_MT * LS_MT + ar * LS_MT + b] = gxx;
        s_g[2 * LS_MT * LS_MT + ar * LS_MT + b] = gxy;
    }
    __syncthreads();

    // horizontal pass of the coefficient maps (full correlation): Hg[a][j] = sum_v w[v] G[a][j + R - v]
    float* s_hg = s_h;  // 3 x [MT][T], the moment rows are dead
    for (int i = tid; i < MT * LS_T; i += 256) {
        const int ar = i / LS_T, j = i % LS_T;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        for (int v = 0; v < K; ++v) {
            const float wv = s_win[v];
            const int b = j + R - v;
            h0 = fmaf(wv, s_g[0 * LS_MT * LS_MT + ar * LS_MT + b], h0);
            h1 = fmaf(wv, s_g[1 * LS_MT * LS_MT + ar * LS_MT + b], h1);
            h2 = fmaf(wv, s_g[2 * LS_MT * LS_MT + ar * LS_MT + b], h2);
        }
        s_hg[0 * LS_MT * LS_T + ar * LS_T + j] = h0;
        s_hg[1 * LS_MT * LS_T + ar * LS_T + j] = h1;
        s_hg[2 * LS_MT * LS_T + ar * LS_T + j] = h2;
    }
    __syncthreads();

    // vertical pass + combine with the MSE term
    float mse_part = 0.f, abs_part = 0.f;
    for (int i = tid; i < LS_T * LS_T; i += 256) {
        const int r = i / LS_T, j = i % LS_T;
        const int qh = qh0 + r, qw = qw0 + j;
        if (qh >= a.H || qw >= a.W) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        for (int u = 0; u < K; ++u) {
            const float wu = s_win[u];
            const int ar = r + R - u;
            v0 = fmaf(wu, s_hg[0 * LS_MT * LS_T + ar * LS_T + j], v0);
            v1 = fmaf(wu, s_hg[1 * LS_MT * LS_T + ar * LS_T + j], v1);
            v2 = fmaf(wu, s_hg[2 * LS_MT * LS_T + ar * LS_T + j], v2);
        }
        const float x = s_x[(r + R) * LS_IT + j + R], y = s_y[(r + R) * LS_IT + j + R];
        const float d = x - y;
        mse_part = fmaf(d, d, mse_part);
        abs_part += fabsf(d);
        if (a.grad != nullptr)
            a.grad[(static_cast<size_t>(img) * a.H + qh) * a.W + qw] =
                a.g_mse * d + a.g_ssim * (v0 + 2.f * x * v1 + y * v2);
    }

    // block reduction of the two loss partials
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mse_part += __shfl_xor_sync(0xffffffffu, mse_part, o);
        ssim_part += __shfl_xor_sync(0xffffffffu, ssim_part, o);
        abs_part += __shfl_xor_sync(0xffffffffu, abs_part, o);
    }
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = mse_part;
        s_red[1][tid >> 5] = ssim_part;
        s_red[2][tid >> 5] = abs_part;
    }
    __syncthreads();
    loss_finish(a, s_red, tid);
}

// ------------------------------------------------------------------------------------------------
// Register-blocked version for the two frozen window sizes (K = 11 Gaussian, K = 7 uniform). Same algorithm and
// tile as above; every separable pass gives a thread a strip of consecutive outputs along the filter direction,
// so each shared-memory value is loaded once per strip instead of once per tap (LDS per pixel 218 -> ~50; the
// generic kernel was LDS-bound at ~180 GB/s algorithmic). Odd row pitches keep column-parallel accesses
// conflict-free. After this the kernel is FMA-bound (~290 FMA per pixel incl. halo), not HBM-bound.
// ------------------------------------------------------------------------------------------------
template <int K>
struct LsFast {
    static constexpr int R = K - 1;
    static constexpr int IT = LS_T + 2 * R;  // input tile edge
    static constexpr int MT = LS_T + R;      // SSIM-map tile edge
    static constexpr int PX = IT | 1;        // odd pitches
    static constexpr int PM = MT | 1;
    static constexpr int PT = LS_T + 1;
    static constexpr int SMEM_FLOATS = 2 * IT * PX + 5 * IT * PM + 3 * MT * PM;
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

template <int K>
__global__ void __launch_bounds__(256, 2) mse_ssim_fast_kernel(const LossArgs a) {
    using L = LsFast<K>;
    constexpr int R = L::R, IT = L::IT, MT = L::MT, PX = L::PX, PM = L::PM, PT = L::PT;
    extern __shared__ float ls_smem[];
    float* s_x = ls_smem;             // [IT][PX]
    float* s_y = s_x + IT * PX;       // [IT][PX]
    float* s_h = s_y + IT * PX;       // 5 x [IT][PM]; later reused as 3 x [MT][PT]
    float* s_g = s_h + 5 * IT * PM;   // 3 x [MT][PM]
    float* s_hg = s_h;
    __shared__ float s_red[3][8];

    const int tid = threadIdx.x;
    const int tiles_w = (a.W + LS_T - 1) / LS_T;
    const int tiles_h = (a.H + LS_T - 1) / LS_T;
    const int img = blockIdx.x / (tiles_w * tiles_h);
    const int t_in = blockIdx.x % (tiles_w * tiles_h);
    const int qh0 = (t_in / tiles_w) * LS_T, qw0 = (t_in % tiles_w) * LS_T;
    const int ih0 = qh0 - R, iw0 = qw0 - R;
    const float* px = a.pred + static_cast<size_t>(img) * a.H * a.W;
    const float* py = a.target + static_cast<size_t>(img) * a.H * a.W;
    const int MH = a.H - K + 1, MW = a.W - K + 1;  // valid map size
    float w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] = a.win[i];

    for (int i = tid; i < IT * IT; i += 256) {
        const int r = i / IT, c = i % IT;
        const int hh = ih0 + r, ww = iw0 + c;
        float vx = 0.f, vy = 0.f;
        if (hh >= 0 && hh < a.H && ww >= 0 && ww < a.W) {
            vx = __ldg(px + hh * a.W + ww);
            vy = __ldg(py + hh * a.W + ww);
        }
        s_x[r * PX + c] = vx;
        s_y[r * PX + c] = vy;
    }
    __syncthreads();

    // ---- pass H: 5 moment images, strips of 6 outputs along a row; consecutive threads take consecutive rows ----
    {
        constexpr int SW = 6, NS = (MT + SW - 1) / SW;
        for (int item = tid; item < IT * NS; item += 256) {
            const int r = item % IT, b0 = (item / IT) * SW;
            float xi[SW + R], yi[SW + R];
#pragma unroll
            for (int i = 0; i < SW + R; ++i) {
                const int c = min(b0 + i, IT - 1);
                xi[i] = s_x[r * PX + c];
                yi[i] = s_y[r * PX + c];
            }
#pragma unroll
            for (int o = 0; o < SW; ++o) {
                float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
#pragma unroll
                for (int v = 0; v < K; ++v) {
                    const float x = xi[o + v], y = yi[o + v], wx = w[v] * x, wy = w[v] * y;
                    hx += wx;
                    hy += wy;
                    hxx = fmaf(wx, x, hxx);
                    hyy = fmaf(wy, y, hyy);
                    hxy = fmaf(wx, y, hxy);
                }
                if (b0 + o < MT) {
                    float* d = s_h + r * PM + b0 + o;
                    d[0 * IT * PM] = hx;
                    d[1 * IT * PM] = hy;
                    d[2 * IT * PM] = hxx;
                    d[3 * IT * PM] = hyy;
                    d[4 * IT * PM] = hxy;
                }
            }
        }
    }
    __syncthreads();

    // ---- pass V: vertical filter -> SSIM value + gradient coefficient maps; strips of 6 rows per column ----
    float ssim_part = 0.f;
    {
        constexpr int SV = 6, NS = (MT + SV - 1) / SV;
        for (int item = tid; item < MT * NS; item += 256) {
            const int b = item % MT, ar0 = (item / MT) * SV;
            float m[5][SV];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float val[SV + R];
#pragma unroll
                for (int i = 0; i < SV + R; ++i) val[i] = s_h[q * IT * PM + min(ar0 + i, IT - 1) * PM + b];
#pragma unroll
                for (int o = 0; o < SV; ++o) {
                    float acc = 0.f;
#pragma unroll
                    for (int u = 0; u < K; ++u) acc = fmaf(w[u], val[o + u], acc);
                    m[q][o] = acc;
                }
            }
#pragma unroll
            for (int o = 0; o < SV; ++o) {
                const int ar = ar0 + o;
                if (ar >= MT) break;
                const int ph = qh0 - R + ar, pw = qw0 - R + b;
                float gm = 0.f, gxx = 0.f, gxy = 0.f;
                if (ph >= 0 && ph < MH && pw >= 0 && pw < MW) {
                    const float mx = m[0][o], my = m[1][o], mxx = m[2][o], myy = m[3][o], mxy = m[4][o];
                    const float cn = a.cov_norm;
                    const float sxx = cn * (mxx - mx * mx), syy = cn * (myy - my * my), sxy = cn * (mxy - mx * my);
                    const float A1 = 2.f * mx * my + a.C1, A2 = 2.f * sxy + a.C2;
                    const float B1 = mx * mx + my * my + a.C1, B2 = sxx + syy + a.C2;
                    const float inv = 1.f / (B1 * B2);
                    const float S = A1 * A2 * inv;
                    const float dS_dsxy = 2.f * A1 * inv;
                    const float dS_dsxx = -S / B2;
                    gxx = cn * dS_dsxx;
                    gxy = cn * dS_dsxy;
                    gm = 2.f * my * A2 * inv - 2.f * mx * S / B1 - cn * my * dS_dsxy - 2.f * cn * mx * dS_dsxx;
                    if (ar >= R && b >= R) ssim_part += S;  // map pixels owned by this block
                }
                s_g[0 * MT * PM + ar * PM + b] = gm;
                s_g[1 * MT * PM + ar * PM + b] = gxx;
                s_g[2 * MT * PM + ar * PM + b] = gxy;
            }
        }
    }
    __syncthreads();

    // ---- pass H2: full correlation of the coefficient maps along rows, strips of 8 outputs ----
    {
        constexpr int SW = 8, NS = LS_T / SW;
        for (int item = tid; item < MT * NS; item += 256) {
            const int ar = item % MT, j0 = (item / MT) * SW;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float val[SW + R];
#pragma unroll
                for (int i = 0; i < SW + R; ++i) val[i] = s_g[q * MT * PM + ar * PM + j0 + i];
#pragma unroll
                for (int o = 0; o < SW; ++o) {
                    float acc = 0.f;
#pragma unroll
                    for (int t = 0; t <= R; ++t) acc = fmaf(w[R - t], val[o + t], acc);
                    s_hg[q * MT * PT + ar * PT + j0 + o] = acc;
                }
            }
        }
    }
    __syncthreads();

    // ---- pass V2: along columns, strips of 8 rows, then combine with the MSE term ----
    float mse_part = 0.f, abs_part = 0.f;
    {
        constexpr int SV = 8, NS = LS_T / SV;
        for (int item = tid; item < LS_T * NS; item += 256) {
            const int j = item % LS_T, r0 = (item / LS_T) * SV;
            float v3[3][SV];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float val[SV + R];
#pragma unroll
                for (int i = 0; i < SV + R; ++i) val[i] = s_hg[q * MT * PT + (r0 + i) * PT + j];
#pragma unroll
                for (int o = 0; o < SV; ++o) {
                    float acc = 0.f;
#pragma unroll
                    for (int t = 0; t <= R; ++t) acc = fmaf(w[R - t], val[o + t], acc);
                    v3[q][o] = acc;
                }
            }
#pragma unroll
            for (int o = 0; o < SV; ++o) {
                const int r = r0 + o;
                const int qh = qh0 + r, qw = qw0 + j;
                if (qh >= a.H || qw >= a.W) continue;
                const float x = s_x[(r + R) * PX + j + R], y = s_y[(r + R) * PX + j + R];
                const float d = x - y;
                mse_part = fmaf(d, d, mse_part);
                abs_part += fabsf(d);
                if (a.grad != nullptr)
                    a.grad[(static_cast<size_t>(img) * a.H + qh) * a.W + qw] =
                        a.g_mse * d + a.g_ssim * (v3[0][o] + 2.f * x * v3[1][o] + y * v3[2][o]);
            }
        }
    }

#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mse_part += __shfl_xor_sync(0xffffffffu, mse_part, o);
        ssim_part += __shfl_xor_sync(0xffffffffu, ssim_part, o);
        abs_part += __shfl_xor_sync(0xffffffffu, abs_part, o);
    }
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = mse_part;
        s_red[1][tid >> 5] = ssim_part;
        s_red[2][tid >> 5] = abs_part;
    }
    __syncthreads();
    loss_finish(a, s_red, tid);
}

// ------------------------------------------------------------------------------------------------
// Evaluation metrics on the device (reference compute_metrics, src/VolumeVisualization.py:237-269): min / max of the
// original volume, min-max normalisation by the ORIGINAL range with the prediction clipped to [0,1]; the per-slice
// SSIM (7x7 uniform window, sample covariance = skimage defaults) / PSNR / MAE come from the loss kernel above in its
// metrics mode.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) volume_minmax_kernel(const float* __restrict__ x, long long n,
                                                            float* __restrict__ partial, unsigned* __restrict__ counter,
                                                            float* __restrict__ out2) {
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = __ldg(x + i);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            lo = fminf(lo, s_lo[w]);
            hi = fmaxf(hi, s_hi[w]);
        }
        partial[2 * blockIdx.x] = lo;
        partial[2 * blockIdx.x + 1] = hi;
    }
    if (!last_block_ticket(counter, gridDim.x)) return;
    lo = INFINITY;
    hi = -INFINITY;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) {
        lo = fminf(lo, __ldcg(partial + 2 * b));
        hi = fmaxf(hi, __ldcg(partial + 2 * b + 1));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            lo = fminf(lo, s_lo[w]);
            hi = fmaxf(hi, s_hi[w]);
        }
        out2[0] = lo;
        out2[1] = hi;
    }
}

__global__ void __launch_bounds__(256) volume_normalize_kernel(const float* __restrict__ orig,
                                                               const float* __restrict__ pred,
                                                               const float* __restrict__ mm,
                                                               float* __restrict__ orig_norm,
                                                               float* __restrict__ pred_norm, long long n) {
    const float lo = mm[0];
    const float range = mm[1] - lo + 1e-8f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        orig_norm[i] = (orig[i] - lo) / range;
        pred_norm[i] = fminf(fmaxf((pred[i] - lo) / range, 0.f), 1.f);
    }
}

}  // namespace b200sr
