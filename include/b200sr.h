/*
 * b200sr — C ABI of the B200-native (sm_100a) UNet slice-interpolation hot path.
 *
 * The reference (DeivanaiThiyagarajan/Multi-Image-Super-Resolution-for-Medical-Images) is pure
 * Python/PyTorch and has no FFI layer: the arithmetic of its hot path lives in torch.nn modules
 * (src/unet_model.py:22-118, :148-191). Each entry point below names the reference call site it replaces.
 * The Python host side (package `multi-image-super-resolution-for-medical-images_b200`, import name
 * `b200sr`) binds these with ctypes; see INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - every function returns 0 on success, a B200SR_E* code otherwise; b200sr_last_error() gives the text
 *     (thread local). There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - all pointers are DEVICE pointers owned by the caller; the library allocates nothing on the device and
 *     never synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - activations are NHWC bf16. A tensor may be a channel slot of a wider buffer, described by
 *     (ptr, pix_stride, c_off): element (b,h,w,c) lives at ptr[((b*H+h)*W+w)*pix_stride + c_off + c].
 *     pix_stride and c_off must be multiples of 8 (16-byte TMA / vector alignment), ptr 16-byte aligned.
 *   - H must be a multiple of 8 and W a multiple of 16 for the tensor-core entry points; channel counts on
 *     the tensor-core path are multiples of 64.
 */
#ifndef B200SR_H_
#define B200SR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B200SR_OK = 0,
    B200SR_EINVAL = 1,   /* bad shape / alignment / null pointer */
    B200SR_ECUDA = 2,    /* CUDA runtime or driver error (text in last_error) */
    B200SR_ENODEV = 3    /* no sm_100 device */
};

int b200sr_version(void);
const char* b200sr_last_error(void);

/* Launch accounting: number of kernels the last b200sr_*_wgrad_det call of this thread enqueued (split-K kernel, the slice
 * fold when it was needed, the layout kernel): 2 or 3. Used by bench.py's gpu_launches claim. */
int b200sr_last_wgrad_launches(void);
/* 0 when a compute-capability 10.x device is current, B200SR_ENODEV otherwise. */
int b200sr_device_ok(void);

/* ---- tensor-core implicit GEMMs (tcgen05 / TMEM / TMA) ------------------------------------------------ */

/* nn.Conv2d(Cin, Cout, 3, padding=1) forward — unet_model.py:27,30 (inside UNetBlock) — and, with the
 * dgrad-packed weights, its data gradient. x: (B,H,W,Cin) slot; w_packed: [Cout][9*Cin] bf16 from
 * b200sr_pack_jobs (PACK_CONV_FWD / PACK_CONV_DGRAD); out: (B,H,W,Cout) slot, raw conv result (no bias).
 * Optional epilogue: v = v*col_scale[c] + col_shift[c] (either may be NULL), ReLU if relu != 0 (eval-mode
 * folded BatchNorm, unet_model.py:28-29), and per-channel sum / sum-of-squares of the stored values
 * (train-mode BatchNorm statistics; NULL to skip) into stats[stats_replicas][2][Cout]:
 *   - stats_replicas >= number of SMs (148): DETERMINISTIC slot mode — every CTA stores its partial sums into its own
 *     slot, the unused slots are zeroed by the kernel, nothing has to be pre-zeroed and no atomics touch the data;
 *   - fewer replicas: legacy mode, partial sums are atomically ADDED into slot (CTA % stats_replicas) of a buffer the
 *     caller zeroed. The same rule holds for every entry point that takes (stats, stats_replicas).
 * Any B, H, W > 0: the kernel tiles pixels 16 x 8; edge tiles of other sizes reach past the image (TMA zero fill, clipped
 * stores, out-of-image pixels masked out of the statistics). The same holds for every tensor-core entry point below. */
int b200sr_conv3x3_fwd(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout,
                       int B, int H, int W, void* out, int out_pix_stride, int out_c_off,
                       const float* col_scale, const float* col_shift, int relu, float* stats,
                       int stats_replicas, void* stream);

/* Train-mode nn.Conv2d + nn.BatchNorm2d statistics + finalize in ONE launch (unet_model.py:27-28,30-31): the conv
 * as above with deterministic statistic slots (stats_replicas >= number of SMs), and b200sr_bn_finalize executed by the
 * last CTA of every 64..256-channel column block to finish (ticket counter per column block; slots summed in slot order).
 * Writes scale / shift / save_mean / save_invstd, updates running_mean / running_var / num_batches_tracked (all nullable
 * together). HOST struct, DEVICE pointers. counters: >= Cout/64 uint32, zero-initialised once, reset by the kernel. */
typedef struct b200sr_bn_train {
    const float* gamma;
    const float* beta;
    const float* conv_bias; /* nullable: re-added to running_mean only (BatchNorm cancels it in the output) */
    float* scale;
    float* shift;
    float* save_mean;
    float* save_invstd;
    float* running_mean;
    float* running_var;
    int64_t* num_batches_tracked;
    uint32_t* counters;
    double count; /* B*H*W */
    float eps;
    float momentum;
} b200sr_bn_train;
int b200sr_conv3x3_fwd_bn(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout, int B,
                          int H, int W, void* out, int out_pix_stride, int out_c_off, float* stats, int stats_replicas,
                          const b200sr_bn_train* bn, void* stream);

/* Same kernel; dy: (B,H,W,Cout) slot, w_packed: [Cin][9*Cout] (PACK_CONV_DGRAD), dx: (B,H,W,Cin) slot.
 * stats (optional) receives per-channel sums of dx (used for the ConvTranspose2d bias gradient). */
int b200sr_conv3x3_dgrad(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed,
                         int Cin, int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, float* stats,
                         int stats_replicas, void* stream);

/* Data gradient that also leaves the per-channel SUMS of its first `ncols` output channels in deterministic per-CTA slots
 * colsum_slots[slots][2][Cin] (row 0; finish with b200sr_sum_slots): the ConvTranspose2d bias gradient is the column sum
 * of the upsampled half of the decoder concat gradient (unet_model.py:101-113 backward). Cheaper than the full statistics
 * of b200sr_conv3x3_dgrad (no squares, no work for the skip half). slots >= number of SMs; ncols % 32 == 0. */
int b200sr_conv3x3_dgrad_colsum(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed, int Cin,
                                int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, float* colsum_slots,
                                int slots, int ncols, void* stream);

/* Data gradient fused with the ReLU backward of the layer it flows into (Conv+ReLU stacks without BatchNorm: the
 * DoubleConv of ModelLoader.py:521-533, the VGG features of the perceptual loss): dx = dgrad(dy) * [act > 0], act the
 * (B,H,W,Cin) activation slot of that layer. stats (optional) receives the per-channel sums of the stored dx, i.e. that
 * layer's bias gradient. */
int b200sr_conv3x3_dgrad_relu(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed, int Cin,
                              int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, const void* act,
                              int act_pix_stride, int act_c_off, float* stats, int stats_replicas, void* stream);

/* nn.ConvTranspose2d(Cin, Cout, 2, stride=2) forward — unet_model.py:67,70,73,76 — as a GEMM
 * [B*H*W, Cin] x [Cin, 4*Cout] with a pixel-shuffle scatter epilogue that writes (+bias) straight into a
 * channel slot of the decoder's concat buffer (this replaces torch.cat, unet_model.py:101,105,109,113).
 * x: (B,H,W,Cin) slot; w_packed: [4*Cout][Cin] (PACK_CONVT_FWD); out: (B,2H,2W,Cout) slot. */
int b200sr_convT2x2_fwd(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout,
                        const float* bias, int B, int H, int W, void* out, int out_pix_stride, int out_c_off,
                        void* stream);

/* ConvTranspose2d data gradient: dup: (B,2H,2W,Cout) slot, w_packed: [Cin][4*Cout] (PACK_CONVT_DGRAD),
 * dx: (B,H,W,Cin) slot. */
int b200sr_convT2x2_dgrad(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* w_packed,
                          int Cin, int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, void* stream);

/* Conv2d 3x3 weight gradient (split-K over pixels, MN-major UMMA operands). x: (B,H,W,Cin) slot,
 * dz: (B,H,W,Cout) slot; G: fp32 [9][Cin][Cout], ADDED into (caller zeroes it). */
int b200sr_conv3x3_wgrad(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz,
                         int dz_pix_stride, int dz_c_off, int Cout, int B, int H, int W, float* G, void* stream);

/* ConvTranspose2d weight gradient. dup: (B,2H,2W,Cout) slot, x: (B,H,W,Cin) slot;
 * G: fp32 [4][Cout][Cin], ADDED into. */
int b200sr_convT2x2_wgrad(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* x,
                          int x_pix_stride, int x_c_off, int Cin, int B, int H, int W, float* G, void* stream);

/* ---- bandwidth-bound kernels --------------------------------------------------------------------------- */

/* Table-driven weight packing / gradient unpacking: `jobs` is a DEVICE array of b200sr_pack_job. */
typedef struct b200sr_pack_job {
    const void* src;
    void* dst;
    int32_t kind; /* 0 conv fwd, 1 conv dgrad, 2 convT fwd, 3 convT dgrad, 4 conv wgrad unpack, 5 convT wgrad unpack,
                     6 conv1x1 fwd, 7 conv1x1 dgrad, 8 conv1x1 wgrad unpack, 9 conv fwd+dgrad (dst, count = second
                     destination), 10 convT fwd+dgrad likewise, 11 conv fwd bf16x3 split [w_hi|w_hi|w_lo], 12 convT fwd
                     bf16x3 split (fp32-accuracy eval mode) */
    int32_t cout;
    int32_t cin;
    int32_t pad;
    int64_t count; /* elements of dst; kinds 9 / 10: the dgrad-packing destination pointer */
} b200sr_pack_job;
int b200sr_pack_jobs(const b200sr_pack_job* jobs, int njobs, void* stream);

typedef struct b200sr_fold_job {
    const float* gamma;
    const float* beta;
    const float* running_mean;
    const float* running_var;
    const float* conv_bias; /* nullable */
    float* scale;
    float* shift;
    int32_t C;
    int32_t pad;
} b200sr_fold_job;
/* eval-mode BatchNorm fold (running stats + conv bias -> scale/shift), unet_model.py:28,31 in eval(). */
int b200sr_bn_fold_eval(const b200sr_fold_job* jobs, int njobs, float eps, void* stream);

/* First layer Conv2d(2,64,3,p=1) read directly from the fp32 NCHW network input (unet_model.py:49 -> :27).
 * x: (B,2,H,W) f32; w: (64,2,3,3) f32 (the parameter itself); out: (B,H,W,64) bf16 dense. */
int b200sr_conv1_fwd(const float* x, const float* w, const float* col_scale, const float* col_shift, int relu,
                     void* out, float* stats, int stats_replicas, int B, int H, int W, void* stream);
/* Data gradient of the first layer: dx (B,2,H,W) f32 written (not added). Needed when the network input carries a
 * gradient (Progressive UNet stages 2A/2B, ModelLoader.py:258-267). */
int b200sr_conv1_dgrad(const void* dz, const float* w, float* dx, int B, int H, int W, void* stream);
/* dW (64,2,3,3) f32, ADDED into. */
int b200sr_conv1_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream);

/* Train-mode nn.BatchNorm2d statistics -> scale/shift (+ saved mean/invstd, running-stat update with
 * momentum, unbiased variance; conv_bias re-added to running_mean; num_batches_tracked += 1 when not NULL).
 * unet_model.py:28,31. The `replicas` statistic slots are summed in a fixed order (double precision). */
int b200sr_bn_finalize(const float* stats, int replicas, int C, double count, const float* gamma,
                       const float* beta, const float* conv_bias, float eps, float momentum, float* scale,
                       float* shift, float* save_mean, float* save_invstd, float* running_mean,
                       float* running_var, int64_t* num_batches_tracked, void* stream);

/* BatchNorm-apply + ReLU (unet_model.py:28-29,31-32), optionally fused with MaxPool2d(2,2) (:52-61) and
 * writing the activation into a concat slot. z: (B,H,W,C) dense raw conv output. pooled may be NULL. */
int b200sr_bnrelu_apply(const void* z, int C, const float* scale, const float* shift, void* act,
                        int act_pix_stride, int act_c_off, void* pooled, int B, int H, int W, void* stream);

/* b200sr_bn_finalize + b200sr_bnrelu_apply in ONE launch (train-mode nn.BatchNorm2d + ReLU (+ MaxPool2d), unet_model.py:
 * 28-32,52-61): every thread derives the scale/shift of its channels from the statistic replicas; scale/shift/
 * save_mean/save_invstd are published for the backward pass and the running statistics are updated. C/8 must
 * divide 256 or be a multiple of 256. */
int b200sr_bn_train_apply(const void* z, int C, const float* stats, int replicas, double count, const float* gamma,
                          const float* beta, const float* conv_bias, float eps, float momentum, float* scale,
                          float* shift, float* save_mean, float* save_invstd, float* running_mean, float* running_var,
                          void* act, int act_pix_stride, int act_c_off, void* pooled, int B, int H, int W,
                          void* stream);

/* b200sr_bn_bwd_finalize + b200sr_bn_bwd_apply in ONE launch (C/8 must divide 256). */
int b200sr_bn_bwd_apply_fused(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C,
                              const float* scale, const float* shift, const float* mean, const float* invstd,
                              const float* sums, int replicas, double count, float* dgamma, float* dbeta, void* dz,
                              int64_t npix, void* stream);

/* nn.MaxPool2d(2,2) forward / backward (unet_model.py:52,55,58,61). Backward adds the skip-connection
 * gradient (dskip may be NULL) and routes to the first maximum in row-major window order like ATen. */
int b200sr_maxpool2x2_fwd(const void* in, int in_pix_stride, int in_c_off, int C, void* out, int B, int H,
                          int W, void* stream);
int b200sr_maxpool2x2_bwd(const void* act, int act_pix_stride, int act_c_off, const void* dpool,
                          const void* dskip, int dskip_pix_stride, int dskip_c_off, int C, void* dy, int B,
                          int H, int W, void* stream);

/* BatchNorm+ReLU backward. reduce: sums[replicas][2][C] += (sum g, sum g*xhat); finalize: c1 = S1/N,
 * c2 = S2/N, dgamma, dbeta; apply: dz = scale*(g - c1 - xhat*c2). */
int b200sr_bn_bwd_reduce(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C,
                         const float* scale, const float* shift, const float* mean, const float* invstd,
                         float* sums, int replicas, int64_t npix, void* stream);
int b200sr_bn_bwd_finalize(const float* sums, int replicas, int C, double count, float* c1, float* c2,
                           float* dgamma, float* dbeta, void* stream);
int b200sr_bn_bwd_apply(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C,
                        const float* scale, const float* shift, const float* mean, const float* invstd,
                        const float* c1, const float* c2, void* dz, int64_t npix, void* stream);

/* ---- perceptual (VGG feature) loss helpers: README.md:82-86 "MSE + perceptual (VGG) + SSIM"; SURVEY §8(f) row 2 ---- */
/* ReLU backward on bf16 tensors of n elements (n % 8 == 0): out = dy * [act > 0]. */
int b200sr_relu_bwd(const void* dy, const void* act, void* out, int64_t n, void* stream);
/* sums[0] (DEVICE double) += sum (fp-ft)^2 over n bf16 elements; grad (nullable) = gscale*(fp-ft)*[fp > 0]. */
int b200sr_feat_mse_grad(const void* fp, const void* ft, void* grad, double* sums, float gscale, int64_t n,
                         void* stream);

/* ---- DeepCNN residual baseline (reference src/ModelLoader.py:276-377; SURVEY §8(f) row 3, BASELINE configs[1]) ---- */
/* Conv2d(2,64,7,padding=3,bias=False) (:324) from the fp32 NCHW input; out (B,H,W,64) bf16; optional BN statistics. */
int b200sr_conv7_fwd(const float* x, const float* w, void* out, float* stats, int stats_replicas, int B, int H, int W,
                     void* stream);
/* dW (64,2,7,7) f32, ADDED into. */
int b200sr_conv7_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream);
/* nn.MaxPool2d(3, stride=1, padding=1) (:327) on dense NHWC bf16; backward routes to the first maximum (ATen). */
int b200sr_maxpool3x3_fwd(const void* in, void* out, int C, int B, int H, int W, void* stream);
int b200sr_maxpool3x3_bwd(const void* in, const void* dout, void* din, int C, int B, int H, int W, void* stream);
/* ResidualBlock tail (:290-307): out = relu(scale2*z2+shift2 + identity); identity = x (scale_d NULL) or
 * scale_d*z_d+shift_d (1x1 downsample + BatchNorm branch). Dense NHWC bf16, C/8 divides 256. */
int b200sr_bn_add_relu(const void* z2, const float* scale2, const float* shift2, const void* identity,
                       const float* scale_d, const float* shift_d, void* out, int C, int64_t npix, void* stream);
/* BatchNorm backward (reduce + finalize + apply) where the ReLU mask comes from a stored activation (mask_src > 0)
 * instead of scale*z+shift > 0: the ReLU of a residual block follows the skip addition. Dense tensors. */
int b200sr_bn_bwd_masked(const void* dy, const void* z, const void* mask_src, int C, const float* scale,
                         const float* shift, const float* mean, const float* invstd, float* sums, int replicas,
                         double count, float* dgamma, float* dbeta, void* dz, int64_t npix, void* stream);
/* out = a + b*[mask > 0] (mask NULL: a + b); n bf16 elements, n % 8 == 0. Gradient join of a residual block. */
int b200sr_add_masked(const void* a, const void* b, const void* mask, void* out, int64_t n, void* stream);
/* output_conv Conv2d(512,1,1)+bias (:336,375): fp32 (B,1,H,W) output, and its backward. C must be 512. */
int b200sr_headw_fwd(const void* act, int C, const float* w, const float* b, float* out, int64_t npix, void* stream);
int b200sr_headw_bwd(const float* dout, const void* act, int C, const float* w, void* dact, float* dw, float* db,
                     int64_t npix, void* stream);
/* Conv2d 1x1 (downsample branch, :347-351) forward and dgrad: D[pixel,n] = sum_c A[pixel,c]*Wp[n,c]; Wp from
 * b200sr_pack_jobs kind 6 (forward) / 7 (dgrad); wgrad into G[ci][co] (unpack kind 8). */
int b200sr_conv1x1(const void* a, int a_pix_stride, int a_c_off, int Ca, const void* w_packed, int N, int B, int H,
                   int W, void* out, int out_pix_stride, int out_c_off, float* stats, int stats_replicas, void* stream);
int b200sr_conv1x1_wgrad(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                         int dz_c_off, int Cout, int B, int H, int W, float* G, void* stream);

/* ---- Fast-DDPM denoiser (reference src/ModelLoader.py:471-636; SURVEY §8(f) row 4, BASELINE configs[4]) ---- */
/* sinusoidal_timestep_embedding (:475-487) + time_mlp Linear/ReLU/Linear (:547-551, :563-564). t: int64 [B];
 * w1,w2: (256,256) f32 row-major [out][in]; emb, hid (kept for backward) and e: [B][256] f32. */
int b200sr_fd_time_mlp_fwd(const int64_t* t, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* emb, float* hid, float* e, int B, void* stream);
/* Contribution of the spatially tiled time embedding (:565-568) to inc.block.0, as a per-sample bias for each of the 9
 * border classes: tb[b][cls][co] = bias[co] + sum_{taps inside the image for cls} sum_c w[co][3+c][tap]*e[b][c].
 * w: the (64,259,3,3) parameter itself; tb: [B][9][64] f32, cls = 3*rowclass + colclass (0 first, 1 interior, 2 last). */
int b200sr_fd_time_bias(const float* e, const float* w, const float* bias, float* tb, int B, void* stream);
/* inc.block.0 + ReLU (:521-527 as instantiated at :554): 3-channel direct conv + tb. Input channel 0 is
 * coef[b].x*x0 + coef[b].y*noise (q_sample fused, :597-599) when noise != NULL, else x0 (sampling, :624); channels
 * 1,2 = cond (B,2,H,W) f32. coef: DEVICE [B][2] f32. out: (B,H,W,64) bf16 dense, post-ReLU. */
int b200sr_fd_convin_fwd(const float* x0, const float* noise, const float* coef, const float* cond, const float* w,
                         const float* tb, void* out, int B, int H, int W, void* stream);
/* Weight gradient of the 3 image channels, ADDED into dw (64,259,3,3) f32 at [co][0..2][tap]. */
int b200sr_fd_convin_wgrad(const float* x0, const float* noise, const float* coef, const float* cond, const void* dz,
                           float* dw, int B, int H, int W, void* stream);
/* ReLU backward fused with the per-sample bias-gradient sums of a Conv+ReLU layer: dz = dy*[act > 0] (dense),
 * ps[b][c] += sum over the sample's pixels of dz. dy / act may be channel slots. C/8 must divide 256. */
int b200sr_fd_relu_bwd_bias(const void* dy, int dy_pix_stride, int dy_c_off, const void* act, int act_pix_stride,
                            int act_c_off, void* dz, float* ps, int C, int B, int H, int W, void* stream);
/* The same ReLU-mask + bias-sum fusion where the gradient is formed by a bandwidth-bound kernel (dz dense, ps ADDED into):
 * nearest-upsample backward (2x2 block sum of the (B,2h,2w,C) slot `dout`, masked by act (B,h,w,C) > 0); the 1x1 head
 * backward (outc, :558: dz = dout*w masked by act > 0, plus dw (64) and db (1) ADDED into); MaxPool2d(2,2) backward plus
 * the skip-connection gradient, masked by the ReLU of the pooled layer (act is both arg-max source and mask). */
int b200sr_fd_upsample2x_bwd_relu(const void* dout, int dout_pix_stride, int dout_c_off, const void* act, void* dz,
                                  float* ps, int C, int B, int h, int w, void* stream);
int b200sr_fd_head_bwd_relu(const float* dout, const void* act, const float* w, void* dz, float* dw, float* db, float* ps,
                            int B, int H, int W, void* stream);
int b200sr_fd_maxpool2x2_bwd_relu(const void* act, int act_pix_stride, int act_c_off, const void* dpool, const void* dskip,
                                  int dskip_pix_stride, int dskip_c_off, int C, void* dz, float* ps, int B, int H, int W,
                                  void* stream);
typedef struct b200sr_fd_bias_job {
    const float* ps; /* [rows][stride] partial sums */
    float* dst;      /* [C] bias gradient, ADDED into */
    int32_t C;
    int32_t rows;    /* 0: one row per sample (the B argument) */
    int32_t stride;  /* floats between rows; 0: C */
    int32_t pad;
} b200sr_fd_bias_job;
/* dst[c] += sum_b ps[b][c] for a DEVICE table of jobs (all conv biases of the network in one launch). */
int b200sr_fd_bias_finish(const b200sr_fd_bias_job* jobs, int njobs, int B, void* stream);
/* Backward of everything the time embedding touches: tap sums S[b][tap][co] of dz (B,H,W,64) from its border rows /
 * columns and ps; dw_in (64,259,3,3)[co][3+c][tap] += sum_b e*S; de = W^T S; time_mlp weight / bias gradients
 * (ADDED into). S [B][9][64], de and dh [B][256] are caller-provided scratch. */
int b200sr_fd_time_bwd(const void* dz, const float* ps, const float* e, const float* emb, const float* hid,
                       const float* w_in, const float* w2, float* S, float* de, float* dh, float* dw_in, float* dw1,
                       float* db1, float* dw2, float* db2, int B, int H, int W, void* stream);
/* F.interpolate(scale_factor=2) nearest (:579,582) written into a concat slot (replaces torch.cat :580,583), and its
 * backward (2x2 block sum). in / din: (B,h,w,C) dense; out / dout: (B,2h,2w,C) slot. */
int b200sr_fd_upsample2x_fwd(const void* in, int C, void* out, int out_pix_stride, int out_c_off, int B, int h, int w,
                             void* stream);
int b200sr_fd_upsample2x_bwd(const void* dout, int dout_pix_stride, int dout_c_off, int C, void* din, int B, int h, int w,
                             void* stream);
/* FastNoiseScheduler.q_sample (:515-518): out = coef[b].x*x0 + coef[b].y*noise over (B,1,H,W) f32. */
int b200sr_fd_q_sample(const float* x0, const float* noise, const float* coef, float* out, int B, int H, int W,
                       void* stream);
/* One deterministic DDIM update of FastDDPM.sample (:626-633), in place on x; clamp(-1,1) when clamp != 0 (:635). */
int b200sr_fd_ddim_update(float* x, const float* eps, float a_bar, float a_bar_prev, int clamp, int64_t n, void* stream);
/* torch.nn.utils.clip_grad_norm_ over a flat gradient buffer: g *= min(1, max_norm/(pre_scale*||g|| + 1e-6)).
 * sumsq: DEVICE double scratch (zeroed here). */
int b200sr_grad_clip(float* g, int64_t n, double* sumsq, float max_norm, float pre_scale, void* stream);

/* final nn.Conv2d(64,1,1) (unet_model.py:80,117): fp32 (B,1,H,W) output; and its backward. */
int b200sr_head_fwd(const void* act, const float* w, const float* b, float* out, int64_t npix, void* stream);
int b200sr_head_bwd(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                    int64_t npix, void* stream);

/* Fused MSE + windowed SSIM loss with gradient (nn.MSELoss, unet_model.py:156,180; SSIM per SURVEY §8 a11).
 * pred/target/grad: (B,1,H,W) f32 (grad may be NULL); sums: DEVICE double[2], ADDED into:
 * sum((x-y)^2) and sum(SSIM map). win: HOST pointer to K (<= 11) separable window taps.
 * grad = w_mse * dMSE/dpred + w_ssim * d(1 - mean SSIM)/dpred. */
int b200sr_mse_ssim(const float* pred, const float* target, float* grad, double* sums, int B, int H, int W,
                    const float* win, int K, float cov_norm, float C1, float C2, float w_mse, float w_ssim,
                    void* stream);

/* torch.optim.Adam step (unet_model.py:155,185) over flat fp32 buffers; grad is multiplied by grad_scale
 * (1/world_size after a sum all-reduce). `step` is the 1-based step count. */
int b200sr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* Same step with the step-dependent scalars read from DEVICE memory: bias_corr[0] = 1 - beta1^step,
 * bias_corr[1] = sqrt(1 - beta2^step). The launch itself is then step-independent and can be replayed from a CUDA graph. */
int b200sr_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                         float eps, const float* bias_corr, float grad_scale, void* stream);

/* ---- deterministic (bit-reproducible) reductions -------------------------------------------------------
 * Every cross-CTA reduction of the UNet train step has a variant that never orders floating-point additions by
 * scheduling: each CTA / split-K slice STORES its partial result into its own slot of a caller-provided workspace
 * `ws`, and a second stage (the last CTA to finish, chosen by a ticket counter, or a tiny follow-up kernel on the
 * same stream) adds the slots in slot order. Outputs are WRITTEN, not added: nothing needs pre-zeroing.
 * `counters`: DEVICE uint32, zero-initialised ONCE by the caller; the kernels reset them. A workspace / counter set
 * must not be shared by calls that may run concurrently on different streams. */

/* Conv2d 3x3 weight gradient (autograd of unet_model.py:27,30), split-K partials in ws, then reduced in split order and
 * written in the PyTorch parameter layout: dW[co][cin_off + ci][3][3] of a (Cout, cin_total, 3, 3) f32 tensor
 * (cin_total > Cin: a conv whose input channels are split over several launches). The split factor is capped by
 * ws_floats / (9*Cin*Cout). */
int b200sr_conv3x3_wgrad_det(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                             int dz_c_off, int Cout, int B, int H, int W, float* dW, int cin_total, int cin_off,
                             float* ws, int64_t ws_floats, void* stream);
/* ConvTranspose2d k2 s2 weight gradient (unet_model.py:67-76), dW written as (Cin, Cout, 2, 2) f32. */
int b200sr_convT2x2_wgrad_det(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* x,
                              int x_pix_stride, int x_c_off, int Cin, int B, int H, int W, float* dW, float* ws,
                              int64_t ws_floats, void* stream);
/* Conv2d 1x1 weight gradient, dW written as (Cout, Cin, 1, 1) f32. */
int b200sr_conv1x1_wgrad_det(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                             int dz_c_off, int Cout, int B, int H, int W, float* dW, float* ws, int64_t ws_floats,
                             void* stream);
/* First-layer weight gradient, dw (64,2,3,3) f32 WRITTEN; ws >= 1152 floats per CTA (up to 2 CTAs per SM). */
int b200sr_conv1_wgrad_det(const float* x, const void* dz, float* dw, int B, int H, int W, float* ws, int64_t ws_floats,
                           void* stream);
/* dst[i] = sum_s slots[s*slot_stride + i], s ascending, i < n. */
int b200sr_sum_slots(const float* slots, int nslots, int64_t slot_stride, int n, float* dst, void* stream);
/* BatchNorm+ReLU backward, pass 1: sums[0][c] = sum g, sums[1][c] = invstd[c] * sum g*(z-mean) WRITTEN (consume with
 * b200sr_bn_bwd_apply_fused(..., sums, replicas = 1, ...)). mask_src (nullable, dense (npix,C)): ReLU mask from a
 * stored activation instead of scale*z+shift > 0. ws: b200sr_bn_bwd_ws_floats(C) floats; counters: C/64 uint32. */
int64_t b200sr_bn_bwd_ws_floats(int C);
int b200sr_bn_bwd_reduce_det(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C, const float* scale,
                             const float* shift, const float* mean, const float* invstd, float* sums, float* ws,
                             int64_t ws_floats, uint32_t* counters, const void* mask_src, int64_t npix, void* stream);
/* b200sr_maxpool2x2_bwd (unet_model.py:52-61 backward + skip-connection add) fused with b200sr_bn_bwd_reduce_det of the
 * BatchNorm+ReLU the gradient flows into (the encoder block's conv.4/conv.5): dy (B,H,W,C) dense is written AND reduced
 * against z in the same pass; sums[2][C] as b200sr_bn_bwd_reduce_det writes them. C % 64 == 0. */
int b200sr_maxpool2x2_bwd_bnred(const void* act, int act_pix_stride, int act_c_off, const void* dpool, const void* dskip,
                                int dskip_pix_stride, int dskip_c_off, int C, void* dy, const void* z, const float* scale,
                                const float* shift, const float* mean, const float* invstd, float* sums, float* ws,
                                int64_t ws_floats, uint32_t* counters, int B, int H, int W, void* stream);
/* 1x1 head backward with dw (64) / db (1) WRITTEN; ws: 72 floats per CTA (up to 4 CTAs per SM); counter: 1 uint32. */
int b200sr_head_bwd_det(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                        int64_t npix, float* ws, int64_t ws_floats, uint32_t* counter, void* stream);
/* b200sr_head_bwd_det fused with b200sr_bn_bwd_reduce_det of the BatchNorm+ReLU feeding the head (dec1.conv.4/.5,
 * unet_model.py:76-80): dact (npix x 64) is written AND reduced against z in the same pass; sums[2][64] as
 * b200sr_bn_bwd_reduce_det writes them. ws: 200 floats per CTA; counters: 2 uint32, zero-initialised once. */
int b200sr_head_bwd_bnred(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                          const void* z, const float* scale, const float* shift, const float* mean, const float* invstd,
                          float* sums, int64_t npix, float* ws, int64_t ws_floats, uint32_t* counters, void* stream);
/* Fused MSE + SSIM loss: out3 = {loss, mse, mean SSIM} (DEVICE f32) WRITTEN by the last CTA; ws: 2 doubles per CTA
 * (B * ceil(H/32) * ceil(W/32) CTAs); counter: 1 uint32. Nothing is left for the host to combine. */
int b200sr_mse_ssim_det(const float* pred, const float* target, float* grad, float* out3, int B, int H, int W,
                        const float* win, int K, float cov_norm, float C1, float C2, float w_mse, float w_ssim,
                        double* ws, int64_t ws_doubles, uint32_t* counter, void* stream);
/* Adam with the step count resident on the device: step_dev = {completed steps, 0} (DEVICE int32[2], zero-initialised
 * or set to the resumed step count). The kernel derives the bias corrections of step step_dev[0]+1 itself and the last
 * CTA increments step_dev[0]: graph-replayable, and immune to a host that runs several steps ahead of the device. */
int b200sr_adam_step_auto(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, int32_t* step_dev, float grad_scale, void* stream);

/* ---- fp32-accuracy eval mode ---------------------------------------------------------------------------
 * The reference's inference path is an fp32 forward (src/VolumeVisualization.py:932-964; BASELINE configs[0]) and the
 * north star asks for rel-L2 <= 1e-4 in fp32/tf32. These entry points run the SAME tcgen05 bf16 main loop with
 * bf16x3 operand splitting: an fp32 activation is stored as (B,H,W,3C) bf16 [hi | lo | hi] (hi = bf16(v),
 * lo = bf16(v - hi)), a weight as [w_hi | w_hi | w_lo] along K (pack kinds 11 / 12), the accumulation is fp32 and
 * the epilogue (fp32 affine + ReLU) splits the result again. part_stride = channel distance between the three parts
 * of the OUTPUT slot (its logical channel count; 2C for a decoder concat buffer). Cin3 = 3 * logical Cin. */
int b200sr_conv3x3_fwd_split(const void* x, int x_pix_stride, int x_c_off, int Cin3, const void* w_packed, int Cout,
                             int B, int H, int W, void* out, int out_pix_stride, int out_c_off, int part_stride,
                             const float* col_scale, const float* col_shift, int relu, void* stream);
int b200sr_convT2x2_fwd_split(const void* x, int x_pix_stride, int x_c_off, int Cin3, const void* w_packed, int Cout,
                              const float* bias, int B, int H, int W, void* out, int out_pix_stride, int out_c_off,
                              int part_stride, void* stream);
/* First layer in fp32 FMAs from the fp32 NCHW input; out: (B,H,W,192) split. */
int b200sr_conv1_fwd_split(const float* x, const float* w, const float* col_scale, const float* col_shift, int relu,
                           void* out, int B, int H, int W, void* stream);
/* MaxPool2d(2,2) of a split channel slot (parts in_part_stride apart) -> dense split (B,H/2,W/2,3C). */
int b200sr_maxpool2x2_fwd_split(const void* in, int in_pix_stride, int in_c_off, int in_part_stride, int C, void* out,
                                int B, int H, int W, void* stream);
/* final nn.Conv2d(64,1,1) on a split (B,H,W,192) activation, fp32 output. */
int b200sr_head_fwd_split(const void* act, const float* w, const float* b, float* out, int64_t npix, void* stream);

/* Evaluation metrics of the reference (compute_metrics, src/VolumeVisualization.py:237-269) on the device: min-max
 * normalisation by the ORIGINAL volume's range (+1e-8), prediction clipped to [0,1], per-slice SSIM (7x7 uniform window,
 * sample covariance: the skimage defaults the reference calls) and PSNR (data_range 1), MAE over the volume.
 * original / predicted / orig_norm / pred_norm: (S,H,W) f32; per_slice: [S][2] = {SSIM, PSNR};
 * out5 = {ssim_mean, ssim_std, psnr_mean, psnr_std, mae} (population std, like np.std).
 * ws: 2048 + 4*S*ceil(H/32)*ceil(W/32) doubles; counters: 2 zero-initialised uint32. */
int b200sr_volume_metrics(const float* original, const float* predicted, int S, int H, int W, float* orig_norm,
                          float* pred_norm, float* out5, float* per_slice, double* ws, int64_t ws_doubles,
                          uint32_t* counters, void* stream);

/* layout casts at the boundary */
int b200sr_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, void* stream);
int b200sr_nhwc_bf16_to_nchw_f32(const void* in, int in_pix_stride, int in_c_off, float* out, int B, int C,
                                 int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SR_H_ */
