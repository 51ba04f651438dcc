// C ABI of the b200sr hot path (see include/b200sr.h). Host-side argument checks, TMA descriptor cache,
// kernel launches. No device allocation, no synchronisation, no CPU fallback.
#include "../../include/b200sr.h"

#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "elementwise.cuh"
#include "conv3x3.cuh"
#include "deepcnn.cuh"
#include "fastddpm.cuh"
#include "firstconv.cuh"
#include "loss.cuh"
#include "wgrad.cuh"
#include "wgrad3x3.cuh"

using namespace b200sr;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define B2_CHECK_ARG(cond)                                                                      \
    do {                                                                                        \
        if (!(cond)) return fail(B200SR_EINVAL, std::string(__func__) + ": requirement failed: " #cond); \
    } while (0)

// resident-wave sizing of the grid-stride BatchNorm-backward kernels (3 blocks of 256 threads fit per SM)
int bn_blocks_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200SR_BN_BPS");
        v = (e != nullptr && atoi(e) > 0) ? atoi(e) : 8;
    }
    return v;
}

// kernels enqueued by the last *_wgrad_det call of this thread (split-K kernel + 0/1 slice fold + layout kernel)
thread_local int g_last_wgrad_launches = 0;

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return B200SR_OK;
}

// ---------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    uint32_t rank;
    uint32_t promo;
    uint64_t dims[5];
    uint64_t strides[4];
    uint32_t box[5];
    bool operator==(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(MapKey); ++i) {
            h ^= p[i];
            h *= 1099511628211ull;
        }
        return static_cast<size_t>(h);
    }
};

std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// bf16 tiled map with the 128-byte swizzle; out-of-bounds elements read as zero.
// narrow: the box covers only part of each pixel's bytes (a channel slot of a wider buffer): promote L2 requests to 128 B
// instead of 256 B — with 256 B every read of the upsampled HALF of a decoder concat gradient (256 B per pixel at level 0)
// dragged the skip half through DRAM as well (ncu: ConvT dgrad / wgrad at level 0 read 2x their algorithmic bytes)
int make_map(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, bool narrow = false) {
    MapKey key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = ptr;
    key.rank = rank;
    static const bool force_256 = std::getenv("B200SR_L2_PROMO_256") != nullptr;  // A/B switch
    if (force_256) narrow = false;
    key.promo = narrow ? 1u : 0u;
    for (uint32_t i = 0; i < rank; ++i) {
        key.dims[i] = dims[i];
        key.box[i] = box[i];
        if (i + 1 < rank) key.strides[i] = strides_bytes[i];
    }
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) {
            *out = it->second;
            return B200SR_OK;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) return fail(B200SR_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t gdims[5];
    cuuint64_t gstrides[4];
    cuuint32_t gbox[5];
    cuuint32_t estr[5];
    for (uint32_t i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstrides[i] = strides_bytes[i];
    }
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdims, gstrides, gbox, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     narrow ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200SR_ECUDA, "cuTensorMapEncodeTiled failed, CUresult=" + std::to_string(r));
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        if (g_map_cache.size() > 4096) g_map_cache.clear();
        g_map_cache[key] = m;
    }
    *out = m;
    return B200SR_OK;
}

// (B,H,W,C) channel slot viewed as a 4-D tensor (c, w, h, b)
int make_act_map(CUtensorMap* out, const void* base, int pix_stride, int c_off, int C, int B, int H, int W, int box_w,
                 int box_h) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + c_off;
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(B)};
    const uint64_t s = static_cast<uint64_t>(pix_stride) * 2;
    const uint64_t strides[3] = {s, s * W, s * W * H};
    const uint32_t box[4] = {64, static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h), 1};
    return make_map(out, p, 4, dims, strides, box, C < pix_stride);
}

// (B,2H,2W,C) channel slot viewed as the 5-D tensor (c, j, w, i, b*H + h): sub-pixel (i,j) gather
int make_gather_map(CUtensorMap* out, const void* base, int pix_stride, int c_off, int C, int B, int H, int W,
                    int box_w, int box_h) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + c_off;
    const uint64_t dims[5] = {static_cast<uint64_t>(C), 2, static_cast<uint64_t>(W), 2,
                              static_cast<uint64_t>(B) * static_cast<uint64_t>(H)};
    const uint64_t s = static_cast<uint64_t>(pix_stride) * 2;
    const uint64_t strides[4] = {s, 2 * s, 2 * static_cast<uint64_t>(W) * s, 4 * static_cast<uint64_t>(W) * s};
    const uint32_t box[5] = {64, 1, static_cast<uint32_t>(box_w), 1, static_cast<uint32_t>(box_h)};
    return make_map(out, p, 5, dims, strides, box, true);  // every other pixel, and possibly half of its channels
}

// (B,H,W,C) channel slot viewed as the 5-D tensor (c within a 64-channel chunk, w, h, chunk, b): one box can
// carry several channel chunks, each landing as its own [h][w][64 ch] block in shared memory
int make_chunk_map(CUtensorMap* out, const void* base, int pix_stride, int c_off, int C, int B, int H, int W,
                   int box_w, int box_h, int box_chunks) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(base) + c_off;
    const uint64_t dims[5] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(C / 64),
                              static_cast<uint64_t>(B)};
    const uint64_t s = static_cast<uint64_t>(pix_stride) * 2;
    const uint64_t strides[4] = {s, s * W, 128, s * W * H};
    const uint32_t box[5] = {64, static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h),
                             static_cast<uint32_t>(box_chunks), 1};
    return make_map(out, p, 5, dims, strides, box, C < pix_stride);
}

int make_weight_map(CUtensorMap* out, const void* w, int K_total, int N_total, int block_n) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K_total), static_cast<uint64_t>(N_total)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K_total) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(block_n)};
    return make_map(out, w, 2, dims, strides, box);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int grid_for(long long total, int block, int max_blocks = 148 * 16) {
    long long g = (total + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return static_cast<int>(g);
}

// ---------------------------------------------------------------------------------------------
// persistent conv3x3 launch (forward and data gradient)
// ---------------------------------------------------------------------------------------------
int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BLOCK_N, int MODE, bool SPLIT = false, bool PAIR = false>
int launch_conv3_t(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Conv3Args& args, int grid,
                   cudaStream_t st) {
    constexpr int smem = C3Cfg<BLOCK_N, SPLIT>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel<BLOCK_N, MODE, SPLIT, PAIR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string("conv3x3 smem attribute: ") + cudaGetErrorString(e));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(c3_threads<BLOCK_N, MODE, SPLIT>());
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (PAIR) {
        // clusters of two CTAs (one TPC): tcgen05.mma.cta_group::2
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = 2;
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    // programmatic dependent launch: the CTAs may be scheduled while the previous kernel of the stream drains; the kernel
    // does its set-up (mbarriers, TMEM, descriptor prefetch) and then waits for that kernel (griddep_wait, conv3x3.cuh)
    static const bool pdl = getenv("B200SR_NO_PDL") == nullptr;
    if (pdl) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_kernel<BLOCK_N, MODE, SPLIT, PAIR>, ma, mb, mo, args);
    if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string("conv3x3 launch: ") + cudaGetErrorString(e));
    return check_launch("conv3x3_kernel");
}

template <int MODE, bool SPLIT = false>
int dispatch_conv3(int block_n, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo,
                   const Conv3Args& args, int grid, cudaStream_t st) {
    switch (block_n) {
        case 64:
            return launch_conv3_t<64, MODE, SPLIT>(ma, mb, mo, args, grid, st);
        case 128:
            return launch_conv3_t<128, MODE, SPLIT>(ma, mb, mo, args, grid, st);
        default:
            return launch_conv3_t<256, MODE, SPLIT>(ma, mb, mo, args, grid, st);
    }
}

// mode 0: conv 3x3 fwd/dgrad (a = (B,H,W,Ca) slot); mode 1: ConvT fwd (a = (B,H,W,Ca) slot, n_total = 4*cout_t,
// out = (B,2H,2W,cout_t) slot); mode 2: ConvT dgrad (a = (B,2H,2W,Ca) slot gathered per sub-pixel, out (B,H,W,n_total))
int run_conv3(int mode, const void* a, int a_stride, int a_coff, int Ca, const void* w_packed, int n_total, int B, int H, int W,
              void* out, int out_stride, int out_coff, const float* col_scale, const float* col_shift, int relu,
              float* stats, int stats_replicas, int cout_t, cudaStream_t st, const void* mask = nullptr,
              int mask_stride = 0, int mask_coff = 0, int split_stride = 0, const b200sr_bn_train* bn = nullptr,
              int stats_sum_cols = 0) {
    B2_CHECK_ARG(a != nullptr && w_packed != nullptr && out != nullptr);
    // split_stride > 0: fp32-accuracy eval epilogue ([hi | lo | hi] output parts, conv3x3.cuh), modes 0 and 1 only
    B2_CHECK_ARG(split_stride == 0 || ((mode == 0 || mode == 1) && mask == nullptr && stats == nullptr &&
                                       split_stride % 8 == 0 && split_stride >= (mode == 1 ? cout_t : n_total)));
    B2_CHECK_ARG(B > 0 && H > 0 && W > 0);
    // Any H, W: edge tiles of a shape that is not a multiple of the 16 x 8 pixel tile reach past the image — TMA zero-fills
    // those loads (the same zeros as the convolution's padding) and clips the stores, the epilogue masks the statistics.
    // The ConvTranspose modes (no halo) see the batch as ONE image of B*H rows: their sub-pixel views fold (b, h) into one
    // TMA dimension, so a ragged H must not leave a partial tile between two images.
    const bool ragged = H % C3_TILE_H != 0 || W % C3_TILE_W != 0;
    if (ragged && mode != 0) {
        B2_CHECK_ARG(static_cast<long long>(B) * H < (1ll << 30));
        H *= B;
        B = 1;
    }
    const int tiles_w = (W + C3_TILE_W - 1) / C3_TILE_W, tiles_h = (H + C3_TILE_H - 1) / C3_TILE_H;
    B2_CHECK_ARG(mask == nullptr || (mode == 0 && mask_stride % 8 == 0 && mask_coff % 8 == 0 && aligned16(mask)));
    if (mode == 1) B2_CHECK_ARG(cout_t % 32 == 0 && n_total == 4 * cout_t);
    B2_CHECK_ARG(Ca % 64 == 0 && n_total % 64 == 0);
    B2_CHECK_ARG(a_stride % 8 == 0 && a_coff % 8 == 0 && out_stride % 8 == 0 && out_coff % 8 == 0);
    B2_CHECK_ARG(aligned16(a) && aligned16(w_packed) && aligned16(out));
    B2_CHECK_ARG(stats == nullptr || stats_replicas > 0);
    int block_n = n_total % 256 == 0 ? 256 : (n_total % 128 == 0 ? 128 : 64);
    if (block_n == 256) {
        // wave quantisation on the persistent grid: 128-wide tiles cost ~4 % more per column but halve the tail
        const long long m_tiles = static_cast<long long>(B) * tiles_h * tiles_w;
        const int sms = num_sms();
        const long long t256 = m_tiles * (n_total / 256), t128 = m_tiles * (n_total / 128);
        const double c256 = static_cast<double>((t256 + sms - 1) / sms) * 256.0;
        const double c128 = static_cast<double>((t128 + sms - 1) / sms) * 128.0 * 1.04;
        if (c128 < c256) block_n = 128;
    }
    if (const char* env = getenv("B200SR_CONV_BLOCK_N")) {
        const int v = atoi(env);
        if ((v == 64 || v == 128 || v == 256) && n_total % v == 0) block_n = v;
    }
    CUtensorMap ma, mb;
    int rc;
    if (mode == 0)
        rc = make_act_map(&ma, a, a_stride, a_coff, Ca, B, H, W, C3_TILE_W, C3_TILE_H + 2);
    else if (mode == 1 || mode == 3)
        rc = make_act_map(&ma, a, a_stride, a_coff, Ca, B, H, W, C3_TILE_W, C3_TILE_H);
    else
        rc = make_gather_map(&ma, a, a_stride, a_coff, Ca, B, H, W, C3_TILE_W, C3_TILE_H);
    if (rc) return rc;
    const int taps = mode == 0 ? 9 : (mode == 2 ? 4 : 1);
    // CTA pairs (tcgen05.mma.cta_group::2, M = 256) for the N = 64 conv3x3 layers whose weights stay resident in shared
    // memory (K <= 1152): each CTA of a pair stages half of the weight rows
    const long long m_tiles_all = static_cast<long long>(B) * tiles_h * tiles_w;
    // OPT-IN (B200SR_PAIR=1): measured slower than the single-CTA kernel on this part (enc1.3 forward 299 vs 184 us, dec1.0
    // 355 vs 336 us, step +0.3 ms; profiles/r2_cta_pair_experiment.txt) — kept as a tested, documented negative result
    const bool pair_enabled = getenv("B200SR_PAIR") != nullptr;
    const bool pair = pair_enabled && !ragged && mode == 0 && block_n == 64 && n_total == 64 && mask == nullptr && split_stride == 0 &&
                      taps * (Ca / 64) <= C3Cfg<64>::SB && m_tiles_all % 2 == 0 && num_sms() % 2 == 0 &&
                      getenv("B200SR_NO_BRESIDENT") == nullptr;
    rc = make_weight_map(&mb, w_packed, taps * Ca, n_total, pair ? 32 : (block_n < 128 ? block_n : 128));
    if (rc) return rc;
    // output tile store: 128 pixels x 64 channels per TMA store; mode 1 scatters through the sub-pixel view
    CUtensorMap mo;
    if (mode == 1)
        rc = make_gather_map(&mo, out, out_stride, out_coff, 2 * split_stride + cout_t, B, H, W, C3_TILE_W, C3_TILE_H);
    else
        rc = make_act_map(&mo, out, out_stride, out_coff, 2 * split_stride + n_total, B, H, W, C3_TILE_W, C3_TILE_H);
    if (rc) return rc;
    if (mode == 1) B2_CHECK_ARG(cout_t % 64 == 0);
    Conv3Args args;
    args.H = H;
    args.W = W;
    args.tiles_w = tiles_w;
    args.tiles_hw = tiles_h * tiles_w;
    args.ragged = ragged ? 1 : 0;
    args.n_tiles = n_total / block_n;
    const long long tiles = static_cast<long long>(B) * args.tiles_hw * args.n_tiles;
    B2_CHECK_ARG(tiles < (1ll << 31));
    args.num_tiles = static_cast<int>(tiles);
    args.cin_chunks = Ca / 64;
    args.C = Ca;
    args.n_total = n_total;
    args.relu = relu;
    args.out_pix_stride = out_stride;
    args.out_c_off = out_coff;
    args.stats_replicas = stats_replicas > 0 ? stats_replicas : 1;
    args.out = static_cast<__nv_bfloat16*>(out);
    args.col_scale = col_scale;
    args.col_shift = col_shift;
    args.stats = stats;
    args.cout_t = cout_t;
    args.mask = static_cast<const __nv_bfloat16*>(mask);
    args.mask_pix_stride = mask_stride;
    args.mask_c_off = mask_coff;
    args.split_stride = split_stride;
    args.stats_sum_cols = stats_sum_cols;
    args.bn_scale = nullptr;
    {
        const int nh = block_n / (block_n < 128 ? block_n : 128);
        const int slots = taps * args.cin_chunks * nh;
        const int sb = (block_n == 64 && split_stride == 0) ? C3Cfg<64>::SB : C3Cfg<128>::SB;
        args.b_resident = (slots <= sb && getenv("B200SR_NO_BRESIDENT") == nullptr) ? 1 : 0;
        // ring bytes that resident weights leave unused become extra activation stages
        const int b_slot = (block_n < 128 ? block_n : 128) * 128;
        const int ring = block_n == 64 ? (split_stride ? C3Cfg<64, true>::RING_BYTES : C3Cfg<64>::RING_BYTES)
                                       : C3Cfg<128>::RING_BYTES;
        int sa = C3Cfg<64>::SA;
        if (args.b_resident && getenv("B200SR_FIXED_SA") == nullptr) {
            sa = (ring - slots * b_slot) / C3_A_SLOT;
            if (sa > C3_SA_MAX) sa = C3_SA_MAX;
            if (sa < C3Cfg<64>::SA) sa = C3Cfg<64>::SA;
        }
        args.sa = sa;
    }
    // persistent grid: one CTA per SM, rounded down so that a CTA stays on one column block (register statistics)
    int grid = num_sms();
    grid -= grid % args.n_tiles;
    if (grid < args.n_tiles) grid = args.n_tiles;
    if (grid > args.num_tiles) grid = args.num_tiles;
    // deterministic statistics when the caller provides one slot per CTA of a column block (see Conv3Args::stats_slots)
    args.stats_slots = (stats != nullptr && stats_replicas >= grid / args.n_tiles) ? 1 : 0;
    if (bn != nullptr) {
        // fused train-mode BatchNorm finalize in the last CTA of every column block (needs the deterministic slots)
        B2_CHECK_ARG(mode == 0 && mask == nullptr && split_stride == 0 && args.stats_slots == 1);
        B2_CHECK_ARG(bn->gamma && bn->beta && bn->scale && bn->shift && bn->save_mean && bn->save_invstd && bn->counters);
        B2_CHECK_ARG((bn->running_mean == nullptr) == (bn->running_var == nullptr) && bn->count > 1.0);
        args.bn_gamma = bn->gamma;
        args.bn_beta = bn->beta;
        args.bn_conv_bias = bn->conv_bias;
        args.bn_scale = bn->scale;
        args.bn_shift = bn->shift;
        args.bn_mean = bn->save_mean;
        args.bn_invstd = bn->save_invstd;
        args.bn_rmean = bn->running_mean;
        args.bn_rvar = bn->running_var;
        args.bn_nbt = reinterpret_cast<long long*>(bn->num_batches_tracked);
        args.bn_counters = bn->counters;
        args.bn_count = static_cast<float>(bn->count);
        args.bn_eps = bn->eps;
        args.bn_momentum = bn->momentum;
    }
    if (split_stride > 0)
        return mode == 0 ? dispatch_conv3<0, true>(block_n, ma, mb, mo, args, grid, st)
                         : dispatch_conv3<1, true>(block_n, ma, mb, mo, args, grid, st);
    if (pair && args.b_resident && grid % 2 == 0) return launch_conv3_t<64, 0, false, true>(ma, mb, mo, args, grid, st);
    if (pair) return fail(B200SR_EINVAL, "conv3x3: pair mode chosen for a shape it does not cover");
    if (mode == 0 && mask != nullptr) return dispatch_conv3<4>(block_n, ma, mb, mo, args, grid, st);
    if (mode == 0) return dispatch_conv3<0>(block_n, ma, mb, mo, args, grid, st);
    if (mode == 1) return dispatch_conv3<1>(block_n, ma, mb, mo, args, grid, st);
    if (mode == 3) return dispatch_conv3<3>(block_n, ma, mb, mo, args, grid, st);
    return dispatch_conv3<2>(block_n, ma, mb, mo, args, grid, st);
}

// ---------------------------------------------------------------------------------------------
// wgrad launch
// ---------------------------------------------------------------------------------------------
// max_splits > 0: deterministic mode (partials of split s at args.out + s * args.split_stride), at most max_splits splits
template <int N_TILE, int STAGES, int KPIX = 32>
int launch_wgrad_t(const CUtensorMap& mt, const CUtensorMap& mp, WGradArgs& args, cudaStream_t st, int max_splits = 0,
                   int* splits_out = nullptr) {
    constexpr int smem = wg_smem_bytes<N_TILE, STAGES, KPIX>();
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(wgrad_kernel<N_TILE, STAGES, KPIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string("wgrad smem attribute: ") + cudaGetErrorString(e));
        configured = true;
    }
    constexpr int AT = wg_atoms_per_cta<N_TILE>();
    const int groups_x = (args.total_atoms + AT - 1) / AT;
    const int groups_y = args.n_total / N_TILE;
    const int groups = groups_x * groups_y;
    int target = 296;  // two waves of 148 SMs
    if (const char* env = getenv("B200SR_WGRAD_CTAS")) target = atoi(env) > 0 ? atoi(env) : target;
    // whole waves: never more CTAs than the target (upconv4: 32 groups x 10 splits = 320 CTAs was a third, mostly empty wave)
    int splits = groups >= target ? 1 : target / groups;
    if (splits > args.total_chunks) splits = args.total_chunks;
    if (max_splits > 0 && splits > max_splits) splits = max_splits;
    args.chunks_per_cta = (args.total_chunks + splits - 1) / splits;
    splits = (args.total_chunks + args.chunks_per_cta - 1) / args.chunks_per_cta;
    if (splits_out != nullptr) *splits_out = splits;
    dim3 grid(groups_x, groups_y, splits);
    wgrad_kernel<N_TILE, STAGES, KPIX><<<grid, WG_THREADS, smem, st>>>(mt, mp, args);
    return check_launch("wgrad_kernel");
}

// t_mode 0: T = (B,H,W,Ct) slot, taps in {1,9}; t_mode 1: T = (B,2H,2W,Ct) slot gathered (4 taps). P = (B,H,W,Cp).
int run_wgrad(int t_mode, const void* t, int t_stride, int t_coff, int Ct, int num_taps, const void* p, int p_stride,
              int p_coff, int Cp, int B, int H, int W, float* G, cudaStream_t st, long long ws_floats = 0,
              int* splits_out = nullptr) {
    B2_CHECK_ARG(t != nullptr && p != nullptr && G != nullptr);
    // any H, W: K chunks (2 x 16 pixels) that reach past the image load zeros for P through TMA (h and b are separate
    // dimensions of its map), so they add nothing to the sum
    B2_CHECK_ARG(B > 0 && H > 0 && W > 0);
    B2_CHECK_ARG(Ct % 64 == 0 && Cp % 64 == 0);
    B2_CHECK_ARG(t_stride % 8 == 0 && t_coff % 8 == 0 && p_stride % 8 == 0 && p_coff % 8 == 0);
    B2_CHECK_ARG(aligned16(t) && aligned16(p) && aligned16(G));
    // pixel rows per pipeline stage: 4 for the N_TILE = 256 variant (64-pixel stages), 2 otherwise
    static const bool k64 = getenv("B200SR_WGRAD_K32") == nullptr;
    const int rows = (Cp % 256 == 0 && k64) ? 4 : 2;
    CUtensorMap mt, mp;
    int rc;
    if (t_mode == 0)
        rc = make_act_map(&mt, t, t_stride, t_coff, Ct, B, H, W, 16, rows);
    else
        rc = make_gather_map(&mt, t, t_stride, t_coff, Ct, B, H, W, 16, rows);
    if (rc) return rc;
    rc = make_act_map(&mp, p, p_stride, p_coff, Cp, B, H, W, 16, rows);
    if (rc) return rc;
    WGradArgs args;
    args.H = H;
    args.W = W;
    args.chunks_w = (W + 15) / 16;
    args.chunks_hw = ((H + rows - 1) / rows) * args.chunks_w;
    args.total_chunks = B * args.chunks_hw;
    args.chunks_per_cta = args.total_chunks;
    args.t_mode = t_mode;
    args.num_taps = num_taps;
    args.tchunks_per_tap = Ct / 64;
    args.total_atoms = num_taps * (Ct / 64);
    args.n_total = Cp;
    args.out = G;
    args.split_stride = 0;
    int max_splits = 0;
    if (ws_floats > 0) {  // deterministic mode: G is a workspace of per-split partials
        args.split_stride = static_cast<long long>(num_taps) * Ct * Cp;
        max_splits = static_cast<int>(ws_floats / args.split_stride < 4096 ? ws_floats / args.split_stride : 4096);
        B2_CHECK_ARG(max_splits >= 1);
    }
    if (Cp % 256 == 0 && rows == 4) return launch_wgrad_t<256, 3, 64>(mt, mp, args, st, max_splits, splits_out);
    if (Cp % 256 == 0) return launch_wgrad_t<256, 6>(mt, mp, args, st, max_splits, splits_out);
    if (Cp % 128 == 0) return launch_wgrad_t<128, 5>(mt, mp, args, st, max_splits, splits_out);
    return launch_wgrad_t<64, 3>(mt, mp, args, st, max_splits, splits_out);
}

// ---------------------------------------------------------------------------------------------
// second-generation conv3x3 wgrad launch
// ---------------------------------------------------------------------------------------------
template <int N_TILE, int MODE_B>
int launch_wgrad3_t(const CUtensorMap& mx, const CUtensorMap& mz, WG3Args& args, int jobs, cudaStream_t st,
                    int ws_max_splits, int* splits_out) {
    using Cfg = WG3Cfg<N_TILE, MODE_B>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad3x3_kernel<N_TILE, MODE_B>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string("wgrad3x3 smem attribute: ") + cudaGetErrorString(e));
        configured = true;
    }
    // split-K factor: fill whole waves of SMs (one CTA per SM), prefer fewer splits (less reduction traffic)
    const int sms = num_sms();
    int best = 1;
    double best_score = -1.0;
    // up to two waves' worth of splits: a single-job layer (Cin = Cout = 64) must still be able to fill every SM
    // (capping the splits at one wave of CTAs — half the partial traffic of the second stage — was measured: no effect on
    // the step within +-0.05 ms, profiles/r2_ab_stage2.txt)
    const int split_cap = 2 * sms > 64 ? 2 * sms : 64;
    int max_splits = args.total_chunks < split_cap ? args.total_chunks : split_cap;
    if (ws_max_splits > 0 && max_splits > ws_max_splits) max_splits = ws_max_splits;  // partial workspace capacity
    for (int s = 1; s <= max_splits; ++s) {
        const int per = (args.total_chunks + s - 1) / s;
        const int real = (args.total_chunks + per - 1) / per;
        if (real != s) continue;
        const long long ctas = static_cast<long long>(jobs) * s;
        const long long waves = (ctas + sms - 1) / sms;
        // time ~ waves * (per + fixed per-CTA cost of ~6 stages for prologue/epilogue)
        const double t = static_cast<double>(waves) * (per + 6.0);
        const double score = 1.0 / t;
        if (score > best_score * 1.0001) {
            best_score = score;
            best = s;
        }
    }
    if (const char* env = getenv("B200SR_WGRAD_SPLITS")) best = atoi(env) > 0 ? atoi(env) : best;
    args.chunks_per_cta = (args.total_chunks + best - 1) / best;
    const int splits = (args.total_chunks + args.chunks_per_cta - 1) / args.chunks_per_cta;
    if (ws_max_splits > 0 && splits > ws_max_splits) return fail(B200SR_EINVAL, "wgrad3x3: split workspace too small");
    if (splits_out != nullptr) *splits_out = splits;
    dim3 grid(jobs, splits, 1);
    wgrad3x3_kernel<N_TILE, MODE_B><<<grid, WG3_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mz, args);
    return check_launch("wgrad3x3_kernel");
}

// returns -1 when the shape is not covered (caller falls back to the first-generation kernel)
int run_wgrad3(const void* x, int x_stride, int x_coff, int Cin, const void* dz, int dz_stride, int dz_coff, int Cout,
               int B, int H, int W, float* G, cudaStream_t st, long long ws_floats = 0, int* splits_out = nullptr) {
    const bool mode_b = Cin == 64;
    if (!mode_b && Cin % 128 != 0) return -1;
    if (Cout % 64 != 0) return -1;
    const int n_tile = (!mode_b && Cout % 128 == 0) ? 128 : 64;
    CUtensorMap mx, mz;
    int rc = make_chunk_map(&mx, x, x_stride, x_coff, Cin, B, H, W, 16, 6, mode_b ? 1 : 2);
    if (rc) return rc;
    rc = make_chunk_map(&mz, dz, dz_stride, dz_coff, Cout, B, H, W, 16, 4, n_tile / 64);
    if (rc) return rc;
    WG3Args args;
    args.H = H;
    args.W = W;
    // any H, W: K chunks (4 x 16 pixels) past the image edge are zero-filled by TMA for both operands
    args.chunks_w = (W + 15) / 16;
    args.chunks_hw = ((H + 3) / 4) * args.chunks_w;
    args.total_chunks = B * args.chunks_hw;
    args.chunks_per_cta = args.total_chunks;
    args.mode_b = mode_b ? 1 : 0;
    args.jobs_ci = mode_b ? 1 : Cin / 128;
    args.jobs_co = Cout / n_tile;
    args.Cin = Cin;
    args.Cout = Cout;
    args.out = G;
    args.split_stride = 0;
    int ws_max = 0;
    if (ws_floats > 0) {  // deterministic mode: G is a workspace of per-split partials
        args.split_stride = 9LL * Cin * Cout;
        const long long m = ws_floats / args.split_stride;
        if (m < 1) return fail(B200SR_EINVAL, "wgrad3x3: split workspace too small");
        ws_max = static_cast<int>(m < 4096 ? m : 4096);
    }
    const int jobs = (mode_b ? 1 : 3 * args.jobs_ci) * args.jobs_co;
    if (mode_b) return launch_wgrad3_t<64, 1>(mx, mz, args, jobs, st, ws_max, splits_out);
    if (n_tile == 128) return launch_wgrad3_t<128, 0>(mx, mz, args, jobs, st, ws_max, splits_out);
    return launch_wgrad3_t<64, 0>(mx, mz, args, jobs, st, ws_max, splits_out);
}

}  // namespace

// =================================================================================================
// exported C ABI
// =================================================================================================
extern "C" {

int b200sr_version(void) { return 100; }

const char* b200sr_last_error(void) { return g_last_error.c_str(); }

int b200sr_last_wgrad_launches(void) { return g_last_wgrad_launches; }

int b200sr_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200SR_ENODEV, "no CUDA device");
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200SR_ENODEV, "cudaGetDeviceProperties failed");
    }
    if (prop.major != 10) return fail(B200SR_ENODEV, "device is not compute capability 10.x (sm_100a required)");
    return B200SR_OK;
}

int b200sr_conv3x3_fwd(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout, int B,
                       int H, int W, void* out, int out_pix_stride, int out_c_off, const float* col_scale,
                       const float* col_shift, int relu, float* stats, int stats_replicas, void* stream) {
    return run_conv3(0, x, x_pix_stride, x_c_off, Cin, w_packed, Cout, B, H, W, out, out_pix_stride, out_c_off,
                     col_scale, col_shift, relu, stats, stats_replicas, Cout, static_cast<cudaStream_t>(stream));
}

int b200sr_conv3x3_fwd_bn(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout, int B,
                          int H, int W, void* out, int out_pix_stride, int out_c_off, float* stats, int stats_replicas,
                          const b200sr_bn_train* bn, void* stream) {
    B2_CHECK_ARG(bn != nullptr && stats != nullptr);
    return run_conv3(0, x, x_pix_stride, x_c_off, Cin, w_packed, Cout, B, H, W, out, out_pix_stride, out_c_off, nullptr,
                     nullptr, 0, stats, stats_replicas, Cout, static_cast<cudaStream_t>(stream), nullptr, 0, 0, 0, bn);
}

int b200sr_conv3x3_dgrad(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed, int Cin,
                         int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, float* stats,
                         int stats_replicas, void* stream) {
    return run_conv3(0, dy, dy_pix_stride, dy_c_off, Cout, w_packed, Cin, B, H, W, dx, dx_pix_stride, dx_c_off,
                     nullptr, nullptr, 0, stats, stats_replicas, Cin, static_cast<cudaStream_t>(stream));
}

int b200sr_conv3x3_dgrad_colsum(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed, int Cin,
                                int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, float* colsum_slots,
                                int slots, int ncols, void* stream) {
    B2_CHECK_ARG(colsum_slots != nullptr && slots > 0 && ncols > 0 && ncols <= Cin && ncols % 32 == 0);
    return run_conv3(0, dy, dy_pix_stride, dy_c_off, Cout, w_packed, Cin, B, H, W, dx, dx_pix_stride, dx_c_off, nullptr,
                     nullptr, 0, colsum_slots, slots, Cin, static_cast<cudaStream_t>(stream), nullptr, 0, 0, 0, nullptr, ncols);
}

int b200sr_conv3x3_dgrad_relu(const void* dy, int dy_pix_stride, int dy_c_off, int Cout, const void* w_packed, int Cin,
                              int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, const void* act,
                              int act_pix_stride, int act_c_off, float* stats, int stats_replicas, void* stream) {
    B2_CHECK_ARG(act != nullptr);
    return run_conv3(0, dy, dy_pix_stride, dy_c_off, Cout, w_packed, Cin, B, H, W, dx, dx_pix_stride, dx_c_off, nullptr,
                     nullptr, 0, stats, stats_replicas, Cin, static_cast<cudaStream_t>(stream), act, act_pix_stride,
                     act_c_off);
}

int b200sr_convT2x2_fwd(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* w_packed, int Cout,
                        const float* bias, int B, int H, int W, void* out, int out_pix_stride, int out_c_off,
                        void* stream) {
    return run_conv3(1, x, x_pix_stride, x_c_off, Cin, w_packed, 4 * Cout, B, H, W, out, out_pix_stride, out_c_off,
                     nullptr, bias, 0, nullptr, 0, Cout, static_cast<cudaStream_t>(stream));
}

int b200sr_convT2x2_dgrad(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* w_packed, int Cin,
                          int B, int H, int W, void* dx, int dx_pix_stride, int dx_c_off, void* stream) {
    return run_conv3(2, dup, dup_pix_stride, dup_c_off, Cout, w_packed, Cin, B, H, W, dx, dx_pix_stride, dx_c_off,
                     nullptr, nullptr, 0, nullptr, 0, Cin, static_cast<cudaStream_t>(stream));
}

int b200sr_conv3x3_wgrad(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                         int dz_c_off, int Cout, int B, int H, int W, float* G, void* stream) {
    if (getenv("B200SR_WGRAD_V1") == nullptr && x != nullptr && dz != nullptr && G != nullptr && B > 0 &&
        x_pix_stride % 8 == 0 && x_c_off % 8 == 0 && dz_pix_stride % 8 == 0 && dz_c_off % 8 == 0 && aligned16(x) &&
        aligned16(dz) && aligned16(G)) {
        const int rc = run_wgrad3(x, x_pix_stride, x_c_off, Cin, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, G,
                                  static_cast<cudaStream_t>(stream));
        if (rc >= 0) return rc;
    }
    return run_wgrad(0, x, x_pix_stride, x_c_off, Cin, 9, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, G,
                     static_cast<cudaStream_t>(stream));
}

int b200sr_convT2x2_wgrad(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* x,
                          int x_pix_stride, int x_c_off, int Cin, int B, int H, int W, float* G, void* stream) {
    return run_wgrad(1, dup, dup_pix_stride, dup_c_off, Cout, 4, x, x_pix_stride, x_c_off, Cin, B, H, W, G,
                     static_cast<cudaStream_t>(stream));
}

int b200sr_pack_jobs(const b200sr_pack_job* jobs, int njobs, void* stream) {
    B2_CHECK_ARG(jobs != nullptr && njobs > 0);
    static_assert(sizeof(b200sr_pack_job) == sizeof(PackJob), "PackJob ABI mismatch");
    dim3 grid(1024, njobs);  // >= tiles of the largest layer; blocks beyond a job's tile count exit at once
    pack_jobs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const PackJob*>(jobs));
    return check_launch("pack_jobs_kernel");
}

int b200sr_bn_fold_eval(const b200sr_fold_job* jobs, int njobs, float eps, void* stream) {
    B2_CHECK_ARG(jobs != nullptr && njobs > 0);
    static_assert(sizeof(b200sr_fold_job) == sizeof(FoldJob), "FoldJob ABI mismatch");
    dim3 grid(4, njobs);
    bn_fold_eval_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const FoldJob*>(jobs),
                                                                             eps);
    return check_launch("bn_fold_eval_kernel");
}

int b200sr_conv1_fwd(const float* x, const float* w, const float* col_scale, const float* col_shift, int relu,
                     void* out, float* stats, int stats_replicas, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x != nullptr && w != nullptr && out != nullptr);
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0);
    B2_CHECK_ARG((col_scale == nullptr) == (col_shift == nullptr));
    B2_CHECK_ARG(stats == nullptr || stats_replicas > 0);
    B2_CHECK_ARG(aligned16(out));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    FirstConvSrc src{nullptr, nullptr, nullptr, x};  // tensor-core (warp MMA) kernel, csrc/firstconv.cuh
    const int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    first_conv_mma_fwd_kernel<2><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, w, 18, nullptr, col_scale, col_shift, relu, static_cast<__nv_bfloat16*>(out), stats,
        stats_replicas > 0 ? stats_replicas : 1, (stats != nullptr && stats_replicas >= grid) ? 1 : 0, H, W, tiles);
    return check_launch("first_conv_mma_fwd_kernel<2>");
}

int b200sr_conv1_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x != nullptr && dz != nullptr && dw != nullptr);
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(dz));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    FirstConvSrc src{nullptr, nullptr, nullptr, x};
    const int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    first_conv_mma_wgrad_kernel<2><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<const __nv_bfloat16*>(dz), dw, 18, H, W, tiles, nullptr);
    return check_launch("first_conv_mma_wgrad_kernel<2>");
}

int b200sr_conv1_dgrad(const void* dz, const float* w, float* dx, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(dz != nullptr && w != nullptr && dx != nullptr);
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(dz));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    const int grid = tiles < num_sms() * 4 ? tiles : num_sms() * 4;
    conv1_direct_dgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dz), w, dx, H, W, tiles);
    return check_launch("conv1_direct_dgrad_kernel");
}

int b200sr_bn_finalize(const float* stats, int replicas, int C, double count, const float* gamma, const float* beta,
                       const float* conv_bias, float eps, float momentum, float* scale, float* shift, float* save_mean,
                       float* save_invstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                       void* stream) {
    B2_CHECK_ARG(stats && gamma && beta && scale && shift && save_mean && save_invstd);
    B2_CHECK_ARG(replicas > 0 && C > 0 && count > 1.0);
    B2_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr));
    bn_finalize_kernel<<<(C + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, replicas, C, static_cast<float>(count), gamma, beta, conv_bias, eps, momentum, scale, shift, save_mean,
        save_invstd, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked));
    return check_launch("bn_finalize_kernel");
}

int b200sr_bnrelu_apply(const void* z, int C, const float* scale, const float* shift, void* act, int act_pix_stride,
                        int act_c_off, void* pooled, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(z && scale && shift && act);
    B2_CHECK_ARG(C % 8 == 0 && act_pix_stride % 8 == 0 && act_c_off % 8 == 0 && B > 0 && H > 0 && W > 0);
    B2_CHECK_ARG(pooled == nullptr || (H % 2 == 0 && W % 2 == 0));  // MaxPool2d(2,2) of an even-sized level
    B2_CHECK_ARG(aligned16(z) && aligned16(act) && (pooled == nullptr || aligned16(pooled)));
    const long long total = static_cast<long long>(B) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
    if (H % 2 == 0 && W % 2 == 0)
        bnrelu_apply_kernel<false><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(z), C, scale, shift, static_cast<__nv_bfloat16*>(act), act_pix_stride,
            act_c_off, static_cast<__nv_bfloat16*>(pooled), H, W, total);
    else
        bnrelu_apply_kernel<true><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(z), C, scale, shift, static_cast<__nv_bfloat16*>(act), act_pix_stride,
            act_c_off, static_cast<__nv_bfloat16*>(pooled), H, W, total);
    return check_launch("bnrelu_apply_kernel");
}

int b200sr_bn_train_apply(const void* z, int C, const float* stats, int replicas, double count, const float* gamma,
                           const float* beta, const float* conv_bias, float eps, float momentum, float* scale,
                           float* shift, float* save_mean, float* save_invstd, float* running_mean, float* running_var,
                           void* act, int act_pix_stride, int act_c_off, void* pooled, int B, int H, int W,
                           void* stream) {
    B2_CHECK_ARG(z && stats && gamma && beta && scale && shift && save_mean && save_invstd && act);
    B2_CHECK_ARG(replicas > 0 && count > 1.0 && (running_mean == nullptr) == (running_var == nullptr));
    B2_CHECK_ARG(C % 8 == 0 && act_pix_stride % 8 == 0 && act_c_off % 8 == 0 && H % 2 == 0 && W % 2 == 0);
    B2_CHECK_ARG(aligned16(z) && aligned16(act) && (pooled == nullptr || aligned16(pooled)) && aligned16(stats));
    const int c8n = C / 8;
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * c8n;
    // a thread must stay on the same 8 channels across its grid-stride loop: grid * 256 % c8n == 0, and the first c8n
    // threads of the grid publish the per-channel results, so the grid must hold at least c8n threads
    B2_CHECK_ARG(c8n <= 256 ? 256 % c8n == 0 : c8n % 256 == 0);
    int grid = grid_for(total, 256);
    if (c8n > 256) {
        const int m = c8n / 256;
        grid = (grid + m - 1) / m * m;
    }
    bn_train_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(z), C, stats, replicas, static_cast<float>(count), gamma, beta, conv_bias, eps,
        momentum, scale, shift, save_mean, save_invstd, running_mean, running_var, static_cast<__nv_bfloat16*>(act),
        act_pix_stride, act_c_off, static_cast<__nv_bfloat16*>(pooled), H, W, total);
    return check_launch("bn_train_apply_kernel");
}

int b200sr_bn_bwd_apply_fused(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C, const float* scale,
                              const float* shift, const float* mean, const float* invstd, const float* sums,
                              int replicas, double count, float* dgamma, float* dbeta, void* dz, int64_t npix,
                              void* stream) {
    B2_CHECK_ARG(dy && z && scale && shift && mean && invstd && sums && dgamma && dbeta && dz);
    B2_CHECK_ARG(C % 8 == 0 && dy_pix_stride % 8 == 0 && dy_c_off % 8 == 0 && npix > 0 && replicas > 0 && count > 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(z) && aligned16(dz) && aligned16(sums));
    const int CV = C / 8;
    B2_CHECK_ARG(CV <= 256 && 256 % CV == 0);
    const int PB = 256 / CV;
    long long blocks = (npix + static_cast<long long>(PB) * BNB_UNROLL - 1) / (static_cast<long long>(PB) * BNB_UNROLL);
    if (blocks > num_sms() * bn_blocks_per_sm()) blocks = num_sms() * bn_blocks_per_sm();
    bn_bwd_apply_fused_kernel<false><<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C, scale,
        shift, mean, invstd, sums, replicas, static_cast<float>(count), dgamma, dbeta, static_cast<__nv_bfloat16*>(dz),
        npix, nullptr);
    return check_launch("bn_bwd_apply_fused_kernel");
}

// ---- deterministic (bit-reproducible) variants: per-block / per-split partials + fixed-order second stage ----------
namespace {
int launch_reduce_unpack(float* ws, int splits, long long split_stride, int T, int outer_total, int inner_total,
                         int inner_dst, int inner_off, float* dst, cudaStream_t st) {
    B2_CHECK_ARG(outer_total % PK_TILE == 0 && inner_total % PK_TILE == 0 && (T == 9 || T == 4 || T == 1));
    int tiles = (outer_total / PK_TILE) * (inner_total / PK_TILE);
    // Layers with at least one tile per SM and few split-K slices: the layout kernel sums the slices itself (fixed slice
    // order), one launch instead of two. (Round-2 A/B had this at +0.25 ms — that was the run-time-T kernel whose loads
    // were serialised, see wgrad_reduce_unpack_kernel.) B200SR_UNPACK_TWO_STAGE=1 restores the two-stage form.
    static const bool two_stage = getenv("B200SR_UNPACK_TWO_STAGE") != nullptr;
    const bool direct = !two_stage && splits > 1 && splits <= 4 && tiles >= num_sms();
    if (splits > 1 && !direct) {
        // stage 2a: fold the split-K slices into slice 0 (element-parallel; slice lanes when there are many slices)
        B2_CHECK_ARG(split_stride % 4 == 0);
        const long long n4 = split_stride / 4;
        const int lanes = splits >= 32 ? 8 : (splits >= 16 ? 4 : (splits >= 8 ? 2 : 1));
        const long long blocks = (n4 + 256 / lanes - 1) / (256 / lanes);
        reduce_splits_inplace_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(reinterpret_cast<float4*>(ws), splits,
                                                                                   n4, n4, lanes);
    }
    g_last_wgrad_launches = 2 + ((splits > 1 && !direct) ? 1 : 0);
    // stage 2b: kernel layout [t][inner][outer] -> PyTorch parameter layout
    const int ns = direct ? splits : 1;
    if (tiles > num_sms() * 8) tiles = num_sms() * 8;
    if (T == 9)
        wgrad_reduce_unpack_kernel<9><<<tiles, 256, 0, st>>>(ws, ns, split_stride, outer_total, inner_total, inner_dst,
                                                             inner_off, dst);
    else if (T == 4)
        wgrad_reduce_unpack_kernel<4><<<tiles, 256, 0, st>>>(ws, ns, split_stride, outer_total, inner_total, inner_dst,
                                                             inner_off, dst);
    else
        wgrad_reduce_unpack_kernel<1><<<tiles, 256, 0, st>>>(ws, ns, split_stride, outer_total, inner_total, inner_dst,
                                                             inner_off, dst);
    return check_launch("wgrad_reduce_unpack_kernel");
}
}  // namespace

int b200sr_conv3x3_wgrad_det(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                             int dz_c_off, int Cout, int B, int H, int W, float* dW, int cin_total, int cin_off,
                             float* ws, int64_t ws_floats, void* stream) {
    B2_CHECK_ARG(x && dz && dW && ws && ws_floats > 0 && B > 0);
    B2_CHECK_ARG(cin_total >= Cin && cin_off >= 0 && cin_off + Cin <= cin_total);
    B2_CHECK_ARG(x_pix_stride % 8 == 0 && x_c_off % 8 == 0 && dz_pix_stride % 8 == 0 && dz_c_off % 8 == 0);
    B2_CHECK_ARG(aligned16(x) && aligned16(dz) && aligned16(ws));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int splits = 0;
    int rc = run_wgrad3(x, x_pix_stride, x_c_off, Cin, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, ws, st, ws_floats,
                        &splits);
    if (rc < 0)
        rc = run_wgrad(0, x, x_pix_stride, x_c_off, Cin, 9, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, ws, st, ws_floats,
                       &splits);
    if (rc) return rc;
    return launch_reduce_unpack(ws, splits, 9LL * Cin * Cout, 9, Cout, Cin, cin_total, cin_off, dW, st);
}

int b200sr_convT2x2_wgrad_det(const void* dup, int dup_pix_stride, int dup_c_off, int Cout, const void* x,
                              int x_pix_stride, int x_c_off, int Cin, int B, int H, int W, float* dW, float* ws,
                              int64_t ws_floats, void* stream) {
    B2_CHECK_ARG(dW && ws && ws_floats > 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int splits = 0;
    const int rc = run_wgrad(1, dup, dup_pix_stride, dup_c_off, Cout, 4, x, x_pix_stride, x_c_off, Cin, B, H, W, ws, st,
                             ws_floats, &splits);
    if (rc) return rc;
    return launch_reduce_unpack(ws, splits, 4LL * Cout * Cin, 4, Cin, Cout, Cout, 0, dW, st);
}

int b200sr_conv1x1_wgrad_det(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                             int dz_c_off, int Cout, int B, int H, int W, float* dW, float* ws, int64_t ws_floats,
                             void* stream) {
    B2_CHECK_ARG(dW && ws && ws_floats > 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int splits = 0;
    const int rc = run_wgrad(0, x, x_pix_stride, x_c_off, Cin, 1, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, ws, st,
                             ws_floats, &splits);
    if (rc) return rc;
    return launch_reduce_unpack(ws, splits, 1LL * Cin * Cout, 1, Cout, Cin, Cin, 0, dW, st);
}

int b200sr_conv1_wgrad_det(const float* x, const void* dz, float* dw, int B, int H, int W, float* ws, int64_t ws_floats,
                           void* stream) {
    B2_CHECK_ARG(x != nullptr && dz != nullptr && dw != nullptr && ws != nullptr);
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(dz));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    FirstConvSrc src{nullptr, nullptr, nullptr, x};
    int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    const long long cap = ws_floats / (18 * FC_COUT);
    B2_CHECK_ARG(cap >= 1);
    if (grid > cap) grid = static_cast<int>(cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    first_conv_mma_wgrad_kernel<2><<<grid, 256, 0, st>>>(src, static_cast<const __nv_bfloat16*>(dz), dw, 18, H, W, tiles,
                                                         ws);
    // partial [k][co] -> dw[co*18 + k]
    reduce_partials_kernel<<<(18 * FC_COUT + 31) / 32, 256, 0, st>>>(ws, grid, 18LL * FC_COUT, 18 * FC_COUT, dw, FC_COUT,
                                                                   18, 1, 1.f);
    return check_launch("first_conv_mma_wgrad_kernel<2> (deterministic)");
}

/* out[i] = sum_s slots[s*slot_stride + i] (s ascending), i < n: finishes per-CTA column sums (ConvTranspose2d bias
 * gradient from the dgrad epilogue statistics) in a fixed order. */
int b200sr_sum_slots(const float* slots, int nslots, int64_t slot_stride, int n, float* dst, void* stream) {
    B2_CHECK_ARG(slots && dst && nslots > 0 && n > 0 && slot_stride >= n);
    reduce_partials_kernel<<<(n + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(slots, nslots, slot_stride, n, dst,
                                                                                      n, 1, 0, 1.f);
    return check_launch("reduce_partials_kernel");
}

int64_t b200sr_bn_bwd_ws_floats(int C) {
    const int groups = C / 64 > 0 ? C / 64 : 1;
    int slices = (num_sms() * 3) / groups;
    if (slices < 1) slices = 1;
    return static_cast<int64_t>(groups) * slices * 128;
}

int b200sr_bn_bwd_reduce_det(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C, const float* scale,
                             const float* shift, const float* mean, const float* invstd, float* sums, float* ws,
                             int64_t ws_floats, uint32_t* counters, const void* mask_src, int64_t npix, void* stream) {
    B2_CHECK_ARG(dy && z && scale && shift && mean && invstd && sums && ws && counters);
    B2_CHECK_ARG(C % 64 == 0 && dy_pix_stride % 8 == 0 && dy_c_off % 8 == 0 && npix > 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(z) && aligned16(ws) && (mask_src == nullptr || aligned16(mask_src)));
    const int groups = C / 64;
    long long slices = (num_sms() * 3) / groups;
    const long long need = (npix + 32LL * BNB_UNROLL - 1) / (32LL * BNB_UNROLL);
    if (slices > need) slices = need;
    if (slices > ws_floats / (128LL * groups)) slices = ws_floats / (128LL * groups);
    B2_CHECK_ARG(slices >= 1);
    dim3 grid(static_cast<unsigned>(slices), static_cast<unsigned>(groups));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mask_src != nullptr)
        bn_bwd_reduce_det_kernel<true><<<grid, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C, scale,
            shift, mean, invstd, sums, ws, counters, npix, static_cast<const __nv_bfloat16*>(mask_src));
    else
        bn_bwd_reduce_det_kernel<false><<<grid, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C, scale,
            shift, mean, invstd, sums, ws, counters, npix, nullptr);
    return check_launch("bn_bwd_reduce_det_kernel");
}

int b200sr_head_bwd_det(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                        int64_t npix, float* ws, int64_t ws_floats, uint32_t* counter, void* stream) {
    B2_CHECK_ARG(dout && act && w && dact && dw && db && ws && counter && npix > 0);
    B2_CHECK_ARG(aligned16(act) && aligned16(dact) && aligned16(w));
    int grid = grid_for(npix * 8, 256, num_sms() * 4);
    if (grid > ws_floats / 72) grid = static_cast<int>(ws_floats / 72);
    B2_CHECK_ARG(grid >= 1);
    head_bwd_det_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, static_cast<const __nv_bfloat16*>(act), w, static_cast<__nv_bfloat16*>(dact), dw, db, ws, counter, npix);
    return check_launch("head_bwd_det_kernel");
}

int b200sr_head_bwd_bnred(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                          const void* z, const float* scale, const float* shift, const float* mean, const float* invstd,
                          float* sums, int64_t npix, float* ws, int64_t ws_floats, uint32_t* counters, void* stream) {
    B2_CHECK_ARG(dout && act && w && dact && dw && db && z && scale && shift && mean && invstd && sums && ws && counters);
    B2_CHECK_ARG(npix > 0 && aligned16(act) && aligned16(dact) && aligned16(w) && aligned16(z) && aligned16(ws));
    long long slices = num_sms() * 2;
    if (slices > (npix + 31) / 32) slices = (npix + 31) / 32;
    if (slices > ws_floats / 200) slices = ws_floats / 200;  // 72 (head) + 128 (BatchNorm) floats per block
    B2_CHECK_ARG(slices >= 1);  // (72 floats = 288 B per head slot: the BatchNorm slot area behind them stays 16-byte aligned)
    head_bwd_bnred_kernel<<<dim3(static_cast<unsigned>(slices), 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, static_cast<const __nv_bfloat16*>(act), w, static_cast<__nv_bfloat16*>(dact), dw, db,
        static_cast<const __nv_bfloat16*>(z), scale, shift, mean, invstd, sums, ws, counters, npix);
    return check_launch("head_bwd_bnred_kernel");
}

int b200sr_adam_step_auto(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, int32_t* step_dev, float grad_scale, void* stream) {
    B2_CHECK_ARG(p && g && m && v && n > 0 && step_dev);
    B2_CHECK_ARG(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v));
    adam_flat_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n, lr, beta1, beta2, eps, 1.f, 1.f, grad_scale, nullptr, step_dev);
    return check_launch("adam_flat_kernel");
}

// ---- fp32-accuracy eval mode (north star: fp32/tf32 tolerance 1e-4; BASELINE configs[0] is an fp32 forward) -------------
int b200sr_conv3x3_fwd_split(const void* x, int x_pix_stride, int x_c_off, int Cin3, const void* w_packed, int Cout,
                             int B, int H, int W, void* out, int out_pix_stride, int out_c_off, int part_stride,
                             const float* col_scale, const float* col_shift, int relu, void* stream) {
    B2_CHECK_ARG(Cin3 % 192 == 0 && part_stride >= Cout);
    return run_conv3(0, x, x_pix_stride, x_c_off, Cin3, w_packed, Cout, B, H, W, out, out_pix_stride, out_c_off,
                     col_scale, col_shift, relu, nullptr, 0, Cout, static_cast<cudaStream_t>(stream), nullptr, 0, 0,
                     part_stride);
}

int b200sr_convT2x2_fwd_split(const void* x, int x_pix_stride, int x_c_off, int Cin3, const void* w_packed, int Cout,
                              const float* bias, int B, int H, int W, void* out, int out_pix_stride, int out_c_off,
                              int part_stride, void* stream) {
    B2_CHECK_ARG(Cin3 % 192 == 0 && part_stride >= Cout);
    return run_conv3(1, x, x_pix_stride, x_c_off, Cin3, w_packed, 4 * Cout, B, H, W, out, out_pix_stride, out_c_off,
                     nullptr, bias, 0, nullptr, 0, Cout, static_cast<cudaStream_t>(stream), nullptr, 0, 0, part_stride);
}

int b200sr_conv1_fwd_split(const float* x, const float* w, const float* col_scale, const float* col_shift, int relu,
                           void* out, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x && w && col_scale && col_shift && out && B > 0 && H > 0 && W > 0 && aligned16(out));
    const long long total = static_cast<long long>(B) * H * W * 8;
    conv1_split_fwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w, col_scale, col_shift, relu, static_cast<__nv_bfloat16*>(out), H, W, total);
    return check_launch("conv1_split_fwd_kernel");
}

int b200sr_maxpool2x2_fwd_split(const void* in, int in_pix_stride, int in_c_off, int in_part_stride, int C, void* out,
                                int B, int H, int W, void* stream) {
    B2_CHECK_ARG(in && out && C % 8 == 0 && in_pix_stride % 8 == 0 && in_c_off % 8 == 0 && in_part_stride % 8 == 0);
    B2_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && aligned16(in) && aligned16(out));
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    maxpool2x2_split_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), in_pix_stride, in_c_off, in_part_stride, C,
        static_cast<__nv_bfloat16*>(out), H, W, total);
    return check_launch("maxpool2x2_split_kernel");
}

int b200sr_head_fwd_split(const void* act, const float* w, const float* b, float* out, int64_t npix, void* stream) {
    B2_CHECK_ARG(act && w && b && out && npix > 0 && aligned16(act) && aligned16(w));
    head_split_fwd_kernel<<<grid_for(npix * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), w, b, out, npix);
    return check_launch("head_split_fwd_kernel");
}

/* reference compute_metrics (src/VolumeVisualization.py:237-269) on the device: original / predicted (S,H,W) f32 ->
 * orig_norm / pred_norm (S,H,W) f32, per_slice [S][2] = {SSIM, PSNR}, out5 = {ssim_mean, ssim_std, psnr_mean, psnr_std,
 * mae}. ws: 2048 + 4 * S * ceil(H/32) * ceil(W/32) doubles; counters: 2 zero-initialised uint32. */
int b200sr_volume_metrics(const float* original, const float* predicted, int S, int H, int W, float* orig_norm,
                          float* pred_norm, float* out5, float* per_slice, double* ws, int64_t ws_doubles,
                          uint32_t* counters, void* stream) {
    B2_CHECK_ARG(original && predicted && orig_norm && pred_norm && out5 && per_slice && ws && counters);
    B2_CHECK_ARG(S > 0 && H >= 7 && W >= 7);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = static_cast<long long>(S) * H * W;
    const int grid_m = S * ((H + LS_T - 1) / LS_T) * ((W + LS_T - 1) / LS_T);
    B2_CHECK_ARG(ws_doubles >= 2048 + 4LL * grid_m);
    float* mm_part = reinterpret_cast<float*>(ws);  // [1024][2] floats = 1024 doubles
    float* mm = mm_part + 2048;                     // 2 floats
    const int g1 = grid_for(n, 256, 1024);
    volume_minmax_kernel<<<g1, 256, 0, st>>>(original, n, mm_part, counters, mm);
    volume_normalize_kernel<<<grid_for(n, 256), 256, 0, st>>>(original, predicted, mm, orig_norm, pred_norm, n);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mse_ssim_fast_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 LsFast<7>::SMEM_BYTES) != cudaSuccess)
            return fail(B200SR_ECUDA, "mse_ssim_fast smem attribute failed");
        configured = true;
    }
    LossArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pred = pred_norm;
    a.target = orig_norm;
    a.H = H;
    a.W = W;
    a.K = 7;
    for (int i = 0; i < LS_KMAX; ++i) a.win[i] = i < 7 ? 1.f / 7.f : 0.f;
    a.cov_norm = 49.f / 48.f;
    a.C1 = 0.01f * 0.01f;
    a.C2 = 0.03f * 0.03f;
    a.partials = ws + 2048;
    a.counter = counters + 1;
    a.out = out5;
    a.per_slice = per_slice;
    a.blocks_per_image = grid_m / S;
    a.nimages = S;
    mse_ssim_fast_kernel<7><<<grid_m, 256, LsFast<7>::SMEM_BYTES, st>>>(a);
    return check_launch("volume_metrics kernels");
}

/* MaxPool2d(2,2) backward + skip add (b200sr_maxpool2x2_bwd) fused with pass 1 of the BatchNorm+ReLU backward of the layer
 * the gradient flows into (b200sr_bn_bwd_reduce_det on the dy it writes): saves the re-read of dy and one launch. */
int b200sr_maxpool2x2_bwd_bnred(const void* act, int act_pix_stride, int act_c_off, const void* dpool, const void* dskip,
                                int dskip_pix_stride, int dskip_c_off, int C, void* dy, const void* z, const float* scale,
                                const float* shift, const float* mean, const float* invstd, float* sums, float* ws,
                                int64_t ws_floats, uint32_t* counters, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(act && dpool && dskip && dy && z && scale && shift && mean && invstd && sums && ws && counters);
    B2_CHECK_ARG(C % 64 == 0 && act_pix_stride % 8 == 0 && act_c_off % 8 == 0 && dskip_pix_stride % 8 == 0 &&
                 dskip_c_off % 8 == 0 && H % 2 == 0 && W % 2 == 0 && B > 0);
    B2_CHECK_ARG(aligned16(act) && aligned16(dpool) && aligned16(dskip) && aligned16(dy) && aligned16(z) && aligned16(ws));
    const int groups = C / 64;
    const long long nquads = static_cast<long long>(B) * (H / 2) * (W / 2);
    long long slices = (num_sms() * 2) / groups;
    if (slices > (nquads + 31) / 32) slices = (nquads + 31) / 32;
    if (slices > ws_floats / (128LL * groups)) slices = ws_floats / (128LL * groups);
    B2_CHECK_ARG(slices >= 1);
    dim3 grid(static_cast<unsigned>(slices), static_cast<unsigned>(groups));
    maxpool2x2_bwd_bnred_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), act_pix_stride, act_c_off, static_cast<const __nv_bfloat16*>(dpool),
        static_cast<const __nv_bfloat16*>(dskip), dskip_pix_stride, dskip_c_off, C, static_cast<__nv_bfloat16*>(dy),
        static_cast<const __nv_bfloat16*>(z), scale, shift, mean, invstd, sums, ws, counters, H, W, nquads);
    return check_launch("maxpool2x2_bwd_bnred_kernel");
}

// ---- DeepCNN residual baseline -------------------------------------------------------------
int b200sr_bn_bwd_masked(const void* dy, const void* z, const void* mask_src, int C, const float* scale,
                         const float* shift, const float* mean, const float* invstd, float* sums, int replicas,
                         double count, float* dgamma, float* dbeta, void* dz, int64_t npix, void* stream) {
    B2_CHECK_ARG(dy && z && mask_src && scale && shift && mean && invstd && sums && dgamma && dbeta && dz);
    B2_CHECK_ARG(C % 8 == 0 && npix > 0 && replicas > 0 && count > 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(z) && aligned16(dz) && aligned16(mask_src) && aligned16(sums));
    const int CV = C / 8;
    B2_CHECK_ARG(CV <= 256 && 256 % CV == 0);
    const int PB = 256 / CV;
    long long blocks = (npix + static_cast<long long>(PB) * BNB_UNROLL - 1) / (static_cast<long long>(PB) * BNB_UNROLL);
    if (blocks > num_sms() * bn_blocks_per_sm()) blocks = num_sms() * bn_blocks_per_sm();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bn_bwd_reduce_fast_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), C, 0, static_cast<const __nv_bfloat16*>(z), C, scale, shift, mean, invstd,
        sums, replicas, npix, static_cast<const __nv_bfloat16*>(mask_src));
    bn_bwd_apply_fused_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(dy), C, 0, static_cast<const __nv_bfloat16*>(z), C, scale, shift, mean, invstd,
        sums, replicas, static_cast<float>(count), dgamma, dbeta, static_cast<__nv_bfloat16*>(dz), npix,
        static_cast<const __nv_bfloat16*>(mask_src));
    return check_launch("bn_bwd_masked");
}

int b200sr_conv7_fwd(const float* x, const float* w, void* out, float* stats, int stats_replicas, int B, int H, int W,
                     void* stream) {
    B2_CHECK_ARG(x && w && out && B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(out));
    B2_CHECK_ARG(stats == nullptr || stats_replicas > 0);
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    const int grid = tiles < num_sms() * 4 ? tiles : num_sms() * 4;
    conv7_direct_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w, static_cast<__nv_bfloat16*>(out), stats, stats_replicas > 0 ? stats_replicas : 1, H, W, tiles);
    return check_launch("conv7_direct_fwd_kernel");
}

int b200sr_conv7_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x && dz && dw && B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(dz));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    const int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    conv7_direct_wgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, static_cast<const __nv_bfloat16*>(dz), dw, H, W, tiles);
    return check_launch("conv7_direct_wgrad_kernel");
}

int b200sr_maxpool3x3_fwd(const void* in, void* out, int C, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(in && out && C % 8 == 0 && B > 0 && H > 0 && W > 0 && aligned16(in) && aligned16(out));
    const long long total = static_cast<long long>(B) * H * W * (C / 8);
    maxpool3x3_fwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), C, H, W, total);
    return check_launch("maxpool3x3_fwd_kernel");
}

int b200sr_maxpool3x3_bwd(const void* in, const void* dout, void* din, int C, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(in && dout && din && C % 8 == 0 && B > 0 && H > 0 && W > 0);
    const long long total = static_cast<long long>(B) * H * W * (C / 2);
    maxpool3x3_bwd_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(din),
        C, H, W, total);
    return check_launch("maxpool3x3_bwd_kernel");
}

int b200sr_bn_add_relu(const void* z2, const float* scale2, const float* shift2, const void* identity,
                       const float* scale_d, const float* shift_d, void* out, int C, int64_t npix, void* stream) {
    B2_CHECK_ARG(z2 && scale2 && shift2 && identity && out && npix > 0 && (scale_d == nullptr) == (shift_d == nullptr));
    B2_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0);
    B2_CHECK_ARG(aligned16(z2) && aligned16(identity) && aligned16(out));
    const int PB = 256 / (C / 8);
    long long blocks = (npix + PB - 1) / PB;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    bn_add_relu_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(z2), scale2, shift2, static_cast<const __nv_bfloat16*>(identity), scale_d,
        shift_d, static_cast<__nv_bfloat16*>(out), C, npix);
    return check_launch("bn_add_relu_kernel");
}

int b200sr_add_masked(const void* a, const void* b, const void* mask, void* out, int64_t n, void* stream) {
    B2_CHECK_ARG(a && b && out && n > 0 && n % 8 == 0 && aligned16(a) && aligned16(b) && aligned16(out));
    add_masked_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
        static_cast<const __nv_bfloat16*>(mask), static_cast<__nv_bfloat16*>(out), n / 8);
    return check_launch("add_masked_kernel");
}

int b200sr_headw_fwd(const void* act, int C, const float* w, const float* b, float* out, int64_t npix, void* stream) {
    B2_CHECK_ARG(act && w && b && out && npix > 0 && aligned16(act) && C == 512);
    headw_fwd_kernel<512><<<grid_for(npix * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), w, b, out, npix);
    return check_launch("headw_fwd_kernel");
}

int b200sr_headw_bwd(const float* dout, const void* act, int C, const float* w, void* dact, float* dw, float* db,
                     int64_t npix, void* stream) {
    B2_CHECK_ARG(dout && act && w && dact && dw && db && npix > 0 && aligned16(act) && aligned16(dact) && C == 512);
    headw_bwd_kernel<512><<<grid_for(npix * 32, 256, 148 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, static_cast<const __nv_bfloat16*>(act), w, static_cast<__nv_bfloat16*>(dact), dw, db, npix);
    return check_launch("headw_bwd_kernel");
}

/* Conv2d 1x1 forward / dgrad: D[pixel, n] = sum_c A[pixel, c] * Wp[n, c] (Wp: PACK_CONV1X1_FWD / _DGRAD) */
int b200sr_conv1x1(const void* a, int a_pix_stride, int a_c_off, int Ca, const void* w_packed, int N, int B, int H,
                   int W, void* out, int out_pix_stride, int out_c_off, float* stats, int stats_replicas, void* stream) {
    return run_conv3(3, a, a_pix_stride, a_c_off, Ca, w_packed, N, B, H, W, out, out_pix_stride, out_c_off, nullptr,
                     nullptr, 0, stats, stats_replicas, N, static_cast<cudaStream_t>(stream));
}

int b200sr_conv1x1_wgrad(const void* x, int x_pix_stride, int x_c_off, int Cin, const void* dz, int dz_pix_stride,
                         int dz_c_off, int Cout, int B, int H, int W, float* G, void* stream) {
    return run_wgrad(0, x, x_pix_stride, x_c_off, Cin, 1, dz, dz_pix_stride, dz_c_off, Cout, B, H, W, G,
                     static_cast<cudaStream_t>(stream));
}

// ---- Fast-DDPM denoiser (reference src/ModelLoader.py:471-636) -------------------------------------------------
int b200sr_fd_time_mlp_fwd(const int64_t* t, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* emb, float* hid, float* e, int B, void* stream) {
    B2_CHECK_ARG(t && w1 && b1 && w2 && b2 && emb && hid && e && B > 0);
    fd_time_mlp_fwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(t), w1, b1, w2, b2, emb, hid, e);
    return check_launch("fd_time_mlp_fwd_kernel");
}

int b200sr_fd_time_bias(const float* e, const float* w, const float* bias, float* tb, int B, void* stream) {
    B2_CHECK_ARG(e && w && bias && tb && B > 0);
    fd_time_bias_kernel<<<B, 576, 0, static_cast<cudaStream_t>(stream)>>>(e, w, bias, tb);
    return check_launch("fd_time_bias_kernel");
}

int b200sr_fd_convin_fwd(const float* x0, const float* noise, const float* coef, const float* cond, const float* w,
                         const float* tb, void* out, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x0 && cond && w && tb && out && (noise == nullptr || coef != nullptr));
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(out) && aligned16(tb));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    FirstConvSrc src{x0, noise, reinterpret_cast<const float2*>(coef), cond};
    const int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    first_conv_mma_fwd_kernel<3><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, w, FD_CIN0 * 9, tb, nullptr, nullptr, 1, static_cast<__nv_bfloat16*>(out), nullptr, 1, 0, H, W, tiles);
    return check_launch("first_conv_mma_fwd_kernel<3>");
}

int b200sr_fd_convin_wgrad(const float* x0, const float* noise, const float* coef, const float* cond, const void* dz,
                           float* dw, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(x0 && cond && dz && dw && (noise == nullptr || coef != nullptr));
    B2_CHECK_ARG(B > 0 && H % C1_TILE == 0 && W % C1_TILE == 0 && aligned16(dz));
    const int tiles = B * (H / C1_TILE) * (W / C1_TILE);
    FirstConvSrc src{x0, noise, reinterpret_cast<const float2*>(coef), cond};
    const int grid = tiles < num_sms() * 2 ? tiles : num_sms() * 2;
    first_conv_mma_wgrad_kernel<3><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<const __nv_bfloat16*>(dz), dw, FD_CIN0 * 9, H, W, tiles, nullptr);
    return check_launch("first_conv_mma_wgrad_kernel<3>");
}

int b200sr_fd_relu_bwd_bias(const void* dy, int dy_pix_stride, int dy_c_off, const void* act, int act_pix_stride,
                            int act_c_off, void* dz, float* ps, int C, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(dy && act && dz && ps && B > 0 && H > 0 && W > 0);
    B2_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0);
    B2_CHECK_ARG(dy_pix_stride % 8 == 0 && dy_c_off % 8 == 0 && act_pix_stride % 8 == 0 && act_c_off % 8 == 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(act) && aligned16(dz));
    const int HW = H * W;
    const int rows = 256 / (C / 8);
    int bx = (HW + rows * 8 - 1) / (rows * 8);  // >= 8 pixels per thread
    const int cap = (num_sms() * 8 + B - 1) / B;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, B);
    fd_relu_bwd_bias_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(act),
        act_pix_stride, act_c_off, static_cast<__nv_bfloat16*>(dz), ps, C, HW);
    return check_launch("fd_relu_bwd_bias_kernel");
}

namespace {
int fd_grid_x(int work_items, int rows, int B) {
    int bx = (work_items + rows * 8 - 1) / (rows * 8);
    const int cap = (num_sms() * 8 + B - 1) / B;
    if (bx > cap) bx = cap;
    return bx < 1 ? 1 : bx;
}
}  // namespace

int b200sr_fd_upsample2x_bwd_relu(const void* dout, int dout_pix_stride, int dout_c_off, const void* act, void* dz,
                                  float* ps, int C, int B, int h, int w, void* stream) {
    B2_CHECK_ARG(dout && act && dz && ps && B > 0 && h > 0 && w > 0);
    B2_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && dout_pix_stride % 8 == 0 && dout_c_off % 8 == 0);
    B2_CHECK_ARG(aligned16(dout) && aligned16(act) && aligned16(dz));
    dim3 grid(fd_grid_x(h * w, 256 / (C / 8), B), B);
    fd_upsample2x_bwd_relu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dout), dout_pix_stride, dout_c_off, static_cast<const __nv_bfloat16*>(act),
        static_cast<__nv_bfloat16*>(dz), ps, C, h, w);
    return check_launch("fd_upsample2x_bwd_relu_kernel");
}

int b200sr_fd_head_bwd_relu(const float* dout, const void* act, const float* w, void* dz, float* dw, float* db, float* ps,
                            int B, int H, int W, void* stream) {
    B2_CHECK_ARG(dout && act && w && dz && dw && db && ps && B > 0 && H > 0 && W > 0);
    B2_CHECK_ARG(aligned16(act) && aligned16(dz) && aligned16(w));
    dim3 grid(fd_grid_x(H * W, 32, B), B);
    fd_head_bwd_relu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, static_cast<const __nv_bfloat16*>(act), w, static_cast<__nv_bfloat16*>(dz), dw, db, ps, H * W);
    return check_launch("fd_head_bwd_relu_kernel");
}

int b200sr_fd_maxpool2x2_bwd_relu(const void* act, int act_pix_stride, int act_c_off, const void* dpool, const void* dskip,
                                  int dskip_pix_stride, int dskip_c_off, int C, void* dz, float* ps, int B, int H, int W,
                                  void* stream) {
    B2_CHECK_ARG(act && dpool && dskip && dz && ps && B > 0 && H % 2 == 0 && W % 2 == 0);
    B2_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0);
    B2_CHECK_ARG(act_pix_stride % 8 == 0 && act_c_off % 8 == 0 && dskip_pix_stride % 8 == 0 && dskip_c_off % 8 == 0);
    B2_CHECK_ARG(aligned16(act) && aligned16(dpool) && aligned16(dskip) && aligned16(dz));
    dim3 grid(fd_grid_x((H / 2) * (W / 2), 256 / (C / 8), B), B);
    fd_maxpool2x2_bwd_relu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), act_pix_stride, act_c_off, static_cast<const __nv_bfloat16*>(dpool),
        static_cast<const __nv_bfloat16*>(dskip), dskip_pix_stride, dskip_c_off, C, static_cast<__nv_bfloat16*>(dz), ps, H,
        W);
    return check_launch("fd_maxpool2x2_bwd_relu_kernel");
}

int b200sr_fd_bias_finish(const b200sr_fd_bias_job* jobs, int njobs, int B, void* stream) {
    B2_CHECK_ARG(jobs && njobs > 0 && B > 0);
    static_assert(sizeof(b200sr_fd_bias_job) == sizeof(FdBiasJob), "FdBiasJob ABI mismatch");
    fd_bias_finish_kernel<<<njobs, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const FdBiasJob*>(jobs), B);
    return check_launch("fd_bias_finish_kernel");
}

int b200sr_fd_time_bwd(const void* dz, const float* ps, const float* e, const float* emb, const float* hid,
                       const float* w_in, const float* w2, float* S, float* de, float* dh, float* dw_in, float* dw1,
                       float* db1, float* dw2, float* db2, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(dz && ps && e && emb && hid && w_in && w2 && S && de && dh && dw_in && dw1 && db1 && dw2 && db2);
    B2_CHECK_ARG(B > 0 && H >= 2 && W >= 2 && aligned16(dz));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fd_tap_sums_kernel<<<B, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dz), ps, S, H, W);
    fd_time_wgrad_kernel<<<C1_COUT, 256, 0, st>>>(e, S, dw_in, B);
    fd_time_dgrad_kernel<<<B, 256, 0, st>>>(w_in, S, de);
    fd_time_mlp_bwd_dh_kernel<<<B, 256, 0, st>>>(de, hid, w2, dh);
    fd_time_mlp_bwd_w_kernel<<<dim3(FD_TDIM, 2), 256, 0, st>>>(de, dh, hid, emb, dw1, db1, dw2, db2, B);
    return check_launch("fd_time_bwd kernels");
}

int b200sr_fd_upsample2x_fwd(const void* in, int C, void* out, int out_pix_stride, int out_c_off, int B, int h, int w,
                             void* stream) {
    B2_CHECK_ARG(in && out && C % 8 == 0 && out_pix_stride % 8 == 0 && out_c_off % 8 == 0 && B > 0 && h > 0 && w > 0);
    B2_CHECK_ARG(aligned16(in) && aligned16(out));
    const long long total = static_cast<long long>(B) * h * w * (C / 8);
    B2_CHECK_ARG(total < (1ll << 31));
    const long long blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < num_sms() * 16 ? blocks : num_sms() * 16);
    fd_upsample2x_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), C, static_cast<__nv_bfloat16*>(out), out_pix_stride, out_c_off, h, w,
        static_cast<unsigned>(total));
    return check_launch("fd_upsample2x_fwd_kernel");
}

int b200sr_fd_upsample2x_bwd(const void* dout, int dout_pix_stride, int dout_c_off, int C, void* din, int B, int h, int w,
                             void* stream) {
    B2_CHECK_ARG(dout && din && C % 8 == 0 && dout_pix_stride % 8 == 0 && dout_c_off % 8 == 0 && B > 0 && h > 0 && w > 0);
    B2_CHECK_ARG(aligned16(dout) && aligned16(din));
    const long long total = static_cast<long long>(B) * h * w * (C / 8);
    B2_CHECK_ARG(total < (1ll << 31));
    const long long blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < num_sms() * 16 ? blocks : num_sms() * 16);
    fd_upsample2x_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dout), dout_pix_stride, dout_c_off, C, static_cast<__nv_bfloat16*>(din), h, w,
        static_cast<unsigned>(total));
    return check_launch("fd_upsample2x_bwd_kernel");
}

int b200sr_fd_q_sample(const float* x0, const float* noise, const float* coef, float* out, int B, int H, int W,
                       void* stream) {
    B2_CHECK_ARG(x0 && noise && coef && out && B > 0 && H > 0 && W > 0);
    const long long total = static_cast<long long>(B) * H * W;
    const long long blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < num_sms() * 8 ? blocks : num_sms() * 8);
    fd_q_sample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x0, noise, reinterpret_cast<const float2*>(coef),
                                                                            out, H * W, total);
    return check_launch("fd_q_sample_kernel");
}

int b200sr_fd_ddim_update(float* x, const float* eps, float a_bar, float a_bar_prev, int clamp, int64_t n, void* stream) {
    B2_CHECK_ARG(x && eps && n > 0 && a_bar > 0.f && a_bar <= 1.f && a_bar_prev > 0.f && a_bar_prev <= 1.f);
    const long long blocks = (n + 255) / 256;
    const int grid = static_cast<int>(blocks < num_sms() * 8 ? blocks : num_sms() * 8);
    fd_ddim_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, eps, sqrtf(1.f - a_bar), sqrtf(a_bar), sqrtf(a_bar_prev), sqrtf(1.f - a_bar_prev), clamp, n);
    return check_launch("fd_ddim_update_kernel");
}

int b200sr_grad_clip(float* g, int64_t n, double* sumsq, float max_norm, float pre_scale, void* stream) {
    B2_CHECK_ARG(g && sumsq && n > 0 && max_norm > 0.f && pre_scale > 0.f);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(sumsq, 0, sizeof(double), st) != cudaSuccess) return check_launch("grad_clip memset");
    const long long blocks = (n + 255) / 256;
    const int grid = static_cast<int>(blocks < num_sms() * 4 ? blocks : num_sms() * 4);
    fd_sumsq_kernel<<<grid, 256, 0, st>>>(g, n, sumsq);
    fd_clip_scale_kernel<<<grid, 256, 0, st>>>(g, n, sumsq, max_norm, pre_scale);
    return check_launch("grad_clip kernels");
}

int b200sr_maxpool2x2_fwd(const void* in, int in_pix_stride, int in_c_off, int C, void* out, int B, int H, int W,
                          void* stream) {
    B2_CHECK_ARG(in && out && C % 8 == 0 && in_pix_stride % 8 == 0 && in_c_off % 8 == 0 && H % 2 == 0 && W % 2 == 0);
    B2_CHECK_ARG(aligned16(in) && aligned16(out));
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    maxpool2x2_fwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), in_pix_stride, in_c_off, C, static_cast<__nv_bfloat16*>(out), H, W,
        total);
    return check_launch("maxpool2x2_fwd_kernel");
}

int b200sr_maxpool2x2_bwd(const void* act, int act_pix_stride, int act_c_off, const void* dpool, const void* dskip,
                          int dskip_pix_stride, int dskip_c_off, int C, void* dy, int B, int H, int W, void* stream) {
    B2_CHECK_ARG(act && dpool && dy && C % 8 == 0 && act_pix_stride % 8 == 0 && act_c_off % 8 == 0);
    B2_CHECK_ARG(dskip == nullptr || (dskip_pix_stride % 8 == 0 && dskip_c_off % 8 == 0 && aligned16(dskip)));
    B2_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && aligned16(act) && aligned16(dpool) && aligned16(dy));
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    maxpool2x2_bwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), act_pix_stride, act_c_off, static_cast<const __nv_bfloat16*>(dpool),
        static_cast<const __nv_bfloat16*>(dskip), dskip_pix_stride, dskip_c_off, C, static_cast<__nv_bfloat16*>(dy), H,
        W, total);
    return check_launch("maxpool2x2_bwd_kernel");
}

int b200sr_bn_bwd_reduce(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C, const float* scale,
                         const float* shift, const float* mean, const float* invstd, float* sums, int replicas,
                         int64_t npix, void* stream) {
    B2_CHECK_ARG(dy && z && scale && shift && mean && invstd && sums);
    B2_CHECK_ARG(C % 64 == 0 && dy_pix_stride % 8 == 0 && dy_c_off % 8 == 0 && replicas > 0 && npix > 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(z));
    const int CV = C / 8;
    if (CV <= 256 && 256 % CV == 0 && getenv("B200SR_BN_SLOW") == nullptr) {
        const int PB = 256 / CV;
        long long blocks = (npix + static_cast<long long>(PB) * BNB_UNROLL - 1) / (static_cast<long long>(PB) * BNB_UNROLL);
        if (blocks > num_sms() * bn_blocks_per_sm()) blocks = num_sms() * bn_blocks_per_sm();
        bn_bwd_reduce_fast_kernel<false><<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C,
            scale, shift, mean, invstd, sums, replicas, npix, nullptr);
        return check_launch("bn_bwd_reduce_fast_kernel");
    }
    const int cg = C / 64;
    long long slices = (npix + 32 * 8 - 1) / (32 * 8);  // >= 8 pixels per thread row
    const long long max_slices = (148 * 8 + cg - 1) / cg;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    bn_bwd_reduce_kernel<<<static_cast<int>(slices) * cg, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C, scale,
        shift, mean, invstd, sums, replicas, npix);
    return check_launch("bn_bwd_reduce_kernel");
}

int b200sr_bn_bwd_finalize(const float* sums, int replicas, int C, double count, float* c1, float* c2, float* dgamma,
                           float* dbeta, void* stream) {
    B2_CHECK_ARG(sums && c1 && c2 && dgamma && dbeta && replicas > 0 && C > 0 && count > 0);
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        sums, replicas, C, static_cast<float>(count), c1, c2, dgamma, dbeta);
    return check_launch("bn_bwd_finalize_kernel");
}

int b200sr_bn_bwd_apply(const void* dy, int dy_pix_stride, int dy_c_off, const void* z, int C, const float* scale,
                        const float* shift, const float* mean, const float* invstd, const float* c1, const float* c2,
                        void* dz, int64_t npix, void* stream) {
    B2_CHECK_ARG(dy && z && scale && shift && mean && invstd && c1 && c2 && dz);
    B2_CHECK_ARG(C % 8 == 0 && dy_pix_stride % 8 == 0 && dy_c_off % 8 == 0 && npix > 0);
    B2_CHECK_ARG(aligned16(dy) && aligned16(z) && aligned16(dz));
    const int CV = C / 8;
    if (CV <= 256 && 256 % CV == 0 && getenv("B200SR_BN_SLOW") == nullptr) {
        const int PB = 256 / CV;
        long long blocks = (npix + static_cast<long long>(PB) * BNB_UNROLL - 1) / (static_cast<long long>(PB) * BNB_UNROLL);
        if (blocks > num_sms() * bn_blocks_per_sm()) blocks = num_sms() * bn_blocks_per_sm();
        bn_bwd_apply_fast_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C,
            scale, shift, mean, invstd, c1, c2, static_cast<__nv_bfloat16*>(dz), npix);
        return check_launch("bn_bwd_apply_fast_kernel");
    }
    const long long total = npix * (C / 8);
    bn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), dy_pix_stride, dy_c_off, static_cast<const __nv_bfloat16*>(z), C, scale,
        shift, mean, invstd, c1, c2, static_cast<__nv_bfloat16*>(dz), total);
    return check_launch("bn_bwd_apply_kernel");
}

int b200sr_relu_bwd(const void* dy, const void* act, void* out, int64_t n, void* stream) {
    B2_CHECK_ARG(dy && act && out && n > 0 && n % 8 == 0 && aligned16(dy) && aligned16(act) && aligned16(out));
    relu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(act), static_cast<__nv_bfloat16*>(out),
        n / 8);
    return check_launch("relu_bwd_kernel");
}

int b200sr_feat_mse_grad(const void* fp, const void* ft, void* grad, double* sums, float gscale, int64_t n,
                         void* stream) {
    B2_CHECK_ARG(fp && ft && sums && n > 0 && n % 8 == 0 && aligned16(fp) && aligned16(ft));
    B2_CHECK_ARG(grad == nullptr || aligned16(grad));
    feat_mse_grad_kernel<<<grid_for(n / 8, 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(fp), static_cast<const __nv_bfloat16*>(ft), static_cast<__nv_bfloat16*>(grad),
        sums, gscale, n / 8);
    return check_launch("feat_mse_grad_kernel");
}

int b200sr_head_fwd(const void* act, const float* w, const float* b, float* out, int64_t npix, void* stream) {
    B2_CHECK_ARG(act && w && b && out && npix > 0 && aligned16(act) && aligned16(w));
    head_fwd_kernel<<<grid_for(npix * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(act), w, b, out, npix);
    return check_launch("head_fwd_kernel");
}

int b200sr_head_bwd(const float* dout, const void* act, const float* w, void* dact, float* dw, float* db,
                    int64_t npix, void* stream) {
    B2_CHECK_ARG(dout && act && w && dact && dw && db && npix > 0 && aligned16(act) && aligned16(dact) && aligned16(w));
    head_bwd_kernel<<<grid_for(npix * 8, 256, 148 * 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, static_cast<const __nv_bfloat16*>(act), w, static_cast<__nv_bfloat16*>(dact), dw, db, npix);
    return check_launch("head_bwd_kernel");
}

namespace {
int run_mse_ssim(const float* pred, const float* target, float* grad, double* sums, int B, int H, int W, const float* win,
                 int K, float cov_norm, float C1, float C2, float w_mse, float w_ssim, double* partials, int64_t ws_doubles,
                 uint32_t* counter, float* out, cudaStream_t st) {
    B2_CHECK_ARG(pred && target && win && (sums != nullptr || partials != nullptr));
    B2_CHECK_ARG(B > 0 && K >= 1 && K <= LS_KMAX && H >= K && W >= K);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mse_ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM_BYTES);
        if (e != cudaSuccess) return fail(B200SR_ECUDA, std::string("mse_ssim smem attribute: ") + cudaGetErrorString(e));
        configured = true;
    }
    LossArgs a;
    a.pred = pred;
    a.target = target;
    a.grad = grad;
    a.sums = sums;
    a.H = H;
    a.W = W;
    a.K = K;
    for (int i = 0; i < LS_KMAX; ++i) a.win[i] = i < K ? win[i] : 0.f;
    a.cov_norm = cov_norm;
    a.C1 = C1;
    a.C2 = C2;
    a.g_mse = static_cast<float>(static_cast<double>(w_mse) * 2.0 / (static_cast<double>(B) * H * W));
    a.g_ssim = static_cast<float>(-static_cast<double>(w_ssim) /
                                  (static_cast<double>(B) * (H - K + 1) * static_cast<double>(W - K + 1)));
    a.partials = partials;
    a.counter = counter;
    a.out = out;
    a.inv_n_mse = 1.0 / (static_cast<double>(B) * H * W);
    a.inv_n_ssim = 1.0 / (static_cast<double>(B) * (H - K + 1) * static_cast<double>(W - K + 1));
    a.w_mse = w_mse;
    a.w_ssim = w_ssim;
    a.per_slice = nullptr;
    a.blocks_per_image = 0;
    a.nimages = B;
    const int grid = B * ((H + LS_T - 1) / LS_T) * ((W + LS_T - 1) / LS_T);
    if (partials != nullptr) B2_CHECK_ARG(counter != nullptr && out != nullptr && ws_doubles >= 2LL * grid);
    if ((K == 11 || K == 7) && getenv("B200SR_SSIM_GENERIC") == nullptr) {
        static bool configured_fast = false;
        if (!configured_fast) {
            cudaError_t e1 = cudaFuncSetAttribute(mse_ssim_fast_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  LsFast<11>::SMEM_BYTES);
            cudaError_t e2 = cudaFuncSetAttribute(mse_ssim_fast_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  LsFast<7>::SMEM_BYTES);
            if (e1 != cudaSuccess || e2 != cudaSuccess)
                return fail(B200SR_ECUDA, "mse_ssim_fast smem attribute failed");
            configured_fast = true;
        }
        if (K == 11)
            mse_ssim_fast_kernel<11><<<grid, 256, LsFast<11>::SMEM_BYTES, st>>>(a);
        else
            mse_ssim_fast_kernel<7><<<grid, 256, LsFast<7>::SMEM_BYTES, st>>>(a);
        return check_launch("mse_ssim_fast_kernel");
    }
    mse_ssim_kernel<<<grid, 256, LS_SMEM_BYTES, st>>>(a);
    return check_launch("mse_ssim_kernel");
}
}  // namespace

int b200sr_mse_ssim(const float* pred, const float* target, float* grad, double* sums, int B, int H, int W,
                    const float* win, int K, float cov_norm, float C1, float C2, float w_mse, float w_ssim,
                    void* stream) {
    return run_mse_ssim(pred, target, grad, sums, B, H, W, win, K, cov_norm, C1, C2, w_mse, w_ssim, nullptr, 0, nullptr,
                        nullptr, static_cast<cudaStream_t>(stream));
}

int b200sr_mse_ssim_det(const float* pred, const float* target, float* grad, float* out3, int B, int H, int W,
                        const float* win, int K, float cov_norm, float C1, float C2, float w_mse, float w_ssim,
                        double* ws, int64_t ws_doubles, uint32_t* counter, void* stream) {
    B2_CHECK_ARG(ws != nullptr && out3 != nullptr && counter != nullptr);
    return run_mse_ssim(pred, target, grad, nullptr, B, H, W, win, K, cov_norm, C1, C2, w_mse, w_ssim, ws, ws_doubles,
                        counter, out3, static_cast<cudaStream_t>(stream));
}

int b200sr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, int64_t step, float grad_scale, void* stream) {
    B2_CHECK_ARG(p && g && m && v && n > 0 && step >= 1);
    B2_CHECK_ARG(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v));
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
    adam_flat_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n, lr, beta1, beta2, eps, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_scale,
        nullptr, nullptr);
    return check_launch("adam_flat_kernel");
}

int b200sr_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                         float eps, const float* bias_corr, float grad_scale, void* stream) {
    B2_CHECK_ARG(p && g && m && v && n > 0 && bias_corr);
    B2_CHECK_ARG(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v));
    adam_flat_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n, lr, beta1, beta2, eps, 1.f, 1.f, grad_scale, bias_corr, nullptr);
    return check_launch("adam_flat_kernel");
}

int b200sr_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, void* stream) {
    B2_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0);
    const long long total = static_cast<long long>(B) * C * H * W;
    nchw_f32_to_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, static_cast<__nv_bfloat16*>(out), C, H * W, total);
    return check_launch("nchw_f32_to_nhwc_bf16_kernel");
}

int b200sr_nhwc_bf16_to_nchw_f32(const void* in, int in_pix_stride, int in_c_off, float* out, int B, int C, int H,
                                 int W, void* stream) {
    B2_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0);
    const long long total = static_cast<long long>(B) * C * H * W;
    nhwc_bf16_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), in_pix_stride, in_c_off, out, C, H * W, total);
    return check_launch("nhwc_bf16_to_nchw_f32_kernel");
}

}  // extern "C"
