// C ABI of libflite_b200.so (see include/flite_b200.h).  Host-side launch code only: argument checks,
// TMA descriptor construction (driver entry point fetched at run time so the library also loads on a
// machine without libcuda), kernel selection and launch.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/flite_b200.h"
#include "attn_cg2_sm100.cuh"
#include "attn_xres_sm100.cuh"
#include "attn_qtmem_sm100.cuh"
#include "attn_sk_sm100.cuh"
#include "attn_sm100.cuh"
#include "elementwise.cuh"
#include "gemm_sm100.cuh"
#include "groupnorm.cuh"

using namespace flite;

namespace {

thread_local char g_err[512] = "";
int g_tuning[32] = {0};   // FLITE_TUNE_* knobs (A/B switches for benchmarking)

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return fail(FLITE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define LAUNCH_CHECK()                                                                          \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) return fail(FLITE_ERR_CUDA, "launch: %s", cudaGetErrorString(_e)); \
    } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct TmapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ k.rows;
        h = h * 1000003u ^ k.cols;
        h = h * 1000003u ^ k.ld;
        h = h * 1000003u ^ k.box_rows;
        return h;
    }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

// 2-D bf16 row-major tensor [rows, cols] with row stride ld (elements); box = 64 columns (128 B, SWIZZLE_128B)
// x box_rows rows; out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    TmapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(g_tmap_mu);
        auto it = g_tmap_cache.find(key);
        if (it != g_tmap_cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(FLITE_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16)
        return fail(FLITE_ERR_INVALID, "TMA operand must be 16-byte aligned (ptr %p, ld %llu)", ptr,
                    (unsigned long long)ld);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FLITE_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
    return 0;
}

int cur_dev() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < 64) ? dev : 0;
}

// per-DEVICE caches (a process may drive more than one GPU): SM count, and "max dynamic smem attribute already set"
int num_sms() {
    static int n[64] = {0};
    const int dev = cur_dev();
    if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    return n[dev];
}

// Launch attributes of the hot-loop kernels: optional cluster dimension + programmatic dependent launch (PDL).
int fill_launch_attrs(cudaLaunchAttribute* attr, int cluster_x) {
    int n = 0;
    if (cluster_x > 0) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (g_tuning[FLITE_TUNE_PDL] == 1) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    return n;
}

template <int kCtaGroup, int BLOCK_N, int kStages, int kEpi>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const CUtensorMap& tah,
                GemmParams p, cudaStream_t stream) {
    using S = GemmSmem<kCtaGroup, BLOCK_N, kStages>;
    auto kern = gemm_bf16_kernel<kCtaGroup, BLOCK_N, kStages, kEpi>;
    static bool configured[64] = {false};
    if (!configured[cur_dev()]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kEpi == EPI_QKV_ROPE ? S::TOTAL_QKV : S::TOTAL));
        configured[cur_dev()] = true;
    }
    const int tile_m = 128 * kCtaGroup;
    const int num_tiles = ((p.M + tile_m - 1) / tile_m) * (p.N / BLOCK_N);
    const int max_clusters = num_sms() / kCtaGroup;
    // tail balancing: a last partial wave of `tail` tiles is run as 2*tail half-width tiles when that still fits in
    // one wave (needs 128-column granularity for SwiGLU pairs, whole heads for the QKV epilogue => not there)
    p.full_units = num_tiles;
    p.num_units = num_tiles;
    const bool can_split = kEpi != EPI_QKV_ROPE && (BLOCK_N / 2) % (kEpi == EPI_SWIGLU ? 128 : 32) == 0 &&
                           (BLOCK_N / 2 / kCtaGroup) >= 8 && g_tuning[FLITE_TUNE_GEMM_TAIL_SPLIT] == 0;
    const int tail = num_tiles % max_clusters;
    if (can_split && tail > 0 && 2 * tail <= max_clusters) {
        p.full_units = num_tiles - tail;
        p.num_units = p.full_units + 2 * tail;
    }
    // Band height (measured on B200, profiles/r1_gemm_band_sweep.json): when A and W together fit in L2 the order does
    // not matter (single band); when W alone fits (<= 96 MB) thin bands keep W resident while A streams once; otherwise
    // a band of A of ~32 MB stays resident while W streams past it once per band.
    {
        const int num_m_tiles = (p.M + tile_m - 1) / tile_m;
        const long long a_tile_bytes = (long long)tile_m * p.K * 2;
        const long long a_bytes = (long long)num_m_tiles * a_tile_bytes, w_bytes = (long long)p.N * p.K * 2;
        long long g;
        if (a_bytes + w_bytes <= (100ll << 20) || g_tuning[FLITE_TUNE_GEMM_BAND] == 1) g = num_m_tiles;
        else if (g_tuning[FLITE_TUNE_GEMM_BAND] > 1) g = g_tuning[FLITE_TUNE_GEMM_BAND];   // explicit band height (tuning)
        else if (w_bytes <= (96ll << 20)) g = 2;
        else {
            // a band of A of <= ~32 MB stays L2-resident while W streams past it once per band; the bands are BALANCED
            // (33 M-tiles -> 17 + 16, not 16 + 16 + 1: a one-tile last band re-streams the whole of W from DRAM for 3 %
            // of the work -- r1f ncu: 518 MB read per gate|up launch for 201 MB of operands)
            long long g_max = (32ll << 20) / a_tile_bytes;
            if (g_max < 1) g_max = 1;
            if (g_max > 24) g_max = 24;
            const long long n_bands = (num_m_tiles + g_max - 1) / g_max;
            g = (num_m_tiles + n_bands - 1) / n_bands;
        }
        if (g > num_m_tiles) g = num_m_tiles;
        p.band_m = (int)g;
        // L2 hints (FLITE_TUNE_GEMM_HINT_A / _B: 0 auto, 1 none, 2 evict_first, 3 evict_last)
        auto policy = [](int v) -> unsigned long long {
            return v == 2 ? L2_EVICT_FIRST : v == 3 ? L2_EVICT_LAST : 0ull;
        };
        p.hint_a = policy(g_tuning[FLITE_TUNE_GEMM_HINT_A]);
        p.hint_b = policy(g_tuning[FLITE_TUNE_GEMM_HINT_B]);
        p.debug = g_tuning[FLITE_TUNE_GEMM_DEBUG];
        // narrow last M-tile (<= 128 valid rows of the 256): M = 128 MMAs, half the padding cost.  Not for the QKV epilogue
        // (needs a whole head per thread pair), and only when the first band (which holds the narrow tile) is made of
        // whole-width units.
        const int tail_rows = p.M % tile_m;
        p.narrow_m = (kCtaGroup == 2 && kEpi != EPI_QKV_ROPE && tail_rows > 0 && tail_rows <= tile_m / 2 &&
                      num_m_tiles >= 2 && g_tuning[FLITE_TUNE_GEMM_NARROW_M] == 0 &&
                      p.full_units >= (int)g * (p.N / BLOCK_N)) ? 1 : 0;
    }
    int clusters = max_clusters;
    if (clusters > p.num_units) clusters = p.num_units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCtaGroup);
    cfg.blockDim = dim3(gemm_threads(kEpi));
    cfg.dynamicSmemBytes = (kEpi == EPI_QKV_ROPE) ? S::TOTAL_QKV : S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, kCtaGroup);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, tbh, tah, p));
    return 0;
}

template <int kCtaGroup, int BLOCK_N, int kStages>
int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const CUtensorMap& tah,
                 const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case EPI_STORE: return launch_gemm<kCtaGroup, BLOCK_N, kStages, EPI_STORE>(ta, tb, tbh, tah, p, s);
        case EPI_GATED_RES: return launch_gemm<kCtaGroup, BLOCK_N, kStages, EPI_GATED_RES>(ta, tb, tbh, tah, p, s);
        case EPI_SWIGLU:
            if constexpr (BLOCK_N % 128 == 0) return launch_gemm<kCtaGroup, BLOCK_N, kStages, EPI_SWIGLU>(ta, tb, tbh, tah, p, s);
            break;
        case EPI_QKV_ROPE:
            if constexpr (BLOCK_N == 256) return launch_gemm<kCtaGroup, BLOCK_N, kStages, EPI_QKV_ROPE>(ta, tb, tbh, tah, p, s);
            break;
    }
    return fail(FLITE_ERR_INVALID, "epilogue %d not available for N-tile %d", epi, BLOCK_N);
}

}  // namespace

extern "C" {

int flite_abi_version(void) { return FLITE_ABI_VERSION; }
const char* flite_last_error(void) { return g_err; }

int flite_set_tuning(int key, int value) {
    if (key < 0 || key >= 32) return fail(FLITE_ERR_INVALID, "set_tuning: unknown key %d", key);
    g_tuning[key] = value;
    return 0;
}

int flite_get_tuning(int key) { return (key >= 0 && key < 32) ? g_tuning[key] : 0; }

int flite_check_device(void) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(FLITE_ERR_UNSUPPORTED, "device is sm_%d%d, need sm_100", prop.major, prop.minor);
    if (!get_encode_fn()) return fail(FLITE_ERR_UNSUPPORTED, "driver does not export cuTensorMapEncodeTiled");
    return 0;
}

int flite_watchdog_status(unsigned int* code_out) {
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned int code = 0, zero = 0;
    CUDA_TRY(cudaMemcpyFromSymbol(&code, g_flite_abort, sizeof(code)));
    if (code_out) *code_out = code;
    if (code != 0) {
        CUDA_TRY(cudaMemcpyToSymbol(g_flite_abort, &zero, sizeof(zero)));
        const unsigned tag = (code >> 16) & 0x7fff;
        if (tag >= 94 && tag <= 98)
            return fail(FLITE_ERR_WATCHDOG, "kernel precondition failed (tag %u: 94 = stream-K share shorter than a unit, 95 = "
                        "stream-K attention called with non-uniform sequence lengths, 96 = a sequence has more than 256 keys "
                        "in the resident-K/V attention, 97 / 98 = shared-memory window misaligned)", tag);
        return fail(FLITE_ERR_WATCHDOG, "kernel barrier wait timed out: tag %u block %u", tag, code & 0xffff);
    }
    return 0;
}

int flite_cfg_euler(void* acc, int acc_is_fp32, const void* v_uncond, const void* v_cond, float guidance, float dt,
                    int do_cfg, void* lat_out, int64_t numel, void* stream) {
    if (!acc || !v_cond || !lat_out || (do_cfg && !v_uncond)) return fail(FLITE_ERR_INVALID, "cfg_euler: null pointer");
    if (numel <= 0 || numel % 8) return fail(FLITE_ERR_INVALID, "cfg_euler: numel %lld must be a positive multiple of 8", (long long)numel);
    const long long n8 = numel / 8;
    int blocks = (int)((n8 + 255) / 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cfg_euler_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        acc, acc_is_fp32, (const __nv_bfloat16*)v_uncond, (const __nv_bfloat16*)v_cond, guidance, dt, do_cfg,
        (__nv_bfloat16*)lat_out, n8);
    LAUNCH_CHECK();
    return 0;
}

int flite_apg_workspace_bytes(void) { return 4 * APG_BLOCKS * (int)sizeof(double); }

int flite_apg_euler(void* acc, int acc_is_fp32, const void* v_uncond, const void* v_cond, float guidance, float dt,
                    float orthogonal_threshold, void* lat_out, int64_t numel, void* workspace, void* stream) {
    if (!acc || !v_cond || !v_uncond || !lat_out || !workspace) return fail(FLITE_ERR_INVALID, "apg_euler: null pointer");
    if (numel <= 8 || numel % 8) return fail(FLITE_ERR_INVALID, "apg_euler: numel %lld must be a multiple of 8 greater than 8", (long long)numel);
    if ((uintptr_t)workspace % 8) return fail(FLITE_ERR_INVALID, "apg_euler: workspace must be 8-byte aligned");
    const long long n8 = numel / 8;
    cudaStream_t s = (cudaStream_t)stream;
    const __nv_bfloat16* u = (const __nv_bfloat16*)v_uncond;
    const __nv_bfloat16* c = (const __nv_bfloat16*)v_cond;
    double* ws = (double*)workspace;
    const float gm1 = (float)((double)guidance - 1.0);
    apg_dot_kernel<<<APG_BLOCKS, APG_THREADS, 0, s>>>(u, c, n8, ws);
    LAUNCH_CHECK();
    apg_orth_stats_kernel<<<APG_BLOCKS, APG_THREADS, 0, s>>>(u, c, n8, ws);
    LAUNCH_CHECK();
    apg_euler_kernel<<<APG_BLOCKS, APG_THREADS, 0, s>>>(acc, acc_is_fp32, u, c, gm1, dt, orthogonal_threshold,
                                                         (__nv_bfloat16*)lat_out, n8, ws);
    LAUNCH_CHECK();
    return 0;
}

int flite_latent_unscale(const void* latents, void* out, float scaling_factor, float shift_factor, int64_t numel,
                         void* stream) {
    if (!latents || !out) return fail(FLITE_ERR_INVALID, "latent_unscale: null pointer");
    if (numel <= 0 || numel % 8) return fail(FLITE_ERR_INVALID, "latent_unscale: numel must be a positive multiple of 8");
    if (scaling_factor == 0.0f) return fail(FLITE_ERR_INVALID, "latent_unscale: scaling_factor is zero");
    const long long n8 = numel / 8;
    int blocks = (int)((n8 + 255) / 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    latent_unscale_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)latents, (__nv_bfloat16*)out,
                                                                    1.0f / scaling_factor, shift_factor, n8);
    LAUNCH_CHECK();
    return 0;
}

int64_t flite_groupnorm_partials_bytes(int N, int groups, int splits) {
    return (int64_t)N * groups * splits * (int64_t)sizeof(float2);
}

int flite_groupnorm_silu_nhwc(const void* x, void* y, const void* gamma, const void* beta, int N, int64_t HW, int C,
                              int groups, float eps, int apply_silu, void* partials, int splits, void* stream) {
    if (!x || !y || !gamma || !beta || !partials) return fail(FLITE_ERR_INVALID, "groupnorm: null pointer");
    if (N <= 0 || HW <= 0) return 0;
    if (C % 8 || groups <= 0 || groups > 64 || C % groups || (C / groups) % 4 || C / 8 > GN_THREADS || splits <= 0)
        return fail(FLITE_ERR_INVALID, "groupnorm: needs C %% 8 == 0, (C / groups) %% 4 == 0, groups <= 64, C <= 2048 (C %d, groups %d)",
                    C, groups);
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15)
        return fail(FLITE_ERR_INVALID, "groupnorm: pointers must be 16-byte aligned");
    const int chunks = C / 8, lanes = GN_THREADS / chunks;
    const size_t smem = (size_t)lanes * chunks * 16;
    groupnorm_stats_kernel<<<dim3(splits, N), GN_THREADS, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (long long)HW, C, groups, (float2*)partials);
    LAUNCH_CHECK();
    long long blocks = (HW * chunks + GN_THREADS * 4 - 1) / (GN_THREADS * 4);
    const long long cap = (long long)num_sms() * 8 / (N > 0 ? 1 : 1);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    groupnorm_apply_kernel<<<dim3((unsigned)blocks, N), GN_THREADS, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, (long long)HW, C,
        groups, eps, apply_silu, (const float2*)partials, splits);
    LAUNCH_CHECK();
    return 0;
}

int flite_upsample_nearest2x_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream) {
    if (!x || !y) return fail(FLITE_ERR_INVALID, "upsample: null pointer");
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    if (C % 8 || (((uintptr_t)x | (uintptr_t)y) & 15))
        return fail(FLITE_ERR_INVALID, "upsample: C must be a multiple of 8 and the pointers 16-byte aligned");
    long long blocks = ((long long)N * H * W * (C / 8) + 256 * 2 - 1) / (256 * 2);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    upsample_nearest2x_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x,
                                                                                       (__nv_bfloat16*)y, N, H, W, C);
    LAUNCH_CHECK();
    return 0;
}

int flite_bias_residual_add_nhwc(void* y, const void* bias, const void* residual, int64_t rows, int C, void* stream) {
    if (!y || !bias) return fail(FLITE_ERR_INVALID, "bias_residual_add: null pointer");
    if (rows <= 0) return 0;
    if (C % 8 || (((uintptr_t)y | (uintptr_t)bias | (uintptr_t)residual) & 15))
        return fail(FLITE_ERR_INVALID, "bias_residual_add: C must be a multiple of 8 and the pointers 16-byte aligned");
    long long blocks = (rows * (C / 8) + 256 * 4 - 1) / (256 * 4);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    bias_residual_add_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)y, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, (long long)rows, C);
    LAUNCH_CHECK();
    return 0;
}

int flite_image_to_uint8(const void* decoded, int in_is_fp32, void* out_u8, int B, int C, int H, int W, void* stream) {
    if (!decoded || !out_u8) return fail(FLITE_ERR_INVALID, "image_to_uint8: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const long long hw = (long long)H * W, total = hw * B;
    int blocks = (int)((total + 255) / 256);
    const int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    image_to_uint8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(decoded, in_is_fp32, (uint8_t*)out_u8, C, hw, total);
    LAUNCH_CHECK();
    return 0;
}

int flite_rmsnorm_modulate(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w, int weight_mode,
                           const void* scale, const void* shift, int64_t ld_mod, int rows_per_sample, int rows, int d,
                           float eps, void* stream) {
    if (!x || !y) return fail(FLITE_ERR_INVALID, "rmsnorm: null pointer");
    if (d % 8 || ldx % 8 || ldy % 8) return fail(FLITE_ERR_INVALID, "rmsnorm: d/ld must be multiples of 8");
    if (weight_mode != 0 && !w) return fail(FLITE_ERR_INVALID, "rmsnorm: weight_mode %d needs a weight", weight_mode);
    if ((scale == nullptr) != (shift == nullptr)) return fail(FLITE_ERR_INVALID, "rmsnorm: scale and shift go together");
    if (rows <= 0) return 0;
    if (rows_per_sample <= 0) rows_per_sample = rows;
    auto launch = [&](auto kern, int rows_per_block, int threads) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((rows + rows_per_block - 1) / rows_per_block);
        cfg.blockDim = dim3(threads);
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 0);
        const long long ldx_ = ldx, ldy_ = ldy, ld_mod_ = ld_mod;
        cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)x, ldx_, (__nv_bfloat16*)y, ldy_, (const __nv_bfloat16*)w,
                           weight_mode, (const __nv_bfloat16*)scale, (const __nv_bfloat16*)shift, ld_mod_,
                           rows_per_sample, rows, d, eps);
    };
    const int mode = g_tuning[FLITE_TUNE_RMSNORM_KERNEL];   // 0 auto (two-pass), 1 two-pass, 2 register-resident, 3 streaming
    if (mode == 3 && d <= 8 * 32 * 16) {
        // persistent warps: 1 (wide rows) or 2 blocks x 8 warps per SM; shrink the grid so that every warp gets the same
        // number of rows
        const int slots = num_sms() * (d <= 8 * 32 * 4 ? 2 : 1) * 8;
        const int iters = (rows + slots - 1) / slots;
        const int blocks = (rows + 8 * iters - 1) / (8 * iters);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(256);
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 0);
        const long long ldx_ = ldx, ldy_ = ldy, ld_mod_ = ld_mod;
        auto go = [&](auto kern) {
            cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)x, ldx_, (__nv_bfloat16*)y, ldy_, (const __nv_bfloat16*)w,
                               weight_mode, (const __nv_bfloat16*)scale, (const __nv_bfloat16*)shift, ld_mod_,
                               rows_per_sample, rows, d, eps);
        };
        if (d <= 8 * 32 * 4) go(rmsnorm_modulate_stream_kernel<4>);
        else if (d <= 8 * 32 * 12) go(rmsnorm_modulate_stream_kernel<12>);
        else go(rmsnorm_modulate_stream_kernel<16>);
        LAUNCH_CHECK();
        return 0;
    }
    if (mode == 2 && d <= 8 * 32 * 4) launch(rmsnorm_modulate_reg_kernel<4>, 8, 256);
    else if (mode == 2 && d <= 8 * 32 * 12) launch(rmsnorm_modulate_reg_kernel<12>, 8, 256);
    else if (mode == 2 && d <= 8 * 32 * 16) launch(rmsnorm_modulate_reg_kernel<16>, 8, 256);
    else launch(rmsnorm_modulate_kernel, 4, 128);
    LAUNCH_CHECK();
    return 0;
}

int flite_rope_qknorm(void* buf, int64_t ld, int rows, int n_slots, const void* cos_t, const void* sin_t,
                      int rows_per_sample, float eps, void* stream) {
    if (!buf) return fail(FLITE_ERR_INVALID, "rope_qknorm: null pointer");
    if (ld % 8) return fail(FLITE_ERR_INVALID, "rope_qknorm: ld must be a multiple of 8");
    if ((cos_t == nullptr) != (sin_t == nullptr)) return fail(FLITE_ERR_INVALID, "rope_qknorm: cos and sin go together");
    if (rows <= 0 || n_slots <= 0) return 0;
    if (rows_per_sample <= 0) rows_per_sample = rows;
    const long long warps = (long long)rows * n_slots;
    rope_qknorm_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)buf, ld, rows, n_slots, (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t,
        rows_per_sample, eps);
    LAUNCH_CHECK();
    return 0;
}

int flite_patch_embed(const void* x, const void* w, const void* bias, const void* reg_tokens, void* out, int B, int C,
                      int H, int W, int P, int d, int n_reg, int tok_offset, int tok_count, void* stream) {
    if (!x || !w || !bias || !out || (n_reg > 0 && !reg_tokens)) return fail(FLITE_ERR_INVALID, "patch_embed: null pointer");
    if (H % P || W % P) return fail(FLITE_ERR_INVALID, "patch_embed: H, W must be multiples of the patch size");
    const int kdim = C * P * P;
    const int L_full = n_reg + (H / P) * (W / P);
    if (tok_count <= 0) { tok_offset = 0; tok_count = L_full; }
    if (tok_offset < 0 || tok_offset + tok_count > L_full) return fail(FLITE_ERR_INVALID, "patch_embed: token slice out of range");
    const int rows = B * tok_count;
    const int blocks = (rows + PE_TOK - 1) / PE_TOK;
    auto args = [&](auto kern) {
        kern<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                      (const __nv_bfloat16*)bias, (const __nv_bfloat16*)reg_tokens,
                                                      (__nv_bfloat16*)out, B, C, H, W, P, d, n_reg, tok_offset, tok_count);
    };
    if (kdim == 64) args(patch_embed_kernel<64>);
    else if (kdim == 16) args(patch_embed_kernel<16>);
    else return fail(FLITE_ERR_INVALID, "patch_embed: C*P*P = %d unsupported (16 or 64)", kdim);
    LAUNCH_CHECK();
    return 0;
}

int flite_patch_gather(const void* x, const void* reg_tokens, void* A, void* out, int B, int C, int H, int W, int P,
                       int d, int n_reg, int tok_offset, int tok_count, void* stream) {
    if (!x || !A || !out || (n_reg > 0 && !reg_tokens)) return fail(FLITE_ERR_INVALID, "patch_gather: null pointer");
    if (P <= 0 || H % P || W % P || d % 8) return fail(FLITE_ERR_INVALID, "patch_gather: H, W must be multiples of the patch size, d of 8");
    const int L_full = n_reg + (H / P) * (W / P);
    if (tok_count <= 0) { tok_offset = 0; tok_count = L_full; }
    if (tok_offset < 0 || tok_offset + tok_count > L_full) return fail(FLITE_ERR_INVALID, "patch_gather: token slice out of range");
    if (B <= 0) return 0;
    int n_reg_local = (n_reg < tok_offset + tok_count ? n_reg : tok_offset + tok_count) - tok_offset;
    if (n_reg_local < 0) n_reg_local = 0;
    patch_gather_kernel<<<B * tok_count, 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)reg_tokens, (__nv_bfloat16*)A, (__nv_bfloat16*)out, B, C, H, W, P,
        d, n_reg, tok_offset, tok_count, n_reg_local);
    LAUNCH_CHECK();
    return 0;
}

int flite_timestep_embed(const float* t, int t_is_bf16, const float* freqs, void* out, int B, int d, void* stream) {
    if (!t || !freqs || !out) return fail(FLITE_ERR_INVALID, "timestep_embed: null pointer");
    if (d % 2) return fail(FLITE_ERR_INVALID, "timestep_embed: d must be even");
    const int n = B * (d / 2);
    timestep_embed_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, t_is_bf16, freqs, (__nv_bfloat16*)out, B, d);
    LAUNCH_CHECK();
    return 0;
}

int flite_unpatchify(const void* tok, int64_t ldt, void* out, int B, int C, int H, int W, int P, int n_reg,
                     void* stream) {
    if (!tok || !out) return fail(FLITE_ERR_INVALID, "unpatchify: null pointer");
    const long long total = (long long)B * C * H * W;
    unpatchify_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)tok, ldt, (__nv_bfloat16*)out, B, C, H, W, P, n_reg);
    LAUNCH_CHECK();
    return 0;
}

int flite_permute_021(const void* src, void* dst, int n0, int n1, int n2, void* stream) {
    if (!src || !dst) return fail(FLITE_ERR_INVALID, "permute: null pointer");
    if (n2 % 8) return fail(FLITE_ERR_INVALID, "permute: innermost extent must be a multiple of 8");
    const long long total = (long long)n0 * n1 * (n2 / 8);
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    permute_021_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, n0, n1, n2 / 8);
    LAUNCH_CHECK();
    return 0;
}

int flite_pack_context(const void* src, int64_t lds, void* dst, int64_t ldd, const float* mask, int B, int Lc, int d,
                       int* pos_ws, int* seqlens_ws, int* cu_seqlens, void* stream) {
    if (!src || !dst || !mask || !pos_ws || !seqlens_ws || !cu_seqlens) return fail(FLITE_ERR_INVALID, "pack_context: null pointer");
    if (d % 8 || lds % 8 || ldd % 8) return fail(FLITE_ERR_INVALID, "pack_context: d/ld must be multiples of 8");
    cudaStream_t s = (cudaStream_t)stream;
    if (B <= 0 || Lc <= 0) return 0;
    mask_scan_kernel<<<B, 256, 0, s>>>(mask, Lc, pos_ws, seqlens_ws);
    cu_seqlens_kernel<<<1, 32, 0, s>>>(seqlens_ws, B, cu_seqlens);
    pack_rows_kernel<<<B * Lc, 128, 0, s>>>((const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, pos_ws,
                                           cu_seqlens, B, Lc, d);
    LAUNCH_CHECK();
    return 0;
}

struct SpPeers {
    void* const* ptrs;
    int rank;
    int seq_len;
};

static int gemm_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                     const void* bias, int act, int epilogue, const void* resid, int64_t ldr, const void* gate,
                     int64_t ld_gate, int rows_per_sample, const void* rope_cos, const void* rope_sin, int qk_cols,
                     float eps, int sp_ranks, int sp_heads_per_rank, int variant, void* stream, const SpPeers* peers) {
    if (!A || !W || !C) return fail(FLITE_ERR_INVALID, "gemm: null pointer");
    if (M <= 0) return 0;
    if (K <= 0 || K % 64) return fail(FLITE_ERR_INVALID, "gemm: K = %d must be a positive multiple of 64", K);
    if (N <= 0 || N % 64) return fail(FLITE_ERR_INVALID, "gemm: N = %d must be a positive multiple of 64", N);
    if (lda % 8 || ldw % 8 || ldc % 8) return fail(FLITE_ERR_INVALID, "gemm: lda/ldw/ldc must be multiples of 8");
    if (epilogue < 0 || epilogue > 3) return fail(FLITE_ERR_INVALID, "gemm: unknown epilogue %d", epilogue);
    if (epilogue == EPI_GATED_RES && (!resid || !gate || ldr % 8 || ld_gate % 8))
        return fail(FLITE_ERR_INVALID, "gemm: gated-residual epilogue needs resid and gate (ld %% 8 == 0)");
    if ((epilogue == EPI_SWIGLU || epilogue == EPI_QKV_ROPE) && N % 256)
        return fail(FLITE_ERR_INVALID, "gemm: epilogue %d needs N %% 256 == 0", epilogue);
    if (epilogue == EPI_QKV_ROPE && ((rope_cos == nullptr) != (rope_sin == nullptr) || qk_cols % 256))
        return fail(FLITE_ERR_INVALID, "gemm: bad RoPE arguments");
    if (rows_per_sample <= 0) rows_per_sample = M;
    if (sp_ranks > 0 && (epilogue != EPI_QKV_ROPE || sp_heads_per_rank <= 0 || N % 768 ||
                         (N / 768) != sp_ranks * sp_heads_per_rank || M % rows_per_sample))
        return fail(FLITE_ERR_INVALID, "gemm: sequence-parallel head scatter needs the QKV epilogue and N = 3*256*ranks*heads_per_rank");

    if (variant == FLITE_GEMM_AUTO && g_tuning[FLITE_TUNE_GEMM_VARIANT]) variant = g_tuning[FLITE_TUNE_GEMM_VARIANT];
    // skinny GEMM (timestep / modulation path): weight streaming on all SMs instead of N/128 tensor-core tiles
    if ((variant == FLITE_GEMM_AUTO && M <= 8 && epilogue == EPI_STORE && K % 8 == 0) || variant == FLITE_GEMM_GEMV) {
        if (M > 8 || epilogue != EPI_STORE)
            return fail(FLITE_ERR_INVALID, "gemm: the GEMV variant handles M <= 8 with the plain epilogue");
        const int rows = GEMV_ROWS(M <= 2 ? 2 : M <= 4 ? 4 : 8);
        int blocks = (N + rows - 1) / rows;             // output columns per block iteration
        const int cap = num_sms() * 8;
        if (blocks > cap) blocks = (blocks + ((blocks + cap - 1) / cap) - 1) / ((blocks + cap - 1) / cap);   // equal shares
        const __nv_bfloat16* a_ = (const __nv_bfloat16*)A; const __nv_bfloat16* w_ = (const __nv_bfloat16*)W;
        __nv_bfloat16* c_ = (__nv_bfloat16*)C; const __nv_bfloat16* b_ = (const __nv_bfloat16*)bias;
        if (M <= 2) gemv_small_m_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(a_, lda, w_, ldw, c_, ldc, b_, act, M, N, K);
        else if (M <= 4) gemv_small_m_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(a_, lda, w_, ldw, c_, ldc, b_, act, M, N, K);
        else gemv_small_m_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(a_, lda, w_, ldw, c_, ldc, b_, act, M, N, K);
        LAUNCH_CHECK();
        return 0;
    }
    if (variant == FLITE_GEMM_AUTO) {
        if (epilogue == EPI_QKV_ROPE) variant = (M > 128) ? FLITE_GEMM_2CTA_N256 : FLITE_GEMM_1CTA_N256;
        else if (N % 256 == 0 && M > 128) variant = FLITE_GEMM_2CTA_N256;
        else if (N % 128 == 0) variant = FLITE_GEMM_1CTA_N128;
        else variant = FLITE_GEMM_1CTA_N64;
    }
    int block_n = 0, cg = 1;
    switch (variant) {
        case FLITE_GEMM_1CTA_N256: block_n = 256; break;
        case FLITE_GEMM_2CTA_N256: block_n = 256; cg = 2; break;
        case FLITE_GEMM_1CTA_N128: block_n = 128; break;
        case FLITE_GEMM_1CTA_N64: block_n = 64; break;
        default: return fail(FLITE_ERR_INVALID, "gemm: unknown variant %d", variant);
    }
    if (N % block_n) return fail(FLITE_ERR_INVALID, "gemm: N = %d not a multiple of the N-tile %d", N, block_n);

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.C = (__nv_bfloat16*)C; p.ldc = ldc;
    p.bias = (const __nv_bfloat16*)bias; p.act = act;
    p.resid = (const __nv_bfloat16*)resid; p.ldr = ldr;
    p.gate = (const __nv_bfloat16*)gate; p.ld_gate = ld_gate;
    p.rows_per_sample = rows_per_sample;
    p.rope_cos = (const __nv_bfloat16*)rope_cos; p.rope_sin = (const __nv_bfloat16*)rope_sin; p.qk_cols = qk_cols; p.eps = eps;
    p.sp_ranks = sp_ranks; p.sp_hp = sp_heads_per_rank; p.n_heads = N / 768;
    if (peers) {
        if (sp_ranks <= 0 || sp_ranks > 8) return fail(FLITE_ERR_INVALID, "gemm: peer scatter needs 1..8 sequence-parallel ranks");
        for (int i = 0; i < sp_ranks; ++i) {
            if (!peers->ptrs[i]) return fail(FLITE_ERR_INVALID, "gemm: null peer buffer %d", i);
            p.sp_peer[i] = (__nv_bfloat16*)peers->ptrs[i];
        }
        p.sp_rank = peers->rank;
        p.sp_seq = peers->seq_len;
    }
    p.stage_stores = (epilogue == EPI_QKV_ROPE && (peers != nullptr || g_tuning[FLITE_TUNE_QKV_STAGED_STORES])) ? 1 : 0;

    CUtensorMap ta, tb, tbh, tah;
    int rc = make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128);
    if (rc) return rc;
    rc = make_tmap(&tah, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64);     // 64-row boxes of the narrow last M-tile
    if (rc) return rc;
    rc = make_tmap(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(block_n / cg));
    if (rc) return rc;
    rc = make_tmap(&tbh, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(block_n / cg / 2));   // half-width tail tiles
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (variant) {
        case FLITE_GEMM_1CTA_N256: return dispatch_epi<1, 256, 4>(epilogue, ta, tb, tbh, tah, p, s);
        case FLITE_GEMM_2CTA_N256: return dispatch_epi<2, 256, 6>(epilogue, ta, tb, tbh, tah, p, s);
        case FLITE_GEMM_1CTA_N128: return dispatch_epi<1, 128, 6>(epilogue, ta, tb, tbh, tah, p, s);
        case FLITE_GEMM_1CTA_N64: return dispatch_epi<1, 64, 8>(epilogue, ta, tb, tbh, tah, p, s);
    }
    return fail(FLITE_ERR_INVALID, "gemm: unreachable");
}

int flite_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K,
                    const void* bias, int act, int epilogue, const void* resid, int64_t ldr, const void* gate,
                    int64_t ld_gate, int rows_per_sample, const void* rope_cos, const void* rope_sin, int qk_cols,
                    float eps, int sp_ranks, int sp_heads_per_rank, int variant, void* stream) {
    return gemm_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, act, epilogue, resid, ldr, gate, ld_gate, rows_per_sample,
                     rope_cos, rope_sin, qk_cols, eps, sp_ranks, sp_heads_per_rank, variant, stream, nullptr);
}

int flite_gemm_qkv_p2p(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, const void* bias,
                       int tokens_per_sample, const void* rope_cos, const void* rope_sin, float eps, int sp_ranks,
                       int sp_heads_per_rank, int sp_rank, int seq_len, void* const* peer_recv, int variant,
                       void* stream) {
    if (!peer_recv) return fail(FLITE_ERR_INVALID, "gemm_qkv_p2p: null peer table");
    if (sp_rank < 0 || sp_rank >= sp_ranks || seq_len != tokens_per_sample * sp_ranks)
        return fail(FLITE_ERR_INVALID, "gemm_qkv_p2p: inconsistent sequence-parallel geometry");
    const int d = sp_ranks * sp_heads_per_rank * 256;
    SpPeers peers{peer_recv, sp_rank, seq_len};
    const int64_t ldc = 3 * (int64_t)sp_heads_per_rank * 256;
    return gemm_impl(A, lda, W, ldw, peer_recv[sp_rank], ldc, M, 3 * d, K, bias, 0, EPI_QKV_ROPE, nullptr, 0, nullptr, 0,
                     tokens_per_sample, rope_cos, rope_sin, 2 * d, eps, sp_ranks, sp_heads_per_rank, variant, stream,
                     &peers);
}

struct AttnPeers {
    void* const* ptrs;
    int n, lq, head0;
};

static int sk_clusters();

static int attention_impl(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                          int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out, int64_t ldo,
                          const int* cu_q, const int* cu_k, int B, int H, int max_q, float softmax_scale,
                          int variant, void* stream, const AttnPeers* peers) {
    if (!q || !k || !v || !out || !cu_q || !cu_k) return fail(FLITE_ERR_INVALID, "attention: null pointer");
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8)
        return fail(FLITE_ERR_INVALID, "attention: strides / column offsets must be multiples of 8");
    if (B <= 0 || H <= 0 || max_q <= 0 || rows_q <= 0) return 0;
    // Declared width of the q / k / v tensor maps = exactly the columns the kernels address (col0 + H heads of 256), NOT
    // the row stride: q, k, v are usually column-offset views of a wider projection buffer, and a map that claims `ld`
    // columns from the view's first element would extend past the end of the allocation in its last row.
    const int64_t q_cols = (int64_t)q_col0 + 256ll * H, k_cols = (int64_t)k_col0 + 256ll * H, v_cols = (int64_t)v_col0 + 256ll * H;
    if (q_cols > ldq || k_cols > ldk || v_cols > ldv)
        return fail(FLITE_ERR_INVALID, "attention: col0 + 256*H exceeds the row stride");
    static bool configured_dev[64] = {false};
    bool& configured = configured_dev[cur_dev()];
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_cg2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_cg2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_cg2_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_cg2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_qtmem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AQ_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(attn_fwd_qtmem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AQ_SMEM));
        configured = true;
    }
    if (variant < FLITE_ATTN_AUTO || variant > FLITE_ATTN_PERSISTENT)
        return fail(FLITE_ERR_INVALID, "attention: unknown variant %d", variant);
    // short key sequences (cross-attention over the text context: <= 512 keys per sequence on average) may use their own variant
    if (variant == FLITE_ATTN_AUTO && !peers && rows_k <= 512ll * B && g_tuning[FLITE_TUNE_ATTN_VARIANT_SHORT_K] > 0 &&
        g_tuning[FLITE_TUNE_ATTN_VARIANT_SHORT_K] != FLITE_ATTN_XRES)   // XRES needs the per-sequence bound the host knows
        variant = g_tuning[FLITE_TUNE_ATTN_VARIANT_SHORT_K];
    // default: the persistent kernel (whole units round-robin over one wave of clusters; bit-identical to variant 5 and
    // faster at every measured shape: cross-attention 72 -> 48 us at C2, 511 -> 298 us at C3, self-attention 339 -> 323 us);
    // peer-memory output goes through the per-unit kernel
    if (variant == FLITE_ATTN_AUTO)
        variant = g_tuning[FLITE_TUNE_ATTN_VARIANT] ? g_tuning[FLITE_TUNE_ATTN_VARIANT]
                                                    : (peers ? FLITE_ATTN_2CTA_1WG_PTMEM : FLITE_ATTN_PERSISTENT);
    if (variant == FLITE_ATTN_PERSISTENT && peers) variant = FLITE_ATTN_2CTA_1WG_PTMEM;   // peer stores: the per-unit kernel
    if (variant == FLITE_ATTN_PERSISTENT) {
        // One wave of clusters, whole (sequence, head, 256-query tile) units handed out round-robin, per-sequence lengths
        // from cu_q / cu_k (attn_sk_kernel, ragged mode): same arithmetic per unit as variant 5, no per-unit launch cost.
        CUtensorMap tq, tk, tv, to;
        int rc = make_tmap(&tq, q, (uint64_t)rows_q, (uint64_t)q_cols, (uint64_t)ldq, 128);
        if (rc) return rc;
        rc = make_tmap(&tk, k, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)k_cols, (uint64_t)ldk, 64);
        if (rc) return rc;
        rc = make_tmap(&tv, v, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)v_cols, (uint64_t)ldv, 128);
        if (rc) return rc;
        AttnSkParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.cu_q = cu_q; sp.cu_k = cu_k;
        sp.out = (__nv_bfloat16*)out; sp.ldo = ldo;
        sp.q_col0 = q_col0; sp.k_col0 = k_col0; sp.v_col0 = v_col0;
        sp.scale_log2 = softmax_scale * 1.4426950408889634f;
        sp.B = B; sp.H = H;
        sp.QT = (max_q + 255) / 256; sp.NT = 1;
        const long long units = (long long)B * H * sp.QT;
        sp.total = units; sp.rr_units = units; sp.ragged = 1;
        sp.debug = g_tuning[FLITE_TUNE_ATTN_DEBUG];
        sp.sp_lq = 1;
        to = tq;
        sp.tma_out = (g_tuning[FLITE_TUNE_ATTN_TMA_OUT] == 0 && ((uintptr_t)out & 15) == 0) ? 1 : 0;
        if (sp.tma_out) {
            rc = make_tmap(&to, out, (uint64_t)rows_q, (uint64_t)(256ll * H), (uint64_t)ldo, 128);
            if (rc) return rc;
        }
        long long clusters = sk_clusters();
        if (clusters > units) clusters = units;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        cfg.blockDim = dim3(SK_THREADS);
        cfg.dynamicSmemBytes = ATT_SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 2);
        CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_sk_kernel, tq, tk, tv, to, sp));
        return 0;
    }
    const bool cg2 = variant >= FLITE_ATTN_2CTA_1WG && variant <= FLITE_ATTN_2CTA_2WG_PTMEM;
    const bool qtmem = variant == FLITE_ATTN_QTMEM_1WG || variant == FLITE_ATTN_QTMEM_2WG;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.cu_q = cu_q; p.cu_k = cu_k;
    p.out = (__nv_bfloat16*)out; p.ldo = ldo;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.debug = g_tuning[FLITE_TUNE_ATTN_DEBUG];
    for (int i = 0; i < 8; ++i) p.out_peer[i] = nullptr;
    p.sp_lq = 1; p.sp_head0 = 0;
    if (peers) {
        if (!cg2)
            return fail(FLITE_ERR_INVALID, "attention: the peer-memory output path needs a 2-CTA variant (3..6)");
        if (peers->n <= 0 || peers->n > 8 || peers->lq <= 0) return fail(FLITE_ERR_INVALID, "attention: bad peer table");
        for (int i = 0; i < peers->n; ++i) {
            if (!peers->ptrs[i]) return fail(FLITE_ERR_INVALID, "attention: null peer buffer %d", i);
            p.out_peer[i] = (__nv_bfloat16*)peers->ptrs[i];
        }
        p.sp_lq = peers->lq; p.sp_head0 = peers->head0;
    }
    p.stage_out = (peers != nullptr || g_tuning[FLITE_TUNE_ATTN_STAGED_STORES]) ? 1 : 0;
    p.tma_out = (cg2 && !p.stage_out && g_tuning[FLITE_TUNE_ATTN_TMA_OUT] == 0) ? 1 : 0;
    const int q_tiles = (max_q + 127) / 128;
    if (variant == FLITE_ATTN_XRES) {
        // persistent cross-attention with resident K/V: every sequence must have <= 256 keys (checked in the kernel)
        if (peers) return fail(FLITE_ERR_INVALID, "attention: the resident-K/V variant has no peer-memory output path");
        static bool xres_configured_dev[64] = {false};
        bool& xres_configured = xres_configured_dev[cur_dev()];
        if (!xres_configured) {
            CUDA_TRY(cudaFuncSetAttribute(attn_xres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XR_SMEM));
            xres_configured = true;
        }
        CUtensorMap xq, xk, xv;
        int rcx = make_tmap(&xq, q, (uint64_t)rows_q, (uint64_t)q_cols, (uint64_t)ldq, 128);
        if (rcx) return rcx;
        rcx = make_tmap(&xk, k, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)k_cols, (uint64_t)ldk, 128);
        if (rcx) return rcx;
        rcx = make_tmap(&xv, v, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)v_cols, (uint64_t)ldv, 128);
        if (rcx) return rcx;
        XresParams xp;
        xp.cu_q = cu_q; xp.cu_k = cu_k; xp.out = (__nv_bfloat16*)out; xp.ldo = ldo;
        xp.q_col0 = q_col0; xp.k_col0 = k_col0; xp.v_col0 = v_col0;
        xp.scale_log2 = softmax_scale * 1.4426950408889634f;
        xp.B = B; xp.H = H; xp.q_pairs = (max_q + 255) / 256;
        const long long n_units = (long long)B * H * xp.q_pairs;
        long long clusters = num_sms() / 2;
        if (clusters > n_units) clusters = n_units;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        cfg.blockDim = dim3(XR_THREADS);
        cfg.dynamicSmemBytes = XR_SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 2);
        CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_xres_kernel, xq, xk, xv, xp));
        return 0;
    }
    if (qtmem) {
        CUtensorMap tk2, tv2;
        int rc2 = make_tmap(&tk2, k, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)k_cols, (uint64_t)ldk, 32);
        if (rc2) return rc2;
        rc2 = make_tmap(&tv2, v, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)v_cols, (uint64_t)ldv, 64);
        if (rc2) return rc2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * ((q_tiles + 1) / 2), H, B);
        cfg.blockDim = dim3(variant == FLITE_ATTN_QTMEM_1WG ? 192 : 320);
        cfg.dynamicSmemBytes = AQ_SMEM;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const __nv_bfloat16* qp = (const __nv_bfloat16*)q;
        long long ldq_ = ldq;
        if (variant == FLITE_ATTN_QTMEM_1WG) CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_qtmem_kernel<1>, tk2, tv2, qp, ldq_, p));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_qtmem_kernel<2>, tk2, tv2, qp, ldq_, p));
        return 0;
    }
    CUtensorMap tq, tk, tv;
    int rc = make_tmap(&tq, q, (uint64_t)rows_q, (uint64_t)q_cols, (uint64_t)ldq, 128);
    if (rc) return rc;
    rc = make_tmap(&tk, k, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)k_cols, (uint64_t)ldk, cg2 ? 64 : 128);
    if (rc) return rc;
    rc = make_tmap(&tv, v, (uint64_t)(rows_k > 0 ? rows_k : 1), (uint64_t)v_cols, (uint64_t)ldv, 128);
    if (rc) return rc;
    if (!cg2) {
        dim3 grid(q_tiles, H, B);
        if (variant == FLITE_ATTN_1WG) attn_fwd_kernel<1><<<grid, 192, ATT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
        else attn_fwd_kernel<2><<<grid, 320, ATT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
        LAUNCH_CHECK();
        return 0;
    }
    // output tile stores through TMA (whole 128-row tiles only; the map covers exactly the H heads the kernel writes)
    CUtensorMap to = tq;
    if (p.tma_out) {
        if (((uintptr_t)out & 15) != 0) p.tma_out = 0;
        else {
            rc = make_tmap(&to, out, (uint64_t)rows_q, (uint64_t)(256ll * H), (uint64_t)ldo, 128);
            if (rc) return rc;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ((q_tiles + 1) / 2), H, B);
    const bool one_wg = variant == FLITE_ATTN_2CTA_1WG || variant == FLITE_ATTN_2CTA_1WG_PTMEM;
    cfg.blockDim = dim3(one_wg ? 192 : 320);
    cfg.dynamicSmemBytes = ATT_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, 2);
    switch (variant) {
        case FLITE_ATTN_2CTA_1WG: CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_cg2_kernel<1, false>, tq, tk, tv, to, p)); break;
        case FLITE_ATTN_2CTA_2WG: CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_cg2_kernel<2, false>, tq, tk, tv, to, p)); break;
        case FLITE_ATTN_2CTA_1WG_PTMEM: CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_cg2_kernel<1, true>, tq, tk, tv, to, p)); break;
        default: CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_fwd_cg2_kernel<2, true>, tq, tk, tv, to, p)); break;
    }
    return 0;
}

int flite_attention_varlen(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                           int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out, int64_t ldo,
                           const int* cu_q, const int* cu_k, int B, int H, int max_q, float softmax_scale,
                           int variant, void* stream) {
    return attention_impl(q, ldq, rows_q, q_col0, k, ldk, rows_k, k_col0, v, ldv, v_col0, out, ldo, cu_q, cu_k, B, H,
                          max_q, softmax_scale, variant, stream, nullptr);
}

// Persistent stream-K self-attention for uniform sequence lengths (attn_sk_sm100.cuh).  The workspace holds the flags and
// one partial-result slot per cluster; it must be zero-filled once when it is allocated (flags are reset by their reader).
static int sk_clusters() {
    // one wave of co-resident 2-CTA clusters with the kernel's real footprint (the GPC layout can strand an SM)
    static int n_dev[64] = {0};
    int& n = n_dev[cur_dev()];
    if (n == 0) {
        cudaFuncSetAttribute(attn_sk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / 2 * 2);
        cfg.blockDim = dim3(SK_THREADS);
        cfg.dynamicSmemBytes = ATT_SMEM;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int active = 0;
        if (cudaOccupancyMaxActiveClusters(&active, attn_sk_kernel, &cfg) != cudaSuccess || active <= 0) {
            cudaGetLastError();
            active = num_sms() / 2;
        }
        n = active < num_sms() / 2 ? active : num_sms() / 2;
        if (n > SK_MAX_CLUSTERS) n = SK_MAX_CLUSTERS;
    }
    return n;
}

int64_t flite_attention_streamk_workspace_bytes(void) {
    return (int64_t)SK_FLAG_BYTES + (int64_t)(num_sms() / 2 + 1) * SK_SLOT_FLOATS * (int64_t)sizeof(float);
}

struct AttnPeers;
static int attention_streamk_impl(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                                  int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out, int64_t ldo,
                                  const int* cu_q, const int* cu_k, int B, int H, int q_len, int k_len, float softmax_scale,
                                  void* workspace, int64_t workspace_bytes, void* stream, void* const* peer_out,
                                  int n_peers, int tokens_per_rank, int head0) {
    if (!q || !k || !v || !out || !cu_q || !cu_k || !workspace) return fail(FLITE_ERR_INVALID, "attention_streamk: null pointer");
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8)
        return fail(FLITE_ERR_INVALID, "attention_streamk: strides / column offsets must be multiples of 8");
    if (B <= 0 || H <= 0 || q_len <= 0 || k_len <= 0)
        return fail(FLITE_ERR_INVALID, "attention_streamk: needs B, H, q_len, k_len > 0 (uniform sequence lengths)");
    if (rows_q < (int64_t)B * q_len || rows_k < (int64_t)B * k_len)
        return fail(FLITE_ERR_INVALID, "attention_streamk: rows_q / rows_k smaller than B * length");
    if (workspace_bytes < flite_attention_streamk_workspace_bytes() || ((uintptr_t)workspace & 15))
        return fail(FLITE_ERR_INVALID, "attention_streamk: workspace too small or misaligned (flite_attention_streamk_workspace_bytes)");
    const int64_t q_cols = (int64_t)q_col0 + 256ll * H, k_cols = (int64_t)k_col0 + 256ll * H, v_cols = (int64_t)v_col0 + 256ll * H;
    if (q_cols > ldq || k_cols > ldk || v_cols > ldv)
        return fail(FLITE_ERR_INVALID, "attention_streamk: col0 + 256*H exceeds the row stride");
    CUtensorMap tq, tk, tv;
    int rc = make_tmap(&tq, q, (uint64_t)rows_q, (uint64_t)q_cols, (uint64_t)ldq, 128);
    if (rc) return rc;
    rc = make_tmap(&tk, k, (uint64_t)rows_k, (uint64_t)k_cols, (uint64_t)ldk, 64);
    if (rc) return rc;
    rc = make_tmap(&tv, v, (uint64_t)rows_k, (uint64_t)v_cols, (uint64_t)ldv, 128);
    if (rc) return rc;
    AttnSkParams p;
    memset(&p, 0, sizeof(p));
    p.cu_q = cu_q; p.cu_k = cu_k;
    p.out = (__nv_bfloat16*)out; p.ldo = ldo;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.B = B; p.H = H; p.q_len = q_len; p.k_len = k_len;
    p.QT = (q_len + 255) / 256; p.NT = (k_len + 127) / 128;
    const long long units = (long long)B * H * p.QT;
    p.total = units * p.NT;
    p.flags = (unsigned int*)workspace;
    p.slots = (float*)((char*)workspace + SK_FLAG_BYTES);
    for (int i = 0; i < 8; ++i) p.out_peer[i] = nullptr;
    p.sp_lq = 1; p.sp_head0 = 0;
    if (peer_out) {
        if (n_peers <= 0 || n_peers > 8 || tokens_per_rank <= 0) return fail(FLITE_ERR_INVALID, "attention_streamk: bad peer table");
        for (int i = 0; i < n_peers; ++i) {
            if (!peer_out[i]) return fail(FLITE_ERR_INVALID, "attention_streamk: null peer buffer %d", i);
            p.out_peer[i] = (__nv_bfloat16*)peer_out[i];
        }
        p.sp_lq = tokens_per_rank; p.sp_head0 = head0;
    }
    long long clusters = sk_clusters();
    if (clusters > units) clusters = units;     // a share is never shorter than one unit => a unit has at most two parts
    // schedule (FLITE_TUNE_ATTN_SK_MODE): 0 pure stream-K | 1 whole units round-robin (never splits: bit-identical to the
    // one-cluster-per-unit launch) | 2 hybrid: whole rounds in lock step, stream-K shares of 1..2 units over the rest
    p.rr_units = 0;
    p.debug = g_tuning[FLITE_TUNE_ATTN_DEBUG];
    CUtensorMap to = tq;
    p.tma_out = (!peer_out && g_tuning[FLITE_TUNE_ATTN_TMA_OUT] == 0 && ((uintptr_t)out & 15) == 0) ? 1 : 0;
    if (p.tma_out) {
        rc = make_tmap(&to, out, (uint64_t)rows_q, (uint64_t)(256ll * H), (uint64_t)ldo, 128);
        if (rc) return rc;
    }
    if (g_tuning[FLITE_TUNE_ATTN_SK_MODE] == 1) p.rr_units = units;
    else if (g_tuning[FLITE_TUNE_ATTN_SK_MODE] == 2) {
        const long long rounds = units / clusters;
        p.rr_units = (units % clusters == 0) ? units : (rounds >= 1 ? (rounds - 1) * clusters : 0);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = ATT_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, 2);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, attn_sk_kernel, tq, tk, tv, to, p));
    return 0;
}

int flite_attention_streamk(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                            int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0, void* out, int64_t ldo,
                            const int* cu_q, const int* cu_k, int B, int H, int q_len, int k_len, float softmax_scale,
                            void* workspace, int64_t workspace_bytes, void* stream) {
    return attention_streamk_impl(q, ldq, rows_q, q_col0, k, ldk, rows_k, k_col0, v, ldv, v_col0, out, ldo, cu_q, cu_k, B, H,
                                  q_len, k_len, softmax_scale, workspace, workspace_bytes, stream, nullptr, 0, 0, 0);
}

int flite_attention_streamk_p2p(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                                int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0,
                                void* const* peer_out, int n_peers, int tokens_per_rank, int head0, int64_t ldo,
                                const int* cu_q, const int* cu_k, int B, int H, int q_len, int k_len, float softmax_scale,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    if (!peer_out) return fail(FLITE_ERR_INVALID, "attention_streamk_p2p: null peer table");
    return attention_streamk_impl(q, ldq, rows_q, q_col0, k, ldk, rows_k, k_col0, v, ldv, v_col0, peer_out[0], ldo, cu_q, cu_k,
                                  B, H, q_len, k_len, softmax_scale, workspace, workspace_bytes, stream, peer_out, n_peers,
                                  tokens_per_rank, head0);
}

int flite_attention_varlen_p2p(const void* q, int64_t ldq, int64_t rows_q, int q_col0, const void* k, int64_t ldk,
                               int64_t rows_k, int k_col0, const void* v, int64_t ldv, int v_col0,
                               void* const* peer_out, int n_peers, int tokens_per_rank, int head0, int64_t ldo,
                               const int* cu_q, const int* cu_k, int B, int H, int max_q, float softmax_scale,
                               int variant, void* stream) {
    if (!peer_out) return fail(FLITE_ERR_INVALID, "attention_p2p: null peer table");
    AttnPeers peers{peer_out, n_peers, tokens_per_rank, head0};
    return attention_impl(q, ldq, rows_q, q_col0, k, ldk, rows_k, k_col0, v, ldv, v_col0, peer_out[0], ldo, cu_q, cu_k,
                          B, H, max_q, softmax_scale, variant, stream, &peers);
}

// ---------------------------------------------------------------------------------------------------------------
// Peer (NVLink) memory plumbing for the fused exchange: symmetric allocations shared between the ranks of one node
// through CUDA IPC handles, and stream-ordered completion flags.
// ---------------------------------------------------------------------------------------------------------------
int flite_p2p_alloc(int64_t bytes, void** out) {
    if (!out || bytes <= 0) return fail(FLITE_ERR_INVALID, "p2p_alloc: bad arguments");
    CUDA_TRY(cudaMalloc(out, (size_t)bytes));
    CUDA_TRY(cudaMemset(*out, 0, (size_t)bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
int flite_p2p_free(void* p) {
    if (p) CUDA_TRY(cudaFree(p));
    return 0;
}
int flite_ipc_get_handle(const void* p, void* handle64) {
    if (!p || !handle64) return fail(FLITE_ERR_INVALID, "ipc_get_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    CUDA_TRY(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(p)));
    return 0;
}
int flite_ipc_open(const void* handle64, void** out) {
    if (!handle64 || !out) return fail(FLITE_ERR_INVALID, "ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int flite_ipc_close(void* p) {
    if (p) CUDA_TRY(cudaIpcCloseMemHandle(p));
    return 0;
}
int flite_p2p_signal(void* const* peer_flags, int n, int my_slot, unsigned int value, void* stream) {
    if (!peer_flags || n <= 0 || n > 8 || my_slot < 0 || my_slot >= 8) return fail(FLITE_ERR_INVALID, "p2p_signal: bad arguments");
    PeerFlagPtrs pf;
    for (int i = 0; i < 8; ++i) pf.p[i] = i < n ? (unsigned int*)peer_flags[i] : nullptr;
    p2p_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, n, my_slot, value);
    LAUNCH_CHECK();
    return 0;
}
int flite_p2p_wait(const void* my_flags, int n, unsigned int value, void* stream) {
    if (!my_flags || n <= 0 || n > 8) return fail(FLITE_ERR_INVALID, "p2p_wait: bad arguments");
    const int secs = g_tuning[FLITE_TUNE_P2P_TIMEOUT_S] > 0 ? g_tuning[FLITE_TUNE_P2P_TIMEOUT_S] : 120;
    p2p_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned int*)my_flags, n, value,
                                                        (unsigned long long)secs * 1000000000ull);
    LAUNCH_CHECK();
    return 0;
}
int flite_poison_on_abort(void* buf, int64_t numel, void* stream) {
    if (!buf || numel <= 0 || numel % 8 || ((uintptr_t)buf & 15))
        return fail(FLITE_ERR_INVALID, "poison_on_abort: need a 16-byte aligned bf16 buffer with numel %% 8 == 0");
    const long long n16 = numel / 8;
    int blocks = (int)((n16 + 255) / 256);
    if (blocks > num_sms() * 4) blocks = num_sms() * 4;
    poison_on_abort_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((uint4*)buf, n16);
    LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
