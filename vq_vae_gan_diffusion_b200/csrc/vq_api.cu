// vq_api.cu -- C-ABI entry points of libvq_b200.so (declared in include/vq_b200.h).
//
// Host side only: argument checks, workspace carving and kernel launches.  Nothing here
// allocates device memory or synchronises; every launch goes to the caller's stream.  There is no CPU fallback:
// on anything but an sm_100 device the compute entry points return VQ_E_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

#include "../../include/vq_b200.h"
#include "vq_allreduce.cuh"
#include "vq_argmin_sm100.cuh"
#include "vq_backward.cuh"
#include "vq_common.cuh"
#include "vq_prep.cuh"
#include "vq_qconv.cuh"
#include "vq_select.cuh"
#include "vq_rows.cuh"
#include "vq_tokens.cuh"

#define VQ_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define VQ_CUDA(expr)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) return fail((int)e_, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define VQ_LAUNCH_CHECK(name)                                                                    \
    do {                                                                                         \
        cudaError_t e_ = cudaGetLastError();                                                     \
        if (e_ != cudaSuccess) return fail((int)e_, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
        g_launches++;                                                                            \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// The production GEMM runs as clusters of two CTAs sharing the codebook stream (vq_argmin_sm100.cuh, kShare);
// VQ_GEMM_SHARE=0 selects independent CTAs (A/B runs).
bool gemm_share() {
    static const bool v = [] {
        const char* e = getenv("VQ_GEMM_SHARE");
        return e == nullptr || atoi(e) != 0;
    }();
    return v;
}

// Programmatic dependent launch of the kernels inside one call (vq_common.cuh: pdl_wait / pdl_trigger); VQ_PDL=0 launches
// them as ordinary stream-ordered kernels (A/B runs).
bool use_pdl() {
    static const bool v = [] {
        const char* e = getenv("VQ_PDL");
        return e == nullptr || atoi(e) != 0;
    }();
    return v;
}

template <typename Kernel, typename... Args>
cudaError_t launch_chained(Kernel kernel, unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// launch of one GEMM variant: as clusters of two CTAs (one cluster per pair of row tiles, at most one per two SMs) when
// `share`, else one CTA per row tile up to one per SM
template <bool kDebug, bool kTimeline, int kNC = vq::kNumDChunks>
cudaError_t launch_gemm(const vq::GemmParams& gp, bool share, int sms, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(vq::kGemmThreads);
    cfg.dynamicSmemBytes = vq::gemm_smem_bytes<kNC>();
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (use_pdl()) {                                          // chained behind vq_prep_z_kernel
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        n_attr++;
    }
    cfg.attrs = attr;
    if (share) {
        const int pairs = (gp.row_tiles + 1) / 2;
        cfg.gridDim = dim3(2 * (pairs < sms / 2 ? pairs : sms / 2));
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = 2;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        n_attr++;
        cfg.numAttrs = n_attr;
        return cudaLaunchKernelEx(&cfg, vq::vq_argmin_gemm_kernel<kDebug, kTimeline, true, kNC>, gp);
    }
    cfg.numAttrs = n_attr;
    cfg.gridDim = dim3(gp.row_tiles < sms ? gp.row_tiles : sms);
    return cudaLaunchKernelEx(&cfg, vq::vq_argmin_gemm_kernel<kDebug, kTimeline, false, kNC>, gp);
}

// ---- optional timing of the distance-GEMM kernel (bench.py's roofline): a ring of event pairs recorded on the
// caller's stream around that one launch; read back after the caller has synchronised.
constexpr int kProfCap = 512;
struct Prof {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[kProfCap][2];
    int created = 0;
};
thread_local Prof g_prof;
thread_local long long* g_timeline = nullptr;   // debug: device buffer for per-tile clock stamps (vq_debug_timeline)
thread_local int g_timeline_tiles = 0;

// ---- per-device info (SM count, capability), cached
struct DevInfo { int sms = 0; int cc = 0; bool attrs_set = false; };
DevInfo g_dev[64];
std::mutex g_mu;

int device_info(DevInfo** out) {
    int dev = 0;
    VQ_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(VQ_E_DEVICE, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevInfo& d = g_dev[dev];
    if (d.sms == 0) {
        int major = 0, minor = 0, sms = 0;
        VQ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        VQ_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
        VQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        d.cc = major * 10 + minor;
        d.sms = sms;
    }
    if (d.cc != 100)
        return fail(VQ_E_DEVICE, "vq_b200 kernels are built for sm_100a only; current device is sm_%d (no fallback)", d.cc);
    if (!d.attrs_set) {
        const void* gemm_variants[] = {(const void*)vq::vq_argmin_gemm_kernel<false, false, false>,
                                       (const void*)vq::vq_argmin_gemm_kernel<false, false, true>,
                                       (const void*)vq::vq_argmin_gemm_kernel<true, false, false>,
                                       (const void*)vq::vq_argmin_gemm_kernel<false, true, false>,
                                       (const void*)vq::vq_argmin_gemm_kernel<false, true, true>};
        for (const void* fn : gemm_variants)
            VQ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vq::kGemmSmemBytes));
        // the other contraction widths (row-major nearest-code searches)
        const std::pair<const void*, size_t> rows_variants[] = {
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, false, 1>, vq::gemm_smem_bytes<1>()},
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, true, 1>, vq::gemm_smem_bytes<1>()},
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, false, 2>, vq::gemm_smem_bytes<2>()},
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, true, 2>, vq::gemm_smem_bytes<2>()},
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, false, 8>, vq::gemm_smem_bytes<8>()},
            {(const void*)vq::vq_argmin_gemm_kernel<false, false, true, 8>, vq::gemm_smem_bytes<8>()}};
        for (const auto& v : rows_variants)
            VQ_CUDA(cudaFuncSetAttribute(v.first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.second));
        VQ_CUDA(cudaFuncSetAttribute((const void*)vq::vq_qconv_prep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)vq::kQcSmemBytes));
        VQ_CUDA(cudaFuncSetAttribute((const void*)vq::vq_qconv_prep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)vq::kQcSmemBytes));
        d.attrs_set = true;
    }
    *out = &d;
    return VQ_OK;
}

// ---- workspace layout
struct Workspace {
    __half* z_h;            // (N_pad, D)
    float* z2;              // (N)
    float* z_inv_scale;     // (N)
    float* denom;           // (N) L2-normalisation divisors of the rows (cdist recipe)
    int32_t* out_cnt;       // (N, 2) one count per epilogue group
    uint32_t* out_q;        // (N, kOutCap) candidate entries (chunk << 8 | quad mask)
    double* loss_partial;   // (ceil(N/32))
    unsigned int* blocks_done;   // (1) + padding; zeroed by vq_forward
    int32_t* fb_rows;            // (2N) rows needing the exact full scan (the GEMM lists each row once)
    int32_t* fb_count;           // (1)
    float4* fb_part;             // (kFbMaxGroups * kFbGroup, kFbMaxParts)
    unsigned int* fb_arrive;     // (kFbMaxGroups)
    unsigned long long* stats;   // (VQ_STAT_COUNT) internal copy when the caller passes none
    size_t control_bytes;        // blocks_done .. fb_arrive, cleared by one memset per call
    size_t bytes;
};

// chunks of 64 the contraction is padded to
inline int chunks_for(int D) { return D <= 64 ? 1 : D <= 128 ? 2 : D <= 256 ? 4 : 8; }

Workspace carve(void* base, int64_t N, int D = vq::kD) {
    Workspace w;
    const int64_t n_pad = round_up(N > 0 ? N : 1, 2 * vq::kRowTile);      // whole PAIRS of GEMM row tiles
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += (size_t)round_up((int64_t)bytes, 256);
        return p;
    };
    w.z_h = static_cast<__half*>(take((size_t)n_pad * (size_t)(chunks_for(D) * vq::kDChunk) * 2));
    w.z2 = static_cast<float*>(take((size_t)n_pad * 4));
    w.z_inv_scale = static_cast<float*>(take((size_t)n_pad * 4));
    w.denom = static_cast<float*>(take((size_t)n_pad * 4));
    w.out_cnt = static_cast<int32_t*>(take((size_t)n_pad * 2 * 4));
    w.out_q = static_cast<uint32_t*>(take((size_t)n_pad * vq::kOutCap * 8));
    w.loss_partial = static_cast<double*>(take((size_t)(n_pad / vq::kSelRows) * 8));
    w.fb_rows = static_cast<int32_t*>(take((size_t)n_pad * 2 * 4));
    w.fb_part = static_cast<float4*>(take((size_t)vq::kFbMaxGroups * vq::kFbGroup * vq::kFbMaxParts * sizeof(float4)));
    // control words, contiguous so that the first kernel of a call clears them all: [blocks_done | fb_count | fb_arrive]
    w.blocks_done = static_cast<unsigned int*>(take(256));
    w.fb_count = static_cast<int32_t*>(take(256));
    w.fb_arrive = static_cast<unsigned int*>(take((size_t)vq::kFbMaxGroups * sizeof(unsigned int)));
    w.control_bytes = 512 + (size_t)vq::kFbMaxGroups * sizeof(unsigned int);
    w.stats = static_cast<unsigned long long*>(take(256));
    w.bytes = off;
    return w;
}

int check_common(const void* z, int64_t B, int64_t HW, int D, int K) {
    if (D != vq::kD) return fail(VQ_E_UNSUPPORTED, "latent_dim D=%d is not supported (kernels are specialised for D=256)", D);
    if (K < 1) return fail(VQ_E_UNSUPPORTED, "K=%d", K);
    if (B < 0 || HW < 0) return fail(VQ_E_INVALID, "negative shape B=%lld HW=%lld", (long long)B, (long long)HW);
    if (B * HW >= (int64_t)1 << 31) return fail(VQ_E_UNSUPPORTED, "N=%lld latents exceed 2^31", (long long)(B * HW));
    if (B * HW > 0 && z == nullptr) return fail(VQ_E_INVALID, "null z pointer");
    return VQ_OK;
}

// shared front half of vq_argmin / vq_forward / vq_debug_scores: prep z + GEMM with fused candidate argmin
// layout of the latents: row-major (N, D) when `rows`, else NCHW (N / HW, D, HW) with the 16-byte fast path when possible
int pick_layout(const float* z, int64_t HW, bool rows) {
    const bool aligned = (reinterpret_cast<uintptr_t>(z) & 15) == 0;
    if (rows) return aligned ? vq::kLayoutRows : vq::kLayoutGeneric;       // (N, D) == NCHW with HW == 1
    return (HW % vq::kSelRows == 0 && aligned) ? vq::kLayoutVec : vq::kLayoutGeneric;
}

// 2-D tensor map of the NCHW activations for vq_qconv_prep_kernel: [B * 256 channel planes][HW] fp32, box 64 planes x 128
// positions.  The encoder is a driver entry point fetched through the runtime (the library links no libcuda); false when this
// driver has none -- the kernel then falls back to one bulk copy per channel plane.
bool encode_h_tensor_map(const float* h, int64_t B, int64_t HW, CUtensorMap* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeFn>(p);
    }();
    if (fn == nullptr || getenv("VQ_QCONV_NO_TENSORMAP") != nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)(B * vq::kD)};
    const cuuint64_t strides[1] = {(cuuint64_t)HW * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)vq::kRowTile, (cuuint32_t)vq::kQcStageCh};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The folded quant_conv (vq_qconv.cuh): when given, the operand preparation starts from the convolution's INPUT h and also
// produces the latents z that every later kernel of the call reads.
struct QconvArgs {
    const float* h;
    const void* w_img;
    const float* w_scalars;
    const float* bias;
};

int run_gemm(const float* z, int64_t N, int64_t HW, bool rows, int recipe, const void* E_h, const float* e2, const float* cb, int K,
             const Workspace& w, float* dbg_scores, int64_t* hist, unsigned long long* stats, cudaStream_t st,
             const QconvArgs* qc = nullptr) {
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const int64_t n_pad = round_up(N, 2 * vq::kRowTile);
    const int k_pad = vq_padded_codes(K);

    const unsigned pgrid = (unsigned)(n_pad / vq::kSelRows);
    vq::PrepClear clr;
    clr.control = w.blocks_done;                       // loss arrival counter, worklist length, group arrivals
    clr.control_words = (int)(w.control_bytes / sizeof(unsigned int));
    clr.hist = reinterpret_cast<unsigned long long*>(hist);
    clr.K = K;
    clr.stats = stats;
    clr.n_stats = VQ_STAT_COUNT;
    if (qc != nullptr) {
        vq::QconvParams qp;
        qp.h = qc->h; qp.N = N; qp.HW = HW; qp.n_pad = n_pad;
        qp.w_img = static_cast<const __half*>(qc->w_img); qp.w_scalars = qc->w_scalars; qp.bias = qc->bias;
        qp.z = const_cast<float*>(z); qp.z_h = w.z_h; qp.z2 = w.z2; qp.z_inv_scale = w.z_inv_scale;
        qp.row_tiles = (int)(N / vq::kRowTile);
        qp.clr = clr;
        const unsigned qgrid = (unsigned)(qp.row_tiles < dev->sms ? qp.row_tiles : dev->sms);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        if (encode_h_tensor_map(qc->h, N / HW, HW, &tmap)) vq::vq_qconv_prep_kernel<true><<<qgrid, vq::kQcThreads, vq::kQcSmemBytes, st>>>(qp, tmap);
        else                                               vq::vq_qconv_prep_kernel<false><<<qgrid, vq::kQcThreads, vq::kQcSmemBytes, st>>>(qp, tmap);
        VQ_LAUNCH_CHECK("vq_qconv_prep_kernel");
    } else {
    switch (pick_layout(z, HW, rows)) {
        case vq::kLayoutRows:
            vq::vq_prep_z_kernel<vq::kLayoutRows><<<pgrid, vq::kPrepThreads, 0, st>>>(z, N, HW, n_pad, w.z_h, w.z2, w.z_inv_scale, clr);
            break;
        case vq::kLayoutVec:
            vq::vq_prep_z_kernel<vq::kLayoutVec><<<pgrid, vq::kPrepThreads, 0, st>>>(z, N, HW, n_pad, w.z_h, w.z2, w.z_inv_scale, clr);
            break;
        default:
            vq::vq_prep_z_kernel<vq::kLayoutGeneric><<<pgrid, vq::kPrepThreads, 0, st>>>(z, N, HW, n_pad, w.z_h, w.z2, w.z_inv_scale, clr);
    }
    VQ_LAUNCH_CHECK("vq_prep_z_kernel");
    }

    vq::GemmParams gp;
    gp.z_h = w.z_h;
    gp.e_h = static_cast<const __half*>(E_h);
    gp.e2 = e2;
    gp.cb = cb;
    gp.z2 = w.z2;
    gp.z_inv_scale = w.z_inv_scale;
    gp.N = N;
    gp.k_tiles = k_pad / vq::kCodeTile;
    gp.row_tiles = (int)(round_up(N, vq::kRowTile) / vq::kRowTile);
    gp.out_cnt = w.out_cnt;
    gp.out_q = w.out_q;
    gp.fb_rows = w.fb_rows;
    gp.fb_count = w.fb_count;
    gp.dbg_scores = dbg_scores;
    gp.recipe = recipe;
    gp.timeline = g_timeline;
    gp.timeline_tiles = g_timeline_tiles;
    const bool prof = g_prof.on && g_prof.n < kProfCap;
    if (prof) {
        while (g_prof.created <= g_prof.n) {
            VQ_CUDA(cudaEventCreate(&g_prof.ev[g_prof.created][0]));
            VQ_CUDA(cudaEventCreate(&g_prof.ev[g_prof.created][1]));
            g_prof.created++;
        }
        VQ_CUDA(cudaEventRecord(g_prof.ev[g_prof.n][0], st));
    }
    // sharing needs at least one pair of row tiles; the debug-score variant stays on independent CTAs
    const bool share = gemm_share() && gp.row_tiles >= 2;
    if (g_timeline != nullptr) VQ_CUDA((launch_gemm<false, true>(gp, share, dev->sms, st)));
    else if (dbg_scores)       VQ_CUDA((launch_gemm<true, false>(gp, false, dev->sms, st)));
    else                       VQ_CUDA((launch_gemm<false, false>(gp, share, dev->sms, st)));
    VQ_LAUNCH_CHECK("vq_argmin_gemm_kernel");
    if (prof) {
        VQ_CUDA(cudaEventRecord(g_prof.ev[g_prof.n][1], st));
        g_prof.n++;
    }
    return VQ_OK;
}

int check_ws(void* ws, size_t ws_bytes, int64_t N, Workspace* w, int D = vq::kD) {
    *w = carve(ws, N, D);
    if (ws == nullptr) return fail(VQ_E_INVALID, "null workspace");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(VQ_E_INVALID, "workspace must be 256-byte aligned");
    if (ws_bytes < w->bytes)
        return fail(VQ_E_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, w->bytes);
    return VQ_OK;
}

static_assert(VQ_RECIPE_EXPANDED == vq::kRecipeExpanded && VQ_RECIPE_DIFFSQ == vq::kRecipeDiffSq &&
              VQ_RECIPE_CDIST_NORMALIZED == vq::kRecipeCdist, "header and kernels disagree");

// ---- row-major nearest-code search at any width D <= 512 (vq_rows.cuh): prep -> GEMM -> exact-scan fallback -> select
template <int kNC>
int run_rows_path(const float* x, int64_t N, int D, int recipe, const float* E, const void* E_h, const float* e2, const float* cb, int K,
                  void* idx, int idx_bits, unsigned long long* stats, const Workspace& w, cudaStream_t st) {
    using C = vq::RowsCfg<kNC>;
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const int64_t n_pad = round_up(N, 2 * vq::kRowTile);
    const int k_pad = vq_padded_codes(K);
    const bool normalize = recipe == vq::kRecipeCdist;
    vq::PrepClear clr;
    clr.control = w.blocks_done;
    clr.control_words = (int)(w.control_bytes / sizeof(unsigned int));
    clr.hist = nullptr;
    clr.K = K;
    clr.stats = stats;
    clr.n_stats = VQ_STAT_COUNT;
    const unsigned pgrid = (unsigned)(n_pad / C::kRows);
    if (normalize) vq::vq_prep_rows_kernel<kNC, true><<<pgrid, vq::kPrepThreads, 0, st>>>(x, N, D, n_pad, w.z_h, w.z2, w.z_inv_scale, w.denom, nullptr, clr);
    else           vq::vq_prep_rows_kernel<kNC, false><<<pgrid, vq::kPrepThreads, 0, st>>>(x, N, D, n_pad, w.z_h, w.z2, w.z_inv_scale, nullptr, nullptr, clr);
    VQ_LAUNCH_CHECK("vq_prep_rows_kernel");

    vq::GemmParams gp;
    gp.z_h = w.z_h;
    gp.e_h = static_cast<const __half*>(E_h);
    gp.e2 = e2;
    gp.cb = cb;
    gp.z2 = w.z2;
    gp.z_inv_scale = w.z_inv_scale;
    gp.N = N;
    gp.k_tiles = k_pad / vq::kCodeTile;
    gp.row_tiles = (int)(round_up(N, vq::kRowTile) / vq::kRowTile);
    gp.out_cnt = w.out_cnt;
    gp.out_q = w.out_q;
    gp.fb_rows = w.fb_rows;
    gp.fb_count = w.fb_count;
    gp.dbg_scores = nullptr;
    gp.recipe = recipe;
    gp.timeline = nullptr;
    gp.timeline_tiles = 0;
    VQ_CUDA((launch_gemm<false, false, kNC>(gp, gemm_share() && gp.row_tiles >= 2, dev->sms, st)));
    VQ_LAUNCH_CHECK("vq_argmin_gemm_kernel");

    vq::SelectRowsParams sp;
    sp.x = x; sp.denom = normalize ? w.denom : nullptr; sp.E = E; sp.e2 = e2; sp.z2 = w.z2;
    sp.out_cnt = w.out_cnt; sp.out_q = w.out_q;
    sp.N = N; sp.D = D; sp.K = K;
    sp.idx = idx; sp.idx_bits = idx_bits; sp.recipe = recipe; sp.stats = stats;
    sp.fb_rows = w.fb_rows; sp.fb_count = w.fb_count;
    sp.part = w.fb_part; sp.arrive = w.fb_arrive;
    VQ_CUDA(launch_chained(vq::vq_fallback_rows_kernel<kNC>, (unsigned)(2 * dev->sms), vq::kSelThreads, st, sp));
    VQ_LAUNCH_CHECK("vq_fallback_rows_kernel");
    VQ_CUDA(launch_chained(vq::vq_select_rows_kernel<kNC>, (unsigned)((N + vq::kSelRows - 1) / vq::kSelRows), vq::kSelThreads, st, sp));
    VQ_LAUNCH_CHECK("vq_select_rows_kernel");
    return VQ_OK;
}

template <int kNC>
int run_table_prep(const float* E, int K, int D, void* E_h, float* e2, float* cb, cudaStream_t st) {
    const int k_pad = vq_padded_codes(K);
    vq::vq_table_norms_kernel<kNC><<<k_pad / vq::kSelRows, vq::kPrepThreads, 0, st>>>(E, K, D, k_pad, e2, cb);
    VQ_LAUNCH_CHECK("vq_table_norms_kernel");
    vq::vq_table_convert_kernel<kNC><<<k_pad / vq::kSelRows, vq::kPrepThreads, 0, st>>>(E, K, D, k_pad, static_cast<__half*>(E_h), cb);
    VQ_LAUNCH_CHECK("vq_table_convert_kernel");
    return VQ_OK;
}

}  // namespace

VQ_EXPORT int vq_abi_version(void) { return 1; }
VQ_EXPORT const char* vq_last_error(void) { return g_err; }
VQ_EXPORT int vq_last_launch_count(void) { return g_launches; }

VQ_EXPORT int vq_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.n = 0;
    return VQ_OK;
}

VQ_EXPORT int vq_profile_collect(float* ms_host, int cap, int* n_host) {
    if (!ms_host || !n_host) return fail(VQ_E_INVALID, "null pointer");
    int n = g_prof.n < cap ? g_prof.n : cap;
    for (int i = 0; i < n; i++) {
        VQ_CUDA(cudaEventSynchronize(g_prof.ev[i][1]));
        VQ_CUDA(cudaEventElapsedTime(&ms_host[i], g_prof.ev[i][0], g_prof.ev[i][1]));
    }
    *n_host = n;
    g_prof.n = 0;
    return VQ_OK;
}

VQ_EXPORT int vq_debug_timeline(long long* stamps_dev, int tiles) {
    g_timeline = stamps_dev;
    g_timeline_tiles = stamps_dev ? tiles : 0;
    return VQ_OK;
}

VQ_EXPORT int vq_device_check(void) {
    DevInfo* d;
    return device_info(&d);
}

VQ_EXPORT int vq_padded_codes(int K) { return (int)round_up(K > 0 ? K : 1, vq::kCodeTile); }

VQ_EXPORT int vq_workspace_bytes(int64_t N, int K, int D, size_t* out) {
    if (out == nullptr) return fail(VQ_E_INVALID, "null out pointer");
    if (D < 1 || D > 512) return fail(VQ_E_UNSUPPORTED, "width D=%d is not supported (1 <= D <= 512; the NCHW CodeBook entry points need 256)", D);
    if (N < 0 || K < 1) return fail(VQ_E_INVALID, "bad shape N=%lld K=%d", (long long)N, K);
    *out = carve(nullptr, N, D).bytes;
    return VQ_OK;
}

VQ_EXPORT int vq_prepare_codebook(const float* E, int K, int D, void* E_h, float* e_norm2, float* cb_scalars,
                                  vq_stream_t stream) {
    g_launches = 0;
    if (D < 1 || D > 512) return fail(VQ_E_UNSUPPORTED, "width D=%d is not supported (1 <= D <= 512)", D);
    if (K < 1) return fail(VQ_E_UNSUPPORTED, "K=%d", K);
    if (!E || !E_h || !e_norm2 || !cb_scalars) return fail(VQ_E_INVALID, "null pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int k_pad = vq_padded_codes(K);
    VQ_CUDA(cudaMemsetAsync(cb_scalars, 0, 4 * sizeof(float), st));
    if (D != vq::kD) {                                        // lookup tables of other widths (operand image of 64-chunk count chunks_for(D))
        switch (chunks_for(D)) {
            case 1: return run_table_prep<1>(E, K, D, E_h, e_norm2, cb_scalars, st);
            case 2: return run_table_prep<2>(E, K, D, E_h, e_norm2, cb_scalars, st);
            case 4: return run_table_prep<4>(E, K, D, E_h, e_norm2, cb_scalars, st);
            default: return run_table_prep<8>(E, K, D, E_h, e_norm2, cb_scalars, st);
        }
    }
    vq::vq_codebook_norms_kernel<<<k_pad / vq::kSelRows, vq::kPrepThreads, 0, st>>>(E, K, k_pad, e_norm2, cb_scalars);
    VQ_LAUNCH_CHECK("vq_codebook_norms_kernel");
    vq::vq_codebook_convert_kernel<<<k_pad / vq::kSelRows, vq::kPrepThreads, 0, st>>>(E, K, k_pad,
                                                                                      static_cast<__half*>(E_h), cb_scalars);
    VQ_LAUNCH_CHECK("vq_codebook_convert_kernel");
    return VQ_OK;
}

static int forward_impl(bool training, bool rows, int recipe, const float* z, int64_t B, int64_t HW, int D, const float* E, const void* E_h,
                        const float* e2, const float* cb, int K, float beta, float* zq, void* idx, int idx_bits, float* loss,
                        int64_t* hist, unsigned long long* stats, void* ws, size_t ws_bytes, vq_stream_t stream,
                        float* code_diff_sum = nullptr, const QconvArgs* qc = nullptr) {
    g_launches = 0;
    int rc = check_common(z, B, HW, D, K);
    if (rc != VQ_OK) return rc;
    if (!E || !E_h || !e2 || !cb) return fail(VQ_E_INVALID, "null codebook pointer");
    const int64_t N = B * HW;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((reinterpret_cast<uintptr_t>(code_diff_sum) & 15) != 0) return fail(VQ_E_INVALID, "code_diff_sum must be 16-byte aligned");
    if (training && code_diff_sum) VQ_CUDA(cudaMemsetAsync(code_diff_sum, 0, (size_t)K * D * sizeof(float), st));
    if (N == 0) {
        // nothing is launched: clear the outputs directly
        if (stats) VQ_CUDA(cudaMemsetAsync(stats, 0, VQ_STAT_COUNT * sizeof(unsigned long long), st));
        if (training && hist) VQ_CUDA(cudaMemsetAsync(hist, 0, (size_t)K * sizeof(int64_t), st));
        // torch.mean over an empty tensor is NaN (codebook.py:96)
        if (training && loss) {
            const float nanv = __builtin_nanf("");
            VQ_CUDA(cudaMemcpyAsync(loss, &nanv, sizeof(float), cudaMemcpyHostToDevice, st));
        }
        return VQ_OK;
    }
    if (!idx) return fail(VQ_E_INVALID, "null idx pointer");
    if (training && !loss) return fail(VQ_E_INVALID, "null loss pointer");
    // the kernels read code rows and write z_q rows with 16-byte accesses: refuse pointers that would fault instead
    if ((reinterpret_cast<uintptr_t>(E) & 15) != 0 || (reinterpret_cast<uintptr_t>(zq) & 15) != 0)
        return fail(VQ_E_INVALID, "E and zq must be 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(idx) & (uintptr_t)(idx_bits / 8 - 1)) != 0) return fail(VQ_E_INVALID, "idx is not aligned to its element size");
    Workspace w;
    rc = check_ws(ws, ws_bytes, N, &w);
    if (rc != VQ_OK) return rc;

    if (idx_bits != 64 && idx_bits != 32 && idx_bits != 16) return fail(VQ_E_INVALID, "idx_bits must be 16, 32 or 64, got %d", idx_bits);
    if (idx_bits == 16 && K > 65536) return fail(VQ_E_INVALID, "16-bit indices need K <= 65536, got K=%d", K);
    if (recipe != vq::kRecipeExpanded && recipe != vq::kRecipeDiffSq) return fail(VQ_E_INVALID, "unknown distance recipe %d", recipe);
    rc = run_gemm(z, N, HW, rows, recipe, E_h, e2, cb, K, w, nullptr, training ? hist : nullptr, stats, st, qc);
    if (rc != VQ_OK) return rc;

    {   // rows whose candidate list overflowed (rare) or that hold Inf / NaN: exact scan; a no-op when the worklist is empty
        vq::FallbackParams fp;
        fp.z = z; fp.E = E; fp.e2 = e2; fp.z2 = w.z2;
        fp.fb_rows = w.fb_rows; fp.fb_count = w.fb_count;
        fp.HW = HW; fp.K = K; fp.recipe = recipe;
        fp.out_cnt = w.out_cnt; fp.out_q = w.out_q; fp.stats = stats;
        fp.part = w.fb_part; fp.arrive = w.fb_arrive;
        DevInfo* dev;
        rc = device_info(&dev);
        if (rc != VQ_OK) return rc;
        // the kernel sizes its own work split from the worklist length so that one wave of resident CTAs covers it
        const unsigned fgrid = (unsigned)(vq::kFbCtasPerSm * dev->sms);
        if (recipe == vq::kRecipeDiffSq) VQ_CUDA(launch_chained(vq::vq_fallback_kernel<true>, fgrid, vq::kFbThreads, st, fp));
        else                             VQ_CUDA(launch_chained(vq::vq_fallback_kernel<false>, fgrid, vq::kFbThreads, st, fp));
        VQ_LAUNCH_CHECK("vq_fallback_kernel");
    }

    vq::SelectParams sp;
    sp.z = z; sp.E = E; sp.e2 = e2; sp.z2 = w.z2;
    sp.out_cnt = w.out_cnt; sp.out_q = w.out_q;
    sp.N = N; sp.HW = HW; sp.K = K; sp.beta = beta;
    sp.idx = idx; sp.idx_bits = idx_bits; sp.recipe = recipe; sp.zq = zq;
    sp.scat = training ? code_diff_sum : nullptr;
    sp.hist = reinterpret_cast<unsigned long long*>(hist);
    sp.loss_partial = w.loss_partial;
    sp.stats = stats;
    const unsigned grid = (unsigned)((N + vq::kSelRows - 1) / vq::kSelRows);
    const int layout = pick_layout(z, HW, rows);
    if (training) {
        if (layout == vq::kLayoutVec) VQ_CUDA(launch_chained(vq::vq_select_kernel<true, vq::kLayoutVec>, grid, vq::kSelThreads, st, sp));
        else                          VQ_CUDA(launch_chained(vq::vq_select_kernel<true, vq::kLayoutGeneric>, grid, vq::kSelThreads, st, sp));
    } else {
        if (layout == vq::kLayoutRows)     VQ_CUDA(launch_chained(vq::vq_select_kernel<false, vq::kLayoutRows>, grid, vq::kSelThreads, st, sp));
        else if (layout == vq::kLayoutVec) VQ_CUDA(launch_chained(vq::vq_select_kernel<false, vq::kLayoutVec>, grid, vq::kSelThreads, st, sp));
        else                               VQ_CUDA(launch_chained(vq::vq_select_kernel<false, vq::kLayoutGeneric>, grid, vq::kSelThreads, st, sp));
    }
    VQ_LAUNCH_CHECK("vq_select_kernel");
    if (training) {
        VQ_CUDA(launch_chained(vq::vq_loss_finalize_kernel, 1u, vq::kSelThreads, st, (const double*)w.loss_partial, grid, N, beta, loss));
        VQ_LAUNCH_CHECK("vq_loss_finalize_kernel");
    }
    return VQ_OK;
}

VQ_EXPORT int vq_argmin(const float* z_nchw, int64_t B, int64_t HW, int D, const float* E, const void* E_h,
                        const float* e_norm2, const float* cb_scalars, int K, int64_t* idx, unsigned long long* stats,
                        void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    return forward_impl(false, false, VQ_RECIPE_EXPANDED, z_nchw, B, HW, D, E, E_h, e_norm2, cb_scalars, K, 0.0f, nullptr, idx, 64, nullptr,
                        nullptr, stats, workspace, workspace_bytes, stream);
}

VQ_EXPORT int vq_argmin_narrow(const float* z_nchw, int64_t B, int64_t HW, int D, const float* E, const void* E_h,
                               const float* e_norm2, const float* cb_scalars, int K, void* idx, int idx_bits,
                               unsigned long long* stats, void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    return forward_impl(false, false, VQ_RECIPE_EXPANDED, z_nchw, B, HW, D, E, E_h, e_norm2, cb_scalars, K, 0.0f, nullptr, idx, idx_bits,
                        nullptr, nullptr, stats, workspace, workspace_bytes, stream);
}

VQ_EXPORT int vq_argmin_rows(const float* x_rows, int64_t N, int D, const float* E, const void* E_h, const float* e_norm2,
                             const float* cb_scalars, int K, int recipe, void* idx, int idx_bits, unsigned long long* stats,
                             void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    if (D != vq::kD || recipe == vq::kRecipeCdist) {
        // any other width, and the normalise + cdist recipe: the generic row-major path (vq_rows.cuh)
        g_launches = 0;
        if (D < 1 || D > 512) return fail(VQ_E_UNSUPPORTED, "width D=%d is not supported (1 <= D <= 512)", D);
        if (K < 1) return fail(VQ_E_UNSUPPORTED, "K=%d", K);
        if (N < 0 || N >= (int64_t)1 << 31) return fail(VQ_E_INVALID, "bad row count N=%lld", (long long)N);
        if (recipe != vq::kRecipeExpanded && recipe != vq::kRecipeDiffSq && recipe != vq::kRecipeCdist)
            return fail(VQ_E_INVALID, "unknown distance recipe %d", recipe);
        if (idx_bits != 64 && idx_bits != 32 && idx_bits != 16) return fail(VQ_E_INVALID, "idx_bits must be 16, 32 or 64, got %d", idx_bits);
        if (idx_bits == 16 && K > 65536) return fail(VQ_E_INVALID, "16-bit indices need K <= 65536, got K=%d", K);
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        if (N == 0) {
            if (stats) VQ_CUDA(cudaMemsetAsync(stats, 0, VQ_STAT_COUNT * sizeof(unsigned long long), st));
            return VQ_OK;
        }
        if (!x_rows || !E || !E_h || !e_norm2 || !cb_scalars || !idx) return fail(VQ_E_INVALID, "null pointer");
        Workspace w;
        int rc = check_ws(workspace, workspace_bytes, N, &w, D);
        if (rc != VQ_OK) return rc;
        switch (chunks_for(D)) {
            case 1: return run_rows_path<1>(x_rows, N, D, recipe, E, E_h, e_norm2, cb_scalars, K, idx, idx_bits, stats, w, st);
            case 2: return run_rows_path<2>(x_rows, N, D, recipe, E, E_h, e_norm2, cb_scalars, K, idx, idx_bits, stats, w, st);
            case 4: return run_rows_path<4>(x_rows, N, D, recipe, E, E_h, e_norm2, cb_scalars, K, idx, idx_bits, stats, w, st);
            default: return run_rows_path<8>(x_rows, N, D, recipe, E, E_h, e_norm2, cb_scalars, K, idx, idx_bits, stats, w, st);
        }
    }
    return forward_impl(false, true, recipe, x_rows, N, 1, D, E, E_h, e_norm2, cb_scalars, K, 0.0f, nullptr, idx, idx_bits, nullptr,
                        nullptr, stats, workspace, workspace_bytes, stream);
}

VQ_EXPORT int vq_normalize_rows(const float* x_rows, int64_t N, int D, float* out_rows, vq_stream_t stream) {
    g_launches = 0;
    if (D < 1 || D > 512) return fail(VQ_E_UNSUPPORTED, "width D=%d is not supported (1 <= D <= 512)", D);
    if (N < 0) return fail(VQ_E_INVALID, "bad row count N=%lld", (long long)N);
    if (N == 0) return VQ_OK;
    if (!x_rows || !out_rows) return fail(VQ_E_INVALID, "null pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    vq::PrepClear clr = {};
#define VQ_NORM_LAUNCH(NC)                                                                                                   \
    vq::vq_prep_rows_kernel<NC, true><<<(unsigned)((N + vq::RowsCfg<NC>::kRows - 1) / vq::RowsCfg<NC>::kRows), vq::kPrepThreads, 0, st>>>( \
        x_rows, N, D, N, nullptr, nullptr, nullptr, nullptr, out_rows, clr)
    switch (chunks_for(D)) {
        case 1: VQ_NORM_LAUNCH(1); break;
        case 2: VQ_NORM_LAUNCH(2); break;
        case 4: VQ_NORM_LAUNCH(4); break;
        default: VQ_NORM_LAUNCH(8);
    }
#undef VQ_NORM_LAUNCH
    VQ_LAUNCH_CHECK("vq_prep_rows_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_forward(const float* z_nchw, int64_t B, int64_t HW, int D, const float* E, const void* E_h,
                         const float* e_norm2, const float* cb_scalars, int K, float beta, float* zq_nhwc, int64_t* idx,
                         float* loss, int64_t* hist, unsigned long long* stats, void* workspace, size_t workspace_bytes,
                         vq_stream_t stream) {
    return forward_impl(true, false, VQ_RECIPE_EXPANDED, z_nchw, B, HW, D, E, E_h, e_norm2, cb_scalars, K, beta, zq_nhwc, idx, 64, loss, hist,
                        stats, workspace, workspace_bytes, stream);
}

VQ_EXPORT int vq_forward_ex(const float* z_nchw, int64_t B, int64_t HW, int D, const float* E, const void* E_h,
                            const float* e_norm2, const float* cb_scalars, int K, float beta, float* zq_nhwc, int64_t* idx,
                            float* loss, int64_t* hist, float* code_diff_sum, unsigned long long* stats, void* workspace,
                            size_t workspace_bytes, vq_stream_t stream) {
    return forward_impl(true, false, VQ_RECIPE_EXPANDED, z_nchw, B, HW, D, E, E_h, e_norm2, cb_scalars, K, beta, zq_nhwc, idx, 64, loss, hist,
                        stats, workspace, workspace_bytes, stream, code_diff_sum);
}

VQ_EXPORT int vq_prepare_quant_conv(const float* W, void* w_img, float* w_scalars, vq_stream_t stream) {
    g_launches = 0;
    if (!W || !w_img || !w_scalars) return fail(VQ_E_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(w_img) & 15) != 0) return fail(VQ_E_INVALID, "w_img must be 16-byte aligned");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    vq::vq_qconv_weight_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, static_cast<__half*>(w_img), w_scalars);
    VQ_LAUNCH_CHECK("vq_qconv_weight_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_forward_qconv(const float* h_nchw, int64_t B, int64_t HW, int D, const void* w_img, const float* w_scalars,
                               const float* bias, float* z_nchw, const float* E, const void* E_h, const float* e_norm2,
                               const float* cb_scalars, int K, float beta, float* zq_nhwc, int64_t* idx, float* loss, int64_t* hist,
                               unsigned long long* stats, void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    if (D != vq::kD) return fail(VQ_E_UNSUPPORTED, "latent_dim D=%d is not supported (the folded quant_conv is 256 -> 256)", D);
    if (B < 1 || HW < 1) return fail(VQ_E_INVALID, "bad shape B=%lld HW=%lld", (long long)B, (long long)HW);
    if (HW % vq::kRowTile != 0)
        return fail(VQ_E_UNSUPPORTED, "HW=%lld: the folded quant_conv needs HW %% 128 == 0 (a row tile is 128 positions of one image)",
                    (long long)HW);
    if (!h_nchw || !w_img || !w_scalars || !z_nchw) return fail(VQ_E_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(h_nchw) & 15) != 0 || (reinterpret_cast<uintptr_t>(z_nchw) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(w_img) & 15) != 0)
        return fail(VQ_E_INVALID, "h, z and w_img must be 16-byte aligned");
    QconvArgs qc{h_nchw, w_img, w_scalars, bias};
    return forward_impl(true, false, VQ_RECIPE_EXPANDED, z_nchw, B, HW, D, E, E_h, e_norm2, cb_scalars, K, beta, zq_nhwc, idx, 64, loss, hist,
                        stats, workspace, workspace_bytes, stream, nullptr, &qc);
}

VQ_EXPORT int vq_debug_scores(const float* z_nchw, int64_t B, int64_t HW, int D, const void* E_h, const float* e_norm2,
                              const float* cb_scalars, int K, float* scores, void* workspace, size_t workspace_bytes,
                              vq_stream_t stream) {
    g_launches = 0;
    int rc = check_common(z_nchw, B, HW, D, K);
    if (rc != VQ_OK) return rc;
    if (!E_h || !e_norm2 || !cb_scalars || !scores) return fail(VQ_E_INVALID, "null pointer");
    const int64_t N = B * HW;
    if (N == 0) return VQ_OK;
    Workspace w;
    rc = check_ws(workspace, workspace_bytes, N, &w);
    if (rc != VQ_OK) return rc;
    return run_gemm(z_nchw, N, HW, false, VQ_RECIPE_EXPANDED, E_h, e_norm2, cb_scalars, K, w, scores, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

static int backward_impl(const float* gout, const int64_t* gout_strides, float g_loss, const float* g_loss_dev,
                         const float* z_nchw, const int64_t* idx, const float* E, int64_t B, int64_t HW, int D, int K, float beta,
                         int64_t n_global, float grad_E_scale, bool deterministic, const float* code_diff_sum, float* grad_z,
                         float* grad_E, void* ws, size_t ws_bytes, vq_stream_t stream) {
    g_launches = 0;
    int rc = check_common(z_nchw, B, HW, D, K);
    if (rc != VQ_OK) return rc;
    const int64_t N = B * HW;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((reinterpret_cast<uintptr_t>(grad_E) & 15) != 0 || (reinterpret_cast<uintptr_t>(E) & 15) != 0)
        return fail(VQ_E_INVALID, "E and grad_E must be 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(code_diff_sum) & 15) != 0) return fail(VQ_E_INVALID, "code_diff_sum must be 16-byte aligned");
    const bool from_sum = code_diff_sum != nullptr && grad_E != nullptr;      // the forward already accumulated sum (e - z) per code
    if (grad_E && !from_sum) VQ_CUDA(cudaMemsetAsync(grad_E, 0, (size_t)K * D * sizeof(float), st));
    if (!grad_z && !grad_E) return VQ_OK;
    if (n_global <= 0) n_global = N;
    float* grad_E_out = grad_E;
    if (from_sum) {
        const int64_t n_elems = (int64_t)K * D;
        const unsigned sgrid = (unsigned)((n_elems + 1023) / 1024 < 4096 ? (n_elems + 1023) / 1024 : 4096);
        vq::vq_grad_from_sum_kernel<<<sgrid, 256, 0, st>>>(code_diff_sum, n_elems, g_loss, g_loss_dev, 1.0 / ((double)n_global * (double)D), beta,
                                                           grad_E_scale, grad_E_out);
        VQ_LAUNCH_CHECK("vq_grad_from_sum_kernel");
        grad_E = nullptr;                                      // the main kernel only has grad_z left to do
        deterministic = false;
    }
    if (N == 0 || (!grad_z && !grad_E)) return VQ_OK;
    if (!idx || !E) return fail(VQ_E_INVALID, "null idx/E pointer");
    if (gout && !gout_strides) return fail(VQ_E_INVALID, "gout given without strides");
    DevInfo* dev;
    rc = device_info(&dev);
    if (rc != VQ_OK) return rc;

    vq::BackwardParams bp;
    bp.gout = gout;
    bp.gs_b = gout ? gout_strides[0] : 0;
    bp.gs_d = gout ? gout_strides[1] : 0;
    bp.gs_hw = gout ? gout_strides[2] : 0;
    bp.z = z_nchw; bp.idx = idx; bp.E = E;
    bp.N = N; bp.HW = HW; bp.K = K;
    bp.g_loss = g_loss;
    bp.g_loss_dev = g_loss_dev;
    bp.inv_nd = 1.0 / ((double)n_global * (double)D);
    bp.beta = beta;
    bp.e_scale = grad_E_scale;
    bp.grad_z = grad_z; bp.grad_E = grad_E;
    bp.acc_fx = nullptr; bp.fx_shift = nullptr;
    const unsigned grid = (unsigned)((N + vq::kSelRows - 1) / vq::kSelRows);
    const bool det = deterministic && grad_E != nullptr;
    if (det) {
        // workspace: [acc (K, D) int64 | maxbits | shift]
        const size_t need = (size_t)K * D * sizeof(long long) + 256;
        if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 255) != 0)
            return fail(VQ_E_INVALID, "deterministic backward needs a 256-byte aligned workspace");
        if (ws_bytes < need) return fail(VQ_E_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, need);
        bp.acc_fx = static_cast<long long*>(ws);
        unsigned int* maxbits = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + (size_t)K * D * sizeof(long long));
        int* shift = reinterpret_cast<int*>(maxbits + 1);
        bp.fx_shift = shift;
        VQ_CUDA(cudaMemsetAsync(ws, 0, need, st));
        vq::vq_backward_maxdiff_kernel<<<grid, vq::kBwdThreads, 0, st>>>(z_nchw, idx, E, N, HW, K, maxbits);
        VQ_LAUNCH_CHECK("vq_backward_maxdiff_kernel");
        vq::vq_backward_fxshift_kernel<<<1, 1, 0, st>>>(maxbits, shift);
        VQ_LAUNCH_CHECK("vq_backward_fxshift_kernel");
    }
    // channels-last upstream gradient (d contiguous; the layout of the z_q we returned) is read lanes-over-d, an
    // hw-contiguous one like z; 16-byte tile accesses need hw-contiguous, 16-byte aligned 32-latent tiles
    const bool cl = gout != nullptr && bp.gs_d == 1 && !(bp.gs_hw == 1 && HW > 1) && bp.gs_b % 4 == 0 && bp.gs_hw % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
    const bool vec = (HW % vq::kSelRows == 0) && ((reinterpret_cast<uintptr_t>(z_nchw) & 15) == 0) &&
                     (grad_z == nullptr || (reinterpret_cast<uintptr_t>(grad_z) & 15) == 0);
#define VQ_BWD_LAUNCH(V, C)                                                                          \
    do {                                                                                             \
        if (det) vq::vq_backward_kernel<V, C, true><<<grid, vq::kBwdThreads, 0, st>>>(bp);            \
        else     vq::vq_backward_kernel<V, C, false><<<grid, vq::kBwdThreads, 0, st>>>(bp);           \
    } while (0)
    if (vec) {
        if (cl) VQ_BWD_LAUNCH(true, true);
        else    VQ_BWD_LAUNCH(true, false);
    } else {
        if (cl) VQ_BWD_LAUNCH(false, true);
        else    VQ_BWD_LAUNCH(false, false);
    }
#undef VQ_BWD_LAUNCH
    VQ_LAUNCH_CHECK("vq_backward_kernel");
    if (det) {
        const int64_t n_elems = (int64_t)K * D;
        const unsigned fgrid = (unsigned)((n_elems + 256 * 8 - 1) / (256 * 8));
        vq::vq_backward_fxfinish_kernel<<<fgrid, 256, 0, st>>>(bp.acc_fx, bp.fx_shift, n_elems, g_loss, g_loss_dev, bp.inv_nd, beta,
                                                              grad_E_scale, grad_E);
        VQ_LAUNCH_CHECK("vq_backward_fxfinish_kernel");
    }
    return VQ_OK;
}

VQ_EXPORT int vq_backward(const float* gout, const int64_t* gout_strides, float g_loss, const float* g_loss_dev,
                          const float* z_nchw,
                          const int64_t* idx, const float* E, int64_t B, int64_t HW, int D, int K, float beta,
                          int64_t n_global, float* grad_z, float* grad_E, vq_stream_t stream) {
    return backward_impl(gout, gout_strides, g_loss, g_loss_dev, z_nchw, idx, E, B, HW, D, K, beta, n_global, 1.0f, false, nullptr, grad_z,
                         grad_E, nullptr, 0, stream);
}

VQ_EXPORT int vq_backward_workspace_bytes(int K, int D, size_t* out) {
    if (out == nullptr) return fail(VQ_E_INVALID, "null out pointer");
    if (K < 1 || D < 1) return fail(VQ_E_INVALID, "bad shape K=%d D=%d", K, D);
    *out = (size_t)K * D * sizeof(long long) + 256;
    return VQ_OK;
}

VQ_EXPORT int vq_backward_ex(const float* gout, const int64_t* gout_strides, float g_loss, const float* g_loss_dev,
                             const float* z_nchw, const int64_t* idx, const float* E, int64_t B, int64_t HW, int D, int K,
                             float beta, int64_t n_global, float grad_E_scale, int deterministic, const float* code_diff_sum,
                             float* grad_z, float* grad_E, void* workspace, size_t workspace_bytes, vq_stream_t stream) {
    return backward_impl(gout, gout_strides, g_loss, g_loss_dev, z_nchw, idx, E, B, HW, D, K, beta, n_global, grad_E_scale,
                         deterministic != 0, code_diff_sum, grad_z, grad_E, workspace, workspace_bytes, stream);
}

VQ_EXPORT int vq_allreduce_multimem(void* multicast_ptr, void* const* signal_pads_dev, int rank, int world, int64_t n_floats,
                                    unsigned int* local_sync, vq_stream_t stream) {
    g_launches = 0;
    if (!multicast_ptr || !signal_pads_dev || !local_sync) return fail(VQ_E_INVALID, "null pointer");
    if (world < 2 || world > 32 || rank < 0 || rank >= world) return fail(VQ_E_INVALID, "bad rank %d / world %d", rank, world);
    if (n_floats <= 0 || n_floats % (4 * (int64_t)world) != 0)
        return fail(VQ_E_INVALID, "n_floats=%lld must be a positive multiple of 4 * world", (long long)n_floats);
    if ((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) != 0) return fail(VQ_E_INVALID, "multicast pointer must be 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(local_sync) & 7) != 0) return fail(VQ_E_INVALID, "local_sync must be 8-byte aligned");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const int64_t n_vec4 = n_floats / 4, slice = n_vec4 / world;
    int64_t blocks = (slice + vq::kArThreads * vq::kArUnroll - 1) / (vq::kArThreads * vq::kArUnroll);
    int64_t max_blocks = std::min<int64_t>(vq::kArMaxBlocks, dev->sms);                        // the CTAs wait for each other: co-resident
    if (const char* e = getenv("VQ_AR_MAX_BLOCKS")) {                                          // tuning runs (tools/dp_overhead.py)
        const int64_t v = atoll(e);
        if (v >= 1 && v <= dev->sms) max_blocks = v;
    }
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    vq::vq_allreduce_multimem_kernel<<<(unsigned)blocks, vq::kArThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<float*>(multicast_ptr), reinterpret_cast<uint32_t* const*>(signal_pads_dev), rank, world, n_vec4, local_sync);
    VQ_LAUNCH_CHECK("vq_allreduce_multimem_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_pack_stats(const int64_t* hist, const float* loss, int K, float* tail, vq_stream_t stream) {
    g_launches = 0;
    if (!hist || !loss || !tail) return fail(VQ_E_INVALID, "null pointer");
    if (K < 1) return fail(VQ_E_INVALID, "bad K=%d", K);
    if ((reinterpret_cast<uintptr_t>(hist) & 7) != 0 || (reinterpret_cast<uintptr_t>(loss) & 3) != 0 || (reinterpret_cast<uintptr_t>(tail) & 3) != 0)
        return fail(VQ_E_INVALID, "misaligned pointer");
    vq::vq_pack_stats_kernel<<<(unsigned)((K + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(hist), loss, K, tail);
    VQ_LAUNCH_CHECK("vq_pack_stats_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_embed_nchw(const int64_t* idx, const float* E, int64_t B, int64_t HW, int D, int K, float* out,
                            vq_stream_t stream) {
    g_launches = 0;
    if (D != vq::kD) return fail(VQ_E_UNSUPPORTED, "latent_dim D=%d is not supported (D must be 256)", D);
    if (B < 0 || HW < 0 || K < 1) return fail(VQ_E_INVALID, "bad shape");
    const int64_t N = B * HW;
    if (N == 0) return VQ_OK;
    if (!idx || !E || !out) return fail(VQ_E_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(idx) & 7) != 0 || (reinterpret_cast<uintptr_t>(E) & 3) != 0 || (reinterpret_cast<uintptr_t>(out) & 3) != 0)
        return fail(VQ_E_INVALID, "misaligned pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const unsigned grid = (unsigned)((N + vq::kSelRows - 1) / vq::kSelRows);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((HW % vq::kSelRows == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0))
        vq::vq_embed_nchw_kernel<true><<<grid, vq::kBwdThreads, 0, st>>>(idx, E, N, HW, K, out);
    else
        vq::vq_embed_nchw_kernel<false><<<grid, vq::kBwdThreads, 0, st>>>(idx, E, N, HW, K, out);
    VQ_LAUNCH_CHECK("vq_embed_nchw_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_index_to_log_onehot(const int64_t* idx, int64_t B, int64_t L, int num_classes, float clamp_min, float* out,
                                     vq_stream_t stream) {
    g_launches = 0;
    if (B < 0 || L < 0 || num_classes < 1) return fail(VQ_E_INVALID, "bad shape B=%lld L=%lld num_classes=%d", (long long)B, (long long)L, num_classes);
    if (B == 0 || L == 0) return VQ_OK;
    if (!idx || !out) return fail(VQ_E_INVALID, "null pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t threads = vec ? B * (L / 4) : B * L;
    const int64_t gx = (threads + vq::kTokThreads - 1) / vq::kTokThreads;
    const int64_t gy = (num_classes + vq::kTokClassTile - 1) / vq::kTokClassTile;
    if (gx > 0x7fffffffLL || gy > 65535) return fail(VQ_E_UNSUPPORTED, "one-hot grid %lld x %lld too large", (long long)gx, (long long)gy);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)gx, (unsigned)gy);
    if (vec) vq::vq_log_onehot_kernel<true><<<grid, vq::kTokThreads, 0, st>>>(idx, B, L, num_classes, clamp_min, out);
    else     vq::vq_log_onehot_kernel<false><<<grid, vq::kTokThreads, 0, st>>>(idx, B, L, num_classes, clamp_min, out);
    VQ_LAUNCH_CHECK("vq_log_onehot_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_log_onehot_to_index(const float* log_x, int64_t B, int64_t L, int num_classes, int64_t* out, vq_stream_t stream) {
    g_launches = 0;
    if (B < 0 || L < 0 || num_classes < 1) return fail(VQ_E_INVALID, "bad shape B=%lld L=%lld num_classes=%d", (long long)B, (long long)L, num_classes);
    if (B == 0 || L == 0) return VQ_OK;
    if (!log_x || !out) return fail(VQ_E_INVALID, "null pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(log_x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t threads = vec ? B * (L / 4) : B * L;
    const int64_t gx = (threads + vq::kTokThreads - 1) / vq::kTokThreads;
    if (gx > 0x7fffffffLL) return fail(VQ_E_UNSUPPORTED, "argmax grid %lld too large", (long long)gx);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec) vq::vq_argmax_classes_kernel<true><<<(unsigned)gx, vq::kTokThreads, 0, st>>>(log_x, B, L, num_classes, out);
    else     vq::vq_argmax_classes_kernel<false><<<(unsigned)gx, vq::kTokThreads, 0, st>>>(log_x, B, L, num_classes, out);
    VQ_LAUNCH_CHECK("vq_argmax_classes_kernel");
    return VQ_OK;
}

VQ_EXPORT int vq_mask_replace(const int64_t* indices, const float* mask, const int64_t* random_indices, int64_t sos_token,
                              int64_t B, int64_t L, int64_t* out, vq_stream_t stream) {
    g_launches = 0;
    if (B < 0 || L < 0) return fail(VQ_E_INVALID, "bad shape B=%lld L=%lld", (long long)B, (long long)L);
    if (B == 0) return VQ_OK;
    if (!out || (L > 0 && (!indices || !mask || !random_indices))) return fail(VQ_E_INVALID, "null pointer");
    DevInfo* dev;
    int rc = device_info(&dev);
    if (rc != VQ_OK) return rc;
    const int64_t gx = (B * (L + 1) + vq::kTokThreads - 1) / vq::kTokThreads;
    if (gx > 0x7fffffffLL) return fail(VQ_E_UNSUPPORTED, "token grid %lld too large", (long long)gx);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    vq::vq_mask_replace_kernel<<<(unsigned)gx, vq::kTokThreads, 0, st>>>(indices, mask, random_indices, sos_token, B, L, out);
    VQ_LAUNCH_CHECK("vq_mask_replace_kernel");
    return VQ_OK;
}
