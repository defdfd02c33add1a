// vq_qconv.cuh -- quant_conv folded into the operand preparation of the quantiser (the encoder side).
//
// The reference runs z = quant_conv(encoder(x)) -- a 1 x 1 convolution Conv2d(256, 256, 1), network/vqvae/vqvae.py:83,128 --
// right before the CodeBook.  vq_qconv_prep_kernel computes that convolution with fp32 accuracy ON THE TENSOR CORES and, from
// the accumulator tile while it is still on chip, everything vq_prep_z_kernel would derive from z: the fp32 z itself (NCHW:
// the exact stage, z_q, the loss and the backward need it), the fp16 operand image of the distance GEMM, |z|^2 in the
// canonical order and the row scale.  HBM traffic per latent: read h 4 D, write z 4 D + image 2 D -- the unfused pair
// (convolution, then vq_prep_z) reads z once more (4 D).
//
// fp32 accuracy from fp16 tensor-core operands: both operands are split x = hi + lo, hi = fp16(x s), lo = fp16(x s - hi) with
// s an exact power of two that puts the largest magnitude in [2^14, 2^15) (per latent row for h, per tensor for the
// weight): hi + lo carries 22 significant bits, and  h W^T = (hi_h hi_W + lo_h hi_W + hi_h lo_W) / (s_h s_W)  up to the dropped
// lo lo term (2^-22 relative per product) -- three fp16 products into ONE fp32 accumulator in TMEM.  (tf32 operands would
// double the shared-memory footprint of every tile; bf16 needs three pieces per operand and six products.)
//
// One CTA per SM, persistent over row tiles of 128 latents (128 consecutive hw positions of one image: HW % 128 == 0),
// twelve warps:
//   warp 0      h producer: the tile's fp32 source, 64 channels x 128 latents (32 KiB) at a time through a two-slot ring, ONE
//               2-D tensor-map copy per stage (sixty-four 512-byte bulk copies per stage, one per channel plane, kept the TMA
//               unit busy for ~2 us a stage: 268 us for the cfg4 call against 156 us with the tensor map) -- TWICE per tile: the
//               first pass finds every row's maximum (DRAM), the second (L2 hits) is converted
//   warp 1      W producer: the weight's operand images [hi | lo][4 chunks of 64 input channels][256 x 64 fp16] (256 KiB in
//               L2) through a two-stage ring, 32 KiB per stage
//   warp 2      MMA issuer (tcgen05.mma M128 N256 K16, kind::f16): per chunk hi_h hi_W, lo_h hi_W, hi_h lo_W; owns TMEM
//   warps 4-7   converters, thread <-> latent row: row maximum, then hi / lo fp16 rows written as SWIZZLE_128B K-major operand
//               rows (16-byte stores, conflict-free) + fence.proxy.async
//   warps 8-11  epilogue, thread <-> TMEM lane <-> latent row, two passes over the 256 accumulator columns: z = acc / (s_h s_W)
//               + bias (one fma), stored NCHW (a warp store = 128 contiguous bytes of one channel), |z|^2 chains, row
//               maximum; then the fp16 operand row of the distance GEMM.  Two accumulators (all 512 TMEM columns): the
//               epilogue of tile i overlaps the loads / conversions / MMAs of tile i + 1.
#pragma once

#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "ptx_sm100.cuh"
#include "vq_common.cuh"
#include "vq_prep.cuh"

namespace vq {

// 2-D tiled tensor-map load global -> shared, completion on an mbarrier: one instruction moves a [box rows x box columns] tile
// of a strided tensor (here: 64 channel planes x 128 consecutive positions of the NCHW activations).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int32_t x, int32_t y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

// (Measured and dropped, cfg4-sized call with a one-tile codebook, forward 0.477 ms: a third stage in the source ring -- the bias
// then read from global memory -- 0.490 ms; the next TMEM load in flight while the current accumulator chunk is processed,
// 167 registers, 0.482 ms; both, 0.507 ms.  profiles/r2_ab_qconv_variants.jsonl)
constexpr int kQcHsSlots = 2;             // stages of the fp32 source ring
constexpr int kQcThreads = 384;
constexpr int kQcWarpH = 0, kQcWarpW = 1, kQcWarpMma = 2;      // (warp 3 idles: the converter / epilogue groups stay 4-aligned)
constexpr int kQcWarpConv0 = 4, kQcWarpEpi0 = 8;
constexpr int kQcStageCh = 64;                                  // channels per source stage == one 64-wide contraction chunk
constexpr uint32_t kQcBytesHs = kQcStageCh * kRowTile * 4;      // 32 KiB
constexpr uint32_t kQcBytesA = kRowTile * kDChunk * 2;          // 16 KiB: [128 latents][64] fp16
constexpr uint32_t kQcBytesB = kD * kDChunk * 2;                // 32 KiB: [256 output channels][64] fp16
constexpr int kQcWImgElems = 2 * kNumDChunks * kD * kDChunk;    // halves in the weight's operand images (hi | lo): 256 KiB

struct QcSmem {
    alignas(1024) uint8_t a_hi[2][kQcBytesA];
    alignas(1024) uint8_t a_lo[2][kQcBytesA];
    alignas(1024) uint8_t b[2][kQcBytesB];
    alignas(128) float hs[kQcHsSlots][kQcStageCh][kRowTile];
    float bias[kD];
    float hinv[2][kRowTile];                                    // 1 / s_h of the tile's rows (by tile parity)
    alignas(8) uint64_t hs_full[kQcHsSlots];
    uint64_t hs_empty[kQcHsSlots];
    uint64_t a_full[2];
    uint64_t a_empty[2];
    uint64_t b_full[2];
    uint64_t b_empty[2];
    uint64_t t_full[2];
    uint64_t t_empty[2];
    uint64_t hinv_full[2];
    uint64_t hinv_empty[2];
    uint32_t tmem_base;
};
constexpr size_t kQcSmemBytes = sizeof(QcSmem) + 1024;          // + slack for manual 1024 B alignment
static_assert(kQcSmemBytes <= 232448, "exceeds the 227 KiB of shared memory a CTA can opt into");

struct QconvParams {
    const float* h;            // (B, 256, HW) fp32, HW % 128 == 0, 16-byte aligned
    int64_t N, HW, n_pad;
    const __half* w_img;       // operand images of the weight: [hi | lo][chunk][256][64] (vq_qconv_weight_kernel)
    const float* w_scalars;    // [0] = 1 / s_W
    const float* bias;         // (256) or null
    float* z;                  // (B, 256, HW) fp32 out
    __half* z_h;               // operand image of the latents (as vq_prep_z_kernel writes it)
    float* z2;                 // (N)
    float* z_inv_scale;        // (N)
    int row_tiles;             // N / 128
    PrepClear clr;
};

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// W (256 output channels x 256 input channels, fp32 row-major: Conv2d(256, 256, 1).weight) -> hi / lo operand images + 1 / s_W.
// One CTA; runs once per weight update (64 K elements).
__global__ void __launch_bounds__(1024)
vq_qconv_weight_kernel(const float* __restrict__ W, __half* __restrict__ w_img, float* __restrict__ w_scalars) {
    __shared__ float red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float mx = 0.0f;
    for (int e = tid; e < kD * kD; e += 1024) mx = fmaxf(mx, fabsf(__ldg(W + e)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int ex = exponent_of(mx);
    const float sc = pow2f(kOperandTopExp - ex);
    if (tid == 0) w_scalars[0] = pow2f(ex - kOperandTopExp);
    for (int e = tid; e < kD * kD; e += 1024) {
        const int co = e / kD, ci = e % kD;
        const float x = __ldg(W + e) * sc;
        const __half hi = __float2half_rn(x);
        const __half lo = __float2half_rn(x - __half2float(hi));        // x - hi is exact in fp32
        const int64_t off = operand_image_offset(ci / kDChunk, kD, co, ci % kDChunk);
        w_img[off] = hi;
        w_img[(int64_t)kNumDChunks * kD * kDChunk + off] = lo;
    }
}

// kTensorMap: the activations arrive through `tmap` (2-D view [B * 256 channel planes][HW], box 64 x 128); otherwise -- no
// tensor-map encoder in the driver -- through one bulk copy per channel plane.
template <bool kTensorMap>
__global__ void __launch_bounds__(kQcThreads, 1)
vq_qconv_prep_kernel(const QconvParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ uint8_t smem_raw[];
    QcSmem& s = *reinterpret_cast<QcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();                                            // first kernel of the chain (vq_common.cuh): the GEMM may be scheduled

    {   // what vq_prep_z_kernel clears for the call: control words, histogram, counters; and the pad rows of the operand image
        const int64_t gtid = (int64_t)blockIdx.x * kQcThreads + threadIdx.x, gsz = (int64_t)gridDim.x * kQcThreads;
        for (int64_t i = gtid; i < p.clr.control_words; i += gsz) p.clr.control[i] = 0u;
        if (p.clr.hist != nullptr)
            for (int64_t i = gtid; i < p.clr.K; i += gsz) p.clr.hist[i] = 0ull;
        if (p.clr.stats != nullptr && gtid < p.clr.n_stats) p.clr.stats[gtid] = 0ull;
        // N is a multiple of the row tile, so the pad rows are whole row tiles at the end of the image: contiguous
        uint4* pad = reinterpret_cast<uint4*>(p.z_h + p.N * kD);
        const int64_t n_pad16 = (p.n_pad - p.N) * kD / 8;
        for (int64_t i = gtid; i < n_pad16; i += gsz) pad[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kQcHsSlots; i++) {
            mbar_init(&s.hs_full[i], 1);
            mbar_init(&s.hs_empty[i], 4);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.a_full[i], 4);
            mbar_init(&s.a_empty[i], 1);
            mbar_init(&s.b_full[i], 1);
            mbar_init(&s.b_empty[i], 1);
            mbar_init(&s.t_full[i], 1);
            mbar_init(&s.t_empty[i], 4);
            mbar_init(&s.hinv_full[i], 4);
            mbar_init(&s.hinv_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == kQcWarpMma) {
        tmem_alloc(&s.tmem_base, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < kD; i += kQcThreads) s.bias[i] = (p.bias != nullptr) ? __ldg(p.bias + i) : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;

    if (warp == kQcWarpH) {
        // ------------------------------------------------------------------ h producer
        const uint64_t pol_stream = policy_evict_first();     // second pass: the tile is not needed again
        uint32_t slot = 0, ph = 0;                            // ring position of the next stage and its phase
        for (int t = blockIdx.x; t < p.row_tiles; t += gridDim.x) {
            const int64_t n0 = (int64_t)t * kRowTile;
            const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
            const float* src0 = p.h + (b * kD) * p.HW + hw0;
            for (int pass = 0; pass < 2; pass++) {
                for (int dc = 0; dc < kNumDChunks; dc++) {
                    mbar_wait(&s.hs_empty[slot], ph ^ 1);
                    if (lane == 0) mbar_expect_tx(&s.hs_full[slot], kQcBytesHs);
                    __syncwarp();
                    if (kTensorMap) {
                        if (lane == 0)
                            tma_load_2d(&s.hs[slot][0][0], &tmap, (int32_t)hw0, (int32_t)(b * kD + kQcStageCh * dc), &s.hs_full[slot]);
                    } else {
#pragma unroll
                        for (int q = 0; q < kQcStageCh / 32; q++) {
                            const int ch = lane + 32 * q;
                            const float* src = src0 + (int64_t)(kQcStageCh * dc + ch) * p.HW;
                            if (pass == 0) bulk_load_1d(&s.hs[slot][ch][0], src, kRowTile * 4, &s.hs_full[slot]);
                            else bulk_load_1d_hint(&s.hs[slot][ch][0], src, kRowTile * 4, &s.hs_full[slot], pol_stream);
                        }
                    }
                    __syncwarp();
                    if (++slot == kQcHsSlots) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == kQcWarpW) {
        // ------------------------------------------------------------------ W producer
        const uint64_t pol_keep = policy_evict_last();        // every CTA re-reads the weight images for every tile
        const bool leader = elect_one();
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.row_tiles; t += gridDim.x) {
            for (int dc = 0; dc < kNumDChunks; dc++) {
                for (int part = 0; part < 2; part++, it++) {
                    const uint32_t slot = it & 1, ph = (it >> 1) & 1;
                    mbar_wait(&s.b_empty[slot], ph ^ 1);
                    if (leader) {
                        mbar_expect_tx(&s.b_full[slot], kQcBytesB);
                        bulk_load_1d_hint(s.b[slot], p.w_img + (int64_t)(part * kNumDChunks + dc) * (kD * kDChunk), kQcBytesB,
                                          &s.b_full[slot], pol_keep);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kQcWarpMma) {
        // ------------------------------------------------------------------ MMA issuer (one thread issues and commits everything)
        constexpr uint32_t idesc = umma_idesc_f16(kRowTile, kD);
        const bool leader = elect_one();
        uint32_t ita = 0, itb = 0, tile_i = 0;
        for (int t = blockIdx.x; t < p.row_tiles; t += gridDim.x, tile_i++) {
            const uint32_t buf = tile_i & 1, use = tile_i >> 1;
            mbar_wait(&s.t_empty[buf], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * kD;
            for (int dc = 0; dc < kNumDChunks; dc++, ita++) {
                const uint32_t as = ita & 1, aph = (ita >> 1) & 1;
                mbar_wait(&s.a_full[as], aph);
                tc_fence_after();
                const uint64_t ahi = umma_desc_sw128(smem_u32(s.a_hi[as]));
                const uint64_t alo = umma_desc_sw128(smem_u32(s.a_lo[as]));
                {   // stage W_hi[dc]: hi_h hi_W and lo_h hi_W
                    const uint32_t bs = itb & 1, bph = (itb >> 1) & 1;
                    mbar_wait(&s.b_full[bs], bph);
                    tc_fence_after();
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(s.b[bs]));
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kDChunk / 16; k++) umma_f16(d_tmem, ahi + 2 * k, bdesc + 2 * k, idesc, (dc | k) != 0);
#pragma unroll
                        for (int k = 0; k < kDChunk / 16; k++) umma_f16(d_tmem, alo + 2 * k, bdesc + 2 * k, idesc, 1u);
                        umma_commit(&s.b_empty[bs]);
                    }
                    __syncwarp();
                    itb++;
                }
                {   // stage W_lo[dc]: hi_h lo_W
                    const uint32_t bs = itb & 1, bph = (itb >> 1) & 1;
                    mbar_wait(&s.b_full[bs], bph);
                    tc_fence_after();
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(s.b[bs]));
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kDChunk / 16; k++) umma_f16(d_tmem, ahi + 2 * k, bdesc + 2 * k, idesc, 1u);
                        umma_commit(&s.b_empty[bs]);
                        umma_commit(&s.a_empty[as]);           // (a commit covers every MMA this thread has issued so far)
                        if (dc == kNumDChunks - 1) umma_commit(&s.t_full[buf]);
                    }
                    __syncwarp();
                    itb++;
                }
            }
        }
    } else if (warp >= kQcWarpConv0 && warp < kQcWarpEpi0) {
        // ------------------------------------------------------------------ converters: thread <-> latent row of the tile
        const int r = (warp - kQcWarpConv0) * 32 + lane;
        uint32_t slot = 0, ph = 0, ita = 0, tile_i = 0;
        const uint32_t hs_sa = smem_u32(&s.hs[0][0][r]);      // (explicit shared-memory accesses: the manually aligned base is a
        const uint32_t ahi_sa = smem_u32(s.a_hi[0]) + r * 128, alo_sa = smem_u32(s.a_lo[0]) + r * 128;   // generic pointer to the compiler)
        for (int t = blockIdx.x; t < p.row_tiles; t += gridDim.x, tile_i++) {
            // pass 0: max |h| of the row over all 256 channels (the operand scale must be known before the first conversion)
            float mx = 0.0f;
            for (int dc = 0; dc < kNumDChunks; dc++) {
                mbar_wait(&s.hs_full[slot], ph);
                const uint32_t col = hs_sa + slot * kQcBytesHs;
#pragma unroll 16
                for (int c = 0; c < kQcStageCh; c++) mx = fmaxf(mx, fabsf(lds_f32(col + c * (kRowTile * 4))));
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.hs_empty[slot]);
                if (++slot == kQcHsSlots) { slot = 0; ph ^= 1; }
            }
            const int ex = exponent_of(mx);
            const float sc = pow2f(kOperandTopExp - ex);
            {
                const uint32_t par = tile_i & 1, use = tile_i >> 1;
                mbar_wait(&s.hinv_empty[par], (use & 1) ^ 1);
                sts_f32(smem_u32(&s.hinv[par][r]), pow2f(ex - kOperandTopExp));
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.hinv_full[par]);
            }
            // pass 1: hi / lo fp16 operand rows of each 64-channel chunk
            for (int dc = 0; dc < kNumDChunks; dc++, ita++) {
                const uint32_t as = ita & 1, aph = (ita >> 1) & 1;
                mbar_wait(&s.hs_full[slot], ph);
                mbar_wait(&s.a_empty[as], aph ^ 1);
                const uint32_t col = hs_sa + slot * kQcBytesHs;
                const uint32_t hi_row = ahi_sa + as * kQcBytesA, lo_row = alo_sa + as * kQcBytesA;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    uint32_t hv[4], lv[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float x0 = lds_f32(col + (8 * c8 + 2 * j) * (kRowTile * 4)) * sc;
                        const float x1 = lds_f32(col + (8 * c8 + 2 * j + 1) * (kRowTile * 4)) * sc;
                        const __half2 hh = __floats2half2_rn(x0, x1);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);      // exact differences
                        hv[j] = *reinterpret_cast<const uint32_t*>(&hh);
                        lv[j] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    const uint32_t piece = (uint32_t)((c8 ^ (r & 7)) << 4);              // SWIZZLE_128B: 16-byte piece ^ (row & 7)
                    sts128(hi_row + piece, hv[0], hv[1], hv[2], hv[3]);
                    sts128(lo_row + piece, lv[0], lv[1], lv[2], lv[3]);
                }
                fence_proxy_async_smem();                      // generic-proxy stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&s.a_full[as]);
                    mbar_arrive(&s.hs_empty[slot]);
                }
                if (++slot == kQcHsSlots) { slot = 0; ph ^= 1; }
            }
        }
    } else if (warp >= kQcWarpEpi0) {
        // ------------------------------------------------------------------ epilogue: thread <-> TMEM lane <-> latent row
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const float w_inv = __ldg(p.w_scalars);
        const uint32_t bias_sa = smem_u32(&s.bias[0]);
        auto bias_at = [&](int ch) -> float { return lds_f32(bias_sa + ch * 4); };
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < p.row_tiles; t += gridDim.x, tile_i++) {
            const uint32_t buf = tile_i & 1, use = tile_i >> 1;
            const int64_t n0 = (int64_t)t * kRowTile;
            const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
            mbar_wait(&s.hinv_full[buf], use & 1);
            const float scale = lds_f32(smem_u32(&s.hinv[buf][r])) * w_inv;       // 1 / (s_h s_W): exact (powers of two)
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.hinv_empty[buf]);
            mbar_wait(&s.t_full[buf], use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * kD;
            float* zp = p.z + (b * kD) * p.HW + hw0 + r;
            // pass A: z = fl(acc * scale + bias) stored NCHW; |z|^2 in the canonical order (partial j over d == j (mod 4),
            // ascending, one fma each -- vq_prep_z_kernel's chains); row maximum
            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, mx = 0.0f;
            auto pass_a = [&](const uint32_t (&acc)[32], int c) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int ch = 32 * c + i;
                    const float z0 = __fmaf_rn(__uint_as_float(acc[i + 0]), scale, bias_at(ch + 0));
                    const float z1 = __fmaf_rn(__uint_as_float(acc[i + 1]), scale, bias_at(ch + 1));
                    const float z2v = __fmaf_rn(__uint_as_float(acc[i + 2]), scale, bias_at(ch + 2));
                    const float z3 = __fmaf_rn(__uint_as_float(acc[i + 3]), scale, bias_at(ch + 3));
                    zp[(int64_t)(ch + 0) * p.HW] = z0;
                    zp[(int64_t)(ch + 1) * p.HW] = z1;
                    zp[(int64_t)(ch + 2) * p.HW] = z2v;
                    zp[(int64_t)(ch + 3) * p.HW] = z3;
                    p0 = __fmaf_rn(z0, z0, p0); p1 = __fmaf_rn(z1, z1, p1);
                    p2 = __fmaf_rn(z2v, z2v, p2); p3 = __fmaf_rn(z3, z3, p3);
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(z0), fabsf(z1)), fmaxf(fabsf(z2v), fabsf(z3))));
                }
            };
            // pass B: the row of the distance GEMM's operand image ([row tile][chunk][128][64] fp16, SWIZZLE_128B), recomputed
            // from the accumulator with the same fma -> the same z, bit for bit
            __half* img = p.z_h + ((int64_t)t * kNumDChunks) * (kRowTile * kDChunk) + r * kDChunk;
            float sc = 0.0f;
            auto pass_b = [&](const uint32_t (&acc)[32], int c) {
                __half* chunk = img + (int64_t)(c >> 1) * (kRowTile * kDChunk);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int i = 8 * q + 2 * j, ch = 32 * c + i;
                        const float za = __fmaf_rn(__uint_as_float(acc[i]), scale, bias_at(ch));
                        const float zb = __fmaf_rn(__uint_as_float(acc[i + 1]), scale, bias_at(ch + 1));
                        const __half2 hh = __floats2half2_rn(za * sc, zb * sc);
                        pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    const int piece = (c & 1) * 4 + q;
                    *reinterpret_cast<uint4*>(chunk + ((piece ^ (r & 7)) << 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            };
            auto finish_a = [&]() {
                const float zz = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
                const int ex = exponent_of(mx);
                p.z2[n0 + r] = zz;
                p.z_inv_scale[n0 + r] = pow2f(ex - kOperandTopExp);
                sc = pow2f(kOperandTopExp - ex);
            };
            auto release = [&]() {                             // the accumulator is in registers for the last time
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.t_empty[buf]);
            };
#pragma unroll 1
            for (int c = 0; c < kD / 32; c++) {
                uint32_t acc[32];
                tmem_ld32(taddr + 32 * c, acc);
                tmem_ld_wait();
                pass_a(acc, c);
            }
            finish_a();
#pragma unroll 1
            for (int c = 0; c < kD / 32; c++) {
                uint32_t acc[32];
                tmem_ld32(taddr + 32 * c, acc);
                tmem_ld_wait();
                if (c == kD / 32 - 1) release();
                pass_b(acc, c);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kQcWarpMma) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace vq
