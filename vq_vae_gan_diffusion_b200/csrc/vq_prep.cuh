// vq_prep.cuh -- operand preparation for the distance GEMM.
//   vq_prep_z_kernel           z fp32 NCHW (B, D, HW) -> z_h (N_pad, D) fp16 rows scaled by 2^(15-ex_n),
//                              |z_n|^2 (canonical order), per-row inverse scale
//   vq_codebook_norms_kernel   E fp32 (K, D) -> |e_k|^2 (+inf on pad rows), max_k |e_k|^2, max |E|
//   vq_codebook_convert_kernel E fp32 -> E_h (K_pad, D) fp16 scaled by 2^(15-ex_E), inverse scale
// All are single-pass, HBM-bound: every input byte is read once with 128-byte coalesced requests and every
// output byte written once.  (reference: codebook.py:62-66 permute+contiguous+view; codebook.py:71-74 the two
// torch.sum(x**2) calls; the operand conversion has no reference counterpart -- the reference runs sgemm.)
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kPrepThreads = 256;

// cb_scalars layout (4 floats, device)
constexpr int kCbE2Max = 0;      // max_k |e_k|^2
constexpr int kCbMaxAbs = 1;     // max |E[k][d]|
constexpr int kCbInvScale = 2;   // 2^(ex_E - 15): inverse of the fp16 operand scale

// One CTA = 32 consecutive latents, tile in shared memory row-major with swizzled 16-byte pieces (vq_common.cuh
// tile_off): column-form fill (16-byte loads along hw), then everything per latent row -- warp w owns rows 4w..4w+3.
// Per-call scratch that later kernels of the same call accumulate into; cleared here (the first kernel of every call)
// instead of by separate memset nodes: the control words of the workspace, the usage histogram and the counters.
struct PrepClear {
    unsigned int* control;          // blocks_done | fb_count | fb_arrive (contiguous)
    int control_words;
    unsigned long long* hist;       // (K) or null
    int K;
    unsigned long long* stats;      // (VQ_STAT_COUNT) or null
    int n_stats;
};

template <int kLayout>
__global__ void __launch_bounds__(kPrepThreads)
vq_prep_z_kernel(const float* __restrict__ z, int64_t N, int64_t HW, int64_t n_pad,
                 __half* __restrict__ z_h, float* __restrict__ z2, float* __restrict__ z_inv_scale, const PrepClear clr) {
    __shared__ __align__(16) float tile[kSelRows * kD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    pdl_trigger();                                            // first kernel of the chain (vq_common.cuh): the GEMM may be scheduled

    {   // grid-strided clears (a few KiB in total)
        const int64_t gtid = (int64_t)blockIdx.x * kPrepThreads + tid, gsz = (int64_t)gridDim.x * kPrepThreads;
        for (int64_t i = gtid; i < clr.control_words; i += gsz) clr.control[i] = 0u;
        if (clr.hist != nullptr)
            for (int64_t i = gtid; i < clr.K; i += gsz) clr.hist[i] = 0ull;
        if (clr.stats != nullptr && gtid < clr.n_stats) clr.stats[gtid] = 0ull;
    }

    if (kLayout != kLayoutGeneric && n0 >= N) {                                   // pad rows of the last GEMM row tile: zero operand rows
        // (a 32-row slab of a row tile is not contiguous in the operand image: zero it piecewise)
        for (int i = tid; i < kSelRows * kD / 2; i += kPrepThreads) {
            const int r = i / (kD / 2), d = 2 * (i % (kD / 2));
            const int64_t n = n0 + r;
            if (n < n_pad)
                *reinterpret_cast<__half2*>(z_h + operand_image_offset((n / kRowTile) * kNumDChunks + d / kDChunk, kRowTile,
                                                                       (int)(n % kRowTile), d % kDChunk)) = __floats2half2_rn(0.f, 0.f);
        }
        return;
    }
    if (kLayout == kLayoutRows) {
        fill_tile_rows(tile, z, n0, N, warp, lane);
    } else if (kLayout == kLayoutVec) {
        const ColForm cf(warp, lane);
        const float* src = z + col_form_origin(n0, HW, warp, lane);
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(4 * i) * HW));
#pragma unroll
        for (int i = 0; i < 8; i++) cf.store(tile, i, v[i]);
    } else {
        const int64_t n = n0 + lane;
        const bool ok = n < N;
        const int64_t b = ok ? n / HW : 0, hw = ok ? n % HW : 0;
        const float* src = z + (b * kD) * HW + hw;
#pragma unroll 8
        for (int i = 0; i < kD / 8; i++) {
            const int d = warp + 8 * i;
            tile[tile_off(lane, d)] = ok ? __ldg(src + (int64_t)d * HW) : 0.0f;      // rows >= N: zero operand rows
        }
    }
    __syncthreads();

    // |z|^2 in canonical order and max|z|: lanes 0..15 of warp w = (row 4w + (lane >> 2), partial j = lane & 3), each
    // a chain of 64 fma over d == j (mod 4) ascending.  (4 rows x 4 d-offsets of one piece: conflict-free.)
    float sc_row;                                            // operand scale of row 4w + (lane >> 3) ... see below
    {
        const int rr = (lane >> 2) & 3, j = lane & 3;
        const int r = warp * 4 + rr;
        const float* zrow = tile + r * kD + j;
        const int g = tile_swz(r);
        int zo[8];
#pragma unroll
        for (int b = 0; b < 8; b++) zo[b] = (b ^ g) << 2;
        float p = 0.0f, mx = 0.0f;
        if (lane < 16) {
#pragma unroll
            for (int q = 0; q < kD / 4; q++) {
                const float v = zrow[32 * (q >> 3) + zo[q & 7]];
                p = __fmaf_rn(v, v, p);
                mx = fmaxf(mx, fabsf(v));
            }
        }
        const float s = combine4(p);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const int ex = exponent_of(mx);
        if (lane < 16 && j == 0 && n0 + r < N) { z2[n0 + r] = s; z_inv_scale[n0 + r] = pow2f(ex - kOperandTopExp); }
        sc_row = pow2f(kOperandTopExp - ex);                 // valid on lanes 0..15: scale of row 4w + (lane >> 2)
    }

    // fp16 operand rows: lane owns d in [4 lane, 4 lane + 4) and [128 + 4 lane, ...): two 16-byte reads and two 8-byte
    // stores per row; lanes 0..15 / 16..31 of a store cover one full 128-byte row of two D chunks of the operand image
    // ([row tile][64-wide D chunk][128 rows][128 B], 16-byte pieces XOR-swizzled by (row & 7): the exact shared-memory
    // image of a SWIZZLE_128B K-major UMMA operand, so the GEMM loads a chunk with ONE contiguous bulk copy)
    // (lane part of the operand-image offset: D chunk (lane >> 4) + 2 h, piece (lane & 15) >> 1, low half 4 (lane & 1))
    const int img_lane = (lane >> 4) * (kRowTile * kDChunk) + 4 * (lane & 1);
    const int img_piece = (lane & 15) >> 1;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        const float sc = __shfl_sync(0xffffffffu, sc_row, 4 * rr);
        if (n >= n_pad) continue;                            // warp-uniform
        const float4* zrow4 = reinterpret_cast<const float4*>(tile + r * kD);
        const int g = tile_swz(r);
        const int rr_t = (int)((uint32_t)n % kRowTile);
        __half* row_img = z_h + ((int64_t)((uint32_t)n / kRowTile) * kNumDChunks) * (kRowTile * kDChunk) + rr_t * kDChunk + img_lane +
                          ((img_piece ^ (rr_t & 7)) << 3);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float4 v = zrow4[(lane + 32 * h) ^ g];     // d = 4 (lane + 32 h) .. + 3
            __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(row_img + h * 2 * (kRowTile * kDChunk)) = pk;
        }
    }
}

// One CTA = 32 codes.  cb[kCbE2Max], cb[kCbMaxAbs] must be zero on entry.
__global__ void __launch_bounds__(kPrepThreads)
vq_codebook_norms_kernel(const float* __restrict__ E, int K, int k_pad, float* __restrict__ e2, float* __restrict__ cb) {
    __shared__ float t[kSelRows][kD + 1];
    __shared__ float smax[4], sabs[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * kSelRows;

#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr, k = k0 + r;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            t[r][d] = (k < K) ? __ldg(E + (int64_t)k * kD + d) : 0.0f;
        }
    }
    __syncthreads();

    if (tid < 4 * kSelRows) {
        const int r = tid >> 2, j = tid & 3, k = k0 + r;
        float p = 0.0f, mx = 0.0f;
#pragma unroll 16
        for (int q = 0; q < kD / 4; q++) {
            const float v = t[r][4 * q + j];
            p = __fmaf_rn(v, v, p);
            mx = fmaxf(mx, fabsf(v));
        }
        const float s = combine4(p);
        if (j == 0 && k < k_pad) e2[k] = (k < K) ? s : INFINITY;
        // maxima over real rows -> one atomic per CTA (values are >= 0, so uint ordering == float ordering;
        // a NaN has a larger bit pattern than every finite value and therefore propagates)
        float m = (k < K) ? s : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), x2 = __shfl_xor_sync(0xffffffffu, mx, o);
            m = __uint_as_float(max(__float_as_uint(m), __float_as_uint(m2)));
            mx = __uint_as_float(max(__float_as_uint(mx), __float_as_uint(x2)));
        }
        if (lane == 0) { smax[warp] = m; sabs[warp] = mx; }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned a = 0, b = 0;
        for (int w = 0; w < 4; w++) { a = max(a, __float_as_uint(smax[w])); b = max(b, __float_as_uint(sabs[w])); }
        atomicMax(reinterpret_cast<unsigned int*>(cb + kCbE2Max), a);
        atomicMax(reinterpret_cast<unsigned int*>(cb + kCbMaxAbs), b);
    }
}

// One CTA = 32 codes; pure streaming conversion (E is L2 resident after the norms pass).
__global__ void __launch_bounds__(kPrepThreads)
vq_codebook_convert_kernel(const float* __restrict__ E, int K, int k_pad, __half* __restrict__ e_h,
                           float* __restrict__ cb) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * kSelRows;
    const int ex = exponent_of(cb[kCbMaxAbs]);
    const float sc = pow2f(kOperandTopExp - ex);
    if (blockIdx.x == 0 && tid == 0) cb[kCbInvScale] = pow2f(ex - kOperandTopExp);
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int k = k0 + warp * 4 + rr;
        if (k >= k_pad) break;
        const float2* src = reinterpret_cast<const float2*>(E + (int64_t)k * kD);
#pragma unroll
        for (int i = 0; i < kNumDChunks; i++) {
            const int d2 = lane + 32 * i;                       // float2 index: d = 2*d2 = 64*i + 2*lane
            float2 v = make_float2(0.0f, 0.0f);
            if (k < K) v = __ldg(src + d2);
            // operand image: [code tile][64-wide D chunk][256 codes][128 B] with the SWIZZLE_128B pattern (see prep_z)
            __half2* dst = reinterpret_cast<__half2*>(
                e_h + operand_image_offset((int64_t)(k / kCodeTile) * kNumDChunks + i, kCodeTile, k % kCodeTile, 2 * lane));
            *dst = __floats2half2_rn(v.x * sc, v.y * sc);
        }
    }
}

// =====================================================================================================================
// Row-major vectors of any width D <= 512 (the nearest-code searches of SURVEY.md 8(f) n2: gaussian_dim = 96 in
// configs/*.yml, 512 in diffusion_gaussian3d.py's own runs).  The contraction is padded to kNC chunks of 64 inside the
// kernels -- a zero column adds fma(0, 0, p) == p to every canonical chain, so no norm, dot product or tie changes -- and
// nothing is padded in HBM: rows are read at their own pitch D.
// =====================================================================================================================
template <int kNC>
struct RowsCfg {
    static constexpr int kDp = 64 * kNC;                     // padded width
    static constexpr int kRows = (kNC == 8) ? 16 : 32;       // rows per CTA: the fp32 tile stays below 34 KiB of static shared memory
    static constexpr int kRowsPerWarp = kRows / (kPrepThreads / 32);
    static constexpr int kPitch = kDp + 4;                   // floats per tile row: consecutive rows start 4 banks apart, so the
                                                             // (row, j) chains of a warp read distinct banks
    static constexpr int kPieces = kDp / 4;                  // float4 pieces per row
};

// F.normalize(x, p=2, dim=-1) as diffusion_gaussian3d.py:560-563 applies it: x / max(|x|_2, 1e-12), with |x|_2^2 in the
// canonical order.  (clamp_min keeps a NaN norm, fmaxf would drop it.)
__device__ __forceinline__ float normalize_denom(float norm2) {
    const float nrm = sqrtf(norm2);
    return (nrm < 1e-12f) ? 1e-12f : nrm;
}

// One CTA = RowsCfg::kRows consecutive rows.  Outputs (each optional): the fp16 operand image + |x|^2 + inverse operand
// scale for the distance GEMM; with kNormalize the rows are L2-normalised first and the divisor / the normalised fp32 rows
// can be kept (`denom`, `x_hat`: the exact stage divides by the same number; the lookup table is normalised once this way).
template <int kNC, bool kNormalize>
__global__ void __launch_bounds__(kPrepThreads)
vq_prep_rows_kernel(const float* __restrict__ x, int64_t N, int D, int64_t n_pad, __half* __restrict__ z_h,
                    float* __restrict__ z2, float* __restrict__ z_inv_scale, float* __restrict__ denom,
                    float* __restrict__ x_hat, const PrepClear clr) {
    using C = RowsCfg<kNC>;
    __shared__ __align__(16) float tile[C::kRows * C::kPitch];
    __shared__ float denom_s[C::kRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * C::kRows;
    pdl_trigger();

    {   // grid-strided clears (a few KiB in total)
        const int64_t gtid = (int64_t)blockIdx.x * kPrepThreads + tid, gsz = (int64_t)gridDim.x * kPrepThreads;
        for (int64_t i = gtid; i < clr.control_words; i += gsz) clr.control[i] = 0u;
        if (clr.hist != nullptr)
            for (int64_t i = gtid; i < clr.K; i += gsz) clr.hist[i] = 0ull;
        if (clr.stats != nullptr && gtid < clr.n_stats) clr.stats[gtid] = 0ull;
    }

    // fill: warp w owns rows kRowsPerWarp * w ..; 16-byte loads when the rows are 16-byte aligned, columns >= D and rows >= N zero
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
#pragma unroll
    for (int rr = 0; rr < C::kRowsPerWarp; rr++) {
        const int r = warp * C::kRowsPerWarp + rr;
        const int64_t n = n0 + r;
        float* trow = tile + r * C::kPitch;
        if (vec) {
            const float4* src = reinterpret_cast<const float4*>(x + (n < N ? n : 0) * (int64_t)D);
#pragma unroll
            for (int h = 0; h < (C::kPieces + 31) / 32; h++) {
                const int pi = lane + 32 * h;
                if (pi < C::kPieces)
                    *reinterpret_cast<float4*>(trow + 4 * pi) = (n < N && 4 * pi < D) ? __ldg(src + pi) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const float* src = x + (n < N ? n : 0) * (int64_t)D;
            for (int d = lane; d < C::kDp; d += 32) trow[d] = (n < N && d < D) ? __ldg(src + d) : 0.0f;
        }
    }
    __syncthreads();

    // canonical |x|^2 and max |x|: lane (row rr = lane >> 2, partial j = lane & 3) runs one chain of kDp / 4 fma
    auto chains = [&](float& s, float& mx) {
        const int rr = (lane >> 2) % C::kRowsPerWarp, j = lane & 3;
        const float* trow = tile + (warp * C::kRowsPerWarp + rr) * C::kPitch + j;
        float p = 0.0f;
        mx = 0.0f;
        if (lane < 4 * C::kRowsPerWarp) {
#pragma unroll 16
            for (int q = 0; q < C::kPieces; q++) {
                const float v = trow[4 * q];
                p = __fmaf_rn(v, v, p);
                mx = fmaxf(mx, fabsf(v));
            }
        }
        s = combine4(p);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    };
    float s, mx;
    chains(s, mx);
    if (kNormalize) {
        if (lane < 4 * C::kRowsPerWarp && (lane & 3) == 0) denom_s[warp * C::kRowsPerWarp + (lane >> 2)] = normalize_denom(s);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < C::kRowsPerWarp; rr++) {
            const int r = warp * C::kRowsPerWarp + rr;
            const int64_t n = n0 + r;
            const float dn = denom_s[r];
            float* trow = tile + r * C::kPitch;
            for (int d = lane; d < C::kDp; d += 32) {
                const float v = __fdiv_rn(trow[d], dn);            // pad columns: 0 / dn == 0
                trow[d] = v;
                if (x_hat != nullptr && n < N && d < D) x_hat[n * (int64_t)D + d] = v;
            }
            if (denom != nullptr && lane == 0 && n < N) denom[n] = dn;
        }
        __syncwarp();
        chains(s, mx);                                             // norms / operand scale of the normalised rows
    }
    const int ex = exponent_of(mx);
    {
        const int r = warp * C::kRowsPerWarp + (lane >> 2);
        if (lane < 4 * C::kRowsPerWarp && (lane & 3) == 0 && n0 + r < N) {
            if (z2 != nullptr) z2[n0 + r] = s;
            if (z_inv_scale != nullptr) z_inv_scale[n0 + r] = pow2f(ex - kOperandTopExp);
        }
    }
    if (z_h == nullptr) return;
    const float sc_row = pow2f(kOperandTopExp - ex);               // valid on lanes < 4 kRowsPerWarp: scale of row (lane >> 2)

    // fp16 operand rows into the image [row tile][64-wide D chunk][128 rows][128 B] (16-byte pieces XOR-swizzled by row & 7)
#pragma unroll
    for (int rr = 0; rr < C::kRowsPerWarp; rr++) {
        const int r = warp * C::kRowsPerWarp + rr;
        const int64_t n = n0 + r;
        const float sc = __shfl_sync(0xffffffffu, sc_row, 4 * rr);
        if (n >= n_pad) continue;                                  // warp-uniform
        const float* trow = tile + r * C::kPitch;
        const int rr_t = (int)((uint32_t)n % kRowTile);
        __half* tile_img = z_h + ((int64_t)((uint32_t)n / kRowTile) * kNC) * (kRowTile * kDChunk) + rr_t * kDChunk;
#pragma unroll
        for (int h = 0; h < (C::kPieces + 31) / 32; h++) {
            const int pi = lane + 32 * h;                          // float4 piece: d = 4 pi .. 4 pi + 3
            if (pi >= C::kPieces) continue;
            const float4 v = *reinterpret_cast<const float4*>(trow + 4 * pi);
            __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            const int c = pi >> 4, dl = 4 * (pi & 15);             // D chunk, first element inside the chunk
            *reinterpret_cast<uint2*>(tile_img + (int64_t)c * (kRowTile * kDChunk) + ((((dl >> 3) ^ (rr_t & 7)) << 3) | (dl & 7))) = pk;
        }
    }
}

// Codebook / lookup-table preparation at width D <= 64 kNC (row pitch D).  One CTA = 32 codes.  cb[kCbE2Max], cb[kCbMaxAbs]
// must be zero on entry.
template <int kNC>
__global__ void __launch_bounds__(kPrepThreads)
vq_table_norms_kernel(const float* __restrict__ E, int K, int D, int k_pad, float* __restrict__ e2, float* __restrict__ cb) {
    constexpr int kDp = 64 * kNC;
    constexpr int kRowsT = (kNC == 8) ? 16 : 32;                  // codes per pass over the shared tile
    __shared__ float t[kRowsT][kDp + 1];
    __shared__ float smax[4], sabs[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float m_all = 0.0f, mx_all = 0.0f;
    for (int pass = 0; pass < kSelRows / kRowsT; pass++) {
        const int k0 = blockIdx.x * kSelRows + pass * kRowsT;
        __syncthreads();
        for (int rr = 0; rr < kRowsT / 8; rr++) {
            const int r = warp * (kRowsT / 8) + rr, k = k0 + r;
            for (int d = lane; d < kDp; d += 32) t[r][d] = (k < K && d < D) ? __ldg(E + (int64_t)k * D + d) : 0.0f;
        }
        __syncthreads();
        if (tid < 4 * kRowsT) {
            const int r = tid >> 2, j = tid & 3, k = k0 + r;
            float p = 0.0f, mx = 0.0f;
#pragma unroll 16
            for (int q = 0; q < kDp / 4; q++) {
                const float v = t[r][4 * q + j];
                p = __fmaf_rn(v, v, p);
                mx = fmaxf(mx, fabsf(v));
            }
            const float s = combine4(p);
            if (j == 0 && k < k_pad) e2[k] = (k < K) ? s : INFINITY;
            const float m = (k < K) ? s : 0.0f;
            m_all = __uint_as_float(max(__float_as_uint(m_all), __float_as_uint(m)));
            mx_all = __uint_as_float(max(__float_as_uint(mx_all), __float_as_uint(mx)));
        }
    }
    // maxima over real rows -> one atomic per CTA (values are >= 0, so uint ordering == float ordering; a NaN has a larger
    // bit pattern than every finite value and therefore propagates)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m_all, o), x2 = __shfl_xor_sync(0xffffffffu, mx_all, o);
        m_all = __uint_as_float(max(__float_as_uint(m_all), __float_as_uint(m2)));
        mx_all = __uint_as_float(max(__float_as_uint(mx_all), __float_as_uint(x2)));
    }
    if (lane == 0 && warp < 4) { smax[warp] = m_all; sabs[warp] = mx_all; }
    __syncthreads();
    if (tid == 0) {
        unsigned a = 0, b = 0;
        for (int w = 0; w < 4; w++) { a = max(a, __float_as_uint(smax[w])); b = max(b, __float_as_uint(sabs[w])); }
        atomicMax(reinterpret_cast<unsigned int*>(cb + kCbE2Max), a);
        atomicMax(reinterpret_cast<unsigned int*>(cb + kCbMaxAbs), b);
    }
}

// One CTA = 32 codes; E fp32 (K, D) -> operand image [code tile][kNC chunks][256 codes][128 B], pad rows / columns zero.
template <int kNC>
__global__ void __launch_bounds__(kPrepThreads)
vq_table_convert_kernel(const float* __restrict__ E, int K, int D, int k_pad, __half* __restrict__ e_h, float* __restrict__ cb) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * kSelRows;
    const int ex = exponent_of(cb[kCbMaxAbs]);
    const float sc = pow2f(kOperandTopExp - ex);
    if (blockIdx.x == 0 && tid == 0) cb[kCbInvScale] = pow2f(ex - kOperandTopExp);
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int k = k0 + warp * 4 + rr;
        if (k >= k_pad) break;
        const float* src = E + (int64_t)k * D;
#pragma unroll
        for (int i = 0; i < kNC; i++) {
            const int d = 64 * i + 2 * lane;
            const float v0 = (k < K && d < D) ? __ldg(src + d) : 0.0f;
            const float v1 = (k < K && d + 1 < D) ? __ldg(src + d + 1) : 0.0f;
            __half2* dst = reinterpret_cast<__half2*>(
                e_h + operand_image_offset((int64_t)(k / kCodeTile) * kNC + i, kCodeTile, k % kCodeTile, 2 * lane));
            *dst = __floats2half2_rn(v0 * sc, v1 * sc);
        }
    }
}

}  // namespace vq
