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

// One CTA = 32 consecutive latents.  Tile held in shared memory as t[d][row] (+1 pad).
template <bool kVec>
__global__ void __launch_bounds__(kPrepThreads)
vq_prep_z_kernel(const float* __restrict__ z, int64_t N, int64_t HW, int64_t n_pad,
                 __half* __restrict__ z_h, float* __restrict__ z2, float* __restrict__ z_inv_scale) {
    __shared__ TileRow t[kD];
    __shared__ float scale_s[kSelRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;

    if (kVec && n0 >= N) {                                   // pad rows of the last GEMM row tile: zero operand rows
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
    load_tile_nchw<kVec, false>(t, z, n0, N, HW, warp, lane);
    __syncthreads();

    // |z|^2 in canonical order and max|z|: 4 threads per row, thread j owns the terms d == j (mod 4)
    if (tid < 4 * kSelRows) {
        const int r = tid >> 2, j = tid & 3;
        float p = 0.0f, mx = 0.0f;
#pragma unroll 16
        for (int q = 0; q < kD / 4; q++) {
            const float v = t[4 * q + j][r];
            p = __fmaf_rn(v, v, p);
            mx = fmaxf(mx, fabsf(v));
        }
        const float s = combine4(p);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        if (j == 0) {
            const int ex = exponent_of(mx);
            scale_s[r] = pow2f(kOperandTopExp - ex);
            if (n0 + r < N) { z2[n0 + r] = s; z_inv_scale[n0 + r] = pow2f(ex - kOperandTopExp); }
        }
    }
    __syncthreads();

    // fp16 rows: warp w writes rows 4w..4w+3, lane covers d = 2*lane + 64*i (4-byte stores, 128 B per warp request)
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        if (n >= n_pad) break;
        const float sc = scale_s[r];
        // operand image: [row tile][64-wide D chunk][128 rows][128 B], 16-byte pieces XOR-swizzled by (row & 7) -- the
        // exact shared-memory image of a SWIZZLE_128B K-major UMMA operand, so a chunk is ONE contiguous bulk copy
        const int64_t rt = n / kRowTile;
        const int rr_t = (int)(n % kRowTile);
#pragma unroll
        for (int i = 0; i < kNumDChunks; i++) {
            const int d = 2 * lane + 64 * i;
            __half2* dst = reinterpret_cast<__half2*>(z_h + operand_image_offset(rt * kNumDChunks + i, kRowTile, rr_t, 2 * lane));
            *dst = __floats2half2_rn(t[d][r] * sc, t[d + 1][r] * sc);   // rows >= N were zero-filled above
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

}  // namespace vq
