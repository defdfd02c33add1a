// vq_backward.cuh -- fused straight-through backward + codebook scatter-add, and the NCHW embedding lookup.
//
// vq_backward_kernel: one CTA = 32 latents, ONE [32 x 256] fp32 tile in shared memory (row-major, 16-byte pieces
// XOR-swizzled: vq_common.cuh tile_off), two transposes in total:
//   1. z tile (NCHW: 16-byte loads along hw, 4 latents x 1 d per request) -> shared memory, column form -> row form
//   2. row form, warp w owns latents 4w..4w+3, lane owns d in [4 lane, 4 lane + 4) and [128 + 4 lane, ...), everything in
//      16-byte requests: e = E[idx[n]] (L2 resident), diff = z - e, grad_E[idx[n]] += -coef * beta * diff
//      (red.global.add.v4.f32), grad = g_out + coef * diff with a channels-last g_out (the layout of the z_q we
//      returned) read straight from global memory in row form; grad (or diff) goes back into the tile
//   3. row form -> column form: grad_z written as NCHW with 16-byte stores; an hw-contiguous g_out (e.g. NCHW) is
//      added here, read in column form like z                                    [autograd of codebook.py:96-106]
// HBM traffic per latent: read g_out 4D + z 4D + idx 8, write grad_z 4D -- the algorithmic minimum; the codebook
// and its gradient (16 MiB each at K = 16384) stay in the 126 MB L2.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kBwdThreads = 256;
#ifndef VQ_BWD_MIN_BLOCKS
#define VQ_BWD_MIN_BLOCKS 3               // resident CTAs per SM the register allocation is tuned for
#endif

struct BackwardParams {
    const float* gout;        // may be null
    int64_t gs_b, gs_d, gs_hw;   // element strides of gout's logical (B, D, HW)
    const float* z;           // (B, D, HW)
    const int64_t* idx;       // (N)
    const float* E;           // (K, D)
    int64_t N, HW;
    int K;
    float g_loss;             // upstream gradient on the loss (host value) ...
    const float* g_loss_dev;  // ... or, when non-null, a device scalar holding it (no host sync in autograd)
    double inv_nd;            // 1 / (n_global * D)
    float beta;
    float e_scale;            // extra factor on the codebook gradient (1 / world size when the ranks' gradients are summed)
    float* grad_z;            // (B, D, HW) or null
    float* grad_E;            // (K, D) or null, zeroed before launch
    // deterministic mode (kDet): the scatter-add runs in 64-bit fixed point -- integer addition is associative, so the
    // result does not depend on the order in which the atomics land -- into acc_fx, scaled by 2^fx_shift[0]
    long long* acc_fx;        // (K, D) zeroed before launch
    const int* fx_shift;      // (1) device scalar written by vq_backward_maxdiff_kernel
};


__device__ __forceinline__ void red_add_s64(long long* addr, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// Fixed-point scale of the deterministic scatter-add: |z - e| < 2^ex for every element (ex from the pre-pass below), so
// with shift = 38 - ex a term is below 2^38 in magnitude and up to 2^24 of them (more latents than the API accepts per
// code) stay inside 63 bits; the quantum 2^-shift is 2^-38 of the largest difference, far below fp32 resolution.
constexpr int kFxTopBit = 38;

// kVec: HW % 32 == 0 and 16-byte aligned z / grad_z (the 32 latents of a tile are 32 consecutive hw positions of one
// batch item).  kGoutCL: g_out is channels-last (gs_d == 1) with 16-byte aligned rows.  kDet: deterministic scatter-add.
template <bool kVec, bool kGoutCL, bool kDet = false>
__global__ void __launch_bounds__(kBwdThreads, VQ_BWD_MIN_BLOCKS)
vq_backward_kernel(const BackwardParams p) {
    __shared__ __align__(16) float tile[kSelRows * kD];       // 32 KiB: z, then grad (kGoutCL) or diff
    __shared__ int idx_s[kSelRows];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dsub = lane >> 3, hq = lane & 7;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    // coef = 2 * g_loss / (n_global * D), evaluated like the oracle (double, rounded once)
    const float coef = (float)(2.0 * (double)(p.g_loss_dev != nullptr ? __ldg(p.g_loss_dev) : p.g_loss) * p.inv_nd);
    const bool has_g = p.gout != nullptr && p.grad_z != nullptr;

    if (tid < kSelRows) {
        const int64_t n = n0 + tid;
        int k = -1;
        if (n < p.N) {
            const int64_t kk = __ldg(p.idx + n);
            k = (kk < 0 || kk >= p.K) ? -1 : (int)kk;       // out-of-range index: contributes nothing
        }
        idx_s[tid] = k;
    }

    // 1. z tile, column form -> shared memory; the channels-last upstream gradient (row form) is requested right
    //    behind it so that both streams are in flight together
    if (kVec) {
        const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
        const float* src = p.z + (b * kD + dsub) * p.HW + hw0 + 4 * hq;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)((warp * 8 + i) * 4) * p.HW));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int d = (warp * 8 + i) * 4 + dsub;
            tile[tile_off(4 * hq + 0, d)] = v[i].x;
            tile[tile_off(4 * hq + 1, d)] = v[i].y;
            tile[tile_off(4 * hq + 2, d)] = v[i].z;
            tile[tile_off(4 * hq + 3, d)] = v[i].w;
        }
    } else {
        const int64_t n = n0 + lane;
        const bool ok = n < p.N;
        const int64_t b = ok ? n / p.HW : 0, hw = ok ? n % p.HW : 0;
        const float* src = p.z + (b * kD) * p.HW + hw;
#pragma unroll 8
        for (int i = 0; i < kD / 8; i++) {
            const int d = warp + 8 * i;
            tile[tile_off(lane, d)] = ok ? __ldcs(src + (int64_t)d * p.HW) : 0.0f;
        }
    }
    float4 gv[4][2];
    if (kGoutCL && has_g) {
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int64_t n = n0 + warp * 4 + rr;
            const float4* g4 = reinterpret_cast<const float4*>(p.gout + (n < p.N ? (n / p.HW) * p.gs_b + (n % p.HW) * p.gs_hw : 0));
#pragma unroll
            for (int h = 0; h < 2; h++) gv[rr][h] = (n < p.N) ? __ldcs(g4 + lane + 32 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();

    // 2. row form: diff = z - e, scatter-add into the codebook gradient, grad (or diff) back into the tile
    {
        const float ce = -(p.beta * coef) * p.e_scale;
        const float fx_mul = kDet ? pow2f(__ldg(p.fx_shift)) : 0.0f;     // exact power of two
        float4 ev[4][2];
        int kk[4];
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int r = warp * 4 + rr;
            kk[rr] = (n0 + r < p.N) ? idx_s[r] : -1;
            const float4* e4 = reinterpret_cast<const float4*>(p.E + (int64_t)max(kk[rr], 0) * kD);
#pragma unroll
            for (int h = 0; h < 2; h++) ev[rr][h] = (kk[rr] >= 0) ? __ldg(e4 + lane + 32 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int r = warp * 4 + rr;
            const bool live = (n0 + r < p.N) && kk[rr] >= 0;
            float4* zrow4 = reinterpret_cast<float4*>(tile + r * kD);
            const int g = tile_swz(r);
            float* ge = (!kDet && p.grad_E != nullptr && live) ? p.grad_E + (int64_t)kk[rr] * kD : nullptr;
            long long* gx = (kDet && p.acc_fx != nullptr && live) ? p.acc_fx + (int64_t)kk[rr] * kD : nullptr;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int q = lane + 32 * h;
                const float4 zv = zrow4[q ^ g];
                const float4 e = ev[rr][h];
                float4 diff = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    diff.x = __fsub_rn(zv.x, e.x); diff.y = __fsub_rn(zv.y, e.y);
                    diff.z = __fsub_rn(zv.z, e.z); diff.w = __fsub_rn(zv.w, e.w);
                }
                if (ge != nullptr) red_add_v4(ge + 4 * q, ce * diff.x, ce * diff.y, ce * diff.z, ce * diff.w);
                if (kDet && gx != nullptr) {
                    // diff * 2^shift is exact in fp32 (a power-of-two scaling) and below 2^38: the conversion is exact too
                    red_add_s64(gx + 4 * q + 0, __float2ll_rn(diff.x * fx_mul));
                    red_add_s64(gx + 4 * q + 1, __float2ll_rn(diff.y * fx_mul));
                    red_add_s64(gx + 4 * q + 2, __float2ll_rn(diff.z * fx_mul));
                    red_add_s64(gx + 4 * q + 3, __float2ll_rn(diff.w * fx_mul));
                }
                if (kGoutCL) {
                    float4 o;
                    const float4 gg = has_g ? gv[rr][h] : make_float4(0.f, 0.f, 0.f, 0.f);
                    o.x = __fmaf_rn(coef, diff.x, gg.x); o.y = __fmaf_rn(coef, diff.y, gg.y);
                    o.z = __fmaf_rn(coef, diff.z, gg.z); o.w = __fmaf_rn(coef, diff.w, gg.w);
                    zrow4[q ^ g] = o;
                } else {
                    zrow4[q ^ g] = diff;
                }
            }
        }
    }
    if (p.grad_z == nullptr) return;
    __syncthreads();

    // 3. column form: grad_z as NCHW.  With a channels-last g_out the tile already holds the result; otherwise it holds
    //    diff and the upstream gradient is read here with its own strides (hw-contiguous: 16-byte loads like z).
    if (kVec) {
        const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
        float* dstb = p.grad_z + (b * kD + dsub) * p.HW + hw0 + 4 * hq;
        const bool g_vec = !kGoutCL && has_g && p.gs_hw == 1 && p.gs_d % 4 == 0 && p.gs_b % 4 == 0 &&
                           (reinterpret_cast<uintptr_t>(p.gout) & 15) == 0;
        const float* gsrc = has_g ? p.gout + b * p.gs_b + dsub * p.gs_d + (hw0 + 4 * hq) * p.gs_hw : nullptr;
        float4 gq[8];
        if (!kGoutCL) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                gq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g_vec) {
                    gq[i] = __ldcs(reinterpret_cast<const float4*>(gsrc + (int64_t)((warp * 8 + i) * 4) * p.gs_d));
                } else if (has_g) {
                    const float* gp = gsrc + (int64_t)((warp * 8 + i) * 4) * p.gs_d;
                    gq[i].x = __ldcs(gp); gq[i].y = __ldcs(gp + p.gs_hw);
                    gq[i].z = __ldcs(gp + 2 * p.gs_hw); gq[i].w = __ldcs(gp + 3 * p.gs_hw);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int d = (warp * 8 + i) * 4 + dsub;
            float4 o;
            o.x = tile[tile_off(4 * hq + 0, d)];
            o.y = tile[tile_off(4 * hq + 1, d)];
            o.z = tile[tile_off(4 * hq + 2, d)];
            o.w = tile[tile_off(4 * hq + 3, d)];
            if (!kGoutCL) {
                o.x = __fmaf_rn(coef, o.x, gq[i].x); o.y = __fmaf_rn(coef, o.y, gq[i].y);
                o.z = __fmaf_rn(coef, o.z, gq[i].z); o.w = __fmaf_rn(coef, o.w, gq[i].w);
            }
            __stcs(reinterpret_cast<float4*>(dstb + (int64_t)((warp * 8 + i) * 4) * p.HW), o);
        }
    } else {
        const int64_t n = n0 + lane;
        if (n < p.N) {
            const int64_t b = n / p.HW, hw = n % p.HW;
            float* dst = p.grad_z + (b * kD) * p.HW + hw;
            const float* g = (!kGoutCL && has_g) ? p.gout + b * p.gs_b + hw * p.gs_hw : nullptr;
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                float o = tile[tile_off(lane, d)];
                if (!kGoutCL) o = __fmaf_rn(coef, o, g != nullptr ? __ldcs(g + (int64_t)d * p.gs_d) : 0.0f);
                __stcs(dst + (int64_t)d * p.HW, o);
            }
        }
    }
}

// ---- deterministic mode helpers -------------------------------------------------------------------------------------
// Pre-pass: max |z - e| over all latents -> the fixed-point shift.  maxbits must be zero on entry (non-negative floats
// order like their bit patterns; a NaN / Inf difference has the largest pattern and makes the shift 0 -- the sums are
// garbage then, exactly like the NaN gradient of the float path).  One CTA = 32 latents, lanes over latents.
__global__ void __launch_bounds__(kBwdThreads)
vq_backward_maxdiff_kernel(const float* __restrict__ z, const int64_t* __restrict__ idx, const float* __restrict__ E,
                           int64_t N, int64_t HW, int K, unsigned int* __restrict__ maxbits) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n = (int64_t)blockIdx.x * kSelRows + lane;
    float mx = 0.0f;
    if (n < N) {
        const int64_t kk = __ldg(idx + n);
        if (kk >= 0 && kk < K) {
            const float* src = z + ((n / HW) * kD) * HW + n % HW;
            const float* e = E + kk * kD;
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                float v = fabsf(__fsub_rn(__ldg(src + (int64_t)d * HW), __ldg(e + d)));
                if (!(v == v)) v = INFINITY;                 // (fmaxf would drop a NaN)
                mx = fmaxf(mx, v);
            }
        }
    }
    unsigned int b = __float_as_uint(mx);
    b = __reduce_max_sync(0xffffffffu, b);
    if (lane == 0 && b != 0u) atomicMax(maxbits, b);
}

// shift = kFxTopBit - ex with |maxdiff| < 2^ex (single thread)
__global__ void vq_backward_fxshift_kernel(const unsigned int* __restrict__ maxbits, int* __restrict__ shift) {
    const float mx = __uint_as_float(*maxbits);
    int ex = exponent_of(mx);
    *shift = max(-100, min(100, kFxTopBit - ex));
}

// grad_E[k][d] = fl( acc * 2^-shift * ce ), ce = -(beta * coef) * e_scale evaluated exactly like the float path
__global__ void __launch_bounds__(256)
vq_backward_fxfinish_kernel(const long long* __restrict__ acc, const int* __restrict__ shift, int64_t n_elems,
                            float g_loss, const float* __restrict__ g_loss_dev, double inv_nd, float beta, float e_scale,
                            float* __restrict__ grad_E) {
    const float coef = (float)(2.0 * (double)(g_loss_dev != nullptr ? __ldg(g_loss_dev) : g_loss) * inv_nd);
    const float ce = -(beta * coef) * e_scale;
    const double unit = (double)ce * (double)pow2f(-__ldg(shift));
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * 256)
        grad_E[i] = (float)((double)acc[i] * unit);
}

// grad_E = (beta * coef * e_scale) * S, S[k] = sum over the latents of code k of (e_k - z_n), accumulated by the FORWARD
// (vq_select_kernel with SelectParams::scat): the scatter-add then needs no second pass over z, and a data-parallel wrapper can
// exchange S while the rest of the step runs (dist.py).
__global__ void __launch_bounds__(256)
vq_grad_from_sum_kernel(const float* __restrict__ S, int64_t n_elems, float g_loss, const float* __restrict__ g_loss_dev, double inv_nd,
                        float beta, float e_scale, float* __restrict__ grad_E) {
    const float coef = (float)(2.0 * (double)(g_loss_dev != nullptr ? __ldg(g_loss_dev) : g_loss) * inv_nd);
    const float ce = (beta * coef) * e_scale;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n_elems; i += (int64_t)gridDim.x * 1024) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(S + i));
        *reinterpret_cast<float4*>(grad_E + i) = make_float4(ce * v.x, ce * v.y, ce * v.z, ce * v.w);
    }
}

// out[b, d, hw] = E[idx[b*HW + hw]][d]  (decode side: worker/vqganVqvaeWorker.py:459, vqTransformer.py:98)
template <bool kVec>
__global__ void __launch_bounds__(kBwdThreads)
vq_embed_nchw_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E, int64_t N, int64_t HW, int K,
                     float* __restrict__ out) {
    __shared__ float et[kD][kSelRows + 1];
    __shared__ int idx_s[kSelRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    if (tid < kSelRows) {
        const int64_t n = n0 + tid;
        int64_t kk = (n < N) ? __ldg(idx + n) : -1;
        idx_s[tid] = (kk < 0 || kk >= K) ? -1 : (int)kk;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int k = idx_s[r];
        const float* e = E + (int64_t)max(k, 0) * kD;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            et[d][r] = (k >= 0) ? __ldg(e + d) : 0.0f;
        }
    }
    __syncthreads();
    if (kVec) {
        const int64_t b = n0 / HW, hw0 = n0 % HW;
        const int dsub = lane >> 3, hq = lane & 7;
        float* dstb = out + (b * kD + dsub) * HW + hw0 + 4 * hq;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int d = (warp * 8 + i) * 4 + dsub;
            const float4 o = make_float4(et[d][4 * hq], et[d][4 * hq + 1], et[d][4 * hq + 2], et[d][4 * hq + 3]);
            __stcs(reinterpret_cast<float4*>(dstb + (int64_t)((warp * 8 + i) * 4) * HW), o);
        }
    } else {
        const int64_t n = n0 + lane;
        if (n < N) {
            float* dst = out + ((n / HW) * kD) * HW + (n % HW);
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                __stcs(dst + (int64_t)d * HW, et[d][lane]);
            }
        }
    }
}

}  // namespace vq
