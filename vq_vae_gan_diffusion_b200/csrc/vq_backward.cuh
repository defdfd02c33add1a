// vq_backward.cuh -- fused straight-through backward + codebook scatter-add, and the NCHW embedding lookup.
//
// vq_backward_kernel: one CTA = 32 latents, two [256 x 32] fp32 tiles in shared memory.
//   1. z tile (NCHW, 16-byte loads, lanes over hw) and the upstream-gradient tile: channels-last g_out (d contiguous,
//      the layout of the z_q we returned) is read lanes-over-d, an hw-contiguous g_out like z; any other layout with
//      strided scalar loads
//   2. lanes over d: gather e = E[idx[n]] rows (coalesced 1 KiB reads, L2 resident), diff = z - e kept in the z tile,
//      grad_E[idx[n]][d] += -coef * beta * diff  (red.global.add.f32, 128-byte coalesced)
//   3. lanes over hw: grad_z = g_out + coef * diff written as NCHW with 16-byte stores   [autograd of codebook.py:96-106]
// HBM traffic per latent: read g_out 4D + z 4D + idx 8, write grad_z 4D -- the algorithmic minimum; the codebook
// and its gradient (16 MiB each at K = 16384) stay in the 126 MB L2.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kBwdThreads = 256;
constexpr size_t kBwdTileBytes = (size_t)kD * (kSelRows + 1) * sizeof(float);

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
    float* grad_z;            // (B, D, HW) or null
    float* grad_E;            // (K, D) or null, zeroed before launch
};

// kVec: HW % 32 == 0 (see load_tile_nchw).  kGoutCL: g_out is channels-last (gs_d == 1).
template <bool kVec, bool kGoutCL>
__global__ void __launch_bounds__(kBwdThreads)
vq_backward_kernel(const BackwardParams p) {
    extern __shared__ __align__(16) float bsm[];
    TileRow* zt = reinterpret_cast<TileRow*>(bsm);                          // z, then z - e
    TileRow* gt = reinterpret_cast<TileRow*>(bsm + kD * (kSelRows + 1));    // upstream gradient
    __shared__ int idx_s[kSelRows];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    // coef = 2 * g_loss / (n_global * D), evaluated like the oracle (double, rounded once)
    const float coef = (float)(2.0 * (double)(p.g_loss_dev != nullptr ? __ldg(p.g_loss_dev) : p.g_loss) * p.inv_nd);

    if (tid < kSelRows) {
        const int64_t n = n0 + tid;
        int k = -1;
        if (n < p.N) {
            const int64_t kk = __ldg(p.idx + n);
            k = (kk < 0 || kk >= p.K) ? -1 : (int)kk;       // out-of-range index: contributes nothing
        }
        idx_s[tid] = k;
    }

    // 1. tiles
    const bool has_g = p.gout != nullptr && p.grad_z != nullptr;
    load_tile_nchw<kVec, true>(zt, p.z, n0, p.N, p.HW, warp, lane);
    if (has_g) {
        if (kGoutCL) {
            // rows of 256 contiguous floats: warp w stages rows 4w..4w+3, lanes over d (8 requests per row in flight)
            float v[4][kD / 32];
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                const int64_t n = n0 + warp * 4 + rr;
                const float* g = p.gout + (n < p.N ? (n / p.HW) * p.gs_b + (n % p.HW) * p.gs_hw : 0);
#pragma unroll
                for (int i = 0; i < kD / 32; i++) v[rr][i] = (n < p.N) ? __ldcs(g + lane + 32 * i) : 0.0f;
            }
#pragma unroll
            for (int rr = 0; rr < 4; rr++)
#pragma unroll
                for (int i = 0; i < kD / 32; i++) gt[lane + 32 * i][warp * 4 + rr] = v[rr][i];
        } else if (kVec && p.gs_hw == 1 && p.gs_d % 4 == 0 && p.gs_b % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(p.gout) & 15) == 0) {
            // hw-contiguous (e.g. NCHW-contiguous) gradient: same access pattern as z, with its own strides
            const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
            const int dsub = lane >> 3, hq = lane & 7;
            const float* src = p.gout + b * p.gs_b + dsub * p.gs_d + hw0 + 4 * hq;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)((warp * 8 + i) * 4) * p.gs_d));
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float* dst = &gt[(warp * 8 + i) * 4 + dsub][4 * hq];
                dst[0] = v[i].x; dst[1] = v[i].y; dst[2] = v[i].z; dst[3] = v[i].w;
            }
        } else {
            const int64_t n = n0 + lane;
            const bool ok = n < p.N;
            const float* g = p.gout + (ok ? (n / p.HW) * p.gs_b + (n % p.HW) * p.gs_hw : 0);
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                gt[d][lane] = ok ? __ldcs(g + (int64_t)d * p.gs_d) : 0.0f;
            }
        }
    }
    __syncthreads();

    // 2. lanes over d: diff = z - e (kept in the tile), scatter-add into the codebook gradient
    {
        const float ce = -(p.beta * coef);
        float ev[4][kD / 32];
        int kk[4];
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int r = warp * 4 + rr;
            kk[rr] = (n0 + r < p.N) ? idx_s[r] : -1;
            const float* e = p.E + (int64_t)max(kk[rr], 0) * kD;
#pragma unroll
            for (int i = 0; i < kD / 32; i++) ev[rr][i] = (kk[rr] >= 0) ? __ldg(e + lane + 32 * i) : 0.0f;
        }
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int r = warp * 4 + rr;
            float* ge = (p.grad_E != nullptr && kk[rr] >= 0) ? p.grad_E + (int64_t)kk[rr] * kD : nullptr;
#pragma unroll
            for (int i = 0; i < kD / 32; i++) {
                const int d = lane + 32 * i;
                const float diff = (n0 + r < p.N) ? __fsub_rn(zt[d][r], ev[rr][i]) : 0.0f;
                zt[d][r] = diff;
                if (ge != nullptr) atomicAdd(ge + d, ce * diff);      // result unused -> RED.E.ADD.F32
            }
        }
    }
    if (p.grad_z == nullptr) return;
    __syncthreads();

    // 3. grad_z = g_out + coef * diff, NCHW
    if (kVec) {
        const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
        const int dsub = lane >> 3, hq = lane & 7;
        float* dstb = p.grad_z + (b * kD + dsub) * p.HW + hw0 + 4 * hq;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int d = (warp * 8 + i) * 4 + dsub;
            float4 o;
            if (has_g) {
                o.x = __fmaf_rn(coef, zt[d][4 * hq + 0], gt[d][4 * hq + 0]);
                o.y = __fmaf_rn(coef, zt[d][4 * hq + 1], gt[d][4 * hq + 1]);
                o.z = __fmaf_rn(coef, zt[d][4 * hq + 2], gt[d][4 * hq + 2]);
                o.w = __fmaf_rn(coef, zt[d][4 * hq + 3], gt[d][4 * hq + 3]);
            } else {
                o.x = __fmaf_rn(coef, zt[d][4 * hq + 0], 0.0f);
                o.y = __fmaf_rn(coef, zt[d][4 * hq + 1], 0.0f);
                o.z = __fmaf_rn(coef, zt[d][4 * hq + 2], 0.0f);
                o.w = __fmaf_rn(coef, zt[d][4 * hq + 3], 0.0f);
            }
            __stcs(reinterpret_cast<float4*>(dstb + (int64_t)((warp * 8 + i) * 4) * p.HW), o);
        }
    } else {
        const int64_t n = n0 + lane;
        if (n < p.N) {
            float* dst = p.grad_z + ((n / p.HW) * kD) * p.HW + (n % p.HW);
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                __stcs(dst + (int64_t)d * p.HW, __fmaf_rn(coef, zt[d][lane], has_g ? gt[d][lane] : 0.0f));
            }
        }
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
