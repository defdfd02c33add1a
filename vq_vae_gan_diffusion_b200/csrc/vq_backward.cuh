// vq_backward.cuh -- fused straight-through backward + codebook scatter-add, and the NCHW embedding lookup.
//
// vq_backward_kernel: one CTA = 32 latents.
//   A. lanes over d:  gather e = E[idx[n]] rows (coalesced 1 KiB reads, L2 resident) into a shared tile;
//                     when the upstream gradient is channels-last (d contiguous, the layout of z_q itself) it is
//                     staged through a second shared tile here as well
//   B. lanes over hw: read z (NCHW, coalesced), form diff = z - e, write
//                     grad_z = g_out + coef * diff   (NCHW, coalesced)          [autograd of codebook.py:96-106]
//                     and keep diff in the tile
//   C. lanes over d:  grad_E[idx[n]][d] += -coef * beta * diff                  (red.global.add.f32, coalesced)
// HBM traffic per latent: read g_out 4D + z 4D + idx 8, write grad_z 4D -- the algorithmic minimum; the codebook
// and its gradient (16 MiB each at K = 16384) stay in the 126 MB L2.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kBwdThreads = 256;

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

template <bool kGoutChannelsLast>
__global__ void __launch_bounds__(kBwdThreads)
vq_backward_kernel(const BackwardParams p) {
    extern __shared__ float bsm[];
    float (*et)[kSelRows + 1] = reinterpret_cast<float (*)[kSelRows + 1]>(bsm);                         // [kD][33]
    float (*gt)[kSelRows + 1] = reinterpret_cast<float (*)[kSelRows + 1]>(bsm + kD * (kSelRows + 1));   // [kD][33]
    __shared__ int idx_s[kSelRows];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    // coef = 2 * g_loss / (n_global * D), evaluated like the oracle (double, rounded once)
    const float coef = (float)(2.0 * (double)(p.g_loss_dev != nullptr ? __ldg(p.g_loss_dev) : p.g_loss) * p.inv_nd);

    if (tid < kSelRows) {
        const int64_t n = n0 + tid;
        int k = 0;
        if (n < p.N) {
            const int64_t kk = __ldg(p.idx + n);
            k = (kk < 0 || kk >= p.K) ? -1 : (int)kk;       // out-of-range index: contributes nothing
        }
        idx_s[tid] = k;
    }
    __syncthreads();

    // A. gather code rows (and channels-last g_out rows): warp w handles rows 4w..4w+3, lanes over d
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        if (n >= p.N) break;
        const int k = idx_s[r];
        const float* e = p.E + (int64_t)max(k, 0) * kD;
        const float* g = nullptr;
        if (kGoutChannelsLast && p.gout != nullptr) g = p.gout + (n / p.HW) * p.gs_b + (n % p.HW) * p.gs_hw;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            et[d][r] = (k >= 0) ? __ldg(e + d) : 0.0f;
            if (kGoutChannelsLast) gt[d][r] = (g != nullptr) ? __ldcs(g + d) : 0.0f;
        }
    }
    __syncthreads();

    // B. lanes over latents (hw contiguous): grad_z and diff
    {
        const int64_t n = n0 + lane;
        const bool ok = n < p.N;
        const int64_t b = ok ? n / p.HW : 0, hw = ok ? n % p.HW : 0;
        const int64_t base = (b * kD) * p.HW + hw;
        const float* gsrc = (!kGoutChannelsLast && p.gout != nullptr) ? p.gout + b * p.gs_b + hw * p.gs_hw : nullptr;
#pragma unroll 8
        for (int i = 0; i < kD / 8; i++) {
            const int d = warp + 8 * i;
            float diff = 0.0f;
            if (ok) {
                const float zv = __ldcs(p.z + base + (int64_t)d * p.HW);
                diff = __fsub_rn(zv, et[d][lane]);
                if (p.grad_z != nullptr) {
                    float g;
                    if (kGoutChannelsLast) g = gt[d][lane];
                    else g = (gsrc != nullptr) ? __ldcs(gsrc + (int64_t)d * p.gs_d) : 0.0f;
                    __stcs(p.grad_z + base + (int64_t)d * p.HW, __fmaf_rn(coef, diff, g));
                }
            }
            et[d][lane] = diff;
        }
    }
    if (p.grad_E == nullptr) return;
    __syncthreads();

    // C. scatter-add into the codebook gradient: lanes over d -> 128-byte coalesced reductions
    const float ce = -(p.beta * coef);
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        if (n0 + r >= p.N) break;
        const int k = idx_s[r];
        if (k < 0) continue;
        float* ge = p.grad_E + (int64_t)k * kD;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            atomicAdd(ge + d, ce * et[d][r]);     // result unused -> RED.E.ADD.F32
        }
    }
}

// out[b, d, hw] = E[idx[b*HW + hw]][d]  (decode side: worker/vqganVqvaeWorker.py:459, vqTransformer.py:98)
__global__ void __launch_bounds__(kBwdThreads)
vq_embed_nchw_kernel(const int64_t* __restrict__ idx, const float* __restrict__ E, int64_t N, int64_t HW, int K,
                     float* __restrict__ out) {
    __shared__ float et[kD][kSelRows + 1];
    __shared__ int idx_s[kSelRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    if (tid < kSelRows) {
        const int64_t n = n0 + tid;
        int64_t kk = (n < N) ? __ldg(idx + n) : 0;
        idx_s[tid] = (kk < 0 || kk >= K) ? -1 : (int)kk;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        if (n0 + r >= N) break;
        const int k = idx_s[r];
        const float* e = E + (int64_t)max(k, 0) * kD;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            et[d][r] = (k >= 0) ? __ldg(e + d) : 0.0f;
        }
    }
    __syncthreads();
    const int64_t n = n0 + lane;
    if (n < N) {
        const int64_t b = n / HW, hw = n % HW;
        float* dst = out + (b * kD) * HW + hw;
#pragma unroll 8
        for (int i = 0; i < kD / 8; i++) {
            const int d = warp + 8 * i;
            __stcs(dst + (int64_t)d * HW, et[d][lane]);
        }
    }
}

}  // namespace vq
