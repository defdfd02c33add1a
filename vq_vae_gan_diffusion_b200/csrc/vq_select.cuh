// vq_select.cuh -- exact fp32 decision + fused forward tail.
//
// vq_select_kernel<kForward>: one CTA = 32 latents (8 warps, 4 rows per warp).
//   1. load the fp32 z tile once (coalesced along hw) into shared memory
//   2. per row: recompute the distances of the candidate quads handed over by the GEMM epilogue with the
//      reference's fp32 formula (codebook.py:70-79) in the oracle's canonical accumulation order and take the
//      first minimum (torch.argmin semantics, codebook.py:82); out_cnt < 0 -> exact scan of the whole row
//   3. kForward only: gather e = E[idx] (codebook.py:85), write z_q = fl(z + fl(e - z)) as NHWC rows
//      (codebook.py:106-109), accumulate sum (e - z)^2 for the loss (codebook.py:96-103) and the usage histogram.
// HBM traffic per latent: read z 4D (+ 36 B of candidates), write idx 8 (+ z_q 4D when kForward).
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kSelThreads = 256;

struct SelectParams {
    const float* z;            // (B, D, HW) fp32
    const float* E;            // (K, D) fp32
    const float* e2;           // (K_pad)
    const float* z2;           // (N)
    const int32_t* out_cnt;    // (N)
    const int32_t* out_q;      // (N, kOutCap)
    int64_t N, HW;
    int K;
    float beta;
    int64_t* idx;              // (N)
    float* zq;                 // (N, D) or null
    unsigned long long* hist;  // (K) or null
    double* loss_partial;      // (gridDim.x)
    unsigned int* blocks_done; // (1) zero-initialised, re-armed by the kernel
    float* loss;               // (1)
    unsigned long long* stats; // (VQ_STAT_COUNT) or null
};

// lexicographic (distance, index) minimum across the warp -> first minimum
__device__ __forceinline__ void warp_argmin(float& d, int& k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
        const int k2 = __shfl_xor_sync(0xffffffffu, k, o);
        if (d2 < d || (d2 == d && k2 < k)) { d = d2; k = k2; }
    }
}

template <bool kForward>
__global__ void __launch_bounds__(kSelThreads)
vq_select_kernel(const SelectParams p) {
    __shared__ float t[kD][kSelRows + 1];
    __shared__ int idx_s[kSelRows];
    __shared__ double red_s[kSelThreads / 32];
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;

    {   // 1. z tile
        const int64_t n = n0 + lane;
        const bool ok = n < p.N;
        const int64_t b = ok ? n / p.HW : 0, hw = ok ? n % p.HW : 0;
        const float* src = p.z + (b * kD) * p.HW + hw;
#pragma unroll 8
        for (int i = 0; i < kD / 8; i++) {
            const int d = warp + 8 * i;
            t[d][lane] = ok ? __ldg(src + (int64_t)d * p.HW) : 0.0f;
        }
    }
    __syncthreads();

    // 2. decide the index of rows 4*warp .. 4*warp+3.  lane = 4*g + j: code slot g of the pass, partial sum j.
    const int j = lane & 3, g = lane >> 2;
    unsigned long long st_tie = 0, st_rerank = 0, st_fallback = 0, st_cand = 0;
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        if (n >= p.N) break;                                  // warp-uniform
        const float z2 = __ldg(p.z2 + n);
        const int nq = __ldg(p.out_cnt + n);
        const bool scan_all = (nq <= 0 || nq > kOutCap);
        const int my_q = (!scan_all && lane < nq) ? __ldg(p.out_q + n * kOutCap + lane) : 0;
        const int total = scan_all ? p.K : nq * kQuad;        // codes to evaluate (some may be >= K: skipped)

        float best_d = INFINITY;
        int best_k = 0x7fffffff, n_at_min = 0;
        for (int base = 0; base < total; base += 8) {
            const int slot = base + g;                        // code slot of this lane group
            const int qsrc = __shfl_sync(0xffffffffu, my_q, (slot / kQuad) & 31);
            int k = scan_all ? slot : qsrc * kQuad + (slot % kQuad);
            if (slot >= total || k >= p.K) k = -1;
            float acc = 0.0f;
            if (k >= 0) {
                const float* e = p.E + (int64_t)k * kD + j;
#pragma unroll 8
                for (int q = 0; q < kD / 4; q++) acc = __fmaf_rn(t[4 * q + j][r], __ldg(e + 4 * q), acc);
            }
            const float dot = combine4(acc);
            const float dist = (k >= 0) ? ref_distance(z2, __ldg(p.e2 + k), dot) : INFINITY;
            float pd = dist;
            int pk = (k >= 0) ? k : 0x7fffffff;
            warp_argmin(pd, pk);
            const int eq = __popc(__ballot_sync(0xffffffffu, j == 0 && k >= 0 && dist == pd));
            if (pd < best_d) { best_d = pd; best_k = pk; n_at_min = eq; }
            else if (pd == best_d) { n_at_min += eq; best_k = min(best_k, pk); }
        }
        if (best_k == 0x7fffffff) best_k = 0;                 // every distance NaN: torch.argmin -> 0 as well
        if (n_at_min > 1) st_tie++;
        if (scan_all) st_fallback++; else if (nq > 1) st_rerank++;
        st_cand += scan_all ? (unsigned long long)p.K : (unsigned long long)nq;
        if (lane == 0) {
            idx_s[r] = best_k;
            p.idx[n] = (int64_t)best_k;
        }
    }
    if (p.stats != nullptr && lane == 0) {
        if (st_tie) atomicAdd(p.stats + 0, st_tie);
        if (st_rerank) atomicAdd(p.stats + 1, st_rerank);
        if (st_fallback) atomicAdd(p.stats + 2, st_fallback);
        if (st_cand) atomicAdd(p.stats + 3, st_cand);
    }
    if (!kForward) return;

    // 3. forward tail
    __syncthreads();
    float sq = 0.0f;
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        if (n >= p.N) break;
        const int k = idx_s[r];
        const float* e = p.E + (int64_t)k * kD;
        float* out = p.zq + n * kD;
#pragma unroll
        for (int i = 0; i < kD / 32; i++) {
            const int d = lane + 32 * i;
            const float zv = t[d][r];
            const float diff = __fsub_rn(__ldg(e + d), zv);    // fl(e - z)
            __stcs(out + d, __fadd_rn(zv, diff));              // fl(z + fl(e - z)), codebook.py:106
            sq = __fmaf_rn(diff, diff, sq);
        }
        if (p.hist != nullptr && lane == 0) atomicAdd(p.hist + k, 1ull);
    }
    // loss: fp32 per thread (<= 32 terms), fp64 from there on; fixed-order final sum by the last CTA
    double v = (double)sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red_s[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < kSelThreads / 32; w++) sum += red_s[w];
        p.loss_partial[blockIdx.x] = sum;
        __threadfence();
        is_last = (atomicAdd(p.blocks_done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double sum = 0.0;
        for (unsigned i = tid; i < gridDim.x; i += kSelThreads) sum += __ldcg(p.loss_partial + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncthreads();
        if (lane == 0) red_s[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < kSelThreads / 32; w++) tot += red_s[w];
            const double m = tot / ((double)p.N * (double)kD);
            *p.loss = (float)(m + (double)p.beta * m);       // mean(a + beta*mean(b)), a == b elementwise
            *p.blocks_done = 0;                              // re-arm for the next call on this workspace
        }
    }
}

}  // namespace vq
