// vq_select.cuh -- exact fp32 decision + fused forward tail, and the exact full-row fallback.
//
// vq_select_kernel<kForward>: one CTA = 32 latents (8 warps, 4 rows per warp).
//   1. expand the candidate entries, then load the fp32 z tile once (16-byte loads along hw) into shared memory
//      (skipped in tokeniser mode when every row of the CTA has a single candidate)
//   2. expand the candidate entries (32-code chunk + quad mask) written by the GEMM epilogue into quads, then
//      recompute the distance of every candidate code with the reference's fp32 formula (codebook.py:70-79) in
//      the oracle's canonical accumulation order -- one thread per (row, code): a warp pass covers 4 rows x 8 codes,
//      each thread streaming its code row with 128-bit loads into the four canonical partial sums -- and take the
//      first minimum (torch.argmin semantics, codebook.py:82)
//   3. kForward only: gather e = E[idx] (codebook.py:85), write z_q = fl(z + fl(e - z)) as NHWC rows
//      (codebook.py:106-109), accumulate sum (e - z)^2 for the loss (codebook.py:96-103) and the usage histogram.
// HBM traffic per latent: read z 4D (+ 72 B of candidates), write idx 8 (+ z_q 4D when kForward).
//
// vq_fallback_kernel: the (rare) rows whose candidate ring overflowed in the GEMM epilogue get an exact scan of the
// whole codebook, one 1024-thread CTA per row, one thread per code; the winner is written back as a one-quad
// candidate entry so vq_select_kernel finishes the row like any other.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kMaxCands = 64;           // codes per row the exact stage evaluates (GEMM hands over <= 32 per group)
constexpr int kFbThreads = 256;
constexpr int kFbGroup = 8;             // overflowed rows scanned together (they share every code-row load)
constexpr int kFbMaxGroups = 512;       // row groups whose scan is split over code blocks (4096 rows)
constexpr int kFbMaxParts = 256;        // code blocks per group

struct SelectParams {
    const float* z;            // (B, D, HW) fp32
    const float* E;            // (K, D) fp32
    const float* e2;           // (K_pad)
    const float* z2;           // (N)
    const int32_t* out_cnt;    // (N, 2)
    const uint32_t* out_q;     // (N, kOutCap)
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

// canonical-order distance of latent row (shared tile column r) to code k; one thread does the whole dot product
__device__ __forceinline__ float exact_distance_tile(const TileRow* t, int r, const float* __restrict__ E,
                                                     const float* __restrict__ e2, int k, float z2) {
    const float4* e4 = reinterpret_cast<const float4*>(E + (int64_t)k * kD);
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
#pragma unroll 16
    for (int q = 0; q < kD / 4; q++) {
        const float4 e = __ldg(e4 + q);
        p0 = __fmaf_rn(t[4 * q + 0][r], e.x, p0);
        p1 = __fmaf_rn(t[4 * q + 1][r], e.y, p1);
        p2 = __fmaf_rn(t[4 * q + 2][r], e.z, p2);
        p3 = __fmaf_rn(t[4 * q + 3][r], e.w, p3);
    }
    const float dot = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
    return ref_distance(z2, __ldg(e2 + k), dot);
}

template <bool kForward, bool kVec>
__global__ void __launch_bounds__(kSelThreads)
vq_select_kernel(const SelectParams p) {
    __shared__ TileRow t[kD];
    __shared__ int clist[kSelWarps][4][kMaxCands];            // candidate codes of each row of each warp
    __shared__ int idx_s[kSelRows];
    __shared__ double red_s[kSelWarps];
    __shared__ unsigned int st_s[4];                          // per-CTA counters (same-address global atomics are slow)
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    if (tid < 4) st_s[tid] = 0;

    // 2a. expand this warp's candidate entries into quads (independent of the z tile)
    int nq[4];
    unsigned resolved_mask = 0;                               // rows decided by vq_fallback_kernel (stats counted there)
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t n = n0 + warp * 4 + rr;
        nq[rr] = 0;
        if (n >= p.N) continue;                               // warp-uniform
        const int c0 = __ldg(p.out_cnt + 2 * n), c1 = __ldg(p.out_cnt + 2 * n + 1);
        const bool resolved = (c0 == -2);
        if (resolved) resolved_mask |= 1u << rr;
        const int g = lane >> 3, i = lane & 7;                // slot = lane: group g owns slots [8g, 8g + 8)
        const bool valid = resolved ? (lane == 0) : (lane < kOutCap && i < (g ? c1 : c0));
        uint2 e = make_uint2(0u, 0u);                         // (chunk id, mask of candidate codes in the chunk)
        if (valid) e = __ldg(reinterpret_cast<const uint2*>(p.out_q) + n * kOutCap + lane);
        const int pc = __popc(e.y);
        int incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int pos = incl - pc;
        uint32_t bits = e.y;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (pos < kMaxCands) clist[warp][rr][pos] = (int)e.x * 32 + b;
            pos++;
        }
        nq[rr] = min(kMaxCands, __shfl_sync(0xffffffffu, incl, 31));
    }

    // 1. z tile (fp32): always needed by the forward tail; in tokeniser mode only when some row of this CTA has more
    //    than one candidate and therefore needs exact distances
    const int need_exact = (nq[0] > 1) | (nq[1] > 1) | (nq[2] > 1) | (nq[3] > 1);
    const bool want_tile = kForward ? true : (__syncthreads_or(need_exact) != 0);
    if (want_tile) load_tile_nchw<kVec, false>(t, p.z, n0, p.N, p.HW, warp, lane);
    __syncthreads();

    // 2b. exact distances: lane = 8*rr + c -> row rr of this warp, code slot c (two quads per pass and row)
    {
        const int rr = lane >> 3, c = lane & 7;
        const int r = warp * 4 + rr;
        const int64_t n = n0 + r;
        const bool row_ok = n < p.N;
        const int my_nq = (rr == 0) ? nq[0] : (rr == 1) ? nq[1] : (rr == 2) ? nq[2] : nq[3];
        // a row with a single candidate is decided without any arithmetic (the margin argument guarantees the
        // oracle's argmin is among the candidates), so only rows with >= 2 candidates take part in the passes
        const int max_nq = max(max(nq[0] > 1 ? nq[0] : 0, nq[1] > 1 ? nq[1] : 0), max(nq[2] > 1 ? nq[2] : 0, nq[3] > 1 ? nq[3] : 0));
        const float z2 = row_ok ? __ldg(p.z2 + n) : 0.0f;
        const unsigned seg = 0xffu << (rr * 8);
        float best_d = INFINITY;
        int best_k = 0x7fffffff, n_at_min = 0;
        if (my_nq == 1) {
            best_k = clist[warp][rr][0];
            if (best_k >= p.K) best_k = 0;
            n_at_min = 1;
        }
        for (int base = 0; base < max_nq; base += 8) {
            int k = (my_nq > 1 && base + c < my_nq) ? clist[warp][rr][base + c] : -1;
            if (k >= p.K) k = -1;                              // pad codes of the last chunk
            float dist = INFINITY;
            if (k >= 0) dist = exact_distance_tile(t, r, p.E, p.e2, k, z2);
            // first minimum within the row's 8 lanes (lexicographic on (distance, index))
            float pd = dist;
            int pk = (k >= 0) ? k : 0x7fffffff;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float d2 = __shfl_xor_sync(0xffffffffu, pd, o);
                const int k2 = __shfl_xor_sync(0xffffffffu, pk, o);
                if (d2 < pd || (d2 == pd && k2 < pk)) { pd = d2; pk = k2; }
            }
            const int eq = __popc(__ballot_sync(0xffffffffu, k >= 0 && dist == pd) & seg);
            if (pd < best_d) { best_d = pd; best_k = pk; n_at_min = eq; }
            else if (pd == best_d) { n_at_min += eq; best_k = min(best_k, pk); }
        }
        if (best_k == 0x7fffffff) best_k = 0;                 // every distance NaN: torch.argmin -> 0 as well
        if (row_ok && c == 0) {
            idx_s[r] = best_k;
            p.idx[n] = (int64_t)best_k;
            if (p.stats != nullptr && !((resolved_mask >> rr) & 1u)) {
                if (n_at_min > 1) atomicAdd(&st_s[0], 1u);
                if (my_nq > 1) atomicAdd(&st_s[1], 1u);
                atomicAdd(&st_s[3], (unsigned)my_nq);
            }
        }
    }
    __syncthreads();
    if (p.stats != nullptr && tid < 4 && st_s[tid] != 0) atomicAdd(p.stats + tid, (unsigned long long)st_s[tid]);
    if (!kForward) return;

    // 3. forward tail: all four code rows of this warp are requested before the first one is consumed
    float sq = 0.0f;
    {
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
            if (kk[rr] < 0) continue;
            const int r = warp * 4 + rr;
            float* out = p.zq + (n0 + r) * kD;
#pragma unroll
            for (int i = 0; i < kD / 32; i++) {
                const int d = lane + 32 * i;
                const float zv = t[d][r];
                const float diff = __fsub_rn(ev[rr][i], zv);       // fl(e - z)
                __stcs(out + d, __fadd_rn(zv, diff));              // fl(z + fl(e - z)), codebook.py:106
                sq = __fmaf_rn(diff, diff, sq);
            }
            if (p.hist != nullptr && lane == 0) atomicAdd(p.hist + kk[rr], 1ull);
        }
    }
    // loss: fp32 per thread (<= 32 terms), fp64 from there on; fixed-order final sum by the last CTA
    double v = (double)sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red_s[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < kSelWarps; w++) sum += red_s[w];
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
            for (int w = 0; w < kSelWarps; w++) tot += red_s[w];
            const double m = tot / ((double)p.N * (double)kD);
            *p.loss = (float)(m + (double)p.beta * m);       // mean(a + beta*mean(b)), a == b elementwise
            *p.blocks_done = 0;                              // re-arm for the next call on this workspace
        }
    }
}

struct FallbackParams {
    const float* z;
    const float* E;
    const float* e2;
    const float* z2;
    const int32_t* fb_rows;
    const int32_t* fb_count;
    int64_t HW;
    int K;
    int32_t* out_cnt;
    uint32_t* out_q;
    float4* part;              // (kFbMaxGroups * kFbGroup, parts) partial (distance, index, multiplicity) results
    unsigned int* arrive;      // (kFbMaxGroups) arrival counters, zero on entry, re-armed by the kernel
    int parts, per_part;       // code blocks per group and codes per block (multiple of kFbThreads)
    unsigned long long* stats;
};

// merge (distance, first index, multiplicity) triples: lexicographic minimum, multiplicities of equal minima add up
__device__ __forceinline__ void merge_min(float& d, int& k, int& c, float d2, int k2, int c2) {
    if (d2 < d) { d = d2; k = k2; c = c2; }
    else if (d2 == d) { c += c2; k = min(k, k2); }
}

// Work item = (group of kFbGroup worklist entries, block of codes): one thread per code streams its code row once
// with 128-bit loads and applies it to all rows of the group (held in shared memory), so the codebook traffic of the
// fallback is 1/kFbGroup of a row-by-row scan and a handful of overflowed rows is spread over the whole chip.  The
// last code block of a group to arrive merges the per-block minima.  Groups beyond kFbMaxGroups (a degenerate
// codebook: every row overflows) are scanned block after block by a single CTA each.
__global__ void __launch_bounds__(kFbThreads)
vq_fallback_kernel(const FallbackParams p) {
    __shared__ float4 zr4[kFbGroup][kD / 4];
    __shared__ float sd[kFbGroup][kFbThreads / 32];
    __shared__ int sk[kFbGroup][kFbThreads / 32], sn[kFbGroup][kFbThreads / 32];
    __shared__ int64_t row_s[kFbGroup];
    __shared__ int is_final;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int count = __ldg(p.fb_count);
    const int groups = (count + kFbGroup - 1) / kFbGroup;
    const int parts = p.parts, per_part = p.per_part;
    const int64_t items = (int64_t)groups * parts;
    for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int g = (int)(w / parts), part0 = (int)(w % parts);
        const bool split = g < kFbMaxGroups;
        if (!split && part0 != 0) continue;                      // block-uniform
        __syncthreads();
        if (tid < kFbGroup) row_s[tid] = (g * kFbGroup + tid < count) ? (int64_t)__ldg(p.fb_rows + g * kFbGroup + tid) : -1;
        __syncthreads();
        for (int r = 0; r < kFbGroup; r++) {
            const int64_t n = row_s[r];
            if (tid < kD) reinterpret_cast<float*>(zr4[r])[tid] = (n >= 0) ? __ldg(p.z + ((n / p.HW) * kD + tid) * p.HW + n % p.HW) : 0.0f;
        }
        __syncthreads();
        float z2[kFbGroup], best_d[kFbGroup];
        int best_k[kFbGroup], n_at_min[kFbGroup];
#pragma unroll
        for (int r = 0; r < kFbGroup; r++) {
            z2[r] = (row_s[r] >= 0) ? __ldg(p.z2 + row_s[r]) : 0.0f;
            best_d[r] = INFINITY; best_k[r] = 0x7fffffff; n_at_min[r] = 0;
        }
        const int part_lo = split ? part0 : 0, part_hi = split ? part0 + 1 : parts;
        for (int part = part_lo; part < part_hi; part++) {
            const int k_hi = min(p.K, (part + 1) * per_part);
            for (int k = part * per_part + tid; k < k_hi; k += kFbThreads) {   // ascending k per thread
                const float4* e4 = reinterpret_cast<const float4*>(p.E + (int64_t)k * kD);
                float acc[kFbGroup][4];
#pragma unroll
                for (int r = 0; r < kFbGroup; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
#pragma unroll 4
                for (int q = 0; q < kD / 4; q++) {
                    const float4 e = __ldg(e4 + q);
#pragma unroll
                    for (int r = 0; r < kFbGroup; r++) {
                        const float4 zv = zr4[r][q];
                        acc[r][0] = __fmaf_rn(zv.x, e.x, acc[r][0]);
                        acc[r][1] = __fmaf_rn(zv.y, e.y, acc[r][1]);
                        acc[r][2] = __fmaf_rn(zv.z, e.z, acc[r][2]);
                        acc[r][3] = __fmaf_rn(zv.w, e.w, acc[r][3]);
                    }
                }
                const float e2k = __ldg(p.e2 + k);
#pragma unroll
                for (int r = 0; r < kFbGroup; r++) {
                    const float dot = __fadd_rn(__fadd_rn(acc[r][0], acc[r][1]), __fadd_rn(acc[r][2], acc[r][3]));
                    merge_min(best_d[r], best_k[r], n_at_min[r], ref_distance(z2[r], e2k, dot), k, 1);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kFbGroup; r++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float d2 = __shfl_xor_sync(0xffffffffu, best_d[r], o);
                const int k2 = __shfl_xor_sync(0xffffffffu, best_k[r], o);
                const int c2 = __shfl_xor_sync(0xffffffffu, n_at_min[r], o);
                merge_min(best_d[r], best_k[r], n_at_min[r], d2, k2, c2);
            }
            if (lane == 0) { sd[r][warp] = best_d[r]; sk[r][warp] = best_k[r]; sn[r][warp] = n_at_min[r]; }
        }
        __syncthreads();
        // thread r finishes row r of the group
        if (tid < kFbGroup && row_s[tid] >= 0) {
            const int r = tid;
            float d = sd[r][0];
            int k = sk[r][0], cnt = sn[r][0];
            for (int v = 1; v < kFbThreads / 32; v++) merge_min(d, k, cnt, sd[r][v], sk[r][v], sn[r][v]);
            if (split) p.part[((int64_t)g * kFbGroup + r) * parts + part0] = make_float4(d, __int_as_float(k), __int_as_float(cnt), 0.0f);
            sd[r][0] = d; sk[r][0] = k; sn[r][0] = cnt;
        }
        if (split) {
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                is_final = (atomicAdd(p.arrive + g, 1u) == (unsigned)parts - 1);
                if (is_final) p.arrive[g] = 0;                   // re-arm for the next call on this workspace
            }
            __syncthreads();
            if (!is_final) continue;
            __threadfence();
        } else {
            __syncthreads();
        }
        if (tid < kFbGroup && row_s[tid] >= 0) {
            const int r = tid;
            const int64_t n = row_s[r];
            float d = sd[r][0];
            int k = sk[r][0], cnt = sn[r][0];
            if (split) {
                d = INFINITY; k = 0x7fffffff; cnt = 0;
                for (int q = 0; q < parts; q++) {
                    const float4 v = __ldcg(p.part + ((int64_t)g * kFbGroup + r) * parts + q);
                    merge_min(d, k, cnt, v.x, __float_as_int(v.y), __float_as_int(v.z));
                }
            }
            // a row can be listed twice (once per epilogue group): the first finisher publishes and counts it
            if (atomicExch(p.out_cnt + 2 * n, -2) != -2) {
                if (k == 0x7fffffff) k = 0;                      // every distance NaN: torch.argmin -> 0 as well
                reinterpret_cast<uint2*>(p.out_q)[n * kOutCap] = make_uint2((uint32_t)(k >> 5), 1u << (k & 31));
                if (p.stats != nullptr) {
                    if (cnt > 1) atomicAdd(p.stats + 0, 1ull);
                    atomicAdd(p.stats + 2, 1ull);
                    atomicAdd(p.stats + 3, (unsigned long long)p.K);
                }
            }
        }
    }
}

}  // namespace vq
