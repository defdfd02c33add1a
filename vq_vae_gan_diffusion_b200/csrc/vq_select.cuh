// vq_select.cuh -- exact fp32 decision + fused forward tail, and the exact full-row fallback.
//
// vq_select_kernel<kForward, kLayout>: one CTA = 32 latents (8 warps, 4 rows per warp), tile in shared memory row-major
// with swizzled 16-byte pieces (vq_common.cuh tile_off).
//   1. forward mode: the fp32 z tile's global loads are issued first (16-byte loads along hw / along d for row-major
//      input); tokeniser mode loads it only when some row of the CTA has more than one candidate
//   2. expand the candidate entries (32-code chunk + code mask) written by the GEMM epilogue -- counts and entries are
//      fetched together, one round trip --, then evaluate every candidate of the rows that need a decision with the
//      reference's fp32 formula in the oracle's canonical accumulation order: the warp's candidates are pooled, four
//      lanes share a (row, code) pair, each lane one canonical partial sum with all its 64 code-row loads independent;
//      first minimum by torch.argmin's rules (NaN is the minimum; lowest index on ties) with redux.sync on distance keys
//   3. kForward only: gather e = E[idx] (codebook.py:85), write z_q = fl(z + fl(e - z)) as NHWC rows
//      (codebook.py:106-109), accumulate sum (e - z)^2 for the loss (codebook.py:96-103) and the usage histogram;
//      16 bytes per lane and request throughout.
// HBM traffic per latent: read z 4D (+ candidates), write idx 8 (+ z_q 4D when kForward).
//
// vq_fallback_kernel: the (rare) rows whose candidate ring overflowed in the GEMM epilogue, and rows / codebooks with
// Inf or NaN, get an exact scan of the whole codebook (see the kernel); the winner is written back as a one-code
// candidate entry so vq_select_kernel finishes the row like any other.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int kSelThreads = 256;
#ifndef VQ_SEL_MIN_BLOCKS
#define VQ_SEL_MIN_BLOCKS 4               // resident CTAs per SM the register allocation is tuned for
#endif
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kMaxCands = 64;           // codes per row the exact stage evaluates (GEMM hands over <= 32 per group)
constexpr int kFbThreads = 128;
constexpr int kFbWarps = kFbThreads / 32;
constexpr int kFbGroup = 8;             // overflowed rows scanned together (they share every code-row load)
constexpr int kFbMaxGroups = 512;       // row groups whose scan is split over code blocks (4096 rows)
constexpr int kFbMaxParts = 256;        // code blocks per group
constexpr int kFbCtasPerSm = 4;         // resident CTAs per SM the kernel is tuned for (128 registers, 27 KiB of shared memory)
constexpr int kFbCodes = 4;             // codes per lane and pass: a lane owns a 4 codes x 8 rows register tile of ONE
                                        // canonical partial sum (32 accumulators), a warp pass covers 32 codes
constexpr int kFbEPitch = 36;           // floats per staged 32-column code-row segment (32 + 4: conflict-free both ways)
constexpr int kFbZPitch = 68;           // floats per (row, partial sum) run of the de-interleaved latents (64 + 4)
constexpr int kFbZRow = 4 * kFbZPitch;  // floats per de-interleaved latent row

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
    void* idx;                 // (N) int64 / int32 / uint16 according to idx_bits
    int idx_bits;
    int recipe;                // kRecipeExpanded / kRecipeDiffSq
    float* zq;                 // (N, D) or null (tokeniser mode; or a forward whose z_q nobody reads)
    float* scat;               // (K, D) or null: per-code sums of (e - z) for the codebook gradient (zeroed before launch)
    unsigned long long* hist;  // (K) or null
    double* loss_partial;      // (gridDim.x) per-CTA sums of (e - z)^2, finished by vq_loss_finalize_kernel
    unsigned long long* stats; // (VQ_STAT_COUNT) or null
};

// One canonical partial sum of the dot product of latent row r (tile row) with code k: the terms d == j (mod 4) in
// ascending d, one fma each (oracle/vq_oracle.c: vqo_dot).  Four lanes (j = 0..3, a "quad") share a (row, code) pair.
// Code-row loads: the quad reads the row as float4s, lane j taking float4 number 4 s + j of step s (64 contiguous bytes per
// quad and step, 16 steps), and a 4 x 4 register transpose inside the quad (two butterfly rounds of shuffles) hands lane
// j the four values d = 16 s + 4 j' + j, j' = 0..3 -- exactly the next four terms of its chain, in order.  A warp request
// then touches 8 lines for 512 useful bytes instead of 8 lines for 128: the scalar version (one 4-byte load per term, 64
// requests of 8 wavefronts each per pass) kept the L1 data pipe busy 70 % of the kernel (ncu, round 1: 224 k of 320 k
// wavefronts per SM came from these loads); the transpose costs 4 shuffles and 12 selects per step on otherwise idle
// issue slots.  The latent values come from the shared-memory tile with 4-byte loads (four rows per warp: one wavefront).
// kDiffSq: the terms are fl(fl(z - e)^2) added one by one (v_vq_diffusion.py:121) instead of fused z * e products.
__device__ __forceinline__ float4 quad_transpose(const float4 a, const bool b0, const bool b1) {
    // lanes = rows j, components = columns c; returns column j of rows 0..3 (x, y, z, w = rows 0, 1, 2, 3)
    const float s1 = b1 ? a.x : a.z, s2 = b1 ? a.y : a.w;                  // round 1: 2 x 2 blocks with lane j ^ 2
    const float r1 = __shfl_xor_sync(0xffffffffu, s1, 2), r2 = __shfl_xor_sync(0xffffffffu, s2, 2);
    const float x0 = b1 ? r1 : a.x, x1 = b1 ? r2 : a.y, x2 = b1 ? a.z : r1, x3 = b1 ? a.w : r2;
    const float t1 = b0 ? x0 : x1, t2 = b0 ? x2 : x3;                      // round 2: single elements with lane j ^ 1
    const float q1 = __shfl_xor_sync(0xffffffffu, t1, 1), q2 = __shfl_xor_sync(0xffffffffu, t2, 1);
    return make_float4(b0 ? q1 : x0, b0 ? x1 : q1, b0 ? q2 : x2, b0 ? x3 : q2);
}

// All 32 lanes must call this (shuffles); lanes whose pair is invalid pass k < 0 and get 0.
template <bool kDiffSq>
__device__ __forceinline__ float exact_partial_tile(const float* tile, int r, const float* __restrict__ E, int k, int j) {
    const bool live = k >= 0;
    const float4* e4 = reinterpret_cast<const float4*>(E + (int64_t)(live ? k : 0) * kD) + j;
    const bool b0 = (j & 1) != 0, b1 = (j & 2) != 0;
    const int g = tile_swz(r);
    const float* zrow = tile + r * kD + j;
    int zo[8];
#pragma unroll
    for (int b = 0; b < 8; b++) zo[b] = (b ^ g) << 2;
    float p = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        float4 v[8];
#pragma unroll
        for (int s = 0; s < 8; s++) v[s] = live ? __ldg(e4 + 4 * (8 * h + s)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const float4 e = quad_transpose(v[s], b0, b1);                 // e_d for d = 16 (8 h + s) + 4 j' + j
            const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int i = 4 * (8 * h + s) + jj;                        // term index of the chain: d = 4 i + j
                const float zv = zrow[32 * (i >> 3) + zo[i & 7]];
                if (kDiffSq) {
                    const float diff = __fsub_rn(zv, ev[jj]);
                    p = __fadd_rn(p, __fmul_rn(diff, diff));
                } else {
                    p = __fmaf_rn(zv, ev[jj], p);
                }
            }
        }
    }
    return live ? p : 0.0f;
}

// Distances as unsigned keys whose integer order is torch.argmin's order (codebook.py:82): NaN -> 0, smaller than every
// number (torch treats a NaN as the minimum and returns the first one); -inf < ... < -0.0 == +0.0 < ... < +inf otherwise.
// Real keys lie in [0x007fffff, 0xff800000]; 0xffffffff is free as the "no candidate" sentinel.
__device__ __forceinline__ uint32_t dist_key(float d) {
    if (!(d == d)) return 0u;
    const uint32_t b = __float_as_uint(d + 0.0f);                          // -0.0 -> +0.0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <bool kForward, int kLayout>
__global__ void __launch_bounds__(kSelThreads, VQ_SEL_MIN_BLOCKS)
vq_select_kernel(const SelectParams p) {
    __shared__ __align__(16) float tile[kSelRows * kD];       // 32 KiB, swizzled row-major (vq_common.cuh tile_off)
    __shared__ int clist[kSelWarps][4][kMaxCands];            // candidate codes of each row of each warp
    __shared__ int idx_s[kSelRows];
    __shared__ double red_s[kSelWarps];
    __shared__ unsigned int st_s[4];                          // per-CTA counters (same-address global atomics are slow)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    if (tid < 4) st_s[tid] = 0;

    // 1a. forward mode always needs the fp32 z tile: put its loads in flight before anything else (16-byte loads along
    //     hw, 8 per thread) so that the candidate expansion below runs in their shadow
    const int dsub = lane >> 3, hq = lane & 7;
    float4 zreg[8];
    constexpr bool kVec = (kLayout == kLayoutVec);
    if (kForward && kVec) {
        const int64_t b = n0 / p.HW, hw0 = n0 % p.HW;
        const float* src = p.z + (b * kD + dsub) * p.HW + hw0 + 4 * hq;
#pragma unroll
        for (int i = 0; i < 8; i++) zreg[i] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)((warp * 8 + i) * 4) * p.HW));
    }

    // PDL (vq_common.cuh): z is an input of the call, so the loads above were issued while the GEMM / fallback kernels were
    // finishing; the candidate lists below are theirs
    pdl_wait();

    // 2a. expand this warp's candidate entries (32-code chunk + code mask) into code lists
    //     (counts and entry slots are fetched together, unconditionally: one DRAM round trip instead of two; the 16
    //     slots of a row are one 128-byte line)
    int nq[4];
    unsigned resolved_mask = 0;                               // rows decided by vq_fallback_kernel (stats counted there)
    int2 cnt2[4];
    uint2 eq[4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t n = n0 + warp * 4 + rr;
        cnt2[rr] = make_int2(0, 0);
        eq[rr] = make_uint2(0u, 0u);
        if (n < p.N) {
            cnt2[rr] = __ldg(reinterpret_cast<const int2*>(p.out_cnt) + n);
            if (lane < kOutCap) eq[rr] = __ldg(reinterpret_cast<const uint2*>(p.out_q) + n * kOutCap + lane);
        }
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t n = n0 + warp * 4 + rr;
        nq[rr] = 0;
        if (n >= p.N) continue;                               // warp-uniform
        const int c0 = cnt2[rr].x, c1 = cnt2[rr].y;
        const bool resolved = (c0 == -2);
        if (resolved) resolved_mask |= 1u << rr;
        const int g = lane >> 3, i = lane & 7;                // slot = lane: group g owns slots [8g, 8g + 8)
        const bool valid = resolved ? (lane == 0) : (lane < kOutCap && i < (g ? c1 : c0));
        uint2 e = eq[rr];                                     // (chunk id, mask of candidate codes in the chunk)
        if (!valid) e = make_uint2(0u, 0u);
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

    // 1b. z tile into shared memory.  Tokeniser mode loads it only when some row of this CTA has more than one
    //     candidate and therefore needs exact distances.
    const int need_exact = (nq[0] > 1) | (nq[1] > 1) | (nq[2] > 1) | (nq[3] > 1);
    const bool want_tile = kForward ? true : (__syncthreads_or(need_exact) != 0);
    if (want_tile) {
        if (kLayout == kLayoutRows) {
            fill_tile_rows(tile, p.z, n0, p.N, warp, lane);
        } else if (kVec) {
            if (!kForward) {
                const float* src = p.z + col_form_origin(n0, p.HW, warp, lane);
#pragma unroll
                for (int i = 0; i < 8; i++) zreg[i] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(4 * i) * p.HW));
            }
            const ColForm cf(warp, lane);
#pragma unroll
            for (int i = 0; i < 8; i++) cf.store(tile, i, zreg[i]);
        } else {
            const int64_t n = n0 + lane;
            const bool ok = n < p.N;
            const int64_t b = ok ? n / p.HW : 0, hw = ok ? n % p.HW : 0;
            const float* src = p.z + (b * kD) * p.HW + hw;
#pragma unroll 8
            for (int i = 0; i < kD / 8; i++) {
                const int d = warp + 8 * i;
                tile[tile_off(lane, d)] = ok ? __ldg(src + (int64_t)d * p.HW) : 0.0f;
            }
        }
    }
    __syncthreads();

    // 2b. exact distances.  The candidates of the warp's rows that need a decision (>= 2 candidates) are pooled into
    //     one list and dealt out one (row, code) pair per lane, so a pass keeps all 32 lanes busy whatever the split
    //     between rows.  Per pass and row: first minimum over the row's lanes with two redux.sync (distance key, then
    //     lowest index at that key) and a ballot for the multiplicity; lane rr keeps the running result of row rr.
    {
        int off[5];
        off[0] = 0;
#pragma unroll
        for (int rr = 0; rr < 4; rr++) off[rr + 1] = off[rr] + (nq[rr] > 1 ? nq[rr] : 0);
        const int total = off[4];
        // running result of row rr of this warp, kept by lane rr (the per-pass minima below are warp-uniform)
        uint32_t best_u = 0xffffffffu;
        int best_k = 0x7fffffff, n_at_min = 0;
        const int my_nq = (lane == 0) ? nq[0] : (lane == 1) ? nq[1] : (lane == 2) ? nq[2] : nq[3];
        if (lane < 4 && my_nq == 1) {
            // a single candidate is decided without any arithmetic: the margin argument guarantees that the
            // oracle's argmin is among the candidates
            best_k = clist[warp][lane][0];
            if (best_k >= p.K) best_k = 0;
            n_at_min = 1;
        }
        for (int base = 0; base < total; base += 8) {
            const int f = base + (lane >> 2), j = lane & 3;    // pair index, canonical partial sum of this lane
            const bool active = f < total;
            const int rr = (f >= off[3]) ? 3 : (f >= off[2]) ? 2 : (f >= off[1]) ? 1 : 0;
            const int pos = f - ((rr == 3) ? off[3] : (rr == 2) ? off[2] : (rr == 1) ? off[1] : 0);
            int k = active ? clist[warp][rr][pos] : -1;
            if (k >= p.K) k = -1;                              // pad codes of the last chunk
            const int r = warp * 4 + rr;
            // (every lane calls: the code-row transpose inside the quad uses shuffles)
            const float pj = (p.recipe == kRecipeDiffSq) ? exact_partial_tile<true>(tile, r, p.E, k, j)
                                                         : exact_partial_tile<false>(tile, r, p.E, k, j);
            const float dot = combine4(pj);                    // (p0 + p1) + (p2 + p3) on all four lanes
            const bool lead = (k >= 0) && (j == 0);
            uint32_t u = 0xffffffffu;
            // (requesting the two norms BEFORE the fma chains, to take one dependent round trip out of the pass, was measured:
            // +180 us on cfg4 -- profiles/r2_ab_select_tail.jsonl)
            if (lead) u = dist_key(p.recipe == kRecipeDiffSq ? dot : ref_distance(__ldg(p.z2 + n0 + r), __ldg(p.e2 + k), dot));
#pragma unroll
            for (int r2 = 0; r2 < 4; r2++) {
                const bool mine = lead && (rr == r2);
                if (__ballot_sync(0xffffffffu, mine) == 0u) continue;          // warp-uniform
                const uint32_t um = __reduce_min_sync(0xffffffffu, mine ? u : 0xffffffffu);
                const bool at = mine && (u == um);
                const int km = (int)__reduce_min_sync(0xffffffffu, at ? (uint32_t)k : 0x7fffffffu);
                const int c = __popc(__ballot_sync(0xffffffffu, at));
                if (lane == r2) {
                    if (um < best_u) { best_u = um; best_k = km; n_at_min = c; }
                    else if (um == best_u) { n_at_min += c; best_k = min(best_k, km); }
                }
            }
        }
        // lane rr publishes row rr of this warp
        if (lane < 4) {
            const int rr = lane;
            const int r = warp * 4 + rr;
            const int64_t n = n0 + r;
            int bk = best_k;
            const int na = n_at_min;
            if (bk == 0x7fffffff) bk = 0;                     // no candidate at all (cannot happen for a row the GEMM listed)
            if (n < p.N) {
                idx_s[r] = bk;
                if (p.idx_bits == 64) reinterpret_cast<int64_t*>(p.idx)[n] = (int64_t)bk;
                else if (p.idx_bits == 32) reinterpret_cast<int32_t*>(p.idx)[n] = bk;
                else reinterpret_cast<uint16_t*>(p.idx)[n] = (uint16_t)bk;
                if (p.stats != nullptr && !((resolved_mask >> rr) & 1u)) {
                    if (na > 1) atomicAdd(&st_s[0], 1u);
                    if (my_nq > 1) atomicAdd(&st_s[1], 1u);
                    atomicAdd(&st_s[3], (unsigned)my_nq);
                }
            }
        }
    }
    if (!kForward) {
        __syncthreads();
        if (p.stats != nullptr && tid < 4 && st_s[tid] != 0) atomicAdd(p.stats + tid, (unsigned long long)st_s[tid]);
        return;
    }
    // No CTA-wide barrier here: a warp's tail only needs its OWN rows' decisions (idx_s entries written by its own lanes),
    // so warps without re-rank work stream their z_q rows while others are still evaluating candidates (ncu, round 2:
    // 3.9 warps per issue slot were parked at this barrier).  The per-CTA counters are flushed after the loss barrier below.
    __syncwarp();

    // 3. forward tail, 16 bytes per lane and request: lane <-> d in [4 lane, 4 lane + 4) and [128 + 4 lane, ...); all four
    //    code rows of this warp are requested before the first one is consumed
    float sq = 0.0f;
    {
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
            if (kk[rr] < 0) continue;
            const int r = warp * 4 + rr;
            const float4* zrow4 = reinterpret_cast<const float4*>(tile + r * kD);
            const int g = tile_swz(r);
            float4* out4 = reinterpret_cast<float4*>(p.zq + (n0 + r) * kD);
            const bool want_zq = p.zq != nullptr;             // a caller that folds post_quant_conv needs only idx / loss / hist
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float4 zv = zrow4[(lane + 32 * h) ^ g];
                const float4 e = ev[rr][h];
                float4 diff, o;
                diff.x = __fsub_rn(e.x, zv.x); diff.y = __fsub_rn(e.y, zv.y);     // fl(e - z)
                diff.z = __fsub_rn(e.z, zv.z); diff.w = __fsub_rn(e.w, zv.w);
                o.x = __fadd_rn(zv.x, diff.x); o.y = __fadd_rn(zv.y, diff.y);     // fl(z + fl(e - z)), codebook.py:106
                o.z = __fadd_rn(zv.z, diff.z); o.w = __fadd_rn(zv.w, diff.w);
                if (want_zq) __stcs(out4 + lane + 32 * h, o);
                if (p.scat != nullptr) red_add_v4(p.scat + (int64_t)kk[rr] * kD + 4 * (lane + 32 * h), diff.x, diff.y, diff.z, diff.w);
                sq = __fmaf_rn(diff.x, diff.x, sq); sq = __fmaf_rn(diff.y, diff.y, sq);
                sq = __fmaf_rn(diff.z, diff.z, sq); sq = __fmaf_rn(diff.w, diff.w, sq);
            }
            if (p.hist != nullptr && lane == 0) atomicAdd(p.hist + kk[rr], 1ull);
        }
    }
    // loss: fp32 per thread (<= 32 terms), fp64 from there on; the per-CTA partial goes to global memory and
    // vq_loss_finalize_kernel (next in the chain) adds the partials in a fixed order.  (Until round 2 the last CTA to arrive did
    // that, which put a fence + an atomic round trip at the end of EVERY CTA's latency chain.)
    double v = (double)sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red_s[warp] = v;
    __syncthreads();
    if (p.stats != nullptr && tid >= 32 && tid < 36 && st_s[tid - 32] != 0) atomicAdd(p.stats + (tid - 32), (unsigned long long)st_s[tid - 32]);
    if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < kSelWarps; w++) sum += red_s[w];
        p.loss_partial[blockIdx.x] = sum;
    }
}

// loss = mean((e - z)^2) * (1 + beta) (codebook.py:96-103: mean(a + beta * mean(b)) with a == b elementwise) from the per-CTA
// partial sums of vq_select_kernel, added in a fixed order (run-to-run reproducible): one CTA, chained behind the select kernel.
__global__ void __launch_bounds__(kSelThreads)
vq_loss_finalize_kernel(const double* __restrict__ loss_partial, unsigned n_parts, int64_t N, float beta, float* __restrict__ loss) {
    __shared__ double red_s[kSelWarps];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_wait();                                               // the select kernel has completed: every partial is visible
    double sum = 0.0;
    for (unsigned i = tid; i < n_parts; i += kSelThreads) sum += __ldcg(loss_partial + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red_s[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kSelWarps; w++) tot += red_s[w];
        const double m = tot / ((double)N * (double)kD);
        *loss = (float)(m + (double)beta * m);
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
    int recipe;                // kRecipeExpanded / kRecipeDiffSq
    int32_t* out_cnt;
    uint32_t* out_q;
    float4* part;              // (kFbMaxGroups * kFbGroup, parts) partial (distance, index, multiplicity) results
    unsigned int* arrive;      // (kFbMaxGroups) arrival counters, zero on entry, re-armed by the kernel
    unsigned long long* stats;
};

// merge (distance key, first index, multiplicity) triples: lexicographic minimum, multiplicities of equal minima add up
// (keys: dist_key -- a NaN distance is the smallest key, as in torch.argmin)
__device__ __forceinline__ void merge_min(uint32_t& d, int& k, int& c, uint32_t d2, int k2, int c2) {
    if (d2 < d) { d = d2; k = k2; c = c2; }
    else if (d2 == d) { c += c2; k = min(k, k2); }
}

// Work item = (group of kFbGroup worklist entries, block of codes).  The scan is bound by shared-memory -> register
// bandwidth (128 B per cycle and SM), so it is register-tiled: lane <-> (code block cb = lane >> 2, canonical partial
// sum j = lane & 3) owns the terms d == j (mod 4) of FOUR codes (kb + cb + 8 t) for all EIGHT rows of the group -- 32
// accumulators; per 16 columns it reads its codes' values with 4-byte loads (16 of them) and each row's four terms with
// one 16-byte load (8 of them) for 128 FMAs, about a third of the shared-memory traffic per FMA of a one-code-per-lane
// scheme (two of those were built first; ncu: l1tex 49-55 % busy, short-scoreboard stalls on top).  A warp
// pass covers 32 codes: their rows are fetched coalesced, 32 columns at a time (8 lanes x 16 bytes per row, the next segment in flight
// while the current one is used), into a per-warp buffer whose 144-byte pitch makes both the 16-byte fills and the
// 4-byte reads conflict-free; the group's latent rows sit in shared memory DE-INTERLEAVED by j, which is what turns
// four consecutive terms of a partial sum into one 16-byte load (broadcast to the 8 lanes of the same j).  After a
// pass the partial sums of every (row, code) pair are combined as (p0 + p1) + (p2 + p3), each lane reduces its four
// codes, the warp reduces with redux.sync on distance keys, and lane r keeps row r's running result.  The number of
// code blocks per group is chosen ON THE DEVICE from the worklist length so that groups x blocks fills the resident
// CTAs about once: a handful of overflowed rows is spread over the whole chip, while a degenerate codebook (every row
// overflows) gets one CTA per group scanning all codes with no merge step.  The last code block of a group to arrive
// merges the per-block minima (a warp per row).
template <bool kDiffSq>
__global__ void __launch_bounds__(kFbThreads, kFbCtasPerSm)
vq_fallback_kernel(const FallbackParams p) {
    __shared__ __align__(16) float zt[kFbGroup * kFbZRow];                    // 8.5 KiB: zt[r][j][i] = z[row r][4 i + j]
    __shared__ __align__(16) float stage[kFbWarps][32 * kFbEPitch];           // 18 KiB: per-warp [32 codes][32 d]
    __shared__ float z2_s[kFbGroup];
    __shared__ uint32_t sd[kFbGroup][kFbWarps];
    __shared__ int sk[kFbGroup][kFbWarps], sn[kFbGroup][kFbWarps];
    __shared__ int64_t row_s[kFbGroup], zoff_s[kFbGroup];
    __shared__ int is_final;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();                                            // PDL (vq_common.cuh): vq_select_kernel may be scheduled ...
    pdl_wait();                                               // ... and the GEMM's worklist is complete
    const int count = __ldcg(p.fb_count);
    if (count == 0) return;
    const int groups = (count + kFbGroup - 1) / kFbGroup;
    // code blocks per group: fill the grid once; a multiple of one CTA pass (128 codes) each; never more than kFbMaxParts
    constexpr int kPass = kFbWarps * 32;
    int parts = max(1, min(min(kFbMaxParts, (int)gridDim.x / groups), (p.K + kPass - 1) / kPass));
    if (groups > kFbMaxGroups) parts = 1;
    const int per_part = ((p.K + parts * kPass - 1) / (parts * kPass)) * kPass;
    parts = (p.K + per_part - 1) / per_part;
    const bool split = parts > 1;
    const int64_t items = (int64_t)groups * parts;
    const int cb = lane >> 2, j = lane & 3;                   // compute role: code block, canonical partial sum
    const int lc = lane >> 3, ls = lane & 7;                  // staging role: request i fetches code rows 4 i + lc
    float* const stg = &stage[warp][0];
    for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int g = (int)(w / parts), part0 = (int)(w % parts);
        __syncthreads();
        if (tid < kFbGroup) {
            const int64_t n = (g * kFbGroup + tid < count) ? (int64_t)__ldg(p.fb_rows + g * kFbGroup + tid) : -1;
            row_s[tid] = n;
            zoff_s[tid] = (n >= 0) ? (n / p.HW) * kD * p.HW + n % p.HW : -1;      // element (row n, d = 0) of the NCHW latents
            z2_s[tid] = (n >= 0) ? __ldg(p.z2 + n) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kFbGroup; r++) {                  // thread <-> d, d + 128; all loads independent
            const int64_t o = zoff_s[r];
#pragma unroll
            for (int h = 0; h < kD / kFbThreads; h++) {
                const int d = tid + kFbThreads * h;
                zt[r * kFbZRow + (d & 3) * kFbZPitch + (d >> 2)] = (o >= 0) ? __ldg(p.z + o + (int64_t)d * p.HW) : 0.0f;
            }
        }
        __syncthreads();
        uint32_t best_d = 0xffffffffu;                        // lane r < kFbGroup: this warp's running result of row r
        int best_k = 0x7fffffff, best_c = 0;
        const int k_lo = part0 * per_part, k_hi = min(p.K, k_lo + per_part);
        float4 pre[8];
        auto fetch = [&](int kb, int db) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int kk = kb + 4 * i + lc;
                pre[i] = (kk < k_hi) ? __ldg(reinterpret_cast<const float4*>(p.E + (int64_t)kk * kD + 32 * db) + ls)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int kb = k_lo + warp * 32;
        if (kb < k_hi) fetch(kb, 0);
        for (; kb < k_hi; kb += kPass) {
            float acc[kFbCodes][kFbGroup];
#pragma unroll
            for (int t = 0; t < kFbCodes; t++)
#pragma unroll
                for (int r = 0; r < kFbGroup; r++) acc[t][r] = 0.0f;
#pragma unroll 1
            for (int db = 0; db < kD / 32; db++) {
#pragma unroll
                for (int i = 0; i < 8; i++) *reinterpret_cast<float4*>(stg + (4 * i + lc) * kFbEPitch + 4 * ls) = pre[i];
                __syncwarp();
                if (db < kD / 32 - 1) fetch(kb, db + 1);      // next segment (or the next pass's first) in flight
                else if (kb + kPass < k_hi) fetch(kb + kPass, 0);
                const float* zq = zt + j * kFbZPitch + 8 * db;
                const float* eq = stg + cb * kFbEPitch + j;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float e[kFbCodes][4];
#pragma unroll
                    for (int t = 0; t < kFbCodes; t++)
#pragma unroll
                        for (int u = 0; u < 4; u++) e[t][u] = eq[8 * t * kFbEPitch + 4 * (4 * h + u)];
#pragma unroll
                    for (int r = 0; r < kFbGroup; r++) {
                        const float4 zv = *reinterpret_cast<const float4*>(zq + r * kFbZRow + 4 * h);
#pragma unroll
                        for (int t = 0; t < kFbCodes; t++) {
                            if (kDiffSq) {
                                const float d0 = __fsub_rn(zv.x, e[t][0]), d1 = __fsub_rn(zv.y, e[t][1]);
                                const float d2 = __fsub_rn(zv.z, e[t][2]), d3 = __fsub_rn(zv.w, e[t][3]);
                                acc[t][r] = __fadd_rn(acc[t][r], __fmul_rn(d0, d0));
                                acc[t][r] = __fadd_rn(acc[t][r], __fmul_rn(d1, d1));
                                acc[t][r] = __fadd_rn(acc[t][r], __fmul_rn(d2, d2));
                                acc[t][r] = __fadd_rn(acc[t][r], __fmul_rn(d3, d3));
                            } else {
                                acc[t][r] = __fmaf_rn(zv.x, e[t][0], acc[t][r]);
                                acc[t][r] = __fmaf_rn(zv.y, e[t][1], acc[t][r]);
                                acc[t][r] = __fmaf_rn(zv.z, e[t][2], acc[t][r]);
                                acc[t][r] = __fmaf_rn(zv.w, e[t][3], acc[t][r]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            // this pass's minimum per row: over the lane's four codes (ascending k), then over the warp
            float e2k[kFbCodes];
#pragma unroll
            for (int t = 0; t < kFbCodes; t++) {
                const int k = kb + cb + 8 * t;
                e2k[t] = (!kDiffSq && k < k_hi) ? __ldg(p.e2 + k) : 0.0f;
            }
#pragma unroll
            for (int r = 0; r < kFbGroup; r++) {
                uint32_t u_loc = 0xffffffffu;
                int k_loc = 0x7fffffff, c_loc = 0;
#pragma unroll
                for (int t = 0; t < kFbCodes; t++) {
                    const float dot = combine4(acc[t][r]);    // (p0 + p1) + (p2 + p3) on all four lanes
                    const int k = kb + cb + 8 * t;
                    if (j == 0 && k < k_hi)
                        merge_min(u_loc, k_loc, c_loc, dist_key(kDiffSq ? dot : ref_distance(z2_s[r], e2k[t], dot)), k, 1);
                }
                const uint32_t um = __reduce_min_sync(0xffffffffu, u_loc);
                const bool at = (c_loc > 0) && (u_loc == um);
                const int km = (int)__reduce_min_sync(0xffffffffu, at ? (uint32_t)k_loc : 0x7fffffffu);
                const int cn = (int)__reduce_add_sync(0xffffffffu, at ? (uint32_t)c_loc : 0u);
                if (lane == r) merge_min(best_d, best_k, best_c, um, km, cn);
            }
        }
        // block-level merge of the per-warp results: thread r finishes row r
        if (lane < kFbGroup) { sd[lane][warp] = best_d; sk[lane][warp] = best_k; sn[lane][warp] = best_c; }
        __syncthreads();
        if (tid < kFbGroup) {
            uint32_t d = sd[tid][0];
            int k = sk[tid][0], cnt = sn[tid][0];
            for (int v = 1; v < kFbWarps; v++) merge_min(d, k, cnt, sd[tid][v], sk[tid][v], sn[tid][v]);
            sd[tid][0] = d; sk[tid][0] = k; sn[tid][0] = cnt;
        }
        if (split) {
            if (tid < kFbGroup && row_s[tid] >= 0)
                p.part[((int64_t)g * kFbGroup + tid) * parts + part0] =
                    make_float4(__uint_as_float(sd[tid][0]), __int_as_float(sk[tid][0]), __int_as_float(sn[tid][0]), 0.0f);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                is_final = (atomicAdd(p.arrive + g, 1u) == (unsigned)parts - 1);
                if (is_final) p.arrive[g] = 0;                   // re-arm for the next call on this workspace
            }
            __syncthreads();
            if (!is_final) continue;
            __threadfence();
            // the last block to arrive merges the per-block results of the group: a warp per row, lanes over the blocks
            for (int r = warp; r < kFbGroup; r += kFbWarps) {
                if (row_s[r] < 0) continue;                      // warp-uniform
                uint32_t d = 0xffffffffu;
                int k = 0x7fffffff, cnt = 0;
                for (int q = lane; q < parts; q += 32) {
                    const float4 v = __ldcg(p.part + ((int64_t)g * kFbGroup + r) * parts + q);
                    merge_min(d, k, cnt, __float_as_uint(v.x), __float_as_int(v.y), __float_as_int(v.z));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint32_t d2 = __shfl_xor_sync(0xffffffffu, d, o);
                    const int k2 = __shfl_xor_sync(0xffffffffu, k, o);
                    const int c2 = __shfl_xor_sync(0xffffffffu, cnt, o);
                    merge_min(d, k, cnt, d2, k2, c2);
                }
                if (lane == 0) { sd[r][0] = d; sk[r][0] = k; sn[r][0] = cnt; }
            }
        }
        __syncthreads();
        if (tid < kFbGroup && row_s[tid] >= 0) {
            const int64_t n = row_s[tid];
            int k = sk[tid][0];
            const int cnt = sn[tid][0];
            // publish the winner as a one-code candidate entry (count -2 = "decided"): the select kernel finishes the row
            if (atomicExch(p.out_cnt + 2 * n, -2) != -2) {
                if (k == 0x7fffffff) k = 0;                      // (K >= 1: cannot happen)
                reinterpret_cast<uint2*>(p.out_q)[n * kOutCap] = make_uint2((uint32_t)(k >> 5), 1u << (k & 31));
                if (p.stats != nullptr) {
                    if (cnt > 1) atomicAdd(p.stats + 0, 1ull);
                    atomicAdd(p.stats + 2, 1ull);
                }
            }
        }
    }
}

}  // namespace vq
