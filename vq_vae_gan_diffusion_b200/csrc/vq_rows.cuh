// vq_rows.cuh -- exact stage of the row-major nearest-code searches at any width D <= 512 and for all three distance
// recipes of the reference:
//   kRecipeExpanded  |x|^2 + |e|^2 - 2 x.e                    diffusion_gaussian2d.py:334-339 (the CodeBook's formula)
//   kRecipeDiffSq    sum_d (x_d - e_d)^2                      v_vq_diffusion.py:114-123
//   kRecipeCdist     torch.cdist(normalize(x), normalize(T))  diffusion_gaussian3d.py:543-570
// The candidates come from the same tcgen05 distance GEMM (vq_argmin_sm100.cuh, contraction over kNC chunks of 64) fed by
// vq_prep_rows_kernel; this file holds what decides among them in fp32, in the oracle's canonical order:
//   vq_select_rows_kernel    candidate expansion + exact distances + first minimum (no shared-memory tile: the quad of a
//                            (row, code) pair reads both rows as float4s and transposes them in registers)
//   vq_fallback_rows_kernel  exact scan of the whole table for the rows the GEMM flagged (more than 64 candidates, Inf / NaN)
//
// torch.cdist's default path for more than 25 rows (_euclidean_dist, ATen/native/Distance.cpp) evaluates the squared distance
// as ONE matrix product of augmented vectors, [-2 x, |x|^2, 1] . [y, 1, |y|^2], followed by clamp_min(0).sqrt(); the canonical
// restatement runs the same D + 2 terms through the four fma chains (term d goes to chain d mod 4, ascending d), so the two
// norm terms are the LAST terms of chains D mod 4 and (D + 1) mod 4.  sqrt is monotone but merges neighbouring fp32 values,
// so the first minimum is taken over the square roots, exactly as argmin(cdist(...)) does.
#pragma once
#include "vq_common.cuh"
#include "vq_select.cuh"

namespace vq {

struct SelectRowsParams {
    const float* x;            // (N, D) fp32 query rows
    const float* denom;        // (N) L2-normalisation divisors (vq_prep_rows_kernel) or null: rows used as they are
    const float* E;            // (K, D) fp32 table (for kRecipeCdist: the normalised table)
    const float* e2;           // (K_pad) canonical |e_k|^2
    const float* z2;           // (N)     canonical |x_n|^2 (of the normalised row when denom is given)
    int32_t* out_cnt;          // (N, 2)  candidate counts per epilogue group (vq_argmin_sm100.cuh); -2: decided by the fallback
    uint32_t* out_q;           // (N, kOutCap, 2)
    int64_t N;
    int D, K;
    void* idx;                 // (N) int64 / int32 / uint16
    int idx_bits;
    int recipe;
    unsigned long long* stats;
    // fallback only
    const int32_t* fb_rows;
    const int32_t* fb_count;
    float4* part;              // (kFbMaxGroups * kFbGroup, parts) per-block (key, index, multiplicity)
    unsigned int* arrive;      // (kFbMaxGroups * kFbGroup) arrival counters, zero on entry, re-armed
};

// float4 number `f` of a row of width D (zero beyond the row's end).  vec: D % 4 == 0 and a 16-byte aligned base.
__device__ __forceinline__ float4 row_piece(const float* __restrict__ row, int f, int D, bool vec) {
    if (vec) return (4 * f < D) ? __ldg(reinterpret_cast<const float4*>(row) + f) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v;
    v.x = (4 * f + 0 < D) ? __ldg(row + 4 * f + 0) : 0.0f;
    v.y = (4 * f + 1 < D) ? __ldg(row + 4 * f + 1) : 0.0f;
    v.z = (4 * f + 2 < D) ? __ldg(row + 4 * f + 2) : 0.0f;
    v.w = (4 * f + 3 < D) ? __ldg(row + 4 * f + 3) : 0.0f;
    return v;
}

// Distance key of pair (query row xr, table row er) as seen by the quad's lane j; all 32 lanes call (shuffles), lanes of
// a dead pair pass live = false.  Returns the distance on all four lanes of the quad.
template <int kNC>
__device__ __forceinline__ float rows_pair_distance(const float* __restrict__ xr, const float* __restrict__ er, bool live, int j,
                                                    int D, bool vec_x, bool vec_e, float dn, bool normalize, int recipe,
                                                    float z2n, float e2k) {
    const bool b0 = (j & 1) != 0, b1 = (j & 2) != 0;
    float p = 0.0f;
#pragma unroll 2
    for (int s = 0; s < 4 * kNC; s++) {                       // 16 columns per step: the quad reads 64 contiguous bytes of each row
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), ev = xv;
        if (live && 16 * s < D) {
            xv = row_piece(xr, 4 * s + j, D, vec_x);
            ev = row_piece(er, 4 * s + j, D, vec_e);
        }
        if (normalize) {                                       // x / max(|x|, 1e-12), the value vq_prep_rows_kernel produced
            xv.x = __fdiv_rn(xv.x, dn); xv.y = __fdiv_rn(xv.y, dn); xv.z = __fdiv_rn(xv.z, dn); xv.w = __fdiv_rn(xv.w, dn);
        }
        const float4 xt = quad_transpose(xv, b0, b1), et = quad_transpose(ev, b0, b1);   // columns 16 s + 4 j' + j, j' = 0..3
        const float xa[4] = {xt.x, xt.y, xt.z, xt.w}, ea[4] = {et.x, et.y, et.z, et.w};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            if (recipe == kRecipeDiffSq) {
                const float diff = __fsub_rn(xa[jj], ea[jj]);
                p = __fadd_rn(p, __fmul_rn(diff, diff));
            } else if (recipe == kRecipeCdist) {
                p = __fmaf_rn(-2.0f * xa[jj], ea[jj], p);      // (-2 x) is exact
            } else {
                p = __fmaf_rn(xa[jj], ea[jj], p);
            }
        }
    }
    if (recipe == kRecipeCdist) {
        if (j == (D & 3)) p = __fadd_rn(p, z2n);               // fma(|x|^2, 1, p): term D of the augmented product
        if (j == ((D + 1) & 3)) p = __fadd_rn(p, e2k);         // fma(1, |y|^2, p): term D + 1
    }
    const float dot = combine4(p);
    if (recipe == kRecipeDiffSq) return dot;
    if (recipe == kRecipeCdist) return sqrtf(dot < 0.0f ? 0.0f : dot);    // clamp_min(0).sqrt(); a NaN stays a NaN
    return ref_distance(z2n, e2k, dot);
}

// One CTA = 32 rows (8 warps x 4 rows).  Rows with a single candidate are decided without arithmetic (the margin argument
// guarantees that the oracle's argmin is among the candidates); the others pool their (row, code) pairs per warp, one pair
// per quad and pass.
template <int kNC>
__global__ void __launch_bounds__(kSelThreads)
vq_select_rows_kernel(const SelectRowsParams p) {
    __shared__ int clist[kSelWarps][4][kMaxCands];
    __shared__ unsigned int st_s[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * kSelRows;
    if (tid < 4) st_s[tid] = 0;
    pdl_wait();                                               // PDL (vq_common.cuh): the candidate lists are the GEMM's / fallback's

    int nq[4];
    unsigned resolved_mask = 0;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t n = n0 + warp * 4 + rr;
        nq[rr] = 0;
        if (n >= p.N) continue;                               // warp-uniform
        const int2 cnt2 = __ldcg(reinterpret_cast<const int2*>(p.out_cnt) + n);
        uint2 e = make_uint2(0u, 0u);
        if (lane < kOutCap) e = __ldcg(reinterpret_cast<const uint2*>(p.out_q) + n * kOutCap + lane);
        const int c0 = cnt2.x, c1 = cnt2.y;
        const bool resolved = (c0 == -2);
        if (resolved) resolved_mask |= 1u << rr;
        const int g = lane >> 3, i = lane & 7;
        const bool valid = resolved ? (lane == 0) : (lane < kOutCap && i < (g ? c1 : c0));
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
    __syncwarp();

    int off[5];
    off[0] = 0;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) off[rr + 1] = off[rr] + (nq[rr] > 1 ? nq[rr] : 0);
    const int total = off[4];
    uint32_t best_u = 0xffffffffu;                            // lane rr < 4 keeps the running result of row rr of this warp
    int best_k = 0x7fffffff, n_at_min = 0;
    const int my_nq = (lane == 0) ? nq[0] : (lane == 1) ? nq[1] : (lane == 2) ? nq[2] : nq[3];
    if (lane < 4 && my_nq == 1) {
        best_k = clist[warp][lane][0];
        if (best_k >= p.K) best_k = 0;
        n_at_min = 1;
    }
    const bool vec_x = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const bool vec_e = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0);
    for (int base = 0; base < total; base += 8) {
        const int f = base + (lane >> 2), j = lane & 3;
        const bool active = f < total;
        const int rr = (f >= off[3]) ? 3 : (f >= off[2]) ? 2 : (f >= off[1]) ? 1 : 0;
        const int pos = f - ((rr == 3) ? off[3] : (rr == 2) ? off[2] : (rr == 1) ? off[1] : 0);
        int k = active ? clist[warp][rr][pos] : -1;
        if (k >= p.K) k = -1;                                  // pad codes of the last chunk
        const int64_t n = n0 + warp * 4 + rr;
        const bool live = k >= 0;
        const float dn = (live && p.denom != nullptr) ? __ldg(p.denom + n) : 1.0f;
        const float z2n = live ? __ldg(p.z2 + n) : 0.0f, e2k = live ? __ldg(p.e2 + k) : 0.0f;
        const float dist = rows_pair_distance<kNC>(p.x + (live ? n : 0) * (int64_t)p.D, p.E + (int64_t)(live ? k : 0) * p.D, live, j,
                                                   p.D, vec_x, vec_e, dn, p.denom != nullptr, p.recipe, z2n, e2k);
        const bool lead = live && (j == 0);
        const uint32_t u = lead ? dist_key(dist) : 0xffffffffu;
#pragma unroll
        for (int r2 = 0; r2 < 4; r2++) {
            const bool mine = lead && (rr == r2);
            if (__ballot_sync(0xffffffffu, mine) == 0u) continue;              // warp-uniform
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
    if (lane < 4) {
        const int64_t n = n0 + warp * 4 + lane;
        int bk = best_k;
        if (bk == 0x7fffffff) bk = 0;
        if (n < p.N) {
            if (p.idx_bits == 64) reinterpret_cast<int64_t*>(p.idx)[n] = (int64_t)bk;
            else if (p.idx_bits == 32) reinterpret_cast<int32_t*>(p.idx)[n] = bk;
            else reinterpret_cast<uint16_t*>(p.idx)[n] = (uint16_t)bk;
            if (p.stats != nullptr && !((resolved_mask >> lane) & 1u)) {
                if (n_at_min > 1) atomicAdd(&st_s[0], 1u);
                if (my_nq > 1) atomicAdd(&st_s[1], 1u);
                atomicAdd(&st_s[3], (unsigned)my_nq);
            }
        }
    }
    __syncthreads();
    if (p.stats != nullptr && tid < 4 && st_s[tid] != 0) atomicAdd(p.stats + tid, (unsigned long long)st_s[tid]);
}

// Exact scan for the flagged rows.  Work item = (worklist entry, block of codes); a warp pass covers 8 codes (one per quad),
// a CTA pass 64.  The last block of a row to arrive merges the per-block (key, first index, multiplicity) triples and
// publishes the winner as a one-code candidate entry (count -2 = "decided") for vq_select_rows_kernel.  This is the rare
// path (rows with more than 64 candidates or with Inf / NaN), written for simplicity, not speed.
template <int kNC>
__global__ void __launch_bounds__(kSelThreads)
vq_fallback_rows_kernel(const SelectRowsParams p) {
    __shared__ uint32_t sd[kSelWarps];
    __shared__ int sk[kSelWarps], sn[kSelWarps];
    __shared__ int is_final;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();
    pdl_wait();
    const int count = __ldcg(p.fb_count);
    if (count == 0) return;
    constexpr int kPass = kSelWarps * 8;
    const int max_rows = kFbMaxGroups * kFbGroup;
    int parts = max(1, min(min(kFbMaxParts, (int)gridDim.x / count), (p.K + kPass - 1) / kPass));
    if (count > max_rows) parts = 1;
    const int per_part = ((p.K + parts * kPass - 1) / (parts * kPass)) * kPass;
    parts = (p.K + per_part - 1) / per_part;
    const bool vec_x = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const bool vec_e = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0);
    const int j = lane & 3, quad = lane >> 2;
    for (int64_t w = blockIdx.x; w < (int64_t)count * parts; w += gridDim.x) {
        const int item = (int)(w / parts), part0 = (int)(w % parts);
        const int64_t n = __ldcg(p.fb_rows + item);
        const float dn = (p.denom != nullptr) ? __ldg(p.denom + n) : 1.0f;
        const float z2n = __ldg(p.z2 + n);
        const int k_lo = part0 * per_part, k_hi = min(p.K, k_lo + per_part);
        uint32_t best_d = 0xffffffffu;
        int best_k = 0x7fffffff, best_c = 0;
        for (int kb = k_lo + warp * 8; kb < k_hi; kb += kPass) {
            const int k = kb + quad;
            const bool live = k < k_hi;
            const float e2k = live ? __ldg(p.e2 + k) : 0.0f;
            const float dist = rows_pair_distance<kNC>(p.x + n * (int64_t)p.D, p.E + (int64_t)(live ? k : 0) * p.D, live, j, p.D,
                                                       vec_x, vec_e, dn, p.denom != nullptr, p.recipe, z2n, e2k);
            if (live && j == 0) merge_min(best_d, best_k, best_c, dist_key(dist), k, 1);
        }
        // warp, then CTA reduction of (key, first index, multiplicity)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t d2 = __shfl_xor_sync(0xffffffffu, best_d, o);
            const int k2 = __shfl_xor_sync(0xffffffffu, best_k, o), c2 = __shfl_xor_sync(0xffffffffu, best_c, o);
            merge_min(best_d, best_k, best_c, d2, k2, c2);
        }
        __syncthreads();
        if (lane == 0) { sd[warp] = best_d; sk[warp] = best_k; sn[warp] = best_c; }
        __syncthreads();
        if (tid == 0) {
            uint32_t d = sd[0];
            int k = sk[0], c = sn[0];
            for (int v = 1; v < kSelWarps; v++) merge_min(d, k, c, sd[v], sk[v], sn[v]);
            bool final_block = true;
            if (parts > 1) {
                p.part[(int64_t)item * parts + part0] = make_float4(__uint_as_float(d), __int_as_float(k), __int_as_float(c), 0.0f);
                __threadfence();
                final_block = (atomicAdd(p.arrive + item, 1u) == (unsigned)parts - 1);
                if (final_block) {
                    p.arrive[item] = 0;                          // re-arm for the next call on this workspace
                    __threadfence();
                    d = 0xffffffffu; k = 0x7fffffff; c = 0;
                    for (int q = 0; q < parts; q++) {
                        const float4 v = __ldcg(p.part + (int64_t)item * parts + q);
                        merge_min(d, k, c, __float_as_uint(v.x), __float_as_int(v.y), __float_as_int(v.z));
                    }
                }
            }
            if (final_block && atomicExch(p.out_cnt + 2 * n, -2) != -2) {
                if (k == 0x7fffffff) k = 0;
                reinterpret_cast<uint2*>(p.out_q)[n * kOutCap] = make_uint2((uint32_t)(k >> 5), 1u << (k & 31));
                if (p.stats != nullptr) {
                    if (c > 1) atomicAdd(p.stats + 0, 1ull);
                    atomicAdd(p.stats + 2, 1ull);
                }
            }
            is_final = 0;
        }
    }
}

}  // namespace vq
