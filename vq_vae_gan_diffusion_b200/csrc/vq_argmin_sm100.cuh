// vq_argmin_sm100.cuh -- fused distance GEMM + candidate argmin for sm_100a (tcgen05 / TMEM / TMA).
//
// Computes, for every latent row n, the approximate scores  s[n,k] = |e_k|^2 - 2 * fp16(z_n) . fp16(e_k)
// (the k-dependent part of codebook.py:70-79) on the 5th-gen tensor cores and reduces them IN THE EPILOGUE to a
// short list of candidate "quads" (4 consecutive codes) whose minimum is within margin[n] of the row minimum,
// so the N x K distance matrix never leaves the SM.  vq_select_kernel (vq_select.cuh) recomputes the distances of
// the surviving quads exactly in fp32 and takes the first minimum.
//
// CTA = 6 warps, persistent over row tiles (128 latents each):
//   warp 0      TMA producer: A = z tile (4 chunks of [128 x 64] fp16, SWIZZLE_128B) once per row tile,
//               B = codebook tile ([256 codes x 64] fp16 per stage) through a 4-stage ring.
//   warp 1      TMEM allocator + MMA issuer: per code tile 16 x tcgen05.mma (M128 N256 K16) into one of two
//               256-column fp32 accumulators (double buffered: the epilogue of tile j overlaps the MMAs of j+1).
//   warps 2..5  epilogue: thread <-> TMEM lane <-> latent row; tcgen05.ld 32 columns at a time, one FFMA per
//               element for the score, a 3-input-min tree per 32-column chunk whose 4-wide partial minima are the
//               quad minima, and a short slow path that pushes quads within the running threshold into a
//               per-row ring in shared memory.
#pragma once
#include <cuda.h>

#include "ptx_sm100.cuh"
#include "vq_common.cuh"

namespace vq {

constexpr int kStagesB = 4;
constexpr int kGemmThreads = 192;
constexpr uint32_t kBytesAChunk = kRowTile * kDChunk * 2;    // 16 KiB
constexpr uint32_t kBytesBStage = kCodeTile * kDChunk * 2;   // 32 KiB
constexpr uint32_t kTmemCols = 512;

struct GemmSmem {
    alignas(1024) uint8_t a[kNumDChunks][kBytesAChunk];      // 64 KiB
    alignas(1024) uint8_t b[kStagesB][kBytesBStage];         // 128 KiB
    int32_t ring_q[kRingCap][kRowTile];                      // 8 KiB   quad ids,    [slot][row]: conflict-free
    float ring_s[kRingCap][kRowTile];                        // 8 KiB   quad minima
    alignas(8) uint64_t a_full[kNumDChunks];
    uint64_t a_empty[kNumDChunks];
    uint64_t b_full[kStagesB];
    uint64_t b_empty[kStagesB];
    uint64_t t_full[2];
    uint64_t t_empty[2];
    uint32_t tmem_base;
};
constexpr size_t kGemmSmemBytes = sizeof(GemmSmem) + 1024;   // + slack for manual 1024 B alignment

struct GemmParams {
    const float* e2;           // (K_pad) |e_k|^2, +inf on pad rows
    const float* cb;           // codebook scalars (vq_prep.cuh)
    const float* z2;           // (N)
    const float* z_inv_scale;  // (N)
    int64_t N;
    int k_tiles;               // K_pad / 256
    int row_tiles;             // N_pad / 128
    int32_t* out_cnt;          // (N)  surviving quads, or -1: list unusable -> exact scan of the whole row
    int32_t* out_q;            // (N, kOutCap) quad ids (code / 4)
    float* dbg_scores;         // (N, K_pad) or null
};

__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <bool kDebugScores>
__global__ void __launch_bounds__(kGemmThreads, 1)
vq_argmin_gemm_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_e,
                      const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem& s = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_z);
        tma_prefetch_desc(&tmap_e);
        for (int i = 0; i < kNumDChunks; i++) { mbar_init(&s.a_full[i], 1); mbar_init(&s.a_empty[i], 1); }
        for (int i = 0; i < kStagesB; i++) { mbar_init(&s.b_full[i], 1); mbar_init(&s.b_empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&s.t_full[i], 1); mbar_init(&s.t_empty[i], 4); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(&s.tmem_base, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            const uint64_t pol_keep = policy_evict_last();    // codebook tiles are re-read by every CTA
            const uint64_t pol_stream = policy_evict_first(); // z tiles are read exactly once
            uint32_t stage = 0, b_phase = 0, a_phase = 0;
            for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
                for (int kt = 0; kt < p.k_tiles; kt++) {
                    for (int dc = 0; dc < kNumDChunks; dc++) {
                        if (kt == 0) {
                            mbar_wait(&s.a_empty[dc], a_phase ^ 1);
                            mbar_expect_tx(&s.a_full[dc], kBytesAChunk);
                            tma_load_2d_hint(s.a[dc], &tmap_z, dc * kDChunk, rt * kRowTile, &s.a_full[dc], pol_stream);
                        }
                        mbar_wait(&s.b_empty[stage], b_phase ^ 1);
                        mbar_expect_tx(&s.b_full[stage], kBytesBStage);
                        tma_load_2d_hint(s.b[stage], &tmap_e, dc * kDChunk, kt * kCodeTile, &s.b_full[stage], pol_keep);
                        if (++stage == kStagesB) { stage = 0; b_phase ^= 1; }
                    }
                }
                a_phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(kRowTile, kCodeTile);
            uint32_t stage = 0, b_phase = 0, a_phase = 0, buf = 0, t_phase = 0;
            for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
                for (int kt = 0; kt < p.k_tiles; kt++) {
                    mbar_wait(&s.t_empty[buf], t_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * kCodeTile;
                    for (int dc = 0; dc < kNumDChunks; dc++) {
                        if (kt == 0) mbar_wait(&s.a_full[dc], a_phase);
                        mbar_wait(&s.b_full[stage], b_phase);
                        tc_fence_after();
                        const uint64_t adesc = umma_desc_sw128(smem_u32(s.a[dc]));
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(s.b[stage]));
#pragma unroll
                        for (int k = 0; k < kDChunk / 16; k++) {
                            // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dc | k) != 0);
                        }
                        umma_commit(&s.b_empty[stage]);
                        if (kt == p.k_tiles - 1) umma_commit(&s.a_empty[dc]);
                        if (++stage == kStagesB) { stage = 0; b_phase ^= 1; }
                    }
                    umma_commit(&s.t_full[buf]);
                    buf ^= 1;
                    if (buf == 0) t_phase ^= 1;
                }
                a_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;                       // TMEM lanes [32*quarter, 32*quarter + 32)
        const int trow = quarter * 32 + lane;               // row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float e2max = __ldg(p.cb + 0);
        const float e_inv = __ldg(p.cb + 2);
        uint32_t buf = 0, t_phase = 0;
        for (int rt = blockIdx.x; rt < p.row_tiles; rt += gridDim.x) {
            const int64_t row = (int64_t)rt * kRowTile + trow;
            const bool row_ok = row < p.N;
            const float margin = candidate_margin(row_ok ? __ldg(p.z2 + row) : 0.0f, e2max);
            // score = e2 + cscale * acc,  acc = (z * 2^a) . (e * 2^b)  ->  cscale = -2 * 2^-a * 2^-b  (exact)
            const float cscale = -2.0f * (row_ok ? __ldg(p.z_inv_scale + row) : 1.0f) * e_inv;
            float m_run = INFINITY, thr = INFINITY, lost_min = INFINITY;
            int cnt = 0;

            for (int kt = 0; kt < p.k_tiles; kt++) {
                mbar_wait(&s.t_full[buf], t_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_addr + buf * kCodeTile;
                const float4* e2v = reinterpret_cast<const float4*>(p.e2 + (int64_t)kt * kCodeTile);

                uint32_t acc[2][32];
                tmem_ld32(taddr, acc[0]);
#pragma unroll 1
                for (int c2 = 0; c2 < kCodeTile / 64; c2++) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int c = 2 * c2 + h;
                        tmem_ld_wait();
                        // prefetch the next 32 columns into the other register buffer while this one is reduced
                        if (c + 1 < kCodeTile / 32) tmem_ld32(taddr + (c + 1) * 32, acc[h ^ 1]);
                        float sc[32];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 e = __ldg(e2v + c * 8 + q);
                            sc[4 * q + 0] = __fmaf_rn(cscale, __uint_as_float(acc[h][4 * q + 0]), e.x);
                            sc[4 * q + 1] = __fmaf_rn(cscale, __uint_as_float(acc[h][4 * q + 1]), e.y);
                            sc[4 * q + 2] = __fmaf_rn(cscale, __uint_as_float(acc[h][4 * q + 2]), e.z);
                            sc[4 * q + 3] = __fmaf_rn(cscale, __uint_as_float(acc[h][4 * q + 3]), e.w);
                        }
                        if (kDebugScores) {
                            if (row_ok) {
                                float* dst = p.dbg_scores + row * ((int64_t)p.k_tiles * kCodeTile) + kt * kCodeTile + c * 32;
#pragma unroll
                                for (int i = 0; i < 32; i++) dst[i] = sc[i];
                            }
                        }
                        // quad minima (4 consecutive codes) and the chunk minimum: 3-input min tree
                        float m8[8];
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            m8[q] = fminf(min3(sc[4 * q], sc[4 * q + 1], sc[4 * q + 2]), sc[4 * q + 3]);
                        const float cm = min3(min3(m8[0], m8[1], m8[2]), min3(m8[3], m8[4], m8[5]), fminf(m8[6], m8[7]));
                        if (cm <= thr) {
                            // slow path: this chunk holds a quad within the running threshold
                            m_run = fminf(m_run, cm);
                            thr = m_run + margin;
                            const int qbase = (kt * kCodeTile + c * 32) / kQuad;
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                if (m8[q] <= thr) {
                                    const int slot = cnt & (kRingCap - 1);
                                    if (cnt >= kRingCap) lost_min = fminf(lost_min, s.ring_s[slot][trow]);
                                    s.ring_q[slot][trow] = qbase + q;
                                    s.ring_s[slot][trow] = m8[q];
                                    cnt++;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.t_empty[buf]);
                buf ^= 1;
                if (buf == 0) t_phase ^= 1;
            }

            // hand the quads that survive the FINAL threshold to the exact stage
            if (row_ok) {
                int n_out = 0;
                bool bad = (cnt == 0) || (lost_min <= thr);      // nothing recorded (NaN row) or a survivor was overwritten
                const int live = min(cnt, kRingCap);
                for (int i = 0; i < live && !bad; i++) {
                    if (s.ring_s[i][trow] <= thr) {
                        if (n_out < kOutCap) p.out_q[row * kOutCap + n_out] = s.ring_q[i][trow];
                        else bad = true;
                        n_out++;
                    }
                }
                p.out_cnt[row] = bad ? -1 : n_out;
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace vq
