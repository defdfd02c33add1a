// vq_argmin_sm100.cuh -- fused distance GEMM + candidate argmin for sm_100a (tcgen05 / TMEM / TMA).
//
// Computes, for every latent row n, the approximate scores  s[n,k] = |e_k|^2 - 2 * fp16(z_n) . fp16(e_k)
// (the k-dependent part of codebook.py:70-79) on the 5th-gen tensor cores and reduces them IN THE EPILOGUE to a
// short list of candidate entries -- a 32-code chunk plus a 32-bit mask of its codes whose score is within the
// candidate threshold of the running row minimum (vq_common.cuh: candidate_threshold) -- so the N x K distance matrix
// never leaves the SM.  vq_select_kernel (vq_select.cuh) recomputes the distances of the surviving codes exactly in
// fp32 and takes the first minimum.
//
// CTA = 11 warps, persistent over row tiles (128 latents each).  In the production configuration (kShare) the CTAs run as
// clusters of two that SHARE the codebook stream: the two CTAs work on different row tiles but the same code tiles, each
// fetches half of every 32 KiB codebook stage and multicasts it into both rings (cp.async.bulk .multicast::cluster), which
// halves the L2 -> SMEM operand traffic per SM -- the kernel runs against the 1 kW power cap, so bytes moved are clock
// (measured: same cycles per tile, 2.5 % less time when the kernel is timed alone, neutral inside a sustained loop).  MMA issue, TMEM hand-off and epilogue stay local to the CTA; only the
// stage-release barrier collects a commit from both CTAs.  Two alternatives were built and measured at K = 16384 and
// dropped: cta_group::2 pairs (one M = 256 MMA over both SMs, each CTA holding half of B: two cross-CTA hops land in the
// TMEM buffer cycle, 1.70 ms vs 1.46 ms) and 128-code half tiles with four accumulator quarters (N = 128 MMAs re-read the
// A operand twice per tile and saturate shared-memory bandwidth, 1.59 ms).
//   warp 0      bulk-copy producer (cp.async.bulk): A = z tile (4 chunks of [128 x 64] fp16) once per row tile, B = codebook
//               tile ([256 codes x 64] fp16 = 32 KiB per stage) through a 4-stage ring, and the |e|^2 slice of every
//               code tile (1 KiB) into a double buffer.  Both operands are stored in global memory as ready-made
//               SWIZZLE_128B shared-memory images (vq_prep.cuh), so every stage is ONE contiguous bulk copy instead of
//               256 strided 128-byte rows of a tensor-map box.
//   warps 1,10  MMA issuers (warp 1 also allocates TMEM), alternating code tiles: per tile 16 x tcgen05.mma (M128 N256 K16)
//               into one of two 256-column fp32 accumulators (the epilogue of tile j overlaps the MMAs of tile j+1).
//   warps 2..5  epilogue group 0: columns [0, 128) of every accumulator tile
//   warps 6..9  epilogue group 1: columns [128, 256)
//               thread <-> TMEM lane <-> latent row; tcgen05.ld 32 columns at a time (prefetched one chunk ahead,
//               like the |e|^2 values), one FFMA per element for the score, a 3-input-min tree per chunk whose
//               4-wide partial minima are reused, and a short slow path that pushes (chunk, code mask, chunk minimum)
//               into a per-(row, group) ring in shared memory.  The accumulator buffer is released as soon as its
//               last columns are in registers.  A buffer cycles MMA -> drain -> MMA, so the tensor pipe stays busy
//               only while drain + hand-off latency <= one MMA tile time; splitting the columns over two warps per
//               SM sub-partition halves the drain and lets the two hide each other's latencies.
#pragma once

#include "ptx_sm100.cuh"
#include "vq_common.cuh"

namespace vq {

// Operand-ring depth for a contraction of kNC chunks of 64: four 32 KiB stages next to a resident latent tile of up to
// 64 KiB (kNC <= 4); D = 512 (kNC = 8) keeps a 128 KiB latent tile resident and is left with two stages.
__host__ __device__ constexpr int gemm_stages(int nc) { return nc == 8 ? 2 : 4; }
constexpr int kEpiGroups = 2;
constexpr int kGroupCols = kCodeTile / kEpiGroups;           // 128 accumulator columns per epilogue group
constexpr int kChunk = 32;                                   // codes per tcgen05.ld / per candidate entry
constexpr int kGemmThreads = 64 + kEpiGroups * 128 + 32;    // 352: TMA, MMA-even, 8 epilogue warps, MMA-odd
constexpr int kMmaWarpB = 2 + 4 * kEpiGroups;                // warp 10
constexpr uint32_t kBytesAChunk = kRowTile * kDChunk * 2;    // 16 KiB
constexpr uint32_t kBytesBStage = kCodeTile * kDChunk * 2;   // 32 KiB
constexpr uint32_t kBytesE2Tile = kCodeTile * 4;             // 1 KiB
constexpr uint32_t kTmemCols = 512;
constexpr int kOutPerGroup = kOutCap / kEpiGroups;           // candidate entries a group may hand over per row
constexpr int kMaxCodesPerGroup = 32;                        // ... and codes (the exact stage lists <= 64 per row)

template <int kNC>
struct GemmSmemT {
    static constexpr int kStagesB = gemm_stages(kNC);
    alignas(1024) uint8_t a[kNC][kBytesAChunk];              // 64 KiB at D = 256
    alignas(1024) uint8_t b[kStagesB][kBytesBStage];         // 128 KiB
    alignas(16) float e2s[2][kCodeTile];                     // 2 KiB   |e|^2 of the code tile in accumulator buffer b
    uint32_t ring_q[kEpiGroups][kRingCap][kRowTile];         // 8 KiB   chunk ids, [slot][row]: conflict-free
    uint32_t ring_m[kEpiGroups][kRingCap][kRowTile];         // 8 KiB   32-bit masks of the chunk's candidate codes
    float ring_s[kEpiGroups][kRingCap][kRowTile];            // 8 KiB   chunk minima
    float m_part[2][kEpiGroups][kRowTile];                   // 2 KiB   per-group running minima (double buffered)
    int32_t c_part[2][kEpiGroups][kRowTile];                 // 2 KiB   per-group push counts
    float m_live[kEpiGroups][kRowTile];                      // 1 KiB   running minima, refreshed once per code tile
    uint8_t bad_part[kEpiGroups][kRowTile];                  //         per-group "row needs the exact fallback" verdicts
    alignas(8) uint64_t a_full[kNC];
    uint64_t a_empty[kNC];
    uint64_t b_full[2][kStagesB];        // one set per tile parity: each MMA warp sees every phase of its own set
    uint64_t b_empty[kStagesB];
    uint64_t t_full[2];
    uint64_t t_empty[2];
    uint64_t e2_full[2];
    uint64_t e2_empty[2];
    uint32_t tmem_base;
};
template <int kNC>
constexpr size_t gemm_smem_bytes() { return sizeof(GemmSmemT<kNC>) + 1024; }   // + slack for manual 1024 B alignment
constexpr size_t kGemmSmemBytes = gemm_smem_bytes<kNumDChunks>();
static_assert(gemm_smem_bytes<1>() <= 232448 && gemm_smem_bytes<2>() <= 232448 && gemm_smem_bytes<4>() <= 232448 &&
              gemm_smem_bytes<8>() <= 232448, "exceeds the 227 KiB of shared memory a CTA can opt into");

struct GemmParams {
    const __half* z_h;         // operand image of the latents  [row tile][D chunk][128][64] (vq_prep.cuh)
    const __half* e_h;         // operand image of the codebook [code tile][D chunk][256][64]
    const float* e2;           // (K_pad) |e_k|^2, +inf on pad rows
    const float* cb;           // codebook scalars (vq_prep.cuh)
    const float* z2;           // (N)
    const float* z_inv_scale;  // (N)
    int64_t N;
    int k_tiles;               // K_pad / 256
    int row_tiles;             // N_pad / 128
    int32_t* out_cnt;          // (N, 2) candidate entries per epilogue group, or -1: list unusable -> exact row scan
    uint32_t* out_q;           // (N, kOutCap, 2) entries (chunk id, 32-bit code mask); group g owns slots [g*8, g*8+8)
    int32_t* fb_rows;          // (2N) worklist of the rows flagged -1 (for vq_fallback_kernel)
    int32_t* fb_count;         // (1)  its length, zeroed before launch
    float* dbg_scores;         // (N, K_pad) or null
    int recipe;                // kRecipeExpanded / kRecipeDiffSq: selects the candidate threshold (vq_common.cuh)
    long long* timeline;       // debug: per-tile clock64 stamps of CTA 0 (kTimeline builds), [tile][8]
    int timeline_tiles;
};

__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ void epi_barrier() {               // the 256 epilogue threads only
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

// kShare: launched as clusters of two CTAs that share the codebook stream (see the header); a work unit is then a PAIR of
// row tiles (row tile 2u + rank for the CTA of that rank; z_h is padded to whole pairs), otherwise one row tile.
// kNC: the contraction runs over kNC chunks of 64 (D padded to 64 kNC; 4 = the CodeBook's 256, the others serve the
// row-major nearest-code searches at their native widths, vq_rows.cuh).
template <bool kDebugScores, bool kTimeline = false, bool kShare = false, int kNC = kNumDChunks>
__global__ void __launch_bounds__(kGemmThreads, 1)
vq_argmin_gemm_kernel(const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    using GemmSmem = GemmSmemT<kNC>;
    constexpr int kStagesB = GemmSmem::kStagesB;
    constexpr int kTilesInRing = (kNC < kStagesB) ? kStagesB / kNC : 1;     // whole code tiles the operand ring holds
    GemmSmem& s = *reinterpret_cast<GemmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = kShare ? cluster_ctarank() : 0u;
    const int unit0 = kShare ? (int)cluster_id_x() : (int)blockIdx.x;
    const int unit_step = kShare ? (int)cluster_count_x() : (int)gridDim.x;
    const int n_units = kShare ? (p.row_tiles + 1) / 2 : p.row_tiles;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kNC; i++) { mbar_init(&s.a_full[i], 1); mbar_init(&s.a_empty[i], 1); }
        // (a shared stage is refilled when BOTH CTAs' MMAs on it have committed)
        for (int i = 0; i < kStagesB; i++) { mbar_init(&s.b_full[0][i], 1); mbar_init(&s.b_full[1][i], 1); mbar_init(&s.b_empty[i], kShare ? 2 : 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.t_full[i], 1);
            mbar_init(&s.t_empty[i], 4 * kEpiGroups);
            mbar_init(&s.e2_full[i], 1);
            mbar_init(&s.e2_empty[i], 4 * kEpiGroups);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(&s.tmem_base, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (kShare) cluster_sync_all();                          // the peer's barriers exist before anything is multicast to it
    tc_fence_after();
    const uint32_t tmem_base = s.tmem_base;
    // PDL (vq_common.cuh): everything above ran while vq_prep_z_kernel was finishing; its operand image is needed from here
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // (whole warp in the loop, one elected lane issues: see the note on the MMA warps)
        {
            const uint64_t pol_keep = policy_evict_last();    // codebook tiles are re-read by every CTA
            const uint64_t pol_stream = policy_evict_first(); // z tiles are read exactly once
            uint32_t stage = 0, b_phase = 0, a_phase = 0, buf = 0, e_phase = 0;
            long long it = 0;                                   // global code-tile counter of this CTA
            for (int u = unit0; u < n_units; u += unit_step) {
                const int64_t rt = kShare ? 2 * (int64_t)u + rank : u;
                for (int kt = 0; kt < p.k_tiles; kt++, it++) {
                    uint64_t* const bfull = s.b_full[it & 1];
                    for (int dc = 0; dc < kNC; dc++) {
                        if (kt == 0) {
                            mbar_wait(&s.a_empty[dc], a_phase ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(&s.a_full[dc], kBytesAChunk);
                                bulk_load_1d_hint(s.a[dc], p.z_h + (rt * kNC + dc) * (kRowTile * kDChunk),
                                                  kBytesAChunk, &s.a_full[dc], pol_stream);
                            }
                            __syncwarp();
                        }
                        const __half* src = p.e_h + ((int64_t)kt * kNC + dc) * (kCodeTile * kDChunk);
                        if (kShare) {
                            // this CTA fetches the codes [128 rank, 128 rank + 128) of the stage for both CTAs; each CTA arms
                            // its own barrier for the whole stage
                            mbar_wait_cluster(&s.b_empty[stage], b_phase ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(&bfull[stage], kBytesBStage);
                                bulk_load_1d_multicast(s.b[stage] + rank * (kBytesBStage / 2), src + rank * (kCodeTile / 2 * kDChunk),
                                                       kBytesBStage / 2, &bfull[stage], (uint16_t)3, pol_keep);
                            }
                        } else {
                            mbar_wait(&s.b_empty[stage], b_phase ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(&bfull[stage], kBytesBStage);
                                bulk_load_1d_hint(s.b[stage], src, kBytesBStage, &bfull[stage], pol_keep);
                            }
                        }
                        __syncwarp();
                        if (++stage == kStagesB) { stage = 0; b_phase ^= 1; }
                    }
                    // |e|^2 of this code tile, issued after its operand stages so that waiting for the epilogue to
                    // release the buffer (two tiles back) never delays an operand load
                    mbar_wait(&s.e2_empty[buf], e_phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&s.e2_full[buf], kBytesE2Tile);
                        bulk_load_1d(s.e2s[buf], p.e2 + (int64_t)kt * kCodeTile, kBytesE2Tile, &s.e2_full[buf]);
                    }
                    __syncwarp();
                    buf ^= 1;
                    if (buf == 0) e_phase ^= 1;
                }
                a_phase ^= 1;
            }
        }
    } else if (warp == 1 || warp == kMmaWarpB) {
        // ------------------------------------------------------------------ MMA issuers (two warps, alternating tiles)
        // The issuing thread paces the tensor pipe, and every barrier it has to look at (TMEM buffer, operand stages)
        // costs ~100 cycles even when already complete.  Two warps therefore alternate code tiles -- warp 1 the even
        // ones (accumulator buffer 0), warp kMmaWarpB the odd ones (buffer 1): while one issues its 16 MMAs the other
        // is already through the waits of the next tile.  No ordering is needed between the two: they write different
        // accumulators, tcgen05.commit tracks only the committing thread's MMAs, and the operand ring itself orders
        // tile j+1's use of a stage after tile j's (the stage is re-filled only after tile j's MMA on it completed).
        // Each whole warp walks its loop (warp-uniform control flow) and one elected lane issues: issuing from an
        // `if (lane == 0)` region makes the compiler wrap every uniform-datapath instruction in an elect/broadcast loop.
        {
            constexpr uint32_t idesc = umma_idesc_f16(kRowTile, kCodeTile);
            const uint32_t buf = (warp == 1) ? 0u : 1u;
            const uint32_t d_tmem = tmem_base + buf * kCodeTile;
            long long it = 0;                                   // global code-tile counter of this CTA
            int rti = 0, tl_seq = 0;
            for (int u = unit0; u < n_units; u += unit_step, rti++) {
                // Both warps look at every row tile's z chunks (even when a short codebook gives a warp no code tile in
                // this row tile): an mbarrier parity wait must never fall a whole phase behind.  With an operand ring at
                // least one code tile deep (kNC <= kStagesB) all chunks are awaited up front.  (Looking at them lazily, chunk
                // by chunk inside the first code tile, so that the MMAs start when chunk 0 has landed, was measured there: no
                // gain at K = 2048 and 3 % slower at K = 16384 -- one more predicated wait per chunk in the issue loop.)
                // D = 512 (kNC = 8, two stages) MUST look lazily: the producer interleaves latent chunks with codebook stages
                // and cannot deliver chunk 2 before the MMAs have freed a stage.
                constexpr bool kLazyA = kNC > kStagesB;
                bool a_seen = !kLazyA;
                if (!kLazyA) {
#pragma unroll
                    for (int dc = 0; dc < kNC; dc++) mbar_wait(&s.a_full[dc], rti & 1);
                }
                for (int kt = 0; kt < p.k_tiles; kt++, it++) {
                    if ((uint32_t)(it & 1) != buf) continue;
                    const uint32_t use = (uint32_t)(it >> 1);    // how often this buffer was used before
                    long long tl0 = 0;
                    if (kTimeline) tl0 = clock64();
                    mbar_wait(&s.t_empty[buf], (use & 1) ^ 1);
                    tc_fence_after();
                    if (kTimeline && lane == 0 && blockIdx.x == 0 && it < p.timeline_tiles) {
                        p.timeline[it * 12 + 0] = tl0;
                        p.timeline[it * 12 + 1] = clock64();
                    }
                    long long tl_bwait = 0;
#pragma unroll
                    for (int dc = 0; dc < kNC; dc++) {
                        // Where chunk (it, dc) sits in the operand ring and how often THIS warp's barrier set has seen that
                        // stage before (its parity).  The ring holds kTilesInRing whole code tiles (slots it % kTilesInRing;
                        // an even number of them, or one: a stage then alternates between the two warps' sets), or -- D = 512
                        // -- a quarter of one, each stage being used kNC / kStagesB times per tile.
                        const int stage = (kNC <= kStagesB) ? (int)(it % kTilesInRing) * kNC + dc : dc % kStagesB;
                        const uint32_t seen = (kNC < kStagesB) ? (uint32_t)(it / kTilesInRing)
                                            : (kNC == kStagesB) ? use : use * (uint32_t)(kNC / kStagesB) + (uint32_t)(dc / kStagesB);
                        long long tb0 = 0;
                        if (kTimeline) tb0 = clock64();
                        if (kLazyA && !a_seen) mbar_wait(&s.a_full[dc], rti & 1);
                        if (kShare) mbar_wait_cluster(&s.b_full[buf][stage], seen & 1);    // half of it was written by the peer's copy
                        else mbar_wait(&s.b_full[buf][stage], seen & 1);
                        tc_fence_after();
                        if (kTimeline) tl_bwait += clock64() - tb0;
                        const uint64_t adesc = umma_desc_sw128(smem_u32(s.a[dc]));
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(s.b[stage]));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kDChunk / 16; k++) {
                                // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                                umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dc | k) != 0);
                            }
                            if (kShare) umma_commit_multicast(&s.b_empty[stage], (uint16_t)3);
                            else umma_commit(&s.b_empty[stage]);
                            if (kt == p.k_tiles - 1) umma_commit(&s.a_empty[dc]);
                            if (dc == kNC - 1) umma_commit(&s.t_full[buf]);
                        }
                        __syncwarp();
                    }
                    a_seen = true;
                    if (kTimeline && lane == 0 && blockIdx.x == 0 && it < p.timeline_tiles) {
                        p.timeline[it * 12 + 2] = clock64();
                        p.timeline[it * 12 + 8] = tl_bwait;
                    }
                    (void)tl_seq;
                }
                if (kLazyA && !a_seen) {                            // this warp had no code tile in the row tile
#pragma unroll
                    for (int dc = 0; dc < kNC; dc++) mbar_wait(&s.a_full[dc], rti & 1);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int grp = (warp - 2) >> 2;                    // which half of the accumulator columns
        const int quarter = warp & 3;                       // TMEM lanes [32*quarter, 32*quarter + 32)
        const int trow = quarter * 32 + lane;               // row inside the tile == TMEM lane
        const uint32_t col0 = grp * kGroupCols;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + col0;
        const uint32_t e2_sa[2] = {smem_u32(&s.e2s[0][col0]), smem_u32(&s.e2s[1][col0])};
        const uint32_t ring_q_sa = smem_u32(&s.ring_q[grp][0][trow]);
        const uint32_t ring_m_sa = smem_u32(&s.ring_m[grp][0][trow]);
        const uint32_t ring_s_sa = smem_u32(&s.ring_s[grp][0][trow]);
        const uint32_t live_own_sa = smem_u32(&s.m_live[grp][trow]);
        const uint32_t live_other_sa = smem_u32(&s.m_live[grp ^ 1][trow]);
        sts_f32(live_own_sa, INFINITY);
        epi_barrier();
        constexpr uint32_t kSlotStride = kRowTile * 4;
        const float e2max = __ldg(p.cb + 0);
        const float e_inv = __ldg(p.cb + 2);
        uint32_t buf = 0, phase = 0;
        int rti = 0, tl_seq = 0;
        for (int u = unit0; u < n_units; u += unit_step, rti++) {
            const int64_t row = (kShare ? 2 * (int64_t)u + rank : (int64_t)u) * kRowTile + trow;
            const bool row_ok = row < p.N;
            const float z2row = row_ok ? __ldg(p.z2 + row) : 0.0f;
            // Inf / NaN in the row or anywhere in the codebook: distances become NaN / inf, where torch.argmin's rule (first
            // NaN wins) cannot be decided from approximate scores -> the row takes the exact full scan
            const bool nonfinite = !(fabsf(z2row) < INFINITY) || !(fabsf(e2max) < INFINITY);
            // threshold for a minimum m: fma(m, cmul, margin0); for the CodeBook's recipe cmul == 1 and this is m + margin
            float cmul, margin;
            candidate_threshold(p.recipe, z2row, e2max, cmul, margin);
            // score = e2 + cscale * acc,  acc = (z * 2^a) . (e * 2^b)  ->  cscale = -2 * 2^-a * 2^-b  (exact)
            const float cscale = -2.0f * (row_ok ? __ldg(p.z_inv_scale + row) : 1.0f) * e_inv;
            float m_run = INFINITY, thr = INFINITY, lost_min = INFINITY;
            int cnt = 0;
            uint32_t slot_off = 0;

            for (int kt = 0; kt < p.k_tiles; kt++) {
                mbar_wait(&s.t_full[buf], phase);
                mbar_wait(&s.e2_full[buf], phase);
                tc_fence_after();
                const bool tl_on = kTimeline && blockIdx.x == 0 && lane == 0 && quarter == 2 && tl_seq < p.timeline_tiles;
                if (tl_on && grp == 0) p.timeline[tl_seq * 12 + 3] = clock64();
                const uint32_t taddr = t_lane + buf * kCodeTile;
                const uint32_t e2a = e2_sa[buf];
                // the other column group's running minimum (possibly one tile stale: still an upper bound of the row
                // minimum) tightens this group's threshold
                thr = fminf(thr, __fmaf_rn(fminf(m_run, lds_f32(live_other_sa)), cmul, margin));

                // The accumulator buffer must go back to the MMA warps EARLY: a buffer cycles MMA -> drain -> MMA and the
                // next MMA on it is due one tile time after the previous one ended.  All four 32-column loads of this
                // warp are therefore issued within its first two chunks (three register buffers; tcgen05.wait::ld
                // waits for everything outstanding), and the buffer is released at the start of the third chunk.
                static_assert(kGroupCols / kChunk == 4, "load schedule below is written for four chunks per warp");
                uint32_t acc[3][32];
                float4 ev[2][8];
                tmem_ld32(taddr, acc[0]);
                tmem_ld32(taddr + kChunk, acc[1]);
#pragma unroll
                for (int q = 0; q < 8; q++) ev[0][q] = lds128(e2a + q * 16);
#pragma unroll
                for (int c = 0; c < kGroupCols / kChunk; c++) {
                    const int h = c & 1;
                    const int ab = c % 3;
                    if (c == 0) {
                        tmem_ld_wait();                                 // chunks 0 and 1 are in registers
                        tmem_ld32(taddr + 2 * kChunk, acc[2]);
                    } else if (c == 1) {
                        tmem_ld32(taddr + 3 * kChunk, acc[0]);          // acc[0] was consumed by chunk 0
                    } else if (c == 2) {
                        tmem_ld_wait();                                 // chunks 2 and 3 are in registers:
                        tc_fence_before();                              // release the accumulator buffer to the MMA warps
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s.t_empty[buf]);
                        if (tl_on) p.timeline[tl_seq * 12 + (grp == 0 ? 4 : 6)] = clock64();
                    }
                    if (c + 1 < kGroupCols / kChunk) {
#pragma unroll
                        for (int q = 0; q < 8; q++) ev[h ^ 1][q] = lds128(e2a + (c + 1) * kChunk * 4 + q * 16);
                    } else {
                        __syncwarp();                                   // every lane has read its last |e|^2 values
                        if (lane == 0) mbar_arrive(&s.e2_empty[buf]);
                    }
                    float sc[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        sc[4 * q + 0] = __fmaf_rn(cscale, __uint_as_float(acc[ab][4 * q + 0]), ev[h][q].x);
                        sc[4 * q + 1] = __fmaf_rn(cscale, __uint_as_float(acc[ab][4 * q + 1]), ev[h][q].y);
                        sc[4 * q + 2] = __fmaf_rn(cscale, __uint_as_float(acc[ab][4 * q + 2]), ev[h][q].z);
                        sc[4 * q + 3] = __fmaf_rn(cscale, __uint_as_float(acc[ab][4 * q + 3]), ev[h][q].w);
                    }
                    if (kDebugScores) {
                        if (row_ok) {
                            float* dst = p.dbg_scores + row * ((int64_t)p.k_tiles * kCodeTile) + kt * kCodeTile + col0 + c * kChunk;
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
                        // slow path: this chunk holds a code within the running threshold -> one ring entry with the
                        // mask of all such codes (only quads whose minimum passes are looked into)
                        m_run = fminf(m_run, cm);
                        thr = fminf(thr, __fmaf_rn(m_run, cmul, margin));
                        uint32_t cmask = 0;
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            if (m8[q] <= thr) {
                                cmask |= ((sc[4 * q + 0] <= thr ? 1u : 0u) | (sc[4 * q + 1] <= thr ? 2u : 0u) |
                                          (sc[4 * q + 2] <= thr ? 4u : 0u) | (sc[4 * q + 3] <= thr ? 8u : 0u)) << (4 * q);
                            }
                        }
                        // (Reusing the slot of an entry that has gone stale -- chunk minimum above the current threshold -- instead
                        // of overwriting the oldest one takes the overflowed rows of cfg4 / init from 2 to 0, bit-exact, and was
                        // dropped: the eight extra shared-memory reads on this path slow the kernel by 2.5 %, +35 us against the
                        // 20 us the scan of two rows costs -- profiles/r2_ab_ring_reuse.jsonl.)
                        if (cnt >= kRingCap) lost_min = fminf(lost_min, lds_f32(ring_s_sa + slot_off));
                        sts_u32(ring_q_sa + slot_off, (uint32_t)(kt * (kCodeTile / kChunk) + grp * (kGroupCols / kChunk) + c));
                        sts_u32(ring_m_sa + slot_off, cmask);
                        sts_f32(ring_s_sa + slot_off, cm);
                        cnt++;
                        slot_off = (slot_off + kSlotStride == kRingCap * kSlotStride) ? 0u : slot_off + kSlotStride;
                    }
                    __syncwarp();
                }
                sts_f32(live_own_sa, m_run);
                if (tl_on) p.timeline[tl_seq * 12 + (grp == 0 ? 5 : 7)] = clock64();
                tl_seq++;
                buf ^= 1;
                if (buf == 0) phase ^= 1;
            }

            // combine the two groups' running minima, then hand the entries that survive the FINAL threshold to the
            // exact stage (each group filters its own ring into its half of the output slots)
            const int pb = rti & 1;
            s.m_part[pb][grp][trow] = m_run;
            s.c_part[pb][grp][trow] = cnt;
            sts_f32(live_own_sa, INFINITY);                     // next row tile starts from scratch (before the barrier)
            epi_barrier();
            const float m_fin = fminf(m_run, s.m_part[pb][grp ^ 1][trow]);
            const int cnt_all = cnt + s.c_part[pb][grp ^ 1][trow];
            if (row_ok) {
                const float thr_fin = __fmaf_rn(m_fin, cmul, margin);
                int n_out = 0, n_codes = 0;
                // nothing recorded at all (NaN row: group 0 reports) or a possible survivor was overwritten
                bool bad = (cnt_all == 0 && grp == 0) || (lost_min <= thr_fin) || nonfinite;
                const int live = min(cnt, kRingCap);
                uint2* dst = reinterpret_cast<uint2*>(p.out_q) + row * kOutCap + grp * kOutPerGroup;
                for (int i = 0; i < live && !bad; i++) {
                    if (lds_f32(ring_s_sa + i * kSlotStride) <= thr_fin) {
                        const uint32_t chunk = lds_u32(ring_q_sa + i * kSlotStride);
                        const uint32_t cmask = lds_u32(ring_m_sa + i * kSlotStride);
                        n_codes += __popc(cmask);
                        if (n_out < kOutPerGroup && n_codes <= kMaxCodesPerGroup) dst[n_out] = make_uint2(chunk, cmask);
                        else bad = true;
                        n_out++;
                    }
                }
                p.out_cnt[row * kEpiGroups + grp] = bad ? -1 : n_out;
                s.bad_part[grp][trow] = bad ? 1 : 0;
            }
            // a row goes on the fallback worklist ONCE, whichever group(s) flagged it: group 0 lists it after seeing
            // group 1's verdict
            epi_barrier();
            if (row_ok && grp == 0 && (s.bad_part[0][trow] | s.bad_part[1][trow]) != 0)
                p.fb_rows[atomicAdd(p.fb_count, 1)] = (int32_t)row;
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (kShare) cluster_sync_all();      // nobody leaves while the peer may still multicast into this ring or signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace vq
