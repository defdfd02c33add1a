// vq_common.cuh -- constants and small device helpers shared by all kernels of the VQ hot path.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vq {

constexpr int kD = 256;             // latent_dim of every reference config (configs/*.yml: latent_channels 256)
constexpr int kRowTile = 128;       // latents per GEMM tile  (UMMA M, one TMEM lane per latent)
constexpr int kCodeTile = 256;      // codes per GEMM tile    (UMMA N)
constexpr int kDChunk = 64;         // 16-bit elements per 128-byte swizzle row
constexpr int kNumDChunks = kD / kDChunk;
constexpr int kRingCap = 8;      // candidate ring per (row, epilogue group) in shared memory inside the GEMM epilogue
constexpr int kOutCap = 16;         // surviving candidate quads handed to the exact stage, per row (half per group)
constexpr int kSelRows = 32;        // latents per CTA in the fp32 kernels (prep / select / backward)

// Programmatic dependent launch (PDL).  The kernels of one call form a chain prep_z -> GEMM -> fallback -> select; all but the
// first are launched with cudaLaunchAttributeProgrammaticStreamSerialization, every kernel calls pdl_trigger() when it starts
// (its successor may then be scheduled as soon as all of THIS grid's CTAs are resident and resources free up: block
// scheduling, barrier / TMEM set-up and loads of call inputs overlap this grid's tail) and pdl_wait() before it touches
// anything a predecessor in the chain wrote (returns once the predecessor grid has completed and its writes are visible;
// every kernel of the chain calls it, so completion is transitive).  The first kernel of a call is launched normally, so
// nothing of a call starts before the previous work in the stream -- e.g. the layer that produced z -- has finished.
// Both are no-ops for a kernel launched without the attribute / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Tensor-core operands are fp16 scaled by exact powers of two so that the largest magnitude lands in
// [2^14, 2^15): per latent row for z, per tensor for the codebook.  exponent_of() returns ex with |x| < 2^ex.
constexpr int kOperandTopExp = 15;
__device__ __forceinline__ int exponent_of(float maxabs) {
    int ex;
    (void)frexpf(maxabs, &ex);                 // maxabs = m * 2^ex, m in [0.5, 1)
    if (!(maxabs > 0.0f) || !(maxabs < INFINITY)) ex = 0;   // zero / inf / nan rows: scale 1 (handled downstream)
    return max(-100, min(100, ex));
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((e + 127) << 23); }   // -126 <= e <= 127

// ---------------------------------------------------------------------------------------------------------------
// Row-major [32 latents x 256] fp32 tile in shared memory with the 16-byte pieces of row r XOR-swizzled by
// g(r) = (r >> 2) ^ (2 (r & 3)).  Conflict-free for (a) the column-form fill / drain (lanes over 8 row groups x 4 d,
// one row of each group per scalar access), (b) one warp reading the same piece of its 4 rows (8 lanes per row
// broadcast) and (c) lanes over the pieces of one row (16-byte accesses, four wavefronts per 512 bytes).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_swz(int r) { return ((r >> 2) ^ ((r & 3) << 1)) & 7; }
__device__ __forceinline__ int tile_off(int r, int d) {                    // word offset of element (r, d)
    return r * kD + ((((d >> 2) ^ tile_swz(r)) << 2) | (d & 3));
}

// Column-form side of the tile (NCHW, HW % 32 == 0): thread (warp, lane) handles the latents 4 hq .. 4 hq + 3 (hq = lane & 7)
// at d = 4 (8 warp + i) + dsub (dsub = lane >> 3) for i = 0 .. 7 -- one 16-byte global access along hw per i, a warp
// request covering four full 128-byte lines.  ColForm precomputes everything that does not depend on i, so that a
// shared-memory access costs one LOP3 and one add on top of the LDS / STS (the kernels are issue-bound otherwise).
struct ColForm {
    int base[4];      // word offset of (row 4 hq + c, d = 32 warp + dsub) with the piece index zeroed
    int gx[4];        // swizzle of row 4 hq + c, pre-shifted to word units (<< 2)
    __device__ __forceinline__ ColForm(int warp, int lane) {
        const int hq = lane & 7, dsub = lane >> 3;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int r = 4 * hq + c;
            base[c] = r * kD + warp * 32 + dsub;
            gx[c] = tile_swz(r) << 2;
        }
    }
    __device__ __forceinline__ int off(int c, int i) const { return base[c] + ((i << 2) ^ gx[c]); }   // i in [0, 8)
    __device__ __forceinline__ void store(float* tile, int i, const float4& v) const {
        tile[off(0, i)] = v.x; tile[off(1, i)] = v.y; tile[off(2, i)] = v.z; tile[off(3, i)] = v.w;
    }
    __device__ __forceinline__ float4 load(const float* tile, int i) const {
        return make_float4(tile[off(0, i)], tile[off(1, i)], tile[off(2, i)], tile[off(3, i)]);
    }
};
// element offset of this thread's first column-form access (i = 0) in an NCHW tensor with `HW` positions per channel
// plane; access i is (4 i) channel planes further.  (N < 2^31 is enforced by the API: 32-bit division.)
__device__ __forceinline__ int64_t col_form_origin(int64_t n0, int64_t HW, int warp, int lane) {
    const uint32_t b = (uint32_t)n0 / (uint32_t)HW, hw0 = (uint32_t)n0 % (uint32_t)HW;
    return ((int64_t)b * kD + warp * 32 + (lane >> 3)) * HW + hw0 + 4 * (lane & 7);
}

// Input layouts of the latent tensor, as seen by the tile kernels.
constexpr int kLayoutGeneric = 0;   // NCHW (B, D, HW), any HW: lanes over latents, scalar accesses
constexpr int kLayoutVec = 1;       // NCHW with HW % 32 == 0 and 16-byte aligned base: 16-byte accesses along hw
constexpr int kLayoutRows = 2;      // row-major (N, D) vectors, 16-byte aligned: 16-byte accesses along d

// Tile fill for kLayoutRows: warp w loads rows 4w..4w+3, lane owns pieces (lane, 32 + lane) of each (8 requests of 16 bytes
// in flight per thread).  Rows >= N are zero-filled.
__device__ __forceinline__ void fill_tile_rows(float* tile, const float* __restrict__ x, int64_t n0, int64_t N, int warp, int lane) {
    float4 v[4][2];
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int64_t n = n0 + warp * 4 + rr;
        const float4* src = reinterpret_cast<const float4*>(x + (n < N ? n : 0) * kD);
#pragma unroll
        for (int h = 0; h < 2; h++) v[rr][h] = (n < N) ? __ldg(src + lane + 32 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const int r = warp * 4 + rr;
        float4* row4 = reinterpret_cast<float4*>(tile + r * kD);
        const int g = tile_swz(r);
#pragma unroll
        for (int h = 0; h < 2; h++) row4[(lane + 32 * h) ^ g] = v[rr][h];
    }
}

// Offset (in 16-bit elements) of element (row, dl) -- dl in [0, 64) -- of block `block` in an "operand image": blocks of
// `rows` x 64 elements stored exactly as a SWIZZLE_128B K-major UMMA operand sits in shared memory (row pitch 128 B,
// the eight 16-byte pieces of a row XOR-ed with (row & 7)).  A block is contiguous, so it is loaded with one bulk copy.
__host__ __device__ __forceinline__ int64_t operand_image_offset(int64_t block, int rows, int row, int dl) {
    return block * (int64_t)rows * kDChunk + (int64_t)row * kDChunk + ((((dl >> 3) ^ (row & 7)) << 3) | (dl & 7));
}

// Canonical-order dot product pieces (oracle/vq_oracle.c: vqo_dot): partial j sums the terms d == j (mod 4)
// in ascending d with one fma each; the result is (p0 + p1) + (p2 + p3).
__device__ __forceinline__ float combine4(float p) {
    // p holds partial j on lane (4*g + j); returns (p0+p1)+(p2+p3) on all four lanes (fp add is commutative).
    float q = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 1));
    return __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 2));
}

// Reference distance formula in fp32, codebook.py:70-79: fl( fl(|z|^2 + |e|^2) - fl(2 * dot) ).
__device__ __forceinline__ float ref_distance(float z2, float e2, float dot) {
    return __fsub_rn(__fadd_rn(z2, e2), __fmul_rn(2.0f, dot));
}

// Rigorous bound eps on | (score_approx + |z|^2) - d_oracle | valid for every code of a row, and the candidate
// margin derived from it.  score_approx = fl(e2[k] - 2 * (fp16(z) . fp16(e_k))) with fp32 accumulation in the tensor
// core; d_oracle = ref_distance() with a canonical-order fp32 dot.  With a = |z| * max_k |e_k| (Cauchy-Schwarz
// bound of sum |z_d e_d|), r = |z|^2 + max|e|^2 + 2a (bounds every intermediate of the ORACLE's formula, which carries
// |z|^2) and r' = max|e|^2 + 2a (bounds the approximate scores and the thresholds, which do not):
//   fp16 rounding of both operands (unit roundoff 2^-11, + subnormal slack)   2 (2^-10 + 2^-22 + 2^-35) a
//   tensor-core fp32 accumulation over D = 256 products                       <= 2^-13 a (budget, checked on HW)
//   oracle's own fp32 dot                                                     2 D 2^-24 a = 2^-15 a
//   the oracle's two roundings, fl(|z|^2 + |e|^2) and the subtraction         <= 2 * 2^-24 r  = 2^-23 r
//   the score's fma rounding and the rounding of the threshold add            <= 2 * 2^-24 r' = 2^-23 r'
//   => eps <= (2^-9 + 2^-12) a + 2^-23 (r + r')
// (Round 1 charged all four roundings at magnitude r: 2^-22 r.  On the reference's init distribution -- |e| ~ 1e-3 |z|, so
// r' << r and the r term is the largest -- that doubled eps and with it the rows that need the exact re-rank: the term that
// remains, 2^-23 r = one ulp of |z|^2, is the oracle's own quantisation and cannot shrink.)
// The oracle's argmin k* then satisfies score[k*] <= min_k score + 2 eps.  See DESIGN.md "candidate margin".
__device__ __forceinline__ float candidate_margin(float z2, float e2max) {
    const float a = sqrtf(z2) * sqrtf(e2max) * 1.000001f;
    const float rp = e2max + 2.0f * a;
    const float r = z2 + rp;
    const float eps = a * (0.001953125f + 0.000244140625f) + (r + rp) * 1.1920928955078125e-07f;
    return 2.0f * eps * 1.0625f + 1e-37f;
}


// ---------------------------------------------------------------------------------------------------------------
// Distance recipes.  The reference computes nearest codes with two different fp32 formulas:
//   kRecipeExpanded  |x|^2 + |e|^2 - 2 x.e     codebook.py:70-79, diffusion_gaussian2d.py:334-339
//   kRecipeDiffSq    sum_d (x_d - e_d)^2       v_vq_diffusion.py:114-123 (broadcast difference, squared, torch.sum)
// Both are decided by the same approximate scores from the tensor cores plus an exact fp32 stage that evaluates the
// recipe in the oracle's canonical order; only the exact stage and the candidate threshold differ.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRecipeExpanded = 0;
constexpr int kRecipeDiffSq = 1;
constexpr int kRecipeCdist = 2;     // sqrt(clamp_min([-2x, |x|^2, 1] . [e, 1, |e|^2], 0)) on L2-normalised rows, diffusion_gaussian3d.py:543-570

// Candidate threshold as a function of the (running or final) minimal score m:  thr(m) = fma(m, cmul, margin0).
//   expanded: cmul = 1, margin0 = candidate_margin()  -- thr = m + margin, bit-identical to the plain add.
//   diffsq:   with D_k the true squared distance, the canonical fp32 value d_k has |d_k - D_k| <= g D_k, g <= 69 * 2^-24
//             (difference, square, 64 + 2 additions), and S_k = score_k + |x|^2 has |S_k - D_k| <= eps1,
//             eps1 = (2^-9 + 2^-12) a + 2^-17.9 (|x|^2 + max|e|^2) + 2^-22 r  (operand rounding and tensor-core accumulation
//             as above; the canonical norms' own rounding, 66 * 2^-24 relative, no longer cancels; score arithmetic).
//             For k* = argmin d and j = argmin S:  D_k* (1-g) <= d_k* <= d_j <= (S_j + eps1)(1+g), hence
//             score_k* <= score_j + 2 eps1 + c (score_j + |x|^2 + eps1) with c = 2^-16 >= ((1+g)/(1-g) - 1) -- every code that can
//             be the exact argmin (or tie with it) satisfies it.
//   cdist:    the squared distance is ONE canonical product of the augmented vectors, so its fp32 error is that of the
//             dot product (2^-15 a) plus the roundings of the last fma of chains D mod 4 and (D + 1) mod 4 and of the two
//             combining adds, each at a magnitude <= r: 4 * 2^-24 r -- the same budget as the expanded formula's -- and the
//             square root can map two squared distances up to 3 ulp apart (<= 3 * 2^-23 r) onto one value, which must all
//             stay candidates because the first of them wins: eps <= (2^-9 + 2^-12) a + 2^-20 r.
__device__ __forceinline__ void candidate_threshold(int recipe, float z2, float e2max, float& cmul, float& margin0) {
    if (recipe == kRecipeExpanded) {
        cmul = 1.0f;
        margin0 = candidate_margin(z2, e2max);
        return;
    }
    if (recipe == kRecipeCdist) {
        const float a = sqrtf(z2) * sqrtf(e2max) * 1.000001f;
        const float r = z2 + e2max + 2.0f * a;
        const float eps = a * (0.001953125f + 0.000244140625f) + r * 9.5367431640625e-07f;
        cmul = 1.0f;
        margin0 = 2.0f * eps * 1.0625f + 1e-37f;
        return;
    }
    const float a = sqrtf(z2) * sqrtf(e2max) * 1.000001f;
    const float r = z2 + e2max + 2.0f * a;
    const float eps1 = a * (0.001953125f + 0.000244140625f) + (z2 + e2max) * 4.1e-6f + r * 2.384185791015625e-07f;
    const float c = 1.52587890625e-05f;                      // 2^-16
    cmul = 1.0f + c;
    margin0 = 2.0f * eps1 * 1.0625f + c * (z2 + eps1) * 1.0625f + 1e-37f;
}

}  // namespace vq
