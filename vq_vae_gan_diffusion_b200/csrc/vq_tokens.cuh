// vq_tokens.cuh -- what the reference does with the token stream right after the tokeniser.
//
// vq_log_onehot_kernel: index_to_log_onehot (network/vq_diffusion/vq_diffusion.py:29-35, diffusion_vq_official.py:53-60):
//     log(one_hot(x, C).permute(0, -1, 1..).float().clamp(min = 1e-30))   -> (B, C, L) fp32
//   The reference materialises an int64 (B, L, C) one-hot, a float copy, the clamp and the log (about 36 bytes of HBM
//   traffic per output element); here the (B, C, L) tensor is written once: 4 bytes per element, HBM-write bound.
//   A thread owns 4 consecutive positions l of one batch row and a tile of classes: it fills its column group with
//   log(clamp(0)) using 16-byte streaming stores (two instructions per 16 bytes, so the store stream and not the
//   issue rate is the bound) and then overwrites the entries of its own indices that fall into the tile with
//   log(clamp(1)) -- the same thread wrote the fill value to that address, so program order makes the fix-up win.
//   Both values are computed ON THE DEVICE with logf (the clamp minimum is a kernel argument, not a constant the
//   compiler could fold with the host's libm), which is what torch.log evaluates on a GPU.
//
// vq_argmax_classes_kernel: log_onehot_to_index (vq_diffusion.py:37-38), the inverse format change (see the kernel).
//
// vq_mask_replace_kernel: the arithmetic of VQTransformer.forward's input corruption (vqTransformer.py:117-141):
//     mask = bernoulli(pkeep).round().long(); new = mask * indices + (1 - mask) * random; cat(sos, new)
//   The two random draws stay in torch (same generator, same order, so the stream is the reference's); the round /
//   cast / blend / concatenation -- five elementwise kernels and a cat in the reference -- are one launch.
#pragma once
#include <cstdint>

namespace vq {

constexpr int kTokThreads = 256;
constexpr int kTokClassTile = 32;       // classes per CTA row of the one-hot kernel

template <bool kVec>
__global__ void __launch_bounds__(kTokThreads)
vq_log_onehot_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t L, int C, float clamp_min, float* __restrict__ out) {
    const float lo = logf(fmaxf(0.0f, clamp_min)), hi = logf(fmaxf(1.0f, clamp_min));
    const int k0 = (int)blockIdx.y * kTokClassTile, k1 = min(C, k0 + kTokClassTile);
    const int64_t t = (int64_t)blockIdx.x * kTokThreads + threadIdx.x;
    if (kVec) {
        const int64_t L4 = L >> 2;                             // L % 4 == 0 and 16-byte aligned pointers (host-checked)
        if (t >= B * L4) return;
        const int64_t b = t / L4, l = (t - b * L4) * 4;
        const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(idx + b * L + l));
        const longlong2 c = __ldg(reinterpret_cast<const longlong2*>(idx + b * L + l) + 1);
        float* o = out + (b * C + k0) * L + l;
        const float4 fill = make_float4(lo, lo, lo, lo);
        for (int k = k0; k < k1; k++, o += L) __stcs(reinterpret_cast<float4*>(o), fill);
        float* base = out + b * C * L + l;
        if (a.x >= k0 && a.x < k1) base[a.x * L + 0] = hi;
        if (a.y >= k0 && a.y < k1) base[a.y * L + 1] = hi;
        if (c.x >= k0 && c.x < k1) base[c.x * L + 2] = hi;
        if (c.y >= k0 && c.y < k1) base[c.y * L + 3] = hi;
    } else {
        if (t >= B * L) return;
        const int64_t b = t / L, l = t - b * L;
        const int64_t i = __ldg(idx + t);
        float* o = out + (b * C + k0) * L + l;
        for (int k = k0; k < k1; k++, o += L) *o = (i == k) ? hi : lo;
    }
}

// log_onehot_to_index (network/vq_diffusion/vq_diffusion.py:37-38): log_x.argmax(1) over the class axis of a (B, C, L)
// tensor -> (B, L) int64.  torch.argmax semantics: the first maximal value wins, a NaN counts as the maximum (the first
// NaN wins).  Thread <-> one position l of one batch row (kVec: four consecutive positions, 16-byte loads); a warp
// reads 128 / 512 contiguous bytes per class, classes streamed with eight loads in flight.  HBM-read bound: 4 bytes per
// input element.
__device__ __forceinline__ void argmax_step(float v, int k, float& best, int& best_k) {
    // v beats best when best is not NaN and (v is NaN or v > best)
    if (!(best != best) && ((v != v) || v > best)) { best = v; best_k = k; }
}

template <bool kVec>
__global__ void __launch_bounds__(kTokThreads)
vq_argmax_classes_kernel(const float* __restrict__ x, int64_t B, int64_t L, int C, int64_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * kTokThreads + threadIdx.x;
    if (kVec) {
        const int64_t L4 = L >> 2;
        if (t >= B * L4) return;
        const int64_t b = t / L4, l = (t - b * L4) * 4;
        const float* src = x + b * C * L + l;
        float4 best = __ldcs(reinterpret_cast<const float4*>(src));
        int k0 = 0, k1 = 0, k2 = 0, k3 = 0;
        int k = 1;
        for (; k + 8 <= C; k += 8) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)(k + i) * L));
#pragma unroll
            for (int i = 0; i < 8; i++) {
                argmax_step(v[i].x, k + i, best.x, k0); argmax_step(v[i].y, k + i, best.y, k1);
                argmax_step(v[i].z, k + i, best.z, k2); argmax_step(v[i].w, k + i, best.w, k3);
            }
        }
        for (; k < C; k++) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)k * L));
            argmax_step(v.x, k, best.x, k0); argmax_step(v.y, k, best.y, k1);
            argmax_step(v.z, k, best.z, k2); argmax_step(v.w, k, best.w, k3);
        }
        longlong2* o = reinterpret_cast<longlong2*>(out + b * L + l);
        o[0] = make_longlong2(k0, k1);
        o[1] = make_longlong2(k2, k3);
    } else {
        if (t >= B * L) return;
        const int64_t b = t / L, l = t - b * L;
        const float* src = x + b * C * L + l;
        float best = __ldcs(src);
        int bk = 0;
        for (int k = 1; k < C; k++) argmax_step(__ldcs(src + (int64_t)k * L), k, best, bk);
        out[t] = bk;
    }
}

__global__ void __launch_bounds__(kTokThreads)
vq_mask_replace_kernel(const int64_t* __restrict__ indices, const float* __restrict__ mask, const int64_t* __restrict__ random_indices,
                       int64_t sos, int64_t B, int64_t L, int64_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * kTokThreads + threadIdx.x;      // element of out (B, L + 1)
    if (t >= B * (L + 1)) return;
    const int64_t b = t / (L + 1), p = t - b * (L + 1);
    if (p == 0) { out[t] = sos; return; }
    const int64_t s = b * L + p - 1;
    const int64_t m = (int64_t)rintf(__ldg(mask + s));         // .round() (half to even) then .to(int64)
    out[t] = m * __ldg(indices + s) + (1 - m) * __ldg(random_indices + s);
}

}  // namespace vq
