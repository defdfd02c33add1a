// vq_allreduce.cuh -- one-shot-per-slice SUM all-reduce of the data-parallel exchange buffer over NVLink / NVSwitch with the
// switch doing the arithmetic (NVLS): every rank owns 1/W of the buffer, pulls the SUM of that slice from all ranks with
// multimem.ld_reduce on the buffer's multicast address (the reduction happens inside the NVSwitch) and pushes it back to all
// ranks with multimem.st -- each byte crosses a rank's links once in and once out.  The buffer is symmetric memory (same
// allocation on every rank, bound to one multicast object; torch.distributed._symmetric_memory does the allocation and the
// handle exchange: plumbing), the ranks meet at two device-side barriers on their signal pads (`world` words of each pad): one
// before the first load (every rank's contribution is complete: stream order on each rank puts its producer kernel before this
// one) and one after the last store (every slice has landed everywhere).  For the CodeBook's 16.9 MB buffer on 8 B200s NCCL's all-reduce
// takes ~100 us (latency-bound at this size); this path is bounded by 2 x 2.1 MB per rank over the links plus two barriers.
//
// The barriers spin on flags written by OTHER GPUs: each rank is its own process on its own GPU (never several ranks on one
// device), and the spins are bounded -- a rank that never arrives turns into a trap, not a hung box.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx_sm100.cuh"

namespace vq {

constexpr int kArThreads = 512;
constexpr int kArUnroll = 4;             // 16-byte switch reductions in flight per thread (one round trip through the NVSwitch is
                                         // several microseconds: a single outstanding load per thread leaves the links idle)
constexpr int kArMaxBlocks = 16;         // CTAs per rank (all co-resident: they wait for each other).  Measured, 16.9 MB buffer:
                                         // 8 B200s: 61 / 63 / 65 / 72 / 70 us at 8 / 16 / 32 / 64 / 128 CTAs (NCCL: 107 us);
                                         // 2 B200s: 67 / 72 / 74 us at 32 / 64 / 128 (NCCL: 61 us) -- more requests in flight only queue up

__device__ __forceinline__ void ar_put_signal(uint32_t* addr) {               // flag 0 -> 1 on a peer's pad, release
    const long long t0 = clock64();
    uint32_t old;
    do {
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
        if (old != 0u && clock64() - t0 > VQ_MBAR_TIMEOUT_CYCLES) {
            printf("vq_b200: all-reduce barrier (put) timed out, block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    } while (old != 0u);
}
__device__ __forceinline__ void ar_wait_signal(uint32_t* addr) {              // flag 1 -> 0 on my pad, acquire
    const long long t0 = clock64();
    uint32_t old;
    do {
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
        if (old != 1u && clock64() - t0 > VQ_MBAR_TIMEOUT_CYCLES) {
            printf("vq_b200: all-reduce barrier (wait) timed out, block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    } while (old != 1u);
}

// barrier among the ranks, run by ONE CTA per rank: thread t < world signals rank t and waits for rank t (slot t of the pads)
__device__ __forceinline__ void ar_barrier(uint32_t* const* pads, int rank, int world) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        ar_put_signal(pads[threadIdx.x] + rank);
        ar_wait_signal(pads[rank] + threadIdx.x);
    }
    __syncthreads();
}

// local_sync: two words of this rank's own device memory, zero when idle: [0] "the ranks have met, go" flag set by CTA 0,
// [1] count of CTAs whose stores are out.  Only one CTA per rank talks to the other GPUs (remote atomics over NVLink cost
// microseconds each and serialise: with one barrier per CTA the kernel got SLOWER with more CTAs, 66 us at 16 CTAs against
// 76 us at 64 on 8 B200s); the other CTAs synchronise through local memory.  All CTAs are co-resident (<= kArMaxBlocks <= SMs).
__global__ void __launch_bounds__(kArThreads)
vq_allreduce_multimem_kernel(float* __restrict__ mc, uint32_t* const* __restrict__ pads, int rank, int world, int64_t n_vec4,
                             unsigned int* __restrict__ local_sync) {
    __shared__ int is_last;
    // 1. every rank's contribution is in its buffer (stream order puts each rank's producer before this kernel)
    if (blockIdx.x == 0) {
        ar_barrier(pads, rank, world);
        if (threadIdx.x == 0) asm volatile("st.global.release.gpu.u32 [%0], %1;" ::"l"(local_sync), "r"(1u) : "memory");
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            uint32_t go;
            do {
                asm volatile("ld.global.acquire.gpu.u32 %0, [%1];" : "=r"(go) : "l"(local_sync) : "memory");
                if (go == 0u && clock64() - t0 > VQ_MBAR_TIMEOUT_CYCLES) {
                    printf("vq_b200: all-reduce start flag timed out, block %d\n", (int)blockIdx.x);
                    __trap();
                }
            } while (go == 0u);
        }
        __syncthreads();
    }
    // 2. reduce-scatter + all-gather through the switch: this rank owns slice `rank`
    const int64_t slice = n_vec4 / world;                // (host guarantees divisibility)
    float4* base = reinterpret_cast<float4*>(mc) + (int64_t)rank * slice;
    const int64_t stride = (int64_t)gridDim.x * kArThreads;
    for (int64_t i0 = (int64_t)blockIdx.x * kArThreads + threadIdx.x; i0 < slice; i0 += stride * kArUnroll) {
        float4 v[kArUnroll];
#pragma unroll
        for (int u = 0; u < kArUnroll; u++) {
            const int64_t i = i0 + u * stride;
            if (i < slice)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(base + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < kArUnroll; u++) {
            const int64_t i = i0 + u * stride;
            if (i < slice)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             :: "l"(base + i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
        }
    }
    // 3. every slice has landed in every buffer: each CTA fences its stores and checks in; the last one meets the other ranks
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        is_last = (atomicAdd(local_sync + 1, 1u) == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    if (!is_last) return;
    ar_barrier(pads, rank, world);
    if (threadIdx.x == 0) { local_sync[0] = 0u; local_sync[1] = 0u; }      // re-arm for the next call
}

// Tail of the data-parallel exchange buffer in one launch: [hist low 16 bits (K) | hist high bits (K) | loss | 1] as fp32 (each
// count word is exactly representable and exactly summable over ranks, dist.py) -- replaces eight small tensor kernels per step.
__global__ void __launch_bounds__(256)
vq_pack_stats_kernel(const long long* __restrict__ hist, const float* __restrict__ loss, int K, float* __restrict__ tail) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < K) {
        const long long c = hist[i];
        tail[i] = (float)(c & 0xFFFF);
        tail[K + i] = (float)(c >> 16);
    }
    if (i == 0) {
        tail[2 * K] = loss[0];
        tail[2 * K + 1] = 1.0f;
    }
}

}  // namespace vq
