// tcgen05.mma issue-rate microbenchmark, CTA-pair form (cta_group::2, kind::f16, SS operands, M = 256 over two SMs):
// clusters of two CTAs, the leader (cluster rank 0) issues ITERS x KSTEPS MMAs of shape 256 x N x 16 -- each CTA holds
// its own 128 A rows and its half (N/2 rows) of B in shared memory at the same offsets, each CTA's tensor memory
// receives its 128 rows of D -- and commits with the multicast form to a barrier in BOTH CTAs.  Garbage operands: only
// the pipe and the protocol (cluster launch, cta_group::2 alloc / mma / commit / dealloc) are exercised.
// Compare with umma_rate.cu (cta_group::1: 128 x N x 16 takes ~N clk = half the dense peak).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate2 umma_rate2.cu && ./umma_rate2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0; !ok; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && (it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000LL) __trap();       // fault instead of hanging the device
        }
    }
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, 8-row groups of 1024 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int N, int KSTEPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate2_kernel(int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];         // A: [128 x 64] fp16 (16 KiB) | B half: [N/2 x 64] fp16 (<= 16 KiB)
    __shared__ uint64_t bar[2];                               // one per accumulator; the multicast commit arrives in both CTAs
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar[0]), 1);
        mbar_init(smem_u32(&bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                          // (every CTA of the pair runs the allocation)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();                                           // both CTAs' barriers and tensor memory are ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // instruction descriptor: D fp32 (bit 4), A/B fp16 K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    if (threadIdx.x == 32 && rank == 0) {                     // the pair's MMA thread
        const uint64_t a0 = umma_desc(base), b0 = umma_desc(base + 16384);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (uint32_t)((it & 1) * N);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t ad = a0 + (uint64_t)((k & 3) * 2), bd = b0 + (uint64_t)((k & 3) * 2);
                const uint32_t acc = k ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                             : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar[it & 1])), "h"((uint16_t)3) : "memory");
            if (it >= 1) mbar_wait(smem_u32(&bar[(it - 1) & 1]), (uint32_t)(((it - 1) >> 1) & 1));
        }
        mbar_wait(smem_u32(&bar[(iters - 1) & 1]), (uint32_t)(((iters - 1) >> 1) & 1));
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    }
    if (threadIdx.x == 32 && rank == 1) {                     // the peer sees every commit through the multicast
        for (int it = 0; it < iters; ++it) mbar_wait(smem_u32(&bar[it & 1]), (uint32_t)((it >> 1) & 1));
        if (blockIdx.x == 1) cycles[1] = iters;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

template <int N, int KSTEPS>
static void run(int iters, long long* d_cycles) {
    const size_t smem = 1024 + 16384 + 16384;
    cudaFuncSetAttribute(umma_rate2_kernel<N, KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_cycles, 0, 2 * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    umma_rate2_kernel<N, KSTEPS><<<148, 128, smem>>>(iters, d_cycles);
    cudaEventRecord(e0);
    umma_rate2_kernel<N, KSTEPS><<<148, 128, smem>>>(iters, d_cycles);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[2] = {0, 0};
    cudaMemcpy(cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 256 * N * 16 * (double)KSTEPS * iters * 74;
    printf("N %3d  K-steps/commit %2d : %.1f us  %.0f TFLOP/s  (%.1f cycles per 256 x %d x 16 MMA on pair 0; peer saw %lld commits)  %s\n",
           N, KSTEPS, ms * 1e3, flop / (ms * 1e-3) / 1e12, (double)cyc[0] / ((double)KSTEPS * iters), N, cyc[1],
           cudaGetErrorString(err));
}

int main() {
    long long* d_cycles;
    cudaMalloc(&d_cycles, 2 * sizeof(long long));
    run<128, 8>(2000, d_cycles);
    run<128, 24>(700, d_cycles);
    run<224, 8>(2000, d_cycles);
    run<224, 24>(700, d_cycles);
    run<256, 8>(2000, d_cycles);
    run<256, 24>(700, d_cycles);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
