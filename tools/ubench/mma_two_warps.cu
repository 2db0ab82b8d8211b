// Microbenchmark: do two warps that issue tcgen05.mma concurrently (different TMEM accumulators) overlap?
// Mode 1: one warp issues 2*iters MMAs alternating between two accumulators.  Mode 2: two warps issue `iters` each.
// Printed: cycles per MMA over the whole CTA.  (Round-2 finding: see profiles/r2_notes.md.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
#define MMA(D, A, B, I) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D), "l"(A), "l"(B), "r"(I), "r"(1) : "memory")
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int mode, int same_operands, int dist, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    __shared__ long long t_end[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    const long long t0 = clock64();
    const bool active = lane == 0 && (warp == 2 || (warp == 3 && mode == 2));
    if (active) {
        const int w = warp - 2;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem) + (same_operands ? 0 : w * 48 * 1024), b0 = smem_u32(smem) + 96 * 1024 + (same_operands ? 0 : w * 32 * 1024);
        const uint64_t da = umma_desc(a0, 2896, 160), db = umma_desc(b0, (uint32_t)N * 16, 128);
        const uint32_t d0 = tb + w * dist, d1 = d0 + (mode == 1 ? dist : 0);
        const int n = mode == 1 ? 2 * iters : iters;
        for (int i = 0; i < n; i += 4) {
            MMA(d0, da, db, idesc); MMA(d1, da + 8, db + 16, idesc); MMA(d0, da + 16, db, idesc); MMA(d1, da + 24, db + 16, idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar[w])) : "memory");
        t_end[w] = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = (mode == 2 && t_end[1] > t_end[0] ? t_end[1] : t_end[0]) - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("N mode(1 = one warp, 2 = two warps) accumulator_distance_cols cycles_per_mma (4000 MMAs per CTA in both modes)\n");
    for (int N : {16, 32, 64, 128})
        for (int mode : {1, 2})
            for (int dist : {N, 2 * N, 64, 128, 256}) {
                if (dist < N || dist + N > 512) continue;
                k<<<148, 128, 160 * 1024>>>(N, iters, mode, 1, dist, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("%3d %d %3d %8.1f %s\n", N, mode, dist, (double)c / (2 * iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
