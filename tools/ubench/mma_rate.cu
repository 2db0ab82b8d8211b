// Microbenchmark: cycles per tcgen05.mma (kind::f16, cta_group::1, M=128, K=16) for the no-swizzle
// K-major layout as a function of N, of the A-operand stride (SBO 128 B = aligned core matrices vs
// 160 B = the shifted-window halo patch of conv_tc.cu) and of how many distinct accumulators are cycled.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128, 1) k(int N, uint32_t sbo_a, uint32_t lbo_a, int iters, int nacc, int distinct_desc, uint32_t layout, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // 1.0 halves
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        const uint64_t da0 = umma_desc(a0, lbo_a, sbo_a, layout), da1 = umma_desc(a0 + 2 * lbo_a, lbo_a, sbo_a, layout);
        const uint64_t db0 = umma_desc(b0, layout ? 16 : (uint32_t)N * 16, layout ? 1024 : 128, layout);
        const uint64_t db1 = umma_desc(b0 + 2 * N * 16, layout ? 16 : (uint32_t)N * 16, layout ? 1024 : 128, layout);
        const uint32_t d0 = tb, d1 = tb + (nacc > 1 ? N : 0);
        long long t0 = clock64();
#define MMA(D, A, B) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(D), "l"(A), "l"(B), "r"(idesc), "r"(1) : "memory")
        if (distinct_desc == 2) {
            // fresh operands for every MMA pair, like the conv kernel: A advances 2*LBO per K-step over a 47 KB
            // patch (two planes), B advances through 64 KB of weights
            const uint32_t a_step = (2 * lbo_a) >> 4, plane = (8 * lbo_a) >> 4, b_step = (2u * N * 32) >> 4;
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
            const uint64_t dbb = umma_desc(b0, (uint32_t)N * 32, 128, 0);
            for (int i = 0; i < iters; i += 8) {
                uint32_t ao = ((uint32_t)(i / 8) % 9) * 1, bo = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    MMA(d0, da0 + ao, dbb + bo);   /* N here is idesc's N; uses idesc (N) for simplicity */
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d0), "l"(da0 + ao + plane), "l"(dbb + bo), "r"(idesc2), "r"(1) : "memory");
                    ao += a_step; bo += b_step;
                }
            }
        } else
        for (int i = 0; i < iters; i += 8) {
            if (distinct_desc) { MMA(d0, da0, db0); MMA(d1, da1, db0); MMA(d0, da0, db1); MMA(d1, da1, db1); MMA(d0, da1, db0); MMA(d1, da0, db1); MMA(d0, da1, db1); MMA(d1, da0, db0); }
            else { MMA(d0, da0, db0); MMA(d1, da0, db0); MMA(d0, da0, db0); MMA(d1, da0, db0); MMA(d0, da0, db0); MMA(d1, da0, db0); MMA(d0, da0, db0); MMA(d1, da0, db0); }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) *out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main(int argc, char** argv) {
    if (argc > 1) {
        // LBO sweep of the conv pattern (dd = 2): mma_rate sweep
        long long* d; cudaMalloc(&d, 8);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        printf("N lbo16 (mod 8) cycles/mma\n");
        for (int N : {32, 64, 128})
            for (uint32_t l16 : {176u, 177u, 178u, 179u, 180u, 181u, 182u, 183u, 184u, 185u})
                for (uint32_t sbo : {128u, 160u}) {
                    k<<<148, 128, 160 * 1024>>>(N, sbo, l16 * 16, 2000, 1, 2, 0, d);
                    cudaDeviceSynchronize();
                    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("%3d %3u %u sbo %u %8.1f\n", N, l16, l16 % 8, sbo, (double)c / 2000);
                }
        return 0;
    }
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("layout N sbo_a nacc distinct cycles/mma  MAC/clk\n");
    for (uint32_t layout : {0u, 2u})
    for (int N : {16, 32, 64, 128, 256})
        for (uint32_t sbo : {128u, 160u})
            for (int nacc : {1, 2})
                for (int dd : {0, 1, 2}) {
                    if (dd == 2 && (layout != 0 || N > 128 || nacc != 1)) continue;
                    if (layout == 2 && sbo != 128) continue;
                    if (nacc * N > 512) continue;
                    const uint32_t lbo = layout ? 16 : 2896, sb = layout ? 1024 : sbo;
                    k<<<148, 128, 160 * 1024>>>(N, sb, lbo, iters, nacc, dd, layout, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                    printf("%s %3d %3u %d %d %8.1f %8.0f %s\n", layout ? "SW128" : "NONE ", N, sbo, nacc, dd, (double)c / iters, 128.0 * N * 16 * iters / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
