// Microbenchmark: issue-side cost of the conv_tc MMA loop.  One warp runs the tile/tap/K-step loops of the
// conv kernel (no barriers, operands are whatever shared memory holds) in several code shapes; prints
// cycles per tcgen05.mma so the loop shape that keeps the tensor pipe fed can be chosen.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_pred(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc, uint32_t issue) {
    asm volatile("{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\tsetp.ne.b32 q, %7, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
struct Prm { int cout, ktot, taps, pitch, slots_p, n_tiles, variant, patch_stages, acc_stages; long long* out; };

__global__ void __launch_bounds__(128, 1) k(const Prm p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
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
    if (warp == 1) {
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * p.cout) >> 3) << 17) | (8u << 24);
        const uint32_t idesc1 = (1u << 4) | ((uint32_t)(p.cout >> 3) << 17) | (8u << 24);
        const uint32_t lbo_a = (uint32_t)p.slots_p * 16, sbo_a = (uint32_t)p.pitch * 16;
        const uint32_t lbo_b = (uint32_t)p.cout * 32;
        const uint32_t n_chunks = 2 * p.ktot;
        const uint32_t plane_bytes = n_chunks * p.slots_p * 16, patch_bytes = 2 * plane_bytes;
        const uint64_t da_base = umma_desc(0, lbo_a, sbo_a), db_base = umma_desc(0, lbo_b, 128);
        const uint32_t da_hi = (uint32_t)(da_base >> 32), db_hi = (uint32_t)(db_base >> 32);
        const uint32_t wst = smem_u32(smem) + p.patch_stages * patch_bytes;
        const uint32_t b016 = (uint32_t)db_base + (wst >> 4);
        const uint32_t plane16 = plane_bytes >> 4, a_step16 = (2 * lbo_a) >> 4, b_step16 = (2 * lbo_b) >> 4;
        const uint32_t patch016 = (uint32_t)da_base + (smem_u32(smem) >> 4), patch_stride16 = patch_bytes >> 4;
        const uint32_t acc_stride = 2 * p.cout < 32 ? 32 : 2 * p.cout;
        const uint32_t leader = elect_one() ? 1u : 0u;
        long long t0 = clock64();
        if (p.variant == 0) {
            // the kernel's resident loop: whole warp, predicated issue, 32-bit progressions
            uint32_t ps = 0, as = 0;
            for (int it = 0; it < p.n_tiles; ++it) {
                const uint32_t patch16 = patch016 + ps * patch_stride16;
                const uint32_t d_tmem = tb + as * acc_stride;
                uint32_t acc = 0, b16 = b016;
                for (int tap = 0; tap < p.taps; ++tap) {
                    uint32_t a16 = patch16;
                    if (p.taps == 9) { const int ky = tap / 3, kx = tap - ky * 3; a16 += (uint32_t)(ky * p.pitch + kx); }
                    for (int kk = 0; kk < p.ktot; ++kk) {
                        umma_pred(d_tmem, a16, da_hi, b16, db_hi, idesc2, acc, leader);
                        umma_pred(d_tmem, a16 + plane16, da_hi, b16, db_hi, idesc1, 1, leader);
                        acc = 1; a16 += a_step16; b16 += b_step16;
                    }
                }
                __syncwarp();
                if (++ps == (uint32_t)p.patch_stages) ps = 0;
                if (++as == (uint32_t)p.acc_stages) as = 0;
            }
        } else if (p.variant == 1) {
            // single elected thread, 64-bit descriptors
            if (leader) {
                uint32_t ps = 0, as = 0;
                const uint64_t da0 = da_base + (smem_u32(smem) >> 4), db0 = db_base + (wst >> 4);
                for (int it = 0; it < p.n_tiles; ++it) {
                    const uint64_t dpatch = da0 + ps * patch_stride16;
                    const uint32_t d_tmem = tb + as * acc_stride;
                    uint32_t acc = 0; uint64_t db = db0;
                    for (int tap = 0; tap < p.taps; ++tap) {
                        uint64_t da = dpatch;
                        if (p.taps == 9) { const int ky = tap / 3, kx = tap - ky * 3; da += (uint32_t)(ky * p.pitch + kx); }
                        for (int kk = 0; kk < p.ktot; ++kk) {
                            umma1(d_tmem, da, db, idesc2, acc);
                            umma1(d_tmem, da + plane16, db, idesc1, 1);
                            acc = 1; da += a_step16; db += b_step16;
                        }
                    }
                    if (++ps == (uint32_t)p.patch_stages) ps = 0;
                    if (++as == (uint32_t)p.acc_stages) as = 0;
                }
            }
            __syncwarp();
        } else if (p.variant == 2) {
            // same MMAs with constant operands (no address arithmetic at all): the pure pipe rate
            if (leader) {
                const uint64_t da0 = da_base + (smem_u32(smem) >> 4), db0 = db_base + (wst >> 4);
                const int n = p.n_tiles * p.taps * p.ktot;
                for (int i = 0; i < n; ++i) {
                    umma1(tb, da0, db0, idesc2, 1);
                    umma1(tb, da0 + plane16, db0, idesc1, 1);
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (leader) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0 && leader) { p.out[0] = t1 - t0; p.out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("variant cout ktot taps | issue cyc/MMA | total cyc/MMA\n");
    for (int variant : {0, 1, 2})
        for (int taps : {9, 1})
            for (int cout : {16, 32, 64, 128})
                for (int ktot : {1, 2, 4}) {
                    Prm p{cout, ktot, taps, 10, 185, 22, variant, 2, 512 / (2 * cout < 32 ? 32 : 2 * cout) > 4 ? 4 : 512 / (2 * cout < 32 ? 32 : 2 * cout), d};
                    if (2 * (2 * ktot) * 185 * 16 * 2 + taps * ktot * cout * 64 > 200 * 1024) continue;
                    k<<<148, 128, 200 * 1024>>>(p);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c[2]; cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
                    const double n = 22.0 * taps * ktot * 2;
                    printf("%d %3d %d %d | %7.1f | %7.1f %s\n", variant, cout, ktot, taps, c[0] / n, c[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    return 0;
}
