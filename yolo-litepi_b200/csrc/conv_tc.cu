// K2: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// GEMM view per tile: D[128 pixels x Cout] += A[128 x K] * B[K x Cout], K = taps * Cin, fp32
// accumulator in TMEM.  Operands are split-f16 (value = hi + lo): every K-step issues THREE
// tcgen05.mma (Ahi*Bhi + Alo*Bhi + Ahi*Blo), which keeps ~22 mantissa bits -- single-pass
// fp16/bf16/tf32 miss the reference's 1e-2 px box tolerance.
//
// A operand (activations): no im2col is ever materialised.  A tile stages ONE halo patch of its
// 16x8 output pixels in shared memory, laid out [plane][8-channel chunk][patch pixel][16 B].  In the
// UMMA K-major no-swizzle canonical layout ((8,m),2):((16 B,SBO),LBO) the 8 pixels of an output row
// are 8 consecutive 16-B rows of a core matrix, output rows are SBO = patch_pitch*16 B apart and
// channel chunks LBO apart -- so the operand of filter tap (ky,kx) is the SAME patch with a
// different descriptor start address.  Stride-2 convs split the patch into its 4 (row,col) parity
// phases so that a tap again reads 8 consecutive rows.  1x1 convs use a flat 128-pixel "patch".
// B operand (weights): pre-split, pre-packed on the host into the canonical layout per K-block
// (tap x <=64 channels), moved with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx).  When the
// whole layer fits next to the patches it is loaded once per CTA and stays resident; otherwise it
// streams through a ring of stages.
// Epilogue: tcgen05.ld (TMEM lane = pixel) -> bias -> act -> (+residual) -> split -> NHWC stores,
// written at a channel offset of the destination buffer (concat / C2f views are store patterns).
//
// Persistent, warp-specialised CTA (320 threads, one per SM, static round-robin over tiles):
//   warps 0-3  epilogue (TMEM quadrant = warp id)            <- acc_full / -> acc_empty
//   warp  4    weight producer (one lane, bulk TMA)          <- w_empty   / -> w_full
//   warp  5    TMEM allocator + MMA issuer (one lane)        <- patch_full, w_full, acc_empty
//   warps 6-9  patch loaders (cp.async 16 B, zero-fill halo) <- patch_empty / -> patch_full
// Two accumulator stages in TMEM and up to two patch stages let tile i+1 load and tile i-1 drain
// while tile i is in the tensor core.
#include "common.cuh"

namespace {

constexpr int TC_THREADS = 320;
constexpr int TILE_M = 128;
constexpr int TCT_H = 16, TCT_W = 8;       // spatial output tile (rows x cols)
constexpr int MAX_WST = 16;                // weight stages / resident K-blocks

struct TcParams {
    const __half* in;  long long in_plane, in_img;  int in_C, in_coff;
    void* out;         long long out_plane, out_img; int out_C, out_coff, out_fmt;
    const __half* res; long long res_plane, res_img; int res_C, res_coff;
    const uint8_t* wtc;        // packed split weights of this op
    const float* bias;
    int cin, cout, act;
    int H, W, Ho, Wo;          // input / output spatial size
    int n_img, tiles_x, tiles_y, n_tiles;
    int ksize, stride;
    int slots, slots_p;        // patch pixels, padded count (LBO_A = slots_p * 16 B)
    int pitch;                 // patch row pitch in pixels (SBO_A = pitch * 16 B); 1x1: 8
    int phase_slots;           // stride 2: slots per parity phase
    int kb_ch, n_cb, n_kb;     // channels per K-block, channel blocks, total K-blocks (taps * n_cb)
    int w_stages, stage_bytes, resident;
    int patch_stages;
    int tmem_cols, acc_stride;  // TMEM columns allocated; column stride between the two accumulator stages
    unsigned magic_chunks, magic_pitch;   // ceil(2^32 / n) for division by n_chunks / pitch
    long long total_pix;       // 1x1: n_img*H*W
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;       // src-size 0 -> 16 bytes of zeros (halo / padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SmemDescriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | no swizzle
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == LP_ACT_SILU) return v / (1.f + __expf(-v));
    if (act == LP_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

struct TileCoord { int img, oy0, ox0; long long pix0; };

__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int tile) {
    TileCoord t{0, 0, 0, 0};
    if (p.ksize == 1) {
        t.pix0 = (long long)tile * TILE_M;
    } else {
        const int per_img = p.tiles_x * p.tiles_y;
        t.img = tile / per_img;
        const int r = tile - t.img * per_img;
        t.oy0 = (r / p.tiles_x) * TCT_H;
        t.ox0 = (r % p.tiles_x) * TCT_W;
    }
    return t;
}

// smem carve-up: [barriers 512 B][patch ring][weight stages]
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);        // [MAX_WST]
    uint64_t* w_empty = w_full + MAX_WST;                        // [MAX_WST]
    uint64_t* patch_full = w_empty + MAX_WST;                    // [2]
    uint64_t* patch_empty = patch_full + 2;                      // [2]
    uint64_t* acc_full = patch_empty + 2;                        // [2]
    uint64_t* acc_empty = acc_full + 2;                          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* patch0 = smem + 512;
    const int n_chunks = p.cin >> 3;
    const uint32_t plane_bytes = (uint32_t)n_chunks * p.slots_p * 16;
    const uint32_t patch_bytes = 2 * plane_bytes;
    uint8_t* wst = patch0 + (size_t)p.patch_stages * patch_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw_in = p.H * p.W, hw_out = p.Ho * p.Wo;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.w_stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&patch_full[s], 128); mbar_init(&patch_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 6) {
        // ================= patch loaders (128 threads) =================
        const int lt = threadIdx.x - 192;
        const int items_per_plane = n_chunks * p.slots;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int ps = it % p.patch_stages;
            if (it >= p.patch_stages) mbar_wait(&patch_empty[ps], ((it / p.patch_stages) - 1) & 1);
            const TileCoord tc = tile_coord(p, tile);
            const uint32_t dst0 = smem_u32(patch0 + (size_t)ps * patch_bytes);
            for (int e = lt; e < items_per_plane; e += 128) {
                const int slot = (int)__umulhi((unsigned)e, p.magic_chunks);       // e / n_chunks
                const int chunk = e - slot * n_chunks;
                bool valid;
                long long goff;                               // element offset of the pixel inside a plane
                if (p.ksize == 1) {
                    const long long gpix = tc.pix0 + slot;
                    valid = gpix < p.total_pix;
                    const int gi = valid ? (int)(gpix / hw_in) : 0;
                    const int pin = valid ? (int)(gpix - (long long)gi * hw_in) : 0;
                    goff = (long long)gi * p.in_img + (long long)pin * p.in_C;
                } else {
                    int py, px;
                    if (p.stride == 1) {
                        py = (int)__umulhi((unsigned)slot, p.magic_pitch);           // slot / pitch
                        px = slot - py * p.pitch;
                    } else {
                        const int ph = slot / p.phase_slots, q = slot - ph * p.phase_slots;
                        const int sr = (int)__umulhi((unsigned)q, p.magic_pitch), sc = q - sr * p.pitch;
                        py = 2 * sr + (ph >> 1);
                        px = 2 * sc + (ph & 1);
                    }
                    const int iy = tc.oy0 * p.stride - 1 + py, ix = tc.ox0 * p.stride - 1 + px;
                    valid = (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W);
                    goff = (long long)tc.img * p.in_img + (valid ? ((long long)iy * p.W + ix) * p.in_C : 0);
                }
                const __half* src = p.in + goff + p.in_coff + chunk * 8;
                const uint32_t dst = dst0 + ((uint32_t)chunk * p.slots_p + slot) * 16;
                cp_async16(dst, src, valid);
                cp_async16(dst + plane_bytes, src + p.in_plane, valid);
            }
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&patch_full[ps]);
        }
    } else if (warp == 4) {
        // ================= weight producer (bulk TMA) =================
        if (lane == 0) {
            int g = 0;                                        // running K-block counter across tiles
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                if (p.resident && tile != (int)blockIdx.x) break;
                for (int kb = 0; kb < p.n_kb; ++kb, ++g) {
                    const int s = g % p.w_stages;
                    if (g >= p.w_stages) mbar_wait(&w_empty[s], ((g / p.w_stages) - 1) & 1);
                    mbar_expect_tx(&w_full[s], (uint32_t)p.stage_bytes);
                    bulk_g2s(wst + (size_t)s * p.stage_bytes, p.wtc + (size_t)kb * p.stage_bytes, (uint32_t)p.stage_bytes, &w_full[s]);
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.cout >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            const uint32_t lbo_a = (uint32_t)p.slots_p * 16, sbo_a = (uint32_t)p.pitch * 16;
            const uint32_t lbo_b = (uint32_t)p.cout * 16, sbo_b = 128;
            const uint32_t wst_s = smem_u32(wst);
            const uint32_t b_plane = (uint32_t)p.kb_ch * p.cout * 2;              // bytes of one plane inside a stage
            int it = 0, g = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int ps = it % p.patch_stages, as = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);
                mbar_wait(&patch_full[ps], (it / p.patch_stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t patch_s = smem_u32(patch0 + (size_t)ps * patch_bytes);
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride);
                uint32_t acc = 0;
                for (int kb = 0; kb < p.n_kb; ++kb, ++g) {
                    const int s = p.resident ? kb : g % p.w_stages;
                    if (!p.resident || it == 0) {
                        mbar_wait(&w_full[s], p.resident ? 0 : (g / p.w_stages) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const int tap = kb / p.n_cb, cb = kb - tap * p.n_cb;
                    int tap_slot = 0;
                    if (p.ksize == 3) {
                        const int ky = tap / 3, kx = tap - ky * 3;
                        if (p.stride == 1) tap_slot = ky * p.pitch + kx;
                        else tap_slot = ((ky & 1) * 2 + (kx & 1)) * p.phase_slots + (ky >> 1) * p.pitch + (kx >> 1);
                    }
                    const uint32_t a_hi = patch_s + ((uint32_t)(cb * (p.kb_ch >> 3)) * p.slots_p + tap_slot) * 16;
                    const uint32_t a_lo = a_hi + plane_bytes;
                    const uint32_t b_hi = wst_s + (uint32_t)s * p.stage_bytes;
                    const uint32_t b_lo = b_hi + b_plane;
                    for (int ks = 0; ks < (p.kb_ch >> 4); ++ks) {
                        const uint32_t ao = (uint32_t)ks * 2 * lbo_a, bo = (uint32_t)ks * 2 * lbo_b;
                        const uint64_t dah = umma_desc(a_hi + ao, lbo_a, sbo_a), dal = umma_desc(a_lo + ao, lbo_a, sbo_a);
                        const uint64_t dbh = umma_desc(b_hi + bo, lbo_b, sbo_b), dbl = umma_desc(b_lo + bo, lbo_b, sbo_b);
                        umma_f16(d_tmem, dah, dbh, idesc, acc);
                        acc = 1;
                        umma_f16(d_tmem, dal, dbh, idesc, 1);
                        umma_f16(d_tmem, dah, dbl, idesc, 1);
                    }
                    if (!p.resident) umma_commit(&w_empty[s]);   // frees the weight stage once these MMAs retire
                }
                umma_commit(&patch_empty[ps]);
                umma_commit(&acc_full[as]);
            }
        }
    } else {
        // ================= epilogue (warps 0-3) =================
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const TileCoord tc = tile_coord(p, tile);
            const int r = warp * 32 + lane;                  // accumulator row == TMEM lane == tile pixel
            bool valid;
            int oimg, opin;                                  // image and pixel-in-image of this row
            if (p.ksize == 1) {
                const long long gp = tc.pix0 + r;
                valid = gp < p.total_pix;
                oimg = valid ? (int)(gp / hw_out) : 0;
                opin = valid ? (int)(gp - (long long)oimg * hw_out) : 0;
            } else {
                const int oy = tc.oy0 + (r >> 3), ox = tc.ox0 + (r & 7);
                valid = (oy < p.Ho && ox < p.Wo);
                oimg = tc.img;
                opin = oy * p.Wo + ox;
            }
            const long long obase = (long long)oimg * p.out_img + (long long)opin * p.out_C + p.out_coff;
            const long long rbase = (long long)oimg * p.res_img + (long long)opin * p.res_C + p.res_coff;
            mbar_wait(&acc_full[as], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.acc_stride);
            for (int c0 = 0; c0 < p.cout; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!valid) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int cc = c0 + 8 * h;
                    float f[8];
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cc));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cc + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = act_fn(__uint_as_float(v[8 * h + i]) + bb[i], p.act);
                    if (p.res) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(p.res + rbase + cc);
                        const uint4 rl = *reinterpret_cast<const uint4*>(p.res + p.res_plane + rbase + cc);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
                        const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = __half22float2(h2[i]), b = __half22float2(l2[i]);
                            f[2 * i] += a.x + b.x;
                            f[2 * i + 1] += a.y + b.y;
                        }
                    }
                    if (p.out_fmt == LP_FMT_SPLIT16) {
                        uint4 oh, ol;
                        __half2* h2 = reinterpret_cast<__half2*>(&oh);
                        __half2* l2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            __half a0, a1, b0h, b1h;
                            split_make(f[2 * i], a0, b0h);
                            split_make(f[2 * i + 1], a1, b1h);
                            h2[i] = __halves2half2(a0, a1);
                            l2[i] = __halves2half2(b0h, b1h);
                        }
                        __half* o = reinterpret_cast<__half*>(p.out) + obase + cc;
                        *reinterpret_cast<uint4*>(o) = oh;
                        *reinterpret_cast<uint4*>(o + p.out_plane) = ol;
                    } else {
                        float* o = reinterpret_cast<float*>(p.out) + obase + cc;
                        *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                        *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[as]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

unsigned magic_u32(int n) { return (unsigned)((0x100000000ull + (unsigned)n - 1) / (unsigned)n); }

}  // namespace

// returns 1 if the op ran on the tensor cores, 0 if it is not eligible, <0 on error
int lp_conv_tc_try(lp_ctx* ctx, lp_net_plan& net, const lp_op_desc& op, int batch, uint8_t* ws, cudaStream_t st) {
    const lp_buf_desc& ib = net.bufs[op.in_buf];
    const lp_buf_desc& ob = net.bufs[op.out_buf];
    if (op.kind != LP_OP_CONV || ib.fmt != LP_FMT_SPLIT16) return 0;
    if (!((op.ksize == 1 && op.stride == 1) || (op.ksize == 3 && (op.stride == 1 || op.stride == 2)))) return 0;
    if (op.cin % 16 || op.cout % 16 || op.cout > 128 || op.cin > 512 || op.in_coff % 8 || op.out_coff % 4) return 0;
    if (op.out_cstride > 1) return 0;
    if (ob.fmt == LP_FMT_SPLIT16 && op.out_coff % 8) return 0;

    TcParams p{};
    const long long in_img = ib.image_bytes / 2;
    p.in = reinterpret_cast<const __half*>(ws + ib.offset);
    p.in_plane = (long long)net.max_batch * in_img; p.in_img = in_img; p.in_C = ib.c; p.in_coff = op.in_coff;
    const int oesz = ob.fmt == LP_FMT_SPLIT16 ? 2 : 4;
    p.out = ws + ob.offset + (size_t)op.row_off * ob.c * oesz;
    p.out_img = ob.image_bytes / oesz; p.out_plane = (long long)net.max_batch * p.out_img;
    p.out_C = ob.c; p.out_coff = op.out_coff; p.out_fmt = ob.fmt;
    if (op.res_buf >= 0) {
        const lp_buf_desc& rb = net.bufs[op.res_buf];
        if (rb.fmt != LP_FMT_SPLIT16 || op.res_coff % 8) return 0;
        p.res = reinterpret_cast<const __half*>(ws + rb.offset);
        p.res_img = rb.image_bytes / 2; p.res_plane = (long long)net.max_batch * p.res_img; p.res_C = rb.c; p.res_coff = op.res_coff;
    }
    p.wtc = net.weights_tc + op.wtc_off;
    p.bias = net.weights + op.b_off;
    p.cin = op.cin; p.cout = op.cout; p.act = op.act;
    p.H = ib.h; p.W = ib.w; p.n_img = batch; p.ksize = op.ksize; p.stride = op.stride;
    p.Ho = (ib.h + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1;
    p.Wo = (ib.w + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1;
    p.kb_ch = 0;
    for (int d : {64, 48, 32, 16}) if (op.cin % d == 0) { p.kb_ch = d; break; }     // same rule as plan.py pack_tc_weights
    if (!p.kb_ch) return 0;
    p.n_cb = op.cin / p.kb_ch;
    const int taps = op.ksize * op.ksize;
    p.n_kb = taps * p.n_cb;
    if (op.ksize == 1) { p.slots = TILE_M; p.pitch = 8; }
    else if (op.stride == 1) { p.pitch = TCT_W + 2; p.slots = (TCT_H + 2) * p.pitch; }
    else { p.pitch = TCT_W + 1; p.phase_slots = (TCT_H + 1) * p.pitch; p.slots = 4 * p.phase_slots; }
    p.slots_p = p.slots + ((9 - (p.slots & 7)) & 7);                 // == 1 (mod 8): conflict-free chunk stride
    p.magic_chunks = magic_u32(op.cin / 8);
    p.magic_pitch = magic_u32(p.pitch);
    p.stage_bytes = p.kb_ch * p.cout * 4;                            // 2 planes x kb_ch x cout x 2 B
    const size_t patch_bytes = (size_t)2 * (op.cin / 8) * p.slots_p * 16;
    const size_t budget = 220 * 1024 - 512;
    const size_t w_all = (size_t)p.n_kb * p.stage_bytes;
    // preference: 2 patch stages + resident weights > 2 patch stages + >=2 streaming stages > 1 patch stage
    p.patch_stages = 0;
    for (int ps = 2; ps >= 1 && !p.patch_stages; --ps) {
        if (ps * patch_bytes > budget) continue;
        const size_t left = budget - ps * patch_bytes;
        if (w_all <= left && p.n_kb <= MAX_WST) { p.patch_stages = ps; p.resident = 1; p.w_stages = p.n_kb; }
        else if (left >= 2 * (size_t)p.stage_bytes) {
            p.patch_stages = ps; p.resident = 0;
            int s = (int)(left / p.stage_bytes);
            p.w_stages = s > 4 ? 4 : s;
            if (p.w_stages > p.n_kb) p.w_stages = p.n_kb;
        }
    }
    if (!p.patch_stages) return 0;
    p.acc_stride = op.cout < 32 ? 32 : op.cout;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.acc_stride) p.tmem_cols <<= 1;
    const size_t smem = 512 + (size_t)p.patch_stages * patch_bytes + (size_t)p.w_stages * p.stage_bytes;

    if (op.ksize == 1) {
        p.total_pix = (long long)batch * ib.h * ib.w;
        p.n_tiles = (int)((p.total_pix + TILE_M - 1) / TILE_M);
    } else {
        p.tiles_x = (p.Wo + TCT_W - 1) / TCT_W;
        p.tiles_y = (p.Ho + TCT_H - 1) / TCT_H;
        p.n_tiles = p.tiles_x * p.tiles_y * batch;
    }
    static bool attr_set = false;
    if (!attr_set) {
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    const int grid = p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count;
    conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        lp_set_error("conv_tc launch failed: %s (smem %zu)", cudaGetErrorString(e), smem);
        return -2;
    }
    return 1;
}
