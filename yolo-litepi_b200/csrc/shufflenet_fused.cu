// K7 fused: the whole ShuffleNetV2 x1.0 forward inside ONE persistent CTA per SM.
// Replaces self.model(batch) (src/vntsr/pipeline/e2e.py:393; torchvision shufflenetv2.py) for the 64x64
// classifier input.  The layer-by-layer plan (71 launches, a few thousand pixels each) is launch- and
// latency-bound; here every activation of a ROI lives in shared memory (<= 30 KB per ROI after the stem),
// the folded weights stream from L2 into a two-stage shared-memory ring with 1-D bulk TMA (cp.async.bulk +
// mbarrier) issued by a dedicated producer warp that walks the same step list ahead of the compute warps,
// and the only other global traffic is the 12 KB u8 crop in, 15 KB parked between middle and tail, and C
// logits out.
//
// The CTA executes a host-built program (plan.py build_fused_classifier -> FusedProgram), three step lists:
//   front,  per ROI : conv1 3x3 s2 (+ToTensor/Normalize via a 256-entry table) -> maxpool 3x3 s2 ->
//                     stage2 unit 0 (the only unit whose intermediates exceed 30 KB); fp32 FMA, 2 x 24 KB stages
//   middle, per ROI : stage2 units 1-3, stage3.  Pointwise layers on the tensor cores (mma.sync m16n8k16,
//                     split-f16 three-product, fp32 accumulate); weight stages grow over the dead front buffers;
//                     the 4x4x232 result is parked in global memory
//   tail,   per CTA : stage4, conv5, global mean, fc for the CTA's ROIs (up to `tail_group`) stacked as GEMM
//                     rows, so 73 % of the weight bytes are read once per CTA instead of once per ROI
// channel_shuffle is the store pattern of the producers (dst channel = off + j * 2), chunk is a view.
// A phase mbarrier keeps the producer from starting a pass before the compute threads finished the previous
// one (the stage regions of different passes overlap buffers of other passes).
#include "common.cuh"
#include <stdlib.h>

enum { FS_CONV1 = 0, FS_MAXPOOL = 1, FS_PW = 2, FS_DW = 3, FS_COPY = 4, FS_MEANFC = 5, FS_STORE = 6, FS_LOAD = 7 };

struct FStep {
    int op;
    int src, dst;              // float offsets into shared memory
    int src_C, src_off;        // channels per pixel of the source tensor, first channel read
    int dst_C, dst_off, dst_cs;
    int cin, cout;
    int H, W, stride;          // input spatial size
    int relu;
    int w_off, b_off;          // float offsets into the weight blob
    int dst_roi_stride;        // FS_STORE / FS_LOAD: floats per ROI parked in global memory between middle and tail
    int w16_off;               // > 0: split-f16 weights of this pointwise layer in the fp16 blob, in units of 16 B (+1); 0: none
};

constexpr int FUSED_THREADS = 512;            // compute threads; one more warp streams weights
constexpr int FUSED_BLOCK = FUSED_THREADS + 32;
constexpr int WBUF_FLOATS = 6144;             // 24 KB per weight stage, two stages
constexpr int FUSED_MAX_STEPS = 96;
// barrier among the compute threads only (the producer warp never joins it)
#define CSYNC() asm volatile("bar.sync 1, %0;" ::"n"(512) : "memory")

__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FD_%=;\n\t"
        "bra FW_%=;\n\t"
        "FD_%=:\n\t}" ::"r"(f_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void f_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(f_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(f_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t f_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void f_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same-named mbarrier of CTA `rank` of the cluster
__device__ __forceinline__ void f_mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(f_smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void f_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "CW_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CD_%=;\n\t"
        "bra CW_%=;\n\t"
        "CD_%=:\n\t}" ::"r"(f_smem_u32(bar)), "r"(parity) : "memory");
}
// one L2 read, delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void f_bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(f_smem_u32(dst)), "l"(src), "r"(bytes), "r"(f_smem_u32(bar)), "h"(mask) : "memory");
}

// ---- tensor-core pointwise layers (back end) --------------------------------------------------------------
// fp16 weight rows are [cout_p8][plane hi|lo][L] halves with L = cin_p16 + pad, L == 4 (mod 32): the stride
// between output channels is L 32-bit words, so the 8 channels x 4 k-pairs a warp reads for a B fragment hit
// 32 different banks.
__device__ __forceinline__ int w16_row_halves(int cin) {
    const int cin_p = (cin + 15) & ~15;
    return cin_p + ((4 - cin_p) & 31);
}
// output channels per weight stage (multiple of 8)
__device__ __forceinline__ int chunk_couts(int cin, int cout, int slot_floats) {
    const int cout_p = (cout + 7) & ~7;
    const int n = ((slot_floats * 4) / (w16_row_halves(cin) * 4)) & ~7;
    return n < cout_p ? n : cout_p;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(f_smem_u32(p)));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Pointwise conv on the tensor cores (legacy warp-level mma.sync: the tiles are 16 x 8, far below a tcgen05
// tile).  fp32 parity comes from the same split as the detector: x = hi + lo in fp16, and
// Ahi*Bhi + Alo*Bhi + Ahi*Blo accumulated in fp32.  Step 1: all threads convert the fp32 activations
// [rows][cin] to two fp16 planes [rows_p16][cin_p16 + 8] (row stride == 4 words mod 32: conflict-free
// ldmatrix).  Step 2: weights arrive in chunks of output channels (every chunk holds complete K), a warp owns
// 16x8 output tiles.  Step 3: bias/ReLU and the channel-shuffle store pattern straight from the fragments.
__device__ __forceinline__ void pw_layer_mma(const float* __restrict__ in, int in_C, float* __restrict__ out, int out_C, int dst_cs,
                                             int rows, int cin, int cout, const float* __restrict__ bias, int relu,
                                             __half* __restrict__ astage, const float* __restrict__ wbuf, int slot_floats,
                                             uint64_t* full, uint64_t* empty, uint32_t& chunk_ctr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_p = (rows + 15) & ~15, cin_p = (cin + 15) & ~15, cout_p = (cout + 7) & ~7;
    const int LA = cin_p + 8, L = w16_row_halves(cin);
    __half* Ah = astage;
    __half* Al = astage + rows_p * LA;
    for (int e = tid; e < rows_p * (cin_p >> 1); e += FUSED_THREADS) {
        const int r = e / (cin_p >> 1), c = (e - r * (cin_p >> 1)) * 2;
        float x0 = 0.f, x1 = 0.f;
        if (r < rows) {
            if (c < cin) x0 = in[r * in_C + c];
            if (c + 1 < cin) x1 = in[r * in_C + c + 1];
        }
        const __half2 hi = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(Ah + r * LA + c) = hi;
        *reinterpret_cast<__half2*>(Al + r * LA + c) = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    }
    CSYNC();
    const int m_tiles = rows_p >> 4, ksteps = cin_p >> 4;
    const int NN = chunk_couts(cin, cout, slot_floats);
    for (int n0 = 0; n0 < cout_p; n0 += NN, ++chunk_ctr) {
        const int nn = min(NN, cout_p - n0);
        const uint32_t slot = chunk_ctr & 1;
        f_mbar_wait(&full[slot], (chunk_ctr >> 1) & 1);
        const __half* Wc = reinterpret_cast<const __half*>(wbuf + slot * slot_floats);      // [nn][2][L]
        const int n_tiles = nn >> 3;
        for (int item = warp; item < m_tiles * n_tiles; item += FUSED_THREADS / 32) {
            const int mt = item / n_tiles, nt = item - mt * n_tiles;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const __half* arow_h = Ah + (mt * 16 + (lane & 15)) * LA + (lane >> 4) * 8;
            const __half* arow_l = arow_h + rows_p * LA;
            const __half* brow = Wc + (size_t)(nt * 8 + (lane >> 2)) * 2 * L + (lane & 3) * 2;
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks) {
                uint32_t ah[4], al[4];
                ldmatrix_x4(ah, arow_h + ks * 16);
                ldmatrix_x4(al, arow_l + ks * 16);
                const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(brow + ks * 16);
                const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(brow + ks * 16 + 8);
                const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(brow + L + ks * 16);
                const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(brow + L + ks * 16 + 8);
                mma_f16(acc, ah, bh0, bh1);
                mma_f16(acc, al, bh0, bh1);
                mma_f16(acc, ah, bl0, bl1);
            }
            const int c = n0 + nt * 8 + (lane & 3) * 2;
            const int r0 = mt * 16 + (lane >> 2);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = r0 + 8 * h;
                if (r >= rows) continue;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (c + j < cout) {
                        const float v = acc[2 * h + j] + __ldg(bias + c + j);
                        out[(size_t)r * out_C + (c + j) * dst_cs] = relu ? fmaxf(v, 0.f) : v;
                    }
                }
            }
        }
        f_mbar_arrive(&empty[slot]);              // 512 arrivals free the stage for the producer
    }
}

// Tensor-core pointwise layer of the TAIL pass: few rows (the CTA's stacked ROIs x 1..16 pixels = one to three 16-row tiles) and
// long K, 73 % of the network's weight bytes.  No shared-memory weight stages: the fp16 hi|lo weights are stored in MMA-fragment
// order ([n_tile][k_step][lane] x 16 bytes, plan.py), every warp owns whole 8-channel output tiles and streams their fragments
// from L2 straight into registers with coalesced 512-byte loads, four K-steps in flight.  One warp sums a tile's whole K in a fixed
// order, so a ROI's result does not depend on how many ROIs are stacked.  (A first version chunked the weights along N through the
// two 44 KB ring stages -- 43 chunks of 3 tiles for conv5, each with two block barriers and a K-slice reduction -- and was
// slower than the fp32 path.)  MEASURED (tools/cls_steps.py, tools/cls_tail.py): this path and the fp32 path take the same time
// (stage-4 layers 16.8 k vs 18.4 k cycles, conv5 112 k vs 103 k): the tail is bound by its weight stream -- 148 SMs each pulling the
// same 3.8 MB from L2, ~30 KB/us per SM = 4.4 TB/s aggregate -- not by the arithmetic, so it stays OFF by default (LP_CLS_TAIL_MMA=1
// enables it; tail group <= 2, its activation staging needs the room of the third ROI).
__device__ __forceinline__ void pw_layer_mma_direct(const float* __restrict__ in, int in_C, float* __restrict__ out, int out_C, int dst_cs,
                                                    int rows, int cin, int cout, const float* __restrict__ bias, int relu,
                                                    __half* __restrict__ astage, const uint4* __restrict__ wfrag) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_p = (rows + 15) & ~15, cin_p = (cin + 15) & ~15, cout_p = (cout + 7) & ~7;
    const int LA = cin_p + 8;
    __half* Ah = astage;
    __half* Al = astage + rows_p * LA;
    for (int e = tid; e < rows_p * (cin_p >> 1); e += FUSED_THREADS) {
        const int r = e / (cin_p >> 1), c = (e - r * (cin_p >> 1)) * 2;
        float x0 = 0.f, x1 = 0.f;
        if (r < rows) {
            if (c < cin) x0 = in[r * in_C + c];
            if (c + 1 < cin) x1 = in[r * in_C + c + 1];
        }
        const __half2 hi = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(Ah + r * LA + c) = hi;
        *reinterpret_cast<__half2*>(Al + r * LA + c) = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    }
    CSYNC();
    const int m_tiles = rows_p >> 4, ksteps = cin_p >> 4, n_tiles = cout_p >> 3;
    constexpr int NW = FUSED_THREADS / 32, MT = 3, PF = 4;        // row tiles per weight pass, K-steps of weights in flight
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int mt0 = 0; mt0 < m_tiles; mt0 += MT) {
        for (int nt = warp; nt < n_tiles; nt += NW) {
            float acc[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
            const uint4* wp = wfrag + (size_t)nt * ksteps * 32 + lane;
            const __half* arow_h = Ah + (mt0 * 16 + (lane & 15)) * LA + (lane >> 4) * 8;
            uint4 cur[PF];
#pragma unroll
            for (int i = 0; i < PF; ++i) cur[i] = i < ksteps ? __ldg(wp + i * 32) : zero4;
            for (int ks0 = 0; ks0 < ksteps; ks0 += PF) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    if (ks0 + i < ksteps) {
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            if (mt0 + m < m_tiles) {
                                uint32_t ah[4], al[4];
                                ldmatrix_x4(ah, arow_h + m * 16 * LA + (ks0 + i) * 16);
                                ldmatrix_x4(al, arow_h + m * 16 * LA + rows_p * LA + (ks0 + i) * 16);
                                mma_f16(acc[m], ah, cur[i].x, cur[i].y);
                                mma_f16(acc[m], al, cur[i].x, cur[i].y);
                                mma_f16(acc[m], ah, cur[i].z, cur[i].w);
                            }
                        }
                    }
                    // the register set of this K-step is free again: fetch the step PF further on into it
                    if (ks0 + PF + i < ksteps) cur[i] = __ldg(wp + (ks0 + PF + i) * 32);
                }
            }
            const int c = nt * 8 + (lane & 3) * 2;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (mt0 + m >= m_tiles) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = (mt0 + m) * 16 + (lane >> 2) + 8 * h;
                    if (r >= rows) continue;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (c + j < cout) {
                            const float v = acc[m][2 * h + j] + __ldg(bias + c + j);
                            out[(size_t)r * out_C + (c + j) * dst_cs] = relu ? fmaxf(v, 0.f) : v;
                        }
                }
            }
        }
    }
}

// rows of W that fit one weight stage
__device__ __forceinline__ int chunk_rows(int cin, int cout, int slot_floats) {
    const int cout_p = (cout + 3) & ~3;
    const int r = slot_floats / cout_p;
    return r < cin ? r : cin;
}

// Pointwise conv as a shared-memory GEMM.  The activations have few rows (4..256), so the layer is bound by
// how often a weight is re-read from shared memory: a thread owns RT rows x 4 output channels (each weight
// float4 feeds 4*RT FMAs) and K is split over KS adjacent lanes (thread = tile * KS + ks handles the rows
// ci == ks mod KS of every chunk) so that all 512 threads have work even when rows * cout / (4 RT) is small;
// the KS partial sums are combined with warp shuffles.  W arrives chunk by chunk ([rows][cout_p]) in the
// two-stage ring: wait full -> FMAs from shared memory -> every thread arrives on empty.
template <int RT>
__device__ __forceinline__ void pw_layer(const float* __restrict__ in, int in_C, float* __restrict__ out, int out_C, int dst_cs,
                                         int rows, int cin, int cout, const float* __restrict__ bias, int relu, int ks_log2,
                                         const float* __restrict__ wbuf, int slot_floats, uint64_t* full, uint64_t* empty,
                                         uint32_t& chunk_ctr) {
    const int cout_p = (cout + 3) & ~3;
    const int ncg = cout_p >> 2, nrg = (rows + RT - 1) / RT;
    const int n_tiles = nrg * ncg;
    const int KS = 1 << ks_log2;
    const int ks = threadIdx.x & (KS - 1), tile = threadIdx.x >> ks_log2;
    const bool active = tile < n_tiles;
    const int tt = active ? tile : 0;
    const int rg = tt / ncg, c0 = (tt - rg * ncg) * 4, r0 = rg * RT;
    float acc[RT][4];
#pragma unroll
    for (int r = 0; r < RT; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
    int roff[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) roff[r] = min(r0 + r, rows - 1) * in_C;
    const int R = chunk_rows(cin, cout, slot_floats);
    for (int k0 = 0; k0 < cin; k0 += R, ++chunk_ctr) {
        const int nr = min(R, cin - k0);
        const uint32_t slot = chunk_ctr & 1;
        f_mbar_wait(&full[slot], (chunk_ctr >> 1) & 1);
        if (active) {
            const float* wp = wbuf + slot * slot_floats + c0;
            const float* ap = in + k0;
#pragma unroll 2
            for (int ci = ks; ci < nr; ci += KS) {
                const float4 w = *reinterpret_cast<const float4*>(wp + ci * cout_p);
                float a[RT];
#pragma unroll
                for (int r = 0; r < RT; ++r) a[r] = ap[roff[r] + ci];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    acc[r][0] = fmaf(a[r], w.x, acc[r][0]);
                    acc[r][1] = fmaf(a[r], w.y, acc[r][1]);
                    acc[r][2] = fmaf(a[r], w.z, acc[r][2]);
                    acc[r][3] = fmaf(a[r], w.w, acc[r][3]);
                }
            }
        }
        f_mbar_arrive(&empty[slot]);              // 512 arrivals free the stage for the producer
    }
    // combine the K slices (adjacent lanes)
    for (int off = 1; off < KS; off <<= 1) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] += __shfl_xor_sync(0xffffffffu, acc[r][j], off);
    }
    if (active && ks == 0) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0));
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (r0 + r >= rows) break;
            float* op = out + (size_t)(r0 + r) * out_C;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c0 + j < cout) {
                    const float v = acc[r][j] + bb[j];
                    op[(c0 + j) * dst_cs] = relu ? fmaxf(v, 0.f) : v;
                }
            }
        }
    }
}

// Depthwise 3x3 (pad 1, bias, no activation), task = (ROI, channel, output row): the three input rows are read once as
// sliding windows (3.75 shared-memory loads per output instead of 9) and the index arithmetic is per row, not per tap
// (the per-tap version spent ~100 instructions per output on it).  Same FMA order per output as a tap loop 0..8.
template <int WO, int S>
__device__ __forceinline__ void dw_rows(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ w,
                                        const float* __restrict__ b, int rois, int C, int H, int W, int src_C, int src_off,
                                        int dst_C, int dst_off, int dst_cs) {
    constexpr int NIN = (WO - 1) * S + 3;
    const int Ho = (H + 2 - 3) / S + 1;
    const int n_tasks = rois * Ho * C;
    for (int task = threadIdx.x; task < n_tasks; task += FUSED_THREADS) {
        const int c = task % C, t2 = task / C, oy = t2 % Ho, g = t2 / Ho;
        float wv[9];
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) wv[t9] = __ldg(w + t9 * C + c);
        const float bias = __ldg(b + c);
        float out[WO];
#pragma unroll
        for (int ox = 0; ox < WO; ++ox) out[ox] = bias;
        const float* ip = src + (size_t)g * H * W * src_C + src_off + c;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * S - 1 + ky;
            if (iy < 0 || iy >= H) continue;
            float row[NIN];
#pragma unroll
            for (int j = 0; j < NIN; ++j) {
                const int ix = j - 1;
                row[j] = (ix >= 0 && ix < W) ? ip[(iy * W + ix) * src_C] : 0.f;
            }
#pragma unroll
            for (int ox = 0; ox < WO; ++ox)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) out[ox] = fmaf(row[ox * S + kx], wv[ky * 3 + kx], out[ox]);
        }
        float* op = dst + ((size_t)g * Ho * WO + (size_t)oy * WO) * dst_C + dst_off + c * dst_cs;
#pragma unroll
        for (int ox = 0; ox < WO; ++ox) op[(size_t)ox * dst_C] = out[ox];
    }
}

__global__ void __launch_bounds__(FUSED_BLOCK, 1)
shufflenet_fused_kernel(const uint8_t* __restrict__ in, int n_rois_cap, const int* __restrict__ n_dev, const float* __restrict__ W,
                        const uint4* __restrict__ W16, int astage_off,
                        const FStep* __restrict__ steps, int n_front, int n_mid, int n_tail, int GT, float* park,
                        int tail_astage_off, int tail_red_off, int tail_off, int tail_floats, int in_hw,
                        float mean, float stdv, float* __restrict__ logits, int n_classes, int wbuf_off,
                        int back_off, int back_floats, long long* dbg, int cs) {
    extern __shared__ __align__(16) float sm[];
    __shared__ FStep s_steps[FUSED_MAX_STEPS];
    __shared__ float s_norm[256];
    __shared__ __align__(16) float s_w1[27 * 24 + 24];        // conv1 weights + bias (24-channel fast path)
    __shared__ uint64_t s_full[2], s_empty[2], s_cready[2];   // weights landed / consumed (local) / every CTA armed (leader)
    __shared__ uint64_t s_phase;                              // the compute threads finished a pass (front end of a ROI / back end)
    const int tid = threadIdx.x;
    // Weight stages.  Front end: two 24-KB stages behind the activations.  Back end: the front end's buffers are
    // dead, so the two stages grow over them ([back_off, back_off + 2*back_floats)): the stream is bound by the
    // bytes in flight per SM, and 2 x 61 KB in flight instead of 2 x 24 KB is what the larger stages buy.  The
    // regions overlap, so the producer starts a pass only after the compute threads finished the previous one.
    float* wbuf = sm + wbuf_off;                            // [2][WBUF_FLOATS]
    float* wbig = sm + back_off;                            // [2][back_floats]
    for (int i = tid; i < (n_front + n_mid + n_tail) * (int)(sizeof(FStep) / 4); i += FUSED_BLOCK)
        reinterpret_cast<int*>(s_steps)[i] = reinterpret_cast<const int*>(steps)[i];
    if (tid == 0) {
        f_mbar_init(&s_full[0], 1); f_mbar_init(&s_full[1], 1);
        f_mbar_init(&s_empty[0], FUSED_THREADS); f_mbar_init(&s_empty[1], FUSED_THREADS);
        f_mbar_init(&s_cready[0], cs); f_mbar_init(&s_cready[1], cs);
        f_mbar_init(&s_phase, FUSED_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ToTensor + Normalize exactly as torchvision computes them: (u8 / 255 - mean) / std, IEEE divisions
    if (tid < 256) s_norm[tid] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)tid, 255.f), mean), stdv);
    __syncthreads();
    f_cluster_sync();                                        // every CTA's barriers exist before any remote arrive / multicast
    const int img_bytes = in_hw * in_hw * 3;
    const int n_rois = n_dev ? min(n_rois_cap, *n_dev) : n_rois_cap;     // count produced on the device by roi_select
    // Schedule of a CTA: its ROIs are b, b + grid, b + 2 grid, ...; "rounds" that exist for CTA 0 exist for every CTA
    // (a CTA without a ROI in the last round computes on stale data and writes nothing: the pass structure, and with
    // it the weight stream, is uniform over the grid / cluster).  Per chunk of GT rounds: front + middle of each ROI
    // (the middle parks its result in global memory), then ONE tail pass over the stacked ROIs.
    const int n_rounds = (n_rois + (int)gridDim.x - 1) / (int)gridDim.x;
    float* my_park = park + (size_t)blockIdx.x * GT * s_steps[n_front + n_mid - 1].dst_roi_stride;
    if (tid >= FUSED_THREADS) {
        // ================= weight producer warp: replays the step list, one bulk copy per K chunk =========
        // Cluster protocol per chunk: every CTA's producer waits until its own consumers released the stage,
        // arms its local `full` barrier with the byte count and tells the leader (remote arrive on cready);
        // the leader waits for all CTAs and issues ONE multicast bulk copy that lands in every CTA's stage
        // and completes every CTA's `full` barrier.  L2 weight traffic per ROI drops by the cluster size.
        if (tid == FUSED_THREADS) {
            const uint32_t rank = f_cluster_rank();
            const uint16_t mask = (uint16_t)((1u << cs) - 1u);
            uint32_t ctr = 0, passes = 0;
            for (int rnd0 = 0; rnd0 < n_rounds; rnd0 += GT) {
                const int ng = min(GT, n_rounds - rnd0);
                for (int pass = 0; pass <= 2 * ng; ++pass, ++passes) {
                    // passes 2j, 2j+1: front and middle of stacked ROI j; pass 2*ng: tail
                    const int kind = pass == 2 * ng ? 2 : (pass & 1);
                    const bool back = kind == 1;
                    const int s_begin = kind == 0 ? 0 : (kind == 1 ? n_front : n_front + n_mid);
                    const int s_end = kind == 0 ? n_front : (kind == 1 ? n_front + n_mid : n_front + n_mid + n_tail);
                    if (passes) f_mbar_wait(&s_phase, (passes - 1) & 1);      // previous pass done: its buffers and stages are free
                    float* wdst = kind == 0 ? wbuf : (kind == 1 ? wbig : sm + tail_off);
                    const int slot_floats = kind == 0 ? WBUF_FLOATS : (kind == 1 ? back_floats : tail_floats);
                    for (int si = s_begin; si < s_end; ++si) {
                        const FStep& st = s_steps[si];
                        if (st.op != FS_PW) continue;
                        if (kind == 2 && st.w16_off > 0) continue;       // tail tensor-core layers read their weights straight from L2
                        if (kind != 0 && st.w16_off > 0) {
                            // tensor-core layer: chunks of output channels, rows of 2 planes x L halves (4*L bytes)
                            const int NN = chunk_couts(st.cin, st.cout, slot_floats), cout_p8 = (st.cout + 7) & ~7;
                            const int row_b = w16_row_halves(st.cin) * 4;
                            for (int n0 = 0; n0 < cout_p8; n0 += NN, ++ctr) {
                                const int nn = min(NN, cout_p8 - n0);
                                const uint32_t slot = ctr & 1;
                                if (ctr >= 2) f_mbar_wait(&s_empty[slot], ((ctr >> 1) - 1) & 1);
                                const uint32_t bytes = (uint32_t)nn * row_b;
                                f_mbar_expect_tx(&s_full[slot], bytes);
                                f_mbar_arrive_remote(&s_cready[slot], 0);
                                if (rank == 0) {
                                    f_mbar_wait_cluster(&s_cready[slot], (ctr >> 1) & 1);
                                    f_bulk_g2s_multicast(wdst + slot * slot_floats,
                                                         reinterpret_cast<const uint8_t*>(W16 + (st.w16_off - 1)) + (size_t)n0 * row_b,
                                                         bytes, &s_full[slot], mask);
                                }
                            }
                            continue;
                        }
                        const int cout_p = (st.cout + 3) & ~3;
                        const int R = chunk_rows(st.cin, st.cout, slot_floats);
                        for (int k0 = 0; k0 < st.cin; k0 += R, ++ctr) {
                            const int nr = min(R, st.cin - k0);
                            const uint32_t slot = ctr & 1;
                            if (ctr >= 2) f_mbar_wait(&s_empty[slot], ((ctr >> 1) - 1) & 1);
                            const uint32_t bytes = (uint32_t)nr * cout_p * 4;
                            f_mbar_expect_tx(&s_full[slot], bytes);
                            f_mbar_arrive_remote(&s_cready[slot], 0);
                            if (rank == 0) {
                                f_mbar_wait_cluster(&s_cready[slot], (ctr >> 1) & 1);
                                f_bulk_g2s_multicast(wdst + slot * slot_floats, W + st.w_off + (size_t)k0 * cout_p, bytes, &s_full[slot], mask);
                            }
                        }
                    }
                }
            }
        }
    } else {
    uint32_t chunk_ctr = 0;

    for (int rnd0 = 0; rnd0 < n_rounds; rnd0 += GT) {
        const int ng = min(GT, n_rounds - rnd0);             // ROIs stacked in this chunk's tail
        for (int pass = 0; pass <= 2 * ng; ++pass) {
            const int kind = pass == 2 * ng ? 2 : (pass & 1);   // 0 front, 1 middle, 2 tail
            const bool back = kind == 1;
            const int jroi = pass >> 1;                           // stacked index of the ROI of a front / middle pass
            const int roi = (int)blockIdx.x + (rnd0 + jroi) * (int)gridDim.x;
            const int s_begin = kind == 0 ? 0 : (kind == 1 ? n_front : n_front + n_mid);
            const int s_end = kind == 0 ? n_front : (kind == 1 ? n_front + n_mid : n_front + n_mid + n_tail);
            if (kind == 0 && roi < n_rois) {
                // stage the u8 crop (16-B copies); the CONV1 step's src is its float offset
                const uint4* g4 = reinterpret_cast<const uint4*>(in + (size_t)roi * img_bytes);
                uint4* s4 = reinterpret_cast<uint4*>(sm + s_steps[0].src);
                for (int i = tid; i < img_bytes / 16; i += FUSED_THREADS) s4[i] = __ldg(g4 + i);
                CSYNC();
            }
            const float* wst = kind == 0 ? wbuf : (kind == 1 ? wbig : sm + tail_off);
            const int wst_floats = kind == 0 ? WBUF_FLOATS : (kind == 1 ? back_floats : tail_floats);
            for (int si = s_begin; si < s_end; ++si) {
                const FStep& st = s_steps[si];
                const long long t_step = dbg ? clock64() : 0;
                const int rois = kind == 2 ? ng : 1;
                float* dst = sm + st.dst;
                const float* src = sm + st.src;
                const int Ho = (st.op == FS_PW || st.op == FS_COPY) ? st.H : (st.H + 2 - 3) / st.stride + 1;
                const int Wo = (st.op == FS_PW || st.op == FS_COPY) ? st.W : (st.W + 2 - 3) / st.stride + 1;
                switch (st.op) {
                case FS_CONV1: {
                    // 3x3 stride 2 pad 1 on the u8 RGB crop, 3 -> cout, ReLU.
                    const uint8_t* img = reinterpret_cast<const uint8_t*>(src);
                    const int ncg = (st.cout + 3) >> 2;
                    const float* w = W + st.w_off;           // [27][cout_p]
                    const int cout_p = ncg * 4;
                    if (cout_p == 24) {
                        // thread = one pixel x all 24 channels: each input byte is fetched and normalised once for 24 FMAs
                        // (the (pixel, 4 channels) mapping below spent 3 memory operations per 4 FMAs: 46 k cycles per ROI);
                        // weights are staged in shared memory and read as broadcast float4
                        for (int i = tid; i < 27 * 24 + 24; i += FUSED_THREADS) s_w1[i] = i < 27 * 24 ? __ldg(w + i) : __ldg(W + st.b_off + i - 27 * 24);
                        CSYNC();
                        for (int pix = tid; pix < Ho * Wo; pix += FUSED_THREADS) {
                            const int oy = pix / Wo, ox = pix - oy * Wo;
                            float acc[24];
#pragma unroll
                            for (int j = 0; j < 24; ++j) acc[j] = s_w1[27 * 24 + j];
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
                                const int iy = oy * 2 - 1 + ky;
                                if (iy < 0 || iy >= st.H) continue;
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    const int ix = ox * 2 - 1 + kx;
                                    if (ix < 0 || ix >= st.W) continue;
                                    const uint8_t* px = img + (iy * st.W + ix) * 3;
#pragma unroll
                                    for (int c = 0; c < 3; ++c) {
                                        const float x = s_norm[px[c]];
                                        const float4* wr = reinterpret_cast<const float4*>(s_w1 + ((ky * 3 + kx) * 3 + c) * 24);
#pragma unroll
                                        for (int cg = 0; cg < 6; ++cg) {
                                            const float4 wv = wr[cg];
                                            acc[4 * cg] = fmaf(x, wv.x, acc[4 * cg]); acc[4 * cg + 1] = fmaf(x, wv.y, acc[4 * cg + 1]);
                                            acc[4 * cg + 2] = fmaf(x, wv.z, acc[4 * cg + 2]); acc[4 * cg + 3] = fmaf(x, wv.w, acc[4 * cg + 3]);
                                        }
                                    }
                                }
                            }
                            float* o = dst + (size_t)pix * st.dst_C;
#pragma unroll
                            for (int j = 0; j < 24; ++j) if (j < st.cout) o[j] = fmaxf(acc[j], 0.f);
                        }
                        break;
                    }
                    for (int t = tid; t < Ho * Wo * ncg; t += FUSED_THREADS) {
                        const int cg = t % ncg, pix = t / ncg, oy = pix / Wo, ox = pix - oy * Wo;
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(W + st.b_off + cg * 4));
                        float a0 = b4.x, a1 = b4.y, a2 = b4.z, a3 = b4.w;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int iy = oy * 2 - 1 + ky;
                            if (iy < 0 || iy >= st.H) continue;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int ix = ox * 2 - 1 + kx;
                                if (ix < 0 || ix >= st.W) continue;
                                const uint8_t* px = img + (iy * st.W + ix) * 3;
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const float x = s_norm[px[c]];
                                    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + ((ky * 3 + kx) * 3 + c) * cout_p + cg * 4));
                                    a0 = fmaf(x, wv.x, a0); a1 = fmaf(x, wv.y, a1); a2 = fmaf(x, wv.z, a2); a3 = fmaf(x, wv.w, a3);
                                }
                            }
                        }
                        float* o = dst + (size_t)pix * st.dst_C + cg * 4;
                        const float v[4] = {a0, a1, a2, a3};
                        for (int j = 0; j < 4; ++j) if (cg * 4 + j < st.cout) o[j] = fmaxf(v[j], 0.f);
                    }
                    break;
                }
                case FS_MAXPOOL: {
                    if (Wo == 16 && st.W == 32) {
                        // task = (channel, output row): the three input rows are scanned once (33 loads for 16 outputs)
                        for (int task = tid; task < Ho * st.cout; task += FUSED_THREADS) {
                            const int c = task % st.cout, oy = task / st.cout;
                            float m[16];
#pragma unroll
                            for (int ox = 0; ox < 16; ++ox) m[ox] = -INFINITY;
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
                                const int iy = oy * 2 - 1 + ky;
                                if (iy < 0 || iy >= st.H) continue;
                                const float* rp = src + (size_t)iy * 32 * st.src_C + st.src_off + c;
                                float prev = -INFINITY;                       // column -1 is padding
#pragma unroll
                                for (int ox = 0; ox < 16; ++ox) {
                                    const float a = rp[(2 * ox) * st.src_C], b = rp[(2 * ox + 1) * st.src_C];
                                    m[ox] = fmaxf(m[ox], fmaxf(prev, fmaxf(a, b)));
                                    prev = b;
                                }
                            }
                            float* op = dst + (size_t)oy * 16 * st.dst_C + st.dst_off + c;
#pragma unroll
                            for (int ox = 0; ox < 16; ++ox) op[(size_t)ox * st.dst_C] = m[ox];
                        }
                        break;
                    }
                    for (int t = tid; t < Ho * Wo * st.cout; t += FUSED_THREADS) {
                        const int c = t % st.cout, pix = t / st.cout, oy = pix / Wo, ox = pix - oy * Wo;
                        float m = -INFINITY;
                        for (int ky = 0; ky < 3; ++ky) {
                            const int iy = oy * 2 - 1 + ky;
                            if (iy < 0 || iy >= st.H) continue;
                            for (int kx = 0; kx < 3; ++kx) {
                                const int ix = ox * 2 - 1 + kx;
                                if (ix < 0 || ix >= st.W) continue;
                                m = fmaxf(m, src[(size_t)(iy * st.W + ix) * st.src_C + st.src_off + c]);
                            }
                        }
                        dst[(size_t)pix * st.dst_C + st.dst_off + c] = m;
                    }
                    break;
                }
                case FS_PW: {
                    const int rows = rois * st.H * st.W;
                    const float* ip = src + st.src_off;
                    float* op = dst + st.dst_off;
                    if (back && st.w16_off > 0) {
                        pw_layer_mma(ip, st.src_C, op, st.dst_C, st.dst_cs, rows, st.cin, st.cout, W + st.b_off, st.relu,
                                     reinterpret_cast<__half*>(sm + astage_off), wbig, back_floats, s_full, s_empty, chunk_ctr);
                        break;
                    }
                    if (kind == 2 && st.w16_off > 0) {
                        pw_layer_mma_direct(ip, st.src_C, op, st.dst_C, st.dst_cs, rows, st.cin, st.cout, W + st.b_off, st.relu,
                                            reinterpret_cast<__half*>(sm + tail_astage_off), W16 + (st.w16_off - 1));
                        break;
                    }
                    // RT = min(8, rows) rows per thread; K split over the largest power of two <= 8 that keeps
                    // tiles * KS within the 512 compute threads (plan.py guarantees tiles <= 512)
                    const int ncg = (st.cout + 3) >> 2;
                    const float* bias = W + st.b_off;
                    // the tiling (and with it the K split = the summation order of every output) is derived from the
                    // CAPACITY of the pass, not from how many ROIs are stacked this time: a ROI's logits do not depend
                    // on what else is in the batch
                    const int rows_cap = (kind == 2 ? GT : 1) * st.H * st.W;
                    // 12 rows (three stacked ROIs of 2x2 pixels): three groups of 4, not 8 + 4 padded to 8
                    // (only while the tiles still fit the 512 threads: conv5 with its 256 channel groups keeps 8-row groups)
                    const int rt = (rows_cap == 12 && 3 * ncg <= FUSED_THREADS) ? 4 : (rows_cap >= 8 ? 8 : (rows_cap >= 4 ? 4 : (rows_cap >= 2 ? 2 : 1)));
                    const int tiles = ((rows_cap + rt - 1) / rt) * ncg;
                    int ksl = 0;
                    while (ksl < 3 && (tiles << (ksl + 1)) <= FUSED_THREADS) ++ksl;
#define PW_CALL(RT_) pw_layer<RT_>(ip, st.src_C, op, st.dst_C, st.dst_cs, rows, st.cin, st.cout, bias, st.relu, ksl, wst, wst_floats, s_full, s_empty, chunk_ctr)
                    if (rt == 8) PW_CALL(8);
                    else if (rt == 4) PW_CALL(4);
                    else if (rt == 2) PW_CALL(2);
                    else PW_CALL(1);
#undef PW_CALL
                    break;
                }
                case FS_DW: {
                    // depthwise 3x3, pad 1, bias, no activation.  weights [9][C].  (row, channel) advance
                    // incrementally -- an integer division per element would dominate this tiny layer.
                    {
                        const float* w9 = W + st.w_off;
                        const float* b9 = W + st.b_off;
                        bool done = true;
#define DW_CALL(WO_, S_) dw_rows<WO_, S_>(src, dst, w9, b9, rois, st.cout, st.H, st.W, st.src_C, st.src_off, st.dst_C, st.dst_off, st.dst_cs)
                        if (st.H == st.W && Wo == 8 && st.stride == 2) DW_CALL(8, 2);
                        else if (st.H == st.W && Wo == 8 && st.stride == 1) DW_CALL(8, 1);
                        else if (st.H == st.W && Wo == 4 && st.stride == 2) DW_CALL(4, 2);
                        else if (st.H == st.W && Wo == 4 && st.stride == 1) DW_CALL(4, 1);
                        else if (st.H == st.W && Wo == 2 && st.stride == 2) DW_CALL(2, 2);
                        else if (st.H == st.W && Wo == 2 && st.stride == 1) DW_CALL(2, 1);
                        else done = false;
#undef DW_CALL
                        if (done) break;
                    }
                    const int C = st.cout, hw_in = st.H * st.W, hw_out = Ho * Wo;
                    const int wo_shift = __ffs(Wo) - 1, hw_shift = __ffs(hw_out) - 1;   // Ho, Wo are powers of two
                    const float* w = W + st.w_off;
                    const float* b = W + st.b_off;
                    // thread = one channel, a strided set of rows: the nine weights and the bias are fetched from
                    // global memory ONCE per thread (they were re-read for every output: ~600 cycles of L2 latency
                    // per row iteration, 4000 cycles for a layer with 30 k multiply-adds)
                    const int lanes_c = FUSED_THREADS / C;           // thread groups of C channels
                    const int c = tid % C, grp = tid / C;
                    const int rows = rois * hw_out;
                    if (grp < lanes_c) {
                        float wv[9];
#pragma unroll
                        for (int t9 = 0; t9 < 9; ++t9) wv[t9] = __ldg(w + t9 * C + c);
                        const float bias = __ldg(b + c);
                        for (int row = grp; row < rows; row += lanes_c) {
                            const int g = row >> hw_shift, r = row & (hw_out - 1);
                            const int oy = r >> wo_shift, ox = r & (Wo - 1);
                            const float* ip = src + (size_t)g * hw_in * st.src_C + st.src_off + c;
                            float xv[9];
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) {         // all loads first (clamped address, zeroed value): no branches
                                const int iy = oy * st.stride - 1 + t9 / 3, ix = ox * st.stride - 1 + t9 % 3;
                                const bool ok = ((unsigned)iy < (unsigned)st.H) && ((unsigned)ix < (unsigned)st.W);
                                const int cy = min(max(iy, 0), st.H - 1), cx = min(max(ix, 0), st.W - 1);
                                const float v = ip[(cy * st.W + cx) * st.src_C];
                                xv[t9] = ok ? v : 0.f;
                            }
                            float acc = bias;
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) acc = fmaf(xv[t9], wv[t9], acc);
                            dst[(size_t)row * st.dst_C + st.dst_off + c * st.dst_cs] = acc;
                        }
                    }
                    break;
                }
                case FS_COPY: {
                    const int rows = rois * st.H * st.W, C = st.cout;
                    const int step_r = FUSED_THREADS / C, step_c = FUSED_THREADS - step_r * C;
                    int c = tid % C, row = tid / C;
                    while (row < rows) {
                        dst[(size_t)row * st.dst_C + st.dst_off + c * st.dst_cs] = src[(size_t)row * st.src_C + st.src_off + c];
                        row += step_r; c += step_c;
                        if (c >= C) { c -= C; ++row; }
                    }
                    break;
                }
                case FS_STORE: {          // park the middle's result of stacked ROI `jroi` (read back by the tail pass)
                    const int n4 = (st.H * st.W * st.src_C) >> 2;
                    const float4* s4 = reinterpret_cast<const float4*>(src);
                    float4* g4 = reinterpret_cast<float4*>(my_park + (size_t)jroi * st.dst_roi_stride);
                    for (int i = tid; i < n4; i += FUSED_THREADS) __stcg(g4 + i, s4[i]);
                    break;
                }
                case FS_LOAD: {           // stacked rows: ROI g at dst + g * stride
                    const int n4 = (st.H * st.W * st.dst_C) >> 2;
                    for (int g = 0; g < rois; ++g) {
                        const float4* g4 = reinterpret_cast<const float4*>(my_park + (size_t)g * st.dst_roi_stride);
                        float4* d4 = reinterpret_cast<float4*>(dst + (size_t)g * st.dst_roi_stride);
                        for (int i = tid; i < n4; i += FUSED_THREADS) d4[i] = __ldcg(g4 + i);   // written by this CTA earlier: never the read-only path
                    }
                    break;
                }
                case FS_MEANFC: {
                    // x.mean([2,3]) then fc.  dst = scratch: [rois][cin] means, then [8][rois][64] partial sums.
                    // K is split over 8 thread groups (64 lanes = classes each) so the weight reads are
                    // coalesced and 8 rows of W are in flight per thread.
                    const int hw = st.H * st.W;
                    for (int t = tid; t < rois * st.cin; t += FUSED_THREADS) {
                        const int c = t % st.cin, g = t / st.cin;
                        float s = 0.f;
                        for (int q = 0; q < hw; ++q) s += src[(size_t)(g * hw + q) * st.src_C + c];
                        dst[t] = s / (float)hw;
                    }
                    CSYNC();
                    const float* w = W + st.w_off;           // [cin][cout]
                    float* part = dst + rois * st.cin;
                    for (int j0 = 0; j0 < st.cout; j0 += 64) {
                        const int j = j0 + (tid & 63), ks = tid >> 6;      // 8 K-slices
                        const int kper = (st.cin + 7) / 8, k0 = ks * kper, k1 = min(st.cin, k0 + kper);
                        for (int g = 0; g < rois; ++g) {
                            float acc = 0.f;
                            if (j < st.cout) {
                                const float* m = dst + (size_t)g * st.cin;
                                int c = k0;
                                for (; c + 8 <= k1; c += 8) {
                                    float wv[8];
#pragma unroll
                                    for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (size_t)(c + u) * st.cout + j);
#pragma unroll
                                    for (int u = 0; u < 8; ++u) acc = fmaf(m[c + u], wv[u], acc);
                                }
                                for (; c < k1; ++c) acc = fmaf(m[c], __ldg(w + (size_t)c * st.cout + j), acc);
                            }
                            part[(ks * rois + g) * 64 + (tid & 63)] = acc;
                        }
                        CSYNC();
                        for (int t = tid; t < ng * 64; t += FUSED_THREADS) {
                            const int jj = t & 63, g = t >> 6;
                            const int rid = (int)blockIdx.x + (rnd0 + g) * (int)gridDim.x;
                            if (j0 + jj < st.cout && rid < n_rois) {
                                float acc = __ldg(W + st.b_off + j0 + jj);
                                for (int k = 0; k < 8; ++k) acc += part[(k * rois + g) * 64 + jj];
                                logits[(size_t)rid * n_classes + j0 + jj] = acc;
                            }
                        }
                        CSYNC();
                    }
                    break;
                }
                }
                CSYNC();
                if (dbg && blockIdx.x == 0 && tid == 0) dbg[si] += clock64() - t_step;
            }
            f_mbar_arrive(&s_phase);          // pass finished: the producer may start the next pass's weight stream
        }
    }
    }   // compute threads
    f_cluster_sync();        // no CTA leaves while a peer may still multicast into it or arrive on its barriers
}

struct lp_fused_cls {
    const FStep* steps_dev = nullptr;
    const float* weights = nullptr;
    const uint4* weights16 = nullptr;   // split-f16 pointwise weights of the back end (may be null: SIMT pointwise layers)
    int astage_off = 0;                 // float offset of the fp16 activation staging of the tensor-core pointwise layers
    int n_front = 0, n_mid = 0, n_tail = 0, GT = 1, in_hw = 0, n_classes = 0;
    float* park = nullptr;              // [sm_count][GT][park_floats]: the middle's results waiting for the tail pass
    int tail_off = 0, tail_floats = 0;  // tail weight stages (floats): start and size of one
    int tail_astage_off = 0, tail_red_off = 0;   // tail: fp16 activation staging and K-slice partial sums of the tensor-core layers
    size_t smem_bytes = 0;
    int wbuf_off = 0;
    int back_off = 0, back_floats = 0;   // back-end weight stages (floats): start and size of one
    int cluster = 1;            // CTAs sharing one multicast weight stream (env LP_CLS_CLUSTER; measured: 2 = no gain, the
                                // stream is bound by bytes in flight per SM, not by L2; 4+ halves the resident CTAs)
    float mean = 0.f, stdv = 1.f;
    bool loaded = false;
};
void lp_fused_free(lp_fused_cls* f) { delete f; }     // owns no device memory: weights, steps and the park buffer are the caller's

extern "C" int lp_fused_classifier_load(lp_ctx* ctx, const void* steps_dev, int n_front, int n_mid, int n_tail,
                                        const float* weights, const void* weights16, int tail_group, int in_hw, int n_classes,
                                        size_t smem_bytes, size_t back_bytes, size_t astage_bytes, size_t tail_bytes,
                                        size_t tail_astage_bytes, int park_floats, float* park, size_t park_bytes,
                                        float mean, float stdv) {
    LP_CHECK(ctx && steps_dev && weights && park, "lp_fused_classifier_load: null argument");
    lp_device_guard dev_guard(ctx);
    LP_CHECK(n_front + n_mid + n_tail <= FUSED_MAX_STEPS && n_front > 0 && n_mid > 0 && n_tail > 0,
             "lp_fused_classifier_load: bad step counts");
    LP_CHECK(tail_group >= 1 && tail_group <= 8 && park_floats > 0 && park_floats % 4 == 0, "lp_fused_classifier_load: bad tail group");
    LP_CHECK(smem_bytes <= 227 * 1024, "lp_fused_classifier_load: %zu B shared memory exceeds 227 KB", smem_bytes);
    LP_CHECK(park_bytes >= (size_t)ctx->sm_count * tail_group * park_floats * sizeof(float),
             "lp_fused_classifier_load: park buffer of %zu B < %d SMs x %d ROIs x %d floats", park_bytes, ctx->sm_count, tail_group, park_floats);
    if (!ctx->fused) ctx->fused = new lp_fused_cls();
    lp_fused_cls& f = *ctx->fused;
    f.loaded = false;
    f.steps_dev = (const FStep*)steps_dev; f.weights = weights; f.weights16 = (const uint4*)weights16;
    f.n_front = n_front; f.n_mid = n_mid; f.n_tail = n_tail; f.GT = tail_group;
    f.in_hw = in_hw; f.n_classes = n_classes; f.mean = mean; f.stdv = stdv;
    f.park = park;
    // the host-built map covers the activations; the two weight stages are appended here
    f.wbuf_off = (int)((smem_bytes + 15) / 16 * 4);
    f.smem_bytes = (size_t)f.wbuf_off * 4 + 2 * WBUF_FLOATS * 4;
    // back-end stages: everything between the end of the back end's activations and the end of the allocation,
    // which is grown to 224 KB (one CTA per SM anyway)
    LP_CHECK(back_bytes > 0 && back_bytes <= smem_bytes, "lp_fused_classifier_load: bad back-end extent");
    {
        cudaFuncAttributes fa{};
        LP_CUDA(cudaFuncGetAttributes(&fa, shufflenet_fused_kernel));
        const size_t dyn_max = ((size_t)227 * 1024 - fa.sharedSizeBytes) & ~(size_t)1023;     // static tables share the 227 KB
        LP_CHECK(f.smem_bytes <= dyn_max, "lp_fused_classifier_load: %zu B dynamic shared memory do not fit", f.smem_bytes);
        f.smem_bytes = dyn_max;
    }
    // [back end activations | fp16 activation staging of the tensor-core pointwise layers | stage 0 | stage 1]
    f.astage_off = (int)((back_bytes + 15) / 16 * 4);
    f.back_off = f.astage_off + (int)((astage_bytes + 15) / 16 * 4);
    LP_CHECK((size_t)f.back_off * 4 + 2 * WBUF_FLOATS * 4 <= f.smem_bytes, "lp_fused_classifier_load: no room for the back-end weight stages");
    f.back_floats = (int)(((f.smem_bytes / 4 - f.back_off) / 2) & ~(size_t)3);
    // tail: [stacked activations | stage 0 | stage 1]
    // tail: [stacked activations | fp16 staging | K-slice partial sums (16 warps x 128 floats) | stage 0 | stage 1]
    f.tail_astage_off = (int)((tail_bytes + 15) / 16 * 4);
    f.tail_red_off = f.tail_astage_off + (int)((tail_astage_bytes + 15) / 16 * 4);
    f.tail_off = f.tail_red_off + (tail_astage_bytes ? 16 * 128 : 0);
    LP_CHECK((size_t)f.tail_off * 4 + 2 * WBUF_FLOATS * 4 <= f.smem_bytes, "lp_fused_classifier_load: tail group %d does not fit", tail_group);
    f.tail_floats = (int)(((f.smem_bytes / 4 - f.tail_off) / 2) & ~(size_t)3);
    smem_bytes = f.smem_bytes;
    LP_CUDA(cudaFuncSetAttribute(shufflenet_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    { const char* e = getenv("LP_CLS_CLUSTER"); f.cluster = e ? atoi(e) : 1; if (f.cluster < 1 || f.cluster > 8) f.cluster = 1; }
    f.loaded = true;
    return 0;
}

// returns 1 if it ran, 0 if no fused classifier is loaded
int lp_fused_classify(lp_ctx* ctx, const uint8_t* in, int n, float* logits, cudaStream_t st) {
    if (!ctx->fused || !ctx->fused->loaded || !ctx->use_fused) return 0;
    const lp_fused_cls& f = *ctx->fused;
    const int groups = n;
    const int cs = f.cluster;
    int grid = groups < ctx->sm_count ? groups : ctx->sm_count;
    grid = (grid + cs - 1) / cs * cs;
    if (grid > ctx->sm_count) grid = ctx->sm_count / cs * cs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FUSED_BLOCK); cfg.dynamicSmemBytes = f.smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, shufflenet_fused_kernel, in, n, ctx->roi_count_dev, f.weights, f.weights16, f.astage_off, f.steps_dev, f.n_front, f.n_mid, f.n_tail, f.GT,
                                        f.park, f.tail_astage_off, f.tail_red_off, f.tail_off, f.tail_floats, f.in_hw, f.mean, f.stdv, logits, f.n_classes, f.wbuf_off, f.back_off, f.back_floats, ctx->tc_dbg, cs);
    if (le != cudaSuccess) { lp_set_error("shufflenet_fused launch failed: %s", cudaGetErrorString(le)); return -2; }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { lp_set_error("shufflenet_fused launch failed: %s", cudaGetErrorString(e)); return -2; }
    return 1;
}
