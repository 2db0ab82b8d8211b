// K2: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// GEMM view per tile: D[128 pixels x Cout] += A[128 x K] * B[K x Cout], K = taps * Cin, fp32
// accumulator in TMEM.  Operands are split-f16 (value = hi + lo): the product keeps the three terms
// Ahi*Bhi + Alo*Bhi + Ahi*Blo (~22 mantissa bits) -- single-pass fp16/bf16/tf32 miss the reference's
// 1e-2 px box tolerance.  Measured on B200 (tools/ubench/mma_rate.cu) an M=128,
// K=16 tcgen05.mma costs max(46, N/2) cycles: below N=128 it is bound by the shared-memory read of A,
// not by N.  So each K-step issues TWO instructions instead of three: Ahi x [Bhi|Blo] (N = 2*Cout, two
// accumulator halves) and Alo x Bhi (N = Cout); the epilogue adds the halves.
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
// Persistent, warp-specialised CTA (640 threads = 20 warps = 5 per SM sub-partition, which keeps 96 registers per
// thread; one CTA per SM, static round-robin over tiles):
//   warps 0-3   store warps: staging tile -> global, coalesced               <- stage_full / -> stage_empty
//               (two staging buffers; layers with room for only one copy it out with the epilogue warps)
//   warps 4-8   patch loaders.  Default: ONE thread issues TMA tensor copies (cp.async.bulk.tensor tile mode, zero fill by the copy
//               engine; 16-byte chunk rows, or whole 128-byte pixel rows into SWIZZLE_128B tiles on the 1x1 layers with cin % 64 == 0;
//               element strides 2 x 2 pick one parity phase per copy on the stride-2 layers) and patch_full counts bytes; the other
//               loader threads idle.  LP_TC_TMA=0: all five warps issue cp.async 16 B with zero-fill  <- patch_empty / -> patch_full
//   warp  9     weight producer (one lane, bulk TMA)                         <- w_empty   / -> w_full
//   warps 10-17 epilogue: TMEM -> bias/act/residual/split -> staging tile    <- acc_full, stage_empty / -> acc_empty, stage_full
//               (TMEM quadrant = warp % 4, column half = (warp - 10) / 4)
//   warp  19    TMEM allocator + MMA issuer A (one lane): Ahi x [Bhi|Blo]     <- patch_full, w_full, acc_empty
//   warp  18    MMA issuer B: Alo x Bhi into its own accumulator columns (split_mma: cout <= 32, and cout <= 85 on long layers)
//   (ids ascend with how critical the role's instruction stream is: the sub-partition arbiter prefers the highest id)
// Up to 8 patch stages and 4 TMEM accumulator stages keep several tiles in flight: the small-channel
// layers are HBM/latency-bound, so tiles i+1.. load and tile i-1 drains while tile i is in the tensor core.
// Consecutive launches are chained with programmatic dependent launch (griddepcontrol).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int TC_THREADS = 640;             // 20 warps = 5 per SM sub-partition: 96 registers per thread (22 warps would cap them at 80)
// Warp roles, ordered by how critical their instruction stream is: the SM sub-partition's arbiter prefers the HIGHEST
// warp id among eligible warps (B300_MICROARCH.md "hi-wid-first"), so the single MMA-issuing warp gets the top id, the
// epilogue warps (the bound of the 1x1 layers) come next, and the latency-tolerant copy roles get the low ids.  The
// epilogue warp's TMEM lane quadrant is warp % 4 whatever its id.
constexpr int STORE_WARPS = 4, W_STORE0 = 0, STORE_THREADS = STORE_WARPS * 32;             // staged tile -> global
constexpr int LOADER_WARPS = 5, W_LOADER0 = 4;
constexpr int W_PRODUCER = 9;
constexpr int EPI_WARPS = 8, W_EPI0 = 10, W_MMA2 = 18, W_MMA = 19;
constexpr int LOADER_THREADS = LOADER_WARPS * 32;
constexpr int MAX_PST = 8, MAX_AST = 4;    // patch / accumulator stages
constexpr int TILE_M = 128;
constexpr int TCT_H = 16, TCT_W = 8;       // spatial output tile (rows x cols)
constexpr int MAX_WST = 16;                // weight stages / resident K-blocks

struct TcParams {
    const __half* in;  long long in_plane, in_img;  int in_C, in_coff;
    void* out;         long long out_plane, out_img; int out_C, out_coff, out_fmt;
    const __half* res; long long res_plane, res_img; int res_C, res_coff;
    const uint8_t* wtc;        // packed split weights of this op
    const float* bias;
    int cin, cout, act, res_first;
    int cout_real, out_cstride, seg_l0, seg_len, seg_pad;   // segmented (channel-shuffle) destination; seg_len == 0: plain
    int H, W, Ho, Wo;          // input / output spatial size
    int n_img, tiles_x, tiles_y, n_tiles;
    int ksize, stride;
    int slots, slots_p;        // patch pixels, padded count (LBO_A = slots_p * 16 B)
    int pitch;                 // patch row pitch in pixels (SBO_A = pitch * 16 B); 1x1: 8
    int phase_slots;           // stride 2: slots per parity phase
    int kb_ch, n_cb, n_kb;     // channels per K unit, channel blocks per tap, weight blocks per tile
    unsigned magic_per_img, magic_tiles_x;   // tile -> (image, tile row, tile column) by multiply-high
    float inv_hw_out;
    int tap_off[9];            // A start offset of each tap inside a patch, in 16-B slots (constant-bank operands of the issue loop)
    int n_units, upb, unit_bytes;   // K units (tap x channel block) per tile, units per weight block, bytes per unit
    int w_stages, stage_bytes, resident;
    int patch_stages, acc_stages;
    int ld_depth;              // patch tiles a loader thread keeps in flight before it waits for (and publishes) the oldest
    int n_mma, mtab_bytes;     // MMA issue table: one uint2 per tcgen05.mma of a tile
    int fills_per_tile;        // streaming weights: n_kb / w_stages (stage pattern repeats every tile)
    int epi_pitch, epi_bytes;  // epilogue staging: bytes per pixel row (+16 pad) and total (0 = direct stores)
    int epi_bufs, epi_buf_bytes;   // staging buffers (2 = conversion and copy-out overlap, 1 = they alternate) and bytes of one
    int tab_bytes;             // 3x3: per-slot geometry table (py | px<<8) in shared memory
    int tmem_cols, acc_stride;  // TMEM columns allocated; column stride between the two accumulator stages
    int split_mma;             // 1: the two MMAs of a K-step are issued by TWO warps (A: Ahi x [Bhi|Blo] -> columns [0, 2*cout),
                               // B: Alo x Bhi -> columns [2*cout, 3*cout)); 0: one warp issues both, Alo x Bhi accumulates into [0, cout)
    unsigned magic_chunks, magic_pitch;   // ceil(2^32 / n) for division by n_chunks / pitch
    long long total_pix;       // 1x1: n_img*H*W
    int a_sw128;               // 1x1 layers with cin % 64 == 0 (TMA only): A tiles are [64-channel block][128 pixels][128 B] in the
                               // SWIZZLE_128B K-major layout -- the tensor copy moves whole 128-byte pixel rows instead of 16-byte chunks
    int use_tma;               // 1: patches arrive by TMA tensor copies (one elected loader thread), 0: cp.async by the loader warps
    int plane_bytes;           // bytes of one plane (hi or lo) of a patch stage
    int tma_box_bytes;         // bytes one tensor copy delivers (one 8-channel chunk of one plane of a patch)
    int max_batch;             // the lo plane of image i is image max_batch + i of the tensor map
    alignas(64) CUtensorMap tmap;   // 3x3: {C, W, H, 2*max_batch} box {8, 10, 18, 1};  1x1: {C, 2*max_batch*H*W} box {8, 128}
    int dbg_flags;             // debugging (env LP_TC_DEBUG): 1 = loaders skip copies, 2 = epilogue skips math/stores, 4 = weights loaded once
    long long* dbg;            // optional: per-role cycle counters of CTA 0 (tools/op_times.py --tc-timing)
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;       // src-size 0 -> 16 bytes of zeros (halo / padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
    // SmemDescriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64)
    // (0 = no swizzle, 2 = SWIZZLE_128B)
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// descriptors passed as (lo, hi) 32-bit halves: only the low word (start address) changes between MMAs
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, executed by the whole warp: only the lane with issue != 0 issues the instruction (no divergent
// branch around it, so the surrounding loop stays in the uniform datapath).
__device__ __forceinline__ void umma_f16_pred(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate, uint32_t issue) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
        "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
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
    if (act == LP_ACT_SILU) return lp_silu(v);
    if (act == LP_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// role timing exists only in the DBG instantiation: on the single-warp MMA issue path every instruction
// costs ~9 cycles, so the production kernel carries no clock reads, flag tests or 64-bit counters
#define TCLK() (DBG ? clock64() : 0ll)
#define DFLAG(bit) (DBG && (p.dbg_flags & (bit)))
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

// MMA issuer role.  which == 3: one warp issues both MMAs of a K-step (Ahi x [Bhi|Blo], then Alo x Bhi accumulating into
// the first cout columns).  which == 1: ONE MMA per K-step, used by the two issuing warps of the SPLIT kernel with the SAME
// instruction stream: `sel` (0 = warp A: Ahi x [Bhi|Blo] into columns [0, 2*cout); 1 = warp B: Alo x Bhi into columns
// [2*cout, 3*cout)) only selects operand values.  `sel` must be a value ptxas knows to be warp-uniform (it comes from a
// redux.sync): with a second copy of the loops, or anything derived from threadIdx in them, ptxas leaves the uniform
// datapath (R2UR + vector adds per tcgen05.mma: measured 2x slower).
template <int which, bool DBG, bool SW>
__device__ __forceinline__ void mma_role(const TcParams& p, const uint32_t tmem_base, const bool pure, uint8_t* const patch0, uint8_t* const wst,
                                         const uint32_t plane_bytes, const uint32_t patch_bytes, uint64_t* const w_full, uint64_t* const w_empty,
                                         uint64_t* const patch_full, uint64_t* const patch_empty, uint64_t* const acc_full,
                                         uint64_t* const acc_empty, const uint32_t sel) {
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * p.cout) >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(p.cout >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    const uint32_t lbo_a = SW ? 16u : (uint32_t)p.slots_p * 16, sbo_a = SW ? 1024u : (uint32_t)p.pitch * 16;
    const uint32_t lbo_b = (uint32_t)p.cout * 32, sbo_b = 128;            // chunk stride = [hi|lo] rows
    const uint64_t da_base = umma_desc(0, lbo_a, sbo_a, SW ? 2u : 0u), db_base = umma_desc(0, lbo_b, sbo_b);
    const uint32_t da_hi = (uint32_t)(da_base >> 32), da_lo0 = (uint32_t)da_base;
    const uint32_t db_hi = (uint32_t)(db_base >> 32);
    const uint32_t b016 = (uint32_t)db_base + (smem_u32(wst) >> 4);
    const uint32_t plane16 = plane_bytes >> 4;
    // A per K-step (16 channels): two chunk planes further; swizzled rows: 32 bytes further inside the 128-byte row, and
    // a_unit16 more at the end of each 64-channel block (the next block is 16 KB further)
    const uint32_t a_step16 = SW ? 2u : (2 * lbo_a) >> 4, b_step16 = (2 * lbo_b) >> 4;
    constexpr uint32_t a_unit16 = SW ? (16384u >> 4) - 8u : 0u;
    const uint32_t patch016 = da_lo0 + (smem_u32(patch0) >> 4), patch_stride16 = patch_bytes >> 4;
    const int ksteps = p.kb_ch >> 4, ktot = p.cin >> 4, taps = p.ksize * p.ksize;
    const uint32_t leader = elect_one() ? 1u : 0u;

    const uint32_t d_sel = sel ? (uint32_t)(2 * p.cout) : 0u;      // warp B: its own accumulator columns
    const uint32_t a_sel16 = sel ? plane16 : 0u;                   // warp B: the lo plane of the patch
    const uint32_t idesc_sel = sel ? idesc1 : idesc2;              // warp B: N = cout (reads the Bhi rows only)
    int it = 0;
    // ring positions advance incrementally: a division by a runtime stage count costs ~20 dependent
    // instructions on this warp, more than the MMAs of a whole K unit
    uint32_t ps = 0, ps_phase = 0, as = 0, as_phase = 0, st = 0, st_phase = 0;
    long long m_wait_acc = 0, m_wait_patch = 0, m_wait_w = 0, m_total0 = TCLK();
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        long long t0 = TCLK();
        if (!pure && it >= p.acc_stages) mbar_wait(&acc_empty[as], as_phase ^ 1);
        long long t1 = TCLK();
        if (!pure) mbar_wait(&patch_full[ps], ps_phase);
        long long t2 = TCLK();
        m_wait_acc += t1 - t0; m_wait_patch += t2 - t1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t patch16 = patch016 + (uint32_t)ps * patch_stride16;
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride) + d_sel;
        uint32_t acc = 0;
        // Taps are unrolled so that each tap offset is a constant-bank operand; per tap the warp spends one
        // add, per K-step two MMAs and three adds.
        if (p.resident) {
            if (it == 0 && !pure) {                 // the whole layer lands once
                for (int kb = 0; kb < p.n_kb; ++kb) mbar_wait(&w_full[kb], 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            uint32_t b16 = b016;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                if (tap < taps) {
                    uint32_t a16 = patch16 + (uint32_t)p.tap_off[tap];
                    if constexpr (SW) {                    // swizzled 1x1 tiles (never split): 64-channel blocks of four K-steps, 16 KB apart
                        for (int cb = 0; cb < p.n_cb; ++cb) {
                            for (int ks = 0; ks < 4; ++ks) {
                                umma_f16_pred(d_tmem, a16, da_hi, b16, db_hi, idesc2, acc, leader);
                                umma_f16_pred(d_tmem, a16 + plane16, da_hi, b16, db_hi, idesc1, 1, leader);
                                acc = 1;
                                a16 += a_step16;
                                b16 += b_step16;
                            }
                            a16 += a_unit16;
                        }
                    } else if constexpr (which == 3) {
                        for (int kk = 0; kk < ktot; ++kk) {
                            umma_f16_pred(d_tmem, a16, da_hi, b16, db_hi, idesc2, acc, leader);           // Ahi x [Bhi|Blo]
                            umma_f16_pred(d_tmem, a16 + plane16, da_hi, b16, db_hi, idesc1, 1, leader);   // Alo x Bhi
                            acc = 1;
                            a16 += a_step16;
                            b16 += b_step16;
                        }
                    } else {
                        a16 += a_sel16;
                        for (int kk = 0; kk < ktot; ++kk) {
                            umma_f16_pred(d_tmem, a16, da_hi, b16, db_hi, idesc_sel, acc, leader);
                            acc = 1;
                            a16 += a_step16;
                            b16 += b_step16;
                        }
                    }
                }
            }
        } else {
            // weight blocks of `upb` K units (tap x channel block) stream through the ring: one barrier
            // wait and one commit per BLOCK, B advances linearly inside a block
            const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
            uint32_t b16 = 0;
            int u_in_blk = 0, units_left = p.n_units;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                if (tap < taps) {
                    uint32_t a16 = patch16 + (uint32_t)p.tap_off[tap];
                    for (int cb = 0; cb < p.n_cb; ++cb) {
                        if (u_in_blk == 0) {
                            long long t3 = TCLK();
                            if (!pure) mbar_wait(&w_full[st], st_phase);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            m_wait_w += TCLK() - t3;
                            b16 = b016 + st * stage16;
                        }
                        if constexpr (which == 3) {
                            for (int ks = 0; ks < ksteps; ++ks) {
                                umma_f16_pred(d_tmem, a16, da_hi, b16, db_hi, idesc2, acc, leader);
                                umma_f16_pred(d_tmem, a16 + plane16, da_hi, b16, db_hi, idesc1, 1, leader);
                                acc = 1;
                                a16 += a_step16;
                                b16 += b_step16;
                            }
                        } else {
                            for (int ks = 0; ks < ksteps; ++ks) {
                                umma_f16_pred(d_tmem, a16 + a_sel16, da_hi, b16, db_hi, idesc_sel, acc, leader);
                                acc = 1;
                                a16 += a_step16;
                                b16 += b_step16;
                            }
                        }
                        if constexpr (SW) a16 += a_unit16;
                        --units_left;
                        if (++u_in_blk == p.upb || units_left == 0) {
                            if (leader && !pure) umma_commit(&w_empty[st]);  // frees the weight stage once these MMAs retire
                            u_in_blk = 0;
                            if (++st == (uint32_t)p.w_stages) { st = 0; st_phase ^= 1; }
                        }
                    }
                }
            }
        }
        if (leader && !pure) {
            umma_commit(&patch_empty[ps]);
            umma_commit(&acc_full[as]);
        }
        __syncwarp();
        if (++ps == (uint32_t)p.patch_stages) { ps = 0; ps_phase ^= 1; }
        if (++as == (uint32_t)p.acc_stages) { as = 0; as_phase ^= 1; }
    }
    if (pure) {
        if (leader) umma_commit(&acc_full[0]);
        __syncwarp();
        mbar_wait(&acc_full[0], 0);
    }
    if (DBG && p.dbg && blockIdx.x == 0 && leader && sel == 0) {
        p.dbg[3] = m_wait_acc; p.dbg[4] = m_wait_patch; p.dbg[5] = m_wait_w; p.dbg[6] = TCLK() - m_total0; p.dbg[7] = it;
    }
}

// smem carve-up: [barriers 512 B][patch ring][weight stages]
template <bool DBG, bool SPLIT, bool SW>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);        // [MAX_WST]
    uint64_t* w_empty = w_full + MAX_WST;                        // [MAX_WST]
    uint64_t* patch_full = w_empty + MAX_WST;                    // [MAX_PST]
    uint64_t* patch_empty = patch_full + MAX_PST;                // [MAX_PST]
    uint64_t* acc_full = patch_empty + MAX_PST;                  // [MAX_AST]
    uint64_t* acc_empty = acc_full + MAX_AST;                    // [MAX_AST]
    uint64_t* stage_full = acc_empty + MAX_AST;                  // [2] epilogue staging tile written
    uint64_t* stage_empty = stage_full + 2;                      // [2] ... and copied out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_empty + 2);
    uint8_t* patch0 = smem + 1024;            // 1024-byte aligned: swizzled A tiles need it
    const int n_chunks = p.cin >> 3;
    const uint32_t plane_bytes = (uint32_t)p.plane_bytes;
    const uint32_t patch_bytes = 2 * plane_bytes;
    uint8_t* wst = patch0 + (size_t)p.patch_stages * patch_bytes;
    uint32_t* tab = reinterpret_cast<uint32_t*>(wst + (size_t)p.w_stages * p.stage_bytes);
    uint2* mtab = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(tab) + p.tab_bytes);
    uint8_t* epi = reinterpret_cast<uint8_t*>(mtab) + p.mtab_bytes;     // [plane][128 rows][epi_pitch] + row bases

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw_in = p.H * p.W, hw_out = p.Ho * p.Wo;

    // Programmatic dependent launch: the next kernel of the plan may start its prologue (barriers, TMEM,
    // weight stream) while this grid drains; every access to activations happens after griddepcontrol.wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (threadIdx.x == 0) {
        const uint32_t n_issuers = SPLIT ? 2 : 1;      // every MMA-issuing warp commits to the barriers it consumes through
        for (int s = 0; s < p.w_stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], n_issuers); }
        for (int s = 0; s < MAX_PST; ++s) { mbar_init(&patch_full[s], p.use_tma ? 1 : LOADER_THREADS); mbar_init(&patch_empty[s], n_issuers); }
        for (int s = 0; s < MAX_AST; ++s) { mbar_init(&acc_full[s], n_issuers); mbar_init(&acc_empty[s], EPI_WARPS * 32); }
        for (int s = 0; s < 2; ++s) { mbar_init(&stage_full[s], EPI_WARPS * 32); mbar_init(&stage_empty[s], STORE_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const bool pure = DFLAG(8);      // debugging: MMA warp alone, no barriers (measures the raw MMA rate)
    if (pure && warp != W_MMA) {
    } else if (warp < W_STORE0 + STORE_WARPS) {
        // ================= store warps (128 threads) =================
        // Copy each staged tile [plane][row][epi_pitch] to global memory with consecutive threads on consecutive
        // 16-B pieces (whole pixel rows per warp store) while the epilogue warps already convert the next tile
        // into the other staging buffer.  The loop is branch-free: the loads of four pieces are issued before the
        // first (predicated) store, so one shared-memory latency is paid per four pieces, not per piece.
        if (p.epi_bytes > 0 && p.epi_bufs == 2) {          // with one staging buffer the epilogue warps copy it out themselves
            const int stt = threadIdx.x - W_STORE0 * 32;
            const int esz = p.out_fmt == LP_FMT_SPLIT16 ? 2 : 4;
            const int n_planes = p.out_fmt == LP_FMT_SPLIT16 ? 2 : 1;
            const int cpr = (p.cout * esz) >> 4;                 // 16-B pieces per pixel row
            const unsigned magic_cpr = (unsigned)((0x100000000ull + cpr - 1) / cpr);
            const uint32_t epi_plane = (uint32_t)TILE_M * p.epi_pitch;
            const int per_plane = TILE_M * cpr;
            const int n_items = n_planes * per_plane;            // a multiple of 128 (TILE_M rows)
            asm volatile("griddepcontrol.wait;" ::: "memory");
            uint32_t sb = 0, sb_phase = 0;
            long long s_wait = 0, s_copy = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const uint8_t* stg = epi + (size_t)sb * p.epi_buf_bytes;
                const long long* row_base = reinterpret_cast<const long long*>(stg + (size_t)n_planes * epi_plane);
                const long long t0 = TCLK();
                mbar_wait(&stage_full[sb], sb_phase);
                const long long t1 = TCLK();
                for (int q0 = stt; q0 < n_items; q0 += 4 * STORE_THREADS) {
                    uint4 val[4];
                    long long base[4];
                    uint32_t off[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        int q = q0 + u * STORE_THREADS;
                        q = q < n_items ? q : stt;                 // clamp: the tail re-reads a valid piece and is not stored
                        const int pl = q >= per_plane ? 1 : 0;
                        const int qq = q - pl * per_plane;
                        const int row = (int)__umulhi((unsigned)qq, magic_cpr);
                        const int ch = qq - row * cpr;
                        base[u] = row_base[row];
                        val[u] = *reinterpret_cast<const uint4*>(stg + (uint32_t)pl * epi_plane + (uint32_t)row * (uint32_t)p.epi_pitch + ch * 16);
                        off[u] = (uint32_t)ch * 16;
                        if (pl) base[u] = base[u] < 0 ? base[u] : base[u] + p.out_plane;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (base[u] >= 0 && q0 + u * STORE_THREADS < n_items)
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + base[u] * esz + off[u]) = val[u];
                }
                mbar_arrive(&stage_empty[sb]);                   // staging buffer free again
                s_wait += t1 - t0; s_copy += TCLK() - t1;
                if (++sb == (uint32_t)p.epi_bufs) { sb = 0; sb_phase ^= 1; }
            }
            if (DBG && p.dbg && blockIdx.x == 0 && stt == 0) { p.dbg[12] = s_copy; p.dbg[15] = s_wait; }
        }
    } else if (warp < W_LOADER0 + LOADER_WARPS) {
        // ================= patch loaders (LOADER_THREADS threads) =================
        // The loader's instruction stream is on the critical path of the small-channel layers, so the
        // tile-independent geometry of every 16-B item is tabulated once; per tile an item costs ~15
        // instructions (bounds test, one multiply-add, two cp.async).
        const int lt = threadIdx.x - W_LOADER0 * 32;
        const int items_per_plane = n_chunks * p.slots;
        if (p.use_tma) {
            // ---- TMA variant: ONE thread issues 2 * n_chunks tensor copies per tile (a chunk = the 8-channel slice of the halo patch,
            // [18][10][16 B] for a 3x3 tile, [128][16 B] for a 1x1 tile; out-of-image coordinates are zero-filled by the copy engine =
            // the conv's padding); the patch barrier counts the bytes (expect_tx), so the other loader threads have nothing to do.
            if (lt == 0) {
                const void* tmap = &p.tmap;
                asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
                asm volatile("griddepcontrol.wait;" ::: "memory");
                uint32_t ps = 0, ps_phase = 0;
                int it = 0;
                const uint32_t chunk_bytes = (uint32_t)p.slots_p * 16;
                const int c_base = p.in_coff;
                for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                    if (it >= p.patch_stages) mbar_wait(&patch_empty[ps], ps_phase ^ 1);
                    const TileCoord tc = tile_coord(p, tile);
                    const uint32_t dst0 = smem_u32(patch0 + (size_t)ps * patch_bytes);
                    mbar_expect_tx(&patch_full[ps], (p.stride == 2 ? 8u : 2u) * (uint32_t)n_chunks * (uint32_t)p.tma_box_bytes);
                    if (SW) {
                        const long long lo_pix = (long long)p.max_batch * hw_in;
                        const int n_blk = n_chunks >> 3;                       // 64-channel blocks
                        for (int pl = 0; pl < 2; ++pl)
                            for (int kb = 0; kb < n_blk; ++kb)
                                tma_load_2d(dst0 + (uint32_t)pl * plane_bytes + (uint32_t)kb * 16384u, tmap, c_base + 64 * kb,
                                            (int)(tc.pix0 + pl * lo_pix), &patch_full[ps]);
                    } else if (p.ksize == 1) {
                        const long long lo_pix = (long long)p.max_batch * hw_in;
                        for (int pl = 0; pl < 2; ++pl)
                            for (int ch = 0; ch < n_chunks; ++ch)
                                tma_load_2d(dst0 + (uint32_t)pl * plane_bytes + (uint32_t)ch * chunk_bytes, tmap, c_base + 8 * ch,
                                            (int)(tc.pix0 + pl * lo_pix), &patch_full[ps]);
                    } else if (p.stride == 1) {
                        for (int pl = 0; pl < 2; ++pl)
                            for (int ch = 0; ch < n_chunks; ++ch)
                                tma_load_4d(dst0 + (uint32_t)pl * plane_bytes + (uint32_t)ch * chunk_bytes, tmap, c_base + 8 * ch,
                                            tc.ox0 - 1, tc.oy0 - 1, pl * p.max_batch + tc.img, &patch_full[ps]);
                    } else {
                        // stride 2: phase (row parity, column parity) = the window's pixels (2*sr + pr, 2*sc + pc), sampled by the map's element strides
                        const uint32_t phase_bytes = (uint32_t)p.phase_slots * 16;
                        for (int pl = 0; pl < 2; ++pl)
                            for (int ch = 0; ch < n_chunks; ++ch)
                                for (int ph = 0; ph < 4; ++ph)
                                    tma_load_4d(dst0 + (uint32_t)pl * plane_bytes + (uint32_t)ch * chunk_bytes + (uint32_t)ph * phase_bytes, tmap,
                                                c_base + 8 * ch, 2 * tc.ox0 - 1 + (ph & 1), 2 * tc.oy0 - 1 + (ph >> 1),
                                                pl * p.max_batch + tc.img, &patch_full[ps]);
                    }
                    if (++ps == (uint32_t)p.patch_stages) { ps = 0; ps_phase ^= 1; }
                }
            }
        } else {
        if (p.ksize == 3) {
            for (int slot = lt; slot < p.slots; slot += LOADER_THREADS) {
                int py, px;
                if (p.stride == 1) {
                    py = slot / p.pitch;
                    px = slot - py * p.pitch;
                } else {
                    const int ph = slot / p.phase_slots, q = slot - ph * p.phase_slots;
                    const int sr = q / p.pitch, sc = q - sr * p.pitch;
                    py = 2 * sr + (ph >> 1);
                    px = 2 * sc + (ph & 1);
                }
                tab[slot] = (uint32_t)py | ((uint32_t)px << 8);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(LOADER_THREADS) : "memory");      // loaders only
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");      // activations of the previous layer are complete and visible
        // Up to `depth` tiles are in flight per thread (one cp.async group each): a tile costs a full
        // L2/HBM round trip, so the loader must not wait for tile i before issuing tile i+1.
        const int depth = p.ld_depth;
        int issued = 0, arrived = 0;
        long long t_wait_empty = 0, t_issue = 0, t_wait_cp = 0;
        const __half* in_c = p.in + p.in_coff;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int ps = issued % p.patch_stages;
            long long t0 = TCLK();
            if (issued >= p.patch_stages) {
                // about to (possibly) block on a stage the MMA warp still reads: first publish every tile that
                // is in flight, or the MMA warp would wait for a patch that has long landed
                if (arrived < issued) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    for (; arrived < issued; ++arrived) mbar_arrive(&patch_full[arrived % p.patch_stages]);
                }
                mbar_wait(&patch_empty[ps], ((issued / p.patch_stages) - 1) & 1);
            }
            long long t1 = TCLK();
            t_wait_empty += t1 - t0;
            const TileCoord tc = tile_coord(p, tile);
            const uint32_t dst0 = smem_u32(patch0 + (size_t)ps * patch_bytes);
            if (DFLAG(1)) {
            } else if (p.ksize == 1) {
                // images are contiguous (host checks in_img == H*W*C): flattened pixel index addresses directly
                const __half* base = in_c + tc.pix0 * p.in_C;
                const int n_valid = (int)((p.total_pix - tc.pix0) < TILE_M ? (p.total_pix - tc.pix0) : TILE_M);
                for (int e = lt; e < items_per_plane; e += LOADER_THREADS) {
                    const int slot = (int)__umulhi((unsigned)e, p.magic_chunks);
                    const int chunk = e - slot * n_chunks;
                    const bool valid = slot < n_valid;
                    const __half* src = base + (valid ? slot * p.in_C : 0) + chunk * 8;
                    const uint32_t dst = dst0 + ((uint32_t)chunk * p.slots_p + slot) * 16;
                    cp_async16(dst, src, valid);
                    cp_async16(dst + plane_bytes, src + p.in_plane, valid);
                }
            } else {
                const __half* base = in_c + (long long)tc.img * p.in_img;
                const int y_base = tc.oy0 * p.stride - 1, x_base = tc.ox0 * p.stride - 1;
                for (int e = lt; e < items_per_plane; e += LOADER_THREADS) {
                    const uint32_t slot = __umulhi((unsigned)e, p.magic_chunks);       // e / n_chunks
                    const uint32_t chunk = (uint32_t)e - slot * (uint32_t)n_chunks;
                    const uint32_t t = tab[slot];
                    const int iy = y_base + (int)(t & 255), ix = x_base + (int)(t >> 8);
                    const bool valid = ((unsigned)iy < (unsigned)p.H) && ((unsigned)ix < (unsigned)p.W);
                    const __half* src = base + (valid ? (iy * p.W + ix) * p.in_C : 0) + chunk * 8;
                    const uint32_t dst = dst0 + (chunk * p.slots_p + slot) * 16;
                    cp_async16(dst, src, valid);
                    cp_async16(dst + plane_bytes, src + p.in_plane, valid);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            ++issued;
            long long t2 = TCLK();
            t_issue += t2 - t1;
            if (issued - arrived > depth) {
                switch (depth) {
                    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                    default: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                mbar_arrive(&patch_full[arrived % p.patch_stages]);
                ++arrived;
                t_wait_cp += TCLK() - t2;
            }
        }
        if (DBG && p.dbg && blockIdx.x == 0 && lt == 0) { p.dbg[0] = t_wait_empty; p.dbg[1] = t_issue; p.dbg[2] = t_wait_cp; }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (; arrived < issued; ++arrived) mbar_arrive(&patch_full[arrived % p.patch_stages]);
        }
    } else if (warp == W_PRODUCER) {
        // ================= weight producer (bulk TMA) =================
        if (lane == 0) {
            int g = 0;                                        // running K-block counter across tiles
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                if (p.resident && tile != (int)blockIdx.x) break;
                for (int kb = 0; kb < p.n_kb; ++kb, ++g) {
                    const int s = g % p.w_stages;
                    if (g >= p.w_stages) mbar_wait(&w_empty[s], ((g / p.w_stages) - 1) & 1);
                    if (DFLAG(4) && g >= p.w_stages) { mbar_arrive(&w_full[s]); continue; }
                    const int units = min(p.upb, p.n_units - kb * p.upb);        // the last block of a tile may be short
                    const uint32_t bytes = (uint32_t)units * p.unit_bytes;
                    mbar_expect_tx(&w_full[s], bytes);
                    bulk_g2s(wst + (size_t)s * p.stage_bytes, p.wtc + (size_t)kb * p.stage_bytes, bytes, &w_full[s]);
                }
            }
        }
    } else if (warp == W_MMA || (SPLIT && warp == W_MMA2)) {
        // ================= MMA issuers =================
        // The issue path of ONE warp is the bound of the 3x3 layers (~87 cycles per tcgen05.mma against a pipe floor of
        // 40-64: ptxas needs ~5 uniform-datapath instructions per UTCHMMA for fresh descriptor register pairs).  With
        // split_mma the two MMAs of a K-step therefore come from two warps that walk the same tiles, patches and
        // weight stages in lockstep: A issues Ahi x [Bhi|Blo] into columns [0, 2*cout), B issues Alo x Bhi into columns
        // [2*cout, 3*cout) (its own columns: two threads must not accumulate into the same TMEM cells); the epilogue
        // adds the three pieces.  Every barrier the issuers signal through tcgen05.commit counts two arrivals.
        // The whole warp runs this loop with warp-uniform values (uniform datapath); one elected lane
        // issues.  Measured: a dependent scalar instruction costs ~6 cycles with a single warp and a
        // tcgen05.mma ~46-64, so the loop body between two MMAs must be a handful of uniform adds.  The
        // operand addresses of a tile are two arithmetic progressions: for each tap, A starts at
        // patch + tap_offset and advances 2*LBO per K-step; B advances 2*LBO_B per K-step through the
        // (contiguous) weight stages.
        // redux.sync: the result lives in a uniform register, so the operand selection below stays in the uniform datapath
        const uint32_t sel = SPLIT ? (uint32_t)__reduce_max_sync(0xffffffffu, warp == W_MMA2 ? 1 : 0) : 0u;
        mma_role<SPLIT ? 1 : 3, DBG, SW>(p, tmem_base, pure, patch0, wst, plane_bytes, patch_bytes, w_full, w_empty, patch_full, patch_empty,
                                     acc_full, acc_empty, sel);
    } else if (warp == W_MMA2) {
        // idle in the one-issuer kernel
    } else {
        // ================= epilogue (warps 10-17) =================
        // Phase 1 (thread = accumulator row, warp half = column half): TMEM -> bias/act/residual -> split ->
        // staging tile in shared memory [plane][row][epi_pitch].  Phase 2: all 256 threads copy the staged
        // tile to global with consecutive threads on consecutive 16-B chunks, so a warp store covers whole
        // pixel rows (4-8 cache lines) instead of 32 scattered 16-B pieces.
        const int quad = warp & 3, half = (warp - W_EPI0) >> 2;        // TMEM lane quadrant (= warp % 4, a hardware rule), column half
        const int etid = threadIdx.x - W_EPI0 * 32;
        const int c_split = (((p.cout >> 4) + 1) >> 1) << 4;   // 16-column groups: first ceil(n/2) to half 0, rest to half 1
        const int c_begin = half ? c_split : 0, c_end = half ? p.cout : c_split;
        const int esz = p.out_fmt == LP_FMT_SPLIT16 ? 2 : 4;
        const int n_planes = p.out_fmt == LP_FMT_SPLIT16 ? 2 : 1;
        const uint32_t epi_plane = (uint32_t)TILE_M * p.epi_pitch;
        const bool staged = p.epi_bytes > 0;
        const bool handoff = p.epi_bufs == 2;               // two staging buffers: dedicated store warps copy them out
        uint32_t sb = 0, sb_phase = 0;                      // staging buffer of this tile
        asm volatile("griddepcontrol.wait;" ::: "memory");      // before any residual read / output write
        int it = 0;
        uint32_t as = 0, as_phase = 0;
        long long e_wait = 0, e_total0 = TCLK(), e_p1 = 0, e_bar = 0, e_p2 = 0, e_ld = 0, e_pre = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const long long e_top = TCLK();
            const int r = quad * 32 + lane;                  // accumulator row == TMEM lane == tile pixel
            bool valid;
            int oimg, opin;                                  // image and pixel-in-image of this row
            if (p.ksize == 1) {
                // hw_out <= 2^16 and the pixel index < 2^31: the float estimate is off by at most one
                const long long gp = (long long)tile * TILE_M + r;
                valid = gp < p.total_pix;
                const unsigned gpu = valid ? (unsigned)gp : 0u;
                unsigned q = (unsigned)((float)gpu * p.inv_hw_out);
                if (q * (unsigned)hw_out > gpu) --q;
                if ((q + 1) * (unsigned)hw_out <= gpu) ++q;
                oimg = (int)q;
                opin = (int)(gpu - q * (unsigned)hw_out);
            } else {
                // division by multiply-high; ceil(2^32 / 1) does not fit 32 bits, so a divisor of 1 (maps no larger than one
                // 16x8 tile: the 8x8 .. 2x2 stages of the classifier backbones) is the identity
                const unsigned per_img = (unsigned)(p.tiles_x * p.tiles_y);
                const unsigned img = per_img == 1 ? (unsigned)tile : __umulhi((unsigned)tile, p.magic_per_img);
                const unsigned rr = (unsigned)tile - img * per_img;
                const unsigned ty = p.tiles_x == 1 ? rr : __umulhi(rr, p.magic_tiles_x), tx = rr - ty * (unsigned)p.tiles_x;
                const int oy = (int)ty * TCT_H + (r >> 3), ox = (int)tx * TCT_W + (r & 7);
                valid = (oy < p.Ho && ox < p.Wo);
                oimg = (int)img;
                opin = oy * p.Wo + ox;
            }
            const long long obase = (long long)oimg * p.out_img + (long long)opin * p.out_C + p.out_coff;
            const long long rbase = (long long)oimg * p.res_img + (long long)opin * p.res_C + p.res_coff;
            uint8_t* stg = epi + (size_t)sb * p.epi_buf_bytes;
            if (staged) {
                const long long t0 = TCLK();
                if (handoff && it >= p.epi_bufs) mbar_wait(&stage_empty[sb], sb_phase ^ 1);      // the store warps have drained it
                e_bar += TCLK() - t0;
                if (half == 0) reinterpret_cast<long long*>(stg + (size_t)n_planes * epi_plane)[r] = valid ? obase : -1;
            }
            { long long t0 = TCLK();
              e_pre += t0 - e_top;
              mbar_wait(&acc_full[as], as_phase);
              e_wait += TCLK() - t0; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long tp1 = TCLK();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (uint32_t)p.acc_stride;
            for (int c0 = c_begin; c0 < (DFLAG(2) ? c_begin : c_end); c0 += 16) {
                uint32_t v[16], v2[16];
                tmem_ld16(trow + c0, v);                       // Ahi*Bhi (+ Alo*Bhi when one warp issues both)
                tmem_ld16(trow + p.cout + c0, v2);             // Ahi*Blo
                if (SPLIT) {
                    uint32_t v3[16];
                    tmem_ld16(trow + 2 * p.cout + c0, v3);     // Alo*Bhi from the second issuing warp
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v3[i]));
                }
                const long long tl0 = TCLK();
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                e_ld += TCLK() - tl0;
                if (!valid) continue;
                float f[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * q));
                    f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + __uint_as_float(v2[4 * q + 0]) + b.x;
                    f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + __uint_as_float(v2[4 * q + 1]) + b.y;
                    f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + __uint_as_float(v2[4 * q + 2]) + b.z;
                    f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + __uint_as_float(v2[4 * q + 3]) + b.w;
                }
                // residual: after the activation (C2f / Bottleneck shortcut) or, with res_first, before it (torchvision BasicBlock)
                for (int pass = 0; pass < 2; ++pass) {
                if (pass == (p.res_first ? 1 : 0)) {
                if (p.act == LP_ACT_SILU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = lp_silu(f[i]);
                } else if (p.act == LP_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
                } else if (p.act == LP_ACT_RELU6) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = fminf(fmaxf(f[i], 0.f), 6.f);
                }
                } else if (p.res) {
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(p.res + rbase + c0 + 8 * h8);
                        const uint4 rl = *reinterpret_cast<const uint4*>(p.res + p.res_plane + rbase + c0 + 8 * h8);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
                        const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 a = __half22float2(h2[i]), b = __half22float2(l2[i]);
                            f[8 * h8 + 2 * i] += a.x + b.x;
                            f[8 * h8 + 2 * i + 1] += a.y + b.y;
                        }
                    }
                }
                }
                if (p.out_fmt == LP_FMT_SPLIT16) {
                    uint4 oh[2], ol[2];
                    __half2* h2 = reinterpret_cast<__half2*>(oh);
                    __half2* l2 = reinterpret_cast<__half2*>(ol);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {            // packed split: hi = rn(v), lo = rn(v - hi)
                        const __half2 hi = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                        const float2 hf = __half22float2(hi);
                        h2[i] = hi;
                        l2[i] = __floats2half2_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
                    }
                    if (p.seg_len > 0) {                   // element scatter (ShuffleNetV2 channel shuffle)
                        __half* o = reinterpret_cast<__half*>(p.out) + (long long)oimg * p.out_img + (long long)opin * p.out_C;
                        const __half* hh = reinterpret_cast<const __half*>(oh);
                        const __half* ll = reinterpret_cast<const __half*>(ol);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int j = c0 + i;
                            if (j < p.cout_real) {
                                const int l = p.seg_l0 + j * p.out_cstride;
                                const int ph = (l / p.seg_len) * p.seg_pad + l % p.seg_len;
                                o[ph] = hh[i];
                                o[ph + p.out_plane] = ll[i];
                            }
                        }
                    } else if (staged) {
                        uint8_t* sp = stg + (size_t)r * p.epi_pitch + c0 * 2;
                        *reinterpret_cast<uint4*>(sp) = oh[0];
                        *reinterpret_cast<uint4*>(sp + 16) = oh[1];
                        *reinterpret_cast<uint4*>(sp + epi_plane) = ol[0];
                        *reinterpret_cast<uint4*>(sp + epi_plane + 16) = ol[1];
                    } else {
                        __half* o = reinterpret_cast<__half*>(p.out) + obase + c0;
                        *reinterpret_cast<uint4*>(o) = oh[0];
                        *reinterpret_cast<uint4*>(o + 8) = oh[1];
                        *reinterpret_cast<uint4*>(o + p.out_plane) = ol[0];
                        *reinterpret_cast<uint4*>(o + p.out_plane + 8) = ol[1];
                    }
                } else if (staged) {
                    uint8_t* sp = stg + (size_t)r * p.epi_pitch + c0 * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(sp + 16 * q) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                } else {
                    float* o = reinterpret_cast<float*>(p.out) + obase + c0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(o + 4 * q) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[as]);                     // accumulator drained: the MMA warp may reuse it
            if (++as == (uint32_t)p.acc_stages) { as = 0; as_phase ^= 1; }
            const long long tp2 = TCLK();
            e_p1 += tp2 - tp1;
            if (staged && handoff) {
                mbar_arrive(&stage_full[sb]);                    // release: the staged tile is visible to the store warps
                if (++sb == (uint32_t)p.epi_bufs) { sb = 0; sb_phase ^= 1; }
            } else if (staged) {
                // single staging buffer (the MMA-bound streaming layers): all 256 epilogue threads copy the tile out,
                // consecutive threads on consecutive 16-B pieces
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const long long tp3 = TCLK();
                const int cpr = (p.cout * esz) >> 4;
                const unsigned magic_cpr = (unsigned)((0x100000000ull + cpr - 1) / cpr);
                const int per_plane = TILE_M * cpr;
                const long long* row_base = reinterpret_cast<const long long*>(stg + (size_t)n_planes * epi_plane);
                for (int q = etid; q < n_planes * per_plane; q += 256) {
                    const int pl = q >= per_plane ? 1 : 0;
                    const int qq = q - pl * per_plane;
                    const int row = (int)__umulhi((unsigned)qq, magic_cpr);
                    const int ch = qq - row * cpr;
                    const long long base = row_base[row];
                    if (base < 0) continue;
                    const uint4 val = *reinterpret_cast<const uint4*>(stg + (size_t)pl * epi_plane + (size_t)row * p.epi_pitch + ch * 16);
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + ((base + (long long)pl * p.out_plane) * esz) + ch * 16) = val;
                }
                e_p2 += TCLK() - tp3;
                asm volatile("bar.sync 2, 256;" ::: "memory");   // staging tile free for the next tile
            }
        }
        if (DBG && p.dbg && blockIdx.x == 0 && etid == 0) { p.dbg[8] = e_wait; p.dbg[9] = TCLK() - e_total0; p.dbg[10] = e_p1; p.dbg[11] = e_bar; p.dbg[13] = e_ld; p.dbg[14] = e_pre; if (!handoff) p.dbg[12] = e_p2; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

unsigned magic_u32(int n) { return (unsigned)((0x100000000ull + (unsigned)n - 1) / (unsigned)n); }

}  // namespace

static int conv_tc_block(lp_ctx* ctx, lp_net_plan& net, const lp_op_desc& op, int batch, uint8_t* ws, cudaStream_t st,
                         int n0, int nb, size_t wtc_off);

// cuTensorMapEncodeTiled through the runtime's driver entry point table: the library must load (and export its symbols) on a
// machine without libcuda.so.1 -- the CPU-only build / test box -- so it carries no link-time dependency on the driver.
typedef CUresult (*lp_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static lp_encode_tiled_fn lp_encode_tiled() {
    static lp_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<lp_encode_tiled_fn>(p);
    }
    return fn;
}

// returns 1 if the op ran on the tensor cores, 0 if it is not eligible, <0 on error
int lp_conv_tc_try(lp_ctx* ctx, lp_net_plan& net, const lp_op_desc& op, int batch, uint8_t* ws, cudaStream_t st) {
    const lp_buf_desc& ib = net.bufs[op.in_buf];
    if (op.kind != LP_OP_CONV || ib.fmt != LP_FMT_SPLIT16 || op.cout % 16) return 0;
    int kb_ch = 0;
    for (int d : {64, 48, 32, 16}) if (op.cin % d == 0) { kb_ch = d; break; }
    if (!kb_ch) return 0;
    // output channels in blocks of <= 128 (N of the [Bhi|Blo] MMA is 2*block <= 256); the packed weights
    // hold the blocks back to back (plan.py pack_tc_weights)
    size_t woff = 0;
    for (int n0 = 0; n0 < op.cout; n0 += 128) {
        const int nb = op.cout - n0 < 128 ? op.cout - n0 : 128;
        const int r = conv_tc_block(ctx, net, op, batch, ws, st, n0, nb, woff);
        if (r != 1) {
            if (n0 == 0) return r;
            lp_set_error("conv_tc: output block %d of an op became ineligible", n0);
            return -1;
        }
        woff += (size_t)op.ksize * op.ksize * op.cin * nb * 4;
        if (n0 + 128 < op.cout) ctx->launches++;
    }
    return 1;
}

static int conv_tc_block(lp_ctx* ctx, lp_net_plan& net, const lp_op_desc& op, int batch, uint8_t* ws, cudaStream_t st,
                         int n0, int nb, size_t wtc_off) {
    const lp_buf_desc& ib = net.bufs[op.in_buf];
    const lp_buf_desc& ob = net.bufs[op.out_buf];
    if (op.kind != LP_OP_CONV || ib.fmt != LP_FMT_SPLIT16) return 0;
    if (!((op.ksize == 1 && op.stride == 1) || (op.ksize == 3 && (op.stride == 1 || op.stride == 2)))) return 0;
    const bool seg = op.out_seg_len > 0;
    if (op.cin % 16 || nb % 16 || op.cin > 512 || op.in_coff % 8) return 0;
    if (!seg && (op.out_cstride > 1 || op.out_coff % 4)) return 0;
    if (seg && ob.fmt != LP_FMT_SPLIT16) return 0;
    if (!seg && ob.fmt == LP_FMT_SPLIT16 && op.out_coff % 8) return 0;

    TcParams p{};
    const long long in_img = ib.image_bytes / 2;
    p.in = reinterpret_cast<const __half*>(ws + ib.offset);
    p.in_plane = (long long)net.max_batch * in_img; p.in_img = in_img; p.in_C = ib.c; p.in_coff = op.in_coff;
    const int oesz = ob.fmt == LP_FMT_SPLIT16 ? 2 : 4;
    p.out = ws + ob.offset + (size_t)op.row_off * ob.c * oesz;
    p.out_img = ob.image_bytes / oesz; p.out_plane = (long long)net.max_batch * p.out_img;
    p.out_C = ob.c; p.out_coff = seg ? 0 : op.out_coff; p.out_fmt = ob.fmt;
    if (op.res_buf >= 0) {
        const lp_buf_desc& rb = net.bufs[op.res_buf];
        if (rb.fmt != LP_FMT_SPLIT16 || op.res_coff % 8) return 0;
        p.res = reinterpret_cast<const __half*>(ws + rb.offset);
        p.res_img = rb.image_bytes / 2; p.res_plane = (long long)net.max_batch * p.res_img; p.res_C = rb.c;
        p.res_coff = op.res_coff + n0;                // this launch covers output channels [n0, n0 + nb) of the op
    }
    p.dbg = ctx->tc_dbg;
    { static int f = -1; if (f < 0) { const char* e = getenv("LP_TC_DEBUG"); f = e ? atoi(e) : 0; } p.dbg_flags = f; }
    p.wtc = net.weights_tc + op.wtc_off + wtc_off;
    p.bias = net.weights + op.b_off + n0;
    p.cin = op.cin; p.cout = nb; p.act = op.act; p.res_first = (op.flags & LP_OPF_RES_BEFORE_ACT) ? 1 : 0;
    if (op.act != LP_ACT_NONE && op.act != LP_ACT_SILU && op.act != LP_ACT_RELU && op.act != LP_ACT_RELU6) return 0;
    p.out_cstride = op.out_cstride > 0 ? op.out_cstride : 1;
    p.seg_len = op.out_seg_len; p.seg_pad = op.out_seg_pad;
    p.seg_l0 = seg ? op.out_coff + n0 * p.out_cstride : 0;
    const int real_all = op.cout_real > 0 ? op.cout_real : n0 + nb;
    p.cout_real = real_all - n0 < nb ? (real_all - n0 < 0 ? 0 : real_all - n0) : nb;
    if (!seg) p.out_coff += n0;
    p.H = ib.h; p.W = ib.w; p.n_img = batch; p.ksize = op.ksize; p.stride = op.stride;
    p.Ho = (ib.h + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1;
    p.Wo = (ib.w + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1;
    const int taps = op.ksize * op.ksize;
    // Patch loads by TMA tensor copies.  ctx->tc_tma (env LP_TC_TMA): bit 0 = 3x3 stride 1, bit 1 = 1x1, bit 2 = 3x3 stride 2; default all.
    { const int f = ctx->tc_tma;
      p.use_tma = ((op.ksize == 3 && op.stride == 1 && (f & 1)) || (op.ksize == 1 && (f & 2)) || (op.ksize == 3 && op.stride == 2 && (f & 4))) ? 1 : 0; }
    if (p.use_tma && (p.dbg_flags & 1)) p.use_tma = 0;
    if (op.ksize == 1) { p.slots = TILE_M; p.pitch = 8; }
    else if (op.stride == 1) { p.pitch = TCT_W + 2; p.slots = (TCT_H + 2) * p.pitch; }
    else {
        // stride 2: the four (row, column) parity phases of the 33 x 17 patch, 17 x 9 slots each; a tensor copy (one per phase) needs a
        // 128-byte aligned destination, so with TMA a phase occupies 160 slots
        p.pitch = TCT_W + 1;
        p.phase_slots = p.use_tma ? ((TCT_H + 1) * p.pitch + 7) / 8 * 8 : (TCT_H + 1) * p.pitch;
        p.slots = 4 * p.phase_slots;
    }
    for (int t = 0; t < taps; ++t) {
        const int ky = t / 3, kx = t % 3;
        p.tap_off[t] = op.ksize == 1 ? 0 : op.stride == 1 ? ky * p.pitch + kx : ((ky & 1) * 2 + (kx & 1)) * p.phase_slots + (ky >> 1) * p.pitch + (kx >> 1);
    }
    // chunk stride: == 1 (mod 8) slots keeps the cp.async writes conflict-free; a tensor copy needs a 128-byte aligned destination
    p.slots_p = p.use_tma ? (p.slots + 7) / 8 * 8 : p.slots + ((9 - (p.slots & 7)) & 7);
    p.plane_bytes = (op.cin / 8) * p.slots_p * 16;
    p.tma_box_bytes = (op.ksize == 3 && op.stride == 2) ? (TCT_H + 1) * p.pitch * 16 : p.slots * 16;     // stride 2: one box per phase
    p.max_batch = net.max_batch;
    p.magic_chunks = magic_u32(op.cin / 8);
    p.magic_pitch = magic_u32(p.pitch);
    const size_t patch_bytes = (size_t)2 * (op.cin / 8) * p.slots_p * 16;
    p.tab_bytes = op.ksize == 3 ? (p.slots * 4 + 127) / 128 * 128 : 0;
    if (op.ksize == 1 && ib.image_bytes != (int64_t)ib.h * ib.w * ib.c * 2) return 0;
    p.n_mma = taps * (op.cin / 16) * 2;
    p.mtab_bytes = 0;
    const int out_row_bytes = nb * (ob.fmt == LP_FMT_SPLIT16 ? 2 : 4);
    p.epi_pitch = out_row_bytes + 16;
    const size_t epi_full = (size_t)(ob.fmt == LP_FMT_SPLIT16 ? 2 : 1) * TILE_M * p.epi_pitch + TILE_M * 8;
    const size_t total = 220 * 1024 - 1024 - p.tab_bytes - p.mtab_bytes;
    const size_t w_all = (size_t)taps * op.cin * nb * 4;
    // Shared-memory plan.  Weights: resident if the whole layer fits beside >= 2 patch stages, else a ring
    // of 2-4 stages.  Epilogue staging
    // (coalesced stores) if >= 2 patch stages still fit.  Everything left goes to patch stages.
    auto plan = [&](int epi_bufs) -> bool {
        const size_t epi_need = (size_t)epi_bufs * epi_full;
        if (total < epi_need) return false;
        const size_t budget = total - epi_need;
        if (w_all + 2 * patch_bytes <= budget && p.n_kb <= MAX_WST) { p.resident = 1; p.w_stages = p.n_kb; }
        else {
            p.resident = 0;
            const int min_ps = (2 * patch_bytes + 2 * (size_t)p.stage_bytes <= budget) ? 2 : 1;
            if (min_ps * patch_bytes + 2 * (size_t)p.stage_bytes > budget) return false;
            int cap = (int)((budget - min_ps * patch_bytes) / p.stage_bytes);
            if (cap > 4) cap = 4;
            if (cap > p.n_kb) cap = p.n_kb;
            if (cap < 2) return false;
            p.w_stages = cap;
        }
        const size_t left = budget - (size_t)p.w_stages * p.stage_bytes;
        int ps = (int)(left / patch_bytes);
        p.patch_stages = ps > MAX_PST ? MAX_PST : ps;
        if (p.patch_stages < 1) return false;
        p.epi_bufs = epi_bufs > 0 ? epi_bufs : 1;
        p.epi_buf_bytes = (int)epi_full;
        p.epi_bytes = (int)epi_need;
        return epi_bufs == 0 || p.patch_stages >= 2;
    };
    // K-block size: the packed weights are [tap][8-channel chunk][plane][n][8], so any multiple of 16 that
    // divides Cin is a valid block; take the largest whose ring fits.  Staging: two buffers (conversion of
    // tile i+1 overlaps the copy-out of tile i) if they fit, else one, else direct stores.  Streaming layers
    // are MMA-bound: they prefer a larger weight block over the second staging buffer.
    const int max_bufs = seg ? 0 : 2;
    bool ok = false;
    for (int d : {64, 48, 32, 16}) {
        if (op.cin % d) continue;
        p.kb_ch = d; p.n_cb = op.cin / d; p.n_units = taps * p.n_cb;
        p.unit_bytes = d * nb * 4;                                   // 2 planes x kb_ch x cout x 2 B
        auto set_upb = [&](int upb) {
            p.upb = upb; p.n_kb = (p.n_units + upb - 1) / upb;
            p.stage_bytes = upb * p.unit_bytes;
        };
        // resident weights: as few blocks as the barrier count allows
        set_upb((p.n_units + MAX_WST - 1) / MAX_WST);
        for (int bufs = max_bufs; bufs >= 0 && !ok; --bufs)
            if (plan(bufs) && p.resident) ok = true;
        for (int bufs = max_bufs > 1 ? 1 : max_bufs; bufs >= 0 && !ok; --bufs)
            for (int upb : {3, 2, 1}) {
                if (upb > p.n_units) continue;
                set_upb(upb);
                if (plan(bufs) && (upb == 1 || p.patch_stages >= 2)) {
                    if (bufs == 1 && max_bufs == 2 && plan(2) && (upb == 1 || p.patch_stages >= 2)) { ok = true; break; }   // second buffer for free
                    plan(bufs);
                    ok = true; break;
                }
            }
        if (ok) break;
    }
    if (!ok) return 0;
    p.fills_per_tile = p.resident ? 0 : p.n_kb / p.w_stages;
    // swizzled A tiles: 1x1 layers whose K splits into 64-channel blocks (same bytes per patch stage as the chunk layout)
    p.a_sw128 = (ctx->tc_sw128 && p.use_tma && op.ksize == 1 && op.cin % 64 == 0 && p.kb_ch == 64) ? 1 : 0;
    // Two issuing warps (split_mma) need a third accumulator piece per stage.  Measured (profiles/r2_notes.md): with
    // cout <= 32 (four accumulator stages of 3 * cout columns still fit the 512 TMEM columns) the 3x3 layers gain 8-10 %.
    // At cout = 64 only two stages fit: with cp.async loaders that lost (conv_48: 110 -> 122 us); with TMA loads (five warps
    // fewer on the issue ports) it wins on the LONG layers (conv_48 118 -> 108 us, 1x1 96 -> 64 at 40x40 27 -> 23) and still loses where
    // a CTA sees only a few tiles (20x20 64 -> 64: 24 -> 27), so those need >= 5 tiles per CTA.  LP_TC_SPLIT=0 disables, =2 forces it
    // wherever two stages fit.
    { static int f = -1; if (f < 0) { const char* e = getenv("LP_TC_SPLIT"); f = e ? atoi(e) : 1; }
      // from the plan's CAPACITY (max_batch), never from the batch of this call: one and two issuers sum the three products in a
      // different order, and a frame's result must not depend on how many frames travel with it (test_full_batch_properties)
      const long long out_px = (long long)net.max_batch * p.Ho * p.Wo;
      const long long tiles = op.ksize == 1 ? (out_px + TILE_M - 1) / TILE_M
                                            : (long long)net.max_batch * ((p.Wo + TCT_W - 1) / TCT_W) * ((p.Ho + TCT_H - 1) / TCT_H);
      const bool long_layer = p.use_tma && tiles >= 5ll * ctx->sm_count;
      p.split_mma = (!p.a_sw128 && (f == 2 ? 3 * nb <= 256 : (f == 1 && (nb <= 32 || (3 * nb <= 256 && long_layer))))) ? 1 : 0; }
    const int acc_cols = (p.split_mma ? 3 : 2) * nb;
    p.acc_stride = (acc_cols + 31) / 32 * 32;        // [Ahi*Bhi(+Alo*Bhi) | Ahi*Blo | Alo*Bhi]
    p.acc_stages = 512 / p.acc_stride > MAX_AST ? MAX_AST : 512 / p.acc_stride;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.acc_stages * p.acc_stride) p.tmem_cols <<= 1;
    const size_t smem = 1024 + (size_t)p.patch_stages * patch_bytes + (size_t)p.w_stages * p.stage_bytes + p.tab_bytes + p.mtab_bytes + p.epi_bytes;

    p.inv_hw_out = 1.0f / (float)(p.Ho * p.Wo);
    if (op.ksize == 1) {
        p.total_pix = (long long)batch * ib.h * ib.w;
        p.n_tiles = (int)((p.total_pix + TILE_M - 1) / TILE_M);
    } else {
        p.tiles_x = (p.Wo + TCT_W - 1) / TCT_W;
        p.tiles_y = (p.Ho + TCT_H - 1) / TCT_H;
        p.n_tiles = p.tiles_x * p.tiles_y * batch;
        p.magic_per_img = magic_u32(p.tiles_x * p.tiles_y);
        p.magic_tiles_x = magic_u32(p.tiles_x);
    }
    if (!(ctx->attr_set & 1)) {          // per context (= per device): the opt-in is a per-device function attribute
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        ctx->attr_set |= 1;
    }
    // Grid: one persistent CTA per SM.  (Tried: fewer CTAs with >= 4 / 6 / 10 tiles each on the layers with few tiles per SM, so that the
    // SMs left free run the other batches in flight: throughput unchanged at 4 and 6, -8 % at 10, single-batch latency +9 .. +46 %.)
    const int grid = p.n_tiles < ctx->sm_count ? p.n_tiles : ctx->sm_count;
    if (p.use_tma) {
        // the tensor map is a view of the whole input buffer; the lo plane is image max_batch + i (make_ref: plane = max_batch images)
        CUresult cr;
        const lp_encode_tiled_fn encode = lp_encode_tiled();
        LP_CHECK(encode != nullptr, "conv_tc: the driver does not export cuTensorMapEncodeTiled");
        if (op.ksize == 1) {
            const cuuint64_t dims[2] = {(cuuint64_t)ib.c, (cuuint64_t)2 * net.max_batch * ib.h * ib.w};
            const cuuint64_t strides[1] = {(cuuint64_t)ib.c * 2};
            const cuuint32_t box[2] = {p.a_sw128 ? 64u : 8u, (cuuint32_t)TILE_M}, es[2] = {1, 1};
            cr = encode(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(p.in), dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, p.a_sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t dims[4] = {(cuuint64_t)ib.c, (cuuint64_t)ib.w, (cuuint64_t)ib.h, (cuuint64_t)2 * net.max_batch};
            const cuuint64_t strides[3] = {(cuuint64_t)ib.c * 2, (cuuint64_t)ib.w * ib.c * 2, (cuuint64_t)ib.image_bytes};
            // stride 1: the 18 x 10 halo patch; stride 2: every other pixel of a 33 x 17 window = one 17 x 9 parity phase per copy
            const cuuint32_t s2 = op.stride == 2 ? 1u : 0u;
            const cuuint32_t box[4] = {8, s2 ? 2u * TCT_W + 1u : (cuuint32_t)p.pitch, s2 ? 2u * TCT_H + 1u : (cuuint32_t)(TCT_H + 2), 1};
            const cuuint32_t es[4] = {1, s2 ? 2u : 1u, s2 ? 2u : 1u, 1};
            cr = encode(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(p.in), dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        LP_CHECK(cr == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed (%d) for %dx%d cin %d", (int)cr, op.ksize, op.ksize, op.cin);
    }
    {   // loader look-ahead: deep for the long HBM-bound layers; CTAs that only see a few tiles publish the first one early
        static int f = -2; if (f == -2) { const char* e = getenv("LP_TC_DEPTH"); f = e ? atoi(e) : -1; }
        const int tiles_per_cta = (p.n_tiles + grid - 1) / grid;
        // measured (tools/op_times.py): publishing every tile as soon as it has landed (0) wins wherever the loaders are not the
        // bound; only the long 1x1 layers, whose tile period is close to the loaders' issue + latency time, want look-ahead
        int d = f >= 0 ? f : ((tiles_per_cta > 8 && p.ksize == 1) ? 4 : 0);
        if (d > p.patch_stages - 1) d = p.patch_stages - 1;
        if (d > 4) d = 4;
        p.ld_depth = d;
    }
    { static int f = -1; if (f < 0) { const char* e = getenv("LP_TC_PLAN"); f = e ? atoi(e) : 0; }
      if (f) fprintf(stderr, "conv_tc plan: %dx%d s%d cin %d cout %d(+%d) %dx%d | %s kb_ch %d units %d upb %d blocks %d w_stages %d stage %d B | patches %d x %zu B | epi %d B | acc %d x %d cols | tiles %d grid %d smem %zu\n",
                     op.ksize, op.ksize, op.stride, op.cin, nb, n0, p.H, p.W, p.resident ? "resident" : "stream", p.kb_ch, p.n_units, p.upb, p.n_kb,
                     p.w_stages, p.stage_bytes, p.patch_stages, patch_bytes, p.epi_bytes, p.acc_stages, p.acc_stride, p.n_tiles, grid, smem); }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = ctx->use_pdl ? 1 : 0;
    const bool dbg_kernel = p.dbg != nullptr || p.dbg_flags != 0;
    cudaError_t e;
    if (p.a_sw128) e = dbg_kernel ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false, true>, p) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, true>, p);
    else if (p.split_mma) e = dbg_kernel ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true, false>, p) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true, false>, p);
    else e = dbg_kernel ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false, false>, p) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, false>, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        lp_set_error("conv_tc launch failed: %s (smem %zu)", cudaGetErrorString(e), smem);
        return -2;
    }
    return 1;
}
