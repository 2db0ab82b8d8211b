// Small-channel convolutions on the warp-level tensor-core path (mma.sync m16n8k16, split-f16 operands).
//
// The first detector layers (model.ncnn.param:6-18: 3x3 s2 8->16, C2f cv1 1x1 16->16, bottleneck 3x3 8->8 twice, cv2 1x1
// 24->16, all at 160x160) have GEMM shapes N = 8..16, K = 16..72.  A tcgen05 tile (M = 128, accumulators in TMEM, five warp
// roles and six barrier hand-offs per tile) costs ~2500 fixed cycles per 128 pixels there -- slower than CUDA cores (r1
// notes) -- and the fp32-FMA kernels that ran them instead are bound by instruction issue (ncu, profiles/r2_ncu_stage_kernels.txt:
// 88-90 % issue-slot utilisation at 14-47 % of DRAM throughput).  A warp-level MMA needs no hand-off at all: a warp owns
// 2 x 16 output pixels, takes its A fragments with ldmatrix straight from the shared-memory halo patch (one 16-byte row =
// the 8 channels of one pixel of one tap, so im2col is only an address), keeps every weight fragment of the layer in
// registers, and finishes bias / SiLU / residual / split in registers.  ~5x fewer instructions per pixel than the FMA
// kernel; the layers move to their HBM bound.
//
// Numerics = the tcgen05 path's: value = hi + lo (fp16 each), product = Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulate.
// Optional fused 1x1 (COUT -> COUT) on the activated result: the accumulator fragment of m16n8k16 IS the A fragment of the
// next MMA (rows = pixels, k = channels), so the C2f cv1 behind the down-sampling conv never leaves the registers.
#include "common.cuh"

namespace {

constexpr int MM_THREADS = 128;            // 4 warps; block tile = 8 output rows x 16 output columns, 2 rows per warp
constexpr int MM_TH = 8, MM_TW = 16;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;               // src-size 0: 16 bytes of zeros (halo / padding)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float act1(float v, int act) {
    if (act == LP_ACT_SILU) return lp_silu(v);
    if (act == LP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == LP_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}

// (a, b) fp32 -> packed hi half2 and lo half2 (hi = rn(v), lo = rn(v - hi)): the split-f16 storage format
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// weight fragment of one k-step (two 8-channel slots) and one n-tile for this lane, split into hi / lo halves.
// w = fp32 [tap][cin][cout] of the plan; slot s = tap * CH + chunk; B[k][n]: k 0-7 -> slot s0, 8-15 -> slot s1.
__device__ __forceinline__ void load_bfrag(const float* __restrict__ w, int cin, int cout, int slot0, int slot1, int n_slots, int CH, int nt,
                                           int lane, uint32_t (&bh)[2], uint32_t (&bl)[2]) {
    const int n = nt * 8 + (lane >> 2), kc = (lane & 3) * 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int slot = h ? slot1 : slot0;
        float x0 = 0.f, x1 = 0.f;
        if (slot < n_slots && n < cout) {
            const int tap = slot / CH, c = (slot - tap * CH) * 8 + kc;
            x0 = __ldg(w + ((long long)tap * cin + c) * cout + n);
            x1 = __ldg(w + ((long long)tap * cin + c + 1) * cout + n);
        }
        split2(x0, x1, bh[h], bl[h]);
    }
}

// KS x KS conv, stride STRIDE, CH 8-channel chunks of input, NT 8-channel tiles of output; POST: fused 1x1 (8*NT -> 8*NT).
template <int KS, int STRIDE, int CH, int NT, bool POST>
__global__ void __launch_bounds__(MM_THREADS) conv_mma_kernel(const ConvParams p, const ConvParams q, int tiles_x, int tiles_y, int n_tiles) {
    constexpr int PH = (MM_TH - 1) * STRIDE + KS, PW = (MM_TW - 1) * STRIDE + KS;
    constexpr int SLOTS = KS * KS * CH, KSTEPS = (SLOTS + 1) / 2;
    constexpr int PSTEPS = (NT + 1) / 2;                       // k-steps of the fused 1x1 (K = 8 * NT)
    constexpr int PLANE = PH * PW * CH * 8;                    // halves per plane
    __shared__ __align__(16) __half s_patch[2][2 * PLANE];      // two tiles: the next tile's patch loads while this one is in the MMAs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pad = KS / 2;

    // ---- weight fragments of the whole layer, once per block
    uint32_t bh[KSTEPS][NT][2], bl[KSTEPS][NT][2];
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) load_bfrag(p.w, p.cin, p.cout, 2 * s, 2 * s + 1, SLOTS, CH, nt, lane, bh[s][nt], bl[s][nt]);
    uint32_t qh[POST ? PSTEPS : 1][NT][2], ql[POST ? PSTEPS : 1][NT][2];
    if (POST) {
#pragma unroll
        for (int s = 0; s < PSTEPS; ++s)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) load_bfrag(q.w, q.cin, q.cout, 2 * s, 2 * s + 1, NT, NT, nt, lane, qh[s][nt], ql[s][nt]);
    }
    // ldmatrix row of this lane: matrix m = lane / 8 (m & 1: pixels 8-15, m >> 1: second slot of the k-step), row r = lane % 8
    const int lm = lane >> 3, lr = lane & 7;
    const int lj = lr + 8 * (lm & 1);                          // output column (inside the 16) this lane addresses
    const int per_img = tiles_x * tiles_y;
    // halo patch of `tile` -> shared memory buffer `buf`: [plane][py][px][chunk][8 halves]; one cp.async group per tile
    auto load_patch = [&](int tile, int buf) {
        const int img = tile / per_img, tr = tile - img * per_img;
        const int oy0 = (tr / tiles_x) * MM_TH, ox0 = (tr % tiles_x) * MM_TW;
        const __half* base = (const __half*)p.in.base + (long long)img * p.in.img + p.in.coff;
        const int iy0 = oy0 * STRIDE - pad, ix0 = ox0 * STRIDE - pad;
        const uint32_t pb = smem_addr(s_patch[buf]);
        for (int e = tid; e < PH * PW * CH; e += MM_THREADS) {
            const int ch = e % CH, pxl = e / CH;
            const int py = pxl / PW, px = pxl - py * PW;
            const int iy = iy0 + py, ix = ix0 + px;
            const bool valid = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
            const __half* src = base + (valid ? (iy * p.W + ix) * p.in.C : 0) + ch * 8;
            const uint32_t dst = pb + (uint32_t)e * 16;
            cp16(dst, src, valid);
            cp16(dst + PLANE * 2, src + p.in.plane, valid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    if ((int)blockIdx.x < n_tiles) load_patch(blockIdx.x, 0);

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int img = tile / per_img, tr = tile - img * per_img;
        const int oy0 = (tr / tiles_x) * MM_TH, ox0 = (tr % tiles_x) * MM_TW;
        const uint32_t patch0 = smem_addr(s_patch[buf]);
        const bool more = tile + (int)gridDim.x < n_tiles;
        if (more) load_patch(tile + gridDim.x, buf ^ 1);       // the other buffer was released by the barrier that ended the previous tile
        // residual operands of this lane (C2f shortcut): issued now, consumed after the MMAs
        uint32_t rres[2][2][NT][2];
        const int cpair = (lane & 3) * 2;
        if (!POST && p.res.base) {
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int oy = min(oy0 + 2 * warp + rt, p.Ho - 1), ox = min(ox0 + (lane >> 2) + 8 * hrow, p.Wo - 1);
                    const __half* rh = (const __half*)p.res.base + (long long)img * p.res.img + (long long)(oy * p.Wo + ox) * p.res.C + p.res.coff;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        rres[rt][hrow][nt][0] = __ldg(reinterpret_cast<const unsigned*>(rh + nt * 8 + cpair));
                        rres[rt][hrow][nt][1] = __ldg(reinterpret_cast<const unsigned*>(rh + p.res.plane + nt * 8 + cpair));
                    }
                }
        }
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- MMAs: this warp's two output rows
        float acc[2][NT][4];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[rt][nt][i] = 0.f;
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int oyl = 2 * warp + rt;
            const uint32_t lane_base = patch0 + (uint32_t)(((oyl * STRIDE) * PW + lj * STRIDE) * CH) * 16;
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
                const int s0 = 2 * s, s1 = (2 * s + 1 < SLOTS) ? 2 * s + 1 : 2 * s;      // a padded slot re-reads real data; its weights are zero
                const int t0 = s0 / CH, c0 = s0 % CH, t1 = s1 / CH, c1 = s1 % CH;
                const uint32_t off0 = (uint32_t)((((t0 / KS) * PW + (t0 % KS)) * CH + c0) * 16);
                const uint32_t off1 = (uint32_t)((((t1 / KS) * PW + (t1 % KS)) * CH + c1) * 16);
                const uint32_t addr = lane_base + ((lm >> 1) ? off1 : off0);
                uint32_t ah[4], al[4];
                ldsm4(addr, ah);
                ldsm4(addr + PLANE * 2, al);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma16816(acc[rt][nt], ah, bh[s][nt][0], bh[s][nt][1]);
                    mma16816(acc[rt][nt], al, bh[s][nt][0], bh[s][nt][1]);
                    mma16816(acc[rt][nt], ah, bl[s][nt][0], bl[s][nt][1]);
                }
            }
        }
        __syncthreads();                                       // this buffer may be overwritten by the prefetch of the tile after next
        // ---- epilogue.  Accumulator fragment: c0,c1 -> pixel lane/4, channels nt*8 + (lane%4)*2 + {0,1}; c2,c3 -> pixel lane/4 + 8
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int oy = oy0 + 2 * warp + rt;
            float v[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float b0 = __ldg(p.bias + nt * 8 + cpair), b1 = __ldg(p.bias + nt * 8 + cpair + 1);
                v[nt][0] = act1(acc[rt][nt][0] + b0, p.act); v[nt][1] = act1(acc[rt][nt][1] + b1, p.act);
                v[nt][2] = act1(acc[rt][nt][2] + b0, p.act); v[nt][3] = act1(acc[rt][nt][3] + b1, p.act);
            }
            const ConvParams& o = POST ? q : p;
            float w[NT][4];
            if (POST) {
                // activated result -> split -> A fragments of the 1x1: a0 = (row, k 0-7) = tile 2s c0,c1; a1 = rows + 8; a2, a3 = tile 2s + 1
                float acc2[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc2[nt][i] = 0.f;
#pragma unroll
                for (int s = 0; s < PSTEPS; ++s) {
                    uint32_t ah[4], al[4];
                    split2(v[2 * s][0], v[2 * s][1], ah[0], al[0]);
                    split2(v[2 * s][2], v[2 * s][3], ah[1], al[1]);
                    if (2 * s + 1 < NT) {
                        split2(v[2 * s + 1 < NT ? 2 * s + 1 : 0][0], v[2 * s + 1 < NT ? 2 * s + 1 : 0][1], ah[2], al[2]);
                        split2(v[2 * s + 1 < NT ? 2 * s + 1 : 0][2], v[2 * s + 1 < NT ? 2 * s + 1 : 0][3], ah[3], al[3]);
                    } else {
                        ah[2] = ah[3] = al[2] = al[3] = 0u;
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        mma16816(acc2[nt], ah, qh[s][nt][0], qh[s][nt][1]);
                        mma16816(acc2[nt], al, qh[s][nt][0], qh[s][nt][1]);
                        mma16816(acc2[nt], ah, ql[s][nt][0], ql[s][nt][1]);
                    }
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float b0 = __ldg(q.bias + nt * 8 + cpair), b1 = __ldg(q.bias + nt * 8 + cpair + 1);
                    w[nt][0] = act1(acc2[nt][0] + b0, q.act); w[nt][1] = act1(acc2[nt][1] + b1, q.act);
                    w[nt][2] = act1(acc2[nt][2] + b0, q.act); w[nt][3] = act1(acc2[nt][3] + b1, q.act);
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[nt][i] = v[nt][i];
            }
            if (oy >= p.Ho) continue;
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int ox = ox0 + (lane >> 2) + 8 * hrow;
                if (ox >= p.Wo) continue;
                const int pix = oy * p.Wo + ox;
                __half* oh = (__half*)o.out.base + (long long)img * o.out.img + (long long)pix * o.out.C + o.out.coff;
                const bool has_res = !POST && p.res.base != nullptr;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    float x0 = w[nt][2 * hrow], x1 = w[nt][2 * hrow + 1];
                    const int c = nt * 8 + cpair;
                    if (has_res) {                              // C2f bottleneck shortcut: added after the activation
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&rres[rt][hrow][nt][0]));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&rres[rt][hrow][nt][1]));
                        x0 += a.x + b.x; x1 += a.y + b.y;
                    }
                    uint32_t hi, lo;
                    split2(x0, x1, hi, lo);
                    *reinterpret_cast<uint32_t*>(oh + c) = hi;
                    *reinterpret_cast<uint32_t*>(oh + o.out.plane + c) = lo;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Stem: 3x3 s2 conv on the u8 letterboxed image (model.ncnn.param:5, /255 of e2e.py:228 fused) on the warp-level tensor cores.
// The fp32-FMA kernel that ran it spent ~940 thread instructions per output pixel for 216 FMAs (ncu: 88 % issue slots at 14 % of
// DRAM throughput, profiles/r2_ncu_stages.txt).  As a GEMM the layer is M = pixels, N = 8 / 16, K = 27 -- what made a first
// MMA version slower was gathering 27 scattered bytes per pixel into fragments.  The gather disappears with the right K order:
// for one output pixel and one filter row ky, the 9 inputs (kx, c) are 9 CONTIGUOUS bytes of the NHWC u8 image, starting at byte
// 6*ox - 3 of input row 2*oy + ky - 1.  K is laid out as three runs of 10: k = 10*ky + j, byte 6*ox - 4 + j of that row, where
// j = 0 (the last channel of the pixel before the window) carries a zero weight; k = 30, 31 are zero too.  Every (k, k+1) pair
// of an A fragment is then one ALIGNED 16-bit shared-memory load, and u8 -> f16 is a byte permute into 0x64xx (= 1024 + x) and
// one HSUB2.  A is exact in f16 (integers 0..255), so the product needs two MMAs per k-step, not three: A x Bhi + A x Blo, with
// B = w * 2^12 / 255 split into hi | lo (the scale keeps the lo halves out of the f16 subnormals) and the fp32 accumulator
// rescaled by 2^-12 before bias and SiLU.  Zero padding = zero bytes (mean 0, std 1 only; other stems use the generic kernel).
constexpr int ST_TH = 8, ST_TW = 64;                   // output tile: one row per warp, four 16-pixel segments per row
constexpr int ST_PR = 2 * ST_TH + 1;                   // patch rows
constexpr int ST_PITCH = 416;                          // bytes per patch row: 16 (alignment lead) + 6 * 64 + 6, rounded up to 16

template <int NT>
__global__ void __launch_bounds__(256) stem_mma_kernel(const ConvParams p, int tiles_x, int tiles_y) {
    __shared__ __align__(16) uint8_t s_patch[ST_PR * ST_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int img = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int oy0 = ty * ST_TH, ox0 = tx * ST_TW;
    const int row_bytes = p.W * 3;
    // ---- patch: rows 2*oy0 - 1 .. 2*oy0 + 15, bytes [6*ox0 - 16, 6*ox0 - 16 + PITCH) of each (16-byte chunks: inside the row or zeros)
    {
        const uint8_t* base = (const uint8_t*)p.in.base + (long long)img * p.in.img;
        const uint32_t pb = smem_addr(s_patch);
        constexpr int CPR = ST_PITCH / 16;
        for (int e = tid; e < ST_PR * CPR; e += 256) {
            const int r = e / CPR, c = e - r * CPR;
            const int iy = 2 * oy0 - 1 + r, off = 6 * ox0 - 16 + 16 * c;
            const bool valid = (unsigned)iy < (unsigned)p.H && off >= 0 && off + 16 <= row_bytes;
            cp16(pb + (uint32_t)(r * ST_PITCH + 16 * c), base + (valid ? (long long)iy * row_bytes + off : 0), valid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- weight fragments (while the patch is in flight): B[k][n], k = 10 * ky + j, j = 1 + 3 * kx + c
    uint32_t bh[2][NT][2], bl[2][NT][2];
    const float wscale = 4096.f / 255.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float x[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int k = 16 * ks + 8 * h + 2 * tq + i, ky = k / 10, j = k - 10 * ky;
                    const int n = nt * 8 + g;
                    x[i] = (k < 30 && j >= 1 && n < p.cout) ? __ldg(p.w + ((ky * 3 + (j - 1) / 3) * 3 + (j - 1) % 3) * p.cout + n) * wscale : 0.f;
                }
                split2(x[0], x[1], bh[ks][nt][h], bl[ks][nt][h]);
            }
    // byte offset of this lane's (k, k+1) pair inside a pixel's window, per (k-step, half); pads re-read pair 0
    int koff[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * ks + 8 * h + 2 * tq, ky = k / 10, j = k - 10 * ky;
            koff[ks][h] = k < 30 ? ky * ST_PITCH + j : 0;
        }
    float bias[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        bias[nt][0] = nt * 8 + tq * 2 < p.cout ? __ldg(p.bias + nt * 8 + tq * 2) : 0.f;
        bias[nt][1] = nt * 8 + tq * 2 + 1 < p.cout ? __ldg(p.bias + nt * 8 + tq * 2 + 1) : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int oy = oy0 + warp;
    if (oy >= p.Ho) return;
    const uint8_t* rowp = s_patch + (2 * warp) * ST_PITCH + 12;       // window of local pixel 0, filter row 0, j = 0
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
    __half* outp = (__half*)p.out.base + (long long)img * p.out.img + (long long)oy * p.Wo * p.out.C + p.out.coff + tq * 2;
#pragma unroll
    for (int sg = 0; sg < ST_TW / 16; ++sg) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
        const uint8_t* px0 = rowp + 6 * (16 * sg + g);                // rows g; rows g + 8 are 48 bytes further
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t a[4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const uint32_t raw = *reinterpret_cast<const uint16_t*>(px0 + koff[ks][h] + 48 * hr);
                    uint32_t v;
                    asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(v) : "r"(raw), "r"(0x64646464u));     // {x0, 0x64, x1, 0x64} = half2(1024 + x0, 1024 + x1)
                    const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&v), k1024);
                    a[2 * h + hr] = *reinterpret_cast<const uint32_t*>(&hv);
                }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                mma16816(acc[nt], a, bh[ks][nt][0], bh[ks][nt][1]);
                mma16816(acc[nt], a, bl[ks][nt][0], bl[ks][nt][1]);
            }
        }
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int ox = ox0 + 16 * sg + g + 8 * hr;
            if (ox >= p.Wo) continue;
            __half* o = outp + (long long)ox * p.out.C;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                if (nt * 8 + tq * 2 >= p.cout) continue;
                const float x0 = act1(fmaf(acc[nt][2 * hr], 1.f / 4096.f, bias[nt][0]), p.act);
                const float x1 = act1(fmaf(acc[nt][2 * hr + 1], 1.f / 4096.f, bias[nt][1]), p.act);
                uint32_t hi, lo;
                split2(x0, x1, hi, lo);
                *reinterpret_cast<uint32_t*>(o + nt * 8) = hi;
                *reinterpret_cast<uint32_t*>(o + p.out.plane + nt * 8) = lo;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Stem + first down-sampling conv + C2f cv1 in ONE kernel (model.ncnn.param:5-8: 3x3 s2 3->8, 3x3 s2 8->16, 1x1 16->16).
// Layer by layer the stem's output (320x320x8 split-f16 = 210 MB per 64 frames) is written and read back once: 420 MB of the
// ~600 MB the three layers move.  Here a block owns an 8 x 16 tile of the SECOND conv's output: it stages the 35 x 67 u8 input
// pixels the tile depends on (cp.async, 16-byte chunks, zero fill = padding; the next tile's bytes load under this tile's math),
// computes the 17 x 33 stem outputs it needs into a shared-memory patch (stem_mma_kernel's arithmetic: contiguous-run K order,
// exact u8 operands; positions outside the 320 x 320 stem map are stored as zeros = the second conv's padding), and then runs
// conv_mma_kernel<3, 2, 1, 2, POST>'s MMA phase on that patch.  The stem is recomputed on the one-pixel halo (x1.1).
constexpr int SC_TH = 8, SC_TW = 16;                          // output tile of the 3x3 s2 conv (rows x cols)
constexpr int SC_SH = 2 * SC_TH + 1, SC_SW = 2 * SC_TW + 1;   // stem outputs the tile reads: 17 x 33
constexpr int SC_SPIX = SC_SH * SC_SW, SC_SEG = (SC_SPIX + 15) / 16, SC_SPAD = SC_SEG * 16;   // 561 pixels = 36 MMA segments
constexpr int SC_IR = 2 * SC_SH + 1;                          // 35 input rows
constexpr int SC_PITCH = 208;                                 // bytes per staged input row: lead 6 + 6 * 33 + 9, rounded up to 16

template <bool POST>
__global__ void __launch_bounds__(MM_THREADS) stem_conv_kernel(const ConvParams ps, const ConvParams p, const ConvParams q,
                                                                int tiles_x, int tiles_y, int n_tiles) {
    constexpr int NT = 2, KSTEPS = 5, PW = SC_SW;
    __shared__ __align__(16) uint8_t s_in[2][SC_IR * SC_PITCH];
    __shared__ __align__(16) __half s_stem[2 * SC_SPAD * 8];     // [plane][stem pixel][8 channels]
    constexpr int PLANE = SC_SPAD * 8;                          // halves per plane
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int per_img = tiles_x * tiles_y;
    const int row_bytes = ps.W * 3;

    // ---- weights of all three layers in registers
    uint32_t sbh[2][2], sbl[2][2];                              // stem: B[k][n], k = 10 * ky + j, j = 1 + 3 * kx + c (stem_mma_kernel)
    const float wscale = 4096.f / 255.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float x[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int k = 16 * ks + 8 * h + 2 * tq + i, ky = k / 10, j = k - 10 * ky;
                x[i] = (k < 30 && j >= 1) ? __ldg(ps.w + ((ky * 3 + (j - 1) / 3) * 3 + (j - 1) % 3) * 8 + g) * wscale : 0.f;
            }
            split2(x[0], x[1], sbh[ks][h], sbl[ks][h]);
        }
    int koff[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * ks + 8 * h + 2 * tq, ky = k / 10, j = k - 10 * ky;
            koff[ks][h] = k < 30 ? ky * SC_PITCH + j : 0;
        }
    const float sbias0 = __ldg(ps.bias + tq * 2), sbias1 = __ldg(ps.bias + tq * 2 + 1);
    uint32_t bh[KSTEPS][NT][2], bl[KSTEPS][NT][2];
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) load_bfrag(p.w, 8, 16, 2 * s, 2 * s + 1, 9, 1, nt, lane, bh[s][nt], bl[s][nt]);
    uint32_t qh[NT][2], ql[NT][2];
    if (POST) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) load_bfrag(q.w, 16, 16, 0, 1, NT, NT, nt, lane, qh[nt], ql[nt]);
    }
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
    const int lm = lane >> 3, lr = lane & 7;
    const int lj = lr + 8 * (lm & 1);

    // staged input of `tile`: rows 4*oy0 - 3 .. +34, bytes [12*ox0 - 16, 12*ox0 - 16 + PITCH) of each
    auto load_in = [&](int tile, int buf) {
        const int img = tile / per_img, tr = tile - img * per_img;
        const int oy0 = (tr / tiles_x) * SC_TH, ox0 = (tr % tiles_x) * SC_TW;
        const uint8_t* base = (const uint8_t*)ps.in.base + (long long)img * ps.in.img;
        const uint32_t pb = smem_addr(s_in[buf]);
        constexpr int CPR = SC_PITCH / 16;
        for (int e = tid; e < SC_IR * CPR; e += MM_THREADS) {
            const int r = e / CPR, c = e - r * CPR;
            const int iy = 4 * oy0 - 3 + r, off = 12 * ox0 - 16 + 16 * c;
            const bool valid = (unsigned)iy < (unsigned)ps.H && off >= 0 && off + 16 <= row_bytes;
            cp16(pb + (uint32_t)(r * SC_PITCH + 16 * c), base + (valid ? (long long)iy * row_bytes + off : 0), valid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int buf = 0;
    if ((int)blockIdx.x < n_tiles) load_in(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int img = tile / per_img, tr = tile - img * per_img;
        const int oy0 = (tr / tiles_x) * SC_TH, ox0 = (tr % tiles_x) * SC_TW;
        const bool more = tile + (int)gridDim.x < n_tiles;
        if (more) load_in(tile + gridDim.x, buf ^ 1);
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                       // input bytes landed; the previous tile's MMA phase is through with s_stem
        // ---- stem on the 17 x 33 region, 16 flat pixels per MMA segment
        for (int seg = warp; seg < SC_SEG; seg += MM_THREADS / 32) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const uint8_t* pxp[2];
            bool inside[2];
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int f = min(seg * 16 + g + 8 * hr, SC_SPIX - 1);
                const int r = f / SC_SW, c = f - r * SC_SW;
                pxp[hr] = s_in[buf] + (2 * r) * SC_PITCH + 6 + 6 * c;            // window of this stem pixel, filter row 0, j = 0
                const int sy = 2 * oy0 - 1 + r, sx = 2 * ox0 - 1 + c;
                inside[hr] = (unsigned)sy < (unsigned)p.H && (unsigned)sx < (unsigned)p.W;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) {
                        const uint32_t raw = *reinterpret_cast<const uint16_t*>(pxp[hr] + koff[ks][h]);
                        uint32_t v;
                        asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(v) : "r"(raw), "r"(0x64646464u));
                        const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&v), k1024);
                        a[2 * h + hr] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                mma16816(acc, a, sbh[ks][0], sbh[ks][1]);
                mma16816(acc, a, sbl[ks][0], sbl[ks][1]);
            }
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int f = seg * 16 + g + 8 * hr;                             // < SC_SPAD: the tail of the last segment lands in padding
                const float x0 = lp_silu(fmaf(acc[2 * hr], 1.f / 4096.f, sbias0));
                const float x1 = lp_silu(fmaf(acc[2 * hr + 1], 1.f / 4096.f, sbias1));
                uint32_t hi, lo;
                split2(x0, x1, hi, lo);
                *reinterpret_cast<uint32_t*>(s_stem + f * 8 + tq * 2) = inside[hr] ? hi : 0u;
                *reinterpret_cast<uint32_t*>(s_stem + PLANE + f * 8 + tq * 2) = inside[hr] ? lo : 0u;
            }
        }
        __syncthreads();
        // ---- 3x3 s2 8 -> 16 on the stem patch (+ the fused 1x1): conv_mma_kernel<3, 2, 1, 2, POST>'s MMA phase
        const uint32_t patch0 = smem_addr(s_stem);
        const int cpair = tq * 2;
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
            const int oyl = 2 * warp + rt;
            const uint32_t lane_base = patch0 + (uint32_t)((oyl * 2) * PW + lj * 2) * 16;
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
                const int t0 = 2 * s, t1 = (2 * s + 1 < 9) ? 2 * s + 1 : 2 * s;
                const uint32_t off0 = (uint32_t)(((t0 / 3) * PW + (t0 % 3)) * 16);
                const uint32_t off1 = (uint32_t)(((t1 / 3) * PW + (t1 % 3)) * 16);
                const uint32_t addr = lane_base + ((lm >> 1) ? off1 : off0);
                uint32_t ah[4], al[4];
                ldsm4(addr, ah);
                ldsm4(addr + PLANE * 2, al);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma16816(acc[nt], ah, bh[s][nt][0], bh[s][nt][1]);
                    mma16816(acc[nt], al, bh[s][nt][0], bh[s][nt][1]);
                    mma16816(acc[nt], ah, bl[s][nt][0], bl[s][nt][1]);
                }
            }
            const int oy = oy0 + oyl;
            float v[NT][4], w[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float b0 = __ldg(p.bias + nt * 8 + cpair), b1 = __ldg(p.bias + nt * 8 + cpair + 1);
                v[nt][0] = lp_silu(acc[nt][0] + b0); v[nt][1] = lp_silu(acc[nt][1] + b1);
                v[nt][2] = lp_silu(acc[nt][2] + b0); v[nt][3] = lp_silu(acc[nt][3] + b1);
            }
            if (POST) {
                float acc2[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc2[nt][i] = 0.f;
                uint32_t ah[4], al[4];
                split2(v[0][0], v[0][1], ah[0], al[0]);
                split2(v[0][2], v[0][3], ah[1], al[1]);
                split2(v[1][0], v[1][1], ah[2], al[2]);
                split2(v[1][2], v[1][3], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma16816(acc2[nt], ah, qh[nt][0], qh[nt][1]);
                    mma16816(acc2[nt], al, qh[nt][0], qh[nt][1]);
                    mma16816(acc2[nt], ah, ql[nt][0], ql[nt][1]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float b0 = __ldg(q.bias + nt * 8 + cpair), b1 = __ldg(q.bias + nt * 8 + cpair + 1);
                    w[nt][0] = lp_silu(acc2[nt][0] + b0); w[nt][1] = lp_silu(acc2[nt][1] + b1);
                    w[nt][2] = lp_silu(acc2[nt][2] + b0); w[nt][3] = lp_silu(acc2[nt][3] + b1);
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[nt][i] = v[nt][i];
            }
            const ConvParams& o = POST ? q : p;
            if (oy >= p.Ho) continue;
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int ox = ox0 + g + 8 * hrow;
                if (ox >= p.Wo) continue;
                __half* oh = (__half*)o.out.base + (long long)img * o.out.img + ((long long)oy * p.Wo + ox) * o.out.C + o.out.coff;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    uint32_t hi, lo;
                    split2(w[nt][2 * hrow], w[nt][2 * hrow + 1], hi, lo);
                    *reinterpret_cast<uint32_t*>(oh + nt * 8 + cpair) = hi;
                    *reinterpret_cast<uint32_t*>(oh + o.out.plane + nt * 8 + cpair) = lo;
                }
            }
        }
        // the next iteration's first barrier separates this MMA phase from the next stem phase
    }
}

template <int KS, int STRIDE, int CH, int NT, bool POST>
int launch_mma(lp_ctx* ctx, const ConvParams& p, const ConvParams* post, cudaStream_t st) {
    const int tiles_x = (p.Wo + MM_TW - 1) / MM_TW, tiles_y = (p.Ho + MM_TH - 1) / MM_TH;
    const long long n_tiles = (long long)tiles_x * tiles_y * p.n_img;
    if (n_tiles <= 0 || n_tiles > 0x7fffffff) return 0;
    // weights live in registers: a block amortises their load over several tiles; smem is small, so many blocks per SM hide latency
    long long grid = (long long)ctx->sm_count * 8;
    if (grid > n_tiles) grid = n_tiles;
    const ConvParams q = post ? *post : p;
    conv_mma_kernel<KS, STRIDE, CH, NT, POST><<<(unsigned)grid, MM_THREADS, 0, st>>>(p, q, tiles_x, tiles_y, (int)n_tiles);
    return 1;
}

}  // namespace

// Shapes covered: split-f16 in and out, channel views on 8-channel boundaries, cout == 8 * NT exactly.
int lp_conv_mma_try(lp_ctx* ctx, const ConvParams& p, const ConvParams* post, cudaStream_t st) {
    if (p.in.fmt != LP_FMT_SPLIT16 || p.out.fmt != LP_FMT_SPLIT16 || p.in.coff % 8 || p.out.coff % 8 || p.cin % 8 || p.cout % 8) return 0;
    if (p.seg_len != 0 || p.out_cstride != 1 || p.res_first) return 0;
    if (p.res.base && (p.res.fmt != LP_FMT_SPLIT16 || p.res.coff % 8 || post)) return 0;
    if (post && (post->cin != p.cout || post->cout != p.cout || post->out.fmt != LP_FMT_SPLIT16 || post->out.coff % 8 || post->res.base ||
                 post->seg_len != 0 || post->out_cstride != 1))
        return 0;
    const int ch = p.cin / 8, nt = p.cout / 8;
#define LP_MMA_CASE(K, S, C, N)                                                                   \
    if (p.ksize == K && p.stride == S && ch == C && nt == N)                                       \
        return post ? launch_mma<K, S, C, N, true>(ctx, p, post, st) : launch_mma<K, S, C, N, false>(ctx, p, nullptr, st);
    LP_MMA_CASE(3, 2, 1, 2)      // v1 model.1   3x3 s2 8 -> 16 (+ model.2.cv1 1x1 16 -> 16)
    LP_MMA_CASE(3, 1, 1, 1)      // v1 model.2.m 3x3 8 -> 8
    LP_MMA_CASE(1, 1, 3, 2)      // v1 model.2.cv2 1x1 24 -> 16
    LP_MMA_CASE(1, 1, 2, 2)      // v1 model.2.cv1 alone (per-op probing)
    LP_MMA_CASE(1, 1, 5, 3)      // v2 model.2.cv2 1x1 36(40) -> 24
    LP_MMA_CASE(1, 1, 3, 3)      // v2 model.2.cv1 alone
#undef LP_MMA_CASE
    return 0;
}

// u8 stem (3x3 s2, 3 -> 8 / 16 channels, x / 255, zero mean / unit std) on the warp-level tensor cores; 0 if the shape is not covered
int lp_stem_mma_try(lp_ctx* ctx, const ConvParams& p, cudaStream_t st) {
    if (p.in.fmt != LP_FMT_U8 || p.out.fmt != LP_FMT_SPLIT16 || p.ksize != 3 || p.stride != 2 || p.cin != 3) return 0;
    if (p.cout != 8 && p.cout != 16) return 0;
    if (p.in_scale_mean != 0.f || p.in_scale_std != 1.f || p.res.base || p.seg_len != 0 || p.out_cstride != 1 || p.out.coff % 8) return 0;
    // 16-byte chunks of an image row: the rows and the images must start on 16-byte boundaries
    if ((p.W * 3) % 16 || (p.in.img % 16) || ((uintptr_t)p.in.base % 16) || p.W % 2 || p.H % 2) return 0;
    const int tiles_x = (p.Wo + ST_TW - 1) / ST_TW, tiles_y = (p.Ho + ST_TH - 1) / ST_TH;
    if (p.n_img > 65535 || tiles_y > 65535) return 0;
    dim3 grid(tiles_x, tiles_y, p.n_img);
    if (p.cout == 8) stem_mma_kernel<1><<<grid, 256, 0, st>>>(p, tiles_x, tiles_y);
    else stem_mma_kernel<2><<<grid, 256, 0, st>>>(p, tiles_x, tiles_y);
    (void)ctx;
    return 1;
}

// u8 stem (3 -> 8) + 3x3 s2 conv (8 -> 16) [+ 1x1 16 -> 16] in one kernel; the stem's output tensor is never materialised.
// ps = the stem, p = the conv that reads ONLY the stem's output, post = its fused 1x1 or nullptr.  0: shapes not covered.
int lp_stem_conv_try(lp_ctx* ctx, const ConvParams& ps, const ConvParams& p, const ConvParams* post, cudaStream_t st) {
    if (ps.in.fmt != LP_FMT_U8 || ps.ksize != 3 || ps.stride != 2 || ps.cin != 3 || ps.cout != 8) return 0;
    if (ps.act != LP_ACT_SILU || p.act != LP_ACT_SILU || (post && post->act != LP_ACT_SILU)) return 0;       // the activation is compiled in
    if (ps.in_scale_mean != 0.f || ps.in_scale_std != 1.f || ps.res.base || ps.seg_len != 0 || ps.out_cstride != 1) return 0;
    if ((ps.W * 3) % 16 || (ps.in.img % 16) || ((uintptr_t)ps.in.base % 16) || ps.W % 4 || ps.H % 4) return 0;
    if (p.ksize != 3 || p.stride != 2 || p.cin != 8 || p.cout != 16 || p.H != ps.Ho || p.W != ps.Wo || p.res.base || p.res_first ||
        p.seg_len != 0 || p.out_cstride != 1 || p.out.fmt != LP_FMT_SPLIT16 || p.out.coff % 8)
        return 0;
    if (post && (post->cin != 16 || post->cout != 16 || post->out.fmt != LP_FMT_SPLIT16 || post->out.coff % 8 || post->res.base ||
                 post->seg_len != 0 || post->out_cstride != 1))
        return 0;
    const int tiles_x = (p.Wo + SC_TW - 1) / SC_TW, tiles_y = (p.Ho + SC_TH - 1) / SC_TH;
    const long long n_tiles = (long long)tiles_x * tiles_y * p.n_img;
    if (n_tiles <= 0 || n_tiles > 0x7fffffff) return 0;
    long long grid = (long long)ctx->sm_count * 4;
    if (grid > n_tiles) grid = n_tiles;
    const ConvParams qq = post ? *post : p;
    if (post) stem_conv_kernel<true><<<(unsigned)grid, MM_THREADS, 0, st>>>(ps, p, qq, tiles_x, tiles_y, (int)n_tiles);
    else stem_conv_kernel<false><<<(unsigned)grid, MM_THREADS, 0, st>>>(ps, p, qq, tiles_x, tiles_y, (int)n_tiles);
    return 1;
}
