// Fused C2f body: the bottleneck chain and cv2 of one C2f block in ONE kernel, intermediates in shared memory.
//
// Ultralytics C2f (model.ncnn.param:8-18 for model.2):  y0|y1 = cv1(x);  y_{i+2} = y_{i+1} + m_i.cv2(m_i.cv1(y_{i+1}))
// (both 3x3 c -> c, SiLU);  out = cv2(cat(y0 .. y_{n+1}))  (1x1, SiLU).  Layer by layer that is 2n + 1 launches whose
// activations each make an HBM round trip, through a concat buffer that every 3x3 reads as a strided channel slice
// (16 of every 48 bytes at the 160x160 level: ncu 157 MB read for a 52 MB tensor, profiles/r2_notes.md item 5), and at
// c = 8 / 16 the GEMMs (N = 8 / 16) are far too narrow for a tcgen05 tile.  Here one CTA owns a 16 x 16 output tile:
//   * it loads y0|y1 for the tile plus a halo of 2n pixels once (cp.async, zero fill outside the image),
//   * runs the 2n 3x3 convs on the warp-level tensor cores (mma.sync m16n8k16, split-f16: Ahi*Bhi + Alo*Bhi + Ahi*Blo in
//     fp32 -- the arithmetic of conv_tc.cu / conv_mma.cu), each result re-split to hi|lo and kept in a shared-memory frame
//     (positions outside the image are stored as zeros: they are the next conv's padding),
//   * and applies cv2 to the centre: the chunks y0 .. y_n come from shared memory, y_{n+1} straight from the accumulator
//     registers of the last 3x3 (the C fragment of m16n8k16 is the A fragment of the next MMA).
// The halo is recomputed per tile (2n rings); nothing but y0|y1 is read and nothing but the block's output is written.
//
// Frames are FLAT: pixel f = row * P + col of a P-column frame, every conv output f reads inputs f + (ky-1)*P + (kx-1).
// A 16-pixel MMA segment is any 16 consecutive f, whatever rows it spans; the columns that wrap around a row end compute
// garbage that only ever reaches the outer rings (ring k is dead after conv k), never the centre.  A-fragments are
// ldmatrix rows of [plane][8-channel chunk][pixel][16 B], so im2col is an address offset, as in conv_mma.cu.
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int CF_TW = 16;                   // output tile width = one MMA segment per centre row

struct C2fParams {
    TensorRef y;                 // cv1 output: channels [coff, coff + 2c) = y0 | y1
    TensorRef out;               // cv2 destination
    int H, W, n_img;
    int tiles_x, tiles_y, n_tiles;
    int shortcut;                // bottleneck residuals (model.ncnn.param BinaryOp add after each m.i.cv2)
    const float* w[5];           // [0, 2n): the 3x3 convs in execution order; [2n]: cv2.  fp32 [tap][cin][cout]
    const float* b[5];
    int cout2;
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 join_pair(uint32_t hi, uint32_t lo) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2(a.x + b.x, a.y + b.y);
}

__device__ __forceinline__ float silu(float v) { return lp_silu(v); }

// One 3x3 conv (8*C8 -> 8*C8 channels) on MS segments of 16 flat pixels starting at f0[m].  in_hi = shared-memory address of
// chunk 0 of the conv's input frame (hi plane; lo plane `plane_stride` further), wf = the conv's weight fragments
// [k-step][n-tile][lane] {bh0, bh1, bl0, bl1}, bs = its bias (the accumulators start from it).
template <int C8, int P, int SL, int CHUNK, int MS>
__device__ __forceinline__ void conv3_multi(uint32_t in_hi, uint32_t plane_stride, const uint4* __restrict__ wf, const float* __restrict__ bs,
                                            const int (&f0)[MS], int lane, float (&acc)[MS][C8][4]) {
    constexpr int NT = C8, SLOTS = 9 * C8, KS3 = (SLOTS + 1) / 2;
    // ldmatrix lane roles: matrix m = lane / 8 (m & 1: pixels 8-15 of the segment, m >> 1: second slot / chunk of the k-step)
    const int lm = lane >> 3, lj = (lane & 7) + 8 * (lm & 1), tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float b0 = bs[nt * 8 + tq * 2], b1 = bs[nt * 8 + tq * 2 + 1];
#pragma unroll
        for (int m = 0; m < MS; ++m) { acc[m][nt][0] = b0; acc[m][nt][1] = b1; acc[m][nt][2] = b0; acc[m][nt][3] = b1; }
    }
    uint32_t lane_base[MS];
#pragma unroll
    for (int m = 0; m < MS; ++m) lane_base[m] = in_hi + (uint32_t)(SL + f0[m] + lj) * 16;
#pragma unroll
    for (int s = 0; s < KS3; ++s) {
        const int s0 = 2 * s, s1 = (2 * s + 1 < SLOTS) ? 2 * s + 1 : 2 * s;       // a padded slot re-reads real data; its weights are zero
        const int t0 = s0 / C8, c0 = s0 % C8, t1 = s1 / C8, c1 = s1 % C8;
        const int off0 = c0 * CHUNK + ((t0 / 3 - 1) * P + (t0 % 3 - 1)) * 16;
        const int off1 = c1 * CHUNK + ((t1 / 3 - 1) * P + (t1 % 3 - 1)) * 16;
        const uint32_t off = (uint32_t)((lm >> 1) ? off1 : off0);
        uint32_t ah[MS][4], al[MS][4];
#pragma unroll
        for (int m = 0; m < MS; ++m) {
            ldsm_x4(lane_base[m] + off, ah[m]);
            ldsm_x4(lane_base[m] + off + plane_stride, al[m]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint4 b = wf[(s * NT + nt) * 32 + lane];
#pragma unroll
            for (int m = 0; m < MS; ++m) {
                hmma(acc[m][nt], ah[m], b.x, b.y);
                hmma(acc[m][nt], al[m], b.x, b.y);
                hmma(acc[m][nt], ah[m], b.z, b.w);
            }
        }
    }
}

template <int C8, int NB, int TH>
struct C2fGeom {
    static constexpr int HALO = 2 * NB, P = CF_TW + 2 * HALO, R = TH + 2 * HALO, F = R * P;
    static constexpr int SL = P + 1;                         // slack pixels in front of and behind a frame (taps of the first / last row)
    static constexpr int FP = F + 2 * SL + 16;               // + one segment: the last segment of a stage may overrun its row range
    static constexpr int CHUNK = FP * 16;                    // bytes of one 8-channel chunk plane
    static constexpr int NT = C8, SLOTS = 9 * C8, KS3 = (SLOTS + 1) / 2;
    static constexpr int K2C = (2 + NB) * C8, KS2 = (K2C + 1) / 2;
};

template <int C8, int NB, int TH, int NT2>
constexpr size_t c2f_smem_bytes() {
    using G = C2fGeom<C8, NB, TH>;
    size_t w = (size_t)(2 * NB * G::KS3 * G::NT + G::KS2 * NT2) * 32 * 16;
    size_t b = (size_t)(2 * NB * 8 * C8 + 8 * NT2) * 4;
    size_t frames = (size_t)2 * G::CHUNK * (2 * C8 + C8 + (NB == 2 ? C8 : 0));
    return w + ((b + 15) / 16) * 16 + frames;
}

// C8 = c / 8, NB bottlenecks, TH tile rows, NT2 = cv2 outputs / 8, NW warps
template <int C8, int NB, int TH, int NT2, int NW>
__global__ void __launch_bounds__(NW * 32) c2f_fused_kernel(const C2fParams p) {
    using G = C2fGeom<C8, NB, TH>;
    constexpr int HALO = G::HALO, P = G::P, R = G::R, F = G::F, SL = G::SL, CHUNK = G::CHUNK;
    constexpr int NT = G::NT, SLOTS = G::SLOTS, KS3 = G::KS3, K2C = G::K2C, KS2 = G::KS2;
    constexpr int NTHR = NW * 32;
    extern __shared__ __align__(16) uint8_t sm[];
    uint4* wf3 = reinterpret_cast<uint4*>(sm);                               // [2NB][KS3][NT][32] {bh0, bh1, bl0, bl1}
    uint4* wf2 = wf3 + 2 * NB * KS3 * NT * 32;                               // [KS2][NT2][32]
    float* bias = reinterpret_cast<float*>(wf2 + KS2 * NT2 * 32);            // [2NB][8*C8] | [8*NT2]
    constexpr int BIAS_BYTES = (((2 * NB * 8 * C8 + 8 * NT2) * 4 + 15) / 16) * 16;
    uint8_t* fr_y = reinterpret_cast<uint8_t*>(bias) + BIAS_BYTES;           // [2 planes][2*C8 chunks][FP][16 B]
    uint8_t* fr_t = fr_y + (size_t)2 * (2 * C8) * CHUNK;                     // [2][C8][FP][16 B]   m.i.cv1 output
    uint8_t* fr_2 = fr_t + (size_t)2 * C8 * CHUNK;                           // [2][C8][FP][16 B]   y2 (NB == 2)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;

    // ---- once per CTA: weight fragments (split to hi | lo) and biases; zero the frames (slack and never-written rings stay finite)
    for (int cv = 0; cv < 2 * NB; ++cv) {
        const float* __restrict__ w = p.w[cv];
        for (int idx = tid; idx < KS3 * NT * 32; idx += NTHR) {
            const int ln = idx & 31, nt = (idx >> 5) % NT, s = (idx >> 5) / NT;
            const int n = nt * 8 + (ln >> 2), kc = (ln & 3) * 2;
            uint32_t bh[2], bl[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = 2 * s + h;
                float x0 = 0.f, x1 = 0.f;
                if (slot < SLOTS) {
                    const int tap = slot / C8, c = (slot - tap * C8) * 8 + kc;
                    x0 = __ldg(w + ((long long)tap * (8 * C8) + c) * (8 * C8) + n);
                    x1 = __ldg(w + ((long long)tap * (8 * C8) + c + 1) * (8 * C8) + n);
                }
                split_pair(x0, x1, bh[h], bl[h]);
            }
            wf3[cv * KS3 * NT * 32 + idx] = make_uint4(bh[0], bh[1], bl[0], bl[1]);
        }
        for (int i = tid; i < 8 * C8; i += NTHR) bias[cv * 8 * C8 + i] = __ldg(p.b[cv] + i);
    }
    {
        const float* __restrict__ w = p.w[2 * NB];
        for (int idx = tid; idx < KS2 * NT2 * 32; idx += NTHR) {
            const int ln = idx & 31, nt = (idx >> 5) % NT2, s = (idx >> 5) / NT2;
            const int n = nt * 8 + (ln >> 2), kc = (ln & 3) * 2;
            uint32_t bh[2], bl[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ci = 2 * s + h;
                float x0 = 0.f, x1 = 0.f;
                if (ci < K2C && n < p.cout2) {
                    x0 = __ldg(w + (long long)(ci * 8 + kc) * p.cout2 + n);
                    x1 = __ldg(w + (long long)(ci * 8 + kc + 1) * p.cout2 + n);
                }
                split_pair(x0, x1, bh[h], bl[h]);
            }
            wf2[idx] = make_uint4(bh[0], bh[1], bl[0], bl[1]);
        }
        for (int i = tid; i < 8 * NT2; i += NTHR) bias[2 * NB * 8 * C8 + i] = i < p.cout2 ? __ldg(p.b[2 * NB] + i) : 0.f;
    }
    {
        constexpr int N16 = (2 * CHUNK * (3 * C8 + (NB == 2 ? C8 : 0))) / 16;
        uint4* z = reinterpret_cast<uint4*>(fr_y);
        for (int i = tid; i < N16; i += NTHR) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    const int per_img = p.tiles_x * p.tiles_y;

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int img = tile / per_img, tr = tile - img * per_img;
        const int ty0 = (tr / p.tiles_x) * TH, tx0 = (tr % p.tiles_x) * CF_TW;
        // ---- y0 | y1 frame (tile + halo), zero outside the image
        {
            const __half* base = (const __half*)p.y.base + (long long)img * p.y.img + p.y.coff;
            const uint32_t dst0 = s_addr(fr_y);
            for (int e = tid; e < F * 2 * C8; e += NTHR) {
                const int ch = e % (2 * C8), f = e / (2 * C8);
                const int ry = f / P, rx = f - ry * P;
                const int gy = ty0 - HALO + ry, gx = tx0 - HALO + rx;
                const bool valid = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
                const __half* src = base + (valid ? ((long long)gy * p.W + gx) * p.y.C : 0) + ch * 8;
                const uint32_t dst = dst0 + (uint32_t)ch * CHUNK + (uint32_t)(SL + f) * 16;
                cp16z(dst, src, valid);
                cp16z(dst + 2 * C8 * CHUNK, src + p.y.plane, valid);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();

        // ---- the 3x3 convs but the last: stage k writes rows [k, R - k) of its output frame
#pragma unroll
        for (int k = 1; k < 2 * NB; ++k) {
            // input / output / residual frames of stage k (1-based): odd k = m.cv1 (y_{(k+1)/2} -> t), even k = m.cv2 (t -> y_{k/2+1})
            const bool is_cv1 = (k & 1) != 0;
            const uint32_t in_hi = is_cv1 ? (k == 1 ? s_addr(fr_y) + C8 * CHUNK : s_addr(fr_2)) : s_addr(fr_t);
            const uint32_t in_pl = is_cv1 ? (k == 1 ? 2 * C8 * CHUNK : C8 * CHUNK) : C8 * CHUNK;
            uint8_t* out_f = is_cv1 ? fr_t : fr_2;                               // an even k < 2NB only exists for NB == 2: y2
            const uint8_t* res_f = fr_y + C8 * CHUNK;                            // residual of stage 2: y1
            constexpr int res_pl = 2 * C8 * CHUNK;
            const uint4* wf = wf3 + (k - 1) * KS3 * NT * 32;
            const float* bs = bias + (k - 1) * 8 * C8;
            const int f_begin = k * P, n_seg = ((R - 2 * k) * P + 15) / 16;
            auto stage_item = [&](auto ms_c, int seg0) {
                constexpr int MS = decltype(ms_c)::value;
                int f0[MS];
#pragma unroll
                for (int m = 0; m < MS; ++m) f0[m] = f_begin + (seg0 + m) * 16;
                float acc[MS][NT][4];
                conv3_multi<C8, P, SL, CHUNK, MS>(in_hi, in_pl, wf, bs, f0, lane, acc);
#pragma unroll
                for (int m = 0; m < MS; ++m)
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) {
                        const int f = f0[m] + g + 8 * hr;
                        const int ry = f / P, rx = f - ry * P;
                        const int gy = ty0 - HALO + ry, gx = tx0 - HALO + rx;
                        const bool inside = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            float x0 = silu(acc[m][nt][2 * hr]), x1 = silu(acc[m][nt][2 * hr + 1]);
                            const uint32_t o = (uint32_t)nt * CHUNK + (uint32_t)(SL + f) * 16 + tq * 4;
                            if (!is_cv1 && p.shortcut) {
                                const float2 r = join_pair(*reinterpret_cast<const uint32_t*>(res_f + o), *reinterpret_cast<const uint32_t*>(res_f + res_pl + o));
                                x0 += r.x; x1 += r.y;
                            }
                            uint32_t hi, lo;
                            split_pair(x0, x1, hi, lo);
                            *reinterpret_cast<uint32_t*>(out_f + o) = inside ? hi : 0u;
                            *reinterpret_cast<uint32_t*>(out_f + C8 * CHUNK + o) = inside ? lo : 0u;
                        }
                    }
            };
            // every warp takes a contiguous share of the stage's segments (sizes differ by at most one), two at a time
            int sg = (warp * n_seg) / NW;
            const int sg_end = ((warp + 1) * n_seg) / NW;
            for (; sg + 1 < sg_end; sg += 2) stage_item(std::integral_constant<int, 2>{}, sg);
            if (sg < sg_end) stage_item(std::integral_constant<int, 1>{}, sg);
            __syncthreads();
        }

        // ---- centre rows: last 3x3 (t -> y_{NB+1}, kept in registers) and cv2 over cat(y0 .. y_{NB+1})
        {
            const uint4* wf = wf3 + (2 * NB - 1) * KS3 * NT * 32;
            const float* bs = bias + (2 * NB - 1) * 8 * C8;
            const float* b2 = bias + 2 * NB * 8 * C8;
            const uint8_t* res_f = NB == 1 ? fr_y + C8 * CHUNK : fr_2;
            constexpr int res_pl = NB == 1 ? 2 * C8 * CHUNK : C8 * CHUNK;
            const int lm = lane >> 3, lj = (lane & 7) + 8 * (lm & 1);
            constexpr int MS = (TH / NW >= 2) ? 2 : 1;      // centre rows per work item: every warp busy, two MMA chains when rows allow
            static_assert(TH % MS == 0, "tile rows per work item");
            for (int r0 = warp * MS; r0 < TH; r0 += NW * MS) {
                int f0[MS];
#pragma unroll
                for (int m = 0; m < MS; ++m) f0[m] = (HALO + r0 + m) * P + HALO;
                float acc[MS][NT][4];
                conv3_multi<C8, P, SL, CHUNK, MS>(s_addr(fr_t), C8 * CHUNK, wf, bs, f0, lane, acc);
                // y_last as A fragments: chunk nt -> {a0 = rows g (k 2tq..), a1 = rows g + 8}
                uint32_t yh[MS][NT][2], yl[MS][NT][2];
#pragma unroll
                for (int m = 0; m < MS; ++m)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int hr = 0; hr < 2; ++hr) {
                            float x0 = silu(acc[m][nt][2 * hr]), x1 = silu(acc[m][nt][2 * hr + 1]);
                            if (p.shortcut) {
                                const uint32_t o = (uint32_t)nt * CHUNK + (uint32_t)(SL + f0[m] + g + 8 * hr) * 16 + tq * 4;
                                const float2 rr = join_pair(*reinterpret_cast<const uint32_t*>(res_f + o), *reinterpret_cast<const uint32_t*>(res_f + res_pl + o));
                                x0 += rr.x; x1 += rr.y;
                            }
                            split_pair(x0, x1, yh[m][nt][hr], yl[m][nt][hr]);
                        }
                float acc2[MS][NT2][4];
#pragma unroll
                for (int m = 0; m < MS; ++m)
#pragma unroll
                    for (int nt = 0; nt < NT2; ++nt) {
                        const float b0 = b2[nt * 8 + tq * 2], b1 = b2[nt * 8 + tq * 2 + 1];
                        acc2[m][nt][0] = b0; acc2[m][nt][1] = b1; acc2[m][nt][2] = b0; acc2[m][nt][3] = b1;
                    }
                // smem address of chunk ci (< K2C - C8) of the concat, hi plane, and its plane stride
                auto chunk_addr = [&](int ci, uint32_t& pl) -> uint32_t {
                    if (ci < 2 * C8) { pl = 2 * C8 * CHUNK; return s_addr(fr_y) + (uint32_t)ci * CHUNK; }
                    pl = C8 * CHUNK; return s_addr(fr_2) + (uint32_t)(ci - 2 * C8) * CHUNK;
                };
#pragma unroll
                for (int s = 0; s < KS2; ++s) {
                    constexpr int NSM = K2C - C8;                // chunks that live in shared memory
                    const int ca = 2 * s, cb = 2 * s + 1;
                    uint32_t ah[MS][4], al[MS][4];
#pragma unroll
                    for (int m = 0; m < MS; ++m) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) ah[m][i] = al[m][i] = 0u;
                        if (cb < NSM) {                                // both halves from shared memory: one ldmatrix.x4 per plane
                            uint32_t pla, plb;
                            const uint32_t aa = chunk_addr(ca, pla), ab = chunk_addr(cb, plb);
                            const uint32_t addr = ((lm >> 1) ? ab : aa) + (uint32_t)(SL + f0[m] + lj) * 16;
                            ldsm_x4(addr, ah[m]);
                            ldsm_x4(addr + ((lm >> 1) ? plb : pla), al[m]);
                        } else {
                            if (ca < NSM) {
                                uint32_t pla;
                                const uint32_t aa = chunk_addr(ca, pla) + (uint32_t)(SL + f0[m] + lj) * 16;   // lanes 16-31: addresses ignored by .x2
                                ldsm_x2(aa, ah[m][0], ah[m][1]);
                                ldsm_x2(aa + pla, al[m][0], al[m][1]);
                            } else if (ca < K2C) {
                                ah[m][0] = yh[m][ca - NSM][0]; ah[m][1] = yh[m][ca - NSM][1]; al[m][0] = yl[m][ca - NSM][0]; al[m][1] = yl[m][ca - NSM][1];
                            }
                            if (cb >= NSM && cb < K2C) {
                                ah[m][2] = yh[m][cb - NSM][0]; ah[m][3] = yh[m][cb - NSM][1]; al[m][2] = yl[m][cb - NSM][0]; al[m][3] = yl[m][cb - NSM][1];
                            }
                        }
                    }
#pragma unroll
                    for (int nt = 0; nt < NT2; ++nt) {
                        const uint4 b = wf2[(s * NT2 + nt) * 32 + lane];
#pragma unroll
                        for (int m = 0; m < MS; ++m) {
                            hmma(acc2[m][nt], ah[m], b.x, b.y);
                            hmma(acc2[m][nt], al[m], b.x, b.y);
                            hmma(acc2[m][nt], ah[m], b.z, b.w);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < MS; ++m) {
                    const int gy = ty0 + r0 + m;
                    if (gy >= p.H) continue;
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) {
                        const int gx = tx0 + g + 8 * hr;
                        if (gx >= p.W) continue;
                        __half* oh = (__half*)p.out.base + (long long)img * p.out.img + ((long long)gy * p.W + gx) * p.out.C + p.out.coff + tq * 2;
#pragma unroll
                        for (int nt = 0; nt < NT2; ++nt) {
                            uint32_t hi, lo;
                            split_pair(silu(acc2[m][nt][2 * hr]), silu(acc2[m][nt][2 * hr + 1]), hi, lo);
                            *reinterpret_cast<uint32_t*>(oh + nt * 8) = hi;
                            *reinterpret_cast<uint32_t*>(oh + p.out.plane + nt * 8) = lo;
                        }
                    }
                }
            }
        }
        __syncthreads();          // the frames are rewritten by the next tile
    }
}

template <int C8, int NB, int TH, int NT2, int NW>
int launch_c2f(lp_ctx* ctx, C2fParams& p, int bit, cudaStream_t st) {
    constexpr size_t smem = c2f_smem_bytes<C8, NB, TH, NT2>();
    static_assert(smem <= 227 * 1024, "c2f frame does not fit shared memory");
    auto kern = c2f_fused_kernel<C8, NB, TH, NT2, NW>;
    if (!(ctx->attr_set & bit)) {
        LP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_set |= bit;
    }
    p.tiles_x = (p.W + CF_TW - 1) / CF_TW;
    p.tiles_y = (p.H + TH - 1) / TH;
    const long long n_tiles = (long long)p.tiles_x * p.tiles_y * p.n_img;
    if (n_tiles <= 0 || n_tiles > 0x7fffffff) return 0;
    p.n_tiles = (int)n_tiles;
    static int per_sm = 0;              // resident CTAs per SM of this instantiation (same on every B200)
    if (per_sm == 0) {
        LP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
        if (per_sm < 1) per_sm = 1;
    }
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, NW * 32, smem, st>>>(p);
    return 1;
}

}  // namespace

// Pattern: ops[oi .. oi + 2n] = n x (3x3 c->c into a scratch buffer, 3x3 c->c back into the concat buffer [+ residual]) and the
// 1x1 over the whole concat buffer.  Returns the number of ops the fused kernel covered (2n + 1), 0 if the pattern or the
// shape is not covered, < 0 on error.
int lp_c2f_fused_try(lp_ctx* ctx, lp_net_plan& net, size_t oi, int batch, uint8_t* ws, cudaStream_t st) {
    const std::vector<lp_op_desc>& ops = net.ops;
    const lp_op_desc& a = ops[oi];
    if (a.kind != LP_OP_CONV || a.ksize != 3 || a.stride != 1 || a.cin != a.cout || a.res_buf >= 0 || a.flags || a.out_seg_len) return 0;
    if (a.act != LP_ACT_SILU) return 0;          // the kernel's activation is compiled in
    const int c = a.cin;
    if (c != 8 && c != 16) return 0;
    const int cat = a.in_buf;
    const lp_buf_desc& cb = net.bufs[cat];
    if (cb.fmt != LP_FMT_SPLIT16 || a.in_coff < c || (a.in_coff - c) % 8) return 0;
    const int coff0 = a.in_coff - c;                         // y0 starts here, y1 = the first bottleneck's input
    int nb = 0;
    size_t k = oi;
    bool shortcut = false;
    while (nb < 2 && k + 1 < ops.size()) {
        const lp_op_desc& m1 = ops[k];
        const lp_op_desc& m2 = ops[k + 1];
        const bool ok1 = m1.kind == LP_OP_CONV && m1.ksize == 3 && m1.stride == 1 && m1.cin == c && m1.cout == c && m1.res_buf < 0 && !m1.flags &&
                         !m1.out_seg_len && m1.out_cstride <= 1 && m1.in_buf == cat && m1.in_coff == coff0 + (1 + nb) * c && m1.out_buf != cat &&
                         m1.out_coff == 0 && net.bufs[m1.out_buf].c == c && net.bufs[m1.out_buf].fmt == LP_FMT_SPLIT16 && m1.act == a.act;
        if (!ok1) break;
        const bool has_res = m2.res_buf >= 0;
        const bool ok2 = m2.kind == LP_OP_CONV && m2.ksize == 3 && m2.stride == 1 && m2.cin == c && m2.cout == c && !m2.flags && !m2.out_seg_len &&
                         m2.out_cstride <= 1 && m2.in_buf == m1.out_buf && m2.in_coff == 0 && m2.out_buf == cat &&
                         m2.out_coff == coff0 + (2 + nb) * c && m2.act == a.act &&
                         (!has_res || (m2.res_buf == cat && m2.res_coff == coff0 + (1 + nb) * c)) && (nb == 0 || has_res == shortcut);
        if (!ok2) break;
        // the scratch buffer has no other reader or writer
        for (size_t j = 0; j < ops.size(); ++j)
            if (j != k && j != k + 1 && (ops[j].in_buf == m1.out_buf || ops[j].out_buf == m1.out_buf || ops[j].res_buf == m1.out_buf)) return 0;
        shortcut = has_res;
        ++nb;
        k += 2;
    }
    if (nb == 0 || k >= ops.size()) return 0;
    const lp_op_desc& o = ops[k];
    if (!(o.kind == LP_OP_CONV && o.ksize == 1 && o.stride == 1 && o.in_buf == cat && o.in_coff == coff0 && o.cin == (2 + nb) * c && o.res_buf < 0 && o.act == LP_ACT_SILU &&
          !o.flags && !o.out_seg_len && o.out_cstride <= 1 && o.out_coff % 8 == 0 && o.cout % 8 == 0 && o.cout <= 32 &&
          (o.cout_real <= 0 || o.cout_real == o.cout) && net.bufs[o.out_buf].fmt == LP_FMT_SPLIT16 && o.out_buf != cat))
        return 0;
    // channels [coff0 + 2c, ...) of the concat buffer are written and read by this group only
    for (size_t j = 0; j < ops.size(); ++j) {
        if (j >= oi && j <= k) continue;
        const lp_op_desc& x = ops[j];
        if (x.res_buf == cat) return 0;
        if (x.in_buf == cat) return 0;
        if (x.out_buf == cat && x.out_coff + x.cout > coff0 + 2 * c) return 0;
    }
    const lp_buf_desc& ob = net.bufs[o.out_buf];
    if (ob.h != cb.h || ob.w != cb.w) return 0;

    C2fParams p{};
    const int esz = 2;
    p.y.base = ws + cb.offset; p.y.img = cb.image_bytes / esz; p.y.plane = (long long)net.max_batch * p.y.img; p.y.C = cb.c; p.y.coff = coff0; p.y.fmt = cb.fmt;
    p.out.base = ws + ob.offset + (size_t)o.row_off * ob.c * esz; p.out.img = ob.image_bytes / esz; p.out.plane = (long long)net.max_batch * p.out.img;
    p.out.C = ob.c; p.out.coff = o.out_coff; p.out.fmt = ob.fmt;
    p.H = cb.h; p.W = cb.w; p.n_img = batch;
    p.shortcut = shortcut ? 1 : 0;
    for (int i = 0; i < 2 * nb + 1; ++i) { p.w[i] = net.weights + ops[oi + i].w_off; p.b[i] = net.weights + ops[oi + i].b_off; }
    p.cout2 = o.cout;
    const int nt2 = o.cout / 8;
    int r = 0;
#define LP_C2F_CASE(C8_, NB_, NT2_, NW_, BIT_)                                              \
    if (c == 8 * C8_ && nb == NB_ && nt2 == NT2_) r = launch_c2f<C8_, NB_, 16, NT2_, NW_>(ctx, p, BIT_, st);
    LP_C2F_CASE(1, 1, 2, 8, 1 << 8)        // v1 model.2  (160x160): c = 8,  n = 1, cv2 24 -> 16
    else LP_C2F_CASE(2, 2, 4, 16, 1 << 9)  // v1 model.4  (80x80):   c = 16, n = 2, cv2 64 -> 32
    else LP_C2F_CASE(2, 1, 4, 8, 1 << 10)  // v1 model.15 (80x80):   c = 16, n = 1, cv2 48 -> 32
    else LP_C2F_CASE(2, 1, 3, 8, 1 << 11)  // v2 model.2  (160x160): c = 16, n = 1, cv2 48 -> 24
#undef LP_C2F_CASE
    if (r <= 0) return r;
    return 2 * nb + 1;
}
