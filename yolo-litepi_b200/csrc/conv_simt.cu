// CUDA-core (SIMT) layer kernels: the generic fp32-FMA convolution family plus the memory-bound
// glue ops.  They implement every layer of the detector (model.ncnn.param:4-183) and of
// ShuffleNetV2 (torchvision shufflenetv2.py) and are the reference implementation the tcgen05
// kernels in conv_tc.cu are checked against on the GPU.  Layers whose GEMM shape cannot feed the
// tensor core (Cin = 3 or 8, N = 1, depthwise) stay here permanently: they are HBM-bound.
//
// Tensors are NHWC.  SPLIT16 buffers hold two fp16 planes (hi | lo), value = hi + lo.
#include "common.cuh"

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == LP_ACT_SILU) return lp_silu(v);
    if (act == LP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == LP_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (act == LP_ACT_SIGMOID) return lp_sigmoid(v);
    return v;
}

// physical channel (relative to the pixel) of output j
__device__ __forceinline__ long long out_chan(const ConvParams& p, int j) {
    if (p.seg_len > 0) {
        const int l = p.seg_l0 + j * p.out_cstride;
        return (long long)(l / p.seg_len) * p.seg_pad + l % p.seg_len;
    }
    return (long long)j * p.out_cstride;
}

__device__ __forceinline__ float ld_elem(const TensorRef& t, long long idx) {
    if (t.fmt == LP_FMT_SPLIT16) {
        const __half* hi = (const __half*)t.base;
        return __half2float(hi[idx]) + __half2float(hi[idx + t.plane]);
    }
    return ((const float*)t.base)[idx];
}

__device__ __forceinline__ void st_elem(const TensorRef& t, long long idx, float v) {
    if (t.fmt == LP_FMT_SPLIT16) {
        __half* hi = (__half*)t.base;
        __half h, l;
        split_make(v, h, l);
        hi[idx] = h;
        hi[idx + t.plane] = l;
    } else {
        ((float*)t.base)[idx] = v;
    }
}

// Load 8 consecutive channels (16-B aligned for SPLIT16) as floats.
__device__ __forceinline__ void ld8(const TensorRef& t, long long idx, float* v) {
    if (t.fmt == LP_FMT_SPLIT16) {
        const __half* hi = (const __half*)t.base;
        uint4 a = *reinterpret_cast<const uint4*>(hi + idx);
        uint4 b = *reinterpret_cast<const uint4*>(hi + idx + t.plane);
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 x = __half22float2(ah[i]), y = __half22float2(bh[i]);
            v[2 * i] = x.x + y.x;
            v[2 * i + 1] = x.y + y.y;
        }
    } else {
        const float* p = (const float*)t.base + idx;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = p[i];
    }
}

__device__ __forceinline__ void st8(const TensorRef& t, long long idx, const float* v) {
    if (t.fmt == LP_FMT_SPLIT16) {
        __half* hi = (__half*)t.base;
        uint4 a, b;
        __half2* ah = reinterpret_cast<__half2*>(&a);
        __half2* bh = reinterpret_cast<__half2*>(&b);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half h0, l0, h1, l1;
            split_make(v[2 * i], h0, l0);
            split_make(v[2 * i + 1], h1, l1);
            ah[i] = __halves2half2(h0, h1);
            bh[i] = __halves2half2(l0, l1);
        }
        *reinterpret_cast<uint4*>(hi + idx) = a;
        *reinterpret_cast<uint4*>(hi + idx + t.plane) = b;
    } else {
        float* p = (float*)t.base + idx;
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Generic conv: one thread = one output pixel x COB output channels; the block stages an input
// patch [CK][PH][PW] (fp32) and a weight slab [taps][CK][COB] in shared memory per 8-channel chunk.
// KS==1: the block covers 128 consecutive pixels of the flattened (image, y, x) index space.
// KS==3: the block covers a 8x16 output tile of one image.
// ---------------------------------------------------------------------------------------------
constexpr int CONV_THREADS = 128;
constexpr int TILE_W = 16, TILE_H = 8, CK = 8;

template <int KS, int STRIDE, int COB>
__global__ void __launch_bounds__(CONV_THREADS) conv_simt_kernel(ConvParams p) {
    constexpr int PH = (KS == 1) ? 1 : (TILE_H - 1) * STRIDE + KS;
    constexpr int PW = (KS == 1) ? CONV_THREADS : (TILE_W - 1) * STRIDE + KS;
    constexpr int TAPS = KS * KS;
    __shared__ float s_patch[CK][PH][PW + 1];
    __shared__ __align__(16) float s_w[TAPS][CK][COB];

    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * COB;
    int img, oy, ox;            // this thread's output pixel
    int ty = 0, tx = tid;       // position inside the tile
    long long pix0 = 0;         // KS==1: first flattened pixel of the block
    int tile_y0 = 0, tile_x0 = 0;
    bool valid;
    if (KS == 1) {
        pix0 = (long long)blockIdx.x * CONV_THREADS;
        long long m = pix0 + tid;
        long long total = (long long)p.n_img * p.Ho * p.Wo;
        valid = m < total;
        long long mm = valid ? m : 0;
        img = (int)(mm / ((long long)p.Ho * p.Wo));
        int r = (int)(mm - (long long)img * p.Ho * p.Wo);
        oy = r / p.Wo;
        ox = r - oy * p.Wo;
    } else {
        const int tiles_x = (p.Wo + TILE_W - 1) / TILE_W;
        img = blockIdx.z;
        tile_y0 = (blockIdx.x / tiles_x) * TILE_H;
        tile_x0 = (blockIdx.x % tiles_x) * TILE_W;
        ty = tid / TILE_W;
        tx = tid % TILE_W;
        oy = tile_y0 + ty;
        ox = tile_x0 + tx;
        valid = (oy < p.Ho && ox < p.Wo);
    }

    float acc[COB];
#pragma unroll
    for (int i = 0; i < COB; ++i) acc[i] = 0.f;

    const int pad = KS / 2;
    for (int c0 = 0; c0 < p.cin; c0 += CK) {
        const int ck = min(CK, p.cin - c0);
        // ---- stage input patch
        if (KS == 1) {
            // flattened OUTPUT pixels; input pixel = output pixel * STRIDE (1x1 stride 2 = ResNet's downsample branch)
            float v[CK];
#pragma unroll
            for (int i = 0; i < CK; ++i) v[i] = 0.f;
            if (valid) {
                long long idx = (long long)img * p.in.img + ((long long)(oy * STRIDE) * p.W + ox * STRIDE) * p.in.C + p.in.coff + c0;
                if (ck == CK && p.in.fmt == LP_FMT_SPLIT16) ld8(p.in, idx, v);
                else for (int i = 0; i < ck; ++i) v[i] = ld_elem(p.in, idx + i);
            }
#pragma unroll
            for (int i = 0; i < CK; ++i) s_patch[i][0][tid] = v[i];
        } else {
            const int iy0 = tile_y0 * STRIDE - pad, ix0 = tile_x0 * STRIDE - pad;
            for (int e = tid; e < PH * PW; e += CONV_THREADS) {
                const int py = e / PW, px = e - py * PW;
                const int iy = iy0 + py, ix = ix0 + px;
                float v[CK];
#pragma unroll
                for (int i = 0; i < CK; ++i) v[i] = 0.f;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                    long long idx = (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff + c0;
                    if (ck == CK && p.in.fmt == LP_FMT_SPLIT16) ld8(p.in, idx, v);
                    else for (int i = 0; i < ck; ++i) v[i] = ld_elem(p.in, idx + i);
                }
#pragma unroll
                for (int i = 0; i < CK; ++i) s_patch[i][py][px] = v[i];
            }
        }
        // ---- stage weights [tap][ci][co]
        for (int e = tid; e < TAPS * CK * COB; e += CONV_THREADS) {
            const int co = e % COB, ci = (e / COB) % CK, t = e / (COB * CK);
            float wv = 0.f;
            if (ci < ck && co0 + co < p.cout) wv = __ldg(p.w + ((long long)t * p.cin + c0 + ci) * p.cout + co0 + co);
            s_w[t][ci][co] = wv;
        }
        __syncthreads();
        // ---- FMA
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            const int ky = t / KS, kx = t % KS;
#pragma unroll
            for (int ci = 0; ci < CK; ++ci) {
                const float a = (KS == 1) ? s_patch[ci][0][tid] : s_patch[ci][ty * STRIDE + ky][tx * STRIDE + kx];
                const float4* wr = reinterpret_cast<const float4*>(&s_w[t][ci][0]);
#pragma unroll
                for (int q = 0; q < COB / 4; ++q) {
                    const float4 w4 = wr[q];
                    acc[4 * q + 0] = fmaf(a, w4.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(a, w4.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(a, w4.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(a, w4.w, acc[4 * q + 3]);
                }
            }
        }
        __syncthreads();
    }
    if (!valid) return;
    // ---- epilogue: bias, activation, residual, store
    const long long opix = (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff;
    const long long rpix = p.res.base ? (long long)img * p.res.img + ((long long)oy * p.Wo + ox) * p.res.C + p.res.coff : 0;
#pragma unroll
    for (int g = 0; g < COB / 8; ++g) {
        const int cbase = co0 + g * 8;
        if (cbase >= p.cout) break;
        float v[8];
        const int n = min(8, p.cout - cbase);
        float r[8];
        if (p.res.base) {
            if (n == 8 && p.res.fmt == LP_FMT_SPLIT16) ld8(p.res, rpix + cbase, r);
            else for (int i = 0; i < n; ++i) r[i] = ld_elem(p.res, rpix + cbase + i);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = acc[g * 8 + i] + (i < n ? __ldg(p.bias + cbase + i) : 0.f);
            if (p.res.base && i < n && p.res_first) x += r[i];
            x = act_apply(x, p.act);
            if (p.res.base && i < n && !p.res_first) x += r[i];
            v[i] = x;
        }
        if (n == 8 && p.out_cstride == 1 && p.seg_len == 0 && (p.out.fmt == LP_FMT_SPLIT16)) st8(p.out, opix + cbase, v);
        else for (int i = 0; i < n; ++i) if (cbase + i < p.cout_real) st_elem(p.out, opix + out_chan(p, cbase + i), v[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// Stem: 3x3 stride-2 conv straight from the u8 RGB image (letterbox / ROI-resize output).
// x = (u8/255 - mean)/std with IEEE divisions exactly as ToTensor/Normalize (e2e.py:368-369)
// and the ORT twin's x/255 (evaluation_tsd_single_img.ipynb:98-110) compute it.
// ---------------------------------------------------------------------------------------------
template <int COB>
__global__ void __launch_bounds__(CONV_THREADS) stem_u8_kernel(ConvParams p) {
    constexpr int PH = (TILE_H - 1) * 2 + 3, PW = (TILE_W - 1) * 2 + 3;
    __shared__ float s_patch[3][PH][PW + 1];
    __shared__ __align__(16) float s_w[9][3][COB];
    const int tid = threadIdx.x;
    const int tiles_x = (p.Wo + TILE_W - 1) / TILE_W;
    const int img = blockIdx.z;
    const int tile_y0 = (blockIdx.x / tiles_x) * TILE_H, tile_x0 = (blockIdx.x % tiles_x) * TILE_W;
    const int ty = tid / TILE_W, tx = tid % TILE_W;
    const int oy = tile_y0 + ty, ox = tile_x0 + tx;
    const int co0 = blockIdx.y * COB;
    const uint8_t* src = (const uint8_t*)p.in.base + (long long)img * p.in.img;
    const int iy0 = tile_y0 * 2 - 1, ix0 = tile_x0 * 2 - 1;
    for (int e = tid; e < PH * PW; e += CONV_THREADS) {
        const int py = e / PW, px = e - py * PW;
        const int iy = iy0 + py, ix = ix0 + px;
        float v[3] = {0.f, 0.f, 0.f};
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            const uint8_t* q = src + ((long long)iy * p.W + ix) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = __fdiv_rn((float)q[c], 255.f);
                if (p.in_scale_std != 1.f || p.in_scale_mean != 0.f)
                    x = __fdiv_rn(__fsub_rn(x, p.in_scale_mean), p.in_scale_std);
                v[c] = x;
            }
        }
        s_patch[0][py][px] = v[0]; s_patch[1][py][px] = v[1]; s_patch[2][py][px] = v[2];
    }
    for (int e = tid; e < 9 * 3 * COB; e += CONV_THREADS) {
        const int co = e % COB, ci = (e / COB) % 3, t = e / (COB * 3);
        s_w[t][ci][co] = (co0 + co < p.cout) ? __ldg(p.w + ((long long)t * 3 + ci) * p.cout + co0 + co) : 0.f;
    }
    __syncthreads();
    if (oy >= p.Ho || ox >= p.Wo) return;
    float acc[COB];
#pragma unroll
    for (int i = 0; i < COB; ++i) acc[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float a = s_patch[ci][ty * 2 + t / 3][tx * 2 + t % 3];
#pragma unroll
            for (int co = 0; co < COB; ++co) acc[co] = fmaf(a, s_w[t][ci][co], acc[co]);
        }
    }
    const long long opix = (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff;
#pragma unroll
    for (int g = 0; g < COB / 8; ++g) {
        const int cbase = co0 + g * 8;
        if (cbase >= p.cout) break;
        const int n = min(8, p.cout - cbase);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_apply(acc[g * 8 + i] + (i < n ? __ldg(p.bias + cbase + i) : 0.f), p.act);
        if (n == 8 && p.out.fmt == LP_FMT_SPLIT16) st8(p.out, opix + cbase, v);
        else for (int i = 0; i < n; ++i) st_elem(p.out, opix + cbase + i, v[i]);
    }
}

// Any odd kernel size / stride from the u8 image (ResNet18's 7x7 stride-2 conv1, torchvision resnet.py): thread =
// output pixel x 8 output channels, taps straight from global memory (the image is L1/L2 resident; 9.6 MMAC per ROI).
__global__ void __launch_bounds__(128) stem_u8_generic_kernel(ConvParams p) {
    const int groups = (p.cout + 7) / 8;
    const long long total = (long long)p.n_img * p.Ho * p.Wo * groups;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    const int pad = p.ksize / 2, c0 = g * 8, n = min(8, p.cout - c0);
    const uint8_t* src = (const uint8_t*)p.in.base + (long long)img * p.in.img;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int ky = 0; ky < p.ksize; ++ky) {
        const int iy = oy * p.stride - pad + ky;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.ksize; ++kx) {
            const int ix = ox * p.stride - pad + kx;
            if (ix < 0 || ix >= p.W) continue;
            const uint8_t* q = src + ((long long)iy * p.W + ix) * 3;
            const float* w = p.w + (long long)((ky * p.ksize + kx) * 3) * p.cout + c0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = __fdiv_rn((float)q[c], 255.f);
                if (p.in_scale_std != 1.f || p.in_scale_mean != 0.f) x = __fdiv_rn(__fsub_rn(x, p.in_scale_mean), p.in_scale_std);
                for (int k = 0; k < n; ++k) acc[k] = fmaf(x, __ldg(w + (long long)c * p.cout + k), acc[k]);
            }
        }
    }
    const long long opix = (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = act_apply(acc[k] + (k < n ? __ldg(p.bias + c0 + k) : 0.f), p.act);
    if (n == 8 && p.out.fmt == LP_FMT_SPLIT16) st8(p.out, opix + c0, v);
    else for (int k = 0; k < n; ++k) st_elem(p.out, opix + c0 + k, v[k]);
}

// ---------------------------------------------------------------------------------------------
// Small-channel convs (Cin 3 / 8 / 24: the first five detector layers).  They cannot feed a tensor core
// and the generic kernel is bound by its shared-memory weight reads (one LDS per 4 FMAs), so here the
// weights live in CONSTANT memory: an FFMA takes its weight operand straight from the constant bank, the
// only shared-memory traffic is one input read per Cout FMAs.  One thread = one output pixel x all Cout.
// ---------------------------------------------------------------------------------------------
// Weights travel BY VALUE as a kernel parameter: every FFMA then names its weight as a constant-bank operand with a
// static offset.  (A __constant__ array indexed by a runtime slot cost one LDC per weight -- ncu: 1350 instructions per
// pixel for 576 FMAs.)  Kernel parameters may be up to 32 KB since CUDA 12.1; the largest set here is 5.8 KB.
constexpr int SMALL_W_FLOATS = 2048;
template <int N> struct SmallW { float v[N]; };      // [tap][cin][cout] then bias[cout]

template <int KS, int STRIDE, int CIN, int COUT, bool U8IN>
__global__ void __launch_bounds__(CONV_THREADS) conv_small_kernel(const ConvParams p, const SmallW<KS * KS * CIN * COUT + COUT> wt,
                                                                  const SmallW<COUT * COUT + COUT> pw, int post_on, int post_act) {
    constexpr int NWT = KS * KS * CIN * COUT;
    constexpr int PH = (KS == 1) ? 1 : (TILE_H - 1) * STRIDE + KS;
    constexpr int PW = (KS == 1) ? CONV_THREADS : (TILE_W - 1) * STRIDE + KS;
    constexpr int CINP = U8IN ? CIN : ((CIN + 7) / 8) * 8;
    __shared__ float s_patch[CINP][PH][PW + 1];
    __shared__ float s_lut[U8IN ? 256 : 1];             // u8 -> normalised float, the reference's own operations (bit-identical)
    const int tid = threadIdx.x;
    if (U8IN) {
        for (int i = tid; i < 256; i += CONV_THREADS) {
            float x = __fdiv_rn((float)i, 255.f);
            if (p.in_scale_std != 1.f || p.in_scale_mean != 0.f) x = __fdiv_rn(__fsub_rn(x, p.in_scale_mean), p.in_scale_std);
            s_lut[i] = x;
        }
        __syncthreads();
    }
    int img, oy, ox, ty = 0, tx = tid, tile_y0 = 0, tile_x0 = 0;
    bool valid;
    if (KS == 1) {
        const long long m = (long long)blockIdx.x * CONV_THREADS + tid;
        const long long total = (long long)p.n_img * p.Ho * p.Wo;
        valid = m < total;
        const long long mm = valid ? m : 0;
        img = (int)(mm / ((long long)p.Ho * p.Wo));
        const int r = (int)(mm - (long long)img * p.Ho * p.Wo);
        oy = r / p.Wo; ox = r - oy * p.Wo;
        float v[CINP];
#pragma unroll
        for (int i = 0; i < CINP; ++i) v[i] = 0.f;
        if (valid) {
            const long long idx = (long long)img * p.in.img + ((long long)oy * p.W + ox) * p.in.C + p.in.coff;
#pragma unroll
            for (int c8 = 0; c8 < CINP; c8 += 8) ld8(p.in, idx + c8, v + c8);
        }
#pragma unroll
        for (int i = 0; i < CIN; ++i) s_patch[i][0][tid] = v[i];
    } else {
        const int tiles_x = (p.Wo + TILE_W - 1) / TILE_W;
        img = blockIdx.z;
        tile_y0 = (blockIdx.x / tiles_x) * TILE_H; tile_x0 = (blockIdx.x % tiles_x) * TILE_W;
        ty = tid / TILE_W; tx = tid % TILE_W;
        oy = tile_y0 + ty; ox = tile_x0 + tx;
        valid = (oy < p.Ho && ox < p.Wo);
        const int iy0 = tile_y0 * STRIDE - KS / 2, ix0 = tile_x0 * STRIDE - KS / 2;
        for (int e = tid; e < PH * PW; e += CONV_THREADS) {
            const int py = e / PW, px = e - py * PW;
            const int iy = iy0 + py, ix = ix0 + px;
            const bool in = (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W);
            if (U8IN) {
                float v[3] = {0.f, 0.f, 0.f};
                if (in) {
                    const uint8_t* q = (const uint8_t*)p.in.base + (long long)img * p.in.img + ((long long)iy * p.W + ix) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = s_lut[q[c]];
                }
                s_patch[0][py][px] = v[0]; s_patch[1][py][px] = v[1]; s_patch[2][py][px] = v[2];
            } else {
                float v[CINP];
#pragma unroll
                for (int i = 0; i < CINP; ++i) v[i] = 0.f;
                if (in) {
                    const long long idx = (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff;
#pragma unroll
                    for (int c8 = 0; c8 < CINP; c8 += 8) ld8(p.in, idx + c8, v + c8);
                }
#pragma unroll
                for (int i = 0; i < CIN; ++i) s_patch[i][py][px] = v[i];
            }
        }
    }
    __syncthreads();
    if (!valid) return;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = wt.v[NWT + co];
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float a = (KS == 1) ? s_patch[ci][0][tid] : s_patch[ci][ty * STRIDE + t / KS][tx * STRIDE + t % KS];
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] = fmaf(a, wt.v[(t * CIN + ci) * COUT + co], acc[co]);
        }
    }
    const long long opix = (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff;
    const long long rpix = (long long)img * p.res.img + ((long long)oy * p.Wo + ox) * p.res.C + p.res.coff;
    if (post_on) {
        // fused COUT -> COUT 1x1 conv on this pixel's activated outputs (the C2f cv1 that follows the down-sampling
        // conv): the intermediate tensor is never written.  Weights [cin][cout] + bias in the second parameter block.
        float y[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) y[co] = pw.v[COUT * COUT + co];
#pragma unroll
        for (int ci = 0; ci < COUT; ++ci) {
            const float a = act_apply(acc[ci], p.act);
#pragma unroll
            for (int co = 0; co < COUT; ++co) y[co] = fmaf(a, pw.v[ci * COUT + co], y[co]);
        }
#pragma unroll
        for (int g = 0; g < COUT / 8; ++g) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = act_apply(y[g * 8 + i], post_act);
            st8(p.out, opix + g * 8, v);
        }
        return;
    }
#pragma unroll
    for (int g = 0; g < COUT / 8; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_apply(acc[g * 8 + i], p.act);
        if (p.res.base) {                              // C2f bottleneck shortcut: added after the activation
            float r[8];
            ld8(p.res, rpix + g * 8, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i];
        }
        st8(p.out, opix + g * 8, v);
    }
}

// Calls f.template run<K,S,CI,CO,U8>() for the specialisation of this shape; false if there is none
// (v1 widths 8/16, v2 widths 16/24 stem).
template <typename F> static bool small_dispatch(int ks, int stride, int cin, int cout, bool u8, F&& f) {
#define LP_SMALL(K, S, CI, CO, U) if (ks == K && stride == S && cin == CI && cout == CO && u8 == U) { f.template run<K, S, CI, CO, U>(); return true; }
    LP_SMALL(3, 2, 3, 8, true)
    LP_SMALL(3, 2, 3, 16, true)
    LP_SMALL(3, 2, 8, 16, false)
    LP_SMALL(3, 1, 8, 8, false)
    LP_SMALL(1, 1, 24, 16, false)
    LP_SMALL(1, 1, 16, 16, false)
#undef LP_SMALL
    return false;
}
struct SmallProbe { template <int K, int S, int CI, int CO, bool U> void run() {} };
struct SmallLaunch {
    const ConvParams* p; const float* w_host; const float* post_host; int post_act; int batch; cudaStream_t st;
    template <int K, int S, int CI, int CO, bool U> void run() {
        SmallW<K * K * CI * CO + CO> w;
        memcpy(w.v, w_host, sizeof(w.v));
        SmallW<CO * CO + CO> pw;
        if (post_host) memcpy(pw.v, post_host, sizeof(pw.v)); else memset(pw.v, 0, sizeof(pw.v));
        dim3 grid;
        if (K == 1) grid = dim3((unsigned)(((long long)batch * p->Ho * p->Wo + CONV_THREADS - 1) / CONV_THREADS), 1, 1);
        else grid = dim3(((p->Wo + TILE_W - 1) / TILE_W) * ((p->Ho + TILE_H - 1) / TILE_H), 1, batch);
        conv_small_kernel<K, S, CI, CO, U><<<grid, CONV_THREADS, 0, st>>>(*p, w, pw, post_host ? 1 : 0, post_act);
    }
};

// ---------------------------------------------------------------------------------------------
// Memory-bound glue: depthwise 3x3, max pool, nearest x2 upsample, channel-slice copy, mean+FC.
// One thread per (pixel, channel); channels fastest so warps read/write contiguous NHWC runs.
// ---------------------------------------------------------------------------------------------
__global__ void dwconv3_kernel(ConvParams p) {
    const long long total = (long long)p.n_img * p.Ho * p.Wo * p.cout;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.cout);
    long long r = i / p.cout;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    float acc = __ldg(p.bias + c);
    const int K = p.ksize, pad = K / 2;                     // 3 (ShuffleNetV2, MobileNetV2) or 5 (EfficientNet)
    for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * p.stride - pad + ky;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < K; ++kx) {
            const int ix = ox * p.stride - pad + kx;
            if (ix < 0 || ix >= p.W) continue;
            const float a = ld_elem(p.in, (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff + c);
            acc = fmaf(a, __ldg(p.w + (ky * K + kx) * p.cout + c), acc);
        }
    }
    acc = act_apply(acc, p.act);
    st_elem(p.out, (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff + out_chan(p, c), acc);
}

// 1x1 conv with a handful of outputs (the Detect head's class-logit convs: cout = nc, model.ncnn.param:160,171,182).  The
// generic kernel stages a 32-wide weight slab and a patch per 8 input channels for one useful column; here a thread owns a
// pixel, reads its cin channels with 16-byte loads and keeps the <= 4 accumulators in registers: the layer runs at the
// speed of its input read.  Same fp32 operation order as the generic kernel (channels ascending, bias last).
__global__ void __launch_bounds__(256) conv1x1_few_kernel(ConvParams p) {
    extern __shared__ float s_wf[];                          // [cin][cout] then bias[cout]
    for (int i = threadIdx.x; i < p.cin * p.cout; i += blockDim.x) s_wf[i] = __ldg(p.w + i);
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) s_wf[p.cin * p.cout + i] = __ldg(p.bias + i);
    __syncthreads();
    const long long total = (long long)p.n_img * p.Ho * p.Wo;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= total) return;
    const int hw = p.Ho * p.Wo;
    const int img = (int)(m / hw), q = (int)(m - (long long)img * hw);
    const long long ipix = (long long)img * p.in.img + (long long)q * p.in.C + p.in.coff;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < p.cin; c0 += 8) {
        float v[8];
        ld8(p.in, ipix + c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < p.cout) acc[j] = fmaf(v[i], s_wf[(c0 + i) * p.cout + j], acc[j]);
    }
    const long long opix = (long long)img * p.out.img + (long long)q * p.out.C + p.out.coff;
    for (int j = 0; j < p.cout; ++j)
        if (j < p.cout_real) st_elem(p.out, opix + out_chan(p, j), act_apply(acc[j] + s_wf[p.cin * p.cout + j], p.act));
}

// SqueezeExcitation pieces (torchvision ops/misc.py): mean over H x W per channel, and x * gate[c]
__global__ void global_mean_kernel(ConvParams p) {
    const long long total = (long long)p.n_img * p.cout;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.cout), img = (int)(i / p.cout);
    float s = 0.f;
    for (int q = 0; q < p.H * p.W; ++q) s += ld_elem(p.in, (long long)img * p.in.img + (long long)q * p.in.C + p.in.coff + c);
    st_elem(p.out, (long long)img * p.out.img + p.out.coff + c, s / (float)(p.H * p.W));
}

__global__ void scale_kernel(ConvParams p) {
    const long long total = (long long)p.n_img * p.Ho * p.Wo * p.cout;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.cout);
    const long long pix = i / p.cout;
    const int img = (int)(pix / ((long long)p.Ho * p.Wo));
    const long long q = pix - (long long)img * p.Ho * p.Wo;
    const float g = ld_elem(p.res, (long long)img * p.res.img + p.res.coff + c);
    const float x = ld_elem(p.in, (long long)img * p.in.img + q * p.in.C + p.in.coff + c);
    st_elem(p.out, (long long)img * p.out.img + q * p.out.C + p.out.coff + c, x * g);
}

__global__ void maxpool_kernel(ConvParams p) {
    const long long total = (long long)p.n_img * p.Ho * p.Wo * p.cout;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.cout);
    long long r = i / p.cout;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    const int pad = p.ksize / 2;
    float m = -INFINITY;
    for (int ky = 0; ky < p.ksize; ++ky) {
        const int iy = oy * p.stride - pad + ky;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.ksize; ++kx) {
            const int ix = ox * p.stride - pad + kx;
            if (ix < 0 || ix >= p.W) continue;
            m = fmaxf(m, ld_elem(p.in, (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff + c));
        }
    }
    st_elem(p.out, (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff + c, m);
}

// upsample x2 (stride field = 2) or plain copy (stride = 1); out pixel (oy,ox) <- in (oy/s, ox/s)
__global__ void resample_copy_kernel(ConvParams p) {
    const long long total = (long long)p.n_img * p.Ho * p.Wo * p.cout;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.cout);
    long long r = i / p.cout;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    const int iy = oy / p.stride, ix = ox / p.stride;
    const long long src = (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff + c;
    const long long dst = (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff + out_chan(p, c);
    if (p.in.fmt == LP_FMT_SPLIT16 && p.out.fmt == LP_FMT_SPLIT16) {   // bit-preserving
        const __half* s = (const __half*)p.in.base;
        __half* d = (__half*)p.out.base;
        d[dst] = s[src];
        d[dst + p.out.plane] = s[src + p.in.plane];
    } else {
        st_elem(p.out, dst, ld_elem(p.in, src));
    }
}

// Vectorised split-f16 variants: one thread = one 16-byte chunk (8 channels) of one plane of one output
// pixel; both the nearest-x2 upsample (stride 2) and the copy (stride 1) are bit-preserving.
__global__ void resample_copy8_kernel(ConvParams p) {
    // 32-bit index arithmetic (the host checks the item count fits): 64-bit divisions were most of this kernel
    const unsigned cpp = (unsigned)p.cout >> 3;                       // chunks per pixel
    const unsigned total = 2u * (unsigned)p.n_img * (unsigned)p.Ho * (unsigned)p.Wo * cpp;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned ch = i % cpp;
    unsigned r = i / cpp;
    const unsigned ox = r % (unsigned)p.Wo; r /= (unsigned)p.Wo;
    const unsigned oy = r % (unsigned)p.Ho; r /= (unsigned)p.Ho;
    const unsigned img = r % (unsigned)p.n_img;
    const unsigned plane = r / (unsigned)p.n_img;
    const unsigned iy = oy / (unsigned)p.stride, ix = ox / (unsigned)p.stride;
    const __half* s = (const __half*)p.in.base + (long long)plane * p.in.plane + (long long)img * p.in.img +
                      ((long long)iy * p.W + ix) * p.in.C + p.in.coff + ch * 8;
    __half* d = (__half*)p.out.base + (long long)plane * p.out.plane + (long long)img * p.out.img +
                ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff + ch * 8;
    *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(s));
}

// Nearest x2 upsample, one thread per SOURCE chunk (16 bytes = 8 channels of one plane of one input pixel): one load, four
// stores, and the index arithmetic (five divisions by runtime divisors: most of the per-output-chunk kernel above, which ran
// at 2 TB/s) is paid once per four output chunks.  Consecutive threads cover the consecutive chunks of a pixel row.
__global__ void upsample2x8_kernel(ConvParams p) {
    const unsigned cpp = (unsigned)p.cout >> 3;
    const unsigned total = 2u * (unsigned)p.n_img * (unsigned)p.H * (unsigned)p.W * cpp;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned ch = i % cpp;
    unsigned r = i / cpp;
    const unsigned ix = r % (unsigned)p.W; r /= (unsigned)p.W;
    const unsigned iy = r % (unsigned)p.H; r /= (unsigned)p.H;
    const unsigned img = r % (unsigned)p.n_img;
    const unsigned plane = r / (unsigned)p.n_img;
    const __half* s = (const __half*)p.in.base + (long long)plane * p.in.plane + (long long)img * p.in.img +
                      ((long long)iy * p.W + ix) * p.in.C + p.in.coff + ch * 8;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(s));
    __half* d = (__half*)p.out.base + (long long)plane * p.out.plane + (long long)img * p.out.img +
                ((long long)(2 * iy) * p.Wo + 2 * ix) * p.out.C + p.out.coff + ch * 8;
    const long long row = (long long)p.Wo * p.out.C;
    *reinterpret_cast<uint4*>(d) = v;
    *reinterpret_cast<uint4*>(d + p.out.C) = v;
    *reinterpret_cast<uint4*>(d + row) = v;
    *reinterpret_cast<uint4*>(d + row + p.out.C) = v;
}

// max pool on split-f16, 8 channels per thread.  max(hi+lo) is taken on the reconstructed fp32 values.
__global__ void maxpool8_kernel(ConvParams p) {
    const int cpp = p.cout >> 3;
    const long long total = (long long)p.n_img * p.Ho * p.Wo * cpp;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % cpp);
    long long r = i / cpp;
    const int ox = (int)(r % p.Wo); r /= p.Wo;
    const int oy = (int)(r % p.Ho);
    const int img = (int)(r / p.Ho);
    const int pad = p.ksize / 2;
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int ky = 0; ky < p.ksize; ++ky) {
        const int iy = oy * p.stride - pad + ky;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.ksize; ++kx) {
            const int ix = ox * p.stride - pad + kx;
            if (ix < 0 || ix >= p.W) continue;
            float v[8];
            ld8(p.in, (long long)img * p.in.img + ((long long)iy * p.W + ix) * p.in.C + p.in.coff + ch * 8, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
        }
    }
    st8(p.out, (long long)img * p.out.img + ((long long)oy * p.Wo + ox) * p.out.C + p.out.coff + ch * 8, m);
}

// SPPF: three cascaded 5x5 stride-1 max pools (= 5x5, 9x9, 13x13 windows) of one 8-channel chunk of one
// image per block, the map held in shared memory; writes the three pooled slices next to each other in
// the concat buffer.  Replaces three launches that each re-read the map from L2.
__global__ void __launch_bounds__(256) sppf3_kernel(ConvParams p, int slice_stride) {
    extern __shared__ float s_map[];                    // [3][H*W][8]: ping, pong, row maxima
    const int hw = p.H * p.W;
    float* a = s_map;
    float* b = s_map + hw * 8;
    const int cpp = p.cout >> 3;
    const int img = blockIdx.x / cpp, ch = blockIdx.x - img * cpp;
    const long long in0 = (long long)img * p.in.img + p.in.coff + ch * 8;
    for (int px = threadIdx.x; px < hw; px += blockDim.x) {
        float v[8];
        ld8(p.in, in0 + (long long)px * p.in.C, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[px * 8 + k] = v[k];
    }
    __syncthreads();
    float* tmp = s_map + 2 * hw * 8;                    // row maxima (the 5x5 window is separable: 5 + 5 taps instead of 25)
    // thread = pixel, all 8 channels of the chunk in two float4 (one element per thread cost ~60 instructions of index
    // arithmetic and bounds tests per tap pass; per pixel they are paid once for eight channels)
    auto max8 = [](float4& m0, float4& m1, const float* q) {
        const float4 v0 = *reinterpret_cast<const float4*>(q), v1 = *reinterpret_cast<const float4*>(q + 4);
        m0.x = fmaxf(m0.x, v0.x); m0.y = fmaxf(m0.y, v0.y); m0.z = fmaxf(m0.z, v0.z); m0.w = fmaxf(m0.w, v0.w);
        m1.x = fmaxf(m1.x, v1.x); m1.y = fmaxf(m1.y, v1.y); m1.z = fmaxf(m1.z, v1.z); m1.w = fmaxf(m1.w, v1.w);
    };
    for (int pass = 0; pass < 3; ++pass) {
        for (int px = threadIdx.x; px < hw; px += blockDim.x) {
            const int oy = px / p.W, ox = px - oy * p.W;
            const int x_lo = max(ox - 2, 0), x_hi = min(ox + 2, p.W - 1);
            float4 m0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), m1 = m0;
            for (int ix = x_lo; ix <= x_hi; ++ix) max8(m0, m1, a + (oy * p.W + ix) * 8);
            *reinterpret_cast<float4*>(tmp + px * 8) = m0;
            *reinterpret_cast<float4*>(tmp + px * 8 + 4) = m1;
        }
        __syncthreads();
        const long long out0 = (long long)img * p.out.img + p.out.coff + (long long)pass * slice_stride + ch * 8;
        for (int px = threadIdx.x; px < hw; px += blockDim.x) {
            const int oy = px / p.W, ox = px - oy * p.W;
            const int y_lo = max(oy - 2, 0), y_hi = min(oy + 2, p.H - 1);
            float4 m0 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), m1 = m0;
            for (int iy = y_lo; iy <= y_hi; ++iy) max8(m0, m1, tmp + (iy * p.W + ox) * 8);
            *reinterpret_cast<float4*>(b + px * 8) = m0;
            *reinterpret_cast<float4*>(b + px * 8 + 4) = m1;
            st8(p.out, out0 + (long long)px * p.out.C, b + px * 8);      // this thread's own pixel: no barrier needed before the store
        }
        float* t = a; a = b; b = t;
        __syncthreads();
    }
}

// global mean over HxW then FC: logits[img][j] = bias[j] + sum_c mean_c * w[c][j]   (w stored [cin][cout])
__global__ void mean_fc_kernel(ConvParams p, float* __restrict__ logits) {
    extern __shared__ float s_mean[];
    const int img = blockIdx.x;
    const int hw = p.H * p.W;
    for (int c = threadIdx.x; c < p.cin; c += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < hw; ++q) s += ld_elem(p.in, (long long)img * p.in.img + (long long)q * p.in.C + p.in.coff + c);
        s_mean[c] = s / (float)hw;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.cout; j += blockDim.x) {
        float acc = __ldg(p.bias + j);
        for (int c = 0; c < p.cin; ++c) acc = fmaf(s_mean[c], __ldg(p.w + (long long)c * p.cout + j), acc);
        logits[(long long)img * p.cout + j] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Plan executor
// ---------------------------------------------------------------------------------------------
// Mark the small-channel convs of a plan that the parameter-weight kernels cover and keep host copies of their
// weights ([weights | bias]), which travel as kernel parameters at every launch.
int lp_assign_small_slots(lp_net_plan& net, cudaStream_t st) {
    net.small_slot.assign(net.ops.size(), -1);
    net.small_host.assign(net.ops.size(), std::vector<float>());
    for (size_t i = 0; i < net.ops.size(); ++i) {
        const lp_op_desc& op = net.ops[i];
        if (op.kind != LP_OP_STEM_U8 && op.kind != LP_OP_CONV) continue;
        const int nw = op.ksize * op.ksize * op.cin * op.cout;
        if (nw > SMALL_W_FLOATS || op.cout > 32) continue;
        if (op.out_seg_len > 0 || op.out_cstride > 1) continue;      // shapes the small kernels do not cover
        if (!small_dispatch(op.ksize, op.stride, op.cin, op.cout, op.kind == LP_OP_STEM_U8, SmallProbe{})) continue;
        std::vector<float>& h = net.small_host[i];
        h.resize((size_t)nw + op.cout);
        LP_CUDA(cudaMemcpyAsync(h.data(), net.weights + op.w_off, (size_t)nw * 4, cudaMemcpyDeviceToHost, st));
        LP_CUDA(cudaMemcpyAsync(h.data() + nw, net.weights + op.b_off, (size_t)op.cout * 4, cudaMemcpyDeviceToHost, st));
        net.small_slot[i] = (int)i;
    }
    LP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static TensorRef make_ref(const lp_net_plan& net, int buf, int coff, uint8_t* ws, int row_off) {
    TensorRef t{};
    if (buf < 0) { t.base = nullptr; return t; }
    const lp_buf_desc& b = net.bufs[buf];
    const int esz = b.fmt == LP_FMT_SPLIT16 ? 2 : (b.fmt == LP_FMT_F32 ? 4 : 1);
    t.base = ws + b.offset + (size_t)row_off * b.c * esz;
    t.img = b.image_bytes / esz;
    t.plane = (b.fmt == LP_FMT_SPLIT16) ? (long long)net.max_batch * t.img : 0;
    t.C = b.c;
    t.coff = coff;
    t.fmt = b.fmt;
    return t;
}

template <int KS, int STRIDE>
static void launch_conv(const ConvParams& p, cudaStream_t st) {
    const int cob = (p.cout % 32 == 0 || p.cout > 32) ? 32 : (p.cout > 8 ? 16 : 8);
    dim3 grid;
    if (KS == 1) grid = dim3((unsigned)(((long long)p.n_img * p.Ho * p.Wo + CONV_THREADS - 1) / CONV_THREADS), (p.cout + cob - 1) / cob, 1);
    else grid = dim3(((p.Wo + TILE_W - 1) / TILE_W) * ((p.Ho + TILE_H - 1) / TILE_H), (p.cout + cob - 1) / cob, p.n_img);
    if (cob == 32) conv_simt_kernel<KS, STRIDE, 32><<<grid, CONV_THREADS, 0, st>>>(p);
    else if (cob == 16) conv_simt_kernel<KS, STRIDE, 16><<<grid, CONV_THREADS, 0, st>>>(p);
    else conv_simt_kernel<KS, STRIDE, 8><<<grid, CONV_THREADS, 0, st>>>(p);
}

int lp_run_plan(lp_ctx* ctx, lp_net_plan& net, const uint8_t* in, int batch, void* workspace,
                size_t workspace_bytes, float* logits, cudaStream_t st) {
    LP_CHECK(net.loaded, "lp_run_plan: network not loaded");
    LP_CHECK(batch >= 1 && batch <= net.max_batch, "lp_run_plan: batch %d outside [1, %d]", batch, net.max_batch);
    LP_CHECK(workspace_bytes >= net.workspace_bytes, "lp_run_plan: workspace %zu B < required %zu B",
             workspace_bytes, net.workspace_bytes);
    uint8_t* ws = (uint8_t*)workspace;
    const int net_id = (&net == &ctx->nets[0]) ? 0 : 1;
    if (net.last_path.size() != net.ops.size()) net.last_path.assign(net.ops.size(), 0);
    for (size_t oi = 0; oi < net.ops.size(); ++oi) {
        net.last_path[oi] = 0;
        const lp_op_desc& op = net.ops[oi];
        // probe_op >= 0: ring of samples of that op; probe_op == -2: one sample of EVERY op (slot = op index)
        const bool probe_all = (ctx->probe_net == net_id && ctx->probe_op == -2 && !ctx->probe_ev.empty() && (int)oi < LP_PROBE_RING);
        const bool probe = probe_all || (ctx->probe_net == net_id && ctx->probe_op == (int)oi && !ctx->probe_ev.empty());
        const int slot = probe_all ? (int)oi : ctx->probe_n % LP_PROBE_RING;
        if (probe) LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * slot], st));
        struct ProbeStop {
            lp_ctx* c; bool on; int slot; cudaStream_t st;
            bool all;
            ~ProbeStop() { if (on) { cudaEventRecord(c->probe_ev[2 * slot + 1], st); if (all) { if (c->probe_n < slot + 1) c->probe_n = slot + 1; } else c->probe_n++; } }
        } probe_stop{ctx, probe, slot, st, probe_all};
        ConvParams p{};
        const lp_buf_desc& ob = net.bufs[op.out_buf >= 0 ? op.out_buf : op.in_buf];
        p.cin = op.cin; p.cout = op.cout; p.out_cstride = op.out_cstride > 0 ? op.out_cstride : 1;
        p.cout_real = op.cout_real > 0 ? op.cout_real : op.cout;
        p.seg_len = op.out_seg_len; p.seg_pad = op.out_seg_pad; p.seg_l0 = op.out_seg_len > 0 ? op.out_coff : 0;
        p.ksize = op.ksize; p.stride = op.stride; p.act = op.act; p.n_img = batch;
        p.res_first = (op.flags & LP_OPF_RES_BEFORE_ACT) ? 1 : 0;
        p.w = net.weights + op.w_off; p.bias = net.weights + op.b_off;
        p.in_scale_mean = 0.f; p.in_scale_std = 1.f;
        if (op.kind == LP_OP_STEM_U8) {
            // the network input is the caller's u8 image tensor, not a workspace buffer
            const lp_buf_desc& ib = net.bufs[op.in_buf];
            p.in.base = in; p.in.img = (long long)ib.h * ib.w * 3; p.in.C = 3; p.in.coff = 0; p.in.fmt = LP_FMT_U8;
            p.H = ib.h; p.W = ib.w;
            p.in_scale_mean = op.in_mean; p.in_scale_std = op.in_std;
        } else {
            const lp_buf_desc& ib = net.bufs[op.in_buf];
            p.in = make_ref(net, op.in_buf, op.in_coff, ws, 0);
            p.H = ib.h; p.W = ib.w;
            if (op.kind == LP_OP_CONV || op.kind == LP_OP_SCALE) p.res = make_ref(net, op.res_buf, op.res_coff, ws, 0);
        }
        if (op.kind != LP_OP_MEAN_FC) {
            p.out = make_ref(net, op.out_buf, op.out_seg_len > 0 ? 0 : op.out_coff, ws, op.row_off);
            const bool same_hw = op.kind == LP_OP_COPY || op.kind == LP_OP_SCALE;
            p.Ho = (op.kind == LP_OP_UPSAMPLE2) ? p.H * 2 : op.kind == LP_OP_GLOBAL_MEAN ? 1 : (same_hw ? p.H : (p.H + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1);
            p.Wo = (op.kind == LP_OP_UPSAMPLE2) ? p.W * 2 : op.kind == LP_OP_GLOBAL_MEAN ? 1 : (same_hw ? p.W : (p.W + 2 * (op.ksize / 2) - op.ksize) / op.stride + 1);
            if (!(ob.w == 1 && ob.h > 1))     // Detect-head row buffers ([anchors][C]) are addressed through row_off
                LP_CHECK(ob.h == p.Ho && ob.w == p.Wo, "op %zu: output buffer %dx%d != computed %dx%d", oi, ob.h, ob.w, p.Ho, p.Wo);
        }
        const long long total = (long long)batch * p.Ho * p.Wo * p.cout;
        // A 1x1 conv of the same width that is the ONLY consumer of this conv's output (the C2f cv1 behind a down-sampling
        // conv) can be applied in registers by the producing kernel: its input tensor is then never written or read.
        auto post_candidate = [&]() -> int {
            if (!(op.kind == LP_OP_CONV && op.ksize == 3 && op.res_buf < 0 && oi + 1 < net.ops.size())) return -1;
            if (ctx->probe_net == net_id && (ctx->probe_op == -2 || ctx->probe_op == (int)oi || ctx->probe_op == (int)oi + 1)) return -1;
            const lp_op_desc& o2 = net.ops[oi + 1];
            bool ok2 = o2.kind == LP_OP_CONV && o2.ksize == 1 && o2.stride == 1 && o2.cin == op.cout && o2.cout == op.cout && o2.flags == 0 &&
                       o2.in_buf == op.out_buf && o2.in_coff == op.out_coff && o2.res_buf < 0 && o2.out_seg_len == 0 &&
                       o2.out_cstride <= 1 && o2.out_coff % 8 == 0 && net.bufs[o2.out_buf].fmt == LP_FMT_SPLIT16 &&
                       net.bufs[o2.out_buf].h == ob.h && net.bufs[o2.out_buf].w == ob.w;
            for (size_t k = 0; ok2 && k < net.ops.size(); ++k)
                if (k != oi && k != oi + 1 && (net.ops[k].in_buf == op.out_buf || net.ops[k].res_buf == op.out_buf || net.ops[k].out_buf == op.out_buf))
                    ok2 = false;
            return ok2 ? (int)oi + 1 : -1;
        };
        // whole C2f body (bottleneck 3x3 chain + cv2) in one kernel, intermediates in shared memory (c2f_mma.cu)
        if (ctx->use_tc && ctx->use_mma && ctx->use_c2f && op.kind == LP_OP_CONV && op.ksize == 3 && op.stride == 1 && op.cin == op.cout &&
            !(ctx->probe_net == net_id && ctx->probe_op >= (int)oi && ctx->probe_op <= (int)oi + 4)) {
            const int n_cov = lp_c2f_fused_try(ctx, net, oi, batch, ws, st);
            if (n_cov < 0) return n_cov;
            if (n_cov > 0) {
                LP_LAUNCH_OK(ctx);
                net.last_path[oi] = 5;
                for (int j = 1; j < n_cov; ++j) {
                    net.last_path[oi + j] = 3;
                    if (probe_all && (int)(oi + j) < LP_PROBE_RING) {     // absorbed ops: an empty interval, so every slot of the pass is recorded
                        LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * (oi + j)], st));
                        LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * (oi + j) + 1], st));
                        if (ctx->probe_n < (int)(oi + j) + 1) ctx->probe_n = (int)(oi + j) + 1;
                    }
                }
                oi += n_cov - 1;
                continue;
            }
        }
        // warp-level tensor-core path for the small-channel layers (conv_mma.cu)
        if (ctx->use_tc && ctx->use_mma && op.kind == LP_OP_CONV && op.flags == 0 && op.out_seg_len == 0 && p.out_cstride == 1 &&
            p.in.fmt == LP_FMT_SPLIT16 && p.out.fmt == LP_FMT_SPLIT16 && op.cin <= 40 && op.cout <= 24 && (size_t)op.cout * 8 == (size_t)(op.cout / 8) * 64) {
            const int post_idx = post_candidate();
            ConvParams pq{};
            if (post_idx >= 0) {
                const lp_op_desc& o2 = net.ops[post_idx];
                pq.cin = o2.cin; pq.cout = o2.cout; pq.act = o2.act; pq.out_cstride = 1; pq.cout_real = o2.cout;
                pq.w = net.weights + o2.w_off; pq.bias = net.weights + o2.b_off;
                pq.out = make_ref(net, o2.out_buf, o2.out_coff, ws, o2.row_off);
                pq.Ho = p.Ho; pq.Wo = p.Wo; pq.n_img = batch; pq.ksize = 1; pq.stride = 1;
            }
            int r = lp_conv_mma_try(ctx, p, post_idx >= 0 ? &pq : nullptr, st);
            if (r == 0 && post_idx >= 0) r = lp_conv_mma_try(ctx, p, nullptr, st) ? 2 : 0;      // shape covered, fusion not
            if (r) {
                LP_LAUNCH_OK(ctx);
                net.last_path[oi] = 4;
                if (r == 1 && post_idx >= 0) { ++oi; net.last_path[oi] = 3; }
                continue;
            }
        }
        // stem + the down-sampling conv behind it (+ its fused 1x1) in one kernel: the stem's output tensor is never materialised
        if (ctx->use_tc && ctx->use_mma && ctx->use_c2f && op.kind == LP_OP_STEM_U8 && op.flags == 0 && oi + 1 < net.ops.size() &&
            !(ctx->probe_net == net_id && ctx->probe_op >= (int)oi && ctx->probe_op <= (int)oi + 2)) {
            const lp_op_desc& o1 = net.ops[oi + 1];
            bool ok1 = o1.kind == LP_OP_CONV && o1.ksize == 3 && o1.stride == 2 && o1.in_buf == op.out_buf && o1.in_coff == op.out_coff &&
                       o1.cin == op.cout && o1.res_buf < 0 && o1.flags == 0 && o1.out_seg_len == 0 && o1.out_cstride <= 1 &&
                       net.bufs[o1.out_buf].fmt == LP_FMT_SPLIT16 && o1.out_buf != op.out_buf;
            for (size_t k = 0; ok1 && k < net.ops.size(); ++k)
                if (k != oi && k != oi + 1 && (net.ops[k].in_buf == op.out_buf || net.ops[k].res_buf == op.out_buf || net.ops[k].out_buf == op.out_buf))
                    ok1 = false;
            if (ok1) {
                auto conv_params = [&](const lp_op_desc& o) {
                    ConvParams c{};
                    const lp_buf_desc& ib2 = net.bufs[o.in_buf];
                    c.cin = o.cin; c.cout = o.cout; c.out_cstride = 1; c.cout_real = o.cout; c.ksize = o.ksize; c.stride = o.stride; c.act = o.act;
                    c.n_img = batch; c.w = net.weights + o.w_off; c.bias = net.weights + o.b_off; c.in_scale_std = 1.f;
                    c.in = make_ref(net, o.in_buf, o.in_coff, ws, 0);
                    c.out = make_ref(net, o.out_buf, o.out_coff, ws, o.row_off);
                    c.H = ib2.h; c.W = ib2.w;
                    c.Ho = (c.H + 2 * (o.ksize / 2) - o.ksize) / o.stride + 1;
                    c.Wo = (c.W + 2 * (o.ksize / 2) - o.ksize) / o.stride + 1;
                    return c;
                };
                const ConvParams p1 = conv_params(o1);
                // the 1x1 of the same width that is the only consumer of o1's output (post_candidate's rule, one op further)
                int post_idx = -1;
                if (oi + 2 < net.ops.size() && !(ctx->probe_net == net_id && ctx->probe_op == -2)) {
                    const lp_op_desc& o2 = net.ops[oi + 2];
                    bool ok2 = o2.kind == LP_OP_CONV && o2.ksize == 1 && o2.stride == 1 && o2.cin == o1.cout && o2.cout == o1.cout && o2.flags == 0 &&
                               o2.in_buf == o1.out_buf && o2.in_coff == o1.out_coff && o2.res_buf < 0 && o2.out_seg_len == 0 && o2.out_cstride <= 1 &&
                               o2.out_coff % 8 == 0 && net.bufs[o2.out_buf].fmt == LP_FMT_SPLIT16 && net.bufs[o2.out_buf].h == net.bufs[o1.out_buf].h &&
                               net.bufs[o2.out_buf].w == net.bufs[o1.out_buf].w;
                    for (size_t k = 0; ok2 && k < net.ops.size(); ++k)
                        if (k != oi + 1 && k != oi + 2 && (net.ops[k].in_buf == o1.out_buf || net.ops[k].res_buf == o1.out_buf || net.ops[k].out_buf == o1.out_buf))
                            ok2 = false;
                    if (ok2) post_idx = (int)oi + 2;
                }
                ConvParams pq{};
                if (post_idx >= 0) pq = conv_params(net.ops[post_idx]);
                if (lp_stem_conv_try(ctx, p, p1, post_idx >= 0 ? &pq : nullptr, st)) {
                    LP_LAUNCH_OK(ctx);
                    net.last_path[oi] = 4;
                    const int n_abs = post_idx >= 0 ? 2 : 1;
                    for (int j = 1; j <= n_abs; ++j) {
                        net.last_path[oi + j] = 3;
                        if (probe_all && (int)(oi + j) < LP_PROBE_RING) {
                            LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * (oi + j)], st));
                            LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * (oi + j) + 1], st));
                            if (ctx->probe_n < (int)(oi + j) + 1) ctx->probe_n = (int)(oi + j) + 1;
                        }
                    }
                    oi += n_abs;
                    continue;
                }
            }
        }
        if (ctx->use_tc && ctx->use_mma && op.kind == LP_OP_STEM_U8 && op.flags == 0 && lp_stem_mma_try(ctx, p, st)) {
            LP_LAUNCH_OK(ctx);
            net.last_path[oi] = 4;
            continue;
        }
        if ((op.kind == LP_OP_STEM_U8 || op.kind == LP_OP_CONV) && net.small_slot.size() > oi && net.small_slot[oi] >= 0 && op.flags == 0 &&
            (p.res.base == nullptr || (p.res.fmt == LP_FMT_SPLIT16 && p.res.coff % 8 == 0)) && p.seg_len == 0 &&
            p.out_cstride == 1 && p.out.fmt == LP_FMT_SPLIT16 &&
            (op.kind == LP_OP_STEM_U8 || p.in.fmt == LP_FMT_SPLIT16) && p.in.coff % 8 == 0 && p.out.coff % 8 == 0) {
                    int post_slot = -1, post_act = 0;
            {
                const int cand = post_candidate();
                if (cand >= 0 && net.small_slot[cand] >= 0) {
                    const lp_op_desc& o2 = net.ops[cand];
                    post_slot = cand;
                    post_act = o2.act;
                    p.out = make_ref(net, o2.out_buf, o2.out_coff, ws, o2.row_off);
                }
            }
            SmallLaunch sl{&p, net.small_host[oi].data(), post_slot >= 0 ? net.small_host[post_slot].data() : nullptr, post_act, batch, st};
            const bool ran = small_dispatch(op.ksize, op.stride, op.cin, op.cout, op.kind == LP_OP_STEM_U8, sl);
            if (ran) {
                LP_LAUNCH_OK(ctx);
                net.last_path[oi] = 1;
                if (post_slot >= 0) { ++oi; net.last_path[oi] = 3; }   // the 1x1 conv is done
                continue;
            }
            LP_CHECK(post_slot < 0, "small conv dispatch failed after a fusion decision");
        }
        switch (op.kind) {
        case LP_OP_STEM_U8: {
            if (!(op.ksize == 3 && op.stride == 2 && op.cout <= 32)) {          // ResNet conv1 (7x7 s2, 64 channels) and friends
                LP_CHECK(op.ksize % 2 == 1 && op.stride >= 1, "stem: kernel size must be odd");
                const long long threads = (long long)batch * p.Ho * p.Wo * ((op.cout + 7) / 8);
                stem_u8_generic_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p);
                break;
            }
            dim3 grid(((p.Wo + TILE_W - 1) / TILE_W) * ((p.Ho + TILE_H - 1) / TILE_H), 1, batch);
            if (op.cout <= 8) stem_u8_kernel<8><<<grid, CONV_THREADS, 0, st>>>(p);
            else if (op.cout <= 16) stem_u8_kernel<16><<<grid, CONV_THREADS, 0, st>>>(p);
            else if (op.cout <= 24) stem_u8_kernel<24><<<grid, CONV_THREADS, 0, st>>>(p);
            else stem_u8_kernel<32><<<grid, CONV_THREADS, 0, st>>>(p);
            break;
        }
        case LP_OP_CONV: {
            if (ctx->use_tc && op.wtc_off >= 0 && net.weights_tc) {
                int r = lp_conv_tc_try(ctx, net, op, batch, ws, st);
                if (r < 0) return r;
                if (r == 1) { ctx->launches++; net.last_path[oi] = 2; continue; }
            }
            if (op.ksize == 1 && op.stride == 1 && op.cout <= 4 && op.cin % 8 == 0 && p.in.fmt == LP_FMT_SPLIT16 && p.in.coff % 8 == 0 &&
                p.res.base == nullptr && (size_t)(op.cin + 1) * op.cout * 4 <= 40 * 1024) {
                const long long px = (long long)batch * p.Ho * p.Wo;
                conv1x1_few_kernel<<<(unsigned)((px + 255) / 256), 256, (size_t)(op.cin + 1) * op.cout * sizeof(float), st>>>(p);
            } else if (op.ksize == 1 && op.stride == 1) launch_conv<1, 1>(p, st);
            else if (op.ksize == 1 && op.stride == 2) launch_conv<1, 2>(p, st);
            else if (op.ksize == 3 && op.stride == 1) launch_conv<3, 1>(p, st);
            else if (op.ksize == 3 && op.stride == 2) launch_conv<3, 2>(p, st);
            else LP_CHECK(false, "conv: unsupported ksize %d stride %d", op.ksize, op.stride);
            break;
        }
        case LP_OP_DWCONV3:
            dwconv3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
            break;
        case LP_OP_MAXPOOL: {
            const bool v8 = p.in.fmt == LP_FMT_SPLIT16 && p.out.fmt == LP_FMT_SPLIT16 && p.cout % 8 == 0 &&
                            p.in.coff % 8 == 0 && p.out.coff % 8 == 0 && p.out_cstride == 1 && p.seg_len == 0;
            // SPPF cascade: this 5x5 s1 pool feeds a second and a third one, all slices of one buffer
            if (v8 && op.ksize == 5 && op.stride == 1 && oi + 2 < net.ops.size() &&
                !(ctx->probe_net == net_id && (ctx->probe_op == -2 || (ctx->probe_op >= (int)oi && ctx->probe_op <= (int)oi + 2))) &&
                (size_t)p.H * p.W * 96 <= 48 * 1024) {
                const lp_op_desc& o1 = net.ops[oi + 1];
                const lp_op_desc& o2 = net.ops[oi + 2];
                const int step = o1.out_coff - op.out_coff;
                auto chained = [&](const lp_op_desc& a, const lp_op_desc& b) {
                    return b.kind == LP_OP_MAXPOOL && b.ksize == 5 && b.stride == 1 && b.cout == a.cout && b.in_buf == a.out_buf &&
                           b.in_coff == a.out_coff && b.out_buf == a.out_buf && b.out_seg_len == 0 && b.out_cstride <= 1;
                };
                if (chained(op, o1) && chained(o1, o2) && o2.out_coff - o1.out_coff == step && step >= op.cout) {
                    sppf3_kernel<<<batch * (p.cout / 8), 256, (size_t)p.H * p.W * 96, st>>>(p, step);
                    net.last_path[oi + 1] = net.last_path[oi + 2] = 3;
                    oi += 2;                          // the two downstream pools are done
                    break;
                }
            }
            if (v8) maxpool8_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, st>>>(p);
            else maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
            break;
        }
        case LP_OP_UPSAMPLE2:
        case LP_OP_COPY: {
            p.stride = op.kind == LP_OP_UPSAMPLE2 ? 2 : 1;
            const bool v8 = p.in.fmt == LP_FMT_SPLIT16 && p.out.fmt == LP_FMT_SPLIT16 && p.cout % 8 == 0 &&
                            p.in.coff % 8 == 0 && p.out.coff % 8 == 0 && p.out_cstride == 1 && p.seg_len == 0;
            if (v8 && op.kind == LP_OP_UPSAMPLE2 && total / 16 < (1ll << 31) && p.Ho == 2 * p.H && p.Wo == 2 * p.W)
                upsample2x8_kernel<<<(unsigned)((total / 16 + 255) / 256), 256, 0, st>>>(p);
            else if (v8 && total / 4 < (1ll << 31)) resample_copy8_kernel<<<(unsigned)((total / 4 + 255) / 256), 256, 0, st>>>(p);
            else resample_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
            break;
        }
        case LP_OP_GLOBAL_MEAN:
            global_mean_kernel<<<(unsigned)(((long long)batch * p.cout + 127) / 128), 128, 0, st>>>(p);
            break;
        case LP_OP_SCALE:
            LP_CHECK(p.res.base != nullptr, "scale: no gate buffer");
            scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
            break;
        case LP_OP_MEAN_FC:
            LP_CHECK(logits != nullptr, "mean_fc: logits pointer is null");
            mean_fc_kernel<<<batch, 256, op.cin * sizeof(float), st>>>(p, logits);
            break;
        default:
            LP_CHECK(false, "unknown op kind %d", op.kind);
        }
        LP_LAUNCH_OK(ctx);
    }
    return 0;
}
