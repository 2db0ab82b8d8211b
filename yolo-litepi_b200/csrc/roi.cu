// K6: batched ROI crop + Pillow-exact antialiased BILINEAR resize to SxS, BGR->RGB.
// Replaces, per ROI, image[y1:y2,x1:x2] -> cv2.cvtColor -> Image.fromarray -> transforms.Resize((64,64))
// (src/vntsr/pipeline/e2e.py:471, :385-388).  Pillow's resize lives outside the reference repo (Pillow,
// unpinned in requirements.txt); this restates its published algorithm (src/libImaging/Resample.c:
// precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc):
//   scale = in/out; fs = max(scale,1); support = fs; center = (x+0.5)*scale;
//   xmin = max(int(center-support+0.5),0); xmax = min(int(center+support+0.5),in) - xmin
//   w_k = max(0, 1-|(k+xmin-center+0.5)/fs|), normalised to sum 1 in double, kk = int(0.5 + w*2^22)
//   pixel = clip8((2^21 + sum kk*p) >> 22); horizontal pass first into a u8 intermediate, then vertical.
// The ToTensor/Normalize step ((u8/255-0.18)/0.34, e2e.py:368-369) is fused into the classifier stem.
#include "common.cuh"

#define ROI_PRECISION_BITS 22
constexpr int ROI_SPLIT = 4;       // blocks per ROI (bands of output rows)

struct RoiFrames {
    const uint8_t* ptr[LP_MAX_TABLE];
    long long pitch[LP_MAX_TABLE];
};

// One thread computes one output index of one axis; all double ops are explicit _rn (no FMA).
__device__ void pil_coeffs(int in_size, int out_size, int xx, int ksize, int* bounds, int* kk) {
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double fs = scale < 1.0 ? 1.0 : scale;
    const double support = fs;                       // bilinear support 1.0 * filterscale
    const double center = __dmul_rn((double)xx + 0.5, scale);
    const double ss = __ddiv_rn(1.0, fs);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
    }
    for (int x = 0; x < ksize; ++x) {
        int q = 0;
        if (x < xmax) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            q = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << ROI_PRECISION_BITS)));
        }
        kk[x] = q;
    }
    bounds[0] = xmin;
    bounds[1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= ROI_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Block per ROI.  Shared memory: coefficient tables for both axes + a ring of horizontally
// resampled rows (u8, S*3 bytes each).
__global__ void __launch_bounds__(256) roi_resize_kernel(RoiFrames fr, int img_base, const int* __restrict__ roi_xyxy,
                                                         const int* __restrict__ roi_src, int n_rois, const int* __restrict__ n_dev, int S, int kmax,
                                                         int tmp_rows, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* kx = reinterpret_cast<int*>(smem);            // [S][kmax]
    int* ky = kx + S * kmax;                           // [S][kmax]
    int* bx = ky + S * kmax;                           // [S][2]
    int* by = bx + S * 2;                              // [S][2]
    int* s_misc = by + S * 2;                          // [4]
    uint8_t* tmp = reinterpret_cast<uint8_t*>(s_misc + 4);   // [tmp_rows][S*3]
    // ROI_SPLIT blocks per ROI, each a band of output rows: a block's critical path (dependent byte loads) was the
    // kernel's duration, and 333 ROIs leave most of the 148 SMs' warp slots empty anyway
    const int r = blockIdx.x / ROI_SPLIT, band = blockIdx.x - r * ROI_SPLIT;
    if (r >= n_rois || (n_dev && r >= *n_dev)) return;
    const int img = roi_src[2 * r] - img_base;
    if (img < 0 || img >= LP_MAX_TABLE) return;        // ROI of another 64-image launch chunk
    const int x1 = roi_xyxy[4 * r], y1 = roi_xyxy[4 * r + 1], x2 = roi_xyxy[4 * r + 2], y2 = roi_xyxy[4 * r + 3];
    const int rw = x2 - x1, rh = y2 - y1;
    const uint8_t* __restrict__ src = fr.ptr[img] + (long long)y1 * fr.pitch[img] + (long long)x1 * 3;
    const long long pitch = fr.pitch[img];
    const int tid = threadIdx.x;
    {
        const double sx = (double)rw / S, sy = (double)rh / S;
        const int ksx = (int)ceil(sx < 1.0 ? 1.0 : sx) * 2 + 1, ksy = (int)ceil(sy < 1.0 ? 1.0 : sy) * 2 + 1;
        for (int t = tid; t < 2 * S; t += blockDim.x) {
            if (t < S) pil_coeffs(rw, S, t, min(ksx, kmax), bx + 2 * t, kx + t * kmax);
            else pil_coeffs(rh, S, t - S, min(ksy, kmax), by + 2 * (t - S), ky + (t - S) * kmax);
        }
    }
    __syncthreads();
    uint8_t* dst = out + (long long)r * S * S * 3;
    const int row_elems = S * 3;
    const int band_rows = (S + ROI_SPLIT - 1) / ROI_SPLIT;
    int y0 = band * band_rows;
    const int y_end = min(S, y0 + band_rows);
    while (y0 < y_end) {
        // group of output rows [y0, y0+G) whose input rows fit in the tmp ring
        if (tid == 0) {
            const int rmin = by[2 * y0];
            int g = 1;
            while (y0 + g < y_end && by[2 * (y0 + g)] + by[2 * (y0 + g) + 1] - rmin <= tmp_rows) ++g;
            s_misc[0] = g;
            s_misc[1] = rmin;
            s_misc[2] = by[2 * (y0 + g - 1)] + by[2 * (y0 + g - 1) + 1] - rmin;   // rows to resample
        }
        __syncthreads();
        const int G = s_misc[0], rmin = s_misc[1], nrows = s_misc[2];
        // horizontal pass: tmp[ry][xx][c], channel order already swapped to RGB
        for (int e = tid; e < nrows * row_elems; e += blockDim.x) {
            const int ry = e / row_elems, q = e - ry * row_elems;
            const int xx = q / 3, c = q - xx * 3;
            const int xmin = bx[2 * xx], cnt = bx[2 * xx + 1];
            const int* k = kx + xx * kmax;
            const uint8_t* p = src + (long long)(rmin + ry) * pitch + (long long)xmin * 3 + (2 - c);
            int ss = 1 << (ROI_PRECISION_BITS - 1);
            for (int t = 0; t < cnt; ++t) ss += (int)__ldg(p + t * 3) * k[t];
            tmp[e] = clip8(ss);
        }
        __syncthreads();
        // vertical pass
        for (int e = tid; e < G * row_elems; e += blockDim.x) {
            const int gy = e / row_elems, q = e - gy * row_elems;
            const int y = y0 + gy;
            const int ymin = by[2 * y] - rmin, cnt = by[2 * y + 1];
            const int* k = ky + y * kmax;
            int ss = 1 << (ROI_PRECISION_BITS - 1);
            for (int t = 0; t < cnt; ++t) ss += (int)tmp[(ymin + t) * row_elems + q] * k[t];
            dst[y * row_elems + q] = clip8(ss);
        }
        __syncthreads();
        y0 += G;
    }
}

// e2e_optimize.py:391-393: cv2.cvtColor(BGR2RGB) + cv2.resize(INTER_LINEAR) of the ROI, bit-exact with OpenCV's
// 11-bit fixed point (same arithmetic as the letterbox kernel).  Block per ROI; coefficients of the S columns
// and S rows are tabulated in shared memory once, then each thread produces output pixels.
__global__ void __launch_bounds__(256) roi_resize_linear_kernel(RoiFrames fr, int img_base, const int* __restrict__ roi_xyxy,
                                                                const int* __restrict__ roi_src, int n_rois,
                                                                const int* __restrict__ n_dev, int S, uint8_t* __restrict__ out) {
    extern __shared__ int s_lin[];                      // [S][3] x: (sx, a0, a1) then [S][3] y: (sy, b0, b1)
    const int r = blockIdx.x;
    if (r >= n_rois || (n_dev && r >= *n_dev)) return;
    const int img = roi_src[2 * r] - img_base;
    if (img < 0 || img >= LP_MAX_TABLE) return;
    const int x1 = roi_xyxy[4 * r], y1 = roi_xyxy[4 * r + 1], rw = roi_xyxy[4 * r + 2] - x1, rh = roi_xyxy[4 * r + 3] - y1;
    const long long pitch = fr.pitch[img];
    const uint8_t* __restrict__ src = fr.ptr[img] + (long long)y1 * pitch + (long long)x1 * 3;
    uint8_t* dst = out + (long long)r * S * S * 3;
    if (rw == S && rh == S) {                            // cv2.resize returns a copy when the size is unchanged
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
            const int y = e / S, x = e - y * S;
            const uint8_t* p = src + (long long)y * pitch + x * 3;
            dst[e * 3 + 0] = p[2]; dst[e * 3 + 1] = p[1]; dst[e * 3 + 2] = p[0];
        }
        return;
    }
    const double scale_x = 1.0 / ((double)S / (double)rw), scale_y = 1.0 / ((double)S / (double)rh);
    for (int t = threadIdx.x; t < 2 * S; t += blockDim.x) {
        int s0, c0, c1;
        if (t < S) lin_coef(t, scale_x, rw, true, s0, c0, c1);
        else lin_coef(t - S, scale_y, rh, false, s0, c0, c1);
        s_lin[t * 3] = s0; s_lin[t * 3 + 1] = c0; s_lin[t * 3 + 2] = c1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        const int y = e / S, x = e - y * S;
        const int sx = s_lin[x * 3], a0 = s_lin[x * 3 + 1], a1 = s_lin[x * 3 + 2];
        const int sy = s_lin[(S + y) * 3], b0 = s_lin[(S + y) * 3 + 1], b1 = s_lin[(S + y) * 3 + 2];
        const int sx1 = min(sx + 1, rw - 1);
        const int y0 = min(max(sy, 0), rh - 1), yy1 = min(max(sy + 1, 0), rh - 1);
        const uint8_t *r0 = src + (long long)y0 * pitch, *r1 = src + (long long)yy1 * pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int t0 = (int)__ldg(r0 + sx * 3 + c) * a0 + (int)__ldg(r0 + sx1 * 3 + c) * a1;
            const int t1 = (int)__ldg(r1 + sx * 3 + c) * a0 + (int)__ldg(r1 + sx1 * 3 + c) * a1;
            dst[e * 3 + (2 - c)] = (uint8_t)((((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2);
        }
    }
}

extern "C" size_t lp_roi_resize_scratch_bytes(int, int, int) { return 0; }   // coefficients live in shared memory

extern "C" int lp_roi_resize(lp_ctx* ctx, const uint8_t* const* frames_h, const int64_t* pitch_h, int batch,
                             const int32_t* roi_xyxy, const int32_t* roi_src, int n_rois, int out_size,
                             int max_side, uint8_t* out, void* stream) {
    LP_CHECK(ctx && frames_h && pitch_h && roi_xyxy && roi_src && out, "lp_roi_resize: null argument");
    lp_device_guard dev_guard(ctx);
    LP_CHECK(out_size > 0 && out_size <= 128 && max_side > 0, "lp_roi_resize: bad out_size/max_side");
    if (n_rois <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->roi_mode == 1) {
        for (int base = 0; base < batch; base += LP_MAX_TABLE) {
            const int n = batch - base < LP_MAX_TABLE ? batch - base : LP_MAX_TABLE;
            RoiFrames fr;
            for (int i = 0; i < n; ++i) { fr.ptr[i] = frames_h[base + i]; fr.pitch[i] = pitch_h[base + i]; }
            for (int i = n; i < LP_MAX_TABLE; ++i) { fr.ptr[i] = nullptr; fr.pitch[i] = 0; }
            roi_resize_linear_kernel<<<n_rois, 256, (size_t)out_size * 6 * sizeof(int), st>>>(fr, base, roi_xyxy, roi_src, n_rois,
                                                                                            ctx->roi_count_dev, out_size, out);
            LP_LAUNCH_OK(ctx);
        }
        return 0;
    }
    const double sc = (double)max_side / out_size;
    const int kmax = (int)ceil(sc < 1.0 ? 1.0 : sc) * 2 + 1;
    int tmp_rows = 3 * kmax > 96 ? 3 * kmax : 96;
    size_t smem = (size_t)(2 * out_size * kmax + 4 * out_size + 4) * 4 + (size_t)tmp_rows * out_size * 3;
    LP_CHECK(smem <= 200 * 1024, "lp_roi_resize: max_side %d needs %zu B shared memory", max_side, smem);
    if (!(ctx->attr_set & 4)) {          // per context (= per device): the opt-in is a per-device function attribute
        LP_CUDA(cudaFuncSetAttribute(roi_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->attr_set |= 4;
    }
    for (int base = 0; base < batch; base += LP_MAX_TABLE) {
        const int n = batch - base < LP_MAX_TABLE ? batch - base : LP_MAX_TABLE;
        RoiFrames fr;
        for (int i = 0; i < n; ++i) { fr.ptr[i] = frames_h[base + i]; fr.pitch[i] = pitch_h[base + i]; }
        for (int i = n; i < LP_MAX_TABLE; ++i) { fr.ptr[i] = nullptr; fr.pitch[i] = 0; }
        roi_resize_kernel<<<n_rois * ROI_SPLIT, 256, smem, st>>>(fr, base, roi_xyxy, roi_src, n_rois, ctx->roi_count_dev, out_size, kmax, tmp_rows, out);
        LP_LAUNCH_OK(ctx);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// softmax + argmax over classifier logits (torch.softmax(dim=1), np.argmax; e2e.py:394-396).
// One warp per ROI.
// ---------------------------------------------------------------------------------------------
__global__ void softmax_argmax_kernel(const float* __restrict__ logits, int n, const int* __restrict__ n_dev, int C,
                                      float* __restrict__ probs, long long* __restrict__ argmax) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n || (n_dev && r >= *n_dev)) return;
    const float* x = logits + (long long)r * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float pr = expf(x[c] - m) / s;
        probs[(long long)r * C + c] = pr;
        if (pr > best) { best = pr; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) argmax[r] = bi;
}

int lp_launch_softmax_argmax(lp_ctx* ctx, const float* logits, int n, int C, float* probs, int64_t* argmax, cudaStream_t st) {
    if (n <= 0) return 0;
    softmax_argmax_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(logits, n, ctx->roi_count_dev, C, probs, (long long*)argmax);
    LP_LAUNCH_OK(ctx);
    return 0;
}
