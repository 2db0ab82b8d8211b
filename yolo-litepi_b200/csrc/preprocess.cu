// K1: fused letterbox = cv2.resize(INTER_LINEAR, u8 fixed point) + copyMakeBorder(114) + BGR->RGB.
// Replaces letterbox() and cv2.cvtColor in the reference (src/vntsr/pipeline/e2e.py:66-86, :224-225).
// Output is the u8 RGB letterboxed image; the x/255 normalisation (e2e.py:234-236) is fused into the
// detector's stem conv, so the 640x640 float tensor is never materialised.
//
// Bit-exact restatement of OpenCV's 8-bit bilinear resize (11-bit coefficients):
//   scale = 1/(dst/src) (double);  f = float((d+0.5)*scale - 0.5);  s = floor(f);  a = f - s
//   x: s<0 -> s=0,a=0;  s>=W-1 -> s=W-1,a=0;   y: rows clamped, coefficient kept
//   c = rint(coef*2048) (int16);  t = S[s]*c0 + S[s+1]*c1
//   dst = (((b0*(t0>>4))>>16) + ((b1*(t1>>4))>>16) + 2) >> 2
#include "common.cuh"

struct FrameTable {
    const uint8_t* ptr[LP_MAX_TABLE];
    long long pitch[LP_MAX_TABLE];
    int h[LP_MAX_TABLE];
    int w[LP_MAX_TABLE];
    int new_w[LP_MAX_TABLE];   // resized size inside the letterbox
    int new_h[LP_MAX_TABLE];
    int left[LP_MAX_TABLE];
    int top[LP_MAX_TABLE];
};

// grid (S/4/blockDim.x.., S/LB_ROWS, B); each thread produces 4 consecutive output pixels (12 bytes) of LB_ROWS
// consecutive rows: the column coefficients (two fp64 operations each) are computed once and reused for every row.
constexpr int LB_ROWS = 8;
template <int PX>
__global__ void __launch_bounds__(160) letterbox_kernel(FrameTable tab, int S, uint8_t* __restrict__ out) {
    const int b = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (x0 >= S) return;
    const int nw = tab.new_w[b], nh = tab.new_h[b], left = tab.left[b], top = tab.top[b];
    const int W = tab.w[b], H = tab.h[b];
    const uint8_t* __restrict__ src = tab.ptr[b];
    const long long pitch = tab.pitch[b];
    const bool same = (nw == W && nh == H);   // reference skips cv2.resize when the size is unchanged
    const double scale_x = same ? 1.0 : 1.0 / ((double)nw / (double)W);
    const double scale_y = same ? 1.0 : 1.0 / ((double)nh / (double)H);
    int sxo[PX], sx1o[PX], a0[PX], a1[PX];
    bool col_in[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        const int rx = x0 + i - left;
        col_in[i] = rx >= 0 && rx < nw;
        sxo[i] = 0; sx1o[i] = 0; a0[i] = 2048; a1[i] = 0;
        if (col_in[i]) {
            if (same) { sxo[i] = rx * 3; }
            else {
                int sx;
                lin_coef(rx, scale_x, W, true, sx, a0[i], a1[i]);
                sxo[i] = sx * 3;
                sx1o[i] = min(sx + 1, W - 1) * 3;
            }
        }
    }
    for (int rr = 0; rr < LB_ROWS; ++rr) {
        const int dy = blockIdx.y * LB_ROWS + rr;
        if (dy >= S) return;
        uint8_t px[PX * 3];
        const int ry = dy - top;
        const bool row_in = (ry >= 0 && ry < nh);
        int sy = 0, b0 = 2048, b1 = 0;
        const uint8_t *r0 = src, *r1 = src;
        if (row_in) {
            if (!same) {
                lin_coef(ry, scale_y, H, false, sy, b0, b1);
                const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
                r0 = src + (long long)y0 * pitch;
                r1 = src + (long long)y1 * pitch;
            } else {
                r0 = src + (long long)ry * pitch;
            }
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            uint8_t B = 114, G = 114, R = 114;
            if (row_in && col_in[i]) {
                if (same) {
                    const uint8_t* p = r0 + sxo[i];
                    B = p[0]; G = p[1]; R = p[2];
                } else {
                    const uint8_t *p00 = r0 + sxo[i], *p01 = r0 + sx1o[i], *p10 = r1 + sxo[i], *p11 = r1 + sx1o[i];
                    uint8_t o[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int t0 = (int)__ldg(p00 + c) * a0[i] + (int)__ldg(p01 + c) * a1[i];
                        const int t1 = (int)__ldg(p10 + c) * a0[i] + (int)__ldg(p11 + c) * a1[i];
                        const int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
                        o[c] = (uint8_t)v;
                    }
                    B = o[0]; G = o[1]; R = o[2];
                }
            }
            px[i * 3 + 0] = R; px[i * 3 + 1] = G; px[i * 3 + 2] = B;   // BGR -> RGB
        }
        uint8_t* dst = out + ((size_t)b * S + dy) * (size_t)S * 3 + (size_t)x0 * 3;
        if (PX == 4 && x0 + 4 <= S) {          // 12 bytes, 4-byte aligned because S*3 and x0*3 are multiples of 4
            uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
            d32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
            d32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
            d32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
        } else {
            for (int i = 0; i < PX && x0 + i < S; ++i)
                for (int c = 0; c < 3; ++c) dst[i * 3 + c] = px[i * 3 + c];
        }
    }
}

// Host-side geometry: exactly the reference's python arithmetic (e2e.py:72-83), in double.
static inline long py_round(double v) {   // Python round(): half to even
    return (long)nearbyint(v);
}

extern "C" int lp_letterbox(lp_ctx* ctx, const uint8_t* const* frames_h, const int32_t* h_h, const int32_t* w_h,
                            const int64_t* pitch_h, int batch, int out_size, uint8_t* out, double* ratio_h,
                            double* pad_h, void* stream) {
    LP_CHECK(ctx && frames_h && h_h && w_h && out, "lp_letterbox: null argument");
    lp_device_guard dev_guard(ctx);
    LP_CHECK(batch >= 0 && out_size > 0 && out_size % 4 == 0, "lp_letterbox: bad batch/out_size");
    cudaStream_t st = (cudaStream_t)stream;
    for (int base = 0; base < batch; base += LP_MAX_TABLE) {
        const int n = batch - base < LP_MAX_TABLE ? batch - base : LP_MAX_TABLE;
        FrameTable tab;
        for (int i = 0; i < n; ++i) {
            const int H = h_h[base + i], W = w_h[base + i];
            LP_CHECK(H > 0 && W > 0, "lp_letterbox: frame %d has empty shape", base + i);
            const double r = fmin((double)out_size / H, (double)out_size / W);
            const int nw = (int)py_round(W * r), nh = (int)py_round(H * r);
            LP_CHECK(nw > 0 && nh > 0 && nw <= out_size && nh <= out_size, "lp_letterbox: degenerate resize");
            const double dw = (out_size - nw) / 2.0, dh = (out_size - nh) / 2.0;
            tab.ptr[i] = frames_h[base + i];
            tab.pitch[i] = pitch_h ? pitch_h[base + i] : (long long)W * 3;
            tab.h[i] = H; tab.w[i] = W; tab.new_w[i] = nw; tab.new_h[i] = nh;
            tab.left[i] = (int)py_round(dw - 0.1);
            tab.top[i] = (int)py_round(dh - 0.1);
            if (ratio_h) ratio_h[base + i] = r;
            if (pad_h) { pad_h[2 * (base + i)] = dw; pad_h[2 * (base + i) + 1] = dh; }
        }
        dim3 block(160, 1, 1);
        dim3 grid((out_size / 4 + 159) / 160, (out_size + LB_ROWS - 1) / LB_ROWS, n);
        letterbox_kernel<4><<<grid, block, 0, st>>>(tab, out_size, out + (size_t)base * out_size * out_size * 3);
        LP_LAUNCH_OK(ctx);
    }
    return 0;
}
