// K4 + K5: Detect tail (DFL softmax-expectation, dist2bbox, sigmoid), then the reference's numpy
// post-processing bit for bit: threshold, xywh->xyxy, un-letterbox, clip, per-class greedy NMS.
//   Detect tail      : model.ncnn.param:184-208 (reshape/concat/softmax/conv_65/anchor math/sigmoid)
//   postprocess      : src/vntsr/pipeline/e2e.py:240-296
//   nms_numpy        : src/vntsr/pipeline/e2e.py:89-119
//   ROI int/clip/area: src/vntsr/pipeline/e2e.py:459-475
// Everything after out0 is float32 with numpy's operation order and IEEE division; every op is an
// explicit round-to-nearest intrinsic so nvcc cannot contract multiplies and adds into FMAs.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Detect tail: head_raw [B][A][HC] f32 (64 DFL logits + nc class logits per anchor) -> out0 [B][4+nc][A]
// 4 lanes per anchor (one per box side), 8 anchors per warp: the warp reads 8 contiguous rows.
// ---------------------------------------------------------------------------------------------
struct LevelTable {
    int n;
    int start[4];
    int gw[4];
    float stride[4];
};

__global__ void __launch_bounds__(256) detect_tail_kernel(const float* __restrict__ head, int B, int A, int HC, int nc,
                                                          LevelTable lv, float* __restrict__ out0) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid >> 2;            // (image, anchor)
    const int side = (int)(gid & 3);
    const bool valid = row < (long long)B * A;
    const long long rr = valid ? row : 0;
    const int b = (int)(rr / A), a = (int)(rr % A);
    const float* p = head + rr * HC + side * 16;
    float x[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    float m = x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, x[i]);
    float s = 0.f, d = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float e;                                  // ex2.approx of a non-positive argument: relative error 2^-22, 1e-4 px on the box after
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x[i] - m) * 1.4426950408889634f));   // the x stride / ratio scaling (expf: 5x the instructions)
        s += e;
        d += e * (float)i;
    }
    d = d / s;                                   // expected bin = distance in stride units
    const unsigned full = 0xffffffffu;
    const int lane0 = (threadIdx.x & 31) & ~3;
    const float dl = __shfl_sync(full, d, lane0 + 0), dt = __shfl_sync(full, d, lane0 + 1);
    const float dr = __shfl_sync(full, d, lane0 + 2), db = __shfl_sync(full, d, lane0 + 3);
    if (!valid) return;
    int l = 0;
    for (int i = 1; i < lv.n; ++i) if (a >= lv.start[i]) l = i;
    const int li = a - lv.start[l];
    const float ax = (float)(li % lv.gw[l]) + 0.5f, ay = (float)(li / lv.gw[l]) + 0.5f;
    const float x1 = ax - dl, y1 = ay - dt, x2 = ax + dr, y2 = ay + db;
    float v;
    if (side == 0) v = (x1 + x2) / 2.f;
    else if (side == 1) v = (y1 + y2) / 2.f;
    else if (side == 2) v = x2 - x1;
    else v = y2 - y1;
    float* o = out0 + (long long)b * (4 + nc) * A;
    o[(long long)side * A + a] = v * lv.stride[l];
    for (int c = side; c < nc; c += 4) {
        const float z = head[rr * HC + 64 + c];
        o[(long long)(4 + c) * A + a] = 1.f / (1.f + expf(-z));
    }
}

int lp_launch_detect_tail(lp_ctx* ctx, const float* head_raw, int batch, int head_c, int in_size, int nc, int n_anchors,
                          float* out0, cudaStream_t st) {
    // geometry of the Detect head: strides 8/16/32 (model.ncnn.param:150, 184-186) on an in_size x in_size input
    LevelTable lv{};
    lv.n = 3;
    const int S = in_size;
    LP_CHECK(S >= 32 && S % 32 == 0, "detect tail: input size %d is not a positive multiple of 32", S);
    int start = 0;
    for (int i = 0; i < 3; ++i) {
        const int s = 8 << i;
        lv.start[i] = start; lv.gw[i] = S / s; lv.stride[i] = (float)s;
        start += (S / s) * (S / s);
    }
    const int A = start;
    LP_CHECK(A == n_anchors, "detect tail: head buffer holds %d anchors, input size %d gives %d", n_anchors, S, A);
    LP_CHECK(nc >= 1 && head_c >= 64 + nc, "detect tail: %d head channels cannot hold 64 DFL logits + %d classes", head_c, nc);
    const long long threads = (long long)batch * A * 4;
    detect_tail_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(head_raw, batch, A, head_c, nc, lv, out0);
    LP_LAUNCH_OK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Candidates: threshold + decode + ordered compaction.  One block per image, anchors scanned in
// ascending order so the candidate list has numpy's boolean-mask order (e2e.py:258-261).
// ---------------------------------------------------------------------------------------------
struct ImageTable {
    int h[LP_MAX_TABLE];
    int w[LP_MAX_TABLE];
    float ratio[LP_MAX_TABLE];
    float padw[LP_MAX_TABLE];
    float padh[LP_MAX_TABLE];
};

struct Cand {          // 24 B per candidate
    float x1, y1, x2, y2, score;
    int cls;
};

__global__ void __launch_bounds__(1024) candidates_kernel(const float* __restrict__ out0, int nc, int A, ImageTable tab,
                                                          float conf, Cand* __restrict__ cands, int* __restrict__ n_cand) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* o = out0 + (long long)b * (4 + nc) * A;
    Cand* out = cands + (long long)b * A;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const float ratio = tab.ratio[b], padw = tab.padw[b], padh = tab.padh[b];
    const float fw = (float)tab.w[b], fh = (float)tab.h[b];
    for (int a0 = 0; a0 < A; a0 += 1024) {
        const int a = a0 + tid;
        float best = -INFINITY;
        int bc = 0;
        if (a < A) {
            for (int c = 0; c < nc; ++c) {           // np.max / np.argmax: first maximum wins
                const float s = o[(long long)(4 + c) * A + a];
                if (s > best) { best = s; bc = c; }
            }
        }
        const bool flag = (a < A) && (best > conf);
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane], incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            s_warp[lane] = incl - v;                 // exclusive prefix over warps
        }
        __syncthreads();
        const int base = s_base;
        if (flag) {
            const int pos = base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
            const float cx = o[a], cy = o[(long long)A + a], w = o[2LL * A + a], h = o[3LL * A + a];
            const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);     // width / 2
            float x1 = __fsub_rn(cx, hw), y1 = __fsub_rn(cy, hh), x2 = __fadd_rn(cx, hw), y2 = __fadd_rn(cy, hh);
            x1 = __fdiv_rn(__fsub_rn(x1, padw), ratio);
            x2 = __fdiv_rn(__fsub_rn(x2, padw), ratio);
            y1 = __fdiv_rn(__fsub_rn(y1, padh), ratio);
            y2 = __fdiv_rn(__fsub_rn(y2, padh), ratio);
            Cand c;
            c.x1 = fminf(fmaxf(x1, 0.f), fw); c.x2 = fminf(fmaxf(x2, 0.f), fw);
            c.y1 = fminf(fmaxf(y1, 0.f), fh); c.y2 = fminf(fmaxf(y2, 0.f), fh);
            c.score = best; c.cls = bc;
            out[pos] = c;
        }
        __syncthreads();
        if (tid == 1023) s_base = base + s_warp[31] + __popc(bal);   // last warp: prefix + own count
        __syncthreads();
    }
    if (tid == 0) n_cand[b] = s_base;
}

// ---------------------------------------------------------------------------------------------
// NMS: one block per image.  Sort keys (class asc, score desc, candidate index desc) with a
// shared-memory bitonic network, then the reference's greedy rule (nms_numpy, e2e.py:89-119): in sorted
// order a box is kept iff no EARLIER KEPT box of its class has iou > thr with it.
//   n <= NMS_BITMASK_MAX: IoU-bitmask form.  The sorted boxes are staged in shared memory; for 32 rows at
//     a time the 32 warps compute the suppression bits of their row against every later box (one IoU per
//     lane, one ballot per 32 boxes) into a shared-memory tile; warp 0 then walks the 32 rows serially,
//     OR-ing the rows of kept boxes into the "removed" bitmap that its lanes hold in registers.  Cost:
//     n^2/2 IoUs fully parallel + n serial steps, no barrier per kept box (the conf = 0.001 mAP pass of
//     e2e.py:987 keeps 10^2..10^3 boxes per frame).
//   larger n: the greedy loop with one block barrier per kept box (n * K IoUs instead of n^2 / 2).
// Identical float32 arithmetic in both (explicit _rn operations, IEEE division), identical results.
// ---------------------------------------------------------------------------------------------
constexpr int NMS_BITMASK_MAX = 4096;

__device__ __forceinline__ bool nms_suppresses(const float4 bi, const float area_i, const float4 bj, const float iou_thr) {
    const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
    const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    const float den = __fadd_rn(__fsub_rn(__fadd_rn(area_i, area_j), inter), 1e-6f);
    return __fdiv_rn(inter, den) > iou_thr;
}
__device__ __forceinline__ unsigned float_sortable(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // ascending unsigned == ascending float
}

__global__ void __launch_bounds__(1024) nms_kernel(const Cand* __restrict__ cands, const int* __restrict__ n_cand, int A,
                                                   float iou_thr, int max_det, float* __restrict__ boxes,
                                                   float* __restrict__ scores, long long* __restrict__ classes,
                                                   int* __restrict__ keep_idx, int* __restrict__ counts, int sort_cap) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* removed = smem + (size_t)sort_cap * 8;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = n_cand[b];
    const Cand* cd = cands + (long long)b * A;
    if (n == 0) { if (tid == 0) counts[b] = 0; return; }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = tid; i < np2; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const Cand c = cd[i];
            k = ((unsigned long long)(unsigned)c.cls << 46) |
                ((unsigned long long)(0xffffffffu - float_sortable(c.score)) << 14) |
                (unsigned long long)(0x3fff - i);
        }
        keys[i] = k;
        removed[i] = 0;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > c) == up) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int nw = (n + 31) >> 5;
    const size_t box_off = ((size_t)np2 * 8 + 15) & ~(size_t)15;
    if (n <= NMS_BITMASK_MAX && box_off + (size_t)n * 24 + (size_t)32 * nw * 4 + 16 <= (size_t)sort_cap * 9) {
        // ---- IoU-bitmask NMS.  Shared memory behind the np2 sort keys: sorted boxes (x1,y1,x2,y2) | class | original index
        //      | 32-row bit tile | kept count
        float4* sbox = reinterpret_cast<float4*>(smem + box_off);
        int* scls = reinterpret_cast<int*>(sbox + n);
        int* sidx = scls + n;
        unsigned* tile = reinterpret_cast<unsigned*>(sidx + n);          // [32][nw]
        int* s_kept = reinterpret_cast<int*>(tile + 32 * nw);           // [1]
        for (int i = tid; i < n; i += blockDim.x) {
            const unsigned long long ki = keys[i];
            const int idx = 0x3fff - (int)(ki & 0x3fff);
            const Cand c = cd[idx];
            sbox[i] = make_float4(c.x1, c.y1, c.x2, c.y2);
            scls[i] = c.cls;
            sidx[i] = idx;
        }
        if (tid == 0) *s_kept = 0;
        __syncthreads();
        const int lane = tid & 31, wid = tid >> 5;
        unsigned rem[NMS_BITMASK_MAX / 32 / 32];                         // warp 0: lane l holds words l, l + 32, ...
#pragma unroll
        for (int k = 0; k < NMS_BITMASK_MAX / 1024; ++k) rem[k] = 0u;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + wid;                                       // this warp's row
            if (i < n) {
                const float4 bi = sbox[i];
                const int cls_i = scls[i];
                const float area_i = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
                for (int wj = i0 >> 5; wj < nw; ++wj) {
                    const int j = (wj << 5) + lane;
                    bool s = false;
                    if (j > i && j < n && scls[j] == cls_i) s = nms_suppresses(bi, area_i, sbox[j], iou_thr);
                    const unsigned bits = __ballot_sync(0xffffffffu, s);
                    if (lane == 0) tile[wid * nw + wj] = bits;
                }
            }
            __syncthreads();
            if (wid == 0) {
                int kept = *s_kept;
                const int rows = min(32, n - i0);
                for (int r = 0; r < rows; ++r) {
                    const int ii = i0 + r;
                    const int w_ = ii >> 5;
                    // removed bit of box ii lives in lane (w_ & 31), register (w_ >> 5)
                    unsigned word = 0u;
#pragma unroll
                    for (int k = 0; k < NMS_BITMASK_MAX / 1024; ++k) if (k == (w_ >> 5)) word = rem[k];
                    word = __shfl_sync(0xffffffffu, word, w_ & 31);
                    if ((word >> (ii & 31)) & 1u) continue;
                    if (lane == 0 && kept < max_det) {
                        const long long o = (long long)b * max_det + kept;
                        const float4 bx = sbox[ii];
                        const int idx = sidx[ii];
                        boxes[o * 4 + 0] = bx.x; boxes[o * 4 + 1] = bx.y; boxes[o * 4 + 2] = bx.z; boxes[o * 4 + 3] = bx.w;
                        scores[o] = cd[idx].score; classes[o] = (long long)scls[ii]; keep_idx[o] = idx;
                    }
                    ++kept;
#pragma unroll
                    for (int k = 0; k < NMS_BITMASK_MAX / 1024; ++k) {
                        const int wj = k * 32 + lane;
                        if (wj >= (i0 >> 5) && wj < nw) rem[k] |= tile[r * nw + wj];
                    }
                }
                if (lane == 0) *s_kept = kept;
            }
            __syncthreads();
        }
        if (tid == 0) counts[b] = *s_kept;
        return;
    }
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (removed[i]) continue;                       // uniform: written before the last barrier
        const unsigned long long ki = keys[i];
        const int idx_i = 0x3fff - (int)(ki & 0x3fff);
        const unsigned cls_i = (unsigned)(ki >> 46);
        const Cand ci = cd[idx_i];
        if (tid == 0 && kept < max_det) {
            const long long o = (long long)b * max_det + kept;
            boxes[o * 4 + 0] = ci.x1; boxes[o * 4 + 1] = ci.y1; boxes[o * 4 + 2] = ci.x2; boxes[o * 4 + 3] = ci.y2;
            scores[o] = ci.score; classes[o] = (long long)ci.cls; keep_idx[o] = idx_i;
        }
        ++kept;
        const float area_i = __fmul_rn(__fsub_rn(ci.x2, ci.x1), __fsub_rn(ci.y2, ci.y1));
        bool any = false;
        for (int j = i + 1 + tid; j < n; j += blockDim.x) {
            if (removed[j]) continue;
            const unsigned long long kj = keys[j];
            if ((unsigned)(kj >> 46) != cls_i) continue;
            const Cand cj = cd[0x3fff - (int)(kj & 0x3fff)];
            const float xx1 = fmaxf(ci.x1, cj.x1), yy1 = fmaxf(ci.y1, cj.y1);
            const float xx2 = fminf(ci.x2, cj.x2), yy2 = fminf(ci.y2, cj.y2);
            const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
            const float inter = __fmul_rn(w, h);
            const float area_j = __fmul_rn(__fsub_rn(cj.x2, cj.x1), __fsub_rn(cj.y2, cj.y1));
            const float den = __fadd_rn(__fsub_rn(__fadd_rn(area_i, area_j), inter), 1e-6f);
            if (__fdiv_rn(inter, den) > iou_thr) { removed[j] = 1; any = true; }
        }
        (void)any;
        __syncthreads();
    }
    if (tid == 0) counts[b] = kept;
}

extern "C" size_t lp_decode_nms_scratch_bytes(int batch, int n_anchors) {
    return (size_t)batch * n_anchors * sizeof(Cand) + 256;
}

extern "C" int lp_decode_nms(lp_ctx* ctx, const float* out0, int nc, int n_anchors, const int32_t* h_h,
                             const int32_t* w_h, const float* ratio_h, const float* pad_h, int batch, float conf,
                             float iou, int max_det, float* boxes, float* scores, int64_t* classes,
                             int32_t* keep_idx, int32_t* counts, int32_t* n_cand, void* scratch,
                             size_t scratch_bytes, void* stream) {
    LP_CHECK(ctx && out0 && h_h && w_h && ratio_h && pad_h && boxes && scores && classes && keep_idx && counts && n_cand && scratch,
             "lp_decode_nms: null argument");
    lp_device_guard dev_guard(ctx);
    LP_CHECK(nc >= 1 && n_anchors >= 1 && n_anchors <= 16384, "lp_decode_nms: nc/n_anchors out of range (anchors <= 16384)");
    LP_CHECK(max_det >= 1, "lp_decode_nms: max_det must be >= 1");
    LP_CHECK(scratch_bytes >= lp_decode_nms_scratch_bytes(batch, n_anchors), "lp_decode_nms: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    int cap = 1;
    while (cap < n_anchors) cap <<= 1;
    const size_t smem = (size_t)cap * 8 + cap;
    if (!(ctx->attr_set & 2)) {          // per context (= per device): the opt-in is a per-device function attribute
        LP_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 9));
        ctx->attr_set |= 2;
    }
    Cand* cands = (Cand*)scratch;
    for (int base = 0; base < batch; base += LP_MAX_TABLE) {
        const int n = batch - base < LP_MAX_TABLE ? batch - base : LP_MAX_TABLE;
        ImageTable tab;
        for (int i = 0; i < n; ++i) {
            tab.h[i] = h_h[base + i]; tab.w[i] = w_h[base + i];
            tab.ratio[i] = ratio_h[base + i];
            tab.padw[i] = pad_h[2 * (base + i)]; tab.padh[i] = pad_h[2 * (base + i) + 1];
        }
        candidates_kernel<<<n, 1024, 0, st>>>(out0 + (size_t)base * (4 + nc) * n_anchors, nc, n_anchors, tab, conf,
                                              cands + (size_t)base * n_anchors, n_cand + base);
        LP_LAUNCH_OK(ctx);
        nms_kernel<<<n, 1024, smem, st>>>(cands + (size_t)base * n_anchors, n_cand + base, n_anchors, iou, max_det,
                                          boxes + (size_t)base * max_det * 4, scores + (size_t)base * max_det,
                                          (long long*)classes + (size_t)base * max_det, keep_idx + (size_t)base * max_det,
                                          counts + base, cap);
        LP_LAUNCH_OK(ctx);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// ROI select (e2e.py:459-475): astype(int) truncation, clip, area filter, ordered compaction.
// ---------------------------------------------------------------------------------------------
struct SizeTable {
    int h[LP_MAX_TABLE];
    int w[LP_MAX_TABLE];
};

__global__ void __launch_bounds__(1024) roi_select_kernel(const float* __restrict__ boxes, const int* __restrict__ counts,
                                                          int max_det, SizeTable tab, int batch, int img_base,
                                                          int min_area, int max_rois, int mode, int* __restrict__ roi_xyxy,
                                                          int* __restrict__ roi_src, int* __restrict__ n_rois) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    __shared__ int s_first[LP_MAX_TABLE + 1];           // exclusive prefix of min(counts, max_det) over images
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_base = (img_base == 0) ? 0 : *n_rois;
        int acc = 0;
        for (int i = 0; i < batch; ++i) { s_first[i] = acc; acc += min(counts[i], max_det); }
        s_first[batch] = acc;
    }
    __syncthreads();
    const int total = s_first[batch];                   // detections actually present (not batch * max_det slots)
    for (int s0 = 0; s0 < total; s0 += 1024) {
        const int s = s0 + tid;
        bool flag = false;
        int x1 = 0, y1 = 0, x2 = 0, y2 = 0, img = 0, k = 0;
        if (s < total) {
            int lo = 0, hi = batch;                     // image of detection s: last i with s_first[i] <= s
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_first[mid] <= s) lo = mid; else hi = mid; }
            img = lo; k = s - s_first[img];
            const float* bx = boxes + ((long long)img * max_det + k) * 4;
            const int w = tab.w[img], h = tab.h[img];
            x1 = (int)bx[0]; y1 = (int)bx[1]; x2 = (int)bx[2]; y2 = (int)bx[3];   // trunc toward zero
            if (mode == 0) {                            // e2e.py:462-469
                x1 = min(max(x1, 0), w - 1); y1 = min(max(y1, 0), h - 1);
                x2 = min(max(x2, x1 + 1), w); y2 = min(max(y2, y1 + 1), h);
            } else {                                    // e2e_optimize.py:480-483: plain clip to the frame
                x1 = min(max(x1, 0), w); x2 = min(max(x2, 0), w);
                y1 = min(max(y1, 0), h); y2 = min(max(y2, 0), h);
            }
            const long long area = (long long)(x2 - x1) * (y2 - y1);
            flag = (area >= (long long)min_area) && x2 > x1 && y2 > y1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane], incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            s_warp[lane] = incl - v;
        }
        __syncthreads();
        const int base = s_base;
        if (flag) {
            const int pos = base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
            if (pos < max_rois) {
                roi_xyxy[pos * 4 + 0] = x1; roi_xyxy[pos * 4 + 1] = y1; roi_xyxy[pos * 4 + 2] = x2; roi_xyxy[pos * 4 + 3] = y2;
                roi_src[pos * 2 + 0] = img_base + img; roi_src[pos * 2 + 1] = k;
            }
        }
        __syncthreads();
        if (tid == 1023) s_base = base + s_warp[31] + __popc(bal);
        __syncthreads();
    }
    if (tid == 0) *n_rois = s_base;
}

extern "C" int lp_roi_select(lp_ctx* ctx, const float* boxes, const int32_t* counts, int max_det, const int32_t* h_h,
                             const int32_t* w_h, int batch, int min_area, int max_rois, int32_t* roi_xyxy,
                             int32_t* roi_src, int32_t* n_rois, void* stream) {
    LP_CHECK(ctx && boxes && counts && h_h && w_h && roi_xyxy && roi_src && n_rois, "lp_roi_select: null argument");
    lp_device_guard dev_guard(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    if (batch == 0) { LP_CUDA(cudaMemsetAsync(n_rois, 0, 4, st)); return 0; }
    for (int base = 0; base < batch; base += LP_MAX_TABLE) {
        const int n = batch - base < LP_MAX_TABLE ? batch - base : LP_MAX_TABLE;
        SizeTable tab;
        for (int i = 0; i < n; ++i) { tab.h[i] = h_h[base + i]; tab.w[i] = w_h[base + i]; }
        roi_select_kernel<<<1, 1024, 0, st>>>(boxes + (size_t)base * max_det * 4, counts + base, max_det, tab, n, base,
                                              min_area, max_rois, ctx->roi_mode, roi_xyxy, roi_src, n_rois);
        LP_LAUNCH_OK(ctx);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Records for the multi-GPU gather: 9 x int32 per detection.
// ---------------------------------------------------------------------------------------------
__global__ void pack_records_kernel(const int* __restrict__ roi_src, const int* __restrict__ frame_ids,
                                    const float* __restrict__ boxes, const float* __restrict__ scores,
                                    const long long* __restrict__ classes, int max_det,
                                    const long long* __restrict__ cls_argmax, const float* __restrict__ probs,
                                    int n_classes, int n_rois, const int* __restrict__ n_dev, int* __restrict__ rec) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rois || (n_dev && r >= *n_dev)) return;
    const int img = roi_src[2 * r], k = roi_src[2 * r + 1];
    const long long o = (long long)img * max_det + k;
    int* q = rec + (long long)r * 9;
    q[0] = frame_ids ? frame_ids[img] : img;
    q[1] = __float_as_int(boxes[o * 4 + 0]); q[2] = __float_as_int(boxes[o * 4 + 1]);
    q[3] = __float_as_int(boxes[o * 4 + 2]); q[4] = __float_as_int(boxes[o * 4 + 3]);
    q[5] = __float_as_int(scores[o]);
    q[6] = (int)classes[o];
    const long long c = cls_argmax[r];
    q[7] = (int)c;
    q[8] = __float_as_int(probs[(long long)r * n_classes + c]);
}

extern "C" int lp_pack_records(lp_ctx* ctx, const int32_t* roi_src, const int32_t* frame_ids, const float* boxes,
                               const float* scores, const int64_t* classes, int max_det, const int64_t* cls_argmax,
                               const float* probs, int n_classes, int n_rois, int32_t* records, void* stream) {
    LP_CHECK(ctx && roi_src && boxes && scores && classes && cls_argmax && probs && records, "lp_pack_records: null argument");
    lp_device_guard dev_guard(ctx);
    if (n_rois <= 0) return 0;
    pack_records_kernel<<<(n_rois + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        roi_src, frame_ids, boxes, scores, (const long long*)classes, max_det, (const long long*)cls_argmax, probs,
        n_classes, n_rois, ctx->roi_count_dev, records);
    LP_LAUNCH_OK(ctx);
    return 0;
}
