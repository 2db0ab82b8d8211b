// Shared declarations for the litepi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/litepi_b200.h"

#define LP_MAX_TABLE 64   // per-image metadata entries carried as kernel arguments per launch

void lp_set_error(const char* fmt, ...);

#define LP_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            lp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return -2;                                                                        \
        }                                                                                     \
    } while (0)

#define LP_CHECK(cond, ...)                  \
    do {                                     \
        if (!(cond)) {                       \
            lp_set_error(__VA_ARGS__);       \
            return -1;                       \
        }                                    \
    } while (0)

#define LP_LAUNCH_OK(ctx)                                                                \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            lp_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

struct lp_net_plan {
    std::vector<lp_buf_desc> bufs;
    std::vector<lp_op_desc> ops;
    const float* weights = nullptr;
    size_t n_floats = 0;
    const uint8_t* weights_tc = nullptr;
    size_t tc_bytes = 0;
    int max_batch = 0;
    size_t workspace_bytes = 0;     // max over bufs of offset + max_batch * image_bytes
    bool loaded = false;
    std::vector<int> small_slot;    // per op: >= 0 if the small-channel conv path (weights as kernel parameters) covers it
    std::vector<std::vector<float>> small_host;   // per op: host copy [weights | bias] for those ops
    std::vector<int8_t> last_path;  // per op, which kernel family ran it last (lp_op_paths): 0 generic SIMT, 1 parameter-weight
                                    // small conv, 2 tcgen05 conv, 3 absorbed by the previous op's kernel, 4 warp-level MMA conv, 5 fused C2f body
};

struct lp_fused_cls;                // fused ShuffleNetV2 program of a context (shufflenet_fused.cu)
void lp_fused_free(lp_fused_cls* f);

struct lp_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    int use_tc = 1;
    lp_net_plan nets[2];
    // probe: CUDA events around one op of one plan (bench.py roofline of the dominant kernel)
    int probe_net = -1, probe_op = -1;
    std::vector<cudaEvent_t> probe_ev;   // pairs (start, stop), ring
    int probe_n = 0;
    lp_fused_cls* fused = nullptr;   // owned: fused-classifier program (lp_fused_classifier_load), freed by lp_destroy
    int attr_set = 0;                // bit per kernel family whose dynamic shared-memory opt-in was made on ctx->device
    int use_fused = 1;
    int use_mma = 1;                 // warp-level MMA kernels for the small-channel layers (env LP_NO_MMA=1: fp32-FMA kernels instead)
    int use_c2f = 1;                 // fused C2f-body kernel for c = 8 / 16 (env LP_NO_C2F=1: layer by layer)
    int tc_tma = 7;                  // conv_tc patch loads by TMA tensor copies: bit 0 = 3x3 stride 1, bit 1 = 1x1, bit 2 = 3x3 stride 2 (env LP_TC_TMA, read by lp_create)
    int tc_sw128 = 1;                // SWIZZLE_128B A tiles for the 1x1 layers with cin % 64 == 0 (env LP_TC_SW128)
    int use_pdl = 1;                 // programmatic dependent launch between tensor-core conv kernels (env LP_NO_PDL=1 disables)
    int roi_mode = 0;                // 0: e2e.py ROI rules + Pillow resize; 1: e2e_optimize.py rules + cv2 INTER_LINEAR
    const int* roi_count_dev = nullptr;   // lp_set_roi_count_device: ROI-side calls take their count from the device
    long long* tc_dbg = nullptr;     // device buffer (16 x int64) for conv_tc role timing; debugging only
};
#define LP_PROBE_RING 512

// Entry points run on the context's device whatever the caller's current device is, and restore it on return.
struct lp_device_guard {
    int prev = -1;
    explicit lp_device_guard(const lp_ctx* c) {
        if (c && cudaGetDevice(&prev) == cudaSuccess && prev != c->device) cudaSetDevice(c->device); else prev = -1;
    }
    ~lp_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- OpenCV 8-bit INTER_LINEAR coefficient of output index d (11-bit fixed point; see preprocess.cu)
__device__ __forceinline__ void lin_coef(int d, double scale, int n, bool clamp_coef, int& s, int& c0, int& c1) {
    // explicit _rn ops: no FMA contraction, or the double result rounds differently from the CPU
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    int si = (int)floorf(f);
    f -= (float)si;
    if (clamp_coef) {
        if (si < 0) { si = 0; f = 0.f; }
        if (si >= n - 1) { si = n - 1; f = 0.f; }
    }
    s = si;
    c0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c1 = (int)rintf(__fmul_rn(f, 2048.f));
}

// ---- SiLU / sigmoid in five instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL).  __expf / __fdividef wrap the same two MUFU
// ops in range fix-ups (3 more instructions per value) that cannot change the result here: where ex2 would be denormal,
// 1 + e rounds to 1 either way.  The epilogues apply this to every activation of the network.
__device__ __forceinline__ float lp_sigmoid(float v) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
__device__ __forceinline__ float lp_silu(float v) { return v * lp_sigmoid(v); }

// ---- split-f16 helpers -------------------------------------------------------
__device__ __forceinline__ float split_load(const __half* hi, const __half* lo, size_t i) {
    return __half2float(hi[i]) + __half2float(lo[i]);
}
__device__ __forceinline__ void split_make(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// ---- layer-kernel parameter blocks (conv_simt.cu, conv_mma.cu) ----------------
struct TensorRef {
    const void* base;        // SPLIT16: hi plane; F32/U8: the data
    long long plane;         // SPLIT16: element offset from hi to lo plane
    long long img;           // elements per image (per plane)
    int C;                   // total channels of the buffer
    int coff;                // first channel of the view
    int fmt;
};

struct ConvParams {
    TensorRef in, out, res;  // res.base == nullptr -> no residual
    int cin, cout, out_cstride;
    int cout_real, seg_l0, seg_len, seg_pad;   // segmented destination (lp_op_desc.out_seg_len); seg_len == 0: plain
    int H, W, Ho, Wo;        // input / output spatial size
    int ksize, stride, act;
    int res_first;           // LP_OPF_RES_BEFORE_ACT: act(conv + bias + residual)
    int n_img;
    const float* w;          // [tap][cin][cout]
    const float* bias;       // [cout]
    float in_scale_mean, in_scale_std;   // STEM_U8: x = (u8/255 - mean)/std ; detector: mean 0, std 1
};

// warp-level tensor-core path for the small-channel layers (conv_mma.cu): returns 1 if it ran the op (and, when *fused_next,
// the 1x1 conv that follows it), 0 if the shape is not covered
int lp_conv_mma_try(lp_ctx* ctx, const ConvParams& p, const ConvParams* post, cudaStream_t st);
int lp_stem_mma_try(lp_ctx* ctx, const ConvParams& p, cudaStream_t st);
int lp_stem_conv_try(lp_ctx* ctx, const ConvParams& ps, const ConvParams& p, const ConvParams* post, cudaStream_t st);

// fused C2f body (c2f_mma.cu): number of plan ops covered starting at op `oi` (0: pattern / shape not covered, < 0: error)
int lp_c2f_fused_try(lp_ctx* ctx, lp_net_plan& net, size_t oi, int batch, uint8_t* ws, cudaStream_t st);

// kernels implemented per translation unit (host launchers)
int lp_run_plan(lp_ctx* ctx, lp_net_plan& net, const uint8_t* in, int batch, void* workspace,
                size_t workspace_bytes, float* logits_or_head, cudaStream_t st);
int lp_launch_detect_tail(lp_ctx* ctx, const float* head_raw, int batch, int head_c, int in_size, int nc, int n_anchors,
                          float* out0, cudaStream_t st);

// tensor-core path (conv_tc.cu); returns 1 if it handled the op, 0 if not applicable, <0 on error
int lp_assign_small_slots(lp_net_plan& net, cudaStream_t st);
int lp_fused_classify(lp_ctx* ctx, const uint8_t* in, int n, float* logits, cudaStream_t st);
int lp_conv_tc_try(lp_ctx* ctx, lp_net_plan& net, const lp_op_desc& op, int batch, uint8_t* ws, cudaStream_t st);
