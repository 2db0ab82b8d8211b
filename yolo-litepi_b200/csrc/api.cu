// C-ABI entry points that are not tied to one kernel file: context lifetime, plan registration,
// detector / classifier forward.  See include/litepi_b200.h for the contract.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void lp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int lp_launch_softmax_argmax(lp_ctx* ctx, const float* logits, int n, int C, float* probs, int64_t* argmax, cudaStream_t st);

extern "C" const char* lp_last_error(void) { return g_err; }
extern "C" int lp_abi_version(void) { return LP_ABI_VERSION; }

extern "C" int lp_create(lp_ctx** out, int device) {
    LP_CHECK(out != nullptr, "lp_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        lp_set_error("lp_create: no CUDA device (%s); litepi_b200 has no CPU fallback", cudaGetErrorString(e));
        return -3;
    }
    LP_CHECK(device >= 0 && device < n, "lp_create: device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;                 // the caller's current device is left alone
    LP_CUDA(cudaGetDeviceProperties(&prop, device));
    LP_CHECK(prop.major == 10, "lp_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
             prop.major, prop.minor);
    lp_ctx* c = new lp_ctx();
    c->device = device;
    { const char* e = getenv("LP_NO_PDL"); c->use_pdl = (e && e[0] == '1') ? 0 : 1; }
    { const char* e = getenv("LP_NO_MMA"); c->use_mma = (e && e[0] == '1') ? 0 : 1; }
    { const char* e = getenv("LP_NO_C2F"); c->use_c2f = (e && e[0] == '1') ? 0 : 1; }
    { const char* e = getenv("LP_TC_TMA"); c->tc_tma = e ? atoi(e) : 7; }
    { const char* e = getenv("LP_TC_SW128"); c->tc_sw128 = e ? atoi(e) : 1; }
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return 0;
}

extern "C" int lp_destroy(lp_ctx* ctx) {
    if (ctx) {
        for (cudaEvent_t e : ctx->probe_ev) cudaEventDestroy(e);
        lp_fused_free(ctx->fused);
    }
    delete ctx;
    return 0;
}

extern "C" int lp_debug_tc_timing(lp_ctx* ctx, void* dev_buf16) {
    LP_CHECK(ctx, "lp_debug_tc_timing: null ctx");
    ctx->tc_dbg = (long long*)dev_buf16;
    return 0;
}

extern "C" int lp_probe_set(lp_ctx* ctx, int net, int op_index) {
    LP_CHECK(ctx, "lp_probe_set: null ctx");
    if (ctx->probe_ev.empty() && (op_index >= 0 || op_index == -2)) {
        ctx->probe_ev.resize(2 * LP_PROBE_RING);
        for (auto& e : ctx->probe_ev) LP_CUDA(cudaEventCreate(&e));
    }
    ctx->probe_net = net; ctx->probe_op = op_index; ctx->probe_n = 0;
    return 0;
}

extern "C" int lp_probe_read(lp_ctx* ctx, float* ms_h, int cap) {
    LP_CHECK(ctx && ms_h, "lp_probe_read: null argument");
    const int n = ctx->probe_n < LP_PROBE_RING ? ctx->probe_n : LP_PROBE_RING;
    int m = 0;
    for (int i = 0; i < n && m < cap; ++i, ++m) {
        LP_CUDA(cudaEventSynchronize(ctx->probe_ev[2 * i + 1]));
        LP_CUDA(cudaEventElapsedTime(&ms_h[m], ctx->probe_ev[2 * i], ctx->probe_ev[2 * i + 1]));
    }
    return m;
}

extern "C" int lp_set_fused_classifier(lp_ctx* ctx, int enable) {
    LP_CHECK(ctx, "lp_set_fused_classifier: null ctx");
    ctx->use_fused = enable ? 1 : 0;
    return 0;
}

extern "C" int lp_op_paths(lp_ctx* ctx, int net, int8_t* out_h, int cap) {
    LP_CHECK(ctx && out_h && (net == 0 || net == 1), "lp_op_paths: bad argument");
    const lp_net_plan& P = ctx->nets[net];
    int n = 0;
    for (; n < (int)P.last_path.size() && n < cap; ++n) out_h[n] = P.last_path[n];
    return n;
}

extern "C" int lp_set_pdl(lp_ctx* ctx, int enable) {
    LP_CHECK(ctx, "lp_set_pdl: null ctx");
    ctx->use_pdl = enable ? 1 : 0;
    return 0;
}

extern "C" int lp_set_tensor_core(lp_ctx* ctx, int enable) {
    LP_CHECK(ctx, "lp_set_tensor_core: null ctx");
    ctx->use_tc = enable ? 1 : 0;
    return 0;
}

extern "C" int lp_sm_count(lp_ctx* ctx) { return ctx ? ctx->sm_count : -1; }
extern "C" int64_t lp_launch_count(lp_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" size_t lp_workspace_bytes(lp_ctx* ctx, int net) {
    if (!ctx || net < 0 || net > 1 || !ctx->nets[net].loaded) return 0;
    return ctx->nets[net].workspace_bytes;
}

extern "C" int lp_net_load(lp_ctx* ctx, int net, const lp_buf_desc* bufs_h, int n_bufs, const lp_op_desc* ops_h,
                           int n_ops, const float* weights, size_t n_floats, const void* weights_tc, size_t tc_bytes,
                           int max_batch) {
    LP_CHECK(ctx && bufs_h && ops_h && weights, "lp_net_load: null argument");
    lp_device_guard dev_guard(ctx);
    LP_CHECK(net == LP_NET_DETECTOR || net == LP_NET_CLASSIFIER, "lp_net_load: bad net id %d", net);
    LP_CHECK(n_bufs > 0 && n_ops > 0 && max_batch > 0, "lp_net_load: empty plan");
    lp_net_plan& P = ctx->nets[net];
    P.bufs.assign(bufs_h, bufs_h + n_bufs);
    P.ops.assign(ops_h, ops_h + n_ops);
    P.weights = weights; P.n_floats = n_floats;
    P.weights_tc = (const uint8_t*)weights_tc; P.tc_bytes = tc_bytes;
    P.max_batch = max_batch;
    size_t need = 0;
    for (int i = 0; i < n_bufs; ++i) {
        const lp_buf_desc& b = P.bufs[i];
        LP_CHECK(b.h > 0 && b.w > 0 && b.c > 0 && b.offset >= 0 && b.offset % 1024 == 0, "lp_net_load: buffer %d malformed", i);
        const int esz = b.fmt == LP_FMT_SPLIT16 ? 2 : (b.fmt == LP_FMT_F32 ? 4 : 1);
        LP_CHECK(b.image_bytes == (int64_t)b.h * b.w * b.c * esz, "lp_net_load: buffer %d image_bytes mismatch", i);
        LP_CHECK(b.fmt != LP_FMT_SPLIT16 || b.c % 8 == 0, "lp_net_load: split-f16 buffer %d needs channels %% 8 == 0", i);
        const size_t planes = b.fmt == LP_FMT_SPLIT16 ? 2 : 1;
        const size_t end = (size_t)b.offset + planes * (size_t)max_batch * (size_t)b.image_bytes;
        if (b.fmt != LP_FMT_U8 && end > need) need = end;     // the U8 input image is caller-owned
    }
    for (int i = 0; i < n_ops; ++i) {
        const lp_op_desc& o = P.ops[i];
        LP_CHECK(o.in_buf >= 0 && o.in_buf < n_bufs, "lp_net_load: op %d in_buf out of range", i);
        LP_CHECK(o.kind == LP_OP_MEAN_FC || (o.out_buf >= 0 && o.out_buf < n_bufs), "lp_net_load: op %d out_buf out of range", i);
        LP_CHECK(o.res_buf < n_bufs, "lp_net_load: op %d res_buf out of range", i);
        LP_CHECK(o.in_coff >= 0 && o.in_coff + o.cin <= P.bufs[o.in_buf].c, "lp_net_load: op %d input slice exceeds buffer", i);
        if (o.kind != LP_OP_MEAN_FC) {
            const int cs = o.out_cstride > 0 ? o.out_cstride : 1;
            const int n_out = o.cout_real > 0 ? o.cout_real : o.cout;
            if (o.out_seg_len > 0) {
                const int l = o.out_coff + (n_out - 1) * cs;
                LP_CHECK(o.out_seg_pad >= o.out_seg_len && (l / o.out_seg_len) * o.out_seg_pad + l % o.out_seg_len < P.bufs[o.out_buf].c,
                         "lp_net_load: op %d segmented output exceeds buffer", i);
            } else {
                LP_CHECK(o.out_coff >= 0 && o.out_coff + (n_out - 1) * cs < P.bufs[o.out_buf].c, "lp_net_load: op %d output slice exceeds buffer", i);
            }
        }
        const size_t wn = o.kind == LP_OP_CONV || o.kind == LP_OP_STEM_U8 ? (size_t)o.ksize * o.ksize * o.cin * o.cout
                          : o.kind == LP_OP_DWCONV3 ? (size_t)o.ksize * o.ksize * o.cout
                          : o.kind == LP_OP_MEAN_FC ? (size_t)o.cin * o.cout : 0;
        if (wn) LP_CHECK(o.w_off >= 0 && (size_t)o.w_off + wn <= n_floats && o.b_off >= 0 && (size_t)o.b_off + o.cout <= n_floats,
                         "lp_net_load: op %d weights outside the blob", i);
    }
    P.workspace_bytes = need;
    P.loaded = true;
    if (net == LP_NET_DETECTOR) return lp_assign_small_slots(P, 0);     // classifier: the fused kernel is the product path
    return 0;
}

// Detector: plan (stem .. Detect convs) then the Detect tail.  The last buffer of the plan is the
// head buffer [A][HC] f32 written by the six final 1x1 convs.
extern "C" int lp_detect_forward(lp_ctx* ctx, const uint8_t* in, int batch, void* workspace, size_t workspace_bytes,
                                 float* out0, void* stream) {
    LP_CHECK(ctx && in && workspace && out0, "lp_detect_forward: null argument");
    lp_device_guard dev_guard(ctx);
    lp_net_plan& P = ctx->nets[LP_NET_DETECTOR];
    LP_CHECK(P.loaded, "lp_detect_forward: detector not loaded");
    cudaStream_t st = (cudaStream_t)stream;
    int r = lp_run_plan(ctx, P, in, batch, workspace, workspace_bytes, nullptr, st);
    if (r) return r;
    const lp_buf_desc& hb = P.bufs.back();
    LP_CHECK(hb.fmt == LP_FMT_F32 && hb.w == 1, "lp_detect_forward: last plan buffer is not the Detect head");
    // geometry from the plan: input size = the u8 image buffer of the stem, nc = width of the class-logit convs
    // (the ops that write the head buffer at channel 64), anchors = rows of the head buffer
    const int in_size = P.bufs[P.ops[0].in_buf].h;
    int nc = 0;
    for (const lp_op_desc& o : P.ops)
        if (o.kind == LP_OP_CONV && o.out_buf == (int)P.bufs.size() - 1 && o.out_coff == 64) nc = o.cout_real > 0 ? o.cout_real : o.cout;
    LP_CHECK(nc >= 1 && hb.c >= 64 + nc, "lp_detect_forward: head buffer has %d channels for nc=%d", hb.c, nc);
    const int tail_slot = (int)P.ops.size();
    const bool probe_tail = ctx->probe_net == LP_NET_DETECTOR && ctx->probe_op == -2 && !ctx->probe_ev.empty() && tail_slot < LP_PROBE_RING;
    if (probe_tail) LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * tail_slot], st));
    r = lp_launch_detect_tail(ctx, (const float*)((uint8_t*)workspace + hb.offset), batch, hb.c, in_size, nc, hb.h, out0, st);
    if (probe_tail) {                                  // probe slot n_ops = the Detect tail (bench.py stage table)
        LP_CUDA(cudaEventRecord(ctx->probe_ev[2 * tail_slot + 1], st));
        if (ctx->probe_n < tail_slot + 1) ctx->probe_n = tail_slot + 1;
    }
    return r;
}

extern "C" int lp_set_roi_mode(lp_ctx* ctx, int mode) {
    LP_CHECK(ctx && (mode == 0 || mode == 1), "lp_set_roi_mode: mode must be 0 (e2e.py) or 1 (e2e_optimize.py)");
    ctx->roi_mode = mode;
    return 0;
}

extern "C" int lp_set_roi_count_device(lp_ctx* ctx, const int32_t* n_rois_dev) {
    LP_CHECK(ctx, "lp_set_roi_count_device: null context");
    ctx->roi_count_dev = n_rois_dev;
    return 0;
}

extern "C" int lp_classify(lp_ctx* ctx, const uint8_t* in, int n, void* workspace, size_t workspace_bytes,
                           float* logits, float* probs, int64_t* argmax, void* stream) {
    LP_CHECK(ctx && workspace && logits && probs && argmax, "lp_classify: null argument");
    lp_device_guard dev_guard(ctx);
    if (n == 0) return 0;
    LP_CHECK(in != nullptr, "lp_classify: null input");
    lp_net_plan& P = ctx->nets[LP_NET_CLASSIFIER];
    LP_CHECK(P.loaded, "lp_classify: classifier not loaded");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = P.ops.back().cout;
    {
        const int r = lp_fused_classify(ctx, in, n, logits, st);       // whole network in one persistent kernel
        if (r < 0) return r;
        if (r == 1) return lp_launch_softmax_argmax(ctx, logits, n, C, probs, argmax, st);
    }
    LP_CHECK(ctx->roi_count_dev == nullptr, "lp_classify: a device-side ROI count needs the fused classifier");
    const size_t img_bytes = (size_t)P.bufs[P.ops[0].in_buf].h * P.bufs[P.ops[0].in_buf].w * 3;
    for (int base = 0; base < n; base += P.max_batch) {
        const int nb = n - base < P.max_batch ? n - base : P.max_batch;
        int r = lp_run_plan(ctx, P, in + (size_t)base * img_bytes, nb, workspace, workspace_bytes, logits + (size_t)base * C, st);
        if (r) return r;
    }
    return lp_launch_softmax_argmax(ctx, logits, n, C, probs, argmax, st);
}
