/* litepi_b200 -- C-ABI of the B200-native YOLO-LitePi hot path.
 *
 * The reference (vinhisreal/YOLO-LitePi) has NO FFI: its plugin surface is the
 * Python class API of src/vntsr/pipeline/e2e.py (NCNNDetector :195-316,
 * PyTorchClassifier :350-396, HybridPipeline :399-531) and every heavy op is a
 * call into a third-party runtime (ncnn / onnxruntime / torch / cv2 / Pillow).
 * This header is the boundary a new detector backend binds instead of those
 * runtimes: plain pointers and sizes, no torch types.  Each entry point cites
 * the reference call it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (lp_last_error() has the
 *     message); nothing falls back to the CPU.
 *   - pointers are DEVICE pointers unless the name ends in _h (host).  Per-image
 *     metadata tables (*_h) travel as kernel arguments, 64 images per launch, so
 *     no call does a host->device copy or a synchronisation of its own.
 *   - the library allocates no device memory: the caller (torch, in the Python
 *     host) owns weights, workspace and I/O buffers and passes them in.
 *   - one lp_ctx per process/GPU; not thread-safe; one call in flight.
 *   - `stream` is a cudaStream_t passed as void*.
 *
 * Activation format ("split-f16"): an NHWC tensor of C channels is stored as two
 * fp16 planes hi|lo with value = float(hi) + float(lo) (22-bit mantissa).  It is
 * what the tcgen05 implicit-GEMM kernels consume directly (3 MMAs per K-step:
 * Ahi*Bhi + Alo*Bhi + Ahi*Blo, fp32 accumulate in TMEM); DESIGN.md section 3 has
 * the precision experiment that rules out single-pass fp16/bf16/tf32.
 */
#ifndef LITEPI_B200_H
#define LITEPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lp_ctx lp_ctx;

#define LP_ABI_VERSION 5

/* ---- network plan (built on the host from the reference's model.ncnn.param) ---- */

enum lp_op_kind {
    LP_OP_STEM_U8   = 0,  /* conv kxk (3x3 s2; 7x7 s2 for ResNet) on the u8 letterboxed/ROI image, x/255 (and optional mean/std) fused */
    LP_OP_CONV      = 1,  /* conv kxk (k in {1,3}), stride 1|2, bias, act, optional residual add */
    LP_OP_DWCONV3   = 2,  /* depthwise k x k (3; 5 in EfficientNet) stride 1|2, bias, act (none in ShuffleNetV2, ReLU6 / SiLU in MobileNetV2 / EfficientNet) */
    LP_OP_MAXPOOL   = 3,  /* max pool k x k, stride s, pad k/2 (SPPF 5/1, ShuffleNetV2 3/2) */
    LP_OP_UPSAMPLE2 = 4,  /* nearest x2 (model.10 / model.13, model.ncnn.param:88,103) */
    LP_OP_COPY      = 5,  /* channel-slice copy (ShuffleNetV2 pass-through half) */
    LP_OP_MEAN_FC   = 6,  /* global mean over HxW + FC (torchvision shufflenetv2.py forward tail) */
    LP_OP_GLOBAL_MEAN = 7, /* [H,W,C] -> [1,1,C] mean (SqueezeExcitation.avgpool, torchvision ops/misc.py) */
    LP_OP_SCALE     = 8   /* out = in * gate, gate = the [1,1,C] map in res_buf (SqueezeExcitation scale) */
};

enum lp_act { LP_ACT_NONE = 0, LP_ACT_SILU = 1, LP_ACT_RELU = 2, LP_ACT_RELU6 = 3 /* MobileNetV2 */, LP_ACT_SIGMOID = 4 };
/* lp_op_desc.flags */
#define LP_OPF_RES_BEFORE_ACT 1   /* out = act(conv + bias + residual): torchvision BasicBlock (resnet.py forward); default is
                                     act(conv + bias) + residual, the Ultralytics Bottleneck / C2f shortcut */
enum lp_fmt { LP_FMT_SPLIT16 = 0, LP_FMT_F32 = 1, LP_FMT_U8 = 2 };

/* One NHWC activation buffer inside the caller-provided workspace. */
typedef struct lp_buf_desc {
    int32_t h, w, c;        /* spatial size and TOTAL channels (multiple of 8 for SPLIT16) */
    int32_t fmt;            /* lp_fmt */
    int64_t offset;         /* byte offset of image 0 (hi plane) inside the workspace (1024-B aligned) */
    int64_t image_bytes;    /* bytes per image per plane; SPLIT16: lo plane starts max_batch*image_bytes later */
} lp_buf_desc;

/* One fused layer.  Convs read channels [in_coff, in_coff+cin) of in_buf and
 * write channels out_coff + out_cstride*j (j < cout) of out_buf, so C2f / SPPF /
 * neck concats (model.ncnn.param cat_0..cat_12) and ShuffleNetV2's channel
 * shuffle are store patterns, never copies. */
typedef struct lp_op_desc {
    int32_t kind;                       /* lp_op_kind */
    int32_t in_buf, in_coff, cin;
    int32_t out_buf, out_coff, cout;
    int32_t out_cstride;                /* 1, or 2 for shuffle-interleaved stores */
    int32_t cout_real;                  /* outputs actually stored (<= cout; the rest is tensor-core N padding) */
    int32_t out_seg_len, out_seg_pad;   /* > 0: segmented destination.  Output j is LOGICAL channel
                                           l = out_coff + j*out_cstride and lands on physical channel
                                           (l / out_seg_len) * out_seg_pad + l % out_seg_len (ShuffleNetV2 halves
                                           of 58/116/232 channels padded to 64/128/256) */
    int32_t res_buf, res_coff;          /* residual added AFTER the activation (-1 = none) */
    int32_t ksize, stride, act;
    int32_t row_off;                    /* Detect head: first anchor row this level writes */
    int32_t flags;                      /* LP_OPF_* */
    float   in_mean, in_std;            /* STEM_U8: x = (u8/255 - in_mean)/in_std (detector: 0, 1) */
    int64_t w_off, b_off;               /* offsets (in floats) into the fp32 weight blob:
                                           weights [tap][cin][cout], bias [cout] */
    int64_t wtc_off;                    /* byte offset into the tensor-core weight blob, or -1 */
} lp_op_desc;

/* ---- lifetime ---- */
int  lp_abi_version(void);
const char* lp_last_error(void);
int  lp_create(lp_ctx** out, int device);
int  lp_destroy(lp_ctx* ctx);

/* Network kinds a context holds. */
enum lp_net { LP_NET_DETECTOR = 0, LP_NET_CLASSIFIER = 1 };

/* Replaces ncnn.Net.load_param/load_model (e2e.py:209-216) and
 * build_classifier + load_state_dict (e2e.py:320-347): register a plan.
 * `weights` = device fp32 blob, `weights_tc` = device blob of pre-split fp16
 * tensor-core operands (may be NULL: every conv then runs on the SIMT kernels). */
int lp_net_load(lp_ctx* ctx, int net, const lp_buf_desc* bufs_h, int n_bufs,
                const lp_op_desc* ops_h, int n_ops,
                const float* weights, size_t n_floats,
                const void* weights_tc, size_t tc_bytes, int max_batch);
/* 1 = use tcgen05 kernels where an op has wtc_off >= 0 (default), 0 = SIMT only. */
int lp_set_tensor_core(lp_ctx* ctx, int enable);
/* 1 = chain consecutive tensor-core conv kernels with programmatic dependent launch (default; env LP_NO_PDL=1
 * starts with 0). */
int lp_set_pdl(lp_ctx* ctx, int enable);

/* Fused classifier: the whole ShuffleNetV2 forward inside one persistent CTA per SM, driven by a host-built
 * program (plan.py build_fused_classifier -> FusedProgram; struct FStep in csrc/shufflenet_fused.cu, 18 x int32
 * per step, device memory) over an fp32 weight blob.  Three step lists: front (n_front steps, per ROI), middle
 * (n_mid, per ROI; ends by parking park_floats floats per ROI in global memory), tail (n_tail, the CTA's ROIs
 * stacked tail_group at a time so that 73 % of the weight stream is read once per group).  When loaded,
 * lp_classify uses it instead of the layer-by-layer plan; lp_set_fused_classifier(ctx, 0) switches back.
 * smem_bytes = extent of the front/middle activation map, back_bytes = extent the middle still uses (behind it:
 * astage_bytes of fp16 activation staging and the middle's weight stages), tail_bytes = extent of the tail's map
 * (behind it tail_astage_bytes of fp16 staging, when its pointwise layers also have fp16 weights).
 * weights16 (device, may be NULL) = split-f16 weights [cout_p8][hi|lo][L] of the middle's pointwise layers
 * (FStep.w16_off), which then run on the tensor cores (mma.sync, Ahi*Bhi + Alo*Bhi + Ahi*Blo in fp32).
 * park (device, caller-owned like every other buffer) = lp_sm_count(ctx) x tail_group x park_floats floats. */
int lp_fused_classifier_load(lp_ctx* ctx, const void* steps_dev, int n_front, int n_mid, int n_tail,
                             const float* weights, const void* weights16, int tail_group, int in_hw, int n_classes,
                             size_t smem_bytes, size_t back_bytes, size_t astage_bytes, size_t tail_bytes,
                             size_t tail_astage_bytes, int park_floats, float* park, size_t park_bytes,
                             float mean, float stdv);
int lp_set_fused_classifier(lp_ctx* ctx, int enable);
int lp_sm_count(lp_ctx* ctx);

/* ---- K1: letterbox.  Replaces letterbox() + cvtColor (e2e.py:66-86, :224-225).
 * frames[i]: HWC BGR u8 image i (pitch[i] bytes per row).  out: B x S x S x 3 RGB u8,
 * bit-exact with cv2.resize(INTER_LINEAR) + copyMakeBorder(114).  ratio[i],
 * pad[2i..2i+1] = (dw, dh) as the reference returns them (f64 -> stored as f64). */
int lp_letterbox(lp_ctx* ctx, const uint8_t* const* frames_h, const int32_t* h_h, const int32_t* w_h,
                 const int64_t* pitch_h, int batch, int out_size, uint8_t* out,
                 double* ratio_h, double* pad_h, void* stream);

/* ---- K2/K3: detector forward.  Replaces ex.input/ex.extract (e2e.py:305-307).
 * in: B x S x S x 3 RGB u8 (lp_letterbox output).  out0: B x 5 x 8400 f32 in the
 * reference layout (rows cx,cy,w,h,score; e2e.py:244-253). */
int lp_detect_forward(lp_ctx* ctx, const uint8_t* in, int batch, void* workspace, size_t workspace_bytes,
                      float* out0, void* stream);

/* ---- K4+K5: decode, threshold, un-letterbox, clip, per-class NMS.
 * Replaces NCNNDetector.postprocess + nms_numpy (e2e.py:240-296, :89-119).
 * out0: B x (4+nc) x A.  Per image i (orig size h_h[i] x w_h[i], f32(ratio), f32(pad) 2 per image):
 * boxes[i][k][4] xyxy original px, scores[i][k], classes[i][k], keep_idx[i][k]
 * (index into the thresholded candidate list, reference order), counts[i] = K_i,
 * n_cand[i] = candidates above conf.  Capacity max_det per image; counts[i] is the
 * true K_i even when > max_det (caller checks).  Bit-exact with the reference. */
int lp_decode_nms(lp_ctx* ctx, const float* out0, int nc, int n_anchors,
                  const int32_t* h_h, const int32_t* w_h, const float* ratio_h, const float* pad_h, int batch,
                  float conf, float iou, int max_det,
                  float* boxes, float* scores, int64_t* classes, int32_t* keep_idx,
                  int32_t* counts, int32_t* n_cand, void* scratch, size_t scratch_bytes, void* stream);
size_t lp_decode_nms_scratch_bytes(int batch, int n_anchors);

/* ---- K6: ROI clip/filter + PIL-exact antialiased bilinear resize.
 * Replaces the ROI loop of HybridPipeline.run (e2e.py:459-475) and
 * cvtColor + transforms.Resize((64,64)) (e2e.py:385-388).
 * lp_roi_select: per detection int-truncate, clip, area filter; compacts in
 * (image, detection) order.  roi_xyxy[r][4] int32, roi_src[r] = {image, det idx};
 * n_rois = total (device int32[1]); capacity max_rois.
 * lp_roi_resize: rois -> out_size x out_size x 3 RGB u8, bit-exact with Pillow. */
int lp_roi_select(lp_ctx* ctx, const float* boxes, const int32_t* counts, int max_det,
                  const int32_t* h_h, const int32_t* w_h, int batch, int min_area, int max_rois,
                  int32_t* roi_xyxy, int32_t* roi_src, int32_t* n_rois, void* stream);
int lp_roi_resize(lp_ctx* ctx, const uint8_t* const* frames_h, const int64_t* pitch_h, int batch,
                  const int32_t* roi_xyxy, const int32_t* roi_src, int n_rois, int out_size,
                  int max_side /* largest frame side: bounds the filter support */,
                  uint8_t* out, void* stream);

/* ---- K7: ShuffleNetV2 forward + softmax + argmax.  Replaces ToTensor/Normalize,
 * self.model(batch), torch.softmax and np.argmax (e2e.py:366-370, :391-396).
 * in: R x S x S x 3 RGB u8.  logits/probs: R x C f32, argmax: R int64. */
int lp_classify(lp_ctx* ctx, const uint8_t* in, int n, void* workspace, size_t workspace_bytes,
                float* logits, float* probs, int64_t* argmax, void* stream);

/* Pack per-detection records for the multi-GPU gather:
 * rec[r] = {frame_id, x1,y1,x2,y2 (f32 bits), det_conf, det_cls, cls_cls, cls_conf} 9 x 4 B. */
int lp_pack_records(lp_ctx* ctx, const int32_t* roi_src, const int32_t* frame_ids,
                    const float* boxes, const float* scores, const int64_t* classes, int max_det,
                    const int64_t* cls_argmax, const float* probs, int n_classes, int n_rois,
                    int32_t* records, void* stream);

/* ROI semantics.  mode 0 (default) = src/vntsr/pipeline/e2e.py: clip x1,y1 to
 * [0, w-1]/[0, h-1] and x2,y2 to [x1+1, w]/[y1+1, h] (:462-469), Pillow antialiased BILINEAR resize
 * (:385-388).  mode 1 = src/tt100k/pipeline/e2e_optimize.py: clip all four to [0, w]/[0, h] (:480-483),
 * keep iff area >= min_area and the box is non-empty (:486-496), cv2.resize INTER_LINEAR (:391-393).
 * Affects lp_roi_select and lp_roi_resize of this context. */
int lp_set_roi_mode(lp_ctx* ctx, int mode);

/* Device-side ROI count.  The reference sizes its classifier batches on the host (len(rois),
 * e2e.py:477-485); a GPU pipeline would have to read the count back in the middle of a step and
 * leave the device idle meanwhile.  After this call with a non-null pointer (the int32 that
 * lp_roi_select writes), lp_roi_resize / lp_classify (fused classifier only) / lp_pack_records
 * treat their count argument as a CAPACITY and process min(*n_rois_dev, capacity) ROIs, so a whole
 * step is enqueued without a host synchronisation.  Null restores host counts. */
int lp_set_roi_count_device(lp_ctx* ctx, const int32_t* n_rois_dev);

/* ---- Evaluation matching.  Replaces section 1 of evaluate_predictions
 * (e2e.py:687-731, box_iou :663-676): for every frame and every IoU threshold, each prediction keeps
 * its best ground truth with iou >= t, each ground truth keeps the lowest-index prediction that chose
 * it, and correct[p][t] = 1 iff their classes agree.  float64, numpy's operation order.
 * pred_box [P][4] / gt_box [G][4] xyxy; *_off [n_frames + 1] prefix offsets; max_per_frame >=
 * max over frames of (predictions + ground truths); correct [P][n_thr] u8.  All device pointers. */
int lp_eval_match(lp_ctx* ctx, const double* pred_box, const int32_t* pred_cls, const int32_t* pred_off,
                  const double* gt_box, const int32_t* gt_cls, const int32_t* gt_off, int n_frames,
                  int max_per_frame, const double* thresholds, int n_thr, uint8_t* correct, void* stream);

/* ---- Frame ingest: baseline JPEG decode on the device.  Replaces cv2.imread (e2e.py:962) for frames
 * that arrive as JPEG bytes; bit-exact with cv2.imdecode (libjpeg-turbo defaults: islow IDCT, fancy up-sampling).
 * All images of a call share one header (size, sampling, tables): `desc` (host) + `tables` (device blob of
 * lp_jpeg_tables_bytes() bytes, layout = struct JpegTables in csrc/jpeg.cu, built by jpeg.py pack_tables).
 * data (device) = the entropy-coded scans back to back (everything after SOS, stuffing and RSTn markers intact);
 * img_off (device, batch + 1 int64) = their byte offsets.  frames_out = batch x height x width x 3 BGR u8.
 * One thread decodes one restart interval, so throughput needs an encoder that emits RSTn every few MCUs. */
typedef struct lp_jpeg_desc {
    int32_t width, height, ncomp;       /* ncomp 1 (grey) or 3 (YCbCr) */
    int32_t h[3], v[3];                 /* sampling factors: luma 1x1 / 2x1 / 2x2, chroma 1x1 */
    int32_t tq[3], td[3], ta[3];        /* quantisation / DC / AC table selectors per component */
    int32_t restart_interval;           /* MCUs per restart interval (DRI), 0 = none */
} lp_jpeg_desc;
size_t lp_jpeg_tables_bytes(void);
size_t lp_jpeg_scratch_bytes(const lp_jpeg_desc* desc, int batch);
int lp_jpeg_decode(lp_ctx* ctx, const uint8_t* data, const int64_t* img_off, int batch, int max_batch /* the scratch was sized for */,
                   const lp_jpeg_desc* desc, const void* tables, void* scratch, size_t scratch_bytes, uint8_t* frames_out,
                   void* stream);

/* Workspace bytes lp_detect_forward / lp_classify need for the loaded plan (0 if not loaded). */
size_t lp_workspace_bytes(lp_ctx* ctx, int net);

/* Probe: CUDA events around op `op_index` of plan `net` on the launching stream (op_index < 0
 * disables).  lp_probe_read returns how many samples it wrote (ms each, oldest first, <= 512). */
int lp_probe_set(lp_ctx* ctx, int net, int op_index);
int lp_probe_read(lp_ctx* ctx, float* ms_h, int cap);

/* Which kernel family ran each op of plan `net` in the last forward (host array, one int8 per op): 0 generic SIMT
 * kernel, 1 parameter-weight small-channel conv, 2 tcgen05 implicit-GEMM conv, 3 absorbed by the previous op's
 * kernel, 4 warp-level MMA small-channel conv.  Returns the number of entries written.  With lp_probe_set(net, -2) every op gets one event pair
 * (slot = op index; for the detector slot n_ops is the Detect tail). */
int lp_op_paths(lp_ctx* ctx, int net, int8_t* out_h, int cap);

/* Debugging: per-role cycle counters of the tensor-core conv kernel (CTA 0) into a device buffer of
 * 16 int64 (NULL disables): [0..2] loader wait-empty/issue/wait-copy, [3..7] MMA wait-acc/wait-patch/
 * wait-weights/total/tiles, [8..9] epilogue wait/total. */
int lp_debug_tc_timing(lp_ctx* ctx, void* dev_buf16);

/* Counters: number of kernels this library launched since lp_create (bench gpu_launches). */
int64_t lp_launch_count(lp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LITEPI_B200_H */
