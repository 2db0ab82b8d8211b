// Evaluation matching: predictions vs ground truth of every frame at T IoU thresholds.
// Replaces section 1 of evaluate_predictions (src/vntsr/pipeline/e2e.py:687-731): box_iou (:663-676) in
// float64 with numpy's operation order, then per threshold
//   1. every prediction keeps its best pair among iou >= t        (argsort desc + unique over predictions)
//   2. every ground truth keeps the LOWEST-INDEX such prediction   (unique over ground truths after np.unique
//                                                                   re-sorted the pairs by prediction index)
//   3. the prediction is correct at t iff the classes agree.
// Exact IoU ties between two ground truths of one prediction go to the higher ground-truth index (what
// numpy's argsort()[::-1] yields for up to 16 candidate pairs; larger tie sets are unspecified in the reference).
// One block per frame; the per-frame counts are tens, so everything lives in shared memory.
#include "common.cuh"

namespace {

__device__ __forceinline__ double iou_f64(const double* a, const double* b) {
    // explicit _rn operations: no FMA contraction, same rounding sequence as the numpy expression
    const double area_a = __dmul_rn(__dsub_rn(a[2], a[0]), __dsub_rn(a[3], a[1]));
    const double area_b = __dmul_rn(__dsub_rn(b[2], b[0]), __dsub_rn(b[3], b[1]));
    double w = __dsub_rn(fmin(a[2], b[2]), fmax(a[0], b[0]));
    double h = __dsub_rn(fmin(a[3], b[3]), fmax(a[1], b[1]));
    w = w < 0.0 ? 0.0 : w;
    h = h < 0.0 ? 0.0 : h;
    const double inter = __dmul_rn(w, h);
    const double uni = __dsub_rn(__dadd_rn(area_a, area_b), inter);
    return __ddiv_rn(inter, __dadd_rn(uni, 1e-7));
}

__global__ void __launch_bounds__(128) eval_match_kernel(const double* __restrict__ pred_box, const int* __restrict__ pred_cls,
                                                         const int* __restrict__ pred_off, const double* __restrict__ gt_box,
                                                         const int* __restrict__ gt_cls, const int* __restrict__ gt_off,
                                                         const double* __restrict__ thr, int n_thr,
                                                         unsigned char* __restrict__ correct) {
    extern __shared__ int s_i[];
    const int f = blockIdx.x;
    const int p0 = pred_off[f], P = pred_off[f + 1] - p0;
    const int g0 = gt_off[f], G = gt_off[f + 1] - g0;
    int* best_gt = s_i;              // [P]
    int* owner = s_i + P;            // [G]
    for (int i = threadIdx.x; i < P * n_thr; i += blockDim.x) correct[(long long)p0 * n_thr + i] = 0;
    if (P == 0 || G == 0) return;
    __syncthreads();
    for (int t = 0; t < n_thr; ++t) {
        const double th = thr[t];
        for (int g = threadIdx.x; g < G; g += blockDim.x) owner[g] = 0x7fffffff;
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const double* pb = pred_box + (long long)(p0 + p) * 4;
            double best = -1.0;
            int bg = -1;
            for (int g = 0; g < G; ++g) {
                const double v = iou_f64(pb, gt_box + (long long)(g0 + g) * 4);
                if (v >= th && v >= best) { best = v; bg = g; }
            }
            best_gt[p] = bg;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < P; p += blockDim.x)
            if (best_gt[p] >= 0) atomicMin(&owner[best_gt[p]], p);
        __syncthreads();
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const int p = owner[g];
            if (p != 0x7fffffff && pred_cls[p0 + p] == gt_cls[g0 + g]) correct[(long long)(p0 + p) * n_thr + t] = 1;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int lp_eval_match(lp_ctx* ctx, const double* pred_box, const int32_t* pred_cls, const int32_t* pred_off,
                             const double* gt_box, const int32_t* gt_cls, const int32_t* gt_off, int n_frames,
                             int max_per_frame, const double* thresholds, int n_thr, uint8_t* correct, void* stream) {
    LP_CHECK(ctx && pred_off && gt_off && thresholds && n_thr > 0, "lp_eval_match: null argument");
    lp_device_guard dev_guard(ctx);
    if (n_frames <= 0) return 0;
    LP_CHECK(pred_box && pred_cls && gt_box && gt_cls && correct, "lp_eval_match: null array");
    const size_t smem = (size_t)max_per_frame * sizeof(int);
    LP_CHECK(smem <= 48 * 1024, "lp_eval_match: %d predictions + ground truths in one frame exceed shared memory", max_per_frame);
    eval_match_kernel<<<n_frames, 128, smem, (cudaStream_t)stream>>>(pred_box, pred_cls, pred_off, gt_box, gt_cls, gt_off,
                                                                     thresholds, n_thr, correct);
    LP_LAUNCH_OK(ctx);
    return 0;
}
