"""Evaluation of pipeline output against ground truth: the B200 counterpart of
``evaluate_predictions`` (/root/reference/src/vntsr/pipeline/e2e.py:656-824), same arguments and the
same result dictionary.

The matching of predictions to ground truth at the ten IoU thresholds (the per-image python loops of
the reference, e2e.py:687-731) runs on the GPU for all frames at once (csrc/eval_match.cu through
``lp_eval_match``); the per-class curves (cumulative TP/FP, 101-point AP, best-F1 operating point,
e2e.py:744-824) are a few vector operations on the host.  There is no CPU fallback for the matching.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

IOU_THRESHOLDS = np.arange(0.5, 1.0, 0.05)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)


class Evaluator:
    """Owns a library context on ``device``; reusable across calls."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("litepi_b200.Evaluator needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device)
        self.ctx = L.Context(device)

    # ------------------------------------------------------------------ matching (GPU)
    def match(self, pred_box: np.ndarray, pred_cls: np.ndarray, pred_n: np.ndarray, gt_box: np.ndarray,
              gt_cls: np.ndarray, gt_n: np.ndarray, thresholds: np.ndarray = IOU_THRESHOLDS) -> np.ndarray:
        """Flat inputs: ``pred_box`` [P,4], ``pred_cls`` [P], ``pred_n`` [F] predictions per frame (same for gt).
        Returns ``correct`` [P, T] bool."""
        F = int(len(pred_n))
        assert len(gt_n) == F
        P, T = int(np.sum(pred_n)), int(len(thresholds))
        if P == 0:
            return np.zeros((0, T), dtype=bool)
        dev = self.device
        off = lambda n: torch.from_numpy(np.concatenate(([0], np.cumsum(n))).astype(np.int32)).to(dev)
        f64 = lambda a, w: torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, w))).to(dev)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(np.int32))).to(dev)
        with torch.cuda.device(dev):
            pb, pc, po = f64(pred_box, 4), i32(pred_cls), off(pred_n)
            gb, gc, go = f64(gt_box, 4), i32(gt_cls), off(gt_n)
            th = torch.from_numpy(np.asarray(thresholds, np.float64).copy()).to(dev)
            correct = torch.empty((P, T), dtype=torch.uint8, device=dev)
            most = int(np.max(np.asarray(pred_n) + np.asarray(gt_n))) if F else 0
            L.check(L.lib().lp_eval_match(self.ctx.handle, _ptr(pb), _ptr(pc), _ptr(po), _ptr(gb), _ptr(gc), _ptr(go), F,
                                          most, _ptr(th), T, _ptr(correct),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lp_eval_match")
            return correct.cpu().numpy().astype(bool)

    # ------------------------------------------------------------------ reference-shaped API
    def evaluate_predictions(self, all_preds: Sequence[List[dict]], all_gts: Sequence[Sequence[Sequence[float]]],
                             num_classes: int, iou_threshold: float = 0.5,
                             iou_thresholds: np.ndarray = IOU_THRESHOLDS) -> Dict[str, object]:
        """e2e.py:656.  ``all_preds[i]`` = list of {'bbox','conf','cls_class'}; ``all_gts[i]`` = rows
        [cls, x1, y1, x2, y2].  ``iou_threshold`` is accepted and unused, as in the reference."""
        keep = [i for i, (p, g) in enumerate(zip(all_preds, all_gts)) if len(p) or len(g)]     # :690-697
        if not keep:
            z = np.zeros(num_classes)
            return {"precision": z, "recall": z.copy(), "f1": z.copy(), "tp": z.copy(), "fp": z.copy(), "fn": z.copy(),
                    "mAP50": 0.0, "mAP50_95": 0.0, "classes_present": np.zeros(num_classes, dtype=bool)}
        preds = [all_preds[i] for i in keep]
        gts = [np.asarray(all_gts[i], np.float64).reshape(-1, 5) for i in keep]
        pred_n = np.array([len(p) for p in preds], np.int64)
        gt_n = np.array([len(g) for g in gts], np.int64)
        pb = np.array([q["bbox"] for p in preds for q in p], np.float64).reshape(-1, 4)
        conf = np.array([q["conf"] for p in preds for q in p], np.float64)
        pcls = np.array([q["cls_class"] for p in preds for q in p], np.int64)
        g_all = np.concatenate(gts, 0) if gts else np.zeros((0, 5))
        correct = self.match(pb, pcls, pred_n, g_all[:, 1:], g_all[:, 0], gt_n, iou_thresholds)
        return metrics_from_stats(correct, conf, pcls, g_all[:, 0], num_classes)

    def evaluate_records(self, records: np.ndarray, n_frames: int, all_gts, num_classes: int) -> Dict[str, object]:
        """Packed pipeline records (``B200Pipeline.fetch_records`` / ``runner.gather_records``) -> metrics, with the
        prediction dictionaries ``process_image`` builds (e2e.py:996-1001: int-truncated bbox, det_conf, cls_class)."""
        f = records.view(np.float32)
        preds: List[List[dict]] = [[] for _ in range(n_frames)]
        for i in range(records.shape[0]):
            preds[int(records[i, 0])].append({"bbox": tuple(f[i, 1:5].astype(int)), "conf": float(f[i, 5]),
                                              "cls_class": int(records[i, 7])})
        return self.evaluate_predictions(preds, all_gts, num_classes)


def _average_precision(recall: np.ndarray, precision: np.ndarray) -> float:
    """e2e.py:679-685: precision envelope, 101 recall points, trapezoid rule."""
    r = np.concatenate(([0.0], recall, [1.0]))
    p = np.concatenate(([1.0], precision, [0.0]))
    p = np.maximum.accumulate(p[::-1])[::-1]
    x = np.linspace(0, 1, 101)
    y = np.interp(x, r, p)
    return float((np.diff(x) * (y[1:] + y[:-1]) / 2.0).sum())


def metrics_from_stats(correct: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray,
                       num_classes: int) -> Dict[str, object]:
    """Per-class curves from the matching result (e2e.py:744-824); float64 host arithmetic."""
    order = np.argsort(-conf)
    tp_sorted, cls_sorted = correct[order], pred_cls[order]
    present, counts = np.unique(target_cls, return_counts=True)
    n_gt_of = {c: n for c, n in zip(present, counts)}
    out = {k: np.zeros(num_classes) for k in ("precision", "recall", "f1", "tp", "fp", "fn")}
    ap50, ap_all = np.zeros(num_classes), np.zeros(num_classes)
    for c in range(num_classes):
        n_gt = n_gt_of.get(c, 0)
        mine = cls_sorted == c
        n_pred = int(mine.sum())
        if n_pred == 0 or n_gt == 0:
            if n_pred == 0 and n_gt == 0:
                continue
            out["fn"][c] = n_gt
            continue
        hits = tp_sorted[mine]
        tpc, fpc = hits.cumsum(0), (1 - hits).cumsum(0)
        recall = tpc / (n_gt + 1e-16)
        precision = tpc / (tpc + fpc + 1e-16)
        aps = [_average_precision(recall[:, j], precision[:, j]) for j in range(hits.shape[1])]
        ap50[c], ap_all[c] = aps[0], np.mean(aps)
        p0, r0 = precision[:, 0], recall[:, 0]
        f1 = 2 * p0 * r0 / (p0 + r0 + 1e-16)
        k = int(np.argmax(f1))
        out["precision"][c], out["recall"][c], out["f1"][c] = p0[k], r0[k], f1[k]
        out["tp"][c], out["fp"][c] = tpc[k, 0], fpc[k, 0]
        out["fn"][c] = n_gt - out["tp"][c]
    idx = present.astype(int)
    out["mAP50"] = float(np.mean(ap50[idx])) if len(idx) else 0.0
    out["mAP50_95"] = float(np.mean(ap_all[idx])) if len(idx) else 0.0
    out["ap50_per_class"] = ap50
    out["classes_present"] = np.isin(np.arange(num_classes), present)
    return out
