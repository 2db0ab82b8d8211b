"""B200Pipeline -- drop-in for the reference's ``HybridPipeline``
(``src/vntsr/pipeline/e2e.py:399-531``) plus the batched / device-resident entry
points the B200 needs (``run_batch``, ``run_device``).

Stage order per batch, all on one CUDA stream:
  H2D -> K1 letterbox -> K2/K3 detector -> K4 decode+threshold -> K5 NMS -> ROI select
      -> (4-byte D2H of the ROI count) -> K6 crop+resize -> K7 ShuffleNetV2 -> pack -> D2H records
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .classifier import B200Classifier
from .detector import B200Detector, FrameBatch, _ptr, _stream


@dataclass
class PipelineMetrics:
    """Field-for-field the reference's dataclass (e2e.py:34-62)."""
    t_detection: float = 0.0
    t_roi_extract: float = 0.0
    t_classification: float = 0.0
    t_postprocess: float = 0.0
    t_total: float = 0.0
    fps: float = 0.0
    num_detections: int = 0
    det_confidence_avg: float = 0.0
    cls_confidence_avg: float = 0.0
    cpu_percent: float = 0.0
    memory_mb: float = 0.0
    temperature: float = 0.0
    precision: float = 0.0
    recall: float = 0.0
    f1: float = 0.0
    level: str = "B200"


REC_WORDS = 9   # frame_id, x1,y1,x2,y2, det_conf, det_cls, cls_cls, cls_conf


class B200Pipeline:
    def __init__(self, detector_param: str, detector_bin: Optional[str], classifier_path: Optional[str],
                 classifier_arch: str = "shufflenetv2", num_classes: int = 58, det_input_size: int = 640,
                 cls_input_size: int = 64, use_gpu_detector: bool = False, detector_threads: int = 4,
                 classifier_device: str = "cpu", batch_size: int = 8,
                 device: int = 0, max_batch: int = 64, max_det: int = 1024, max_rois: Optional[int] = None,
                 classifier_state_dict: Optional[dict] = None, seed: Optional[int] = None,
                 roi_mode: str = "reference"):
        self._ctor = dict(detector_param=detector_param, detector_bin=detector_bin, classifier_path=classifier_path,
                          classifier_arch=classifier_arch, num_classes=num_classes, det_input_size=det_input_size,
                          cls_input_size=cls_input_size, use_gpu_detector=use_gpu_detector,
                          detector_threads=detector_threads, classifier_device=classifier_device, batch_size=batch_size,
                          device=device, max_batch=max_batch, max_det=max_det, max_rois=max_rois, seed=seed,
                          roi_mode=roi_mode)
        self.detector = B200Detector(detector_param, detector_bin, input_size=det_input_size,
                                     use_gpu=use_gpu_detector, num_threads=detector_threads,
                                     device=device, max_batch=max_batch, max_det=max_det, seed=seed or 0)
        # classify all ROIs of a frame batch in one pass: the classifier is launch/latency-bound, so its
        # time barely depends on the ROI count (workspace = 0.77 MB per ROI slot)
        self.classifier = B200Classifier(classifier_path, classifier_arch, num_classes=num_classes,
                                         input_size=cls_input_size, device=classifier_device,
                                         state_dict=classifier_state_dict, cuda_device=device, seed=seed,
                                         max_batch=min(max(256, 16 * int(max_batch)), 1024))
        # ROI semantics: "reference" = src/vntsr/pipeline/e2e.py (clip rules :462-469, Pillow resize); "optimized" =
        # src/tt100k/pipeline/e2e_optimize.py (plain clip :480-483, cv2 INTER_LINEAR resize :391-393)
        if roi_mode not in ("reference", "optimized"):
            raise ValueError(f"Unknown roi_mode: {roi_mode}")
        self.roi_mode = roi_mode
        for ctx in (self.detector.ctx, self.classifier.ctx):
            L.check(L.lib().lp_set_roi_mode(ctx.handle, 1 if roi_mode == "optimized" else 0), "lp_set_roi_mode")
        self.batch_size = batch_size          # reference's classifier mini-batch; ROIs are classified in one pass here
        self.device = self.detector.device
        self.ctx = self.detector.ctx
        self.max_batch = int(max_batch)

        class _Launches:                      # kernels launched by both contexts (bench.py gpu_launches)
            def __init__(s, a, b): s.a, s.b = a, b
            def launch_count(s): return s.a.launch_count() + s.b.launch_count()
        self.counters = _Launches(self.detector.ctx, self.classifier.ctx)
        self._stream_runner = None
        self.max_rois = int(max_rois) if max_rois else self.max_batch * 64
        with torch.cuda.device(self.device):
            R = self.max_rois
            self.roi_xyxy = torch.empty((R, 4), dtype=torch.int32, device=self.device)
            self.roi_src = torch.empty((R, 2), dtype=torch.int32, device=self.device)
            self.n_rois = torch.zeros((1,), dtype=torch.int32, device=self.device)
            self.records = torch.empty((R, REC_WORDS), dtype=torch.int32, device=self.device)
            # host mirrors are double-buffered ("slots") so step s+1 can be enqueued while step s is read back
            self._n_rois_h = [torch.zeros((1,), dtype=torch.int32).pin_memory() for _ in range(2)]
            self._records_h = [torch.empty((R, REC_WORDS), dtype=torch.int32).pin_memory() for _ in range(2)]
            self._counts_h = [torch.zeros((self.max_batch,), dtype=torch.int32).pin_memory() for _ in range(2)]
            self._fetched = [torch.cuda.Event() for _ in range(2)]
            self._pending = [0, 0]

    # ------------------------------------------------------------------ device path
    def enqueue_device(self, fb: FrameBatch, conf_threshold: float = 0.5, iou_threshold: float = 0.45,
                       min_area: int = 100, frame_ids: Optional[torch.Tensor] = None, slot: int = 0) -> None:
        """Enqueue the whole hot path for device-resident frames WITHOUT a host synchronisation: the ROI
        count stays on the device (lp_set_roi_count_device) and the ROI-side kernels are launched at
        capacity ``max_rois``.  ``finish`` waits and validates; records are in ``self.records`` (device)."""
        det, clf, lib = self.detector, self.classifier, L.lib()
        det.detect_device(fb, conf_threshold, iou_threshold)
        L.check(lib.lp_roi_select(self.ctx.handle, _ptr(det.boxes), _ptr(det.counts), det.max_det, fb.h, fb.w, fb.n,
                                  int(min_area), self.max_rois, _ptr(self.roi_xyxy), _ptr(self.roi_src),
                                  _ptr(self.n_rois), _stream()), "lp_roi_select")
        self._n_rois_h[slot].copy_(self.n_rois, non_blocking=True)
        self._counts_h[slot][:fb.n].copy_(det.counts[:fb.n], non_blocking=True)
        self._pending[slot] = fb.n
        if not clf.fused:
            return                                    # layer-by-layer classifier: sized on the host in finish()
        cap = self.max_rois
        for ctx in (clf.ctx, self.ctx):
            L.check(lib.lp_set_roi_count_device(ctx.handle, _ptr(self.n_rois)), "lp_set_roi_count_device")
        try:
            cls_in = clf.resize_device(fb, self.roi_xyxy, self.roi_src, cap)
            clf.classify_device(cls_in)
            L.check(lib.lp_pack_records(self.ctx.handle, _ptr(self.roi_src), _ptr(frame_ids), _ptr(det.boxes),
                                        _ptr(det.scores), _ptr(det.classes), det.max_det, _ptr(clf.argmax),
                                        _ptr(clf.probs), clf.num_classes, cap, _ptr(self.records), _stream()),
                    "lp_pack_records")
        finally:
            for ctx in (clf.ctx, self.ctx):
                lib.lp_set_roi_count_device(ctx.handle, None)

    def _validate(self, slot: int) -> int:
        n, nf = int(self._n_rois_h[slot][0]), int(self._pending[slot])
        if n > self.max_rois:
            raise RuntimeError(f"litepi_b200: {n} ROIs exceed max_rois={self.max_rois}")
        if nf and int(self._counts_h[slot][:nf].max()) > self.detector.max_det:
            raise RuntimeError(f"litepi_b200: a frame has more than max_det={self.detector.max_det} detections")
        return n

    def enqueue_fetch_copy(self, slot: int = 0) -> None:
        """The D2H copy of the step's records (capacity-sized: the count is not known on the host yet), queued
        behind the step on the current stream.  Capturable in a CUDA graph; ``mark_fetch`` records its event."""
        self._records_h[slot].copy_(self.records, non_blocking=True)

    def mark_fetch(self, slot: int = 0) -> None:
        self._fetched[slot].record(torch.cuda.current_stream())

    def mark_enqueued(self, slot: int, n_frames: int) -> None:
        """Host-side bookkeeping of ``enqueue_device`` when the step was replayed from a CUDA graph."""
        self._pending[slot] = int(n_frames)

    def enqueue_fetch(self, slot: int = 0) -> None:
        self.enqueue_fetch_copy(slot)
        self.mark_fetch(slot)

    def clone(self) -> "B200Pipeline":
        """A second instance with the same models and its own workspace (one per lane of ``stream()``)."""
        return B200Pipeline(**self._ctor, classifier_state_dict=self.classifier.state_dict)

    def stream(self, lanes: int = 2, use_graph: Optional[bool] = None):
        """Host-buffer streaming front end (pinned ring, copy stream, lanes, CUDA graph): see stream.py."""
        from .stream import StreamRunner
        return StreamRunner(self, lanes=lanes, use_graph=use_graph)

    def run_stream(self, batches, conf_threshold: float = 0.5, iou_threshold: float = 0.45, min_area: int = 100,
                   lanes: int = 2):
        """Iterate batches of HOST frames (lists / arrays of HWC BGR uint8, e.g. what cv2.imread returns) through the
        whole hot path; yields one result list per batch (per frame: the dicts of e2e.py:519-529), in order."""
        if getattr(self, "_stream_runner", None) is None or self._stream_runner.n_lanes != lanes:
            self._stream_runner = self.stream(lanes)
        sizes: List[int] = []

        def counted():
            for b in batches:
                sizes.append(len(b))
                yield b
        for rec in self._stream_runner.run_stream(counted(), conf_threshold, iou_threshold, min_area):
            yield self.records_to_results(rec, sizes.pop(0))

    def collect(self, slot: int = 0) -> np.ndarray:
        """Records of the step enqueued with ``slot`` (after ``enqueue_fetch``), as a [n, 9] int32 array."""
        self._fetched[slot].synchronize()
        n = self._validate(slot)
        return self._records_h[slot][:n].numpy().copy()

    def finish(self, fb: Optional[FrameBatch] = None, frame_ids: Optional[torch.Tensor] = None, slot: int = 0) -> int:
        """Wait for the enqueued step; returns its ROI count.  Capacity overruns fail loudly here."""
        det, clf, lib = self.detector, self.classifier, L.lib()
        torch.cuda.current_stream().synchronize()
        n = self._validate(slot)
        if n and not clf.fused:
            assert fb is not None, "finish(fb) is required with the layer-by-layer classifier"
            cls_in = clf.resize_device(fb, self.roi_xyxy, self.roi_src, n)
            clf.classify_device(cls_in)
            L.check(lib.lp_pack_records(self.ctx.handle, _ptr(self.roi_src), _ptr(frame_ids), _ptr(det.boxes),
                                        _ptr(det.scores), _ptr(det.classes), det.max_det, _ptr(clf.argmax),
                                        _ptr(clf.probs), clf.num_classes, n, _ptr(self.records), _stream()),
                    "lp_pack_records")
        return n

    def run_device(self, fb: FrameBatch, conf_threshold: float = 0.5, iou_threshold: float = 0.45,
                   min_area: int = 100, frame_ids: Optional[torch.Tensor] = None) -> int:
        """Whole hot path on device-resident frames.  Returns the ROI count; records for them are in
        ``self.records[:n]`` (device)."""
        self.enqueue_device(fb, conf_threshold, iou_threshold, min_area, frame_ids)
        return self.finish(fb, frame_ids)

    def fetch_records(self, n: int) -> np.ndarray:
        """D2H of ``n`` packed records -> structured numpy view [n] (blocking)."""
        if n == 0:
            return np.zeros((0, REC_WORDS), np.int32)
        self._records_h[0][:n].copy_(self.records[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._records_h[0][:n].numpy().copy()

    @staticmethod
    def records_to_results(rec: np.ndarray, n_frames: int) -> List[List[Dict]]:
        """Packed records -> per-frame result dict lists (keys of e2e.py:519-529)."""
        out: List[List[Dict]] = [[] for _ in range(n_frames)]
        if rec.shape[0] == 0:
            return out
        f = rec.view(np.float32)
        for i in range(rec.shape[0]):
            box = f[i, 1:5]
            out[int(rec[i, 0])].append({
                "bbox": tuple(box.astype(int)),
                "box_f32": box.copy(),
                "det_class": int(rec[i, 6]),
                "det_conf": float(f[i, 5]),
                "cls_class": int(rec[i, 7]),
                "cls_conf": float(f[i, 8]),
            })
        return out

    # ------------------------------------------------------------------ batched host API
    def run_batch(self, images: Sequence[np.ndarray], conf_threshold: float = 0.5, iou_threshold: float = 0.45,
                  min_area: int = 100) -> List[List[Dict]]:
        """Batched form of ``run``: list of frames -> list (per frame) of result dicts."""
        results: List[List[Dict]] = []
        for i in range(0, len(images), self.max_batch):
            chunk = images[i:i + self.max_batch]
            fb = FrameBatch.from_host(chunk, self.device)
            n = self.run_device(fb, conf_threshold, iou_threshold, min_area)
            results.extend(self.records_to_results(self.fetch_records(n), len(chunk)))
        return results

    # ------------------------------------------------------------------ reference API
    def run(self, image: np.ndarray, conf_threshold: float = 0.5, iou_threshold: float = 0.45,
            min_area: int = 100) -> Tuple[List[Dict], PipelineMetrics]:
        """e2e.py:443-531.  Stage times come from CUDA events on the launching stream."""
        m = PipelineMetrics(level="B200")
        t_start = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        det, clf, lib = self.detector, self.classifier, L.lib()
        fb = FrameBatch.from_host([image], self.device)
        ev[0].record()
        det.detect_device(fb, conf_threshold, iou_threshold)
        ev[1].record()
        L.check(lib.lp_roi_select(self.ctx.handle, _ptr(det.boxes), _ptr(det.counts), det.max_det, fb.h, fb.w, 1,
                                  int(min_area), self.max_rois, _ptr(self.roi_xyxy), _ptr(self.roi_src),
                                  _ptr(self.n_rois), _stream()), "lp_roi_select")
        ev[2].record()
        dets = det._collect(1)[0]                      # syncs; reference counts pre-filter detections (e2e.py:454)
        m.num_detections = len(dets[0])
        if len(dets[1]) > 0:
            m.det_confidence_avg = float(np.mean(dets[1]))
        n = int(self.n_rois.cpu()[0])
        if n > self.max_rois:
            raise RuntimeError(f"litepi_b200: {n} ROIs exceed max_rois={self.max_rois}")
        if n:
            cls_in = clf.resize_device(fb, self.roi_xyxy, self.roi_src, n)
            clf.classify_device(cls_in)
        ev[3].record()
        ev[3].synchronize()
        m.t_detection = ev[0].elapsed_time(ev[1])
        m.t_roi_extract = ev[1].elapsed_time(ev[2])
        m.t_classification = ev[2].elapsed_time(ev[3])
        results: List[Dict] = []
        if n:
            src = self.roi_src[:n].cpu().numpy()
            probs = clf.probs[:n].cpu().numpy()
            cls_ids = np.argmax(probs, axis=1)
            m.cls_confidence_avg = float(np.mean(probs.max(axis=1)))
            for r in range(n):
                k = int(src[r, 1])
                results.append({
                    "bbox": tuple(dets[0][k].astype(int)),
                    "det_class": int(dets[2][k]),
                    "det_conf": float(dets[1][k]),
                    "cls_class": int(cls_ids[r]),
                    "cls_conf": float(np.max(probs[r])),
                    "time_det": m.t_detection / n,
                    "time_cls": m.t_classification / n,
                })
        m.t_total = (time.perf_counter() - t_start) * 1000
        m.fps = 1000.0 / m.t_total if m.t_total > 0 else 0
        try:
            import psutil
            m.cpu_percent = psutil.cpu_percent()
            m.memory_mb = psutil.Process().memory_info().rss / 1024 / 1024
        except Exception:
            pass
        return results, m
