"""B200Detector -- drop-in for the reference's ``NCNNDetector``
(``src/vntsr/pipeline/e2e.py:195-316``) running on one B200.

Same constructor arguments and the same ``detect(image, conf, iou) -> (boxes, scores,
class_ids)`` contract (xyxy float32 in original pixels, class asc / score desc order,
float64 empties), plus ``detect_batch`` / device-resident entry points.  PyTorch is
used only to own device memory; every computation is a CUDA kernel behind the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .ncnn_model import load_ncnn
from .plan import build_detector_plan


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FrameBatch:
    """Device-resident frames + the host-side tables the C-ABI takes.  Keeps the tensors
    alive; ``frames`` are HWC BGR uint8 (cv2.imread layout, e2e.py:962)."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.tensors = list(tensors)
        n = len(self.tensors)
        self.n = n
        self.ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in self.tensors])
        self.h = (C.c_int32 * max(n, 1))(*[int(t.shape[0]) for t in self.tensors])
        self.w = (C.c_int32 * max(n, 1))(*[int(t.shape[1]) for t in self.tensors])
        self.pitch = (C.c_int64 * max(n, 1))(*[int(t.stride(0)) for t in self.tensors])
        self.max_side = max([max(int(t.shape[0]), int(t.shape[1])) for t in self.tensors], default=1)
        self.nbytes = sum(int(t.numel()) for t in self.tensors)

    @staticmethod
    def from_host(frames: Sequence[np.ndarray], device: torch.device) -> "FrameBatch":
        out = []
        for f in frames:
            if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise ValueError("frames must be HWC BGR uint8")
            out.append(torch.from_numpy(np.ascontiguousarray(f)).to(device, non_blocking=True))
        return FrameBatch(out)

    @staticmethod
    def from_device(batch: torch.Tensor) -> "FrameBatch":
        """[B,H,W,3] uint8 CUDA tensor (or a list of [H,W,3] tensors)."""
        if isinstance(batch, torch.Tensor):
            if batch.dtype != torch.uint8 or batch.dim() != 4 or batch.shape[3] != 3 or batch.stride(2) != 3:
                raise ValueError("device batch must be [B,H,W,3] uint8 with packed pixels")
            return FrameBatch([batch[i] for i in range(batch.shape[0])])
        return FrameBatch(list(batch))


class B200Detector:
    def __init__(self, param_path: str, bin_path: Optional[str], input_size: int = 640,
                 use_gpu: bool = False, num_threads: int = 4,
                 input_name: str = "in0", output_name: str = "out0",
                 device: int = 0, max_batch: int = 64, max_det: int = 1024, seed: int = 0,
                 tensor_cores: bool = True):
        # use_gpu / num_threads / input_name / output_name are accepted for signature parity with
        # NCNNDetector (e2e.py:198-200) and ignored: there is one device path.
        self.input_size = int(input_size)
        if self.input_size < 32 or self.input_size % 32:
            raise ValueError(f"input_size must be a positive multiple of 32 (Detect strides 8/16/32), got {input_size}")
        self.input_name, self.output_name = input_name, output_name
        if not torch.cuda.is_available():
            raise RuntimeError("litepi_b200: no CUDA device; the B200 backend has no CPU fallback")
        self.device = torch.device("cuda", device)
        self.ctx = L.Context(device)              # one lp_ctx per detector object (it holds the plan)
        self.model = load_ncnn(param_path, bin_path, seed=seed)       # RuntimeError on failure (e2e.py:213-216)
        self.plan = build_detector_plan(self.model, self.input_size)
        self.max_batch, self.max_det = int(max_batch), int(max_det)
        self.n_anchors = self.plan.meta["n_anchors"]
        self.nc = self.plan.meta["nc"]
        ws_bytes = self.plan.layout(self.max_batch)
        tc_blob = self.plan.pack_tc_weights() if tensor_cores else np.zeros(0, np.uint8)
        self.tc_ops = sum(1 for o in self.plan.ops if o["wtc_off"] >= 0)
        with torch.cuda.device(self.device):
            self.weights = torch.from_numpy(self.plan.weights()).to(self.device)
            self.weights_tc = torch.from_numpy(tc_blob).to(self.device) if tc_blob.size else None
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
            bufs, ops = self.plan.c_arrays()
            L.check(L.lib().lp_net_load(self.ctx.handle, L.NET_DETECTOR, bufs, len(bufs), ops, len(ops),
                                        _ptr(self.weights), self.weights.numel(), _ptr(self.weights_tc),
                                        tc_blob.size, self.max_batch),
                    "lp_net_load(detector)")
            B, S, A, D = self.max_batch, self.input_size, self.n_anchors, self.max_det
            self.lb = torch.empty((B, S, S, 3), dtype=torch.uint8, device=self.device)
            self.out0 = torch.empty((B, 4 + self.nc, A), dtype=torch.float32, device=self.device)
            self.boxes = torch.empty((B, D, 4), dtype=torch.float32, device=self.device)
            self.scores = torch.empty((B, D), dtype=torch.float32, device=self.device)
            self.classes = torch.empty((B, D), dtype=torch.int64, device=self.device)
            self.keep_idx = torch.empty((B, D), dtype=torch.int32, device=self.device)
            self.counts = torch.zeros((B,), dtype=torch.int32, device=self.device)
            self.n_cand = torch.zeros((B,), dtype=torch.int32, device=self.device)
            self.nms_scratch = torch.empty(L.lib().lp_decode_nms_scratch_bytes(B, A), dtype=torch.uint8,
                                           device=self.device)
        self.ratio = (C.c_double * B)()
        self.pad = (C.c_double * (2 * B))()

    # ------------------------------------------------------------------ stages (device, async)
    def letterbox_device(self, fb: FrameBatch) -> torch.Tensor:
        """K1 on ``fb``; returns the [n,S,S,3] RGB uint8 view and fills self.ratio / self.pad."""
        if fb.n > self.max_batch:
            raise ValueError(f"batch {fb.n} > max_batch {self.max_batch}")
        L.check(L.lib().lp_letterbox(self.ctx.handle, fb.ptrs, fb.h, fb.w, fb.pitch, fb.n, self.input_size,
                                     _ptr(self.lb), self.ratio, self.pad, _stream()), "lp_letterbox")
        return self.lb[:fb.n]

    def forward_device(self, lb: torch.Tensor) -> torch.Tensor:
        """K2/K3: [n,S,S,3] RGB u8 -> out0 [n,4+nc,A] f32 (reference layout)."""
        n = int(lb.shape[0])
        L.check(L.lib().lp_detect_forward(self.ctx.handle, _ptr(lb), n, _ptr(self.workspace),
                                          self.workspace.numel(), _ptr(self.out0), _stream()), "lp_detect_forward")
        return self.out0[:n]

    def decode_nms_device(self, out0: torch.Tensor, h, w, ratio, pad, conf: float, iou: float) -> None:
        """K4+K5 into self.boxes/scores/classes/keep_idx/counts/n_cand."""
        n = int(out0.shape[0])
        hh = (C.c_int32 * n)(*[int(v) for v in h])
        ww = (C.c_int32 * n)(*[int(v) for v in w])
        rr = (C.c_float * n)(*[float(np.float32(v)) for v in ratio])
        pp = (C.c_float * (2 * n))(*[float(np.float32(v)) for v in pad])
        L.check(L.lib().lp_decode_nms(self.ctx.handle, _ptr(out0), int(out0.shape[1]) - 4, int(out0.shape[2]),
                                      hh, ww, rr, pp, n, float(conf), float(iou), self.max_det,
                                      _ptr(self.boxes), _ptr(self.scores), _ptr(self.classes), _ptr(self.keep_idx),
                                      _ptr(self.counts), _ptr(self.n_cand), _ptr(self.nms_scratch),
                                      self.nms_scratch.numel(), _stream()), "lp_decode_nms")

    def detect_device(self, fb: FrameBatch, conf: float, iou: float) -> None:
        lb = self.letterbox_device(fb)
        out0 = self.forward_device(lb)
        self.decode_nms_device(out0, fb.h[:fb.n], fb.w[:fb.n], self.ratio[:fb.n], self.pad[:2 * fb.n], conf, iou)

    def _collect(self, n: int):
        counts = self.counts[:n].cpu().numpy()
        if counts.size and int(counts.max()) > self.max_det:
            raise RuntimeError(f"litepi_b200: {int(counts.max())} detections in one frame exceed max_det="
                               f"{self.max_det}; construct the detector with a larger max_det")
        boxes, scores, classes = self.boxes[:n].cpu().numpy(), self.scores[:n].cpu().numpy(), self.classes[:n].cpu().numpy()
        out = []
        for i in range(n):
            k = int(counts[i])
            if k == 0:        # reference returns float64 empties (e2e.py:264, 292-294)
                out.append((np.empty((0, 4)), np.empty((0,)), np.empty((0,))))
            else:
                out.append((boxes[i, :k].copy(), scores[i, :k].copy(), classes[i, :k].copy()))
        return out

    # ------------------------------------------------------------------ reference-shaped API
    def preprocess(self, image: np.ndarray):
        """e2e.py:222-238: returns (float32 [3,S,S] RGB/255 CUDA tensor, ratio, (dw, dh))."""
        fb = FrameBatch.from_host([image], self.device)
        lb = self.letterbox_device(fb)
        x = (lb[0].permute(2, 0, 1).to(torch.float32) / 255.0).contiguous()
        return x, float(self.ratio[0]), (float(self.pad[0]), float(self.pad[1]))

    def letterbox(self, image: np.ndarray):
        """K1 alone: (RGB u8 [S,S,3] numpy, ratio, (dw,dh)) for stage-wise parity tests."""
        fb = FrameBatch.from_host([image], self.device)
        lb = self.letterbox_device(fb)
        return lb[0].cpu().numpy(), float(self.ratio[0]), (float(self.pad[0]), float(self.pad[1]))

    def forward(self, x) -> np.ndarray:
        """Detector forward on letterboxed RGB u8 [n,S,S,3] (numpy or CUDA tensor) -> out0 numpy."""
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x)).to(self.device)
        outs = []
        for i in range(0, x.shape[0], self.max_batch):
            outs.append(self.forward_device(x[i:i + self.max_batch].contiguous()).cpu().numpy().copy())
        return np.concatenate(outs, 0)

    def postprocess(self, output, orig_shape: Tuple[int, int], ratio: float, pad: Tuple[float, float],
                    conf_threshold: float = 0.5, iou_threshold: float = 0.45):
        """e2e.py:240-296 on a [4+nc, A] (or [1,4+nc,A]) array."""
        o = np.asarray(output, dtype=np.float32)
        if o.ndim == 2:
            o = o[None]
        if o.shape[-1] == 84:                      # e2e.py:248-249
            o = o.transpose(0, 2, 1)
        t = torch.from_numpy(np.ascontiguousarray(o[:1])).to(self.device)
        if t.shape[2] > 16384:
            raise ValueError("at most 16384 anchors")
        if t.shape[2] > self.n_anchors or t.shape[1] != 4 + self.nc:
            scratch = torch.empty(L.lib().lp_decode_nms_scratch_bytes(1, t.shape[2]), dtype=torch.uint8, device=self.device)
            keep, self.nms_scratch = self.nms_scratch, scratch
            try:
                self.decode_nms_device(t, [orig_shape[0]], [orig_shape[1]], [ratio], list(pad), conf_threshold, iou_threshold)
            finally:
                self.nms_scratch = keep
        else:
            self.decode_nms_device(t, [orig_shape[0]], [orig_shape[1]], [ratio], list(pad), conf_threshold, iou_threshold)
        return self._collect(1)[0]

    def detect(self, image: np.ndarray, conf_threshold: float = 0.5, iou_threshold: float = 0.45):
        """e2e.py:298-316."""
        return self.detect_batch([image], conf_threshold, iou_threshold)[0]

    def detect_batch(self, images: Sequence[np.ndarray], conf_threshold: float = 0.5, iou_threshold: float = 0.45):
        out = []
        for i in range(0, len(images), self.max_batch):
            fb = FrameBatch.from_host(images[i:i + self.max_batch], self.device)
            self.detect_device(fb, conf_threshold, iou_threshold)
            out.extend(self._collect(fb.n))
        return out
