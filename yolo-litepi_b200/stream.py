"""Host-buffer streaming front end of ``B200Pipeline`` (public API: ``B200Pipeline.stream()`` /
``run_stream``).

The reference processes one ``cv2.imread`` frame at a time (``src/vntsr/pipeline/e2e.py:962-973``,
``main`` loop ``:1108-1116``).  On a B200 the same call pattern would leave the GPU idle behind the
PCIe copy, so the streaming entry point owns everything a host-resident frame source needs:

* a ring of PINNED host staging buffers (``host_buffer(i)``: a capture loop can decode straight into
  them -- ``cv2.VideoCapture.read(dst)``, ``np.copyto``) and matching device frame buffers;
* one copy stream that keeps the PCIe link busy back to back;
* ``lanes`` pipeline instances on their own CUDA streams (batch k+1 starts while the partially filled
  last waves of batch k's persistent kernels drain);
* the whole device step of a lane captured ONCE per (buffer, thresholds, frame shapes) in a CUDA graph
  (all ~75 kernel launches of a step, the ROI count stays on the device) and replayed per batch;
* the read-back of the packed detection records (pinned, double-buffered).

Frames handed in as ordinary pageable ``np.ndarray`` (what ``cv2.imread`` returns) are first copied into
the pinned ring by a small thread pool (numpy releases the GIL while copying).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .detector import FrameBatch
from .jpeg import JpegBatchDecoder, is_jpeg

Frames = Union[np.ndarray, Sequence[np.ndarray]]


class JpegBatch:
    """A batch of same-header JPEGs already packed in PINNED host memory (``StreamRunner.pack_jpeg_batch``): the
    compressed counterpart of a frame batch in the pinned ring -- the copy engine reads it directly."""

    def __init__(self, hit, data: torch.Tensor, off: torch.Tensor, used: int, n: int):
        self.hit, self.data, self.off, self.used, self.n = hit, data, off, int(used), int(n)

    def __len__(self):
        return self.n


class StreamRunner:
    def __init__(self, pipe, lanes: int = 2, use_graph: Optional[bool] = None, copy_threads: Optional[int] = None):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.device = pipe.device
        self.pipes = [pipe] + [pipe.clone() for _ in range(lanes - 1)]
        self.n_lanes = lanes
        self.n_buf = 2 * lanes
        self.B = pipe.max_batch
        if use_graph is None:
            use_graph = os.environ.get("LP_NO_GRAPH", "0") != "1"
        self.use_graph = bool(use_graph) and all(p.classifier.fused for p in self.pipes)
        with torch.cuda.device(self.device):
            self.lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
            self.copy_stream = torch.cuda.Stream(device=self.device)
        self.pool = ThreadPoolExecutor(max_workers=copy_threads or max(1, min(8, (os.cpu_count() or 2) - 1)))
        self.shape: Optional[Tuple[int, int]] = None
        self.graph_failed: Optional[str] = None
        self.steps_graph = self.steps_direct = 0
        self.graph_kernels = {}             # kernels inside each captured graph (lp_launch_count delta of its capture)
        self.replayed_kernels = 0
        # JPEG ingest (frames handed in as encoded bytes): one decoder (tables + scratch) per lane, pinned byte ring
        self.jpeg = [JpegBatchDecoder(p.ctx, self.device, self.B) for p in self.pipes]
        self.jcap = 0
        self.h2d_bytes = 0                  # bytes copied host -> device so far (frames or JPEG scans)

    # ------------------------------------------------------------------ buffers
    def _ensure(self, h: int, w: int) -> None:
        if self.shape == (h, w):
            return
        torch.cuda.synchronize(self.device)
        B, nb = self.B, self.n_buf
        with torch.cuda.device(self.device):
            self.host = [torch.empty((B, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(nb)]
            self.host_np = [t.numpy() for t in self.host]
            self.dev = [torch.empty((B, h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(nb)]
            self.fid_h = [torch.zeros((B,), dtype=torch.int32).pin_memory() for _ in range(nb)]
            self.fid = [torch.zeros((B,), dtype=torch.int32, device=self.device) for _ in range(nb)]
            self.ready = [torch.cuda.Event() for _ in range(nb)]       # H2D of buffer b landed
            self.consumed = [torch.cuda.Event() for _ in range(nb)]    # the step that read device buffer b is done
        self.fb = [FrameBatch.from_device(d) for d in self.dev]
        self.used = [False] * nb
        self.graphs = {}
        self.shape = (h, w)

    def _ensure_jpeg(self, need: int) -> None:
        if need <= self.jcap:
            return
        torch.cuda.synchronize(self.device)
        cap = int(need * 1.25) + 4096
        with torch.cuda.device(self.device):
            self.jhost = [torch.empty((cap,), dtype=torch.uint8).pin_memory() for _ in range(self.n_buf)]
            self.jhost_np = [t.numpy() for t in self.jhost]
            self.jdev = [torch.empty((cap,), dtype=torch.uint8, device=self.device) for _ in range(self.n_buf)]
            self.joff_h = [torch.zeros((self.B + 1,), dtype=torch.int64).pin_memory() for _ in range(self.n_buf)]
            self.joff_np = [t.numpy() for t in self.joff_h]
            self.joff = [torch.zeros((self.B + 1,), dtype=torch.int64, device=self.device) for _ in range(self.n_buf)]
        self.jcap = cap

    def pack_jpeg_batch(self, jpegs: Sequence) -> JpegBatch:
        """Concatenate the entropy-coded scans of same-header JPEG byte strings into one pinned buffer (+ offsets)."""
        n = len(jpegs)
        if n < 1 or n > self.B:
            raise ValueError(f"a batch holds 1..{self.B} frames, got {n}")
        total = sum(len(j) for j in jpegs)
        data = torch.empty((max(total, 1),), dtype=torch.uint8).pin_memory()
        off = torch.zeros((self.B + 1,), dtype=torch.int64).pin_memory()
        hit, used = self.jpeg[0].stage(jpegs, data.numpy(), off.numpy())
        off[n + 1:] = used
        return JpegBatch(hit, data, off, used, n)

    def host_buffer(self, i: int, h: int, w: int) -> np.ndarray:
        """Pinned staging buffer ``i % n_buf`` as a [max_batch, h, w, 3] uint8 array.  Blocks until the copy engine has
        finished reading its previous contents, so a producer may overwrite it."""
        self._ensure(h, w)
        b = i % self.n_buf
        if self.used[b]:
            self.ready[b].synchronize()
        return self.host_np[b]

    # ------------------------------------------------------------------ one step
    def _stage(self, b: int, frames: Frames):
        """frames -> the tensor the copy engine reads for this step: pinned ring buffer b after a host copy, the ring
        buffer itself when the frames were produced in place, or the caller's own tensor when it is already pinned
        host memory or device-resident ([n,H,W,3] uint8 torch tensor).  Returns (frame count, source tensor)."""
        n = len(frames)
        if n < 1 or n > self.B:
            raise ValueError(f"a batch holds 1..{self.B} frames, got {n}")
        if isinstance(frames, JpegBatch):
            hd = frames.hit[0]
            self._ensure(hd.height, hd.width)
            self._ensure_jpeg(frames.used)
            # tables / scratch are per lane: look the header up in this lane's decoder
            hit = self.jpeg[b % self.n_lanes].for_header(hd)
            return n, ("jpegbatch", hit, frames)
        f0 = frames[0]
        if not isinstance(frames, (np.ndarray, torch.Tensor)) and is_jpeg(f0):
            # encoded frames: only the entropy-coded bytes cross PCIe, the decode runs on the lane's stream
            dec = self.jpeg[b % self.n_lanes]
            hit = dec.header(f0)
            self._ensure(hit[0].height, hit[0].width)
            self._ensure_jpeg(sum(len(j) for j in frames))
            if self.used[b]:
                self.ready[b].synchronize()
            hit, used = dec.stage(frames, self.jhost_np[b], self.joff_np[b])
            return n, ("jpeg", hit, used)
        if isinstance(frames, torch.Tensor):
            if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_contiguous():
                raise ValueError("tensor batches must be contiguous [n,H,W,3] uint8")
            self._ensure(int(frames.shape[1]), int(frames.shape[2]))
            if frames.is_cuda or frames.is_pinned():
                return n, frames
            frames = frames.numpy()
            f0 = frames[0]
        if f0.dtype != np.uint8 or f0.ndim != 3 or f0.shape[2] != 3:
            raise ValueError("frames must be HWC BGR uint8")
        h, w = int(f0.shape[0]), int(f0.shape[1])
        self._ensure(h, w)
        dst = self.host_np[b]
        if isinstance(frames, np.ndarray) and frames.ctypes.data == dst.ctypes.data:
            return n, self.host[b]                           # produced in place
        if self.used[b]:
            self.ready[b].synchronize()                      # the previous H2D out of this buffer has finished
        for f in frames:
            if f.shape != (h, w, 3):
                raise ValueError("run_stream: all frames of a stream share one shape (use run_batch for ragged batches)")
        list(self.pool.map(lambda i: np.copyto(dst[i], frames[i]), range(n)))
        return n, self.host[b]

    def _capture(self, b: int, key) -> Optional[torch.cuda.CUDAGraph]:
        g = self._capture_once(b, key)
        if g is None and self.graph_failed and not getattr(self, "_pdl_off", False):
            # a driver that cannot capture programmatic-dependent-launch edges: retry without them
            self._pdl_off = True
            for p in self.pipes:
                p.detector.ctx.set_pdl(False)
                p.classifier.ctx.set_pdl(False)
            first = self.graph_failed
            self.use_graph, self.graph_failed = True, None
            g = self._capture_once(b, key)
            self.graph_note = f"captured without PDL after: {first}" if g is not None else None
        return g

    def _capture_once(self, b: int, key) -> Optional[torch.cuda.CUDAGraph]:
        conf, iou, min_area = key
        ln = b % self.n_lanes
        pipe, st = self.pipes[ln], self.lane_streams[ln]
        slot = b // self.n_lanes
        try:
            with torch.cuda.device(self.device):
                with torch.cuda.stream(st):                    # warm-up outside the capture: lazy one-time setup
                    pipe.enqueue_device(self.fb[b], conf, iou, min_area, self.fid[b], slot=slot)
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                k0 = pipe.counters.launch_count()
                with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                    pipe.enqueue_device(self.fb[b], conf, iou, min_area, self.fid[b], slot=slot)
                    pipe.enqueue_fetch_copy(slot)
                self.graph_kernels[(b, key)] = pipe.counters.launch_count() - k0
            return g
        except Exception as e:                                 # capture unsupported here: keep the direct path
            self.graph_failed = f"{type(e).__name__}: {e}"
            self.use_graph = False
            torch.cuda.synchronize(self.device)
            return None

    def _launch(self, s: int, n: int, conf: float, iou: float, min_area: int, frame_ids, h2d: bool, src=None) -> None:
        b, ln = s % self.n_buf, s % self.n_lanes
        slot = b // self.n_lanes
        pipe, st = self.pipes[ln], self.lane_streams[ln]
        with torch.cuda.device(self.device):
            if h2d:
                if self.used[b]:
                    self.ready[b].synchronize()             # the previous H2D out of the pinned id table has finished
                if frame_ids is None:
                    self.fid_h[b][:n] = torch.arange(n, dtype=torch.int32)
                else:
                    self.fid_h[b][:n] = torch.as_tensor(frame_ids, dtype=torch.int32)
                jpeg = isinstance(src, tuple) and src[0] in ("jpeg", "jpegbatch")
                with torch.cuda.stream(self.copy_stream):
                    if self.used[b]:
                        self.copy_stream.wait_event(self.consumed[b])
                    if jpeg and src[0] == "jpegbatch":
                        jb = src[2]
                        used = max(jb.used, 1)
                        self.jdev[b][:used].copy_(jb.data[:used], non_blocking=True)
                        self.joff[b].copy_(jb.off, non_blocking=True)
                        self.h2d_bytes += used + 8 * (self.B + 1)
                    elif jpeg:
                        used = max(int(src[2]), 1)
                        self.jdev[b][:used].copy_(self.jhost[b][:used], non_blocking=True)
                        self.joff[b].copy_(self.joff_h[b], non_blocking=True)
                        self.h2d_bytes += used + 8 * (self.B + 1)
                    else:
                        s_t = (self.host[b] if src is None else src)[:n]
                        self.dev[b][:n].copy_(s_t, non_blocking=True)
                        self.h2d_bytes += 0 if s_t.is_cuda else s_t.numel()
                    self.fid[b][:n].copy_(self.fid_h[b][:n], non_blocking=True)
                    self.ready[b].record(self.copy_stream)
                self.used[b] = True
            with torch.cuda.stream(st):
                if h2d:
                    st.wait_event(self.ready[b])
                    if jpeg:                               # JPEG bytes -> BGR frames in the device ring (csrc/jpeg.cu)
                        import ctypes as _C
                        self.jpeg[ln].decode_device(src[1], self.jdev[b], self.joff[b], n, self.dev[b], _C.c_void_p(st.cuda_stream))
                g = None
                if self.use_graph and n == self.B:
                    key = (float(conf), float(iou), int(min_area))
                    g = self.graphs.get((b, key))
                    if g is None:
                        g = self._capture(b, key)
                        if g is not None:
                            self.graphs[(b, key)] = g
                if g is not None:
                    g.replay()
                    pipe.mark_enqueued(slot, n)
                    self.steps_graph += 1
                    self.replayed_kernels += self.graph_kernels[(b, key)]
                else:
                    fb = self.fb[b] if n == self.B else FrameBatch([self.dev[b][i] for i in range(n)])
                    pipe.enqueue_device(fb, conf, iou, min_area, self.fid[b], slot=slot)
                    if not pipe.classifier.fused:          # layer-by-layer classifiers are sized on the host (one sync per step)
                        pipe.finish(fb, self.fid[b], slot)
                    pipe.enqueue_fetch_copy(slot)
                    self.steps_direct += 1
                pipe.mark_fetch(slot)
                if h2d:
                    self.consumed[b].record(st)

    def _collect(self, s: int) -> np.ndarray:
        b, ln = s % self.n_buf, s % self.n_lanes
        return self.pipes[ln].collect(b // self.n_lanes)

    # ------------------------------------------------------------------ public
    def run_stream(self, batches: Iterable[Frames], conf_threshold: float = 0.5, iou_threshold: float = 0.45,
                   min_area: int = 100, frame_ids: Optional[Iterable[Sequence[int]]] = None) -> Iterator[np.ndarray]:
        """For each batch of host frames (<= max_batch, one shape per stream) yield its packed detection records
        ([n, 9] int32, ``B200Pipeline.records_to_results`` turns them into the reference's dicts), in order.  The
        results of batch s are yielded while batches s+1 .. s+lanes are already in flight."""
        ids_it = iter(frame_ids) if frame_ids is not None else None
        s = 0
        pending: List[int] = []
        for frames in batches:
            n, src = self._stage(s % self.n_buf, frames)
            self._launch(s, n, conf_threshold, iou_threshold, min_area, next(ids_it) if ids_it is not None else None, True, src)
            pending.append(s)
            s += 1
            if len(pending) > self.n_lanes:
                yield self._collect(pending.pop(0))
        while pending:
            yield self._collect(pending.pop(0))

    def run_one(self, frames: Frames, conf_threshold: float = 0.5, iou_threshold: float = 0.45,
                min_area: int = 100) -> np.ndarray:
        """Synchronous single batch (latency path): stage, H2D, graph replay, read back."""
        n, src = self._stage(0, frames)
        self._launch(0, n, conf_threshold, iou_threshold, min_area, None, True, src)
        return self._collect(0)

    def replay_resident(self, steps: int, conf_threshold: float, iou_threshold: float, min_area: int) -> None:
        """Enqueue ``steps`` steps on frames that are already resident in the device ring (no H2D): the
        device-resident throughput measurement.  Collect with ``drain_resident``."""
        if self.shape is None:
            raise RuntimeError("fill the ring first (run_stream / run_one)")
        for s in range(steps):
            if s >= self.n_lanes:
                self._collect(s - self.n_lanes)
            self._launch(s, self.B, conf_threshold, iou_threshold, min_area, None, False)

    def drain_resident(self, steps: int) -> List[np.ndarray]:
        return [self._collect(s) for s in range(max(0, steps - self.n_lanes), steps)]

    def launch_count(self) -> int:
        """Kernels launched so far: direct launches of every lane + the kernels inside replayed graphs."""
        return sum(p.counters.launch_count() for p in self.pipes) + self.replayed_kernels

    def close(self) -> None:
        self.pool.shutdown(wait=False)
