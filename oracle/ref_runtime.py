"""Runs the reference's OWN pipeline code on the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py).

``/root/reference/src/vntsr/pipeline/e2e.py`` is imported UNMODIFIED -- from /root/reference in the
authoring container, or from the copy ``__graft_entry__.build()`` stages under ``oracle/_ref/e2e.py``
(git-ignored, travels to the GPU box) -- and its ``HybridPipeline.run`` (e2e.py:443-531) is driven
exactly as ``process_image`` drives it (e2e.py:973).  Three of its imports are absent from this image:

* ``matplotlib`` / ``seaborn``: plotting only, stubbed with empty modules;
* ``ncnn`` ("manual_build", requirements.txt:54-58; not installable offline): replaced by the small
  shim below, which implements the five calls ``NCNNDetector`` makes (e2e.py:209-216, 227-236,
  305-307) on top of **OpenCV-DNN executing the reference's own exported graph and trained weights**
  (``yolo_plus.onnx``, the ORT twin of the ncnn export: same weights, verified bit-identical on
  model.0).  When no ONNX export is available for a graph (the TT100K v2 weights are missing from the
  reference repo) the shim executes ``model.ncnn.param`` with the torch fp32 graph oracle instead.

So every line of Python that runs is the reference's; the detector's native runtime is the stand-in
named in ``runtime``.  Used by ``bench.py --impl reference`` / ``cpu_baseline`` and by the pinning tests.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
STAGE = os.path.join(ROOT, "oracle", "_ref")

_state = {"onnx_hint": None}


def e2e_source_path() -> Optional[str]:
    for p in (os.path.join(REF, "src/vntsr/pipeline/e2e.py"), os.path.join(STAGE, "e2e.py")):
        if os.path.exists(p):
            return p
    return None


def _find_onnx(param_path: str) -> Optional[str]:
    """The ONNX export that belongs to a ``model.ncnn.param``: reference layout
    ``.../yolo_plus/yolo_plus_ncnn_model/model.ncnn.param`` <-> ``.../yolo_plus/yolo_plus.onnx``;
    staged layout ``oracle/_ref/<set>.model.ncnn.param`` <-> ``oracle/_ref/<set>.yolo_plus.onnx``."""
    d, base = os.path.dirname(os.path.abspath(param_path)), os.path.basename(param_path)
    cands = [os.path.join(os.path.dirname(d), "yolo_plus.onnx")]
    if base.endswith(".model.ncnn.param"):
        cands.append(os.path.join(d, base[:-len(".model.ncnn.param")] + ".yolo_plus.onnx"))
    for c in cands:
        if os.path.exists(c):
            return c
    return None


# ------------------------------------------------------------------------------------ ncnn shim
class _Mat:
    """ncnn.Mat as NCNNDetector uses it: planar float32 [c, h, w] (from_pixels, e2e.py:227-232),
    in-place normalisation (:234-236), ``np.array(mat)`` (:242)."""

    class PixelType:
        PIXEL_RGB = 1
        PIXEL_BGR = 2

    def __init__(self, data: np.ndarray):
        self.data = np.ascontiguousarray(data, dtype=np.float32)

    @staticmethod
    def from_pixels(pixels: np.ndarray, pixel_type: int, w: int, h: int) -> "_Mat":
        assert pixels.dtype == np.uint8 and pixels.shape[:2] == (h, w)
        return _Mat(pixels.transpose(2, 0, 1).astype(np.float32))

    def substract_mean_normalize(self, mean_vals, norm_vals) -> None:          # ncnn: (x - mean) * norm, float32
        for c in range(self.data.shape[0]):
            if mean_vals:
                self.data[c] -= np.float32(mean_vals[c])
            if norm_vals:
                self.data[c] *= np.float32(norm_vals[c])

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class _Opt:
    use_vulkan_compute = False
    num_threads = 4


class _Extractor:
    def __init__(self, net: "_Net"):
        self.net, self.x = net, None

    def input(self, name: str, mat: _Mat) -> int:
        self.x = mat.data[None]
        return 0

    def extract(self, name: str):
        return 0, _Mat(self.net._forward(self.x)[0])


class _Net:
    """ncnn.Net: load_param / load_model / create_extractor (e2e.py:209-216, :305)."""

    def __init__(self):
        self.opt = _Opt()
        self._param = self._bin = None
        self._dnn = self._graph = None
        self.runtime = None

    def load_param(self, path: str) -> int:
        if not os.path.exists(path):
            return -1
        self._param = path
        return 0

    def load_model(self, path) -> int:
        if path is not None and not os.path.exists(path):
            return -1
        self._bin = path
        return 0

    def _build(self):
        import cv2
        onnx = _state["onnx_hint"] or _find_onnx(self._param)
        if onnx and self._bin is not None:
            self._dnn = cv2.dnn.readNetFromONNX(onnx)
            self.runtime = "OpenCV-DNN(yolo_plus.onnx)"
        else:
            from oracle.ncnn_graph import DetectorOracle
            self._graph = DetectorOracle(self._param, self._bin, seed=0)
            self.runtime = "torch-fp32 graph executor(model.ncnn.param)"

    def _forward(self, x: np.ndarray) -> np.ndarray:
        import cv2
        if self._dnn is None and self._graph is None:
            self._build()
        if self._dnn is not None:
            cv2.setNumThreads(int(self.opt.num_threads))       # net.opt.num_threads (e2e.py:211)
            self._dnn.setInput(np.ascontiguousarray(x))
            return self._dnn.forward()
        import torch
        torch.set_num_threads(int(self.opt.num_threads))
        return self._graph.forward(x).numpy()

    def create_extractor(self) -> _Extractor:
        return _Extractor(self)


def _install_stubs() -> None:
    for n in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    shim = types.ModuleType("ncnn")
    shim.Net, shim.Mat, shim.Extractor = _Net, _Mat, _Extractor
    shim.__doc__ = "OpenCV-DNN-backed stand-in for the ncnn Python binding (oracle/ref_runtime.py)"
    sys.modules["ncnn"] = shim


_ref_module = None


def load_reference():
    """Import the reference's e2e.py unmodified; returns the module or None if it is nowhere to be found."""
    global _ref_module
    if _ref_module is None:
        src = e2e_source_path()
        if src is None:
            return None
        _install_stubs()
        spec = importlib.util.spec_from_file_location("ref_e2e", src)
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        _ref_module = mod
    return _ref_module


class ReferencePipeline:
    """The reference's ``HybridPipeline`` (unmodified class) with the classifier's seeded random-init
    weights shared with the GPU path (the reference ships no classifier weights: e2e.py:337-343 keeps the
    random init when the file is missing)."""

    def __init__(self, param: str, binp: Optional[str], num_classes: int, state_dict=None, threads: int = 4,
                 onnx: Optional[str] = None, arch: str = "shufflenetv2"):
        ref = load_reference()
        if ref is None:
            raise FileNotFoundError("reference e2e.py not available (neither /root/reference nor oracle/_ref/e2e.py)")
        import torch
        self.ref, self.threads = ref, int(threads)
        _state["onnx_hint"] = onnx
        torch.set_num_threads(self.threads)
        with contextlib.redirect_stdout(io.StringIO()):           # the constructors print banners; bench prints ONE line
            self.pipe = ref.HybridPipeline(param, binp, "/nonexistent/shufflenetv2.pth", arch, num_classes=num_classes,
                                           detector_threads=self.threads, classifier_device="cpu", batch_size=8)
        self.pipe.detector.net._build()
        _state["onnx_hint"] = None
        if state_dict is not None:
            self.pipe.classifier.model.load_state_dict(state_dict)
        self.runtime = self.pipe.detector.net.runtime

    def set_threads(self, n: int) -> None:
        import torch
        self.threads = int(n)
        self.pipe.detector.net.opt.num_threads = self.threads
        torch.set_num_threads(self.threads)

    def run(self, frame: np.ndarray, conf: float, iou: float, min_area: int):
        return self.pipe.run(frame, conf_threshold=conf, iou_threshold=iou, min_area=min_area)
