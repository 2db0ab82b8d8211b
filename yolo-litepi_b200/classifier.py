"""B200Classifier -- drop-in for the reference's ``PyTorchClassifier``
(``src/vntsr/pipeline/e2e.py:350-396``): ShuffleNetV2 x1.0 at 64x64 on one B200.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .detector import FrameBatch, _ptr, _stream
from .plan import build_classifier_plan, build_fused_classifier


def _random_state_dict(num_classes: int, seed: Optional[int], arch: str = "shufflenetv2"):
    """The reference builds the torchvision model with ``weights=None`` and swaps the classification head
    (e2e.py:322-333); a missing weight file silently keeps that random init (e2e.py:337-343)."""
    from .cls_archs import torchvision_model
    if seed is not None:
        torch.manual_seed(seed)
    return torchvision_model(arch, num_classes).state_dict()


class B200Classifier:
    def __init__(self, model_path: Optional[str], arch: str = "shufflenetv2", num_classes: int = 58,
                 input_size: int = 64, device="cpu", state_dict: Optional[dict] = None,
                 cuda_device: int = 0, max_batch: int = 256, seed: Optional[int] = None,
                 tensor_cores: bool = True, fused: bool = True, fused_group: int = 3):
        # `device` is the reference's torch device string (e2e.py:354); this backend always runs on
        # cuda:`cuda_device`.  The four architectures of build_classifier (e2e.py:322-335); ValueError like :335 otherwise.
        if arch not in ("shufflenetv2", "resnet18", "mobilenetv2", "efficientnet"):
            raise ValueError(f"Unknown architecture: {arch}")
        if not torch.cuda.is_available():
            raise RuntimeError("litepi_b200: no CUDA device; the B200 backend has no CPU fallback")
        self.arch, self.num_classes, self.input_size = arch, int(num_classes), int(input_size)
        self.device = torch.device("cuda", cuda_device)
        self.ctx = L.Context(cuda_device)         # one lp_ctx per classifier object
        sd = state_dict
        if sd is None:
            sd = _random_state_dict(self.num_classes, seed, arch)
            if model_path and os.path.exists(model_path):
                try:
                    loaded = torch.load(model_path, map_location="cpu")
                    ref = {k: v.shape for k, v in sd.items()}
                    if set(loaded) != set(ref) or any(tuple(loaded[k].shape) != tuple(ref[k]) for k in ref):
                        raise RuntimeError(f"state_dict does not match the torchvision {arch} model")
                    sd = loaded
                    print(f"✓ Loaded classifier weights from {model_path}")
                except Exception as e:                                   # e2e.py:342-343
                    print(f"⚠ Warning: Could not load weights: {e}")
        self.state_dict = sd
        if arch == "shufflenetv2":
            self.plan = build_classifier_plan(sd, self.input_size)
        else:
            from .cls_archs import PLAN_BUILDERS
            self.plan = PLAN_BUILDERS[arch](sd, self.input_size)
        if self.plan.meta["num_classes"] != self.num_classes:
            raise ValueError("state_dict fc size does not match num_classes")
        self.max_batch = int(max_batch)
        ws_bytes = self.plan.layout(self.max_batch)
        tc_blob = self.plan.pack_tc_weights() if tensor_cores else np.zeros(0, np.uint8)
        self.tc_ops = sum(1 for o in self.plan.ops if o["wtc_off"] >= 0)
        with torch.cuda.device(self.device):
            self.weights = torch.from_numpy(self.plan.weights()).to(self.device)
            self.weights_tc = torch.from_numpy(tc_blob).to(self.device) if tc_blob.size else None
            # zero-initialised: the padding channels of the shuffled halves are never written
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
            bufs, ops = self.plan.c_arrays()
            L.check(L.lib().lp_net_load(self.ctx.handle, L.NET_CLASSIFIER, bufs, len(bufs), ops, len(ops),
                                        _ptr(self.weights), self.weights.numel(), _ptr(self.weights_tc),
                                        tc_blob.size, self.max_batch),
                    "lp_net_load(classifier)")
        # the production path: the whole network in one persistent kernel (csrc/shufflenet_fused.cu); the
        # layer-by-layer plan above stays loaded as the cross-check (set_fused(False))
        self.fused = self._has_fused = bool(fused) and self.input_size == 64 and arch == "shufflenetv2"
        if self.fused:
            prog = build_fused_classifier(sd, in_size=self.input_size, tail_group=fused_group,
                                          tail_mma=os.environ.get("LP_CLS_TAIL_MMA", "0") == "1")
            with torch.cuda.device(self.device):
                self.fused_steps = torch.from_numpy(prog.steps).to(self.device)
                self.fused_weights = torch.from_numpy(prog.weights).to(self.device)
                self.fused_weights16 = torch.from_numpy(prog.weights16.view(np.int16)).to(self.device)
                n_sm = int(L.lib().lp_sm_count(self.ctx.handle))
                self.fused_park = torch.empty(n_sm * prog.tail_group * prog.park_floats, dtype=torch.float32, device=self.device)
            L.check(L.lib().lp_fused_classifier_load(self.ctx.handle, _ptr(self.fused_steps), prog.n_front, prog.n_mid,
                                                     prog.n_tail, _ptr(self.fused_weights), _ptr(self.fused_weights16),
                                                     prog.tail_group, self.input_size, self.num_classes, prog.smem_bytes,
                                                     prog.back_bytes, prog.astage_bytes, prog.tail_bytes, prog.tail_astage_bytes,
                                                     prog.park_floats, _ptr(self.fused_park), self.fused_park.numel() * 4,
                                                     0.18, 0.34), "lp_fused_classifier_load")
        else:
            L.check(L.lib().lp_set_fused_classifier(self.ctx.handle, 0), "lp_set_fused_classifier")
        self._cap = 0
        self._alloc(self.max_batch)

    def set_preprocess(self, mode: str):
        """"pil" (e2e.py:385-388, default) or "cv2" (e2e_optimize.py:391-393) resize in preprocess_batch/predict_batch."""
        if mode not in ("pil", "cv2"):
            raise ValueError(f"Unknown preprocess mode: {mode}")
        L.check(L.lib().lp_set_roi_mode(self.ctx.handle, 1 if mode == "cv2" else 0), "lp_set_roi_mode")

    def set_fused(self, enable: bool):
        L.check(L.lib().lp_set_fused_classifier(self.ctx.handle, 1 if enable else 0))
        self.fused = bool(enable) and self._has_fused

    def _alloc(self, n: int):
        if n <= self._cap:
            return
        S, Cn = self.input_size, self.num_classes
        with torch.cuda.device(self.device):
            self.cls_in = torch.empty((n, S, S, 3), dtype=torch.uint8, device=self.device)
            self.logits = torch.empty((n, Cn), dtype=torch.float32, device=self.device)
            self.probs = torch.empty((n, Cn), dtype=torch.float32, device=self.device)
            self.argmax = torch.empty((n,), dtype=torch.int64, device=self.device)
        self._cap = n

    # ------------------------------------------------------------------ device stages
    def resize_device(self, fb: FrameBatch, roi_xyxy: torch.Tensor, roi_src: torch.Tensor, n_rois: int) -> torch.Tensor:
        """K6: crop + Pillow-exact resize of ``n_rois`` ROIs into self.cls_in[:n_rois] (RGB u8)."""
        self._alloc(n_rois)
        L.check(L.lib().lp_roi_resize(self.ctx.handle, fb.ptrs, fb.pitch, fb.n, _ptr(roi_xyxy), _ptr(roi_src),
                                      int(n_rois), self.input_size, fb.max_side, _ptr(self.cls_in), _stream()),
                "lp_roi_resize")
        return self.cls_in[:n_rois]

    def classify_device(self, cls_in: torch.Tensor) -> None:
        """K7 on [n,S,S,3] RGB u8 -> self.logits/probs/argmax[:n]."""
        n = int(cls_in.shape[0])
        self._alloc(n)
        L.check(L.lib().lp_classify(self.ctx.handle, _ptr(cls_in), n, _ptr(self.workspace), self.workspace.numel(),
                                    _ptr(self.logits), _ptr(self.probs), _ptr(self.argmax), _stream()), "lp_classify")

    # ------------------------------------------------------------------ reference-shaped API
    def preprocess_batch(self, images: Sequence[np.ndarray]) -> np.ndarray:
        """ROI list (HWC BGR u8, ragged) -> [n,S,S,3] RGB u8 exactly as cvtColor + PIL Resize((64,64))."""
        n = len(images)
        fb = FrameBatch.from_host(images, self.device)
        out = []
        for base in range(0, n, 64):                # each ROI image is its own "frame"
            m = min(64, n - base)
            sub = FrameBatch(fb.tensors[base:base + m])
            xy = torch.tensor([[0, 0, t.shape[1], t.shape[0]] for t in sub.tensors], dtype=torch.int32, device=self.device)
            src = torch.tensor([[i, 0] for i in range(m)], dtype=torch.int32, device=self.device)
            out.append(self.resize_device(sub, xy, src, m).clone())
        return torch.cat(out, 0)

    @torch.no_grad()
    def predict_batch(self, images: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        """e2e.py:378-396."""
        if len(images) == 0:
            return np.array([]), np.array([])
        x = self.preprocess_batch(images)
        self.classify_device(x)
        n = len(images)
        probs = self.probs[:n].cpu().numpy().copy()
        return np.argmax(probs, axis=1), probs

    def logits_for(self, cls_in) -> np.ndarray:
        """Raw logits for [n,S,S,3] RGB u8 input (numpy or CUDA tensor); parity tests."""
        if isinstance(cls_in, np.ndarray):
            cls_in = torch.from_numpy(np.ascontiguousarray(cls_in)).to(self.device)
        self.classify_device(cls_in.contiguous())
        return self.logits[:cls_in.shape[0]].cpu().numpy().copy()
