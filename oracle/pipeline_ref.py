"""Oracle: numpy restatement of the reference's pre/post-processing hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

Each function follows the cited lines of ``/root/reference/src/vntsr/pipeline/e2e.py``.
Where the reference calls a third-party library (OpenCV 4.9.0.80 ``cv2.resize``,
Pillow ``Image.resize`` via ``torchvision.transforms.Resize``; ``requirements.txt:5,23``)
the library's published fixed-point algorithm is restated in numpy so that the
oracle is self-contained; ``tests/test_oracle_pinning.py`` pins the restatement to
the installed cv2 / Pillow on many shapes and to the unmodified ``e2e.py`` run in
the authoring container (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


# --------------------------------------------------------------------------- letterbox
def cv_resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """cv2.resize(img, (new_w,new_h), INTER_LINEAR) on uint8 HWC, restated
    (OpenCV imgproc/resize.cpp: resizeGeneric_ with HResizeLinear / VResizeLinear,
    INTER_RESIZE_COEF_BITS = 11).  Called at e2e.py:80."""
    h, w = img.shape[:2]

    def coefs(n_out, n_in, clamp):
        scale = 1.0 / (np.float64(n_out) / np.float64(n_in))
        d = np.arange(n_out, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        a = (f - s.astype(np.float32)).astype(np.float32)
        if clamp:
            lo = s < 0
            s[lo] = 0
            a[lo] = 0
            hi = s >= n_in - 1
            s[hi] = n_in - 1
            a[hi] = 0
        c1 = np.rint(a * np.float32(2048)).astype(np.int64)
        c0 = np.rint((np.float32(1) - a) * np.float32(2048)).astype(np.int64)
        return s, c0, c1

    sx, a0, a1 = coefs(new_w, w, True)
    sy, b0, b1 = coefs(new_h, h, False)
    sx1 = np.minimum(sx + 1, w - 1)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int64)
    # horizontal pass on the rows that are used
    t = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # [h, new_w, c]
    t0, t1 = t[y0], t[y1]
    out = (((b0[:, None, None] * (t0 >> 4)) >> 16) + ((b1[:, None, None] * (t1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_ref(img: np.ndarray, new_shape=(640, 640), color=114):
    """e2e.py:66-86 -- returns (BGR u8 letterboxed image, ratio, (dw, dh))."""
    shape = img.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        img = cv_resize_linear_u8(img, new_unpad[0], new_unpad[1])
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    out = np.full((img.shape[0] + top + bottom, img.shape[1] + left + right, 3), color, np.uint8)
    out[top:top + img.shape[0], left:left + img.shape[1]] = img
    return out, r, (dw, dh)


def preprocess_ref(img: np.ndarray, size: int = 640):
    """e2e.py:222-238 (+ ORT twin evaluation_tsd_single_img.ipynb:98-110):
    letterbox, BGR->RGB, float32 / 255, HWC->CHW, batch dim."""
    lb, r, pad = letterbox_ref(img, (size, size))
    x = lb[:, :, ::-1].astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None], r, pad, lb


# --------------------------------------------------------------------------- postprocess + NMS
def nms_ref(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float = 0.45) -> List[int]:
    """e2e.py:89-119 with the tie order DEFINED as (score desc, index desc): the
    reference's ``scores.argsort()[::-1]`` uses numpy's unstable default sort, so
    its order among exactly equal scores is implementation-defined."""
    if len(boxes) == 0:
        return []
    x1, y1, x2, y2 = boxes.T
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(scores, kind="stable")[::-1]
    keep = []
    thr = np.float32(iou_threshold)
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        if len(order) == 1:
            break
        rest = order[1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0.0), xx2 - xx1)
        h = np.maximum(np.float32(0.0), yy2 - yy1)
        inter = w * h
        iou = inter / (areas[i] + areas[rest] - inter + np.float32(1e-6))
        order = rest[np.where(iou <= thr)[0]]
    return keep


def postprocess_ref(out0: np.ndarray, orig_shape: Tuple[int, int], ratio: float, pad: Tuple[float, float],
                    conf_threshold: float = 0.5, iou_threshold: float = 0.45, return_candidates: bool = False):
    """e2e.py:240-296 on ``out0`` [4+nc, A] float32.  float32 arithmetic with numpy's
    weak-scalar rules made explicit."""
    pred = np.asarray(out0, dtype=np.float32)
    boxes = pred[:4].T
    scores = pred[4:].T
    class_scores = np.max(scores, axis=1)
    class_ids = np.argmax(scores, axis=1)
    mask = class_scores > np.float32(conf_threshold)
    boxes, scores1, class_ids = boxes[mask], class_scores[mask], class_ids[mask]
    if len(boxes) == 0:
        e = (np.empty((0, 4)), np.empty((0,)), np.empty((0,)))
        return e + ((np.empty((0, 4), np.float32), np.empty((0,), np.float32), np.empty((0,), np.int64),
                     np.empty((0,), np.int64)),) if return_candidates else e
    xc, yc, w, h = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    two = np.float32(2)
    xyxy = np.stack([xc - w / two, yc - h / two, xc + w / two, yc + h / two], axis=1).astype(np.float32)
    xyxy[:, [0, 2]] -= np.float32(pad[0])
    xyxy[:, [1, 3]] -= np.float32(pad[1])
    xyxy /= np.float32(ratio)
    xyxy[:, [0, 2]] = np.clip(xyxy[:, [0, 2]], np.float32(0), np.float32(orig_shape[1]))
    xyxy[:, [1, 3]] = np.clip(xyxy[:, [1, 3]], np.float32(0), np.float32(orig_shape[0]))
    idx = []
    for cls in np.unique(class_ids):
        m = class_ids == cls
        keep = nms_ref(xyxy[m], scores1[m], iou_threshold)
        idx.extend(np.where(m)[0][keep])
    idx = np.array(idx, dtype=np.int64)
    res = (xyxy[idx], scores1[idx], class_ids[idx].astype(np.int64))
    if return_candidates:
        return res + ((xyxy, scores1, class_ids.astype(np.int64), idx),)
    return res


# --------------------------------------------------------------------------- ROI extraction
def roi_select_ref(boxes: np.ndarray, img_shape: Tuple[int, int], min_area: int = 100):
    """e2e.py:459-475 -- returns (int ROIs [K',4] as (x1,y1,x2,y2), valid indices)."""
    h, w = img_shape
    rois, valid = [], []
    for idx, box in enumerate(boxes):
        x1, y1, x2, y2 = box.astype(int)
        x1, y1 = np.clip(x1, 0, w - 1), np.clip(y1, 0, h - 1)
        x2, y2 = np.clip(x2, x1 + 1, w), np.clip(y2, y1 + 1, h)
        area = (x2 - x1) * (y2 - y1)
        if area >= min_area and x2 > x1 and y2 > y1:
            rois.append((int(x1), int(y1), int(x2), int(y2)))
            valid.append(idx)
    return np.array(rois, dtype=np.int32).reshape(-1, 4), valid


# --------------------------------------------------------------------------- Pillow resize
_PB = 22


def _pil_coeffs(in_size: int, out_size: int):
    """Pillow src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc, BILINEAR."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / fs
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        ws = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            ws.append(w)
            ww += w
        for x in range(xmax):
            w = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx, x] = int(0.5 + w * (1 << _PB))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear_u8(img: np.ndarray, out_size: int = 64) -> np.ndarray:
    """``PIL.Image.fromarray(img).resize((S,S), BILINEAR)`` restated for uint8 HWC
    (what ``transforms.Resize((64,64))`` does, e2e.py:367,386-388).  Horizontal pass
    first, uint8 intermediate, then vertical."""
    h, w = img.shape[:2]
    bx, kx = _pil_coeffs(w, out_size)
    by, ky = _pil_coeffs(h, out_size)
    src = img.astype(np.int64)
    tmp = np.empty((h, out_size, img.shape[2]), np.uint8)
    for xx in range(out_size):
        x0, n = bx[xx]
        acc = (1 << (_PB - 1)) + np.tensordot(src[:, x0:x0 + n], kx[xx, :n], axes=([1], [0]))
        tmp[:, xx] = np.clip(acc >> _PB, 0, 255)
    out = np.empty((out_size, out_size, img.shape[2]), np.uint8)
    t64 = tmp.astype(np.int64)
    for yy in range(out_size):
        y0, n = by[yy]
        acc = (1 << (_PB - 1)) + np.tensordot(ky[yy, :n], t64[y0:y0 + n], axes=([0], [0]))
        out[yy] = np.clip(acc >> _PB, 0, 255)
    return out


def classifier_input_ref(roi_bgr: np.ndarray, size: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    """e2e.py:385-388 with transform :366-370 -> (u8 RGB SxSx3, float32 CHW normalised)."""
    rgb = roi_bgr[:, :, ::-1]
    u8 = pil_resize_bilinear_u8(rgb, size)
    x = u8.astype(np.float32) / np.float32(255)
    x = (x - np.float32(0.18)) / np.float32(0.34)
    return u8, np.ascontiguousarray(x.transpose(2, 0, 1))


# --------------------------------------------------------------------------- classifier
def build_shufflenet(num_classes: int, seed: int = 0):
    """e2e.py:331-333 -- torchvision shufflenet_v2_x1_0(weights=None) with fc -> C,
    random-init under ``torch.manual_seed(seed)`` (the reference ships no classifier
    weights: '../weight/shufflenetv2.pth' at e2e.py:1019 is never committed)."""
    import torch
    import torch.nn as nn
    from torchvision import models
    torch.manual_seed(seed)
    m = models.shufflenet_v2_x1_0(weights=None)
    m.fc = nn.Linear(m.fc.in_features, num_classes)
    # un-trained BatchNorm has trivial statistics; give it seeded non-trivial ones so that BN
    # folding in the product is actually exercised
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
            mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=g)
            mod.running_mean.data = 0.2 * torch.randn(mod.running_mean.shape, generator=g)
            mod.running_var.data = 0.5 + torch.rand(mod.running_var.shape, generator=g)
    m.eval()
    return m


def build_classifier_ref(arch: str, num_classes: int, seed: int = 0):
    """e2e.py:320-335 ``build_classifier`` for every ``--clf_arch`` the reference accepts (resnet18, efficientnet = b0,
    mobilenetv2, shufflenetv2), random-init under ``torch.manual_seed(seed)`` with seeded non-trivial BatchNorm
    statistics (so that BN folding in the product is exercised)."""
    import torch
    import torch.nn as nn
    from torchvision import models
    torch.manual_seed(seed)
    if arch == "resnet18":
        m = models.resnet18(weights=None)
        m.fc = nn.Linear(m.fc.in_features, num_classes)
    elif arch == "efficientnet":
        m = models.efficientnet_b0(weights=None)
        m.classifier[1] = nn.Linear(m.classifier[1].in_features, num_classes)
    elif arch == "mobilenetv2":
        m = models.mobilenet_v2(weights=None)
        m.classifier[1] = nn.Linear(m.classifier[1].in_features, num_classes)
    elif arch == "shufflenetv2":
        return build_shufflenet(num_classes, seed)
    else:
        raise ValueError(f"Unknown architecture: {arch}")
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data = 0.8 + 0.4 * torch.rand(mod.weight.shape, generator=g)
            mod.bias.data = 0.1 * torch.randn(mod.bias.shape, generator=g)
            mod.running_mean.data = 0.1 * torch.randn(mod.running_mean.shape, generator=g)
            mod.running_var.data = 0.8 + 0.4 * torch.rand(mod.running_var.shape, generator=g)
    m.eval()
    return m


def classify_ref(model, rois_bgr: List[np.ndarray], size: int = 64):
    """e2e.py:378-396 -- returns (argmax int64 [B], probs f32 [B,C], logits f32 [B,C])."""
    import torch
    if len(rois_bgr) == 0:
        return np.array([]), np.array([]), np.array([])
    batch = torch.from_numpy(np.stack([classifier_input_ref(r, size)[1] for r in rois_bgr]))
    with torch.no_grad():
        logits = model(batch)
        probs = torch.softmax(logits, dim=1).numpy()
    return np.argmax(probs, axis=1), probs, logits.numpy()


# --------------------------------------------------------------------------- library-backed twins
# The same two steps through the third-party libraries the reference itself calls (cv2, Pillow).
# tests/test_oracle_pinning.py asserts they are bit-identical to the restatements above; bench.py's
# CPU arm uses them so the CPU baseline runs at the libraries' native speed, as the reference does.
def letterbox_lib(img: np.ndarray, new_shape=(640, 640), color=(114, 114, 114)):
    """e2e.py:66-86 verbatim in behaviour, via cv2.resize / cv2.copyMakeBorder."""
    import cv2
    shape = img.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    if shape[::-1] != new_unpad:
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)
    return img, r, (dw, dh)


def preprocess_lib(img: np.ndarray, size: int = 640):
    import cv2
    lb, r, pad = letterbox_lib(img, (size, size))
    x = cv2.cvtColor(lb, cv2.COLOR_BGR2RGB).astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None], r, pad, lb


def classifier_input_lib(roi_bgr: np.ndarray, size: int = 64):
    """e2e.py:385-388 via cv2.cvtColor + PIL resize (what transforms.Resize does on a PIL image)."""
    import cv2
    from PIL import Image
    rgb = cv2.cvtColor(np.ascontiguousarray(roi_bgr), cv2.COLOR_BGR2RGB)
    u8 = np.asarray(Image.fromarray(rgb).resize((size, size), Image.BILINEAR))
    x = (u8.astype(np.float32) / np.float32(255) - np.float32(0.18)) / np.float32(0.34)
    return u8, np.ascontiguousarray(x.transpose(2, 0, 1))


def classify_lib(model, rois_bgr: List[np.ndarray], size: int = 64):
    import torch
    if len(rois_bgr) == 0:
        return np.array([]), np.array([]), np.array([])
    batch = torch.from_numpy(np.stack([classifier_input_lib(r, size)[1] for r in rois_bgr]))
    with torch.no_grad():
        logits = model(batch)
        probs = torch.softmax(logits, dim=1).numpy()
    return np.argmax(probs, axis=1), probs, logits.numpy()


# ------------------------------------------------------------------ e2e_optimize.py variant
def roi_select_opt_ref(boxes: np.ndarray, shape, min_area: int):
    """src/tt100k/pipeline/e2e_optimize.py:476-499: int32 truncation, clip to [0,w]x[0,h], area filter, and the
    non-empty test of the ROI list comprehension.  Returns (rois [K,4] int32, indices of the kept detections)."""
    h, w = shape[:2]
    if len(boxes) == 0:
        return np.zeros((0, 4), np.int32), []
    b = np.asarray(boxes).astype(np.int32)
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, w)
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, h)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep = (area >= min_area) & (b[:, 2] > b[:, 0]) & (b[:, 3] > b[:, 1])
    return b[keep], [int(i) for i in np.nonzero(keep)[0]]


def classifier_input_opt_ref(roi_bgr: np.ndarray, size: int = 64) -> np.ndarray:
    """e2e_optimize.py:391-393: BGR->RGB then cv2.resize(INTER_LINEAR) (restated, no antialias) -> [size,size,3] u8."""
    rgb = roi_bgr[:, :, ::-1]
    if rgb.shape[0] == size and rgb.shape[1] == size:
        return np.ascontiguousarray(rgb)
    return cv_resize_linear_u8(np.ascontiguousarray(rgb), size, size)
