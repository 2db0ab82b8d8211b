"""Synthetic traffic-scene frames (datasets are unavailable offline).

SURVEY.md section 8(d): blurred uniform-noise background with pasted sign-like
glyphs (red ring + white disc + digits, blue disc + arrow, red/yellow triangle)
so that the reference's trained v1 detector produces real detections.  Frames
are HWC BGR uint8 exactly like ``cv2.imread`` output (reference
``src/vntsr/pipeline/e2e.py:962``).  Seeded per frame: ``default_rng(seed)``.
"""
from __future__ import annotations

import cv2
import numpy as np

VN_SHAPE = (681, 1198)      # H, W of VN-Signs frames (BASELINE.json configs[0], [1])
TT_SHAPE = (2048, 2048)     # TT100K frames (configs[2])


def _glyph(img, cx, cy, r, kind, rng):
    if kind == 0:      # prohibition / speed limit: red ring, white disc, digits
        cv2.circle(img, (cx, cy), r, (20, 20, 200), -1, cv2.LINE_AA)
        cv2.circle(img, (cx, cy), int(r * 0.72), (245, 245, 245), -1, cv2.LINE_AA)
        txt = str(int(rng.integers(2, 13)) * 10)
        fs = r / 28.0
        (tw, th), _ = cv2.getTextSize(txt, cv2.FONT_HERSHEY_SIMPLEX, fs, max(1, r // 12))
        cv2.putText(img, txt, (cx - tw // 2, cy + th // 2), cv2.FONT_HERSHEY_SIMPLEX, fs,
                    (10, 10, 10), max(1, r // 12), cv2.LINE_AA)
    elif kind == 1:    # mandatory: blue disc, white arrow
        cv2.circle(img, (cx, cy), r, (200, 90, 20), -1, cv2.LINE_AA)
        cv2.arrowedLine(img, (cx - r // 2, cy + r // 3), (cx + r // 2, cy - r // 3), (250, 250, 250),
                        max(1, r // 6), cv2.LINE_AA, tipLength=0.45)
    else:              # warning: red triangle, yellow fill
        pts = np.array([[cx, cy - r], [cx - r, cy + int(0.8 * r)], [cx + r, cy + int(0.8 * r)]], np.int32)
        cv2.fillPoly(img, [pts], (20, 20, 210), cv2.LINE_AA)
        inner = ((pts - [cx, cy]) * 0.68 + [cx, cy + r * 0.06]).astype(np.int32)
        cv2.fillPoly(img, [inner], (40, 220, 245), cv2.LINE_AA)
        cv2.line(img, (cx, cy - r // 4), (cx, cy + r // 3), (10, 10, 10), max(1, r // 8), cv2.LINE_AA)


def synth_frame(h: int, w: int, n_signs: int, seed: int, rmin: int = 18, rmax: int = 45) -> np.ndarray:
    """One HWC BGR uint8 frame with ``n_signs`` glyphs of radius U[rmin, rmax]."""
    rng = np.random.default_rng(seed)
    # low-resolution noise upsampled + blurred: cheap at 2048^2 and statistically
    # the same as blurring full-resolution noise with sigma 8
    small = rng.integers(60, 200, size=((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
    img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
    img = cv2.GaussianBlur(img, (0, 0), 3)
    for _ in range(n_signs):
        r = int(rng.integers(rmin, rmax + 1))
        cx = int(rng.integers(r + 2, w - r - 2))
        cy = int(rng.integers(r + 2, h - r - 2))
        _glyph(img, cx, cy, r, int(rng.integers(0, 3)), rng)
    noise = rng.integers(-6, 7, size=img.shape, dtype=np.int16)
    return np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def vn_frame(seed: int) -> np.ndarray:
    """VN-Signs-shape frame, 2..10 signs (BASELINE.json configs[1])."""
    n = int(np.random.default_rng(seed + 100003).integers(2, 11))
    return synth_frame(VN_SHAPE[0], VN_SHAPE[1], n, seed)


def tt_frame(seed: int) -> np.ndarray:
    """TT100K-shape frame, 20..40 small signs (BASELINE.json configs[2])."""
    n = int(np.random.default_rng(seed + 100003).integers(20, 41))
    return synth_frame(TT_SHAPE[0], TT_SHAPE[1], n, seed, rmin=24, rmax=60)


def roi_crops(n: int, seed: int = 0, smin: int = 10, smax: int = 90):
    """``n`` synthetic sign crops (HWC BGR u8, ragged sizes U[smin,smax]^2) for
    the classifier-alone workload (BASELINE.json configs[3])."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        hh, ww = int(rng.integers(smin, smax + 1)), int(rng.integers(smin, smax + 1))
        img = np.full((hh, ww, 3), rng.integers(60, 200, 3, dtype=np.uint8), np.uint8)
        _glyph(img, ww // 2, hh // 2, max(3, min(hh, ww) // 2 - 1), int(rng.integers(0, 3)), rng)
        noise = rng.integers(-8, 9, size=img.shape, dtype=np.int16)
        out.append(np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8))
    return out
