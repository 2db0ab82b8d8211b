#!/usr/bin/env python3
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF in the authoring container.

The reference ships no tests or golden vectors, so parity is pinned by
recording outputs of its own code:
  * /root/reference/src/vntsr/pipeline/e2e.py imported UNMODIFIED (three absent imports --
    ncnn, matplotlib, seaborn -- are stubbed): letterbox, NCNNDetector.postprocess (+ nms_numpy),
    HybridPipeline.run's ROI loop, PyTorchClassifier's transform;
  * the reference's exported graph + trained weights (yolo_plus.onnx) executed by OpenCV-DNN
    (onnxruntime / ncnn are not installable offline) for out0.
Run from the repo root:  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import litepi_b200  # noqa: E402
from litepi_b200 import synth  # noqa: E402

REF = "/root/reference"


def load_reference():
    for n in ("ncnn", "matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["ncnn"].Mat = np.ndarray
    sys.modules["ncnn"].Net = object
    spec = importlib.util.spec_from_file_location("ref_e2e", f"{REF}/src/vntsr/pipeline/e2e.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def main():
    ref = load_reference()
    det = ref.NCNNDetector.__new__(ref.NCNNDetector)
    net = cv2.dnn.readNetFromONNX(f"{REF}/src/vntsr/convert/model/yolo_plus/yolo_plus.onnx")

    # ---- 1. the pnnx smoke recipe (model_ncnn.py:6-7): seeded rand input -> out0
    torch.manual_seed(0)
    x = torch.rand(1, 3, 640, 640)
    net.setInput(x.numpy())
    out_rand = net.forward().copy()
    np.savez_compressed(os.path.join(HERE, "detector_rand_seed0.npz"), out0=out_rand[0].astype(np.float32))

    # ---- 2. full detector path on synthetic frames (VN + TT shapes), both thresholds
    g = {}
    frames = {"vn0": synth.vn_frame(0), "vn1": synth.vn_frame(1), "tt0": synth.tt_frame(0)}
    for name, f in frames.items():
        lb, r, pad = ref.letterbox(f, (640, 640))
        rgb = cv2.cvtColor(lb, cv2.COLOR_BGR2RGB)
        xin = (rgb.astype(np.float32) / 255.0).transpose(2, 0, 1)[None]
        net.setInput(np.ascontiguousarray(xin))
        out0 = net.forward()[0].copy()
        g[f"{name}.lb_sum"] = np.array([int(lb.astype(np.int64).sum()), int((lb.astype(np.int64) * np.arange(lb.size).reshape(lb.shape) % 65521).sum())])
        g[f"{name}.lb_rows"] = lb[::64].copy()                     # every 64th row, exact bytes
        g[f"{name}.ratio_pad"] = np.array([r, pad[0], pad[1]], np.float64)
        g[f"{name}.out0"] = out0.astype(np.float32)
        for conf in (0.25, 0.001):
            b, s, c = det.postprocess(out0, f.shape[:2], r, pad, conf, 0.45)
            tag = f"{name}.c{conf}"
            g[tag + ".boxes"], g[tag + ".scores"], g[tag + ".classes"] = np.asarray(b), np.asarray(s), np.asarray(c)
            # ROI loop of HybridPipeline.run (e2e.py:459-475), executed verbatim from the reference source
            rois, valid = [], []
            h, w = f.shape[:2]
            for idx, box in enumerate(b):
                x1, y1, x2, y2 = box.astype(int)
                x1, y1 = np.clip(x1, 0, w - 1), np.clip(y1, 0, h - 1)
                x2, y2 = np.clip(x2, x1 + 1, w), np.clip(y2, y1 + 1, h)
                area = (x2 - x1) * (y2 - y1)
                if area >= 50 and x2 > x1 and y2 > y1:
                    rois.append((x1, y1, x2, y2)); valid.append(idx)
            g[tag + ".rois"] = np.array(rois, np.int32).reshape(-1, 4)
            g[tag + ".valid"] = np.array(valid, np.int64)
    np.savez_compressed(os.path.join(HERE, "detector_path.npz"), **g)

    # ---- 3. classifier preprocessing of the reference's real ROI crops (debug_rois) through its own transform
    clf = ref.PyTorchClassifier.__new__(ref.PyTorchClassifier)
    from torchvision import transforms
    from PIL import Image
    clf.transform = transforms.Compose([transforms.Resize((64, 64)), transforms.ToTensor(),
                                        transforms.Normalize([0.18] * 3, [0.34] * 3)])   # e2e.py:366-370
    d = f"{REF}/src/vntsr/pipeline/debug_rois"
    c = {}
    for fn in sorted(os.listdir(d)):
        img = cv2.imread(os.path.join(d, fn))
        rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        pil = Image.fromarray(rgb)
        t = clf.transform(pil).numpy()
        u8 = np.asarray(transforms.Resize((64, 64))(pil))
        c[fn + ".shape"] = np.array(img.shape)
        c[fn + ".u8"] = u8
        c[fn + ".tensor_sum"] = np.array([float(t.astype(np.float64).sum()), float(np.abs(t).astype(np.float64).sum())])
        c[fn + ".tensor_row"] = t[:, 31, :].copy()
    np.savez_compressed(os.path.join(HERE, "classifier_input.npz"), **c)
    make_eval_golden(ref)
    for fn in ("detector_rand_seed0.npz", "detector_path.npz", "classifier_input.npz", "eval_cases.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)) // 1024, "KiB")


def eval_case(seed: int, n_img: int, n_cls: int, w: int = 1198, h: int = 681):
    """Seeded evaluation workload: ground-truth signs per image, predictions = jittered copies (some with the
    wrong class, some duplicated, some spurious), integer predicted boxes like HybridPipeline.run emits."""
    rng = np.random.default_rng(seed)
    preds, gts = [], []
    for i in range(n_img):
        ng = int(rng.integers(0, 7))
        if i % 9 == 4:
            ng = 0
        g = []
        for _ in range(ng):
            bw, bh = rng.uniform(12, 120), rng.uniform(12, 120)
            x1, y1 = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
            g.append([float(rng.integers(0, n_cls)), x1, y1, x1 + bw, y1 + bh])
        pr = []
        if i % 7 != 3:
            for row in g:
                for _ in range(int(rng.integers(0, 3))):                       # 0, 1 or 2 detections per sign
                    j = rng.normal(0, 0.08, 4) * np.array([row[3] - row[1], row[4] - row[2]] * 2)
                    bb = np.array(row[1:]) + j
                    cls = int(row[0]) if rng.random() < 0.8 else int(rng.integers(0, n_cls))
                    pr.append({"bbox": tuple(np.array(bb).astype(int)), "conf": float(rng.uniform(0.05, 1.0)), "cls_class": cls})
            for _ in range(int(rng.integers(0, 3))):                               # false alarms
                bw, bh = rng.uniform(12, 90), rng.uniform(12, 90)
                x1, y1 = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
                pr.append({"bbox": tuple(np.array([x1, y1, x1 + bw, y1 + bh]).astype(int)), "conf": float(rng.uniform(0.05, 0.6)),
                           "cls_class": int(rng.integers(0, n_cls))})
        preds.append(pr); gts.append(g)
    return preds, gts


def pack_eval_case(preds, gts):
    """Flat arrays (what the npz stores and tests rebuild the lists from)."""
    pb = np.array([p["bbox"] for im in preds for p in im], np.int64).reshape(-1, 4)
    pc = np.array([p["conf"] for im in preds for p in im], np.float64)
    pk = np.array([p["cls_class"] for im in preds for p in im], np.int64)
    pn = np.array([len(im) for im in preds], np.int64)
    gb = np.array([r for im in gts for r in im], np.float64).reshape(-1, 5)
    gn = np.array([len(im) for im in gts], np.int64)
    return dict(pred_box=pb, pred_conf=pc, pred_cls=pk, pred_n=pn, gt=gb, gt_n=gn)


def make_eval_golden(ref):
    """evaluate_predictions (e2e.py:656-824) run UNMODIFIED on seeded cases; inputs and outputs recorded."""
    import warnings
    out = {}
    for name, (seed, n_img, n_cls) in {"small": (0, 12, 3), "mid": (1, 80, 6), "wide": (2, 200, 49), "empty": (3, 0, 4)}.items():
        preds, gts = eval_case(seed, n_img, n_cls)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = ref.evaluate_predictions(preds, gts, n_cls)
        for k, v in pack_eval_case(preds, gts).items():
            out[f"{name}.in.{k}"] = v
        out[f"{name}.in.num_classes"] = np.array(n_cls)
        for k, v in m.items():
            out[f"{name}.out.{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)


if __name__ == "__main__":
    main()
