"""Pins the CPU oracle (oracle/) to the reference: against the installed third-party libraries the
reference calls (cv2, Pillow), against the unmodified e2e.py when /root/reference is mounted, and
against the golden vectors recorded from it (tests/golden/make_golden.py).  CPU only."""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import pytest
import torch
from PIL import Image

from helpers import GOLDEN, REF, debug_roi_paths, onnx_path
from oracle import pipeline_ref as PR
from oracle.ncnn_graph import DetectorOracle
import litepi_b200
from litepi_b200 import synth

SHAPES = [(681, 1198), (2048, 2048), (720, 1280), (480, 640), (640, 640), (333, 517), (1280, 1280),
          (100, 37), (641, 639), (1000, 250), (64, 64), (37, 1400)]


@pytest.fixture(scope="module")
def ref_e2e():
    path = f"{REF}/src/vntsr/pipeline/e2e.py"
    if not os.path.exists(path):
        pytest.skip("/root/reference not mounted (GPU box): covered by the golden vectors instead")
    for n in ("ncnn", "matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["ncnn"].Mat = np.ndarray
    sys.modules["ncnn"].Net = object
    spec = importlib.util.spec_from_file_location("ref_e2e", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("shape", SHAPES)
def test_letterbox_restatement_equals_cv2(shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    a, r, pad = PR.letterbox_ref(img)
    b, r2, pad2 = PR.letterbox_lib(img)
    assert a.shape == b.shape == (640, 640, 3)
    assert np.array_equal(a, b)
    assert r == r2 and tuple(pad) == tuple(pad2)


def test_letterbox_equals_reference(ref_e2e):
    for shape in SHAPES[:6]:
        img = np.random.default_rng(1).integers(0, 256, shape + (3,), dtype=np.uint8)
        a, r, pad = ref_e2e.letterbox(img, (640, 640))
        b, r2, pad2 = PR.letterbox_ref(img)
        assert np.array_equal(a, b) and r == r2 and tuple(pad) == tuple(pad2)


@pytest.mark.parametrize("shape", [(10, 10), (72, 84), (45, 31), (64, 64), (64, 100), (130, 64), (200, 333),
                                   (17, 90), (500, 480), (3, 5), (1, 1), (2, 300), (1, 77)])
def test_pil_restatement_equals_pillow(shape):
    img = np.random.default_rng(shape[0] + 1000 * shape[1]).integers(0, 256, shape + (3,), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((64, 64), Image.BILINEAR))
    assert np.array_equal(PR.pil_resize_bilinear_u8(img, 64), want)
    u8, x = PR.classifier_input_ref(img)
    u8b, xb = PR.classifier_input_lib(img)
    assert np.array_equal(u8, u8b) and np.array_equal(x, xb)


def test_postprocess_equals_reference(ref_e2e):
    det = ref_e2e.NCNNDetector.__new__(ref_e2e.NCNNDetector)
    g = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    for name, shape in (("vn0", synth.VN_SHAPE), ("tt0", synth.TT_SHAPE)):
        out0 = g[f"{name}.out0"]
        r, pw, ph = (float(v) for v in g[f"{name}.ratio_pad"])   # python floats, as letterbox() returns them
        for conf in (0.25, 0.001):
            wb, ws, wc = det.postprocess(out0, shape, r, (pw, ph), conf, 0.45)
            gb, gs, gc = PR.postprocess_ref(out0, shape, r, (pw, ph), conf, 0.45)
            assert np.array_equal(wb, gb) and np.array_equal(ws, gs) and np.array_equal(wc, gc)
    # empty result: float64 empties like the reference
    e = PR.postprocess_ref(g["vn0.out0"], synth.VN_SHAPE, 0.5, (0.0, 0.0), 0.9999, 0.45)
    r = det.postprocess(g["vn0.out0"], synth.VN_SHAPE, 0.5, (0.0, 0.0), 0.9999, 0.45)
    for a, b in zip(e, r):
        assert a.shape == b.shape and a.dtype == b.dtype == np.float64


def test_nms_equals_reference_random(ref_e2e):
    rng = np.random.default_rng(5)
    for n in (1, 2, 17, 300):
        xy = rng.uniform(0, 600, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 120, (n, 2)).astype(np.float32)
        boxes = np.concatenate([xy, xy + wh], 1)
        scores = rng.permutation(n).astype(np.float32) / n          # distinct -> no tie ambiguity
        assert list(ref_e2e.nms_numpy(boxes, scores, 0.45)) == PR.nms_ref(boxes, scores, 0.45)
    assert PR.nms_ref(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32)) == []


def test_golden_detector_path():
    """Oracle vs vectors recorded from the reference code (runs anywhere)."""
    g = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    frames = {"vn0": synth.vn_frame(0), "vn1": synth.vn_frame(1), "tt0": synth.tt_frame(0)}
    for name, f in frames.items():
        lb, r, pad = PR.letterbox_ref(f)
        assert np.array_equal(lb[::64], g[f"{name}.lb_rows"])
        assert int(lb.astype(np.int64).sum()) == int(g[f"{name}.lb_sum"][0])
        assert np.array_equal(np.array([r, pad[0], pad[1]]), g[f"{name}.ratio_pad"])
        for conf in (0.25, 0.001):
            tag = f"{name}.c{conf}"
            b, s, c = PR.postprocess_ref(g[f"{name}.out0"], f.shape[:2], r, pad, conf, 0.45)
            assert np.array_equal(b, g[tag + ".boxes"]) and np.array_equal(s, g[tag + ".scores"])
            assert np.array_equal(c, g[tag + ".classes"])
            rois, valid = PR.roi_select_ref(b, f.shape[:2], 50)
            assert np.array_equal(rois, g[tag + ".rois"]) and np.array_equal(np.array(valid, np.int64), g[tag + ".valid"])


def test_golden_detector_graph(v1_paths):
    """torch-fp32 graph oracle vs OpenCV-DNN running the reference's ONNX (recorded)."""
    param, binp = v1_paths
    if binp is None:
        pytest.skip("trained v1 weights not staged")
    g = np.load(os.path.join(GOLDEN, "detector_rand_seed0.npz"))
    torch.manual_seed(0)
    x = torch.rand(1, 3, 640, 640)                       # model_ncnn.py:6-7
    out = DetectorOracle(param, binp).forward(x)[0].numpy()
    d = np.abs(out - g["out0"])
    assert d[:4].max() < 2e-3 and d[4].max() < 1e-5      # two fp32 runtimes: re-association noise only
    p = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    xin, _, _, _ = PR.preprocess_ref(synth.vn_frame(0))
    d = np.abs(DetectorOracle(param, binp).forward(xin)[0].numpy() - p["vn0.out0"])
    assert d[:4].max() < 2e-3 and d[4].max() < 1e-5


def test_golden_detector_graph_live_onnx(v1_paths):
    onnx = onnx_path()
    param, binp = v1_paths
    if onnx is None or binp is None:
        pytest.skip("reference ONNX not available")
    net = cv2.dnn.readNetFromONNX(onnx)
    xin, _, _, _ = PR.preprocess_ref(synth.vn_frame(2))
    net.setInput(xin)
    ref = net.forward()[0]
    d = np.abs(DetectorOracle(param, binp).forward(xin)[0].numpy() - ref)
    assert d[:4].max() < 2e-3 and d[4].max() < 1e-5


def test_golden_classifier_input():
    paths = debug_roi_paths()
    if not paths:
        pytest.skip("reference debug_rois not staged")
    g = np.load(os.path.join(GOLDEN, "classifier_input.npz"))
    for p in paths:
        fn = os.path.basename(p)
        img = cv2.imread(p)
        assert tuple(img.shape) == tuple(g[fn + ".shape"])
        u8, x = PR.classifier_input_ref(img)
        assert np.array_equal(u8, g[fn + ".u8"])
        assert np.array_equal(x[:, 31, :], g[fn + ".tensor_row"])


def test_shufflenet_oracle_is_torchvision():
    m = PR.build_shufflenet(49, seed=0)
    assert m.fc.out_features == 49 and sum(p.numel() for p in m.parameters()) == 1303829
    crops = synth.roi_crops(5, seed=3)
    a = PR.classify_ref(m, crops)
    b = PR.classify_lib(m, crops)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert PR.classify_ref(m, [])[0].shape == (0,)


# ------------------------------------------------------------------ evaluation
@pytest.mark.parametrize("case", ["small", "mid", "wide", "empty"])
def test_eval_restatement_equals_reference_golden(case):
    """oracle/eval_ref.py against outputs of the UNMODIFIED evaluate_predictions (e2e.py:656-824)."""
    from helpers import load_eval_case
    from oracle import eval_ref as ER
    preds, gts, nc, want = load_eval_case(case)
    got = ER.evaluate_predictions_ref(preds, gts, nc)
    assert set(got) == set(want)
    for k, v in want.items():
        assert np.array_equal(np.asarray(got[k]), v), k


def test_eval_restatement_equals_reference_live(ref_e2e):
    """fresh seeds through the reference itself (only where /root/reference is mounted)"""
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "golden"))
    from make_golden import eval_case
    from oracle import eval_ref as ER
    for seed, n_img, nc in ((11, 30, 4), (12, 150, 58), (13, 5, 2)):
        preds, gts = eval_case(seed, n_img, nc)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = ref_e2e.evaluate_predictions(preds, gts, nc)
        got = ER.evaluate_predictions_ref(preds, gts, nc)
        for k, v in want.items():
            assert np.array_equal(np.asarray(got[k]), np.asarray(v)), (seed, k)


# ------------------------------------------------------------------ e2e_optimize.py variant
def test_optimized_variant_restatements():
    """classifier_input_opt_ref against cv2 itself (what e2e_optimize.py:391-393 calls) on ROI-like shapes, and
    roi_select_opt_ref against the reference's own numpy lines (:480-496) executed verbatim on the same boxes."""
    rng = np.random.default_rng(3)
    shapes = [(128, 128), (64, 64), (32, 32), (127, 129), (2, 2), (3, 200), (200, 3), (10, 90), (255, 31)] + \
             [(int(rng.integers(2, 260)), int(rng.integers(2, 260))) for _ in range(60)]
    for h, w in shapes:
        roi = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = cv2.resize(cv2.cvtColor(roi, cv2.COLOR_BGR2RGB), (64, 64), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(PR.classifier_input_opt_ref(roi), want), (h, w)
    for trial in range(50):
        h, w = int(rng.integers(50, 2100)), int(rng.integers(50, 2100))
        n = int(rng.integers(0, 40))
        xy = rng.uniform(-30, max(h, w) + 30, (n, 2)); wh = rng.uniform(0, 200, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        min_area = int(rng.choice([1, 50, 100]))
        rois, valid = PR.roi_select_opt_ref(boxes, (h, w), min_area)
        if n:
            boxes_int = boxes.astype(np.int32)                                  # e2e_optimize.py:480-487
            boxes_int[:, [0, 2]] = np.clip(boxes_int[:, [0, 2]], 0, w)
            boxes_int[:, [1, 3]] = np.clip(boxes_int[:, [1, 3]], 0, h)
            areas = (boxes_int[:, 2] - boxes_int[:, 0]) * (boxes_int[:, 3] - boxes_int[:, 1])
            m = areas >= min_area
            want = [tuple(b) for b in boxes_int[m] if b[2] > b[0] and b[3] > b[1]]
            assert [tuple(r) for r in rois] == want
        else:
            assert len(rois) == 0 and valid == []
