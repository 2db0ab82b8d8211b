"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded
inputs, the committed golden vectors, and size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star): NMS keep indices and ROI integers bit-exact given identical
pre-NMS input; detector out0 within 1e-2 px (boxes) / 1e-3 (scores); classifier top-1 equal and
logits within 1e-2.  Integer/byte stages (letterbox, ROI resize) are bit-exact."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, debug_roi_paths, oracle_pipeline_run
from oracle import pipeline_ref as PR
from oracle.ncnn_graph import DetectorOracle

pytestmark = pytest.mark.gpu

BOX_TOL, SCORE_TOL, LOGIT_TOL = 1e-2, 1e-3, 1e-2


@pytest.fixture(scope="module")
def lp():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import litepi_b200
    return litepi_b200


@pytest.fixture(scope="module")
def det_v1(lp, v1_paths):
    d = lp.B200Detector(v1_paths[0], v1_paths[1], max_batch=8, max_det=8400, seed=0)   # max_det = anchors: never truncates
    orc = DetectorOracle(v1_paths[0], v1_paths[1], seed=0)
    _sync(orc, d.model)
    return d, orc


@pytest.fixture(scope="module")
def clf(lp):
    ref = PR.build_shufflenet(49, seed=0)
    return lp.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=64), ref


def _sync(orc, model):
    ci = iter(model.convs)
    for ly in orc.layers:
        if ly.type == "Convolution":
            c = next(ci)
            ly.weight, ly.bias = c.weight, c.bias


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("shape", [(681, 1198), (2048, 2048), (720, 1280), (480, 640), (640, 640), (333, 517),
                                   (100, 37), (1280, 1280), (641, 639), (37, 1400), (1, 1)])
def test_letterbox_bit_exact(det_v1, shape):
    det, _ = det_v1
    img = np.random.default_rng(shape[0] * 31 + shape[1]).integers(0, 256, shape + (3,), dtype=np.uint8)
    got, r, pad = det.letterbox(img)
    want, r2, pad2 = PR.letterbox_ref(img)
    assert np.array_equal(got, want[:, :, ::-1])
    assert r == r2 and pad == tuple(pad2)


def test_letterbox_golden_and_ragged_batch(lp, det_v1):
    det, _ = det_v1
    from litepi_b200 import synth
    from litepi_b200.detector import FrameBatch
    g = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    frames = [synth.vn_frame(0), synth.tt_frame(0), synth.vn_frame(1)]          # ragged shapes in one launch
    lb = det.letterbox_device(FrameBatch.from_host(frames, det.device)).cpu().numpy()
    for i, name in enumerate(["vn0", "tt0", "vn1"]):
        assert np.array_equal(lb[i, ::64, :, ::-1], g[f"{name}.lb_rows"])
        assert det.ratio[i] == g[f"{name}.ratio_pad"][0]


# ------------------------------------------------------------------------------------------ K2/K3
def test_detector_forward_v1_vs_oracle_and_golden(det_v1):
    det, orc = det_v1
    from litepi_b200 import synth
    frames = [synth.vn_frame(0), synth.vn_frame(1), synth.tt_frame(0)]
    lbs = np.stack([PR.letterbox_ref(f)[0][:, :, ::-1] for f in frames])
    got = det.forward(lbs)
    ref = orc.forward(torch.from_numpy(lbs.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
    d = np.abs(got - ref)
    assert d[:, :4].max() < BOX_TOL and d[:, 4].max() < SCORE_TOL


def test_detector_forward_v1_vs_recorded_reference_run(det_v1):
    """The independent check: out0 of the reference's own graph + TRAINED weights as recorded by
    tests/golden/make_golden.py (OpenCV-DNN on yolo_plus.onnx).  Needs the trained model.ncnn.bin."""
    det, _ = det_v1
    from litepi_b200 import synth
    from helpers import model_paths
    if model_paths("vntsr")[1] is None:
        pytest.skip("trained v1 weights (model.ncnn.bin) are neither mounted nor staged under oracle/_ref: "
                    "the goldens were recorded with them")
    g = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    frames = [synth.vn_frame(0), synth.vn_frame(1), synth.tt_frame(0)]
    lbs = np.stack([PR.letterbox_ref(f)[0][:, :, ::-1] for f in frames])
    got = det.forward(lbs)
    for i, name in enumerate(["vn0", "vn1", "tt0"]):
        dg = np.abs(got[i] - g[f"{name}.out0"])
        assert dg[:4].max() < BOX_TOL and dg[4].max() < SCORE_TOL


@pytest.mark.parametrize("nc,size", [(3, 640), (1, 320), (5, 416)])
def test_detector_other_class_counts_and_input_sizes(lp, v1_paths, tmp_path, nc, size):
    """SURVEY 8(f)4 / ADVICE: the Detect tail takes its geometry (input size, anchors, nc) from the plan.  A variant
    of the reference graph with nc classes and another input size (random weights) must match the graph oracle,
    and decode + NMS must stay bit-exact on its multi-class out0."""
    from helpers import make_variant_param
    p = make_variant_param(v1_paths[0], str(tmp_path / f"nc{nc}_{size}.param"), nc=nc, in_size=size)
    det = lp.B200Detector(p, None, input_size=size, max_batch=2, max_det=4096, seed=11)
    orc = DetectorOracle(p, None, seed=11, in_size=size)
    _sync(orc, det.model)
    assert det.nc == nc and det.n_anchors == sum((size // s) ** 2 for s in (8, 16, 32))
    x = np.random.default_rng(nc * 1000 + size).integers(0, 256, (2, size, size, 3), dtype=np.uint8)
    got = det.forward(x)
    ref = orc.forward(torch.from_numpy(x.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
    assert got.shape == ref.shape == (2, 4 + nc, det.n_anchors)
    d = np.abs(got - ref)
    assert d[:, :4].max() < BOX_TOL and d[:, 4:].max() < SCORE_TOL
    # post-processing on the GPU's own out0 (identical input to both sides): bit-exact incl. per-class NMS groups
    thr = float(np.quantile(got[0, 4:].max(0), 0.97))
    _check_post(det, got[0], (size, size), 1.0, (0.0, 0.0), thr, 0.45)


def test_input_size_must_be_multiple_of_32(lp, v1_paths):
    with pytest.raises(ValueError):
        lp.B200Detector(v1_paths[0], None, input_size=500, max_batch=1)


def test_detector_forward_v2_random_weights(lp, v2_paths):
    det = lp.B200Detector(v2_paths[0], None, max_batch=2, seed=5)
    orc = DetectorOracle(v2_paths[0], None, seed=5)
    _sync(orc, det.model)
    x = np.random.default_rng(2).integers(0, 256, (2, 640, 640, 3), dtype=np.uint8)
    got = det.forward(x)
    ref = orc.forward(torch.from_numpy(x.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
    d = np.abs(got - ref)
    assert d[:, :4].max() < BOX_TOL and d[:, 4].max() < SCORE_TOL


def test_detector_batch_composition_invariant(det_v1):
    det, _ = det_v1
    x = np.random.default_rng(3).integers(0, 256, (5, 640, 640, 3), dtype=np.uint8)
    a = det.forward(x)
    b = np.concatenate([det.forward(x[i:i + 1]) for i in range(5)])
    assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------ K4/K5
def _check_post(det, out0, shape, r, pad, conf, iou):
    wb, ws, wc, (cb, cs, cc, widx) = PR.postprocess_ref(out0, shape, r, pad, conf, iou, return_candidates=True)
    gb, gs, gc = det.postprocess(out0, shape, r, pad, conf, iou)
    assert len(gb) == len(wb)
    assert int(det.n_cand[0]) == len(cb)
    if len(wb) == 0:
        assert gb.dtype == np.float64 and gb.shape == (0, 4) and gs.shape == (0,) and gc.shape == (0,)
        return 0
    assert gb.dtype == np.float32 and gs.dtype == np.float32 and gc.dtype == np.int64
    assert np.array_equal(gb, wb) and np.array_equal(gs, ws) and np.array_equal(gc, wc)
    assert np.array_equal(det.keep_idx[0, :len(gb)].cpu().numpy(), widx)
    return len(wb)


def test_decode_nms_bit_exact_golden(det_v1):
    det, _ = det_v1
    from litepi_b200 import synth
    g = np.load(os.path.join(GOLDEN, "detector_path.npz"))
    for name, shape in (("vn0", synth.VN_SHAPE), ("vn1", synth.VN_SHAPE), ("tt0", synth.TT_SHAPE)):
        r, pw, ph = (float(v) for v in g[f"{name}.ratio_pad"])
        for conf in (0.25, 0.001):
            _check_post(det, g[f"{name}.out0"], shape, r, (pw, ph), conf, 0.45)
            gb, gs, gc = det.postprocess(g[f"{name}.out0"], shape, r, (pw, ph), conf, 0.45)
            tag = f"{name}.c{conf}"                                   # recorded from the reference's own postprocess
            assert np.array_equal(gb, g[tag + ".boxes"]) and np.array_equal(gs, g[tag + ".scores"])


@pytest.mark.parametrize("seed,n_hot,nc", [(0, 0, 1), (1, 1, 1), (2, 40, 1), (3, 900, 1), (4, 8400, 1), (5, 300, 3)])
def test_decode_nms_bit_exact_synthetic(det_v1, seed, n_hot, nc):
    """empty, single, dense, every-anchor (8400 candidates) and multi-class inputs"""
    det, _ = det_v1
    rng = np.random.default_rng(seed)
    A = 8400
    out0 = np.zeros((4 + nc, A), np.float32)
    out0[0] = rng.uniform(0, 640, A); out0[1] = rng.uniform(138, 502, A)
    out0[2] = rng.uniform(4, 200, A); out0[3] = rng.uniform(4, 200, A)
    hot = rng.permutation(A)[:n_hot]
    sc = rng.permutation(A * nc).astype(np.float32).reshape(nc, A) / (A * nc) * 0.2          # distinct, < 0.25
    sc[rng.integers(0, nc, n_hot), hot] += 0.5
    out0[4:] = sc
    k = _check_post(det, out0, (681, 1198), 640 / 1198, (0.0, 138.0), 0.25, 0.45)
    assert (k > 0) == (n_hot > 0)


def test_nms_tie_order_and_idempotence(det_v1):
    det, _ = det_v1
    A = 64
    out0 = np.zeros((5, A), np.float32)
    out0[0] = 50 + 60 * (np.arange(A) % 8); out0[1] = 200 + 40 * (np.arange(A) // 8); out0[2] = 30; out0[3] = 30
    out0[4] = 0.5                                   # all tied: defined order = higher index first
    gb, gs, gc = det.postprocess(out0, (681, 1198), 640 / 1198, (0.0, 138.0), 0.25, 0.45)
    _check_post(det, out0, (681, 1198), 640 / 1198, (0.0, 138.0), 0.25, 0.45)
    kidx = det.keep_idx[0, :len(gb)].cpu().numpy()
    assert kidx[0] == A - 1 and np.all(np.diff(kidx) < 0)
    # idempotence: NMS of the kept set keeps everything (kept boxes pairwise iou <= thr)
    o2 = np.zeros((5, len(gb)), np.float32)
    o2[0] = (gb[:, 0] + gb[:, 2]) / 2; o2[1] = (gb[:, 1] + gb[:, 3]) / 2
    o2[2] = gb[:, 2] - gb[:, 0]; o2[3] = gb[:, 3] - gb[:, 1]; o2[4] = gs
    b2, _, _ = det.postprocess(o2, (2000, 2000), 1.0, (0.0, 0.0), 0.25, 0.45)
    assert len(b2) == len(gb)


# ------------------------------------------------------------------------------------------ K6
def test_roi_resize_bit_exact(clf):
    c, _ = clf
    from litepi_b200 import synth
    rng = np.random.default_rng(7)
    crops = synth.roi_crops(70, seed=1) + [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in
                                           [(64, 64), (10, 10), (200, 333), (64, 100), (130, 64), (500, 480), (3, 5),
                                            (1, 1), (1, 90), (700, 2)]]
    got = c.preprocess_batch(crops).cpu().numpy()
    for i, cr in enumerate(crops):
        assert np.array_equal(got[i], PR.classifier_input_ref(cr)[0]), f"crop {i} {cr.shape}"


def test_roi_resize_reference_crops_golden(clf):
    import cv2
    c, _ = clf
    paths = debug_roi_paths()
    if not paths:
        pytest.skip("reference debug_rois not staged")
    g = np.load(os.path.join(GOLDEN, "classifier_input.npz"))
    imgs = [cv2.imread(p) for p in paths]
    got = c.preprocess_batch(imgs).cpu().numpy()
    for i, p in enumerate(paths):
        assert np.array_equal(got[i], g[os.path.basename(p) + ".u8"])


def test_roi_select_bit_exact(lp, det_v1):
    det, _ = det_v1
    import ctypes as C
    from litepi_b200 import _lib as L
    from litepi_b200.detector import _ptr, _stream
    rng = np.random.default_rng(11)
    B, D = 3, det.max_det
    shapes = [(681, 1198), (2048, 2048), (50, 60)]
    counts = [37, 0, 9]
    boxes = np.zeros((B, D, 4), np.float32)
    for i, (h, w) in enumerate(shapes):
        x1 = rng.uniform(-5, w + 5, D); y1 = rng.uniform(-5, h + 5, D)
        bw = rng.choice([0.2, 3, 9, 40, 400], D); bh = rng.choice([0.2, 3, 9, 40, 400], D)
        boxes[i] = np.clip(np.stack([x1, y1, x1 + bw, y1 + bh], 1), 0, [w, h, w, h])
    tb = torch.from_numpy(boxes).to(det.device)
    tc = torch.tensor(counts, dtype=torch.int32, device=det.device)
    rx = torch.zeros((512, 4), dtype=torch.int32, device=det.device)
    rs = torch.zeros((512, 2), dtype=torch.int32, device=det.device)
    nr = torch.zeros((1,), dtype=torch.int32, device=det.device)
    hh = (C.c_int32 * B)(*[s[0] for s in shapes]); ww = (C.c_int32 * B)(*[s[1] for s in shapes])
    for min_area in (50, 100, 0):
        L.check(L.lib().lp_roi_select(det.ctx.handle, _ptr(tb), _ptr(tc), D, hh, ww, B, min_area, 512, _ptr(rx), _ptr(rs),
                                      _ptr(nr), _stream()))
        want_r, want_s = [], []
        for i in range(B):
            rois, valid = PR.roi_select_ref(boxes[i, :counts[i]], shapes[i], min_area)
            want_r.extend(rois.tolist()); want_s.extend([[i, k] for k in valid])
        n = int(nr.cpu()[0])
        assert n == len(want_r)
        assert rx[:n].cpu().numpy().tolist() == want_r and rs[:n].cpu().numpy().tolist() == want_s


# ------------------------------------------------------------------------------------------ K7
def test_classifier_logits_and_top1(clf):
    c, ref = clf
    from litepi_b200 import synth
    crops = synth.roi_crops(150, seed=2)                   # > max_batch: exercises chunking
    u8 = np.stack([PR.classifier_input_ref(cr)[0] for cr in crops])
    lg = c.logits_for(u8)
    x = (torch.from_numpy(u8.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        rl = ref(x.permute(0, 3, 1, 2)).numpy()
    assert np.abs(lg - rl).max() < LOGIT_TOL
    assert np.array_equal(lg.argmax(1), rl.argmax(1))
    cls, probs = c.predict_batch(crops)
    wcls, wprobs, _ = PR.classify_ref(ref, crops)
    assert np.array_equal(cls, wcls) and np.abs(probs - wprobs).max() < 1e-4
    assert probs.dtype == np.float32 and probs.shape == (150, 49)
    e = c.predict_batch([])
    assert e[0].shape == (0,) and e[1].shape == (0,)


def test_more_than_16_contexts_keep_classifiers_alive(lp):
    """the fused-classifier state used to live in a 16-slot global table indexed by a wrapping counter, so
    the 17th context overwrote (and freed the park buffer of) a live classifier.  It now lives in the context."""
    ref = PR.build_shufflenet(49, seed=8)
    x = np.random.default_rng(8).integers(0, 256, (5, 64, 64, 3), dtype=np.uint8)
    first = lp.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=8)
    want = first.logits_for(x)
    keep = []
    for i in range(20):
        other = PR.build_shufflenet(49, seed=100 + i)
        c = lp.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=other.state_dict(), max_batch=8,
                              fused=(i % 3 != 0))
        c.logits_for(x)
        keep.append(c)
        if i % 5 == 4:
            keep.pop(0)                       # destroy some while others stay alive
    assert np.array_equal(first.logits_for(x), want)
    with torch.no_grad():
        rl = ref((torch.from_numpy(x.astype(np.float32)) / 255 - 0.18).div(0.34).permute(0, 3, 1, 2)).numpy()
    assert np.abs(want - rl).max() < LOGIT_TOL


@pytest.mark.parametrize("mode", ["fused", "layered_tc", "layered_simt"])
def test_classifier_paths_agree(lp, mode):
    """the three classifier execution paths (fused persistent kernel, layer-by-layer plan on the tensor
    cores, layer-by-layer plan on the SIMT kernels) all reproduce torchvision"""
    from litepi_b200 import synth
    ref = PR.build_shufflenet(49, seed=2)
    c = lp.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=32,
                          tensor_cores=(mode != "layered_simt"), fused=(mode == "fused"))
    u8 = np.stack([PR.classifier_input_ref(cr)[0] for cr in synth.roi_crops(45, seed=5)])     # 45: ragged last group/chunk
    lg = c.logits_for(u8)
    x = (torch.from_numpy(u8.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        rl = ref(x.permute(0, 3, 1, 2)).numpy()
    assert np.abs(lg - rl).max() < LOGIT_TOL and np.array_equal(lg.argmax(1), rl.argmax(1))


@pytest.mark.parametrize("tc", [True, False])
def test_detector_tensor_core_and_simt_paths(lp, v1_paths, tc):
    det = lp.B200Detector(v1_paths[0], v1_paths[1], max_batch=3, seed=0, tensor_cores=tc)
    orc = DetectorOracle(v1_paths[0], v1_paths[1], seed=0)
    _sync(orc, det.model)
    assert (det.tc_ops > 40) == tc
    x = np.random.default_rng(9).integers(0, 256, (3, 640, 640, 3), dtype=np.uint8)
    got = det.forward(x)
    ref = orc.forward(torch.from_numpy(x.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
    d = np.abs(got - ref)
    assert d[:, :4].max() < BOX_TOL and d[:, 4].max() < SCORE_TOL


@pytest.mark.parametrize("which,size", [("v1", 640), ("v1", 416), ("v2", 640)])
def test_fused_c2f_body_equals_layer_by_layer(lp, v1_paths, v2_paths, monkeypatch, tmp_path, which, size):
    """csrc/c2f_mma.cu runs the bottleneck chain + cv2 of the c = 8 / 16 C2f blocks in one kernel.  Same arithmetic as the
    layer-by-layer kernels (split-f16 three-product MMAs, fp32 accumulate, re-split between layers): out0 differs only by
    fp32 summation order.  416 = 52 x 52 / 104 x 104 maps: tiles overhang the image on both axes."""
    from litepi_b200 import _lib as L
    param, binp = (v1_paths if which == "v1" else (v2_paths[0], None))
    if size != 640:
        from helpers import make_variant_param
        param = make_variant_param(param, str(tmp_path / f"c2f_{size}.param"), nc=1, in_size=size)
    x = np.random.default_rng(21).integers(0, 256, (3, size, size, 3), dtype=np.uint8)
    det = lp.B200Detector(param, binp if size == 640 else None, input_size=size, max_batch=3, seed=3)
    fused = det.forward(x)
    paths = det.ctx.op_paths(L.NET_DETECTOR)
    n_fused = sum(1 for v in paths if v == 5)
    assert n_fused >= (3 if which == "v1" else 1), paths
    monkeypatch.setenv("LP_NO_C2F", "1")
    det2 = lp.B200Detector(param, binp if size == 640 else None, input_size=size, max_batch=3, seed=3)
    monkeypatch.delenv("LP_NO_C2F")
    layered = det2.forward(x)
    assert 5 not in det2.ctx.op_paths(L.NET_DETECTOR)
    d = np.abs(fused - layered)
    assert d[:, :4].max() < 2e-3 and d[:, 4:].max() < 1e-4, (d[:, :4].max(), d[:, 4:].max())
    orc = DetectorOracle(param, binp if size == 640 else None, seed=3, in_size=size)
    _sync(orc, det.model)
    ref = orc.forward(torch.from_numpy(x.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
    d = np.abs(fused - ref)
    assert d[:, :4].max() < BOX_TOL and d[:, 4:].max() < SCORE_TOL


@pytest.mark.parametrize("which", ["v1", "v2"])
def test_tma_patch_loads_equal_cp_async_loads(lp, v1_paths, v2_paths, monkeypatch, which):
    """conv_tc.cu moves the activation patches by TMA tensor copies (chunk-major boxes for the 3x3 layers, SWIZZLE_128B pixel rows
    for the 1x1 layers with cin % 64 == 0) or, with LP_TC_TMA=0, by cp.async from five loader warps.  Only the transport and the
    shared-memory layout differ: K order and accumulation order are the same, so out0 must be IDENTICAL bit for bit -- including
    the zero fill of the halo, the partial last tile of a 1x1 layer (batch 3 of a max_batch 4 plan) and frames of different
    content in one batch."""
    param, binp = (v1_paths if which == "v1" else (v2_paths[0], None))
    x = np.random.default_rng(33).integers(0, 256, (3, 640, 640, 3), dtype=np.uint8)
    a = lp.B200Detector(param, binp, max_batch=4, seed=2).forward(x)
    monkeypatch.setenv("LP_TC_TMA", "0")
    b = lp.B200Detector(param, binp, max_batch=4, seed=2).forward(x)
    monkeypatch.setenv("LP_TC_TMA", "3")
    monkeypatch.setenv("LP_TC_SW128", "0")
    c = lp.B200Detector(param, binp, max_batch=4, seed=2).forward(x)
    monkeypatch.delenv("LP_TC_TMA"); monkeypatch.delenv("LP_TC_SW128")
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_classifier_default_init_state_dict(lp):
    """the reference's own construction: torchvision random init + fc swap, default BatchNorm"""
    import torch.nn as nn
    from torchvision import models
    torch.manual_seed(3)
    m = models.shufflenet_v2_x1_0(weights=None)
    m.fc = nn.Linear(m.fc.in_features, 91)
    m.eval()
    c = lp.B200Classifier(None, "shufflenetv2", num_classes=91, state_dict=m.state_dict(), max_batch=16)
    u8 = np.random.default_rng(4).integers(0, 256, (16, 64, 64, 3), dtype=np.uint8)
    x = (torch.from_numpy(u8.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        rl = m(x.permute(0, 3, 1, 2)).numpy()
    lg = c.logits_for(u8)
    assert np.abs(lg - rl).max() < LOGIT_TOL and np.array_equal(lg.argmax(1), rl.argmax(1))


# ------------------------------------------------------------------------------------------ pipeline
def _compare(got, want):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.abs(a["box_f32"] - b["box_f32"]).max() < BOX_TOL
        assert abs(a["det_conf"] - b["det_conf"]) < SCORE_TOL
        assert a["det_class"] == b["det_class"] and a["cls_class"] == b["cls_class"]
        assert abs(a["cls_conf"] - b["cls_conf"]) < 1e-3


def test_pipeline_vs_oracle(lp, v1_paths, clf):
    from litepi_b200 import synth
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=4,
                           classifier_state_dict=ref.state_dict(), seed=0)
    orc = DetectorOracle(v1_paths[0], v1_paths[1], seed=0)
    _sync(orc, pipe.detector.model)
    frames = [synth.vn_frame(i) for i in range(5)] + [synth.tt_frame(0)]        # 6 frames, max_batch 4: two chunks
    got = pipe.run_batch(frames, 0.25, 0.45, 50)
    for f, g in zip(frames, got):
        _compare(g, oracle_pipeline_run(orc, ref, f, 0.25, 0.45, 50))
    # reference-shaped single-frame API
    res, m = pipe.run(frames[0], conf_threshold=0.25, iou_threshold=0.45, min_area=50)
    want = oracle_pipeline_run(orc, ref, frames[0], 0.25, 0.45, 50)
    assert len(res) == len(want) and m.num_detections >= len(res) and m.t_total > 0 and m.fps > 0
    for a, b in zip(res, want):
        assert a["bbox"] == b["bbox"] and a["cls_class"] == b["cls_class"]
        assert set(a) == {"bbox", "det_class", "det_conf", "cls_class", "cls_conf", "time_det", "time_cls"}
    # a frame with nothing in it
    blank = np.full((681, 1198, 3), 127, np.uint8)
    res, m = pipe.run(blank, 0.25, 0.45, 50)
    assert res == [] and m.num_detections == 0
    assert pipe.run_batch([blank], 0.25, 0.45, 50) == [[]]
    b, s, c = pipe.detector.detect(blank, 0.25, 0.45)
    assert b.shape == (0, 4) and b.dtype == np.float64


def test_full_batch_properties(lp, v1_paths, clf):
    """BASELINE configs[1] size (batch 64 VN frames): batch result == per-frame results, and the
    gather/sort of records is a permutation-invariant of the sharding."""
    from litepi_b200 import synth, runner
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=64,
                           classifier_state_dict=ref.state_dict(), seed=0)
    frames = [synth.vn_frame(i) for i in range(64)]
    whole = runner.run_sharded(pipe, frames, 0.25, 0.45, 50, 0, 1)
    parts = np.concatenate([runner.run_sharded(pipe, frames, 0.25, 0.45, 50, r, 4) for r in range(4)])
    parts = runner.sort_records(parts)
    assert whole.shape == parts.shape and np.array_equal(whole, parts)
    assert whole.shape[0] > 100
    got = pipe.records_to_results(whole, 64)
    one = pipe.run_batch([frames[17]], 0.25, 0.45, 50)[0]
    assert [d["bbox"] for d in got[17]] == [d["bbox"] for d in one]
    assert [d["cls_class"] for d in got[17]] == [d["cls_class"] for d in one]


def test_config3_tt100k_dense_batch(lp, v1_paths):
    """BASELINE configs[2] shape: 32 frames of 2048x2048 with dense small signs, 91 classes.  Two frames are
    checked against the oracle end to end; at full size the batch must equal its per-frame runs, every ROI
    must be a valid integer box of the frame, and records must survive the shard/gather/sort round trip."""
    from litepi_b200 import synth, runner
    ref = PR.build_shufflenet(91, seed=3)
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=91, max_batch=32,
                           classifier_state_dict=ref.state_dict(), seed=0)
    orc = DetectorOracle(v1_paths[0], v1_paths[1], seed=0)
    _sync(orc, pipe.detector.model)
    frames = [synth.tt_frame(i) for i in range(32)]
    got = pipe.run_batch(frames, 0.25, 0.45, 50)
    assert sum(len(g) for g in got) > 32 * 5                       # dense: well over 5 signs per frame survive
    for i in (0, 19):
        _compare(got[i], oracle_pipeline_run(orc, ref, frames[i], 0.25, 0.45, 50))
    for i in (3, 31):
        one = pipe.run_batch([frames[i]], 0.25, 0.45, 50)[0]
        assert [d["bbox"] for d in one] == [d["bbox"] for d in got[i]]
        assert [d["cls_class"] for d in one] == [d["cls_class"] for d in got[i]]
    for g in got:
        for d in g:
            x1, y1, x2, y2 = d["bbox"]
            assert all(isinstance(v, (int, np.integer)) for v in d["bbox"]) and 0 <= d["cls_class"] < 91   # e2e.py:522 yields numpy ints
            assert x2 >= x1 and y2 >= y1 and x1 >= 0 and y1 >= 0 and x2 <= 2048 and y2 <= 2048    # postprocess clips to the frame
    whole = runner.run_sharded(pipe, frames, 0.25, 0.45, 50, 0, 1)
    parts = runner.sort_records(np.concatenate([runner.run_sharded(pipe, frames, 0.25, 0.45, 50, r, 8) for r in range(8)]))
    assert np.array_equal(whole, parts)


def test_config4_classifier_1024_crops(lp):
    """BASELINE configs[3]: ShuffleNetV2 alone on 1024 crops (the 15 real debug_rois of the reference + synthetic
    glyph crops of 10..90 px): Pillow-exact resize, logits within 1e-2 and top-1 equal to torchvision for all."""
    import cv2
    from litepi_b200 import synth
    ref = PR.build_shufflenet(49, seed=4)
    c = lp.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=1024)
    real = [cv2.imread(p) for p in debug_roi_paths()]
    assert len(real) == 15 and all(r is not None for r in real)
    crops = real + synth.roi_crops(1024 - len(real), seed=7)
    u8 = np.stack([PR.classifier_input_ref(cr)[0] for cr in crops])
    assert np.array_equal(c.preprocess_batch(crops).cpu().numpy(), u8)      # K6 bit-exact on all 1024
    lg = c.logits_for(u8)
    x = (torch.from_numpy(u8.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        rl = ref(x.permute(0, 3, 1, 2)).numpy()
    assert np.abs(lg - rl).max() < LOGIT_TOL
    assert np.array_equal(lg.argmax(1), rl.argmax(1))
    cls, probs = c.predict_batch(crops)
    assert np.array_equal(cls, rl.argmax(1)) and np.allclose(probs.sum(1), 1.0, atol=1e-5)
    # composition invariance: any sub-batch gives the same rows
    cls2, probs2 = c.predict_batch(crops[100:133])
    assert np.array_equal(cls2, cls[100:133]) and np.abs(probs2 - probs[100:133]).max() < 1e-6


# ------------------------------------------------------------------------------------------ evaluation (8f.1)
@pytest.mark.parametrize("case", ["small", "mid", "wide", "empty"])
def test_eval_equals_reference_golden(lp, case):
    """Evaluator.evaluate_predictions (GPU matching + host curves) against the recorded outputs of the reference's
    evaluate_predictions: every key bit-exact (bool/int matching, float64 curves)."""
    from helpers import load_eval_case
    preds, gts, nc, want = load_eval_case(case)
    got = lp.Evaluator().evaluate_predictions(preds, gts, nc)
    assert set(got) == set(want)
    for k, v in want.items():
        assert np.array_equal(np.asarray(got[k]), v), k


def test_eval_match_vs_oracle_and_edges(lp):
    from oracle import eval_ref as ER
    ev = lp.Evaluator()
    rng = np.random.default_rng(5)
    # 300 frames, up to 60 predictions and 40 ground truths each, integer boxes on both sides (many exact overlaps)
    pn = rng.integers(0, 61, 300); gn = rng.integers(0, 41, 300)
    pn[7] = 0; gn[9] = 0; pn[11] = gn[11] = 0

    def boxes(n):
        xy = rng.integers(0, 600, (n, 2)); wh = rng.integers(8, 120, (n, 2))
        return np.concatenate([xy, xy + wh], 1).astype(np.float64)
    pb, gb = boxes(int(pn.sum())), boxes(int(gn.sum())) + rng.uniform(0, 0.5, (int(gn.sum()), 4))
    pc, gc = rng.integers(0, 5, int(pn.sum())), rng.integers(0, 5, int(gn.sum()))
    got = ev.match(pb, pc, pn, gb, gc, gn)
    a = b = 0
    for f in range(300):
        want = ER.match_image_ref(pb[a:a + pn[f]], pc[a:a + pn[f]], gb[b:b + gn[f]], gc[b:b + gn[f]])
        assert np.array_equal(got[a:a + pn[f]], want), f
        a += pn[f]; b += gn[f]
    assert got.any() and not got.all()
    # identical boxes: iou = 1 - eps >= 0.95 at every threshold; the LOWEST-index duplicate prediction owns the ground truth
    one = np.array([[10., 10., 50., 50.]])
    c = ev.match(np.repeat(one, 3, 0), np.array([1, 1, 1]), np.array([3]), one, np.array([1]), np.array([1]))
    assert c[0].all() and not c[1:].any()
    # an exact IoU tie between two ground truths goes to the higher ground-truth index (documented tie rule)
    g2 = np.array([[0., 0., 10., 10.], [20., 0., 30., 10.]])
    c = ev.match(np.array([[5., 0., 25., 10.]]), np.array([2]), np.array([1]), g2, np.array([1, 2]), np.array([2]),
                 thresholds=np.array([0.1]))
    assert c[0, 0]
    assert ev.match(np.zeros((0, 4)), np.zeros(0), np.array([0, 0]), one, np.array([1]), np.array([1, 0])).shape == (0, 10)


def test_eval_of_pipeline_records(lp, v1_paths, clf):
    """End to end: records of a batch evaluated against ground truth made from the pipeline's own output score
    the maximum mAP50 = mAP50-95 on the classes present; dropping every second ground truth turns the unmatched
    detections into false positives and lowers it."""
    from litepi_b200 import synth
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=16,
                           classifier_state_dict=ref.state_dict(), seed=0)
    frames = [synth.vn_frame(i) for i in range(16)]
    fb = lp.detector.FrameBatch.from_host(frames, pipe.device)
    rec = pipe.fetch_records(pipe.run_device(fb, 0.25, 0.45, 50))
    res = pipe.records_to_results(rec, 16)
    gts = [[[d["cls_class"], *[float(v) for v in d["bbox"]]] for d in r] for r in res]
    ev = lp.Evaluator()
    m = ev.evaluate_records(rec, 16, gts, 49)
    # a perfect detector scores 0.995 under the reference's 101-point rule (the sentinel (recall 1, precision 0)
    # costs half of the last 0.01-wide trapezoid)
    assert m["mAP50"] == pytest.approx(0.995, abs=1e-9) and m["mAP50_95"] == pytest.approx(0.995, abs=1e-9)
    assert m["fp"].sum() == 0 and m["fn"].sum() == 0 and m["tp"].sum() == rec.shape[0]
    half = [g[::2] for g in gts]
    m2 = ev.evaluate_records(rec, 16, half, 49)
    assert m2["mAP50"] < 0.995 and m2["tp"].sum() <= sum(len(g) for g in half)


# ------------------------------------------------------------------------------------------ e2e_optimize.py mode (8f.3)
def test_optimized_roi_mode(lp, v1_paths, clf):
    """roi_mode="optimized": ROI integers and the cv2-INTER_LINEAR classifier input are bit-exact with the
    restatement of e2e_optimize.py; the default mode is untouched."""
    from litepi_b200 import synth
    c, ref = clf
    crops = synth.roi_crops(120, seed=9) + [np.random.default_rng(1).integers(0, 256, (64, 64, 3), dtype=np.uint8)]
    c.set_preprocess("cv2")
    try:
        got = c.preprocess_batch(crops).cpu().numpy()
    finally:
        c.set_preprocess("pil")
    want = np.stack([PR.classifier_input_opt_ref(cr) for cr in crops])
    assert np.array_equal(got, want)
    assert np.array_equal(c.preprocess_batch(crops[:8]).cpu().numpy(), np.stack([PR.classifier_input_ref(cr)[0] for cr in crops[:8]]))
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=8,
                           classifier_state_dict=ref.state_dict(), seed=0, roi_mode="optimized")
    frames = [synth.vn_frame(i) for i in range(6)] + [synth.tt_frame(1)]
    fb = lp.detector.FrameBatch.from_host(frames, pipe.device)
    n = pipe.run_device(fb, 0.25, 0.45, 50)
    rois = pipe.roi_xyxy[:n].cpu().numpy(); src = pipe.roi_src[:n].cpu().numpy()
    dets = pipe.detector._collect(len(frames))
    want_r, want_s, want_in = [], [], []
    for i, f in enumerate(frames):
        r, valid = PR.roi_select_opt_ref(dets[i][0], f.shape, 50)
        want_r.extend(r.tolist()); want_s.extend([[i, k] for k in valid])
        want_in.extend(PR.classifier_input_opt_ref(f[y1:y2, x1:x2]) for x1, y1, x2, y2 in r)
    assert n == len(want_r) and n > 10
    assert rois.tolist() == want_r and src.tolist() == want_s
    assert np.array_equal(pipe.classifier.cls_in[:n].cpu().numpy(), np.stack(want_in))
    with pytest.raises(ValueError):
        lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, roi_mode="fast")


def test_async_step_api_equals_synchronous(lp, v1_paths, clf):
    """enqueue_device / enqueue_fetch / collect (no host round trip inside a step, device-side ROI count, double-buffered
    host mirrors) return exactly the records of run_device + fetch_records, also when two steps are in flight."""
    from litepi_b200 import synth
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=8,
                           classifier_state_dict=ref.state_dict(), seed=0)
    fa = lp.detector.FrameBatch.from_host([synth.vn_frame(i) for i in range(8)], pipe.device)
    fb = lp.detector.FrameBatch.from_host([synth.vn_frame(i) for i in range(8, 14)] + [np.full((681, 1198, 3), 127, np.uint8)], pipe.device)
    want_a = pipe.fetch_records(pipe.run_device(fa, 0.25, 0.45, 50))
    want_b = pipe.fetch_records(pipe.run_device(fb, 0.25, 0.45, 50))
    assert want_a.shape[0] > 0 and want_b.shape[0] > 0
    pipe.enqueue_device(fa, 0.25, 0.45, 50, slot=0); pipe.enqueue_fetch(0)
    pipe.enqueue_device(fb, 0.25, 0.45, 50, slot=1); pipe.enqueue_fetch(1)
    got_a, got_b = pipe.collect(0), pipe.collect(1)
    assert np.array_equal(got_a, want_a) and np.array_equal(got_b, want_b)
    # a batch with nothing in it
    blank = lp.detector.FrameBatch.from_host([np.full((480, 640, 3), 90, np.uint8)] * 2, pipe.device)
    pipe.enqueue_device(blank, 0.25, 0.45, 50, slot=0); pipe.enqueue_fetch(0)
    assert pipe.collect(0).shape == (0, 9)


def test_stream_api_equals_run_batch(lp, v1_paths, clf):
    """B200Pipeline.stream() / run_stream (pinned ring, copy stream, two lanes, one CUDA graph per step) returns, batch by
    batch, exactly what the synchronous run_batch returns: for frames produced in place in the pinned ring, for
    ordinary pageable numpy frames, for pinned torch tensors, with a short last batch, and on graph replays."""
    from litepi_b200 import synth
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=4,
                           classifier_state_dict=ref.state_dict(), seed=0)
    frames = [synth.vn_frame(i) for i in range(22)]
    batches = [frames[i:i + 4] for i in range(0, 22, 4)]            # 5 full batches + one of 2
    want = [pipe.run_batch(b, 0.25, 0.45, 50) for b in batches]
    strip = lambda res: [[(d["bbox"], d["det_class"], d["cls_class"], d["det_conf"], d["cls_conf"]) for d in fr] for fr in res]
    # (1) public generator API on pageable frames, twice (second pass replays the captured graphs)
    for _ in range(2):
        got = list(pipe.run_stream(batches, 0.25, 0.45, 50, lanes=2))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert strip(g) == strip(w)
    sr = pipe._stream_runner
    assert sr.graph_failed is None, sr.graph_failed
    assert sr.steps_graph >= 8 and sr.steps_direct >= 2              # full batches replay graphs, the short one launches directly
    # (2) frames produced in place in the pinned ring + explicit global frame ids
    recs = []
    def produce():
        for s, b in enumerate(batches[:5]):
            buf = sr.host_buffer(s, 681, 1198)
            for i, f in enumerate(b):
                buf[i] = f
            yield buf
    ids = [[100 * s + i for i in range(4)] for s in range(5)]
    for s, rec in enumerate(sr.run_stream(produce(), 0.25, 0.45, 50, frame_ids=ids)):
        local = rec.copy(); local[:, 0] -= 100 * s
        assert strip(pipe.records_to_results(local, 4)) == strip(want[s])
        assert set(rec[:, 0].tolist()) <= set(ids[s])
    # (3) pinned torch tensors are read by the copy engine directly
    pinned = [torch.from_numpy(np.stack(b)).pin_memory() for b in batches[:3]]
    for s, rec in enumerate(sr.run_stream(pinned, 0.25, 0.45, 50)):
        assert strip(pipe.records_to_results(rec, 4)) == strip(want[s])
    # (4) the synchronous latency path
    one = sr.run_one(batches[1], 0.25, 0.45, 50)
    assert strip(pipe.records_to_results(one, 4)) == strip(want[1])
    with pytest.raises(ValueError):
        list(sr.run_stream([frames[:5]], 0.25, 0.45, 50))           # more than max_batch


@pytest.mark.parametrize("arch", ["resnet18", "mobilenetv2", "efficientnet"])
def test_other_classifier_archs(lp, v1_paths, arch):
    """the reference's other --clf_arch choices (build_classifier, e2e.py:322-335) on the GPU: logits within
    1e-2 of torchvision, top-1 equal, predict_batch contract; and the whole pipeline with that classifier against the oracle."""
    from litepi_b200 import synth
    ref = PR.build_classifier_ref(arch, 49, seed=3)
    c = lp.B200Classifier(None, arch, num_classes=49, state_dict=ref.state_dict(), max_batch=64)
    assert not c.fused
    crops = synth.roi_crops(70, seed=11)                        # 70 > max_batch: two chunks
    u8 = np.stack([PR.classifier_input_ref(cr)[0] for cr in crops])
    lg = c.logits_for(u8)
    with torch.no_grad():
        rl = ref(((torch.from_numpy(u8.astype(np.float32)) / 255 - 0.18) / 0.34).permute(0, 3, 1, 2)).numpy()
    assert np.abs(lg - rl).max() < LOGIT_TOL
    assert np.array_equal(lg.argmax(1), rl.argmax(1))
    cls, probs = c.predict_batch(crops)
    assert np.array_equal(cls, rl.argmax(1)) and probs.shape == (70, 49) and np.allclose(probs.sum(1), 1.0, atol=1e-5)
    if arch == "resnet18":
        pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, arch, num_classes=49, max_batch=2,
                               classifier_state_dict=ref.state_dict(), seed=0)
        orc = DetectorOracle(v1_paths[0], v1_paths[1], seed=0)
        _sync(orc, pipe.detector.model)
        frames = [synth.vn_frame(i) for i in range(3)]
        got = pipe.run_batch(frames, 0.25, 0.45, 50)
        for f, g in zip(frames, got):
            rois_want = oracle_pipeline_run(orc, ref, f, 0.25, 0.45, 50)
            _compare(g, rois_want)
        streamed = list(pipe.run_stream([frames[:2], frames[2:]], 0.25, 0.45, 50))
        assert [[d["bbox"] for d in fr] for b in streamed for fr in b] == [[d["bbox"] for d in fr] for fr in got]
        assert [[d["cls_class"] for d in fr] for b in streamed for fr in b] == [[d["cls_class"] for d in fr] for fr in got]


def test_unknown_classifier_arch_raises(lp):
    with pytest.raises(ValueError):
        lp.B200Classifier(None, "vgg16", num_classes=49)


# ------------------------------------------------------------------------------------------ frame ingest (JPEG)
def _jpeg(a, q=90, sf=None, rst=0):
    import cv2
    par = [cv2.IMWRITE_JPEG_QUALITY, q]
    if sf is not None:
        par += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf]
    if rst:
        par += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    ok, b = cv2.imencode(".jpg", a, par)
    assert ok
    return bytes(b)


@pytest.mark.parametrize("h,w,q,sf,rst", [(64, 48, 90, None, 0), (37, 53, 75, None, 0), (120, 160, 95, None, 4), (8, 8, 100, None, 0),
                                          (33, 17, 90, "444", 3), (40, 72, 85, "422", 2), (9, 200, 90, None, 1), (1, 1, 90, None, 0),
                                          (681, 1198, 90, None, 4), (512, 512, 30, None, 7), (100, 100, 10, None, 0)])
def test_jpeg_decode_bit_exact_vs_cv2(lp, h, w, q, sf, rst):
    """device JPEG decode == cv2.imdecode (what the reference's cv2.imread returns, e2e.py:962), bit for bit:
    sampling 4:2:0 / 4:2:2 / 4:4:4, odd sizes, with and without restart intervals, several qualities, batches."""
    import cv2
    from litepi_b200 import _lib as L
    from litepi_b200.jpeg import JpegBatchDecoder
    sfv = {None: None, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422}[sf]
    rng = np.random.default_rng(h * 7 + w)
    imgs = []
    for k in range(3):
        small = rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
        a = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
        imgs.append(np.clip(a.astype(np.int16) + rng.integers(-8, 9, a.shape), 0, 255).astype(np.uint8))
    js = [_jpeg(a, q, sfv, rst) for a in imgs]
    dec = JpegBatchDecoder(L.context(0), torch.device("cuda", 0), max_batch=4)
    got = dec.decode(js).cpu().numpy()
    for k, j in enumerate(js):
        want = cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(got[k], want), f"image {k}: {int((got[k] != want).sum())} bytes differ"


def test_jpeg_grey_and_errors(lp):
    import cv2
    from litepi_b200 import _lib as L
    from litepi_b200.jpeg import JpegBatchDecoder
    dec = JpegBatchDecoder(L.context(0), torch.device("cuda", 0), max_batch=2)
    g = np.random.default_rng(1).integers(0, 256, (50, 70), dtype=np.uint8)
    j = _jpeg(g, 85)
    assert np.array_equal(dec.decode([j])[0].cpu().numpy(), cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_COLOR))
    a = np.random.default_rng(2).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    with pytest.raises(ValueError, match="share one header"):
        dec.decode([_jpeg(a, 90), _jpeg(a, 50)])
    ok, prog = cv2.imencode(".jpg", a, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError, match="baseline"):
        dec.decode([bytes(prog)])


def test_stream_from_jpeg_bytes_equals_decoded_frames(lp, v1_paths, clf):
    """run_stream on JPEG byte strings (decode on the device, only the entropy-coded bytes cross PCIe) returns exactly what
    run_batch returns on the frames cv2.imdecode produces from the same bytes."""
    import cv2
    from litepi_b200 import synth
    _, ref = clf
    pipe = lp.B200Pipeline(v1_paths[0], v1_paths[1], None, "shufflenetv2", num_classes=49, max_batch=4,
                           classifier_state_dict=ref.state_dict(), seed=0)
    js = [_jpeg(synth.vn_frame(i), 90, None, 4) for i in range(10)]
    frames = [cv2.imdecode(np.frombuffer(j, np.uint8), cv2.IMREAD_COLOR) for j in js]
    batches = [js[i:i + 4] for i in range(0, 10, 4)]
    want = [pipe.run_batch(frames[i:i + 4], 0.25, 0.45, 50) for i in range(0, 10, 4)]
    strip = lambda res: [[(d["bbox"], d["det_class"], d["cls_class"], d["det_conf"], d["cls_conf"]) for d in fr] for fr in res]
    for _ in range(2):
        got = list(pipe.run_stream(batches, 0.25, 0.45, 50))
        assert [strip(g) for g in got] == [strip(w) for w in want]
    assert sum(len(fr) for w in want for fr in w) > 10
