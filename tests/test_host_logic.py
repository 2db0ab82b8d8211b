"""CPU tests of the host side: model reader, plan lowering (proved against the oracle through the
plan interpreter), C-ABI surface, loud failure without a GPU, multi-rank gather over gloo."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import ROOT
import litepi_b200
from litepi_b200 import _lib as L, ncnn_model, plan, runner, synth
from litepi_b200.pipeline import B200Pipeline, PipelineMetrics
from oracle import pipeline_ref as PR
from oracle.ncnn_graph import DetectorOracle
from plan_interp import detect_tail_cpu, run_plan_cpu


def _sync_weights(orc, model):
    ci = iter(model.convs)
    for ly in orc.layers:
        if ly.type == "Convolution":
            c = next(ci)
            ly.weight, ly.bias = c.weight, c.bias


def test_ncnn_reader(v1_paths, v2_paths):
    m1 = ncnn_model.load_ncnn(*v1_paths)
    m2 = ncnn_model.load_ncnn(v2_paths[0], None, seed=1)
    for m in (m1, m2):
        assert len(m.convs) == 64 and m.c2f_depths == [1, 2, 2, 1, 1, 1, 1, 1] and m.n_anchors == 8400
    assert [m1.convs[i].cout for i in (0, 1, 6, 13, 20)] == [8, 16, 32, 64, 128]
    assert [m2.convs[i].cout for i in (0, 1, 6, 13, 20)] == [16, 24, 48, 96, 192]
    n1 = sum(c.weight.size + (c.bias.size if c.bias is not None else 0) for c in m1.convs)
    assert n1 == 966355                                  # weight count of v1
    with pytest.raises(RuntimeError, match="Failed to load param"):
        ncnn_model.load_ncnn("/nonexistent.param", None)
    with pytest.raises(RuntimeError, match="Failed to load bin"):
        ncnn_model.load_ncnn(v1_paths[0], "/nonexistent.bin")


@pytest.mark.parametrize("which", ["v1", "v2"])
def test_detector_plan_reproduces_oracle(which, v1_paths, v2_paths):
    param, binp = v1_paths if which == "v1" else v2_paths
    model = ncnn_model.load_ncnn(param, binp, seed=3)
    P = plan.build_detector_plan(model)
    assert sum(P.macs) == (1418713600 if which == "v1" else 2542483200)       # SURVEY App. A.2 totals
    x = np.random.default_rng(0).integers(0, 256, (1, 640, 640, 3), dtype=np.uint8)
    bufs, _ = run_plan_cpu(P, x)
    out = detect_tail_cpu(bufs[-1][:, :, 0, :])
    orc = DetectorOracle(param, binp, seed=3)
    _sync_weights(orc, model)
    ref = orc.forward(torch.from_numpy(x.astype(np.float32) / 255).permute(0, 3, 1, 2))
    d = (out - ref).abs()
    assert float(d[:, :4].max()) < 2e-3 and float(d[:, 4].max()) < 1e-5
    # split-f16 buffers: every view starts on an 8-channel (16-byte) boundary
    for op in P.ops:
        if P.bufs[op["in_buf"]]["fmt"] == L.FMT_SPLIT16:
            assert op["in_coff"] % 8 == 0 and op["cin"] % 8 == 0
        if op["out_buf"] >= 0 and P.bufs[op["out_buf"]]["fmt"] == L.FMT_SPLIT16:
            assert op["out_coff"] % 8 == 0


def test_classifier_plan_reproduces_torchvision():
    model = PR.build_shufflenet(49, seed=0)
    P = plan.build_classifier_plan(model.state_dict())
    assert abs(sum(P.macs) - 11.80e6) < 0.01e6
    x = np.random.default_rng(1).integers(0, 256, (3, 64, 64, 3), dtype=np.uint8)
    _, logits = run_plan_cpu(P, x)
    xin = (torch.from_numpy(x.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        ref = model(xin.permute(0, 3, 1, 2))
    assert float((logits - ref).abs().max()) < 1e-5


def test_fused_classifier_program_reproduces_torchvision():
    """the step lists the fused kernel executes (front / middle / tail with 3 ROIs stacked), run on the CPU with the
    same shared-memory maps: offsets, overlays and the park/load hand-off are right if torchvision is reproduced"""
    from plan_interp import run_fused_cpu
    model = PR.build_shufflenet(49, seed=5)
    x = np.random.default_rng(2).integers(0, 256, (7, 64, 64, 3), dtype=np.uint8)
    xin = (torch.from_numpy(x.astype(np.float32)) / 255 - 0.18) / 0.34
    with torch.no_grad():
        ref = model(xin.permute(0, 3, 1, 2)).numpy()
    for gt in (1, 3):
        prog = plan.build_fused_classifier(model.state_dict(), tail_group=gt)
        assert prog.n_front == 7 and prog.n_front + prog.n_mid + prog.n_tail == len(prog.steps)
        got = run_fused_cpu(prog, x, grid=2)
        assert float(np.abs(got - ref).max()) < 1e-4


def test_plan_rejects_foreign_graph(v1_paths, tmp_path):
    model = ncnn_model.load_ncnn(*v1_paths)
    model.c2f_depths = [1, 2, 2]
    with pytest.raises(RuntimeError, match="unsupported detector graph"):
        plan.build_detector_plan(model)


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "litepi_b200.h")).read()
    declared = set(re.findall(r"\b(lp_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert os.path.exists(L.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/litepi_b200.h but not exported"
    assert set(L.EXPORTS) <= declared
    assert lib.lp_abi_version() == L.ABI_VERSION


def test_struct_layout_matches_header():
    assert ctypes.sizeof(L.BufDesc) == 32
    assert ctypes.sizeof(L.OpDesc) == 17 * 4 + 2 * 4 + 4 + 3 * 8      # 17 int32, 2 float, pad, 3 int64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(v1_paths):
    h = ctypes.c_void_p()
    rc = L.lib().lp_create(ctypes.byref(h), 0)
    assert rc != 0 and b"no CPU fallback" in L.lib().lp_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        litepi_b200.B200Detector(*v1_paths)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        litepi_b200.Evaluator()


def test_classifier_arch_errors():
    with pytest.raises(ValueError, match="Unknown architecture"):          # e2e.py:335
        litepi_b200.B200Classifier(None, "vgg16")
    with pytest.raises(RuntimeError, match="no CPU fallback"):             # a supported arch still needs the GPU
        litepi_b200.B200Classifier(None, "resnet18")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo-litepi_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, f"{fn} mentions the oracle"
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        if fn.endswith((".cu", ".cuh")):
            assert "oracle" not in open(os.path.join(pkg, "csrc", fn)).read()


def test_records_roundtrip_and_metrics_fields():
    rec = np.zeros((3, 9), np.int32)
    f = rec.view(np.float32)
    rec[:, 0] = [0, 2, 2]
    f[:, 1:5] = [[1.9, 2.2, 30.7, 40.1], [5, 6, 7, 8], [9.5, 1.5, 20.5, 30.5]]
    f[:, 5] = [0.9, 0.8, 0.7]
    rec[:, 6] = 0
    rec[:, 7] = [4, 5, 6]
    f[:, 8] = [0.5, 0.6, 0.7]
    out = B200Pipeline.records_to_results(rec, 3)
    assert [len(o) for o in out] == [1, 0, 2]
    assert out[0][0]["bbox"] == (1, 2, 30, 40) and out[2][1]["cls_class"] == 6
    ref_fields = ["t_detection", "t_roi_extract", "t_classification", "t_postprocess", "t_total", "fps",
                  "num_detections", "det_confidence_avg", "cls_confidence_avg", "cpu_percent", "memory_mb",
                  "temperature", "precision", "recall", "f1", "level"]               # e2e.py:34-62
    assert list(PipelineMetrics.__dataclass_fields__) == ref_fields


def test_shard_indices():
    all_ids = sorted(sum((runner.shard_indices(4096, r, 8) for r in range(8)), []))
    assert all_ids == list(range(4096))
    assert runner.shard_indices(10, 3, 4) == [3, 7]
    with pytest.raises(ValueError):
        runner.shard_indices(4, 4, 4)


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = [3, 0][rank]                                    # ragged, one rank empty
    local = torch.arange(n * 9, dtype=torch.int32).reshape(n, 9) + 1000 * rank
    local[:, 0] = torch.tensor([4, 0, 2][:n], dtype=torch.int32) if n else local[:, 0]
    out = runner.gather_records(local)
    q.put((rank, runner.sort_records(out.numpy()).tolist()))
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]
    assert [r[0] for r in res[0]] == [0, 2, 4] and len(res[0]) == 3


def test_eval_host_metrics_equal_reference_golden():
    """The product's host half of the evaluation (per-class curves, AP, best-F1: litepi_b200.evaluate.metrics_from_stats)
    on a matching result computed by the oracle must reproduce the reference's recorded outputs bit for bit."""
    from helpers import load_eval_case
    from litepi_b200.evaluate import metrics_from_stats
    from oracle import eval_ref as ER
    for case in ("small", "mid", "wide"):
        preds, gts, nc, want = load_eval_case(case)
        cs, conf, pcs, tcs = [], [], [], []
        for p, g in zip(preds, gts):
            g = np.asarray(g, np.float64).reshape(-1, 5)
            tcs.append(g[:, 0])
            if not p:
                continue
            pb = np.array([q["bbox"] for q in p]); pc = np.array([q["cls_class"] for q in p])
            cs.append(ER.match_image_ref(pb, pc, g[:, 1:], g[:, 0]))
            conf.append(np.array([q["conf"] for q in p])); pcs.append(pc)
        got = metrics_from_stats(np.concatenate(cs), np.concatenate(conf), np.concatenate(pcs), np.concatenate(tcs), nc)
        for k, v in want.items():
            assert np.array_equal(np.asarray(got[k]), v), (case, k)


@pytest.mark.parametrize("arch", ["resnet18", "mobilenetv2", "efficientnet"])
def test_other_classifier_archs_plan_equals_torchvision(arch):
    """the layer plans of the reference's other --clf_arch choices (e2e.py:322-335), executed by the CPU
    plan interpreter with the CUDA executor's buffer / channel-padding / residual semantics, reproduce torchvision."""
    import torch
    from litepi_b200.cls_archs import PLAN_BUILDERS
    from oracle import pipeline_ref as PR
    ref = PR.build_classifier_ref(arch, 49, seed=3)
    plan = PLAN_BUILDERS[arch](ref.state_dict(), 64)
    assert plan.meta["num_classes"] == 49
    x = np.random.default_rng(5).integers(0, 256, (3, 64, 64, 3), dtype=np.uint8)
    _, logits = run_plan_cpu(plan, x)
    with torch.no_grad():
        want = ref(((torch.from_numpy(x.astype(np.float32)) / 255 - 0.18) / 0.34).permute(0, 3, 1, 2)).numpy()
    assert np.abs(logits.numpy() - want).max() < 2e-4 * max(1.0, np.abs(want).max())
    assert np.array_equal(logits.numpy().argmax(1), want.argmax(1))
    plan.layout(4)
    tc = plan.pack_tc_weights()
    assert tc.size > 0 and sum(1 for o in plan.ops if o["wtc_off"] >= 0) > 5      # most convs are tensor-core eligible
