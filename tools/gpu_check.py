"""Stage-by-stage GPU vs oracle report (development tool; the asserts live in tests/)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import synth, _lib as L
from litepi_b200.detector import FrameBatch
from oracle.ncnn_graph import DetectorOracle
from oracle import pipeline_ref as PR
from helpers import model_paths, oracle_pipeline_run
from plan_interp import run_plan_cpu, detect_tail_cpu

def section(name):
    print("\n=== " + name, flush=True)

def read_buf(net_obj, plan, bi, n):
    """decode workspace buffer bi -> float32 [n,h,w,c]"""
    b = plan.bufs[bi]; ws = net_obj.workspace
    if b["fmt"] == L.FMT_SPLIT16:
        per = b["image_bytes"] // 2
        hi = ws[b["offset"]:b["offset"] + n * b["image_bytes"]].view(torch.float16).float()
        lo_off = b["offset"] + net_obj.max_batch * b["image_bytes"]
        lo = ws[lo_off:lo_off + n * b["image_bytes"]].view(torch.float16).float()
        return (hi + lo).reshape(n, b["h"], b["w"], b["c"]).cpu()
    if b["fmt"] == L.FMT_F32:
        return ws[b["offset"]:b["offset"] + n * b["image_bytes"]].view(torch.float32).reshape(n, b["h"], b["w"], b["c"]).cpu()
    return None

def main():
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    param, binp = model_paths("vntsr")
    print("weights:", "trained v1" if binp else "RANDOM (staged weights missing)")
    det = litepi_b200.B200Detector(param, binp, max_batch=4, seed=0, tensor_cores=("--simt" not in sys.argv))
    print("tensor-core ops:", det.tc_ops)
    orc = DetectorOracle(param, binp, seed=0)
    if binp is None:
        ci = iter(det.model.convs)
        for Ly in orc.layers:
            if Ly.type == "Convolution":
                c = next(ci); Ly.weight, Ly.bias = c.weight, c.bias
    rng = np.random.default_rng(0)

    section("K1 letterbox vs oracle (bit-exact)")
    try:
        for (h, w) in [(681, 1198), (2048, 2048), (720, 1280), (480, 640), (640, 640), (333, 517), (100, 37), (1280, 1280)]:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            got, r, pad = det.letterbox(img)
            want, r2, pad2 = PR.letterbox_ref(img)
            print((h, w), "mismatch", int((got != want[:, :, ::-1]).sum()), r == r2, pad == tuple(pad2))
    except Exception: traceback.print_exc()

    section("K2/K3 detector forward vs oracle")
    try:
        frames = [synth.vn_frame(0), synth.vn_frame(1), synth.tt_frame(0)]
        lbs = np.stack([PR.letterbox_ref(f)[0][:, :, ::-1] for f in frames])
        got = det.forward(lbs)
        ref = orc.forward(torch.from_numpy(lbs.astype(np.float32) / 255).permute(0, 3, 1, 2).contiguous()).numpy()
        d = np.abs(got - ref)
        print("out0 max diff boxes %.3e scores %.3e  (tolerance 1e-2 / 1e-3)" % (d[:, :4].max(), d[:, 4].max()))
        cand = ref[:, 4] > 0.25
        print("candidates>0.25:", int(cand.sum()), " box diff at candidates %.3e" % (d[:, :4].transpose(0, 2, 1)[cand].max() if cand.any() else 0))
        if d[:, :4].max() > 1e-2 or d[:, 4].max() > 1e-3:
            bufs, _ = run_plan_cpu(det.plan, lbs)
            for bi, b in enumerate(det.plan.bufs):
                g = read_buf(det, det.plan, bi, lbs.shape[0])
                if g is None: continue
                e = (g - bufs[bi]).abs()
                print("  buf %2d %4dx%-4d c%-4d maxdiff %.3e (max |ref| %.2f)" % (bi, b["h"], b["w"], b["c"], float(e.max()), float(bufs[bi].abs().max())))
    except Exception: traceback.print_exc()

    section("K4+K5 decode/NMS vs oracle on identical out0 (bit-exact)")
    try:
        for fi, f in enumerate(frames):
            _, r, pad, _ = PR.preprocess_ref(f)
            for conf in (0.25, 0.001):
                wb, wsc, wc, (cb, cs, cc, widx) = PR.postprocess_ref(ref[fi], f.shape[:2], r, pad, conf, 0.45, return_candidates=True)
                gb, gs, gc = det.postprocess(ref[fi], f.shape[:2], r, pad, conf, 0.45)
                kidx = det.keep_idx[0, :len(gb)].cpu().numpy()
                ok = len(gb) == len(wb) and np.array_equal(gb, wb.astype(np.float32)) and np.array_equal(gs, wsc) and np.array_equal(kidx, widx)
                print("frame %d conf %.3f: cand %d (gpu %d) kept %d (gpu %d) bit-exact %s" % (fi, conf, len(cb), int(det.n_cand[0]), len(wb), len(gb), ok))
    except Exception: traceback.print_exc()

    section("K6 ROI select + PIL-exact resize vs oracle (bit-exact)")
    try:
        clf_ref = PR.build_shufflenet(49, seed=0)
        clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=clf_ref.state_dict(), max_batch=64)
        crops = synth.roi_crops(40, seed=1) + [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in [(64, 64), (10, 10), (200, 333), (64, 100), (130, 64), (500, 480), (3, 5), (1, 1)]]
        got = clf.preprocess_batch(crops).cpu().numpy()
        bad = 0
        for i, c in enumerate(crops):
            want, _ = PR.classifier_input_ref(c)
            bad += int((got[i] != want).sum())
        print("resize mismatching bytes over %d crops: %d" % (len(crops), bad))
        section("K7 ShuffleNetV2 logits vs torchvision fp32")
        lg = clf.logits_for(got)
        x = (torch.from_numpy(got.astype(np.float32)) / 255 - 0.18) / 0.34
        with torch.no_grad(): rl = clf_ref(x.permute(0, 3, 1, 2)).numpy()
        print("fused  logits max diff %.3e (tolerance 1e-2), top-1 agree %d/%d" % (np.abs(lg - rl).max(), int((lg.argmax(1) == rl.argmax(1)).sum()), len(lg)))
        clf.set_fused(False)
        lg2 = clf.logits_for(got)
        clf.set_fused(True)
        print("layered logits max diff %.3e, top-1 agree %d/%d" % (np.abs(lg2 - rl).max(), int((lg2.argmax(1) == rl.argmax(1)).sum()), len(lg2)))
        if np.abs(lg - rl).max() > 1e-3:
            bufs, _ = run_plan_cpu(clf.plan, got[:clf.max_batch])
            n = min(len(got), clf.max_batch); clf.logits_for(got[:n])
            for bi, b in enumerate(clf.plan.bufs):
                g = read_buf(clf, clf.plan, bi, n)
                if g is None: continue
                print("  buf %2d %3dx%-3d c%-4d maxdiff %.3e" % (bi, b["h"], b["w"], b["c"], float((g - bufs[bi][:n]).abs().max())))
    except Exception: traceback.print_exc()

    section("pipeline end-to-end vs oracle")
    try:
        pipe = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=49, max_batch=4,
                                        classifier_state_dict=clf_ref.state_dict(), seed=0)
        fr = [synth.vn_frame(i) for i in range(3)] + [synth.tt_frame(0)]
        got = pipe.run_batch(fr, 0.25, 0.45, 50)
        for f, g in zip(fr, got):
            want = oracle_pipeline_run(orc, clf_ref, f, 0.25, 0.45, 50)
            same = len(want) == len(g)
            bd = max([np.abs(a["box_f32"] - b["box_f32"]).max() for a, b in zip(g, want)], default=0)
            sd = max([abs(a["det_conf"] - b["det_conf"]) for a, b in zip(g, want)], default=0)
            cl = sum(a["cls_class"] == b["cls_class"] for a, b in zip(g, want))
            ints = sum(a["bbox"] == b["bbox"] for a, b in zip(g, want))
            print("frame %s: dets gpu %d oracle %d | box %.2e score %.2e | cls top-1 %d/%d | int bbox %d/%d" % (f.shape, len(g), len(want), bd, sd, cl, len(want), ints, len(want)))
        res, m = pipe.run(fr[0], 0.25, 0.45, 50)
        print("run(): %d results, t_det %.2f ms t_roi %.2f t_cls %.2f total %.2f" % (len(res), m.t_detection, m.t_roi_extract, m.t_classification, m.t_total))
    except Exception: traceback.print_exc()

    section("timing, batch 64 VN frames (device resident)")
    try:
        B = 64
        pipe = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=49, max_batch=B,
                                        classifier_state_dict=clf_ref.state_dict(), seed=0)
        fr = [synth.vn_frame(i) for i in range(B)]
        fb = FrameBatch.from_host(fr, pipe.device)
        d = pipe.detector
        def timed(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); e1.synchronize()
            return e0.elapsed_time(e1) / n
        t_lb = timed(lambda: d.letterbox_device(fb))
        lb = d.letterbox_device(fb)
        t_fw = timed(lambda: d.forward_device(lb))
        out0 = d.forward_device(lb)
        t_nms = timed(lambda: d.decode_nms_device(out0, fb.h[:B], fb.w[:B], d.ratio[:B], d.pad[:2 * B], 0.25, 0.45))
        t_all = timed(lambda: pipe.run_device(fb, 0.25, 0.45, 50))
        n = pipe.run_device(fb, 0.25, 0.45, 50)
        print("letterbox %.3f ms | detector %.3f ms | decode+nms %.3f ms | whole run_device %.3f ms (%d rois) -> %.0f frames/s"
              % (t_lb, t_fw, t_nms, t_all, n, B / t_all * 1e3))
        print("detector %.2f TFLOP/s algorithmic" % (2 * sum(d.plan.macs) * B / (t_fw * 1e-3) / 1e12))
    except Exception: traceback.print_exc()

if __name__ == "__main__":
    main()
