#!/usr/bin/env python3
"""Headline benchmark: E2E frames/s (detect + NMS + classify) of the YOLO-LitePi hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the whole hot path over one batch of synthetic frames
(BASELINE.json configs[1]: 64 VN-Signs-shape 1198x681 frames, detector at 640, conf 0.25,
IoU 0.45, min_area 50, ShuffleNetV2 x1.0 with 49 classes).  Prints ONE JSON line:

  value      frames/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the public API with HOST frames: pinned H2D copy of every
             batch and D2H of the detection records inside the timed region
  roofline   the dominant kernel (largest Detect-head 3x3 conv), timed live with CUDA events
  cpu_baseline  the CPU oracle port of the reference path on this host's cores (bounded sample)

`--impl reference` times the reference's own CPU path (oracle port; OpenCV-DNN on the
reference's yolo_plus.onnx when it was staged) instead.  Multi-GPU: one process per GPU
(torchrun), frames sharded across ranks, NCCL only for the final detection gather.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

BATCH = 64
CONF, IOU, MIN_AREA, NUM_CLASSES = 0.25, 0.45, 50, 49
METRIC = "E2E frames/sec (detect+NMS+classify)"
WORKLOAD = "configs[1]: full two-stage pipeline, batch 64 VN-Signs-shape 1198x681 frames, YOLO-LitePi v1 @640 + ShuffleNetV2 x1.0 (49 cls)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark(self):
        """call at the start of the timed region: only samples after this point are reported"""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.lines = self.lines[getattr(self, "first", 0):]
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(n_frames: int, warm: int = 3):
    """The reference path on host cores: oracle port of e2e.py HybridPipeline.run per frame
    (cv2 letterbox, numpy postprocess/NMS/ROI as in e2e.py, PIL resize, torchvision ShuffleNetV2 CPU); the detector
    forward runs OpenCV-DNN on the reference's yolo_plus.onnx when staged, else the torch fp32 graph oracle."""
    import torch
    import cv2
    from litepi_b200 import synth
    from oracle import pipeline_ref as PR
    from oracle.ncnn_graph import DetectorOracle
    from helpers import model_paths, onnx_path
    param, binp = model_paths("vntsr")
    onnx = onnx_path()
    net = cv2.dnn.readNetFromONNX(onnx) if onnx else None
    orc = None if net is not None else DetectorOracle(param, binp, seed=0)
    clf = PR.build_shufflenet(NUM_CLASSES, seed=0)
    frames = [synth.vn_frame(i) for i in range(max(n_frames, 1))]

    def one(f):
        x, r, pad, _ = PR.preprocess_lib(f)
        if net is not None:
            net.setInput(x)
            out0 = net.forward()[0]
        else:
            out0 = orc.forward(x)[0].numpy()
        boxes, scores, classes = PR.postprocess_ref(out0, f.shape[:2], r, pad, CONF, IOU)
        rois, valid = PR.roi_select_ref(boxes, f.shape[:2], MIN_AREA)
        crops = [f[y1:y2, x1:x2] for (x1, y1, x2, y2) in rois]
        for i in range(0, len(crops), 8):                   # reference batch_size 8 (e2e.py:413)
            PR.classify_lib(clf, crops[i:i + 8])
        return len(valid)

    for f in frames[:warm]:
        one(f)
    lat = []
    t0 = time.perf_counter()
    for f in frames:
        t = time.perf_counter()
        one(f)
        lat.append((time.perf_counter() - t) * 1e3)
    dt = time.perf_counter() - t0
    cores = max(torch.get_num_threads(), cv2.getNumThreads())
    runtime = "OpenCV-DNN(yolo_plus.onnx)" if net is not None else "torch-fp32 graph oracle"
    return {"fps": len(frames) / dt, "p50_ms": statistics.median(lat), "cores": cores, "n": len(frames),
            "runtime": runtime, "host_cpus": os.cpu_count()}


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    n = 32
    t0 = time.perf_counter()
    r = None
    for _ in range(args.warmup):
        r = cpu_reference_run(8, warm=1)
    vals = []
    for _ in range(args.steps):
        r = cpu_reference_run(n, warm=0)
        vals.append(r["fps"])
    fps = len(vals) * n / sum(n / v for v in vals)
    sample = f"{n} VN-shape frames per step through the oracle port ({r['runtime']} + torchvision ShuffleNetV2 CPU)"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / fps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": n, "conf": CONF, "iou": IOU, "min_area": MIN_AREA},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "p50_latency_ms": r["p50_ms"], "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--lanes", type=int, default=2, help="batches in flight per GPU (pipeline instances on their own CUDA streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true",
                    help="device-resident steps only (no e2e / latency / CPU legs): the command ncu wraps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import litepi_b200
    from litepi_b200 import synth, _lib as L
    from litepi_b200.detector import FrameBatch
    from litepi_b200.runner import gather_records
    from oracle import pipeline_ref as PR          # only for the classifier's seeded state_dict + cpu_baseline leg
    from helpers import model_paths

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]                # these levels print NCCL's version banner on stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    B = args.batch
    param, binp = model_paths("vntsr")
    clf_sd = PR.build_shufflenet(NUM_CLASSES, seed=0).state_dict()
    # Two pipeline instances on two CUDA streams: consecutive batches are independent, so batch k+1 starts while the
    # tail of batch k (partially filled last waves of its persistent kernels) still runs.  Each owns its workspace.
    n_lanes = max(1, args.lanes)
    pipes = [litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=NUM_CLASSES, device=local_rank,
                                      max_batch=B, classifier_state_dict=clf_sd, seed=0) for _ in range(n_lanes)]
    lanes = [torch.cuda.Stream(device=torch.device("cuda", local_rank)) for _ in range(n_lanes)]
    pipe = pipes[0]
    # rank r owns frames i with i % world == r  (frame ids are global)
    ids = [rank + world * i for i in range(B)]
    frames = np.stack([synth.vn_frame(i) for i in ids])
    host = torch.from_numpy(frames).pin_memory()
    dev_frames = [torch.empty_like(host, device=dev) for _ in range(n_lanes)]
    for d in dev_frames:
        d.copy_(host)
    frame_ids = torch.tensor(ids, dtype=torch.int32, device=dev)
    fbs = [FrameBatch.from_device(d) for d in dev_frames]
    fb0 = fbs[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the dominant kernel: the largest conv of the plan (model.22.cv2.0.1, 64->64 3x3 on 80x80)
    macs = pipe.detector.plan.macs
    dom = int(np.argmax(macs))
    dom_flops = 2.0 * macs[dom] * B

    # ---------------- device-resident throughput (value)
    sampler = ClockSampler(local_rank)
    sampler.start()                                     # nvidia-smi needs a few hundred ms to produce its first line
    for _ in range(args.warmup):
        for ln in range(n_lanes):
            pipes[ln].run_device(fbs[ln], CONF, IOU, MIN_AREA, frame_ids)
    barrier()
    t_spin = time.time()
    while len(sampler.lines) == 0 and time.time() - t_spin < 3.0:
        time.sleep(0.05)
    # the wait for the sampler's first line left the GPU idle (clocks drop): one more untimed step per lane right before
    # the timed region, so that a short run (small --steps) is not dominated by the clock ramp
    for ln in range(n_lanes):
        pipes[ln].run_device(fbs[ln], CONF, IOU, MIN_AREA, frame_ids)
    barrier()
    sampler.mark()
    launches0 = sum(p_.counters.launch_count() for p_ in pipes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_rois = 0
    rec_ring = torch.empty((args.steps,) + tuple(pipe.records.shape), dtype=torch.int32, device=dev)
    cnt_ring = torch.zeros((args.steps,), dtype=torch.int32, device=dev)
    barrier()
    main = torch.cuda.current_stream()
    e0.record(main)
    for st_ in lanes:
        st_.wait_event(e0)
    # steps are enqueued back to back, alternating between the lanes: the ROI count stays on the device
    # (lp_set_roi_count_device), so a step needs no host round trip; each step's records and count are kept on the
    # device for the final gather
    for k in range(args.steps):
        ln = k % n_lanes
        with torch.cuda.stream(lanes[ln]):
            pipes[ln].enqueue_device(fbs[ln], CONF, IOU, MIN_AREA, frame_ids)
            rec_ring[k].copy_(pipes[ln].records, non_blocking=True)      # preallocated: no allocator call (= no implicit sync) in the timed region
            cnt_ring[k:k + 1].copy_(pipes[ln].n_rois, non_blocking=True)
    for st_ in lanes:
        main.wait_stream(st_)
    for ln in range(n_lanes):
        with torch.cuda.stream(lanes[ln]):
            pipes[ln].finish(fbs[ln], frame_ids)         # waits; validates capacities of the lane's last step
    counts_h = cnt_ring.cpu().tolist()
    if counts_h and max(counts_h) > pipe.max_rois:
        raise RuntimeError("bench: a step exceeded max_rois")
    n_rois = sum(counts_h)
    local = (torch.cat([rec_ring[k, :c] for k, c in enumerate(counts_h)]) if counts_h
             else torch.zeros((0, 9), dtype=torch.int32, device=dev))
    gathered = gather_records(local)                    # the one collective: final detection gather (NCCL)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(p_.counters.launch_count() for p_ in pipes) - launches0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * B * args.steps / (ms * 1e-3)

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "value": value, "ms_per_step": ms / args.steps,
                              "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel: a single-lane pass of the same steps with CUDA events around every launch of
    # that kernel on its stream (with several batches in flight the kernels of different batches share the SMs, which
    # stretches any one launch and says nothing about the kernel)
    pipe.ctx.probe_set(L.NET_DETECTOR, dom)
    for _ in range(args.steps):
        pipe.enqueue_device(fb0, CONF, IOU, MIN_AREA, frame_ids)
    pipe.finish(fb0, frame_ids)
    probe = pipe.ctx.probe_read()
    pipe.ctx.probe_set(L.NET_DETECTOR, -1)

    # ---------------- per-stage figures (SURVEY.md 8d "roofline per stage"): each stage alone, single lane, CUDA events
    def timed(fn, n=20):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / n
    det_, clf_ = pipe.detector, pipe.classifier
    t_lb = timed(lambda: det_.letterbox_device(fb0))
    lb_ = det_.letterbox_device(fb0)
    t_fw = timed(lambda: det_.forward_device(lb_))
    out0_ = det_.forward_device(lb_)
    t_nms = timed(lambda: det_.decode_nms_device(out0_, fb0.h[:B], fb0.w[:B], det_.ratio[:B], det_.pad[:2 * B], CONF, IOU))
    n_r = pipe.run_device(fb0, CONF, IOU, MIN_AREA, frame_ids)
    t_rs = timed(lambda: clf_.resize_device(fb0, pipe.roi_xyxy, pipe.roi_src, n_r)) if n_r else 0.0
    cls_in_ = clf_.resize_device(fb0, pipe.roi_xyxy, pipe.roi_src, n_r) if n_r else None
    t_cl = timed(lambda: clf_.classify_device(cls_in_)) if n_r else 0.0
    src_bytes = int(sum(int(h) * int(w) * 3 for h, w in zip(fb0.h[:B], fb0.w[:B])))
    rx = pipe.roi_xyxy[:n_r].cpu().numpy().astype(np.int64) if n_r else np.zeros((0, 4), np.int64)
    roi_px = int(((rx[:, 2] - rx[:, 0]) * (rx[:, 3] - rx[:, 1])).sum())
    stage_rows = [
        ("K1 letterbox", t_lb, "hbm", (src_bytes + B * 640 * 640 * 3) / 1e9, "GB"),
        ("K2 detector convs", t_fw, "tensor", 2.0 * sum(macs) * B / 1e12, "TFLOP"),
        ("K4+K5 decode+NMS", t_nms, "hbm", B * (8400 * 65 * 4) / 1e9, "GB"),
        ("K6 ROI resize", t_rs, "hbm", (roi_px * 3 + n_r * 64 * 64 * 3) / 1e9, "GB"),
        ("K7 ShuffleNetV2", t_cl, "tensor", n_r * 23.59e6 / 1e12, "TFLOP"),
    ]
    # ---------------- end-to-end with host frames (e2e): pinned H2D + hot path + D2H of records
    # One copy stream keeps the PCIe link busy back to back (2 device frame buffers per lane); lane s % n_lanes runs
    # step s on its own stream once its frames have landed, then queues the D2H of the records; the host reads step
    # s - n_lanes back while step s runs.
    copy_stream = torch.cuda.Stream(device=dev)
    n_buf = 2 * n_lanes
    e2e_frames = dev_frames + [torch.empty_like(host, device=dev) for _ in range(n_buf - n_lanes)]
    e2e_fbs = [FrameBatch.from_device(d) for d in e2e_frames]
    ready = [torch.cuda.Event() for _ in range(n_buf)]
    consumed = [torch.cuda.Event() for _ in range(n_buf)]
    n_out = [0]

    def e2e_loop(steps):
        d2h = 0
        ahead = 0                                            # copies issued so far

        def issue_copy(i):
            b = i % n_buf
            with torch.cuda.stream(copy_stream):
                if i >= n_buf:
                    copy_stream.wait_event(consumed[b])     # the step that read this buffer last is done with it
                e2e_frames[b].copy_(host, non_blocking=True)
                ready[b].record(copy_stream)

        while ahead < min(n_buf, steps):
            issue_copy(ahead); ahead += 1
        for s_ in range(steps):
            ln, b = s_ % n_lanes, s_ % n_buf
            if s_ >= n_lanes:
                n_out[0] += pipes[ln].collect(0).shape[0]
            with torch.cuda.stream(lanes[ln]):
                lanes[ln].wait_event(ready[b])
                pipes[ln].enqueue_device(e2e_fbs[b], CONF, IOU, MIN_AREA, frame_ids, slot=0)
                consumed[b].record(lanes[ln])
                pipes[ln].enqueue_fetch(0)               # D2H of the step's records, queued behind the step
            if ahead < steps:
                issue_copy(ahead); ahead += 1
            d2h += pipe.records.numel() * 4 + 4 + 4 * B
        for ln in range(min(n_lanes, steps)):
            n_out[0] += pipes[ln].collect(0).shape[0]
        return d2h

    e2e_loop(args.warmup)
    barrier()
    e0.record()
    d2h_total = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e = float(t)
    e2e_value = world * B * args.steps / (ms_e * 1e-3)

    # ---------------- p50 single-frame latency through the public API (host frame in, results out)
    lat = []
    for i in range(30):
        t0 = time.perf_counter()
        pipe.run_batch([frames[i % B]], CONF, IOU, MIN_AREA)
        lat.append((time.perf_counter() - t0) * 1e3)
    p50 = statistics.median(lat[5:])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_src = peaks()
    dom_ms = statistics.mean(probe) if probe else None
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms else None
    peak = pk["bf16_tflops_sustained"]
    traffic = None                                   # measured once with ncu --set full for this kernel at this workload
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("kernel") == pipe.detector.plan.names[dom] and B == 64:
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16x2-split operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "conf": CONF, "iou": IOU, "min_area": MIN_AREA,
                   "weights": "reference trained v1 (model.ncnn.bin)" if binp else "random-init (weights not staged)",
                   "classifier_weights": "random-init seed 0 (reference ships none)",
                   "l2": "per-step inputs (157 MB frames) + 2.2 GB activation workspace exceed the 126 MB L2",
                   "rois_per_step": n_rois / max(args.steps, 1), "parallelism": f"frames sharded over {world} GPU(s); {n_lanes} batches in flight per GPU (CUDA streams)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(host.numel()),
                "d2h_bytes_per_step": int(d2h_total / max(args.steps, 1)), "ms_per_step": ms_e / args.steps},
        "gpu_launches": int(launches),
        "p50_latency_ms": p50,
        "roofline": {"bound": "tensor", "kernel": pipe.detector.plan.names[dom], "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_unit": "bytes per launch (ncu dram read+write, profiles/dominant_kernel_traffic.json)",
                     "peak_source": pk_src + " bf16_tflops_sustained", "launch_ms": dom_ms,
                     "measured_in": "single-lane pass of the same steps, CUDA events around each launch on its stream",
                     "flops_per_launch": dom_flops},
    }
    line["stages"] = [{"stage": nm, "ms": round(t, 4), "bound": bd,
                       "achieved": (work / (t * 1e-3)) if t else None, "unit": "GB/s" if unit == "GB" else "TFLOP/s",
                       "peak": pk["hbm_gbs"] if bd == "hbm" else peak,
                       "frac": ((work / (t * 1e-3)) / (pk["hbm_gbs"] if bd == "hbm" else peak)) if t else None}
                      for nm, t, bd, work, unit in stage_rows]
    if not args.no_cpu_baseline:
        r = cpu_reference_run(48)
        line["cpu_baseline"] = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                                "sample": f"{r['n']} frames of the same workload, one at a time as e2e.py does, "
                                          f"{r['runtime']} + numpy post-processing + torchvision ShuffleNetV2 CPU; "
                                          f"p50 {r['p50_ms']:.1f} ms/frame"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
