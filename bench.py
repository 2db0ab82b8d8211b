#!/usr/bin/env python3
"""Headline benchmark: E2E frames/s (detect + NMS + classify) of the YOLO-LitePi hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {0,1,2,3,4}]

Default = BASELINE.json configs[1]: one step = one pass of the whole hot path over one batch of 64 synthetic
VN-Signs-shape 1198x681 frames (detector at 640, conf 0.25, IoU 0.45, min_area 50, ShuffleNetV2 x1.0, 49 classes).
Prints ONE JSON line:

  value        frames/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the public host API (`B200Pipeline.stream().run_stream`): frames in pinned
               HOST memory, H2D copy of every batch and D2H of the detection records inside the timed region
  roofline     the dominant kernel (largest Detect-head 3x3 conv) timed live with CUDA events, plus the same
               figure over ALL launches of that kernel family (`kernel_all_launches`)
  cpu_baseline the reference's own CPU path on this host's cores (bounded sample)

Other configs (BASELINE.json configs[0..4]; one committed line each under profiles/):
  0  detector + NMS on ONE 1198x681 frame            2  TT100K-shape 2048x2048 frames, batch 32, 91 classes
  3  ShuffleNetV2 alone on 1024 crops (unit: crops)  4  4096 distinct frames sharded i % W over the ranks + NCCL gather

`--impl reference` times the reference's own CPU implementation: the UNMODIFIED e2e.py `HybridPipeline.run`
(oracle/ref_runtime.py; OpenCV-DNN executes the reference's yolo_plus.onnx in place of the uninstallable ncnn).
Multi-GPU: one process per GPU (torchrun), frames sharded across ranks, NCCL only for the final detection gather.
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

try:
    _ORIG_AFFINITY = os.sched_getaffinity(0)
except Exception:
    _ORIG_AFFINITY = None
CONF, IOU, MIN_AREA = 0.25, 0.45, 50
METRIC = "E2E frames/sec (detect+NMS+classify)"
CONFIGS = {
    0: dict(name="configs[0]: YOLO-LitePi v1 detector + NMS, one synthetic 1198x681 frame at 640 input",
            batch=1, classes=49, shape="vn", unit="frames/s"),
    1: dict(name="configs[1]: full two-stage pipeline, batch 64 VN-Signs-shape 1198x681 frames, YOLO-LitePi v1 @640 + "
                 "ShuffleNetV2 x1.0 (49 cls)", batch=64, classes=49, shape="vn", unit="frames/s"),
    2: dict(name="configs[2]: TT100K-shape 2048x2048 frames with dense small signs, batch 32, batched classification "
                 "(91 cls)", batch=32, classes=91, shape="tt", unit="frames/s"),
    3: dict(name="configs[3]: ShuffleNetV2 x1.0 alone on 1024 synthetic 64x64 ROI crops (49 cls)",
            batch=1024, classes=49, shape="crops", unit="crops/s"),
    4: dict(name="configs[4]: frame-sharded sweep, 4096 distinct VN-shape frames (frame i on rank i % W), batch 64, "
                 "NCCL gather of the detection records", batch=64, classes=49, shape="vn", unit="frames/s"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout=3.0):
        t0 = time.time()
        while self.proc is not None and len(self.lines) == 0 and time.time() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        """call at the start of the timed region: only samples after this point are reported"""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        lines = self.lines[self.first:]
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
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


# ======================================================================================= frames
def make_frames(cfg_id: int, ids):
    from litepi_b200 import synth
    if CONFIGS[cfg_id]["shape"] == "tt":
        return np.stack([synth.tt_frame(i) for i in ids])
    if cfg_id == 4:
        # 4096 DISTINCT frames from 256 generated scenes: frame i = scene (i % 256) rolled horizontally by 37 * (i // 256)
        # pixels (generating 4096 scenes costs a minute of host time per rank and says nothing about the path)
        base = {}
        out = []
        for i in ids:
            s = i % 256
            if s not in base:
                base[s] = synth.vn_frame(s)
            out.append(np.roll(base[s], 37 * (i // 256), axis=1) if i >= 256 else base[s])
        return np.stack(out)
    return np.stack([synth.vn_frame(i) for i in ids])


# ======================================================================================= CPU arm
class CpuArm:
    """The reference's CPU implementation on the host cores.  kind "reference": the UNMODIFIED e2e.py classes
    (HybridPipeline.run, e2e.py:443-531) through oracle/ref_runtime.py; kind "port": the oracle port, only when the
    reference source is neither mounted nor staged."""

    def __init__(self, cfg_id: int, state_dict=None, threads=None):
        import torch
        import cv2
        from helpers import model_paths, onnx_path
        from oracle import ref_runtime as RR
        self.cfg_id, self.cfg = cfg_id, CONFIGS[cfg_id]
        self.host_cpus = os.cpu_count() or 1
        # "all the host threads it can use" = the runtimes' own defaults for this process (cgroup / affinity aware)
        self.threads = int(threads or max(torch.get_num_threads(), cv2.getNumThreads(), 1))
        param, binp = model_paths("vntsr")
        if state_dict is None:
            from litepi_b200.classifier import _random_state_dict
            state_dict = _random_state_dict(self.cfg["classes"], 0)
        self.kind = "reference" if RR.e2e_source_path() else "port"
        if self.kind == "reference":
            self.rp = RR.ReferencePipeline(param, binp, self.cfg["classes"], state_dict, threads=self.threads, onnx=onnx_path())
            self.runtime = self.rp.runtime
        else:
            from oracle import pipeline_ref as PR
            from oracle.ncnn_graph import DetectorOracle
            self.PR = PR
            onnx = onnx_path()
            self.net = cv2.dnn.readNetFromONNX(onnx) if onnx else None
            self.orc = None if self.net is not None else DetectorOracle(param, binp, seed=0)
            self.clf = PR.build_shufflenet(self.cfg["classes"], seed=0)
            self.clf.load_state_dict(state_dict)
            self.runtime = "OpenCV-DNN(yolo_plus.onnx)" if self.net is not None else "torch-fp32 graph oracle"
            torch.set_num_threads(self.threads); cv2.setNumThreads(self.threads)

    def set_threads(self, n: int):
        import torch
        import cv2
        self.threads = int(n)
        if self.kind == "reference":
            self.rp.set_threads(n)
        else:
            torch.set_num_threads(n); cv2.setNumThreads(n)

    def frame(self, f) -> int:
        """one frame through the pipeline the config names; returns the number of results"""
        if self.kind == "reference":
            if self.cfg_id == 0:                                   # detector + NMS only (NCNNDetector.detect, e2e.py:298)
                return len(self.rp.pipe.detector.detect(f, CONF, IOU)[0])
            return len(self.rp.run(f, CONF, IOU, MIN_AREA)[0])
        PR = self.PR
        x, r, pad, _ = PR.preprocess_lib(f)
        if self.net is not None:
            self.net.setInput(x)
            out0 = self.net.forward()[0]
        else:
            out0 = self.orc.forward(x)[0].numpy()
        boxes, scores, classes = PR.postprocess_ref(out0, f.shape[:2], r, pad, CONF, IOU)
        if self.cfg_id == 0:
            return len(boxes)
        rois, valid = PR.roi_select_ref(boxes, f.shape[:2], MIN_AREA)
        crops = [f[y1:y2, x1:x2] for (x1, y1, x2, y2) in rois]
        for i in range(0, len(crops), 8):                          # reference batch_size 8 (e2e.py:413)
            PR.classify_lib(self.clf, crops[i:i + 8])
        return len(valid)

    def crops(self, crops) -> int:
        """config 3: the reference classifier wrapper on a list of crops, batch 64 (evaluation-tsr.ipynb:407)"""
        if self.kind == "reference":
            for i in range(0, len(crops), 64):
                self.rp.pipe.classifier.predict_batch(crops[i:i + 64])
        else:
            for i in range(0, len(crops), 64):
                self.PR.classify_lib(self.clf, crops[i:i + 64])
        return len(crops)

    def run(self, items, warm: int = 0):
        """time `items` (frames, or ONE list of crops for config 3) one at a time as e2e.py does; returns units/s, p50 ms"""
        fn = self.crops if self.cfg_id == 3 else self.frame
        for it in items[:warm]:
            fn(it)
        lat, units = [], 0
        t0 = time.perf_counter()
        for it in items:
            t = time.perf_counter()
            r = fn(it)
            units += r if self.cfg_id == 3 else 1
            lat.append((time.perf_counter() - t) * 1e3)
        dt = time.perf_counter() - t0
        return units / dt, statistics.median(lat)

    def describe(self, n_units: int, p50=None) -> str:
        what = {0: "NCNNDetector.detect (letterbox + forward + postprocess/NMS)", 3: "PyTorchClassifier.predict_batch, batch 64"}.get(
            self.cfg_id, "HybridPipeline.run, one frame at a time as e2e.py:1108-1116 does")
        src = "the reference's unmodified e2e.py" if self.kind == "reference" else "oracle port of e2e.py"
        s = f"{n_units} {self.cfg['unit'].split('/')[0]} of the same workload through {src}: {what}; detector runtime {self.runtime}, " \
            f"torchvision ShuffleNetV2 on CPU"
        return s + (f"; p50 {p50:.1f} ms" if p50 is not None else "")


def cpu_items(cfg_id: int, n: int):
    if cfg_id == 3:
        from litepi_b200 import synth
        return [synth.roi_crops(n, seed=7)]
    return list(make_frames(cfg_id, list(range(n))))


def run_reference_arm(args, rank: int):
    """`--impl reference`: each step = the reference CPU path over a bounded sample of the config's batch."""
    if rank != 0:
        return
    cfg_id, cfg = args.config, CONFIGS[args.config]
    t_wall = time.perf_counter()
    arm = CpuArm(cfg_id)
    arm.threads_all = arm.threads
    per_step = {0: 16, 1: 64, 2: 16, 3: 1024, 4: 64}[cfg_id]      # config 1/3/4: the GPU arm's own batch; 0/2: bounded
    items = cpu_items(cfg_id, per_step if cfg_id != 3 else per_step)
    arm.run(items[:5] if cfg_id != 3 else items, warm=0)            # 5 warm frames (evaluation_tsd_single_img.ipynb cell [5])
    for _ in range(args.warmup):
        arm.run(items[:8] if cfg_id != 3 else items)
    vals, p50s = [], []
    for _ in range(args.steps):
        v, p = arm.run(items)
        vals.append(v); p50s.append(p)
    fps = len(vals) / sum(1.0 / v for v in vals)
    # the reference pins its runtimes to 4 threads (net.opt.num_threads e2e.py:211, run.bash:31-39): same sample once
    arm.set_threads(4)
    arm.run(items[:3] if cfg_id != 3 else items)
    fps4, _ = arm.run(items[:16] if cfg_id != 3 else items)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": cfg["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step / fps, "higher_is_better": True,
            "scaling": "strong" if cfg_id == 4 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": gpu_config_dict(cfg_id, cfg["batch"]),
            "cpu_baseline": {"value": fps, "unit": cfg["unit"], "cores": arm.threads_all, "kind": arm.kind,
                             "sample": arm.describe(per_step) + f"; {per_step} per step", "host_cpus": arm.host_cpus,
                             "value_4_threads": fps4},
            "e2e": {"value": fps, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "p50_latency_ms": statistics.median(p50s), "wall_s": time.perf_counter() - t_wall}
    print(json.dumps(line), flush=True)


def gpu_config_dict(cfg_id: int, batch: int):
    """The `config` object BOTH arms print, identical for a given --config (the driver compares them); everything
    arm-specific goes to `config_detail`."""
    per_frame = {"vn": 1198 * 681 * 3, "tt": 2048 * 2048 * 3, "crops": 64 * 64 * 3}[CONFIGS[cfg_id]["shape"]]
    l2 = (f"inputs of one step = {batch * per_frame / 1e6:.1f} MB of frames" +
          (" (+ GB-sized activation workspaces): larger than the 126 MB L2, no flush needed" if batch * per_frame > 126e6 else
           ": smaller than the 126 MB L2 -- this is a latency configuration; the step's activations (>= 35 MB/frame) are rewritten every step"))
    return {"workload": CONFIGS[cfg_id]["name"], "batch": batch, "conf": CONF, "iou": IOU, "min_area": MIN_AREA, "l2": l2}


# ======================================================================================= GPU arm helpers
def pin_to_gpu_numa(local_rank: int):
    """Bind this rank to the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated (first touch puts the
    staging buffers next to the GPU's PCIe root).  Returns a description or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/local_cpulist") as f:
            cl = f.read().strip()
        node = None
        if os.path.exists(base + "/numa_node"):
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        cpus = set()
        for part in cl.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "cpus": len(cpus) if cpus else len(allowed), "bound": bool(cpus and cpus != allowed)}
    except Exception as e:                                   # containers without sysfs: run unbound, say so
        return {"bound": False, "error": f"{type(e).__name__}: {e}"}


def h2d_ceiling(host_t, dev_t, stream, n=8):
    """GB/s of back-to-back cudaMemcpyAsync of this step's input buffer, pinned host -> device (the PCIe bound of e2e)."""
    import torch
    with torch.cuda.stream(stream):
        dev_t.copy_(host_t, non_blocking=True)
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            dev_t.copy_(host_t, non_blocking=True)
        b.record(stream)
        b.synchronize()
    return host_t.numel() * n / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cpu-leg"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--detector", default="v1", choices=["v1", "v2"],
                    help="v1 = the reference's trained VN-Signs export; v2 = the paper's YOLO-LitePi widths (TT100K export, random-init: weights are not in the reference repo)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=3, help="batches in flight per GPU (pipeline instances on their own CUDA streams)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="direct launches instead of the captured CUDA graph")
    ap.add_argument("--profile-mode", action="store_true",
                    help="device-resident steps only, direct launches (no e2e / latency / CPU legs): the command ncu wraps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.impl == "cpu-leg":
        cpu_leg_main(args.config)
        return
    if args.config == 0:
        return bench_detector_single(args)
    if args.config == 3:
        return bench_classifier_alone(args)
    return bench_pipeline(args)


# ======================================================================================= configs 1, 2, 4
def bench_pipeline(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback")
    affinity = pin_to_gpu_numa(local_rank)                  # before the first pinned allocation
    import litepi_b200
    from litepi_b200 import _lib as L
    from litepi_b200.runner import gather_records
    from helpers import model_paths

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]                # these levels print NCCL's version banner on stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    cfg_id, cfg = args.config, CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    NC = cfg["classes"]
    param, binp = model_paths("vntsr" if args.detector == "v1" else "tt100k")
    # random-init weights (the v2 / paper-width export ships without weights) score every anchor around 0.5: at conf 0.25 that
    # is thousands of meaningless "detections" per frame.  The v2 line measures the ARCHITECTURE's throughput: a threshold no
    # random score passes, so the ROI stages see an empty list (said in config_detail).
    conf_t = CONF if binp else 0.999
    n_lanes = max(1, args.lanes)
    pipe = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=NC, device=local_rank, max_batch=B, seed=0)
    sr = pipe.stream(lanes=n_lanes, use_graph=False if (args.no_graph or args.profile_mode) else None)
    pipes = sr.pipes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- frames.  configs 1/2: rank r owns frames i with i % world == r (ids are global), one batch re-used every step.
    # config 4: all of the rank's share of the 4096 distinct frames, in pinned host memory.
    if cfg_id == 4:
        n_total = 4096
        ids_all = list(range(rank, n_total, world))
        local = make_frames(4, ids_all)
        n_local_steps = (len(ids_all) + B - 1) // B
        host_all = torch.from_numpy(local).pin_memory()
        del local
        frames = host_all[:B].numpy()
    else:
        ids = [rank + world * i for i in range(B)]
        frames = make_frames(cfg_id, ids)
    H, W = int(frames.shape[1]), int(frames.shape[2])
    for b in range(sr.n_buf):
        sr.host_buffer(b, H, W)[:] = frames                 # the pinned ring holds this rank's batch (config 4: its first)
    ids_step = [rank + world * i for i in range(B)]

    def ring_batches(n):
        for s in range(n):
            yield sr.host_np[s % sr.n_buf]                  # produced in place: no host copy, H2D straight from the ring

    macs = pipe.detector.plan.macs
    dom = int(np.argmax(macs))
    dom_flops = 2.0 * macs[dom] * B

    # ---------------- warm-up through the public API (captures the CUDA graphs)
    sampler = ClockSampler(local_rank)
    sampler.start()
    n_warm = max(args.warmup, sr.n_buf)                     # every ring buffer has its own graph
    for _ in sr.run_stream(ring_batches(n_warm), conf_t, IOU, MIN_AREA, frame_ids=[ids_step] * n_warm):
        pass
    barrier()
    sampler.wait_first()
    sr.replay_resident(n_lanes, conf_t, IOU, MIN_AREA); sr.drain_resident(n_lanes)   # the wait above left the GPU idle: ramp the clocks again
    barrier()

    # ---------------- device-resident throughput (value): K steps replayed on the frames resident in the device ring
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_st = torch.cuda.current_stream()
    steps = args.steps
    if cfg_id == 4:
        # the rank's share, resident in HBM; every step stages its own 64 frames device-to-device into the ring
        dev_all = host_all.to(dev)
        steps = n_local_steps
    sampler.mark()
    launches0 = sr.launch_count()
    barrier()
    e0.record(main_st)
    for st_ in sr.lane_streams + [sr.copy_stream]:
        st_.wait_event(e0)
    if cfg_id == 4:
        recs4 = list(sr.run_stream((dev_all[i * B:(i + 1) * B] for i in range(steps)), conf_t, IOU, MIN_AREA,
                                   frame_ids=[ids_all[i * B:(i + 1) * B] for i in range(steps)]))
        local_rec = torch.from_numpy(np.concatenate(recs4)).to(dev) if recs4 else torch.zeros((0, 9), dtype=torch.int32, device=dev)
        gathered = gather_records(local_rec)                # inside the timed region: config 4 is the whole job
    else:
        sr.replay_resident(steps, conf_t, IOU, MIN_AREA)
        last = sr.drain_resident(steps)
    for st_ in sr.lane_streams:
        main_st.wait_stream(st_)
    e1.record(main_st)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sr.launch_count() - launches0
    clocks = sampler.stop()
    frames_done = (4096 if cfg_id == 4 else world * B * steps)
    value = frames_done / (ms * 1e-3)
    # the one collective: final gather of the detection records (NCCL), checked
    t_g = time.perf_counter()
    if cfg_id != 4:
        local_rec = torch.from_numpy(np.concatenate(last) if last else np.zeros((0, 9), np.int32)).to(dev)
        gathered = gather_records(local_rec)
    torch.cuda.synchronize()
    gather_ms = (time.perf_counter() - t_g) * 1e3
    cnt = torch.tensor([local_rec.shape[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    if int(cnt) != int(gathered.shape[0]):
        raise RuntimeError(f"bench: gathered {int(gathered.shape[0])} records, ranks hold {int(cnt)}")
    gather_info = {"records": int(gathered.shape[0]), "ms": None if cfg_id == 4 else gather_ms, "checked": "count == sum over ranks",
                   "in_timed_region": cfg_id == 4}
    if cfg_id == 4:
        fids = gathered[:, 0].cpu().numpy()
        gather_info["frames_with_detections"] = int(np.unique(fids).size)
        if fids.size and (fids.min() < 0 or fids.max() >= 4096):
            raise RuntimeError("bench: gathered frame ids outside [0, 4096)")
    rois_per_step = (local_rec.shape[0] / max(steps if cfg_id == 4 else len(last), 1))

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "value": value, "ms_per_step": ms / steps, "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline: single-lane pass with direct launches and CUDA events around (a) every launch of the
    # dominant conv, (b) once every op of the plan (fusions of neighbouring ops are off in that pass)
    fb0 = sr.fb[0]
    fid0 = sr.fid[0]
    n_probe = min(args.steps, 50)
    pipe.ctx.probe_set(L.NET_DETECTOR, dom)
    for _ in range(n_probe):
        pipe.enqueue_device(fb0, conf_t, IOU, MIN_AREA, fid0)
    pipe.finish(fb0, fid0)
    probe = pipe.ctx.probe_read()
    pipe.ctx.probe_set(L.NET_DETECTOR, -2)
    pipe.enqueue_device(fb0, conf_t, IOU, MIN_AREA, fid0); pipe.finish(fb0, fid0)     # warm (unfused variants)
    pipe.enqueue_device(fb0, conf_t, IOU, MIN_AREA, fid0); pipe.finish(fb0, fid0)
    per_op = pipe.ctx.probe_read()
    paths = pipe.ctx.op_paths(L.NET_DETECTOR)
    pipe.ctx.probe_set(L.NET_DETECTOR, -1)
    n_ops = len(pipe.detector.plan.ops)
    tc_ms = sum(per_op[i] for i in range(min(n_ops, len(per_op))) if paths[i] == 2)
    tc_flops = sum(2.0 * macs[i] * B for i in range(n_ops) if paths[i] == 2)
    tc_launches = sum(1 for i in range(n_ops) if paths[i] == 2)
    tail_ms = per_op[n_ops] if len(per_op) > n_ops else None

    # ---------------- per-stage figures: each stage alone, single lane, CUDA events
    def timed(fn, n=20):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / n
    det_, clf_ = pipe.detector, pipe.classifier
    A, ncd = det_.n_anchors, det_.nc
    t_lb = timed(lambda: det_.letterbox_device(fb0))
    lb_ = det_.letterbox_device(fb0)
    t_fw = timed(lambda: det_.forward_device(lb_))
    out0_ = det_.forward_device(lb_)
    t_nms = timed(lambda: det_.decode_nms_device(out0_, fb0.h[:B], fb0.w[:B], det_.ratio[:B], det_.pad[:2 * B], conf_t, IOU))
    n_cand = int(det_.n_cand[:B].sum())
    n_keep = int(torch.clamp(det_.counts[:B], max=det_.max_det).sum())
    n_r = pipe.run_device(fb0, conf_t, IOU, MIN_AREA, fid0)
    t_rs = timed(lambda: clf_.resize_device(fb0, pipe.roi_xyxy, pipe.roi_src, n_r)) if n_r else 0.0
    cls_in_ = clf_.resize_device(fb0, pipe.roi_xyxy, pipe.roi_src, n_r) if n_r else None
    t_cl = timed(lambda: clf_.classify_device(cls_in_)) if n_r else 0.0
    src_bytes = B * H * W * 3
    rx = pipe.roi_xyxy[:n_r].cpu().numpy().astype(np.int64) if n_r else np.zeros((0, 4), np.int64)
    roi_px = int(((rx[:, 2] - rx[:, 0]) * (rx[:, 3] - rx[:, 1])).sum())
    S = det_.input_size
    stage_rows = [
        ("K1 letterbox", t_lb, "hbm", (src_bytes + B * S * S * 3) / 1e9, "GB", None),
        ("K2 detector convs + Detect tail (lp_detect_forward)", t_fw, "tensor", 2.0 * sum(macs) * B / 1e12, "TFLOP", None),
        ("Detect tail (DFL softmax-expectation, dist2bbox, sigmoid)", tail_ms, "hbm",
         B * A * (pipe.detector.plan.meta["head_c"] + 4 + ncd) * 4 / 1e9, "GB", "reads the f32 head [A][64+nc], writes out0"),
        ("K4+K5 decode + threshold + NMS", t_nms, "hbm", (B * (4 + ncd) * A * 4 + 24 * n_cand * 2 + 36 * n_keep) / 1e9, "GB",
         f"latency-bound: {n_cand} candidates, {n_keep} kept in {B} frames; one block per frame"),
        ("K6 ROI resize", t_rs, "hbm", (roi_px * 3 + n_r * 64 * 64 * 3) / 1e9, "GB", f"{n_r} ROIs; latency-bound"),
        ("K7 ShuffleNetV2", t_cl, "tensor", n_r * 23.59e6 / 1e12, "TFLOP", f"{n_r} ROIs"),
    ]

    # ---------------- end to end through the public host API (e2e): frames in pinned HOST memory -> run_stream
    def e2e_run(batches, n, fids):
        barrier()
        e0.record(main_st)
        for st_ in sr.lane_streams + [sr.copy_stream]:
            st_.wait_event(e0)
        n_rec = 0
        for rec in sr.run_stream(batches, conf_t, IOU, MIN_AREA, frame_ids=fids):
            n_rec += rec.shape[0]
        for st_ in sr.lane_streams:
            main_st.wait_stream(st_)
        e1.record(main_st)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), n_rec

    g0, d0 = sr.steps_graph, sr.steps_direct
    if cfg_id == 4:
        e_steps = n_local_steps
        ms_e, n_rec_e = e2e_run((host_all[i * B:(i + 1) * B] for i in range(e_steps)), e_steps,
                                [ids_all[i * B:(i + 1) * B] for i in range(e_steps)])
        e2e_value = 4096 / (ms_e * 1e-3)
    else:
        e_steps = args.steps
        for _ in sr.run_stream(ring_batches(sr.n_buf), conf_t, IOU, MIN_AREA, frame_ids=[ids_step] * sr.n_buf):
            pass
        ms_e, n_rec_e = e2e_run(ring_batches(e_steps), e_steps, [ids_step] * e_steps)
        e2e_value = world * B * e_steps / (ms_e * 1e-3)
    e2e_graph_steps, e2e_direct_steps = sr.steps_graph - g0, sr.steps_direct - d0
    d2h_per_step = pipe.records.numel() * 4 + 4 + 4 * B
    h2d_per_step = B * H * W * 3 + 4 * B
    # pageable frames (what cv2.imread returns): staged into the pinned ring by the runner's copy threads first
    p_steps = max(2, min(e_steps, 20))
    pageable = np.array(frames, copy=True)
    ms_p, _ = e2e_run((pageable for _ in range(p_steps)), p_steps, [ids_step] * p_steps)
    e2e_pageable = world * B * p_steps / (ms_p * 1e-3)
    # frames that arrive ENCODED: JPEG quality 90, 4:2:0, restart interval 2 MCUs, decoded on the device.
    # (a) the packed scans of the batch sit in pinned host memory like the raw frames above; (b) a plain list of bytes objects
    import cv2
    js = [bytes(cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])[1]) for f in frames]
    jb = sr.pack_jpeg_batch(js)
    j_steps = e_steps if cfg_id != 4 else 20
    for _ in sr.run_stream((jb for _ in range(sr.n_buf)), conf_t, IOU, MIN_AREA, frame_ids=[ids_step] * sr.n_buf):
        pass
    h0 = sr.h2d_bytes
    ms_j, n_rec_j = e2e_run((jb for _ in range(j_steps)), j_steps, [ids_step] * j_steps)
    jpeg_h2d = (sr.h2d_bytes - h0) / j_steps
    e2e_jpeg = world * B * j_steps / (ms_j * 1e-3)
    ms_jl, _ = e2e_run((js for _ in range(p_steps)), p_steps, [ids_step] * p_steps)
    e2e_jpeg_list = world * B * p_steps / (ms_jl * 1e-3)
    dec_ = sr.jpeg[0]
    hit_ = dec_.for_header(jb.hit[0])
    import ctypes as _C
    t_dec = timed(lambda: dec_.decode_device(hit_, sr.jdev[0], sr.joff[0], B, sr.dev[0], _C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for b_ in range(sr.n_buf):
        sr.host_buffer(b_, H, W)[:] = frames                # the ring holds raw frames again
    ceiling = h2d_ceiling(sr.host[0], sr.dev[0], sr.copy_stream)
    h2d_gbs = h2d_per_step / (ms_e / e_steps * 1e-3) / 1e9

    # ---------------- p50 single-frame latency through the public API (host frame in, records out): a max_batch=1
    # instance, one captured graph per step
    pipe1 = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=NC, device=local_rank, max_batch=1, seed=0)
    sr1 = pipe1.stream(lanes=1, use_graph=False if args.no_graph else None)
    lat, lat_run = [], []
    for i in range(40):
        f1 = [frames[i % B]]
        t0 = time.perf_counter()
        sr1.run_one(f1, conf_t, IOU, MIN_AREA)
        lat.append((time.perf_counter() - t0) * 1e3)
    for i in range(15):
        t0 = time.perf_counter()
        pipe1.run(frames[i % B], conf_t, IOU, MIN_AREA)           # the reference-shaped call (e2e.py:443), per-stage events + dict building
        lat_run.append((time.perf_counter() - t0) * 1e3)
    p50, p50_run = statistics.median(lat[8:]), statistics.median(lat_run[5:])

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
        if tj.get("kernel") == pipe.detector.plan.names[dom] and B == 64 and args.detector == "v1":
            traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
    ws_gb = (pipe.detector.workspace.numel() + pipe.classifier.workspace.numel()) / 1e9
    conf_d = gpu_config_dict(cfg_id, B)
    detail = {}
    detail.update({
        "batch_per_gpu": B, "detector": args.detector,
        "weights": ("reference trained v1 (model.ncnn.bin)" if binp else "random-init seed 0 (weights not in the reference repo / not staged)"),
        "classifier_weights": "random-init seed 0 (reference ships none)",
        "conf_used": conf_t,
        "workspace_gb": ws_gb,
        "rois_per_step": rois_per_step,
        "parallelism": f"frames sharded over {world} GPU(s); {n_lanes} batches in flight per GPU (CUDA streams), step = one CUDA graph",
        "cuda_graph": {"used": sr.use_graph, "failed": sr.graph_failed, "note": getattr(sr, "graph_note", None)},
        "affinity": affinity})
    if cfg_id == 4:
        detail["frames_total"] = 4096
        detail["frames"] = "4096 distinct: 256 generated scenes, frame i = scene i % 256 rolled horizontally by 37 * (i // 256) px"
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if cfg_id == 4 else "weak",
        "vs_baseline": None, "dtype": "f16x2-split operands, f32 accumulate", "data": "synthetic",
        "config": conf_d, "config_detail": detail,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_per_step),
                "d2h_bytes_per_step": int(d2h_per_step), "ms_per_step": ms_e / e_steps, "steps": e_steps,
                "api": "B200Pipeline.stream().run_stream (pinned ring, copy stream, lanes, CUDA graph per step)",
                "graph_steps": e2e_graph_steps, "direct_steps": e2e_direct_steps, "records": n_rec_e,
                "h2d_gbs": h2d_gbs, "h2d_ceiling_gbs": ceiling, "frac_of_h2d_ceiling": h2d_gbs / ceiling if ceiling else None,
                "jpeg_frames": {"value": e2e_jpeg, "steps": j_steps, "h2d_bytes_per_step": int(jpeg_h2d), "ms_per_step": ms_j / j_steps,
                                "records": n_rec_j, "decode_ms_per_batch": t_dec,
                                "value_from_bytes_objects": e2e_jpeg_list,
                                "note": "same call with JPEG-encoded frames (quality 90, 4:2:0, restart interval 2 MCUs; "
                                        f"{sum(len(j) for j in js) / len(js) / 1e3:.0f} kB/frame): scans packed in pinned host memory, decoded on the "
                                        "device bit-exactly with cv2.imdecode; detections are those of the DECODED frames"},
                "pageable_frames": {"value": e2e_pageable, "steps": p_steps,
                                    "note": "same call with ordinary (pageable) numpy frames: host copy into the pinned ring by "
                                            f"{sr.pool._max_workers} threads included"}},
        "gpu_launches": int(launches),
        "final_gather": gather_info,
        "p50_latency_ms": p50, "p50_latency_ms_run_api": p50_run,
        "roofline": {"bound": "tensor", "kernel": pipe.detector.plan.names[dom], "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_unit": "bytes per launch (ncu dram read+write, profiles/dominant_kernel_traffic.json)",
                     "peak_source": pk_src + " bf16_tflops_sustained", "launch_ms": dom_ms,
                     "measured_in": "single-lane pass of the same steps, direct launches, CUDA events around each launch on its stream",
                     "flops_per_launch": dom_flops,
                     "kernel_all_launches": {"kernel": "conv_tc_kernel", "launches_per_step": tc_launches, "ms": tc_ms,
                                             "flops": tc_flops, "achieved": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                             "frac": (tc_flops / (tc_ms * 1e-3) / 1e12 / peak) if tc_ms else None,
                                             "note": "every op probed once in an unfused single-lane pass; sum over the ops that ran on the tcgen05 kernel"}},
    }
    hb = pk["hbm_gbs"]
    line["stages"] = [{"stage": nm, "ms": round(t, 4) if t else t, "bound": bd,
                       "achieved": (work / (t * 1e-3)) if t else None, "unit": "GB/s" if unit == "GB" else "TFLOP/s",
                       "peak": hb if bd == "hbm" else peak,
                       "frac": ((work / (t * 1e-3)) / (hb if bd == "hbm" else peak)) if t else None, "note": note}
                      for nm, t, bd, work, unit, note in stage_rows]
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(cfg_id)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(cfg_id: int, state_dict=None, n: int = 32):
    """GPU arm: the CPU leg runs in a fresh process with this process's ORIGINAL CPU affinity (the GPU arm binds itself
    to one NUMA node, and the runtimes size their thread pools when they are first imported)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-leg", "--config", str(cfg_id)],
                             capture_output=True, text=True, timeout=600,
                             preexec_fn=(lambda: os.sched_setaffinity(0, _ORIG_AFFINITY)) if _ORIG_AFFINITY else None)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": "cpu leg printed no JSON", "stderr": out.stderr[-400:]}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def cpu_leg_main(cfg_id: int, n: int = 32):
    arm = CpuArm(cfg_id, None)
    threads_all = arm.threads
    items = cpu_items(cfg_id, n if cfg_id != 3 else 1024)
    if cfg_id == 2:
        items = items[:12]
    arm.run(items[:5] if cfg_id != 3 else items)
    v, p50 = arm.run(items)
    nn = len(items) if cfg_id != 3 else 1024
    arm.set_threads(4)
    arm.run(items[:2] if cfg_id != 3 else items)
    v4, _ = arm.run(items[:12] if cfg_id != 3 else items)
    print(json.dumps({"value": v, "unit": CONFIGS[cfg_id]["unit"], "cores": threads_all, "kind": arm.kind,
                      "sample": arm.describe(nn, p50), "host_cpus": arm.host_cpus, "value_4_threads": v4}), flush=True)


# ======================================================================================= config 0
def bench_detector_single(args):
    """configs[0]: detector + NMS on one 1198x681 frame (`B200Detector.detect`, the NCNNDetector.detect drop-in)."""
    import torch
    import litepi_b200
    from litepi_b200 import _lib as L
    from litepi_b200.detector import FrameBatch
    from helpers import model_paths
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback")
    affinity = pin_to_gpu_numa(0)
    torch.cuda.set_device(0)
    param, binp = model_paths("vntsr" if args.detector == "v1" else "tt100k")
    det = litepi_b200.B200Detector(param, binp, max_batch=1, seed=0)
    frame = make_frames(0, [1])[0]
    fb = FrameBatch.from_host([frame], det.device)
    sampler = ClockSampler(0); sampler.start()
    for _ in range(max(args.warmup, 10)):
        det.detect_device(fb, CONF, IOU)
    torch.cuda.synchronize()
    sampler.wait_first()
    for _ in range(10):
        det.detect_device(fb, CONF, IOU)
    torch.cuda.synchronize()
    # device-resident: the frame stays in HBM, K detect steps captured in one CUDA graph launch each
    g = None
    if not args.no_graph:
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                det.detect_device(fb, CONF, IOU)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                det.detect_device(fb, CONF, IOU)
        except Exception as e:
            g = None
            graph_err = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize()
    l0 = det.ctx.launch_count()
    det.detect_device(fb, CONF, IOU)
    per_step = det.ctx.launch_count() - l0
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        if g is not None:
            g.replay()
        else:
            det.detect_device(fb, CONF, IOU)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # e2e: the drop-in call, host frame in -> numpy boxes/scores/classes out (e2e.py:298-316)
    for _ in range(5):
        det.detect(frame, CONF, IOU)
    lat = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        boxes, scores, classes = det.detect(frame, CONF, IOU)
        lat.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = (time.perf_counter() - t_all) * 1e3 / args.steps
    macs = det.plan.macs
    dom = int(np.argmax(macs))
    det.ctx.probe_set(L.NET_DETECTOR, dom)
    for _ in range(min(args.steps, 50)):
        det.detect_device(fb, CONF, IOU)
    torch.cuda.synchronize()
    probe = det.ctx.probe_read()
    det.ctx.probe_set(L.NET_DETECTOR, -1)
    pk, pk_src = peaks()
    dom_ms = statistics.mean(probe)
    achieved = 2.0 * macs[dom] / (dom_ms * 1e-3) / 1e12
    line = {"metric": METRIC.replace("detect+NMS+classify", "detect+NMS"), "value": args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16x2-split operands, f32 accumulate", "data": "synthetic",
            "config": gpu_config_dict(0, 1),
            "config_detail": {"detector": args.detector, "detections": int(len(boxes)), "affinity": affinity, "cuda_graph": g is not None},
            "clocks": clocks,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": int(frame.nbytes),
                    "d2h_bytes_per_step": int(4 + (16 + 4 + 8) * det.max_det), "ms_per_step": e2e_ms, "p50_ms": statistics.median(lat),
                    "api": "B200Detector.detect(image, conf, iou): pageable numpy frame in, numpy boxes/scores/class_ids out"},
            "gpu_launches": int(per_step * args.steps), "p50_latency_ms": statistics.median(lat),
            "roofline": {"bound": "tensor", "kernel": det.plan.names[dom], "achieved": achieved, "peak": pk["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"], "traffic": None, "launch_ms": dom_ms,
                         "peak_source": pk_src + " bf16_tflops_sustained",
                         "note": "batch 1: 50 tiles of 128 pixels on 148 SMs -- a latency-bound launch by construction"}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(0)
    print(json.dumps(line), flush=True)


# ======================================================================================= config 3
def bench_classifier_alone(args):
    """configs[3]: ShuffleNetV2 x1.0 alone on 1024 synthetic crops (15 real debug_rois when staged + glyph crops)."""
    import torch
    import litepi_b200
    from litepi_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback")
    affinity = pin_to_gpu_numa(0)
    torch.cuda.set_device(0)
    N = args.batch or 1024
    clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, max_batch=N, seed=0)
    crops = synth.roi_crops(N, seed=7)
    x = clf.preprocess_batch(crops).contiguous()             # [N,64,64,3] RGB u8, resident
    sampler = ClockSampler(0); sampler.start()
    for _ in range(max(args.warmup, 5)):
        clf.classify_device(x)
    torch.cuda.synchronize()
    sampler.wait_first()
    for _ in range(5):
        clf.classify_device(x)
    l0 = clf.ctx.launch_count()
    clf.classify_device(x)
    per_step = clf.ctx.launch_count() - l0
    torch.cuda.synchronize()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        clf.classify_device(x)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    for _ in range(2):
        clf.predict_batch(crops)
    t0 = time.perf_counter()
    n_e = max(3, min(args.steps, 20))
    for _ in range(n_e):
        cls, probs = clf.predict_batch(crops)                # host crops in (ragged), probs out
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e
    pk, pk_src = peaks()
    flops = N * 23.59e6
    ach = flops / (ms / args.steps * 1e-3) / 1e12
    line = {"metric": "ShuffleNetV2 classifier crops/sec", "value": N * args.steps / (ms * 1e-3), "unit": "crops/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 / f16x2-split tensor-core pointwise layers", "data": "synthetic",
            "config": gpu_config_dict(3, N),
            "config_detail": {"affinity": affinity,
                              "reference_published": "279.2 img/s PyTorch CPU batch 64 (src/vntsr/evaluation-tsr.ipynb:407)"},
            "clocks": clocks,
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": int(sum(c.nbytes for c in crops)),
                    "d2h_bytes_per_step": int(N * 49 * 4), "ms_per_step": e2e_ms,
                    "api": "B200Classifier.predict_batch(list of ragged BGR crops): per-crop pageable H2D, Pillow-exact resize on the device, probs out"},
            "gpu_launches": int(per_step * args.steps),
            "roofline": {"bound": "tensor", "kernel": "shufflenet_fused_kernel", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                         "peak_source": pk_src + " bf16_tflops_sustained", "flops_per_launch": flops,
                         "note": "23.59 MFLOP per crop (SURVEY.md 8d); the kernel is bound by its weight stream and fp32 FMA tail, not by the tensor pipe"}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(3)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
