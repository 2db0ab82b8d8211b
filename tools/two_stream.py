"""Throughput of S pipelines on S CUDA streams, steps issued round-robin without host synchronisation
(development tool: does cross-step overlap fill the tails of the persistent kernels?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import synth
from litepi_b200.detector import FrameBatch
from helpers import model_paths
B = 64
frames = [synth.vn_frame(i) for i in range(B)]
dev = torch.device("cuda:0")
for S in (1, 2, 3):
    pipes = [litepi_b200.B200Pipeline(*model_paths("vntsr"), None, "shufflenetv2", num_classes=49, max_batch=B, seed=0) for _ in range(S)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    fbs = [FrameBatch.from_host(frames, dev) for _ in range(S)]
    torch.cuda.synchronize()
    def step(i):
        s = i % S
        with torch.cuda.stream(streams[s]):
            pipes[s].enqueue_device(fbs[s], 0.25, 0.45, 50)
    for i in range(3 * S): step(i)
    torch.cuda.synchronize()
    K = 60
    t0 = time.perf_counter()
    for i in range(K): step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ns = []
    for s in range(S):
        with torch.cuda.stream(streams[s]):
            ns.append(pipes[s].finish(fbs[s]))
    print(f"streams {S}: {dt / K * 1e3:.3f} ms/step -> {B * K / dt:.0f} frames/s (rois {ns})")
    del pipes, fbs
    torch.cuda.empty_cache()
