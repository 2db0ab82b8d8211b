"""Detector forward time at batch 64 (CUDA events over 30 back-to-back forwards; development tool for env-switch sweeps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import litepi_b200
from helpers import model_paths
B = int(os.environ.get("LP_B", "64"))
which = sys.argv[1] if len(sys.argv) > 1 else "vntsr"
det = litepi_b200.B200Detector(*model_paths(which), max_batch=B)
x = torch.randint(0, 255, (B, 640, 640, 3), dtype=torch.uint8, device=det.device)
for _ in range(5):
    det.forward_device(x)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        det.forward_device(x)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
print(f"detector forward {best * 1e3:.1f} us  (LP_TC_DEPTH={os.environ.get('LP_TC_DEPTH')}, LP_NO_C2F={os.environ.get('LP_NO_C2F')})")
