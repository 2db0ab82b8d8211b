"""Two detector forwards at batch 64 on random frames (development tool: the command ncu wraps to capture detector kernels)."""
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
for _ in range(2):
    det.forward_device(x)
torch.cuda.synchronize()
print("ok", det.ctx.op_paths(0)[:66])
