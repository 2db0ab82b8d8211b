"""Classifier-alone timing: fused persistent kernel vs layer-by-layer plan (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from oracle import pipeline_ref as PR
ref = PR.build_shufflenet(49, seed=0)
for G in (1, 2, 3, 4):
    clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=1024, fused_group=G)
    for n in (148, 296, 333, 444, 1024):
        x = torch.randint(0, 255, (n, 64, 64, 3), dtype=torch.uint8, device=clf.device)
        for fused in (True,):
            clf.set_fused(fused)
            for _ in range(3): clf.classify_device(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): clf.classify_device(x)
            e1.record(); e1.synchronize()
            print(f"G={G} n={n} fused={fused}: {e0.elapsed_time(e1)/10*1e3:.0f} us  ({n/(e0.elapsed_time(e1)/10)*1e3:.0f} ROIs/s)")
