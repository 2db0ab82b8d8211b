"""Per-step cycle counts of the fused classifier kernel (CTA 0) -- development tool."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import _lib as L, plan
from oracle import pipeline_ref as PR
ref = PR.build_shufflenet(49, seed=0)
clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=512, fused_group=int(os.environ.get("LP_FG", "3")))
n = 148
x = torch.randint(0, 255, (n, 64, 64, 3), dtype=torch.uint8, device=clf.device)
clf.classify_device(x); torch.cuda.synchronize()
dbg = torch.zeros(128, dtype=torch.int64, device=clf.device)
L.check(L.lib().lp_debug_tc_timing(clf.ctx.handle, C.c_void_p(dbg.data_ptr())))
clf.classify_device(x); torch.cuda.synchronize()
d = dbg.cpu().numpy()
steps = clf.fused_steps.cpu().numpy()
names = ["CONV1", "MAXPOOL", "PW", "DW", "COPY", "MEANFC", "STORE", "LOAD"]
tot = d[:len(steps)].sum()
agg = {}
for i, st in enumerate(steps):
    agg[names[st[0]]] = agg.get(names[st[0]], 0) + int(d[i])
print("total cycles per ROI (CTA 0, 1 ROI):", int(tot), "=", tot / 1.85e3, "us")
for k, v in agg.items(): print(f"  {k:8s} {v:9d} cycles {v/tot*100:5.1f}%")
for i, st in enumerate(steps):
    print(i, names[st[0]], "cin", st[8], "cout", st[9], "HW", st[10], "cycles", int(d[i]))
