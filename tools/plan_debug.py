"""Per-op comparison of a classifier layer plan on the GPU against the CPU plan interpreter (development tool):
prints, for every op in order, the max abs difference of its output buffer slice.  usage: plan_debug.py <arch> [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import _lib as L
from oracle import pipeline_ref as PR
from plan_interp import run_plan_cpu

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ref = PR.build_classifier_ref(arch, 49, seed=3)
c = litepi_b200.B200Classifier(None, arch, num_classes=49, state_dict=ref.state_dict(), max_batch=8, fused=False)
x = np.random.default_rng(5).integers(0, 256, (n, 64, 64, 3), dtype=np.uint8)
lg = c.logits_for(x)
bufs, logits = run_plan_cpu(c.plan, x)
ws = c.workspace.cpu().numpy()
print("logits diff", np.abs(lg - logits.numpy()).max())
for i, (nm, op) in enumerate(zip(c.plan.names, c.plan.ops)):
    if op["out_buf"] < 0:
        continue
    b = c.plan.bufs[op["out_buf"]]
    if b["fmt"] != L.FMT_SPLIT16:
        continue
    per = b["image_bytes"] // 2
    hi = ws[b["offset"]:b["offset"] + c.max_batch * b["image_bytes"]].view(np.float16).reshape(c.max_batch, b["h"], b["w"], b["c"])
    lo_off = b["offset"] + c.max_batch * b["image_bytes"]
    lo = ws[lo_off:lo_off + c.max_batch * b["image_bytes"]].view(np.float16).reshape(c.max_batch, b["h"], b["w"], b["c"])
    got = hi[:n].astype(np.float32) + lo[:n].astype(np.float32)
    want = bufs[op["out_buf"]].numpy()
    sl = slice(op["out_coff"], op["out_coff"] + op["cout"])
    d = np.abs(got[..., sl] - want[..., sl])
    per_img = d.reshape(n, -1).max(1)
    print(f"{i:3d} {nm:28s} kind {op['kind']} k{op['ksize']} s{op['stride']} {b['h']:2d}x{b['w']:<2d} cin {op['cin']:4d} cout {op['cout']:4d} "
          f"path {c.ctx.op_paths(L.NET_CLASSIFIER)[i]} max|d| {d.max():.3e} per image {np.array2string(per_img, precision=2)} ref max {np.abs(want[..., sl]).max():.2f}")
