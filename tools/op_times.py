"""Per-op CUDA-event times of the detector / classifier plans at batch 64 (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import synth, _lib as L
from litepi_b200.detector import FrameBatch
from oracle import pipeline_ref as PR
from helpers import model_paths
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
param, binp = model_paths("vntsr")
clf_ref = PR.build_shufflenet(49, seed=0)
pipe = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=49, max_batch=B, classifier_state_dict=clf_ref.state_dict(), seed=0)
fb = FrameBatch.from_host([synth.vn_frame(i) for i in range(B)], pipe.device)
for _ in range(3): n = pipe.run_device(fb, 0.25, 0.45, 50)
for name, obj, net in (("detector", pipe.detector, L.NET_DETECTOR), ("classifier", pipe.classifier, L.NET_CLASSIFIER)):
    acc = None
    R = 5
    for _ in range(R):
        obj.ctx.probe_set(net, -2)
        pipe.run_device(fb, 0.25, 0.45, 50)
        torch.cuda.synchronize()
        t = np.array(obj.ctx.probe_read())
        acc = t if acc is None else acc + t[:len(acc)]
    obj.ctx.probe_set(net, -1)
    acc = acc / R * 1e3
    P = obj.plan
    print(f"== {name}: {len(acc)} ops, total {acc.sum():.0f} us (batch {B}, rois {n})")
    for i, (nm, op, us, mac) in enumerate(zip(P.names, P.ops, acc, P.macs)):
        hb = P.bufs[op['in_buf']]
        units = B if name == "detector" else n
        tf = 2 * mac * units / (us * 1e-6) / 1e12 if mac else 0
        tc = "TC" if op.get("wtc_off", -1) >= 0 else "  "
        print(f"{i:3d} {nm:28s} {tc} k{op['ksize']} s{op['stride']} {hb['h']:3d}x{hb['w']:<3d} cin{op['cin']:4d} cout{op['cout']:4d} {us:8.1f} us {tf:7.1f} TF/s")
