"""Evaluation matching throughput at the 4096-frame size of BASELINE configs[4]:
lp_eval_match on the GPU (device-resident inputs, CUDA events) next to the oracle port on one host core
(bounded sample).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ctypes as C
import numpy as np, torch
import litepi_b200
from litepi_b200 import _lib as L
from make_golden import eval_case, pack_eval_case
from oracle import eval_ref as ER

F = 4096
preds, gts = eval_case(21, F, 49)
d = pack_eval_case(preds, gts)
ev = litepi_b200.Evaluator()
dev = ev.device
pn, gn = d["pred_n"], d["gt_n"]
off = lambda n: torch.from_numpy(np.concatenate(([0], np.cumsum(n))).astype(np.int32)).to(dev)
pb = torch.from_numpy(d["pred_box"].astype(np.float64)).to(dev); pc = torch.from_numpy(d["pred_cls"].astype(np.int32)).to(dev)
gb = torch.from_numpy(np.ascontiguousarray(d["gt"][:, 1:])).to(dev); gc = torch.from_numpy(d["gt"][:, 0].astype(np.int32)).to(dev)
po, go = off(pn), off(gn)
th = torch.from_numpy(ER.IOU_THRESHOLDS.copy()).to(dev)
P = int(pn.sum())
correct = torch.empty((P, 10), dtype=torch.uint8, device=dev)
most = int((pn + gn).max())
ptr = lambda t: C.c_void_p(t.data_ptr())
def run():
    L.check(L.lib().lp_eval_match(ev.ctx.handle, ptr(pb), ptr(pc), ptr(po), ptr(gb), ptr(gc), ptr(go), F, most, ptr(th), 10,
                                  ptr(correct), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 50
# CPU port on a bounded sample, and parity on that sample
n_s = 512
t0 = time.perf_counter()
a = b = 0; ok = True
got = correct.cpu().numpy().astype(bool)
for f in range(n_s):
    g = np.asarray(gts[f], np.float64).reshape(-1, 5)
    want = ER.match_image_ref(d["pred_box"][a:a + pn[f]], d["pred_cls"][a:a + pn[f]], g[:, 1:], g[:, 0])
    ok &= np.array_equal(got[a:a + pn[f]], want)
    a += pn[f]; b += gn[f]
cpu_s = time.perf_counter() - t0
algo_bytes = P * 4 * 8 + int(gn.sum()) * 4 * 8 + P * 4 + int(gn.sum()) * 4 + P * 10
print(json.dumps({"metric": "evaluation matching frames/s", "frames": F, "predictions": P, "ground_truths": int(gn.sum()),
                  "gpu_ms": ms, "gpu_frames_per_s": F / (ms * 1e-3), "algorithmic_bytes": algo_bytes,
                  "achieved_GBps": algo_bytes / (ms * 1e-3) / 1e9, "bound": "latency (one small block per frame)",
                  "cpu_port_frames_per_s": n_s / cpu_s, "cpu_sample_frames": n_s, "parity_on_sample": bool(ok)}))
