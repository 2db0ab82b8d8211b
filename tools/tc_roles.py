"""Per-role cycle breakdown of conv_tc_kernel for selected detector ops (development tool)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from litepi_b200 import synth, _lib as L
from helpers import model_paths
B = int(os.environ.get("LP_B", "64"))
det = litepi_b200.B200Detector(*model_paths("vntsr"), max_batch=B)
x = torch.randint(0, 255, (B, 640, 640, 3), dtype=torch.uint8, device=det.device)
det.forward_device(x); torch.cuda.synchronize()
dbg = torch.zeros(16, dtype=torch.int64, device=det.device)
names = sys.argv[1:] or ["conv_4", "conv_9", "conv_10", "conv_47", "conv_48", "conv_49", "conv_53", "conv_15"]
# run the plan op by op is not exposed; instead enable timing for the whole forward but only keep ops by probing each:
# trick: tc_dbg is overwritten by every TC launch, so run the forward with the plan truncated after the op of interest
import copy
full_ops = det.plan.ops
for nm in names:
    k = det.plan.names.index(nm)
    plan_ops = full_ops[:k + 1]
    bufs, _ = det.plan.c_arrays()
    ops = (L.OpDesc * len(plan_ops))()
    keys = [f[0] for f in L.OpDesc._fields_]
    for i, o in enumerate(plan_ops): ops[i] = L.OpDesc(*[o[kk] for kk in keys])
    L.check(L.lib().lp_net_load(det.ctx.handle, L.NET_DETECTOR, bufs, len(bufs), ops, len(plan_ops), C.c_void_p(det.weights.data_ptr()),
                                det.weights.numel(), C.c_void_p(det.weights_tc.data_ptr()), det.weights_tc.numel(), B))
    L.check(L.lib().lp_debug_tc_timing(det.ctx.handle, C.c_void_p(dbg.data_ptr())))
    dbg.zero_()
    # lp_detect_forward needs the head buffer last; call the plan through detect_forward anyway (tail reads garbage, harmless)
    L.lib().lp_detect_forward(det.ctx.handle, C.c_void_p(x.data_ptr()), B, C.c_void_p(det.workspace.data_ptr()), det.workspace.numel(),
                              C.c_void_p(det.out0.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    t = max(int(d[7]), 1)
    o = full_ops[k]; n_mma = o["ksize"] ** 2 * (o["cin"] // 16) * 2
    print(f"[{o['ksize']}x{o['ksize']} s{o['stride']} {o['cin']}->{o['cout']}, {n_mma} MMA/tile, {d[6]/t/n_mma:.0f} cyc/MMA] ", end="")
    print(f"{nm:8s} tiles/CTA {t:3d} | loader/tile: wait_empty {d[0]/t:7.0f} issue {d[1]/t:7.0f} wait_cp {d[2]/t:7.0f} | "
          f"MMA/tile: wait_acc {d[3]/t:7.0f} wait_patch {d[4]/t:7.0f} wait_w {d[5]/t:7.0f} total {d[6]/t:7.0f} | "
          f"epi/tile: wait {d[8]/t:6.0f} total {d[9]/t:6.0f} p1 {d[10]/t:6.0f} stage_wait {d[11]/t:6.0f} prologue {d[14]/t:6.0f} | store/tile: wait {d[15]/t:6.0f} copy {d[12]/t:6.0f}")
