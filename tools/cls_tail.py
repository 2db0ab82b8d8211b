"""Fused classifier: fp32 tail vs tensor-core tail (LP_CLS_TAIL_MMA), time and logits against torchvision (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import litepi_b200
from oracle import pipeline_ref as PR
ref = PR.build_shufflenet(49, seed=0).eval()
x = torch.randint(0, 255, (333, 64, 64, 3), dtype=torch.uint8, device="cuda")
xin = ((x.float().cpu() / 255 - 0.18) / 0.34).permute(0, 3, 1, 2).contiguous()
with torch.no_grad():
    want = ref(xin).numpy()
for mode in ("0", "1"):
    os.environ["LP_CLS_TAIL_MMA"] = mode
    for G in (2, 3):
        try:
            clf = litepi_b200.B200Classifier(None, "shufflenetv2", num_classes=49, state_dict=ref.state_dict(), max_batch=1024, fused_group=G)
        except Exception as e:
            print(f"tail_mma={mode} G={G}: {str(e)[:120]}"); continue
        for n in (148, 333, 1024):
            xs = x[:n] if n <= 333 else torch.randint(0, 255, (n, 64, 64, 3), dtype=torch.uint8, device="cuda")
            for _ in range(3): clf.classify_device(xs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): clf.classify_device(xs)
            e1.record(); e1.synchronize()
            msg = f"tail_mma={mode} G={G} n={n}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us"
            if n == 333:
                got = clf.logits[:n].cpu().numpy()
                msg += f"  max|dlogit| {np.abs(got - want).max():.2e} top1 equal {np.array_equal(got.argmax(1), want.argmax(1))}"
            print(msg)
