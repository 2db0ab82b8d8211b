"""Per-layer precision budget -- TEST INFRASTRUCTURE.

Which detector convs could run with single-pass fp16 operands (1 tcgen05.mma per K-step, 2 B/element) while out0 stays
inside the north_star tolerance (boxes 1e-2 px in ORIGINAL pixels at candidates, scores 1e-3) with a 2x margin?
Formats per conv (fp32 accumulation everywhere; baseline = both operands hi+lo fp16 = 3 products):
  s   activations AND weights rounded to one fp16        (1 MMA, N = cout, reads/writes one plane)
  a1  activations one fp16, weights hi+lo                (1 MMA, N = 2*cout)
  w1  weights one fp16, activations hi+lo                (2 MMAs, N = cout)
Step 1: each conv alone in the cheaper format, error of out0 vs the all-baseline run.  Step 2: greedy -- add convs in
order of their solo error while the joint error stays below tolerance / 2.
Run:  python -m oracle.experiments.precision_budget [n_vn_frames n_tt_frames]
"""
import os
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "yolo-litepi_b200"))
import synth  # noqa: E402
from oracle.ncnn_graph import DetectorOracle, run_graph  # noqa: E402

R = "/root/reference/src/vntsr/convert/model/yolo_plus/yolo_plus_ncnn_model/"
BOX_TOL, SCORE_TOL = 1e-2, 1e-3


def q1(t): return t.half().float()
def q2(t):
    hi = t.half().float()
    return hi + (t - hi).half().float()


FMT = {"x2": (q2, q2), "s": (q1, q1), "a1": (q1, q2), "w1": (q2, q1)}


def letterbox_in(frame):
    h, w = frame.shape[:2]; r = min(640 / h, 640 / w); nw, nh = round(w * r), round(h * r)
    im = cv2.resize(frame, (nw, nh)) if (nw, nh) != (w, h) else frame
    top, left = int(round((640 - nh) / 2 - 0.1)), int(round((640 - nw) / 2 - 0.1))
    lb = np.full((640, 640, 3), 114, np.uint8); lb[top:top + nh, left:left + nw] = im
    return torch.from_numpy(lb[:, :, ::-1].copy()).permute(2, 0, 1)[None].float() / 255, r


def main():
    n_vn = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_tt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    det = DetectorOracle(R + "model.ncnn.param", R + "model.ncnn.bin")
    convs = [L.name for L in det.layers if L.type == "Convolution" and L.bias is not None]
    frames = [synth.vn_frame(s) for s in range(n_vn)] + [synth.tt_frame(s) for s in range(n_tt)]
    xs, rs = zip(*[letterbox_in(f) for f in frames])
    x = torch.cat(xs)
    ratio = np.array(rs)

    def run(assign):
        return run_graph(det.layers, x, quant_for=lambda name: FMT[assign.get(name, "x2")])["out0"].numpy()

    base = run({})
    fp32 = run_graph(det.layers, x)["out0"].numpy()
    cand = fp32[:, 4] > 0.25

    def err(out, ref):
        d = np.abs(out - ref)
        box = d[:, :4].max(1) / ratio[:, None]                       # original-image pixels
        return float(box[cand].max()) if cand.any() else 0.0, float(box.max()), float(d[:, 4].max())

    print(f"frames: {n_vn} VN + {n_tt} TT, candidates (score > 0.25): {int(cand.sum())}")
    print("baseline (every conv hi+lo) vs fp32: box@cand %.2e  box@all %.2e  score %.2e" % err(base, fp32))
    solo = {}
    macs = {}
    for L in det.layers:
        if L.type == "Convolution" and L.bias is not None:
            macs[L.name] = int(np.prod(L.weight.shape))
    print(f"\n{'conv':10s} {'fmt':3s}  box@cand   box@all    score      (that conv alone in the cheaper format, vs fp32)")
    for name in convs:
        for fmt in ("s", "a1", "w1"):
            e = err(run({name: fmt}), fp32)
            solo[(name, fmt)] = e
            print(f"{name:10s} {fmt:3s}  {e[0]:9.2e}  {e[1]:9.2e}  {e[2]:9.2e}", flush=True)
    for fmt in ("s", "a1", "w1"):
        order = sorted(convs, key=lambda n: max(solo[(n, fmt)][0] / BOX_TOL, solo[(n, fmt)][2] / SCORE_TOL))
        chosen = {}
        for n in order:
            trial = dict(chosen); trial[n] = fmt
            e = err(run(trial), fp32)
            if e[0] <= BOX_TOL / 2 and e[2] <= SCORE_TOL / 2:
                chosen = trial
        e = err(run(chosen), fp32) if chosen else err(base, fp32)
        print(f"\ngreedy set for format {fmt} (joint error <= tolerance/2): {len(chosen)} of {len(convs)} convs: {sorted(chosen)}")
        print("  joint error: box@cand %.2e  box@all %.2e  score %.2e" % e)


if __name__ == "__main__":
    main()
