"""Precision-budget experiment -- TEST INFRASTRUCTURE.

Emulates candidate tensor-core operand formats for the detector convs by
rounding every conv's input activations and weights, keeping fp32 accumulation,
and reports the error of out0 against the fp32 oracle in the units of the
north_star tolerance: boxes in ORIGINAL-image pixels (letterbox px / ratio) and
scores absolute.  Run:  python -m oracle.experiments.precision
"""
import sys, os
import numpy as np, torch, cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "yolo-litepi_b200"))
import synth  # noqa: E402
from oracle.ncnn_graph import DetectorOracle, run_graph  # noqa: E402

R = "/root/reference/src/vntsr/convert/model/yolo_plus/yolo_plus_ncnn_model/"


def q_fp16(t): return t.half().float()
def q_bf16(t): return t.bfloat16().float()
def q_tf32(t):  # round-to-nearest-even to 10 mantissa bits
    i = t.contiguous().view(torch.int32)
    r = ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)
def q_fp16x2(t):  # hi + lo, both fp16 (3-MMA split scheme)
    hi = t.half().float()
    return hi + (t - hi).half().float()
def q_bf16x2(t):
    hi = t.bfloat16().float()
    return hi + (t - hi).bfloat16().float()
def q_bf16x3(t):
    hi = t.bfloat16().float(); r = t - hi
    mid = r.bfloat16().float()
    return hi + mid + (r - mid).bfloat16().float()


def letterbox_in(frame):
    h, w = frame.shape[:2]; r = min(640 / h, 640 / w); nw, nh = round(w * r), round(h * r)
    im = cv2.resize(frame, (nw, nh)) if (nw, nh) != (w, h) else frame
    top, left = int(round((640 - nh) / 2 - 0.1)), int(round((640 - nw) / 2 - 0.1))
    lb = np.full((640, 640, 3), 114, np.uint8); lb[top:top + nh, left:left + nw] = im
    return torch.from_numpy(lb[:, :, ::-1].copy()).permute(2, 0, 1)[None].float() / 255, r


def main():
    det = DetectorOracle(R + "model.ncnn.param", R + "model.ncnn.bin")
    frames = [synth.vn_frame(s) for s in range(4)] + [synth.tt_frame(s) for s in range(2)]
    print(f"{'format':10s} {'box err px (orig) max/p99 @score>0.25':42s} {'all-anchor box max':20s} score max")
    for name, q in (("fp16", q_fp16), ("bf16", q_bf16), ("tf32", q_tf32), ("bf16x2", q_bf16x2),
                    ("fp16x2", q_fp16x2), ("bf16x3", q_bf16x3)):
        eb_c, eb_a, es = [], [], []
        for f in frames:
            x, r = letterbox_in(f)
            ref = run_graph(det.layers, x)["out0"][0].numpy()
            got = run_graph(det.layers, x, quant=q)["out0"][0].numpy()
            d = np.abs(got - ref)
            box = d[:4].max(0) / r
            cand = ref[4] > 0.25
            eb_c.append(box[cand]); eb_a.append(box.max()); es.append(d[4].max())
        ec = np.concatenate(eb_c)
        print(f"{name:10s} max {ec.max():9.2e}  p99 {np.percentile(ec, 99):9.2e}  mean {ec.mean():9.2e}"
              f"   {max(eb_a):9.2e}          {max(es):9.2e}")


if __name__ == "__main__":
    main()
