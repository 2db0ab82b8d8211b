"""Shared test helpers (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF = "/root/reference"
STAGE = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def model_paths(which: str = "vntsr"):
    """(param, bin) of the reference's exported detector.  Looks in /root/reference (authoring
    container) then oracle/_ref (staged by build() for the GPU box).  bin is None when the
    trained weights are unavailable (tt100k always; vntsr if nothing was staged): callers then
    use seeded random weights of the same architecture."""
    ref_dir = os.path.join(REF, f"src/{which}/convert/model/yolo_plus/yolo_plus_ncnn_model")
    cands = [(os.path.join(ref_dir, "model.ncnn.param"), os.path.join(ref_dir, "model.ncnn.bin")),
             (os.path.join(STAGE, f"{which}.model.ncnn.param"), os.path.join(STAGE, f"{which}.model.ncnn.bin")),
             (os.path.join(GOLDEN, f"{which}.model.ncnn.param"), None)]
    for p, b in cands:
        if os.path.exists(p):
            return p, (b if b and os.path.exists(b) else None)
    raise FileNotFoundError(f"no model.ncnn.param for {which}")


def onnx_path():
    for p in (os.path.join(REF, "src/vntsr/convert/model/yolo_plus/yolo_plus.onnx"),
              os.path.join(STAGE, "vntsr.yolo_plus.onnx")):
        if os.path.exists(p):
            return p
    return None


def debug_roi_paths():
    for d in (os.path.join(REF, "src/vntsr/pipeline/debug_rois"), os.path.join(STAGE, "debug_rois")):
        if os.path.isdir(d):
            return [os.path.join(d, f) for f in sorted(os.listdir(d))]
    return []


def oracle_pipeline_run(det_oracle, clf_model, frame, conf, iou, min_area):
    """HybridPipeline.run (e2e.py:443-531) restated with the oracle pieces; returns result dicts
    with the extra float32 box under 'box_f32'."""
    from oracle import pipeline_ref as PR
    x, r, pad, _ = PR.preprocess_ref(frame)
    out0 = det_oracle.forward(x)[0].numpy()
    boxes, scores, classes = PR.postprocess_ref(out0, frame.shape[:2], r, pad, conf, iou)
    rois, valid = PR.roi_select_ref(boxes, frame.shape[:2], min_area)
    crops = [frame[y1:y2, x1:x2] for (x1, y1, x2, y2) in rois]
    cls, probs, _ = PR.classify_ref(clf_model, crops)
    res = []
    for j, k in enumerate(valid):
        res.append({"bbox": tuple(boxes[k].astype(int)), "box_f32": boxes[k].astype(np.float32),
                    "det_class": int(classes[k]), "det_conf": float(scores[k]),
                    "cls_class": int(cls[j]), "cls_conf": float(np.max(probs[j]))})
    return res


def load_eval_case(name: str):
    """(all_preds, all_gts, num_classes, expected dict) of tests/golden/eval_cases.npz -- inputs and outputs of the
    reference's evaluate_predictions recorded by tests/golden/make_golden.py."""
    import numpy as np
    d = np.load(os.path.join(GOLDEN, "eval_cases.npz"))
    pn, gn = d[f"{name}.in.pred_n"], d[f"{name}.in.gt_n"]
    pb, pc, pk, gb = d[f"{name}.in.pred_box"], d[f"{name}.in.pred_conf"], d[f"{name}.in.pred_cls"], d[f"{name}.in.gt"]
    preds, gts, a, b = [], [], 0, 0
    for i in range(len(pn)):
        preds.append([{"bbox": tuple(pb[a + j]), "conf": float(pc[a + j]), "cls_class": int(pk[a + j])} for j in range(pn[i])])
        gts.append([list(gb[b + j]) for j in range(gn[i])])
        a += pn[i]; b += gn[i]
    want = {k.split(".out.")[1]: d[k] for k in d.files if k.startswith(name + ".out.")}
    return preds, gts, int(d[f"{name}.in.num_classes"]), want


def make_variant_param(src_param: str, dst_param: str, nc: int = 1, in_size: int = 640) -> str:
    """Rewrite the reference's model.ncnn.param into a variant of the same graph family with ``nc`` classes and/or
    another input size: the three class-logit convs (1x1, cout 1 -> nc), the per-level reshapes, the 64|nc slice and
    the anchor-count constants (6400/1600/400/8400 at 640).  The product reads only the Convolution records; the
    graph oracle executes every line, so both see the same architecture."""
    import re
    sizes = [(in_size // s) ** 2 for s in (8, 16, 32)]
    a_old, a_new = 8400, sum(sizes)
    out = []
    with open(src_param) as f:
        lines = f.read().splitlines()
    for ln in lines:
        tok = ln.split()
        if tok and tok[0] == "Convolution" and "0=1" in tok and "1=1" in tok and "5=1" in tok:
            cin = int([t for t in tok if t.startswith("6=")][0][2:])
            ln = re.sub(r"\b0=1\b", f"0={nc}", ln, count=1)
            ln = re.sub(r"\b6=%d\b" % cin, f"6={cin * nc}", ln)
        elif tok and tok[0] == "Reshape":
            lvl = dict(zip(("6400", "1600", "400"), sizes))
            ln = re.sub(r"\b0=(6400|1600|400) 1=65\b", lambda m: f"0={lvl[m.group(1)]} 1={64 + nc}", ln)   # one pass: 6400->1600 must not cascade
            ln = re.sub(r"\b0=%d\b" % a_old, f"0={a_new}", ln)
        elif tok and tok[0] == "Slice" and "-23300=2,64,1" in ln:
            ln = ln.replace("-23300=2,64,1", f"-23300=2,64,{nc}")
        elif tok and tok[0] == "MemoryData":
            ln = re.sub(r"\b0=%d\b" % a_old, f"0={a_new}", ln)
        out.append(ln)
    with open(dst_param, "w") as f:
        f.write("\n".join(out) + "\n")
    return dst_param
