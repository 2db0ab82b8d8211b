"""Oracle: fp32 CPU execution of the reference's exported detector graph.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This walks the reference's
``model.ncnn.param`` line by line and executes every layer with plain torch
fp32 CPU ops, i.e. it restates what ``ex.extract("out0")`` computes at
``src/vntsr/pipeline/e2e.py:305-307`` for the graph in
``src/vntsr/convert/model/yolo_plus/yolo_plus_ncnn_model/model.ncnn.param:1-208``.
The arithmetic itself lives in the un-vendored ncnn runtime ("manual_build",
``requirements.txt:54-58``); the layer semantics restated here are ncnn's
published ones (Convolution/Swish/Slice/Split/BinaryOp/Concat/Pooling/Interp/
Reshape/Permute/Softmax/Sigmoid/MemoryData) and are pinned by comparing the
whole graph against OpenCV-DNN running the reference's ``yolo_plus.onnx`` (same
weights) in ``tests/test_oracle_detector.py``.

Weight container layout (``model.ncnn.bin``): layers in ``.param`` order; each
Convolution = uint32 flag (0 = fp32) + weights (OIHW) + bias; each MemoryData =
raw fp32 without flag.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclasses.dataclass
class Layer:
    type: str
    name: str
    inputs: List[str]
    outputs: List[str]
    params: Dict[int, object]
    weight: Optional[np.ndarray] = None   # Convolution: OIHW fp32; MemoryData: the tensor
    bias: Optional[np.ndarray] = None


def parse_param(path: str) -> List[Layer]:
    """Parse an ncnn text ``.param`` file into a layer list (file order)."""
    with open(path, "r") as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    if lines[0] != "7767517":
        raise ValueError(f"not an ncnn param file (magic {lines[0]!r})")
    n_layers, _n_blobs = (int(v) for v in lines[1].split())
    layers: List[Layer] = []
    for ln in lines[2:]:
        tok = ln.split()
        ltype, name, n_in, n_out = tok[0], tok[1], int(tok[2]), int(tok[3])
        ins = tok[4:4 + n_in]
        outs = tok[4 + n_in:4 + n_in + n_out]
        params: Dict[int, object] = {}
        for kv in tok[4 + n_in + n_out:]:
            k, v = kv.split("=")
            k = int(k)
            if k <= -23300:                       # array parameter: count,v0,v1,...
                vals = v.split(",")
                params[-23300 - k] = [_num(x) for x in vals[1:1 + int(vals[0])]]
            else:
                params[k] = _num(v)
        layers.append(Layer(ltype, name, ins, outs, params))
    if len(layers) != n_layers:
        raise ValueError(f"layer count mismatch: header {n_layers}, parsed {len(layers)}")
    return layers


def _num(s: str):
    try:
        return int(s)
    except ValueError:
        return float(s)


def conv_input_channels(layer: Layer) -> int:
    p = layer.params
    return p[6] // (p[0] * p[1] * p.get(11, p[1]))


def load_bin(layers: List[Layer], path: str) -> None:
    """Attach weights from ``model.ncnn.bin`` (SURVEY.md App. D.2 layout)."""
    raw = np.fromfile(path, dtype=np.uint8)
    off = 0

    def take_f32(n: int) -> np.ndarray:
        nonlocal off
        a = raw[off:off + 4 * n].view(np.float32).copy()
        if a.size != n:
            raise ValueError("model.ncnn.bin truncated")
        off += 4 * n
        return a

    for L in layers:
        if L.type == "Convolution":
            flag = int(raw[off:off + 4].view(np.uint32)[0])
            off += 4
            if flag != 0:
                raise ValueError(f"{L.name}: unsupported weight flag {flag:#x} (only fp32)")
            p = L.params
            cout, kw, kh = p[0], p[1], p.get(11, p[1])
            cin = conv_input_channels(L)
            L.weight = take_f32(p[6]).reshape(cout, cin, kh, kw)
            L.bias = take_f32(cout) if p.get(5, 0) else None
        elif L.type == "MemoryData":
            p = L.params
            w, h, c = p.get(0, 0), p.get(1, 0), p.get(2, 0)
            shape = [d for d in (c, h, w) if d]
            L.weight = take_f32(int(np.prod(shape))).reshape(shape)
    if off != raw.size:
        raise ValueError(f"model.ncnn.bin: {raw.size - off} trailing bytes")


def random_init(layers: List[Layer], seed: int = 0, in_size: int = 640) -> None:
    """Seeded random weights of the architecture the ``.param`` names (used for
    the v2 / TT100K graph whose weights are missing, ``.MISSING_LARGE_BLOBS:10-12``).
    Scale keeps activations O(1) through the SiLU stack; Detect constants
    (strides, anchor points, DFL arange) are the architectural ones."""
    g = np.random.default_rng(seed)
    sizes = [in_size // 8, in_size // 16, in_size // 32]
    anchors, strides = [], []
    for s, n in zip((8, 16, 32), sizes):
        ys, xs = np.meshgrid(np.arange(n) + 0.5, np.arange(n) + 0.5, indexing="ij")
        anchors.append(np.stack([xs.ravel(), ys.ravel()], 0))
        strides.append(np.full(n * n, s, np.float32))
    anchors = np.concatenate(anchors, 1).astype(np.float32)
    strides = np.concatenate(strides).astype(np.float32)
    for L in layers:
        if L.type == "Convolution":
            p = L.params
            cout, kw, kh = p[0], p[1], p.get(11, p[1])
            cin = conv_input_channels(L)
            if not p.get(5, 0) and cout == 1 and cin == 16:          # DFL projection
                L.weight = np.arange(16, dtype=np.float32).reshape(1, 16, 1, 1)
                L.bias = None
                continue
            fan = cin * kh * kw
            L.weight = (g.standard_normal((cout, cin, kh, kw)) * np.sqrt(2.0 / fan)).astype(np.float32)
            L.bias = (g.standard_normal(cout) * 0.1).astype(np.float32) if p.get(5, 0) else None
        elif L.type == "MemoryData":
            p = L.params
            if p.get(1, 0) == 2:
                L.weight = anchors.copy()
            else:
                L.weight = strides.copy()


@torch.no_grad()
def run_graph(layers: List[Layer], x: torch.Tensor, want: Optional[List[str]] = None,
              dtype: torch.dtype = torch.float32, quant=None, quant_for=None) -> Dict[str, torch.Tensor]:
    """Execute the graph on ``x`` [B,3,H,W]; returns {blob name: tensor}.

    Tensors carry a leading batch dim; ncnn axis k maps to torch dim k+1.
    ``want`` = blob names to keep (default: just the last layer's outputs).
    ``quant`` = optional f(tensor)->tensor applied to every conv's input and
    weights (precision-budget experiments, oracle/experiments/precision.py); ``quant_for`` = optional
    f(layer name) -> (activation rounding, weight rounding) for a per-layer budget (precision_budget.py).
    """
    blobs: Dict[str, torch.Tensor] = {}
    keep = set(want or [])
    last = layers[-1].outputs
    for L in layers:
        p = L.params
        t = L.type
        ins = [blobs[n] for n in L.inputs]
        if t == "Input":
            outs = [x.to(dtype)]
        elif t == "Convolution":
            w = torch.from_numpy(L.weight).to(dtype)
            b = torch.from_numpy(L.bias).to(dtype) if L.bias is not None else None
            a = ins[0]
            squeeze = False
            if a.dim() == 3:                       # [B,H,W] treated as 1 channel? no: [B,C,W] -> [B,C,1,W]
                a = a.unsqueeze(2)
                squeeze = True
            if quant is not None and L.bias is not None:       # not the DFL projection
                a, w = quant(a), quant(w)
            if quant_for is not None and L.bias is not None:
                qa, qw = quant_for(L.name)
                a, w = qa(a), qw(w)
            y = F.conv2d(a, w, b, stride=(p.get(13, p.get(3, 1)), p.get(3, 1)),
                         padding=(p.get(14, p.get(4, 0)), p.get(4, 0)),
                         dilation=(p.get(12, p.get(2, 1)), p.get(2, 1)))
            outs = [y.squeeze(2) if squeeze else y]
        elif t == "Swish":
            outs = [ins[0] * torch.sigmoid(ins[0])]
        elif t == "Sigmoid":
            outs = [torch.sigmoid(ins[0])]
        elif t == "Split":
            outs = [ins[0]] * len(L.outputs)
        elif t == "Slice":
            axis = p.get(1, 0) + 1
            sizes = list(p[0])
            total = ins[0].shape[axis]
            known = sum(s for s in sizes if s > 0)
            nfree = sum(1 for s in sizes if s <= 0)
            sizes = [s if s > 0 else (total - known) // nfree for s in sizes]
            outs = list(torch.split(ins[0], sizes, dim=axis))
        elif t == "Concat":
            outs = [torch.cat(ins, dim=p.get(0, 0) + 1)]
        elif t == "BinaryOp":
            op = p.get(0, 0)
            a = ins[0]
            b = ins[1] if len(ins) > 1 else torch.tensor(float(p[2]), dtype=a.dtype)
            if len(ins) > 1 and b.dim() < a.dim():
                b = b.unsqueeze(0)
            if len(ins) > 1 and a.dim() < b.dim():
                a = a.unsqueeze(0)
            outs = [(a + b, a - b, a * b, a / b)[op]]
        elif t == "Pooling":
            if p.get(0, 0) != 0:
                raise NotImplementedError("only max pooling")
            k = (p.get(11, p[1]), p[1])
            s = (p.get(12, p.get(2, 1)), p.get(2, 1))
            pad = (p.get(14, p.get(3, 0)), p.get(3, 0))
            outs = [F.max_pool2d(ins[0], k, s, pad)]
        elif t == "Interp":
            if p.get(0, 0) != 1:
                raise NotImplementedError("only nearest interp")
            outs = [F.interpolate(ins[0], scale_factor=(float(p[1]), float(p[2])), mode="nearest")]
        elif t == "Reshape":
            dims = [p[k] for k in (2, 1, 0) if p.get(k, 0)]
            outs = [ins[0].reshape(ins[0].shape[0], *dims)]
        elif t == "Permute":
            if p.get(0, 0) != 2 or ins[0].dim() != 4:
                raise NotImplementedError("only 3-D permute order 2 (w c h)")
            outs = [ins[0].permute(0, 2, 1, 3).contiguous()]   # new c = old h, new h = old c
        elif t == "Softmax":
            outs = [torch.softmax(ins[0], dim=p.get(0, 0) + 1)]
        elif t == "MemoryData":
            outs = [torch.from_numpy(L.weight).to(dtype).unsqueeze(0)]
        else:
            raise NotImplementedError(f"ncnn layer type {t}")
        for n, o in zip(L.outputs, outs):
            blobs[n] = o
    if want is None:
        return {n: blobs[n] for n in last}
    return {n: blobs[n] for n in list(keep) + list(last)}


class DetectorOracle:
    """fp32 CPU detector = the reference graph + weights.  ``forward`` maps a
    float RGB tensor [B,3,640,640] in [0,1] to ``out0`` [B,5,8400]
    (rows cx,cy,w,h,score in letterbox pixels; e2e.py:244-253 consumes it)."""

    def __init__(self, param_path: str, bin_path: Optional[str] = None, seed: int = 0, in_size: int = 640):
        self.layers = parse_param(param_path)
        if bin_path is not None:
            load_bin(self.layers, bin_path)
        else:
            random_init(self.layers, seed, in_size)

    def forward(self, x, want=None):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        out = run_graph(self.layers, x, want)
        return out if want is not None else out["out0"]

    def convs(self) -> List[Layer]:
        return [L for L in self.layers if L.type == "Convolution"]
