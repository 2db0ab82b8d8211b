"""Reader for the reference's exported detector: ``model.ncnn.param`` + ``model.ncnn.bin``.

Replaces ``ncnn.Net.load_param`` / ``load_model`` (reference
``src/vntsr/pipeline/e2e.py:209-216``).  Only what the YOLO-LitePi graph family
uses is understood: the Convolution records (fp32 weights, OIHW, folded-BN bias)
in file order, and where the C2f blocks' residual adds sit, from which
``plan.py`` derives channel widths and block depths.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class ConvRec:
    name: str
    cout: int
    cin: int
    ksize: int
    stride: int
    pad: int
    has_bias: bool
    silu: bool                     # followed by a Swish layer
    weight: Optional[np.ndarray]   # [cout, cin, k, k] fp32
    bias: Optional[np.ndarray]


@dataclass
class NcnnModel:
    convs: List[ConvRec]
    c2f_depths: List[int]          # bottlenecks per C2f block, graph order
    n_anchors: int
    layer_types: List[str]


def _kv(tokens):
    out = {}
    for t in tokens:
        k, v = t.split("=")
        k = int(k)
        if k <= -23300:
            vals = v.split(",")
            out[-23300 - k] = [float(x) if "." in x or "e" in x else int(x) for x in vals[1:1 + int(vals[0])]]
        else:
            out[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
    return out


def load_ncnn(param_path: str, bin_path: Optional[str] = None, seed: int = 0) -> NcnnModel:
    """Parse the graph; read weights from ``bin_path`` or, when it is None
    (the TT100K export's weights are not in the reference repo), draw seeded
    random weights of the same shapes."""
    try:
        with open(param_path, "r") as f:
            lines = [ln.split() for ln in f if ln.strip()]
    except OSError as e:
        raise RuntimeError(f"Failed to load param: {param_path}") from e
    if not lines or lines[0] != ["7767517"]:
        raise RuntimeError(f"Failed to load param: {param_path} (bad magic)")
    raw = None
    if bin_path is not None:
        try:
            raw = np.fromfile(bin_path, dtype=np.uint8)
        except OSError as e:
            raise RuntimeError(f"Failed to load bin: {bin_path}") from e
    rng = np.random.default_rng(seed)
    off = 0
    convs: List[ConvRec] = []
    depths: List[int] = []
    types: List[str] = []
    n_anchors = 0
    in_c2f = False
    head_started = False               # the first MemoryData (strides) precedes the Detect head
    body = lines[2:]
    for i, tok in enumerate(body):
        ltype, name, n_in, n_out = tok[0], tok[1], int(tok[2]), int(tok[3])
        p = _kv(tok[4 + n_in + n_out:])
        types.append(ltype)
        if ltype == "Convolution":
            cout, k = p[0], p[1]
            cin = p[6] // (cout * k * k)
            has_bias = bool(p.get(5, 0))
            w = b = None
            if raw is not None:
                flag = int(raw[off:off + 4].view(np.uint32)[0]); off += 4
                if flag != 0:
                    raise RuntimeError(f"Failed to load bin: {bin_path} ({name}: non-fp32 weights)")
                w = raw[off:off + 4 * p[6]].view(np.float32).reshape(cout, cin, k, k).copy(); off += 4 * p[6]
                if has_bias:
                    b = raw[off:off + 4 * cout].view(np.float32).copy(); off += 4 * cout
            elif has_bias:
                w = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
                b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            else:                                   # DFL projection: arange(16), architectural constant
                w = np.arange(cin, dtype=np.float32).reshape(1, cin, 1, 1)
            silu = i + 1 < len(body) and body[i + 1][0] == "Swish"
            convs.append(ConvRec(name, cout, cin, k, p.get(3, 1), p.get(4, 0), has_bias, silu, w, b))
        elif ltype == "MemoryData":
            n = 1
            for key in (0, 1, 2):
                if p.get(key, 0):
                    n *= p[key]
            if raw is not None:
                off += 4 * n
            n_anchors = max(n_anchors, p.get(0, 0))
            head_started = True
        elif ltype == "Slice" and not in_c2f and not head_started:
            # C2f: Slice (chunk 2) ... one BinaryOp add per bottleneck ... Concat
            if p.get(1, 0) == 0 and len(p.get(0, [])) == 2:
                in_c2f = True
                depths.append(0)
        elif ltype == "BinaryOp" and in_c2f and p.get(0, 0) == 0:
            depths[-1] += 1
        elif ltype == "Concat" and in_c2f:
            in_c2f = False
    if raw is not None and off != raw.size:
        raise RuntimeError(f"Failed to load bin: {bin_path} ({raw.size - off} trailing bytes)")
    return NcnnModel(convs, depths, n_anchors, types)
