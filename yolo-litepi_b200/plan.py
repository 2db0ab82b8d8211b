"""Lowers the two networks of the hot path to flat launch plans (``lp_op_desc`` lists).

Detector: the reference's exported graph (``model.ncnn.param``; Ultralytics YAML at
``train_model/train-yolo-custom-vntsr..ipynb:1918-1960``) is a YOLOv8-family template:
Conv, Conv, C2f, Conv, C2f, Conv, C2f, Conv, C2f, SPPF, (Upsample, Concat, C2f) x2,
(Conv, Concat, C2f) x2, Detect.  Widths and C2f depths are read from the file; the
template consumes the Convolution records in file order and checks every shape.

Everything that is a view in the graph stays a view here: C2f chunk/concat, SPPF
concat, the neck concats and ShuffleNetV2's chunk + channel_shuffle become channel
offsets / strides of the producing kernel's store (no concat or shuffle kernels).

Classifier: torchvision ``shufflenet_v2_x1_0`` (``e2e.py:331-333``) from its
``state_dict``, BatchNorm folded into the convs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .ncnn_model import ConvRec, NcnnModel


def p8(c: int) -> int:
    return (c + 7) // 8 * 8


def pad_to(c: int, g: int) -> int:
    return (c + g - 1) // g * g


@dataclass
class View:
    """A run of logical channel segments inside one buffer; segment i holds ``lens[i]``
    real channels starting at physical channel ``offs[i]`` (padded to 8 for split-f16)."""
    buf: int
    offs: Tuple[int, ...]
    lens: Tuple[int, ...]
    pad8: bool = True
    gran: int = 8                       # padding granularity of a segment when pad8 (8 detector, 16 classifier)

    @property
    def start(self) -> int:
        return self.offs[0]

    @property
    def phys(self) -> int:
        last = self.lens[-1]
        return self.offs[-1] + (pad_to(last, self.gran) if self.pad8 else last) - self.offs[0]

    @property
    def logical(self) -> int:
        return sum(self.lens)

    def chan_map(self) -> np.ndarray:
        """physical index (relative to start) of every logical channel"""
        idx = []
        for o, n in zip(self.offs, self.lens):
            idx.extend(range(o - self.offs[0], o - self.offs[0] + n))
        return np.asarray(idx, dtype=np.int64)

    def seg(self, i: int, j: Optional[int] = None) -> "View":
        j = i + 1 if j is None else j
        return View(self.buf, self.offs[i:j], self.lens[i:j], self.pad8, self.gran)


@dataclass
class Plan:
    bufs: List[dict] = field(default_factory=list)
    ops: List[dict] = field(default_factory=list)
    blobs: List[np.ndarray] = field(default_factory=list)
    n_floats: int = 0
    names: List[str] = field(default_factory=list)
    macs: List[int] = field(default_factory=list)          # per op, per image
    meta: Dict[str, object] = field(default_factory=dict)

    # ---- buffers
    def buf(self, h: int, w: int, c: int, fmt: int) -> int:
        self.bufs.append(dict(h=h, w=w, c=c, fmt=fmt))
        return len(self.bufs) - 1

    def new_view(self, h: int, w: int, lens: Sequence[int], fmt: int = L.FMT_SPLIT16, gran: int = 8) -> View:
        pad8 = fmt == L.FMT_SPLIT16
        offs, o = [], 0
        for n in lens:
            offs.append(o)
            o += pad_to(n, gran) if pad8 else n
        return View(self.buf(h, w, o, fmt), tuple(offs), tuple(lens), pad8, gran)

    # ---- weights
    def _push(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        off = self.n_floats
        self.blobs.append(a)
        self.n_floats += a.size
        pad = (-self.n_floats) % 4                      # keep every tensor 16-B aligned
        if pad:
            self.blobs.append(np.zeros(pad, np.float32))
            self.n_floats += pad
        return off

    def weights(self) -> np.ndarray:
        return np.concatenate(self.blobs) if self.blobs else np.zeros(0, np.float32)

    # ---- ops
    def conv(self, name: str, w: np.ndarray, b: Optional[np.ndarray], src: View, dst: View, stride: int, act: int,
             res: Optional[View] = None, row_off: int = 0, out_cstride: int = 1, kind: int = L.OP_CONV,
             in_mean: float = 0.0, in_std: float = 1.0, seg: Optional[Tuple[int, int, int]] = None, flags: int = 0) -> None:
        """w: [cout, cin, k, k] (logical channels) -> packed [tap][cin_phys][cout_phys], zero padded.
        ``seg`` = (logical offset, segment length, padded segment length): output j goes to LOGICAL channel
        off + j*out_cstride of a buffer whose logical channels are stored in padded segments."""
        cout, cin, k, _ = w.shape
        assert cin == src.logical, f"{name}: cin {cin} != view {src.logical}"
        assert seg is not None or cout == dst.logical, f"{name}: cout {cout} != view {dst.logical}"
        hi, wi = self.bufs[src.buf]["h"], self.bufs[src.buf]["w"]
        cinp = src.phys if kind != L.OP_STEM_U8 else 3
        if seg is not None:
            coutp = pad_to(cout, 16)                   # N padding for the tensor core; only `cout` are stored
        else:
            coutp = dst.phys if out_cstride == 1 else cout
        wp = np.zeros((k * k, cinp, coutp), np.float32)
        im = src.chan_map() if kind != L.OP_STEM_U8 else np.arange(3)
        om = dst.chan_map() if (out_cstride == 1 and seg is None) else np.arange(cout)
        wp[:, im[:, None], om[None, :]] = w.transpose(2, 3, 1, 0).reshape(k * k, cin, cout)
        bp = np.zeros(coutp, np.float32)
        if b is not None:
            bp[om] = b
        if res is not None:
            assert res.logical == cout and res.phys == coutp
        ho = (hi + 2 * (k // 2) - k) // stride + 1
        wo = (wi + 2 * (k // 2) - k) // stride + 1
        self.ops.append(dict(kind=kind, in_buf=src.buf, in_coff=src.start if kind != L.OP_STEM_U8 else 0, cin=cinp,
                             out_buf=dst.buf, out_coff=seg[0] if seg else dst.start, cout=coutp, out_cstride=out_cstride,
                             cout_real=cout if seg else coutp, out_seg_len=seg[1] if seg else 0,
                             out_seg_pad=seg[2] if seg else 0,
                             res_buf=res.buf if res is not None else -1, res_coff=res.start if res is not None else 0,
                             ksize=k, stride=stride, act=act, row_off=row_off, flags=flags, in_mean=in_mean, in_std=in_std,
                             w_off=self._push(wp), b_off=self._push(bp), wtc_off=-1))
        self.names.append(name)
        self.macs.append(ho * wo * k * k * cin * cout)

    def simple(self, kind: int, name: str, src: View, dst: View, ksize: int = 1, stride: int = 1,
               out_cstride: int = 1, w: Optional[np.ndarray] = None, b: Optional[np.ndarray] = None,
               act: int = L.ACT_NONE, seg: Optional[Tuple[int, int, int]] = None, res: Optional[View] = None) -> None:
        n = src.phys if (out_cstride == 1 and seg is None) else src.logical
        w_off = self._push(w) if w is not None else 0
        b_off = self._push(b) if b is not None else 0
        self.ops.append(dict(kind=kind, in_buf=src.buf, in_coff=src.start, cin=n, out_buf=dst.buf,
                             out_coff=seg[0] if seg else dst.start,
                             cout=n, out_cstride=out_cstride, cout_real=n, out_seg_len=seg[1] if seg else 0,
                             out_seg_pad=seg[2] if seg else 0, res_buf=res.buf if res is not None else -1,
                             res_coff=res.start if res is not None else 0, ksize=ksize, stride=stride,
                             act=act, row_off=0, flags=0, in_mean=0.0, in_std=1.0, w_off=w_off, b_off=b_off, wtc_off=-1))
        self.names.append(name)
        self.macs.append(0)

    def mean_fc(self, src: View, fcw: np.ndarray, fcb: np.ndarray) -> None:
        """global mean over HxW + Linear (torchvision classifiers' avgpool + flatten + fc); fcw [classes, channels]."""
        cin_phys = self.bufs[src.buf]["c"]
        wp = np.zeros((cin_phys, fcw.shape[0]), np.float32)
        wp[src.chan_map() + src.start] = fcw.T
        self.ops.append(dict(kind=L.OP_MEAN_FC, in_buf=src.buf, in_coff=0, cin=cin_phys, out_buf=-1, out_coff=0,
                             cout=fcw.shape[0], out_cstride=1, cout_real=fcw.shape[0], out_seg_len=0, out_seg_pad=0,
                             res_buf=-1, res_coff=0, ksize=1, stride=1, act=L.ACT_NONE, row_off=0, flags=0,
                             in_mean=0.0, in_std=1.0, w_off=self._push(wp), b_off=self._push(fcb), wtc_off=-1))
        self.names.append("mean_fc")
        self.macs.append(int(fcw.size))

    # ---- finalisation
    def layout(self, max_batch: int) -> int:
        """Assign workspace offsets; returns the workspace size in bytes."""
        off = 0
        for b in self.bufs:
            esz = {L.FMT_SPLIT16: 2, L.FMT_F32: 4, L.FMT_U8: 1}[b["fmt"]]
            b["image_bytes"] = b["h"] * b["w"] * b["c"] * esz
            if b["fmt"] == L.FMT_U8:
                b["offset"] = 0
                continue
            b["offset"] = off
            planes = 2 if b["fmt"] == L.FMT_SPLIT16 else 1
            off += planes * max_batch * b["image_bytes"]
            off = (off + 1023) // 1024 * 1024
        return off

    def pack_tc_weights(self) -> np.ndarray:
        """Pre-split, pre-packed tensor-core operands for every conv the tcgen05 kernel can run
        (csrc/conv_tc.cu ``lp_conv_tc_try`` applies the same eligibility test).  Per K-block
        (tap x <=64 input channels) the stage image is [8-channel chunk][plane hi|lo][cout][8] fp16 --
        the UMMA K-major no-swizzle canonical layout, so one bulk copy lands a ready B operand.
        Sets ``wtc_off`` of the eligible ops; returns the blob as uint8."""
        W = self.weights()
        parts, off = [], 0
        for op in self.ops:
            op["wtc_off"] = -1
            if op["kind"] != L.OP_CONV:
                continue
            seg = op["out_seg_len"] > 0
            if (op["ksize"], op["stride"]) not in ((1, 1), (3, 1), (3, 2)) or (op["out_cstride"] != 1 and not seg):
                continue
            cin, cout = op["cin"], op["cout"]
            if self.bufs[op["in_buf"]]["fmt"] != L.FMT_SPLIT16 or cin % 16 or cout % 16 or cin > 512:
                continue
            if seg and self.bufs[op["out_buf"]]["fmt"] != L.FMT_SPLIT16:
                continue
            kb = next((d for d in (64, 48, 32, 16) if cin % d == 0), 0)      # channels per K-block
            if not kb:
                continue
            taps, ncb = op["ksize"] ** 2, cin // kb
            w = W[op["w_off"]:op["w_off"] + taps * cin * cout].reshape(taps, cin, cout)
            hi = w.astype(np.float16)
            lo = (w - hi.astype(np.float32)).astype(np.float16)
            op["wtc_off"] = off
            for n0 in range(0, cout, 128):                                  # output blocks of <= 128 channels
                nb = min(128, cout - n0)
                planes = [a[:, :, n0:n0 + nb].reshape(taps, ncb, kb // 8, 8, nb).transpose(0, 1, 2, 4, 3) for a in (hi, lo)]
                raw = np.ascontiguousarray(np.stack(planes, axis=3)).view(np.uint8).ravel()   # [tap][cb][chunk][plane][n][8]
                parts.append(raw)
                off += raw.size
            pad = (-off) % 128
            if pad:
                parts.append(np.zeros(pad, np.uint8))
                off += pad
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)

    def c_arrays(self):
        bufs = (L.BufDesc * len(self.bufs))()
        for i, b in enumerate(self.bufs):
            bufs[i] = L.BufDesc(b["h"], b["w"], b["c"], b["fmt"], b["offset"], b["image_bytes"])
        ops = (L.OpDesc * len(self.ops))()
        keys = [f[0] for f in L.OpDesc._fields_]
        for i, o in enumerate(self.ops):
            ops[i] = L.OpDesc(*[o[k] for k in keys])
        return bufs, ops


# ============================================================================ detector
def build_detector_plan(model: NcnnModel, in_size: int = 640) -> Plan:
    P = Plan()
    convs = list(model.convs)
    depths = list(model.c2f_depths)
    if len(depths) != 8:
        raise RuntimeError(f"unsupported detector graph: expected 8 C2f blocks, found {len(depths)}")
    it = iter(convs)

    def nxt(k: int, s: int, cin: Optional[int] = None) -> ConvRec:
        c = next(it)
        if c.ksize != k or c.stride != s or (cin is not None and c.cin != cin):
            raise RuntimeError(f"unsupported detector graph at {c.name}: got k{c.ksize} s{c.stride} cin{c.cin}, "
                               f"template expects k{k} s{s} cin{cin}")
        return c

    SILU, NONE = L.ACT_SILU, L.ACT_NONE

    def conv(c: ConvRec, src: View, dst: View, res=None, row_off=0, kind=L.OP_CONV):
        P.conv(c.name, c.weight, c.bias, src, dst, c.stride, SILU if c.silu else NONE, res, row_off, kind=kind)

    def c2f(src: View, n: int, dst_of) -> View:
        """cv1 -> [a|b]; b_{i+1} = b_i + cvb(cva(b_i)); cv2(cat[a, b_0..b_n]) -> dst.
        ``dst_of(cout)`` returns the destination view."""
        cv1 = nxt(1, 1, src.logical)
        c = cv1.cout // 2
        h, w = P.bufs[src.buf]["h"], P.bufs[src.buf]["w"]
        cat = P.new_view(h, w, [c] * (2 + n))
        conv(cv1, src, cat.seg(0, 2))
        for i in range(n):
            cva, cvb = nxt(3, 1, c), nxt(3, 1, c)
            tmp = P.new_view(h, w, [c])
            conv(cva, cat.seg(1 + i), tmp)
            conv(cvb, tmp, cat.seg(2 + i), res=cat.seg(1 + i))
        cv2 = nxt(1, 1, (2 + n) * c)
        dst = dst_of(cv2.cout)
        conv(cv2, cat, dst)
        return dst

    S = in_size
    img = View(P.buf(S, S, 3, L.FMT_U8), (0,), (3,), False)
    # ---- backbone
    c0 = nxt(3, 2, 3)
    t0 = P.new_view(S // 2, S // 2, [c0.cout])
    conv(c0, img, t0, kind=L.OP_STEM_U8)
    c1 = nxt(3, 2, c0.cout)
    t1 = P.new_view(S // 4, S // 4, [c1.cout])
    conv(c1, t0, t1)
    t2 = c2f(t1, depths[0], lambda co: P.new_view(S // 4, S // 4, [co]))
    c3 = nxt(3, 2, t2.logical)
    t3 = P.new_view(S // 8, S // 8, [c3.cout])
    conv(c3, t2, t3)
    # model.4 output lands in the P3 neck concat [up(model.12) | model.4]; widths are only known after
    # walking further, so peek: model.12's width = cv2 cout of the 6th C2f.  Build lazily via placeholders.
    # Simpler: first compute widths by a dry walk over the conv list.
    widths = _dry_widths(convs, depths)
    cat_p3 = P.new_view(S // 8, S // 8, [widths["m12"], widths["m4"]])      # cat_7 = [up(m12), m4]
    cat_p4 = P.new_view(S // 16, S // 16, [widths["m9"], widths["m6"]])     # cat_5 = [up(m9), m6]
    m4 = c2f(t3, depths[1], lambda co: cat_p3.seg(1))
    c5 = nxt(3, 2, m4.logical)
    t5 = P.new_view(S // 16, S // 16, [c5.cout])
    conv(c5, m4, t5)
    m6 = c2f(t5, depths[2], lambda co: cat_p4.seg(1))
    c7 = nxt(3, 2, m6.logical)
    t7 = P.new_view(S // 32, S // 32, [c7.cout])
    conv(c7, m6, t7)
    m8 = c2f(t7, depths[3], lambda co: P.new_view(S // 32, S // 32, [co]))
    # ---- SPPF: cv1 -> [x | mp(x) | mp(mp(x)) | mp^3(x)] -> cv2
    s1 = nxt(1, 1, m8.logical)
    sp = P.new_view(S // 32, S // 32, [s1.cout] * 4)
    conv(s1, m8, sp.seg(0))
    for i in range(3):
        P.simple(L.OP_MAXPOOL, f"sppf.maxpool{i}", sp.seg(i), sp.seg(i + 1), ksize=5, stride=1)
    s2 = nxt(1, 1, 4 * s1.cout)
    cat_p5n = P.new_view(S // 32, S // 32, [widths["m19"], s2.cout])        # cat_11 = [m19, m9]
    m9 = cat_p5n.seg(1)
    conv(s2, sp, m9)
    # ---- top-down
    P.simple(L.OP_UPSAMPLE2, "upsample.m10", m9, cat_p4.seg(0))
    cat_p4n = P.new_view(S // 16, S // 16, [widths["m16"], widths["m12"]])  # cat_9 = [m16, m12]
    m12 = c2f(cat_p4, depths[4], lambda co: cat_p4n.seg(1))
    P.simple(L.OP_UPSAMPLE2, "upsample.m13", m12, cat_p3.seg(0))
    p3 = c2f(cat_p3, depths[5], lambda co: P.new_view(S // 8, S // 8, [co]))
    # ---- bottom-up
    c16 = nxt(3, 2, p3.logical)
    conv(c16, p3, cat_p4n.seg(0))
    p4 = c2f(cat_p4n, depths[6], lambda co: P.new_view(S // 16, S // 16, [co]))
    c19 = nxt(3, 2, p4.logical)
    conv(c19, p4, cat_p5n.seg(0))
    p5 = c2f(cat_p5n, depths[7], lambda co: P.new_view(S // 32, S // 32, [co]))
    # ---- Detect: per level box (3x3, 3x3, 1x1 -> 64 raw) and cls (3x3, 3x3, 1x1 -> nc raw)
    n_anchors = sum((S // s) ** 2 for s in (8, 16, 32))
    head_specs = []
    row = 0
    for lvl, feat in enumerate((p3, p4, p5)):
        h = P.bufs[feat.buf]["h"]
        b0, b1, b2 = nxt(3, 1, feat.logical), nxt(3, 1), nxt(1, 1)
        k0, k1, k2 = nxt(3, 1, feat.logical), nxt(3, 1), nxt(1, 1)
        if b2.cout != 64 or b2.silu or k2.silu:
            raise RuntimeError("unsupported Detect head (reg_max must be 16)")
        head_specs.append((feat, h, (b0, b1, b2), (k0, k1, k2), row))
        row += h * h
    nc = head_specs[0][3][2].cout
    hc = 64 + (nc + 3) // 4 * 4
    # buffers for the branches first, the head buffer LAST (lp_detect_forward reads bufs.back())
    # The first conv of the box branch and of the class branch read the same feature map with the same
    # geometry: they run as ONE conv with concatenated output channels (one patch load, one launch, wider N).
    staged = []
    for feat, h, box, cls, row in head_specs:
        first = P.new_view(h, h, [box[0].cout, cls[0].cout])
        hb2, hc2 = P.new_view(h, h, [box[1].cout]), P.new_view(h, h, [cls[1].cout])
        staged.append((feat, box, cls, row, first, hb2, hc2))
    head = P.buf(n_anchors, 1, hc, L.FMT_F32)
    for feat, box, cls, row, first, hb2, hc2 in staged:
        b0, k0 = box[0], cls[0]
        if b0.silu == k0.silu and b0.has_bias == k0.has_bias:
            both = ConvRec(f"{b0.name}+{k0.name}", b0.cout + k0.cout, b0.cin, 3, 1, b0.pad, b0.has_bias, b0.silu,
                           np.concatenate([b0.weight, k0.weight], 0),
                           np.concatenate([b0.bias, k0.bias]) if b0.has_bias else None)
            conv(both, feat, first)
        else:
            conv(b0, feat, first.seg(0))
            conv(k0, feat, first.seg(1))
        conv(box[1], first.seg(0), hb2)
        conv(box[2], hb2, View(head, (0,), (64,), False), row_off=row)
        conv(cls[1], first.seg(1), hc2)
        conv(cls[2], hc2, View(head, (64,), (nc,), False), row_off=row)
    dfl = next(it)
    if dfl.has_bias or dfl.cin != 16 or not np.array_equal(dfl.weight.ravel(), np.arange(16, dtype=np.float32)):
        raise RuntimeError("unsupported Detect head: DFL projection is not arange(16)")
    if next(it, None) is not None:
        raise RuntimeError("unsupported detector graph: trailing convolutions")
    P.meta.update(n_anchors=n_anchors, nc=nc, head_c=hc, in_size=S,
                  widths=[c0.cout, c1.cout, c3.cout, c5.cout, c7.cout])
    return P


def _dry_widths(convs: List[ConvRec], depths: List[int]) -> Dict[str, int]:
    """Output widths of the modules whose results are concat operands (model.4/6/9/12/16/19)."""
    i = 2                                   # after model.0, model.1
    out = {}

    def skip_c2f(n):
        nonlocal i
        i += 1 + 2 * n
        w = convs[i].cout
        i += 1
        return w

    skip_c2f(depths[0])                     # model.2
    i += 1                                  # model.3
    out["m4"] = skip_c2f(depths[1])
    i += 1                                  # model.5
    out["m6"] = skip_c2f(depths[2])
    i += 1                                  # model.7
    skip_c2f(depths[3])                     # model.8
    i += 1                                  # sppf cv1
    out["m9"] = convs[i].cout
    i += 1
    out["m12"] = skip_c2f(depths[4])
    skip_c2f(depths[5])                     # model.15
    out["m16"] = convs[i].cout
    i += 1
    skip_c2f(depths[6])                     # model.18
    out["m19"] = convs[i].cout
    return out


# ============================================================================ classifier
def _fold_bn(w: np.ndarray, sd: dict, bn: str, eps: float = 1e-5):
    g = sd[bn + ".weight"].astype(np.float64)
    beta = sd[bn + ".bias"].astype(np.float64)
    mu = sd[bn + ".running_mean"].astype(np.float64)
    var = sd[bn + ".running_var"].astype(np.float64)
    s = g / np.sqrt(var + eps)
    return (w.astype(np.float64) * s.reshape(-1, 1, 1, 1)).astype(np.float32), (beta - mu * s).astype(np.float32)


def build_classifier_plan(state_dict: dict, in_size: int = 64, mean: float = 0.18, std: float = 0.34) -> Plan:
    """torchvision ShuffleNetV2 x1.0 (shufflenetv2.py) -> plan, split-f16 activations so that the 1x1
    convs run on the tensor-core kernel.

    A stage tensor with 2h logical channels is stored as two halves of h channels, each padded to a
    multiple of 16 (58->64, 116->128, 232->240): logical channel l lives at physical (l // h) * hp + l % h.
    ``chunk`` is then a contiguous physical view, and ``channel_shuffle(cat[a, b], 2)`` (a to even, b to
    odd logical channels) is the store pattern of the two producers (lp_op_desc.out_seg_len/out_seg_pad).
    Padding channels are never written and stay zero (the workspace is zero-initialised)."""
    sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state_dict.items()}
    P = Plan()
    RELU, NONE, G = L.ACT_RELU, L.ACT_NONE, 16
    S = in_size
    img = View(P.buf(S, S, 3, L.FMT_U8), (0,), (3,), False)
    w, b = _fold_bn(sd["conv1.0.weight"], sd, "conv1.1")
    h = S // 2
    x = P.new_view(h, h, [w.shape[0]], gran=G)
    P.conv("conv1", w, b, img, x, 2, RELU, kind=L.OP_STEM_U8, in_mean=mean, in_std=std)
    h //= 2
    y = P.new_view(h, h, [x.logical], gran=G)
    P.simple(L.OP_MAXPOOL, "maxpool", x, y, ksize=3, stride=2)
    x = y                                              # View: segments of the current feature map

    def dw(name, prefix_conv, prefix_bn, src, stride):
        wdw, bdw = _fold_bn(sd[prefix_conv + ".weight"], sd, prefix_bn)        # [C,1,3,3]
        c = wdw.shape[0]
        assert c == src.logical
        hs = P.bufs[src.buf]["h"]
        ho = (hs + 2 - 3) // stride + 1
        dst = P.new_view(ho, ho, list(src.lens), gran=G)                          # same segment structure
        cm = src.chan_map()
        wp = np.zeros((9, src.phys), np.float32)
        bp = np.zeros(src.phys, np.float32)
        wp[:, cm] = wdw.reshape(c, 9).T
        bp[cm] = bdw
        P.simple(L.OP_DWCONV3, name, src, dst, ksize=3, stride=stride, w=wp, b=bp)
        P.macs[-1] = ho * ho * 9 * c
        return dst

    def pw(name, prefix_conv, prefix_bn, src, dst, seg=None):
        wp, bp = _fold_bn(sd[prefix_conv + ".weight"], sd, prefix_bn)
        P.conv(name, wp, bp, src, dst, 1, RELU, out_cstride=2 if seg else 1, seg=seg)

    stage = 2
    while f"stage{stage}.0.branch2.0.weight" in sd:
        u = 0
        while f"stage{stage}.{u}.branch2.0.weight" in sd:
            pre = f"stage{stage}.{u}"
            hh = P.bufs[x.buf]["h"]
            if u == 0:                                  # down-sampling unit: both branches see all channels
                bf = sd[pre + ".branch2.5.weight"].shape[0]
                ho = (hh + 2 - 3) // 2 + 1
                out = P.new_view(ho, ho, [bf, bf], gran=G)
                hp = pad_to(bf, G)
                t = dw(pre + ".branch1.dw", pre + ".branch1.0", pre + ".branch1.1", x, 2)
                pw(pre + ".branch1.pw", pre + ".branch1.2", pre + ".branch1.3", t, out, seg=(0, bf, hp))
                t = P.new_view(hh, hh, [bf], gran=G)
                pw(pre + ".branch2.pw1", pre + ".branch2.0", pre + ".branch2.1", x, t)
                t = dw(pre + ".branch2.dw", pre + ".branch2.3", pre + ".branch2.4", t, 2)
                pw(pre + ".branch2.pw2", pre + ".branch2.5", pre + ".branch2.6", t, out, seg=(1, bf, hp))
            else:                                       # basic unit: x1 passes through, x2 -> branch2
                bf = x.logical // 2
                hp = pad_to(bf, G)
                out = P.new_view(hh, hh, [bf, bf], gran=G)
                x1, x2 = x.seg(0), x.seg(1)
                P.simple(L.OP_COPY, pre + ".passthrough", x1, out, out_cstride=2, seg=(0, bf, hp))
                t = P.new_view(hh, hh, [bf], gran=G)
                pw(pre + ".branch2.pw1", pre + ".branch2.0", pre + ".branch2.1", x2, t)
                t = dw(pre + ".branch2.dw", pre + ".branch2.3", pre + ".branch2.4", t, 1)
                pw(pre + ".branch2.pw2", pre + ".branch2.5", pre + ".branch2.6", t, out, seg=(1, bf, hp))
            x = out
            u += 1
        stage += 1
    w5, b5 = _fold_bn(sd["conv5.0.weight"], sd, "conv5.1")
    hh = P.bufs[x.buf]["h"]
    y = P.new_view(hh, hh, [w5.shape[0]], gran=G)
    P.conv("conv5", w5, b5, x, y, 1, RELU)
    fcw, fcb = sd["fc.weight"], sd["fc.bias"]                       # [C, 1024]
    P.ops.append(dict(kind=L.OP_MEAN_FC, in_buf=y.buf, in_coff=0, cin=fcw.shape[1], out_buf=-1, out_coff=0,
                      cout=fcw.shape[0], out_cstride=1, cout_real=fcw.shape[0], out_seg_len=0, out_seg_pad=0,
                      res_buf=-1, res_coff=0, ksize=1, stride=1, act=NONE,
                      row_off=0, flags=0, in_mean=0.0, in_std=1.0, w_off=P._push(fcw.T.copy()), b_off=P._push(fcb), wtc_off=-1))
    P.names.append("mean_fc")
    P.macs.append(fcw.size)
    P.meta.update(num_classes=int(fcw.shape[0]), in_size=S)
    return P


# ============================================================================ fused classifier
FS_CONV1, FS_MAXPOOL, FS_PW, FS_DW, FS_COPY, FS_MEANFC, FS_STORE, FS_LOAD = range(8)
FSTEP_WORDS = 18        # struct FStep in csrc/shufflenet_fused.cu


@dataclass
class FusedProgram:
    """What csrc/shufflenet_fused.cu executes: three step lists in one array.  front (per ROI: conv1, max pool,
    stage2 unit 0), middle (per ROI: stage2 units 1-3, stage3; tensor-core pointwise layers; ends by parking the
    4x4x232 result in global memory), tail (the CTA's `tail_group` ROIs stacked as rows: stage4, conv5, mean, fc)."""
    steps: np.ndarray          # int32 [n, 18] (struct FStep)
    weights: np.ndarray        # fp32 blob
    weights16: np.ndarray      # split-f16 pointwise weights of the middle
    n_front: int
    n_mid: int
    n_tail: int
    tail_group: int
    smem_bytes: int            # extent of the front/middle activation map
    back_bytes: int            # extent the middle still uses (behind it: fp16 staging + weight stages)
    astage_bytes: int
    tail_bytes: int            # extent of the tail's activation map (behind it: weight stages)
    park_floats: int           # floats per ROI parked in global memory between middle and tail
    tail_astage_bytes: int = 0 # fp16 activation staging of the tail's tensor-core layers (0: fp32 FMA tail)


def build_fused_classifier(state_dict: dict, group: int = 1, in_size: int = 64, tail_group: int = 3,
                           tail_mma: bool = False) -> FusedProgram:
    """Step list + fp32 weight blob for the persistent fused ShuffleNetV2 kernel
    (csrc/shufflenet_fused.cu).  Returns a FusedProgram.

    Shared-memory map (floats): Y = G x 7424 (stage tensor A) | scratch.  Back end: B = G x 7424,
    T1 = G x 7424, T2 = G x 3712.  Front end (per ROI) overlays the scratch: u8 crop, conv1 output,
    pooled tensor, and -- once conv1's output is dead -- the stage2.0 intermediates."""
    if in_size != 64:
        raise ValueError("the fused classifier is laid out for 64x64 inputs")
    sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state_dict.items()}
    G = int(group)
    if G != 1:
        raise ValueError("front end and middle run one ROI at a time; stack ROIs with tail_group")
    blobs: List[np.ndarray] = []
    nf = [0]

    def push(a):
        a = np.ascontiguousarray(a, np.float32).ravel()
        off = nf[0]
        blobs.append(a)
        nf[0] += a.size
        pad = (-nf[0]) % 4
        if pad:
            blobs.append(np.zeros(pad, np.float32)); nf[0] += pad
        return off

    h16: List[np.ndarray] = []
    nh = [0]
    last16 = [0]
    frag_mode = [False]     # True while the tail is emitted with tail_mma: fp16 weights in MMA-fragment order (read straight from L2)

    def pw_w(conv, bn):
        w, b = _fold_bn(sd[conv + ".weight"], sd, bn)                # [cout, cin, 1, 1]
        cout, cin = w.shape[:2]
        cp = (cout + 3) // 4 * 4
        wp = np.zeros((cin, cp), np.float32); wp[:, :cout] = w.reshape(cout, cin).T
        bp = np.zeros(cp, np.float32); bp[:cout] = b
        # tensor-core copy: [cout_p8][hi|lo][L] fp16, L = cin_p16 + pad with L == 4 (mod 32) (csrc w16_row_halves)
        last16[0] = 0
        if frag_mode[0]:
            # [n_tile][k_step][lane][bh0, bh1, bl0, bl1]: lane (g, t) holds B[k = 16*ks + 2t (+1, +8, +9)][n = 8*nt + g] of the hi and
            # of the lo plane -- the operand registers of mma.m16n8k16, so a warp loads one coalesced 512-byte line per (n_tile, k_step)
            cin_p = (cin + 15) // 16 * 16
            cout_p8 = (cout + 7) // 8 * 8
            w2 = np.zeros((cout_p8, cin_p), np.float32); w2[:cout, :cin] = w.reshape(cout, cin)
            hi = w2.astype(np.float16)
            lo = (w2 - hi.astype(np.float32)).astype(np.float16)
            nt_i, ks_i, ln_i = np.meshgrid(np.arange(cout_p8 // 8), np.arange(cin_p // 16), np.arange(32), indexing="ij")
            n_i = nt_i * 8 + ln_i // 4
            k_i = ks_i * 16 + 2 * (ln_i % 4)
            blk = np.stack([hi[n_i, k_i], hi[n_i, k_i + 1], hi[n_i, k_i + 8], hi[n_i, k_i + 9],
                            lo[n_i, k_i], lo[n_i, k_i + 1], lo[n_i, k_i + 8], lo[n_i, k_i + 9]], axis=-1)
            assert nh[0] % 8 == 0
            last16[0] = nh[0] // 8 + 1
            h16.append(np.ascontiguousarray(blk, np.float16).ravel()); nh[0] += blk.size
        elif True:                      # every pointwise layer gets an fp16 copy; which ones use it is decided per pass
            cin_p = (cin + 15) // 16 * 16
            L = cin_p + ((4 - cin_p) % 32)
            w2 = w.reshape(cout, cin).astype(np.float32)
            hi = w2.astype(np.float16)
            lo = (w2 - hi.astype(np.float32)).astype(np.float16)
            blk = np.zeros(((cout + 7) // 8 * 8, 2, L), np.float16)
            blk[:cout, 0, :cin] = hi
            blk[:cout, 1, :cin] = lo
            assert nh[0] % 8 == 0
            last16[0] = nh[0] // 8 + 1                                # units of 16 B, +1 so that 0 means "none"
            h16.append(blk.ravel()); nh[0] += blk.size
            pad = (-nh[0]) % 8
            if pad:
                h16.append(np.zeros(pad, np.float16)); nh[0] += pad
        return push(wp), push(bp), cin, cout

    def dw_w(conv, bn):
        w, b = _fold_bn(sd[conv + ".weight"], sd, bn)                # [C,1,3,3]
        c = w.shape[0]
        return push(w.reshape(c, 9).T.copy()), push(b), c

    steps: List[List[int]] = []

    def step(op, src, dst, src_C=0, src_off=0, dst_C=0, dst_off=0, dst_cs=1, cin=0, cout=0, H=0, W=0, stride=1, relu=0,
             w_off=0, b_off=0, roi_stride=0):
        steps.append([op, src, dst, src_C, src_off, dst_C, dst_off, dst_cs, cin, cout, H, W, stride, relu, w_off, b_off,
                      roi_stride, last16[0] if op == FS_PW else 0])
        last16[0] = 0

    c1 = sd["conv1.0.weight"].shape[0]                                # 24
    widths = [sd[f"stage{s}.0.branch2.5.weight"].shape[0] for s in (2, 3, 4)]   # 58, 116, 232
    c5 = sd["conv5.0.weight"].shape[0]
    ncls = sd["fc.weight"].shape[0]
    ymax = 8 * 8 * 2 * widths[0]                                      # 7424 floats per ROI: largest stage tensor
    Y, S = 0, G * ymax
    B, T1, T2 = S, S + G * ymax, S + 2 * G * ymax
    t2_roi = 8 * 8 * widths[0]                                        # 3712
    IMG, C1 = S, S + (in_size * in_size * 3 + 3) // 4
    X0 = C1 + 32 * 32 * c1
    F_T2 = C1
    F_T3 = F_T2 + 16 * 16 * widths[0]
    F_T1 = F_T3 + 8 * 8 * widths[0]
    assert F_T1 + 8 * 8 * c1 <= X0 and X0 + 16 * 16 * c1 <= S + max(G * (2 * ymax + t2_roi), 0) + 10 ** 9
    scratch = max(X0 + 16 * 16 * c1 - S, G * (2 * ymax + t2_roi))
    total_floats = S + scratch
    # ---- front end (per ROI)
    w, b = _fold_bn(sd["conv1.0.weight"], sd, "conv1.1")             # [24,3,3,3]
    cp = (c1 + 3) // 4 * 4
    w1 = np.zeros((27, cp), np.float32); w1[:, :c1] = w.transpose(2, 3, 1, 0).reshape(27, c1)
    b1 = np.zeros(cp, np.float32); b1[:c1] = b
    step(FS_CONV1, IMG, C1, dst_C=c1, cout=c1, H=64, W=64, stride=2, relu=1, w_off=push(w1), b_off=push(b1))
    step(FS_MAXPOOL, C1, X0, src_C=c1, dst_C=c1, cout=c1, H=32, W=32, stride=2)
    bf = widths[0]
    wo, bo, c = dw_w("stage2.0.branch1.0", "stage2.0.branch1.1")
    step(FS_DW, X0, F_T1, src_C=c1, dst_C=c1, cout=c, H=16, W=16, stride=2, w_off=wo, b_off=bo)
    wo, bo, ci, co = pw_w("stage2.0.branch1.2", "stage2.0.branch1.3")
    step(FS_PW, F_T1, Y, src_C=c1, dst_C=2 * bf, dst_off=0, dst_cs=2, cin=ci, cout=co, H=8, W=8, relu=1, w_off=wo, b_off=bo,
         roi_stride=ymax)
    wo, bo, ci, co = pw_w("stage2.0.branch2.0", "stage2.0.branch2.1")
    step(FS_PW, X0, F_T2, src_C=c1, dst_C=bf, cin=ci, cout=co, H=16, W=16, relu=1, w_off=wo, b_off=bo)
    wo, bo, c = dw_w("stage2.0.branch2.3", "stage2.0.branch2.4")
    step(FS_DW, F_T2, F_T3, src_C=bf, dst_C=bf, cout=c, H=16, W=16, stride=2, w_off=wo, b_off=bo)
    wo, bo, ci, co = pw_w("stage2.0.branch2.5", "stage2.0.branch2.6")
    step(FS_PW, F_T3, Y, src_C=bf, dst_C=2 * bf, dst_off=1, dst_cs=2, cin=ci, cout=co, H=8, W=8, relu=1, w_off=wo, b_off=bo,
         roi_stride=ymax)
    n_front = len(steps)

    def emit_units(stage, bf, X, OUT, T1, T2, hw, skip_first):
        """units of one stage on ping-pong buffers X/OUT with temporaries T1/T2; returns (X, OUT, hw)"""
        u = 0
        while f"stage{stage}.{u}.branch2.0.weight" in sd:
            pre = f"stage{stage}.{u}"
            if u == 0 and skip_first:
                u += 1
                continue                                              # done in the front end
            if u == 0:                                                # down-sampling unit
                cin_all = bf                                          # previous stage has 2*(bf/2) = bf channels
                wo, bo, c = dw_w(pre + ".branch1.0", pre + ".branch1.1")
                step(FS_DW, X, T2, src_C=cin_all, dst_C=cin_all, cout=c, H=hw, W=hw, stride=2, w_off=wo, b_off=bo)
                wo, bo, ci, co = pw_w(pre + ".branch1.2", pre + ".branch1.3")
                step(FS_PW, T2, OUT, src_C=cin_all, dst_C=2 * bf, dst_off=0, dst_cs=2, cin=ci, cout=co, H=hw // 2, W=hw // 2,
                     relu=1, w_off=wo, b_off=bo)
                wo, bo, ci, co = pw_w(pre + ".branch2.0", pre + ".branch2.1")
                step(FS_PW, X, T1, src_C=cin_all, dst_C=bf, cin=ci, cout=co, H=hw, W=hw, relu=1, w_off=wo, b_off=bo)
                wo, bo, c = dw_w(pre + ".branch2.3", pre + ".branch2.4")
                step(FS_DW, T1, T2, src_C=bf, dst_C=bf, cout=c, H=hw, W=hw, stride=2, w_off=wo, b_off=bo)
                hw //= 2
                wo, bo, ci, co = pw_w(pre + ".branch2.5", pre + ".branch2.6")
                step(FS_PW, T2, OUT, src_C=bf, dst_C=2 * bf, dst_off=1, dst_cs=2, cin=ci, cout=co, H=hw, W=hw, relu=1,
                     w_off=wo, b_off=bo)
            else:                                                     # basic unit
                step(FS_COPY, X, OUT, src_C=2 * bf, src_off=0, dst_C=2 * bf, dst_off=0, dst_cs=2, cout=bf, H=hw, W=hw)
                wo, bo, ci, co = pw_w(pre + ".branch2.0", pre + ".branch2.1")
                step(FS_PW, X, T1, src_C=2 * bf, src_off=bf, dst_C=bf, cin=ci, cout=co, H=hw, W=hw, relu=1, w_off=wo, b_off=bo)
                wo, bo, c = dw_w(pre + ".branch2.3", pre + ".branch2.4")
                step(FS_DW, T1, T2, src_C=bf, dst_C=bf, cout=c, H=hw, W=hw, stride=1, w_off=wo, b_off=bo)
                wo, bo, ci, co = pw_w(pre + ".branch2.5", pre + ".branch2.6")
                step(FS_PW, T2, OUT, src_C=bf, dst_C=2 * bf, dst_off=1, dst_cs=2, cin=ci, cout=co, H=hw, W=hw, relu=1,
                     w_off=wo, b_off=bo)
            X, OUT = OUT, X
            u += 1
        return X, OUT, hw

    # ---- middle (per ROI): stage2 units 1-3 and stage3 on the tensor cores, result parked in global memory
    X, OUT, hw = emit_units(2, widths[0], Y, B, T1, T2, 8, True)
    X, OUT, hw = emit_units(3, widths[1], X, OUT, T1, T2, hw, False)
    park = hw * hw * 2 * widths[1]                                    # 4 x 4 x 232 floats per ROI
    step(FS_STORE, X, 0, src_C=2 * widths[1], cout=2 * widths[1], H=hw, W=hw, roi_stride=park)
    n_mid = len(steps) - n_front
    # ---- tail (GT ROIs of this CTA stacked as rows): stage4, conv5, mean, fc.  Its weights are 73 % of the stream,
    # read once for the stacked ROIs.  Own map: IN | T1 | T2 | A | B (conv5's output overlays IN and the head of T1).
    GT = int(tail_group)
    bf4 = widths[2]
    t_in = 0
    t_t1 = t_in + GT * park
    t_t2 = t_t1 + GT * hw * hw * bf4                                   # T1 at stage4.0: 4 x 4 x 232 per ROI
    t_a = t_t2 + GT * (hw // 2) ** 2 * bf4
    tail_floats = t_a + GT * (hw // 2) ** 2 * 2 * bf4
    step(FS_LOAD, 0, t_in, dst_C=2 * widths[1], cout=2 * widths[1], H=hw, W=hw, roi_stride=park)
    frag_mode[0] = bool(tail_mma)
    # unit 0 reads IN and writes A; the basic units ping-pong between A and the (now dead) IN region
    X, OUT, hw = emit_units(4, bf4, t_in, t_a, t_t1, t_t2, hw, False)
    wo, bo, ci, co = pw_w("conv5.0", "conv5.1")
    c5 = co
    assert GT * hw * hw * c5 <= t_a - t_t1                            # conv5's output spans T1 and T2 (both dead)
    step(FS_PW, X, t_t1, src_C=2 * bf4, dst_C=c5, cin=ci, cout=co, H=hw, W=hw, relu=1, w_off=wo, b_off=bo)
    mscr = t_a if X == t_in else t_in                                  # means + K-split partial sums: the buffer conv5 did not read
    assert GT * c5 + 8 * GT * 64 <= GT * (hw * hw * 2 * bf4)
    step(FS_MEANFC, t_t1, mscr, src_C=c5, cin=c5, cout=ncls, H=hw, W=hw, w_off=push(sd["fc.weight"].T.copy()), b_off=push(sd["fc.bias"]))
    n_tail = len(steps) - n_front - n_mid
    arr = np.asarray(steps, dtype=np.int32)
    assert arr.shape[1] == FSTEP_WORDS
    back_floats = S + G * (2 * ymax + t2_roi)                        # Y | B | T1 | T2: all the middle touches
    # fp16 activation staging of the tensor-core pointwise layers (middle only): two planes [rows_p16][cin_p16 + 8]
    astage = 0
    for st in steps[n_front:n_front + n_mid]:
        if st[0] == FS_PW and st[17] > 0:
            rows_p = (G * st[10] * st[11] + 15) // 16 * 16
            astage = max(astage, 2 * rows_p * ((st[8] + 15) // 16 * 16 + 8) * 2)
    # every fp32 pointwise layer must fit its (row group x 4-channel) tiles into the kernel's 512 compute threads
    # (mirror of the tiling rule in csrc/shufflenet_fused.cu, case FS_PW)
    for k, st in enumerate(steps):
        if st[0] != FS_PW:
            continue
        cap = (GT if k >= n_front + n_mid else 1) * st[10] * st[11]
        ncg = (st[9] + 3) // 4
        rt = 4 if (cap == 12 and 3 * ncg <= 512) else (8 if cap >= 8 else 4 if cap >= 4 else 2 if cap >= 2 else 1)
        assert -(-cap // rt) * ncg <= 512, f"fused classifier: step {k} needs {-(-cap // rt) * ncg} tiles"
    # tail: rows = the stacked ROIs.  tail_mma: its pointwise layers on the tensor cores (csrc pw_layer_mma_direct): every warp owns
    # whole output tiles and streams their fragment-ordered weights straight from L2 into registers -- no shared-memory weight
    # stages, no K split, no partial sums (the first tensor-core tail chunked the weights along N through the 2 x 44 KB ring:
    # 43 chunks of 3 output tiles for conv5, each with two barriers and a reduction, and measured slower than the fp32 path).
    tail_astage = 0
    for st in steps[n_front + n_mid:]:
        if st[0] == FS_PW and st[17] > 0:
            if not tail_mma:
                st[17] = 0
                continue
            rows_p = (GT * st[10] * st[11] + 15) // 16 * 16
            tail_astage = max(tail_astage, 2 * rows_p * ((st[8] + 15) // 16 * 16 + 8) * 2)
    arr = np.asarray(steps, dtype=np.int32)
    w16 = np.concatenate(h16) if h16 else np.zeros(8, np.float16)
    return FusedProgram(arr, np.concatenate(blobs), w16, n_front, n_mid, n_tail, GT, total_floats * 4, back_floats * 4,
                        astage, tail_floats * 4, park, tail_astage)
