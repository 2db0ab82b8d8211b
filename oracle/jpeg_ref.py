"""Oracle: baseline JPEG decoder, bit-exact with what the reference's ``cv2.imread`` produces
(``src/vntsr/pipeline/e2e.py:962``) -- TEST INFRASTRUCTURE (see oracle/__init__.py).

``cv2.imread`` / ``cv2.imdecode`` hand the file to the JPEG library bundled with OpenCV (libjpeg-turbo; pinned OpenCV
4.9.0.80 in requirements.txt:23, 4.13.0 / libjpeg-turbo 3.1.2 in this image) with its defaults: ``JDCT_ISLOW``,
``do_fancy_upsampling = TRUE``, output colour space BGR.  The algorithm restated here is libjpeg's published one
(ITU-T T.81 entropy coding; jidctint.c ``jpeg_idct_islow``; jdsample.c ``h2v1/h2v2_fancy_upsample``; jdcolor.c
``ycc_rgb_convert`` fixed-point tables); the SIMD paths of libjpeg-turbo are bit-identical to those C routines by
design.  ``tests/test_oracle_jpeg.py`` pins this restatement to ``cv2.imdecode`` itself on many images.

Scope (what the GPU ingest path supports): baseline sequential DCT (SOF0), 8-bit, one scan,
3 components YCbCr with luma sampling 1x1 / 2x1 / 2x2 (4:4:4, 4:2:2, 4:2:0) or 1 component (grey), optional
restart intervals.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                   21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60,
                   61, 54, 47, 55, 62, 63], dtype=np.int64)       # zigzag index -> natural (row-major) index


@dataclass
class HuffTable:
    bits: List[int]                  # number of codes of length 1..16
    vals: List[int]
    # derived (T.81 Annex C / F.2.2.3)
    mincode: List[int] = field(default_factory=list)
    maxcode: List[int] = field(default_factory=list)
    valptr: List[int] = field(default_factory=list)

    def build(self):
        code, k = 0, 0
        self.mincode, self.maxcode, self.valptr = [0] * 17, [-1] * 17, [0] * 17
        for ln in range(1, 17):
            n = self.bits[ln - 1]
            self.valptr[ln] = k
            self.mincode[ln] = code
            if n:
                code += n
                self.maxcode[ln] = code - 1
                k += n
            code <<= 1
        return self


@dataclass
class JpegHeader:
    width: int
    height: int
    comps: List[Tuple[int, int, int, int]]          # (id, h, v, quant table id)
    qt: Dict[int, np.ndarray]                        # table id -> 64 values in ZIGZAG order
    dc: Dict[int, HuffTable]
    ac: Dict[int, HuffTable]
    scan: List[Tuple[int, int, int]]                # per scan component: (component index, dc table, ac table)
    restart_interval: int
    data_offset: int                                 # first entropy-coded byte
    header_bytes: bytes = b""


def parse_header(buf: bytes) -> JpegHeader:
    if buf[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG (no SOI)")
    i, qt, dc, ac, dri = 2, {}, {}, {}, 0
    sof = None
    while True:
        if buf[i] != 0xFF:
            raise ValueError("marker expected")
        m = buf[i + 1]
        if m == 0xFF:
            i += 1
            continue
        L = (buf[i + 2] << 8) | buf[i + 3]
        seg = buf[i + 4:i + 2 + L]
        if m == 0xDB:
            p = 0
            while p < len(seg):
                pq, tq = seg[p] >> 4, seg[p] & 15
                if pq:
                    raise ValueError("16-bit quantisation tables are not supported")
                qt[tq] = np.frombuffer(seg[p + 1:p + 65], np.uint8).astype(np.int32)
                p += 65
        elif m == 0xC4:
            p = 0
            while p < len(seg):
                tc, th = seg[p] >> 4, seg[p] & 15
                bits = list(seg[p + 1:p + 17])
                n = sum(bits)
                (ac if tc else dc)[th] = HuffTable(bits, list(seg[p + 17:p + 17 + n])).build()
                p += 17 + n
        elif m == 0xC0:
            if seg[0] != 8:
                raise ValueError("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            sof = (w, h, [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(nc)])
        elif m in (0xC1, 0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential JPEG (SOF0) is supported")
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if sof is None or ns != len(sof[2]):
                raise ValueError("one interleaved scan with all components is required")
            ids = [c[0] for c in sof[2]]
            scan = [(ids.index(seg[1 + 2 * k]), seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(ns)]
            off = i + 2 + L
            return JpegHeader(sof[0], sof[1], sof[2], qt, dc, ac, scan, dri, off, bytes(buf[:off]))
        i += 2 + L


class _Bits:
    """MSB-first bit reader over entropy-coded data with 0xFF00 unstuffing; stops at any other marker."""

    def __init__(self, buf: bytes, pos: int):
        self.buf, self.pos, self.acc, self.n = buf, pos, 0, 0

    def _fill(self):
        while self.n <= 24:
            b = self.buf[self.pos] if self.pos < len(self.buf) else 0
            if b == 0xFF:
                nxt = self.buf[self.pos + 1] if self.pos + 1 < len(self.buf) else 0xD9
                if nxt == 0:
                    self.pos += 2
                else:
                    b = 0                                    # marker: feed zeros, do not advance
                    self.acc = (self.acc << 8) | b
                    self.n += 8
                    continue
            else:
                self.pos += 1
            self.acc = (self.acc << 8) | b
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def decode(self, t: HuffTable) -> int:
        code = 0
        for ln in range(1, 17):
            code = (code << 1) | self.get(1)
            if t.maxcode[ln] >= 0 and code <= t.maxcode[ln]:
                return t.vals[t.valptr[ln] + code - t.mincode[ln]]
        raise ValueError("bad Huffman code")

    def align_after_restart(self):
        """discard the partial byte, skip the RSTn marker"""
        self.acc, self.n = 0, 0
        while not (self.buf[self.pos] == 0xFF and 0xD0 <= self.buf[self.pos + 1] <= 0xD7):
            self.pos += 1
        self.pos += 2


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def decode_coefficients(buf: bytes, hd: JpegHeader):
    """Entropy decoding (T.81 F.2): returns per component an int16 array [blocks_y, blocks_x, 64] of QUANTISED
    coefficients in natural order (padded to whole MCUs) and the MCU geometry."""
    hmax = max(c[1] for c in hd.comps)
    vmax = max(c[2] for c in hd.comps)
    mcux, mcuy = (hd.width + 8 * hmax - 1) // (8 * hmax), (hd.height + 8 * vmax - 1) // (8 * vmax)
    coefs = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in hd.comps]
    br = _Bits(buf, hd.data_offset)
    pred = [0] * len(hd.comps)
    n_mcu, ri = mcux * mcuy, hd.restart_interval
    for m in range(n_mcu):
        if ri and m and m % ri == 0:
            br.align_after_restart()
            pred = [0] * len(hd.comps)
        my, mx = divmod(m, mcux)
        for ci, tdc, tac in hd.scan:
            _, h, v, _ = hd.comps[ci]
            for by in range(v):
                for bx in range(h):
                    blk = coefs[ci][my * v + by, mx * h + bx]
                    s = br.decode(hd.dc[tdc])
                    pred[ci] += _extend(br.get(s), s) if s else 0
                    blk[0] = pred[ci]
                    k = 1
                    while k < 64:
                        rs = br.decode(hd.ac[tac])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k]] = _extend(br.get(s), s)
                        k += 1
    return coefs, (mcux, mcuy, hmax, vmax)


# jidctint.c constants (CONST_BITS = 13, PASS1_BITS = 2)
_C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(x, shift):
    """one pass of jpeg_idct_islow over the LAST axis of x (int64 [..., 8]); DESCALE by `shift`"""
    c = _C
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * c["f0_541"]
    tmp2 = z1 - z3 * c["f1_847"]
    tmp3 = z1 + z2 * c["f0_765"]
    tmp0 = (x[..., 0] + x[..., 4]) << 13
    tmp1 = (x[..., 0] - x[..., 4]) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * c["f1_175"]
    t0, t1, t2, t3 = t0 * c["f0_298"], t1 * c["f2_053"], t2 * c["f3_072"], t3 * c["f1_501"]
    z1, z2, z3, z4 = -z1 * c["f0_899"], -z2 * c["f2_562"], -z3 * c["f1_961"] + z5, -z4 * c["f0_390"] + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    r = (1 << (shift - 1))
    out = np.stack([tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3], -1)
    return (out + r) >> shift


def idct_blocks(coef: np.ndarray, qt_zigzag: np.ndarray) -> np.ndarray:
    """[by, bx, 64] quantised coefficients (natural order) -> u8 samples [by*8, bx*8] (jidctint.c jpeg_idct_islow)."""
    q = np.empty(64, np.int64)
    q[ZIGZAG] = qt_zigzag
    x = coef.astype(np.int64) * q                                   # dequantise
    by, bx = x.shape[:2]
    x = x.reshape(by, bx, 8, 8)
    ws = _idct_1d(np.swapaxes(x, 2, 3), 13 - 2)                      # pass 1: columns (last axis = row index)
    ws = np.swapaxes(ws, 2, 3)
    px = _idct_1d(ws, 13 + 2 + 3)                                    # pass 2: rows
    px = np.clip(px + 128, 0, 255).astype(np.uint8)
    return px.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def upsample_h2v1_fancy(c: np.ndarray, out_w: int) -> np.ndarray:
    """jdsample.c h2v1_fancy_upsample on rows of c [h, w]"""
    c = c.astype(np.int32)
    w = c.shape[1]
    left = np.concatenate([c[:, :1], c[:, :-1]], 1)
    right = np.concatenate([c[:, 1:], c[:, -1:]], 1)
    even = (3 * c + left + 1) >> 2
    odd = (3 * c + right + 2) >> 2
    even[:, 0] = c[:, 0]
    odd[:, w - 1] = c[:, w - 1]
    out = np.empty((c.shape[0], 2 * w), np.int32)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out[:, :out_w].astype(np.uint8)


def upsample_h2v2_fancy(c: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample on c [h, w] (the REAL down-sampled size: edges replicate the last real row/column)"""
    c = c.astype(np.int32)
    h, w = c.shape
    up = np.concatenate([c[:1], c[:-1]], 0)
    dn = np.concatenate([c[1:], c[-1:]], 0)
    rows = np.empty((2 * h, w), np.int32)
    rows[0::2] = 3 * c + up                                          # colsum of output rows 2i (nearer the row above)
    rows[1::2] = 3 * c + dn
    left = np.concatenate([rows[:, :1], rows[:, :-1]], 1)
    right = np.concatenate([rows[:, 1:], rows[:, -1:]], 1)
    even = (3 * rows + left + 8) >> 4
    odd = (3 * rows + right + 7) >> 4
    even[:, 0] = (4 * rows[:, 0] + 8) >> 4
    odd[:, w - 1] = (4 * rows[:, w - 1] + 7) >> 4
    out = np.empty((2 * h, 2 * w), np.int32)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out[:out_h, :out_w].astype(np.uint8)


def ycc_to_bgr(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    """jdcolor.c build_ycc_rgb_table + ycc_rgb_convert (SCALEBITS 16), output B,G,R"""
    x = np.arange(256, dtype=np.int64) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    cr_r = (fix(1.40200) * x + 32768) >> 16
    cb_b = (fix(1.77200) * x + 32768) >> 16
    cr_g = -fix(0.71414) * x
    cb_g = -fix(0.34414) * x + 32768
    yy = y.astype(np.int64)
    r = yy + cr_r[cr]
    g = yy + ((cb_g[cb] + cr_g[cr]) >> 16)
    b = yy + cb_b[cb]
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def decode(buf: bytes) -> np.ndarray:
    """JPEG bytes -> HWC BGR uint8 (grey JPEGs are replicated to 3 channels like cv2.IMREAD_COLOR)."""
    hd = parse_header(buf)
    coefs, (mcux, mcuy, hmax, vmax) = decode_coefficients(buf, hd)
    planes = [idct_blocks(coefs[i], hd.qt[c[3]]) for i, c in enumerate(hd.comps)]
    H, W = hd.height, hd.width
    if len(planes) == 1:
        g = planes[0][:H, :W]
        return np.stack([g, g, g], -1)
    y = planes[0][:H, :W]
    ch = []
    for i in (1, 2):
        _, h, v, _ = hd.comps[i]
        cw, chh = -(-W * h // hmax), -(-H * v // vmax)             # real down-sampled size (ceil)
        p = planes[i][:chh, :cw]
        if hmax // h == 1 and vmax // v == 1:
            ch.append(p[:H, :W])
        elif hmax // h == 2 and vmax // v == 1:
            ch.append(upsample_h2v1_fancy(p, W)[:H])
        elif hmax // h == 2 and vmax // v == 2:
            ch.append(upsample_h2v2_fancy(p, H, W))
        else:
            raise ValueError("unsupported chroma sampling")
    return ycc_to_bgr(y, ch[0], ch[1])
