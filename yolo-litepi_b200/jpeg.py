"""Frame ingest from JPEG bytes: the B200 counterpart of ``cv2.imread`` in the reference's
``process_image`` (``src/vntsr/pipeline/e2e.py:962``).  The host only parses the few hundred header bytes (cached per
distinct header: a camera / encoder emits the same tables for every frame); Huffman decoding, IDCT, chroma
up-sampling and colour conversion run on the device (``csrc/jpeg.cu``) and are bit-exact with ``cv2.imdecode``.

Supported: baseline sequential JPEG (SOF0), 8 bit, one interleaved scan, YCbCr 4:4:4 / 4:2:2 / 4:2:0 or grey, optional
restart intervals (one device thread per restart interval: encode with a restart interval of a few MCUs, e.g.
``cv2.imencode('.jpg', img, [cv2.IMWRITE_JPEG_RST_INTERVAL, 4])``).  Anything else raises ``ValueError``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

_ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42,
           49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


@dataclass
class JpegHeader:
    width: int
    height: int
    comps: List[Tuple[int, int, int, int]]      # (id, h, v, tq)
    qt: Dict[int, np.ndarray]                    # zigzag order
    huff: Dict[Tuple[int, int], Tuple[List[int], List[int]]]     # (class, id) -> (bits[16], vals)
    scan: List[Tuple[int, int, int]]            # per component index: (comp index, td, ta)
    restart_interval: int
    data_offset: int
    header_bytes: bytes


def parse_header(buf) -> JpegHeader:
    """Marker segments up to and including SOS (ITU-T T.81 Annex B)."""
    b = bytes(buf[:4096]) if len(buf) > 4096 else bytes(buf)
    if b[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG (no SOI marker)")
    i, qt, huff, dri, sof = 2, {}, {}, 0, None
    while True:
        if i + 4 > len(b):
            b = bytes(buf)                                   # unusually long header (EXIF): read all of it
            if i + 4 > len(b):
                raise ValueError("truncated JPEG header")
        if b[i] != 0xFF:
            raise ValueError("corrupt JPEG header")
        m = b[i + 1]
        if m == 0xFF:
            i += 1
            continue
        ln = (b[i + 2] << 8) | b[i + 3]
        if i + 2 + ln > len(b):
            b = bytes(buf)
        seg = b[i + 4:i + 2 + ln]
        if m == 0xDB:
            p = 0
            while p < len(seg):
                if seg[p] >> 4:
                    raise ValueError("16-bit quantisation tables are not supported")
                qt[seg[p] & 15] = np.frombuffer(seg[p + 1:p + 65], np.uint8).astype(np.int32)
                p += 65
        elif m == 0xC4:
            p = 0
            while p < len(seg):
                bits = list(seg[p + 1:p + 17])
                n = sum(bits)
                huff[(seg[p] >> 4, seg[p] & 15)] = (bits, list(seg[p + 17:p + 17 + n]))
                p += 17 + n
        elif m == 0xC0:
            if seg[0] != 8:
                raise ValueError("only 8-bit JPEG samples are supported")
            sof = ((seg[3] << 8) | seg[4], (seg[1] << 8) | seg[2],
                   [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(seg[5])])
        elif 0xC1 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise ValueError("only baseline sequential JPEG (SOF0) is supported (this file is progressive / lossless / arithmetic)")
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            if sof is None or seg[0] != len(sof[2]):
                raise ValueError("JPEG must have one interleaved scan with all components")
            ids = [c[0] for c in sof[2]]
            scan = [(ids.index(seg[1 + 2 * k]), seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(seg[0])]
            off = i + 2 + ln
            return JpegHeader(sof[0], sof[1], sof[2], qt, huff, scan, dri, off, b[:off])
        i += 2 + ln


def pack_tables(hd: JpegHeader) -> np.ndarray:
    """struct JpegTables of csrc/jpeg.cu: qt[4][64] int32 natural order | lut[4][512] u16 | maxcode[4][18] | mincode[4][17] |
    valptr[4][17] int32 | vals[4][256] u8.  Table slots: 0/1 = DC 0/1, 2/3 = AC 0/1."""
    qt = np.zeros((4, 64), np.int32)
    for k, z in hd.qt.items():
        if k < 4:
            qt[k][_ZIGZAG] = z
    lut = np.zeros((4, 512), np.uint16)
    maxcode = np.full((4, 18), -1, np.int32)
    mincode = np.zeros((4, 17), np.int32)
    valptr = np.zeros((4, 17), np.int32)
    vals = np.zeros((4, 256), np.uint8)
    for (tc, th), (bits, v) in hd.huff.items():
        if th > 1:
            raise ValueError("baseline JPEG uses Huffman table ids 0 and 1")
        t = 2 * tc + th
        vals[t, :len(v)] = v
        code, k = 0, 0
        for ln in range(1, 17):
            n = bits[ln - 1]
            valptr[t, ln], mincode[t, ln] = k, code
            for j in range(n):
                if ln <= 9:
                    lo = (code + j) << (9 - ln)
                    lut[t, lo:lo + (1 << (9 - ln))] = (ln << 8) | v[k + j]
            if n:
                maxcode[t, ln] = code + n - 1
                code += n
                k += n
            code <<= 1
        maxcode[t, 17] = 0x7FFFFFFF
    blob = b"".join(a.tobytes() for a in (qt, lut, maxcode, mincode, valptr, vals))
    need = L.lib().lp_jpeg_tables_bytes()
    if len(blob) != need:
        raise RuntimeError(f"JpegTables layout mismatch: packed {len(blob)} B, library expects {need} B")
    return np.frombuffer(blob, np.uint8).copy()


def make_desc(hd: JpegHeader) -> L.JpegDesc:
    d = L.JpegDesc()
    d.width, d.height, d.ncomp, d.restart_interval = hd.width, hd.height, len(hd.comps), hd.restart_interval
    if len(hd.comps) not in (1, 3):
        raise ValueError("JPEG must have 1 or 3 components")
    for k, (ci, td, ta) in enumerate(hd.scan):
        if ci != k:
            raise ValueError("JPEG scan components must be in frame order")
        _, h, v, tq = hd.comps[ci]
        d.h[k], d.v[k], d.tq[k], d.td[k], d.ta[k] = h, v, tq, td, ta
    if len(hd.comps) == 3:
        hv = (hd.comps[0][1], hd.comps[0][2])
        if hv not in ((1, 1), (2, 1), (2, 2)) or any(c[1] != 1 or c[2] != 1 for c in hd.comps[1:]):
            raise ValueError("JPEG chroma sampling must be 4:4:4, 4:2:2 or 4:2:0")
    return d


def is_jpeg(x) -> bool:
    if isinstance(x, (bytes, bytearray, memoryview)):
        return len(x) > 3 and x[0] == 0xFF and x[1] == 0xD8
    return isinstance(x, np.ndarray) and x.ndim == 1 and x.dtype == np.uint8 and x.size > 3 and x[0] == 0xFF and x[1] == 0xD8


class JpegBatchDecoder:
    """Decodes batches of same-header JPEGs into a device frame tensor.  Owns the per-header device tables and scratch."""

    def __init__(self, ctx: L.Context, device: torch.device, max_batch: int):
        self.ctx, self.device, self.max_batch = ctx, device, int(max_batch)
        self._cache: Dict[bytes, Tuple[JpegHeader, L.JpegDesc, torch.Tensor, torch.Tensor]] = {}

    def header(self, buf) -> Tuple[JpegHeader, L.JpegDesc, torch.Tensor, torch.Tensor]:
        last = getattr(self, "_last", None)                 # a stream repeats one header: compare bytes before parsing
        if last is not None:
            hb = last[0].header_bytes
            if len(buf) >= len(hb) and bytes(buf[:len(hb)]) == hb:
                return last
        hd = parse_header(buf)
        hit = self._cache.get(hd.header_bytes)
        if hit is None:
            desc = make_desc(hd)
            with torch.cuda.device(self.device):
                tables = torch.from_numpy(pack_tables(hd)).to(self.device)
                scratch = torch.zeros(L.lib().lp_jpeg_scratch_bytes(C.byref(desc), self.max_batch), dtype=torch.uint8, device=self.device)
            hit = (hd, desc, tables, scratch)
            self._cache[hd.header_bytes] = hit
        self._last = hit
        return hit

    def for_header(self, hd: JpegHeader):
        """tables + scratch of THIS decoder for an already parsed header (a lane picks up a batch packed by another)"""
        hit = self._cache.get(hd.header_bytes)
        return hit if hit is not None else self.header(hd.header_bytes + b"\xff\xd9")

    def stage(self, jpegs: Sequence, host_bytes: np.ndarray, host_off: np.ndarray):
        """Concatenate the entropy-coded scans of `jpegs` into `host_bytes` (pinned u8), offsets into `host_off` (int64 [B+1]).
        Returns (header tuple, total bytes).  All files must share one header."""
        hit = self.header(jpegs[0])
        hd = hit[0]
        hb, off = hd.header_bytes, hd.data_offset
        pos = 0
        host_off[0] = 0
        for k, j in enumerate(jpegs):
            a = np.frombuffer(j, np.uint8) if not isinstance(j, np.ndarray) else j
            if a.size < off or bytes(a[:off]) != hb:
                raise ValueError("all JPEGs of a batch must share one header (size, sampling, tables); frame %d differs" % k)
            n = a.size - off
            if pos + n > host_bytes.size:
                raise ValueError("JPEG batch exceeds the staging buffer")
            host_bytes[pos:pos + n] = a[off:]
            pos += n
            host_off[k + 1] = pos
        return hit, pos

    def decode_device(self, hit, data_dev: torch.Tensor, off_dev: torch.Tensor, n: int, out_frames: torch.Tensor, stream) -> None:
        hd, desc, tables, scratch = hit
        if out_frames.shape[1] != hd.height or out_frames.shape[2] != hd.width:
            raise ValueError("output frame buffer does not match the JPEG size")
        L.check(L.lib().lp_jpeg_decode(self.ctx.handle, C.c_void_p(data_dev.data_ptr()), C.c_void_p(off_dev.data_ptr()), int(n),
                                       self.max_batch, C.byref(desc), C.c_void_p(tables.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                       scratch.numel(), C.c_void_p(out_frames.data_ptr()), stream), "lp_jpeg_decode")

    def decode(self, jpegs: Sequence) -> torch.Tensor:
        """Convenience: list of JPEG byte strings -> [n, H, W, 3] BGR uint8 CUDA tensor (synchronous staging)."""
        n = len(jpegs)
        if n > self.max_batch:
            raise ValueError(f"batch {n} > max_batch {self.max_batch}")
        total = sum(len(j) for j in jpegs)
        hb = np.empty(total, np.uint8)
        ho = np.zeros(n + 1, np.int64)
        hit, used = self.stage(jpegs, hb, ho)
        with torch.cuda.device(self.device):
            d = torch.from_numpy(hb[:max(used, 1)]).to(self.device)
            o = torch.from_numpy(ho).to(self.device)
            out = torch.empty((n, hit[0].height, hit[0].width, 3), dtype=torch.uint8, device=self.device)
            self.decode_device(hit, d, o, n, out, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out
