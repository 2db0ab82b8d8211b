"""Pins the JPEG-decode oracle (oracle/jpeg_ref.py) to the library the reference itself calls: cv2.imread / cv2.imdecode
(src/vntsr/pipeline/e2e.py:962).  Bit-exact on every case; also the host-side header parser / table packer of the product."""
import cv2
import numpy as np
import pytest

from oracle import jpeg_ref as J


def _img(h, w, seed):
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8)
    a = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
    cv2.circle(a, (w // 2, h // 2), max(2, min(h, w) // 3), (20, 20, 220), -1)
    return np.clip(a.astype(np.int16) + rng.integers(-8, 9, a.shape), 0, 255).astype(np.uint8)


CASES = [(64, 48, 90, None, 0), (37, 53, 75, None, 0), (120, 160, 95, None, 4), (16, 16, 50, None, 0), (8, 8, 100, None, 0),
         (33, 17, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 3), (40, 72, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 2),
         (9, 200, 90, None, 1), (1, 1, 90, None, 0), (17, 31, 30, None, 5), (100, 100, 10, None, 0)]


def encode(a, q, sf, rst):
    par = [cv2.IMWRITE_JPEG_QUALITY, q]
    if sf is not None:
        par += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf]
    if rst:
        par += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
    ok, b = cv2.imencode(".jpg", a, par)
    assert ok
    return bytes(b)


@pytest.mark.parametrize("h,w,q,sf,rst", CASES)
def test_oracle_equals_cv2_imdecode(h, w, q, sf, rst):
    b = encode(_img(h, w, h * 1000 + w), q, sf, rst)
    want = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(J.decode(b), want)


def test_oracle_grey_and_full_frame():
    from litepi_b200 import synth
    f = synth.vn_frame(3)
    g = cv2.cvtColor(f[:100, :130], cv2.COLOR_BGR2GRAY)
    b = encode(g, 80, None, 0)
    assert np.array_equal(J.decode(b), cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR))
    b = encode(f[:240, :400], 90, None, 4)                    # a crop of a VN-shape frame keeps the CPU suite fast
    assert np.array_equal(J.decode(b), cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR))


def test_product_header_parser_and_table_packer():
    from litepi_b200 import jpeg
    b = encode(_img(120, 160, 1), 90, None, 4)
    hd, ref = jpeg.parse_header(b), J.parse_header(b)
    assert (hd.width, hd.height, hd.restart_interval, hd.data_offset) == (ref.width, ref.height, ref.restart_interval, ref.data_offset)
    assert hd.comps == ref.comps and hd.scan == ref.scan and jpeg.is_jpeg(b) and not jpeg.is_jpeg(b"\x00\x01\x02\x03")
    d = jpeg.make_desc(hd)
    assert (d.width, d.height, d.ncomp, list(d.h), list(d.v), d.restart_interval) == (160, 120, 3, [2, 1, 1], [2, 1, 1], 4)
    blob = jpeg.pack_tables(hd)
    qt = blob[:1024].view(np.int32).reshape(4, 64)
    for k, z in ref.qt.items():
        nat = np.empty(64, np.int32); nat[J.ZIGZAG] = z
        assert np.array_equal(qt[k], nat)
    lut = blob[1024:1024 + 4096].view(np.uint16).reshape(4, 512)
    # every code of length <= 9 decodes through the look-up table exactly as the canonical tables say
    for (tc, th), t in list({(0, k): v for k, v in ref.dc.items()}.items()) + list({(1, k): v for k, v in ref.ac.items()}.items()):
        for ln in range(1, 10):
            for j in range(t.bits[ln - 1]):
                code = t.mincode[ln] + j
                e = int(lut[2 * tc + th][code << (9 - ln)])
                assert (e >> 8, e & 255) == (ln, t.vals[t.valptr[ln] + j])
    ok, prog = cv2.imencode(".jpg", _img(32, 32, 2), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError, match="baseline"):
        jpeg.parse_header(bytes(prog))
    with pytest.raises(ValueError, match="SOI"):
        jpeg.parse_header(b"not a jpeg at all")
