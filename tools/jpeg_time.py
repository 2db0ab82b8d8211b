"""Time of the device JPEG decode of one batch (development tool).  LP_JPEG_STOP=1|2 ends after the marker scan / the
Huffman+IDCT kernel, so three runs give the per-kernel split.  usage: jpeg_time.py [batch] [restart_interval] [quality]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, cv2, ctypes as C
import litepi_b200
from litepi_b200 import synth, _lib as L
from litepi_b200.jpeg import JpegBatchDecoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ri = int(sys.argv[2]) if len(sys.argv) > 2 else 4
q = int(sys.argv[3]) if len(sys.argv) > 3 else 90
frames = [synth.vn_frame(i) for i in range(B)]
js = [bytes(cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, ri])[1]) for f in frames]
dec = JpegBatchDecoder(L.context(0), torch.device("cuda", 0), B)
total = sum(len(j) for j in js)
hb = np.empty(total, np.uint8); ho = np.zeros(B + 1, np.int64)
hit, used = dec.stage(js, hb, ho)
d = torch.from_numpy(hb).cuda(); o = torch.from_numpy(ho).cuda()
out = torch.empty((B, hit[0].height, hit[0].width, 3), dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): dec.decode_device(hit, d, o, B, out, st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): dec.decode_device(hit, d, o, B, out, st)
b.record(); b.synchronize()
print(f"batch {B} ri {ri} q {q} bytes/frame {total / B:.0f} stop={os.environ.get('LP_JPEG_STOP', '0')}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per batch")
