"""ctypes binding of ``liblitepi_b200.so`` (C-ABI in ``include/litepi_b200.h``).

There is no CPU fallback: if the shared library is missing or no B200 is present,
every entry point raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblitepi_b200.so")

# enums (include/litepi_b200.h)
OP_STEM_U8, OP_CONV, OP_DWCONV3, OP_MAXPOOL, OP_UPSAMPLE2, OP_COPY, OP_MEAN_FC, OP_GLOBAL_MEAN, OP_SCALE = range(9)
ACT_NONE, ACT_SILU, ACT_RELU, ACT_RELU6, ACT_SIGMOID = range(5)
OPF_RES_BEFORE_ACT = 1
FMT_SPLIT16, FMT_F32, FMT_U8 = range(3)
NET_DETECTOR, NET_CLASSIFIER = 0, 1
ABI_VERSION = 5


class BufDesc(C.Structure):
    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("fmt", C.c_int32),
                ("offset", C.c_int64), ("image_bytes", C.c_int64)]


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_buf", C.c_int32), ("in_coff", C.c_int32), ("cin", C.c_int32),
                ("out_buf", C.c_int32), ("out_coff", C.c_int32), ("cout", C.c_int32),
                ("out_cstride", C.c_int32), ("cout_real", C.c_int32), ("out_seg_len", C.c_int32),
                ("out_seg_pad", C.c_int32), ("res_buf", C.c_int32), ("res_coff", C.c_int32),
                ("ksize", C.c_int32), ("stride", C.c_int32), ("act", C.c_int32),
                ("row_off", C.c_int32), ("flags", C.c_int32), ("in_mean", C.c_float), ("in_std", C.c_float),
                ("w_off", C.c_int64), ("b_off", C.c_int64), ("wtc_off", C.c_int64)]


class JpegDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("h", C.c_int32 * 3), ("v", C.c_int32 * 3),
                ("tq", C.c_int32 * 3), ("td", C.c_int32 * 3), ("ta", C.c_int32 * 3), ("restart_interval", C.c_int32)]


_lib = None

_PROTOS = {
    "lp_abi_version": (C.c_int, []),
    "lp_last_error": (C.c_char_p, []),
    "lp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "lp_destroy": (C.c_int, [C.c_void_p]),
    "lp_net_load": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(BufDesc), C.c_int, C.POINTER(OpDesc), C.c_int,
                              C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "lp_set_tensor_core": (C.c_int, [C.c_void_p, C.c_int]),
    "lp_set_fused_classifier": (C.c_int, [C.c_void_p, C.c_int]),
    "lp_set_pdl": (C.c_int, [C.c_void_p, C.c_int]),
    "lp_op_paths": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "lp_fused_classifier_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                           C.c_int, C.c_void_p, C.c_size_t, C.c_float, C.c_float]),
    "lp_sm_count": (C.c_int, [C.c_void_p]),
    "lp_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "lp_letterbox": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                               C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_void_p,
                               C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]),
    "lp_detect_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "lp_decode_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "lp_decode_nms_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "lp_roi_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lp_roi_resize": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lp_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "lp_pack_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lp_set_roi_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "lp_set_roi_count_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lp_eval_match": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "lp_jpeg_tables_bytes": (C.c_size_t, []),
    "lp_jpeg_scratch_bytes": (C.c_size_t, [C.POINTER(JpegDesc), C.c_int]),
    "lp_jpeg_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(JpegDesc), C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p, C.c_void_p]),
    "lp_launch_count": (C.c_int64, [C.c_void_p]),
    "lp_debug_tc_timing": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lp_probe_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "lp_probe_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
}

EXPORTS = tuple(_PROTOS)


def lib():
    """Load the shared library once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"litepi_b200: {LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C yolo-litepi_b200/csrc`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.lp_abi_version() != ABI_VERSION:
            raise RuntimeError("litepi_b200: shared library ABI version mismatch; rebuild it")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().lp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"litepi_b200 {what} failed ({rc}): {msg}")


class Context:
    """One per process / GPU (``lp_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().lp_create(C.byref(self._h), int(device)), "lp_create")
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(lib().lp_launch_count(self._h))

    def probe_set(self, net: int, op_index: int):
        check(lib().lp_probe_set(self._h, net, op_index), "lp_probe_set")

    def probe_read(self):
        buf = (C.c_float * 512)()
        n = lib().lp_probe_read(self._h, buf, 512)
        if n < 0:
            check(n, "lp_probe_read")
        return [float(buf[i]) for i in range(n)]

    def set_tensor_core(self, enable: bool):
        check(lib().lp_set_tensor_core(self._h, 1 if enable else 0))

    def op_paths(self, net: int):
        buf = (C.c_int8 * 512)()
        n = lib().lp_op_paths(self._h, net, buf, 512)
        if n < 0:
            check(n, "lp_op_paths")
        return [int(buf[i]) for i in range(n)]

    def set_pdl(self, enable: bool):
        check(lib().lp_set_pdl(self._h, 1 if enable else 0))

    def close(self):
        if self._h:
            lib().lp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def context(device: int = 0) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
