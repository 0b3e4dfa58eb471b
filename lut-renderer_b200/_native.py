"""ctypes binding of the C-ABI (include/lutr.h) -- the only way Python reaches the CUDA code.

There is no fallback: if ``liblutr.so`` is missing this module raises, and every compute call raises
``LutrError`` when the library reports a failure (e.g. no sm_100 device).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# LUTR_LIB: another build of the same library (A/B measurements of kernel variants in one GPU session)
LIB_PATH = os.environ.get("LUTR_LIB") or os.path.join(_HERE, "liblutr.so")

OK, ERR_INVALID_ARG, ERR_INVALID_DATA, ERR_LUT_SIZE, ERR_IO, ERR_NOMEM, ERR_CUDA, ERR_UNSUPPORTED, ERR_CANCELLED = 0, -1, -2, -3, -4, -5, -6, -7, -8

PIX_FMT = {"yuv420p": 0, "yuv422p10le": 1, "yuv420p10le": 2, "yuv422p": 3, "yuv444p": 4, "yuv444p10le": 5}
RANGE = {"tv": 0, "pc": 1}
MATRIX = {"bt709": 0, "bt601": 1, "smpte170m": 1, "bt470bg": 1, "bt2020": 2, "bt2020nc": 2, "bt2020c": 2, "fcc": 3,
          "smpte240m": 4}
INTERP = {"nearest": 0, "trilinear": 1, "tetrahedral": 2, "pyramid": 3, "prism": 4, "none": 5}
FLOAT_MODE = {"c": 0, "fma": 1}


class LutrError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"lutr error {status}: {message}")
        self.status = status
        self.message = message


class StageDesc(Structure):
    _fields_ = [("struct_size", c_uint32), ("width", c_int32), ("height", c_int32), ("in_fmt", c_int32), ("out_fmt", c_int32),
                ("in_range", c_int32), ("norm_range", c_int32), ("in_matrix", c_int32), ("out_matrix", c_int32),
                ("out_range", c_int32), ("interp", c_int32), ("float_mode", c_int32), ("via_8bit", c_int32), ("reserved", c_int32 * 3)]


class Planes(Structure):
    _fields_ = [("data", c_void_p * 3), ("stride", c_int32 * 3), ("frame_step", c_int64)]


class PumpStats(Structure):
    _fields_ = [("frames_in", c_int64), ("frames_out", c_int64), ("bytes_in", c_int64), ("bytes_out", c_int64), ("seconds", ctypes.c_double)]


PROGRESS_FN = ctypes.CFUNCTYPE(None, c_void_p, c_int64)


# every symbol include/lutr.h declares: (restype, argtypes)
SIGNATURES = {
    "lutr_version": (ctypes.c_int, []),
    "lutr_status_string": (c_char_p, [ctypes.c_int]),
    "lutr_last_error": (c_char_p, []),
    "lutr_cube_parse": (ctypes.c_int, [c_char_p, c_size_t, POINTER(c_int32), POINTER(c_float), POINTER(POINTER(c_float))]),
    "lutr_cube_parse_file": (ctypes.c_int, [c_char_p, POINTER(c_int32), POINTER(c_float), POINTER(POINTER(c_float))]),
    "lutr_free": (None, [c_void_p]),
    "lutr_device_count": (ctypes.c_int, [POINTER(c_int32)]),
    "lutr_ctx_create": (ctypes.c_int, [c_int32, POINTER(c_void_p)]),
    "lutr_ctx_destroy": (None, [c_void_p]),
    "lutr_ctx_info": (ctypes.c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "lutr_host_alloc": (ctypes.c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "lutr_host_free": (None, [c_void_p, c_void_p]),
    "lutr_lut_load_cube": (ctypes.c_int, [c_void_p, c_char_p, POINTER(c_void_p)]),
    "lutr_lut_load_file": (ctypes.c_int, [c_void_p, c_char_p, POINTER(c_void_p)]),
    "lutr_lut_parse_text": (ctypes.c_int, [c_char_p, c_size_t, c_int32, POINTER(c_int32), POINTER(c_float), POINTER(POINTER(c_float))]),
    "lutr_lut_parse_file": (ctypes.c_int, [c_char_p, POINTER(c_int32), POINTER(c_float), POINTER(POINTER(c_float))]),
    "lutr_lut_load_cube_memory": (ctypes.c_int, [c_void_p, c_char_p, c_size_t, POINTER(c_void_p)]),
    "lutr_lut_from_table": (ctypes.c_int, [c_void_p, c_int32, c_void_p, c_void_p, POINTER(c_void_p)]),
    "lutr_lut_size": (ctypes.c_int, [c_void_p]),
    "lutr_lut_destroy": (None, [c_void_p]),
    "lutr_stage_create": (ctypes.c_int, [c_void_p, POINTER(StageDesc), POINTER(c_void_p)]),
    "lutr_stage_destroy": (None, [c_void_p]),
    "lutr_stage_frame_bytes": (ctypes.c_int, [c_void_p, POINTER(c_size_t), POINTER(c_size_t)]),
    "lutr_stage_packed_planes": (ctypes.c_int, [c_void_p, c_int32, c_void_p, POINTER(Planes)]),
    "lutr_stage_run": (ctypes.c_int, [c_void_p, c_void_p, POINTER(Planes), POINTER(Planes), c_int32, c_void_p]),
    "lutr_stage_run_host": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32]),
    "lutr_stage_pump": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_int32, POINTER(c_int32), PROGRESS_FN, c_void_p,
                                       POINTER(PumpStats)]),
    "lutr_stage_sync": (ctypes.c_int, [c_void_p]),
    "lutr_stage_launch_count": (c_int64, [c_void_p]),
    "lutr_stage_kernel_name": (c_char_p, [c_void_p]),
    "lutr_stage_set_option": (ctypes.c_int, [c_void_p, c_char_p, c_int32]),
    "lutr_tables_yuv2rgb": (ctypes.c_int, [c_int32, c_int32, POINTER(c_int32)]),
    "lutr_tables_rgb2yuv": (ctypes.c_int, [c_int32, POINTER(c_int32)]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises (never substitutes anything) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python lut-renderer_b200/build.py` "
                              "(there is no CPU fallback for the colour stage)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int) -> int:
    if status < 0:
        raise LutrError(status, (lib().lutr_last_error() or b"").decode("utf-8", "replace")
                        or lib().lutr_status_string(status).decode())
    return status


LUT_DIALECT = {"cube": 0, "dat": 1, "3dl": 2, "m3d": 3, "csp": 4}   # lutr_lut_dialect


def parse_lut_file(path) -> tuple:
    """Host-only parse of any LUT file lut3d reads, the dialect chosen by the extension -> (size, scale[3], table [r][g][b][3])."""
    import os
    size = c_int32()
    scale = (c_float * 3)()
    table = POINTER(c_float)()
    check(lib().lutr_lut_parse_file(os.fsencode(str(path)), byref(size), scale, byref(table)))
    return _take_table(size, scale, table)


def parse_cube_text(text, dialect: str = "cube") -> tuple:
    """Host-only .cube (or other dialect) parse -> (size, scale[3], float32 table [r][g][b][3])."""
    data = text if isinstance(text, (bytes, bytearray)) else str(text).encode()
    size = c_int32()
    scale = (c_float * 3)()
    table = POINTER(c_float)()
    if dialect == "cube":
        check(lib().lutr_cube_parse(bytes(data), len(data), byref(size), scale, byref(table)))
    else:
        check(lib().lutr_lut_parse_text(bytes(data), len(data), LUT_DIALECT[dialect], byref(size), scale, byref(table)))
    return _take_table(size, scale, table)


def _take_table(size, scale, table) -> tuple:
    try:
        n = size.value ** 3 * 3
        values = (c_float * n).from_address(ctypes.addressof(table.contents))
        import numpy as np
        arr = np.frombuffer(values, dtype=np.float32).reshape(size.value, size.value, size.value, 3).copy()
    finally:
        lib().lutr_free(table)
    return size.value, [scale[0], scale[1], scale[2]], arr
