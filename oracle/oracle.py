"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/liblutr_oracle.so (lutr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblutr_oracle.so")

MATRIX_ID = {"bt709": 0, "bt601": 1, "smpte170m": 1, "bt470bg": 1, "default": 1, "bt2020nc": 2, "bt2020c": 2,
             "bt2020": 2, "fcc": 3, "smpte240m": 4}
INTERP_ID = {"nearest": 0, "trilinear": 1, "tetrahedral": 2, "pyramid": 3, "prism": 4}
PIX_ID = {"yuv420p": 0, "yuv422p10le": 1, "yuv420p10le": 2, "yuv422p": 3, "yuv444p": 4, "yuv444p10le": 5}

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lutr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_cube_parse_path.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.orc_cube_parse_memory.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        L.orc_lut_free.argtypes = [ctypes.c_void_p]
        L.orc_lut_free.restype = None
        L.orc_lut_size.argtypes = [ctypes.c_void_p]
        L.orc_lut_data.argtypes = [ctypes.c_void_p]
        L.orc_lut_data.restype = ctypes.POINTER(ctypes.c_float)
        L.orc_lut_scale.argtypes = [ctypes.c_void_p]
        L.orc_lut_scale.restype = ctypes.POINTER(ctypes.c_float)
        L.orc_lut_from_table.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_lut_from_table.restype = ctypes.c_void_p
        L.orc_inv_table.restype = ctypes.POINTER(ctypes.c_int32)
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class OracleLut:
    """A parsed 3D LUT (in-memory order [r][g][b], float32 triplets) -- restates vf_lut3d's LUT3DContext."""

    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)
        L = lib()
        self.size = L.orc_lut_size(self._h)
        n = self.size ** 3 * 3
        self.table = np.ctypeslib.as_array(L.orc_lut_data(self._h), shape=(n,)).reshape(
            self.size, self.size, self.size, 3).copy()
        self.scale = np.ctypeslib.as_array(L.orc_lut_scale(self._h), shape=(3,)).copy()

    @classmethod
    def from_path(cls, path):
        h = ctypes.c_void_p()
        r = lib().orc_cube_parse_path(os.fsencode(str(path)), ctypes.byref(h))
        if r != 0:
            raise ValueError(f"cube parse failed: {r}")
        return cls(h.value)

    @classmethod
    def from_text(cls, text):
        data = text if isinstance(text, bytes) else text.encode()
        h = ctypes.c_void_p()
        r = lib().orc_cube_parse_memory(data, len(data), ctypes.byref(h))
        if r != 0:
            raise ValueError(f"cube parse failed: {r}")
        return cls(h.value)

    @classmethod
    def from_table(cls, table, scale=(1.0, 1.0, 1.0)):
        t = np.ascontiguousarray(table, dtype=np.float32)
        s = np.asarray(scale, dtype=np.float32)
        return cls(lib().orc_lut_from_table(t.shape[0], _p(t), _p(s)))

    def __del__(self):
        try:
            lib().orc_lut_free(self._h)
        except Exception:
            pass


def yuv2rgb_coeffs(matrix, full_range):
    out = (ctypes.c_int32 * 7)()
    lib().orc_yuv2rgb_setup(MATRIX_ID[matrix], int(full_range), out)
    return dict(zip(("yC", "yO8", "vr", "ub", "ug", "vg", "yOff10"), list(out)))


def rgb2yuv_coeffs(matrix):
    out = (ctypes.c_int32 * 9)()
    lib().orc_rgb2yuv_setup(MATRIX_ID[matrix], out)
    return list(out)


def range_tables():
    tl, tc = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    lib().orc_range_pc_to_tv_tables(_p(tl), _p(tc))
    return tl, tc


def yuv420p_to_rgb24(Y, U, V, matrix="bt709", full_range=False):
    h, w = Y.shape
    rgb = np.empty((h, 3 * w), np.uint8)
    lib().orc_yuv420p_to_rgb24(_p(Y), Y.strides[0], _p(U), _p(V), U.strides[0], _p(rgb), rgb.strides[0], w, h,
                               MATRIX_ID[matrix], int(full_range))
    return rgb


def rgb24_to_yuv420p(rgb, matrix="bt709"):
    h, w = rgb.shape[0], rgb.shape[1] // 3
    Y, U, V = np.empty((h, w), np.uint8), np.empty((h // 2, w // 2), np.uint8), np.empty((h // 2, w // 2), np.uint8)
    lib().orc_rgb24_to_yuv420p(_p(rgb), rgb.strides[0], _p(Y), w, _p(U), _p(V), w // 2, w, h, MATRIX_ID[matrix])
    return Y, U, V


def rgb24_to_yuv422p10(rgb, matrix="bt709"):
    h, w = rgb.shape[0], rgb.shape[1] // 3
    Y, U, V = np.empty((h, w), np.uint16), np.empty((h, w // 2), np.uint16), np.empty((h, w // 2), np.uint16)
    lib().orc_rgb24_to_yuv422p10(_p(rgb), rgb.strides[0], _p(Y), w, _p(U), _p(V), w // 2, w, h, MATRIX_ID[matrix])
    return Y, U, V


def yuv422p10_to_gbrp10(Y, U, V, matrix="bt709", full_range=False):
    h, w = Y.shape
    R, G, B = (np.empty((h, w), np.uint16) for _ in range(3))
    lib().orc_yuv422p10_to_gbrp10(_p(Y), w, _p(U), _p(V), w // 2, _p(R), _p(G), _p(B), w, w, h, MATRIX_ID[matrix],
                                  int(full_range))
    return R, G, B


def yuv420p10_to_gbrp10(Y, U, V, matrix="bt709", full_range=False):
    h, w = Y.shape
    R, G, B = (np.empty((h, w), np.uint16) for _ in range(3))
    lib().orc_yuv420p10_to_gbrp10(_p(Y), w, _p(U), _p(V), w // 2, _p(R), _p(G), _p(B), w, w, h, MATRIX_ID[matrix],
                                  int(full_range))
    return R, G, B


def gbrp10_to_yuv422p10(R, G, B, matrix="bt709"):
    h, w = R.shape
    Y, U, V = np.empty((h, w), np.uint16), np.empty((h, w // 2), np.uint16), np.empty((h, w // 2), np.uint16)
    lib().orc_gbrp10_to_yuv422p10(_p(R), _p(G), _p(B), w, _p(Y), w, _p(U), _p(V), w // 2, w, h, MATRIX_ID[matrix])
    return Y, U, V


def gbrp10_to_yuv420p10(R, G, B, matrix="bt709"):
    h, w = R.shape
    Y, U, V = np.empty((h, w), np.uint16), np.empty((h // 2, w // 2), np.uint16), np.empty((h // 2, w // 2), np.uint16)
    lib().orc_gbrp10_to_yuv420p10(_p(R), _p(G), _p(B), w, _p(Y), w, _p(U), _p(V), w // 2, w, h, MATRIX_ID[matrix])
    return Y, U, V


def lut3d_rgb24(lut: OracleLut, rgb, interp="tetrahedral"):
    h, w = rgb.shape[0], rgb.shape[1] // 3
    out = np.empty_like(rgb)
    lib().orc_lut3d_rgb24(lut._h, INTERP_ID[interp], _p(rgb), rgb.strides[0], _p(out), out.strides[0], w, h)
    return out


def lut3d_planar16(lut: OracleLut, R, G, B, depth=10, interp="tetrahedral"):
    h, w = R.shape
    Ro, Go, Bo = (np.empty((h, w), np.uint16) for _ in range(3))
    lib().orc_lut3d_planar16(lut._h, INTERP_ID[interp], depth, _p(R), _p(G), _p(B), w, _p(Ro), _p(Go), _p(Bo), w, w, h)
    return Ro, Go, Bo


def stage_yuv422p10(lut, Y, U, V, interp="tetrahedral", in_matrix="bt709", in_full_range=False, out_matrix="bt709"):
    """cfg 3/4 chain on one frame; planes are C-contiguous uint16 (h,w),(h,w/2),(h,w/2)."""
    h, w = Y.shape
    Yo, Uo, Vo = np.empty_like(Y), np.empty_like(U), np.empty_like(V)
    r = lib().orc_stage_yuv422p10(lut._h, INTERP_ID[interp], MATRIX_ID[in_matrix], int(in_full_range),
                                  MATRIX_ID[out_matrix], _p(Y), _p(U), _p(V), _p(Yo), _p(Uo), _p(Vo), w, h)
    if r:
        raise RuntimeError(f"orc_stage_yuv422p10: {r}")
    return Yo, Uo, Vo


def stage_yuv420p(lut, Y, U, V, interp="tetrahedral", range_pass=False, in_matrix="bt709", in_full_range=False,
                  out_matrix="bt709"):
    """cfg 1/2 chain on one frame; planes are C-contiguous uint8 (h,w),(h/2,w/2),(h/2,w/2)."""
    h, w = Y.shape
    Yo, Uo, Vo = np.empty_like(Y), np.empty_like(U), np.empty_like(V)
    r = lib().orc_stage_yuv420p(lut._h, INTERP_ID[interp], int(range_pass), MATRIX_ID[in_matrix], int(in_full_range),
                                MATRIX_ID[out_matrix], _p(Y), _p(U), _p(V), _p(Yo), _p(Uo), _p(Vo), w, h)
    if r:
        raise RuntimeError(f"orc_stage_yuv420p: {r}")
    return Yo, Uo, Vo


def stage_yuv420p_to_yuv422p10(lut, Y, U, V, interp="tetrahedral", range_pass=False, in_matrix="bt709", in_full_range=False,
                               out_matrix="bt709"):
    """Pro-mode stage 1 on an 8-bit 4:2:0 source: planes uint8 (h,w),(h/2,w/2),(h/2,w/2) -> uint16 (h,w),(h,w/2),(h,w/2)."""
    h, w = Y.shape
    Yo, Uo, Vo = np.empty((h, w), np.uint16), np.empty((h, w // 2), np.uint16), np.empty((h, w // 2), np.uint16)
    r = lib().orc_stage_yuv420p_to_yuv422p10(lut._h, INTERP_ID[interp], int(range_pass), MATRIX_ID[in_matrix], int(in_full_range),
                                             MATRIX_ID[out_matrix], _p(Y), _p(U), _p(V), _p(Yo), _p(Uo), _p(Vo), w, h)
    if r:
        raise RuntimeError(f"orc_stage_yuv420p_to_yuv422p10: {r}")
    return Yo, Uo, Vo


def stage_yuv420p10_to_yuv422p10(lut, Y, U, V, interp="tetrahedral", in_matrix="bt709", in_full_range=False, out_matrix="bt709"):
    """Pro-mode stage 1 on a 10-bit 4:2:0 source: planes uint16 (h,w),(h/2,w/2),(h/2,w/2) -> uint16 (h,w),(h,w/2),(h,w/2)."""
    h, w = Y.shape
    Yo, Uo, Vo = np.empty((h, w), np.uint16), np.empty((h, w // 2), np.uint16), np.empty((h, w // 2), np.uint16)
    r = lib().orc_stage_yuv420p10_to_yuv422p10(lut._h, INTERP_ID[interp], MATRIX_ID[in_matrix], int(in_full_range),
                                               MATRIX_ID[out_matrix], _p(Y), _p(U), _p(V), _p(Yo), _p(Uo), _p(Vo), w, h)
    if r:
        raise RuntimeError(f"orc_stage_yuv420p10_to_yuv422p10: {r}")
    return Yo, Uo, Vo


def stage_10bit(lut, Y, U, V, in_420=False, out_420=False, interp="tetrahedral", in_matrix="bt709", in_full_range=False,
                out_matrix="bt709"):
    """10-bit chain through gbrp10le; chroma planes uint16 (h/2 or h, w/2) on each side as in_420 / out_420 say."""
    h, w = Y.shape
    co = (h // 2 if out_420 else h, w // 2)
    Yo, Uo, Vo = np.empty((h, w), np.uint16), np.empty(co, np.uint16), np.empty(co, np.uint16)
    r = lib().orc_stage_10bit(lut._h, INTERP_ID[interp], int(in_420), int(out_420), MATRIX_ID[in_matrix], int(in_full_range),
                              MATRIX_ID[out_matrix], _p(Y), _p(U), _p(V), _p(Yo), _p(Uo), _p(Vo), w, h)
    if r:
        raise RuntimeError(f"orc_stage_10bit: {r}")
    return Yo, Uo, Vo


# ---- the whole format matrix (generic legs; lutr_oracle.c "orc_stage_any") -------------------------------------------------
def pix_geometry(pix_fmt, w, h):
    """[(rows, cols)] of the three planes and the sample dtype."""
    L = lib()
    sub, depth = L.orc_pix_sub(PIX_ID[pix_fmt]), L.orc_pix_depth(PIX_ID[pix_fmt])
    cw, ch = w >> (sub & 1), h >> (sub >> 1)
    return [(h, w), (ch, cw), (ch, cw)], (np.uint16 if depth > 8 else np.uint8)


def range_coeffs(depth, to_jpeg, chroma):
    c, o = ctypes.c_int64(), ctypes.c_int64()
    lib().orc_range_coeffs(depth, int(to_jpeg), int(chroma), ctypes.byref(c), ctypes.byref(o))
    return c.value, o.value


def yuv8_to_rgb24(pix_fmt, Y, U, V, matrix="bt709", full_range=False):
    h, w = Y.shape
    rgb = np.empty((h, 3 * w), np.uint8)
    L = lib()
    L.orc_yuv8_to_rgb24(L.orc_pix_sub(PIX_ID[pix_fmt]), _p(Y), w, _p(U), _p(V), U.shape[1], _p(rgb), 3 * w, w, h,
                        MATRIX_ID[matrix], int(full_range))
    return rgb


def yuv444p10_to_gbrp10(Y, U, V, matrix="bt709", full_range=False):
    h, w = Y.shape
    R, G, B = (np.empty((h, w), np.uint16) for _ in range(3))
    lib().orc_yuv444p10_to_gbrp10(_p(Y), _p(U), _p(V), w, _p(R), _p(G), _p(B), w, w, h, MATRIX_ID[matrix], int(full_range))
    return R, G, B


def rgb_to_yuv_any(src_depth, out_fmt, R, G, B, matrix="bt709", out_full_range=False):
    """R, G, B: uint16 planes holding rgb24 (src_depth 8) or gbrp10le (src_depth 10) values."""
    h, w = R.shape
    shapes, dt = pix_geometry(out_fmt, w, h)
    out = [np.empty(s, dt) for s in shapes]
    R, G, B = (np.ascontiguousarray(p, dtype=np.uint16) for p in (R, G, B))
    lib().orc_rgb_to_yuv_any(src_depth, PIX_ID[out_fmt], int(out_full_range), _p(R), _p(G), _p(B), w, _p(out[0]), _p(out[1]),
                             _p(out[2]), w, h, MATRIX_ID[matrix])
    return out


def prepass_10_to_8(mode, plane, src):
    """swscale's yuv4xxp10le(pc) -> yuv4xxp 8-bit re-quantisation of one plane (mode 1: with pc->tv, 2: range kept)."""
    h, w = src.shape
    dst = np.empty((h, w), np.uint8)
    src = np.ascontiguousarray(src, dtype=np.uint16)
    lib().orc_prepass_10_to_8(int(mode), int(plane), _p(src), w, _p(dst), w, w, h)
    return dst


def stage_any(lut, in_fmt, out_fmt, Y, U, V, interp="tetrahedral", range_pass=False, in_matrix="bt709", in_full_range=False,
              out_matrix="bt709", out_full_range=False):
    """Any supported source -> [range pass] -> RGB -> lut3d -> any supported output (one frame, contiguous planes).
    range_pass: 8-bit sources True = pc->tv pass; 10-bit sources 1 / 2 = the reference's 8-bit re-quantisation of a full-range
    source with / without the range conversion (the rest of the chain is then the 8-bit one)."""
    h, w = Y.shape
    shapes, dt = pix_geometry(out_fmt, w, h)
    out = [np.empty(s, dt) for s in shapes]
    r = lib().orc_stage_any(lut._h, INTERP_ID[interp], PIX_ID[in_fmt], PIX_ID[out_fmt], int(range_pass), MATRIX_ID[in_matrix],
                            int(in_full_range), MATRIX_ID[out_matrix], int(out_full_range), _p(Y), _p(U), _p(V), _p(out[0]),
                            _p(out[1]), _p(out[2]), w, h)
    if r:
        raise RuntimeError(f"orc_stage_any: {r}")
    return out


def convert_yuv422p10(out_fmt, Y, U, V):
    """Pro mode stage 2 without a LUT: yuv422p10le -> yuv420p / yuv420p10le / yuv422p as the encoding ffmpeg's auto-inserted scaler
    converts it (YUV -> YUV, same matrix and range; orc_convert_yuv422p10)."""
    h, w = Y.shape
    shapes, dt = pix_geometry(out_fmt, w, h)
    out = [np.empty(s, dt) for s in shapes]
    Y, U, V = (np.ascontiguousarray(p, dtype=np.uint16) for p in (Y, U, V))
    r = lib().orc_convert_yuv422p10(PIX_ID[out_fmt], _p(Y), _p(U), _p(V), _p(out[0]), _p(out[1]), _p(out[2]), w, h)
    if r:
        raise RuntimeError(f"orc_convert_yuv422p10: {out_fmt} is not modelled")
    return out


def set_float_mode(mode: str) -> None:
    """"c" (default) or "fma": evaluation order of the trilinear / tetrahedral blend (see orc_set_float_mode).  Process-wide."""
    lib().orc_set_float_mode(1 if mode == "fma" else 0)
