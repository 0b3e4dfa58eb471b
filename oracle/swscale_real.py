"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

ctypes driver for the *real* FFmpeg 8.0.1 libswscale 9.1.100 that ships inside this
image's ``opencv_python_headless.libs``.  libswscale is the actual
implementation the reference invokes (through the ``scale``/``format`` filters and the
auto-inserted scalers around ``lut3d``: /root/reference/src/lut_renderer/ffmpeg.py:225-235,
:246, :309-310) for every arithmetic leg of the colour stage except the LUT itself.

It is used to (1) pin ``oracle/swscale_model`` (the closed-form integer restatement the CUDA
kernels implement) and (2) as the conversion legs of the CPU baseline.

Context creation follows vf_scale (alloc -> av_opt_set_int -> sws_init_context ->
sws_setColorspaceDetails); ``sws_getContext`` ignores range on YUV->YUV (App. D pitfall 2).
"""
from __future__ import annotations

import ctypes
import glob
import os
import threading

import numpy as np

SWS_BICUBIC = 4
# SWS_CS_* ids accepted by sws_getCoefficients()
SWS_CS = {"bt709": 1, "fcc": 4, "bt470bg": 5, "smpte170m": 5, "bt601": 5, "default": 5,
          "smpte240m": 7, "bt2020nc": 9, "bt2020c": 9, "bt2020": 9}

_lock = threading.Lock()
_libs = None


def _find_libdir():
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    if spec is None or not spec.origin:
        return None
    site = os.path.dirname(os.path.dirname(spec.origin))
    d = os.path.join(site, "opencv_python_headless.libs")
    return d if os.path.isdir(d) else None


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def _load():
    global _libs
    with _lock:
        if _libs is not None:
            return _libs
        d = _find_libdir()
        if d is None:
            raise RuntimeError("bundled FFmpeg libs (opencv_python_headless.libs) not found")
        import cv2  # noqa: F401  (its RPATH pulls in libdrm/libcrypto deps; App. D pitfall 1)
        avu = ctypes.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0])
        sws = ctypes.CDLL(glob.glob(os.path.join(d, "libswscale-*.so*"))[0])
        avu.av_version_info.restype = ctypes.c_char_p
        avu.av_get_pix_fmt.argtypes = [ctypes.c_char_p]
        avu.av_get_pix_fmt.restype = ctypes.c_int
        avu.av_opt_set_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
        avu.av_opt_set_int.restype = ctypes.c_int
        sws.swscale_version.restype = ctypes.c_uint
        sws.sws_alloc_context.restype = ctypes.c_void_p
        sws.sws_init_context.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        sws.sws_init_context.restype = ctypes.c_int
        sws.sws_freeContext.argtypes = [ctypes.c_void_p]
        sws.sws_freeContext.restype = None
        sws.sws_getCoefficients.argtypes = [ctypes.c_int]
        sws.sws_getCoefficients.restype = ctypes.POINTER(ctypes.c_int)
        sws.sws_setColorspaceDetails.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int]
        sws.sws_setColorspaceDetails.restype = ctypes.c_int
        sws.sws_scale.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                  ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                  ctypes.POINTER(ctypes.c_int)]
        sws.sws_scale.restype = ctypes.c_int
        _libs = (avu, sws)
        return _libs


def versions():
    avu, sws = _load()
    v = sws.swscale_version()
    return {"ffmpeg": avu.av_version_info().decode(), "libswscale": f"{v >> 16}.{(v >> 8) & 255}.{v & 255}"}


def coefficients(matrix: str):
    """The 4-entry inverse table {crv, cbu, cgu, cgv} swscale uses for ``matrix``."""
    _, sws = _load()
    p = sws.sws_getCoefficients(SWS_CS[matrix])
    return [int(p[i]) for i in range(4)]


class SwsConverter:
    """One same-size pixel-format conversion context (single-threaded, re-usable)."""

    def __init__(self, w, h, src_fmt, dst_fmt, src_range=0, dst_range=0,
                 src_matrix="bt709", dst_matrix="bt709", flags=SWS_BICUBIC):
        avu, sws = _load()
        self._sws = sws
        self.w, self.h, self.src_fmt, self.dst_fmt = w, h, src_fmt, dst_fmt
        ctx = sws.sws_alloc_context()
        if not ctx:
            raise MemoryError("sws_alloc_context")
        self.ctx = ctypes.c_void_p(ctx)
        sf, df = avu.av_get_pix_fmt(src_fmt.encode()), avu.av_get_pix_fmt(dst_fmt.encode())
        if sf < 0 or df < 0:
            raise ValueError(f"unknown pix_fmt {src_fmt}/{dst_fmt}")
        for k, v in (("srcw", w), ("srch", h), ("dstw", w), ("dsth", h), ("src_format", sf), ("dst_format", df),
                     ("src_range", src_range), ("dst_range", dst_range), ("sws_flags", flags)):
            if avu.av_opt_set_int(self.ctx, k.encode(), v, 0) < 0:
                raise RuntimeError(f"av_opt_set_int({k})")
        if sws.sws_init_context(self.ctx, None, None) < 0:
            raise RuntimeError("sws_init_context failed")
        sws.sws_setColorspaceDetails(self.ctx, sws.sws_getCoefficients(SWS_CS[src_matrix]), src_range,
                                     sws.sws_getCoefficients(SWS_CS[dst_matrix]), dst_range, 0, 1 << 16, 1 << 16)

    def close(self):
        if self.ctx:
            self._sws.sws_freeContext(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _padded(rows, row_bytes):
        """A 64-byte aligned, row-padded scratch plane (swscale's SIMD reads/writes past the row end)."""
        stride = (row_bytes + 63) // 64 * 64 + 64
        raw = np.zeros(stride * (rows + 2) + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        return raw, raw[off:off + stride * rows].reshape(rows, stride), stride

    def convert(self, src_planes, dst_planes):
        """src/dst: lists of 2-D numpy arrays (planes; packed rgb24 = one (h, 3w) plane)."""
        keep, sp_, ss_, dp_, ds_, dviews = [], [], [], [], [], []
        for p in src_planes:
            rb = p.shape[1] * p.itemsize
            raw, view, stride = self._padded(p.shape[0], rb)
            view[:, :rb] = np.ascontiguousarray(p).view(np.uint8).reshape(p.shape[0], rb)
            keep.append(raw)
            sp_.append(view.ctypes.data)
            ss_.append(stride)
        for p in dst_planes:
            rb = p.shape[1] * p.itemsize
            raw, view, stride = self._padded(p.shape[0], rb)
            keep.append(raw)
            dviews.append((view, rb))
            dp_.append(view.ctypes.data)
            ds_.append(stride)
        sp = (ctypes.c_void_p * 4)(*(sp_ + [None] * (4 - len(sp_))))
        ss = (ctypes.c_int * 4)(*(ss_ + [0] * (4 - len(ss_))))
        dp = (ctypes.c_void_p * 4)(*(dp_ + [None] * (4 - len(dp_))))
        ds = (ctypes.c_int * 4)(*(ds_ + [0] * (4 - len(ds_))))
        r = self._sws.sws_scale(self.ctx, sp, ss, 0, self.h, dp, ds)
        if r != self.h:
            raise RuntimeError(f"sws_scale returned {r}, expected {self.h}")
        for p, (view, rb) in zip(dst_planes, dviews):
            p[...] = view[:, :rb].copy().view(p.dtype).reshape(p.shape)
        return dst_planes


def plane_shapes(fmt, w, h):
    """[(rows, cols, dtype)] of each plane of ``fmt``; gbrp* planes are in FFmpeg order G,B,R."""
    t = {
        "yuv420p": [(h, w, np.uint8), (h // 2, w // 2, np.uint8), (h // 2, w // 2, np.uint8)],
        "yuv422p": [(h, w, np.uint8), (h, w // 2, np.uint8), (h, w // 2, np.uint8)],
        "yuv444p": [(h, w, np.uint8)] * 3,
        "yuv420p10le": [(h, w, np.uint16), (h // 2, w // 2, np.uint16), (h // 2, w // 2, np.uint16)],
        "yuv422p10le": [(h, w, np.uint16), (h, w // 2, np.uint16), (h, w // 2, np.uint16)],
        "yuv444p10le": [(h, w, np.uint16)] * 3,
        "rgb24": [(h, 3 * w, np.uint8)],
        "gbrp": [(h, w, np.uint8)] * 3,
        "gbrp10le": [(h, w, np.uint16)] * 3,
    }
    return t[fmt]


def alloc_planes(fmt, w, h):
    return [np.zeros((r, c), dtype=d) for r, c, d in plane_shapes(fmt, w, h)]
