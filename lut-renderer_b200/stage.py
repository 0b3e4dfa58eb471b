"""Python face of the fused GPU colour stage: device context, LUT cache, stage handle.

``ColourStage`` is what ``task_manager.TaskRunner`` drives in place of the ``lut3d`` filter-graph node
(reference src/lut_renderer/ffmpeg.py:246 executed at task_manager.py:145-151).  Frames are raw planar
buffers; device buffers are passed as integer addresses (``tensor.data_ptr()``), host buffers as anything
exposing the buffer protocol.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import byref, c_int32, c_size_t, c_void_p
from pathlib import Path
from typing import Dict, Optional, Tuple

from . import _native as N
from .ffmpeg import GPU_INTERPS, ColourStagePlan

_FMT_BYTES_PER_PIXEL = {"yuv420p": 1.5, "yuv422p": 2.0, "yuv444p": 3.0, "yuv420p10le": 3.0, "yuv422p10le": 4.0, "yuv444p10le": 6.0}


def frame_bytes(pix_fmt: str, width: int, height: int) -> int:
    return int(_FMT_BYTES_PER_PIXEL[pix_fmt] * width * height)


class DeviceContext:
    """One per GPU and process (lutr_ctx)."""
    _lock = threading.Lock()
    _by_device: Dict[int, "DeviceContext"] = {}

    def __init__(self, device: int = 0):
        h = c_void_p()
        N.check(N.lib().lutr_ctx_create(int(device), byref(h)))
        self._h = h
        self.device = int(device)
        dev, sms, maj, mnr = c_int32(), c_int32(), c_int32(), c_int32()
        N.check(N.lib().lutr_ctx_info(h, byref(dev), byref(sms), byref(maj), byref(mnr)))
        self.sm_count, self.compute_capability = sms.value, (maj.value, mnr.value)
        self._luts: Dict[Tuple, "DeviceLut"] = {}

    @classmethod
    def get(cls, device: int = 0) -> "DeviceContext":
        with cls._lock:
            ctx = cls._by_device.get(device)
            if ctx is None:
                ctx = cls._by_device[device] = cls(device)
            return ctx

    def lut_from_cube(self, path) -> "DeviceLut":
        """Upload a .cube once per device; cached by (path, mtime, size) so clips sharing a LUT share the upload."""
        p = Path(path)
        st = os.stat(p)
        key = (str(p.resolve()), st.st_mtime_ns, st.st_size)
        with self._lock:
            lut = self._luts.get(key)
        if lut is None:
            h = c_void_p()
            N.check(N.lib().lutr_lut_load_cube(self._h, os.fsencode(str(p)), byref(h)))
            lut = DeviceLut(self, h)
            with self._lock:
                self._luts[key] = lut
        return lut

    def lut_from_file(self, path) -> "DeviceLut":
        """Any LUT file lut3d reads (.cube .dat .3dl .m3d .csp), the reader chosen by the extension as lut3d does; same cache."""
        p = Path(path)
        if p.suffix.lower() == ".cube":
            return self.lut_from_cube(p)
        st = os.stat(p)
        key = (str(p.resolve()), st.st_mtime_ns, st.st_size)
        with self._lock:
            lut = self._luts.get(key)
        if lut is None:
            h = c_void_p()
            N.check(N.lib().lutr_lut_load_file(self._h, os.fsencode(str(p)), byref(h)))
            lut = DeviceLut(self, h)
            with self._lock:
                self._luts[key] = lut
        return lut

    def prefetch_lut(self, path, cached: bool = True):
        """Start parsing + uploading a .cube on a background thread (the C calls release the GIL) and return a Future of the
        DeviceLut: a batch queue loads the next clip's LUT while the GPU works on the current clip.  cached=False always
        re-reads the file (a private LUT the caller closes itself)."""
        with self._lock:
            if getattr(self, "_pool", None) is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="lutr-lut")
        if cached:
            return self._pool.submit(self.lut_from_cube, path)

        def load():
            with open(path, "rb") as f:
                return self.lut_from_text(f.read())
        return self._pool.submit(load)

    def lut_from_text(self, text) -> "DeviceLut":
        data = text if isinstance(text, (bytes, bytearray)) else str(text).encode()
        h = c_void_p()
        N.check(N.lib().lutr_lut_load_cube_memory(self._h, bytes(data), len(data), byref(h)))
        return DeviceLut(self, h)

    def lut_from_table(self, table, scale=(1.0, 1.0, 1.0)) -> "DeviceLut":
        """table: float32 array [n][n][n][3] in lut3d order [r][g][b]."""
        import numpy as np
        t = np.ascontiguousarray(table, dtype=np.float32)
        s = np.asarray(scale, dtype=np.float32)
        h = c_void_p()
        N.check(N.lib().lutr_lut_from_table(self._h, int(t.shape[0]), t.ctypes.data, s.ctypes.data, byref(h)))
        return DeviceLut(self, h)

    def pinned_buffer(self, nbytes: int):
        """A pinned host buffer (numpy uint8 view) for the raw-pipe ring; freed with the returned object."""
        return PinnedBuffer(self, nbytes)


class DeviceLut:
    def __init__(self, ctx: DeviceContext, handle: c_void_p):
        self.ctx, self._h = ctx, handle
        self.size = N.lib().lutr_lut_size(handle)

    def close(self):
        if self._h:
            N.lib().lutr_lut_destroy(self._h)
            self._h = None


class PinnedBuffer:
    def __init__(self, ctx: DeviceContext, nbytes: int):
        import numpy as np
        p = c_void_p()
        N.check(N.lib().lutr_host_alloc(ctx._h, nbytes, byref(p)))
        self._ctx, self._p, self.nbytes = ctx, p, nbytes
        self.array = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(p.value), dtype=np.uint8)

    @property
    def address(self) -> int:
        return self._p.value

    def close(self):
        if self._p:
            self.array = None
            N.lib().lutr_host_free(self._ctx._h, self._p)
            self._p = None


def _stream_arg(stream):
    """None -> the stage's own stream; a cudaStream_t handle otherwise (0 = the legacy default stream, cudaStreamLegacy)."""
    if stream is None:
        return c_void_p(None)
    handle = int(getattr(stream, "cuda_stream", stream))
    return c_void_p(handle if handle else 1)


def _address(buf) -> int:
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        return int(buf.data_ptr())
    if hasattr(buf, "ctypes"):
        return int(buf.ctypes.data)
    if isinstance(buf, PinnedBuffer):
        return buf.address
    return ctypes.addressof(ctypes.c_char.from_buffer(buf))


class ColourStage:
    """A configured fused stage (lutr_stage) for one clip geometry / format / option set."""

    def __init__(self, width: int, height: int, pix_fmt: str = "yuv422p10le", out_pix_fmt: Optional[str] = None,
                 in_range: str = "tv", norm_range: Optional[str] = None, in_matrix: str = "bt709", out_matrix: str = "bt709",
                 out_range: str = "tv", interp: str = "tetrahedral", float_mode: str = "c", device: int = 0, via_8bit: bool = False):
        out_pix_fmt = out_pix_fmt or pix_fmt
        if interp not in GPU_INTERPS and interp != "none":   # "none": LUT-less conversion of the master (pro mode stage 2)
            raise N.LutrError(N.ERR_UNSUPPORTED, f"interp={interp} is not implemented on the GPU stage (have {GPU_INTERPS})")
        for name, table, value in (("pix_fmt", N.PIX_FMT, pix_fmt), ("out_pix_fmt", N.PIX_FMT, out_pix_fmt)):
            if value not in table:
                raise N.LutrError(N.ERR_UNSUPPORTED, f"{name}={value} is not implemented on the GPU stage (have {sorted(table)})")
        self.ctx = DeviceContext.get(device)
        self.width, self.height, self.pix_fmt, self.out_pix_fmt = width, height, pix_fmt, out_pix_fmt
        d = N.StageDesc()
        d.struct_size = ctypes.sizeof(N.StageDesc)
        d.width, d.height = width, height
        d.in_fmt, d.out_fmt = N.PIX_FMT[pix_fmt], N.PIX_FMT[out_pix_fmt]
        d.in_range, d.norm_range = N.RANGE[in_range], N.RANGE[norm_range or in_range]
        d.in_matrix, d.out_matrix, d.out_range = N.MATRIX[in_matrix], N.MATRIX[out_matrix], N.RANGE[out_range]
        d.interp, d.float_mode = N.INTERP[interp], N.FLOAT_MODE[float_mode]
        d.via_8bit = 1 if via_8bit else 0
        h = c_void_p()
        N.check(N.lib().lutr_stage_create(self.ctx._h, byref(d), byref(h)))
        self._h = h
        a, b = c_size_t(), c_size_t()
        N.check(N.lib().lutr_stage_frame_bytes(h, byref(a), byref(b)))
        self.in_frame_bytes, self.out_frame_bytes = a.value, b.value

    @classmethod
    def from_plan(cls, plan: ColourStagePlan, device: int = 0, float_mode: str = "c") -> "ColourStage":
        return cls(plan.width, plan.height, plan.in_pix_fmt, plan.out_pix_fmt, plan.in_range, plan.norm_range,
                   plan.in_matrix, plan.out_matrix, plan.out_range, plan.interp, float_mode, device, via_8bit=plan.via_8bit)

    # -- device-resident frames ------------------------------------------------------------------------------
    def run(self, lut: DeviceLut, src, dst, nframes: int = 1, stream=None) -> None:
        """Enqueue the fused kernel on tightly packed device frames (src/dst: tensors or addresses). Asynchronous."""
        pin, pout = N.Planes(), N.Planes()
        N.check(N.lib().lutr_stage_packed_planes(self._h, 0, _address(src), byref(pin)))
        N.check(N.lib().lutr_stage_packed_planes(self._h, 1, _address(dst), byref(pout)))
        N.check(N.lib().lutr_stage_run(self._h, (lut._h if lut is not None else None), byref(pin), byref(pout), int(nframes), _stream_arg(stream)))

    def run_planes(self, lut: DeviceLut, in_planes, in_strides, out_planes, out_strides, nframes: int = 1,
                   in_frame_step: int = 0, out_frame_step: int = 0, stream=None) -> None:
        pin, pout = N.Planes(), N.Planes()
        for i in range(3):
            pin.data[i], pin.stride[i] = _address(in_planes[i]), int(in_strides[i])
            pout.data[i], pout.stride[i] = _address(out_planes[i]), int(out_strides[i])
        pin.frame_step, pout.frame_step = int(in_frame_step), int(out_frame_step)
        N.check(N.lib().lutr_stage_run(self._h, (lut._h if lut is not None else None), byref(pin), byref(pout), int(nframes), _stream_arg(stream)))

    # -- host frames (the raw-pipe path) ---------------------------------------------------------------------
    def run_host(self, lut: DeviceLut, src, dst, nframes: int = 1) -> None:
        """H2D -> kernel -> D2H for tightly packed host frames; returns when ``dst`` is complete."""
        N.check(N.lib().lutr_stage_run_host(self._h, (lut._h if lut is not None else None), _address(src), _address(dst), int(nframes)))

    def pump(self, lut: DeviceLut, fd_in: int, fd_out: int, frames_per_batch: int = 0, cancel=None, on_frames=None) -> dict:
        """Stream raw frames from ``fd_in`` through the stage into ``fd_out`` until end of file (lutr_stage_pump: reader thread,
        pinned ring, H2D / kernel / D2H, writer thread, all inside the library and outside the interpreter lock).

        ``cancel``: a ``ctypes.c_int32`` the caller may set to 1 from another thread; ``on_frames(total)`` is called after every
        batch.  Returns the pump statistics; raises ``LutrError`` (status ERR_CANCELLED / ERR_IO / ERR_CUDA) otherwise."""
        stats = N.PumpStats()
        cb = N.PROGRESS_FN((lambda user, total: on_frames(int(total))) if on_frames else (lambda user, total: None))
        flag = cancel if cancel is not None else ctypes.c_int32(0)
        N.check(N.lib().lutr_stage_pump(self._h, (lut._h if lut is not None else None), int(fd_in), int(fd_out), int(frames_per_batch), byref(flag), cb, None, byref(stats)))
        return {"frames_in": stats.frames_in, "frames_out": stats.frames_out, "bytes_in": stats.bytes_in, "bytes_out": stats.bytes_out,
                "seconds": stats.seconds}

    def sync(self) -> None:
        N.check(N.lib().lutr_stage_sync(self._h))

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().lutr_stage_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(N.lib().lutr_stage_launch_count(self._h))

    @property
    def kernel_name(self) -> str:
        """Kernel family the most recent enqueue launched (diagnostics / bench.py's roofline.kernel)."""
        return (N.lib().lutr_stage_kernel_name(self._h) or b"").decode()

    def close(self):
        if self._h:
            N.lib().lutr_stage_destroy(self._h)
            self._h = None
