"""ORACLE / TEST INFRASTRUCTURE ONLY -- the reference's CPU path, timed as the reported baseline.

The literal reference path is `ffmpeg -threads N ... -vf <chain>,lut3d=...` (task_manager.py:145-151); no ffmpeg
binary exists in this image, so the same chain is assembled from what does exist:
  leg 1  YUV->RGB     real libswscale 9.1.100 (FFmpeg 8.0.1, x86 SIMD), vf_scale-style context
  leg 2  lut3d        oracle/lutr_oracle.c restatement of vf_lut3d's C code (gcc -O2, no FMA contraction)
  leg 3  RGB->YUV     real libswscale
frame-parallel over worker threads (ctypes releases the GIL), one set of contexts and scratch per worker.
If libswscale cannot be loaded the legs fall back to the oracle's own integer models (kind stays "port").
"""
from __future__ import annotations

import os
import threading
import time

import numpy as np

from . import oracle as O
from . import swscale_real as S


class CpuStage:
    """One worker's contexts for a given configuration."""

    def __init__(self, pix_fmt, w, h, lut, interp="tetrahedral", in_matrix="bt709", out_matrix="bt709", use_swscale=True,
                 out_pix_fmt=None):
        self.pix_fmt, self.w, self.h, self.lut, self.interp = pix_fmt, w, h, lut, interp
        self.out_pix_fmt = out_pix_fmt or pix_fmt
        self.in_matrix, self.out_matrix = in_matrix, out_matrix
        self.use_swscale = use_swscale and S.available()
        if self.use_swscale:
            mid = "gbrp10le" if "10" in pix_fmt else "rgb24"
            self.fwd = S.SwsConverter(w, h, pix_fmt, mid, src_matrix=in_matrix, dst_matrix=in_matrix)
            self.back = S.SwsConverter(w, h, mid, self.out_pix_fmt, src_matrix=out_matrix, dst_matrix=out_matrix)
            self.mid = S.alloc_planes(mid, w, h)
            self.out = S.alloc_planes(self.out_pix_fmt, w, h)

    def run(self, planes):
        Y, U, V = planes
        if not self.use_swscale:
            return O.stage_any(self.lut, self.pix_fmt, self.out_pix_fmt, Y, U, V, self.interp, False, self.in_matrix, False, self.out_matrix)
        self.fwd.convert([Y, U, V], self.mid)
        if "10" in self.pix_fmt:
            g, b, r = self.mid
            r2, g2, b2 = O.lut3d_planar16(self.lut, r, g, b, 10, self.interp)
            self.back.convert([g2, b2, r2], self.out)
        else:
            rgb = O.lut3d_rgb24(self.lut, self.mid[0], self.interp)
            self.back.convert([rgb], self.out)
        return self.out


def time_cpu_path(pix_fmt, w, h, lut, frames, seconds=12.0, threads=None, interp="tetrahedral", use_swscale=True, out_pix_fmt=None):
    """Run the CPU chain frame-parallel for about `seconds`; returns dict(fps, frames, seconds, threads, legs)."""
    threads = threads or os.cpu_count() or 1
    from lut_renderer_b200 import synth  # frame splitting only (numpy views)
    plane_sets = [[np.ascontiguousarray(p) for p in synth.split_planes(f, pix_fmt, w, h)] for f in frames]
    stages = [CpuStage(pix_fmt, w, h, lut, interp, use_swscale=use_swscale, out_pix_fmt=out_pix_fmt) for _ in range(threads)]
    for st in stages[:1]:
        st.run(plane_sets[0])  # warm-up: page in code and tables
    counts = [0] * threads
    start_evt = threading.Event()
    deadline = [0.0]

    def worker(i):
        start_evt.wait()
        k = i
        while time.perf_counter() < deadline[0]:
            stages[i].run(plane_sets[k % len(plane_sets)])
            counts[i] += 1
            k += threads

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    t0 = time.perf_counter()
    deadline[0] = t0 + seconds
    start_evt.set()
    for t in ths:
        t.join()
    elapsed = time.perf_counter() - t0
    total = sum(counts)
    return dict(fps=total / elapsed, frames=total, seconds=elapsed, threads=threads,
                legs="libswscale-9.1.100+lut3d-port" if stages[0].use_swscale else "oracle-port")
