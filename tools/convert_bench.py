#!/usr/bin/env python
"""Device-resident throughput of the LUT-less conversion stage (k_convert.cu, pro mode stage 2).   python tools/convert_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from lut_renderer_b200.stage import ColourStage  # noqa: E402

w, h, F = 3840, 2160, 64
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
frame = torch.from_numpy(synth.natural_frame("yuv422p10le", w, h, t=1)).cuda()
src = frame.repeat(F)
for out_fmt in ("yuv420p", "yuv420p10le", "yuv422p"):
    st = ColourStage(w, h, "yuv422p10le", out_pix_fmt=out_fmt, interp="none")
    dst = torch.empty(F * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    for _ in range(3):
        st.run(None, src, dst, F, stream=s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 50
    for _ in range(n):
        st.run(None, src, dst, F, stream=s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = F * (st.in_frame_bytes + st.out_frame_bytes) / ms / 1e6
    print(json.dumps({"stage": f"yuv422p10le -> {out_fmt} (no LUT)", "kernel": st.kernel_name, "fps": round(F / ms * 1e3, 1), "GBs": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / peak, 3)}))
    st.close()
