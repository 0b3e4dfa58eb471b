#!/usr/bin/env python
"""A handful of small launches of every kernel family, for one compute-sanitizer tool at a time:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from lut_renderer_b200.stage import ColourStage  # noqa: E402

CASES = [  # (source, output, w, h, interp, stage kwargs, options)
    ("yuv422p10le", "yuv422p10le", 496, 22, "tetrahedral", {}, {}),
    ("yuv422p10le", "yuv422p10le", 496, 22, "tetrahedral", {}, {"wk_tma": 1}),
    ("yuv422p10le", "yuv422p10le", 496, 22, "trilinear", {}, {"wk_tma": 2}),
    ("yuv422p10le", "yuv422p10le", 16, 2, "tetrahedral", {}, {"wk_tma": 2}),
    ("yuv444p10le", "yuv422p10le", 248, 9, "tetrahedral", {}, {}),
    ("yuv420p", "yuv420p", 328, 38, "tetrahedral", dict(in_range="pc", norm_range="pc", out_range="pc"), {}),
    ("yuv420p", "yuv420p", 328, 38, "trilinear", dict(in_range="pc", norm_range="tv"), {}),
    ("yuv420p", "yuv422p10le", 256, 32, "tetrahedral", {}, {}),
    ("yuv422p", "yuv444p", 264, 16, "tetrahedral", {}, {}),
    ("yuv420p10le", "yuv422p10le", 256, 32, "tetrahedral", {}, {}),
    ("yuv420p10le", "yuv420p10le", 256, 32, "tetrahedral", {}, {}),
    ("yuv422p10le", "yuv420p10le", 496, 30, "trilinear", {}, {}),
    ("yuv420p10le", "yuv420p", 256, 32, "tetrahedral", {}, {}),
    ("yuv444p", "yuv422p", 134, 14, "prism", {}, {}),
    ("yuv422p10le", "yuv420p", 264, 40, "none", {}, {}),
    ("yuv422p10le", "yuv420p10le", 72, 16, "none", {}, {}),
    ("yuv422p10le", "yuv422p", 16, 14, "none", {}, {}),
]
table = synth.look_lut(17).astype(np.float32)
for in_fmt, out_fmt, w, h, interp, kw, opts in CASES:
    st = ColourStage(w, h, in_fmt, out_pix_fmt=out_fmt, interp=interp, **kw)
    for k, v in opts.items():
        st.set_option(k, v)
    lut = None if interp == "none" else st.ctx.lut_from_table(table)
    frames = np.concatenate([synth.noise_frame(in_fmt, w, h, seed=3), synth.natural_frame(in_fmt, w, h, t=2)])
    src = torch.from_numpy(frames).cuda()
    dst = torch.zeros(2 * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(lut, src, dst, 2, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    print(in_fmt, out_fmt, w, h, interp, opts, st.kernel_name, int(dst.sum().item()))
    st.close()
print("done")
