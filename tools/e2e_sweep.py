#!/usr/bin/env python
"""Host-path (ColourStage.run_host) throughput against the staging chunk size.  python tools/e2e_sweep.py [--frames 32]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from lut_renderer_b200 import synth
from lut_renderer_b200.stage import ColourStage

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
w, h, fmt = 3840, 2160, "yuv422p10le"
table = synth.look_lut(65).astype(np.float32)
base = [synth.natural_frame(fmt, w, h, t=t) for t in range(4)]
for mb in (32, 64, 128, 256, 512):
    st = ColourStage(w, h, fmt, device=0)
    st.set_option("host_chunk_mb", mb)
    lut = st.ctx.lut_from_table(table)
    n = args.frames
    pin_in, pin_out = st.ctx.pinned_buffer(n * st.in_frame_bytes), st.ctx.pinned_buffer(n * st.out_frame_bytes)
    for i in range(n):
        pin_in.array[i * st.in_frame_bytes:(i + 1) * st.in_frame_bytes] = base[i % 4]
    for _ in range(2):
        st.run_host(lut, pin_in, pin_out, n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.run_host(lut, pin_in, pin_out, n)
    dt = time.perf_counter() - t0
    fps = n * args.steps / dt
    print(json.dumps({"host_chunk_mb": mb, "fps": round(fps, 1), "GBs_each_way": round(fps * st.in_frame_bytes / 1e9, 2)}), flush=True)
    pin_in.close(); pin_out.close(); st.close(); lut.close()
