#!/usr/bin/env python
"""Throughput of lutr_stage_pump alone (no TaskRunner, start-up excluded): a generator child writes N raw frames into a pipe, the
pump runs them through the stage, a sink child discards them.  Also measures the bare pipe (generator -> sink, no pump) as the
ceiling of the host side.   python tools/pump_bench.py [--frames 480] [--size 3840x2160] [--fmt yuv422p10le] [--batch 0]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from lut_renderer_b200.stage import ColourStage  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=480)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--fmt", default="yuv422p10le")
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
w, h = (int(v) for v in args.size.split("x"))
tmp = Path(tempfile.mkdtemp(prefix="lutr_pump_"))
frame = synth.natural_frame(args.fmt, w, h, t=1)
(tmp / "frame.raw").write_bytes(frame.tobytes())
gen = f"import sys; b=open({str(tmp / 'frame.raw')!r},'rb').read(); o=sys.stdout.buffer\nfor _ in range({args.frames}): o.write(b)"
sink = "f=open(0,'rb',buffering=0); b=bytearray(1<<22)\nwhile f.readinto(b): pass"

# bare pipe: generator -> sink
t0 = time.perf_counter()
g = subprocess.Popen([sys.executable, "-c", gen], stdout=subprocess.PIPE, bufsize=0)
s = subprocess.Popen([sys.executable, "-c", sink], stdin=g.stdout)
g.stdout.close()
s.wait(), g.wait()
bare = args.frames * frame.nbytes / (time.perf_counter() - t0) / 1e9

st = ColourStage(w, h, args.fmt, device=0)
lut = st.ctx.lut_from_table(synth.look_lut(65 if "10" in args.fmt else 33).astype(np.float32))
g = subprocess.Popen([sys.executable, "-c", gen], stdout=subprocess.PIPE, bufsize=0)
s = subprocess.Popen([sys.executable, "-c", sink], stdin=subprocess.PIPE, bufsize=0)
stats = st.pump(lut, g.stdout.fileno(), s.stdin.fileno(), args.batch)
s.stdin.close()
s.wait(), g.wait()
print(json.dumps({"frames": stats["frames_out"], "size": args.size, "fmt": args.fmt, "batch": args.batch, "pipe_mb": os.environ.get("LUTR_PIPE_MB", "max"),
                  "pump_fps": round(stats["frames_out"] / stats["seconds"], 1), "pump_GBs_each_way": round(stats["bytes_in"] / stats["seconds"] / 1e9, 2),
                  "bare_pipe_GBs": round(bare, 2)}))
