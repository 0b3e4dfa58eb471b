#!/usr/bin/env python
"""Throughput of the reference-facing runner (TaskRunner: decode process -> raw pipe -> GPU stage -> raw pipe -> encode process)
with stand-in child processes: a generator that writes N raw frames to stdout as fast as the pipe takes them and a sink that
discards its stdin.  What is measured is everything the product adds between the two ffmpeg processes of a real run (pipes, pinned
staging, H2D, kernel, D2H).   python tools/runner_bench.py [--frames 240] [--size 3840x2160] [--fmt yuv422p10le] [--batch 8]"""
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

from lut_renderer_b200 import ProcessingParams, Task, VideoInfo, synth  # noqa: E402
from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=480)
ap.add_argument("--size", default="3840x2160")
ap.add_argument("--fmt", default="yuv422p10le")
ap.add_argument("--batch", type=int, default=0, help="frames per batch of the pump (0 = the library picks ~64 MB)")
ap.add_argument("--shards", type=int, default=1, help="cut the clip into this many frame ranges, each with its own decoder / encoder chain (all on GPU 0)")
args = ap.parse_args()
w, h = (int(v) for v in args.size.split("x"))
tmp = Path(tempfile.mkdtemp(prefix="lutr_runner_"))
cube = synth.write_cube(tmp / "look.cube", synth.look_lut(65 if "10" in args.fmt else 33).astype(np.float32))
frame = synth.natural_frame(args.fmt, w, h, t=1)
(tmp / "frame.raw").write_bytes(frame.tobytes())
depth = 10 if "10" in args.fmt else 8
info = VideoInfo(width=w, height=h, pix_fmt=args.fmt, bit_depth=depth, colorspace="bt709", color_range="tv", fps=25.0, duration=args.frames / 25.0)
params = ProcessingParams(video_codec="prores_ks" if depth == 10 else "libx264")
task = Task("bench", tmp / "in.mov", tmp / "out.mov", cube, None, params, source_info=info)
sink = "f=open(0,'rb',buffering=0); b=bytearray(1<<22)\nwhile f.readinto(b): pass"


def popen(cmd, **kw):
    if "pipe:1" in cmd:      # generator: as many frames as this decoder was asked for
        import math
        start = math.ceil(float(cmd[cmd.index("-ss") + 1]) * 25.0 - 1e-6) if "-ss" in cmd else 0
        count = int(cmd[cmd.index("-frames:v") + 1]) if "-frames:v" in cmd else args.frames - start
        code = f"import sys; b=open({str(tmp / 'frame.raw')!r},'rb').read(); o=sys.stdout.buffer\nfor _ in range({count}): o.write(b)"
    elif "pipe:0" in cmd:    # sink (a part file must exist for the concat step of a sharded run)
        code = sink + f"\nopen({cmd[-1]!r}, 'wb').close()"
    else:                    # concat step of a sharded run
        code = f"open({cmd[-1]!r}, 'wb').close()"
    return subprocess.Popen([sys.executable, "-c", code], **kw)


class Rec(TaskCallbacks):
    status_text = ""

    def status(self, tid, text):
        self.status_text = text


rec = Rec()
t0 = time.perf_counter()
TaskRunner(task, callbacks=rec, frames_per_batch=args.batch, popen=popen, devices=[0] * args.shards if args.shards > 1 else None).run()
dt = time.perf_counter() - t0
print(json.dumps({"frames": args.frames, "size": args.size, "fmt": args.fmt, "batch": args.batch, "shards": args.shards, "status": rec.status_text,
                  "fps_including_startup": round(args.frames / dt, 1), "GBs_in": round(args.frames * frame.nbytes / dt / 1e9, 2)}))
