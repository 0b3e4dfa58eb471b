#!/usr/bin/env python
"""Host <-> device copy rates of this box (pinned memory, 256 MB buffers): H2D alone, D2H alone, both directions at once on two
streams.  The end-to-end path of bench.py (`e2e`) moves one frame in and one frame out per frame, so the simultaneous figure is
its ceiling."""
import json
import time

import torch

n = 256 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


run(True, True, 3)
res = {"h2d_alone_gbs": round(run(True, False), 2), "d2h_alone_gbs": round(run(False, True), 2),
       "each_direction_when_simultaneous_gbs": round(run(True, True), 2), "device": torch.cuda.get_device_name(0)}
print(json.dumps(res))
