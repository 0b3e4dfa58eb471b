#!/usr/bin/env python
"""What storing the LUT as packed fp16 (north_star's shared-memory option for 33-point tables) would cost in parity: the oracle's
lut3d leg with the table rounded to fp16 against the same leg with the fp32 table, tetrahedral, 2 M random RGB inputs, and the
effect on the final YUV planes of BASELINE cfg 3.  CPU only (oracle/):   python tests/fp16_lut_study.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(0)
res = []
for name, table in (("look33", synth.look_lut(33)), ("look65", synth.look_lut(65)), ("random33", synth.random_lut(33, 2)),
                    ("identity33", synth.identity_lut(33))):
    t32 = table.astype(np.float32)
    t16 = t32.astype(np.float16).astype(np.float32)
    L32, L16 = O.OracleLut.from_table(t32), O.OracleLut.from_table(t16)
    row = {"lut": name}
    for depth in (10, 8):
        hi = (1 << depth) - 1
        R, G, B = (rng.integers(0, hi + 1, (1000, 2000), dtype=np.uint16) for _ in range(3))
        a = np.stack(O.lut3d_planar16(L32, R, G, B, depth, "tetrahedral"))
        b = np.stack(O.lut3d_planar16(L16, R, G, B, depth, "tetrahedral"))
        d = np.abs(a.astype(int) - b.astype(int))
        row[f"rgb{depth}_differing_pct"] = round(100 * float((d != 0).mean()), 3)
        row[f"rgb{depth}_max_abs"] = int(d.max())
    # the whole cfg 3 chain on a natural 4K frame
    f = synth.natural_frame("yuv422p10le", 3840, 2160, t=3)
    Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(f, "yuv422p10le", 3840, 2160)]
    a = O.stage_any(L32, "yuv422p10le", "yuv422p10le", Y, U, V)
    b = O.stage_any(L16, "yuv422p10le", "yuv422p10le", Y, U, V)
    for n_, x, y in zip("YUV", a, b):
        row[f"cfg3_{n_}_identical_pct"] = round(100 * float((x == y).mean()), 3)
    res.append(row)
    print(json.dumps(row), flush=True)
