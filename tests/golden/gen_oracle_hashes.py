#!/usr/bin/env python
"""Writes tests/golden/oracle_hashes.json: SHA-256 of the oracle's output planes for a fixed set of seeded inputs, one entry per
cell of the format matrix x interpolation mode x range handling that matters.  The oracle's swscale legs are pinned live against
libswscale (tests/test_oracle_swscale.py); these hashes additionally freeze the COMPOSITE (incl. the restated lut3d leg), so that
an edit to oracle/ or a different libm / compiler that changes a single output byte is noticed.
    python tests/golden/gen_oracle_hashes.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

W, H = 200, 36
CASES = [(i, o, interp, rp, ofr)
         for i, o in (("yuv420p", "yuv420p"), ("yuv422p10le", "yuv422p10le"), ("yuv420p", "yuv422p10le"), ("yuv420p10le", "yuv420p"),
                      ("yuv420p10le", "yuv420p10le"), ("yuv444p", "yuv422p"), ("yuv422p", "yuv444p10le"), ("yuv444p10le", "yuv420p10le"))
         for interp, rp, ofr in (("tetrahedral", 0, 0), ("trilinear", 1, 0), ("prism", 0, 1), ("nearest", 2, 0), ("pyramid", 0, 0))]


def compute():
    out = {}
    luts = {"look17": synth.look_lut(17), "rand9": synth.random_lut(9, 4) * 1.4 - 0.2}
    for name, t in luts.items():
        L = O.OracleLut.from_table(t.astype(np.float32), (1.0, 1.0, 1.0) if name == "look17" else (0.5, 1.0, 0.75))
        for in_fmt, out_fmt, interp, rp, ofr in CASES:
            if rp == 2 and "10" not in in_fmt:
                rp_eff = 0
            else:
                rp_eff = rp
            frame = synth.noise_frame(in_fmt, W, H, seed=17)
            Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(frame, in_fmt, W, H)]
            planes = O.stage_any(L, in_fmt, out_fmt, Y, U, V, interp=interp, range_pass=rp_eff, in_full_range=(rp_eff == 2),
                                 out_full_range=bool(ofr))
            h = hashlib.sha256()
            for p in planes:
                h.update(np.ascontiguousarray(p).tobytes())
            out[f"{name}|{in_fmt}->{out_fmt}|{interp}|range_pass={rp_eff}|out_full={ofr}"] = h.hexdigest()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_hashes.json")
    with open(path, "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
    print("wrote", path)
