#!/usr/bin/env python
"""Randomised parity soak: many random (layout, geometry, grid, LUT, interp, matrix, range, option) combinations through the C-ABI
against the oracle, bit-exact.  The committed fuzz test (tests/test_gpu_parity.py) is one fixed seed of this; run other seeds here.
    python tests/soak.py [--seed 1] [--trials 200]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import assert_bit_exact, gpu_stage, oracle_lut  # noqa: E402
from lut_renderer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--trials", type=int, default=200)
ap.add_argument("--warp-bias", type=float, default=0.6, help="share of trials drawn from the warp-window kernels' geometry")
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
FMTS = ["yuv420p", "yuv422p", "yuv444p", "yuv420p10le", "yuv422p10le", "yuv444p10le"]
kernels = {}
for trial in range(args.trials):
    in_fmt, out_fmt = str(rng.choice(FMTS)), str(rng.choice(FMTS))
    if rng.random() < 0.5:   # bias towards the combinations that have specialised kernels
        in_fmt, out_fmt = [("yuv420p", "yuv420p"), ("yuv422p10le", "yuv422p10le"), ("yuv420p", "yuv422p10le"), ("yuv420p10le", "yuv420p"),
                           ("yuv420p10le", "yuv420p10le"), ("yuv422p10le", "yuv420p10le"), ("yuv420p10le", "yuv422p10le"),
                           ("yuv422p", "yuv422p"), ("yuv420p", "yuv444p"), ("yuv422p10le", "yuv420p")][int(rng.integers(10))]
    w = int(rng.choice([16, 18, 24, 64, 120, 128, 136, 240, 248, 256, 264, 376, 480, 488, 496, 512, 520, 736, 854, 1032, 1928]))
    h = int(rng.choice([14, 16, 18, 22, 30, 34, 46, 66, 70]))
    interp = str(rng.choice(["tetrahedral", "tetrahedral", "tetrahedral", "trilinear", "trilinear", "nearest", "pyramid", "prism"]))
    if rng.random() < args.warp_bias:   # geometry and interpolation the warp-window kernels take
        w = int(rng.choice([128, 136, 240, 248, 256, 264, 376, 480, 488, 496, 512, 520, 736, 1032, 1928, 2056]))
        h = int(rng.choice([14, 16, 18, 22, 30, 34, 46, 66, 70, 130, 258]))
        interp = str(rng.choice(["tetrahedral", "tetrahedral", "trilinear"]))
    kw = dict(in_matrix=str(rng.choice(["bt709", "bt601", "bt2020"])), out_matrix=str(rng.choice(["bt709", "bt601", "bt2020"])))
    rp = 0
    if rng.random() < 0.3:
        norm = "tv" if rng.random() < 0.6 else "pc"
        kw.update(in_range="pc", norm_range=norm)
        if "10" in in_fmt:
            if rng.random() < 0.6:
                kw.update(via_8bit=True)
                rp = 1 if norm == "tv" else 2
            else:
                kw.update(norm_range="pc")
                norm = "pc"
        else:
            rp = 1 if norm == "tv" else 0
    if rng.random() < 0.2:
        kw.update(out_range="pc")
    options = {}
    if rng.random() < 0.5:
        options["grid_ctas"] = int(rng.integers(1, 12))
    if rng.random() < 0.15:
        options["any_kernel"] = 1
    if rng.random() < 0.1:
        options["wk_minb"] = int(rng.choice([2, 3]))
    if rng.random() < 0.15:
        options["wk_tma"] = int(rng.choice([1, 2]))      # row-staging kernels (where the combination has one)
    if rng.random() < 0.2:
        options["rows_per_cta"] = int(rng.choice([1, 3, 8, 16, 40, 128]))
    n = int(rng.choice([2, 5, 9, 17, 20, 33, 33, 64, 65, 65]))
    table = (synth.look_lut(n) if rng.random() < 0.5 else synth.random_lut(n, int(rng.integers(1 << 20))) * (1.4 if rng.random() < 0.3 else 1.0)
             - (0.2 if rng.random() < 0.2 else 0.0)).astype(np.float32)
    scale = (1.0, 1.0, 1.0) if rng.random() < 0.7 else [(0.5, 1.0, 0.75), (0.5, 0.5, 0.5), (0.8, 0.8, 0.8), (1.0, 0.25, 0.6)][int(rng.integers(4))]
    fm = "fma" if rng.random() < 0.25 else "c"
    nfr = int(rng.integers(1, 4))
    frames = [synth.noise_frame(in_fmt, w, h, seed=int(rng.integers(1 << 30))) if rng.random() < 0.6 else
              synth.natural_frame(in_fmt, w, h, t=trial + i, full_range=kw.get("in_range") == "pc", chroma_grain=0.02) for i in range(nfr)]
    desc = f"trial {trial}: {in_fmt}->{out_fmt} {w}x{h} {interp} n={n} {kw} {options} scale {scale} {fm}"
    got = gpu_stage(in_fmt, w, h, frames, table, scale=scale, out_pix_fmt=out_fmt, interp=interp, options=options, float_mode=fm, info=(info := {}), **kw)
    kernels[info["kernel"]] = kernels.get(info["kernel"], 0) + 1
    L = oracle_lut(table, scale)
    O.set_float_mode(fm)   # the oracle restates both float orders (contracted multiply-adds = FFmpeg's x86 assembly)
    for i, f in enumerate(frames):
        Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(f, in_fmt, w, h)]
        exp = synth.pack_planes(O.stage_any(L, in_fmt, out_fmt, Y, U, V, interp=interp, range_pass=rp, in_matrix=kw["in_matrix"],
                                            in_full_range=kw.get("norm_range") == "pc", out_matrix=kw["out_matrix"],
                                            out_full_range=kw.get("out_range") == "pc"))
        assert_bit_exact(out_fmt, w, h, got[i], exp, desc + f" frame {i}")
    O.set_float_mode("c")
import json  # noqa: E402
print(json.dumps({"soak_seed": args.seed, "trials": args.trials, "result": "bit-exact", "kernels": dict(sorted(kernels.items()))}))
