#!/usr/bin/env python
"""GPU experiment driver for the 4:2:2 10-bit kernels: every variant is first checked bit-for-bit (against the oracle on a
small clip, against the parity-tested TMA kernel on full frames), then timed with CUDA events.

    python tools/sweep422.py [--frames 16] [--iters 10] [--w 3840 --h 2160] [--content natural|noise] [--variants JSON]

Prints one JSON line per variant.  Test infrastructure only (uses oracle/ as the checker).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from lut_renderer_b200.stage import ColourStage  # noqa: E402

DEFAULT_VARIANTS = [
    dict(idx_mode=2),
    dict(idx_mode=3, wk_tex=0, wk_pk=1, wk_minb=3),
    dict(idx_mode=3, wk_tex=0, wk_pk=0, wk_minb=3),
    dict(idx_mode=3, wk_tex=0, wk_pk=1, wk_minb=2),
    dict(idx_mode=3, wk_tex=0, wk_pk=1, wk_minb=4),
    dict(idx_mode=3, wk_tex=0, wk_pk=0, wk_minb=2),
    dict(idx_mode=3, wk_tex=0, wk_pk=0, wk_minb=4),
    dict(idx_mode=3, wk_tex=2, wk_pk=1, wk_minb=3),
    dict(idx_mode=3, wk_tex=4, wk_pk=1, wk_minb=3),
    dict(idx_mode=3, wk_tex=2, wk_pk=1, wk_minb=2),
    dict(idx_mode=3, wk_tex=4, wk_pk=1, wk_minb=2),
]


FMT = "yuv422p10le"
STAGE_KW = {}


def make_stage(w, h, interp, opts):
    st = ColourStage(w, h, FMT, interp=interp, device=0, **STAGE_KW)
    for k, v in opts.items():
        st.set_option(k, v)
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--lut", type=int, default=65)
    ap.add_argument("--interp", default="tetrahedral")
    ap.add_argument("--content", default="natural")
    ap.add_argument("--variants", default=None, help="JSON list of option dicts")
    ap.add_argument("--skip-oracle", action="store_true")
    ap.add_argument("--fmt", default="yuv422p10le", choices=["yuv422p10le", "yuv420p"])
    ap.add_argument("--range-pass", action="store_true", help="full-range source, pc->tv pass in front (cfg 2)")
    args = ap.parse_args()
    global FMT
    FMT = args.fmt
    if args.range_pass:
        STAGE_KW.update(in_range="pc", norm_range="tv")
    variants = json.loads(args.variants) if args.variants else DEFAULT_VARIANTS
    dev = torch.device("cuda:0")
    table = synth.look_lut(args.lut).astype(np.float32)

    # small-clip oracle check geometry: several windows, both row ends, a ragged last window
    small = None
    if not args.skip_oracle:
        from oracle import oracle as O
        sw, sh = 744, 38
        frames = [synth.natural_frame(FMT, sw, sh, t=3, full_range=args.range_pass), synth.noise_frame(FMT, sw, sh, seed=5)]
        L = O.OracleLut.from_table(table)
        exp = []
        for fr in frames:
            Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(fr, FMT, sw, sh)]
            if FMT == "yuv422p10le":
                exp.append(synth.pack_planes(O.stage_yuv422p10(L, Y, U, V, interp=args.interp)))
            else:
                exp.append(synth.pack_planes(O.stage_yuv420p(L, Y, U, V, interp=args.interp, range_pass=args.range_pass)))
        small = (sw, sh, np.concatenate(frames), np.concatenate(exp))

    w, h, n = args.w, args.h, args.frames
    host = []
    for t in range(n):
        if args.content == "natural":
            host.append(synth.natural_frame(FMT, w, h, t=t % 16, full_range=args.range_pass) if t < 16 else np.roll(host[t % 16], 8 * (t // 16)))
        else:
            host.append(synth.noise_frame(FMT, w, h, seed=100 + t))
    src = torch.from_numpy(np.concatenate(host)).to(dev)
    fb = src.numel() // n
    ref = None
    for opts in variants:
        rec = dict(opts=opts)
        try:
            if small is not None:
                sw, sh, sfr, sexp = small
                st = make_stage(sw, sh, args.interp, opts)
                lut = st.ctx.lut_from_table(table)
                d = torch.zeros(sexp.size, dtype=torch.uint8, device=dev)
                st.run(lut, torch.from_numpy(sfr).to(dev), d, 2, stream=torch.cuda.current_stream())
                torch.cuda.synchronize()
                got = d.cpu().numpy()
                rec["oracle_mismatch_bytes"] = int((got != sexp).sum())
                st.close()
                lut.close()
            st = make_stage(w, h, args.interp, opts)
            lut = st.ctx.lut_from_table(table)
            dst = torch.zeros(n * st.out_frame_bytes, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream()
            st.run(lut, src, dst, n, stream=stream)
            torch.cuda.synchronize()
            if ref is None:
                ref = dst.clone()
                rec["vs_first_mismatch_bytes"] = 0
            else:
                rec["vs_first_mismatch_bytes"] = int((dst != ref).sum().item())
            for _ in range(2):
                st.run(lut, src, dst, n, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.iters):
                st.run(lut, src, dst, n, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            rec["ms_per_launch"] = round(ms, 4)
            rec["fps"] = round(n / ms * 1e3, 1)
            rec["gpix_s"] = round(n * w * h / ms / 1e6, 2)
            rec["hbm_gbs"] = round(n * w * h * (8 if FMT == "yuv422p10le" else 3) / ms / 1e6, 1)
            rec["kernel"] = st.kernel_name
            st.close()
            lut.close()
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
