#!/usr/bin/env python
"""Writes tests/golden/lut3d_vectors/: a replay kit for the one leg of the oracle that nothing in this image can pin -- FFmpeg's
`lut3d` filter (reference call site src/lut_renderer/ffmpeg.py:242-247; no libavfilter exists here).

  *.cube                      the tables (identity, an analytic look, a random table with per-channel DOMAIN_MAX, a table that leaves [0,1])
  in_<name>.gbrp10le.raw      64x32 planar G,B,R 10-bit inputs (random; r == g / g == b / r == b / grey: the tetrahedron boundaries where
  in_<name>.rgb24.raw         upstream's strict `>` compares decide; extremes 0 / max)
  manifest.json               every case with the SHA-256 of what oracle/lutr_oracle.c computes for it: float_mode c = FFmpeg's C code
                              (separate multiply and add), fma = contracted multiply-adds (the x86 AVX2 + FMA3 assembly upstream
                              installs for planar formats)
  exp_*.raw                   the expected bytes themselves, only with --write-expected (not committed: 3 MB; for diffing a mismatch)
  replay_with_ffmpeg.sh       runs every case through a real ffmpeg and says which expectation (c / fma / neither) its output equals

    python tests/golden/gen_lut3d_vectors.py [--write-expected]    (tests/test_lut3d_vectors.py checks the oracle against the manifest)"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lut_renderer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lut3d_vectors")
W, H = 64, 32
CUBES = {
    "identity17": dict(table=lambda: synth.identity_lut(17)),
    "look17": dict(table=lambda: synth.look_lut(17)),
    "rand9_domain": dict(table=lambda: synth.random_lut(9, 4), domain=((0.0, 0.0, 0.0), (2.0, 1.0, 4.0))),
    "wide9": dict(table=lambda: synth.random_lut(9, 5) * 1.6 - 0.3),
}
INTERPS = ("tetrahedral", "trilinear", "nearest", "pyramid", "prism")


def inputs(maxv):
    rng = np.random.default_rng(2026 + maxv)
    rnd = [rng.integers(0, maxv + 1, (H, W), dtype=np.uint16) for _ in range(3)]
    a, b = rng.integers(0, maxv + 1, (H, W), dtype=np.uint16), rng.integers(0, maxv + 1, (H, W), dtype=np.uint16)
    edges = rng.choice(np.array([0, 1, maxv // 2, maxv - 1, maxv], np.uint16), (3, H, W))
    bound = [a.copy(), a.copy(), b.copy()]                       # r == g
    q = H // 4
    bound[1][q:2 * q], bound[2][q:2 * q] = b[q:2 * q], b[q:2 * q]              # g == b
    bound[0][2 * q:3 * q], bound[1][2 * q:3 * q], bound[2][2 * q:3 * q] = a[2 * q:3 * q], b[2 * q:3 * q], a[2 * q:3 * q]   # r == b
    bound[1][3 * q:], bound[2][3 * q:] = a[3 * q:], a[3 * q:]                 # grey
    return {"random": rnd, "boundaries": bound, "extremes": [edges[0], edges[1], edges[2]]}


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith((".raw", ".cube", ".json")):
            os.remove(os.path.join(OUT, f))
    manifest = {"geometry": [W, H], "plane_order": {"gbrp10le": "G, B, R planes of little-endian u16", "rgb24": "packed R, G, B bytes"}, "cases": []}
    luts = {}
    for name, spec in CUBES.items():
        table = spec["table"]().astype(np.float32)
        path = os.path.join(OUT, name + ".cube")
        synth.write_cube(path, table, title=name, domain=spec.get("domain"))
        luts[name] = O.OracleLut.from_path(path)          # parsed from the text, as ffmpeg would
    for fmt, maxv in (("gbrp10le", 1023), ("rgb24", 255)):
        for iname, (R, G, B) in inputs(maxv).items():
            if fmt == "gbrp10le":
                raw = np.stack([G, B, R]).astype("<u2").tobytes()
            else:
                raw = np.stack([R, G, B], axis=-1).astype(np.uint8).tobytes()
            with open(os.path.join(OUT, f"in_{iname}.{fmt}.raw"), "wb") as f:
                f.write(raw)
            for cube, L in luts.items():
                for interp in INTERPS:
                    for mode in ("c", "fma") if interp in ("tetrahedral", "trilinear") else ("c",):
                        O.set_float_mode(mode)
                        if fmt == "gbrp10le":
                            Ro, Go, Bo = O.lut3d_planar16(L, R, G, B, depth=10, interp=interp)
                            exp = np.stack([Go, Bo, Ro]).astype("<u2").tobytes()
                        else:
                            rgb = np.stack([R, G, B], axis=-1).astype(np.uint8).reshape(H, 3 * W)
                            exp = np.ascontiguousarray(O.lut3d_rgb24(L, rgb, interp=interp)).tobytes()
                        O.set_float_mode("c")
                        name = f"exp_{cube}_{interp}_{iname}_{fmt}_{mode}.raw"
                        if "--write-expected" in sys.argv:
                            with open(os.path.join(OUT, name), "wb") as f:
                                f.write(exp)
                        manifest["cases"].append({"cube": cube + ".cube", "interp": interp, "input": f"in_{iname}.{fmt}.raw", "pix_fmt": fmt,
                                                  "float_mode": mode, "expected": name, "sha256": hashlib.sha256(exp).hexdigest()})
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(manifest["cases"]), "cases to", OUT)


if __name__ == "__main__":
    main()
