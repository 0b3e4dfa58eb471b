#!/bin/sh
# Confirms or refutes the one unpinned leg of the oracle (lut3d) against a REAL ffmpeg -- run it anywhere one exists:
#     sh tests/golden/lut3d_vectors/replay_with_ffmpeg.sh [path/to/ffmpeg]
# For every case of manifest.json it feeds the committed raw input through `lut3d=file=<cube>:interp=<interp>` with the same pixel
# format on both sides (no scaler gets inserted: lut3d takes gbrp10le and rgb24 as they are) and compares the SHA-256 of the output
# with the two expectations of oracle/lutr_oracle.c: float_mode c (FFmpeg's C code) and fma (its x86 AVX2+FMA3 assembly, planar only).
# Reference call site: src/lut_renderer/ffmpeg.py:242-247.  Needs: ffmpeg, python3 (json + hashlib only).
FFMPEG=${1:-ffmpeg}
cd "$(dirname "$0")" || exit 2
"$FFMPEG" -hide_banner -version | head -1
python3 - "$FFMPEG" <<'PY'
import hashlib, json, subprocess, sys
ffmpeg = sys.argv[1]
m = json.load(open("manifest.json"))
w, h = m["geometry"]
groups = {}
for c in m["cases"]:
    groups.setdefault((c["cube"], c["interp"], c["input"], c["pix_fmt"]), {})[c["float_mode"]] = c["sha256"]
tally = {"c": 0, "fma": 0, "both": 0, "neither": 0}
for (cube, interp, inp, fmt), exp in sorted(groups.items()):
    out = subprocess.run([ffmpeg, "-v", "error", "-f", "rawvideo", "-pix_fmt", fmt, "-s", f"{w}x{h}", "-i", inp, "-vf",
                          f"lut3d=file={cube}:interp={interp}", "-frames:v", "1", "-f", "rawvideo", "-pix_fmt", fmt, "-"],
                         capture_output=True).stdout
    got = hashlib.sha256(out).hexdigest()
    hit = [k for k, v in exp.items() if v == got]
    verdict = "both" if len(hit) == 2 else (hit[0] if hit else "neither")
    if verdict == "c" and "fma" not in exp:
        verdict = "c"
    tally[verdict] += 1
    print(f"{verdict:8s} {fmt:9s} {interp:12s} {cube:18s} {inp}")
print(tally)
print("all 'c' (or 'both'): the C semantics hold for this ffmpeg;  'fma' on gbrp10le tetrahedral/trilinear: this build runs the x86 assembly;")
print("'neither': the restatement is wrong for that case -- regenerate the expected bytes with gen_lut3d_vectors.py --write-expected and diff.")
PY
