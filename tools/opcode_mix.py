#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an `ncu --set full --import-source on` report: thread instructions per pixel and
share of the warp-stall samples, per SASS opcode.   python tools/opcode_mix.py REPORT.ncu-rep PIXELS_IN_LAUNCH > profiles/...txt"""
import collections
import csv
import re
import subprocess
import sys

rep, pixels = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True,
                     check=True).stdout
rows = list(csv.reader(raw.splitlines()))
name, hdr, data = rows[0][1], rows[1], rows[2:]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
total = sum(int(r[ia]) for r in data if r[ia].isdigit())
h, hs = collections.Counter(), collections.Counter()
for r in data:
    if not r[ia].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[isrc])
    op = m.group(2) if m else "?"
    h[op] += int(r[ia])
    hs[op] += int(r[isamp]) if r[isamp].isdigit() else 0
print(f"kernel: {name}")
print(f"warp instructions executed: {total}   pixels in the launch: {int(pixels)}   thread instructions per pixel: {total * 32 / pixels:.2f}")
print(f"{'opcode':10s} {'thr-instr/px':>12s} {'% of instr':>10s} {'% of stall samples':>18s}")
for op, c in h.most_common(30):
    print(f"{op:10s} {c * 32 / pixels:12.2f} {100 * c / total:10.1f} {100 * hs[op] / max(1, sum(hs.values())):18.1f}")
