#!/usr/bin/env python
"""Opcode histogram of a kernel's row loop, read from the SASS of a cubin / shared library (no GPU needed).

    python tools/sass_loop.py <cubin|so> <substring of the mangled kernel name> [--loop N] [--px 8]

Lists the backward branches of the function (a loop = a branch to a lower address) and prints the opcode mix of the N-th
largest loop (default 1 = second largest: the outer persistent loop is the largest), divided by --px pixels per lane and trip.
"""
import collections
import re
import subprocess
import sys


def function_sass(path, pat):
    elf = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout
    names = sorted(set(re.findall(r"_Z[A-Za-z0-9_]+", elf)))
    names = [n for n in names if pat in n and "k" in n]
    if not names:
        raise SystemExit(f"no function matching {pat}")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], path], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    return names[0], ins


def main():
    path, pat = sys.argv[1], sys.argv[2]
    which = int(sys.argv[sys.argv.index("--loop") + 1]) if "--loop" in sys.argv else 1
    px = float(sys.argv[sys.argv.index("--px") + 1]) if "--px" in sys.argv else 8.0
    name, ins = function_sass(path, pat)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr_index:
            loops.append((i - addr_index[int(m.group(1), 16)] + 1, addr_index[int(m.group(1), 16)], i))
    loops.sort(reverse=True)
    print(name)
    print("instructions:", len(ins), " loops (size, from, to):", loops[:4])
    if which >= len(loops):
        which = len(loops) - 1
    n, lo, hi = loops[which]
    hist = collections.Counter()
    for _, text in ins[lo:hi + 1]:
        text = re.sub(r"^@!?U?P\w+\s+", "", text)
        hist[text.split()[0].split(".")[0]] += 1
    print(f"loop {which}: {n} instructions = {n / px:.1f} per pixel")
    print("  ".join(f"{k}:{v}" for k, v in hist.most_common()))
    fma = sum(v for k, v in hist.items() if k in ("IMAD", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "HFMA2"))
    print(f"FMA-pipe instructions {fma} ({fma / px:.1f}/px), packed FP {sum(v for k, v in hist.items() if k in ('FFMA2', 'FMUL2', 'FADD2'))}, "
          f"IMAD {hist['IMAD']}, LSU {sum(v for k, v in hist.items() if k in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'SHFL'))}")


if __name__ == "__main__":
    main()
