"""Build the C-ABI shared library (liblutr.so) in-tree with nvcc for sm_100a.

    python lut-renderer_b200/build.py [--force] [--verbose]

The library is the drop-in boundary (include/lutr.h); it links the static CUDA runtime only.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblutr.so")
SOURCES = ["cube_parser.cpp", "lut_dialects.cpp", "colour_tables.cpp", "k422_warp.cu", "k422_tma.cu", "k420_warp.cu", "k420_warp10.cu", "k420_warp422.cu", "k420_warp444.cu", "k10_warp.cu", "k10_warp8.cu", "k_any.cu", "k_any_fixed8.cu", "k_any_fixed10.cu", "k_convert.cu", "lutr_api.cu"]
HEADERS = ["lutr_internal.h", "warp_common.cuh", "k420_warp.cuh", "k10_warp.cuh", "lut3d_generic.cuh", "k_any.cuh", os.path.join("..", "..", "include", "lutr.h")]
GENCODE = "arch=compute_100a,code=sm_100a"


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


LAST_BUILD = {"nvcc_commands": 0}   # what the most recent build() call actually ran


def build(force: bool = False, verbose: bool = False) -> str:
    LAST_BUILD["nvcc_commands"] = 0
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = [
        "-gencode", GENCODE, "-O3", "-lineinfo", "-std=c++17",
        # lut3d's C semantics round every product and every sum: ptxas must never contract a mul + add pair (the packed
        # mul.rn.f32x2 / add.rn.f32x2 pair is contracted regardless, see addp2 in csrc/warp_common.cuh)
        "-fmad=false",
        "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wextra,-ffp-contract=off",
    ]
    if verbose:
        common += ["-Xptxas", "-v"]
    if os.environ.get("LUTR_NVCC_FLAGS"):  # experiment switches, e.g. LUTR_NVCC_FLAGS="-DLUTR_PACKED_BLEND=0"
        common += os.environ["LUTR_NVCC_FLAGS"].split()
    if os.environ.get("LUTR_WK_EXPERIMENTS"):
        common += ["-DLUTR_WK_EXPERIMENTS"]
    # one nvcc per translation unit, in parallel (the template-heavy kernel files dominate the build time)
    procs = []
    flag_file = os.path.join(objdir, "flags.txt")   # objects are reused only when they were built with the same flags
    same_flags = os.path.exists(flag_file) and open(flag_file).read() == " ".join(common)
    if not force and same_flags and not needs_build():
        return OUT
    hdr_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if not force and same_flags and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_time, os.path.getmtime(os.path.join(CSRC, src))):
            continue   # this translation unit is up to date
        cmd = [nvcc_path(), "-c", "-o", obj] + common + [os.path.join(CSRC, src)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        LAST_BUILD["nvcc_commands"] += 1
    failed = False
    for src, _, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building liblutr.so")
    with open(flag_file, "w") as f:
        f.write(" ".join(common))
    cmd = [nvcc_path(), "-shared", "-o", OUT, "-gencode", GENCODE, "-cudart", "static"] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    LAST_BUILD["nvcc_commands"] += 1
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking liblutr.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
