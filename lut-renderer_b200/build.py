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
SOURCES = ["cube_parser.cpp", "colour_tables.cpp", "stage_kernels.cu", "k422_warp.cu", "lutr_api.cu"]
HEADERS = ["lutr_internal.h", os.path.join("..", "..", "include", "lutr.h")]
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


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [
        nvcc_path(), "-shared", "-o", OUT,
        "-gencode", GENCODE, "-O3", "-lineinfo", "-std=c++17",
        # lut3d's C semantics round every product and every sum: ptxas must never contract a mul + add pair (it does so for the
        # packed mul.rn.f32x2 / add.rn.f32x2 pair unless told not to; wanted fmas are written as fma explicitly)
        "-fmad=false",
        "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wextra,-ffp-contract=off",
        "-cudart", "static",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if os.environ.get("LUTR_WK_EXPERIMENTS"):
        cmd += ["-DLUTR_WK_EXPERIMENTS"]
    if os.environ.get("LUTR_WK_UNROLL"):
        cmd += ["-DLUTR_WK_UNROLL=" + os.environ["LUTR_WK_UNROLL"]]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liblutr.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
