"""The C-ABI library loads without a GPU, exports every symbol include/lutr.h declares, fails loudly (no CPU fallback),
and its integer coefficient tables equal libswscale's."""
import ctypes
import os
import re

import pytest

from lut_renderer_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lutr.h")).read()
    return sorted(set(re.findall(r"LUTR_API\s+[\w\s\*]+?\b(lutr_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in lutr.h but not exported"
    assert sorted(N.SIGNATURES) == names, set(names) ^ set(N.SIGNATURES)


def test_version_and_status_strings():
    L = N.lib()
    assert L.lutr_version() == 100
    assert L.lutr_status_string(0) == b"ok"
    for code in range(-7, 0):
        assert L.lutr_status_string(code) not in (b"ok", b"unknown status")


def test_struct_layout_matches_header():
    assert ctypes.sizeof(N.StageDesc) == 4 * 16
    assert ctypes.sizeof(N.Planes) == 3 * 8 + 3 * 4 + 4 + 8


def test_tables_match_oracle_and_libswscale():
    from oracle import oracle as O
    L = N.lib()
    for name, mid in (("bt709", 0), ("bt601", 1), ("bt2020", 2), ("fcc", 3), ("smpte240m", 4)):
        for full in (0, 1):
            out = (ctypes.c_int32 * 7)()
            assert L.lutr_tables_yuv2rgb(mid, full, out) == 0
            ref = O.yuv2rgb_coeffs(name, full)
            assert list(out) == [ref[k] for k in ("yC", "yO8", "vr", "ub", "ug", "vg", "yOff10")]
        out9 = (ctypes.c_int32 * 9)()
        assert L.lutr_tables_rgb2yuv(mid, out9) == 0
        assert list(out9) == O.rgb2yuv_coeffs(name)
    assert L.lutr_tables_rgb2yuv(9, out9) == N.ERR_INVALID_ARG


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = N.lib().lutr_ctx_create(0, ctypes.byref(h))
    assert rc == N.ERR_CUDA and not h.value
    assert b"no CPU fallback" in N.lib().lutr_last_error() or b"CUDA" in N.lib().lutr_last_error()
    from lut_renderer_b200.stage import ColourStage
    with pytest.raises(N.LutrError):
        ColourStage(64, 16, "yuv420p")


def test_unsupported_requests_are_named_before_touching_the_device():
    from lut_renderer_b200.stage import ColourStage
    with pytest.raises(N.LutrError) as e:
        ColourStage(64, 16, "yuv420p", interp="cubic")        # in the reference's whitelist, not a lut3d mode (ffmpeg.py:243)
    assert e.value.status == N.ERR_UNSUPPORTED
    with pytest.raises(N.LutrError) as e:
        ColourStage(64, 16, "yuv440p")
    assert e.value.status == N.ERR_UNSUPPORTED
    # every layout / interp of the operator maps to an enum of include/lutr.h
    assert set(N.PIX_FMT) == {"yuv420p", "yuv422p", "yuv444p", "yuv420p10le", "yuv422p10le", "yuv444p10le"}
    assert set(N.INTERP) == {"nearest", "trilinear", "tetrahedral", "pyramid", "prism", "none"}   # none: LUT-less pro stage 2 conversion


def test_header_is_plain_c_and_links(tmp_path):
    """include/lutr.h compiles as C99 and a C program that uses the host-only entry points links against liblutr.so and runs
    (no GPU needed: .cube parsing and the coefficient tables are host code)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "lutr.h"
int main(void) {
    const char *cube = "LUT_3D_SIZE 2\n0 0 0\n1 0 0\n0 1 0\n1 1 0\n0 0 1\n1 0 1\n0 1 1\n1 1 1\n";
    int32_t size = 0, t[9];
    float scale[3], *table = NULL;
    lutr_stage_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    if (lutr_version() != LUTR_VERSION) return 1;
    if (lutr_cube_parse(cube, strlen(cube), &size, scale, &table) != LUTR_OK || size != 2) return 2;
    if (table[3 * 4 + 0] != 1.0f) return 3;                  /* node r=1,g=0,b=0 holds the file's second line */
    lutr_free(table);
    if (lutr_cube_parse("LUT_3D_SIZE 1\n", 14, &size, scale, &table) != LUTR_ERR_LUT_SIZE) return 4;
    if (lutr_tables_rgb2yuv(LUTR_MATRIX_BT709, t) != LUTR_OK || t[0] != 5983 || t[1] != 20127 || t[2] != 2032) return 5;
    printf("%s\n", lutr_status_string(LUTR_ERR_UNSUPPORTED));
    return 0;
}
''')
    exe = tmp_path / "t"
    lib_dir = os.path.join(root, "lut-renderer_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-llutr", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "unsupported configuration"
