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
    assert set(N.INTERP) == {"nearest", "trilinear", "tetrahedral", "pyramid", "prism"}
