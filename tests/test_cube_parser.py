"""The product's .cube parser (C++ behind lutr_cube_parse, host only) against the oracle's restatement of vf_lut3d's
parse_cube(), on the dialect corners SURVEY.md App. C lists."""
import numpy as np
import pytest

from lut_renderer_b200 import _native as N
from lut_renderer_b200 import synth
from oracle import oracle as O


def both(text):
    size, scale, table = N.parse_cube_text(text)
    ref = O.OracleLut.from_text(text)
    assert size == ref.size
    assert np.array_equal(table, ref.table)
    assert [np.float32(s) for s in scale] == list(ref.scale)
    return size, scale, table


def fails(text, status):
    with pytest.raises(N.LutrError) as e:
        N.parse_cube_text(text)
    assert e.value.status == status, e.value
    with pytest.raises(ValueError):
        O.OracleLut.from_text(text)


DATA8 = "0 0 0\n1 0 0\n0 1 0\n1 1 0\n0 0 1\n1 0 1\n0 1 1\n1 1 1\n"


def test_plain_and_generated_files():
    both("LUT_3D_SIZE 2\n" + DATA8)
    for n in (2, 5, 17, 33):
        t = synth.random_lut(n, n).astype(np.float32)
        size, _, table = both(synth.cube_text(t, decimals=7))
        assert size == n and np.allclose(table, t, atol=1e-6)
    both(synth.cube_text(synth.look_lut(9), crlf=True))


def test_comments_title_blank_lines_domain_positions():
    txt = ("# header\nTITLE \"x\"\nDOMAIN_MIN 0.5 0.5 0.5\n\nLUT_3D_SIZE 2\n# c\n\n   \nTITLE \"again\"\nDOMAIN_MIN 0 0 0\n"
           "DOMAIN_MAX 2.0 4.0 1.0\n" + DATA8)
    _, scale, _ = both(txt)
    assert scale == [0.5, 0.25, 1.0]          # DOMAIN before SIZE ignored; min parsed but only 1/(max-min) applied
    _, scale, _ = both("DOMAIN_MAX 4 4 4\nLUT_3D_SIZE 2\n" + DATA8)
    assert scale == [1.0, 1.0, 1.0]
    _, scale, _ = both("LUT_3D_SIZE 2\nDOMAIN_MIN 0.5 0 0\nDOMAIN_MAX 1 1 1\n" + DATA8)
    assert scale == [1.0, 1.0, 1.0]           # 1/(1-0.5) = 2 clipped to 1


def test_size_rules():
    fails("LUT_3D_SIZE 1\n0 0 0\n", N.ERR_LUT_SIZE)
    fails("LUT_3D_SIZE 257\n", N.ERR_LUT_SIZE)
    fails("LUT_3D_SIZE\n", N.ERR_LUT_SIZE)
    fails("LUT_3D_SIZE abc\n", N.ERR_LUT_SIZE)
    both("LUT_3D_SIZE 0x2\n" + DATA8)         # strtol base 0
    both("LUT_3D_SIZE 02\n" + DATA8)
    both("LUT_3D_SIZE 2 trailing\n" + DATA8)


def test_invalid_data():
    fails("", N.ERR_INVALID_DATA)                                   # empty LUT
    fails("TITLE \"no size\"\n0 0 0\n", N.ERR_INVALID_DATA)
    fails("LUT_3D_SIZE 2\n" + DATA8[:-6], N.ERR_INVALID_DATA)       # unexpected EOF
    fails("LUT_3D_SIZE 2\n0 0 0\n1 0\n" + DATA8, N.ERR_INVALID_DATA)  # short line
    fails("LUT_3D_SIZE 2\nLUT_3D_INPUT_RANGE 0.0 1.0\n" + DATA8, N.ERR_INVALID_DATA)  # other keywords after SIZE
    fails("LUT_3D_SIZE 2\nDOMAIN_MID 0 0 0\n" + DATA8, N.ERR_INVALID_DATA)


def test_extra_lines_after_table_are_ignored_and_values_outside_unit_range_kept():
    _, _, table = both("LUT_3D_SIZE 2\n" + DATA8.replace("1 1 1", "1.5 -0.25 1e-3") + "9 9 9\ngarbage\n")
    assert np.allclose(table[1, 1, 1], (1.5, -0.25, 1e-3))


def test_long_lines_are_split_like_a_512_byte_fgets():
    pad = " " * 600
    txt = "LUT_3D_SIZE 2\n" + DATA8.replace("1 0 0\n", "1 0 0" + pad + "\n", 1)
    both(txt)   # the tail of the long line is whitespace-only -> skipped as a blank line
    fails("LUT_3D_SIZE 2\n" + "0" + " " * 520 + "0 0\n" + DATA8, N.ERR_INVALID_DATA)


def test_non_finite_rejected_by_product():
    with pytest.raises(N.LutrError) as e:
        N.parse_cube_text("LUT_3D_SIZE 2\n" + DATA8.replace("1 1 1", "nan 1 1"))
    assert e.value.status == N.ERR_INVALID_DATA


def test_parse_file(tmp_path):
    import ctypes
    p = synth.write_cube(tmp_path / "a b'.cube", synth.look_lut(5))
    size = ctypes.c_int32()
    scale = (ctypes.c_float * 3)()
    tab = ctypes.POINTER(ctypes.c_float)()
    assert N.lib().lutr_cube_parse_file(str(p).encode(), ctypes.byref(size), scale, ctypes.byref(tab)) == 0
    assert size.value == 5
    N.lib().lutr_free(tab)
    assert N.lib().lutr_cube_parse_file(str(tmp_path / "missing.cube").encode(), ctypes.byref(size), scale, ctypes.byref(tab)) == N.ERR_IO


def test_large_files_take_the_parallel_reader_with_identical_results():
    """Files above 256 KB are parsed by several threads (pieces cut at line boundaries); every dialect corner must come out as
    from the sequential reader: DOMAIN lines in the middle of the data (also across pieces), a DOMAIN line / garbage behind the
    last entry (never read), comments and blank lines, CRLF, and the error cases (which fall back to the sequential reader)."""
    t = synth.random_lut(40, 6).astype(np.float32)              # 64 000 entries, ~1.7 MB of text
    lines = synth.cube_text(t).split("\n")
    k = len(lines) // 2
    both("\n".join(lines))
    mixed = lines[:]
    mixed.insert(k, "DOMAIN_MAX 2.0 4.0 1.0")
    mixed.insert(k + 5000, "# a comment in the data")
    mixed.insert(k + 9000, "")
    mixed.insert(30, "DOMAIN_MIN 0.5 0 0")
    mixed.insert(k + 20000, 'TITLE "late title"')
    size, scale, _ = both("\n".join(mixed) + "\nDOMAIN_MAX 9 9 9\nthis is not a data line\n")
    assert size == 40 and [round(float(s), 6) for s in scale] == [0.666667, 0.25, 1.0]
    both("\r\n".join(lines))
    bad = lines[:]
    bad[k + 100] = "0.1 oops 0.3"
    fails("\n".join(bad), N.ERR_INVALID_DATA)
    fails("\n".join(lines[:k]), N.ERR_INVALID_DATA)               # unexpected EOF
    bad = lines[:]
    bad[k] = "DOMAIN_MID 0 0 0"
    fails("\n".join(bad), N.ERR_INVALID_DATA)
    long_line = lines[:]
    long_line[k] = " " * 600 + long_line[k]                       # crosses the 511-byte fgets window: sequential rules apply
    size, scale, table = N.parse_cube_text("\n".join(long_line))
    ref = O.OracleLut.from_text("\n".join(long_line))
    assert np.array_equal(table, ref.table)
