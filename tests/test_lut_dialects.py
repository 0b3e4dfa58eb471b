"""The LUT file dialects lut3d reads besides .cube (.dat .3dl .m3d .csp; reference ffmpeg.py:246 hands the path to lut3d, which picks
the reader by extension): the product readers (csrc/lut_dialects.cpp, through the C-ABI) against the independent restatement in
oracle/lut_dialects.py, on files written here.  Both are restated from memory of vf_lut3d.c -- this pins them to each other and
documents the rules, not to FFmpeg (no libavfilter in this image)."""
import numpy as np
import pytest

from lut_renderer_b200 import _native as N
from oracle import lut_dialects as D


def _table(n, seed=0, lo=0.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, (n, n, n, 3)).astype(np.float32)


def _rows(t, red_fastest):
    n = t.shape[0]
    if red_fastest:
        return [t[r, g, b] for b in range(n) for g in range(n) for r in range(n)]
    return list(t.reshape(-1, 3))


def write_dat(path, t, with_size=True, comments=True):
    lines = ["# made by a test", ""] if comments else []
    if with_size:
        lines.append(f"3DLUTSIZE {t.shape[0]}")
    for i, v in enumerate(_rows(t, False)):
        if comments and i == 5:
            lines += ["   ", "# in the middle"]
        lines.append("%.7f %.7f %.7f" % tuple(v))
    path.write_text("\n".join(lines) + "\n")


def write_3dl(path, t12):
    lines = ["# lustre", " ".join(str(i * 64) for i in range(17))]
    lines += ["%d %d %d" % tuple(v) for v in t12.reshape(-1, 3)]
    path.write_text("\r\n".join(lines) + "\r\n")


def write_m3d(path, t, out=4096, order="r g b"):
    n = t.shape[0]
    cols = [order.split().index(c) for c in "rgb"]     # column of each channel in the file
    lines = ["name test", f"in {n ** 3}", f"out {out}", "format lut", f"values {order}"]
    for v in t.reshape(-1, 3):
        row = [0.0] * 3
        for c in range(3):
            row[cols[c]] = float(v[c])
        lines.append("%.3f %.3f %.3f" % tuple(row))
    path.write_text("\n".join(lines) + "\n")


def write_csp(path, t, in_rng=((0, 1), (0, 1), (0, 1)), out_rng=((0, 1), (0, 1), (0, 1)), metadata=True, prelut=False):
    n = t.shape[0]
    lines = ["CSPLUTV100", "3D", ""]
    if metadata:
        lines += ["BEGIN METADATA", "made by a test", "END METADATA", ""]
    for c in range(3):
        if prelut and c == 0:
            lines += ["4", "0.0 0.25 0.5 1.0", "0.0 0.3 0.6 1.0"]
        else:
            lines += ["2", "%g %g" % in_rng[c], "%g %g" % out_rng[c]]
    lines += ["", f"{n} {n} {n}"]
    lines += ["%.6f %.6f %.6f" % tuple(v) for v in _rows(t, True)]
    path.write_text("\n".join(lines) + "\n")


def _both(path):
    size, scale, table = N.parse_lut_file(path)
    osize, oscale, otable = D.READERS[D.dialect_of(str(path))](path.read_bytes())
    assert size == osize
    assert [np.float32(s) for s in scale] == [np.float32(s) for s in oscale]
    assert table.shape == otable.shape and np.array_equal(table.view(np.uint32), otable.view(np.uint32))
    return size, scale, table


@pytest.mark.parametrize("n,with_size", [(33, False), (33, True), (5, True), (17, True)])
def test_dat(tmp_path, n, with_size):
    t = _table(n, seed=n)
    p = tmp_path / "look.DAT"                      # the extension is matched case-insensitively
    write_dat(p, t, with_size=with_size)
    size, scale, table = _both(p)
    assert size == n and scale == [1.0, 1.0, 1.0]
    assert np.array_equal(table, np.float32([["%.7f" % x for x in v] for v in t.reshape(-1, 3)]).reshape(t.shape))   # blue fastest


def test_3dl(tmp_path):
    t12 = np.random.default_rng(3).integers(0, 4096, (17, 17, 17, 3))
    p = tmp_path / "look.3dl"
    write_3dl(p, t12)
    size, scale, table = _both(p)
    assert size == 17 and np.array_equal(table, (t12 / 4096.0).astype(np.float32))


@pytest.mark.parametrize("order", ["r g b", "b g r", "r b g"])   # (what Pandora writes, and the swaps; upstream maps column id -> channel)
def test_m3d(tmp_path, order):
    t = np.round(_table(6, seed=9) * 4095)
    p = tmp_path / "look.m3d"
    write_m3d(p, t, out=4096, order=order)
    size, scale, table = _both(p)
    assert size == 6 and np.array_equal(table, (t.astype(np.float32) * np.float32(1.0 / 4095)))


def test_csp_ranges(tmp_path):
    t = _table(9, seed=4)
    p = tmp_path / "look.csp"
    write_csp(p, t, in_rng=((0, 2), (0, 1), (-1, 3)), out_rng=((0, 1), (0.1, 0.9), (0, 2)))
    size, scale, table = _both(p)
    assert size == 9 and scale == [0.5, 1.0, 0.25]
    exp = np.float32([["%.6f" % x for x in v] for v in t.reshape(-1, 3)]).reshape(t.shape)
    gain = np.float32([1.0, np.float32(0.9) - np.float32(0.1), 2.0])
    assert np.array_equal(table, exp * gain)          # [r][g][b] although the file runs red-fastest


def test_csp_shaper_is_unsupported(tmp_path):
    p = tmp_path / "shaper.csp"
    write_csp(p, _table(3), prelut=True)
    with pytest.raises(N.LutrError) as e:
        N.parse_lut_file(p)
    assert e.value.status == N.ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        D.parse_csp(p.read_bytes())


def test_truncated_and_garbled_files(tmp_path):
    t = _table(4, seed=1)
    p = tmp_path / "short.dat"
    write_dat(p, t, comments=False)
    p.write_text("\n".join(p.read_text().splitlines()[:-3]) + "\n")
    with pytest.raises(N.LutrError) as e:
        N.parse_lut_file(p)
    assert e.value.status == N.ERR_INVALID_DATA and "EOF" in str(e.value)
    with pytest.raises(ValueError):
        D.parse_dat(p.read_bytes())
    q = tmp_path / "bad.3dl"
    write_3dl(q, np.zeros((17, 17, 17, 3), np.int64))
    q.write_text(q.read_text().replace("0 0 0", "0 x 0", 1))
    with pytest.raises(N.LutrError) as e:
        N.parse_lut_file(q)
    assert e.value.status == N.ERR_INVALID_DATA
    m = tmp_path / "noheader.m3d"
    m.write_text("values r g b\n0 0 0\n")
    with pytest.raises(N.LutrError) as e:
        N.parse_lut_file(m)
    assert e.value.status == N.ERR_INVALID_DATA
    d = tmp_path / "huge.dat"
    d.write_text("3DLUTSIZE 300\n0 0 0\n")
    with pytest.raises(N.LutrError) as e:
        N.parse_lut_file(d)
    assert e.value.status == N.ERR_LUT_SIZE


def test_extension_dispatch(tmp_path):
    for name in ("look", "look.txt", "look.cube.bak"):
        f = tmp_path / name
        f.write_text("LUT_3D_SIZE 2\n" + "0 0 0\n" * 8)
        with pytest.raises(N.LutrError) as e:
            N.parse_lut_file(f)
        assert e.value.status == N.ERR_INVALID_ARG
        with pytest.raises(ValueError):
            D.dialect_of(str(f))
    f = tmp_path / "ok.CuBe"
    f.write_text("LUT_3D_SIZE 2\n" + "0 0 0\n" * 7 + "1 1 1\n")
    size, scale, table = N.parse_lut_file(f)
    assert size == 2 and table[1, 1, 1].tolist() == [1.0, 1.0, 1.0]
    # text entry point with an explicit dialect
    size, _, table = N.parse_cube_text(b"3DLUTSIZE 2\n" + b"0.5 0.25 1\n" * 8, dialect="dat")
    assert size == 2 and table[0, 1, 0].tolist() == [0.5, 0.25, 1.0]
