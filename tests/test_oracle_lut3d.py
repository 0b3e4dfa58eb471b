"""Known-answer tests for the oracle's vf_lut3d restatement ("parity unpinned": libavfilter is absent, so these pin the
DEFINED semantics -- SURVEY.md App. C -- with hand-computed values)."""
import numpy as np
import pytest

from lut_renderer_b200 import synth
from oracle import oracle as O


def _px(lut, r, g, b, depth=10, interp="tetrahedral"):
    a = lambda v: np.full((1, 1), v, np.uint16)  # noqa: E731
    R, G, B = O.lut3d_planar16(lut, a(r), a(g), a(b), depth, interp)
    return int(R[0, 0]), int(G[0, 0]), int(B[0, 0])


def test_identity_lut_is_not_bit_identity_because_of_truncation():
    lut = O.OracleLut.from_table(synth.identity_lut(33).astype(np.float32))
    vals = np.arange(1024, dtype=np.uint16).reshape(1, -1)
    R, G, B = O.lut3d_planar16(lut, vals, vals, vals, 10, "tetrahedral")
    d = vals.astype(int) - R.astype(int)
    assert d.min() >= 0 and d.max() == 1 and 0 < (d != 0).mean() < 0.2   # truncating store: never above, sometimes one below
    assert (R == G).all() and (G == B).all()
    assert _px(lut, 0, 0, 0) == (0, 0, 0) and _px(lut, 1023, 1023, 1023) == (1023, 1023, 1023)


def test_lut_memory_order_is_r_g_b_and_file_order_is_red_fastest():
    n = 3
    t = np.zeros((n, n, n, 3), np.float32)
    t[2, 0, 0] = (1.0, 0.25, 0.5)     # node r=2, g=0, b=0
    lut = O.OracleLut.from_text(synth.cube_text(t))
    assert np.array_equal(lut.table, t)
    assert _px(lut, 1023, 0, 0, interp="nearest") == (1023, 255, 511)   # 0.25*1023 = 255.75 -> 255, 0.5*1023 = 511.5 -> 511
    assert _px(lut, 0, 1023, 0, interp="nearest") == (0, 0, 0)


def test_tetrahedral_hand_computed():
    # 2-point LUT = the unit cube corners; tetrahedral on f(r,g,b) = corner values
    t = np.zeros((2, 2, 2, 3), np.float32)
    for r in range(2):
        for g in range(2):
            for b in range(2):
                t[r, g, b] = (0.1 * r + 0.2 * g + 0.4 * b, float(r and g), float(r or b))
    lut = O.OracleLut.from_table(t)
    f32 = np.float32
    x = lambda v: f32(v) * f32(1.0 / 1023)   # noqa: E731
    dr, dg, db = x(700), x(300), x(500)      # dr > db > dg -> vertices c000, c100, c101, c111
    w = [f32(1) - dr, dr - db, db - dg, dg]
    exp = []
    for ch in range(3):
        acc = f32(w[0] * t[0, 0, 0, ch])
        acc = f32(acc + f32(w[1] * t[1, 0, 0, ch]))
        acc = f32(acc + f32(w[2] * t[1, 0, 1, ch]))
        acc = f32(acc + f32(w[3] * t[1, 1, 1, ch]))
        exp.append(int(min(max(int(f32(acc * f32(1023.0))), 0), 1023)))
    assert _px(lut, 700, 300, 500) == tuple(exp)


def test_trilinear_vs_tetrahedral_agree_on_axes_and_differ_off_axis():
    lut = O.OracleLut.from_table(synth.look_lut(17).astype(np.float32))
    assert _px(lut, 512, 0, 0, interp="trilinear") == _px(lut, 512, 0, 0, interp="tetrahedral")
    rng = np.random.default_rng(0)
    R, G, B = (rng.integers(0, 1024, (8, 64), dtype=np.uint16) for _ in range(3))
    a = O.lut3d_planar16(lut, R, G, B, 10, "trilinear")
    b = O.lut3d_planar16(lut, R, G, B, 10, "tetrahedral")
    assert any((x != y).any() for x, y in zip(a, b))
    assert max(int(np.abs(x.astype(int) - y.astype(int)).max()) for x, y in zip(a, b)) < 40


def test_out_of_range_lut_values_clip_and_domain_scale():
    t = np.full((2, 2, 2, 3), -0.5, np.float32)
    t[1, 1, 1] = 2.0
    lut = O.OracleLut.from_table(t)
    assert _px(lut, 0, 0, 0) == (0, 0, 0) and _px(lut, 1023, 1023, 1023) == (1023, 1023, 1023)
    # DOMAIN_MAX 2 -> scale 0.5: full-scale input only reaches the middle of the table
    ident = synth.identity_lut(3).astype(np.float32)
    half = O.OracleLut.from_text(synth.cube_text(ident, domain=((0, 0, 0), (2, 2, 2))))
    assert list(half.scale) == [0.5, 0.5, 0.5]
    assert _px(half, 1023, 1023, 1023)[0] in (510, 511)
    # DOMAIN_MAX 0.5 -> 1/(max-min) = 2 is clipped to 1
    assert list(O.OracleLut.from_text(synth.cube_text(ident, domain=((0, 0, 0), (.5, .5, .5)))).scale) == [1.0, 1.0, 1.0]


def test_8bit_packed_path():
    lut = O.OracleLut.from_table(synth.identity_lut(33).astype(np.float32))
    rgb = np.arange(256, dtype=np.uint8).repeat(3).reshape(1, -1)
    out = O.lut3d_rgb24(lut, rgb, "tetrahedral")
    d = rgb.astype(int) - out.astype(int)
    assert d.min() >= 0 and d.max() <= 1


@pytest.mark.parametrize("interp", ["pyramid", "prism"])
def test_pyramid_prism_reproduce_corners_and_affine_luts(interp):
    """interp=pyramid / prism (reference whitelist ffmpeg.py:243): both reproduce the LUT nodes, are exact on affine LUTs (every
    difference term cancels) and agree with tetrahedral along the cube axes / grey diagonal of a 2-point LUT."""
    rng = np.random.default_rng(5)
    t = rng.random((2, 2, 2, 3)).astype(np.float32)
    lut = O.OracleLut.from_table(t)
    for r in (0, 1):
        for g in (0, 1):
            for b in (0, 1):
                exp = tuple(int(np.float32(v) * np.float32(1023.0)) for v in t[r, g, b])
                assert _px(lut, 1023 * r, 1023 * g, 1023 * b, interp=interp) == exp
    # affine LUT: value = 0.1 + 0.5 r + 0.2 g + 0.15 b on a 9-point grid -> any cell interpolation returns the affine function
    n = 9
    ax = np.linspace(0, 1, n, dtype=np.float64)
    aff = (0.1 + 0.5 * ax[:, None, None] + 0.2 * ax[None, :, None] + 0.15 * ax[None, None, :])
    lut = O.OracleLut.from_table(np.repeat(aff[..., None], 3, -1).astype(np.float32))
    pts = rng.integers(0, 1024, (1, 4096, 3)).astype(np.uint16)
    R, G, B = O.lut3d_planar16(lut, pts[..., 0].copy(), pts[..., 1].copy(), pts[..., 2].copy(), 10, interp)
    ideal = (0.1 + (0.5 * pts[..., 0] + 0.2 * pts[..., 1] + 0.15 * pts[..., 2]) / 1023.0) * 1023.0
    assert np.abs(R.astype(np.float64) - np.floor(ideal)).max() <= 1 and np.array_equal(R, G) and np.array_equal(G, B)
    # on a cube edge (two fractions zero) every interpolation is the same 1-D lerp up to float rounding of the blend order
    lut = O.OracleLut.from_table(synth.look_lut(17).astype(np.float32))
    for px in ((512, 0, 0), (0, 300, 0), (0, 0, 900)):
        a, b = _px(lut, *px, interp=interp), _px(lut, *px, interp="trilinear")
        assert max(abs(x - y) for x, y in zip(a, b)) <= 1


def test_trilinear_agrees_with_independent_implementations():
    """The lut3d leg has no FFmpeg binary to pin against.  Two unrelated implementations at least confirm its
    geometry -- which file column is which axis, node spacing, cell lookup, blend -- to within the rounding each of them uses:
    scipy's RegularGridInterpolator (float64 multilinear) and Pillow's Color3DLUT filter (C, fixed point, 8 bit)."""
    from scipy.interpolate import RegularGridInterpolator
    from PIL import Image, ImageFilter
    n = 17
    t = synth.random_lut(n, 11).astype(np.float32)          # [r][g][b][3]
    lut = O.OracleLut.from_text(synth.cube_text(t))           # through the .cube text: red fastest in the file
    rng = np.random.default_rng(3)
    h, w = 64, 96
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = O.lut3d_rgb24(lut, rgb.reshape(h, 3 * w).copy(), "trilinear").reshape(h, w, 3).astype(np.int32)
    # scipy: float64 multilinear on the same grid; lut3d truncates, so floor() of the ideal value, +-1 for float32 rounding
    ax = np.linspace(0.0, 1.0, n)
    ideal = np.stack([RegularGridInterpolator((ax, ax, ax), t[..., c].astype(np.float64))(rgb.reshape(-1, 3) / 255.0) for c in range(3)], -1)
    ref = np.floor(np.clip(ideal, 0.0, 1.0) * 255.0).reshape(h, w, 3)
    assert np.abs(got - ref).max() <= 1 and (got == ref).mean() > 0.98
    # Pillow: table flattened with the red index fastest (the .cube file order), 8-bit fixed-point blend, rounding to nearest
    flat = np.ascontiguousarray(t.transpose(2, 1, 0, 3)).reshape(-1).tolist()
    pil = np.asarray(Image.fromarray(rgb, "RGB").filter(ImageFilter.Color3DLUT(n, flat))).astype(np.int32)
    assert np.abs(got - pil).max() <= 2
    # a wrong axis order would be off by tens of codes on a random table:
    swapped = np.asarray(Image.fromarray(rgb, "RGB").filter(ImageFilter.Color3DLUT(n, t.reshape(-1).tolist()))).astype(np.int32)
    assert np.abs(got - swapped).max() > 20


def test_tetrahedral_against_the_sorted_coordinate_definition():
    """Tetrahedral interpolation from its mathematical definition (sort the three fractions, walk from c000 to c111 adding one axis
    at a time, weights = successive differences), in float64 numpy: the oracle's six-branch float32 code must agree to the last
    code value except where float32 rounding meets the truncating store (+-1)."""
    n = 33
    t = synth.random_lut(n, 12).astype(np.float32)
    lut = O.OracleLut.from_table(t)
    rng = np.random.default_rng(4)
    pts = rng.integers(0, 1024, (20000, 3))
    R, G, B = (np.ascontiguousarray(pts[:, c].reshape(1, -1).astype(np.uint16)) for c in range(3))
    got = np.stack(O.lut3d_planar16(lut, R, G, B, 10, "tetrahedral"), -1).reshape(-1, 3).astype(np.int64)
    s = pts / 1023.0 * (n - 1)
    p = np.minimum(np.floor(s).astype(int), n - 1)
    d = s - p
    order = np.argsort(-d, axis=1, kind="stable")          # axis with the largest fraction first
    ds = np.take_along_axis(d, order, 1)
    w = np.stack([1 - ds[:, 0], ds[:, 0] - ds[:, 1], ds[:, 1] - ds[:, 2], ds[:, 2]], 1)
    idx = p.copy()
    acc = w[:, :1] * t[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float64)
    for k in range(3):
        step = np.zeros_like(idx)
        step[np.arange(len(idx)), order[:, k]] = 1
        idx = np.minimum(idx + step, n - 1)
        acc += w[:, k + 1:k + 2] * t[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float64)
    ref = np.floor(np.clip(acc, 0, 1) * 1023.0).astype(np.int64)
    assert np.abs(got - ref).max() <= 1 and (got == ref).mean() > 0.99


def test_composite_outputs_match_the_committed_hashes():
    """The oracle composite (swscale legs + restated lut3d) reproduces the committed SHA-256 digests for 80 seeded cases
    (tests/golden/oracle_hashes.json, written by tests/golden/gen_oracle_hashes.py): edits to oracle/, another compiler or libm
    cannot silently move the yardstick the GPU is held to."""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("gen_oracle_hashes", os.path.join(here, "gen_oracle_hashes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(here, "oracle_hashes.json")) as f:
        want = json.load(f)
    got = mod.compute()
    assert set(got) == set(want) and len(want) == 80
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, bad[:5]
