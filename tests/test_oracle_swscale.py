"""Pins the oracle's integer models of the swscale legs against the REAL libswscale 9.1.100 (FFmpeg 8.0.1) that
ships in this image -- the library the reference's `scale` / `format` filters run (ffmpeg.py:225-235, 309-310).
Bit-exact on random full-code-range frames, known answers from SURVEY.md App. B.6."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import swscale_real as S

pytestmark = pytest.mark.skipif(not S.available(), reason="bundled libswscale not loadable")


def test_versions_and_coefficient_tables():
    v = S.versions()
    assert v["libswscale"].startswith("9."), v       # a different FFmpeg would need re-pinning
    for name, mid in (("bt709", 0), ("bt601", 1), ("bt2020", 2), ("fcc", 3), ("smpte240m", 4)):
        assert S.coefficients(name) == [int(O.lib().orc_inv_table(mid)[i]) for i in range(4)]
    assert O.rgb2yuv_coeffs("bt709") == [5983, 20127, 2032, -3298, -11094, 14392, 14392, -13073, -1320]
    assert O.rgb2yuv_coeffs("bt2020") == [7393, 19080, 1669, -4019, -10373, 14392, 14392, -13235, -1158]
    assert O.rgb2yuv_coeffs("bt601") == [8414, 16519, 3208, -4865, -9528, 14392, 14392, -12061, -2332]
    t = O.yuv2rgb_coeffs("bt709", False)
    assert (t["yC"], t["vr"], t["ub"], t["ug"], t["vg"]) == (9539, 14686, 17305, -1747, -4366)


@pytest.mark.parametrize("w,h", [(1920, 1080), (64, 16), (34, 12), (1000, 18)])
@pytest.mark.parametrize("matrix", ["bt709", "bt601", "bt2020"])
def test_yuv420p_legs_match_libswscale(w, h, matrix):
    rng = np.random.default_rng(w * 31 + h)
    for full in (0, 1):
        Y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        U = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        V = rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        ref = S.SwsConverter(w, h, "yuv420p", "rgb24", src_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
            [Y, U, V], S.alloc_planes("rgb24", w, h))[0]
        assert np.array_equal(O.yuv420p_to_rgb24(Y, U, V, matrix, full), ref)
    rgb = rng.integers(0, 256, (h, 3 * w), dtype=np.uint8)
    ref = S.SwsConverter(w, h, "rgb24", "yuv420p", src_matrix=matrix, dst_matrix=matrix).convert([rgb], S.alloc_planes("yuv420p", w, h))
    for a, b in zip(O.rgb24_to_yuv420p(rgb, matrix), ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(3840, 16), (1920, 8), (16, 2), (22, 3), (7680, 2)])
@pytest.mark.parametrize("matrix", ["bt709", "bt601", "bt2020"])
def test_yuv422p10_legs_match_libswscale(w, h, matrix):
    rng = np.random.default_rng(w * 17 + h)
    for full in (0, 1):
        Y = rng.integers(0, 1024, (h, w), dtype=np.uint16)
        U = rng.integers(0, 1024, (h, w // 2), dtype=np.uint16)
        V = rng.integers(0, 1024, (h, w // 2), dtype=np.uint16)
        g, b, r = S.SwsConverter(w, h, "yuv422p10le", "gbrp10le", src_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
            [Y, U, V], S.alloc_planes("gbrp10le", w, h))
        R, G, B = O.yuv422p10_to_gbrp10(Y, U, V, matrix, full)
        assert np.array_equal(R, r) and np.array_equal(G, g) and np.array_equal(B, b)
    R, G, B = (rng.integers(0, 1024, (h, w), dtype=np.uint16) for _ in range(3))
    ref = S.SwsConverter(w, h, "gbrp10le", "yuv422p10le", src_matrix=matrix, dst_matrix=matrix).convert(
        [G, B, R], S.alloc_planes("yuv422p10le", w, h))
    for a, b in zip(O.gbrp10_to_yuv422p10(R, G, B, matrix), ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 32), (64, 16), (34, 12), (1000, 18), (16, 2)])
@pytest.mark.parametrize("matrix", ["bt709", "bt601", "bt2020"])
def test_rgb24_to_yuv422p10_matches_libswscale(w, h, matrix):
    """Return leg of the pro-mode master for 8-bit sources (rgb24 -> yuv422p10le): bit-exact incl. saturated colours."""
    rng = np.random.default_rng(w * 13 + h)
    rgb = rng.integers(0, 256, (h, 3 * w), dtype=np.uint8)
    sel = rng.random(rgb.shape) < 0.3
    rgb[sel] = np.where(rng.random(int(sel.sum())) < 0.5, 0, 255).astype(np.uint8)
    ref = S.SwsConverter(w, h, "rgb24", "yuv422p10le", src_matrix=matrix, dst_matrix=matrix).convert(
        [rgb], S.alloc_planes("yuv422p10le", w, h))
    for a, b in zip(O.rgb24_to_yuv422p10(rgb, matrix), ref):
        assert np.array_equal(a, b)
    grey = np.full((h, 3 * w), 128, np.uint8)          # grey 128 -> (504, 512, 512)
    Y, U, V = O.rgb24_to_yuv422p10(grey)
    assert (int(Y[0, 0]), int(U[0, 0]), int(V[0, 0])) == (504, 512, 512)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 24), (64, 32), (34, 14), (1000, 18), (16, 16), (22, 30), (48, 20)])
@pytest.mark.parametrize("matrix", ["bt709", "bt601", "bt2020"])
def test_yuv420p10_to_gbrp10_matches_libswscale(w, h, matrix):
    """Forward leg of the master for 10-bit 4:2:0 sources: separable 4-tap horizontal x 4-tap vertical chroma upsampling with
    the edge-normalised tap sets on the two top rows / left columns; full code range, both source ranges.  Valid domain found by
    probing the binary: even heights >= 14 (shorter frames make swscale shrink its vertical filters differently)."""
    rng = np.random.default_rng(w * 7 + h)
    for full in (0, 1):
        Y = rng.integers(0, 1024, (h, w), dtype=np.uint16)
        U = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.uint16)
        V = rng.integers(0, 1024, (h // 2, w // 2), dtype=np.uint16)
        g, b, r = S.SwsConverter(w, h, "yuv420p10le", "gbrp10le", src_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
            [Y, U, V], S.alloc_planes("gbrp10le", w, h))
        R, G, B = O.yuv420p10_to_gbrp10(Y, U, V, matrix, full)
        assert np.array_equal(R, r) and np.array_equal(G, g) and np.array_equal(B, b)


@pytest.mark.parametrize("w,h", [(1920, 1080), (3840, 24), (64, 32), (34, 14), (1000, 18), (16, 16), (22, 30), (48, 20)])
@pytest.mark.parametrize("matrix", ["bt709", "bt601", "bt2020"])
def test_gbrp10_to_yuv420p10_matches_libswscale(w, h, matrix):
    """Return leg for 10-bit 4:2:0 outputs (libx265 preserve, ffmpeg.py:297-298): 8-tap horizontal then 8-tap vertical chroma
    decimation, exact accumulation; random full-code-range RGB with saturated patches (hits the 32767 clip of the 15-bit rows)."""
    rng = np.random.default_rng(w * 11 + h)
    R, G, B = (rng.integers(0, 1024, (h, w), dtype=np.uint16) for _ in range(3))
    for P in (R, G, B):
        sel = rng.random(P.shape) < 0.3
        P[sel] = np.where(rng.random(int(sel.sum())) < 0.5, 0, 1023).astype(np.uint16)
    R[: h // 2, : w // 2], G[: h // 2, : w // 2], B[: h // 2, : w // 2] = 0, 0, 1023     # flat saturated blue block: max U
    ref = S.SwsConverter(w, h, "gbrp10le", "yuv420p10le", src_matrix=matrix, dst_matrix=matrix).convert(
        [G, B, R], S.alloc_planes("yuv420p10le", w, h))
    for a, b in zip(O.gbrp10_to_yuv420p10(R, G, B, matrix), ref):
        assert np.array_equal(a, b)


def test_range_pass_tables_match_libswscale():
    tl, tc = O.range_tables()
    w, h = 256, 12
    for fmt in ("yuv420p", "yuv422p", "yuv444p"):
        conv = S.SwsConverter(w, h, fmt, fmt, src_range=1, dst_range=0)
        src = S.alloc_planes(fmt, w, h)
        for p in src:
            p[:] = np.resize(np.arange(256, dtype=np.uint8), p.shape[1])[None, :]
        dst = conv.convert(src, S.alloc_planes(fmt, w, h))
        assert np.array_equal(dst[0], tl[src[0]])
        assert np.array_equal(dst[1], tc[src[1]]) and np.array_equal(dst[2], tc[src[2]])
    assert (tl[0], tl[128], tl[255], tc[0], tc[128], tc[254], tc[255]) == (16, 126, 235, 16, 128, 239, 240)


def test_known_answers_b6():
    w, h = 64, 12
    conv = S.SwsConverter(w, h, "yuv422p10le", "gbrp10le")
    for y, exp in ((64, 0), (65, 1), (512, 522), (940, 1020), (941, 1021), (1023, 1023)):
        planes = [np.full((h, w), y, np.uint16), np.full((h, w // 2), 512, np.uint16), np.full((h, w // 2), 512, np.uint16)]
        g, b, r = conv.convert(planes, S.alloc_planes("gbrp10le", w, h))
        R, G, B = O.yuv422p10_to_gbrp10(*planes)
        assert int(r[3, 9]) == int(R[3, 9]) == exp
    for v, exp in ((0, (64, 512, 512)), (512, (504, 512, 512)), (1023, (943, 512, 512))):
        R = np.full((h, w), v, np.uint16)
        Y, U, V = O.gbrp10_to_yuv422p10(R, R, R)
        assert (int(Y[2, 5]), int(U[2, 5]), int(V[2, 5])) == exp
    rgb = np.full((h, 3 * w), 128, np.uint8)
    Y, U, V = O.rgb24_to_yuv420p(rgb)
    assert (int(Y[4, 4]), int(U[2, 2]), int(V[2, 2])) == (126, 128, 128)
    rgb = np.tile(np.array([200, 30, 77], np.uint8), (h, w)).reshape(h, 3 * w)
    Y, U, V = O.rgb24_to_yuv420p(rgb)
    assert int(U[2, 3]) == 131 and int(U[0, 3]) == 132 and int(U[h // 2 - 1, 3]) == 132   # edge-folded taps / C scaler


def test_composite_with_real_swscale_legs_equals_oracle_composite():
    """The full chain with REAL libswscale around the lut3d restatement == the all-oracle chain (what the GPU is held to)."""
    from lut_renderer_b200 import synth
    from oracle import cpu_baseline as C
    for fmt, out_fmt, w, h, n in (("yuv422p10le", None, 640, 48, 33), ("yuv420p", None, 320, 36, 17), ("yuv420p", "yuv422p10le", 320, 36, 33),
                                  ("yuv420p10le", "yuv422p10le", 320, 36, 33), ("yuv420p10le", "yuv420p10le", 320, 36, 33),
                                  ("yuv422p10le", "yuv420p10le", 328, 30, 17)):
        lut = O.OracleLut.from_table(synth.look_lut(n).astype(np.float32))
        for frame in (synth.natural_frame(fmt, w, h, 2), synth.noise_frame(fmt, w, h, 4)):
            planes = [np.ascontiguousarray(p) for p in synth.split_planes(frame, fmt, w, h)]
            a = C.CpuStage(fmt, w, h, lut, use_swscale=True, out_pix_fmt=out_fmt).run(planes)
            b = C.CpuStage(fmt, w, h, lut, use_swscale=False, out_pix_fmt=out_fmt).run(planes)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


# ---- the whole format matrix: generic legs (orc_yuv8_to_rgb24, orc_yuv444p10_to_gbrp10, orc_rgb_to_yuv_any) ---------------------
def _saturate(rng, planes, hi):
    for P in planes:
        sel = rng.random(P.shape) < 0.25
        P[sel] = np.where(rng.random(int(sel.sum())) < 0.5, 0, hi).astype(P.dtype)


@pytest.mark.parametrize("fmt", ["yuv420p", "yuv422p", "yuv444p"])
@pytest.mark.parametrize("w,h", [(1920, 36), (64, 16), (328, 18), (16, 12)])
def test_any_8bit_sources_match_libswscale(fmt, w, h):
    """yuv4xxp -> rgb24 for every 8-bit layout the reference's range pass can leave (ffmpeg.py:137-143): 4:2:0 / 4:2:2 take the
    SIMD path (replicated chroma), 4:4:4 swscale's C path; 3 matrices x 2 ranges, full code range."""
    rng = np.random.default_rng(w + h)
    shapes, dt = O.pix_geometry(fmt, w, h)
    for matrix in ("bt709", "bt601", "bt2020"):
        for full in (0, 1):
            Y, U, V = (rng.integers(0, 256, s, dtype=np.uint8) for s in shapes)
            ref = S.SwsConverter(w, h, fmt, "rgb24", src_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
                [Y, U, V], S.alloc_planes("rgb24", w, h))[0]
            assert np.array_equal(O.yuv8_to_rgb24(fmt, Y, U, V, matrix, full), ref)


@pytest.mark.parametrize("w,h", [(512, 32), (66, 14), (16, 2)])
def test_any_yuv444p10_source_matches_libswscale(w, h):
    rng = np.random.default_rng(w * 3 + h)
    for matrix in ("bt709", "bt601", "bt2020"):
        for full in (0, 1):
            Y, U, V = (rng.integers(0, 1024, (h, w), dtype=np.uint16) for _ in range(3))
            g, b, r = S.SwsConverter(w, h, "yuv444p10le", "gbrp10le", src_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
                [Y, U, V], S.alloc_planes("gbrp10le", w, h))
            R, G, B = O.yuv444p10_to_gbrp10(Y, U, V, matrix, full)
            assert np.array_equal(R, r) and np.array_equal(G, g) and np.array_equal(B, b)


@pytest.mark.parametrize("src_depth", [8, 10])
@pytest.mark.parametrize("out_fmt", ["yuv420p", "yuv422p", "yuv444p", "yuv420p10le", "yuv422p10le", "yuv444p10le"])
@pytest.mark.parametrize("w,h", [(1920, 36), (64, 16), (328, 18), (22, 30)])
def test_any_output_leg_matches_libswscale(src_depth, out_fmt, w, h):
    """RGB -> every output layout / depth / range: rgb24 (8-bit sources) and gbrp10le (10-bit sources) in front; limited and full
    range behind (swscale's 15-bit range conversion), ordered dither for the 10 -> 8 bit outputs, the 4:2:0 vertical scalers
    (SIMD per-tap truncation for 8 bit incl. the C-path last row, exact for 10 bit)."""
    rng = np.random.default_rng(w * 5 + h + src_depth)
    hi = (1 << src_depth) - 1
    src_fmt = "rgb24" if src_depth == 8 else "gbrp10le"
    for matrix in ("bt709", "bt601", "bt2020"):
        for full in (0, 1):
            R, G, B = (rng.integers(0, hi + 1, (h, w), dtype=np.uint16) for _ in range(3))
            _saturate(rng, (R, G, B), hi)
            if src_depth == 8:
                src = [np.stack([R, G, B], -1).astype(np.uint8).reshape(h, 3 * w)]
            else:
                src = [G, B, R]
            ref = S.SwsConverter(w, h, src_fmt, out_fmt, dst_range=full, src_matrix=matrix, dst_matrix=matrix).convert(
                src, S.alloc_planes(out_fmt, w, h))
            got = O.rgb_to_yuv_any(src_depth, out_fmt, R, G, B, matrix, full)
            for name, a, b in zip("YUV", got, ref):
                assert np.array_equal(a, b), (name, matrix, full, int((a != b).sum()))


def test_any_composite_equals_specialised_composites():
    """orc_stage_any == the specialised composites on the formats both cover, and the range constants reproduce the pc->tv tables."""
    from lut_renderer_b200 import synth
    lut = O.OracleLut.from_table(synth.look_lut(17).astype(np.float32))
    w, h = 96, 20
    for fmt, out_fmt in (("yuv420p", "yuv420p"), ("yuv420p", "yuv422p10le"), ("yuv422p10le", "yuv422p10le"),
                         ("yuv420p10le", "yuv422p10le"), ("yuv422p10le", "yuv420p10le"), ("yuv420p10le", "yuv420p10le")):
        f = synth.noise_frame(fmt, w, h, 3)
        Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(f, fmt, w, h)]
        a = O.stage_any(lut, fmt, out_fmt, Y, U, V, range_pass=(fmt == "yuv420p"))
        if fmt == "yuv420p" and out_fmt == "yuv420p":
            b = O.stage_yuv420p(lut, Y, U, V, range_pass=True)
        elif fmt == "yuv420p":
            b = O.stage_yuv420p_to_yuv422p10(lut, Y, U, V, range_pass=True)
        else:
            b = O.stage_10bit(lut, Y, U, V, in_420=fmt == "yuv420p10le", out_420=out_fmt == "yuv420p10le")
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    tl, tc = O.range_tables()
    x = np.arange(256, dtype=np.int64) << 7
    for chroma, t in ((0, tl), (1, tc)):
        c, o = O.range_coeffs(8, False, chroma)
        assert np.array_equal(np.clip((np.minimum((x * c + o) >> 14, 32767) + 64) >> 7, 0, 255), t)


@pytest.mark.parametrize("fmt", ["yuv420p10le", "yuv422p10le", "yuv444p10le"])
@pytest.mark.parametrize("w,h", [(256, 32), (66, 18), (1920, 36)])
def test_any_10bit_full_range_prepass(fmt, w, h):
    """A full-range source deeper than 8 bit: the reference's `scale=in_range=pc:out_range=R,format=yuv4xxp` is ONE swscale pass
    from yuv4xxp10le(pc) to 8-bit yuv4xxp (ffmpeg.py:211-233).  Both forms (R = tv: range conversion + ordered dither; R = pc: planar
    copy with the 2x2 dither and the full-range luma correction) bit-exact, plus the exhaustive luma ramp for the pc form."""
    rng = np.random.default_rng(w + 3 * h)
    shapes, _ = O.pix_geometry(fmt, w, h)
    P = [rng.integers(0, 1024, s, dtype=np.uint16) for s in shapes]
    dst_fmt = fmt.replace("10le", "")
    for mode, dst_range in ((1, 0), (2, 1)):
        ref = S.SwsConverter(w, h, fmt, dst_fmt, src_range=1, dst_range=dst_range).convert(P, S.alloc_planes(dst_fmt, w, h))
        for plane in range(3):
            assert np.array_equal(O.prepass_10_to_8(mode, plane, P[plane]), ref[plane]), (mode, plane)
    if fmt == "yuv444p10le" and w == 256:
        ramp = np.tile((np.arange(8192) >> 3).astype(np.uint16)[None, :], (16, 1))      # every code value at every dither phase
        mid = np.full((16, 8192), 512, np.uint16)
        ref = S.SwsConverter(8192, 16, fmt, dst_fmt, src_range=1, dst_range=1).convert([ramp, mid, mid], S.alloc_planes(dst_fmt, 8192, 16))
        assert np.array_equal(O.prepass_10_to_8(2, 0, ramp), ref[0])
        ref = S.SwsConverter(8192, 16, fmt, dst_fmt, src_range=1, dst_range=0).convert([ramp, ramp, ramp], S.alloc_planes(dst_fmt, 8192, 16))
        for plane in range(3):
            assert np.array_equal(O.prepass_10_to_8(1, plane, ramp), ref[plane])


@pytest.mark.parametrize("out_fmt", ["yuv420p", "yuv420p10le", "yuv422p"])
@pytest.mark.parametrize("w,h", [(256, 32), (72, 16), (1920, 36), (3840, 18)])
def test_convert_422p10_without_lut_matches_swscale(out_fmt, w, h):
    """Pro mode stage 2 (reference ffmpeg.py:459-471): no LUT, the encoding ffmpeg converts the yuv422p10le master to the delivery
    pix_fmt with an auto-inserted scaler (YUV -> YUV, bicubic flags, same matrix and range).  orc_convert_yuv422p10 is bit-exact
    against the real libswscale on full-code-range noise, on extreme values and on a flat frame (edge rows, last chroma row)."""
    rng = np.random.default_rng(7 * w + h)
    shapes, _ = O.pix_geometry("yuv422p10le", w, h)
    cases = [[rng.integers(0, 1024, s, dtype=np.uint16) for s in shapes],
             [rng.choice(np.array([0, 1, 511, 512, 1022, 1023], np.uint16), s) for s in shapes],
             [np.full(s, v, np.uint16) for s, v in zip(shapes, (940, 64, 960))]]
    conv = S.SwsConverter(w, h, "yuv422p10le", out_fmt)
    for k, P in enumerate(cases):
        ref = conv.convert(P, S.alloc_planes(out_fmt, w, h))
        got = O.convert_yuv422p10(out_fmt, *P)
        for plane in range(3):
            assert np.array_equal(got[plane], ref[plane]), (out_fmt, w, h, k, plane)
