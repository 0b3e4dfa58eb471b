"""GPU parity: the CUDA path (through the C-ABI) against the oracle on the same seeded inputs.

Bar: BIT-EXACT on every plane (the oracle restates the reference's integer legs exactly and its float32 LUT leg
with the same operation order; north_star's tolerance -- <= 1 code value, >= 99.9 % identical -- is therefore met
with margin and asserted as 0 differences here)."""
import numpy as np
import pytest

from lut_renderer_b200 import synth

from helpers import assert_bit_exact, gpu_stage, oracle_lut, oracle_stage

pytestmark = pytest.mark.gpu

LUTS = {
    "look33": lambda: synth.look_lut(33),
    "look65": lambda: synth.look_lut(65),
    "rand17": lambda: synth.random_lut(17, 5),
    "rand65": lambda: synth.random_lut(65, 2),
    "ident33": lambda: synth.identity_lut(33),
    "rand2": lambda: synth.random_lut(2, 9),
    "rand64": lambda: synth.random_lut(64, 3),   # size-1 not a power of two
}


def _frames(pix_fmt, w, h, kinds):
    out = []
    for k in kinds:
        if k == "natural":
            out.append(synth.natural_frame(pix_fmt, w, h, t=3))
        elif k == "noise":
            out.append(synth.noise_frame(pix_fmt, w, h, seed=11))
        elif k == "ramp":
            out.append(synth.ramp_frame(pix_fmt, w, h))
        elif k == "extremes":
            f = synth.noise_frame(pix_fmt, w, h, seed=12)
            hi = 1023 if "10" in pix_fmt else 255
            v = f.view(np.dtype("<u2") if hi > 255 else np.uint8)
            sel = np.random.default_rng(13).random(v.shape) < 0.5
            v[sel] = np.where(np.random.default_rng(14).random(int(sel.sum())) < 0.5, 0, hi).astype(v.dtype)
            out.append(f)
    return out


@pytest.mark.parametrize("lut_name", ["look65", "rand65", "ident33", "rand17", "rand2", "rand64"])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear", "nearest"])
def test_yuv422p10_interp_luts(lut_name, interp):
    w, h = 256, 24
    table = LUTS[lut_name]().astype(np.float32)
    frames = _frames("yuv422p10le", w, h, ["natural", "noise", "ramp", "extremes"])
    got = gpu_stage("yuv422p10le", w, h, frames, table, interp=interp)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv422p10le", w, h, f, L, interp=interp)
        assert_bit_exact("yuv422p10le", w, h, got[i], exp, f"422p10 {lut_name} {interp} frame {i}")


@pytest.mark.parametrize("w,h", [(16, 2), (24, 1), (64, 3), (1920, 4), (3840, 2), (7680, 2), (4104, 1), (16384, 2)])
def test_yuv422p10_geometries(w, h):
    table = LUTS["look33"]().astype(np.float32)
    frames = _frames("yuv422p10le", w, h, ["noise", "natural"])
    got = gpu_stage("yuv422p10le", w, h, frames, table)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv422p10le", w, h, got[i], oracle_stage("yuv422p10le", w, h, f, L), f"422p10 {w}x{h} frame {i}")


@pytest.mark.parametrize("in_matrix,out_matrix,norm_range", [("bt601", "bt709", "tv"), ("bt2020", "bt2020", "tv"),
                                                             ("bt709", "bt601", "tv"), ("bt709", "bt709", "pc")])
def test_yuv422p10_matrices_ranges(in_matrix, out_matrix, norm_range):
    w, h = 128, 8
    table = LUTS["rand17"]().astype(np.float32)
    frames = _frames("yuv422p10le", w, h, ["noise", "extremes"])
    kw = dict(in_matrix=in_matrix, out_matrix=out_matrix, in_range=norm_range, norm_range=norm_range)
    got = gpu_stage("yuv422p10le", w, h, frames, table, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv422p10le", w, h, got[i], oracle_stage("yuv422p10le", w, h, f, L, **kw), f"422p10 {kw} frame {i}")


def test_yuv422p10_rows_per_cta_and_domain_scale():
    w, h = 512, 37
    table = LUTS["rand17"]().astype(np.float32)
    scale = (0.5, 1.0, 0.25)   # DOMAIN_MAX = 2, 1, 4
    frames = _frames("yuv422p10le", w, h, ["noise", "natural", "ramp"])
    L = oracle_lut(table, scale)
    exp = [oracle_stage("yuv422p10le", w, h, f, L) for f in frames]
    for rows in (1, 2, 5, 64):
        got = gpu_stage("yuv422p10le", w, h, frames, table, scale, options={"rows_per_cta": rows})
        for i in range(len(frames)):
            assert_bit_exact("yuv422p10le", w, h, got[i], exp[i], f"422p10 rows_per_cta={rows} frame {i}")


@pytest.mark.parametrize("options", [dict(fast_path=0), dict(), dict(grid_ctas=1), dict(grid_ctas=7), dict(grid_ctas=3, rows_per_cta=8),
                                     dict(rows_per_cta=64), dict(rows_per_cta=128, grid_ctas=2), dict(float_mode="fma"),
                                     # the row-staging kernel (k422_tma.cu): 1 = per-lane cp.async ring, 2 = bulk async copies (TMA + mbarrier)
                                     dict(wk_tma=1), dict(wk_tma=2), dict(wk_tma=1, grid_ctas=1), dict(wk_tma=2, grid_ctas=1),
                                     dict(wk_tma=1, rows_per_cta=8, grid_ctas=5), dict(wk_tma=2, rows_per_cta=8, grid_ctas=5),
                                     dict(wk_tma=2, rows_per_cta=128), dict(wk_tma=1, float_mode="fma"), dict(wk_tma=2, float_mode="fma")])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear"])
def test_yuv422p10_kernel_variants(options, interp):
    """Every code path of the 4:2:2 stage (generic tile kernel, warp-window kernel with register prefetch, row-staging kernel with
    either copy engine, tiny grids whose CTAs walk many units, short and long units) gives the same bits; LUTs: values inside [0,1]
    (no store clip), outside (clip), size-1 not a power of two."""
    w, h = 656, 21
    options = dict(options)
    float_mode = options.pop("float_mode", "c")
    frames = _frames("yuv422p10le", w, h, ["natural", "noise", "extremes"])
    for lut_name, table in (("look33", LUTS["look33"]()), ("look65", LUTS["look65"]()), ("rand64", LUTS["rand64"]()),
                            ("wide9", synth.random_lut(9, 4) * 1.6 - 0.3)):
        table = table.astype(np.float32)
        got = gpu_stage("yuv422p10le", w, h, frames, table, interp=interp, options=options, float_mode=float_mode)
        if float_mode == "fma":  # opt-in contraction: same bits from every kernel that implements it
            ref = gpu_stage("yuv422p10le", w, h, frames, table, interp=interp, options=dict(fast_path=0), float_mode=float_mode)
            for i in range(len(frames)):
                assert np.array_equal(got[i], ref[i]), f"422p10 fma {lut_name} {interp} frame {i}"
            continue
        L = oracle_lut(table)
        for i, f in enumerate(frames):
            assert_bit_exact("yuv422p10le", w, h, got[i], oracle_stage("yuv422p10le", w, h, f, L, interp=interp),
                             f"422p10 {options} {lut_name} {interp} frame {i}")


@pytest.mark.parametrize("w", [16, 232, 240, 248, 256, 472, 480, 488, 496, 728, 1928])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear"])
@pytest.mark.parametrize("extra", [dict(), dict(wk_tma=1), dict(wk_tma=2)])
def test_yuv422p10_window_edges(w, interp, extra):
    """Row widths around the 240-column windows of the warp-window kernel and the 248-column windows of the row-staging kernel
    (exactly one window, one column group more or less, ragged last windows; widths that are not a multiple of 16 take the
    warp-window kernel whatever wk_tma says), several frames, more rows than one CTA unit and a row count that is not a multiple of 8."""
    h = 19
    table = LUTS["look65"]().astype(np.float32)
    frames = _frames("yuv422p10le", w, h, ["natural", "noise", "extremes"])
    got = gpu_stage("yuv422p10le", w, h, frames, table, interp=interp, options=dict(rows_per_cta=8, **extra))
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv422p10le", w, h, got[i], oracle_stage("yuv422p10le", w, h, f, L, interp=interp),
                         f"422p10 window edges w={w} {interp} frame {i}")


@pytest.mark.parametrize("options", [dict(fast_path=0), dict(), dict(grid_ctas=2), dict(grid_ctas=444)])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear"])
def test_yuv420p_kernel_variants(options, interp):
    w, h = 328, 54
    frames = _frames("yuv420p", w, h, ["natural", "noise", "extremes"])
    for lut_name, table in (("look33", LUTS["look33"]()), ("rand64", LUTS["rand64"]()), ("wide9", synth.random_lut(9, 4) * 1.6 - 0.3)):
        table = table.astype(np.float32)
        for kw in (dict(), dict(in_range="pc", norm_range="tv")):
            got = gpu_stage("yuv420p", w, h, frames, table, interp=interp, options=options, **kw)
            L = oracle_lut(table)
            for i, f in enumerate(frames):
                assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L, interp=interp, **kw),
                                 f"420p {options} {lut_name} {interp} {kw} frame {i}")


@pytest.mark.parametrize("w,h,options", [
    (256, 12, dict()), (264, 14, dict()), (320, 38, dict(wk_unroll=2)), (384, 38, dict(wk_unroll=3)), (448, 22, dict(wk_unroll=4)),
    (1928, 16, dict()), (640, 54, dict(grid_ctas=1)), (640, 54, dict(grid_ctas=2, wk_unroll=2)), (640, 54, dict(grid_ctas=5, wk_unroll=3)),
    (704, 30, dict(grid_ctas=3, wk_unroll=4))])
@pytest.mark.parametrize("interp,kw", [("tetrahedral", dict()), ("trilinear", dict(in_range="pc", norm_range="tv"))])
def test_yuv420p_warp_windows(w, h, options, interp, kw):
    """The 4:2:0 warp-window kernel: rows narrower / wider than its windows (shifted last window), every window width, shares
    that span several windows and frames (tiny grids), frame heights down to the 12-row minimum, with and without the pc->tv pass."""
    table = LUTS["look33"]().astype(np.float32)
    frames = _frames("yuv420p", w, h, ["natural", "noise", "extremes", "noise"])
    got = gpu_stage("yuv420p", w, h, frames, table, interp=interp, options=dict(**options), **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L, interp=interp, **kw),
                         f"420p warp windows {w}x{h} {options} {interp} frame {i}")


@pytest.mark.parametrize("w,h,options", [(128, 12, dict()), (256, 14, dict()), (328, 22, dict()), (448, 38, dict(wk_unroll=2)),
                                         (640, 54, dict(grid_ctas=1, wk_unroll=3)), (704, 30, dict(grid_ctas=3, wk_unroll=4)),
                                         (1920, 24, dict())])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look65", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_range="pc", norm_range="tv")),
                                                ("tetrahedral", "wide9", dict(in_matrix="bt601"))])
def test_yuv420p_to_yuv422p10_master(w, h, options, interp, lut_name, kw):
    """Pro-mode stage 1 on an 8-bit 4:2:0 source (reference ffmpeg.py:417-433 with an 8-bit VideoInfo): yuv420p in, rgb24 LUT
    domain, yuv422p10le out.  Bit-exact against the oracle chain whose rgb24->yuv422p10le leg is pinned to libswscale."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    # (rows narrower than 256 columns with a LUT outside [0,1] / a non-power-of-two step fall back to the generic tile kernel)
    frames = _frames("yuv420p", w, h, ["natural", "noise", "extremes"])
    got = gpu_stage("yuv420p", w, h, frames, table, interp=interp, out_pix_fmt="yuv422p10le", options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv420p", w, h, f, L, interp=interp, out_pix_fmt="yuv422p10le", **kw)
        assert_bit_exact("yuv422p10le", w, h, got[i], exp, f"420p->422p10 {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("w,h,options", [(16, 14, dict()), (240, 16, dict()), (248, 18, dict()), (488, 30, dict(grid_ctas=1)),
                                         (728, 54, dict(grid_ctas=2)), (1928, 22, dict()), (656, 38, dict(grid_ctas=7))])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look65", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_matrix="bt2020", out_matrix="bt2020")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc"))])
def test_yuv420p10_to_yuv422p10_master(w, h, options, interp, lut_name, kw):
    """Pro-mode stage 1 on a 10-bit 4:2:0 source: separable 4x4-tap chroma upsampling (edge tap sets on the top rows / left
    columns), gbrp10le LUT domain, yuv422p10le out.  Shares that start anywhere in a window column (tiny grids), several frames."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames("yuv420p10le", w, h, ["natural", "noise", "extremes", "ramp"])
    got = gpu_stage("yuv420p10le", w, h, frames, table, interp=interp, out_pix_fmt="yuv422p10le", options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv420p10le", w, h, f, L, interp=interp, out_pix_fmt="yuv422p10le", **kw)
        assert_bit_exact("yuv422p10le", w, h, got[i], exp, f"420p10->422p10 {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("pix_fmt", ["yuv420p10le", "yuv422p10le"])
@pytest.mark.parametrize("w,h,options", [(16, 14, dict()), (240, 16, dict()), (248, 18, dict()), (488, 30, dict(grid_ctas=1)),
                                         (728, 54, dict(grid_ctas=2)), (1928, 22, dict()), (656, 38, dict(grid_ctas=7)),
                                         (512, 64, dict(grid_ctas=5))])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look65", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_matrix="bt2020", out_matrix="bt2020")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc"))])
def test_10bit_to_yuv420p10(pix_fmt, w, h, options, interp, lut_name, kw):
    """10-bit 4:2:0 OUTPUT (libx265 / hevc_videotoolbox 'preserve', reference ffmpeg.py:297-298) from 10-bit 4:2:0 and 4:2:2 sources:
    8-tap horizontal + 8-tap vertical chroma decimation with the output-row ring; shares that start anywhere in a window column
    (tiny grids: halo rows above / below a share), frame top / bottom clamping, several frames."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames(pix_fmt, w, h, ["natural", "noise", "extremes", "ramp"])
    got = gpu_stage(pix_fmt, w, h, frames, table, interp=interp, out_pix_fmt="yuv420p10le", options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage(pix_fmt, w, h, f, L, interp=interp, out_pix_fmt="yuv420p10le", **kw)
        assert_bit_exact("yuv420p10le", w, h, got[i], exp, f"{pix_fmt}->420p10 {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("pix_fmt", ["yuv420p10le", "yuv422p10le"])
def test_full_4k_frame_to_yuv420p10(pix_fmt):
    w, h = 3840, 2160
    table = synth.look_lut(65).astype(np.float32)
    frames = [synth.natural_frame(pix_fmt, w, h, t=5), synth.noise_frame(pix_fmt, w, h, seed=3)]
    got = gpu_stage(pix_fmt, w, h, frames, table, out_pix_fmt="yuv420p10le")
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage(pix_fmt, w, h, f, L, out_pix_fmt="yuv420p10le")
        assert_bit_exact("yuv420p10le", w, h, got[i], exp, f"{pix_fmt}->420p10 4K frame {i}")


@pytest.mark.parametrize("lut_name", ["look33", "rand65", "ident33", "rand2"])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear", "nearest"])
def test_yuv420p_interp_luts(lut_name, interp):
    w, h = 128, 36
    table = LUTS[lut_name]().astype(np.float32)
    frames = _frames("yuv420p", w, h, ["natural", "noise", "ramp", "extremes"])
    got = gpu_stage("yuv420p", w, h, frames, table, interp=interp)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L, interp=interp), f"420p {lut_name} {interp} frame {i}")


@pytest.mark.parametrize("w,h", [(16, 12), (64, 14), (200, 18), (1920, 32), (4096, 16), (8192, 12), (16384, 12)])
def test_yuv420p_geometries(w, h):
    table = LUTS["look33"]().astype(np.float32)
    frames = _frames("yuv420p", w, h, ["noise", "natural"])
    got = gpu_stage("yuv420p", w, h, frames, table)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L), f"420p {w}x{h} frame {i}")


@pytest.mark.parametrize("kw", [dict(in_range="pc", norm_range="tv", interp="trilinear"),           # BASELINE cfg 2
                                dict(in_range="pc", norm_range="pc"),                              # tags=none on a pc source
                                dict(in_matrix="bt601", out_matrix="bt709"),
                                dict(in_matrix="bt2020", out_matrix="bt601")])
def test_yuv420p_ranges_matrices(kw):
    w, h = 64, 20
    table = LUTS["look33"]().astype(np.float32)
    frames = [synth.noise_frame("yuv420p", w, h, seed=3), synth.natural_frame("yuv420p", w, h, t=1, full_range=kw.get("in_range") == "pc")]
    got = gpu_stage("yuv420p", w, h, frames, table, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L, **kw), f"420p {kw} frame {i}")


def test_full_1080p_frame_cfg1():
    """BASELINE cfg 1 at its real size: 1920x1080 yuv420p, 33-point LUT, tetrahedral."""
    w, h = 1920, 1080
    table = LUTS["look33"]().astype(np.float32)
    frames = [synth.natural_frame("yuv420p", w, h, t=0), synth.noise_frame("yuv420p", w, h, seed=1)]
    got = gpu_stage("yuv420p", w, h, frames, table)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv420p", w, h, got[i], oracle_stage("yuv420p", w, h, f, L), f"cfg1 frame {i}")


def test_full_4k_frame_cfg3():
    """BASELINE cfg 3 at its real size: 3840x2160 yuv422p10le, 65-point LUT, tetrahedral."""
    w, h = 3840, 2160
    table = LUTS["look65"]().astype(np.float32)
    frames = [synth.natural_frame("yuv422p10le", w, h, t=0)]
    got = gpu_stage("yuv422p10le", w, h, frames, table)
    L = oracle_lut(table)
    assert_bit_exact("yuv422p10le", w, h, got[0], oracle_stage("yuv422p10le", w, h, frames[0], L), "cfg3")


def test_strided_planes_and_frame_step():
    """Padded rows and a frame step larger than the frame (what a pooled allocator hands out)."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    w, h, n = 64, 12, 3
    table = LUTS["rand17"]().astype(np.float32)
    st = ColourStage(w, h, "yuv422p10le")
    lut = st.ctx.lut_from_table(table)
    frames = [synth.noise_frame("yuv422p10le", w, h, seed=20 + i) for i in range(n)]
    ystride, cstride = 2 * w + 32, w + 16
    step = ((ystride + 2 * cstride) * h + 255) // 256 * 256
    src = torch.zeros(n * step, dtype=torch.uint8, device="cuda")
    dst = torch.full((n * step,), 0xAB, dtype=torch.uint8, device="cuda")
    offs = [0, ystride * h, ystride * h + cstride * h]
    host = np.zeros(n * step, np.uint8)
    for i, f in enumerate(frames):
        for p, (pl, stride) in enumerate(zip(synth.split_planes(f, "yuv422p10le", w, h), (ystride, cstride, cstride))):
            rows = pl.view(np.uint8).reshape(h, -1)
            for r in range(h):
                o = i * step + offs[p] + r * stride
                host[o:o + rows.shape[1]] = rows[r]
    src.copy_(torch.from_numpy(host))
    st.run_planes(lut, [src.data_ptr() + o for o in offs], (ystride, cstride, cstride),
                  [dst.data_ptr() + o for o in offs], (ystride, cstride, cstride), n, step, step,
                  stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    out = dst.cpu().numpy()
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = synth.split_planes(oracle_stage("yuv422p10le", w, h, f, L), "yuv422p10le", w, h)
        for p, (pl, stride) in enumerate(zip(exp, (ystride, cstride, cstride))):
            rows = pl.view(np.uint8).reshape(h, -1)
            for r in range(h):
                o = i * step + offs[p] + r * stride
                assert np.array_equal(out[o:o + rows.shape[1]], rows[r]), (i, p, r)
                assert (out[o + rows.shape[1]:o + stride] == 0xAB).all(), "padding was written"
    st.close()


def test_run_host_matches_device_path():
    from lut_renderer_b200.stage import ColourStage
    w, h, n = 256, 16, 70   # more frames than one pipeline slot for this size? (chunk is capped at 64)
    table = LUTS["look33"]().astype(np.float32)
    frames = [synth.noise_frame("yuv420p", w, h, seed=100 + i) for i in range(n)]
    st = ColourStage(w, h, "yuv420p")
    lut = st.ctx.lut_from_table(table)
    src = np.concatenate(frames)
    dst = np.zeros(n * st.out_frame_bytes, np.uint8)
    st.run_host(lut, src, dst, n)
    pin_in, pin_out = st.ctx.pinned_buffer(src.nbytes), st.ctx.pinned_buffer(dst.nbytes)
    pin_in.array[:] = src
    st.run_host(lut, pin_in, pin_out, n)
    assert np.array_equal(pin_out.array, dst)
    L = oracle_lut(table)
    for i in (0, 1, 63, 64, 69):
        exp = oracle_stage("yuv420p", w, h, frames[i], L)
        assert_bit_exact("yuv420p", w, h, dst[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes], exp, f"run_host frame {i}")
    assert st.launch_count >= 4
    pin_in.close(), pin_out.close(), st.close(), lut.close()


def test_cube_file_to_pixels(tmp_path):
    """The whole drop-in path: .cube text on disk -> parser -> upload -> kernel, against the oracle's own parser."""
    from oracle import oracle as O
    from lut_renderer_b200.stage import ColourStage
    import torch
    w, h = 128, 12
    path = synth.write_cube(tmp_path / "look.cube", synth.look_lut(33), domain=((0, 0, 0), (1, 1, 1)), crlf=True)
    st = ColourStage(w, h, "yuv422p10le")
    lut = st.ctx.lut_from_cube(path)
    assert st.ctx.lut_from_cube(path) is lut          # cached per device by (path, mtime, size)
    f = synth.natural_frame("yuv422p10le", w, h, t=2)
    src = torch.from_numpy(f).cuda()
    dst = torch.zeros(st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(lut, src, dst, 1, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    exp = oracle_stage("yuv422p10le", w, h, f, O.OracleLut.from_path(path))
    assert_bit_exact("yuv422p10le", w, h, dst.cpu().numpy(), exp, "cube file path")
    st.close()


def test_fma_mode_matches_the_oracle_fma_order():
    """float_mode=fma (the contracted multiply-adds of lut3d's x86 assembly, same operand order) is an opt-in; it is held to the
    oracle's restatement of that order bit for bit (every kernel family that implements it), and stays within one code of the C
    semantics."""
    from oracle import oracle as O
    table = synth.identity_lut(17).astype(np.float32)      # knife-edge results: the two orders differ in ~0.5 % of the RGB samples
    L = oracle_lut(table)
    for fmt, out_fmt, w, h, interp in (("yuv422p10le", "yuv422p10le", 256, 16, "tetrahedral"), ("yuv422p10le", "yuv422p10le", 256, 16, "trilinear"),
                                       ("yuv420p", "yuv420p", 256, 32, "tetrahedral"), ("yuv420p10le", "yuv420p", 256, 32, "tetrahedral"),
                                       ("yuv444p10le", "yuv444p10le", 136, 16, "trilinear")):
        frames = [synth.natural_frame(fmt, w, h, t=5), synth.noise_frame(fmt, w, h, seed=8)]
        c = gpu_stage(fmt, w, h, frames, table, out_pix_fmt=out_fmt, interp=interp, float_mode="c")
        f = gpu_stage(fmt, w, h, frames, table, out_pix_fmt=out_fmt, interp=interp, float_mode="fma")
        try:
            O.set_float_mode("fma")
            exp = [oracle_stage(fmt, w, h, fr, L, out_pix_fmt=out_fmt, interp=interp) for fr in frames]
        finally:
            O.set_float_mode("c")
        dt = "<u2" if "10" in out_fmt else np.uint8
        for i in range(len(frames)):
            assert_bit_exact(out_fmt, w, h, f[i], exp[i], f"fma {fmt}->{out_fmt} {interp} frame {i}")
            d = np.abs(c[i].view(dt).astype(int) - f[i].view(dt).astype(int))
            assert d.max() <= 1 and (d == 0).mean() > 0.98


# ---- the generic tile kernel (k_any.cu): the whole format matrix of the operator -------------------------------------------------
ALL_FMTS = ["yuv420p", "yuv422p", "yuv444p", "yuv420p10le", "yuv422p10le", "yuv444p10le"]


@pytest.mark.parametrize("in_fmt", ALL_FMTS)
@pytest.mark.parametrize("out_fmt", ALL_FMTS)
@pytest.mark.parametrize("out_range", ["tv", "pc"])
def test_any_format_matrix(in_fmt, out_fmt, out_range):
    """Every source layout x every output layout / depth x both output ranges through ONE launch of the generic tile kernel
    (forced with any_kernel=1 where a specialised kernel exists): 8-bit sources go through rgb24, 10-bit through gbrp10le; the
    10 -> 8 bit outputs carry swscale's ordered dither.  Geometry: two tile columns with a ragged right tile, three tile rows."""
    w, h = 200, 70 if "420" not in in_fmt + out_fmt else 72
    table = synth.look_lut(17).astype(np.float32)
    frames = _frames(in_fmt, w, h, ["natural", "noise", "extremes"])
    got = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt, out_range=out_range, options=dict(any_kernel=1))
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt, out_range=out_range)
        assert_bit_exact(out_fmt, w, h, got[i], exp, f"{in_fmt}->{out_fmt} {out_range} frame {i}")


@pytest.mark.parametrize("interp", ["nearest", "trilinear", "tetrahedral", "pyramid", "prism"])
@pytest.mark.parametrize("in_fmt,out_fmt", [("yuv420p", "yuv444p"), ("yuv422p10le", "yuv420p"), ("yuv444p10le", "yuv422p10le")])
def test_any_interps_and_domain_scales(interp, in_fmt, out_fmt):
    """All five interpolation modes of the reference's whitelist (ffmpeg.py:243) incl. pyramid / prism, a LUT that leaves [0,1]
    and per-channel DOMAIN scales."""
    w, h = 136, 36
    table = (synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames(in_fmt, w, h, ["noise", "ramp"])
    for scale in ((1.0, 1.0, 1.0), (0.5, 1.0, 0.25)):
        got = gpu_stage(in_fmt, w, h, frames, table, scale=scale, out_pix_fmt=out_fmt, interp=interp)
        L = oracle_lut(table, scale)
        for i, f in enumerate(frames):
            exp = oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt, interp=interp)
            assert_bit_exact(out_fmt, w, h, got[i], exp, f"{in_fmt}->{out_fmt} {interp} scale {scale} frame {i}")


@pytest.mark.parametrize("in_fmt", ["yuv420p", "yuv422p", "yuv444p"])
@pytest.mark.parametrize("kw", [dict(in_range="pc", norm_range="tv"), dict(in_range="pc", norm_range="pc", out_range="pc"),
                                dict(in_range="pc", norm_range="tv", in_matrix="bt601", out_matrix="bt2020")])
def test_any_range_pass_and_matrices(in_fmt, kw):
    """The reference's full-range chains (ffmpeg.py:211-233): pc -> tv pass re-quantised to the 8-bit layout of the source
    (tags=bt709), and the pass-through pc -> pc chain with a full-range output (tags=inherit / none)."""
    w, h = 264, 40
    table = synth.look_lut(33).astype(np.float32)
    frames = _frames(in_fmt, w, h, ["natural", "extremes"])
    got = gpu_stage(in_fmt, w, h, frames, table, options=dict(any_kernel=1), **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact(in_fmt, w, h, got[i], oracle_stage(in_fmt, w, h, f, L, **kw), f"{in_fmt} {kw} frame {i}")


@pytest.mark.parametrize("in_fmt,out_fmt,w,h", [("yuv422p10le", "yuv420p", 1920, 1080), ("yuv420p", "yuv444p", 1920, 1080),
                                                ("yuv444p10le", "yuv420p10le", 1280, 720), ("yuv420p10le", "yuv420p", 3840, 2160)])
def test_any_full_frames(in_fmt, out_fmt, w, h):
    """Full delivery frames: e.g. a 10-bit camera source to 8-bit H.264 (yuv420p for non-10-bit codecs, ffmpeg.py:299-302)."""
    table = synth.look_lut(33).astype(np.float32)
    frames = [synth.natural_frame(in_fmt, w, h, t=2), synth.noise_frame(in_fmt, w, h, seed=8)]
    got = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact(out_fmt, w, h, got[i], oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt), f"{in_fmt}->{out_fmt} frame {i}")


def test_any_kernel_agrees_with_specialised_kernels():
    """The generic tile kernel and the warp-window kernels are independent implementations of the same stage: identical bytes."""
    table = synth.look_lut(65).astype(np.float32)
    for in_fmt, out_fmt, w, h in (("yuv422p10le", "yuv422p10le", 496, 38), ("yuv420p", "yuv420p", 512, 40),
                                  ("yuv420p", "yuv422p10le", 328, 22), ("yuv420p10le", "yuv420p10le", 488, 30)):
        frames = _frames(in_fmt, w, h, ["natural", "noise"])
        a = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt)
        b = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt, options=dict(any_kernel=1))
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (in_fmt, out_fmt)


@pytest.mark.parametrize("pix_fmt", ["yuv420p10le", "yuv422p10le"])
@pytest.mark.parametrize("w,h,options", [(16, 14, dict()), (240, 16, dict()), (248, 18, dict()), (488, 30, dict(grid_ctas=1)),
                                         (728, 54, dict(grid_ctas=2)), (1928, 22, dict()), (656, 38, dict(grid_ctas=7)),
                                         (512, 64, dict(grid_ctas=5))])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look65", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_matrix="bt2020", out_matrix="bt2020")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc"))])
def test_10bit_to_yuv420p_8bit_delivery(pix_fmt, w, h, options, interp, lut_name, kw):
    """A 10-bit source delivered as 8-bit yuv420p (the reference's choice for codecs without 10-bit support, e.g. the default
    libx264: ffmpeg.py:299-302): gbrp10le intermediate, swscale's ordered dither on luma and chroma (V shifted by 3 columns),
    per-tap truncating vertical scaler with folded edge taps and the exact C-path last row.  Warp-window kernel k10_warp8."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames(pix_fmt, w, h, ["natural", "noise", "extremes", "ramp"])
    got = gpu_stage(pix_fmt, w, h, frames, table, interp=interp, out_pix_fmt="yuv420p", options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage(pix_fmt, w, h, f, L, interp=interp, out_pix_fmt="yuv420p", **kw)
        assert_bit_exact("yuv420p", w, h, got[i], exp, f"{pix_fmt}->420p8 {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("in_fmt,out_fmt,w,h,options", [
    ("yuv422p10le", "yuv422p10le", 488, 10, dict()), ("yuv420p", "yuv420p", 328, 22, dict(grid_ctas=2)),
    ("yuv420p", "yuv422p10le", 328, 22, dict(grid_ctas=2)), ("yuv420p10le", "yuv422p10le", 488, 30, dict(grid_ctas=2)),
    ("yuv420p10le", "yuv420p10le", 488, 30, dict(grid_ctas=2)), ("yuv422p10le", "yuv420p10le", 488, 30, dict(grid_ctas=3)),
    ("yuv420p10le", "yuv420p", 488, 30, dict(grid_ctas=2)), ("yuv422p10le", "yuv420p", 496, 30, dict(grid_ctas=5)),
    ("yuv444p", "yuv420p10le", 200, 36, dict()), ("yuv420p10le", "yuv444p", 200, 36, dict()),
    ("yuv422p10le", "yuv422p10le", 256, 6, dict(fast_path=0)), ("yuv422p10le", "yuv422p10le", 256, 6, dict(wk_tma=1)),
    ("yuv422p10le", "yuv422p10le", 256, 6, dict(wk_tma=2)), ("yuv422p10le", "yuv422p10le", 512, 9, dict(wk_tma=2, rows_per_cta=8)),
    ("yuv420p", "yuv420p", 64, 14, dict())])
def test_no_kernel_writes_outside_its_frames(in_fmt, out_fmt, w, h, options):
    """Guard bands around the output batch stay untouched for every kernel family (compute-sanitizer is not available on this
    pool): the frames sit inside a larger buffer filled with a sentinel, 4 KB before and after."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    table = synth.look_lut(17).astype(np.float32)
    frames = [synth.noise_frame(in_fmt, w, h, seed=s) for s in (1, 2, 3)]
    st = ColourStage(w, h, in_fmt, out_pix_fmt=out_fmt)
    for k, v in options.items():
        st.set_option(k, v)
    lut = st.ctx.lut_from_table(table)
    guard, n = 4096, len(frames)
    src = torch.from_numpy(np.concatenate(frames)).cuda()
    buf = torch.full((guard + n * st.out_frame_bytes + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    st.run(lut, src, buf[guard:guard + n * st.out_frame_bytes], n, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    out = buf.cpu().numpy()
    assert (out[:guard] == 0xA5).all() and (out[-guard:] == 0xA5).all()
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        got = out[guard + i * st.out_frame_bytes: guard + (i + 1) * st.out_frame_bytes]
        assert_bit_exact(out_fmt, w, h, got, oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt), f"{in_fmt}->{out_fmt} frame {i}")
    st.close()
    lut.close()


def test_fuzz_geometries_formats_options():
    """Seeded random sweep over layouts, geometries, grids, LUTs, interps, matrices and ranges: whatever kernel the library picks
    (or is forced to pick) must match the oracle bit for bit.  Complements the hand-picked edge cases above."""
    rng = np.random.default_rng(20261018)
    luts = [synth.look_lut(17), synth.look_lut(33), synth.random_lut(9, 7) * 1.4 - 0.2, synth.random_lut(20, 8), synth.identity_lut(65)]
    cases = 0
    for trial in range(48):
        in_fmt, out_fmt = rng.choice(ALL_FMTS), rng.choice(ALL_FMTS)
        w = int(rng.choice([16, 24, 64, 120, 136, 240, 248, 256, 264, 376, 488, 512, 520, 736, 1032]))
        h = int(rng.choice([14, 16, 18, 22, 30, 34, 46, 66]))
        interp = str(rng.choice(["tetrahedral", "tetrahedral", "trilinear", "nearest", "pyramid", "prism"]))
        kw = dict(in_matrix=str(rng.choice(["bt709", "bt601", "bt2020"])), out_matrix=str(rng.choice(["bt709", "bt601", "bt2020"])))
        if rng.random() < 0.25:
            kw.update(in_range="pc", norm_range="tv" if "10" not in in_fmt and rng.random() < 0.6 else "pc")
        if rng.random() < 0.25:
            kw.update(out_range="pc")
        options = {}
        if rng.random() < 0.5:
            options["grid_ctas"] = int(rng.integers(1, 9))
        if rng.random() < 0.2:
            options["any_kernel"] = 1
        if rng.random() < 0.3:
            options["wk_tma"] = int(rng.integers(1, 3))      # row-staging kernel where the stage and geometry allow it
        scale = (1.0, 1.0, 1.0) if rng.random() < 0.8 else (0.5, 1.0, 0.75)
        table = luts[int(rng.integers(len(luts)))].astype(np.float32)
        frames = [synth.noise_frame(in_fmt, w, h, seed=int(rng.integers(1 << 30))),
                  synth.natural_frame(in_fmt, w, h, t=trial, full_range=kw.get("in_range") == "pc")]
        got = gpu_stage(in_fmt, w, h, frames, table, scale=scale, out_pix_fmt=out_fmt, interp=interp, options=options, **kw)
        L = oracle_lut(table, scale)
        for i, f in enumerate(frames):
            exp = oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt, interp=interp, **kw)
            assert_bit_exact(out_fmt, w, h, got[i], exp, f"fuzz {trial}: {in_fmt}->{out_fmt} {w}x{h} {interp} {kw} {options} scale {scale} frame {i}")
        cases += 1
    assert cases == 48


@pytest.mark.parametrize("in_fmt,out_fmt,w,h", [("yuv420p", "yuv420p", 854, 480), ("yuv420p10le", "yuv420p", 854, 480),
                                                ("yuv422p10le", "yuv422p10le", 1366, 18), ("yuv420p", "yuv422p10le", 18, 16),
                                                ("yuv420p10le", "yuv420p10le", 22, 30), ("yuv444p", "yuv420p", 50, 14),
                                                ("yuv422p", "yuv444p10le", 134, 12), ("yuv420p", "yuv420p", 130, 34)])
def test_widths_that_are_not_multiples_of_8(in_fmt, out_fmt, w, h):
    """Even widths that are not a multiple of 8 (854x480, 1366 wide, ...): the warp-window kernels work on 8-column groups, so the
    generic tile kernel takes these frames -- same oracle, same bit-exact bar, ragged last tile / pair, odd-sized chroma planes."""
    table = synth.look_lut(33).astype(np.float32)
    frames = _frames(in_fmt, w, h, ["natural", "noise", "extremes"])
    got = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        assert_bit_exact(out_fmt, w, h, got[i], oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt), f"{in_fmt}->{out_fmt} {w}x{h} frame {i}")


@pytest.mark.parametrize("in_fmt", ["yuv420p10le", "yuv422p10le", "yuv444p10le"])
@pytest.mark.parametrize("norm_range,out_fmt,out_range", [("tv", "yuv420p", "tv"), ("tv", "yuv420p10le", "tv"), ("pc", "yuv422p", "pc"),
                                                           ("pc", "yuv444p10le", "tv")])
def test_full_range_10bit_source_via_8bit_intermediate(in_fmt, norm_range, out_fmt, out_range):
    """Full-range 10-bit sources as the reference treats them (ffmpeg.py:211-233): one swscale pass to 8-bit yuv4xxp (with the
    pc->tv conversion and ordered dither, or as a dithered planar copy when the range is kept), then the 8-bit chain through rgb24."""
    w, h = 264, 40
    table = synth.look_lut(33).astype(np.float32)
    frames = [synth.natural_frame(in_fmt, w, h, t=2, full_range=True), synth.noise_frame(in_fmt, w, h, 9), synth.ramp_frame(in_fmt, w, h)]
    got = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt, in_range="pc", norm_range=norm_range, out_range=out_range, via_8bit=True)
    L = oracle_lut(table)
    from oracle import oracle as O
    for i, f in enumerate(frames):
        Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(f, in_fmt, w, h)]
        exp = synth.pack_planes(O.stage_any(L, in_fmt, out_fmt, Y, U, V, range_pass=1 if norm_range == "tv" else 2,
                                            in_full_range=(norm_range == "pc"), out_full_range=(out_range == "pc")))
        assert_bit_exact(out_fmt, w, h, got[i], exp, f"{in_fmt} pc->{norm_range} via 8 bit -> {out_fmt} {out_range} frame {i}")


@pytest.mark.parametrize("out_fmt", ["yuv422p", "yuv444p"])
@pytest.mark.parametrize("w,h,options", [(128, 12, dict()), (256, 14, dict()), (328, 22, dict()), (448, 38, dict(wk_unroll=2)),
                                         (640, 30, dict(grid_ctas=3)), (1920, 36, dict()), (200, 16, dict(grid_ctas=1))])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look33", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_range="pc", norm_range="tv")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc"))])
def test_yuv420p_to_8bit_422_444(out_fmt, w, h, options, interp, lut_name, kw):
    """The GUI's pix_fmt choices yuv422p / yuv444p on an 8-bit 4:2:0 source (reference main_window.py:767-771): the 4:2:0 warp-window
    kernel with its row-local 8-bit output sides (pair-sum chroma for 4:2:2, per-pixel chroma for 4:4:4); window widths, tiny grids,
    range pass, a LUT that leaves [0,1]."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames("yuv420p", w, h, ["natural", "noise", "extremes"])
    got = gpu_stage("yuv420p", w, h, frames, table, interp=interp, out_pix_fmt=out_fmt, options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv420p", w, h, f, L, interp=interp, out_pix_fmt=out_fmt, **kw)
        assert_bit_exact(out_fmt, w, h, got[i], exp, f"420p->{out_fmt} {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("out_fmt", ["yuv422p", "yuv420p", "yuv444p", "yuv422p10le"])
@pytest.mark.parametrize("w,h,options", [(128, 12, dict()), (256, 14, dict()), (328, 22, dict()), (448, 38, dict(wk_unroll=2)),
                                         (640, 31, dict(grid_ctas=3)), (1920, 36, dict()), (200, 16, dict(grid_ctas=1))])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look33", dict()), ("trilinear", "look33", dict(in_range="pc", norm_range="tv")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc", out_matrix="bt601")),
                                                ("tetrahedral", "look33", dict(scale=(0.5, 1.0, 0.75)))])
def test_yuv422p_8bit_source(out_fmt, w, h, options, interp, lut_name, kw):
    """8-bit 4:2:2 sources (yuv422p / yuvj422p: MJPEG cameras, the 4:2:2 range-pass intermediate of ffmpeg.py:137-143) through the
    warp-window kernel's IN422 side -- one chroma row per luma row, otherwise the 4:2:0 arithmetic -- to every row-local output and
    to yuv420p; odd heights for the 4:2:2 -> 4:2:2 case; per-channel DOMAIN scales."""
    kw = dict(kw)
    scale = kw.pop("scale", (1.0, 1.0, 1.0))
    if out_fmt == "yuv420p" and h % 2:
        h += 1
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames("yuv422p", w, h, ["natural", "noise", "extremes"])
    got = gpu_stage("yuv422p", w, h, frames, table, scale=scale, interp=interp, out_pix_fmt=out_fmt, options=options, **kw)
    L = oracle_lut(table, scale)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv422p", w, h, f, L, interp=interp, out_pix_fmt=out_fmt, **kw)
        assert_bit_exact(out_fmt, w, h, got[i], exp, f"422p->{out_fmt} {w}x{h} {options} {interp} {lut_name} {kw} frame {i}")


def test_gpu_against_committed_golden_digests():
    """The CUDA path against the COMMITTED golden digests (tests/golden/oracle_hashes.json, 80 seeded cases over layouts, interps,
    range handling, per-channel DOMAIN scales) -- no oracle code runs here, only its frozen outputs."""
    import hashlib
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "oracle_hashes.json")) as f:
        want = json.load(f)
    luts = {"look17": (synth.look_lut(17), (1.0, 1.0, 1.0)), "rand9": (synth.random_lut(9, 4) * 1.4 - 0.2, (0.5, 1.0, 0.75))}
    w, h = 200, 36
    for key, digest in sorted(want.items()):
        lut_name, fmts, interp, rp, ofr = key.split("|")
        in_fmt, out_fmt = fmts.split("->")
        rp, ofr = int(rp.split("=")[1]), int(ofr.split("=")[1])
        table, scale = luts[lut_name]
        kw = dict(out_range="pc" if ofr else "tv")
        if rp == 1:
            kw.update(in_range="pc", norm_range="tv", via_8bit="10" in in_fmt)
        elif rp == 2:
            kw.update(in_range="pc", norm_range="pc", via_8bit=True)
        frame = synth.noise_frame(in_fmt, w, h, seed=17)
        got = gpu_stage(in_fmt, w, h, [frame], table.astype(np.float32), scale=scale, out_pix_fmt=out_fmt, interp=interp, **kw)[0]
        assert hashlib.sha256(got.tobytes()).hexdigest() == digest, key


@pytest.mark.parametrize("n", [129, 256])
def test_largest_lut_sizes(n):
    """LUT_3D_SIZE up to lut3d's maximum of 256 (16.7 M nodes, a 268 MB device table): node index arithmetic of every kernel family
    near its 32-bit limits, noise frames so that the whole table is addressed."""
    rng = np.random.default_rng(n)
    table = rng.random((n, n, n, 3), dtype=np.float32)
    L = oracle_lut(table)
    for in_fmt, out_fmt, w, h, opts in (("yuv422p10le", "yuv422p10le", 496, 6, dict()), ("yuv420p", "yuv420p", 256, 16, dict()),
                                        ("yuv420p10le", "yuv420p", 256, 16, dict()), ("yuv444p", "yuv444p10le", 136, 8, dict()),
                                        ("yuv422p10le", "yuv422p10le", 256, 4, dict(wk_tma=2))):
        frames = [synth.noise_frame(in_fmt, w, h, seed=n)]
        got = gpu_stage(in_fmt, w, h, frames, table, out_pix_fmt=out_fmt, options=opts)
        assert_bit_exact(out_fmt, w, h, got[0], oracle_stage(in_fmt, w, h, frames[0], L, out_pix_fmt=out_fmt), f"n={n} {in_fmt}->{out_fmt} {opts}")


@pytest.mark.parametrize("w,h,options", [(16, 2, dict()), (240, 9, dict()), (248, 16, dict()), (496, 33, dict(grid_ctas=2)),
                                         (728, 24, dict(rows_per_cta=8)), (1928, 12, dict()), (3840, 4, dict())])
@pytest.mark.parametrize("interp,lut_name,kw", [("tetrahedral", "look65", dict()), ("tetrahedral", "rand64", dict(out_matrix="bt601")),
                                                ("trilinear", "look33", dict(in_matrix="bt2020", out_matrix="bt2020")),
                                                ("tetrahedral", "wide9", dict(in_range="pc", norm_range="pc"))])
def test_yuv444p10_source_to_master(w, h, options, interp, lut_name, kw):
    """10-bit 4:4:4 sources (ProRes 4444 and the like) to the pro-mode master yuv422p10le on the headline kernel's IN444 side: per-pixel
    chroma, no upsampling filter; window edges, odd heights, several frames."""
    table = (LUTS[lut_name]() if lut_name in LUTS else synth.random_lut(9, 4) * 1.6 - 0.3).astype(np.float32)
    frames = _frames("yuv444p10le", w, h, ["natural", "noise", "extremes", "ramp"])
    got = gpu_stage("yuv444p10le", w, h, frames, table, interp=interp, out_pix_fmt="yuv422p10le", options=options, **kw)
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage("yuv444p10le", w, h, f, L, interp=interp, out_pix_fmt="yuv422p10le", **kw)
        assert_bit_exact("yuv422p10le", w, h, got[i], exp, f"444p10->422p10 {w}x{h} {options} {interp} {lut_name} frame {i}")


@pytest.mark.parametrize("out_fmt", ["yuv420p", "yuv420p10le", "yuv422p"])
@pytest.mark.parametrize("w,h", [(16, 14), (72, 16), (264, 40), (1920, 36), (3840, 2160)])
def test_convert_master_without_lut(out_fmt, w, h):
    """Pro mode stage 2 (reference ffmpeg.py:459-471): the yuv422p10le master to the delivery pix_fmt without a LUT -- interp="none",
    lut=None -- bit-exact against the oracle's restatement of the encoding ffmpeg's swscale pass (itself pinned to libswscale):
    dithered luma, 8-tap vertical chroma decimation with folded edge rows and the exact last row, several frames."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    from oracle import oracle as O
    frames = _frames("yuv422p10le", w, h, ["noise", "natural", "extremes"] if w < 3000 else ["noise"])
    st = ColourStage(w, h, "yuv422p10le", out_pix_fmt=out_fmt, interp="none")
    src = torch.from_numpy(np.concatenate(frames)).cuda()
    dst = torch.zeros(len(frames) * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(None, src, dst, len(frames), stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert st.kernel_name == "k_convert" and st.launch_count == 1
    out = dst.cpu().numpy()
    for i, f in enumerate(frames):
        Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(f, "yuv422p10le", w, h)]
        exp = synth.pack_planes(O.convert_yuv422p10(out_fmt, Y, U, V))
        assert_bit_exact(out_fmt, w, h, out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes], exp, f"convert -> {out_fmt} {w}x{h} frame {i}")
    # host path of the same stage (what the runner's pump calls)
    pin_in, pin_out = st.ctx.pinned_buffer(len(frames) * st.in_frame_bytes), st.ctx.pinned_buffer(len(frames) * st.out_frame_bytes)
    pin_in.array[:] = np.concatenate(frames)
    st.run_host(None, pin_in, pin_out, len(frames))
    assert np.array_equal(pin_out.array, out)
    pin_in.close(), pin_out.close(), st.close()


def test_convert_stage_rejects_what_it_does_not_model():
    from lut_renderer_b200 import _native as N
    from lut_renderer_b200.stage import ColourStage
    for kw in (dict(pix_fmt="yuv420p10le", out_pix_fmt="yuv420p"), dict(pix_fmt="yuv422p10le", out_pix_fmt="yuv444p"),
               dict(pix_fmt="yuv422p10le", out_pix_fmt="yuv420p", out_matrix="bt601"), dict(pix_fmt="yuv422p10le", out_pix_fmt="yuv420p", out_range="pc")):
        with pytest.raises(N.LutrError) as e:
            ColourStage(256, 16, interp="none", **kw)
        assert e.value.status == N.ERR_UNSUPPORTED


@pytest.mark.parametrize("in_fmt,out_fmt,kernel", [("yuv420p", "yuv420p", "k420_warp"), ("yuv420p", "yuv422p", "k420_warp422"),
                                                   ("yuv420p", "yuv444p", "k420_warp444"), ("yuv422p", "yuv420p", "k420_warp<422>"),
                                                   ("yuv422p", "yuv422p", "k420_warp422<422>")])
@pytest.mark.parametrize("interp,norm_range", [("tetrahedral", "pc"), ("trilinear", "pc"), ("tetrahedral", "tv")])
def test_full_range_output_on_the_8bit_warp_kernels(in_fmt, out_fmt, kernel, interp, norm_range):
    """tags = inherit / none on a full-range source (reference ffmpeg.py:216-223): the stage output is full range.  The 8-bit
    warp-window kernels carry swscale's 15-bit range conversion as a run-time switch instead of handing the stage to the generic
    kernel; several rows per share, window edges, the last chroma row."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    w, h = 328, 38
    table = LUTS["look33"]().astype(np.float32)
    frames = _frames(in_fmt, w, h, ["natural", "noise", "extremes"])
    kw = dict(in_range="pc", norm_range=norm_range, out_range="pc", out_matrix="bt601")
    st = ColourStage(w, h, in_fmt, out_pix_fmt=out_fmt, interp=interp, **kw)
    lut = st.ctx.lut_from_table(table)
    src = torch.from_numpy(np.concatenate(frames)).cuda()
    dst = torch.zeros(len(frames) * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(lut, src, dst, len(frames), stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    # (a pc -> tv pass in front AND a full-range output is outside what the reference's options produce: generic kernel)
    assert st.kernel_name == (kernel if norm_range == "pc" else "k_any"), st.kernel_name
    out = dst.cpu().numpy()
    L = oracle_lut(table)
    for i, f in enumerate(frames):
        exp = oracle_stage(in_fmt, w, h, f, L, interp=interp, out_pix_fmt=out_fmt, **kw)
        assert_bit_exact(out_fmt, w, h, out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes], exp, f"pc out {in_fmt}->{out_fmt} {interp} frame {i}")
    st.close(), lut.close()


@pytest.mark.parametrize("in_fmt,out_fmt,kernel", [("yuv422p10le", "yuv422p10le", "k422_warp"), ("yuv444p10le", "yuv422p10le", "k422_warp"),
                                                   ("yuv420p10le", "yuv422p10le", "k10_warp"), ("yuv420p10le", "yuv420p10le", "k10_warp"),
                                                   ("yuv422p10le", "yuv420p", "k10_warp8"), ("yuv420p", "yuv420p", "k420_warp"),
                                                   ("yuv422p", "yuv422p10le", "k420_warp10"), ("yuv420p", "yuv444p", "k420_warp444")])
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear"])
@pytest.mark.parametrize("scale", [(0.5, 1.0, 0.25), (0.75, 0.5, 1.0)])
def test_per_channel_domain_scales_on_the_warp_kernels(in_fmt, out_fmt, kernel, interp, scale):
    """A .cube whose DOMAIN_MAX differs per channel (vf_lut3d: scale.c = clip(1 / (max.c - min.c), 0, 1), one per channel; reference
    ffmpeg.py:246 hands the file to lut3d untouched): the warp-window kernels' general index path multiplies by a per-channel
    constant, so the stage stays off the generic kernel.  17- and 33-point tables, both float orders."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    from oracle import oracle as O
    w, h = 328, 38
    for lut_name, fm in (("rand17", "c"), ("look33", "fma")):
        table = LUTS[lut_name]().astype(np.float32)
        frames = _frames(in_fmt, w, h, ["natural", "noise", "extremes"])
        st = ColourStage(w, h, in_fmt, out_pix_fmt=out_fmt, interp=interp, float_mode=fm)
        lut = st.ctx.lut_from_table(table, scale)
        src = torch.from_numpy(np.concatenate(frames)).cuda()
        dst = torch.zeros(len(frames) * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
        st.run(lut, src, dst, len(frames), stream=torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert st.kernel_name.startswith(kernel), st.kernel_name
        out = dst.cpu().numpy()
        L = oracle_lut(table, scale)
        O.set_float_mode(fm)
        try:
            for i, f in enumerate(frames):
                exp = oracle_stage(in_fmt, w, h, f, L, interp=interp, out_pix_fmt=out_fmt)
                assert_bit_exact(out_fmt, w, h, out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes], exp,
                                 f"per-channel DOMAIN {scale} {in_fmt}->{out_fmt} {interp} {lut_name} {fm} frame {i}")
        finally:
            O.set_float_mode("c")
        st.close(), lut.close()


FIXED_CELLS = [("yuv422p10le", "yuv422p10le"), ("yuv420p10le", "yuv422p10le"), ("yuv420p10le", "yuv420p10le"), ("yuv420p10le", "yuv420p"),
               ("yuv422p10le", "yuv420p"), ("yuv420p", "yuv420p"), ("yuv420p", "yuv444p"), ("yuv420p", "yuv422p10le")]


@pytest.mark.parametrize("in_fmt,out_fmt", FIXED_CELLS)
@pytest.mark.parametrize("interp", ["tetrahedral", "trilinear"])
def test_any_fixed_format_instantiations(in_fmt, out_fmt, interp):
    """The generic tile kernel exists twice for the cells the reference's options reach most (k_any_fixed8.cu / k_any_fixed10.cu: the
    layouts as compile-time constants; wk_var=4000 keeps the run-time-format instantiation): both bit-exact against the oracle, on a
    width off the 8-column grid (what sends these cells to the generic kernel in the first place), with a pc -> tv pass in front for
    the 8-bit sources, a per-channel DOMAIN and a table that leaves [0, 1]."""
    w, h = 854 if "10" in in_fmt else 206, 72
    table = (synth.random_lut(17, 77) * 1.3 - 0.1).astype(np.float32)
    scale = (0.5, 1.0, 0.75)
    kw = dict(in_range="pc", norm_range="tv") if "10" not in in_fmt else {}
    frames = _frames(in_fmt, w, h, ["natural", "noise", "extremes"])
    L = oracle_lut(table, scale)
    exp = [oracle_stage(in_fmt, w, h, f, L, out_pix_fmt=out_fmt, interp=interp, **kw) for f in frames]
    for options in (dict(), dict(wk_var=4000), dict(grid_ctas=3)):
        info = {}
        got = gpu_stage(in_fmt, w, h, frames, table, scale, out_pix_fmt=out_fmt, interp=interp, options=options, info=info, **kw)
        assert info["kernel"] == "k_any"
        for i in range(len(frames)):
            assert_bit_exact(out_fmt, w, h, got[i], exp[i], f"{in_fmt}->{out_fmt} {interp} {options} frame {i}")
