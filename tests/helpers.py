"""Shared test plumbing: run one configuration through the ORACLE (checker) and through the C-ABI (product)."""
from __future__ import annotations

import numpy as np

from lut_renderer_b200 import synth


def oracle_stage(pix_fmt, w, h, frame, lut, interp="tetrahedral", in_matrix="bt709", out_matrix="bt709",
                 in_range="tv", norm_range=None, out_pix_fmt=None, out_range="tv"):
    """Packed frame in -> packed frame out through oracle/ (CPU restatement of the FFmpeg chain)."""
    from oracle import oracle as O
    norm_range = norm_range or in_range
    Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(frame, pix_fmt, w, h)]
    out = O.stage_any(lut, pix_fmt, out_pix_fmt or pix_fmt, Y, U, V, interp=interp, range_pass=(in_range == "pc" and norm_range == "tv"),
                      in_matrix=in_matrix, in_full_range=(norm_range == "pc"), out_matrix=out_matrix, out_full_range=(out_range == "pc"))
    return synth.pack_planes(out)


def oracle_lut(table, scale=(1.0, 1.0, 1.0)):
    from oracle import oracle as O
    return O.OracleLut.from_table(table, scale)


def gpu_stage(pix_fmt, w, h, frames, table, scale=(1.0, 1.0, 1.0), device=0, options=None, info=None, **stage_kw):
    """List of packed frames -> list of packed frames through liblutr.so on the GPU (device-resident path)."""
    import torch
    from lut_renderer_b200.stage import ColourStage
    st = ColourStage(w, h, pix_fmt, device=device, **stage_kw)
    for k, v in (options or {}).items():
        st.set_option(k, v)
    lut = st.ctx.lut_from_table(table, scale)
    n = len(frames)
    src = torch.from_numpy(np.concatenate(frames)).to(f"cuda:{device}")
    dst = torch.zeros(n * st.out_frame_bytes, dtype=torch.uint8, device=f"cuda:{device}")
    st.run(lut, src, dst, n, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    out = dst.cpu().numpy()
    if info is not None:
        info["kernel"] = st.kernel_name
    st.close()
    lut.close()
    return [out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes] for i in range(n)]


def plane_stats(pix_fmt, w, h, got, exp):
    """Per plane: (identical fraction, max |d|, mean |d|) -- the statistics north_star asks to state."""
    res = {}
    for name, a, b in zip("YUV", synth.split_planes(got, pix_fmt, w, h), synth.split_planes(exp, pix_fmt, w, h)):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        res[name] = (float((d == 0).mean()), int(d.max()), float(d.mean()))
    return res


def assert_bit_exact(pix_fmt, w, h, got, exp, what=""):
    st = plane_stats(pix_fmt, w, h, got, exp)
    bad = {k: v for k, v in st.items() if v[1] != 0}
    if bad:
        lines = []
        for name, a, b in zip("YUV", synth.split_planes(got, pix_fmt, w, h), synth.split_planes(exp, pix_fmt, w, h)):
            idx = np.argwhere(a != b)[:6]
            lines.append(f"{name}: {[(tuple(i), int(a[tuple(i)]), int(b[tuple(i)])) for i in idx]}")
        raise AssertionError(f"{what}: not bit-exact {bad}\n" + "\n".join(lines))
