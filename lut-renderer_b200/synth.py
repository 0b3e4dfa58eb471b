"""Synthetic inputs for the parity harness and the benchmark (no ffmpeg / lavfi here, so `testsrc2` is
replaced by a deterministic procedural clip).  numpy only.

Frames are returned tightly packed (Y plane, U plane, V plane) as one uint8 array per frame, the layout the
raw pipe of the decode process delivers and ``lutr_stage_packed_planes`` describes.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_KR, _KB = 0.2126, 0.0722  # BT.709 luma weights (content generation only; not part of the measured arithmetic)


_SUBSAMPLING = {"yuv420p": (1, 1, 8), "yuv422p": (1, 0, 8), "yuv444p": (0, 0, 8),
                "yuv420p10le": (1, 1, 10), "yuv422p10le": (1, 0, 10), "yuv444p10le": (0, 0, 10)}  # log2 sx, log2 sy, depth


def plane_shapes(pix_fmt: str, w: int, h: int):
    if pix_fmt not in _SUBSAMPLING:
        raise ValueError(pix_fmt)
    sx, sy, depth = _SUBSAMPLING[pix_fmt]
    return [(h, w), (h >> sy, w >> sx), (h >> sy, w >> sx)], (np.dtype("<u2") if depth > 8 else np.uint8)


def split_planes(frame: np.ndarray, pix_fmt: str, w: int, h: int):
    """Views of the three planes of one packed frame."""
    shapes, dt = plane_shapes(pix_fmt, w, h)
    out, off = [], 0
    buf = frame.view(np.uint8).reshape(-1)
    for r, c in shapes:
        n = r * c * np.dtype(dt).itemsize
        out.append(buf[off:off + n].view(dt).reshape(r, c))
        off += n
    return out


def pack_planes(planes) -> np.ndarray:
    return np.concatenate([np.ascontiguousarray(p).view(np.uint8).reshape(-1) for p in planes])


def _rgb_scene(w: int, h: int, t: int, seed: int, grain: float = 0.004, chroma_grain: float = 0.0) -> np.ndarray:
    """float32 RGB in [0,1], (h, w, 3): colour bars, ramps, a moving gradient and disc, mild noise."""
    rng = np.random.default_rng(seed * 7919 + t)
    x = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :]
    y = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None]
    img = np.empty((h, w, 3), np.float32)
    # moving diagonal gradient background
    ph = t * 0.013
    img[..., 0] = 0.5 + 0.45 * np.sin(2 * np.pi * (x * 1.5 + y * 0.5 + ph))
    img[..., 1] = 0.5 + 0.45 * np.sin(2 * np.pi * (x * 0.7 - y * 1.1 + 1.3 * ph) + 1.0)
    img[..., 2] = 0.5 + 0.45 * np.sin(2 * np.pi * (-x * 0.9 + y * 1.7 + 0.7 * ph) + 2.0)
    # top: 8 colour bars (75 %) ; second strip: grey ramp ; third strip: saturated hue sweep
    bars = np.array([[1, 1, 1], [1, 1, 0], [0, 1, 1], [0, 1, 0], [1, 0, 1], [1, 0, 0], [0, 0, 1], [0, 0, 0]], np.float32) * 0.75
    h0, h1, h2 = h // 6, h // 4, h // 3
    idx = np.minimum((x[0] * 8).astype(int), 7)
    img[:h0] = bars[idx][None, :, :]
    img[h0:h1] = np.repeat(x.T, 3, axis=1)[None, :, :]
    hue = x[0] * 6.0
    sweep = np.stack([np.clip(np.abs(hue - 3) - 1, 0, 1), np.clip(2 - np.abs(hue - 2), 0, 1), np.clip(2 - np.abs(hue - 4), 0, 1)], -1)
    img[h1:h2] = sweep[None, :, :]
    # moving disc
    cx, cy = 0.5 + 0.3 * np.cos(0.05 * t), 0.65 + 0.2 * np.sin(0.07 * t)
    disc = ((x - cx) ** 2 * (w / h) ** 2 + (y - cy) ** 2) < 0.02
    img[disc] = np.array([0.9, 0.4, 0.1], np.float32)
    img += rng.normal(0.0, grain, size=(h, w, 1)).astype(np.float32)  # luma-ish grain
    if chroma_grain > 0.0:  # independent per channel: the kind of noise that spreads neighbouring pixels over LUT cells
        img += rng.normal(0.0, chroma_grain, size=(h, w, 3)).astype(np.float32)
    return np.clip(img, 0.0, 1.0)


def _rgb_to_yuv_planes(rgb: np.ndarray, pix_fmt: str, full_range: bool):
    h, w, _ = rgb.shape
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    yf = _KR * r + (1 - _KR - _KB) * g + _KB * b
    uf = (b - yf) / (2 * (1 - _KB))
    vf = (r - yf) / (2 * (1 - _KR))
    depth = 10 if "10" in pix_fmt else 8
    s = 1 << (depth - 8)
    if full_range:
        mx = (1 << depth) - 1
        Y, U, V = yf * mx, uf * mx + 128 * s, vf * mx + 128 * s
    else:
        Y, U, V = (16 + 219 * yf) * s, (128 + 224 * uf) * s, (128 + 224 * vf) * s
    sx, sy, _ = _SUBSAMPLING[pix_fmt]
    U = U.reshape(h >> sy, 1 << sy, w >> sx, 1 << sx).mean(axis=(1, 3))
    V = V.reshape(h >> sy, 1 << sy, w >> sx, 1 << sx).mean(axis=(1, 3))
    dt = np.dtype("<u2") if depth > 8 else np.uint8
    q = lambda a: np.clip(np.rint(a), 0, (1 << depth) - 1).astype(dt)  # noqa: E731
    return [q(Y), q(U), q(V)]


def natural_frame(pix_fmt: str, w: int, h: int, t: int = 0, seed: int = 0, full_range: bool = False, grain: float = 0.004,
                  chroma_grain: float = 0.0) -> np.ndarray:
    """Content class A ("natural", testsrc2-like).  One packed frame.  grain / chroma_grain: standard deviation (fraction of full
    scale) of the luma-like and of the per-channel noise added to the scene."""
    return pack_planes(_rgb_to_yuv_planes(_rgb_scene(w, h, t, seed, grain, chroma_grain), pix_fmt, full_range))


def noise_frame(pix_fmt: str, w: int, h: int, seed: int = 1) -> np.ndarray:
    """Content class B ("worst-case gather"): iid uniform over the full code range of every sample."""
    rng = np.random.default_rng(seed)
    shapes, dt = plane_shapes(pix_fmt, w, h)
    hi = 1024 if "10" in pix_fmt else 256
    return pack_planes([rng.integers(0, hi, size=s, dtype=np.uint16).astype(dt) for s in shapes])


def ramp_frame(pix_fmt: str, w: int, h: int) -> np.ndarray:
    """Content class C: horizontal luma ramp, vertical chroma ramps (hits every code value and the edges)."""
    shapes, dt = plane_shapes(pix_fmt, w, h)
    hi = 1023 if "10" in pix_fmt else 255
    Y = np.broadcast_to(np.linspace(0, hi, shapes[0][1]).round()[None, :], shapes[0]).astype(dt)
    U = np.broadcast_to(np.linspace(0, hi, shapes[1][0]).round()[:, None], shapes[1]).astype(dt)
    V = np.broadcast_to(np.linspace(hi, 0, shapes[2][1]).round()[None, :], shapes[2]).astype(dt)
    return pack_planes([Y, U, V])


# ----------------------------------------------------------------------------------------------------------------
# LUTs
# ----------------------------------------------------------------------------------------------------------------
def _grid(n: int):
    v = np.arange(n, dtype=np.float64) / (n - 1)
    return np.meshgrid(v, v, v, indexing="ij")  # [r][g][b]


def identity_lut(n: int) -> np.ndarray:
    r, g, b = _grid(n)
    return np.stack([r, g, b], -1)


def look_lut(n: int) -> np.ndarray:
    """Analytic 'look': gamma 0.9, saturation 1.2, S-curve; clipped to [0,1]."""
    r, g, b = _grid(n)
    rgb = np.stack([r, g, b], -1) ** 0.9
    luma = (rgb * np.array([_KR, 1 - _KR - _KB, _KB])).sum(-1, keepdims=True)
    rgb = luma + 1.2 * (rgb - luma)
    rgb = np.clip(rgb, 0.0, 1.0)
    rgb = rgb * rgb * (3.0 - 2.0 * rgb) * 0.6 + rgb * 0.4
    return np.clip(rgb, 0.0, 1.0)


def random_lut(n: int, seed: int = 2) -> np.ndarray:
    return np.random.default_rng(seed).random((n, n, n, 3))


def cube_text(table: np.ndarray, title: Optional[str] = "synthetic", domain: Optional[Tuple] = None,
              decimals: int = 6, crlf: bool = False, comments: bool = True) -> str:
    """Serialise a [r][g][b][3] table as Adobe/Resolve .cube text (red index fastest)."""
    n = table.shape[0]
    nl = "\r\n" if crlf else "\n"
    out = []
    if comments:
        out.append("# generated by lut_renderer_b200.synth")
    if title is not None:
        out.append(f'TITLE "{title}"')
    out.append(f"LUT_3D_SIZE {n}")
    if domain is not None:
        out.append("DOMAIN_MIN " + " ".join(f"{v:.6f}" for v in domain[0]))
        out.append("DOMAIN_MAX " + " ".join(f"{v:.6f}" for v in domain[1]))
    if comments:
        out.append("")
    flat = np.transpose(table, (2, 1, 0, 3)).reshape(-1, 3)  # blue-major iteration, red fastest
    fmt = f"%.{decimals}f %.{decimals}f %.{decimals}f"
    out.extend(fmt % (float(a), float(b), float(c)) for a, b, c in flat)
    return nl.join(out) + nl


def write_cube(path, table: np.ndarray, **kw) -> Path:
    p = Path(path)
    p.write_text(cube_text(table, **kw))
    return p
