"""Source description consumed by the colour stage -- API-compatible with the reference's
``lut_renderer.media_info`` (reference src/lut_renderer/media_info.py:12-226).

Only ``VideoInfo`` matters on the hot path: ``pix_fmt``, ``bit_depth``, ``colorspace``, ``color_range``,
``width``/``height`` select the stage descriptor.  ``probe_video`` needs an ``ffprobe`` binary (not part of
the hot path; kept so callers of the reference API keep working).
"""
from __future__ import annotations

import json
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

_VFR_TOLERANCE = 0.1


@dataclass
class VideoInfo:
    width: Optional[int] = None
    height: Optional[int] = None
    sar: Optional[str] = None
    dar: Optional[str] = None
    bitrate: Optional[str] = None
    container_bitrate: Optional[str] = None
    fps: Optional[float] = None
    avg_fps: Optional[float] = None
    r_fps: Optional[float] = None
    is_vfr: bool = False
    duration: Optional[float] = None
    pix_fmt: Optional[str] = None
    bit_depth: Optional[int] = None
    codec_name: Optional[str] = None
    codec_long_name: Optional[str] = None
    profile: Optional[str] = None
    level: Optional[str] = None
    color_primaries: Optional[str] = None
    color_trc: Optional[str] = None
    colorspace: Optional[str] = None
    color_range: Optional[str] = None
    format_name: Optional[str] = None
    format_long_name: Optional[str] = None
    file_size: Optional[int] = None
    audio_codec: Optional[str] = None
    audio_codec_long_name: Optional[str] = None
    audio_channels: Optional[int] = None
    audio_channel_layout: Optional[str] = None
    audio_sample_rate: Optional[int] = None
    audio_bitrate: Optional[str] = None
    format_tags: Optional[dict] = None
    video_tags: Optional[dict] = None
    audio_tags: Optional[dict] = None

    @property
    def resolution(self) -> Optional[str]:
        return f"{self.width}x{self.height}" if self.width and self.height else None


def _ratio(text: Optional[str]) -> Optional[float]:
    """'30000/1001' or '29.97' -> float; None for empty, '0/0' or unparsable input."""
    if not text or not text.strip() or text.strip() == "0/0":
        return None
    num, _, den = text.strip().partition("/")
    try:
        n = float(num)
        d = float(den) if den else 1.0
    except ValueError:
        return None
    return n / d if d else None


def _tag(value: Optional[str]) -> Optional[str]:
    if not value:
        return None
    v = str(value).strip()
    return None if v.lower() in {"", "unknown", "unspecified", "unknown/unknown"} else v


def _infer_bit_depth(pix_fmt: Optional[str], bits_per_raw_sample: Optional[str]) -> Optional[int]:
    """bits_per_raw_sample when ffprobe reports it, else the digits after the 'p' of the pix_fmt (yuv422p10le -> 10)."""
    if bits_per_raw_sample:
        try:
            bits = int(float(bits_per_raw_sample))
            if bits > 0:
                return bits
        except ValueError:
            pass
    for token in (pix_fmt or "").split(":"):
        head, sep, tail = token.partition("p")
        if not sep:
            continue
        digits = ""
        for ch in tail:
            if not ch.isdigit():
                break
            digits += ch
        if digits:
            return int(digits)
    return None


def _kbps(raw) -> Optional[str]:
    try:
        bits = int(float(raw)) if raw else 0
    except ValueError:
        return None
    return f"{max(1, round(bits / 1000))}k" if bits > 0 else None


_STREAM_KEYS = ("index,codec_type,codec_name,codec_long_name,profile,level,width,height,bit_rate,avg_frame_rate,"
                "r_frame_rate,pix_fmt,bits_per_raw_sample,color_primaries,color_transfer,color_space,color_range,"
                "duration,sample_aspect_ratio,display_aspect_ratio,channels,channel_layout,sample_rate")


def probe_video(path: Path) -> VideoInfo:
    cmd = ["ffprobe", "-v", "error", "-show_entries", f"stream={_STREAM_KEYS}", "-show_entries",
           "format=bit_rate,duration,size,format_name,format_long_name:format_tags", "-of", "json", str(path)]
    data = json.loads(subprocess.run(cmd, capture_output=True, text=True, check=True).stdout or "{}")
    streams = data.get("streams") or []
    v = next((s for s in streams if s.get("codec_type") == "video"), {})
    a = next((s for s in streams if s.get("codec_type") == "audio"), {})
    fmt = data.get("format") or {}
    avg, real = _ratio(v.get("avg_frame_rate")), _ratio(v.get("r_frame_rate"))
    pix_fmt = v.get("pix_fmt")
    color_range = _tag(v.get("color_range"))
    if not color_range and pix_fmt and str(pix_fmt).startswith("yuvj"):
        color_range = "pc"  # yuvj* = legacy spelling of full-range yuv*
    duration = None
    for raw in (v.get("duration"), fmt.get("duration")):
        try:
            if raw:
                duration = float(raw)
                break
        except ValueError:
            continue
    size = None
    try:
        size = int(float(fmt["size"])) if fmt.get("size") else None
    except ValueError:
        size = None
    return VideoInfo(
        width=v.get("width"), height=v.get("height"), sar=v.get("sample_aspect_ratio"), dar=v.get("display_aspect_ratio"),
        bitrate=_kbps(v.get("bit_rate") or fmt.get("bit_rate")), container_bitrate=_kbps(fmt.get("bit_rate")),
        fps=avg or real, avg_fps=avg, r_fps=real, is_vfr=bool(avg and real and abs(avg - real) > _VFR_TOLERANCE),
        duration=duration, pix_fmt=pix_fmt, bit_depth=_infer_bit_depth(pix_fmt, v.get("bits_per_raw_sample")),
        codec_name=v.get("codec_name"), codec_long_name=v.get("codec_long_name"), profile=v.get("profile"),
        level=str(v.get("level")) if v.get("level") is not None else None,
        color_primaries=_tag(v.get("color_primaries")), color_trc=_tag(v.get("color_trc") or v.get("color_transfer")),
        colorspace=_tag(v.get("colorspace") or v.get("color_space")), color_range=color_range,
        format_name=fmt.get("format_name"), format_long_name=fmt.get("format_long_name"), file_size=size,
        audio_codec=a.get("codec_name"), audio_codec_long_name=a.get("codec_long_name"), audio_channels=a.get("channels"),
        audio_channel_layout=a.get("channel_layout"),
        audio_sample_rate=int(a["sample_rate"]) if a.get("sample_rate") else None, audio_bitrate=_kbps(a.get("bit_rate")),
        format_tags=fmt.get("tags"), video_tags=v.get("tags"), audio_tags=a.get("tags"),
    )
