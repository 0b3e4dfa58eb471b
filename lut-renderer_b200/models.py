"""Operator parameters of the colour stage -- same public surface as the reference's
``lut_renderer.models`` (reference src/lut_renderer/models.py:11-143): ``TaskStatus``,
``ProcessingParams`` (field names, defaults, ``to_dict``/``from_dict`` JSON round trip) and ``Task``.

Kept API-compatible so presets / task lists written by the reference load unchanged.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from enum import Enum
from pathlib import Path
from typing import Any, Dict, Optional

from .media_info import VideoInfo


class TaskStatus(str, Enum):
    PENDING = "pending"
    RUNNING = "running"
    COMPLETED = "completed"
    FAILED = "failed"
    CANCELED = "canceled"


@dataclass
class ProcessingParams:
    # encoder / container options (passed through to the encode process untouched)
    video_codec: str = "libx264"
    audio_codec: str = "aac"
    pix_fmt: str = ""        # "" = let bit_depth_policy / the encoder decide
    resolution: str = ""     # "" = keep source
    bitrate: str = ""
    fps: str = ""            # "" = keep source
    crf: str = ""
    preset: str = ""
    tune: str = ""
    gop: str = ""
    profile: str = ""
    level: str = ""
    threads: str = ""
    audio_bitrate: str = ""
    sample_rate: str = ""
    channels: str = ""
    faststart: bool = False
    overwrite: bool = True
    generate_cover: bool = False
    processing_mode: str = "fast"        # "fast" | "pro" (ProRes 422 HQ master, then distribution encode)
    bit_depth_policy: str = "preserve"   # "preserve" | "auto" | "force_8bit"
    force_cfr: bool = True
    inherit_color_metadata: bool = True
    # --- colour-stage operator options (the ones the GPU stage consumes) ---
    lut_interp: str = "tetrahedral"      # lut3d interp: nearest | trilinear | tetrahedral (others fall back)
    zscale_dither: str = "none"
    lut_input_matrix: str = "auto"       # auto | bt709 | none | <matrix name>: YUV->RGB matrix in front of the LUT
    lut_output_tags: str = "bt709"       # bt709 | inherit | none: tags written on the output (and its matrix/range)

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "ProcessingParams":
        out = cls()
        for f in dataclasses.fields(cls):
            if f.name not in data:
                continue
            value = data[f.name]
            # booleans are coerced (JSON written by hand may hold 0/1); everything else is taken as stored
            setattr(out, f.name, bool(value) if isinstance(getattr(out, f.name), bool) else value)
        return out


@dataclass
class Task:
    task_id: str
    source_path: Path
    output_path: Path
    lut_path: Optional[Path]
    cover_path: Optional[Path]
    params: ProcessingParams
    source_info: Optional[VideoInfo] = None
    intermediate_path: Optional[Path] = None
    status: TaskStatus = TaskStatus.PENDING
    progress: int = 0
    error: str = ""
    started_at: Optional[float] = None
    finished_at: Optional[float] = None
    metadata: dict = field(default_factory=dict)

    def display_name(self) -> str:
        return self.source_path.name
