"""Operator / command builder of the colour stage.

Keeps the reference's operator surface (reference src/lut_renderer/ffmpeg.py): ``CommandStage``,
``build_command(source, output, params, lut_path, ffmpeg_bin, source_info, notes)`` and
``build_pipeline(task, ffmpeg_bin)`` return exactly what the reference returns, so existing callers and the
reference's own smoke test keep working.  On top of that this module holds the ONE place where the operator
options are resolved into the descriptor of the fused GPU stage:

* ``plan_colour_stage()``  -- ProcessingParams + VideoInfo -> ``ColourStagePlan`` (formats, ranges, matrices, interp)
* ``build_gpu_stage()``    -- a ``CommandStage`` -> decode argv / ``ColourStagePlan`` / encode argv, i.e. the
  ``decode -> raw pipe -> GPU -> raw pipe -> encode`` replacement for the single ffmpeg process with ``-vf ...lut3d...``.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Tuple

from .media_info import VideoInfo
from .models import ProcessingParams, Task

LUT3D_INTERPS = ("nearest", "trilinear", "tetrahedral", "pyramid", "prism", "cubic")
GPU_INTERPS = ("nearest", "trilinear", "tetrahedral", "pyramid", "prism")
_TEN_BIT_ENCODERS = frozenset({"prores_ks", "libx265", "hevc_videotoolbox"})
_FORCEABLE_MATRICES = ("bt709", "smpte170m", "bt470bg", "bt2020nc", "bt2020c")
_BT709_TAGS = ["-color_primaries", "bt709", "-color_trc", "bt709", "-colorspace", "bt709", "-color_range", "tv"]
_BITRATE = re.compile(r"^\s*(\d+(?:\.\d+)?)([kKmMgG]?)\s*$")


@dataclass
class CommandStage:
    name: str
    source_path: Path
    output_path: Path
    params: ProcessingParams
    lut_path: Optional[Path] = None
    cleanup_on_success: bool = False
    notes: List[str] = field(default_factory=list)
    # probe the stage input right before building its command (pro mode: stage 2 reads the intermediate)
    probe_source: bool = False


# ----------------------------------------------------------------------------------------------------------------
# small parsers / formatters
# ----------------------------------------------------------------------------------------------------------------
def _escape_filter_path(path: Path) -> str:
    """Quote a path for use inside '...' in a filtergraph: backslashes first, then single quotes."""
    return str(path).replace("\\", "\\\\").replace("'", "\\'")


def _format_float(value: float) -> str:
    return f"{value:.3f}".rstrip("0").rstrip(".")


def _parse_fraction(value: str) -> Optional[float]:
    text = (value or "").strip()
    if not text:
        return None
    num, slash, den = text.partition("/")
    try:
        n = float(num)
        d = float(den) if slash else 1.0
    except ValueError:
        return None
    return n / d if d else None


def _parse_bitrate(value: str) -> Optional[Tuple[float, str]]:
    m = _BITRATE.match(value) if value else None
    if not m or float(m.group(1)) <= 0:
        return None
    return float(m.group(1)), m.group(2) or ""


def _scale_bitrate(value: str, factor: float) -> Optional[str]:
    parsed = _parse_bitrate(value)
    if not parsed:
        return None
    number, unit = parsed[0] * factor, parsed[1]
    if abs(number - round(number)) < 1e-6:
        return f"{int(round(number))}{unit}"
    return f"{number:g}{unit}"


def _bitrate_to_kbps(value: Optional[str]) -> Optional[float]:
    parsed = _parse_bitrate(value) if value else None
    if not parsed:
        return None
    number, unit = parsed
    return {"k": number, "m": number * 1e3, "g": number * 1e6}.get(unit.lower())


def _supports_10bit(codec: str) -> bool:
    return codec in _TEN_BIT_ENCODERS


def _normalize_scale_matrix(value: Optional[str]) -> Optional[str]:
    text = str(value).strip().lower() if value else ""
    return text if text in _FORCEABLE_MATRICES else None


def _needs_full_range_normalization(info: Optional[VideoInfo]) -> bool:
    if not info:
        return False
    if info.pix_fmt and str(info.pix_fmt).startswith("yuvj"):
        return True
    return bool(info.color_range) and str(info.color_range).lower() == "pc"


def _full_range_intermediate_pix_fmt(info: Optional[VideoInfo]) -> str:
    pix_fmt = str(info.pix_fmt) if info and info.pix_fmt else ""
    for sub in ("444", "422"):
        if sub in pix_fmt:
            return f"yuv{sub}p"
    return "yuv420p"


def _resolve_fps(params: ProcessingParams, info: Optional[VideoInfo]) -> Tuple[Optional[float], Optional[str]]:
    if params.fps:
        return _parse_fraction(params.fps), params.fps
    if info and info.fps:
        return info.fps, _format_float(info.fps)
    return None, None


# ----------------------------------------------------------------------------------------------------------------
# option resolution shared by the argv builder and the GPU planner
# ----------------------------------------------------------------------------------------------------------------
def _policy(params: ProcessingParams, name: str, default: str) -> str:
    return (getattr(params, name, "") or default).strip().lower()


def _forced_matrix(params: ProcessingParams, info: Optional[VideoInfo]) -> Optional[str]:
    """LUT-input matrix the operator forces on the scaler (None = leave it to FFmpeg)."""
    policy = _policy(params, "lut_input_matrix", "auto")
    if policy == "bt709":
        return "bt709"
    if policy == "auto":
        return _normalize_scale_matrix(info.colorspace if info else None)
    if policy == "none":
        return None
    return _normalize_scale_matrix(policy)


def _normalised_range(params: ProcessingParams, info: Optional[VideoInfo]) -> str:
    """out_range of the pc-normalisation pass: tv for bt709 tags, the source's own range for inherit, pc for none."""
    tags = _policy(params, "lut_output_tags", "bt709")
    if tags == "bt709":
        return "tv"
    if tags == "inherit":
        return str(info.color_range).lower().strip() if info and info.color_range else "pc"
    return "pc"


def _resolve_interp(params: ProcessingParams) -> str:
    interp = params.lut_interp or "tetrahedral"
    return interp if interp in LUT3D_INTERPS else "tetrahedral"


def _resolve_pix_fmt(params: ProcessingParams, info: Optional[VideoInfo], notes: List[str]) -> str:
    """Output pixel format after the bit-depth policy ('' = nothing forced, the encoder negotiates)."""
    pix_fmt = params.pix_fmt
    if params.bit_depth_policy == "force_8bit":
        if pix_fmt != "yuv420p":
            notes.append("位深策略=强制8bit: pix_fmt=yuv420p")
        return "yuv420p"
    if params.bit_depth_policy in {"preserve", "auto"} and not pix_fmt:
        if info and info.bit_depth and info.bit_depth >= 10:
            if _supports_10bit(params.video_codec):
                pix_fmt = "yuv422p10le" if params.video_codec == "prores_ks" else "yuv420p10le"
                notes.append(f"位深策略=保持10bit: pix_fmt={pix_fmt}")
            else:
                pix_fmt = "yuv420p"
                notes.append("位深策略=保持10bit: 编码器不支持10bit，回退 yuv420p")
    return pix_fmt


def _colour_filters(params: ProcessingParams, lut_path: Path, info: Optional[VideoInfo], notes: List[str]) -> List[str]:
    """The `scale,format,lut3d` part of the -vf chain the reference emits for an active LUT."""
    chain: List[str] = []
    matrix = _forced_matrix(params, info)
    matrix_policy = _policy(params, "lut_input_matrix", "auto")
    if _needs_full_range_normalization(info):
        out_range = _normalised_range(params, info)
        intermediate = _full_range_intermediate_pix_fmt(info)
        parts = ["in_range=pc", f"out_range={out_range}"]
        notes.append(f"Range: 检测到 full-range(pc)，已按 out_range={out_range} 规范化，避免 yuvj* 旧像素格式（format={intermediate}）")
        if matrix:
            parts += [f"in_color_matrix={matrix}", f"out_color_matrix={matrix}"]
            notes.append(f"LUT 输入矩阵: {matrix}（{matrix_policy}）")
        chain += ["scale=" + ":".join(parts), f"format={intermediate}"]
    elif matrix:
        chain.append(f"scale=in_color_matrix={matrix}:out_color_matrix={matrix}")
        notes.append(f"LUT 输入矩阵: {matrix}（{matrix_policy}）")
    else:
        notes.append("LUT 输入矩阵: 未强制（auto/none 或无法识别源 colorspace）")
    interp = _resolve_interp(params)
    chain.append(f"lut3d=file='{_escape_filter_path(lut_path)}':interp={interp}")
    notes.append(f"LUT: 使用 lut3d（interp={interp}）")
    return chain


def _source_tags(cmd: List[str], info: Optional[VideoInfo], notes: List[str]) -> None:
    if not info:
        return
    shown = []
    for flag, label, value in (("-color_primaries", "primaries", info.color_primaries), ("-color_trc", "trc", info.color_trc),
                               ("-colorspace", "colorspace", info.colorspace), ("-color_range", "range", info.color_range)):
        if value:
            cmd += [flag, value]
            shown.append(f"{label}={value}")
    if shown:
        notes.append(f"继承色彩元数据: {', '.join(shown)}")


def _timing_args(cmd: List[str], params: ProcessingParams, info: Optional[VideoInfo], notes: List[str]) -> Optional[float]:
    fps_value, fps_text = _resolve_fps(params, info)
    if params.fps:
        cmd += ["-fps_mode", "cfr", "-r", params.fps]
        notes.append(f"时间结构: fps_mode=cfr, 输出帧率={params.fps}")
        return fps_value
    vfr = bool(info and info.is_vfr)
    if vfr and params.force_cfr:
        cmd += ["-fps_mode", "cfr"]
        if fps_text:
            cmd += ["-r", fps_text]
            notes.append(f"时间结构: 源为 VFR，已强制 CFR，输出帧率={fps_text}")
        else:
            notes.append("时间结构: 源为 VFR，已强制 CFR（未检测到帧率）")
    elif params.force_cfr and info is None:
        cmd += ["-fps_mode", "cfr"]
        notes.append("时间结构: fps_mode=cfr（未读取源信息）")
    else:
        cmd += ["-fps_mode", "passthrough"]
        notes.append("时间结构: 源为 VFR，fps_mode=passthrough（不重写时间戳）" if vfr
                     else "时间结构: 源为 CFR/未知，fps_mode=passthrough（避免时间戳重写）")
    return fps_value


def _rate_and_codec_args(cmd: List[str], params: ProcessingParams, fps_value: Optional[float], notes: List[str]) -> None:
    if params.resolution:
        cmd += ["-s", params.resolution]
    if params.bitrate:
        cmd += ["-b:v", params.bitrate]
        bufsize = _scale_bitrate(params.bitrate, 2)
        if bufsize:
            cmd += ["-maxrate", params.bitrate, "-bufsize", bufsize]
            notes.append(f"码率稳定: maxrate={params.bitrate}, bufsize={bufsize}")
    for flag, value in (("-crf", params.crf), ("-preset", params.preset), ("-tune", params.tune)):
        if value:
            cmd += [flag, value]
    if params.gop:
        cmd += ["-g", params.gop]
    elif fps_value:
        gop = max(1, round(fps_value))
        cmd += ["-g", str(gop)]
        notes.append(f"自动 GOP={gop} (fps={_format_float(fps_value)})")
    for flag, value in (("-profile:v", params.profile), ("-level", params.level), ("-threads", params.threads)):
        if value:
            cmd += [flag, value]


def _output_tag_args(cmd: List[str], params: ProcessingParams, info: Optional[VideoInfo], lut_active: bool,
                     notes: List[str]) -> None:
    if not lut_active:
        if params.inherit_color_metadata:
            _source_tags(cmd, info, notes)
        return
    policy = _policy(params, "lut_output_tags", "bt709")
    if policy == "inherit":
        if params.inherit_color_metadata:
            _source_tags(cmd, info, notes)
    elif policy == "none":
        notes.append("LUT 输出标记: none（不写色彩元数据）")
    else:
        cmd += _BT709_TAGS
        notes.append("LUT 输出标记: bt709/bt709/bt709, range=tv" + ("" if policy == "bt709" else "（回退）"))


def _audio_and_mux_args(cmd: List[str], params: ProcessingParams) -> None:
    if params.audio_codec and params.audio_codec != "copy":
        for flag, value in (("-b:a", params.audio_bitrate), ("-ar", params.sample_rate), ("-ac", params.channels)):
            if value:
                cmd += [flag, value]
    if params.faststart:
        cmd += ["-movflags", "+faststart"]


def _videotoolbox_hint(params: ProcessingParams, info: Optional[VideoInfo], notes: List[str]) -> None:
    if params.video_codec and "videotoolbox" in params.video_codec:
        kbps = _bitrate_to_kbps(params.bitrate or (info.bitrate if info else ""))
        if kbps and kbps >= 50_000:
            notes.append("提示: h264_videotoolbox 在高码率/重负载时可能出现 PTS 重建/帧重排的节奏错觉；"
                         "如需更稳定建议用 libx264 或切到“专业母带”。")


# ----------------------------------------------------------------------------------------------------------------
# reference-compatible argv builder
# ----------------------------------------------------------------------------------------------------------------
def build_command(
    source: Path,
    output: Path,
    params: ProcessingParams,
    lut_path: Optional[Path] = None,
    ffmpeg_bin: str = "ffmpeg",
    source_info: Optional[VideoInfo] = None,
    notes: Optional[List[str]] = None,
) -> List[str]:
    """The single-process ffmpeg command of the reference (CPU `lut3d` path), argument for argument."""
    notes = notes if notes is not None else []
    cmd = [ffmpeg_bin, "-hide_banner"]
    if params.overwrite:
        cmd.append("-y")
    cmd += ["-i", str(source)]

    filters = _colour_filters(params, lut_path, source_info, notes) if lut_path else []
    if params.video_codec:
        cmd += ["-c:v", params.video_codec]
    if params.audio_codec:
        cmd += ["-c:a", params.audio_codec]
    if filters and params.video_codec == "copy":
        raise ValueError("启用 LUT/滤镜时不能使用视频 copy（streamcopy 与滤镜不可同时使用）。")

    if params.video_codec and params.video_codec != "copy":
        fps_value = _timing_args(cmd, params, source_info, notes)
        pix_fmt = _resolve_pix_fmt(params, source_info, notes)
        if pix_fmt:
            if getattr(params, "zscale_dither", "none") == "error_diffusion":
                filters.append("zscale=dither=error_diffusion")
                notes.append("抖动: zscale=dither=error_diffusion")
            if lut_path:
                filters.append(f"format={pix_fmt}")
            cmd += ["-pix_fmt", pix_fmt]
        _rate_and_codec_args(cmd, params, fps_value, notes)
        _output_tag_args(cmd, params, source_info, bool(lut_path), notes)
        _videotoolbox_hint(params, source_info, notes)

    if filters:
        cmd += ["-vf", ",".join(filters)]
    _audio_and_mux_args(cmd, params)
    cmd.append(str(output))
    return cmd


def _build_master_params(params: ProcessingParams) -> ProcessingParams:
    """Pro mode stage 1: ProRes 422 HQ, yuv422p10le, audio copied, every rate-control knob cleared."""
    master = ProcessingParams.from_dict(params.to_dict())
    master.video_codec, master.audio_codec = "prores_ks", "copy"
    master.pix_fmt, master.profile, master.bit_depth_policy = "yuv422p10le", "3", "preserve"
    for name in ("level", "crf", "preset", "tune", "bitrate", "audio_bitrate", "sample_rate", "channels"):
        setattr(master, name, "")
    master.faststart = False
    return master


def build_pipeline(task: Task, ffmpeg_bin: str = "ffmpeg") -> List[CommandStage]:
    params = task.params
    if params.processing_mode != "pro":
        return [CommandStage(name="快速交付", source_path=task.source_path, output_path=task.output_path, params=params,
                             lut_path=task.lut_path, notes=[])]
    if not task.intermediate_path:
        raise ValueError("专业母带模式需要显式设置中间文件路径（请在界面中设置母带缓存目录）。")
    return [
        CommandStage(name="ProRes 母带", source_path=task.source_path, output_path=task.intermediate_path,
                     params=_build_master_params(params), lut_path=task.lut_path, cleanup_on_success=True,
                     notes=["母带固定为 ProRes 422 HQ (yuv422p10le)"]),
        CommandStage(name="分发编码", source_path=task.intermediate_path, output_path=task.output_path, params=params,
                     lut_path=None, notes=[], probe_source=True),
    ]


# ----------------------------------------------------------------------------------------------------------------
# GPU stage planning: the same options resolved into a descriptor instead of a filter string
# ----------------------------------------------------------------------------------------------------------------
# colourspace names (ffprobe / -colorspace vocabulary) -> matrix family of the stage
_MATRIX_FAMILY = {"bt709": "bt709", "smpte170m": "bt601", "bt470bg": "bt601", "bt601": "bt601", "bt2020nc": "bt2020",
                  "bt2020c": "bt2020", "bt2020": "bt2020", "fcc": "fcc", "smpte240m": "smpte240m"}


@dataclass
class ColourStagePlan:
    """Everything the fused kernel needs to stand in for `scale,format,lut3d,format` on one clip."""
    width: int
    height: int
    in_pix_fmt: str    # what the decode process must deliver on the raw pipe
    out_pix_fmt: str   # what the encode process receives
    in_range: str      # "tv" | "pc": range of the decoded samples
    norm_range: str    # range after the reference's pc-normalisation pass (== in_range for tv sources)
    in_matrix: str     # matrix family for YUV->RGB in front of the LUT
    out_matrix: str    # matrix family for RGB->YUV behind it
    out_range: str
    interp: str
    lut_path: Optional[Path]
    rgb_intermediate: str            # "rgb24" (8-bit sources) | "gbrp10le" (10-bit sources): what FFmpeg negotiates for lut3d
    via_8bit: bool = False   # full-range source deeper than 8 bit: the stage re-quantises it to the reference's 8-bit intermediate
    reference_filters: List[str] = field(default_factory=list)  # the -vf chain this plan replaces
    notes: List[str] = field(default_factory=list)


def _effective_matrix(forced: Optional[str], tagged: Optional[str]) -> str:
    """Matrix FFmpeg >= 7 ends up using for a scaler: the forced one, else the frame's tagged colourspace,
    else swscale's default (BT.601 table)."""
    for name in (forced, (tagged or "").strip().lower() or None):
        if name and name in _MATRIX_FAMILY:
            return _MATRIX_FAMILY[name]
    return "bt601"


def plan_colour_stage(params: ProcessingParams, lut_path: Optional[Path], source_info: Optional[VideoInfo],
                      notes: Optional[List[str]] = None) -> Optional[ColourStagePlan]:
    """Resolve the operator options exactly like ``build_command`` does, but into a stage descriptor.

    Returns None when no LUT is active (nothing for the GPU stage to do).  Raises ValueError for the same
    misuse ``build_command`` rejects (video copy + LUT) and when the source geometry is unknown.
    """
    if not lut_path:
        return None
    if params.video_codec == "copy":
        raise ValueError("启用 LUT/滤镜时不能使用视频 copy（streamcopy 与滤镜不可同时使用）。")
    info = source_info
    if not info or not info.width or not info.height:
        raise ValueError("colour stage needs the source geometry (VideoInfo.width/height)")
    plan_notes: List[str] = notes if notes is not None else []
    reference_filters = _colour_filters(params, lut_path, info, plan_notes)

    src_fmt = str(info.pix_fmt or "yuv420p")
    src_depth = info.bit_depth or (10 if "10" in src_fmt else 8)
    full_range = _needs_full_range_normalization(info)
    via_8bit = False
    if full_range:
        # the reference re-quantises to 8-bit yuv4xxp behind the range pass whatever the source depth (ffmpeg.py:137-143,233)
        in_pix_fmt, work_depth = _full_range_intermediate_pix_fmt(info), 8
        deep = in_pix_fmt + "10le"
        if src_depth == 10 and src_fmt.replace("yuvj", "yuv") == deep:
            # hand the stage the 10-bit samples: it performs swscale's one-pass conversion to that 8-bit intermediate itself
            in_pix_fmt, via_8bit = deep, True
        in_range, norm_range = "pc", _normalised_range(params, info)
        if norm_range not in ("tv", "pc"):
            norm_range = "pc"
    else:
        in_pix_fmt, work_depth = src_fmt.replace("yuvj", "yuv"), src_depth
        in_range = norm_range = "tv"

    quiet: List[str] = []
    out_pix_fmt = _resolve_pix_fmt(params, info, quiet)
    if out_pix_fmt:
        if getattr(params, "zscale_dither", "none") == "error_diffusion":
            reference_filters.append("zscale=dither=error_diffusion")
        reference_filters.append(f"format={out_pix_fmt}")
    else:
        # nothing forced: the reference leaves the choice to the encoder, which sees lut3d's RGB output
        out_pix_fmt = _NEGOTIATED_PIX_FMT.get((params.video_codec, work_depth))
        if out_pix_fmt:
            plan_notes.append(f"pix_fmt 未指定: 按编码器协商结果输出 {out_pix_fmt}（lut3d 输出 RGB，{params.video_codec} 选择无色度损失的格式）")
        else:
            out_pix_fmt = in_pix_fmt
            plan_notes.append(f"pix_fmt 未指定: GPU 阶段按源格式输出 {out_pix_fmt}（与参考实现的编码器协商结果可能不同）")

    tags = _policy(params, "lut_output_tags", "bt709")
    if tags == "inherit" and params.inherit_color_metadata:
        out_matrix = _effective_matrix(None, info.colorspace)
        out_range = "pc" if str(info.color_range or "").lower() == "pc" else "tv"
    elif tags == "none" or tags == "inherit":
        out_matrix, out_range = "bt601", "tv"   # untagged output: swscale's default matrix, limited range
    else:
        out_matrix, out_range = "bt709", "tv"

    return ColourStagePlan(
        width=int(info.width), height=int(info.height), in_pix_fmt=in_pix_fmt, out_pix_fmt=out_pix_fmt,
        in_range=in_range, norm_range=norm_range, in_matrix=_effective_matrix(_forced_matrix(params, info), info.colorspace),
        out_matrix=out_matrix, out_range=out_range, interp=_resolve_interp(params), lut_path=Path(lut_path),
        rgb_intermediate="gbrp10le" if work_depth >= 10 else "rgb24", via_8bit=via_8bit, reference_filters=reference_filters,
        notes=plan_notes)


def plan_conversion_stage(params: ProcessingParams, source_info: Optional[VideoInfo],
                          notes: Optional[List[str]] = None) -> Optional[ColourStagePlan]:
    """Pro mode stage 2 (reference ffmpeg.py:459-471): no LUT, but the encoding ffmpeg still converts the yuv422p10le master to the
    delivery pix_fmt with an auto-inserted swscale pass.  Returns the plan of that conversion for the GPU stage (``interp="none"``)
    when it is one the stage implements -- yuv422p10le, limited range, to yuv420p / yuv420p10le / yuv422p, width a multiple of 8 --
    else None (nothing to convert, or a conversion that stays inside ffmpeg)."""
    info = source_info
    if not info or not info.width or not info.height or str(info.pix_fmt or "") != "yuv422p10le":
        return None
    if str(info.color_range or "tv").lower() == "pc" or params.video_codec in ("", "copy"):
        return None
    quiet: List[str] = []
    out_pix_fmt = _resolve_pix_fmt(params, info, quiet)
    if out_pix_fmt not in ("yuv420p", "yuv420p10le", "yuv422p") or int(info.width) % 8 or int(info.height) % 2 or int(info.height) < 14:
        return None
    if notes is not None:
        notes.append(f"分发编码: 像素格式转换 yuv422p10le -> {out_pix_fmt} 在 GPU 上完成（无 LUT，swscale 等价算法）")
    matrix = _effective_matrix(None, info.colorspace)
    return ColourStagePlan(width=int(info.width), height=int(info.height), in_pix_fmt="yuv422p10le", out_pix_fmt=out_pix_fmt,
                           in_range="tv", norm_range="tv", in_matrix=matrix, out_matrix=matrix, out_range="tv", interp="none",
                           lut_path=None, rgb_intermediate="none", reference_filters=[f"format={out_pix_fmt}"],
                           notes=notes if notes is not None else [])


@dataclass
class GpuCommandStage:
    """decode -> raw pipe -> GPU colour stage -> raw pipe -> encode, for one ``CommandStage`` with a LUT."""
    stage: CommandStage
    plan: ColourStagePlan
    decode_cmd: List[str]
    encode_cmd: List[str]


class RawPipeUnsupported(ValueError):
    """The clip cannot go through the raw-video pipe without changing its timing; the caller runs the reference's
    single-process command for this stage instead (and says so in the task log)."""


# what FFmpeg's format negotiation picks when nothing is forced and lut3d hands RGB to the encoder (the
# reference emits neither format= nor -pix_fmt for 8-bit sources): the encoder's best match for rgb24 / gbrp10le.  Only the
# encoders whose choice is a planar YUV layout of the stage are listed; others keep the source layout (logged as a deviation).
_NEGOTIATED_PIX_FMT = {("libx264", 8): "yuv444p", ("libx264", 10): "yuv444p10le"}


def _pipe_rate_text(fps: float) -> str:
    """A frame rate as ffmpeg's -r takes it, exact for the NTSC family (VideoInfo keeps only the float)."""
    n = round(fps * 1001)
    if n % 1000 == 0 and abs(n / 1001 - fps) < 2e-3 and abs(fps - round(fps)) > 1e-3:
        return f"{n}/1001"
    if abs(fps - round(fps)) < 1e-3:
        return str(int(round(fps)))
    return f"{fps:.6f}".rstrip("0").rstrip(".")


def _pipe_timing(params: ProcessingParams, info: Optional[VideoInfo]) -> Tuple[List[str], str]:
    """(timing options of the decoder, frame rate declared on the encoder's rawvideo input).

    rawvideo carries no timestamps, so the pipe must run at a constant rate that reproduces the SOURCE timing; the conversion to a
    requested output rate stays where the reference has it, on the output side (`-fps_mode cfr -r X`, reference ffmpeg.py:262-286).
    Variable-frame-rate sources are made constant in the decoder when the reference would force CFR anyway; a VFR source in
    passthrough mode, or a source of unknown rate, cannot cross the pipe."""
    vfr = bool(info and info.is_vfr)
    src_fps = _pipe_rate_text(info.fps) if info and info.fps else None
    if vfr:
        if params.fps:       # resample to the requested constant rate in the decoder (same frame selection as cfr on the output)
            return ["-fps_mode", "cfr", "-r", params.fps], params.fps
        if params.force_cfr and src_fps:
            return ["-fps_mode", "cfr", "-r", src_fps], src_fps
        raise RawPipeUnsupported("源为 VFR 且保持时间戳（passthrough）：原始视频管道无法携带时间戳")
    if not src_fps:
        raise RawPipeUnsupported("未知源帧率：原始视频管道需要恒定帧率")
    return ["-fps_mode", "passthrough"], src_fps


def build_gpu_stage(stage: CommandStage, source_info: Optional[VideoInfo], ffmpeg_bin: str = "ffmpeg",
                    convert_without_lut: bool = False) -> Optional[GpuCommandStage]:
    """Split a LUT-carrying stage into a raw-video decoder and a raw-video encoder around the GPU stage.

    The encoder command carries every option ``build_command`` would have put on the single process
    (codec, rate control, timing, tags); only the `-vf` colour chain is gone and its input is the raw pipe.
    Audio and metadata are taken from the original source as a second input.  Raises ``RawPipeUnsupported`` when the
    source timing cannot be reproduced over a raw pipe (see ``_pipe_timing``).  ``convert_without_lut``: also split a stage WITHOUT a
    LUT (pro mode stage 2) when its pixel-format conversion is one the GPU stage implements (``plan_conversion_stage``).
    """
    if stage.lut_path:
        plan = plan_colour_stage(stage.params, stage.lut_path, source_info, stage.notes)
    else:
        plan = plan_conversion_stage(stage.params, source_info, stage.notes) if convert_without_lut else None
    if plan is None:
        return None
    params = stage.params
    decode_timing, pipe_fps = _pipe_timing(params, source_info)
    decode = [ffmpeg_bin, "-hide_banner", "-v", "error", "-i", str(stage.source_path), "-map", "0:v:0", "-f", "rawvideo",
              # untouched samples: 8-bit full-range sources are named yuvj4xxp; deeper ones keep their own name (no yuvj form exists)
              "-pix_fmt", ("yuvj" + plan.in_pix_fmt[3:]) if plan.in_range == "pc" and plan.in_pix_fmt in ("yuv420p", "yuv422p", "yuv444p")
              else plan.in_pix_fmt] + decode_timing + ["pipe:1"]
    encode = [ffmpeg_bin, "-hide_banner"]
    if params.overwrite:
        encode.append("-y")
    encode += ["-f", "rawvideo", "-pix_fmt", plan.out_pix_fmt, "-s", f"{plan.width}x{plan.height}", "-r", pipe_fps]
    encode += ["-i", "pipe:0", "-i", str(stage.source_path), "-map", "0:v:0", "-map", "1:a?", "-map_metadata", "1"]
    if params.video_codec:
        encode += ["-c:v", params.video_codec]
    if params.audio_codec:
        encode += ["-c:a", params.audio_codec]
    quiet: List[str] = []
    fps_value = _timing_args(encode, params, source_info, quiet)
    encode += ["-pix_fmt", plan.out_pix_fmt]
    _rate_and_codec_args(encode, params, fps_value, quiet)
    _output_tag_args(encode, params, source_info, bool(stage.lut_path), quiet)
    _audio_and_mux_args(encode, params)
    encode.append(str(stage.output_path))
    return GpuCommandStage(stage=stage, plan=plan, decode_cmd=decode, encode_cmd=encode)


# ---- one clip over several GPUs: contiguous, GOP-aligned frame ranges (BASELINE.json north_star item 4) -----------------------------
@dataclass
class ShardCommand:
    """One frame range of a LUT-carrying stage: its own decoder and encoder process around the GPU stage of ``device``."""
    index: int
    device: int
    start_frame: int
    frames: int
    decode_cmd: List[str]
    encode_cmd: List[str]
    part_path: Path
    open_ended: bool = False   # the last range: no -frames:v, it may deliver more frames than planned (never fewer)


@dataclass
class ShardedGpuStage:
    base: GpuCommandStage
    shards: List[ShardCommand]
    concat_list_path: Path
    concat_list_text: str
    concat_cmd: List[str]      # stitches the parts (stream copy) and brings audio / metadata over from the source


def build_sharded_gpu_stage(stage: CommandStage, source_info: Optional[VideoInfo], devices: List[int],
                            ffmpeg_bin: str = "ffmpeg") -> Optional[ShardedGpuStage]:
    """Split a LUT-carrying stage into one (decode, GPU, encode) chain per device over contiguous frame ranges.

    Frames are independent in every filter of the reference's chain, so the ranges need no exchange: each device uploads the LUT
    once and works through its range.  Range boundaries fall on multiples of the encoder's GOP length (every part starts on a
    keyframe; the GOP is the one ``build_command`` would set: ``-g``, else round(fps)), parts are stitched with the concat demuxer
    and stream copy, audio and metadata come from the source in that last step.  Returns None when the clip cannot be cut
    reliably -- unknown duration / frame rate, variable frame rate, fewer GOPs than devices -- and the caller falls back to
    ``build_gpu_stage``."""
    from .sharding import gop_aligned_ranges
    base = build_gpu_stage(stage, source_info, ffmpeg_bin)
    info = source_info
    if base is None or len(devices) < 2 or not info or not info.duration or not info.fps or info.is_vfr:
        return None
    params = stage.params
    fps = float(info.fps)
    total = int(round(info.duration * fps))
    try:
        gop = int(params.gop) if params.gop else max(1, round(fps))
    except (TypeError, ValueError):
        gop = max(1, round(fps))
    if total < gop * len(devices):
        return None
    out = Path(stage.output_path)
    shards: List[ShardCommand] = []
    lines = []
    ranges = [(a, b) for a, b in gop_aligned_ranges(total, len(devices), gop) if b > a]
    for k, (a, b) in enumerate(ranges):
        last = k == len(ranges) - 1
        part = out.with_name(f"{out.stem}.part{k:02d}{out.suffix}")
        dec = list(base.decode_cmd)
        i_at = dec.index("-i")
        if a > 0:
            # input seek half a frame early: ffmpeg's accurate seek drops frames whose pts lie before the target, so a target that
            # rounding (of this text or of container timestamps) puts a hair past frame a's pts would lose frame a
            dec[i_at:i_at] = ["-ss", f"{(a - 0.5) / fps:.6f}"]
        if not last:
            # the final range runs to the end of the stream: an under-estimated duration must not drop the tail
            o_at = dec.index("-f")
            dec[o_at:o_at] = ["-frames:v", str(b - a)]
        enc = []
        skip = 0
        src_inputs = 0
        for j, tok in enumerate(base.encode_cmd[:-1]):
            if skip:
                skip -= 1
                continue
            if tok == "-i":
                src_inputs += 1
                if src_inputs == 2:                           # the source as second input (audio / metadata): parts are video only
                    skip = 1
                    continue
            if tok in ("-map", "-map_metadata", "-c:a", "-b:a", "-ar", "-ac", "-movflags"):
                skip = 1
                continue
            enc.append(tok)
        enc += ["-map", "0:v:0", "-an", str(part)]
        shards.append(ShardCommand(index=k, device=devices[k % len(devices)], start_frame=a, frames=b - a, decode_cmd=dec,
                                   encode_cmd=enc, part_path=part, open_ended=last))
        lines.append("file '" + str(part).replace("'", "'\\''") + "'")
    list_path = out.with_name(f"{out.stem}.parts.txt")
    concat = [ffmpeg_bin, "-hide_banner"] + (["-y"] if params.overwrite else []) + \
             ["-f", "concat", "-safe", "0", "-i", str(list_path), "-i", str(stage.source_path), "-map", "0:v:0", "-map", "1:a?",
              "-map_metadata", "1", "-c:v", "copy"]
    if params.audio_codec:
        concat += ["-c:a", params.audio_codec]
    quiet: List[str] = []
    _output_tag_args(concat, params, source_info, True, quiet)
    if getattr(params, "faststart", False):
        concat += ["-movflags", "+faststart"]
    concat.append(str(out))
    return ShardedGpuStage(base=base, shards=shards, concat_list_path=list_path, concat_list_text="\n".join(lines) + "\n",
                           concat_cmd=concat)
