"""Generates tests/golden/build_command.json by RUNNING the reference's own Python operator builder
(/root/reference/src/lut_renderer/ffmpeg.py: build_command, build_pipeline) over a grid of parameters.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/gen_build_command_golden.py
The committed JSON is what tests/test_operator_surface.py checks our build_command / build_pipeline against.
"""
import hashlib
import itertools
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, "/root/reference/src")
from lut_renderer.ffmpeg import build_command, build_pipeline  # noqa: E402
from lut_renderer.media_info import VideoInfo  # noqa: E402
from lut_renderer.models import ProcessingParams, Task  # noqa: E402

SOURCES = {
    "none": None,
    "8bit_709_tv": dict(width=1920, height=1080, pix_fmt="yuv420p", bit_depth=8, colorspace="bt709", color_range="tv",
                        color_primaries="bt709", color_trc="bt709", fps=25.0),
    "8bit_pc_j": dict(width=1920, height=1080, pix_fmt="yuvj420p", bit_depth=8, colorspace="bt470bg", color_range="pc", fps=29.97),
    "8bit_pc_422": dict(width=1280, height=720, pix_fmt="yuv422p", bit_depth=8, colorspace=None, color_range="pc", fps=30.0),
    "10bit_709": dict(width=3840, height=2160, pix_fmt="yuv422p10le", bit_depth=10, colorspace="bt709", color_range="tv",
                      fps=23.976, bitrate="80000k"),
    "10bit_2020_vfr": dict(width=3840, height=2160, pix_fmt="yuv420p10le", bit_depth=10, colorspace="bt2020nc", color_range="tv",
                           fps=59.94, is_vfr=True, color_primaries="bt2020", color_trc="smpte2084"),
    "10bit_444_pc": dict(width=1920, height=1080, pix_fmt="yuv444p10le", bit_depth=10, colorspace="smpte240m", color_range="pc"),
}

PARAM_SETS = []
for codec, interp, matrix, tags, policy, pix in itertools.product(
        ["libx264", "libx265", "prores_ks", "h264_videotoolbox"], ["tetrahedral", "trilinear", "bogus", "cubic"],
        ["auto", "bt709", "none", "bt2020nc", "weird"], ["bt709", "inherit", "none", "other"],
        ["preserve", "force_8bit", "auto"], ["", "yuv420p", "yuv422p"]):
    PARAM_SETS.append(dict(video_codec=codec, lut_interp=interp, lut_input_matrix=matrix, lut_output_tags=tags,
                           bit_depth_policy=policy, pix_fmt=pix))
# thin the grid deterministically, then add hand-picked option mixes
PARAM_SETS = PARAM_SETS[::31]
PARAM_SETS += [
    dict(video_codec="libx264", zscale_dither="error_diffusion", pix_fmt="yuv420p"),
    dict(video_codec="libx264", fps="30000/1001", bitrate="8M", crf="18", preset="slow", tune="film", gop="48",
         profile="high", level="4.1", threads="8", faststart=True, audio_bitrate="192k", sample_rate="48000", channels="2"),
    dict(video_codec="libx264", bitrate="2.5M", force_cfr=False, inherit_color_metadata=False, overwrite=False),
    dict(video_codec="h264_videotoolbox", bitrate="60M"),
    dict(video_codec="libx264", audio_codec="copy", resolution="1280x720"),
    dict(video_codec="", audio_codec=""),
]


def main():
    cases = []
    for src_name, src in SOURCES.items():
        info = VideoInfo(**src) if src else None
        for pi, ps in enumerate(PARAM_SETS):
            for lut in (None, "look's \\dir/look.cube"):
                params = ProcessingParams(**ps)
                notes = []
                try:
                    cmd = build_command(Path("in put.mov"), Path("out.mp4"), params, lut_path=Path(lut) if lut else None,
                                        ffmpeg_bin="ffmpeg", source_info=info, notes=notes)
                    err = None
                except ValueError as exc:
                    cmd, err = None, str(exc)
                digest = hashlib.sha1(json.dumps([cmd, notes, err], ensure_ascii=False).encode("utf-8")).hexdigest()
                case = dict(source=src_name, params=ps, lut=lut, sha1=digest)
                if len(cases) % 25 == 0:  # keep some cases in clear text for readable diffs
                    case.update(cmd=cmd, notes=notes, error=err)
                cases.append(case)
    pipelines = []
    for mode, inter in (("fast", None), ("pro", "cache/master.mov"), ("pro", None)):
        params = ProcessingParams(processing_mode=mode, video_codec="libx264", bitrate="8M", crf="20", preset="fast",
                                  faststart=True, audio_bitrate="128k", level="4.0", tune="film")
        task = Task(task_id="t1", source_path=Path("a.mov"), output_path=Path("b.mp4"), lut_path=Path("l.cube"),
                    cover_path=None, params=params, intermediate_path=Path(inter) if inter else None)
        try:
            stages = build_pipeline(task)
            pipelines.append(dict(mode=mode, intermediate=inter, error=None, stages=[
                dict(name=s.name, source=str(s.source_path), output=str(s.output_path), params=s.params.to_dict(),
                     lut=str(s.lut_path) if s.lut_path else None, cleanup=s.cleanup_on_success, notes=s.notes,
                     probe=s.probe_source) for s in stages]))
        except ValueError as exc:
            pipelines.append(dict(mode=mode, intermediate=inter, error=str(exc), stages=None))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build_command.json")
    with open(out, "w", encoding="utf-8") as f:
        json.dump(dict(sources=SOURCES, cases=cases, pipelines=pipelines), f, ensure_ascii=False, indent=0)
    print(len(cases), "cases,", len(pipelines), "pipelines ->", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
