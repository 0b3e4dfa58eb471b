"""B200-native 3D-LUT colour stage: drop-in for the per-pixel work LUT-renderer delegates to FFmpeg's
scale / lut3d / format filters.  Host-side operator surface mirrors ``lut_renderer`` (models, media_info,
ffmpeg); the pixels go through ``liblutr.so`` (include/lutr.h) -- CUDA for sm_100a only, no CPU fallback."""

__version__ = "0.1.0"

from .media_info import VideoInfo, probe_video  # noqa: F401
from .models import ProcessingParams, Task, TaskStatus  # noqa: F401
from .ffmpeg import (CommandStage, ColourStagePlan, GpuCommandStage, build_command, build_gpu_stage,  # noqa: F401
                     build_pipeline, plan_colour_stage)
