// yuv420p10le | yuv422p10le -> lut3d -> yuv420p (8 bit): a 10-bit source delivered through a codec without 10-bit support (the
// reference picks yuv420p for those, ffmpeg.py:299-302 -- the default libx264 among them).  k10_warp.cuh with the 8-bit output
// stage: ordered dither, per-tap truncating vertical chroma scaler (oracle/lutr_oracle.c orc_rgb_to_yuv_any, src_depth 10).
#include "k10_warp.cuh"

namespace lutr {

int launch_k10_warp8(const StageK &k, bool in420, int interp, int float_mode, bool grid_forced, void *stream) {
    if (in420) return launch_k10<true, true, true>(k, interp, float_mode, grid_forced, stream);
    return launch_k10<false, true, true>(k, interp, float_mode, grid_forced, stream);
}

}  // namespace lutr
