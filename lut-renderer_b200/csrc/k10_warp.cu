// 10-bit outputs of the ordered-row-walk kernel family (k10_warp.cuh): yuv420p10le -> yuv422p10le | yuv420p10le and
// yuv422p10le -> yuv420p10le.  The 8-bit output side is built in k10_warp8.cu (two translation units compile in parallel).
#include "k10_warp.cuh"

namespace lutr {

// K.tiles_per_row = ceil(w / 240); every warp of the grid gets an equal share of chroma rows
int launch_k10_warp(const StageK &k, bool in420, bool out420, int interp, int float_mode, bool grid_forced, void *stream) {
    if (in420 && !out420) return launch_k10<true, false, false>(k, interp, float_mode, grid_forced, stream);
    if (in420 && out420) return launch_k10<true, true, false>(k, interp, float_mode, grid_forced, stream);
    if (!in420 && out420) return launch_k10<false, true, false>(k, interp, float_mode, grid_forced, stream);
    return LUTR_ERR_UNSUPPORTED;  // 4:2:2 -> 4:2:2 is k422_warp.cu
}

}  // namespace lutr
