// yuv420p -> lut3d -> yuv422p10le (pro-mode master from an 8-bit source): the 4:2:0 warp-window kernel with the 10-bit 4:2:2
// output side.  See k420_warp.cu / k420_warp.cuh.
#define LUTR_K420_OUT 1
#include "k420_warp.cu"
