// yuv420p -> lut3d -> yuv444p (8 bit): the GUI's pix_fmt choice yuv444p on an 8-bit source (reference main_window.py:767-771,
// ffmpeg.py:304-310).  The 4:2:0 warp-window kernel with the 8-bit 4:4:4 output side (rgb24ToUV per pixel, no filter).
#define LUTR_K420_OUT 3
#include "k420_warp.cu"
