// Internal declarations shared by the host side (cube_parser.cpp, colour_tables.cpp, lutr_api.cu)
// and the kernels (stage_kernels.cu).  Not part of the C-ABI (include/lutr.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <vector_types.h>

#include "../../include/lutr.h"

namespace lutr {

// ---- error channel ---------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

// ---- .cube ------------------------------------------------------------------------------------------
struct CubeLut {
    int size = 0;
    float scale[3] = {1.f, 1.f, 1.f};
    std::vector<float> rgb;  // size^3 * 3, order [r][g][b]
};
// returns LUTR_OK or a negative lutr_status and sets the error text
int parse_cube(const char *text, size_t len, CubeLut &out);
int parse_cube_file(const char *path, CubeLut &out);

// ---- colour tables (integer coefficients of the swscale legs) --------------------------------------
struct Yuv2RgbK {
    int32_t yC, yO8, vr, ub, ug, vg, yOff10;
};
struct Rgb2YuvK {
    int32_t ry, gy, by, ru, gu, bu, rv, gv, bv;
};
bool make_yuv2rgb(int matrix, bool full_range, Yuv2RgbK &k);
bool make_rgb2yuv(int matrix, Rgb2YuvK &k);

// ---- kernel parameter blocks ------------------------------------------------------------------------
struct LutK {
    const float4 *tab;  // [r][g][b] nodes as float4 (w unused), padded
    int32_t n, n2;      // size, size*size
    float lut_max;      // size - 1
    float in_scale;     // 1.0f / (2^depth - 1)
    float sr, sg, sb;   // scale.c * lut_max
    float out_scale;    // (float)(2^depth - 1)
};

struct PlaneSet {
    const uint8_t *in[3];
    uint8_t *out[3];
    int32_t in_stride[3], out_stride[3];  // bytes
    int64_t in_step, out_step;            // bytes between frames
};

struct StageK {
    int32_t w, h;
    int32_t nframes;
    int32_t range_pass;    // 1: 8-bit pc->tv pass in front of YUV->RGB
    int32_t rows_per_cta;  // 422p10: rows handled by one CTA;  420p: luma rows per band (even)
    int32_t bands;         // 420p: bands per frame
    Yuv2RgbK f;
    Rgb2YuvK b;
    LutK lut;
    PlaneSet p;
};

// launchers (stage_kernels.cu); `stream` is a cudaStream_t
int launch_yuv422p10(const StageK &k, int interp, int float_mode, void *stream);
int launch_yuv420p8(const StageK &k, int interp, int float_mode, void *stream);

}  // namespace lutr
