// Internal declarations shared by the host side (cube_parser.cpp, colour_tables.cpp, lutr_api.cu)
// and the kernels.  Not part of the C-ABI (include/lutr.h).
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
// the other dialects lut3d reads (lut_dialects.cpp): dialect = lutr_lut_dialect, or chosen by the file extension
int lut_dialect_of_path(const char *path);
int parse_lut_text(const char *text, size_t len, int dialect, CubeLut &out);
int parse_lut_file(const char *path, CubeLut &out);

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
    unsigned long long tex;  // experiment builds: cudaTextureObject_t over `tab` (linear, point sampling), 0 otherwise
};

struct PlaneSet {
    const uint8_t *in[3];
    uint8_t *out[3];
    int32_t in_stride[3], out_stride[3];  // bytes
    int64_t in_step, out_step;            // bytes between frames
};

// Constants of the fast kernels, derived on the host so they sit in the constant bank instead of registers.
struct FastK {
    uint32_t yA;                  // 128 * yC (10 bit)
    uint32_t vr4, vg4, ug4, ub4;  // 4 * chroma gains (u17 = 4*c15 - 65536 folded in)
    uint32_t kR, kG, kB;          // luma offset, rounding and the -65536*gain terms
    uint32_t kY, kU, kV;          // rounding/offset constants of RGB->YUV with the 0x4B000000 bias folded in
    // 10-bit 4:2:2 output side: chroma rows of the RGB->YUV matrix and kU / kV times 32 (mod 2^32), so that the 15-bit term
    // S >> 11 is the high half of the 32-bit sum and one PRMT packs the (u, v) pair (k422_warp.cu, pass B)
    uint32_t cu5[3], cv5[3], kU5, kV5;
    int32_t n2b, nb, totalb;      // byte steps of the LUT axes: n*n*16, n*16, sum + 16
    // arithmetic index path: s = min(fl(fl(x*in_scale) * sr), lut_max) computed as fma(2^23+x, idx_mul, idx_add) [* idx_sr[c]]
    float idx_mul, idx_add, idx_sr[3];   // idx_sr: per channel (a .cube DOMAIN may differ per channel); all 1 when folded
    int32_t idx_fold;             // 1: sr is a power of two and was folded into idx_mul (no second multiply)
    int32_t maxv;                 // 2^depth - 1
};

// Parameters of the generic tile kernel (k_any.cu)
struct AnyK {
    int32_t in_sub, in_depth;    // log2 chroma subsampling (sx | sy << 1) and sample depth of the source
    int32_t out_sub, out_depth;  // ... of the output
    int32_t out_full;            // 1: full-range output (15-bit range conversion with the constants below)
    int32_t pre10;               // full-range 10-bit source through the reference's 8-bit intermediate: 1 = pc -> tv, 2 = range kept
    int32_t in_lc, in_lo, in_cc, in_co;      // 8-bit pc -> tv pass: luma / chroma coefficient and offset
    int32_t out_lc, out_lo, out_cc, out_co;  // tv -> pc on the output side (output depth)
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
    FastK fk;
    int32_t clamp_out;        // 1: LUT entries leave [0,1], the store needs its clip
    int32_t grid_ctas;        // persistent grid size of the fast kernels
    int32_t idx_mode;         // 3: warp-window kernels, -1: generic tile kernel
    int32_t min_blocks;       // occupancy variant of the fast kernel (2, 3 or 4 CTAs per SM)
    // warp-window kernel (k422_warp.cu)
    int32_t tiles_per_row;       // ceil(w / 240)
    int32_t wk_pk;               // 1: packed f32x2 instructions
    float wk_one;                // 1.0f the compiler cannot constant-fold (see addp2 in k422_warp.cu)
    int32_t wk_unroll;           // pass B double-iterations unrolled together (1, 2, 4)
    int32_t wk_nomin;            // 1: the clip of the LUT coordinate to lut_max can never bind (host-checked), drop it
    int32_t wk_bulk;             // row-staging kernel: 1 = bulk async copies (TMA), 0 = per-lane cp.async
    int32_t wk_carveout;         // experiments: preferred shared-memory carveout in percent (0 = driver's choice)
    int32_t wk_var;              // experiment builds: variant id of the headline kernel (k422_warp.cu), -1 = default
    AnyK any;                    // generic tile kernel only
};

// pixel format properties (lutr_pix_fmt ids)
inline int pix_depth(int fmt) { return (fmt == LUTR_PIX_YUV422P10LE || fmt == LUTR_PIX_YUV420P10LE || fmt == LUTR_PIX_YUV444P10LE) ? 10 : 8; }
inline int pix_sub(int fmt) {  // log2 chroma subsampling: sx | sy << 1
    switch (fmt) {
        case LUTR_PIX_YUV420P: case LUTR_PIX_YUV420P10LE: return 3;
        case LUTR_PIX_YUV422P: case LUTR_PIX_YUV422P10LE: return 1;
        default: return 0;
    }
}
// swscale's 15-bit range conversion constants (colour_tables.cpp): out15 = min((in15 * coeff + offset) >> 14, 32767)
void range_coeffs(int depth, bool to_jpeg, bool chroma, int32_t *coeff, int32_t *offset);

// launchers; `stream` is a cudaStream_t
int launch_k422_warp(const StageK &k, bool in444, int interp, int float_mode, void *stream);  // yuv422p10le | yuv444p10le -> yuv422p10le
// TMA-staged form of the same stage (k422_tma.cu); k422_tma_ok: the geometry allows 16-byte bulk copies
bool k422_tma_ok(const StageK &k);
int launch_k422_tma(const StageK &k, int interp, int float_mode, void *stream);
int launch_k420_warp(const StageK &k, bool in422, int interp, int float_mode, bool grid_forced, void *stream);
// 10-bit chains with a 4:2:0 side (k10_warp.cu): yuv420p10le in and / or yuv420p10le out
int launch_k10_warp(const StageK &k, bool in420, bool out420, int interp, int float_mode, bool grid_forced, void *stream);
int launch_k10_warp8(const StageK &k, bool in420, int interp, int float_mode, bool grid_forced, void *stream);  // ... -> yuv420p (8 bit)
int launch_k_any(const StageK &k, int interp, int float_mode, void *stream);
int launch_k_convert(const StageK &k, int out_fmt, void *stream);  // LUT-less yuv422p10le -> yuv420p | yuv420p10le | yuv422p (k_convert.cu)  // generic tile kernel, the whole format matrix
int launch_k420_warp10(const StageK &k, bool in422, int interp, int float_mode, bool grid_forced, void *stream);  // yuv422p10le out
int launch_k420_warp422(const StageK &k, bool in422, int interp, int float_mode, bool grid_forced, void *stream);  // yuv422p (8 bit) out
int launch_k420_warp444(const StageK &k, bool in422, int interp, int float_mode, bool grid_forced, void *stream);  // yuv444p (8 bit) out

}  // namespace lutr
