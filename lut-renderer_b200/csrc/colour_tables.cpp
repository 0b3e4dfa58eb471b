// Integer coefficient tables of the two conversion legs around lut3d, as libswscale 9.1.100 (FFmpeg 8.0.1)
// derives them for the reference's `scale=in_color_matrix=M:out_color_matrix=M` / auto-inserted scalers
// (reference src/lut_renderer/ffmpeg.py:229-236) and the final `format=PIXFMT` (ffmpeg.py:309).
// See DESIGN.md "Arithmetic" for the derivation; tests pin these against sws_getCoefficients() and
// against conversions done by the real library.
#include "lutr_internal.h"

namespace lutr {

namespace {

// inverse tables {crv, cbu, cgu, cgv}: V->R, U->B, U->G, V->G gains in 16.16, limited-range chroma
struct Inv {
    int64_t crv, cbu, cgu, cgv;
};
constexpr Inv kInv[5] = {
    {117489, 138438, 13975, 34925},  // bt709
    {104597, 132201, 25675, 53279},  // bt601 family / default
    {110013, 140363, 12277, 42626},  // bt2020
    {104448, 132798, 24759, 53109},  // fcc
    {117579, 136230, 16907, 35559},  // smpte240m
};

int32_t to_q13_s16(int64_t v16) {
    // (v * 2^13) rounded to the nearest multiple of 2^16, kept as a signed 16-bit lane
    int64_t r = ((v16 << 13) + 32768) >> 16;
    if (r > 32767) r = 32767;
    if (r < -32767) r = -32767;
    return (int32_t)r;
}

int64_t div_round(int64_t a, int64_t b) {
    const int64_t half = b >> 1;
    return (a >= 0 ? a + half : a - half) / b;
}

}  // namespace

bool make_yuv2rgb(int matrix, bool full_range, Yuv2RgbK &k) {
    if (matrix < 0 || matrix > 4) return false;
    Inv c = kInv[matrix];
    int64_t luma_gain = 65536, luma_off = 0;
    if (full_range) {
        // full-range chroma spans 255 instead of 224 codes
        c.crv = c.crv * 224 / 255;
        c.cbu = c.cbu * 224 / 255;
        c.cgu = -(c.cgu * 224 / 255);
        c.cgv = -(c.cgv * 224 / 255);
    } else {
        luma_gain = luma_gain * 255 / 219;
        luma_off = 16 << 16;
        c.cgu = -c.cgu;
        c.cgv = -c.cgv;
    }
    k.yC = to_q13_s16(luma_gain);
    k.yO8 = to_q13_s16(luma_off >> 10);  // == round16(luma_off << 3)
    k.vr = to_q13_s16(c.crv);
    k.ub = to_q13_s16(c.cbu);
    k.ug = to_q13_s16(c.cgu);
    k.vg = to_q13_s16(c.cgv);
    k.yOff10 = full_range ? 0 : 8192;
    return true;
}

bool make_rgb2yuv(int matrix, Rgb2YuvK &k) {
    if (matrix < 0 || matrix > 4) return false;
    if (matrix == LUTR_MATRIX_BT601) {
        // swscale keeps a literal BT.601 table for its default colourspace
        const double y = 219.0 / 255 * 32768, c = 224.0 / 255 * 32768;
        k.ry = (int)(0.299 * y + 0.5), k.gy = (int)(0.587 * y + 0.5), k.by = (int)(0.114 * y + 0.5);
        k.ru = -(int)(0.169 * c + 0.5), k.gu = -(int)(0.331 * c + 0.5), k.bu = (int)(0.500 * c + 0.5);
        k.rv = (int)(0.500 * c + 0.5), k.gv = -(int)(0.419 * c + 0.5), k.bv = -(int)(0.081 * c + 0.5);
        return true;
    }
    const Inv &c = kInv[matrix];
    const int64_t one = 65536, one2 = one * one, q15 = 32768;
    const int64_t gain_y = one * 255 / 219;
    // invert the yuv->rgb matrix symbolically: wb / wr are the B / R luma weights in 16.16 of 16.16
    const int64_t wb = div_round(one2 * -c.cgu, c.cbu);
    const int64_t wr = div_round(one2 * -c.cgv, c.crv);
    const int64_t wg = one2 - wb - wr;
    const int64_t ny = div_round(gain_y * wg, one), nu = div_round(c.cbu * wg, one), nv = div_round(c.crv * wg, one);
    k.ry = (int32_t)-div_round(q15 * wr, ny);
    k.gy = (int32_t)div_round(q15 * one2, ny);
    k.by = (int32_t)-div_round(q15 * wb, ny);
    k.ru = (int32_t)div_round(q15 * wr, nu);
    k.gu = (int32_t)-div_round(q15 * one2, nu);
    k.bu = (int32_t)div_round(q15 * (wg + wb), nu);
    k.rv = (int32_t)div_round(q15 * (wr + wg), nv);
    k.gv = (int32_t)-div_round(q15 * one2, nv);
    k.bv = (int32_t)div_round(q15 * wb, nv);
    return true;
}

// libswscale >= 7.1 range conversion on the 15-bit intermediates (lumRangeTo/FromJpeg_c, chrRangeTo/FromJpeg_c with the
// constants of solve_range_convert), pinned against libswscale 9.1.100 through oracle/lutr_oracle.c orc_range_coeffs:
//   coeff = ceil(((dst_max - dst_min) << 14) / (src_max - src_min)),
//   offset = (dst_max << (29 - depth)) - (src_max << (15 - depth)) * coeff + (1 << 13)
void range_coeffs(int depth, bool to_jpeg, bool chroma, int32_t *coeff, int32_t *offset) {
    const int ss = 15 - depth, ts = 14 + ss;
    const int64_t mpeg_min = 16 << (depth - 8), mpeg_max = (chroma ? 240 : 235) << (depth - 8), jpeg_max = (1 << depth) - 1;
    const int64_t smin = to_jpeg ? mpeg_min : 0, smax = to_jpeg ? mpeg_max : jpeg_max;
    const int64_t dmin = to_jpeg ? 0 : mpeg_min, dmax = to_jpeg ? jpeg_max : mpeg_max;
    const int64_t num = (dmax - dmin) << ts, den = (smax - smin) << ss;
    const int64_t c = (num + den - 1) / den;
    *coeff = (int32_t)c;
    *offset = (int32_t)((dmax << ts) - (smax << ss) * c + (1 << 13));
}

}  // namespace lutr
