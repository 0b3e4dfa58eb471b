/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (lut-renderer_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 *
 * CPU restatement, in plain C, of the per-pixel colour stage that the reference delegates to FFmpeg
 * through the filter chain it builds in /root/reference/src/lut_renderer/ffmpeg.py:195-247,304-310
 *     [scale=in_range=pc:out_range=..,format=yuv4xxp,] (auto-)scale -> lut3d=file=..:interp=.. -> format=PIXFMT
 *
 * The arithmetic itself lives in a third-party dependency that is NOT vendored in /root/reference:
 * FFmpeg ("ffmpeg on PATH", un-pinned: readme.md:29,115).  We pin FFmpeg 8.0.1 semantics, because a real
 * libswscale 9.1.100 / libavutil 60.8.100 from that release ships in this image (opencv_python_headless.libs):
 *
 *   PINNED   swscale legs (orc_yuv420p_to_rgb24, orc_rgb24_to_yuv420p, orc_rgb24_to_yuv422p10, orc_yuv422p10_to_gbrp10,
 *            orc_gbrp10_to_yuv422p10, orc_range_pc_to_tv_8): closed-form integer restatements of the code
 *            paths libswscale takes for these conversions with flags=bicubic.  They are
 *            checked bit-exactly against the real binary (oracle/swscale_real.py) in
 *            tests/test_oracle_swscale.py on every run.
 *   UNPINNED vf_lut3d (orc_cube_parse*, orc_lut3d_*): libavfilter is absent from the image, so this is a
 *            restatement of upstream libavfilter/vf_lut3d.c (parse_cube, interp_nearest / interp_trilinear /
 *            interp_tetrahedral, DEFINE_INTERP_FUNC / DEFINE_INTERP_FUNC_PLANAR: C code, float32, separate
 *            mul/add, truncating store) written from its published algorithm.  "parity unpinned" for this
 *            leg: anchored on the reference call site ffmpeg.py:242-247 and on hand-computed KATs only.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, like FFmpeg's C build).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* coefficient tables                                                                                */
/* ------------------------------------------------------------------------------------------------ */

/* sws_getCoefficients() tables {crv, cbu, cgu, cgv} as found in libswscale 9.1.100.
 * id: 0 bt709, 1 bt601/smpte170m/bt470bg/"default", 2 bt2020, 3 fcc, 4 smpte240m */
static const int32_t k_inv_tables[5][4] = {
    {117489, 138438, 13975, 34925},
    {104597, 132201, 25675, 53279},
    {110013, 140363, 12277, 42626},
    {104448, 132798, 24759, 53109},
    {117579, 136230, 16907, 35559},
};

ORC_API const int32_t *orc_inv_table(int matrix_id)
{
    if (matrix_id < 0 || matrix_id > 4)
        return NULL;
    return k_inv_tables[matrix_id];
}

typedef struct OrcYuv2Rgb {
    int32_t yC, yO8; /* luma gain (Q13), luma offset * 8 */
    int32_t vr, ub, ug, vg;
    int32_t yOff10; /* 8192 (tv) or 0 (pc): offset of Y*128 in the 10-bit path */
} OrcYuv2Rgb;

static int32_t round_to_int16(int64_t f)
{
    int64_t r = (f + (1 << 15)) >> 16;
    if (r < -0x7FFF)
        return -0x7FFF; /* libswscale roundToInt16(): clamps to +-0x7FFF */
    if (r > 0x7FFF)
        return 0x7FFF;
    return (int32_t)r;
}

/* libswscale ff_yuv2rgb_c_init_tables() + the x86 coefficient packing (yuv2rgb init), brightness=0,
 * contrast=saturation=1.0 */
ORC_API void orc_yuv2rgb_setup(int matrix_id, int full_range, OrcYuv2Rgb *t)
{
    const int32_t *inv = k_inv_tables[matrix_id];
    int64_t crv = inv[0], cbu = inv[1], cgu = -inv[2], cgv = -inv[3];
    int64_t cy = 1 << 16, oy = 0;
    if (!full_range) {
        cy = (cy * 255) / 219;
        oy = 16 << 16;
    } else {
        crv = (crv * 224) / 255;
        cbu = (cbu * 224) / 255;
        cgu = (cgu * 224) / 255;
        cgv = (cgv * 224) / 255;
    }
    t->yC = round_to_int16(cy * 8192);
    t->yO8 = round_to_int16(oy * 8);
    t->vr = round_to_int16(crv * 8192);
    t->ub = round_to_int16(cbu * 8192);
    t->ug = round_to_int16(cgu * 8192);
    t->vg = round_to_int16(cgv * 8192);
    t->yOff10 = full_range ? 0 : 8192;
}

static int64_t rounded_div(int64_t a, int64_t b)
{
    return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b;
}

/* libswscale fill_rgb2yuv_table(): Q15, always limited range.  c[9] = ry gy by ru gu bu rv gv bv */
ORC_API void orc_rgb2yuv_setup(int matrix_id, int32_t c[9])
{
    const int32_t *inv = k_inv_tables[matrix_id];
    if (matrix_id == 1) { /* SWS_CS_DEFAULT: hard-coded BT.601 */
        c[0] = (int)(0.299 * 219 / 255 * (1 << 15) + 0.5);
        c[1] = (int)(0.587 * 219 / 255 * (1 << 15) + 0.5);
        c[2] = (int)(0.114 * 219 / 255 * (1 << 15) + 0.5);
        c[3] = -(int)(0.169 * 224 / 255 * (1 << 15) + 0.5);
        c[4] = -(int)(0.331 * 224 / 255 * (1 << 15) + 0.5);
        c[5] = (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        c[6] = (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        c[7] = -(int)(0.419 * 224 / 255 * (1 << 15) + 0.5);
        c[8] = -(int)(0.081 * 224 / 255 * (1 << 15) + 0.5);
        return;
    }
    int64_t vr = inv[0], ub = inv[1], ug = -inv[2], vg = -inv[3];
    const int64_t ONE = 65536;
    int64_t cy = ONE * 255 / 219;
    int64_t W = rounded_div(ONE * ONE * ug, ub);
    int64_t V = rounded_div(ONE * ONE * vg, vr);
    int64_t Z = ONE * ONE - W - V;
    int64_t Cy = rounded_div(cy * Z, ONE);
    int64_t Cu = rounded_div(ub * Z, ONE);
    int64_t Cv = rounded_div(vr * Z, ONE);
    const int64_t S = 1 << 15;
    c[0] = (int32_t)-rounded_div(S * V, Cy);
    c[1] = (int32_t)rounded_div(S * ONE * ONE, Cy);
    c[2] = (int32_t)-rounded_div(S * W, Cy);
    c[3] = (int32_t)rounded_div(S * V, Cu);
    c[4] = (int32_t)-rounded_div(S * ONE * ONE, Cu);
    c[5] = (int32_t)rounded_div(S * (Z + W), Cu);
    c[6] = (int32_t)rounded_div(S * (V + Z), Cv);
    c[7] = (int32_t)-rounded_div(S * ONE * ONE, Cv);
    c[8] = (int32_t)rounded_div(S * W, Cv);
}

static inline int clip_u(int v, int hi)
{
    return v < 0 ? 0 : (v > hi ? hi : v);
}

/* ------------------------------------------------------------------------------------------------ */
/* B.5  8-bit range pass  yuv4xxp(pc) -> yuv4xxp(tv)   (scale=in_range=pc:out_range=tv, format=yuv4xxp) */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_range_pc_to_tv_tables(uint8_t luma[256], uint8_t chroma[256])
{
    for (int v = 0; v < 256; v++) {
        int o15 = (int)((((int64_t)(v << 7)) * 14071 + 33561947) >> 14);
        luma[v] = (uint8_t)clip_u((o15 + 64) >> 7, 255);
        o15 = (int)((((int64_t)(v << 7)) * 7197 + 16762880) >> 13); /* fitted to 9.1.100, exact on all 256 inputs */
        chroma[v] = (uint8_t)clip_u((o15 + 64) >> 7, 255);
    }
}

ORC_API void orc_range_pc_to_tv_8(const uint8_t *src, int src_stride, uint8_t *dst, int dst_stride, int w, int h,
                                  int is_chroma)
{
    uint8_t tl[256], tc[256];
    orc_range_pc_to_tv_tables(tl, tc);
    const uint8_t *t = is_chroma ? tc : tl;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            dst[y * dst_stride + x] = t[src[y * src_stride + x]];
}

/* ------------------------------------------------------------------------------------------------ */
/* B.1  yuv420p -> rgb24   (x86 SIMD unscaled path; chroma replicated 2x2)                           */
/* ------------------------------------------------------------------------------------------------ */
ORC_API void orc_yuv420p_to_rgb24(const uint8_t *Y, int ys, const uint8_t *U, const uint8_t *V, int cs, uint8_t *rgb,
                                  int rs, int w, int h, int matrix_id, int full_range)
{
    OrcYuv2Rgb t;
    orc_yuv2rgb_setup(matrix_id, full_range, &t);
    for (int y = 0; y < h; y++) {
        const uint8_t *yr = Y + y * ys, *ur = U + (y >> 1) * cs, *vr_ = V + (y >> 1) * cs;
        uint8_t *o = rgb + y * rs;
        for (int x = 0; x < w; x++) {
            int yy = ((yr[x] * 8 - t.yO8) * t.yC) >> 16;
            int u = ur[x >> 1] * 8 - 1024, v = vr_[x >> 1] * 8 - 1024;
            o[3 * x + 0] = (uint8_t)clip_u(yy + ((v * t.vr) >> 16), 255);
            o[3 * x + 1] = (uint8_t)clip_u(yy + ((u * t.ug) >> 16) + ((v * t.vg) >> 16), 255);
            o[3 * x + 2] = (uint8_t)clip_u(yy + ((u * t.ub) >> 16), 255);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* B.2  rgb24 -> yuv420p   (rgb24ToY / rgb24ToUV_half + 8-tap vertical bicubic chroma decimation)    */
/* ------------------------------------------------------------------------------------------------ */
static const int k_v420_taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58}; /* Q12, sum 4096 */

ORC_API void orc_rgb24_to_yuv420p(const uint8_t *rgb, int rs, uint8_t *Y, int ys, uint8_t *U, uint8_t *V, int cs, int w,
                                  int h, int matrix_id)
{
    int32_t c[9];
    orc_rgb2yuv_setup(matrix_id, c);
    const int cw = w / 2, ch = h / 2;
    int16_t *u15 = (int16_t *)malloc(sizeof(int16_t) * (size_t)cw * h);
    int16_t *v15 = (int16_t *)malloc(sizeof(int16_t) * (size_t)cw * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *p = rgb + y * rs;
        for (int x = 0; x < w; x++) {
            int r = p[3 * x], g = p[3 * x + 1], b = p[3 * x + 2];
            int y14 = (c[0] * r + c[1] * g + c[2] * b + (32 << 14) + (1 << 8)) >> 9;
            Y[y * ys + x] = (uint8_t)clip_u((2 * y14 + 64) >> 7, 255);
        }
        for (int i = 0; i < cw; i++) {
            int r = p[6 * i] + p[6 * i + 3], g = p[6 * i + 1] + p[6 * i + 4], b = p[6 * i + 2] + p[6 * i + 5];
            int u14 = (c[3] * r + c[4] * g + c[5] * b + (256 << 15) + (1 << 9)) >> 10;
            int v14 = (c[6] * r + c[7] * g + c[8] * b + (256 << 15) + (1 << 9)) >> 10;
            u15[y * cw + i] = (int16_t)(2 * u14);
            v15[y * cw + i] = (int16_t)(2 * v14);
        }
    }
    for (int j = 0; j < ch; j++) {
        /* fold taps whose source row lies outside the frame onto the edge row */
        int rows[8], coef[8], n = 0;
        for (int k = 0; k < 8; k++) {
            int row = 2 * j - 3 + k;
            row = row < 0 ? 0 : (row > h - 1 ? h - 1 : row);
            if (n && rows[n - 1] == row)
                coef[n - 1] += k_v420_taps[k];
            else {
                rows[n] = row;
                coef[n] = k_v420_taps[k];
                n++;
            }
        }
        const int c_path = (j == ch - 1); /* swscale drops to the C vertical scaler for the last output rows */
        for (int i = 0; i < cw; i++) {
            int su, sv;
            if (c_path) {
                su = sv = 64 << 12;
                for (int k = 0; k < n; k++) {
                    su += u15[rows[k] * cw + i] * coef[k];
                    sv += v15[rows[k] * cw + i] * coef[k];
                }
                su >>= 19;
                sv >>= 19;
            } else {
                su = sv = 7;
                for (int k = 0; k < n; k++) { /* pmulhw: per-tap truncation */
                    su += (u15[rows[k] * cw + i] * coef[k]) >> 16;
                    sv += (v15[rows[k] * cw + i] * coef[k]) >> 16;
                }
                su >>= 3;
                sv >>= 3;
            }
            U[j * cs + i] = (uint8_t)clip_u(su, 255);
            V[j * cs + i] = (uint8_t)clip_u(sv, 255);
        }
    }
    free(u15);
    free(v15);
}

/* ------------------------------------------------------------------------------------------------ */
/* B.3  yuv422p10le -> gbrp10le   (4-tap horizontal bicubic chroma upsample + yuv2gbrp_full_X)       */
/* ------------------------------------------------------------------------------------------------ */
static inline int clampi(int v, int lo, int hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

static int chroma_up_422(const uint16_t *C, int cw, int x)
{
    int64_t s;
    const int i = x >> 1;
    if (x == 0)
        s = 17729 * (int64_t)C[0] - 1345 * (int64_t)C[cw > 1 ? 1 : 0];
    else if (x == 2 && cw > 2)
        s = 3835 * (int64_t)C[0] + 13894 * (int64_t)C[1] - 1345 * (int64_t)C[2];
    else if (!(x & 1))
        s = -461 * (int64_t)C[clampi(i - 2, 0, cw - 1)] + 3943 * (int64_t)C[clampi(i - 1, 0, cw - 1)] +
            14284 * (int64_t)C[i] - 1382 * (int64_t)C[clampi(i + 1, 0, cw - 1)];
    else
        s = -1382 * (int64_t)C[clampi(i - 1, 0, cw - 1)] + 14284 * (int64_t)C[i] +
            3943 * (int64_t)C[clampi(i + 1, 0, cw - 1)] - 461 * (int64_t)C[clampi(i + 2, 0, cw - 1)];
    s >>= 9;
    return (int)(s > 32767 ? 32767 : s);
}

static inline uint16_t clip30_shift20(int32_t v)
{
    /* av_clip_uintp2(v, 30) >> 20 */
    if (v & ~0x3FFFFFFF)
        v = (~v) >> 31 & 0x3FFFFFFF;
    return (uint16_t)(v >> 20);
}

ORC_API void orc_yuv422p10_to_gbrp10(const uint16_t *Y, int ys, const uint16_t *U, const uint16_t *V, int cs, uint16_t *R,
                                     uint16_t *G, uint16_t *B, int os, int w, int h, int matrix_id, int full_range)
{
    /* strides in elements */
    OrcYuv2Rgb t;
    orc_yuv2rgb_setup(matrix_id, full_range, &t);
    const int cw = w / 2;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int cu = chroma_up_422(U + y * cs, cw, x), cv = chroma_up_422(V + y * cs, cw, x);
            int32_t u17 = (int32_t)(((1 << 9) - (128 << 19) + (int64_t)cu * 4096) >> 10);
            int32_t v17 = (int32_t)(((1 << 9) - (128 << 19) + (int64_t)cv * 4096) >> 10);
            /* wrapping int32 arithmetic, as the C code in libswscale/output.c does */
            uint32_t yy = (uint32_t)(Y[y * ys + x] * 128 - t.yOff10) * (uint32_t)t.yC + (1u << 19);
            uint32_t r = yy + (uint32_t)v17 * (uint32_t)t.vr;
            uint32_t g = yy + (uint32_t)v17 * (uint32_t)t.vg + (uint32_t)u17 * (uint32_t)t.ug;
            uint32_t b = yy + (uint32_t)u17 * (uint32_t)t.ub;
            R[y * os + x] = clip30_shift20((int32_t)r);
            G[y * os + x] = clip30_shift20((int32_t)g);
            B[y * os + x] = clip30_shift20((int32_t)b);
        }
    }
}

/* yuv420p10le -> gbrp10le: forward leg of the pro-mode master for 10-bit 4:2:0 sources (ffmpeg.py:417-433; the intermediate
 * is gbrp10le as for 4:2:2).  PINNED against libswscale 9.1.100 (tests/test_oracle_swscale.py).  swscale scales chroma
 * horizontally first (the 4-tap filter of chroma_up_422, to 15 bits, per chroma row) and vertically inside the output function
 * (yuv2gbrp_full_X): Q12 bicubic taps at the quarter phases over 4 chroma rows, the two top luma rows with edge-normalised taps:
 *   y = 0:              4432 C15[0] - 336 C15[1]
 *   y = 2:              959 C15[0] + 3473 C15[1] - 336 C15[2]
 *   y = 2j   (others):  -115 C15[j-2] + 985 C15[j-1] + 3572 C15[j] - 346 C15[j+1]
 *   y = 2j+1:           -346 C15[j-1] + 3572 C15[j] + 985 C15[j+1] - 115 C15[j+2]      (row indices clamped to the plane)
 * then u17 = ((1<<9) - (128<<19) + sum) >> 10 and the matrix / clip of orc_yuv422p10_to_gbrp10. */
static void vtaps_420(int y, int ch, int rows[4], int taps[4])
{
    const int j = y >> 1;
    if (y & 1) {
        rows[0] = j - 1, rows[1] = j, rows[2] = j + 1, rows[3] = j + 2;
        taps[0] = -346, taps[1] = 3572, taps[2] = 985, taps[3] = -115;
    } else {
        rows[0] = j - 2, rows[1] = j - 1, rows[2] = j, rows[3] = j + 1;
        if (y == 0)
            taps[0] = 0, taps[1] = 0, taps[2] = 4432, taps[3] = -336;
        else if (y == 2)
            taps[0] = 0, taps[1] = 959, taps[2] = 3473, taps[3] = -336;
        else
            taps[0] = -115, taps[1] = 985, taps[2] = 3572, taps[3] = -346;
    }
    for (int k = 0; k < 4; k++)
        rows[k] = clampi(rows[k], 0, ch - 1);
}

ORC_API void orc_yuv420p10_to_gbrp10(const uint16_t *Y, int ys, const uint16_t *U, const uint16_t *V, int cs, uint16_t *R,
                                     uint16_t *G, uint16_t *B, int os, int w, int h, int matrix_id, int full_range)
{
    /* strides in elements */
    OrcYuv2Rgb t;
    orc_yuv2rgb_setup(matrix_id, full_range, &t);
    const int cw = w / 2, ch = h / 2;
    /* horizontally upsampled 15-bit chroma of every chroma row */
    int32_t *hu = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * ch);
    int32_t *hv = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * ch);
    for (int j = 0; j < ch; j++)
        for (int x = 0; x < w; x++) {
            hu[(size_t)j * w + x] = chroma_up_422(U + j * cs, cw, x);
            hv[(size_t)j * w + x] = chroma_up_422(V + j * cs, cw, x);
        }
    for (int y = 0; y < h; y++) {
        int rows[4], taps[4];
        vtaps_420(y, ch, rows, taps);
        for (int x = 0; x < w; x++) {
            int64_t su = 0, sv = 0;
            for (int k = 0; k < 4; k++) {
                su += (int64_t)hu[(size_t)rows[k] * w + x] * taps[k];
                sv += (int64_t)hv[(size_t)rows[k] * w + x] * taps[k];
            }
            int32_t u17 = (int32_t)(((1 << 9) - (128 << 19) + su) >> 10);
            int32_t v17 = (int32_t)(((1 << 9) - (128 << 19) + sv) >> 10);
            uint32_t yy = (uint32_t)(Y[y * ys + x] * 128 - t.yOff10) * (uint32_t)t.yC + (1u << 19);
            uint32_t r = yy + (uint32_t)v17 * (uint32_t)t.vr;
            uint32_t g = yy + (uint32_t)v17 * (uint32_t)t.vg + (uint32_t)u17 * (uint32_t)t.ug;
            uint32_t b = yy + (uint32_t)u17 * (uint32_t)t.ub;
            R[y * os + x] = clip30_shift20((int32_t)r);
            G[y * os + x] = clip30_shift20((int32_t)g);
            B[y * os + x] = clip30_shift20((int32_t)b);
        }
    }
    free(hu);
    free(hv);
}

/* ------------------------------------------------------------------------------------------------ */
/* B.4  gbrp10le -> yuv422p10le   (planar_rgb16_to_y/uv + 8-tap horizontal bicubic + yuv2plane1_10)   */
/* ------------------------------------------------------------------------------------------------ */
static const int k_h422_taps[8] = {-230, -692, 1972, 7142, 7142, 1972, -692, -230}; /* Q14, sum 16384 */

ORC_API void orc_gbrp10_to_yuv422p10(const uint16_t *R, const uint16_t *G, const uint16_t *B, int is, uint16_t *Y, int ys,
                                     uint16_t *U, uint16_t *V, int cs, int w, int h, int matrix_id)
{
    int32_t c[9];
    orc_rgb2yuv_setup(matrix_id, c);
    const int cw = w / 2;
    int32_t *u14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    int32_t *v14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int r = R[y * is + x], g = G[y * is + x], b = B[y * is + x];
            int y14 = (c[0] * r + c[1] * g + c[2] * b + (16 << 17) + (1 << 10)) >> 11;
            int y15 = 2 * y14;
            if (y15 > 32767)
                y15 = 32767;
            Y[y * ys + x] = (uint16_t)clip_u((y15 + 16) >> 5, 1023);
            u14[x] = (c[3] * r + c[4] * g + c[5] * b + (128 << 17) + (1 << 10)) >> 11;
            v14[x] = (c[6] * r + c[7] * g + c[8] * b + (128 << 17) + (1 << 10)) >> 11;
        }
        for (int i = 0; i < cw; i++) {
            int64_t su = 0, sv = 0;
            for (int k = 0; k < 8; k++) {
                int x = clampi(2 * i - 3 + k, 0, w - 1);
                su += (int64_t)u14[x] * k_h422_taps[k];
                sv += (int64_t)v14[x] * k_h422_taps[k];
            }
            su >>= 13;
            sv >>= 13;
            if (su > 32767)
                su = 32767;
            if (sv > 32767)
                sv = 32767;
            U[y * cs + i] = (uint16_t)clip_u((int)((su + 16) >> 5), 1023);
            V[y * cs + i] = (uint16_t)clip_u((int)((sv + 16) >> 5), 1023);
        }
    }
    free(u14);
    free(v14);
}

/* gbrp10le -> yuv420p10le: return leg when the encoder wants 10-bit 4:2:0 (libx265 / hevc_videotoolbox with
 * bit_depth_policy preserve, ffmpeg.py:297-298, 309-310).  PINNED against libswscale 9.1.100
 * (tests/test_oracle_swscale.py).  Luma and the per-column chroma terms as orc_gbrp10_to_yuv422p10; swscale then scales
 * chroma horizontally (the 8-tap Q14 filter, to 15 bits with the 32767 clip, per luma row) and vertically in
 * yuv2planeX_10: Q12 taps {-58,-172,492,1786,1786,492,-172,-58} over luma rows 2j-3 .. 2j+4 (row indices clamped to the
 * frame = swscale's edge-folded coefficients, the accumulation is exact), out = clip10((sum + (1 << 16)) >> 17). */
ORC_API void orc_gbrp10_to_yuv420p10(const uint16_t *R, const uint16_t *G, const uint16_t *B, int is, uint16_t *Y, int ys,
                                     uint16_t *U, uint16_t *V, int cs, int w, int h, int matrix_id)
{
    int32_t c[9];
    orc_rgb2yuv_setup(matrix_id, c);
    const int cw = w / 2, ch = h / 2;
    int32_t *u14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    int32_t *v14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    int32_t *hu = (int32_t *)malloc(sizeof(int32_t) * (size_t)cw * h); /* 15-bit horizontally decimated rows */
    int32_t *hv = (int32_t *)malloc(sizeof(int32_t) * (size_t)cw * h);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            int r = R[y * is + x], g = G[y * is + x], b = B[y * is + x];
            int y14 = (c[0] * r + c[1] * g + c[2] * b + (16 << 17) + (1 << 10)) >> 11;
            int y15 = 2 * y14;
            if (y15 > 32767)
                y15 = 32767;
            Y[y * ys + x] = (uint16_t)clip_u((y15 + 16) >> 5, 1023);
            u14[x] = (c[3] * r + c[4] * g + c[5] * b + (128 << 17) + (1 << 10)) >> 11;
            v14[x] = (c[6] * r + c[7] * g + c[8] * b + (128 << 17) + (1 << 10)) >> 11;
        }
        for (int i = 0; i < cw; i++) {
            int64_t su = 0, sv = 0;
            for (int k = 0; k < 8; k++) {
                int x = clampi(2 * i - 3 + k, 0, w - 1);
                su += (int64_t)u14[x] * k_h422_taps[k];
                sv += (int64_t)v14[x] * k_h422_taps[k];
            }
            su >>= 13;
            sv >>= 13;
            hu[(size_t)y * cw + i] = su > 32767 ? 32767 : (int32_t)su;
            hv[(size_t)y * cw + i] = sv > 32767 ? 32767 : (int32_t)sv;
        }
    }
    for (int j = 0; j < ch; j++)
        for (int i = 0; i < cw; i++) {
            int64_t su = 1 << 16, sv = 1 << 16;
            for (int k = 0; k < 8; k++) {
                int y = clampi(2 * j - 3 + k, 0, h - 1);
                su += (int64_t)hu[(size_t)y * cw + i] * k_v420_taps[k];
                sv += (int64_t)hv[(size_t)y * cw + i] * k_v420_taps[k];
            }
            U[j * cs + i] = (uint16_t)clip_u((int)(su >> 17), 1023);
            V[j * cs + i] = (uint16_t)clip_u((int)(sv >> 17), 1023);
        }
    free(u14);
    free(v14);
    free(hu);
    free(hv);
}

/* ------------------------------------------------------------------------------------------------ */
/* vf_lut3d restatement: .cube parser                                                                */
/* ------------------------------------------------------------------------------------------------ */
#define ORC_MAX_LINE 512
#define ORC_MAX_LEVEL 256

enum { ORC_OK = 0, ORC_EINVAL = -22, ORC_INVALIDDATA = -1094995529 /* AVERROR_INVALIDDATA */, ORC_ENOMEM = -12,
       ORC_ENOENT = -2 };

typedef struct OrcLut {
    int size;
    float scale[3]; /* clip(1/(max-min), 0, 1) */
    float *rgb;     /* size^3 * 3, index [r][g][b] */
} OrcLut;

static int is_space(int ch)
{
    return ch == ' ' || ch == '\f' || ch == '\n' || ch == '\r' || ch == '\t' || ch == '\v';
}

static int skip_line(const char *p)
{
    while (*p && is_space((unsigned char)*p))
        p++;
    return !*p || *p == '#';
}

static float clipf(float v, float lo, float hi)
{
    /* av_clipf: FFMIN(FFMAX(a, amin), amax) */
    float m = v > lo ? v : lo;
    return m > hi ? hi : m;
}

ORC_API void orc_lut_free(OrcLut *l)
{
    if (l) {
        free(l->rgb);
        free(l);
    }
}

/* parse_cube(): scan for LUT_3D_SIZE; then size^3 data lines, blue-major loops (red fastest in the file) */
ORC_API int orc_cube_parse_file(FILE *f, OrcLut **out)
{
    char line[ORC_MAX_LINE];
    float min[3] = {0.f, 0.f, 0.f}, max[3] = {1.f, 1.f, 1.f};
    OrcLut *l = (OrcLut *)calloc(1, sizeof(*l));
    if (!l)
        return ORC_ENOMEM;
    *out = NULL;
    while (fgets(line, sizeof(line), f)) {
        if (!strncmp(line, "LUT_3D_SIZE", 11)) {
            const int size = (int)strtol(line + 12, NULL, 0);
            if (size < 2 || size > ORC_MAX_LEVEL) {
                orc_lut_free(l);
                return ORC_EINVAL; /* "Too large or invalid 3D LUT size" */
            }
            const int size2 = size * size;
            l->size = size;
            l->rgb = (float *)malloc(sizeof(float) * 3 * (size_t)size2 * size);
            if (!l->rgb) {
                orc_lut_free(l);
                return ORC_ENOMEM;
            }
            for (int k = 0; k < size; k++)
                for (int j = 0; j < size; j++)
                    for (int i = 0; i < size; i++) {
                        float *vec = l->rgb + 3 * ((size_t)i * size2 + j * size + k);
                        for (;;) {
                            if (!fgets(line, sizeof(line), f)) {
                                orc_lut_free(l);
                                return ORC_INVALIDDATA; /* "Unexpected EOF" */
                            }
                            if (!strncmp(line, "DOMAIN_", 7)) {
                                float *vals = NULL;
                                if (!strncmp(line + 7, "MIN ", 4))
                                    vals = min;
                                else if (!strncmp(line + 7, "MAX ", 4))
                                    vals = max;
                                if (!vals) {
                                    orc_lut_free(l);
                                    return ORC_INVALIDDATA;
                                }
                                sscanf(line + 11, "%f %f %f", vals, vals + 1, vals + 2);
                                continue;
                            }
                            if (!strncmp(line, "TITLE", 5))
                                continue;
                            if (!skip_line(line))
                                break;
                        }
                        if (sscanf(line, "%f %f %f", vec, vec + 1, vec + 2) != 3) {
                            orc_lut_free(l);
                            return ORC_INVALIDDATA;
                        }
                    }
            break;
        }
    }
    if (!l->size) {
        orc_lut_free(l);
        return ORC_INVALIDDATA; /* "3D LUT is empty" */
    }
    for (int c = 0; c < 3; c++)
        l->scale[c] = clipf((float)(1. / (max[c] - min[c])), 0.f, 1.f);
    *out = l;
    return ORC_OK;
}

ORC_API int orc_cube_parse_path(const char *path, OrcLut **out)
{
    FILE *f = fopen(path, "r");
    if (!f)
        return ORC_ENOENT;
    int r = orc_cube_parse_file(f, out);
    fclose(f);
    return r;
}

ORC_API int orc_cube_parse_memory(const char *buf, size_t len, OrcLut **out)
{
    FILE *f = fmemopen((void *)buf, len, "r");
    if (!f)
        return ORC_ENOMEM;
    int r = orc_cube_parse_file(f, out);
    fclose(f);
    return r;
}

ORC_API int orc_lut_size(const OrcLut *l) { return l->size; }
ORC_API const float *orc_lut_data(const OrcLut *l) { return l->rgb; }
ORC_API const float *orc_lut_scale(const OrcLut *l) { return l->scale; }

ORC_API OrcLut *orc_lut_from_table(int size, const float *rgb, const float scale[3])
{
    OrcLut *l = (OrcLut *)calloc(1, sizeof(*l));
    size_t n = 3 * (size_t)size * size * size;
    l->size = size;
    l->rgb = (float *)malloc(sizeof(float) * n);
    memcpy(l->rgb, rgb, sizeof(float) * n);
    for (int c = 0; c < 3; c++)
        l->scale[c] = scale ? scale[c] : 1.f;
    return l;
}

/* ------------------------------------------------------------------------------------------------ */
/* vf_lut3d restatement: interpolation (C code path: float32, no FMA contraction)                    */
/* ------------------------------------------------------------------------------------------------ */
typedef struct rgbvec {
    float r, g, b;
} rgbvec;

enum { ORC_INTERP_NEAREST = 0, ORC_INTERP_TRILINEAR = 1, ORC_INTERP_TETRAHEDRAL = 2, ORC_INTERP_PYRAMID = 3, ORC_INTERP_PRISM = 4 };

#define LUTAT(l, ri, gi, bi) (*(const rgbvec *)((l)->rgb + 3 * ((size_t)(ri) * (l)->size * (l)->size + (gi) * (l)->size + (bi))))
#define PREV(x) ((int)(x))
#define NEXT(x, l) ((int)(x) + 1 < (l)->size - 1 ? (int)(x) + 1 : (l)->size - 1)
#define NEAR(x) ((int)((x) + .5))

/* Float evaluation order of the blend.  0 = the C code (one rounding per multiply and per add: what packed rgb24 always gets).
 * 1 = contracted multiply-adds in the same operand order, the first product of a sum plain: the shape of the AVX2 + FMA3 assembly
 * upstream installs for planar formats on x86-64 (libavfilter/x86/vf_lut3d.asm, fmaddps) -- restated from memory, unverifiable here.
 * Set once before use (orc_set_float_mode); pyramid / prism / nearest have no assembly and ignore it. */
static int g_float_mode = 0;
ORC_API void orc_set_float_mode(int mode) { g_float_mode = mode ? 1 : 0; }
ORC_API int orc_get_float_mode(void) { return g_float_mode; }

static inline float lerpf(float v0, float v1, float f)
{
    return g_float_mode ? fmaf(v1 - v0, f, v0) : v0 + (v1 - v0) * f;
}

static inline rgbvec lerp(const rgbvec *v0, const rgbvec *v1, float f)
{
    rgbvec v = {lerpf(v0->r, v1->r, f), lerpf(v0->g, v1->g, f), lerpf(v0->b, v1->b, f)};
    return v;
}

static inline rgbvec interp_nearest(const OrcLut *l, const rgbvec *s)
{
    return LUTAT(l, NEAR(s->r), NEAR(s->g), NEAR(s->b));
}

static inline rgbvec interp_trilinear(const OrcLut *l, const rgbvec *s)
{
    const int p[3] = {PREV(s->r), PREV(s->g), PREV(s->b)};
    const int n[3] = {NEXT(s->r, l), NEXT(s->g, l), NEXT(s->b, l)};
    const rgbvec d = {s->r - p[0], s->g - p[1], s->b - p[2]};
    const rgbvec c000 = LUTAT(l, p[0], p[1], p[2]), c001 = LUTAT(l, p[0], p[1], n[2]);
    const rgbvec c010 = LUTAT(l, p[0], n[1], p[2]), c011 = LUTAT(l, p[0], n[1], n[2]);
    const rgbvec c100 = LUTAT(l, n[0], p[1], p[2]), c101 = LUTAT(l, n[0], p[1], n[2]);
    const rgbvec c110 = LUTAT(l, n[0], n[1], p[2]), c111 = LUTAT(l, n[0], n[1], n[2]);
    const rgbvec c00 = lerp(&c000, &c100, d.r), c10 = lerp(&c010, &c110, d.r);
    const rgbvec c01 = lerp(&c001, &c101, d.r), c11 = lerp(&c011, &c111, d.r);
    const rgbvec c0 = lerp(&c00, &c10, d.g), c1 = lerp(&c01, &c11, d.g);
    return lerp(&c0, &c1, d.b);
}

/* w0*c000 + w1*cA + w2*cB + w3*c111, evaluated left to right per channel like the upstream expressions */
static inline rgbvec tetra_blend(float w0, const rgbvec *c000, float w1, const rgbvec *cA, float w2, const rgbvec *cB,
                                 float w3, const rgbvec *c111)
{
    rgbvec c;
    if (g_float_mode) {
        c.r = fmaf(w3, c111->r, fmaf(w2, cB->r, fmaf(w1, cA->r, w0 * c000->r)));
        c.g = fmaf(w3, c111->g, fmaf(w2, cB->g, fmaf(w1, cA->g, w0 * c000->g)));
        c.b = fmaf(w3, c111->b, fmaf(w2, cB->b, fmaf(w1, cA->b, w0 * c000->b)));
        return c;
    }
    c.r = w0 * c000->r + w1 * cA->r + w2 * cB->r + w3 * c111->r;
    c.g = w0 * c000->g + w1 * cA->g + w2 * cB->g + w3 * c111->g;
    c.b = w0 * c000->b + w1 * cA->b + w2 * cB->b + w3 * c111->b;
    return c;
}

static inline rgbvec interp_tetrahedral(const OrcLut *l, const rgbvec *s)
{
    const int p[3] = {PREV(s->r), PREV(s->g), PREV(s->b)};
    const int n[3] = {NEXT(s->r, l), NEXT(s->g, l), NEXT(s->b, l)};
    const rgbvec d = {s->r - p[0], s->g - p[1], s->b - p[2]};
    const rgbvec c000 = LUTAT(l, p[0], p[1], p[2]), c111 = LUTAT(l, n[0], n[1], n[2]);
    if (d.r > d.g) {
        if (d.g > d.b) {
            const rgbvec c100 = LUTAT(l, n[0], p[1], p[2]), c110 = LUTAT(l, n[0], n[1], p[2]);
            return tetra_blend(1 - d.r, &c000, d.r - d.g, &c100, d.g - d.b, &c110, d.b, &c111);
        } else if (d.r > d.b) {
            const rgbvec c100 = LUTAT(l, n[0], p[1], p[2]), c101 = LUTAT(l, n[0], p[1], n[2]);
            return tetra_blend(1 - d.r, &c000, d.r - d.b, &c100, d.b - d.g, &c101, d.g, &c111);
        } else {
            const rgbvec c001 = LUTAT(l, p[0], p[1], n[2]), c101 = LUTAT(l, n[0], p[1], n[2]);
            return tetra_blend(1 - d.b, &c000, d.b - d.r, &c001, d.r - d.g, &c101, d.g, &c111);
        }
    } else {
        if (d.b > d.g) {
            const rgbvec c001 = LUTAT(l, p[0], p[1], n[2]), c011 = LUTAT(l, p[0], n[1], n[2]);
            return tetra_blend(1 - d.b, &c000, d.b - d.g, &c001, d.g - d.r, &c011, d.r, &c111);
        } else if (d.b > d.r) {
            const rgbvec c010 = LUTAT(l, p[0], n[1], p[2]), c011 = LUTAT(l, p[0], n[1], n[2]);
            return tetra_blend(1 - d.g, &c000, d.g - d.b, &c010, d.b - d.r, &c011, d.r, &c111);
        } else {
            const rgbvec c010 = LUTAT(l, p[0], n[1], p[2]), c110 = LUTAT(l, n[0], n[1], p[2]);
            return tetra_blend(1 - d.g, &c000, d.g - d.r, &c010, d.r - d.b, &c110, d.b, &c111);
        }
    }
}

/* interp=pyramid / interp=prism (accepted by the reference's whitelist, ffmpeg.py:243; not offered by its GUI).  Restated from
 * upstream vf_lut3d.c interp_pyramid / interp_prism: one expression per channel, evaluated left to right in float32. */
#define PYR(ch, A1, A0, B1, B0, C1, C0, D3, D2, D1, D0, f1, f2, f3, g1, g2)                                                   \
    (c000.ch + (A1.ch - A0.ch) * (f1) + (B1.ch - B0.ch) * (f2) + (C1.ch - C0.ch) * (f3) +                                      \
     (D3.ch - D2.ch - D1.ch + D0.ch) * (g1) * (g2))

static inline rgbvec interp_pyramid(const OrcLut *l, const rgbvec *s)
{
    const int p[3] = {PREV(s->r), PREV(s->g), PREV(s->b)};
    const int n[3] = {NEXT(s->r, l), NEXT(s->g, l), NEXT(s->b, l)};
    const rgbvec d = {s->r - p[0], s->g - p[1], s->b - p[2]};
    const rgbvec c000 = LUTAT(l, p[0], p[1], p[2]), c111 = LUTAT(l, n[0], n[1], n[2]);
    rgbvec c;
    if (d.g > d.r && d.b > d.r) {
        const rgbvec c001 = LUTAT(l, p[0], p[1], n[2]), c010 = LUTAT(l, p[0], n[1], p[2]), c011 = LUTAT(l, p[0], n[1], n[2]);
        c.r = PYR(r, c111, c011, c010, c000, c001, c000, c011, c001, c010, c000, d.r, d.g, d.b, d.g, d.b);
        c.g = PYR(g, c111, c011, c010, c000, c001, c000, c011, c001, c010, c000, d.r, d.g, d.b, d.g, d.b);
        c.b = PYR(b, c111, c011, c010, c000, c001, c000, c011, c001, c010, c000, d.r, d.g, d.b, d.g, d.b);
    } else if (d.r > d.g && d.b > d.g) {
        const rgbvec c001 = LUTAT(l, p[0], p[1], n[2]), c100 = LUTAT(l, n[0], p[1], p[2]), c101 = LUTAT(l, n[0], p[1], n[2]);
        c.r = PYR(r, c100, c000, c111, c101, c001, c000, c101, c001, c100, c000, d.r, d.g, d.b, d.r, d.b);
        c.g = PYR(g, c100, c000, c111, c101, c001, c000, c101, c001, c100, c000, d.r, d.g, d.b, d.r, d.b);
        c.b = PYR(b, c100, c000, c111, c101, c001, c000, c101, c001, c100, c000, d.r, d.g, d.b, d.r, d.b);
    } else {
        const rgbvec c010 = LUTAT(l, p[0], n[1], p[2]), c110 = LUTAT(l, n[0], n[1], p[2]), c100 = LUTAT(l, n[0], p[1], p[2]);
        c.r = PYR(r, c100, c000, c010, c000, c111, c110, c110, c100, c010, c000, d.r, d.g, d.b, d.r, d.g);
        c.g = PYR(g, c100, c000, c010, c000, c111, c110, c110, c100, c010, c000, d.r, d.g, d.b, d.r, d.g);
        c.b = PYR(b, c100, c000, c010, c000, c111, c110, c110, c100, c010, c000, d.r, d.g, d.b, d.r, d.g);
    }
    return c;
}

#define PRISM(ch, A1, A0, B1, B0, D3, D2, D1, D0, E3, E2, E1, E0)                                                              \
    (c000.ch + (A1.ch - A0.ch) * d.b + (B1.ch - B0.ch) * d.r + (c010.ch - c000.ch) * d.g +                                    \
     (D3.ch - D2.ch - D1.ch + D0.ch) * d.b * d.g + (E3.ch - E2.ch - E1.ch + E0.ch) * d.r * d.g)

static inline rgbvec interp_prism(const OrcLut *l, const rgbvec *s)
{
    const int p[3] = {PREV(s->r), PREV(s->g), PREV(s->b)};
    const int n[3] = {NEXT(s->r, l), NEXT(s->g, l), NEXT(s->b, l)};
    const rgbvec d = {s->r - p[0], s->g - p[1], s->b - p[2]};
    const rgbvec c000 = LUTAT(l, p[0], p[1], p[2]), c010 = LUTAT(l, p[0], n[1], p[2]);
    const rgbvec c101 = LUTAT(l, n[0], p[1], n[2]), c111 = LUTAT(l, n[0], n[1], n[2]);
    rgbvec c;
    if (d.b > d.r) {
        const rgbvec c001 = LUTAT(l, p[0], p[1], n[2]), c011 = LUTAT(l, p[0], n[1], n[2]);
        c.r = PRISM(r, c001, c000, c101, c001, c000, c010, c001, c011, c001, c011, c101, c111);
        c.g = PRISM(g, c001, c000, c101, c001, c000, c010, c001, c011, c001, c011, c101, c111);
        c.b = PRISM(b, c001, c000, c101, c001, c000, c010, c001, c011, c001, c011, c101, c111);
    } else {
        const rgbvec c110 = LUTAT(l, n[0], n[1], p[2]), c100 = LUTAT(l, n[0], p[1], p[2]);
        c.r = PRISM(r, c101, c100, c100, c000, c100, c110, c101, c111, c000, c010, c100, c110);
        c.g = PRISM(g, c101, c100, c100, c000, c100, c110, c101, c111, c000, c010, c100, c110);
        c.b = PRISM(b, c101, c100, c100, c000, c100, c110, c101, c111, c000, c010, c100, c110);
    }
    return c;
}

static inline rgbvec interp_any(const OrcLut *l, int interp, const rgbvec *s)
{
    switch (interp) {
    case ORC_INTERP_NEAREST:
        return interp_nearest(l, s);
    case ORC_INTERP_TRILINEAR:
        return interp_trilinear(l, s);
    case ORC_INTERP_PYRAMID:
        return interp_pyramid(l, s);
    case ORC_INTERP_PRISM:
        return interp_prism(l, s);
    default:
        return interp_tetrahedral(l, s);
    }
}

static inline int trunc_clip(float v, int depth)
{
    /* av_clip_uintp2((int)v, depth) with x86 float->int conversion semantics (cvttss2si: NaN/overflow -> INT_MIN) */
    int a;
    if (!(v > -2147483904.f && v < 2147483648.f))
        a = INT32_MIN;
    else
        a = (int)v;
    const int mask = (1 << depth) - 1;
    if (a & ~mask)
        return (~a) >> 31 & mask;
    return a;
}

/* One pixel, integer in / integer out: the body of DEFINE_INTERP_FUNC[_PLANAR](name, nbits[, depth]) */
static inline void lut3d_pixel(const OrcLut *l, int interp, int depth, int ri, int gi, int bi, int *ro, int *go, int *bo)
{
    const float lut_max = (float)(l->size - 1);
    const float scale_f = 1.0f / ((1 << depth) - 1);
    const float scale_r = l->scale[0] * lut_max, scale_g = l->scale[1] * lut_max, scale_b = l->scale[2] * lut_max;
    const rgbvec rgb = {ri * scale_f, gi * scale_f, bi * scale_f};
    const rgbvec scaled = {clipf(rgb.r * scale_r, 0, lut_max), clipf(rgb.g * scale_g, 0, lut_max),
                           clipf(rgb.b * scale_b, 0, lut_max)};
    const rgbvec v = interp_any(l, interp, &scaled);
    const float om = (float)((1 << depth) - 1);
    *ro = trunc_clip(v.r * om, depth);
    *go = trunc_clip(v.g * om, depth);
    *bo = trunc_clip(v.b * om, depth);
}

/* packed rgb24, in place or out of place: interp_8_<name> */
ORC_API void orc_lut3d_rgb24(const OrcLut *l, int interp, const uint8_t *src, int ss, uint8_t *dst, int ds, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int r, g, b;
            const uint8_t *p = src + y * ss + 3 * x;
            lut3d_pixel(l, interp, 8, p[0], p[1], p[2], &r, &g, &b);
            uint8_t *o = dst + y * ds + 3 * x;
            o[0] = (uint8_t)r, o[1] = (uint8_t)g, o[2] = (uint8_t)b;
        }
}

/* planar 9..16-bit in uint16 containers: interp_16_<name>_p<depth>; strides in elements */
ORC_API void orc_lut3d_planar16(const OrcLut *l, int interp, int depth, const uint16_t *R, const uint16_t *G,
                                const uint16_t *B, int is, uint16_t *Ro, uint16_t *Go, uint16_t *Bo, int os, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int r, g, b;
            lut3d_pixel(l, interp, depth, R[y * is + x], G[y * is + x], B[y * is + x], &r, &g, &b);
            Ro[y * os + x] = (uint16_t)r, Go[y * os + x] = (uint16_t)g, Bo[y * os + x] = (uint16_t)b;
        }
}

/* ------------------------------------------------------------------------------------------------ */
/* composite stages = what one FFmpeg filter chain of the reference does to one frame                */
/* ------------------------------------------------------------------------------------------------ */

/* cfg 3/4: scale=in_color_matrix=M:out_color_matrix=M,lut3d=..,format=yuv422p10le  (ffmpeg.py:235,246,309) */
ORC_API int orc_stage_yuv422p10(const OrcLut *l, int interp, int in_matrix, int in_full_range, int out_matrix,
                                const uint16_t *Y, const uint16_t *U, const uint16_t *V, uint16_t *Yo, uint16_t *Uo,
                                uint16_t *Vo, int w, int h)
{
    size_t n = (size_t)w * h;
    uint16_t *rgb = (uint16_t *)malloc(sizeof(uint16_t) * 3 * n);
    if (!rgb)
        return ORC_ENOMEM;
    uint16_t *R = rgb, *G = rgb + n, *B = rgb + 2 * n;
    orc_yuv422p10_to_gbrp10(Y, w, U, V, w / 2, R, G, B, w, w, h, in_matrix, in_full_range);
    orc_lut3d_planar16(l, interp, 10, R, G, B, w, R, G, B, w, w, h);
    orc_gbrp10_to_yuv422p10(R, G, B, w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    free(rgb);
    return ORC_OK;
}

/* pro-mode stage 1 on a 10-bit 4:2:0 source: (auto-)scale -> gbrp10le, lut3d, format=yuv422p10le (ffmpeg.py:417-433). */
ORC_API int orc_stage_yuv420p10_to_yuv422p10(const OrcLut *l, int interp, int in_matrix, int in_full_range, int out_matrix,
                                             const uint16_t *Y, const uint16_t *U, const uint16_t *V, uint16_t *Yo, uint16_t *Uo,
                                             uint16_t *Vo, int w, int h)
{
    size_t n = (size_t)w * h;
    uint16_t *rgb = (uint16_t *)malloc(3 * n * sizeof(uint16_t));
    if (!rgb)
        return ORC_ENOMEM;
    uint16_t *R = rgb, *G = rgb + n, *B = rgb + 2 * n;
    orc_yuv420p10_to_gbrp10(Y, w, U, V, w / 2, R, G, B, w, w, h, in_matrix, in_full_range);
    orc_lut3d_planar16(l, interp, 10, R, G, B, w, R, G, B, w, w, h);
    orc_gbrp10_to_yuv422p10(R, G, B, w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    free(rgb);
    return ORC_OK;
}

/* rgb24 -> yuv422p10le: the return leg of the pro-mode master when the source is 8 bit (lut3d works on rgb24, the
 * `format=yuv422p10le` behind it is an auto-inserted scaler; ffmpeg.py:417-433, 309-310).  PINNED against libswscale
 * 9.1.100 (tests/test_oracle_swscale.py): rgb24ToY / rgb24ToUV_half as in orc_rgb24_to_yuv420p, no vertical filter
 * (4:2:2 keeps every row), 15-bit intermediates 2*y14 / 2*c14 straight into the 10-bit output stage (+16) >> 5. */
ORC_API void orc_rgb24_to_yuv422p10(const uint8_t *rgb, int rs, uint16_t *Y, int ys, uint16_t *U, uint16_t *V, int cs,
                                    int w, int h, int matrix_id)
{
    int32_t c[9];
    orc_rgb2yuv_setup(matrix_id, c);
    const int cw = w / 2;
    for (int y = 0; y < h; y++) {
        const uint8_t *p = rgb + y * rs;
        for (int x = 0; x < w; x++) {
            int r = p[3 * x], g = p[3 * x + 1], b = p[3 * x + 2];
            int y14 = (c[0] * r + c[1] * g + c[2] * b + (32 << 14) + (1 << 8)) >> 9;
            Y[y * ys + x] = (uint16_t)clip_u((2 * y14 + 16) >> 5, 1023);
        }
        for (int i = 0; i < cw; i++) {
            int r = p[6 * i] + p[6 * i + 3], g = p[6 * i + 1] + p[6 * i + 4], b = p[6 * i + 2] + p[6 * i + 5];
            int u14 = (c[3] * r + c[4] * g + c[5] * b + (256 << 15) + (1 << 9)) >> 10;
            int v14 = (c[6] * r + c[7] * g + c[8] * b + (256 << 15) + (1 << 9)) >> 10;
            U[y * cs + i] = (uint16_t)clip_u((2 * u14 + 16) >> 5, 1023);
            V[y * cs + i] = (uint16_t)clip_u((2 * v14 + 16) >> 5, 1023);
        }
    }
}

/* pro-mode stage 1 on an 8-bit 4:2:0 source: [range pass,] (auto-)scale -> rgb24, lut3d, format=yuv422p10le
 * (ffmpeg.py:417-433 with an 8-bit VideoInfo; plan_colour_stage: in yuv420p, out yuv422p10le, rgb24 intermediate). */
ORC_API int orc_stage_yuv420p_to_yuv422p10(const OrcLut *l, int interp, int range_pass, int in_matrix, int in_full_range,
                                           int out_matrix, const uint8_t *Y, const uint8_t *U, const uint8_t *V, uint16_t *Yo,
                                           uint16_t *Uo, uint16_t *Vo, int w, int h)
{
    size_t n = (size_t)w * h;
    uint8_t *rgb = (uint8_t *)malloc(3 * n);
    uint8_t *tmp = NULL;
    if (!rgb)
        return ORC_ENOMEM;
    if (range_pass) {
        tmp = (uint8_t *)malloc(n + n / 2);
        if (!tmp) {
            free(rgb);
            return ORC_ENOMEM;
        }
        orc_range_pc_to_tv_8(Y, w, tmp, w, w, h, 0);
        orc_range_pc_to_tv_8(U, w / 2, tmp + n, w / 2, w / 2, h / 2, 1);
        orc_range_pc_to_tv_8(V, w / 2, tmp + n + n / 4, w / 2, w / 2, h / 2, 1);
        Y = tmp, U = tmp + n, V = tmp + n + n / 4;
    }
    orc_yuv420p_to_rgb24(Y, w, U, V, w / 2, rgb, 3 * w, w, h, in_matrix, in_full_range);
    orc_lut3d_rgb24(l, interp, rgb, 3 * w, rgb, 3 * w, w, h);
    orc_rgb24_to_yuv422p10(rgb, 3 * w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    free(rgb);
    free(tmp);
    return ORC_OK;
}

/* cfg 1/2: [scale=in_range=pc:out_range=tv..,format=yuv420p,] (auto-)scale,lut3d=..,format=yuv420p
 * (ffmpeg.py:225-235,246,309).  range_pass!=0 applies the 8-bit pc->tv pass first (re-quantised to u8). */
ORC_API int orc_stage_yuv420p(const OrcLut *l, int interp, int range_pass, int in_matrix, int in_full_range,
                              int out_matrix, const uint8_t *Y, const uint8_t *U, const uint8_t *V, uint8_t *Yo,
                              uint8_t *Uo, uint8_t *Vo, int w, int h)
{
    size_t n = (size_t)w * h;
    uint8_t *rgb = (uint8_t *)malloc(3 * n);
    uint8_t *tmp = NULL;
    if (!rgb)
        return ORC_ENOMEM;
    if (range_pass) {
        tmp = (uint8_t *)malloc(n + n / 2);
        if (!tmp) {
            free(rgb);
            return ORC_ENOMEM;
        }
        orc_range_pc_to_tv_8(Y, w, tmp, w, w, h, 0);
        orc_range_pc_to_tv_8(U, w / 2, tmp + n, w / 2, w / 2, h / 2, 1);
        orc_range_pc_to_tv_8(V, w / 2, tmp + n + n / 4, w / 2, w / 2, h / 2, 1);
        Y = tmp, U = tmp + n, V = tmp + n + n / 4;
    }
    orc_yuv420p_to_rgb24(Y, w, U, V, w / 2, rgb, 3 * w, w, h, in_matrix, in_full_range);
    orc_lut3d_rgb24(l, interp, rgb, 3 * w, rgb, 3 * w, w, h);
    orc_rgb24_to_yuv420p(rgb, 3 * w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    free(rgb);
    free(tmp);
    return ORC_OK;
}

/* 10-bit chains through gbrp10le with either chroma layout on each side: in_420 / out_420 select yuv420p10le, otherwise
 * yuv422p10le.  out 4:2:0 = libx265 / hevc_videotoolbox "preserve" outputs (ffmpeg.py:297-298), out 4:2:2 = pro-mode master. */
ORC_API int orc_stage_10bit(const OrcLut *l, int interp, int in_420, int out_420, int in_matrix, int in_full_range,
                            int out_matrix, const uint16_t *Y, const uint16_t *U, const uint16_t *V, uint16_t *Yo, uint16_t *Uo,
                            uint16_t *Vo, int w, int h)
{
    size_t n = (size_t)w * h;
    uint16_t *rgb = (uint16_t *)malloc(3 * n * sizeof(uint16_t));
    if (!rgb)
        return ORC_ENOMEM;
    uint16_t *R = rgb, *G = rgb + n, *B = rgb + 2 * n;
    if (in_420)
        orc_yuv420p10_to_gbrp10(Y, w, U, V, w / 2, R, G, B, w, w, h, in_matrix, in_full_range);
    else
        orc_yuv422p10_to_gbrp10(Y, w, U, V, w / 2, R, G, B, w, w, h, in_matrix, in_full_range);
    orc_lut3d_planar16(l, interp, 10, R, G, B, w, R, G, B, w, w, h);
    if (out_420)
        orc_gbrp10_to_yuv420p10(R, G, B, w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    else
        orc_gbrp10_to_yuv422p10(R, G, B, w, Yo, w, Uo, Vo, w / 2, w, h, out_matrix);
    free(rgb);
    return ORC_OK;
}

/* ================================================================================================ */
/* The whole format matrix of the operator (reference ffmpeg.py:137-143 intermediate formats, :287-310 output formats,  */
/* :213-223 output range): generic legs and one composite, orc_stage_any.  Every leg below is PINNED against libswscale   */
/* 9.1.100 in tests/test_oracle_swscale.py (test_any_*).                                                                  */
/* ================================================================================================ */

/* pixel format ids (= lutr_pix_fmt in include/lutr.h) */
enum { ORC_PIX_YUV420P = 0, ORC_PIX_YUV422P10 = 1, ORC_PIX_YUV420P10 = 2, ORC_PIX_YUV422P = 3, ORC_PIX_YUV444P = 4,
       ORC_PIX_YUV444P10 = 5, ORC_PIX_COUNT = 6 };

ORC_API int orc_pix_depth(int fmt) { return (fmt == ORC_PIX_YUV422P10 || fmt == ORC_PIX_YUV420P10 || fmt == ORC_PIX_YUV444P10) ? 10 : 8; }
/* log2 chroma subsampling: sx | sy << 1 */
ORC_API int orc_pix_sub(int fmt)
{
    switch (fmt) {
    case ORC_PIX_YUV420P: case ORC_PIX_YUV420P10: return 1 | 2;
    case ORC_PIX_YUV422P: case ORC_PIX_YUV422P10: return 1;
    default: return 0;
    }
}

/* libswscale >= 7.1 range conversion on the 15-bit intermediates (swscale.c lumRangeTo/FromJpeg_c with the constants of
 * init_range_convert_constants / solve_range_convert), restated and PINNED: for an output of `depth` bits
 *   coeff  = ceil(((dst_max - dst_min) << 14) / (src_max - src_min))
 *   offset = (dst_max << (29 - depth)) - (src_max << (15 - depth)) * coeff + (1 << 13)
 *   out15  = min((in15 * coeff + offset) >> 14, 32767)
 * with mpeg = [16, 235 | 240] << (depth - 8) and jpeg = [0, 2^depth - 1].  which: 0 luma, 1 chroma. */
ORC_API void orc_range_coeffs(int depth, int to_jpeg, int which, int64_t *coeff, int64_t *offset)
{
    const int ss = 15 - depth, ts = 14 + ss;
    const int64_t mpeg_min = 16 << (depth - 8), mpeg_max = (which ? 240 : 235) << (depth - 8), jpeg_max = (1 << depth) - 1;
    const int64_t smin = to_jpeg ? mpeg_min : 0, smax = to_jpeg ? mpeg_max : jpeg_max;
    const int64_t dmin = to_jpeg ? 0 : mpeg_min, dmax = to_jpeg ? jpeg_max : mpeg_max;
    const int64_t num = (dmax - dmin) << ts, den = (smax - smin) << ss;
    *coeff = (num + den - 1) / den;
    *offset = (dmax << ts) - (smax << ss) * *coeff + (1 << 13);
}

static inline int range_conv15(int v, int64_t coeff, int64_t offset)
{
    int64_t o = ((int64_t)v * coeff + offset) >> 14;
    return (int)(o > 32767 ? 32767 : o);
}

/* 8-bit sources -> rgb24.  4:2:0 / 4:2:2: the x86 SIMD path of orc_yuv420p_to_rgb24 (chroma replicated);
 * 4:4:4: swscale's C path (8-bit samples lifted to 15 bits, yuv2rgb_full: 30-bit clip, >> 22). */
ORC_API void orc_yuv8_to_rgb24(int sub, const uint8_t *Y, int ys, const uint8_t *U, const uint8_t *V, int cs, uint8_t *rgb, int rs,
                               int w, int h, int matrix_id, int full_range)
{
    OrcYuv2Rgb t;
    orc_yuv2rgb_setup(matrix_id, full_range, &t);
    const int sx = sub & 1, sy = sub >> 1;
    for (int y = 0; y < h; y++) {
        const uint8_t *yr = Y + y * ys, *ur = U + (y >> sy) * cs, *vr_ = V + (y >> sy) * cs;
        uint8_t *o = rgb + y * rs;
        for (int x = 0; x < w; x++) {
            if (sx) {
                int yy = ((yr[x] * 8 - t.yO8) * t.yC) >> 16;
                int u = ur[x >> 1] * 8 - 1024, v = vr_[x >> 1] * 8 - 1024;
                o[3 * x + 0] = (uint8_t)clip_u(yy + ((v * t.vr) >> 16), 255);
                o[3 * x + 1] = (uint8_t)clip_u(yy + ((u * t.ug) >> 16) + ((v * t.vg) >> 16), 255);
                o[3 * x + 2] = (uint8_t)clip_u(yy + ((u * t.ub) >> 16), 255);
            } else {
                uint32_t yy = (uint32_t)(yr[x] * 512 - t.yOff10) * (uint32_t)t.yC + (1u << 21);
                uint32_t u = (uint32_t)((ur[x] - 128) * 512), v = (uint32_t)((vr_[x] - 128) * 512);
                int32_t r = (int32_t)(yy + v * (uint32_t)t.vr), g = (int32_t)(yy + v * (uint32_t)t.vg + u * (uint32_t)t.ug);
                int32_t b = (int32_t)(yy + u * (uint32_t)t.ub);
                if ((r | g | b) & (int32_t)0xC0000000) {
                    r = r & ~0x3FFFFFFF ? (~r) >> 31 & 0x3FFFFFFF : r;
                    g = g & ~0x3FFFFFFF ? (~g) >> 31 & 0x3FFFFFFF : g;
                    b = b & ~0x3FFFFFFF ? (~b) >> 31 & 0x3FFFFFFF : b;
                }
                o[3 * x + 0] = (uint8_t)(r >> 22), o[3 * x + 1] = (uint8_t)(g >> 22), o[3 * x + 2] = (uint8_t)(b >> 22);
            }
        }
    }
}

/* yuv444p10le -> gbrp10le: orc_yuv422p10_to_gbrp10 without the chroma filter (c15 = C << 5, u17 = (C - 512) * 128) */
ORC_API void orc_yuv444p10_to_gbrp10(const uint16_t *Y, const uint16_t *U, const uint16_t *V, int is, uint16_t *R, uint16_t *G,
                                     uint16_t *B, int os, int w, int h, int matrix_id, int full_range)
{
    OrcYuv2Rgb t;
    orc_yuv2rgb_setup(matrix_id, full_range, &t);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t u17 = (uint32_t)((U[y * is + x] - 512) * 128), v17 = (uint32_t)((V[y * is + x] - 512) * 128);
            uint32_t yy = (uint32_t)(Y[y * is + x] * 128 - t.yOff10) * (uint32_t)t.yC + (1u << 19);
            R[y * os + x] = clip30_shift20((int32_t)(yy + v17 * (uint32_t)t.vr));
            G[y * os + x] = clip30_shift20((int32_t)(yy + v17 * (uint32_t)t.vg + u17 * (uint32_t)t.ug));
            B[y * os + x] = clip30_shift20((int32_t)(yy + u17 * (uint32_t)t.ub));
        }
}

/* ff_dither_8x8_128: swscale's ordered dither for 8-bit outputs of > 8-bit sources (recovered bit-exactly from the binary) */
static const uint8_t k_dither_8x8_128[8][8] = {
    {36, 68, 60, 92, 34, 66, 58, 90},   {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74},
    {116, 20, 108, 12, 114, 18, 106, 10}, {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30},
    {48, 80, 40, 72, 54, 86, 46, 78},   {112, 16, 104, 8, 118, 22, 110, 14},
};

/* RGB (src_depth 8 = rgb24 values, 10 = gbrp10le values; planes of uint16, stride `is` elements) -> any supported YUV output.
 *   15-bit intermediates:  src 8: y15 = 2*y14 (rgb24ToY), chroma rgb24ToUV (4:4:4) or rgb24ToUV_half pair sums (4:2:2 / 4:2:0);
 *                          src 10: y15 = min(2*y14, 32767) (planar_rgb16_to_y), chroma per column, then for 4:2:2 / 4:2:0 the
 *                          8-tap horizontal decimation (orc_gbrp10_to_yuv422p10), for 4:4:4 min(2*c14, 32767)
 *   out_full:              range conversion of the 15-bit values (orc_range_coeffs with the OUTPUT depth)
 *   output stage:          10 bit: (v + 16) >> 5;  8 bit: (v + d) >> 7 with d = 64, or ff_dither_8x8_128[y & 7][(x + off) & 7]
 *                          (off = 3 for V) when the RGB source is deeper than 8 bit
 *   4:2:0 vertical:        8-tap over luma rows 2j-3 .. 2j+4, edge taps folded; 10 bit: exact, (sum + (1 << 16)) >> 17;
 *                          8 bit: x86 yuv2yuvX: ((d + 56) >> 4) + sum((c15 * tap) >> 16)) >> 3, except the last chroma row
 *                          (C code): ((d << 12) + sum(c15 * tap)) >> 19 */
ORC_API void orc_rgb_to_yuv_any(int src_depth, int out_fmt, int out_full, const uint16_t *R, const uint16_t *G, const uint16_t *B,
                                int is, void *Yo, void *Uo, void *Vo, int w, int h, int matrix_id)
{
    int32_t c[9];
    orc_rgb2yuv_setup(matrix_id, c);
    const int od = orc_pix_depth(out_fmt), sub = orc_pix_sub(out_fmt), sx = sub & 1, sy = sub >> 1;
    const int cw = w >> sx, ch = h >> sy, omax = (1 << od) - 1;
    const int bayer = src_depth > 8 && od == 8;
    int64_t lc, lo, cc, co;
    orc_range_coeffs(od, 1, 0, &lc, &lo);
    orc_range_coeffs(od, 1, 1, &cc, &co);
    int32_t *u14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w), *v14 = (int32_t *)malloc(sizeof(int32_t) * (size_t)w);
    int32_t *hu = (int32_t *)malloc(sizeof(int32_t) * (size_t)cw * h), *hv = (int32_t *)malloc(sizeof(int32_t) * (size_t)cw * h);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const int r = R[y * is + x], g = G[y * is + x], b = B[y * is + x];
            int y15;
            if (src_depth == 8) {
                y15 = 2 * ((c[0] * r + c[1] * g + c[2] * b + (32 << 14) + (1 << 8)) >> 9);
                u14[x] = (c[3] * r + c[4] * g + c[5] * b + (256 << 14) + (1 << 8)) >> 9;
                v14[x] = (c[6] * r + c[7] * g + c[8] * b + (256 << 14) + (1 << 8)) >> 9;
            } else {
                y15 = 2 * ((c[0] * r + c[1] * g + c[2] * b + (16 << 17) + (1 << 10)) >> 11);
                if (y15 > 32767)
                    y15 = 32767;
                u14[x] = (c[3] * r + c[4] * g + c[5] * b + (128 << 17) + (1 << 10)) >> 11;
                v14[x] = (c[6] * r + c[7] * g + c[8] * b + (128 << 17) + (1 << 10)) >> 11;
            }
            if (out_full)
                y15 = range_conv15(y15, lc, lo);
            const int d = bayer ? k_dither_8x8_128[y & 7][x & 7] : 64;
            if (od == 8)
                ((uint8_t *)Yo)[(size_t)y * w + x] = (uint8_t)clip_u((y15 + d) >> 7, 255);
            else
                ((uint16_t *)Yo)[(size_t)y * w + x] = (uint16_t)clip_u((y15 + 16) >> 5, omax);
        }
        for (int i = 0; i < cw; i++) {
            int64_t su, sv;
            if (!sx) {
                su = 2 * u14[i], sv = 2 * v14[i];
            } else if (src_depth == 8) {
                const int r = R[y * is + 2 * i] + R[y * is + 2 * i + 1], g = G[y * is + 2 * i] + G[y * is + 2 * i + 1];
                const int b = B[y * is + 2 * i] + B[y * is + 2 * i + 1];
                su = 2 * ((c[3] * r + c[4] * g + c[5] * b + (256 << 15) + (1 << 9)) >> 10);
                sv = 2 * ((c[6] * r + c[7] * g + c[8] * b + (256 << 15) + (1 << 9)) >> 10);
            } else {
                su = sv = 0;
                for (int k = 0; k < 8; k++) {
                    const int x = clampi(2 * i - 3 + k, 0, w - 1);
                    su += (int64_t)u14[x] * k_h422_taps[k];
                    sv += (int64_t)v14[x] * k_h422_taps[k];
                }
                su >>= 13, sv >>= 13;
            }
            int a = (int)(su > 32767 ? 32767 : su), b_ = (int)(sv > 32767 ? 32767 : sv);
            if (out_full)
                a = range_conv15(a, cc, co), b_ = range_conv15(b_, cc, co);
            hu[(size_t)y * cw + i] = a, hv[(size_t)y * cw + i] = b_;
        }
    }
    for (int j = 0; j < ch; j++) {
        int rows[8], coef[8], n = 0;
        if (sy)
            for (int k = 0; k < 8; k++) { /* fold taps whose source row lies outside the frame onto the edge row */
                const int row = clampi(2 * j - 3 + k, 0, h - 1);
                if (n && rows[n - 1] == row)
                    coef[n - 1] += k_v420_taps[k];
                else
                    rows[n] = row, coef[n] = k_v420_taps[k], n++;
            }
        for (int i = 0; i < cw; i++) {
            const int du = bayer ? k_dither_8x8_128[j & 7][i & 7] : 64, dv = bayer ? k_dither_8x8_128[j & 7][(i + 3) & 7] : 64;
            int ou, ov;
            if (!sy) {
                const int a = hu[(size_t)j * cw + i], b_ = hv[(size_t)j * cw + i];
                ou = od == 8 ? (a + du) >> 7 : (a + 16) >> 5, ov = od == 8 ? (b_ + dv) >> 7 : (b_ + 16) >> 5;
            } else if (od == 10) {
                int64_t su = 1 << 16, sv = 1 << 16;
                for (int k = 0; k < n; k++)
                    su += (int64_t)hu[(size_t)rows[k] * cw + i] * coef[k], sv += (int64_t)hv[(size_t)rows[k] * cw + i] * coef[k];
                ou = (int)(su >> 17), ov = (int)(sv >> 17);
            } else if (j == ch - 1) { /* swscale drops to the C vertical scaler for the last output rows */
                int su = du << 12, sv = dv << 12;
                for (int k = 0; k < n; k++)
                    su += hu[(size_t)rows[k] * cw + i] * coef[k], sv += hv[(size_t)rows[k] * cw + i] * coef[k];
                ou = su >> 19, ov = sv >> 19;
            } else { /* x86 yuv2yuvX: per-tap truncation (pmulhw) */
                int su = (du + 56) >> 4, sv = (dv + 56) >> 4;
                for (int k = 0; k < n; k++)
                    su += (hu[(size_t)rows[k] * cw + i] * coef[k]) >> 16, sv += (hv[(size_t)rows[k] * cw + i] * coef[k]) >> 16;
                ou = su >> 3, ov = sv >> 3;
            }
            if (od == 8)
                ((uint8_t *)Uo)[(size_t)j * cw + i] = (uint8_t)clip_u(ou, 255), ((uint8_t *)Vo)[(size_t)j * cw + i] = (uint8_t)clip_u(ov, 255);
            else
                ((uint16_t *)Uo)[(size_t)j * cw + i] = (uint16_t)clip_u(ou, omax), ((uint16_t *)Vo)[(size_t)j * cw + i] = (uint16_t)clip_u(ov, omax);
        }
    }
    free(u14), free(v14), free(hu), free(hv);
}

/* Full-range sources deeper than 8 bit: the reference's range pass always lands on 8-bit yuv4xxp (`scale=in_range=pc:out_range=R,
 * format=yuv4xxp`, ffmpeg.py:211-233), i.e. swscale converts yuv4xxp10le(pc) -> yuv4xxp 8 bit in one pass.  PINNED against
 * libswscale 9.1.100 (tests/test_oracle_swscale.py test_any_10bit_full_range_prepass):
 *   mode 1 (out_range tv): generic scaler path: 15-bit sample (v << 5) -> range conversion (8-bit constants, from jpeg) -> 8-bit
 *          output stage with the ordered dither ff_dither_8x8_128[y & 7][(x + off) & 7], off = 3 for the V plane
 *   mode 2 (out_range pc): swscale's planar copy with its 2x2 dither {{1,2},{3,0}}: chroma t = (v + d) >> 2, t - (t >> 8);
 *          full-range luma min((v + d - (v >> 8)) >> 2, 255)        (recovered exhaustively from the binary)
 * plane: 0 Y, 1 U, 2 V; coordinates are those of the plane itself. */
ORC_API void orc_prepass_10_to_8(int mode, int plane, const uint16_t *src, int ss, uint8_t *dst, int ds, int w, int h)
{
    static const int d2[2][2] = {{1, 2}, {3, 0}};
    int64_t c, o;
    orc_range_coeffs(8, 0, plane > 0, &c, &o);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int v = src[y * ss + x];
            int r;
            if (mode == 1) {
                const int v15 = range_conv15(v << 5, c, o);
                r = clip_u((v15 + k_dither_8x8_128[y & 7][(x + (plane == 2 ? 3 : 0)) & 7]) >> 7, 255);
            } else {
                const int d = d2[y & 1][x & 1];
                if (plane == 0) {
                    r = (v + d - (v >> 8)) >> 2;
                    if (r > 255)
                        r = 255;
                } else {
                    const int t = (v + d) >> 2;
                    r = t - (t >> 8);
                }
            }
            dst[y * ds + x] = (uint8_t)r;
        }
}

/* Any supported source -> [8-bit pc->tv pass] -> RGB (rgb24 for 8-bit, gbrp10le for 10-bit sources) -> lut3d -> any supported
 * output, one frame, tightly packed planes.  range_pass: 8-bit sources: 1 = pc -> tv pass; 10-bit sources: 1 / 2 = the reference's
 * 8-bit re-quantisation of a full-range source with / without the range conversion (orc_prepass_10_to_8), then the 8-bit chain.
 * in_full_range = range of the samples the YUV->RGB step sees. */
ORC_API int orc_stage_any(const OrcLut *l, int interp, int in_fmt, int out_fmt, int range_pass, int in_matrix, int in_full_range,
                          int out_matrix, int out_full_range, const void *Y, const void *U, const void *V, void *Yo, void *Uo,
                          void *Vo, int w, int h)
{
    if (in_fmt < 0 || in_fmt >= ORC_PIX_COUNT || out_fmt < 0 || out_fmt >= ORC_PIX_COUNT)
        return ORC_EINVAL;
    const int id = orc_pix_depth(in_fmt), sub = orc_pix_sub(in_fmt), cw = w >> (sub & 1), ch = h >> (sub >> 1);
    const size_t n = (size_t)w * h, cn = (size_t)cw * ch;
    uint16_t *rgb = (uint16_t *)malloc(3 * n * sizeof(uint16_t));
    if (!rgb)
        return ORC_ENOMEM;
    uint16_t *R = rgb, *G = rgb + n, *B = rgb + 2 * n;
    uint8_t *pre = NULL;
    int wd = id; /* working depth: 8 = rgb24 intermediate, 10 = gbrp10le */
    if (id == 10 && range_pass) { /* full-range 10-bit source: the reference's pass re-quantises to 8-bit yuv4xxp first */
        pre = (uint8_t *)malloc(n + 2 * cn);
        orc_prepass_10_to_8(range_pass, 0, (const uint16_t *)Y, w, pre, w, w, h);
        orc_prepass_10_to_8(range_pass, 1, (const uint16_t *)U, cw, pre + n, cw, cw, ch);
        orc_prepass_10_to_8(range_pass, 2, (const uint16_t *)V, cw, pre + n + cn, cw, cw, ch);
        Y = pre, U = pre + n, V = pre + n + cn;
        wd = 8, range_pass = 0;
    }
    if (wd == 8) {
        uint8_t *p8 = (uint8_t *)malloc(3 * n), *tmp = NULL;
        const uint8_t *y8 = (const uint8_t *)Y, *u8 = (const uint8_t *)U, *v8 = (const uint8_t *)V;
        if (range_pass) {
            tmp = (uint8_t *)malloc(n + 2 * cn);
            orc_range_pc_to_tv_8(y8, w, tmp, w, w, h, 0);
            orc_range_pc_to_tv_8(u8, cw, tmp + n, cw, cw, ch, 1);
            orc_range_pc_to_tv_8(v8, cw, tmp + n + cn, cw, cw, ch, 1);
            y8 = tmp, u8 = tmp + n, v8 = tmp + n + cn;
        }
        orc_yuv8_to_rgb24(sub, y8, w, u8, v8, cw, p8, 3 * w, w, h, in_matrix, in_full_range);
        orc_lut3d_rgb24(l, interp, p8, 3 * w, p8, 3 * w, w, h);
        for (size_t i = 0; i < n; i++)
            R[i] = p8[3 * i], G[i] = p8[3 * i + 1], B[i] = p8[3 * i + 2];
        free(p8), free(tmp);
    } else {
        const uint16_t *y16 = (const uint16_t *)Y, *u16 = (const uint16_t *)U, *v16 = (const uint16_t *)V;
        if (in_fmt == ORC_PIX_YUV420P10)
            orc_yuv420p10_to_gbrp10(y16, w, u16, v16, cw, R, G, B, w, w, h, in_matrix, in_full_range);
        else if (in_fmt == ORC_PIX_YUV422P10)
            orc_yuv422p10_to_gbrp10(y16, w, u16, v16, cw, R, G, B, w, w, h, in_matrix, in_full_range);
        else
            orc_yuv444p10_to_gbrp10(y16, u16, v16, w, R, G, B, w, w, h, in_matrix, in_full_range);
        orc_lut3d_planar16(l, interp, 10, R, G, B, w, R, G, B, w, w, h);
    }
    orc_rgb_to_yuv_any(wd, out_fmt, out_full_range, R, G, B, w, Yo, Uo, Vo, w, h, out_matrix);
    free(rgb);
    free(pre);
    return ORC_OK;
}

/* ---- pro mode stage 2 without a LUT (reference ffmpeg.py:459-471 + :287-310): the encoding ffmpeg converts the yuv422p10le master
 * to the delivery pix_fmt with an auto-inserted scaler, YUV -> YUV, no RGB in between, same matrix and range on both sides.
 * PINNED against libswscale 9.1.100 (tests/test_oracle_swscale.py test_convert_422p10_*):
 *   -> yuv420p    generic scaler: 15-bit samples v << 5; luma (y15 + ff_dither_8x8_128[y & 7][x & 7]) >> 7; chroma through the 8-tap
 *                 vertical decimation of orc_rgb_to_yuv_any's 8-bit output stage (x86 per-tap truncation, C code on the last row,
 *                 dither column offset 3 for V)
 *   -> yuv420p10le  luma unchanged ((v << 5) + 16) >> 5; chroma exact 8-tap vertical, (sum + (1 << 16)) >> 17
 *   -> yuv422p    swscale's planar copy (same subsampling): every plane min((v + d) >> 2, 255), d = {{1,2},{3,0}}[y & 1][x & 1]
 * Returns ORC_OK or -1 for an output layout that is not modelled. */
ORC_API int orc_convert_yuv422p10(int out_fmt, const uint16_t *Y, const uint16_t *U, const uint16_t *V, void *Yo, void *Uo, void *Vo,
                                  int w, int h)
{
    static const int d2[2][2] = {{1, 2}, {3, 0}};
    const int cw = w / 2;
    if (out_fmt == ORC_PIX_YUV422P) {
        const uint16_t *src[3] = {Y, U, V};
        uint8_t *dst[3] = {(uint8_t *)Yo, (uint8_t *)Uo, (uint8_t *)Vo};
        for (int p = 0; p < 3; p++) {
            const int pw = p ? cw : w;
            for (int y = 0; y < h; y++)
                for (int x = 0; x < pw; x++) {
                    const int r = (src[p][(size_t)y * pw + x] + d2[y & 1][x & 1]) >> 2;
                    dst[p][(size_t)y * pw + x] = (uint8_t)(r > 255 ? 255 : r);
                }
        }
        return ORC_OK;
    }
    if (out_fmt != ORC_PIX_YUV420P && out_fmt != ORC_PIX_YUV420P10)
        return -1;
    const int od = out_fmt == ORC_PIX_YUV420P ? 8 : 10, ch = h / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int y15 = Y[(size_t)y * w + x] << 5;
            if (od == 8)
                ((uint8_t *)Yo)[(size_t)y * w + x] = (uint8_t)clip_u((y15 + k_dither_8x8_128[y & 7][x & 7]) >> 7, 255);
            else
                ((uint16_t *)Yo)[(size_t)y * w + x] = (uint16_t)clip_u((y15 + 16) >> 5, 1023);
        }
    for (int j = 0; j < ch; j++) {
        int rows[8], coef[8], n = 0;
        for (int k = 0; k < 8; k++) { /* fold taps whose source row lies outside the frame onto the edge row */
            const int row = clampi(2 * j - 3 + k, 0, h - 1);
            if (n && rows[n - 1] == row)
                coef[n - 1] += k_v420_taps[k];
            else
                rows[n] = row, coef[n] = k_v420_taps[k], n++;
        }
        for (int p = 1; p < 3; p++) {
            const uint16_t *C = p == 1 ? U : V;
            for (int i = 0; i < cw; i++) {
                int o;
                if (od == 10) {
                    int64_t s = 1 << 16;
                    for (int k = 0; k < n; k++)
                        s += (int64_t)(C[(size_t)rows[k] * cw + i] << 5) * coef[k];
                    o = clip_u((int)(s >> 17), 1023);
                    ((uint16_t *)(p == 1 ? Uo : Vo))[(size_t)j * cw + i] = (uint16_t)o;
                    continue;
                }
                const int d = k_dither_8x8_128[j & 7][(i + (p == 2 ? 3 : 0)) & 7];
                if (j == ch - 1) {
                    int s = d << 12;
                    for (int k = 0; k < n; k++)
                        s += (C[(size_t)rows[k] * cw + i] << 5) * coef[k];
                    o = s >> 19;
                } else {
                    int s = (d + 56) >> 4;
                    for (int k = 0; k < n; k++)
                        s += ((C[(size_t)rows[k] * cw + i] << 5) * coef[k]) >> 16;
                    o = s >> 3;
                }
                ((uint8_t *)(p == 1 ? Uo : Vo))[(size_t)j * cw + i] = (uint8_t)clip_u(o, 255);
            }
        }
    }
    return ORC_OK;
}
