// Generic fused tile kernel: ANY supported source layout -> [8-bit pc->tv pass] -> RGB -> lut3d (any interpolation mode) -> ANY
// supported output layout / depth / range, in one launch.  It is the complete form of the operator the reference builds
// (src/lut_renderer/ffmpeg.py:137-143 range-pass intermediates yuv420p / yuv422p / yuv444p, :242-247 lut3d with the interp
// whitelist, :287-310 output formats, :213-223 + :348-383 output range), for every combination the specialised warp-window
// kernels (k422_warp.cu, k420_warp.cuh, k10_warp.cu) do not cover:
//
//   sources   yuv420p  yuv422p  yuv444p  (8 bit, RGB intermediate rgb24)     yuv420p10le  yuv422p10le  yuv444p10le  (gbrp10le)
//   outputs   yuv420p  yuv422p  yuv444p  yuv420p10le  yuv422p10le  yuv444p10le,  limited or full range,
//             ordered dither (ff_dither_8x8_128) when a 10-bit chain ends in an 8-bit format
//   interp    nearest  trilinear  tetrahedral  pyramid  prism,  per-channel LUT domain scales
//
// Arithmetic: oracle/lutr_oracle.c orc_stage_any (every swscale leg pinned bit-exactly against libswscale 9.1.100).
//
// Decomposition: a CTA owns a tile of 128 x 32 luma pixels of one frame.
//   phase A  one pixel PAIR per thread step over the tile plus the halo the output filters need (4 columns per side for the
//            8-tap horizontal decimation of the 10-bit legs, 3 rows above / 4 below for the 4:2:0 vertical decimation):
//            source leg straight from global memory (read-only path) -> lut3d -> luma out, per-pixel chroma terms to shared memory
//   phase B  per chroma column: pair sum / horizontal 8-tap decimation, range conversion; 4:2:2 / 4:4:4 outputs are written here
//   phase C  4:2:0 outputs: vertical 8-tap decimation of the rows phase B left in shared memory
// lut3d: tetrahedral / trilinear run the packed two-pixel evaluation of the warp-window kernels (warp_common.cuh lut3d_pair, general
// index path, coordinate and output clips kept); nearest / pyramid / prism the scalar restatement in lut3d_generic.cuh.
// This kernel favours coverage over speed (halo recomputation up to 29 %, run-time format switches, no register-level pipelining;
// 271 thread-instructions per pixel, issue-bound; built for 4 CTAs per SM with 64 registers); the headline configurations never
// reach it.
#pragma once
#include <cuda_runtime.h>

#include "lut3d_generic.cuh"
#include "lutr_internal.h"
#include "warp_common.cuh"

namespace lutr {

namespace {

constexpr int kTW = 128, kTH = 32;            // luma tile
constexpr int kCW = kTW + 8, kCH = kTH + 7;   // with halo: columns X0-4 .. X0+TW+3, rows Y0-3 .. Y0+TH+3
constexpr int kThreads = 256;
constexpr int kUpRows = (kCH + 1) / 2 + 4;     // chroma rows the 4-tap vertical upsample of a tile's luma rows can touch

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clamp_hi(int v, int hi) { return __vimin_s32_relu(v, hi); }
// av_clip_uintp2(x, 30) of a wrapped 32-bit sum
__device__ __forceinline__ int clip30(int v) { return __vimin_s32_relu(v, 0x3FFFFFFF); }
__device__ __forceinline__ int range15(int v, int coeff, int offset) { return min((v * coeff + offset) >> 14, 32767); }

__constant__ __align__(8) unsigned char c_dither[8][8] = {
    {36, 68, 60, 92, 34, 66, 58, 90},   {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74},
    {116, 20, 108, 12, 114, 18, 106, 10}, {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30},
    {48, 80, 40, 72, 54, 86, 46, 78},   {112, 16, 104, 8, 118, 22, 110, 14},
};

// ff_dither_8x8_128[row & 7][col & 7].  The row index is (nearly) warp-uniform, the column differs per thread: one 8-byte constant load of
// the row and a shift, instead of a per-thread index into constant memory (which is served address by address)
__device__ __forceinline__ int dither128(int row, int col) {
    const uint2 r = *reinterpret_cast<const uint2 *>(&c_dither[row & 7][0]);
    return (int)((((col & 4) ? r.y : r.x) >> (8 * (col & 3))) & 0xFFu);
}

// horizontal 4-tap chroma upsample of a 10-bit 4:2:2 / 4:2:0 chroma row at luma column x (oracle chroma_up_422)
__device__ __forceinline__ int up422(const uint16_t *C, int cw, int x) {
    const int i = x >> 1;
    int s;
    if (x == 0)
        s = 17729 * (int)__ldg(C) - 1345 * (int)__ldg(C + 1);
    else if (x == 2)
        s = 3835 * (int)__ldg(C) + 13894 * (int)__ldg(C + 1) - 1345 * (int)__ldg(C + 2);
    else if (!(x & 1))
        s = -461 * (int)__ldg(C + max(i - 2, 0)) + 3943 * (int)__ldg(C + max(i - 1, 0)) + 14284 * (int)__ldg(C + i) -
            1382 * (int)__ldg(C + min(i + 1, cw - 1));
    else
        s = -1382 * (int)__ldg(C + max(i - 1, 0)) + 14284 * (int)__ldg(C + i) + 3943 * (int)__ldg(C + min(i + 1, cw - 1)) -
            461 * (int)__ldg(C + min(i + 2, cw - 1));
    return min(s >> 9, 32767);
}

// The same filter for the two luma columns x (even) and x + 1 of one chroma sample pair: five loads instead of eight
__device__ __forceinline__ void up422_pair(const uint16_t *C, int cw, int x, int &even, int &odd) {
    const int i = x >> 1;
    const int a = (int)__ldg(C + max(i - 2, 0)), b = (int)__ldg(C + max(i - 1, 0)), c = (int)__ldg(C + i);
    const int d = (int)__ldg(C + min(i + 1, cw - 1)), e = (int)__ldg(C + min(i + 2, cw - 1));
    // columns 0 and 2 take swscale's edge-normalised taps (there a = b = C[0])
    const int t0 = x <= 2 ? 0 : -461, t1 = x == 0 ? 0 : (x == 2 ? 3835 : 3943), t2 = x == 0 ? 17729 : (x == 2 ? 13894 : 14284);
    const int t3 = x <= 2 ? -1345 : -1382;
    even = min((t0 * a + t1 * b + t2 * c + t3 * d) >> 9, 32767);
    odd = min((-1382 * b + 14284 * c + 3943 * d - 461 * e) >> 9, 32767);
}

// One sample of a full-range 10-bit source through swscale's one-pass conversion to the 8-bit intermediate the reference inserts
// (`scale=in_range=pc:out_range=R,format=yuv4xxp`, ffmpeg.py:211-233; oracle orc_prepass_10_to_8).  (py, px): position in the plane.
__device__ __forceinline__ int pre10_sample(int v, int py, int px, int mode, int plane, int coeff, int offset) {
    if (mode == 1)  // generic scaler: 15-bit sample, range conversion, ordered dither (V shifted by 3 columns)
        return clamp_hi((range15(v << 5, coeff, offset) + dither128(py, px + (plane == 2 ? 3 : 0))) >> 7, 255);
    const int d = (py & 1) ? ((px & 1) ? 0 : 3) : ((px & 1) ? 2 : 1);  // planar copy, 2x2 dither {{1,2},{3,0}}
    if (plane == 0) return min((v + d - (v >> 8)) >> 2, 255);
    const int t = (v + d) >> 2;
    return t - (t >> 8);
}

template <int INTERP, bool FMA, int FMT>
__global__ void __launch_bounds__(kThreads, 4) k_any(const StageK K) {
    __shared__ __align__(16) unsigned s_c[kCH][kCW];       // per luma pixel: (cu | cv << 16) chroma terms (c14, or c15 for 8-bit 4:4:4)
    __shared__ unsigned s_h[kCH][kTW / 2];   // 4:2:0 outputs: per chroma column and luma row, 15-bit (int16 | int16 << 16)
    __shared__ __align__(16) unsigned s_up[kUpRows][kCW];  // 10-bit 4:2:0 sources: horizontally upsampled chroma rows of the tile, (int16 | int16 << 16)

    const int w = K.w, h = K.h;
    // FMT >= 0: the layouts are compile-time constants (any_fmt_code below): the format switches of the loops fold away, measured
    // +23 % on the 4:2:2 10-bit cell.  FMT < 0: run-time switches, one instantiation for the whole format matrix.
    constexpr bool FX = FMT >= 0;
    const int pre10 = FX ? 0 : K.any.pre10;  // != 0: 10-bit full-range samples enter through the reference's 8-bit intermediate
    const int sd = FX ? ((FMT & 1) ? 10 : 8) : (pre10 ? 8 : K.any.in_depth);
    const int in_sub = FX ? ((FMT >> 1) & 3) : K.any.in_sub, out_sub = FX ? ((FMT >> 3) & 3) : K.any.out_sub;
    const int in_sx = in_sub & 1, in_sy = in_sub >> 1, o_sx = out_sub & 1, o_sy = out_sub >> 1;
    const int od = FX ? ((FMT & 32) ? 10 : 8) : K.any.out_depth;
    const bool out_full = FX ? false : K.any.out_full != 0;
    const bool bayer = sd > 8 && od == 8;
    const int hx = (sd > 8 && o_sx) ? 4 : 0;  // halo columns per side (the window stays pair-aligned)
    const int hy0 = o_sy ? 3 : 0, hy1 = o_sy ? 4 : 0;
    const int cwi = w >> in_sx, cho = h >> o_sy;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    const unsigned yA10 = (unsigned)(128 * f.yC), yK10 = (unsigned)(-f.yOff10 * f.yC) + (1u << 19);
    const int tiles_x = (w + kTW - 1) / kTW, tiles_y = (h + kTH - 1) / kTH;
    const long total = (long)tiles_x * tiles_y * K.nframes;
    const int omax = (1 << od) - 1;
    const LutC LC = make_lutc(K);

    for (long t = blockIdx.x; t < total; t += gridDim.x) {
        const int frame = (int)(t / ((long)tiles_x * tiles_y));
        const int tr = (int)(t - (long)frame * tiles_x * tiles_y), ty = tr / tiles_x, tx = tr - ty * tiles_x;
        const int X0 = tx * kTW, Y0 = ty * kTH;
        const int X1 = min(X0 + kTW, w), Y1 = min(Y0 + kTH, h);        // interior (outputs this tile owns)
        const int WX0 = X0 - hx, WX1 = min(X1 + hx, w);                 // window columns actually computed: [max(WX0,0), WX1)
        const int WY0 = max(Y0 - hy0, 0), WY1 = min(Y1 + hy1, h);       // window rows
        const int npairs = (WX1 - max(WX0, 0)) >> 1, nrows = WY1 - WY0;
        const int px0 = max(WX0, 0);
        const uint8_t *inY = K.p.in[0] + (long)frame * K.p.in_step, *inU = K.p.in[1] + (long)frame * K.p.in_step;
        const uint8_t *inV = K.p.in[2] + (long)frame * K.p.in_step;
        uint8_t *outY = K.p.out[0] + (long)frame * K.p.out_step, *outU = K.p.out[1] + (long)frame * K.p.out_step;
        uint8_t *outV = K.p.out[2] + (long)frame * K.p.out_step;

        __syncthreads();  // the previous tile's phases B / C have finished reading shared memory
        // ------------------------------------------------ phase 0 ---------------------------------------------------------
        // 10-bit 4:2:0 source: swscale upsamples chroma horizontally first (per chroma row, to 15 bits) and vertically afterwards, so
        // every chroma row the tile needs is upsampled ONCE here instead of four times per pixel in phase A
        const bool stage_up = sd > 8 && in_sx && in_sy;
        const int chi = h >> 1;
        const int uc0 = max((WY0 >> 1) - 2, 0), uc1 = min(((WY1 - 1) >> 1) + 2, chi - 1);
        if (stage_up) {
            const int ncols = WX1 - px0;
            for (int idx = threadIdx.x; idx < (uc1 - uc0 + 1) * ncols; idx += kThreads) {
                const int cr = idx / ncols, xx = px0 + (idx - cr * ncols);
                const int cu = up422(reinterpret_cast<const uint16_t *>(inU + (long)(uc0 + cr) * K.p.in_stride[1]), cwi, xx);
                const int cv = up422(reinterpret_cast<const uint16_t *>(inV + (long)(uc0 + cr) * K.p.in_stride[2]), cwi, xx);
                s_up[cr][xx - WX0] = ((unsigned)cu & 0xFFFFu) | ((unsigned)cv << 16);
            }
            __syncthreads();
        }
        // ------------------------------------------------ phase A ---------------------------------------------------------
        // idx / n for idx * n < 2^20 (a tile has at most 68 x 39 pairs): multiply by ceil(2^20 / n)
        const unsigned rcp_pairs = 0xFFFFFu / (unsigned)npairs + 1u;
        for (int idx = threadIdx.x; idx < npairs * nrows; idx += kThreads) {
            const int r = (int)(((unsigned)idx * rcp_pairs) >> 20), pc = idx - r * npairs;
            const int y = WY0 + r, x = px0 + 2 * pc;
            int R[2], G[2], B[2];
            if (sd == 8) {
                const uint8_t *yr = inY + (long)y * K.p.in_stride[0];
                const uint8_t *ur = inU + (long)(y >> in_sy) * K.p.in_stride[1], *vr = inV + (long)(y >> in_sy) * K.p.in_stride[2];
                int Yv[2], Uv[2], Vv[2];
                if (pre10) {  // 16-bit samples, re-quantised to the 8-bit intermediate at their own plane coordinates
                    const uint16_t *y16 = reinterpret_cast<const uint16_t *>(yr), *u16 = reinterpret_cast<const uint16_t *>(ur);
                    const uint16_t *v16 = reinterpret_cast<const uint16_t *>(vr);
                    const int cy = y >> in_sy;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int cx = (x + q) >> in_sx;
                        Yv[q] = pre10_sample((int)__ldg(y16 + x + q), y, x + q, pre10, 0, K.any.in_lc, K.any.in_lo);
                        Uv[q] = pre10_sample((int)__ldg(u16 + cx), cy, cx, pre10, 1, K.any.in_cc, K.any.in_co);
                        Vv[q] = pre10_sample((int)__ldg(v16 + cx), cy, cx, pre10, 2, K.any.in_cc, K.any.in_co);
                    }
                } else if (in_sx) {
                    Yv[0] = (int)__ldg(yr + x), Yv[1] = (int)__ldg(yr + x + 1);
                    Uv[0] = Uv[1] = (int)__ldg(ur + (x >> 1)), Vv[0] = Vv[1] = (int)__ldg(vr + (x >> 1));
                } else {
                    Yv[0] = (int)__ldg(yr + x), Yv[1] = (int)__ldg(yr + x + 1);
                    Uv[0] = (int)__ldg(ur + x), Uv[1] = (int)__ldg(ur + x + 1), Vv[0] = (int)__ldg(vr + x), Vv[1] = (int)__ldg(vr + x + 1);
                }
                if (K.range_pass) {  // swscale's 8-bit pc -> tv pass: 15-bit range conversion, then the 8-bit output stage
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        Yv[q] = clamp_hi((range15(Yv[q] << 7, K.any.in_lc, K.any.in_lo) + 64) >> 7, 255);
                        Uv[q] = clamp_hi((range15(Uv[q] << 7, K.any.in_cc, K.any.in_co) + 64) >> 7, 255);
                        Vv[q] = clamp_hi((range15(Vv[q] << 7, K.any.in_cc, K.any.in_co) + 64) >> 7, 255);
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (in_sx) {  // x86 SIMD path of yuv420p / yuv422p -> rgb24
                        const int yy = ((Yv[q] * 8 - f.yO8) * f.yC) >> 16, u = Uv[q] * 8 - 1024, v = Vv[q] * 8 - 1024;
                        R[q] = clamp_hi(yy + ((v * f.vr) >> 16), 255);
                        G[q] = clamp_hi(yy + ((u * f.ug) >> 16) + ((v * f.vg) >> 16), 255);
                        B[q] = clamp_hi(yy + ((u * f.ub) >> 16), 255);
                    } else {      // C path of yuv444p -> rgb24 (yuv2rgb_full: 30-bit clip, >> 22)
                        const unsigned yy = (unsigned)(Yv[q] * 512 - f.yOff10) * (unsigned)f.yC + (1u << 21);
                        const unsigned u = (unsigned)((Uv[q] - 128) * 512), v = (unsigned)((Vv[q] - 128) * 512);
                        R[q] = clip30((int)(yy + v * (unsigned)f.vr)) >> 22;
                        G[q] = clip30((int)(yy + v * (unsigned)f.vg + u * (unsigned)f.ug)) >> 22;
                        B[q] = clip30((int)(yy + u * (unsigned)f.ub)) >> 22;
                    }
                }
            } else {
                const uint16_t *yr = reinterpret_cast<const uint16_t *>(inY + (long)y * K.p.in_stride[0]);
                int u17[2], v17[2];
                if (!in_sx) {         // yuv444p10le
                    const uint16_t *ur = reinterpret_cast<const uint16_t *>(inU + (long)y * K.p.in_stride[1]);
                    const uint16_t *vr = reinterpret_cast<const uint16_t *>(inV + (long)y * K.p.in_stride[2]);
                    const unsigned up = __ldg(reinterpret_cast<const unsigned *>(ur + x)), vp = __ldg(reinterpret_cast<const unsigned *>(vr + x));
                    u17[0] = ((int)(up & 0xFFFFu) - 512) * 128, u17[1] = ((int)(up >> 16) - 512) * 128;
                    v17[0] = ((int)(vp & 0xFFFFu) - 512) * 128, v17[1] = ((int)(vp >> 16) - 512) * 128;
                } else if (!in_sy) {  // yuv422p10le: 4-tap horizontal upsample of the pair's two columns
                    const uint16_t *ur = reinterpret_cast<const uint16_t *>(inU + (long)y * K.p.in_stride[1]);
                    const uint16_t *vr = reinterpret_cast<const uint16_t *>(inV + (long)y * K.p.in_stride[2]);
                    int ue, uo, ve, vo;
                    up422_pair(ur, cwi, x, ue, uo), up422_pair(vr, cwi, x, ve, vo);
                    u17[0] = 4 * ue - 65536, u17[1] = 4 * uo - 65536, v17[0] = 4 * ve - 65536, v17[1] = 4 * vo - 65536;
                } else {              // yuv420p10le: the staged horizontally upsampled rows, then the 4-tap vertical filter
                    const int j = y >> 1;
                    int r0, t0, t1, t2, t3;
                    if (y & 1) {
                        r0 = j - 1, t0 = -346, t1 = 3572, t2 = 985, t3 = -115;
                    } else {
                        r0 = j - 2;
                        if (y == 0) t0 = 0, t1 = 0, t2 = 4432, t3 = -336;
                        else if (y == 2) t0 = 0, t1 = 959, t2 = 3473, t3 = -336;
                        else t0 = -115, t1 = 985, t2 = 3572, t3 = -346;
                    }
                    const int tp[4] = {t0, t1, t2, t3};
                    int su[2], sv[2];
                    su[0] = su[1] = sv[0] = sv[1] = (1 << 9) - (128 << 19);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (tp[k] == 0) continue;
                        // the pair's two columns are neighbours in the staged row: one 64-bit load (x - WX0 is even, rows are 8-byte aligned)
                        const uint2 pk = *reinterpret_cast<const uint2 *>(&s_up[clampi(r0 + k, 0, chi - 1) - uc0][x - WX0]);
                        su[0] += tp[k] * (int)(short)(pk.x & 0xFFFFu), sv[0] += tp[k] * ((int)pk.x >> 16);
                        su[1] += tp[k] * (int)(short)(pk.y & 0xFFFFu), sv[1] += tp[k] * ((int)pk.y >> 16);
                    }
                    u17[0] = su[0] >> 10, v17[0] = sv[0] >> 10, u17[1] = su[1] >> 10, v17[1] = sv[1] >> 10;
                }
                const unsigned ypair = __ldg(reinterpret_cast<const unsigned *>(yr + x));   // two 10-bit luma samples
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const unsigned yy = (q ? ypair >> 16 : ypair & 0xFFFFu) * yA10 + yK10;
                    R[q] = clip30((int)(yy + (unsigned)v17[q] * (unsigned)f.vr)) >> 20;
                    G[q] = clip30((int)(yy + (unsigned)v17[q] * (unsigned)f.vg + (unsigned)u17[q] * (unsigned)f.ug)) >> 20;
                    B[q] = clip30((int)(yy + (unsigned)u17[q] * (unsigned)f.ub)) >> 20;
                }
            }
            // lut3d on both pixels, then the luma output and the chroma terms
            int ro[2], go[2], bo[2], yq[2];
            unsigned cq[2];
            if constexpr (INTERP == LUTR_INTERP_TETRAHEDRAL || INTERP == LUTR_INTERP_TRILINEAR) {
                // the packed two-pixel evaluation of the warp-window kernels (general index, coordinate and output clips kept):
                // codes go in as the floats 2^23 + code and come out as 0x4B000000 | code
                const unsigned mr[2] = {kMagic | (unsigned)R[0], kMagic | (unsigned)R[1]}, mg[2] = {kMagic | (unsigned)G[0], kMagic | (unsigned)G[1]};
                const unsigned mb[2] = {kMagic | (unsigned)B[0], kMagic | (unsigned)B[1]};
                unsigned pr[2], pg[2], pb[2];
                lut3d_pair<INTERP, FMA, true, 2, false, 0, true>(LC, K.lut.tab, mr, mg, mb, pr, pg, pb);
#pragma unroll
                for (int q = 0; q < 2; ++q) ro[q] = (int)(pr[q] & 0xFFFFu), go[q] = (int)(pg[q] & 0xFFFFu), bo[q] = (int)(pb[q] & 0xFFFFu);
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q) lut3d_px<INTERP, FMA>(K.lut, R[q], G[q], B[q], (1 << sd) - 1, ro[q], go[q], bo[q]);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                int y15;
                if (sd == 8) y15 = 2 * ((b.ry * ro[q] + b.gy * go[q] + b.by * bo[q] + (32 << 14) + (1 << 8)) >> 9);
                else y15 = min(2 * ((b.ry * ro[q] + b.gy * go[q] + b.by * bo[q] + (16 << 17) + (1 << 10)) >> 11), 32767);
                if (out_full) y15 = range15(y15, K.any.out_lc, K.any.out_lo);
                const int d = bayer ? dither128(y, x + q) : 64;
                yq[q] = od == 8 ? clamp_hi((y15 + d) >> 7, 255) : clamp_hi((y15 + 16) >> 5, omax);
                unsigned cu, cv;
                if (sd == 8) {
                    if (o_sx) {
                        cu = cv = 0;  // pair sums below
                    } else {
                        cu = 2 * ((b.ru * ro[q] + b.gu * go[q] + b.bu * bo[q] + (256 << 14) + (1 << 8)) >> 9);
                        cv = 2 * ((b.rv * ro[q] + b.gv * go[q] + b.bv * bo[q] + (256 << 14) + (1 << 8)) >> 9);
                    }
                } else {
                    cu = (b.ru * ro[q] + b.gu * go[q] + b.bu * bo[q] + (128 << 17) + (1 << 10)) >> 11;
                    cv = (b.rv * ro[q] + b.gv * go[q] + b.bv * bo[q] + (128 << 17) + (1 << 10)) >> 11;
                }
                cq[q] = (cu & 0xFFFFu) | (cv << 16);
            }
            if (sd == 8 && o_sx) {  // rgb24ToUV_half: matrix on the pair sums
                const int rs = ro[0] + ro[1], gs = go[0] + go[1], bs = bo[0] + bo[1];
                const unsigned cu = 2 * ((b.ru * rs + b.gu * gs + b.bu * bs + (256 << 15) + (1 << 9)) >> 10);
                const unsigned cv = 2 * ((b.rv * rs + b.gv * gs + b.bv * bs + (256 << 15) + (1 << 9)) >> 10);
                cq[0] = (cu & 0xFFFFu) | (cv << 16);
            }
            *reinterpret_cast<uint2 *>(&s_c[r][x - WX0]) = make_uint2(cq[0], cq[1]);  // x - WX0 is even: one 8-byte store per pair
            if (x >= X0 && x < X1 && y >= Y0 && y < Y1) {
                if (od == 8) *reinterpret_cast<unsigned short *>(outY + (long)y * K.p.out_stride[0] + x) = (unsigned short)(yq[0] | yq[1] << 8);
                else *reinterpret_cast<unsigned *>(outY + (long)y * K.p.out_stride[0] + 2 * x) = (unsigned)(yq[0] | yq[1] << 16);
            }
        }
        __syncthreads();
        // ------------------------------------------------ phase B ---------------------------------------------------------
        // chroma columns of the tile: ci = (X0 >> o_sx) + i; rows: all window rows for 4:2:0 outputs, the tile's own rows otherwise
        const int ncol = (X1 - X0) >> o_sx;
        const unsigned rcp_col = 0xFFFFFu / (unsigned)ncol + 1u;   // idx / ncol as in phase A (at most 128 x 39 samples)
        // no tap of this tile's decimation leaves the row: columns X0 - 3 .. X1 + 2 exist (then WX0 = X0 - 4 and tap k of chroma
        // column i sits at word 2 i + 1 + k of the staged row)
        const bool inner = X0 >= 4 && X1 + 3 <= w;
        for (int idx = threadIdx.x; idx < ncol * nrows; idx += kThreads) {
            const int r = (int)(((unsigned)idx * rcp_col) >> 20), i = idx - r * ncol;
            const int y = WY0 + r;
            int cu, cv;
            if (sd > 8 && o_sx) {  // 8-tap horizontal decimation over luma columns 2ci-3 .. 2ci+4, clamped to the row
                // the taps are symmetric and the staged words hold two 14-bit fields: add the mirrored words first (no carry between
                // the fields), four multiplies per channel instead of eight
                unsigned s0, s1, s2, s3;
                if (inner) {
                    const unsigned *row = &s_c[r][2 * i];
                    const uint2 p23 = *reinterpret_cast<const uint2 *>(row + 2), p45 = *reinterpret_cast<const uint2 *>(row + 4);
                    const uint2 p67 = *reinterpret_cast<const uint2 *>(row + 6);
                    s0 = row[1] + row[8], s1 = p23.x + p67.y, s2 = p23.y + p67.x, s3 = p45.x + p45.y;
                } else {
                    unsigned pk[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) pk[k] = s_c[r][clampi(X0 + 2 * i - 3 + k, 0, w - 1) - WX0];
                    s0 = pk[0] + pk[7], s1 = pk[1] + pk[6], s2 = pk[2] + pk[5], s3 = pk[3] + pk[4];
                }
                const int su = -230 * (int)(s0 & 0xFFFFu) - 692 * (int)(s1 & 0xFFFFu) + 1972 * (int)(s2 & 0xFFFFu) + 7142 * (int)(s3 & 0xFFFFu);
                const int sv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16);
                cu = min(su >> 13, 32767), cv = min(sv >> 13, 32767);
            } else {
                const unsigned pk = s_c[r][X0 + (i << o_sx) - WX0];
                cu = (int)(pk & 0xFFFFu), cv = (int)(pk >> 16);
                if (sd > 8) cu = min(2 * cu, 32767), cv = min(2 * cv, 32767);  // 4:4:4 from gbrp10le
            }
            if (out_full) cu = range15(cu, K.any.out_cc, K.any.out_co), cv = range15(cv, K.any.out_cc, K.any.out_co);
            if (o_sy) {
                s_h[r][i] = ((unsigned)cu & 0xFFFFu) | ((unsigned)cv << 16);
            } else {
                const int ci = (X0 >> o_sx) + i;
                const int du = bayer ? dither128(y, ci) : 64, dv = bayer ? dither128(y, ci + 3) : 64;
                if (od == 8) {
                    outU[(long)y * K.p.out_stride[1] + ci] = (uint8_t)clamp_hi((cu + du) >> 7, 255);
                    outV[(long)y * K.p.out_stride[2] + ci] = (uint8_t)clamp_hi((cv + dv) >> 7, 255);
                } else {
                    *reinterpret_cast<uint16_t *>(outU + (long)y * K.p.out_stride[1] + 2 * ci) = (uint16_t)clamp_hi((cu + 16) >> 5, omax);
                    *reinterpret_cast<uint16_t *>(outV + (long)y * K.p.out_stride[2] + 2 * ci) = (uint16_t)clamp_hi((cv + 16) >> 5, omax);
                }
            }
        }
        if (!o_sy) continue;
        __syncthreads();
        // ------------------------------------------------ phase C ---------------------------------------------------------
        const int nj = (Y1 - Y0) >> 1;
        for (int idx = threadIdx.x; idx < ncol * nj; idx += kThreads) {
            const int jr = (int)(((unsigned)idx * rcp_col) >> 20), i = idx - jr * ncol;
            const int j = (Y0 >> 1) + jr, ci = (X0 >> 1) + i;
            const int taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
            const int du = bayer ? dither128(j, ci) : 64, dv = bayer ? dither128(j, ci + 3) : 64;
            const bool c_path = od == 8 && j == cho - 1;  // swscale's C vertical scaler takes the last 8-bit output row
            int su, sv;
            if (od == 10) su = sv = 1 << 16;
            else if (c_path) su = du << 12, sv = dv << 12;
            else su = (du + 56) >> 4, sv = (dv + 56) >> 4;
            int fold = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // taps that land on the same (edge) row are summed into one coefficient before the multiply
                const int row = clampi(2 * j - 3 + k, 0, h - 1), nrow = clampi(2 * j - 2 + k, 0, h - 1);
                fold += taps[k];
                if (k == 7 || nrow != row) {
                    const unsigned pk = s_h[row - WY0][i];
                    const int cu = (int)(short)(pk & 0xFFFFu), cv = (int)pk >> 16;
                    if (od == 10 || c_path) su += cu * fold, sv += cv * fold;
                    else su += (cu * fold) >> 16, sv += (cv * fold) >> 16;
                    fold = 0;
                }
            }
            if (od == 10) {
                *reinterpret_cast<uint16_t *>(outU + (long)j * K.p.out_stride[1] + 2 * ci) = (uint16_t)clamp_hi(su >> 17, omax);
                *reinterpret_cast<uint16_t *>(outV + (long)j * K.p.out_stride[2] + 2 * ci) = (uint16_t)clamp_hi(sv >> 17, omax);
            } else {
                outU[(long)j * K.p.out_stride[1] + ci] = (uint8_t)clamp_hi(su >> (c_path ? 19 : 3), 255);
                outV[(long)j * K.p.out_stride[2] + ci] = (uint8_t)clamp_hi(sv >> (c_path ? 19 : 3), 255);
            }
        }
    }
}

// compile-time format code of a fixed-format instantiation: source depth 10 | source subsampling (0 / 1 / 3) | output subsampling |
// output depth 10; limited-range output, no 10-bit pre-pass
constexpr int any_fmt_code(bool sd10, int in_sub, int out_sub, bool od10) { return (sd10 ? 1 : 0) | (in_sub << 1) | (out_sub << 3) | (od10 ? 32 : 0); }

using AnyKernelFn = void (*)(const StageK);

inline int launch_any_fn(AnyKernelFn fn, const StageK &k, void *stream) {
    const long tiles = (long)((k.w + kTW - 1) / kTW) * ((k.h + kTH - 1) / kTH) * k.nframes;
    long grid = k.grid_ctas > 0 ? k.grid_ctas : 148l * 4 * 4;  // 4 resident CTAs per SM, 4 waves of tiles
    if (grid > tiles) grid = tiles;
    fn<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace

// fixed-format instantiations (k_any_fixed10.cu / k_any_fixed8.cu): the kernel for a format code and tetrahedral / trilinear, or null
AnyKernelFn any_fixed10(int code, int interp);
AnyKernelFn any_fixed8(int code, int interp);

}  // namespace lutr
