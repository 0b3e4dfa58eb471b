// Warp-window fused kernels for the 10-bit chains whose rows must be walked IN ORDER by one warp:
//
//   IN420  : yuv420p10le source   (10-bit HEVC / camera 4:2:0)                 -> vertical 4-tap chroma upsampling needs a row ring
//   OUT420 : yuv420p10le output   (libx265 / hevc_videotoolbox "preserve",     -> vertical 8-tap chroma decimation needs a row ring
//            reference ffmpeg.py:297-298, 309-310)
//   OUT8   : yuv420p output       (10-bit source to a codec without 10-bit support, e.g. the default libx264: ffmpeg.py:299-302)
//            = OUT420 with swscale's 8-bit output stage: ordered dither ff_dither_8x8_128, per-tap truncating vertical scaler
//
// Instantiated: IN420 -> 4:2:2 out (pro-mode master, ffmpeg.py:417-433), IN420 -> OUT420, 4:2:2 in -> OUT420.  (4:2:2 -> 4:2:2 is
// row-local and lives in k422_warp.cu.)  Arithmetic: oracle/lutr_oracle.c orc_yuv420p10_to_gbrp10 / orc_yuv422p10_to_gbrp10 +
// lut3d + orc_gbrp10_to_yuv422p10 / orc_gbrp10_to_yuv420p10, every leg pinned to libswscale 9.1.100.
//
// Same three passes per luma row as k422_warp.cu (one pixel column per lane in the LUT pass, 256-column windows with 8 halo
// columns per side).  Differences:
//   * IN420: swscale upsamples chroma horizontally first (to 15 bits, with a clip) and vertically afterwards, so the horizontally
//     upsampled rows are kept in a per-warp ring of 4 chroma rows of packed (cu | cv << 16) words.  Every chroma row is upsampled
//     once; pass B adds the 4-tap vertical filter per column.
//   * OUT420: pass C leaves the horizontally decimated 15-bit row (swscale's intermediate, 32767 clip included) in a per-warp ring
//     of 8 luma rows; chroma row j is emitted after luma row 2j+4 by the exact 8-tap vertical accumulation of yuv2planeX_10.
//     A share recomputes 3 luma rows above and below its own rows for the filter support (no luma output for those).
//   * the persistent grid is split into equal, contiguous shares of the chroma rows of all (frame, window) columns (like
//     k420_warp.cuh), each share re-warming its rings.
#pragma once
#include <cuda_runtime.h>

#include "lutr_internal.h"
#include "warp_common.cuh"

namespace lutr {

namespace {

constexpr int kORow = kWin / 2;  // words per row of the output ring (120 used)
#ifndef LUTR_K10_WARPS
#define LUTR_K10_WARPS 8
#endif
constexpr int kW10 = LUTR_K10_WARPS;  // warps per CTA (a CTA is only a container)

// ff_dither_8x8_128: swscale's ordered dither of the 8-bit output stage for sources deeper than 8 bit (OUT8)
__constant__ unsigned char c_dither10[8][8] = {
    {36, 68, 60, 92, 34, 66, 58, 90},   {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74},
    {116, 20, 108, 12, 114, 18, 106, 10}, {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30},
    {48, 80, 40, 72, 54, 86, 46, 78},   {112, 16, 104, 8, 118, 22, 110, 14},
};

template <bool IN420, bool OUT420>
struct K10Smem {
    static constexpr int words = (IN420 ? 4 : 1) * kWin + kWin + kWin / 2 + (OUT420 ? 8 * kORow : 0);  // per warp
};

template <bool IN420, bool OUT420, bool OUT8, int INTERP, bool FMA, bool CLAMP, int IDX, bool NOMIN, int NSZ, int MINB>
__global__ void __launch_bounds__(kW10 * 32, MINB) k10_warp(const StageK K) {
    extern __shared__ __align__(16) unsigned s_dyn[];
    constexpr int NR = IN420 ? 4 : 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lane_odd = lane ^ 4;
    unsigned *ring = s_dyn + wib * K10Smem<IN420, OUT420>::words;  // [NR][256] horizontally upsampled chroma rows, (cu | cv << 16)
    unsigned *sC = ring + NR * kWin;                                  // [256] (u14 | v14 << 16) of the current luma row
    unsigned *sYw = sC + kWin;                                        // [128] luma samples, two per word (in: source, out: result)
    unsigned short *sY = reinterpret_cast<unsigned short *>(sYw);    // [256]
    unsigned *oring = sYw + kWin / 2;                                 // OUT420: [8][128] decimated rows, (hu | hv << 16), biased

    const int w = K.w, h = K.h, ch = h >> 1, tpr = K.tiles_per_row;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    const FastK fk = K.fk;
    const LutC LC = make_lutc(K);
    const float4 *tab = K.lut.tab;
    // yy = (Y*128 - yOff10)*yC + (1 << 19) in wrapping 32-bit arithmetic
    const unsigned yA = fk.yA, yK = (unsigned)(-f.yOff10 * f.yC) + (1u << 19);
    // IN420: u17 = ((1<<9) - (128<<19) + sum) >> 10 with sum over BIASED c15 values: the taps always add up to 4096, so the bias
    // contributes kCBias * 4096; both constants ride on the multiply-add chains of the vertical filter
    constexpr int kVK = (1 << 9) - (128 << 19) - kCBias * 4096;
    // 4:2:2 in: the folded constants of k422_warp.cu (u17 = 4 * c15 - 65536 on the biased packed word)
    const unsigned c_vgx = fk.vg4 - 65536u * fk.ug4, c_ubx = 0u - 65536u * fk.ub4;
    const unsigned c_kR = fk.kR - (unsigned)kCBias * fk.vr4, c_kG = fk.kG - (unsigned)kCBias * fk.vg4 - (unsigned)kCBias * fk.ug4;
    const unsigned c_kB = fk.kB - (unsigned)kCBias * fk.ub4;

    const unsigned total_c = (unsigned)K.nframes * (unsigned)tpr * (unsigned)ch;
    const unsigned nwarps = gridDim.x * kW10, gw = blockIdx.x * kW10 + wib;
    const unsigned per = total_c / nwarps, rem = total_c % nwarps;
    unsigned c_pos = gw * per + min(gw, rem);
    const unsigned c_end = c_pos + per + (gw < rem ? 1u : 0u);

    while (c_pos < c_end) {
        const unsigned col = c_pos / (unsigned)ch, j0 = c_pos - col * (unsigned)ch;
        const unsigned j1 = min((unsigned)ch, j0 + (c_end - c_pos));
        const unsigned frame = col / (unsigned)tpr, tx = col - frame * (unsigned)tpr;
        c_pos += j1 - j0;
        const int a = (int)tx * kOut - 8;  // first column of the window
        const bool edge_tile = tx == 0 || tx == (unsigned)tpr - 1;
        const int yo0 = 2 * (int)j0, yo1 = 2 * (int)j1;  // luma rows this share owns
        // rows it computes: OUT420 needs the decimated rows 2j-3 .. 2j+4 of every chroma row j it emits
        const int y0 = OUT420 ? max(yo0 - 3, 0) : yo0, y1 = OUT420 ? min(yo1 + 3, h) : yo1;
        const int x0 = a + 8 * lane, xc = min(max(x0, 0), w - 8);  // this lane's 8 columns in pass A (clamped into the row)
        const uint8_t *pY = K.p.in[0] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[0] + 2 * xc;
        const uint8_t *bU = K.p.in[1] + (long)frame * K.p.in_step + xc;
        const uint8_t *bV = K.p.in[2] + (long)frame * K.p.in_step + xc;
        uint8_t *oY = K.p.out[0] + (long)frame * K.p.out_step + (long)yo0 * K.p.out_stride[0] + (OUT8 ? 1 : 2) * (long)(a + 8 + 8 * lane);
        // chroma output rows: 4:2:2 = luma rows from yo0, 4:2:0 = chroma rows (addressed per emitted row)
        const long c_off = OUT8 ? (long)((a + 8 + 8 * lane) >> 1) : (long)(a + 8 + 8 * lane);  // chroma sample (a+8)/2 + 4*lane, in bytes
        uint8_t *oU = K.p.out[1] + (long)frame * K.p.out_step + (OUT420 ? 0l : (long)yo0 * K.p.out_stride[1]) + c_off;
        uint8_t *oV = K.p.out[2] + (long)frame * K.p.out_step + (OUT420 ? 0l : (long)yo0 * K.p.out_stride[2]) + c_off;
        const int nout = min(kOut, w - (a + 8));  // columns a+8 .. a+8+nout-1 are this window's outputs
        const bool do_c = 8 * lane < nout;

        // horizontal 4-tap chroma upsample of this lane's 8 columns of one chroma row -> ring slot (row & (NR-1)); identical to
        // pass A of k422_warp.cu (neighbour samples by shuffle, edge-normalised taps on the two left-most even columns of a row)
        auto upsample_row = [&](uint2 cu2, uint2 cv2, int crow) {
            unsigned ul = __shfl_up_sync(0xFFFFFFFFu, cu2.y, 1), vl = __shfl_up_sync(0xFFFFFFFFu, cv2.y, 1);
            unsigned ur = __shfl_down_sync(0xFFFFFFFFu, cu2.x, 1), vr = __shfl_down_sync(0xFFFFFFFFu, cv2.x, 1);
            if (edge_tile) {
                if (xc == 0) ul = (cu2.x & 0xFFFFu) * 0x10001u, vl = (cv2.x & 0xFFFFu) * 0x10001u;
                if (xc == w - 8) ur = (cu2.y >> 16) * 0x10001u, vr = (cv2.y >> 16) * 0x10001u;
            }
            const unsigned wu[4] = {ul, cu2.x, cu2.y, ur}, wvv[4] = {vl, cv2.x, cv2.y, vr};
            unsigned wv[8];
            chroma_up8<0>(wu, wvv, edge_tile && xc == 0, wv);  // exact float filter (warp_common.cuh)
            unsigned *dst = ring + (crow & (NR - 1)) * kWin;
            *reinterpret_cast<uint4 *>(dst + 4 * swz_chunk(2 * lane)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            *reinterpret_cast<uint4 *>(dst + 4 * swz_chunk(2 * lane + 1)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        };
        // chroma row `crow` of the source planes (IN420: chroma row index, 4:2:2: luma row index)
        auto load_c = [&](const uint8_t *base, int crow, int cstride) { return ldg_s8(base + (long)crow * cstride); };

        int next_c = 0;  // IN420: next chroma row to upsample; 4:2:2: unused
        uint2 nu = make_uint2(0u, 0u), nv = make_uint2(0u, 0u);
        __syncwarp();  // the previous share's last passes have finished reading the rings
        if (IN420) {
            // warm the ring with the chroma rows the first luma row needs: even y = 2j: j-2 .. j+1, odd: j-1 .. j+2 (clamped)
            const int jf = y0 >> 1, odd = y0 & 1;
            const int c_lo = max(jf - 2 + odd, 0), c_hi = min(jf + 1 + odd, ch - 1);
            for (int c = c_lo; c <= c_hi; ++c) upsample_row(load_c(bU, c, K.p.in_stride[1]), load_c(bV, c, K.p.in_stride[2]), c);
            next_c = c_hi + 1;  // needed first by the odd luma row 2*(next_c - 2) + 1
            if (next_c < ch) nu = load_c(bU, next_c, K.p.in_stride[1]), nv = load_c(bV, next_c, K.p.in_stride[2]);
        } else {
            nu = load_c(bU, y0, K.p.in_stride[1]), nv = load_c(bV, y0, K.p.in_stride[2]);
        }
        uint4 ny = ldg_s16(pY);

        for (int y = y0; y < y1; ++y) {
            const uint4 cy = ny;
            const uint2 cu2 = nu, cv2 = nv;
            if (y + 1 < y1) {
                pY += K.p.in_stride[0];
                ny = ldg_s16(pY);
                if (!IN420) nu = load_c(bU, y + 1, K.p.in_stride[1]), nv = load_c(bV, y + 1, K.p.in_stride[2]);
            }
            const int j = y >> 1;
            __syncwarp();  // pass C of the previous row has finished reading sC / sY; its pass B the ring
            if (IN420) {
                if ((y & 1) && j + 2 == next_c && next_c < ch) {  // the odd row 2j+1 is the first to need chroma row j+2
                    upsample_row(cu2, cv2, next_c);
                    ++next_c;
                    if (next_c < ch) nu = load_c(bU, next_c, K.p.in_stride[1]), nv = load_c(bV, next_c, K.p.in_stride[2]);
                }
            } else {
                upsample_row(cu2, cv2, 0);
            }
            *reinterpret_cast<uint4 *>(sYw + 4 * lane) = cy;
            // IN420: vertical taps of this luma row (warp-uniform) and the ring rows they apply to
            int t0 = 0, t1 = 0, t2 = 0, t3 = 0, r0 = 0;
            if (IN420) {
                if (y & 1) {
                    r0 = j - 1, t0 = -346, t1 = 3572, t2 = 985, t3 = -115;
                } else {
                    r0 = j - 2;
                    if (y == 0) t0 = 0, t1 = 0, t2 = 4432, t3 = -336;
                    else if (y == 2) t0 = 0, t1 = 959, t2 = 3473, t3 = -336;
                    else t0 = -115, t1 = 985, t2 = 3572, t3 = -346;
                }
            }
            const unsigned *q0 = ring + (min(max(r0, 0), ch - 1) & (NR - 1)) * kWin, *q1 = ring + (min(max(r0 + 1, 0), ch - 1) & (NR - 1)) * kWin;
            const unsigned *q2 = ring + (min(max(r0 + 2, 0), ch - 1) & (NR - 1)) * kWin, *q3 = ring + (min(max(r0 + 3, 0), ch - 1) & (NR - 1)) * kWin;
            __syncwarp();
            // OUT8: this lane's luma dither of the row: column a + 32*k + lane has (a + 32*k) % 8 == 0
            const int dY = OUT8 ? (int)c_dither10[y & 7][lane & 7] : 0;

            // ---------------- pass B: one column per lane, columns a + 32*k + lane ---------------------------------------
#pragma unroll 4
            for (int kk = 0; kk < 8; kk += 2) {
                unsigned mr[2], mg[2], mb[2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int jc = 32 * (kk + p) + lane, slot = 32 * (kk + p) + (p ? lane_odd : lane);  // slot == swz_slot(jc)
                    const unsigned Y = sY[jc];
                    int Rv, Gv, Bv;
                    if (IN420) {
                        const unsigned w0 = q0[slot], w1 = q1[slot], w2 = q2[slot], w3 = q3[slot];
                        // vertical 4-tap filter on the packed words: V from the high halves, U = (filter over the words) - 65536 * V
                        const int sv = t0 * (int)(w0 >> 16) + t1 * (int)(w1 >> 16) + t2 * (int)(w2 >> 16) + t3 * (int)(w3 >> 16) + kVK;
                        const unsigned e = (unsigned)t0 * w0 + (unsigned)t1 * w1 + (unsigned)t2 * w2 + (unsigned)t3 * w3 + (unsigned)kVK +
                                           65536u * (unsigned)kVK;
                        const int su = (int)(e - 65536u * (unsigned)sv);
                        const unsigned u17 = (unsigned)(su >> 10), v17 = (unsigned)(sv >> 10);
                        const unsigned yy = Y * yA + yK;
                        Rv = (int)(yy + v17 * (unsigned)f.vr);
                        Gv = (int)(yy + v17 * (unsigned)f.vg + u17 * (unsigned)f.ug);
                        Bv = (int)(yy + u17 * (unsigned)f.ub);
                    } else {
                        // as pass B of k422_warp.cu: the packed word stands in for cu, the surplus is folded into the cv gain
                        const unsigned wv = q0[slot], cvb = wv >> 16;
                        Rv = (int)(cvb * fk.vr4 + (Y * yA + c_kR));
                        Gv = (int)(wv * fk.ug4 + (cvb * c_vgx + (Y * yA + c_kG)));
                        Bv = (int)(cvb * c_ubx + (wv * fk.ub4 + (Y * yA + c_kB)));
                    }
                    mr[p] = clip30_to_magic(Rv);
                    mg[p] = clip30_to_magic(Gv);
                    mb[p] = clip30_to_magic(Bv);
                }
                unsigned ro[2], go[2], bo[2];
                lut3d_pair<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, true>(LC, tab, mr, mg, mb, ro, go, bo);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int jc = 32 * (kk + p) + lane;
                    const int sy = (int)((unsigned)b.ry * ro[p] + (unsigned)b.gy * go[p] + (unsigned)b.by * bo[p] + fk.kY);
                    // 10 bit: clip10((min(2*y14, 32767) + 16) >> 5) with y14 = (S - (8 << 11)) >> 11  ==  S >> 15 (no clip binds, host-checked)
                    // 8 bit:  clip8((2*y14 + d) >> 7), d = ordered dither
                    const unsigned yq = OUT8 ? (unsigned)__vimin_s32_relu((2 * ((sy >> 11) - 8) + dY) >> 7, 255) : (unsigned)(sy >> 15);
                    // u14 = S >> 11 == (32 * S) >> 16: matrix rows pre-scaled by 32 on the host, one PRMT packs (u14 | v14 << 16)
                    const unsigned su5 = fk.cu5[0] * ro[p] + fk.cu5[1] * go[p] + fk.cu5[2] * bo[p] + fk.kU5;
                    const unsigned sv5 = fk.cv5[0] * ro[p] + fk.cv5[1] * go[p] + fk.cv5[2] * bo[p] + fk.kV5;
                    sC[32 * (kk + p) + (p ? lane_odd : lane)] = __byte_perm(su5, sv5, 0x7632u);
                    sY[jc] = (unsigned short)yq;
                }
            }
            __syncwarp();
            if (edge_tile) {  // warp-uniform: columns outside the row take the first / last column's chroma terms
                const int last = w - 1 - a;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int jc = 32 * q + lane;
                    if (a + jc < 0) sC[swz_slot(jc)] = sC[swz_slot(-a)];
                    if (jc > last) sC[swz_slot(jc)] = sC[swz_slot(last)];
                }
                __syncwarp();
            }

            // ---------------- pass C: 8-tap decimation, lane = chroma samples (a+8)/2 + 4*lane .. + 3 (as k422_warp.cu) ------
            const bool own_row = !OUT420 || (y >= yo0 && y < yo1);
            if (do_c) {
                unsigned tt[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 v4 = *reinterpret_cast<const uint4 *>(sC + 4 * swz_chunk(2 * lane + 1 + q));
                    tt[4 * q] = v4.x, tt[4 * q + 1] = v4.y, tt[4 * q + 2] = v4.z, tt[4 * q + 3] = v4.w;
                }
                // 4:2:2 out: clip10((min(S >> 13, 32767) + 16) >> 5) == clamp((S + (16 << 13)) >> 18, 0, 1023)
                // 4:2:0 out: keep swscale's 15-bit row min(S >> 13, 32767), biased by kCBias (never negative: S >> 13 >= -3688)
                // 8-bit out: the same 15-bit row, un-biased, as signed 16-bit halves (the per-tap truncating vertical scaler needs the sign)
                constexpr unsigned kRnd = OUT8 ? 0u : (OUT420 ? ((unsigned)kCBias << 13) : (16u << 13));
                unsigned uo[4], vo[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const unsigned s0 = tt[2 * jj + 1] + tt[2 * jj + 8], s1 = tt[2 * jj + 2] + tt[2 * jj + 7];
                    const unsigned s2 = tt[2 * jj + 3] + tt[2 * jj + 6], s3 = tt[2 * jj + 4] + tt[2 * jj + 5];
                    const int dv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16);
                    const unsigned e = 0u - 230u * s0 - 692u * s1 + 1972u * s2 + 7142u * s3 + kRnd;
                    const int du = (int)(e - 65536u * (unsigned)dv);
                    if (OUT8) {
                        uo[jj] = (unsigned)min(du >> 13, 32767) & 0xFFFFu;
                        vo[jj] = (unsigned)min(dv >> 13, 32767);
                    } else if (OUT420) {
                        uo[jj] = (unsigned)min(du >> 13, 32767 + kCBias);
                        vo[jj] = (unsigned)min((dv + (int)kRnd) >> 13, 32767 + kCBias);
                    } else {
                        uo[jj] = (unsigned)__vimin_s32_relu(du >> 18, 1023);
                        vo[jj] = (unsigned)__vimin_s32_relu((dv + (int)kRnd) >> 18, 1023);
                    }
                }
                if (own_row) {
                    const uint4 y8 = *reinterpret_cast<const uint4 *>(sYw + 4 * lane + 4);  // luma of columns a+8+8*lane .. +7
                    if (OUT8) stg_s8(oY, make_uint2(__byte_perm(y8.x, y8.y, 0x6420), __byte_perm(y8.z, y8.w, 0x6420)));
                    else stg_s16(oY, y8);
                }
                if (OUT420) {
                    *reinterpret_cast<uint4 *>(oring + (y & 7) * kORow + 4 * lane) =
                        make_uint4(uo[0] | vo[0] << 16, uo[1] | vo[1] << 16, uo[2] | vo[2] << 16, uo[3] | vo[3] << 16);
                } else {
                    stg_s8(oU, make_uint2(uo[0] | uo[1] << 16, uo[2] | uo[3] << 16));
                    stg_s8(oV, make_uint2(vo[0] | vo[1] << 16, vo[2] | vo[3] << 16));
                }
            }
            if (own_row) oY += K.p.out_stride[0];
            if (!OUT420) oU += K.p.out_stride[1], oV += K.p.out_stride[2];

            if (OUT420) {
                // chroma row jo is complete once luma row min(2*jo + 4, h - 1) has been decimated: after an even row y >= 4 that is
                // jo = (y - 4) / 2, after the last row of the frame the two bottom chroma rows (their taps clamp to row h - 1)
                int e0 = 0, e1 = 0;  // emit chroma rows [e0, e1)
                if (y == h - 1) e0 = ch - 2, e1 = ch;
                else if (!(y & 1) && y >= 4) e0 = (y - 4) >> 1, e1 = e0 + 1;
                e0 = max(e0, (int)j0), e1 = min(e1, (int)j1);
                for (int jo = e0; jo < e1; ++jo) {
                    if (OUT8) {
                        if (do_c) {
                            // x86 yuv2yuvX: ((d + 56) >> 4) + sum((c15 * tap) >> 16)) >> 3; the last chroma row of the frame goes through
                            // swscale's C scaler: ((d << 12) + sum(c15 * tap)) >> 19.  Taps that land on the same (edge) row are summed
                            // into one coefficient first.  Chroma column (a+8)/2 + 4*lane + q: (a+8)/2 is a multiple of 8.
                            constexpr int tv[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
                            const bool c_path = jo == ch - 1;
                            int au[4], av[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int du = c_dither10[jo & 7][(4 * lane + q) & 7], dv = c_dither10[jo & 7][(4 * lane + q + 3) & 7];
                                au[q] = c_path ? du << 12 : (du + 56) >> 4, av[q] = c_path ? dv << 12 : (dv + 56) >> 4;
                            }
                            if (2 * jo - 3 >= 0 && 2 * jo + 4 <= h - 1) {  // interior (never the last row): one multiply-high per tap
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const uint4 v4 = *reinterpret_cast<const uint4 *>(oring + ((2 * jo - 3 + k) & 7) * kORow + 4 * lane);
                                    const unsigned wv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        au[q] += __mulhi((int)(wv[q] << 16), tv[k]);          // (hu * tap) >> 16
                                        av[q] += __mulhi((int)(wv[q] & 0xFFFF0000u), tv[k]);  // (hv * tap) >> 16
                                    }
                                }
                            } else {
                                int fold = 0;
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const int row = min(max(2 * jo - 3 + k, 0), h - 1), nrow = min(max(2 * jo - 2 + k, 0), h - 1);
                                    fold += tv[k];
                                    if (k == 7 || nrow != row) {
                                        const uint4 v4 = *reinterpret_cast<const uint4 *>(oring + (row & 7) * kORow + 4 * lane);
                                        const unsigned wv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            const int hu = (int)(short)(wv[q] & 0xFFFFu), hv = (int)wv[q] >> 16;
                                            if (c_path) au[q] += hu * fold, av[q] += hv * fold;
                                            else au[q] += (hu * fold) >> 16, av[q] += (hv * fold) >> 16;
                                        }
                                        fold = 0;
                                    }
                                }
                            }
                            unsigned uq = 0, vq = 0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uq |= (unsigned)__vimin_s32_relu(au[q] >> (c_path ? 19 : 3), 255) << (8 * q);
                                vq |= (unsigned)__vimin_s32_relu(av[q] >> (c_path ? 19 : 3), 255) << (8 * q);
                            }
                            stg_s4(oU + (long)jo * K.p.out_stride[1], uq);
                            stg_s4(oV + (long)jo * K.p.out_stride[2], vq);
                        }
                    } else if (do_c) {  // (each lane reads back only what it wrote itself: no warp barrier needed)
                        unsigned su[4] = {0u, 0u, 0u, 0u};
                        int sv[4] = {0, 0, 0, 0};
                        constexpr int tv[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int r = min(max(2 * jo - 3 + k, 0), h - 1);
                            const uint4 v4 = *reinterpret_cast<const uint4 *>(oring + (r & 7) * kORow + 4 * lane);
                            const unsigned wv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                sv[q] += tv[k] * (int)(wv[q] >> 16);
                                su[q] += (unsigned)tv[k] * wv[q];  // mod 2^32: = sum(t * hu) + 65536 * sum(t * hv)
                            }
                        }
                        // clip10((sum + (1 << 16)) >> 17) on the un-biased sum (the taps add up to 4096)
                        constexpr int kVO = (1 << 16) - 4096 * kCBias;
                        unsigned uq[4], vq[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int tu = (int)(su[q] - 65536u * (unsigned)sv[q]);
                            uq[q] = (unsigned)__vimin_s32_relu((tu + kVO) >> 17, 1023);
                            vq[q] = (unsigned)__vimin_s32_relu((sv[q] + kVO) >> 17, 1023);
                        }
                        stg_s8(oU + (long)jo * K.p.out_stride[1], make_uint2(uq[0] | uq[1] << 16, uq[2] | uq[3] << 16));
                        stg_s8(oV + (long)jo * K.p.out_stride[2], make_uint2(vq[0] | vq[1] << 16, vq[2] | vq[3] << 16));
                    }
                }
            }
        }
    }
}

using KernelFn = void (*)(const StageK);

template <bool IN420, bool OUT420, bool OUT8, int INTERP, bool FMA, bool CLAMP, int NSZ, int MINB>
KernelFn pick_idx(bool fold_nomin) {
    if (fold_nomin) return k10_warp<IN420, OUT420, OUT8, INTERP, FMA, CLAMP, 1, true, NSZ, MINB>;
    return k10_warp<IN420, OUT420, OUT8, INTERP, FMA, CLAMP, 2, false, NSZ, MINB>;
}

template <bool IN420, bool OUT420, bool OUT8, int INTERP, int MINB>
KernelFn pick_sem(bool fma, bool clamp, bool fold_nomin, int n) {
    if (!fma && !clamp) {
        if (n == 65) return pick_idx<IN420, OUT420, OUT8, INTERP, false, false, 65, MINB>(fold_nomin);
        return pick_idx<IN420, OUT420, OUT8, INTERP, false, false, 0, MINB>(fold_nomin);
    }
    if (!fma) return pick_idx<IN420, OUT420, OUT8, INTERP, false, true, 0, MINB>(fold_nomin);
    return clamp ? pick_idx<IN420, OUT420, OUT8, INTERP, true, true, 0, MINB>(fold_nomin)
                 : pick_idx<IN420, OUT420, OUT8, INTERP, true, false, 0, MINB>(fold_nomin);
}

template <bool IN420, bool OUT420, bool OUT8>
int launch_k10(const StageK &k, int interp, int float_mode, bool grid_forced, void *stream) {
    const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
    const bool fold_nomin = k.fk.idx_fold && k.wk_nomin;
    // 2 CTAs per SM with 128 registers by default (measured 5-10 % faster than 3 CTAs x 80 registers, which spill); min_blocks == 3
    // selects the 80-register build of the tetrahedral kernel.  Both rings together (IN420 && OUT420: 76 KB of shared memory per
    // CTA) leave room for two CTAs per SM only; trilinear keeps 16 gathers in flight per step and always gets 128 registers.
    constexpr size_t smem = (size_t)kW10 * K10Smem<IN420, OUT420>::words * sizeof(unsigned);
    constexpr bool two_only = 3 * smem > 200 * 1024;
    const int minb = interp == LUTR_INTERP_TRILINEAR || k.min_blocks != 3 || two_only ? 2 : 3;
    KernelFn fn = interp == LUTR_INTERP_TRILINEAR ? pick_sem<IN420, OUT420, OUT8, LUTR_INTERP_TRILINEAR, 2>(fma, clamp, fold_nomin, k.lut.n)
                  : minb == 2                     ? pick_sem<IN420, OUT420, OUT8, LUTR_INTERP_TETRAHEDRAL, 2>(fma, clamp, fold_nomin, k.lut.n)
                                                  : pick_sem<IN420, OUT420, OUT8, LUTR_INTERP_TETRAHEDRAL, 3>(fma, clamp, fold_nomin, k.lut.n);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return LUTR_ERR_CUDA;
    const long total_c = (long)k.nframes * k.tiles_per_row * (k.h / 2);
    long grid = k.grid_ctas > 0 ? k.grid_ctas : 148 * minb;
    // small jobs: a smaller grid so that a share is at least 16 chroma rows and the ring warm-up / halo rows stay cheap
    if (!grid_forced && total_c / (kW10 * grid) < 16) grid = total_c / (kW10 * 16) > 0 ? total_c / (kW10 * 16) : 1;
    fn<<<(unsigned)grid, kW10 * 32, smem, (cudaStream_t)stream>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace

}  // namespace lutr
