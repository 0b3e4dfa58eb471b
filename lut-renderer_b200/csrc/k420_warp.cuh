// Warp-window fused kernel for yuv420p (8 bit) -> [pc->tv pass] -> lut3d -> yuv420p  (BASELINE cfg 1 / 2).
//
// Same arithmetic as k_yuv420p8_fast in stage_kernels.cu (reference chain: src/lut_renderer/ffmpeg.py:211-247, 304-310; integer
// legs as pinned against libswscale 9.1.100, float32 lut3d), decomposed like k422_warp.cu:
//
//  * work unit = one warp x (one window of 64*NK luma columns) x (a contiguous range of rows).  Nothing crosses warps.  4:2:0 has no
//    horizontal filter on either side (replicated chroma in, 2-pixel sums out), so windows need no halo columns; the 8-tap
//    VERTICAL chroma decimation needs 3 rows above and 4 below a band, which the unit recomputes (7 rows per band).
//  * per luma row, per step: lane = one chroma column = the pixel pair (2l, 2l+1).  A warp instruction therefore covers 64
//    adjacent pixels (LUT gathers of neighbouring pixels share cache lines) and the pair sum of rgb24ToUV_half stays in-lane.
//  * the per-row chroma terms go to a per-warp ring of 8 rows in shared memory; after every second row the lanes switch to
//    4 chroma columns each and run the vertical filter for one chroma row.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "lutr_internal.h"
#include "warp_common.cuh"

namespace lutr {

namespace {

constexpr int kWarps420 = 8;

__device__ __forceinline__ void stg_u16(void *p, unsigned v) {
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void stg_u8(void *p, unsigned v) {
    asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void stg_u32(void *p, unsigned v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// swscale's 8-bit full -> limited range pass on one sample (lumRangeFromJpeg / chrRangeFromJpeg on 15-bit intermediates, then
// the 8-bit output stage), as in stage_kernels.cu
__device__ __forceinline__ int pc2tv_y(int v) { return __vimin_s32_relu(((((v << 7) * 14071 + 33561947) >> 14) + 64) >> 7, 255); }
__device__ __forceinline__ int pc2tv_c(int v) { return __vimin_s32_relu(((((v << 7) * 7197 + 16762880) >> 13) + 64) >> 7, 255); }

// OUT: 0 = yuv420p, 1 = yuv422p10le, 2 = yuv422p, 3 = yuv444p (the last three are row-local: no vertical filter, no halo rows)
// IN422: the source is yuv422p (one chroma row per luma row) instead of yuv420p; same SIMD-path arithmetic (chroma replicated)
// RMODE: 0 = limited range throughout, 1 = 8-bit pc -> tv pass in front (BASELINE cfg 2), 2 = full-range source kept full range and
// FULL-RANGE OUTPUT (tags = inherit / none on a pc source, reference ffmpeg.py:216-223; 8-bit outputs only)
template <int INTERP, bool FMA, bool CLAMP, int IDX, bool NOMIN, int NSZ, int NK, int RMODE, int OUT, bool IN422, int MINB>
__global__ void __launch_bounds__(kWarps420 * 32, MINB) k420_warp(const StageK K) {
    constexpr int W = 64 * NK, CW = W / 2;  // luma / chroma columns per window
    constexpr bool ROWLOCAL = OUT != 0, OUT10 = OUT == 1, RANGE = RMODE == 1;
    // per warp: (2*u14 | 2*v14 << 16) of the last 8 luma rows, and one staged input row (W luma + CW + CW chroma bytes)
    __shared__ __align__(16) unsigned s_ring[kWarps420][ROWLOCAL ? 1 : 8][CW];  // unused for the row-local outputs
    __shared__ __align__(16) unsigned char s_in[kWarps420][2 * W];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned(*ring)[CW] = s_ring[wib];
    unsigned char *sY = s_in[wib], *sU = sY + W, *sV = sU + CW;

    const int w = K.w, h = K.h, ch = h >> 1, nw = K.tiles_per_row;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    const FastK fk = K.fk;
    const LutC LC = make_lutc(K);
    const float4 *tab = K.lut.tab;
    // yy = ((Y*8 - yO8) * yC) >> 16
    const int yA = 8 * f.yC, yB = -f.yO8 * f.yC;
    const int taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
    // output constants of the 10-bit 4:2:2 leg, derived from the 8-bit ones (same sums, different rounding term and shift)
    const unsigned kY10 = fk.kY - (24u << 9), kU10 = fk.kU + (8u << 10), kV10 = fk.kV + (8u << 10);
    // yuv444p: per-pixel chroma, rounding (256 << 14) + (1 << 8), one float bias per channel, output rounding (32 << 9)
    const unsigned kU444 = (unsigned)((256 << 14) + (1 << 8) + (32 << 9)) - kMagic * (unsigned)(b.ru + b.gu + b.bu);
    const unsigned kV444 = (unsigned)((256 << 14) + (1 << 8) + (32 << 9)) - kMagic * (unsigned)(b.rv + b.gv + b.bv);
    // Full-range OUTPUT: swscale converts the 15-bit luma / chroma values before its output stage, min((v * coeff + offset) >> 14,
    // 32767) with the 8-bit constants (the result can be NEGATIVE for chroma below the limited range: signed 16-bit fields in the
    // ring).  The limited-range instantiations keep their folded constants, the full-range ones spell the steps out.
    constexpr bool out_full = !OUT10 && RMODE == 2;
    const int lc = K.any.out_lc, lo = K.any.out_lo, cc = K.any.out_cc, co = K.any.out_co;
    auto range15 = [](int v, int coeff, int offset) { return min((v * coeff + offset) >> 14, 32767); };
    // 15-bit luma of the full-range path: 2 * y14, y14 = (S + (32 << 14) + (1 << 8)) >> 9 (fk.kY carries a further 32 << 9)
    auto luma_full = [&](unsigned sy) {
        const int y15 = 2 * ((int)(sy + fk.kY - (32u << 9)) >> 9);
        return (unsigned)__vimin_s32_relu((range15(y15, lc, lo) + 64) >> 7, 255);
    };
    auto chroma_full8 = [&](int c15) { return (unsigned)__vimin_s32_relu((range15(c15, cc, co) + 64) >> 7, 255); };
    // staging loads: NK == 2: every lane brings 4 luma + 2 + 2 chroma bytes; NK == 3, 4: the first W / 8 lanes bring 8 + 4 + 4
    constexpr int YB = NK == 2 ? 4 : 8, CB = YB / 2;
    using YLoad = typename std::conditional<NK == 2, unsigned, uint2>::type;
    using CLoad = typename std::conditional<NK == 2, unsigned short, unsigned>::type;
    const bool loader = YB * lane < W;

    // Work split: the chroma rows of all (frame, window) columns, in the order ((frame * nw + window) * ch + j), are cut into one
    // contiguous range per warp -- equal shares (+-1 row), so the persistent grid has no tail, and each range pays the 6 halo luma
    // rows of the vertical filter once per column it touches (none at the frame's top and bottom, where the filter clamps).
    const unsigned total_c = (unsigned)K.nframes * (unsigned)nw * (unsigned)ch;
    const unsigned nwarps = gridDim.x * kWarps420, gw = blockIdx.x * kWarps420 + wib;
    const unsigned per = total_c / nwarps, rem = total_c % nwarps;
    unsigned c_pos = gw * per + min(gw, rem);
    const unsigned c_end = c_pos + per + (gw < rem ? 1u : 0u);

    while (c_pos < c_end) {
        const unsigned col = c_pos / (unsigned)ch, j0 = c_pos - col * (unsigned)ch;
        const unsigned j1 = min((unsigned)ch, j0 + (c_end - c_pos));
        const unsigned frame = col / (unsigned)nw, win = col - frame * (unsigned)nw;
        c_pos += j1 - j0;
        // the last window is shifted left to end at the row end: it recomputes a few columns of its neighbour (same values), and
        // no lane ever falls outside the row
        const int x0 = min((int)win * W, w - W);
        const int y_begin = 2 * (int)j0, y_end = 2 * (int)j1;
        // 4:2:2 output has no vertical filter: no halo rows
        const int r_first = ROWLOCAL ? y_begin : max(0, y_begin - 3), r_last = ROWLOCAL ? y_end - 1 : min(h - 1, y_end + 2);
        // this lane's share of a row for the staging loads, and its outputs
        const uint8_t *pY = K.p.in[0] + (long)frame * K.p.in_step + (long)r_first * K.p.in_stride[0] + x0 + YB * lane;
        const uint8_t *bU = K.p.in[1] + (long)frame * K.p.in_step + (x0 >> 1) + CB * lane;
        const uint8_t *bV = K.p.in[2] + (long)frame * K.p.in_step + (x0 >> 1) + CB * lane;
        // outputs: yuv420p: luma pair (2 B) per step, 4 chroma bytes per lane and chroma row; yuv422p10le: luma pair (4 B) and one
        // chroma sample per plane (2 B) per step and luma row
        uint8_t *oY = K.p.out[0] + (long)frame * K.p.out_step + (long)r_first * K.p.out_stride[0] + (OUT10 ? 2 : 1) * (x0 + 2 * lane);
        // chroma bytes of this lane per step: yuv422p10le 2 (one sample), yuv422p 1, yuv444p 2 (two samples); yuv420p: 4 per chroma row
        const long c_off = OUT == 1 ? 2 * ((x0 >> 1) + lane) : (OUT == 2 ? (x0 >> 1) + lane : (OUT == 3 ? x0 + 2 * lane : (x0 >> 1) + 4 * lane));
        uint8_t *oU = K.p.out[1] + (long)frame * K.p.out_step + (long)(ROWLOCAL ? r_first : (y_begin >> 1)) * K.p.out_stride[1] + c_off;
        uint8_t *oV = K.p.out[2] + (long)frame * K.p.out_step + (long)(ROWLOCAL ? r_first : (y_begin >> 1)) * K.p.out_stride[2] + c_off;

        auto load_y = [&](const uint8_t *p) { return *reinterpret_cast<const YLoad *>(p); };
        auto load_c = [&](const uint8_t *base, int y, int cstride) {
            return *reinterpret_cast<const CLoad *>(base + (long)(IN422 ? y : (y >> 1)) * cstride);
        };
        YLoad ny{};
        CLoad nu{}, nv{};
        if (loader) ny = load_y(pY), nu = load_c(bU, r_first, K.p.in_stride[1]), nv = load_c(bV, r_first, K.p.in_stride[2]);

        int y = r_first;
        for (int j = y_begin >> 1; j < (y_end >> 1); ++j) {
            const int last_needed = ROWLOCAL ? 2 * j + 1 : min(h - 1, 2 * j + 4);
            for (; y <= last_needed; ++y) {
                __syncwarp();  // everybody is done with the previous row's staged samples
                if (loader) {
                    *reinterpret_cast<YLoad *>(sY + YB * lane) = ny;
                    if (IN422 || !(y & 1) || y == r_first) {  // 4:2:0: an odd row re-uses the chroma row its even neighbour staged
                        *reinterpret_cast<CLoad *>(sU + CB * lane) = nu;
                        *reinterpret_cast<CLoad *>(sV + CB * lane) = nv;
                    }
                }
                if (loader && y < r_last) {  // prefetch the next row into registers
                    pY += K.p.in_stride[0];
                    ny = load_y(pY);
                    if (IN422 || (y & 1)) nu = load_c(bU, y + 1, K.p.in_stride[1]), nv = load_c(bV, y + 1, K.p.in_stride[2]);
                }
                __syncwarp();
                const bool own_row = y >= y_begin && y < y_end;
#pragma unroll(NK == 4 ? 2 : (NK == 3 ? 1 : NK))
                for (int kk = 0; kk < NK; ++kk) {
                    // this lane's pixel pair: columns x0 + 64*kk + 2*lane, +1
                    const unsigned y2 = *reinterpret_cast<const unsigned short *>(sY + 64 * kk + 2 * lane);
                    int Yv[2] = {(int)(y2 & 0xFF), (int)(y2 >> 8)};
                    int Uv = sU[32 * kk + lane], Vv = sV[32 * kk + lane];
                    if (RANGE) Yv[0] = pc2tv_y(Yv[0]), Yv[1] = pc2tv_y(Yv[1]), Uv = pc2tv_c(Uv), Vv = pc2tv_c(Vv);
                    // yuv2rgb (x86 SIMD form): replicated chroma, 13-bit coefficients
                    const int uu = Uv * 8 - 1024, vv = Vv * 8 - 1024;
                    const int cr = (vv * f.vr) >> 16, cg = ((uu * f.ug) >> 16) + ((vv * f.vg) >> 16), cb = (uu * f.ub) >> 16;
                    unsigned mr[2], mg[2], mb[2];
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const int yy = (Yv[p] * yA + yB) >> 16;
                        mr[p] = (unsigned)__viaddmin_s32_relu(yy, cr, 255) | kMagic;
                        mg[p] = (unsigned)__viaddmin_s32_relu(yy, cg, 255) | kMagic;
                        mb[p] = (unsigned)__viaddmin_s32_relu(yy, cb, 255) | kMagic;
                    }
                    unsigned ro[2], go[2], bo[2];
                    lut3d_pair<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, true>(LC, tab, mr, mg, mb, ro, go, bo);
                    const unsigned sy0 = (unsigned)b.ry * ro[0] + (unsigned)b.gy * go[0] + (unsigned)b.by * bo[0];
                    const unsigned sy1 = (unsigned)b.ry * ro[1] + (unsigned)b.gy * go[1] + (unsigned)b.by * bo[1];
                    // rgb24ToUV_half on the pair sums (each sum carries 2 x the 0x4B000000 bias, folded into kU / kV)
                    const unsigned rs = ro[0] + ro[1], gs = go[0] + go[1], bs = bo[0] + bo[1];
                    const unsigned su = (unsigned)b.ru * rs + (unsigned)b.gu * gs + (unsigned)b.bu * bs;
                    const unsigned sv = (unsigned)b.rv * rs + (unsigned)b.gv * gs + (unsigned)b.bv * bs;
                    if (OUT10) {
                        // rgb24 -> yuv422p10le: clip10((2*y14 + 16) >> 5), y14 = S >> 9  ==  (S + (8 << 9)) >> 13;
                        //                       clip10((2*c14 + 16) >> 5), c14 = S >> 10 ==  (S + (8 << 10)) >> 14
                        const unsigned yq0 = (unsigned)__vimin_s32_relu((int)(sy0 + kY10) >> 13, 1023);
                        const unsigned yq1 = (unsigned)__vimin_s32_relu((int)(sy1 + kY10) >> 13, 1023);
                        stg_u32(oY + 128 * kk, yq0 | (yq1 << 16));
                        stg_u16(oU + 64 * kk, (unsigned)__vimin_s32_relu((int)(su + kU10) >> 14, 1023));
                        stg_u16(oV + 64 * kk, (unsigned)__vimin_s32_relu((int)(sv + kV10) >> 14, 1023));
                    } else if (OUT == 2 || OUT == 3) {
                        // 8-bit row-local outputs: luma as for yuv420p; chroma clip8((2*c14 + 64) >> 7) with
                        //   yuv422p: c14 = S_pair >> 10 (rgb24ToUV_half)  ==  (S_pair + kU + (32 << 10)) >> 16
                        //   yuv444p: c14 = S_pixel >> 9 (rgb24ToUV)       ==  (S_pixel + kU444 + (32 << 9)) >> 15
                        const unsigned yq0 = out_full ? luma_full(sy0) : (unsigned)((int)(sy0 + fk.kY) >> 15);
                        const unsigned yq1 = out_full ? luma_full(sy1) : (unsigned)((int)(sy1 + fk.kY) >> 15);
                        stg_u16(oY + 64 * kk, yq0 | (yq1 << 8));
                        if (OUT == 2) {
                            if (out_full) {
                                stg_u8(oU + 32 * kk, chroma_full8(2 * ((int)(su + fk.kU) >> 10)));
                                stg_u8(oV + 32 * kk, chroma_full8(2 * ((int)(sv + fk.kV) >> 10)));
                            } else {
                                stg_u8(oU + 32 * kk, (unsigned)__vimin_s32_relu((int)(su + fk.kU + (32u << 10)) >> 16, 255));
                                stg_u8(oV + 32 * kk, (unsigned)__vimin_s32_relu((int)(sv + fk.kV + (32u << 10)) >> 16, 255));
                            }
                        } else {
                            // per-pixel sums carry ONE 0x4B000000 bias per channel: half of the pair constants' bias term
                            const unsigned kU1 = kU444, kV1 = kV444;
                            unsigned uq[2], vq[2];
#pragma unroll
                            for (int p = 0; p < 2; ++p) {
                                const unsigned s1u = (unsigned)b.ru * ro[p] + (unsigned)b.gu * go[p] + (unsigned)b.bu * bo[p];
                                const unsigned s1v = (unsigned)b.rv * ro[p] + (unsigned)b.gv * go[p] + (unsigned)b.bv * bo[p];
                                if (out_full) {  // c15 = 2 * c14, c14 = (S + (256 << 14) + (1 << 8)) >> 9
                                    uq[p] = chroma_full8(2 * ((int)(s1u + kU1 - (32u << 9)) >> 9));
                                    vq[p] = chroma_full8(2 * ((int)(s1v + kV1 - (32u << 9)) >> 9));
                                } else {
                                    uq[p] = (unsigned)__vimin_s32_relu((int)(s1u + kU1) >> 15, 255);
                                    vq[p] = (unsigned)__vimin_s32_relu((int)(s1v + kV1) >> 15, 255);
                                }
                            }
                            stg_u16(oU + 64 * kk, uq[0] | (uq[1] << 8));
                            stg_u16(oV + 64 * kk, vq[0] | (vq[1] << 8));
                        }
                    } else {
                        // rgb24ToY: clip8((2*y14 + 64) >> 7), y14 = S >> 9  ==  (S + (32 << 9)) >> 15; no clip can bind (host-checked)
                        const unsigned yq0 = out_full ? luma_full(sy0) : (unsigned)((int)(sy0 + fk.kY) >> 15);
                        const unsigned yq1 = out_full ? luma_full(sy1) : (unsigned)((int)(sy1 + fk.kY) >> 15);
                        if (own_row) stg_u16(oY + 64 * kk, yq0 | (yq1 << 8));
                        unsigned u15 = 2u * (unsigned)((int)(su + fk.kU) >> 10), v15 = 2u * (unsigned)((int)(sv + fk.kV) >> 10);
                        if (out_full) u15 = (unsigned)range15((int)u15, cc, co) & 0xFFFFu, v15 = (unsigned)range15((int)v15, cc, co);
                        ring[y & 7][32 * kk + lane] = (v15 << 16) + u15;
                    }
                }
                if (ROWLOCAL) oU += K.p.out_stride[1], oV += K.p.out_stride[2];
                oY += K.p.out_stride[0];
            }
            if (ROWLOCAL) continue;
            __syncwarp();
            // vertical 8-tap decimation for chroma row j over luma rows 2j-3 .. 2j+4; lane = chroma columns 4*lane .. +3
            if (8 * lane < W) {
                int au[4], av[4];
                const bool exact = (j == ch - 1);  // swscale's C vertical scaler handles the last output row: no per-tap truncation
#pragma unroll
                for (int c = 0; c < 4; ++c) au[c] = av[c] = exact ? (64 << 12) : 7;
                if (2 * j - 3 >= 0 && 2 * j + 4 <= h - 1) {
                    // interior: 8 distinct rows, per-tap truncation (the SIMD vertical scaler): (c * tap) >> 16 is the high word of
                    // c * (tap << 16), one multiply-high with accumulate per tap
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(&ring[(2 * j - 3 + k) & 7][4 * lane]);
                        const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            au[c] += __mulhi(out_full ? (int)(short)(qq[c] & 0xFFFF) : (int)(qq[c] & 0xFFFF), taps[k] << 16);
                            av[c] += __mulhi(out_full ? (int)qq[c] >> 16 : (int)(qq[c] >> 16), taps[k] << 16);
                        }
                    }
                } else {
                    // rows clamped to the frame: taps that land on the same edge row are summed into one coefficient first
                    int fold = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int row = max(0, min(h - 1, 2 * j - 3 + k));
                        const int nrow = max(0, min(h - 1, 2 * j - 2 + k));
                        fold += taps[k];
                        if (k == 7 || nrow != row) {
                            const uint4 q = *reinterpret_cast<const uint4 *>(&ring[row & 7][4 * lane]);
                            const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int cu = out_full ? (int)(short)(qq[c] & 0xFFFF) : (int)(qq[c] & 0xFFFF);
                                const int cv = out_full ? (int)qq[c] >> 16 : (int)(qq[c] >> 16);
                                if (exact) au[c] += cu * fold, av[c] += cv * fold;
                                else au[c] += (cu * fold) >> 16, av[c] += (cv * fold) >> 16;
                            }
                            fold = 0;
                        }
                    }
                }
                unsigned uo = 0, vo = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uo |= (unsigned)__vimin_s32_relu(au[c] >> (exact ? 19 : 3), 255) << (8 * c);
                    vo |= (unsigned)__vimin_s32_relu(av[c] >> (exact ? 19 : 3), 255) << (8 * c);
                }
                stg_u32(oU, uo);
                stg_u32(oV, vo);
            }
            oU += K.p.out_stride[1], oV += K.p.out_stride[2];
            // the next rows overwrite ring slots this filter has just read: ordered by the __syncwarp at the top of the row loop
        }
    }
}

}  // namespace

}  // namespace lutr
