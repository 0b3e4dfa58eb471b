// Warp-autonomous fused kernel for yuv422p10le -> lut3d -> yuv422p10le (BASELINE cfg 3 / 4).
//
// Same arithmetic as stage_kernels.cu (reference chain: src/lut_renderer/ffmpeg.py:211-247, 304-310; integer legs and
// float32 lut3d as defined in DESIGN.md "Arithmetic"), different decomposition:
//
//  * work unit = one warp x one 128-column window of one row.  The window owns the outputs of its middle 120 columns and
//    recomputes 4 halo columns on each side, so nothing is ever exchanged between warps: no __syncthreads, no row-sized
//    shared-memory buffers (768 B per warp), the unified L1 stays almost entirely cache for the LUT gathers.
//  * pass A (lane = 4 consecutive columns): 4-tap chroma upsample, packed (cu | cv << 16) per column to shared memory.
//  * pass B (lane = ONE column, 32 consecutive columns per warp instruction): YUV->RGB matrix, lut3d, RGB->YUV.  Adjacent
//    lanes hold adjacent pixels, so one LUT gather instruction touches ~4-5 distinct 128-byte lines instead of ~9 with
//    4-column lanes (L1 wavefronts are the co-limiter of this stage next to instruction issue).
//  * pass C (lane = 2 chroma samples): 8-tap horizontal decimation of the per-column chroma terms.
//  * float work uses Blackwell's packed FP32 instructions (FFMA2 / FADD2 / FMUL2 with a broadcast scalar operand): exact
//    per-element IEEE rounding, half the issue slots.
//  * the next window's samples are prefetched into registers while the current one is computed.
#include <cuda_runtime.h>

#include "lutr_internal.h"
#include "warp_common.cuh"

namespace lutr {

namespace {

#ifndef LUTR_K422_WARPS
#define LUTR_K422_WARPS kWarps
#endif
constexpr int kW422 = LUTR_K422_WARPS;  // warps per CTA (experiments: 7 x 3 CTAs or 10 x 2 CTAs trade warps for registers; 8 x 3 shipped)

// IN444: the source is yuv444p10le (ProRes 4444 and the like): one chroma sample per pixel, no upsampling filter; everything behind
// pass A is unchanged (u17 = (C - 512) * 128 == 4 * (C << 5) - 65536, the same expression pass B applies to the 15-bit value).
// COPY: the memory path alone (same windows, halo loads, register prefetch, shared-memory round trip and 16-byte / 8-byte stores; the
// arithmetic replaced by moving samples) for the roofline study in DESIGN.md section 6 -- never dispatched by the product path.
template <int INTERP, bool FMA, bool CLAMP, int IDX, bool NOMIN, int NSZ, bool PK, int MINB, int UNR, bool IN444, int BM = LUTR_BMODE, int TEXV = 0, bool COPY = false>
__global__ void __launch_bounds__(kW422 * 32, MINB) k422_warp(const StageK K) {
    __shared__ __align__(16) unsigned s_all[kW422][kWin + kWin / 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lane_odd = lane ^ 4;
    unsigned *sC = s_all[wib];                                                  // [256] (cu | cv << 16), later (u14 | v14 << 16)
    unsigned *sYw = sC + kWin;                                                  // [128] luma samples, two per word
    unsigned short *sY = reinterpret_cast<unsigned short *>(sYw);               // [256]

    const int w = K.w, h = K.h, tpr = K.tiles_per_row, R = K.rows_per_cta;
    const int nrb = (h + R - 1) / R;  // row blocks per frame
    const unsigned total = (unsigned)K.nframes * (unsigned)nrb * (unsigned)tpr;
    const FastK fk = K.fk;
    const Rgb2YuvK b = K.b;

    struct {
        unsigned yA, kR, kG, kB, vr4, vgx, ug4, ub4, ubx;                  // YUV->RGB on the biased packed chroma word (pass B)
        unsigned ry, gy, by, ru, gu, bu, rv, gv, bv, kY, kU, kV;           // RGB->YUV with the 0x4B000000 bias folded in
    } C;
    const LutC LC = make_lutc(K);
    C.yA = fk.yA;
    C.vr4 = fk.vr4, C.ug4 = fk.ug4, C.ub4 = fk.ub4;
    C.vgx = fk.vg4 - 65536u * fk.ug4;
    C.ubx = 0u - 65536u * fk.ub4;
    C.kR = fk.kR - (unsigned)kCBias * fk.vr4;
    C.kG = fk.kG - (unsigned)kCBias * fk.vg4 - (unsigned)kCBias * fk.ug4;
    C.kB = fk.kB - (unsigned)kCBias * fk.ub4;
    C.ry = (unsigned)b.ry, C.gy = (unsigned)b.gy, C.by = (unsigned)b.by;
    // u14 = S >> 11 == (32 * S) >> 16 (0 <= 32 * S < 2^31 for 10-bit RGB): the chroma rows of the matrix come pre-scaled by 32 from the
    // host, the two 15-bit terms are then the high halves of two registers and one PRMT builds the packed word pass C reads
    C.ru = fk.cu5[0], C.gu = fk.cu5[1], C.bu = fk.cu5[2], C.rv = fk.cv5[0], C.gv = fk.cv5[1], C.bv = fk.cv5[2];
    C.kU = fk.kU5, C.kV = fk.kV5;
    C.kY = fk.kY;
    const float4 *tab = K.lut.tab;

    // unit = (frame, block of R rows, window), owned by a CTA: its 8 warps take the rows y0 + wib, y0 + wib + 8, ... so that the
    // warps of an SM work on neighbouring rows of the same image region (similar colours -> the LUT lines they gather stay in L1).
    // No synchronisation between the warps; a CTA walks units u, u + gridDim.x, ... and keeps the coordinates incrementally.
    unsigned u = blockIdx.x;
    if (u >= total) return;
    unsigned tx = u % (unsigned)tpr, rb = (u / (unsigned)tpr) % (unsigned)nrb, frame = u / ((unsigned)tpr * (unsigned)nrb);
    const unsigned stride = gridDim.x;
    const unsigned dtx = stride % (unsigned)tpr, drb = (stride / (unsigned)tpr) % (unsigned)nrb, dfr = stride / ((unsigned)tpr * (unsigned)nrb);

    for (; u < total; u += stride) {
        const int a = (int)tx * kOut - 8;  // first column of the window
        const bool edge_tile = tx == 0 || tx == (unsigned)tpr - 1;
        const int y0 = (int)rb * R + wib, y1 = min(h, (int)rb * R + R);
        const int x0 = a + 8 * lane, xc = min(max(x0, 0), w - 8);  // this lane's 8 columns in pass A (clamped into the row)
        const uint8_t *pY = K.p.in[0] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[0] + 2 * xc;
        const uint8_t *pU = K.p.in[1] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[1] + (IN444 ? 2 * xc : xc);
        const uint8_t *pV = K.p.in[2] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[2] + (IN444 ? 2 * xc : xc);
        uint8_t *oY = K.p.out[0] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[0] + 2 * (long)(a + 8 + 8 * lane);
        uint8_t *oU = K.p.out[1] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[1] + (long)(a + 8 + 8 * lane);
        uint8_t *oV = K.p.out[2] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[2] + (long)(a + 8 + 8 * lane);
        const int nout = min(kOut, w - (a + 8));  // columns a+8 .. a+8+nout-1 are this window's outputs
        const bool do_c = 8 * lane < nout;

        uint4 ny = make_uint4(0u, 0u, 0u, 0u);
        uint4 nu = make_uint4(0u, 0u, 0u, 0u), nv = make_uint4(0u, 0u, 0u, 0u);  // 4:2:2: 4 samples in .x .y; 4:4:4: 8 samples
        auto load_c = [](const uint8_t *p) {
            if (IN444) return ldg_s16(p);
            const uint2 t = ldg_s8(p);
            return make_uint4(t.x, t.y, 0u, 0u);
        };
        if (y0 < y1) ny = ldg_s16(pY), nu = load_c(pU), nv = load_c(pV);
        for (int y = y0; y < y1; y += kW422) {
            const uint4 cy = ny;
            const uint4 cu4 = nu, cv4 = nv;
            const uint2 cu2 = make_uint2(cu4.x, cu4.y), cv2 = make_uint2(cv4.x, cv4.y);
            if (y + kW422 < y1) {  // prefetch this warp's next row of the window
                pY += kW422 * (long)K.p.in_stride[0], pU += kW422 * (long)K.p.in_stride[1], pV += kW422 * (long)K.p.in_stride[2];
                ny = ldg_s16(pY), nu = load_c(pU), nv = load_c(pV);
            }

            // ---------------- pass A: chroma upsample for columns a + 8*lane .. + 7 ------------------------------------
            if (IN444) {  // one chroma sample per column: c15 = C << 5 (never clips), biased and packed like the filtered form
                const unsigned uw[4] = {cu4.x, cu4.y, cu4.z, cu4.w}, vw[4] = {cv4.x, cv4.y, cv4.z, cv4.w};
                unsigned wv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned U = (uw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu, V = (vw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                    wv[k] = (((V << 5) + (unsigned)kCBias) << 16) | ((U << 5) + (unsigned)kCBias);
                }
                __syncwarp();  // pass C of the previous row has finished reading sC
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane + 1)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                *reinterpret_cast<uint4 *>(sYw + 4 * lane) = cy;
                __syncwarp();
            } else {
                // chroma words of the neighbouring lanes: (C[i0-2], C[i0-1]) and (C[i0+4], C[i0+5])
                unsigned ul = __shfl_up_sync(0xFFFFFFFFu, cu2.y, 1), vl = __shfl_up_sync(0xFFFFFFFFu, cv2.y, 1);
                unsigned ur = __shfl_down_sync(0xFFFFFFFFu, cu2.x, 1), vr = __shfl_down_sync(0xFFFFFFFFu, cv2.x, 1);
                if (edge_tile) {  // warp-uniform: replicate the first / last chroma sample beyond the row ends
                    if (xc == 0) ul = (cu2.x & 0xFFFFu) * 0x10001u, vl = (cv2.x & 0xFFFFu) * 0x10001u;
                    if (xc == w - 8) ur = (cu2.y >> 16) * 0x10001u, vr = (cv2.y >> 16) * 0x10001u;
                }
                // chroma window C[i0-2 .. i0+5] as packed pairs; exact float filter (chroma_up8, warp_common.cuh)
                const unsigned wu[4] = {ul, cu2.x, cu2.y, ur}, wvv[4] = {vl, cv2.x, cv2.y, vr};
                unsigned wv[8];
                if (COPY) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) wv[k] = __byte_perm(wu[1 + (k >> 2)], wvv[1 + (k >> 2)], ((k >> 1) & 1) ? 0x7632u : 0x5410u);
                } else {
                    chroma_up8<0>(wu, wvv, edge_tile && xc == 0, wv);
                }
                __syncwarp();  // pass C of the previous row has finished reading sC
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane + 1)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                *reinterpret_cast<uint4 *>(sYw + 4 * lane) = cy;
                __syncwarp();
            }

            // ---------------- pass B: one column per lane, columns a + 32*k + lane ---------------------------------------
#pragma unroll(UNR)
            for (int kk = 0; kk < (COPY ? 0 : 8); kk += 2) {
                unsigned mr[2], mg[2], mb[2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int j = 32 * (kk + p) + lane;
                    const unsigned Y = sY[j], wv = sC[32 * (kk + p) + (p ? lane_odd : lane)];  // == sC[swz_slot(j)]: kk is even
                    const unsigned cvb = wv >> 16;
                    // R = Y*yA + cv*vr4 + kR, G = .. + cv*vg4 + cu*ug4, B = .. + cu*ub4 (wrapping 32-bit, as the C code); the
                    // packed word stands in for cu: wv*g == cu*g + 65536*cv*g (mod 2^32), the surplus is folded into the cv gain
                    const int Rv = (int)(cvb * C.vr4 + (Y * C.yA + C.kR));
                    const int Gv = (int)(wv * C.ug4 + (cvb * C.vgx + (Y * C.yA + C.kG)));
                    const int Bv = (int)(cvb * C.ubx + (wv * C.ub4 + (Y * C.yA + C.kB)));
                    // av_clip_uintp2(x, 30) >> 20, then 0x4B000000 | code (one funnel shift)
                    mr[p] = clip30_to_magic(Rv);
                    mg[p] = clip30_to_magic(Gv);
                    mb[p] = clip30_to_magic(Bv);
                }
                unsigned ro[2], go[2], bo[2];
                lut3d_pair<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, PK, PK ? BM : 0, TEXV>(LC, tab, mr, mg, mb, ro, go, bo);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int j = 32 * (kk + p) + lane;
                    // Y: clip10((min(2*y14, 32767) + 16) >> 5), y14 = S >> 11  ==  (S + (8 << 11)) >> 15 (host-checked: no clip binds)
                    const unsigned yq = (unsigned)((int)(C.ry * ro[p] + C.gy * go[p] + C.by * bo[p] + C.kY) >> 15);
                    const unsigned su = C.ru * ro[p] + C.gu * go[p] + C.bu * bo[p] + C.kU, sv = C.rv * ro[p] + C.gv * go[p] + C.bv * bo[p] + C.kV;
                    // (v14 << 16) + u14 with u14 = S >> 11;  the same lane read this slot above: in place
                    sC[32 * (kk + p) + (p ? lane_odd : lane)] = __byte_perm(su, sv, 0x7632u);
                    sY[j] = (unsigned short)yq;  // in place as well; pass C stores the window's own columns 16 bytes per lane
                }
            }
            __syncwarp();
            if (edge_tile) {  // warp-uniform: columns outside the row take the first / last column's chroma terms
                const int last = w - 1 - a;  // window slot of the row's last column (>= 255 when the window ends inside the row)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = 32 * q + lane;
                    if (a + j < 0) sC[swz_slot(j)] = sC[swz_slot(-a)];
                    if (j > last) sC[swz_slot(j)] = sC[swz_slot(last)];
                }
                __syncwarp();
            }

            // ---------------- pass C: 8-tap decimation, lane = chroma samples (a+8)/2 + 4*lane .. + 3 -------------------------
            if (do_c) {
                // tt[0..15] = slots 8*lane+4 .. 8*lane+19; chroma sample j covers luma columns 2i-3 .. 2i+4 = tt[2j+1 .. 2j+8]
                unsigned tt[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 v4 = *reinterpret_cast<const uint4 *>(sC + 4 * swz_chunk(2 * lane + 1 + q));
                    tt[4 * q] = v4.x, tt[4 * q + 1] = v4.y, tt[4 * q + 2] = v4.z, tt[4 * q + 3] = v4.w;
                }
                unsigned uo[2], vo[2];
                int dus[4], dvs[4];
#pragma unroll
                for (int j = 0; j < (COPY ? 0 : 4); ++j) {
                    // symmetric taps: add mirrored columns while packed (u and v sums stay below 2^16)
                    const unsigned s0 = tt[2 * j + 1] + tt[2 * j + 8], s1 = tt[2 * j + 2] + tt[2 * j + 7];
                    const unsigned s2 = tt[2 * j + 3] + tt[2 * j + 6], s3 = tt[2 * j + 4] + tt[2 * j + 5];
                    // V from the high halves; U = (filter over the packed words) - 65536 * V  (mod 2^32; the true value fits)
                    const int dv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16);
                    const unsigned e = 0u - 230u * s0 - 692u * s1 + 1972u * s2 + 7142u * s3 + (16u << 13);
                    const int du = (int)(e - 65536u * (unsigned)dv);
                    // clip10((min(S >> 13, 32767) + 16) >> 5) == clamp((S + (16 << 13)) >> 18, 0, 1023)
                    dus[j] = du >> 18, dvs[j] = (dv + (16 << 13)) >> 18;  // |.| < 2^13: they fit the halves of one register
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {  // pack two samples, then clamp both halves to [0, 1023] with one instruction
                    if (COPY) {
                        uo[q] = __byte_perm(tt[4 * q + 4], tt[4 * q + 6], 0x5410u), vo[q] = __byte_perm(tt[4 * q + 4], tt[4 * q + 6], 0x7632u);
                        continue;
                    }
                    uo[q] = __vimin_s16x2_relu(__byte_perm((unsigned)dus[2 * q], (unsigned)dus[2 * q + 1], 0x5410u), 0x03FF03FFu);
                    vo[q] = __vimin_s16x2_relu(__byte_perm((unsigned)dvs[2 * q], (unsigned)dvs[2 * q + 1], 0x5410u), 0x03FF03FFu);
                }
                stg_s16(oY, *reinterpret_cast<const uint4 *>(sYw + 4 * lane + 4));  // luma of columns a+8+8*lane .. +7
                stg_s8(oU, make_uint2(uo[0], uo[1]));
                stg_s8(oV, make_uint2(vo[0], vo[1]));
            }
            oY += kW422 * (long)K.p.out_stride[0], oU += kW422 * (long)K.p.out_stride[1], oV += kW422 * (long)K.p.out_stride[2];
        }

        // next unit
        tx += dtx;
        unsigned c = tx >= (unsigned)tpr ? 1u : 0u;
        tx -= c * (unsigned)tpr;
        rb += drb + c;
        c = rb >= (unsigned)nrb ? 1u : 0u;
        rb -= c * (unsigned)nrb;
        frame += dfr + c;
    }
}

}  // namespace

// ---- dispatch ---------------------------------------------------------------------------------------------------------------
// Specialised on the LUT size (33 / 65: node strides become immediates) for the default semantics only; everything else runs
// the size-generic instantiation.  IDX 1 = LUT scale folded into the input scale (size-1 a power of two) with the coordinate
// clip proven dead on the host, IDX 2 = general.
using KernelFn = void (*)(const StageK);

template <int INTERP, bool FMA, bool CLAMP, int NSZ, bool PK, int MINB, int UNR>
static KernelFn pick_idx(bool fold_nomin, bool in444) {
    if (in444) {
        if (fold_nomin) return k422_warp<INTERP, FMA, CLAMP, 1, true, NSZ, PK, MINB, UNR, true>;
        return k422_warp<INTERP, FMA, CLAMP, 2, false, NSZ, PK, MINB, UNR, true>;
    }
    if (fold_nomin) return k422_warp<INTERP, FMA, CLAMP, 1, true, NSZ, PK, MINB, UNR, false>;
    return k422_warp<INTERP, FMA, CLAMP, 2, false, NSZ, PK, MINB, UNR, false>;
}

template <int INTERP, bool PK, int MINB, int UNR>
static KernelFn pick_sem(bool fma, bool clamp, bool fold_nomin, int n, bool in444) {
    if (!fma && !clamp) {
        if (n == 65) return pick_idx<INTERP, false, false, 65, PK, MINB, UNR>(fold_nomin, in444);
        if (n == 33) return pick_idx<INTERP, false, false, 33, PK, MINB, UNR>(fold_nomin, in444);
        return pick_idx<INTERP, false, false, 0, PK, MINB, UNR>(fold_nomin, in444);
    }
    if (!fma) return pick_idx<INTERP, false, true, 0, PK, MINB, UNR>(fold_nomin, in444);
    return clamp ? pick_idx<INTERP, true, true, 0, PK, MINB, UNR>(fold_nomin, in444) : pick_idx<INTERP, true, false, 0, PK, MINB, UNR>(fold_nomin, in444);
}

int launch_k422_warp(const StageK &k, bool in444, int interp, int float_mode, void *stream) {
    const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
    const bool fold_nomin = k.fk.idx_fold && k.wk_nomin;
    KernelFn fn = nullptr;
#ifdef LUTR_WK_EXPERIMENTS
    // variants of the headline configuration, selected by the stage option wk_var = 100 * BM (blend mode) / 2001, 2002 (texture path)
    if (interp == LUTR_INTERP_TETRAHEDRAL && !fma && !clamp && fold_nomin && k.lut.n == 65 && !in444 && k.wk_var >= 0) {
#define WK_CASE(B) \
    if (k.wk_var == 100 * B) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 3, 4, false, B>;
        WK_CASE(0) WK_CASE(1) WK_CASE(2)
#undef WK_CASE
        if (k.wk_var == 3002) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 2, 4, false>;  // 2 CTAs / SM, up to 128 registers
        if (k.wk_var == 3012) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 2, 2, false>;  // same, two steps unrolled
        if (k.wk_var == 2001 && k.lut.tex) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 3, 4, false, 1, 1>;  // c111 through TEX
        if (k.wk_var == 2002 && k.lut.tex) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 3, 4, false, 1, 2>;  // + second step
    }
#endif
    if (k.wk_var == 1004 && !in444) fn = k422_warp<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, true, 3, 4, false, LUTR_BMODE, 0, true>;  // memory path only
    if (!fn) {
        // tetrahedral: two columns per step, all four steps unrolled (8 gathers in flight); trilinear has 16 gathers per step
        fn = interp == LUTR_INTERP_TRILINEAR ? pick_sem<LUTR_INTERP_TRILINEAR, true, 2, 1>(fma, clamp, fold_nomin, k.lut.n, in444)
                                             : pick_sem<LUTR_INTERP_TETRAHEDRAL, true, 3, 4>(fma, clamp, fold_nomin, k.lut.n, in444);
    }
    const int minb = k.min_blocks >= 2 && k.min_blocks <= 4 ? k.min_blocks : 3;
    const long total = (long)((k.h + k.rows_per_cta - 1) / k.rows_per_cta) * k.nframes * k.tiles_per_row;
    long grid = k.grid_ctas > 0 ? k.grid_ctas : 148 * minb;
    if (grid > total) grid = total;
    if (k.wk_carveout > 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, k.wk_carveout);  // experiments: % of the unified L1 / shared array
    fn<<<(unsigned)grid, kW422 * 32, 0, (cudaStream_t)stream>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace lutr
