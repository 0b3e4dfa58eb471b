// TMA-staged warp-window kernel for yuv422p10le -> lut3d -> yuv422p10le (BASELINE cfg 3 / 4): the headline kernel.
//
// Arithmetic and pass structure are those of k422_warp.cu (reference chain: src/lut_renderer/ffmpeg.py:211-247, 304-310; integer
// legs and float32 lut3d as defined in DESIGN.md "Arithmetic").  What differs is how rows reach the warp and how work is dealt:
//
//  * every warp owns a two-deep ring of row buffers in shared memory and an mbarrier per buffer; one elected lane issues three 1-D
//    bulk async copies (TMA: cp.async.bulk.shared.global, complete_tx on the mbarrier) for the NEXT row of the warp's window -- luma,
//    U, V -- while the warp computes the current one.  No per-lane load instructions, no prefetch registers, no register rotation
//    at the end of a row, and the luma row is already where pass B / C want it.
//  * because the staged row can start at any 16-byte boundary of the source and be read back at any offset, the window no longer
//    has to start on an 8-column boundary of the lanes: it owns 248 of its 256 computed columns (4 halo columns left, 4 right: what
//    the 8-tap output filter needs) instead of 240 -- 3.2 % recomputed columns instead of 6.7 %.
//  * the warps of a CTA take (unit, row slice) tasks from a shared-memory counter instead of a fixed slice of every unit: a warp the
//    scheduler favours simply does more tasks, all warps of a CTA finish within one task of each other (the static split left a
//    sixth of the warp slots empty over the run: ncu sm__warps_active 20.0 of 24).
//
// Requirements checked by the launcher (otherwise k422_warp runs): width a multiple of 16, source planes and strides 16-byte aligned.
#include <cuda_runtime.h>

#include "lutr_internal.h"
#include "warp_common.cuh"

namespace lutr {

namespace {

constexpr int kOwnT = 248;    // output columns a window owns
constexpr int kYStage = 264;  // luma samples staged per row: columns own - 8 .. own + 255 (33 x 16 bytes)
constexpr int kCStage = 144;  // chroma samples staged per row and plane: from the 16-byte boundary at or below sample own/2 - 4

constexpr int kRingU = 2 * kYStage, kRingV = kRingU + 2 * kCStage, kRing = kRingV + 2 * kCStage;  // byte offsets inside a ring slot

struct __align__(16) WarpSmem {
    unsigned sC[kWin];                   // per column (cu | cv << 16), later (u14 | v14 << 16); 4-word chunks, swizzled
    unsigned char ring[2][kRing];        // row ring: luma | U | V of one row per slot (written by the copy engines only)
    unsigned short sYo[kYStage];         // output luma of the current row (pass B -> pass C), indexed like the staged luma
    unsigned long long bar[2];           // one mbarrier per ring slot (BULK)
};
static_assert(sizeof(WarpSmem) == 3776 && sizeof(WarpSmem) % 16 == 0 && kRing % 16 == 0, "row ring layout");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// COPY: the memory path alone (same staging, windows, halo and stores; the arithmetic replaced by moving samples), for the roofline
// study in DESIGN.md -- never dispatched by the product path.
// BULK: rows staged by bulk async copies issued by one elected lane (TMA, mbarrier completion); otherwise every lane stages its own
// 16-byte chunks with cp.async (LDGSTS, completion by wait_group): same ring, same layout, far fewer instructions per row.
template <int INTERP, bool FMA, bool CLAMP, int IDX, bool NOMIN, int NSZ, int MINB, int UNR, int BM, bool COPY, bool BULK>
__global__ void __launch_bounds__(kWarps * 32, MINB) k422_tma(const StageK K) {
    __shared__ WarpSmem s_all[kWarps];
    __shared__ unsigned s_next;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lane_odd = lane ^ 4;
    WarpSmem &S = s_all[wib];
    unsigned *sC = S.sC;

    const int w = K.w, h = K.h, tpr = K.tiles_per_row, R = K.rows_per_cta;
    const int nrb = (h + R - 1) / R;  // row blocks per frame
    const unsigned units = (unsigned)K.nframes * (unsigned)nrb * (unsigned)tpr;
    const FastK fk = K.fk;

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
    C.ry = (unsigned)K.b.ry, C.gy = (unsigned)K.b.gy, C.by = (unsigned)K.b.by;
    // u14 = S >> 11 == (32 * S) >> 16 (0 <= 32 * S < 2^31 for 10-bit RGB): the chroma rows of the matrix come pre-scaled by 32 from
    // the host, the two 15-bit terms are then the high halves of two registers and one PRMT builds the packed word pass C reads
    C.ru = fk.cu5[0], C.gu = fk.cu5[1], C.bu = fk.cu5[2], C.rv = fk.cv5[0], C.gv = fk.cv5[1], C.bv = fk.cv5[2];
    C.kY = fk.kY, C.kU = fk.kU5, C.kV = fk.kV5;
    const float4 *tab = K.lut.tab;

    const unsigned bar0 = smem_u32(&S.bar[0]), ring0 = smem_u32(&S.ring[0][0]);
    if (threadIdx.x == 0) s_next = 0u;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier: task counter and mbarriers are initialised

    unsigned g = 0;  // rows this warp has staged so far: ring slot = g & 1, mbarrier phase = (g >> 1) & 1
    for (;;) {
        // next task of the CTA: (unit ordinal, slice of 1/8 of the unit's rows); units are dealt to CTAs round-robin
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&s_next, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        const unsigned u = blockIdx.x + (t >> 3) * gridDim.x;
        if (u >= units) break;
        const unsigned tx = u % (unsigned)tpr, rb = (u / (unsigned)tpr) % (unsigned)nrb, frame = u / ((unsigned)tpr * (unsigned)nrb);
        const int y0 = (int)rb * R + (int)(t & 7u), y1 = min(h, (int)rb * R + R);
        if (y0 >= y1) continue;

        const int own = (int)tx * kOwnT;  // first output column of the window
        const int ws = own - 4;           // first computed column
        const bool edge_tile = ws < 0 || ws + kWin > w;
        // staged source ranges, clipped to the row (all bounds are multiples of 8 samples = 16 bytes)
        const int ylo = max(own - 8, 0), yhi = min(own + 256, w);
        const int cs0 = ((own - 8) & ~15) >> 1;        // first staged chroma sample (window coordinates, may be negative)
        const int csh = (own >> 1) - 4 - cs0;          // 0 or 4: where the window's first chroma sample sits in the staged row
        const int cw = w >> 1;
        const int clo = max(cs0, 0), chi = min(cs0 + kCStage, cw);
        const unsigned ybytes = 2u * (unsigned)(yhi - ylo), cbytes = 2u * (unsigned)(chi - clo);
        const uint8_t *pY = K.p.in[0] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[0] + 2 * ylo;
        const uint8_t *pU = K.p.in[1] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[1] + 2 * clo;
        const uint8_t *pV = K.p.in[2] + (long)frame * K.p.in_step + (long)y0 * K.p.in_stride[2] + 2 * clo;
        const int nout = min(kOwnT, w - own);
        const bool do_c = 8 * lane < nout;
        uint8_t *oY = K.p.out[0] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[0] + 2 * (long)(own + 8 * lane);
        uint8_t *oU = K.p.out[1] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[1] + (long)(own + 8 * lane);
        uint8_t *oV = K.p.out[2] + (long)frame * K.p.out_step + (long)y0 * K.p.out_stride[2] + (long)(own + 8 * lane);

        // per-lane staging (!BULK): lane l copies luma chunk l (columns own-8+8l .. +7; lane 0 also chunk 32) and, for l < 18, chunk l
        // of either chroma plane; chunks outside the row are skipped (every row bound is a multiple of 8 samples)
        const int ycol = own - 8 + 8 * lane, ccol = cs0 + 8 * lane;
        const bool y_in = ycol >= 0 && ycol < w, y2_in = lane == 0 && own + 248 < w, c_in = lane < kCStage / 8 && ccol >= 0 && ccol < cw;
        const long lane_y = 2l * (ycol - ylo), lane_c = 2l * (ccol - clo);  // byte offsets from the row pointers
        auto issue = [&](unsigned slot) {  // stage the row the pointers stand on into ring slot `slot`
            // The ring slots are only ever WRITTEN by the copy engines (pass B keeps its output luma in sYo), so the refill has a
            // write-after-read hazard only: every lane's reads of the slot's previous row have returned (their values were consumed
            // a row ago) and the warp meets before the copies are issued -- the consumer-release / producer pattern, which needs
            // no proxy fence (a fence.proxy.async here costs a MEMBAR.ALL.CTA per row and waits for the row's stores).
            __syncwarp();
            if (BULK) {
                if (lane == 0) {
                    const unsigned bar = bar0 + 8u * slot;
                    mbar_expect_tx(bar, ybytes + 2u * cbytes);
                    const unsigned rp = ring0 + slot * kRing;
                    bulk_g2s(rp + 2u * (unsigned)(ylo - (own - 8)), pY, ybytes, bar);
                    bulk_g2s(rp + kRingU + 2u * (unsigned)(clo - cs0), pU, cbytes, bar);
                    bulk_g2s(rp + kRingV + 2u * (unsigned)(clo - cs0), pV, cbytes, bar);
                }
            } else {
                const unsigned rp = ring0 + slot * kRing + 16u * lane;  // this lane's chunk of every plane
                if (y_in) cp_async16(rp, pY + lane_y);
                if (y2_in) cp_async16(rp + 512, pY + lane_y + 512);
                if (c_in) cp_async16(rp + kRingU, pU + lane_c), cp_async16(rp + kRingV, pV + lane_c);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        };
        issue(g & 1u);
        for (int y = y0; y < y1; y += kWarps, ++g) {
            const unsigned slot = g & 1u;
            if (y + kWarps < y1) {  // the warp's next row of the window goes into the other slot while this one is computed
                pY += kWarps * (long)K.p.in_stride[0], pU += kWarps * (long)K.p.in_stride[1], pV += kWarps * (long)K.p.in_stride[2];
                issue(slot ^ 1u);
            }
            if (BULK) {
                mbar_wait(bar0 + 8u * slot, (g >> 1) & 1u);
            } else {
                if (y + kWarps < y1) asm volatile("cp.async.wait_group 1;" ::: "memory");  // all but the row just issued
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();  // every lane's chunks of this row have landed
            }
            unsigned char *rp = S.ring[slot];
            unsigned short *sY = reinterpret_cast<unsigned short *>(rp) + 4;  // sY[j]: luma of window column j (the staged row starts 4 columns earlier)
            unsigned short *sU = reinterpret_cast<unsigned short *>(rp + kRingU), *sV = reinterpret_cast<unsigned short *>(rp + kRingV);
            if (edge_tile) {  // warp-uniform: replicate the row's first / last chroma sample into the staged positions outside it
                for (int p = lane; p < kCStage; p += 32) {
                    const int ci = cs0 + p;
                    if (ci < 0) sU[p] = sU[-cs0], sV[p] = sV[-cs0];
                    if (ci >= cw) sU[p] = sU[cw - 1 - cs0], sV[p] = sV[cw - 1 - cs0];
                }
                __syncwarp();
            }

            // ---------------- pass A: chroma upsample for window columns 8*lane .. 8*lane + 7 ----------------------------------
            {
                // chroma window e[0..7] = C[i0-2 .. i0+5], i0 = (ws + 8*lane) / 2: staged samples csh + 4*lane .. + 7
                const uint2 ua = *reinterpret_cast<const uint2 *>(sU + csh + 4 * lane), ub = *reinterpret_cast<const uint2 *>(sU + csh + 4 * lane + 4);
                const uint2 va = *reinterpret_cast<const uint2 *>(sV + csh + 4 * lane), vb = *reinterpret_cast<const uint2 *>(sV + csh + 4 * lane + 4);
                const unsigned wu[4] = {ua.x, ua.y, ub.x, ub.y}, wvv[4] = {va.x, va.y, vb.x, vb.y};
                unsigned wv[8];
                if (COPY) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const unsigned sel = ((k >> 1) & 1) ? 0x7632u : 0x5410u;
                        wv[k] = __byte_perm(wu[1 + (k >> 2)], wvv[1 + (k >> 2)], sel);
                    }
                } else {
                    chroma_up8<4>(wu, wvv, ws < 0 && lane == 0, wv);  // lane 0 of the row's first window holds row columns 0 and 2 as its columns 4 and 6
                }
                __syncwarp();  // pass C of the previous row has finished reading sC
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                *reinterpret_cast<uint4 *>(sC + 4 * swz_chunk(2 * lane + 1)) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
                __syncwarp();
            }

            // ---------------- pass B: one column per lane, window columns 32*k + lane ------------------------------------------
            if (!COPY) {
#pragma unroll(UNR)
                for (int kk = 0; kk < 8; kk += 2) {
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
                    lut3d_pair<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, true, BM>(LC, tab, mr, mg, mb, ro, go, bo);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const int j = 32 * (kk + p) + lane;
                        // Y: clip10((min(2*y14, 32767) + 16) >> 5), y14 = S >> 11  ==  (S + (8 << 11)) >> 15 (host-checked: no clip binds)
                        const unsigned yq = (unsigned)((int)(C.ry * ro[p] + C.gy * go[p] + C.by * bo[p] + C.kY) >> 15);
                        const unsigned su = C.ru * ro[p] + C.gu * go[p] + C.bu * bo[p] + C.kU, sv = C.rv * ro[p] + C.gv * go[p] + C.bv * bo[p] + C.kV;
                        sC[32 * (kk + p) + (p ? lane_odd : lane)] = __byte_perm(su, sv, 0x7632u);  // (v14 << 16) + u14, in place
                        S.sYo[j + 4] = (unsigned short)yq;  // pass C stores the window's own columns 16 bytes per lane
                    }
                }
                __syncwarp();
            }
            if (edge_tile) {  // warp-uniform: columns outside the row take the first / last column's chroma terms
                const int last = w - 1 - ws;  // window slot of the row's last column (>= 255 when the window ends inside the row)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = 32 * q + lane;
                    if (ws + j < 0) sC[swz_slot(j)] = sC[swz_slot(-ws)];
                    if (j > last) sC[swz_slot(j)] = sC[swz_slot(last)];
                }
                __syncwarp();
            }

            // ---------------- pass C: 8-tap decimation, lane = the 4 chroma samples of window columns 4 + 8*lane .. + 7 --------------
            if (do_c) {
                // tt[0..15] = window slots 8*lane .. 8*lane + 15; chroma sample j sits on slot 4 + 2j and covers slots 2j+1 .. 2j+8
                unsigned tt[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 v4 = *reinterpret_cast<const uint4 *>(sC + 4 * swz_chunk(2 * lane + q));
                    tt[4 * q] = v4.x, tt[4 * q + 1] = v4.y, tt[4 * q + 2] = v4.z, tt[4 * q + 3] = v4.w;
                }
                unsigned uo[2], vo[2];
                if (COPY) {
                    uo[0] = __byte_perm(tt[4], tt[6], 0x5410u), uo[1] = __byte_perm(tt[8], tt[10], 0x5410u);
                    vo[0] = __byte_perm(tt[4], tt[6], 0x7632u), vo[1] = __byte_perm(tt[8], tt[10], 0x7632u);
                } else {
                    int dus[4], dvs[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // symmetric taps: add mirrored columns while packed (u and v sums stay below 2^16)
                        const unsigned s0 = tt[2 * j + 1] + tt[2 * j + 8], s1 = tt[2 * j + 2] + tt[2 * j + 7];
                        const unsigned s2 = tt[2 * j + 3] + tt[2 * j + 6], s3 = tt[2 * j + 4] + tt[2 * j + 5];
                        // V from the high halves; U = (filter over the packed words) - 65536 * V  (mod 2^32; the true value fits)
                        const int dv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16);
                        const unsigned e = 0u - 230u * s0 - 692u * s1 + 1972u * s2 + 7142u * s3 + (16u << 13);
                        const int du = (int)(e - 65536u * (unsigned)dv);
                        // clip10((min(S >> 13, 32767) + 16) >> 5) == clamp((S + (16 << 13)) >> 18, 0, 1023); |S >> 18| < 2^13
                        dus[j] = du >> 18, dvs[j] = (dv + (16 << 13)) >> 18;
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {  // pack two samples, then clamp both halves to [0, 1023] with one instruction
                        uo[q] = __vimin_s16x2_relu(__byte_perm((unsigned)dus[2 * q], (unsigned)dus[2 * q + 1], 0x5410u), 0x03FF03FFu);
                        vo[q] = __vimin_s16x2_relu(__byte_perm((unsigned)dvs[2 * q], (unsigned)dvs[2 * q + 1], 0x5410u), 0x03FF03FFu);
                    }
                }
                stg_s16(oY, *reinterpret_cast<const uint4 *>((COPY ? sY + 4 : S.sYo + 8) + 8 * lane));  // luma of row columns own + 8*lane .. + 7
                stg_s8(oU, make_uint2(uo[0], uo[1]));
                stg_s8(oV, make_uint2(vo[0], vo[1]));
            }
            oY += kWarps * (long)K.p.out_stride[0], oU += kWarps * (long)K.p.out_stride[1], oV += kWarps * (long)K.p.out_stride[2];
        }
    }
}

}  // namespace

// ---- dispatch ---------------------------------------------------------------------------------------------------------------
// Specialised on the LUT size (33 / 65: node strides become immediates) for the default semantics; everything else runs the
// size-generic instantiation.  IDX 1 = LUT scale folded into the input scale (size-1 a power of two) with the coordinate clip
// proven dead on the host, IDX 2 = general.
using KernelFn = void (*)(const StageK);

template <int INTERP, bool FMA, bool CLAMP, int NSZ, int MINB, int UNR, bool BULK>
static KernelFn pick_idx(bool fold_nomin) {
    if (fold_nomin) return k422_tma<INTERP, FMA, CLAMP, 1, true, NSZ, MINB, UNR, LUTR_BMODE, false, BULK>;
    return k422_tma<INTERP, FMA, CLAMP, 2, false, NSZ, MINB, UNR, LUTR_BMODE, false, BULK>;
}

template <int INTERP, int MINB, int UNR, bool BULK>
static KernelFn pick_sem(bool fma, bool clamp, bool fold_nomin, int n) {
    if (!clamp) {
        if (fma) return n == 65 ? pick_idx<INTERP, true, false, 65, MINB, UNR, BULK>(fold_nomin) : pick_idx<INTERP, true, false, 0, MINB, UNR, BULK>(fold_nomin);
        return n == 65 ? pick_idx<INTERP, false, false, 65, MINB, UNR, BULK>(fold_nomin) : pick_idx<INTERP, false, false, 0, MINB, UNR, BULK>(fold_nomin);
    }
    return fma ? pick_idx<INTERP, true, true, 0, MINB, UNR, BULK>(fold_nomin) : pick_idx<INTERP, false, true, 0, MINB, UNR, BULK>(fold_nomin);
}

template <bool BULK>
static KernelFn pick_interp(int interp, bool fma, bool clamp, bool fold_nomin, int n) {
    // tetrahedral: two columns per step, all four steps unrolled (8 gathers in flight); trilinear has 16 gathers per step
    return interp == LUTR_INTERP_TRILINEAR ? pick_sem<LUTR_INTERP_TRILINEAR, 2, 1, BULK>(fma, clamp, fold_nomin, n)
                                           : pick_sem<LUTR_INTERP_TETRAHEDRAL, 3, 4, BULK>(fma, clamp, fold_nomin, n);
}

// true when the geometry allows the bulk copies (16-byte aligned starts and lengths on the source side)
bool k422_tma_ok(const StageK &k) {
    if (k.w % 16 != 0 || k.w < 16) return false;
    for (int p = 0; p < 3; ++p)
        if (((uintptr_t)k.p.in[p] % 16) != 0 || (k.p.in_stride[p] % 16) != 0) return false;
    return k.nframes <= 1 || (k.p.in_step % 16) == 0;
}

int launch_k422_tma(const StageK &k0, int interp, int float_mode, void *stream) {
    StageK k = k0;
    k.tiles_per_row = (k.w + kOwnT - 1) / kOwnT;
    const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
    const bool fold_nomin = k.fk.idx_fold && k.wk_nomin;
    KernelFn fn = nullptr;
    // the memory path alone (roofline study): same staging, windows and stores, no arithmetic
    if (k.wk_var == 1000) fn = k422_tma<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, 3, 4, LUTR_BMODE, true, true>;
    if (k.wk_var == 1001) fn = k422_tma<LUTR_INTERP_TETRAHEDRAL, false, false, 1, true, 65, 3, 4, LUTR_BMODE, true, false>;
    if (!fn) fn = k.wk_bulk ? pick_interp<true>(interp, fma, clamp, fold_nomin, k.lut.n) : pick_interp<false>(interp, fma, clamp, fold_nomin, k.lut.n);
    const int minb = interp == LUTR_INTERP_TRILINEAR ? 2 : 3;
    const long total = (long)((k.h + k.rows_per_cta - 1) / k.rows_per_cta) * k.nframes * k.tiles_per_row;
    long grid = k.grid_ctas > 0 ? k.grid_ctas : 148 * minb;
    if (grid > total) grid = total;
    if (k.wk_carveout > 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, k.wk_carveout);  // experiments: % of the unified L1 / shared array
    fn<<<(unsigned)grid, kWarps * 32, 0, (cudaStream_t)stream>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace lutr
