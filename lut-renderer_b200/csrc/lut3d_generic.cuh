// Generic per-pixel lut3d evaluation (all interpolation modes, per-channel domain scales, clamped "next" nodes): the
// straightforward restatement of FFmpeg vf_lut3d's C code used by the generic tile kernel (k_any.cuh).
// float32, one rounding per C operation unless FMA (the contracted order of FFmpeg's x86 assembly for planar formats).
// Reference call site: src/lut_renderer/ffmpeg.py:242-247.  Oracle: oracle/lutr_oracle.c lut3d_pixel / interp_*.
#pragma once
#include <cuda_runtime.h>

#include "lutr_internal.h"

namespace lutr {
namespace {

template <bool FMA>
__device__ __forceinline__ float madd(float a, float b, float c) {  // a*b + c
    if (FMA) return __fmaf_rn(a, b, c);
    return __fadd_rn(__fmul_rn(a, b), c);
}

template <bool FMA>
__device__ __forceinline__ float lerp1(float v0, float v1, float f) {  // v0 + (v1 - v0) * f
    return madd<FMA>(__fsub_rn(v1, v0), f, v0);
}

template <bool FMA>
__device__ __forceinline__ float4 lerp3(float4 a, float4 b, float f) {
    return make_float4(lerp1<FMA>(a.x, b.x, f), lerp1<FMA>(a.y, b.y, f), lerp1<FMA>(a.z, b.z, f), 0.f);
}

// int (0..2^depth-1) in, int out.  Body of lut3d's per-pixel loop:
//   s = clip(src * (1/max) * (scale*lut_max), 0, lut_max); prev = (int)s; next = min(prev+1, n-1); d = s - prev
//   out = clip((int)(interp(...) * max), 0, max)     (truncation, not rounding)
template <int INTERP, bool FMA>
__device__ __forceinline__ void lut3d_px(const LutK &L, int ri, int gi, int bi, int out_max, int &ro, int &go, int &bo) {
    const float sr = fminf(fmaxf(__fmul_rn(__fmul_rn(__int2float_rn(ri), L.in_scale), L.sr), 0.f), L.lut_max);
    const float sg = fminf(fmaxf(__fmul_rn(__fmul_rn(__int2float_rn(gi), L.in_scale), L.sg), 0.f), L.lut_max);
    const float sb = fminf(fmaxf(__fmul_rn(__fmul_rn(__int2float_rn(bi), L.in_scale), L.sb), 0.f), L.lut_max);
    float4 c;
    if (INTERP == LUTR_INTERP_NEAREST) {
        // NEAR(x) = (int)((double)x + .5); x <= 255 so the double sum is exact and float rz after an exact add matches
        const int nr = __double2int_rz((double)sr + .5), ng = __double2int_rz((double)sg + .5), nb = __double2int_rz((double)sb + .5);
        c = __ldg(L.tab + (nr * L.n2 + ng * L.n + nb));
    } else {
        const int pr = __float2int_rz(sr), pg = __float2int_rz(sg), pb = __float2int_rz(sb);
        const float dr = __fsub_rn(sr, __int2float_rn(pr)), dg = __fsub_rn(sg, __int2float_rn(pg)), db = __fsub_rn(sb, __int2float_rn(pb));
        const int base = pr * L.n2 + pg * L.n + pb;
        // offsets to the "next" node along each axis; 0 when prev is already the last node (next = min(prev+1, n-1))
        const int oR = pr < L.n - 1 ? L.n2 : 0, oG = pg < L.n - 1 ? L.n : 0, oB = pb < L.n - 1 ? 1 : 0;
        if (INTERP == LUTR_INTERP_TRILINEAR) {
            const float4 c000 = __ldg(L.tab + base), c001 = __ldg(L.tab + base + oB);
            const float4 c010 = __ldg(L.tab + base + oG), c011 = __ldg(L.tab + base + oG + oB);
            const float4 c100 = __ldg(L.tab + base + oR), c101 = __ldg(L.tab + base + oR + oB);
            const float4 c110 = __ldg(L.tab + base + oR + oG), c111 = __ldg(L.tab + base + oR + oG + oB);
            const float4 c00 = lerp3<FMA>(c000, c100, dr), c10 = lerp3<FMA>(c010, c110, dr);
            const float4 c01 = lerp3<FMA>(c001, c101, dr), c11 = lerp3<FMA>(c011, c111, dr);
            const float4 c0 = lerp3<FMA>(c00, c10, dg), c1 = lerp3<FMA>(c01, c11, dg);
            c = lerp3<FMA>(c0, c1, db);
        } else if (INTERP == LUTR_INTERP_PYRAMID || INTERP == LUTR_INTERP_PRISM) {
            // interp_pyramid / interp_prism: c000 + sum of (difference of nodes) * fraction terms, left to right, always with
            // separate multiplies and adds (FFmpeg has no assembly for these two modes)
            const float4 c000 = __ldg(L.tab + base), c111 = __ldg(L.tab + base + oR + oG + oB);
            auto term = [](float acc, float a1, float a0, float f) { return __fadd_rn(acc, __fmul_rn(__fsub_rn(a1, a0), f)); };
            auto cross = [](float acc, float d3, float d2, float d1, float d0, float g1, float g2) {
                return __fadd_rn(acc, __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(d3, d2), d1), d0), g1), g2));
            };
            if (INTERP == LUTR_INTERP_PYRAMID) {
                float4 a1, a0, b1, b0, e1, e0, x3, x2, x1;
                float g1, g2;
                if (dg > dr && db > dr) {
                    const float4 c001 = __ldg(L.tab + base + oB), c010 = __ldg(L.tab + base + oG), c011 = __ldg(L.tab + base + oG + oB);
                    a1 = c111, a0 = c011, b1 = c010, b0 = c000, e1 = c001, e0 = c000, x3 = c011, x2 = c001, x1 = c010, g1 = dg, g2 = db;
                } else if (dr > dg && db > dg) {
                    const float4 c001 = __ldg(L.tab + base + oB), c100 = __ldg(L.tab + base + oR), c101 = __ldg(L.tab + base + oR + oB);
                    a1 = c100, a0 = c000, b1 = c111, b0 = c101, e1 = c001, e0 = c000, x3 = c101, x2 = c001, x1 = c100, g1 = dr, g2 = db;
                } else {
                    const float4 c010 = __ldg(L.tab + base + oG), c110 = __ldg(L.tab + base + oR + oG), c100 = __ldg(L.tab + base + oR);
                    a1 = c100, a0 = c000, b1 = c010, b0 = c000, e1 = c111, e0 = c110, x3 = c110, x2 = c100, x1 = c010, g1 = dr, g2 = dg;
                }
                c.x = cross(term(term(term(c000.x, a1.x, a0.x, dr), b1.x, b0.x, dg), e1.x, e0.x, db), x3.x, x2.x, x1.x, c000.x, g1, g2);
                c.y = cross(term(term(term(c000.y, a1.y, a0.y, dr), b1.y, b0.y, dg), e1.y, e0.y, db), x3.y, x2.y, x1.y, c000.y, g1, g2);
                c.z = cross(term(term(term(c000.z, a1.z, a0.z, dr), b1.z, b0.z, dg), e1.z, e0.z, db), x3.z, x2.z, x1.z, c000.z, g1, g2);
            } else {
                const float4 c010 = __ldg(L.tab + base + oG), c101 = __ldg(L.tab + base + oR + oB);
                float4 a1, a0, b1, b0, p3, p2, p1, p0, q3, q2, q1, q0;
                if (db > dr) {
                    const float4 c001 = __ldg(L.tab + base + oB), c011 = __ldg(L.tab + base + oG + oB);
                    a1 = c001, a0 = c000, b1 = c101, b0 = c001, p3 = c000, p2 = c010, p1 = c001, p0 = c011, q3 = c001, q2 = c011, q1 = c101, q0 = c111;
                } else {
                    const float4 c110 = __ldg(L.tab + base + oR + oG), c100 = __ldg(L.tab + base + oR);
                    a1 = c101, a0 = c100, b1 = c100, b0 = c000, p3 = c100, p2 = c110, p1 = c101, p0 = c111, q3 = c000, q2 = c010, q1 = c100, q0 = c110;
                }
                c.x = cross(cross(term(term(term(c000.x, a1.x, a0.x, db), b1.x, b0.x, dr), c010.x, c000.x, dg), p3.x, p2.x, p1.x, p0.x, db, dg), q3.x, q2.x, q1.x, q0.x, dr, dg);
                c.y = cross(cross(term(term(term(c000.y, a1.y, a0.y, db), b1.y, b0.y, dr), c010.y, c000.y, dg), p3.y, p2.y, p1.y, p0.y, db, dg), q3.y, q2.y, q1.y, q0.y, dr, dg);
                c.z = cross(cross(term(term(term(c000.z, a1.z, a0.z, db), b1.z, b0.z, dr), c010.z, c000.z, dg), p3.z, p2.z, p1.z, p0.z, db, dg), q3.z, q2.z, q1.z, q0.z, dr, dg);
            }
        } else {
            // Tetrahedral.  All six branches of the C code have the form
            //   (1-d1)*c000 + (d1-d2)*c[e1] + (d2-d3)*c[e1+e2] + d3*c111   with d1 >= d2 >= d3 the sorted fractions.
            // Ties only ever swap two vertices whose weight is exactly 0, so any stable sort gives identical bits
            // (table entries are finite; the parser guarantees it).  Sort (d, offset) pairs without branches.
            const bool rg = dr > dg;
            const float hi = fmaxf(dr, dg), lo = fminf(dr, dg);
            const int ohi = rg ? oR : oG, olo = rg ? oG : oR;
            const bool bh = db > hi;
            const float d1 = fmaxf(hi, db), md = fminf(hi, db);
            const int o1 = bh ? oB : ohi, om = bh ? ohi : oB;
            const bool ml = md > lo;
            const float d2 = fmaxf(md, lo), d3 = fminf(md, lo);
            const int o2 = ml ? om : olo;
            const float4 c000 = __ldg(L.tab + base), cA = __ldg(L.tab + base + o1);
            const float4 cB = __ldg(L.tab + base + o1 + o2), c111 = __ldg(L.tab + base + oR + oG + oB);
            const float w0 = __fsub_rn(1.f, d1), w1 = __fsub_rn(d1, d2), w2 = __fsub_rn(d2, d3), w3 = d3;
            c.x = madd<FMA>(w3, c111.x, madd<FMA>(w2, cB.x, madd<FMA>(w1, cA.x, __fmul_rn(w0, c000.x))));
            c.y = madd<FMA>(w3, c111.y, madd<FMA>(w2, cB.y, madd<FMA>(w1, cA.y, __fmul_rn(w0, c000.y))));
            c.z = madd<FMA>(w3, c111.z, madd<FMA>(w2, cB.z, madd<FMA>(w1, cA.z, __fmul_rn(w0, c000.z))));
        }
    }
    const float om = L.out_scale;
    // (int) truncation then clip == clamp in float then truncate (NaN -> 0 either way)
    ro = __float2int_rz(fminf(fmaxf(__fmul_rn(c.x, om), 0.f), om));
    go = __float2int_rz(fminf(fmaxf(__fmul_rn(c.y, om), 0.f), om));
    bo = __float2int_rz(fminf(fmaxf(__fmul_rn(c.z, om), 0.f), om));
    (void)out_max;
}


}  // namespace
}  // namespace lutr
