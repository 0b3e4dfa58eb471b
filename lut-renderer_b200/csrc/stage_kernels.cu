// Fused colour-stage kernels for sm_100a.
//
// One launch = unpack -> (range pass) -> YUV->RGB -> 3D-LUT interpolation -> RGB->YUV -> chroma decimation -> pack,
// i.e. the whole `scale,format,lut3d,format` chain the reference hands to FFmpeg
// (reference src/lut_renderer/ffmpeg.py:211-247, 304-310), for a batch of device-resident frames.
//
// The arithmetic is integer fixed point (the swscale legs) and float32 without contraction (lut3d); it is
// defined in DESIGN.md "Arithmetic" and checked bit-for-bit against oracle/ in tests/.  This is an
// HBM-streaming / L1-gather kernel: no tensor cores on purpose.
#include <cuda_runtime.h>

#include "lutr_internal.h"
#include "lut3d_generic.cuh"

namespace lutr {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp_u(int v, int hi) { return __vimin_s32_relu(v, hi); }  // min(max(v,0),hi), 1 instr

__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ldg_stream4(const void *p) {
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream16(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream8(void *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream4(void *p, unsigned v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// yuv422p10le -> yuv422p10le   (BASELINE cfg 3 / 4)
// One CTA owns whole rows; a thread owns 8-pixel groups g = tid, tid + blockDim, ...  (16 B of Y, 8 B of U, 8 B of V).
// Phase 1 per group: 4-tap chroma upsample + matrix -> gbrp10 -> lut3d -> Y out, per-pixel chroma terms to smem.
// Phase 2 per group: 8-tap horizontal chroma decimation from smem -> U, V out.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int up4(int a, int b, int c, int d, int t0, int t1, int t2, int t3) {
    return min((a * t0 + b * t1 + c * t2 + d * t3) >> 9, 32767);
}

template <int INTERP, bool FMA>
__global__ void __launch_bounds__(512, 2) k_yuv422p10(const StageK K) {
    extern __shared__ unsigned s_c14[];  // [2][w]: (u14 | v14 << 16) per luma column, double buffered over rows
    const int w = K.w, cw = w >> 1, groups = w >> 3;
    const long total_rows = (long)K.h * K.nframes;
    const long row0 = (long)blockIdx.x * K.rows_per_cta;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    // constants of the 10-bit yuv2rgb: wrapping 32-bit arithmetic as in the C code
    const unsigned yA = (unsigned)(128 * f.yC), yK = (unsigned)(-f.yOff10 * f.yC) + (1u << 19);

    for (int rr = 0; rr < K.rows_per_cta; ++rr) {
        const long grow = row0 + rr;
        if (grow >= total_rows) break;
        const int frame = (int)(grow / K.h), y = (int)(grow - (long)frame * K.h);
        const uint8_t *inY = K.p.in[0] + frame * K.p.in_step + (long)y * K.p.in_stride[0];
        const uint8_t *inU = K.p.in[1] + frame * K.p.in_step + (long)y * K.p.in_stride[1];
        const uint8_t *inV = K.p.in[2] + frame * K.p.in_step + (long)y * K.p.in_stride[2];
        uint8_t *outY = K.p.out[0] + frame * K.p.out_step + (long)y * K.p.out_stride[0];
        uint8_t *outU = K.p.out[1] + frame * K.p.out_step + (long)y * K.p.out_stride[1];
        uint8_t *outV = K.p.out[2] + frame * K.p.out_step + (long)y * K.p.out_stride[2];
        unsigned *sc = s_c14 + (rr & 1) * w;

        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g << 3, i0 = g << 2;
            const uint4 yv = ldg_stream16(inY + 2 * x0);
            const uint2 uv = ldg_stream8(inU + 2 * i0), vv = ldg_stream8(inV + 2 * i0);
            // chroma window e[0..7] = C[i0-2 .. i0+5], clamped to the row
            int eu[8], ev[8];
            eu[2] = uv.x & 0xFFFF, eu[3] = uv.x >> 16, eu[4] = uv.y & 0xFFFF, eu[5] = uv.y >> 16;
            ev[2] = vv.x & 0xFFFF, ev[3] = vv.x >> 16, ev[4] = vv.y & 0xFFFF, ev[5] = vv.y >> 16;
            if (g > 0) {
                const unsigned lu = __ldg((const unsigned *)(inU + 2 * (i0 - 2))), lv = __ldg((const unsigned *)(inV + 2 * (i0 - 2)));
                eu[0] = lu & 0xFFFF, eu[1] = lu >> 16, ev[0] = lv & 0xFFFF, ev[1] = lv >> 16;
            } else {
                eu[0] = eu[1] = eu[2], ev[0] = ev[1] = ev[2];
            }
            if (g < groups - 1) {
                const unsigned ru = __ldg((const unsigned *)(inU + 2 * (i0 + 4))), rv = __ldg((const unsigned *)(inV + 2 * (i0 + 4)));
                eu[6] = ru & 0xFFFF, eu[7] = ru >> 16, ev[6] = rv & 0xFFFF, ev[7] = rv >> 16;
            } else {
                eu[6] = eu[7] = eu[5], ev[6] = ev[7] = ev[5];
            }
            const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w};
            unsigned yo[4];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int Y = (k & 1) ? (int)(yw[k >> 1] >> 16) : (int)(yw[k >> 1] & 0xFFFF);
                const int j = k >> 1;
                int cu, cv;
                if (!(k & 1)) {  // even column 2i: taps on C[i-2..i+1]
                    cu = up4(eu[j], eu[j + 1], eu[j + 2], eu[j + 3], -461, 3943, 14284, -1382);
                    cv = up4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3], -461, 3943, 14284, -1382);
                    if (g == 0 && k == 0) {  // the two left-most even columns use swscale's edge-normalised taps
                        cu = min((17729 * eu[2] - 1345 * eu[3]) >> 9, 32767);
                        cv = min((17729 * ev[2] - 1345 * ev[3]) >> 9, 32767);
                    } else if (g == 0 && k == 2) {
                        cu = min((3835 * eu[2] + 13894 * eu[3] - 1345 * eu[4]) >> 9, 32767);
                        cv = min((3835 * ev[2] + 13894 * ev[3] - 1345 * ev[4]) >> 9, 32767);
                    }
                } else {  // odd column 2i+1: taps on C[i-1..i+2]
                    cu = up4(eu[j + 1], eu[j + 2], eu[j + 3], eu[j + 4], -1382, 14284, 3943, -461);
                    cv = up4(ev[j + 1], ev[j + 2], ev[j + 3], ev[j + 4], -1382, 14284, 3943, -461);
                }
                // u17 = ((1<<9) - (128<<19) + c15*4096) >> 10 == 4*c15 - 65536
                const unsigned u17 = (unsigned)(4 * cu - 65536), v17 = (unsigned)(4 * cv - 65536);
                const unsigned yy = (unsigned)Y * yA + yK;
                const int R = (int)(yy + v17 * (unsigned)f.vr);
                const int G = (int)(yy + v17 * (unsigned)f.vg + u17 * (unsigned)f.ug);
                const int B = (int)(yy + u17 * (unsigned)f.ub);
                // av_clip_uintp2(x, 30) >> 20  ==  clamp(x >> 20, 0, 1023)
                const int r10 = clamp_u(R >> 20, 1023), g10 = clamp_u(G >> 20, 1023), b10 = clamp_u(B >> 20, 1023);
                int ro, go, bo;
                lut3d_px<INTERP, FMA>(K.lut, r10, g10, b10, 1023, ro, go, bo);
                // planar_rgb16_to_y + yuv2plane1_10
                const int y14 = (b.ry * ro + b.gy * go + b.by * bo + (16 << 17) + (1 << 10)) >> 11;
                const unsigned yq = (unsigned)clamp_u((min(2 * y14, 32767) + 16) >> 5, 1023);
                const int u14 = (b.ru * ro + b.gu * go + b.bu * bo + (128 << 17) + (1 << 10)) >> 11;
                const int v14 = (b.rv * ro + b.gv * go + b.bv * bo + (128 << 17) + (1 << 10)) >> 11;
                sc[x0 + k] = (unsigned)(u14 & 0xFFFF) | ((unsigned)v14 << 16);
                if (k & 1) yo[k >> 1] |= yq << 16;
                else yo[k >> 1] = yq;
            }
            stg_stream16(outY + 2 * x0, make_uint4(yo[0], yo[1], yo[2], yo[3]));
        }
        __syncthreads();
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g << 3, i0 = g << 2;
            // c14 window t[0..13] = columns x0-3 .. x0+10, clamped to the row
            int tu[14], tv[14];
#pragma unroll
            for (int k = 0; k < 14; ++k) {
                int x = x0 - 3 + k;
                x = max(0, min(x, w - 1));
                const unsigned pk = sc[x];
                tu[k] = (int)(short)(pk & 0xFFFF);
                tv[k] = (int)(pk >> 16);
            }
            unsigned uo[2], vo[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // chroma column i0+j covers luma columns 2i-3 .. 2i+4 = t[2j .. 2j+7]
                const int su = -230 * (tu[2 * j] + tu[2 * j + 7]) - 692 * (tu[2 * j + 1] + tu[2 * j + 6]) +
                               1972 * (tu[2 * j + 2] + tu[2 * j + 5]) + 7142 * (tu[2 * j + 3] + tu[2 * j + 4]);
                const int sv = -230 * (tv[2 * j] + tv[2 * j + 7]) - 692 * (tv[2 * j + 1] + tv[2 * j + 6]) +
                               1972 * (tv[2 * j + 2] + tv[2 * j + 5]) + 7142 * (tv[2 * j + 3] + tv[2 * j + 4]);
                const unsigned uq = (unsigned)clamp_u((min(su >> 13, 32767) + 16) >> 5, 1023);
                const unsigned vq = (unsigned)clamp_u((min(sv >> 13, 32767) + 16) >> 5, 1023);
                if (j & 1) uo[j >> 1] |= uq << 16, vo[j >> 1] |= vq << 16;
                else uo[j >> 1] = uq, vo[j >> 1] = vq;
            }
            stg_stream8(outU + 2 * i0, make_uint2(uo[0], uo[1]));
            stg_stream8(outV + 2 * i0, make_uint2(vo[0], vo[1]));
            (void)cw;
        }
        // next row writes the other half of s_c14; the barrier above orders this row's reads before the
        // row after next overwrites them (every thread passes the next barrier only after its reads here)
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path building blocks (used when all three channels share one domain scale and interp != nearest).
//
//  * index table in shared memory: code value (with the YUV->RGB clip folded into the index) -> {prev, d}; the entries are
//    computed on the host with lut3d's own float operations, so the lookup is exact by construction and replaces
//    I2F, 2 FMUL, 2 FMNMX, F2I, I2F, FADD per channel by one LDS.64.
//  * "next" nodes are NOT clamped: prev == n-1 only happens with d == 0, and every vertex that uses the step of a channel
//    with d == 0 has weight exactly 0 (tetrahedral: it is the smallest fraction; trilinear: lerp factor 0), so the
//    out-of-cell node only needs to be finite and inside the (zero-padded) buffer.  x + 0*finite == x bit for bit.
//  * fractions are non-negative floats, so they sort as integers: max3 / min3 / xor-median, no float compares.
//  * truncating store: float -> int by adding 2^23 with round-toward-zero; the integer stays biased by 0x4B000000 and the
//    bias is folded into the constants of the following integer matrix (wrapping 32-bit arithmetic).
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned kMagicBits = 0x4B000000u;  // bits of 8388608.0f

__device__ __forceinline__ float4 ldg_node(const char *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Exact lut3d prologue for one channel without I2F/F2I: x in [0, maxv] -> (prev biased by 0x4B000000, d).
//   fl(x * c) = fma(2^23 + x, c, -2^23 * c)   (the product is exact in the fma, one rounding like the plain multiply)
//   floor(s)  = (s + 2^23 rounded toward zero) - 2^23
template <int IDX>
__device__ __forceinline__ void idx_arith(int x, const FastK &fk, float lut_max, unsigned &prev_biased, unsigned &dbits) {
    const float m = __uint_as_float((unsigned)x | kMagicBits);
    float s = __fmaf_rn(m, fk.idx_mul, fk.idx_add);
    if (IDX == 2) s = __fmul_rn(s, fk.idx_sr);
    s = fminf(s, lut_max);
    const float t = __fadd_rz(s, 8388608.f);
    prev_biased = __float_as_uint(t);
    dbits = __float_as_uint(__fsub_rn(s, __fsub_rn(t, 8388608.f)));
}

// pr/pg/pb: prev node indices (biased by `bias` each, folded out through kbase); dr/dg/db: fraction bits.
template <int INTERP, bool FMA, bool CLAMP>
__device__ __forceinline__ void lut3d_fast(const char *tab, int n, int n2, unsigned kbase, int n2b, int nb, int total, float out_scale,
                                           unsigned pr, unsigned pg, unsigned pb, unsigned dr_, unsigned dg_, unsigned db_,
                                           unsigned &rb, unsigned &gb, unsigned &bb) {
    const unsigned base = pr * (unsigned)n2 + (pg * (unsigned)n + kbase) + pb;
    const char *p000 = tab + (size_t)base * 16u;
    float4 c;
    if (INTERP == LUTR_INTERP_TRILINEAR) {
        const float dr = __uint_as_float(dr_), dg = __uint_as_float(dg_), db = __uint_as_float(db_);
        const float4 c000 = ldg_node(p000), c001 = ldg_node(p000 + 16);
        const float4 c010 = ldg_node(p000 + nb), c011 = ldg_node(p000 + nb + 16);
        const float4 c100 = ldg_node(p000 + n2b), c101 = ldg_node(p000 + n2b + 16);
        const float4 c110 = ldg_node(p000 + n2b + nb), c111 = ldg_node(p000 + n2b + nb + 16);
        const float4 c00 = lerp3<FMA>(c000, c100, dr), c10 = lerp3<FMA>(c010, c110, dr);
        const float4 c01 = lerp3<FMA>(c001, c101, dr), c11 = lerp3<FMA>(c011, c111, dr);
        const float4 c0 = lerp3<FMA>(c00, c10, dg), c1 = lerp3<FMA>(c01, c11, dg);
        c = lerp3<FMA>(c0, c1, db);
    } else {
        const int ar = (int)dr_, ag = (int)dg_, ab = (int)db_;
        const int mx = __vimax3_s32(ar, ag, ab), mn = __vimin3_s32(ar, ag, ab);
        const int md = ar ^ ag ^ ab ^ mx ^ mn;
        // step of the largest fraction first, of the smallest last (tie order is irrelevant: tied steps carry weight 0)
        const int o1 = (ar == mx) ? n2b : ((ag == mx) ? nb : 16);
        const int o3 = (ab == mn) ? 16 : ((ag == mn) ? nb : n2b);
        const float4 c000 = ldg_node(p000), cA = ldg_node(p000 + o1), cB = ldg_node(p000 + (total - o3)), c111 = ldg_node(p000 + total);
        const float d1 = __int_as_float(mx), d2 = __int_as_float(md), d3 = __int_as_float(mn);
        const float w0 = __fsub_rn(1.f, d1), w1 = __fsub_rn(d1, d2), w2 = __fsub_rn(d2, d3), w3 = d3;
        c.x = madd<FMA>(w3, c111.x, madd<FMA>(w2, cB.x, madd<FMA>(w1, cA.x, __fmul_rn(w0, c000.x))));
        c.y = madd<FMA>(w3, c111.y, madd<FMA>(w2, cB.y, madd<FMA>(w1, cA.y, __fmul_rn(w0, c000.y))));
        c.z = madd<FMA>(w3, c111.z, madd<FMA>(w2, cB.z, madd<FMA>(w1, cA.z, __fmul_rn(w0, c000.z))));
    }
    float fr = __fmul_rn(c.x, out_scale), fg = __fmul_rn(c.y, out_scale), fb = __fmul_rn(c.z, out_scale);
    if (CLAMP) {
        fr = fminf(fmaxf(fr, 0.f), out_scale), fg = fminf(fmaxf(fg, 0.f), out_scale), fb = fminf(fmaxf(fb, 0.f), out_scale);
    }
    rb = __float_as_uint(__fadd_rz(fr, 8388608.f));
    gb = __float_as_uint(__fadd_rz(fg, 8388608.f));
    bb = __float_as_uint(__fadd_rz(fb, 8388608.f));
}

// yuv422p10le -> yuv422p10le, fast path.  Persistent CTAs: CTA b owns rows b, b + gridDim.x, ... of the (frame, row) space.
// A thread owns 4-pixel groups (8 B of Y, one 32-bit word = 2 samples of U and of V); the chroma words of the two
// neighbouring groups come from the adjacent lanes by warp shuffle (warp-edge lanes fetch them from L1/L2).
// IDX: 0 = index table in shared memory, 1 = arithmetic with the LUT scale folded (power of two), 2 = arithmetic, general
template <int INTERP, bool FMA, bool CLAMP, int IDX, int MINB>
__global__ void __launch_bounds__(256, MINB) k_yuv422p10_fast(const StageK K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *s_idx = reinterpret_cast<uint2 *>(smem_raw);                                                       // [4096] {prev, d} (IDX == 0)
    unsigned *s_c14 = reinterpret_cast<unsigned *>(smem_raw + (IDX == 0 ? kIdxEntries10 * sizeof(uint2) : 0));  // [2][w + 8], 4 pad each side
    const int w = K.w, groups = w >> 2, pitch = w + 8;
    if (IDX == 0) {
        for (int i = threadIdx.x; i < kIdxEntries10 / 2; i += blockDim.x)
            reinterpret_cast<uint4 *>(s_idx)[i] = __ldg(reinterpret_cast<const uint4 *>(K.idx_tab) + i);
        __syncthreads();
    }

    const Rgb2YuvK b = K.b;
    const FastK fk = K.fk;
    const char *tab = reinterpret_cast<const char *>(K.lut.tab);
    const int n = K.lut.n, n2 = K.lut.n2;
    const float out_scale = K.lut.out_scale, lut_max = K.lut.lut_max;
    const unsigned kbase = IDX == 0 ? 0u : 0u - kMagicBits * (unsigned)(n2 + n + 1);
    const long total_rows = (long)K.h * K.nframes;
    const int lane = threadIdx.x & 31;
    const int gloop = (groups + 31) & ~31;  // whole warps enter the loop so the shuffles are convergent
    int parity = 0;

    for (long grow = blockIdx.x; grow < total_rows; grow += gridDim.x, parity ^= 1) {
        const int frame = (int)(grow / K.h), y = (int)(grow - (long)frame * K.h);
        const uint8_t *inY = K.p.in[0] + frame * K.p.in_step + (long)y * K.p.in_stride[0];
        const uint8_t *inU = K.p.in[1] + frame * K.p.in_step + (long)y * K.p.in_stride[1];
        const uint8_t *inV = K.p.in[2] + frame * K.p.in_step + (long)y * K.p.in_stride[2];
        uint8_t *outY = K.p.out[0] + frame * K.p.out_step + (long)y * K.p.out_stride[0];
        uint8_t *outU = K.p.out[1] + frame * K.p.out_step + (long)y * K.p.out_stride[1];
        uint8_t *outV = K.p.out[2] + frame * K.p.out_step + (long)y * K.p.out_stride[2];
        unsigned *sc = s_c14 + parity * pitch + 4;  // sc[x] for x in [-4, w+4)

        for (int gi = threadIdx.x; gi < gloop; gi += blockDim.x) {
            const bool valid = gi < groups;
            const int g = valid ? gi : groups - 1;
            const int x0 = g << 2;
            const uint2 yv = ldg_stream8(inY + 2 * x0);
            const unsigned uw = ldg_stream4(inU + x0), vw = ldg_stream4(inV + x0);   // samples 2g, 2g+1
            unsigned ul = __shfl_up_sync(0xFFFFFFFFu, uw, 1), vl = __shfl_up_sync(0xFFFFFFFFu, vw, 1);
            unsigned ur = __shfl_down_sync(0xFFFFFFFFu, uw, 1), vr = __shfl_down_sync(0xFFFFFFFFu, vw, 1);
            if (g == 0) {  // clamp at the left edge of the row: C[-2] = C[-1] = C[0]
                ul = (uw & 0xFFFFu) * 0x10001u, vl = (vw & 0xFFFFu) * 0x10001u;
            } else if (lane == 0) {
                ul = __ldg((const unsigned *)(inU + x0 - 4)), vl = __ldg((const unsigned *)(inV + x0 - 4));
            }
            if (g == groups - 1) {
                ur = (uw >> 16) * 0x10001u, vr = (vw >> 16) * 0x10001u;
            } else if (lane == 31) {
                ur = __ldg((const unsigned *)(inU + x0 + 4)), vr = __ldg((const unsigned *)(inV + x0 + 4));
            }
            // chroma window e[0..5] = C[i0-2 .. i0+3]
            const int eu[6] = {(int)(ul & 0xFFFF), (int)(ul >> 16), (int)(uw & 0xFFFF), (int)(uw >> 16), (int)(ur & 0xFFFF), (int)(ur >> 16)};
            const int ev[6] = {(int)(vl & 0xFFFF), (int)(vl >> 16), (int)(vw & 0xFFFF), (int)(vw >> 16), (int)(vr & 0xFFFF), (int)(vr >> 16)};
            unsigned yo[2], pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned yw = k < 2 ? yv.x : yv.y;
                const unsigned Y = (k & 1) ? (yw >> 16) : (yw & 0xFFFF);
                const int j = k >> 1;
                int cu, cv;
                if (!(k & 1)) {  // even column 2i: taps on C[i-2 .. i+1]
                    cu = up4(eu[j], eu[j + 1], eu[j + 2], eu[j + 3], -461, 3943, 14284, -1382);
                    cv = up4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3], -461, 3943, 14284, -1382);
                    if (g == 0) {  // the two left-most even columns use swscale's edge-normalised taps
                        if (k == 0) {
                            cu = min((17729 * eu[2] - 1345 * eu[3]) >> 9, 32767);
                            cv = min((17729 * ev[2] - 1345 * ev[3]) >> 9, 32767);
                        } else {
                            cu = min((3835 * eu[2] + 13894 * eu[3] - 1345 * eu[4]) >> 9, 32767);
                            cv = min((3835 * ev[2] + 13894 * ev[3] - 1345 * ev[4]) >> 9, 32767);
                        }
                    }
                } else {  // odd column 2i+1: taps on C[i-1 .. i+2]
                    cu = up4(eu[j + 1], eu[j + 2], eu[j + 3], eu[j + 4], -1382, 14284, 3943, -461);
                    cv = up4(ev[j + 1], ev[j + 2], ev[j + 3], ev[j + 4], -1382, 14284, 3943, -461);
                }
                const unsigned yy = Y * fk.yA;
                const int R = (int)(yy + (unsigned)cv * fk.vr4 + fk.kR);
                const int G = (int)(yy + (unsigned)cv * fk.vg4 + (unsigned)cu * fk.ug4 + fk.kG);
                const int B = (int)(yy + (unsigned)cu * fk.ub4 + fk.kB);
                uint2 er, eg, eb;
                if (IDX == 0) {
                    // table index = 12-bit two's-complement code of (x >> 20): the clip to [0,1023] is part of the table
                    er = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(s_idx) + ((unsigned)(R >> 17) & 0x7FF8u));
                    eg = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(s_idx) + ((unsigned)(G >> 17) & 0x7FF8u));
                    eb = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(s_idx) + ((unsigned)(B >> 17) & 0x7FF8u));
                } else {
                    // av_clip_uintp2(x, 30) >> 20 == clamp(x >> 20, 0, 1023)
                    idx_arith<IDX>(clamp_u(R >> 20, 1023), fk, lut_max, er.x, er.y);
                    idx_arith<IDX>(clamp_u(G >> 20, 1023), fk, lut_max, eg.x, eg.y);
                    idx_arith<IDX>(clamp_u(B >> 20, 1023), fk, lut_max, eb.x, eb.y);
                }
                unsigned rb, gb, bb;
                lut3d_fast<INTERP, FMA, CLAMP>(tab, n, n2, kbase, fk.n2b, fk.nb, fk.totalb, out_scale, er.x, eg.x, eb.x, er.y, eg.y, eb.y,
                                               rb, gb, bb);
                // Y: clip10((min(2*y14, 32767) + 16) >> 5) with y14 = S >> 11  ==  (S + (8 << 11)) >> 15 (the min and the clip
                // never bind: S < 2^25 for 10-bit RGB with any of the luma rows; checked on the host)
                const unsigned yq = (unsigned)((int)((unsigned)b.ry * rb + (unsigned)b.gy * gb + (unsigned)b.by * bb + fk.kY) >> 15);
                const unsigned u14 = (unsigned)((int)((unsigned)b.ru * rb + (unsigned)b.gu * gb + (unsigned)b.bu * bb + fk.kU) >> 11);
                const unsigned v14 = (unsigned)((int)((unsigned)b.rv * rb + (unsigned)b.gv * gb + (unsigned)b.bv * bb + fk.kV) >> 11);
                pk[k] = (v14 << 16) + u14;
                if (k & 1) yo[k >> 1] |= yq << 16;
                else yo[k >> 1] = yq;
            }
            if (valid) {
                *reinterpret_cast<uint4 *>(sc + x0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (g == 0) *reinterpret_cast<uint4 *>(sc - 4) = make_uint4(pk[0], pk[0], pk[0], pk[0]);          // clamp left
                if (g == groups - 1) *reinterpret_cast<uint4 *>(sc + w) = make_uint4(pk[3], pk[3], pk[3], pk[3]);  // clamp right
                stg_stream8(outY + 2 * x0, make_uint2(yo[0], yo[1]));
            }
        }
        __syncthreads();
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g << 2;
            // t[0..11] = packed c14 of columns x0-4 .. x0+7; chroma column 2g+j needs columns x0+2j-3 .. x0+2j+4 = t[2j+1 .. 2j+8]
            unsigned t[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 v4 = *reinterpret_cast<const uint4 *>(sc + x0 - 4 + 4 * q);
                t[4 * q] = v4.x, t[4 * q + 1] = v4.y, t[4 * q + 2] = v4.z, t[4 * q + 3] = v4.w;
            }
            unsigned uo = 0, vo = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                // symmetric taps: add the mirrored columns while still packed (u and v sums stay below 2^16)
                const unsigned s0 = t[2 * j + 1] + t[2 * j + 8], s1 = t[2 * j + 2] + t[2 * j + 7];
                const unsigned s2 = t[2 * j + 3] + t[2 * j + 6], s3 = t[2 * j + 4] + t[2 * j + 5];
                const int su = -230 * (int)(s0 & 0xFFFF) - 692 * (int)(s1 & 0xFFFF) + 1972 * (int)(s2 & 0xFFFF) + 7142 * (int)(s3 & 0xFFFF) + (16 << 13);
                const int sv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16) + (16 << 13);
                // clip10((min(S >> 13, 32767) + 16) >> 5) == clamp((S + (16 << 13)) >> 18, 0, 1023)
                uo |= (unsigned)clamp_u(su >> 18, 1023) << (16 * j);
                vo |= (unsigned)clamp_u(sv >> 18, 1023) << (16 * j);
            }
            stg_stream4(outU + x0, uo);
            stg_stream4(outV + x0, vo);
        }
        // the next row uses the other half of s_c14; its barrier orders these reads before that half is reused
    }
}


// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant of the fast 4:2:2 kernel.
// One elected thread streams whole input rows (Y, U, V) into a 2-stage shared-memory ring with 1-D bulk async copies
// (cp.async.bulk + mbarrier complete_tx), one to two rows ahead of the math, so no warp ever waits on an HBM load; pixel
// inputs are then plain 32-bit-addressed LDS (no 64-bit address arithmetic, free chroma-window access at any alignment).
// ---------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int INTERP, bool FMA, bool CLAMP, int IDX, int MINB>
__global__ void __launch_bounds__(MINB == 2 ? 320 : 256, MINB) k_yuv422p10_tma(const StageK K) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int w = K.w, groups = w >> 2, pitch = w + 8;
    const unsigned row_bytes = 4u * (unsigned)w;  // Y: 2w, U: w, V: w
    // layout: [in stage 0 | in stage 1 | c14 buffer 0 | c14 buffer 1 | 2 mbarriers]
    unsigned char *s_in = smem_raw;
    unsigned *s_c14 = reinterpret_cast<unsigned *>(smem_raw + 2 * row_bytes);
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + 2 * row_bytes + 2 * pitch * sizeof(unsigned));
    const unsigned bar0 = smem_u32(s_bar), in0 = smem_u32(s_in);
    const long total_rows = (long)K.h * K.nframes;
    const long stride_rows = gridDim.x;

    auto issue_row = [&](long grow, int stage) {  // thread 0 only
        const int frame = (int)(grow / K.h), y = (int)(grow - (long)frame * K.h);
        const unsigned bar = bar0 + 8u * stage, dst = in0 + stage * row_bytes;
        mbar_expect_tx(bar, row_bytes);
        bulk_g2s(dst, K.p.in[0] + frame * K.p.in_step + (long)y * K.p.in_stride[0], 2u * w, bar);
        bulk_g2s(dst + 2u * w, K.p.in[1] + frame * K.p.in_step + (long)y * K.p.in_stride[1], (unsigned)w, bar);
        bulk_g2s(dst + 3u * w, K.p.in[2] + frame * K.p.in_step + (long)y * K.p.in_stride[2], (unsigned)w, bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((long)blockIdx.x < total_rows) issue_row(blockIdx.x, 0);
        if ((long)blockIdx.x + stride_rows < total_rows) issue_row(blockIdx.x + stride_rows, 1);
    }
    __syncthreads();

    const Rgb2YuvK b = K.b;
    const FastK fk = K.fk;
    const char *tab = reinterpret_cast<const char *>(K.lut.tab);
    const int n = K.lut.n, n2 = K.lut.n2;
    const float out_scale = K.lut.out_scale, lut_max = K.lut.lut_max;
    const unsigned kbase = 0u - kMagicBits * (unsigned)(n2 + n + 1);
    int it = 0;

    for (long grow = blockIdx.x; grow < total_rows; grow += stride_rows, ++it) {
        const int stage = it & 1;
        const unsigned phase = (it >> 1) & 1;
        const int frame = (int)(grow / K.h), y = (int)(grow - (long)frame * K.h);
        uint8_t *outY = K.p.out[0] + frame * K.p.out_step + (long)y * K.p.out_stride[0];
        uint8_t *outU = K.p.out[1] + frame * K.p.out_step + (long)y * K.p.out_stride[1];
        uint8_t *outV = K.p.out[2] + frame * K.p.out_step + (long)y * K.p.out_stride[2];
        unsigned *sc = s_c14 + stage * pitch + 4;  // sc[x] for x in [-4, w+4)
        const unsigned char *sY = s_in + stage * row_bytes, *sU = sY + 2 * w, *sV = sY + 3 * w;
        mbar_wait(bar0 + 8u * stage, phase);

        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g << 2;
            const uint2 yv = *reinterpret_cast<const uint2 *>(sY + 2 * x0);
            // chroma words: samples (2g-2, 2g-1), (2g, 2g+1), (2g+2, 2g+3); clamp at the row ends
            const unsigned uw = *reinterpret_cast<const unsigned *>(sU + x0), vw = *reinterpret_cast<const unsigned *>(sV + x0);
            unsigned ul, vl, ur, vr;
            if (g > 0) ul = *reinterpret_cast<const unsigned *>(sU + x0 - 4), vl = *reinterpret_cast<const unsigned *>(sV + x0 - 4);
            else ul = (uw & 0xFFFFu) * 0x10001u, vl = (vw & 0xFFFFu) * 0x10001u;
            if (g < groups - 1) ur = *reinterpret_cast<const unsigned *>(sU + x0 + 4), vr = *reinterpret_cast<const unsigned *>(sV + x0 + 4);
            else ur = (uw >> 16) * 0x10001u, vr = (vw >> 16) * 0x10001u;
            const int eu[6] = {(int)(ul & 0xFFFF), (int)(ul >> 16), (int)(uw & 0xFFFF), (int)(uw >> 16), (int)(ur & 0xFFFF), (int)(ur >> 16)};
            const int ev[6] = {(int)(vl & 0xFFFF), (int)(vl >> 16), (int)(vw & 0xFFFF), (int)(vw >> 16), (int)(vr & 0xFFFF), (int)(vr >> 16)};
            unsigned yo[2], pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned yw = k < 2 ? yv.x : yv.y;
                const unsigned Y = (k & 1) ? (yw >> 16) : (yw & 0xFFFF);
                const int j = k >> 1;
                int cu, cv;
                if (!(k & 1)) {  // even column 2i: taps on C[i-2 .. i+1]
                    cu = up4(eu[j], eu[j + 1], eu[j + 2], eu[j + 3], -461, 3943, 14284, -1382);
                    cv = up4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3], -461, 3943, 14284, -1382);
                    if (g == 0) {  // the two left-most even columns use swscale's edge-normalised taps
                        if (k == 0) {
                            cu = min((17729 * eu[2] - 1345 * eu[3]) >> 9, 32767);
                            cv = min((17729 * ev[2] - 1345 * ev[3]) >> 9, 32767);
                        } else {
                            cu = min((3835 * eu[2] + 13894 * eu[3] - 1345 * eu[4]) >> 9, 32767);
                            cv = min((3835 * ev[2] + 13894 * ev[3] - 1345 * ev[4]) >> 9, 32767);
                        }
                    }
                } else {  // odd column 2i+1: taps on C[i-1 .. i+2]
                    cu = up4(eu[j + 1], eu[j + 2], eu[j + 3], eu[j + 4], -1382, 14284, 3943, -461);
                    cv = up4(ev[j + 1], ev[j + 2], ev[j + 3], ev[j + 4], -1382, 14284, 3943, -461);
                }
                const unsigned yy = Y * fk.yA;
                const int R = (int)(yy + (unsigned)cv * fk.vr4 + fk.kR);
                const int G = (int)(yy + (unsigned)cv * fk.vg4 + (unsigned)cu * fk.ug4 + fk.kG);
                const int B = (int)(yy + (unsigned)cu * fk.ub4 + fk.kB);
                uint2 er, eg, eb;
                idx_arith<IDX>(clamp_u(R >> 20, 1023), fk, lut_max, er.x, er.y);
                idx_arith<IDX>(clamp_u(G >> 20, 1023), fk, lut_max, eg.x, eg.y);
                idx_arith<IDX>(clamp_u(B >> 20, 1023), fk, lut_max, eb.x, eb.y);
                unsigned rb, gb, bb;
                lut3d_fast<INTERP, FMA, CLAMP>(tab, n, n2, kbase, fk.n2b, fk.nb, fk.totalb, out_scale, er.x, eg.x, eb.x, er.y, eg.y, eb.y,
                                               rb, gb, bb);
                const unsigned yq = (unsigned)((int)((unsigned)b.ry * rb + (unsigned)b.gy * gb + (unsigned)b.by * bb + fk.kY) >> 15);
                const unsigned u14 = (unsigned)((int)((unsigned)b.ru * rb + (unsigned)b.gu * gb + (unsigned)b.bu * bb + fk.kU) >> 11);
                const unsigned v14 = (unsigned)((int)((unsigned)b.rv * rb + (unsigned)b.gv * gb + (unsigned)b.bv * bb + fk.kV) >> 11);
                pk[k] = (v14 << 16) + u14;
                if (k & 1) yo[k >> 1] |= yq << 16;
                else yo[k >> 1] = yq;
            }
            *reinterpret_cast<uint4 *>(sc + x0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (g == 0) *reinterpret_cast<uint4 *>(sc - 4) = make_uint4(pk[0], pk[0], pk[0], pk[0]);          // clamp left
            if (g == groups - 1) *reinterpret_cast<uint4 *>(sc + w) = make_uint4(pk[3], pk[3], pk[3], pk[3]);  // clamp right
            stg_stream8(outY + 2 * x0, make_uint2(yo[0], yo[1]));
        }
        __syncthreads();  // c14 of this row complete; nobody reads this input stage any more
        if (threadIdx.x == 0 && grow + 2 * stride_rows < total_rows) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order the generic-proxy reads before the async refill
            issue_row(grow + 2 * stride_rows, stage);
        }
        for (int g = threadIdx.x; g < groups; g += blockDim.x) {
            const int x0 = g << 2;
            unsigned t[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 v4 = *reinterpret_cast<const uint4 *>(sc + x0 - 4 + 4 * q);
                t[4 * q] = v4.x, t[4 * q + 1] = v4.y, t[4 * q + 2] = v4.z, t[4 * q + 3] = v4.w;
            }
            unsigned uo = 0, vo = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const unsigned s0 = t[2 * j + 1] + t[2 * j + 8], s1 = t[2 * j + 2] + t[2 * j + 7];
                const unsigned s2 = t[2 * j + 3] + t[2 * j + 6], s3 = t[2 * j + 4] + t[2 * j + 5];
                const int su = -230 * (int)(s0 & 0xFFFF) - 692 * (int)(s1 & 0xFFFF) + 1972 * (int)(s2 & 0xFFFF) + 7142 * (int)(s3 & 0xFFFF) + (16 << 13);
                const int sv = -230 * (int)(s0 >> 16) - 692 * (int)(s1 >> 16) + 1972 * (int)(s2 >> 16) + 7142 * (int)(s3 >> 16) + (16 << 13);
                uo |= (unsigned)clamp_u(su >> 18, 1023) << (16 * j);
                vo |= (unsigned)clamp_u(sv >> 18, 1023) << (16 * j);
            }
            stg_stream4(outU + x0, uo);
            stg_stream4(outV + x0, vo);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// yuv420p (8 bit) -> yuv420p   (BASELINE cfg 1 / 2)
// A CTA owns one band of luma rows of one frame and walks down it; a thread owns 8-pixel column groups.
// Nothing crosses threads: YUV->RGB uses replicated chroma, the chroma decimation is 2-px horizontal sum (inside the
// group) + 8-tap vertical filter over the same columns.  The per-row chroma terms live in a thread-private 8-row
// ring in shared memory.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pc2tv_luma8(int v) { return clamp_u(((((v << 7) * 14071 + 33561947) >> 14) + 64) >> 7, 255); }
__device__ __forceinline__ int pc2tv_chroma8(int v) { return clamp_u(((((v << 7) * 7197 + 16762880) >> 13) + 64) >> 7, 255); }

__device__ __forceinline__ unsigned map4(unsigned p, bool chroma) {
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = (p >> (8 * k)) & 0xFF;
        r |= (unsigned)(chroma ? pc2tv_chroma8(v) : pc2tv_luma8(v)) << (8 * k);
    }
    return r;
}

template <int INTERP, bool FMA>
__global__ void __launch_bounds__(512, 2) k_yuv420p8(const StageK K) {
    extern __shared__ uint4 s_ring[];  // [8][groups]: 4 x (u15 | v15 << 16) per row and group
    const int w = K.w, h = K.h, groups = w >> 3, ch = h >> 1;
    const int frame = blockIdx.x / K.bands, band = blockIdx.x - frame * K.bands;
    const int y_begin = band * K.rows_per_cta;              // first luma row whose outputs this CTA owns (even)
    const int y_end = min(h, y_begin + K.rows_per_cta);     // one past the last
    if (y_begin >= h) return;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    const uint8_t *inY = K.p.in[0] + frame * K.p.in_step, *inU = K.p.in[1] + frame * K.p.in_step, *inV = K.p.in[2] + frame * K.p.in_step;
    uint8_t *outY = K.p.out[0] + frame * K.p.out_step, *outU = K.p.out[1] + frame * K.p.out_step, *outV = K.p.out[2] + frame * K.p.out_step;
    const int taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};

    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int x0 = g << 3, i0 = g << 2;
        int next_row = max(0, y_begin - 3);  // next luma row to run through the pixel pipeline
        for (int j = y_begin >> 1; j < (y_end >> 1); ++j) {
            const int last_needed = min(h - 1, 2 * j + 4);
            for (; next_row <= last_needed; ++next_row) {
                const int y = next_row;
                uint2 yv = ldg_stream8(inY + (long)y * K.p.in_stride[0] + x0);
                unsigned uv = ldg_stream4(inU + (long)(y >> 1) * K.p.in_stride[1] + i0);
                unsigned vv = ldg_stream4(inV + (long)(y >> 1) * K.p.in_stride[2] + i0);
                if (K.range_pass) {
                    yv.x = map4(yv.x, false), yv.y = map4(yv.y, false);
                    uv = map4(uv, true), vv = map4(vv, true);
                }
                unsigned yo[2] = {0u, 0u};
                unsigned ring[4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {  // pixel pairs sharing one chroma sample
                    const int u = (int)((uv >> (8 * pr)) & 0xFF) * 8 - 1024, v = (int)((vv >> (8 * pr)) & 0xFF) * 8 - 1024;
                    const int cr = (v * f.vr) >> 16, cg = ((u * f.ug) >> 16) + ((v * f.vg) >> 16), cb = (u * f.ub) >> 16;
                    int rs = 0, gs = 0, bs = 0;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int k = 2 * pr + q;
                        const int Y = (int)(((k < 4 ? yv.x : yv.y) >> (8 * (k & 3))) & 0xFF);
                        const int yy = ((Y * 8 - f.yO8) * f.yC) >> 16;
                        const int r8 = clamp_u(yy + cr, 255), g8 = clamp_u(yy + cg, 255), b8 = clamp_u(yy + cb, 255);
                        int ro, go, bo;
                        lut3d_px<INTERP, FMA>(K.lut, r8, g8, b8, 255, ro, go, bo);
                        // rgb24ToY + unit hscale + yuv2plane1_8
                        const int y14 = (b.ry * ro + b.gy * go + b.by * bo + (32 << 14) + (1 << 8)) >> 9;
                        const unsigned yq = (unsigned)clamp_u((2 * y14 + 64) >> 7, 255);
                        yo[k >> 2] |= yq << (8 * (k & 3));
                        rs += ro, gs += go, bs += bo;
                    }
                    // rgb24ToUV_half: sum of the pair, then 15-bit intermediate
                    const int u14 = (b.ru * rs + b.gu * gs + b.bu * bs + (256 << 15) + (1 << 9)) >> 10;
                    const int v14 = (b.rv * rs + b.gv * gs + b.bv * bs + (256 << 15) + (1 << 9)) >> 10;
                    ring[pr] = (unsigned)((2 * u14) & 0xFFFF) | ((unsigned)(2 * v14) << 16);
                }
                s_ring[(y & 7) * groups + g] = make_uint4(ring[0], ring[1], ring[2], ring[3]);
                if (y >= y_begin && y < y_end) stg_stream8(outY + (long)y * K.p.out_stride[0] + x0, make_uint2(yo[0], yo[1]));
            }
            // vertical 8-tap decimation for chroma row j over luma rows 2j-3 .. 2j+4 (clamped; taps that land on the
            // same edge row are summed into one coefficient before the multiply)
            int au[4], av[4];
            const bool exact = (j == ch - 1);  // swscale's C vertical scaler handles the last output row: no per-tap truncation
#pragma unroll
            for (int c = 0; c < 4; ++c) au[c] = av[c] = exact ? (64 << 12) : 7;
            int fold = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = max(0, min(h - 1, 2 * j - 3 + k));
                const int nrow = max(0, min(h - 1, 2 * j - 2 + k));
                fold += taps[k];
                if (k == 7 || nrow != row) {
                    const uint4 q = s_ring[(row & 7) * groups + g];
                    const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cu = (int)(qq[c] & 0xFFFF), cv = (int)(qq[c] >> 16);
                        if (exact) au[c] += cu * fold, av[c] += cv * fold;
                        else au[c] += (cu * fold) >> 16, av[c] += (cv * fold) >> 16;
                    }
                    fold = 0;
                }
            }
            unsigned uo = 0, vo = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uo |= (unsigned)clamp_u(au[c] >> (exact ? 19 : 3), 255) << (8 * c);
                vo |= (unsigned)clamp_u(av[c] >> (exact ? 19 : 3), 255) << (8 * c);
            }
            stg_stream4(outU + (long)j * K.p.out_stride[1] + i0, uo);
            stg_stream4(outV + (long)j * K.p.out_stride[2] + i0, vo);
        }
    }
}

// yuv420p (8 bit) -> yuv420p, fast path: same band / thread-private ring structure as k_yuv420p8, with the exact
// conversion-free LUT prologue / epilogue of the 4:2:2 fast kernel and the next row's samples prefetched into registers
// while the current row is computed (the only long-latency loads left are the LUT gathers).
template <int INTERP, bool FMA, bool CLAMP, int IDX>
__global__ void __launch_bounds__(256, 3) k_yuv420p8_fast(const StageK K) {
    extern __shared__ uint4 s_ring[];  // [8][groups]: 4 x (u15 | v15 << 16) per row and group
    const int w = K.w, h = K.h, groups = w >> 3, ch = h >> 1;
    const int frame = blockIdx.x / K.bands, band = blockIdx.x - frame * K.bands;
    const int y_begin = band * K.rows_per_cta;
    const int y_end = min(h, y_begin + K.rows_per_cta);
    if (y_begin >= h) return;
    const Yuv2RgbK f = K.f;
    const Rgb2YuvK b = K.b;
    const FastK fk = K.fk;
    const char *tab = reinterpret_cast<const char *>(K.lut.tab);
    const int n = K.lut.n, n2 = K.lut.n2;
    const float out_scale = K.lut.out_scale, lut_max = K.lut.lut_max;
    const unsigned kbase = 0u - kMagicBits * (unsigned)(n2 + n + 1);
    const uint8_t *inY = K.p.in[0] + frame * K.p.in_step, *inU = K.p.in[1] + frame * K.p.in_step, *inV = K.p.in[2] + frame * K.p.in_step;
    uint8_t *outY = K.p.out[0] + frame * K.p.out_step, *outU = K.p.out[1] + frame * K.p.out_step, *outV = K.p.out[2] + frame * K.p.out_step;
    const int taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
    const int r_first = max(0, y_begin - 3), r_last = min(h - 1, y_end + 2);

    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int x0 = g << 3, i0 = g << 2;
        int y = r_first;
        uint2 yv = ldg_stream8(inY + (long)y * K.p.in_stride[0] + x0);
        unsigned uv = ldg_stream4(inU + (long)(y >> 1) * K.p.in_stride[1] + i0);
        unsigned vv = ldg_stream4(inV + (long)(y >> 1) * K.p.in_stride[2] + i0);
        for (int j = y_begin >> 1; j < (y_end >> 1); ++j) {
            const int last_needed = min(h - 1, 2 * j + 4);
            for (; y <= last_needed; ++y) {
                // prefetch the next row of this column group
                const int yn = min(y + 1, r_last);
                const uint2 yv_n = ldg_stream8(inY + (long)yn * K.p.in_stride[0] + x0);
                const unsigned uv_n = ldg_stream4(inU + (long)(yn >> 1) * K.p.in_stride[1] + i0);
                const unsigned vv_n = ldg_stream4(inV + (long)(yn >> 1) * K.p.in_stride[2] + i0);
                if (K.range_pass) {
                    yv.x = map4(yv.x, false), yv.y = map4(yv.y, false);
                    uv = map4(uv, true), vv = map4(vv, true);
                }
                unsigned yo[2] = {0u, 0u};
                unsigned ring[4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {  // pixel pairs sharing one chroma sample
                    const int u = (int)((uv >> (8 * pr)) & 0xFF) * 8 - 1024, v = (int)((vv >> (8 * pr)) & 0xFF) * 8 - 1024;
                    const int cr = (v * f.vr) >> 16, cg = ((u * f.ug) >> 16) + ((v * f.vg) >> 16), cb = (u * f.ub) >> 16;
                    unsigned rs = 0, gs = 0, bs = 0;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int k = 2 * pr + q;
                        const int Y = (int)(((k < 4 ? yv.x : yv.y) >> (8 * (k & 3))) & 0xFF);
                        const int yy = ((Y * 8 - f.yO8) * f.yC) >> 16;
                        uint2 er, eg, eb;
                        idx_arith<IDX>(clamp_u(yy + cr, 255), fk, lut_max, er.x, er.y);
                        idx_arith<IDX>(clamp_u(yy + cg, 255), fk, lut_max, eg.x, eg.y);
                        idx_arith<IDX>(clamp_u(yy + cb, 255), fk, lut_max, eb.x, eb.y);
                        unsigned rb, gb, bb;
                        lut3d_fast<INTERP, FMA, CLAMP>(tab, n, n2, kbase, fk.n2b, fk.nb, fk.totalb, out_scale, er.x, eg.x, eb.x, er.y, eg.y,
                                                       eb.y, rb, gb, bb);
                        // clip8((2*y14 + 64) >> 7), y14 = S >> 9  ==  (S + (32 << 9)) >> 15; no clip can bind (host-checked)
                        const unsigned yq = (unsigned)((int)((unsigned)b.ry * rb + (unsigned)b.gy * gb + (unsigned)b.by * bb + fk.kY) >> 15);
                        yo[k >> 2] |= yq << (8 * (k & 3));
                        rs += rb, gs += gb, bs += bb;
                    }
                    // rgb24ToUV_half on the pair sums (each sum carries 2 x the 0x4B000000 bias, folded into kU / kV)
                    const unsigned u14 = (unsigned)((int)((unsigned)b.ru * rs + (unsigned)b.gu * gs + (unsigned)b.bu * bs + fk.kU) >> 10);
                    const unsigned v14 = (unsigned)((int)((unsigned)b.rv * rs + (unsigned)b.gv * gs + (unsigned)b.bv * bs + fk.kV) >> 10);
                    ring[pr] = ((2u * v14) << 16) + 2u * u14;
                }
                s_ring[(y & 7) * groups + g] = make_uint4(ring[0], ring[1], ring[2], ring[3]);
                if (y >= y_begin && y < y_end) stg_stream8(outY + (long)y * K.p.out_stride[0] + x0, make_uint2(yo[0], yo[1]));
                yv = yv_n, uv = uv_n, vv = vv_n;
            }
            int au[4], av[4];
            const bool exact = (j == ch - 1);
#pragma unroll
            for (int c = 0; c < 4; ++c) au[c] = av[c] = exact ? (64 << 12) : 7;
            if (2 * j - 3 >= 0 && 2 * j + 4 <= h - 1) {
                // interior: 8 distinct rows, per-tap truncation (the SIMD vertical scaler)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint4 q = s_ring[((2 * j - 3 + k) & 7) * groups + g];
                    const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        au[c] += ((int)(qq[c] & 0xFFFF) * taps[k]) >> 16;
                        av[c] += ((int)(qq[c] >> 16) * taps[k]) >> 16;
                    }
                }
            } else {
                int fold = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int row = max(0, min(h - 1, 2 * j - 3 + k));
                    const int nrow = max(0, min(h - 1, 2 * j - 2 + k));
                    fold += taps[k];
                    if (k == 7 || nrow != row) {
                        const uint4 q = s_ring[(row & 7) * groups + g];
                        const unsigned qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int cu = (int)(qq[c] & 0xFFFF), cv = (int)(qq[c] >> 16);
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
                uo |= (unsigned)clamp_u(au[c] >> (exact ? 19 : 3), 255) << (8 * c);
                vo |= (unsigned)clamp_u(av[c] >> (exact ? 19 : 3), 255) << (8 * c);
            }
            stg_stream4(outU + (long)j * K.p.out_stride[1] + i0, uo);
            stg_stream4(outV + (long)j * K.p.out_stride[2] + i0, vo);
        }
    }
}

template <int INTERP>
int launch422(const StageK &k, int float_mode, cudaStream_t st) {
    const int groups = k.w >> 3;
    int per = (groups + 511) / 512;                   // groups per thread
    int threads = ((groups + per - 1) / per + 31) / 32 * 32;
    const long total_rows = (long)k.h * k.nframes;
    if (k.idx_mode >= 0 && INTERP != LUTR_INTERP_NEAREST) {
        const int idx = k.idx_mode == 0 ? 0 : (k.fk.idx_fold ? 1 : 2);
        size_t smem = (idx == 0 ? kIdxEntries10 * sizeof(uint2) : 0) + (size_t)2 * (k.w + 8) * sizeof(unsigned);
        void (*fn)(const StageK) = nullptr;
        const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
        int mb = k.min_blocks == 4 ? 4 : 3;
        int fast_threads = 256;
        if (k.idx_mode == 2) {  // TMA-staged input rows
            if (k.min_blocks == 2) mb = 2;
            // threads per CTA: as close to the cap as divides the row's 4-pixel groups evenly (balanced phases, one barrier per row)
            const int cap = mb == 2 ? 320 : 256, groups4 = k.w >> 2;
            const int per = (groups4 + cap - 1) / cap;
            fast_threads = k.fast_threads > 0 ? k.fast_threads : ((groups4 + per - 1) / per + 31) / 32 * 32;
            if (fast_threads > cap) fast_threads = cap;
            smem = (size_t)8 * k.w + (size_t)2 * (k.w + 8) * sizeof(unsigned) + 16;
#define LUTR_PICK_T(F, C, I, M) \
    if (fma == F && clamp == C && idx == I && mb == M) fn = k_yuv422p10_tma<INTERP, F, C, I, M>;
#define LUTR_PICK_TM(F, C, I) LUTR_PICK_T(F, C, I, 2) LUTR_PICK_T(F, C, I, 3) LUTR_PICK_T(F, C, I, 4)
#define LUTR_PICK_TI(F, C) LUTR_PICK_TM(F, C, 1) LUTR_PICK_TM(F, C, 2)
            LUTR_PICK_TI(false, false) LUTR_PICK_TI(false, true) LUTR_PICK_TI(true, false) LUTR_PICK_TI(true, true)
#undef LUTR_PICK_TI
#undef LUTR_PICK_TM
#undef LUTR_PICK_T
        }
        if (!fn) {
#define LUTR_PICK(F, C, I, M) \
    if (fma == F && clamp == C && idx == I && mb == M) fn = k_yuv422p10_fast<INTERP, F, C, I, M>;
#define LUTR_PICK_M(F, C, I) LUTR_PICK(F, C, I, 3) LUTR_PICK(F, C, I, 4)
#define LUTR_PICK_I(F, C) LUTR_PICK_M(F, C, 0) LUTR_PICK_M(F, C, 1) LUTR_PICK_M(F, C, 2)
        LUTR_PICK_I(false, false) LUTR_PICK_I(false, true) LUTR_PICK_I(true, false) LUTR_PICK_I(true, true)
#undef LUTR_PICK_I
#undef LUTR_PICK_M
#undef LUTR_PICK
        }
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        long grid = k.grid_ctas > 0 ? k.grid_ctas : 148 * mb;
        if (grid > total_rows) grid = total_rows;
        fn<<<(unsigned)grid, fast_threads, smem, st>>>(k);
        return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
    }
    const unsigned grid = (unsigned)((total_rows + k.rows_per_cta - 1) / k.rows_per_cta);
    const size_t smem = (size_t)2 * k.w * sizeof(unsigned);
    auto fn = float_mode == LUTR_FLOAT_FMA ? k_yuv422p10<INTERP, true> : k_yuv422p10<INTERP, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fn<<<grid, threads, smem, st>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

template <int INTERP>
int launch420(const StageK &k, int float_mode, cudaStream_t st) {
    const int groups = k.w >> 3;
    const unsigned grid = (unsigned)(k.bands * k.nframes);
    const size_t smem = (size_t)8 * groups * sizeof(uint4);
    if (k.idx_mode >= 0 && INTERP != LUTR_INTERP_NEAREST) {
        const int idx = k.fk.idx_fold ? 1 : 2;
        const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
        void (*fn)(const StageK) = nullptr;
#define LUTR_PICK8(F, C, I) \
    if (fma == F && clamp == C && idx == I) fn = k_yuv420p8_fast<INTERP, F, C, I>;
#define LUTR_PICK8_I(F, C) LUTR_PICK8(F, C, 1) LUTR_PICK8(F, C, 2)
        LUTR_PICK8_I(false, false) LUTR_PICK8_I(false, true) LUTR_PICK8_I(true, false) LUTR_PICK8_I(true, true)
#undef LUTR_PICK8_I
#undef LUTR_PICK8
        const int threads = groups >= 256 ? 256 : (groups + 31) / 32 * 32;
        if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fn<<<grid, threads, smem, st>>>(k);
        return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
    }
    int per = (groups + 511) / 512;
    int threads = ((groups + per - 1) / per + 31) / 32 * 32;
    auto fn = float_mode == LUTR_FLOAT_FMA ? k_yuv420p8<INTERP, true> : k_yuv420p8<INTERP, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fn<<<grid, threads, smem, st>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace

int launch_yuv422p10(const StageK &k, int interp, int float_mode, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (interp) {
        case LUTR_INTERP_NEAREST: return launch422<LUTR_INTERP_NEAREST>(k, float_mode, st);
        case LUTR_INTERP_TRILINEAR: return launch422<LUTR_INTERP_TRILINEAR>(k, float_mode, st);
        default: return launch422<LUTR_INTERP_TETRAHEDRAL>(k, float_mode, st);
    }
}

int launch_yuv420p8(const StageK &k, int interp, int float_mode, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (interp) {
        case LUTR_INTERP_NEAREST: return launch420<LUTR_INTERP_NEAREST>(k, float_mode, st);
        case LUTR_INTERP_TRILINEAR: return launch420<LUTR_INTERP_TRILINEAR>(k, float_mode, st);
        default: return launch420<LUTR_INTERP_TETRAHEDRAL>(k, float_mode, st);
    }
}

}  // namespace lutr
