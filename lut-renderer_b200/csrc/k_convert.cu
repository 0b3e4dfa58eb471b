// LUT-less pixel-format conversion of the pro-mode master: yuv422p10le -> yuv420p | yuv420p10le | yuv422p.
//
// Reference: pro mode stage 2 (src/lut_renderer/ffmpeg.py:459-471 builds a command without lut3d; :287-310 picks the delivery pix_fmt):
// the encoding ffmpeg converts the ProRes 422 HQ intermediate with an auto-inserted swscale pass, YUV -> YUV, same matrix and range
// on both sides.  The arithmetic is the output stage of the colour kernels fed with the 15-bit samples v << 5 (restated in
// oracle/lutr_oracle.c orc_convert_yuv422p10, pinned against libswscale 9.1.100):
//   -> yuv420p      luma clip8(((v << 5) + d) >> 7), d = ff_dither_8x8_128[y & 7][x & 7]; chroma 8-tap vertical decimation over luma
//                   rows 2j-3 .. 2j+4 (edge rows folded into one coefficient), per-tap truncation (x86 yuv2yuvX), the last chroma
//                   row exact (C code), dither column offset 3 for V
//   -> yuv420p10le  luma unchanged, chroma exact 8-tap vertical: (sum + (1 << 16)) >> 17
//   -> yuv422p      swscale's planar copy: every plane min((v + d2) >> 2, 255), d2 = {{1,2},{3,0}}[y & 1][x & 1]
// A streaming kernel (4 B/px in, 1.5 - 3 B/px out, no gathers): one thread owns 8 luma columns (4 chroma columns) of a strip of rows,
// 16-byte luma and 8-byte chroma accesses; the 8 chroma source rows of an output row are re-read from L1 / L2.
#include <cuda_runtime.h>

#include "lutr_internal.h"

namespace lutr {

namespace {

__constant__ __align__(8) unsigned char kDither128[8][8] = {
    {36, 68, 60, 92, 34, 66, 58, 90},     {100, 4, 124, 28, 98, 2, 122, 26}, {52, 84, 44, 76, 50, 82, 42, 74},
    {116, 20, 108, 12, 114, 18, 106, 10}, {32, 64, 56, 88, 38, 70, 62, 94},  {96, 0, 120, 24, 102, 6, 126, 30},
    {48, 80, 40, 72, 54, 86, 46, 78},     {112, 16, 104, 8, 118, 22, 110, 14},
};

constexpr int kStrip = 8;  // output chroma rows (4:2:0 outputs) or plane rows (4:2:2 output) per thread

__device__ __forceinline__ int clip_hi(int v, int hi) { return min(max(v, 0), hi); }

// MODE 0: -> yuv420p, 1: -> yuv420p10le, 2: -> yuv422p
template <int MODE>
__global__ void __launch_bounds__(128) k_convert_422p10(const StageK K) {
    const int w = K.w, h = K.h;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 luma / 4 chroma columns
    if (8 * g >= w) return;
    const int frame = blockIdx.z;
    const uint8_t *inY = K.p.in[0] + (long)frame * K.p.in_step + 16l * g;
    const uint8_t *inU = K.p.in[1] + (long)frame * K.p.in_step + 8l * g, *inV = K.p.in[2] + (long)frame * K.p.in_step + 8l * g;
    constexpr int obps = MODE == 1 ? 2 : 1;
    uint8_t *outY = K.p.out[0] + (long)frame * K.p.out_step + (long)(8 * obps) * g;
    uint8_t *outU = K.p.out[1] + (long)frame * K.p.out_step + (long)(4 * obps) * g, *outV = K.p.out[2] + (long)frame * K.p.out_step + (long)(4 * obps) * g;
    auto ld16 = [](const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); };
    auto ld8 = [](const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); };
    auto sample = [](const unsigned wd[], int k) { return (int)((wd[k >> 1] >> (16 * (k & 1))) & 0xFFFFu); };
    // one row of the dither matrix with ONE uniform 8-byte constant load; a per-thread column index into constant memory would be
    // served address by address.  drow_u / drow_v: the 4 bytes for this thread's chroma columns 4g .. 4g+3 (V shifted by 3 columns)
    auto drow = [](int y) { return *reinterpret_cast<const uint2 *>(&kDither128[y & 7][0]); };
    auto dbyte = [](uint2 row, int k) { return (int)(((k < 4 ? row.x : row.y) >> (8 * (k & 3))) & 0xFFu); };
    auto drow_u = [&](uint2 row) { return (g & 1) ? row.y : row.x; };
    auto drow_v = [&](uint2 row) { return __byte_perm(row.x, row.y, (g & 1) ? 0x2107u : 0x6543u); };

    if (MODE == 2) {  // planar copy: rows of this strip
        const int y0 = blockIdx.y * kStrip;
        for (int y = y0; y < min(y0 + kStrip, h); ++y) {
            const uint4 yv = ld16(inY + (long)y * K.p.in_stride[0]);
            const uint2 uv = ld8(inU + (long)y * K.p.in_stride[1]), vv = ld8(inV + (long)y * K.p.in_stride[2]);
            const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w}, uw[2] = {uv.x, uv.y}, vw[2] = {vv.x, vv.y};
            const int d0 = (y & 1) ? 3 : 1, d1 = (y & 1) ? 0 : 2;  // d2[y & 1][x & 1]; every group starts on an even column
            unsigned yo[2] = {0u, 0u}, uo = 0u, vo = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) yo[k >> 2] |= (unsigned)min((sample(yw, k) + ((k & 1) ? d1 : d0)) >> 2, 255) << (8 * (k & 3));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uo |= (unsigned)min((sample(uw, k) + ((k & 1) ? d1 : d0)) >> 2, 255) << (8 * k);
                vo |= (unsigned)min((sample(vw, k) + ((k & 1) ? d1 : d0)) >> 2, 255) << (8 * k);
            }
            *reinterpret_cast<uint2 *>(outY + (long)y * K.p.out_stride[0]) = make_uint2(yo[0], yo[1]);
            *reinterpret_cast<unsigned *>(outU + (long)y * K.p.out_stride[1]) = uo;
            *reinterpret_cast<unsigned *>(outV + (long)y * K.p.out_stride[2]) = vo;
        }
        return;
    }

    const int ch = h >> 1, j0 = blockIdx.y * kStrip;
    constexpr int taps[8] = {-58, -172, 492, 1786, 1786, 492, -172, -58};
    // luma rows 2j, 2j + 1 of this thread's 8 columns
    auto luma_rows = [&](int j) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int y = 2 * j + r;
            const uint4 yv = ld16(inY + (long)y * K.p.in_stride[0]);
            const unsigned yw[4] = {yv.x, yv.y, yv.z, yv.w};
            if (MODE == 1) {  // ((v << 5) + 16) >> 5 == v for 10-bit samples; clip of the (invalid) codes above 1023 like the scaler
                unsigned o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = (unsigned)min(sample(yw, 2 * q), 1023) | ((unsigned)min(sample(yw, 2 * q + 1), 1023) << 16);
                *reinterpret_cast<uint4 *>(outY + (long)y * K.p.out_stride[0]) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                unsigned yo[2] = {0u, 0u};
                const uint2 dr = drow(y);
#pragma unroll
                for (int k = 0; k < 8; ++k) yo[k >> 2] |= (unsigned)min(((sample(yw, k) << 5) + dbyte(dr, k)) >> 7, 255) << (8 * (k & 3));
                *reinterpret_cast<uint2 *>(outY + (long)y * K.p.out_stride[0]) = make_uint2(yo[0], yo[1]);
            }
        }
    };
    // Interior strips (all 8 source rows of every chroma row inside the frame, no last chroma row): a sliding window of the 8 source
    // rows in registers -- every chroma source row is loaded once per strip (22 row loads for 8 output rows) instead of 8 times.
    if (j0 + kStrip <= ch - 1 && 2 * j0 - 3 >= 0 && 2 * (j0 + kStrip - 1) + 4 <= h - 1) {
        uint2 wu[8], wvv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) wu[t] = ld8(inU + (long)(2 * j0 - 3 + t) * K.p.in_stride[1]), wvv[t] = ld8(inV + (long)(2 * j0 - 3 + t) * K.p.in_stride[2]);
#pragma unroll
        for (int jj = 0; jj < kStrip; ++jj) {
            const int j = j0 + jj;
            uint2 nu0 = wu[0], nu1 = wu[0], nv0 = wvv[0], nv1 = wvv[0];
            if (jj + 1 < kStrip) {  // the two rows the next output row adds, in flight while this one is computed
                nu0 = ld8(inU + (long)(2 * j + 5) * K.p.in_stride[1]), nu1 = ld8(inU + (long)(2 * j + 6) * K.p.in_stride[1]);
                nv0 = ld8(inV + (long)(2 * j + 5) * K.p.in_stride[2]), nv1 = ld8(inV + (long)(2 * j + 6) * K.p.in_stride[2]);
            }
            luma_rows(j);
            const uint2 dj = drow(j);
            const unsigned du4 = drow_u(dj), dv4 = drow_v(dj);
            int su[4], sv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (MODE == 1) su[k] = sv[k] = 1 << 16;
                else su[k] = (int)(((du4 >> (8 * k)) & 0xFFu) + 56) >> 4, sv[k] = (int)(((dv4 >> (8 * k)) & 0xFFu) + 56) >> 4;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const unsigned uw[2] = {wu[t].x, wu[t].y}, vw[2] = {wvv[t].x, wvv[t].y};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pu_ = (sample(uw, k) << 5) * taps[t], pv_ = (sample(vw, k) << 5) * taps[t];
                    if (MODE == 0) su[k] += pu_ >> 16, sv[k] += pv_ >> 16;
                    else su[k] += pu_, sv[k] += pv_;
                }
            }
            if (MODE == 1) {
                const unsigned uo0 = (unsigned)clip_hi(su[0] >> 17, 1023) | ((unsigned)clip_hi(su[1] >> 17, 1023) << 16);
                const unsigned uo1 = (unsigned)clip_hi(su[2] >> 17, 1023) | ((unsigned)clip_hi(su[3] >> 17, 1023) << 16);
                const unsigned vo0 = (unsigned)clip_hi(sv[0] >> 17, 1023) | ((unsigned)clip_hi(sv[1] >> 17, 1023) << 16);
                const unsigned vo1 = (unsigned)clip_hi(sv[2] >> 17, 1023) | ((unsigned)clip_hi(sv[3] >> 17, 1023) << 16);
                *reinterpret_cast<uint2 *>(outU + (long)j * K.p.out_stride[1]) = make_uint2(uo0, uo1);
                *reinterpret_cast<uint2 *>(outV + (long)j * K.p.out_stride[2]) = make_uint2(vo0, vo1);
            } else {
                unsigned uo = 0u, vo = 0u;
#pragma unroll
                for (int k = 0; k < 4; ++k) uo |= (unsigned)clip_hi(su[k] >> 3, 255) << (8 * k), vo |= (unsigned)clip_hi(sv[k] >> 3, 255) << (8 * k);
                *reinterpret_cast<unsigned *>(outU + (long)j * K.p.out_stride[1]) = uo;
                *reinterpret_cast<unsigned *>(outV + (long)j * K.p.out_stride[2]) = vo;
            }
#pragma unroll
            for (int t = 0; t < 6; ++t) wu[t] = wu[t + 2], wvv[t] = wvv[t + 2];
            wu[6] = nu0, wu[7] = nu1, wvv[6] = nv0, wvv[7] = nv1;
        }
        return;
    }
    for (int j = j0; j < min(j0 + kStrip, ch); ++j) {
        luma_rows(j);
        // ---- chroma row j: luma rows 2j-3 .. 2j+4, taps of rows outside the frame folded onto the edge row ------------------------
        int su[4], sv[4];
        const bool last = MODE == 0 && j == ch - 1;  // swscale's C vertical scaler on the last output row: exact accumulation
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (MODE == 1) su[k] = sv[k] = 1 << 16;
            else {
                const int du = (int)((drow_u(drow(j)) >> (8 * k)) & 0xFFu), dv = (int)((drow_v(drow(j)) >> (8 * k)) & 0xFFu);
                su[k] = last ? du << 12 : (du + 56) >> 4, sv[k] = last ? dv << 12 : (dv + 56) >> 4;
            }
        }
        int prev_row = -1, coef = 0;
        uint2 pu = make_uint2(0u, 0u), pv = make_uint2(0u, 0u);
        auto flush = [&]() {  // accumulate the pending (row, folded coefficient)
            const unsigned uw[2] = {pu.x, pu.y}, vw[2] = {pv.x, pv.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int pu_ = (sample(uw, k) << 5) * coef, pv_ = (sample(vw, k) << 5) * coef;
                if (MODE == 0 && !last) su[k] += pu_ >> 16, sv[k] += pv_ >> 16;
                else su[k] += pu_, sv[k] += pv_;
            }
        };
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int row = min(max(2 * j - 3 + t, 0), h - 1);
            if (row == prev_row) {
                coef += taps[t];
                continue;
            }
            if (prev_row >= 0) flush();
            prev_row = row, coef = taps[t];
            pu = ld8(inU + (long)row * K.p.in_stride[1]), pv = ld8(inV + (long)row * K.p.in_stride[2]);
        }
        flush();
        if (MODE == 1) {
            const unsigned uo0 = (unsigned)clip_hi(su[0] >> 17, 1023) | ((unsigned)clip_hi(su[1] >> 17, 1023) << 16);
            const unsigned uo1 = (unsigned)clip_hi(su[2] >> 17, 1023) | ((unsigned)clip_hi(su[3] >> 17, 1023) << 16);
            const unsigned vo0 = (unsigned)clip_hi(sv[0] >> 17, 1023) | ((unsigned)clip_hi(sv[1] >> 17, 1023) << 16);
            const unsigned vo1 = (unsigned)clip_hi(sv[2] >> 17, 1023) | ((unsigned)clip_hi(sv[3] >> 17, 1023) << 16);
            *reinterpret_cast<uint2 *>(outU + (long)j * K.p.out_stride[1]) = make_uint2(uo0, uo1);
            *reinterpret_cast<uint2 *>(outV + (long)j * K.p.out_stride[2]) = make_uint2(vo0, vo1);
        } else {
            unsigned uo = 0u, vo = 0u;
            const int sh = last ? 19 : 3;
#pragma unroll
            for (int k = 0; k < 4; ++k) uo |= (unsigned)clip_hi(su[k] >> sh, 255) << (8 * k), vo |= (unsigned)clip_hi(sv[k] >> sh, 255) << (8 * k);
            *reinterpret_cast<unsigned *>(outU + (long)j * K.p.out_stride[1]) = uo;
            *reinterpret_cast<unsigned *>(outV + (long)j * K.p.out_stride[2]) = vo;
        }
    }
}

}  // namespace

int launch_k_convert(const StageK &k, int out_fmt, void *stream) {
    const int groups = k.w / 8;
    const int rows = out_fmt == LUTR_PIX_YUV422P ? k.h : k.h / 2;
    dim3 grid((unsigned)((groups + 127) / 128), (unsigned)((rows + kStrip - 1) / kStrip), (unsigned)k.nframes), block(128);
    if (out_fmt == LUTR_PIX_YUV420P) k_convert_422p10<0><<<grid, block, 0, (cudaStream_t)stream>>>(k);
    else if (out_fmt == LUTR_PIX_YUV420P10LE) k_convert_422p10<1><<<grid, block, 0, (cudaStream_t)stream>>>(k);
    else if (out_fmt == LUTR_PIX_YUV422P) k_convert_422p10<2><<<grid, block, 0, (cudaStream_t)stream>>>(k);
    else return LUTR_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace lutr
