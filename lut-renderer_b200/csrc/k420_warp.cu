// Instantiations and launcher of the 4:2:0 warp-window kernel (k420_warp.cuh).  LUTR_K420_OUT selects the output side this
// translation unit builds: 0 yuv420p (k420_warp.cu), 1 yuv422p10le (k420_warp10.cu), 2 yuv422p, 3 yuv444p (k420_warp422.cu,
// k420_warp444.cu) -- separate files so that they compile in parallel.
#include "k420_warp.cuh"

#ifndef LUTR_K420_OUT
#define LUTR_K420_OUT 0
#endif

namespace lutr {

namespace {

constexpr int kOutSide = LUTR_K420_OUT;
constexpr bool kRowLocal = kOutSide != 0;  // no vertical chroma filter on the output side

using KernelFn = void (*)(const StageK);

template <int INTERP, bool FMA, bool CLAMP, int IDX, bool NOMIN, int NSZ, int NK, int MINB>
KernelFn pick_range420(int rmode, bool in422) {  // rmode: 0 tv throughout, 1 pc -> tv pass in front, 2 full-range output (8-bit outputs)
    if constexpr (kOutSide != 1) {
        if (rmode == 2)
            return in422 ? k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 2, kOutSide, true, MINB>
                         : k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 2, kOutSide, false, MINB>;
    }
    if (in422)
        return rmode == 1 ? k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 1, kOutSide, true, MINB>
                          : k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 0, kOutSide, true, MINB>;
    return rmode == 1 ? k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 1, kOutSide, false, MINB>
                      : k420_warp<INTERP, FMA, CLAMP, IDX, NOMIN, NSZ, NK, 0, kOutSide, false, MINB>;
}

// Window widths of 128 / 192 / 256 columns (NK = 2 / 3 / 4) and the 33-point specialisation exist for the default semantics
// with the folded index only (C float order, LUT inside [0,1], size-1 a power of two); everything else runs NK = 4, generic size.
template <int INTERP, int MINB>
KernelFn pick420(bool fma, bool clamp, bool fold_nomin, int range, bool in422, int n, int nk) {
    if (!fma && !clamp && fold_nomin) {
        if constexpr (kOutSide == 0) {
            if (n == 33) {
                if (nk == 2) return pick_range420<INTERP, false, false, 1, true, 33, 2, MINB>(range, in422);
                if (nk == 3) return pick_range420<INTERP, false, false, 1, true, 33, 3, MINB>(range, in422);
                return pick_range420<INTERP, false, false, 1, true, 33, 4, MINB>(range, in422);
            }
        }
        if (nk == 2) return pick_range420<INTERP, false, false, 1, true, 0, 2, MINB>(range, in422);
        if (nk == 3) return pick_range420<INTERP, false, false, 1, true, 0, 3, MINB>(range, in422);
        return pick_range420<INTERP, false, false, 1, true, 0, 4, MINB>(range, in422);
    }
    if (!fma && !clamp) return pick_range420<INTERP, false, false, 2, false, 0, 4, MINB>(range, in422);
    if (fold_nomin) {
        if (!fma) return pick_range420<INTERP, false, true, 1, true, 0, 4, MINB>(range, in422);
        return clamp ? pick_range420<INTERP, true, true, 1, true, 0, 4, MINB>(range, in422) : pick_range420<INTERP, true, false, 1, true, 0, 4, MINB>(range, in422);
    }
    if (!fma) return pick_range420<INTERP, false, true, 2, false, 0, 4, MINB>(range, in422);
    return clamp ? pick_range420<INTERP, true, true, 2, false, 0, 4, MINB>(range, in422) : pick_range420<INTERP, true, false, 2, false, 0, 4, MINB>(range, in422);
}

}  // namespace

// Window width and grid: every warp of the persistent grid gets an equal, contiguous share of the chroma rows of all
// (frame, window) columns and recomputes 6 halo luma rows per share.  Wide windows amortise the per-row overhead, narrow ones
// may tile the row without a (recomputing) shifted last window and give longer shares: estimate each.  Small jobs use a smaller
// grid so that a share is at least 16 chroma rows.  k.wk_unroll (2, 3, 4) forces the window width; k.grid_ctas > 0 the grid.
#if LUTR_K420_OUT == 1
int launch_k420_warp10(const StageK &k_in, bool in422, int interp, int float_mode, bool grid_forced, void *stream) {
#elif LUTR_K420_OUT == 2
int launch_k420_warp422(const StageK &k_in, bool in422, int interp, int float_mode, bool grid_forced, void *stream) {
#elif LUTR_K420_OUT == 3
int launch_k420_warp444(const StageK &k_in, bool in422, int interp, int float_mode, bool grid_forced, void *stream) {
#else
int launch_k420_warp(const StageK &k_in, bool in422, int interp, int float_mode, bool grid_forced, void *stream) {
#endif
    StageK k = k_in;
    const bool fma = float_mode == LUTR_FLOAT_FMA, clamp = k.clamp_out != 0;
    const bool fold_nomin = k.fk.idx_fold && k.wk_nomin;
    const bool all_widths = !fma && !clamp && fold_nomin;
    const long ch = k.h / 2;
    int best_nk = 0;
    long best_grid = 1;
    double best = 1e300;
    for (int nk = all_widths ? 2 : 4; nk <= 4; ++nk) {
        if (k.wk_unroll >= 2 && k.wk_unroll <= 4 && all_widths && nk != k.wk_unroll) continue;
        if (64 * nk > k.w) continue;
        const int W = 64 * nk, nw = (k.w + W - 1) / W;
        const long total_c = (long)k.nframes * nw * ch;
        long grid = k.grid_ctas;
        if (!grid_forced && total_c / (kWarps420 * grid) < 16) grid = total_c / (kWarps420 * 16) > 0 ? total_c / (kWarps420 * 16) : 1;
        const double share = (double)total_c / (double)(kWarps420 * grid);  // chroma rows per warp
        const double cost = (2.0 * share + (kRowLocal ? 0.0 : 6.0)) * (W + 24);
        if (cost < best) best = cost, best_nk = nk, best_grid = grid;
    }
    if (!best_nk) return LUTR_ERR_UNSUPPORTED;  // row narrower than the narrowest window: the caller uses the band kernel
    k.tiles_per_row = (k.w + 64 * best_nk - 1) / (64 * best_nk);
    const int rmode = k.range_pass ? 1 : (k.any.out_full ? 2 : 0);
    KernelFn fn = interp == LUTR_INTERP_TRILINEAR ? pick420<LUTR_INTERP_TRILINEAR, 2>(fma, clamp, fold_nomin, rmode, in422, k.lut.n, best_nk)
                                                  : pick420<LUTR_INTERP_TETRAHEDRAL, 3>(fma, clamp, fold_nomin, rmode, in422, k.lut.n, best_nk);
    fn<<<(unsigned)best_grid, kWarps420 * 32, 0, (cudaStream_t)stream>>>(k);
    return cudaGetLastError() == cudaSuccess ? LUTR_OK : LUTR_ERR_CUDA;
}

}  // namespace lutr
