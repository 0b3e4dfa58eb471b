// Fixed-format instantiations of the generic tile kernel (k_any.cuh) for 10-bit sources: the cells the reference's options reach when
// a warp-window kernel does not apply (widths off the 8-column grid, rows narrower than a window) --
//   yuv422p10le -> yuv422p10le   pro-mode master from a 10-bit 4:2:2 source          (reference ffmpeg.py:417-433)
//   yuv420p10le -> yuv422p10le   pro-mode master from 10-bit HEVC / camera 4:2:0
//   yuv420p10le -> yuv420p10le   libx265 "preserve"                                   (ffmpeg.py:297-298)
//   yuv420p10le -> yuv420p       10-bit source to the default libx264                 (ffmpeg.py:299-302)
//   yuv422p10le -> yuv420p
#include "k_any.cuh"

namespace lutr {

namespace {
template <int CODE>
AnyKernelFn both(int interp) {
    return interp == LUTR_INTERP_TRILINEAR ? k_any<LUTR_INTERP_TRILINEAR, false, CODE> : k_any<LUTR_INTERP_TETRAHEDRAL, false, CODE>;
}
}  // namespace

AnyKernelFn any_fixed10(int code, int interp) {
    switch (code) {
        case any_fmt_code(true, 1, 1, true): return both<any_fmt_code(true, 1, 1, true)>(interp);
        case any_fmt_code(true, 3, 1, true): return both<any_fmt_code(true, 3, 1, true)>(interp);
        case any_fmt_code(true, 3, 3, true): return both<any_fmt_code(true, 3, 3, true)>(interp);
        case any_fmt_code(true, 3, 3, false): return both<any_fmt_code(true, 3, 3, false)>(interp);
        case any_fmt_code(true, 1, 3, false): return both<any_fmt_code(true, 1, 3, false)>(interp);
    }
    return nullptr;
}

}  // namespace lutr
