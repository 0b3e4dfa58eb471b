// Fixed-format instantiations of the generic tile kernel (k_any.cuh) for 8-bit sources --
//   yuv420p -> yuv420p       the default delivery of an 8-bit clip with pix_fmt / force_8bit     (reference ffmpeg.py:287-310)
//   yuv420p -> yuv444p       libx264's own choice when the reference forces no pix_fmt           
//   yuv420p -> yuv422p10le   pro-mode master from an 8-bit camera file                           (ffmpeg.py:417-433)
#include "k_any.cuh"

namespace lutr {

namespace {
template <int CODE>
AnyKernelFn both(int interp) {
    return interp == LUTR_INTERP_TRILINEAR ? k_any<LUTR_INTERP_TRILINEAR, false, CODE> : k_any<LUTR_INTERP_TETRAHEDRAL, false, CODE>;
}
}  // namespace

AnyKernelFn any_fixed8(int code, int interp) {
    switch (code) {
        case any_fmt_code(false, 3, 3, false): return both<any_fmt_code(false, 3, 3, false)>(interp);
        case any_fmt_code(false, 3, 0, false): return both<any_fmt_code(false, 3, 0, false)>(interp);
        case any_fmt_code(false, 3, 1, true): return both<any_fmt_code(false, 3, 1, true)>(interp);
    }
    return nullptr;
}

}  // namespace lutr
