// Generic fused tile kernel, dispatch: the run-time-format instantiations (every layout / depth / range / interp combination of
// include/lutr.h) and the hand-over to the fixed-format instantiations of the cells the reference's options reach most
// (k_any_fixed8.cu, k_any_fixed10.cu).  Kernel and arithmetic: k_any.cuh.
#include "k_any.cuh"

namespace lutr {

namespace {

template <int INTERP>
AnyKernelFn pick(bool fma) {
    // FFmpeg has assembly (contracted multiply-adds) for tetrahedral and trilinear only
    if constexpr (INTERP == LUTR_INTERP_TETRAHEDRAL || INTERP == LUTR_INTERP_TRILINEAR) {
        if (fma) return k_any<INTERP, true, -1>;
    }
    return k_any<INTERP, false, -1>;
}

}  // namespace

int launch_k_any(const StageK &k, int interp, int float_mode, void *stream) {
    const bool fma = float_mode == LUTR_FLOAT_FMA;
    // fixed-format instantiations: tetrahedral / trilinear in C float order, limited-range output, no 10-bit pre-pass
    // (stage option wk_var = 4000 keeps the run-time-format kernel, for A / B runs and the parity tests of both)
    if (!fma && !k.any.pre10 && !k.any.out_full && k.wk_var != 4000 && (interp == LUTR_INTERP_TETRAHEDRAL || interp == LUTR_INTERP_TRILINEAR)) {
        const int code = any_fmt_code(k.any.in_depth > 8, k.any.in_sub, k.any.out_sub, k.any.out_depth > 8);
        AnyKernelFn fx = k.any.in_depth > 8 ? any_fixed10(code, interp) : any_fixed8(code, interp);
        if (fx) return launch_any_fn(fx, k, stream);
    }
    AnyKernelFn fn = nullptr;
    switch (interp) {
        case LUTR_INTERP_NEAREST: fn = pick<LUTR_INTERP_NEAREST>(fma); break;
        case LUTR_INTERP_TRILINEAR: fn = pick<LUTR_INTERP_TRILINEAR>(fma); break;
        case LUTR_INTERP_TETRAHEDRAL: fn = pick<LUTR_INTERP_TETRAHEDRAL>(fma); break;
        case LUTR_INTERP_PYRAMID: fn = pick<LUTR_INTERP_PYRAMID>(fma); break;
        case LUTR_INTERP_PRISM: fn = pick<LUTR_INTERP_PRISM>(fma); break;
        default: return LUTR_ERR_INVALID_ARG;
    }
    return launch_any_fn(fn, k, stream);
}

}  // namespace lutr
