// C-ABI implementation (include/lutr.h): contexts, LUT upload, stage set-up, launches, host<->device pipelining.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <poll.h>
#include <pthread.h>
#include <signal.h>
#include <sys/ioctl.h>
#include <sys/uio.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "lutr_internal.h"

using namespace lutr;

struct lutr_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0;
    // freed LUT buffers, kept for the next upload of the same size (a batch queue with per-clip LUTs would otherwise pay a
    // cudaMalloc + cudaFree per clip: 3-4 ms, as long as the GPU needs for five 4K frames)
    std::mutex pool_mu;
    std::vector<std::pair<size_t, void *>> pool;
};

struct lutr_lut {
    lutr_ctx *ctx = nullptr;
    int size = 0;
    float scale[3] = {1.f, 1.f, 1.f};
    float4 *dev = nullptr;   // size^3 nodes + tail padding
    size_t dev_bytes = 0;
    bool needs_clamp = false;  // some entry lies outside [0,1]: the truncating store needs its clip
    mutable cudaTextureObject_t tex = 0;  // experiment builds (wk_var 2001 / 2002): the table as a point-sampled linear texture
};

// device copy of one index table; depends on (bit depth, LUT size, domain scale) only, never overwritten once built

struct lutr_stage {
    lutr_ctx *ctx = nullptr;
    lutr_stage_desc d{};
    Yuv2RgbK f{};
    Rgb2YuvK b{};
    int depth = 8;
    size_t in_bytes = 0, out_bytes = 0;  // packed frame sizes
    cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;
    // host pipeline: NBUF device slots of `chunk` frames each
    static constexpr int NBUF = 3;
    int chunk = 0;
    int chunk_mb = 64;     // host pipeline: input megabytes per slot (option "host_chunk_mb")
    uint8_t *dev_in[NBUF] = {nullptr, nullptr, nullptr}, *dev_out[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[NBUF]{}, ev_done[NBUF]{}, ev_out[NBUF]{};
    bool ev_ok = false;
    // tunables
    int rows_per_cta = 0;  // 0 = auto
    int use_fast = 1;      // 0 forces the generic kernels (testing)
    bool any_only = false; // the combination exists in the generic tile kernel (k_any.cuh) only
    bool convert_only = false;  // interp = LUTR_INTERP_NONE: LUT-less conversion of the master (k_convert.cu)
    int pre10 = 0;         // full-range 10-bit source through the 8-bit intermediate: 1 = with pc -> tv, 2 = range kept
    int force_any = 0;     // option "any_kernel": 1 routes every launch through the generic tile kernel (testing)
    int grid_ctas = 0;     // 0 = min_blocks CTAs per SM
    int min_blocks = 2;    // occupancy variant of the fast kernels (CTAs per SM)
    int wk_pk = 1;         // warp-window kernel: packed f32x2 instructions
    int wk_unroll = 4;     // warp-window kernel: pass B steps unrolled together (experiment builds)
    int wk_var = -1;       // experiment builds: variant of the headline kernel
    int wk_carveout = 0;   // experiments: preferred shared-memory carveout (percent)
    int wk_tma = 0;        // 4:2:2 10-bit stage: 1 = the row-staging kernel (k422_tma.cu: TMA / cp.async ring) when the geometry allows; measured slower
    int wk_nk_forced = 0;  // 4:2:0 warp-window kernel: wk_unroll was set explicitly (2 or 4 steps of 64 columns per window)
    int wk_minb = 3;       // warp-window kernel: CTAs per SM the register budget is set for
    bool wk_minb_set = false;  // ... set explicitly (the ordered-row-walk kernels default to 2 CTAs per SM: 128 registers, no spills)
    std::atomic<int64_t> launches{0};
    const char *last_kernel = "";  // kernel family of the most recent enqueue (bench.py's roofline.kernel)
};

namespace {

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return LUTR_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

int plane_rows(int fmt, int h, int p) { return p > 0 ? h >> (pix_sub(fmt) >> 1) : h; }
int plane_row_bytes(int fmt, int w, int p) {
    const int bps = pix_depth(fmt) > 8 ? 2 : 1;
    return (p == 0 ? w : w >> (pix_sub(fmt) & 1)) * bps;
}
size_t frame_bytes(int fmt, int w, int h) {
    size_t n = 0;
    for (int p = 0; p < 3; ++p) n += (size_t)plane_rows(fmt, h, p) * plane_row_bytes(fmt, w, p);
    return n;
}

void packed_planes(int fmt, int w, int h, uint8_t *base, lutr_planes *pl) {
    size_t off = 0;
    for (int p = 0; p < 3; ++p) {
        pl->data[p] = base + off;
        pl->stride[p] = plane_row_bytes(fmt, w, p);
        off += (size_t)plane_rows(fmt, h, p) * pl->stride[p];
    }
    pl->frame_step = (int64_t)off;
}

int upload_lut(lutr_ctx *ctx, const CubeLut &c, lutr_lut **out) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    lutr_lut *l = new (std::nothrow) lutr_lut;
    if (!l) return LUTR_ERR_NOMEM;
    l->ctx = ctx;
    l->size = c.size;
    std::memcpy(l->scale, c.scale, sizeof(l->scale));
    const size_t n = (size_t)c.size * c.size * c.size;
    const size_t pad = (size_t)c.size * c.size + c.size + 2;  // zeroed tail: any "next" offset stays inside the buffer
    std::vector<float4> host(n + pad, make_float4(0.f, 0.f, 0.f, 0.f));
    for (size_t i = 0; i < n; ++i) {
        if (!std::isfinite(c.rgb[3 * i]) || !std::isfinite(c.rgb[3 * i + 1]) || !std::isfinite(c.rgb[3 * i + 2])) {
            delete l;
            set_error("lut: non-finite table entry %zu", i);
            return LUTR_ERR_INVALID_DATA;
        }
        host[i] = make_float4(c.rgb[3 * i], c.rgb[3 * i + 1], c.rgb[3 * i + 2], 0.f);
        for (int ch = 0; ch < 3; ++ch)
            if (c.rgb[3 * i + ch] < 0.f || c.rgb[3 * i + ch] > 1.f) l->needs_clamp = true;
    }
    l->dev_bytes = host.size() * sizeof(float4);
    {
        std::lock_guard<std::mutex> lock(ctx->pool_mu);
        for (size_t i = 0; i < ctx->pool.size(); ++i)
            if (ctx->pool[i].first == l->dev_bytes) {
                l->dev = (float4 *)ctx->pool[i].second;
                ctx->pool.erase(ctx->pool.begin() + (long)i);
                break;
            }
    }
    cudaError_t e = l->dev ? cudaSuccess : cudaMalloc((void **)&l->dev, l->dev_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(l->dev, host.data(), l->dev_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("lut upload: %s", cudaGetErrorString(e));
        if (l->dev) cudaFree(l->dev);
        delete l;
        return LUTR_ERR_CUDA;
    }
    *out = l;
    return LUTR_OK;
}

int check_planes(const lutr_stage *s, const lutr_planes *pl, int fmt, int nframes, const char *what) {
    for (int p = 0; p < 3; ++p) {
        if (!pl->data[p]) {
            set_error("%s plane %d is NULL", what, p);
            return LUTR_ERR_INVALID_ARG;
        }
        if (pl->stride[p] < plane_row_bytes(fmt, s->d.width, p)) {
            set_error("%s stride[%d]=%d is smaller than the row (%d bytes)", what, p, pl->stride[p], plane_row_bytes(fmt, s->d.width, p));
            return LUTR_ERR_INVALID_ARG;
        }
        // vector accesses: one 8-pixel group of Y (16 B at 10 bit, 8 B at 8 bit), half of that for a chroma group; stages that only
        // ever run the generic tile kernel (pixel pairs) need 4 B / 2 B for Y and one sample for chroma
        const bool wide = pix_depth(fmt) > 8;
        const int ya = s->any_only ? (wide ? 4 : 2) : (wide ? 16 : 8);
        const int a = p == 0 || !(pix_sub(fmt) & 1) ? ya : (s->any_only ? (wide ? 2 : 1) : ya / 2);
        if (((uintptr_t)pl->data[p] % a) || (pl->stride[p] % a) || (nframes > 1 && (pl->frame_step % ya))) {
            set_error("%s plane %d: pointer and stride must be %d-byte aligned, frame_step %d-byte aligned", what, p, a, ya);
            return LUTR_ERR_UNSUPPORTED;
        }
    }
    return LUTR_OK;
}

// The fast kernels need luma rows whose 15-bit sum cannot saturate.
bool fast_path_ok(const lutr_stage *s, const lutr_lut *lut) {
    if (!s->use_fast || s->d.interp == LUTR_INTERP_NEAREST) return false;
    const int maxv = (1 << s->depth) - 1;
    const long ysum = (long)(s->b.ry + s->b.gy + s->b.by) * maxv;
    if (s->b.ry < 0 || s->b.gy < 0 || s->b.by < 0) return false;
    // 10 bit: y14 = (sum + (16<<17) + (1<<10)) >> 11 must keep 2*y14 <= 32767;  8 bit: (sum + (32<<14) + (1<<8)) >> 9 likewise
    const long y14 = s->depth == 10 ? (ysum + (16 << 17) + (1 << 10)) >> 11 : (ysum + (32 << 14) + (1 << 8)) >> 9;
    return 2 * y14 <= 32767;
}

int enqueue(lutr_stage *s, const lutr_lut *lut, const lutr_planes *in, const lutr_planes *out, int nframes, cudaStream_t st) {
    StageK k{};
    k.w = s->d.width, k.h = s->d.height, k.nframes = nframes;
    if (s->convert_only) {  // pro mode stage 2 without a LUT: one streaming kernel, at most 65535 frames per launch (grid z)
        for (int f0 = 0; f0 < nframes; f0 += 65535) {
            k.nframes = std::min(65535, nframes - f0);
            for (int p = 0; p < 3; ++p) {
                k.p.in[p] = (const uint8_t *)in->data[p] + (size_t)f0 * in->frame_step, k.p.out[p] = (uint8_t *)out->data[p] + (size_t)f0 * out->frame_step;
                k.p.in_stride[p] = in->stride[p], k.p.out_stride[p] = out->stride[p];
            }
            k.p.in_step = nframes > 1 ? in->frame_step : 0, k.p.out_step = nframes > 1 ? out->frame_step : 0;
            const int rc = launch_k_convert(k, s->d.out_fmt, st);
            if (rc != LUTR_OK) {
                set_error("kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return rc;
            }
            s->launches.fetch_add(1, std::memory_order_relaxed);
        }
        s->last_kernel = "k_convert";
        return LUTR_OK;
    }
    k.range_pass = (s->d.in_range == LUTR_RANGE_PC && s->d.norm_range == LUTR_RANGE_TV) ? 1 : 0;
    k.f = s->f, k.b = s->b;
    const float lut_max = (float)(lut->size - 1);
    const int maxv = (1 << s->depth) - 1;
    k.lut.tab = lut->dev;
    k.lut.n = lut->size, k.lut.n2 = lut->size * lut->size;
    k.lut.lut_max = lut_max;
    k.lut.in_scale = 1.0f / maxv;
    k.lut.sr = lut->scale[0] * lut_max, k.lut.sg = lut->scale[1] * lut_max, k.lut.sb = lut->scale[2] * lut_max;
    k.lut.out_scale = (float)maxv;
#ifdef LUTR_WK_EXPERIMENTS
    if (s->wk_var >= 2000 && !lut->tex) {  // texture-path experiment: same memory, fetched by element index
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = lut->dev, rd.res.linear.desc = cudaCreateChannelDesc<float4>(), rd.res.linear.sizeInBytes = lut->dev_bytes;
        cudaTextureDesc td{};
        td.filterMode = cudaFilterModePoint, td.readMode = cudaReadModeElementType, td.addressMode[0] = cudaAddressModeClamp;
        if (cudaCreateTextureObject(&lut->tex, &rd, &td, nullptr) != cudaSuccess) lut->tex = 0, cudaGetLastError();
    }
    k.lut.tex = (unsigned long long)lut->tex;
#endif
    for (int p = 0; p < 3; ++p) {
        k.p.in[p] = (const uint8_t *)in->data[p], k.p.out[p] = (uint8_t *)out->data[p];
        k.p.in_stride[p] = in->stride[p], k.p.out_stride[p] = out->stride[p];
    }
    k.p.in_step = nframes > 1 ? in->frame_step : 0, k.p.out_step = nframes > 1 ? out->frame_step : 0;
    int rc;
    k.idx_mode = -1;
    k.min_blocks = s->min_blocks;
    k.clamp_out = lut->needs_clamp ? 1 : 0;
    {   // lut3d index constants of the packed evaluation (warp_common.cuh lut3d_pair), shared by the warp-window kernels and k_any
        FastK &fk = k.fk;
        // arithmetic index: fold a power-of-two LUT scale into the input scale (exact: scaling by 2^k commutes with rounding)
        int exp2 = 0;
        const float sr = k.lut.sr;
        const bool one_scale = k.lut.sg == sr && k.lut.sb == sr;  // per-channel DOMAIN scales take the general (two-multiply) index
        fk.idx_fold = (one_scale && sr > 0.f && std::frexp(sr, &exp2) == 0.5f) ? 1 : 0;
        fk.idx_mul = fk.idx_fold ? k.lut.in_scale * sr : k.lut.in_scale;
        fk.idx_sr[0] = fk.idx_fold ? 1.f : sr, fk.idx_sr[1] = fk.idx_fold ? 1.f : k.lut.sg, fk.idx_sr[2] = fk.idx_fold ? 1.f : k.lut.sb;
        fk.idx_add = -(fk.idx_mul * 8388608.f);
        fk.maxv = maxv;
        k.wk_one = 1.0f;
        // the clip of the LUT coordinate to lut_max is dead code unless some code value overshoots it: try them all
        k.wk_nomin = 1;
        for (int x = 0; x <= maxv; ++x) {
            volatile float a = (float)x * fk.idx_mul;  // volatile: one rounding per operation
            for (int ch = 0; ch < 3; ++ch) {
                volatile float c = a * fk.idx_sr[ch];
                if (c > lut_max) k.wk_nomin = 0;
            }
        }
    }
    bool warp_kernel = false;
    if (fast_path_ok(s, lut)) {
        // warp-window kernels: 8-column lanes (w % 8, 16-byte luma / 8-byte chroma vector accesses: check_planes), 32-bit unit counter
        const long tiles = (long)((k.w + 239) / 240) * k.h * nframes;
        warp_kernel = (k.w % 8) == 0 && k.w >= 16 && tiles < (1l << 31);
        // the 4:2:0 windows are 128 or 256 columns and never wider than the row (the last one is shifted, not clipped)
        if (s->d.in_fmt == LUTR_PIX_YUV420P && k.w < 256 && s->d.out_fmt == s->d.in_fmt) warp_kernel = false;
    }
    if (warp_kernel) {
        k.idx_mode = 3;
        const uint32_t magic = 0x4B000000u;  // bits of 8388608.0f: bias of the truncated LUT outputs
        FastK &fk = k.fk;
        const uint32_t yK = (uint32_t)(-s->f.yOff10 * s->f.yC) + (1u << 19);
        fk.yA = (uint32_t)(128 * s->f.yC);
        fk.vr4 = (uint32_t)(4 * s->f.vr), fk.vg4 = (uint32_t)(4 * s->f.vg), fk.ug4 = (uint32_t)(4 * s->f.ug), fk.ub4 = (uint32_t)(4 * s->f.ub);
        fk.kR = yK - 65536u * (uint32_t)s->f.vr;
        fk.kG = yK - 65536u * (uint32_t)s->f.vg - 65536u * (uint32_t)s->f.ug;
        fk.kB = yK - 65536u * (uint32_t)s->f.ub;
        if (s->depth == 10) {
            fk.kY = (uint32_t)((16 << 17) + (1 << 10) + (8 << 11)) - magic * (uint32_t)(s->b.ry + s->b.gy + s->b.by);
            fk.kU = (uint32_t)((128 << 17) + (1 << 10)) - magic * (uint32_t)(s->b.ru + s->b.gu + s->b.bu);
            fk.kV = (uint32_t)((128 << 17) + (1 << 10)) - magic * (uint32_t)(s->b.rv + s->b.gv + s->b.bv);
            fk.cu5[0] = 32u * (uint32_t)s->b.ru, fk.cu5[1] = 32u * (uint32_t)s->b.gu, fk.cu5[2] = 32u * (uint32_t)s->b.bu;
            fk.cv5[0] = 32u * (uint32_t)s->b.rv, fk.cv5[1] = 32u * (uint32_t)s->b.gv, fk.cv5[2] = 32u * (uint32_t)s->b.bv;
            fk.kU5 = 32u * fk.kU, fk.kV5 = 32u * fk.kV;
        } else {
            fk.kY = (uint32_t)((32 << 14) + (1 << 8) + (32 << 9)) - magic * (uint32_t)(s->b.ry + s->b.gy + s->b.by);
            fk.kU = (uint32_t)((256 << 15) + (1 << 9)) - 2u * magic * (uint32_t)(s->b.ru + s->b.gu + s->b.bu);
            fk.kV = (uint32_t)((256 << 15) + (1 << 9)) - 2u * magic * (uint32_t)(s->b.rv + s->b.gv + s->b.bv);
        }
        fk.n2b = k.lut.n2 * 16, fk.nb = k.lut.n * 16, fk.totalb = fk.n2b + fk.nb + 16;
        {
            k.tiles_per_row = (k.w + 239) / 240;  // 4:2:2; the 4:2:0 windows are sized below
            k.wk_pk = s->wk_pk;
            k.wk_unroll = s->wk_unroll;
            k.wk_var = s->wk_var;
            k.wk_carveout = s->wk_carveout;
            k.min_blocks = s->d.interp == LUTR_INTERP_TRILINEAR ? 2 : s->wk_minb;
        }
    }
    k.grid_ctas = s->grid_ctas > 0 ? s->grid_ctas : k.min_blocks * s->ctx->sm_count;
    // output range conversion constants (full-range outputs of the 8-bit warp-window kernels and of the generic kernel)
    k.any.out_full = s->d.out_range == LUTR_RANGE_PC ? 1 : 0;
    range_coeffs(pix_depth(s->d.out_fmt), true, false, &k.any.out_lc, &k.any.out_lo);
    range_coeffs(pix_depth(s->d.out_fmt), true, true, &k.any.out_cc, &k.any.out_co);
    // generic tile kernel: combinations without a specialised kernel, and specialised combinations whose fast path does not apply
    // (interp = nearest / pyramid / prism, widths that are not a multiple of 8, rows narrower than a window)
    auto run_any = [&]() {
        AnyK &a = k.any;
        a.in_sub = pix_sub(s->d.in_fmt), a.in_depth = pix_depth(s->d.in_fmt);
        a.out_sub = pix_sub(s->d.out_fmt), a.out_depth = pix_depth(s->d.out_fmt);
        a.out_full = s->d.out_range == LUTR_RANGE_PC ? 1 : 0;
        a.pre10 = s->pre10;
        if (s->pre10) k.range_pass = 0;  // the pre-pass already delivers the normalised 8-bit samples
        range_coeffs(8, false, false, &a.in_lc, &a.in_lo);
        range_coeffs(8, false, true, &a.in_cc, &a.in_co);
        range_coeffs(a.out_depth, true, false, &a.out_lc, &a.out_lo);
        range_coeffs(a.out_depth, true, true, &a.out_cc, &a.out_co);
        k.grid_ctas = s->grid_ctas;
        s->last_kernel = "k_any";
        return launch_k_any(k, s->d.interp, s->d.float_mode, st);
    };
    if (s->any_only || s->force_any) {
        rc = run_any();
    } else if (s->depth == 10 && s->d.out_fmt == LUTR_PIX_YUV420P) {
        if (!warp_kernel) {
            rc = run_any();
        } else {
            k.rows_per_cta = 0, k.bands = 1;
            k.min_blocks = s->wk_minb_set ? s->wk_minb : 2;  // measured: +5..10 % over 3 CTAs x 80 registers (spills)
            // more CTAs than resident slots: the hardware deals them out as others retire (profiles/r02_grid_sweep_other_kernels.jsonl);
            // every share pays 6 halo rows of the vertical output filter, which bounds the useful multiple
            k.grid_ctas = s->grid_ctas > 0 ? s->grid_ctas : 6 * k.min_blocks * s->ctx->sm_count;
            rc = launch_k10_warp8(k, s->d.in_fmt == LUTR_PIX_YUV420P10LE, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = s->d.in_fmt == LUTR_PIX_YUV420P10LE ? "k10_warp8<420>" : "k10_warp8<422>";
        }
    } else if (s->d.in_fmt == LUTR_PIX_YUV420P10LE || s->d.out_fmt == LUTR_PIX_YUV420P10LE) {
        if (!warp_kernel) {
            rc = run_any();
        } else {
            const bool in420 = s->d.in_fmt == LUTR_PIX_YUV420P10LE, out420 = s->d.out_fmt == LUTR_PIX_YUV420P10LE;
            k.rows_per_cta = 0, k.bands = 1;
            k.min_blocks = s->wk_minb_set ? s->wk_minb : 2;  // measured: +5..10 % over 3 CTAs x 80 registers (spills)
            // as above; a 4:2:2 output side is row-local (no halo rows per share), so shares can be much shorter
            k.grid_ctas = s->grid_ctas > 0 ? s->grid_ctas : (out420 ? 6 : 24) * k.min_blocks * s->ctx->sm_count;
            rc = launch_k10_warp(k, in420, out420, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = !out420 ? "k10_warp<420,422>" : (in420 ? "k10_warp<420,420>" : "k10_warp<422,420>");
        }
    } else if (!warp_kernel) {
        rc = run_any();
    } else if (s->d.in_fmt == LUTR_PIX_YUV422P10LE || s->d.in_fmt == LUTR_PIX_YUV444P10LE) {
        const bool in444 = s->d.in_fmt == LUTR_PIX_YUV444P10LE;
        {
            // Work unit of a CTA = (frame, block of R rows, window); its 8 warps take interleaved rows.  Measured on the B200
            // (profiles/r02_grid_rows_sweep.jsonl): a persistent grid with a fixed share of units per CTA leaves a sixth of the
            // warp slots idle (the warp scheduler is not fair, CTAs drift apart and the fast ones run out of work), one unit per
            // CTA with the hardware dealing CTAs to SMs as they retire is 9 % faster; long units (R = 128: 16 rows per warp)
            // amortise the row prefetch.  So: R as large as still leaves ~6 units per resident CTA slot, grid = all units.
            long r = s->rows_per_cta;
            if (r <= 0) {
                const long slots = 3l * s->ctx->sm_count;
                r = (long)k.tiles_per_row * k.h * nframes / (6l * slots);
                r = r < 16 ? 16 : (r > 128 ? 128 : r);
            }
            k.rows_per_cta = (int)(r / 8 * 8 < 8 ? 8 : r / 8 * 8);
            if (s->grid_ctas <= 0) k.grid_ctas = (int)std::min<long>((long)k.tiles_per_row * ((k.h + k.rows_per_cta - 1) / k.rows_per_cta) * nframes, 1l << 30);
        }
        k.bands = 1;
        if (!in444 && s->wk_tma && k422_tma_ok(k)) {
            k.wk_bulk = s->wk_tma == 2;
            rc = launch_k422_tma(k, s->d.interp, s->d.float_mode, st);
            s->last_kernel = s->wk_tma == 2 ? "k422_tma" : "k422_ring";
        } else {
            rc = launch_k422_warp(k, in444, s->d.interp, s->d.float_mode, st);
            s->last_kernel = in444 ? "k422_warp<444>" : "k422_warp";
        }
    } else if (s->d.in_fmt == LUTR_PIX_YUV420P || s->d.in_fmt == LUTR_PIX_YUV422P) {
        // 4:2:0 / 4:2:2 8-bit source warp-window kernel (window width and grid are chosen in the launcher)
        const bool in422 = s->d.in_fmt == LUTR_PIX_YUV422P;
        k.wk_unroll = s->wk_nk_forced ? s->wk_unroll : 0;
        k.rows_per_cta = 0, k.bands = 1;
        if (s->grid_ctas <= 0)  // as above: 4 x the resident CTAs with the 6 halo rows of a 4:2:0 output (2 x for trilinear), 16 x without
            k.grid_ctas *= s->d.out_fmt != LUTR_PIX_YUV420P ? 16 : (s->d.interp == LUTR_INTERP_TRILINEAR ? 2 : 4);
        if (s->d.out_fmt == LUTR_PIX_YUV422P10LE) {
            rc = launch_k420_warp10(k, in422, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = in422 ? "k420_warp10<422>" : "k420_warp10";
        } else if (s->d.out_fmt == LUTR_PIX_YUV422P) {
            rc = launch_k420_warp422(k, in422, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = in422 ? "k420_warp422<422>" : "k420_warp422";
        } else if (s->d.out_fmt == LUTR_PIX_YUV444P) {
            rc = launch_k420_warp444(k, in422, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = in422 ? "k420_warp444<422>" : "k420_warp444";
        } else {
            rc = launch_k420_warp(k, in422, s->d.interp, s->d.float_mode, s->grid_ctas > 0, st);
            s->last_kernel = in422 ? "k420_warp<422>" : "k420_warp";
        }
        if (rc == LUTR_ERR_UNSUPPORTED) rc = run_any();  // this LUT / option combination needs a row of at least 256 columns
    } else {
        rc = run_any();
    }
    if (rc != LUTR_OK) {
        set_error("kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    s->launches.fetch_add(1, std::memory_order_relaxed);
    return LUTR_OK;
}

void free_host_pipeline(lutr_stage *s) {
    for (int i = 0; i < lutr_stage::NBUF; ++i) {
        if (s->dev_in[i]) cudaFree(s->dev_in[i]);
        if (s->dev_out[i]) cudaFree(s->dev_out[i]);
        s->dev_in[i] = s->dev_out[i] = nullptr;
    }
    s->chunk = 0;
}

}  // namespace

extern "C" {

int lutr_version(void) { return LUTR_VERSION; }

const char *lutr_status_string(int status) {
    switch (status) {
        case LUTR_OK: return "ok";
        case LUTR_ERR_INVALID_ARG: return "invalid argument";
        case LUTR_ERR_INVALID_DATA: return "invalid data found when processing input";
        case LUTR_ERR_LUT_SIZE: return "too large or invalid 3D LUT size";
        case LUTR_ERR_IO: return "i/o error";
        case LUTR_ERR_NOMEM: return "out of memory";
        case LUTR_ERR_CUDA: return "CUDA error / no usable sm_100 device";
        case LUTR_ERR_UNSUPPORTED: return "unsupported configuration";
        case LUTR_ERR_CANCELLED: return "cancelled";
        default: return "unknown status";
    }
}

const char *lutr_last_error(void) { return get_error(); }

int lutr_cube_parse(const char *text, size_t len, int32_t *size, float scale[3], float **table) {
    if (!size || !table) {
        set_error("lutr_cube_parse: NULL output");
        return LUTR_ERR_INVALID_ARG;
    }
    CubeLut c;
    int rc = parse_cube(text, len, c);
    if (rc != LUTR_OK) return rc;
    float *t = (float *)std::malloc(c.rgb.size() * sizeof(float));
    if (!t) return LUTR_ERR_NOMEM;
    std::memcpy(t, c.rgb.data(), c.rgb.size() * sizeof(float));
    *size = c.size;
    if (scale) std::memcpy(scale, c.scale, 3 * sizeof(float));
    *table = t;
    return LUTR_OK;
}

int lutr_cube_parse_file(const char *path, int32_t *size, float scale[3], float **table) {
    if (!size || !table) {
        set_error("lutr_cube_parse_file: NULL output");
        return LUTR_ERR_INVALID_ARG;
    }
    CubeLut c;
    int rc = parse_cube_file(path, c);
    if (rc != LUTR_OK) return rc;
    float *t = (float *)std::malloc(c.rgb.size() * sizeof(float));
    if (!t) return LUTR_ERR_NOMEM;
    std::memcpy(t, c.rgb.data(), c.rgb.size() * sizeof(float));
    *size = c.size;
    if (scale) std::memcpy(scale, c.scale, 3 * sizeof(float));
    *table = t;
    return LUTR_OK;
}

static int export_table(const CubeLut &c, int32_t *size, float scale[3], float **table) {
    float *t = (float *)std::malloc(c.rgb.size() * sizeof(float));
    if (!t) return LUTR_ERR_NOMEM;
    std::memcpy(t, c.rgb.data(), c.rgb.size() * sizeof(float));
    *size = c.size;
    if (scale) std::memcpy(scale, c.scale, 3 * sizeof(float));
    *table = t;
    return LUTR_OK;
}

int lutr_lut_parse_text(const char *text, size_t len, int32_t dialect, int32_t *size, float scale[3], float **table) {
    if (!size || !table) {
        set_error("lutr_lut_parse_text: NULL output");
        return LUTR_ERR_INVALID_ARG;
    }
    CubeLut c;
    const int rc = parse_lut_text(text, len, dialect, c);
    return rc != LUTR_OK ? rc : export_table(c, size, scale, table);
}

int lutr_lut_parse_file(const char *path, int32_t *size, float scale[3], float **table) {
    if (!size || !table) {
        set_error("lutr_lut_parse_file: NULL output");
        return LUTR_ERR_INVALID_ARG;
    }
    CubeLut c;
    const int rc = parse_lut_file(path, c);
    return rc != LUTR_OK ? rc : export_table(c, size, scale, table);
}

void lutr_free(void *p) { std::free(p); }

int lutr_device_count(int32_t *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        if (count) *count = 0;
        return LUTR_ERR_CUDA;
    }
    if (count) *count = n;
    return LUTR_OK;
}

int lutr_ctx_create(int32_t device, lutr_ctx **out) {
    if (!out) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return LUTR_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (0..%d)", device, n - 1);
        return LUTR_ERR_INVALID_ARG;
    }
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this build contains sm_100a code only", device, prop.major, prop.minor);
        return LUTR_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    lutr_ctx *c = new (std::nothrow) lutr_ctx;
    if (!c) return LUTR_ERR_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major, c->cc_minor = prop.minor;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    // No persisting-L2 window for the LUT: a 0.6-4.4 MB table that every warp gathers from stays in the 126 MB L2 by itself while
    // the frames stream past with L1::no_allocate accesses (measured in round 1: the window changed nothing), and a per-stream
    // access-policy attribute is state that outlives the call on the caller's stream.
    *out = c;
    return LUTR_OK;
}

void lutr_ctx_destroy(lutr_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->pool.empty()) {
        cudaSetDevice(ctx->device);
        for (auto &b : ctx->pool) cudaFree(b.second);
    }
    delete ctx;
}

int lutr_ctx_info(const lutr_ctx *ctx, int32_t *device, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    if (!ctx) return LUTR_ERR_INVALID_ARG;
    if (device) *device = ctx->device;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return LUTR_OK;
}

int lutr_host_alloc(lutr_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return LUTR_ERR_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return LUTR_OK;
}

void lutr_host_free(lutr_ctx *ctx, void *p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
}

int lutr_lut_load_cube(lutr_ctx *ctx, const char *path, lutr_lut **out) {
    if (!ctx || !out) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    CubeLut c;
    int rc = parse_cube_file(path, c);
    if (rc != LUTR_OK) return rc;
    return upload_lut(ctx, c, out);
}

int lutr_lut_load_file(lutr_ctx *ctx, const char *path, lutr_lut **out) {
    if (!ctx || !out) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    CubeLut c;
    const int rc = parse_lut_file(path, c);
    if (rc != LUTR_OK) return rc;
    return upload_lut(ctx, c, out);
}

int lutr_lut_load_cube_memory(lutr_ctx *ctx, const char *text, size_t len, lutr_lut **out) {
    if (!ctx || !out) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    CubeLut c;
    int rc = parse_cube(text, len, c);
    if (rc != LUTR_OK) return rc;
    return upload_lut(ctx, c, out);
}

int lutr_lut_from_table(lutr_ctx *ctx, int32_t size, const float *rgb, const float scale[3], lutr_lut **out) {
    if (!ctx || !out || !rgb) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    if (size < 2 || size > 256) {
        set_error("lut: Too large or invalid 3D LUT size (%d)", size);
        return LUTR_ERR_LUT_SIZE;
    }
    CubeLut c;
    c.size = size;
    c.rgb.assign(rgb, rgb + (size_t)size * size * size * 3);
    for (int i = 0; i < 3; ++i) c.scale[i] = scale ? scale[i] : 1.f;
    return upload_lut(ctx, c, out);
}

int lutr_lut_size(const lutr_lut *lut) { return lut ? lut->size : LUTR_ERR_INVALID_ARG; }

void lutr_lut_destroy(lutr_lut *lut) {
    if (!lut) return;
    if (lut->dev) {
        cudaSetDevice(lut->ctx->device);
        // like cudaFree, wait for whatever still reads the table; then keep the buffer for the next LUT of this size
        cudaDeviceSynchronize();
        std::lock_guard<std::mutex> lock(lut->ctx->pool_mu);
        if (lut->ctx->pool.size() < 8) lut->ctx->pool.emplace_back(lut->dev_bytes, (void *)lut->dev);
        else cudaFree(lut->dev);
    }
    delete lut;
}

int lutr_stage_create(lutr_ctx *ctx, const lutr_stage_desc *desc, lutr_stage **out) {
    if (!ctx || !desc || !out) return LUTR_ERR_INVALID_ARG;
    *out = nullptr;
    if (desc->struct_size != sizeof(lutr_stage_desc)) {
        set_error("lutr_stage_desc.struct_size=%u, expected %zu", desc->struct_size, sizeof(lutr_stage_desc));
        return LUTR_ERR_INVALID_ARG;
    }
    const lutr_stage_desc &d = *desc;
    if (d.width <= 0 || d.height <= 0 || d.interp < 0 || d.interp > LUTR_INTERP_NONE || d.float_mode < 0 || d.float_mode > 1 ||
        d.in_matrix < 0 || d.in_matrix > 4 || d.out_matrix < 0 || d.out_matrix > 4 || d.in_range < 0 || d.in_range > 1 ||
        d.norm_range < 0 || d.norm_range > 1 || d.out_range < 0 || d.out_range > 1 || d.in_fmt < 0 ||
        d.in_fmt > LUTR_PIX_YUV444P10LE || d.out_fmt < 0 || d.out_fmt > LUTR_PIX_YUV444P10LE) {
        set_error("stage: bad geometry or enum value");
        return LUTR_ERR_INVALID_ARG;
    }
    if (d.in_range == LUTR_RANGE_TV && d.norm_range != LUTR_RANGE_TV) {
        set_error("stage: norm_range must equal in_range for a tv source");
        return LUTR_ERR_INVALID_ARG;
    }
    const bool range_pass = d.in_range == LUTR_RANGE_PC && d.norm_range == LUTR_RANGE_TV;
    // a full-range source deeper than 8 bit through the reference's 8-bit intermediate (desc.via_8bit, generic kernel)
    const bool pre10 = pix_depth(d.in_fmt) > 8 && d.in_range == LUTR_RANGE_PC && d.via_8bit != 0;
    if (pix_depth(d.in_fmt) > 8 && range_pass && !pre10) {
        // the reference's pc->tv pass always drops to 8-bit yuv4xxp (ffmpeg.py:137-143,233): hand the stage that 8-bit format
        set_error("stage: pc->tv normalisation of a 10-bit source re-quantises to 8 bit in the reference: set lutr_stage_desc.via_8bit");
        return LUTR_ERR_UNSUPPORTED;
    }
    // geometry the arithmetic models were pinned for against libswscale (tests/test_oracle_swscale.py)
    if ((d.width % 2) || d.width < 16 || d.width > 16384) {
        set_error("stage: width must be even and in [16, 16384]");
        return LUTR_ERR_UNSUPPORTED;
    }
    const bool v420 = (pix_sub(d.in_fmt) | pix_sub(d.out_fmt)) & 2;  // a vertical chroma filter on either side
    const int min_h = pix_depth(d.in_fmt) > 8 ? 14 : 12;
    if (v420 && ((d.height % 2) || d.height < min_h)) {
        set_error("stage: 4:2:0 input / output needs an even height >= %d", min_h);
        return LUTR_ERR_UNSUPPORTED;
    }
    const bool convert_only = d.interp == LUTR_INTERP_NONE;
    if (convert_only) {
        // pro mode stage 2 (reference ffmpeg.py:459-471): the master is yuv422p10le, tv range, and the conversion changes neither matrix nor range
        const bool out_ok = d.out_fmt == LUTR_PIX_YUV420P || d.out_fmt == LUTR_PIX_YUV420P10LE || d.out_fmt == LUTR_PIX_YUV422P;
        if (d.in_fmt != LUTR_PIX_YUV422P10LE || !out_ok || d.in_range != LUTR_RANGE_TV || d.out_range != LUTR_RANGE_TV ||
            d.in_matrix != d.out_matrix || (d.width % 8) || (d.height % 2) || d.height < 14) {
            set_error("stage: interp=none converts yuv422p10le (tv) to yuv420p / yuv420p10le / yuv422p with the same matrix and range; "
                      "width a multiple of 8, even height >= 14");
            return LUTR_ERR_UNSUPPORTED;
        }
    }
    // combinations with a specialised warp-window kernel; everything else runs the generic tile kernel (k_any.cuh)
    const bool tri_tet = d.interp == LUTR_INTERP_TRILINEAR || d.interp == LUTR_INTERP_TETRAHEDRAL;
    const bool same_fmt = d.in_fmt == d.out_fmt && (d.in_fmt == LUTR_PIX_YUV420P || d.in_fmt == LUTR_PIX_YUV422P10LE);
    const bool in8 = d.in_fmt == LUTR_PIX_YUV420P || d.in_fmt == LUTR_PIX_YUV422P;  // 8-bit sources of the SIMD-path arithmetic
    // yuv420p source with a row-local output side: the pro-mode master, and the GUI's 8-bit yuv422p / yuv444p choices
    const bool to_master = in8 && d.width >= 128 && tri_tet && (d.height % 2) == 0 &&  // (the kernel walks pairs of luma rows)
                           (d.out_fmt == LUTR_PIX_YUV422P10LE || d.out_fmt == LUTR_PIX_YUV422P || d.out_fmt == LUTR_PIX_YUV444P ||
                            (d.in_fmt == LUTR_PIX_YUV422P && d.out_fmt == LUTR_PIX_YUV420P));
    const bool in10 = d.in_fmt == LUTR_PIX_YUV420P10LE || d.in_fmt == LUTR_PIX_YUV422P10LE;
    const bool ordered10 = in10 && tri_tet &&
                           ((d.in_fmt == LUTR_PIX_YUV420P10LE && d.out_fmt == LUTR_PIX_YUV422P10LE) || d.out_fmt == LUTR_PIX_YUV420P10LE ||
                            d.out_fmt == LUTR_PIX_YUV420P);  // (10-bit source -> 8-bit 4:2:0 delivery: k10_warp8.cu)
    // (the warp-window kernels work on 8-column groups with 16-byte accesses: widths that are not a multiple of 8, e.g. 854x480,
    // take the generic kernel)
    const bool master444 = d.in_fmt == LUTR_PIX_YUV444P10LE && d.out_fmt == LUTR_PIX_YUV422P10LE && tri_tet;  // ProRes 4444 -> master
    // full-range output: the 8-bit warp-window kernels carry it as a run-time switch (8-bit source, 8-bit output); elsewhere generic
    const bool pc_out_ok = d.out_range == LUTR_RANGE_TV || (in8 && pix_depth(d.out_fmt) == 8 && d.norm_range == LUTR_RANGE_PC);
    const bool specialised = !pre10 && pc_out_ok && d.interp <= LUTR_INTERP_TETRAHEDRAL && (d.width % 8) == 0 &&
                             (same_fmt || to_master || ordered10 || master444);
    CUDA_TRY(cudaSetDevice(ctx->device));
    lutr_stage *s = new (std::nothrow) lutr_stage;
    if (!s) return LUTR_ERR_NOMEM;
    s->ctx = ctx;
    s->d = d;
    s->depth = pre10 ? 8 : pix_depth(d.in_fmt);  // depth of the RGB intermediate the LUT works on
    s->pre10 = pre10 ? (range_pass ? 1 : 2) : 0;
    s->any_only = !specialised && !convert_only;
    s->convert_only = convert_only;
    // YUV->RGB sees the range left by the normalisation pass
    make_yuv2rgb(d.in_matrix, d.norm_range == LUTR_RANGE_PC, s->f);
    make_rgb2yuv(d.out_matrix, s->b);
    s->in_bytes = frame_bytes(d.in_fmt, d.width, d.height);
    s->out_bytes = frame_bytes(d.out_fmt, d.width, d.height);
    cudaError_t e = cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking);
    for (int i = 0; i < lutr_stage::NBUF && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        set_error("stage: stream/event creation failed: %s", cudaGetErrorString(e));
        lutr_stage_destroy(s);
        return LUTR_ERR_CUDA;
    }
    s->ev_ok = true;
    *out = s;
    return LUTR_OK;
}

void lutr_stage_destroy(lutr_stage *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->compute) cudaStreamSynchronize(s->compute);
    free_host_pipeline(s);
    if (s->ev_ok)
        for (int i = 0; i < lutr_stage::NBUF; ++i) {
            cudaEventDestroy(s->ev_in[i]);
            cudaEventDestroy(s->ev_done[i]);
            cudaEventDestroy(s->ev_out[i]);
        }
    if (s->compute) cudaStreamDestroy(s->compute);
    if (s->h2d) cudaStreamDestroy(s->h2d);
    if (s->d2h) cudaStreamDestroy(s->d2h);
    delete s;
}

int lutr_stage_frame_bytes(const lutr_stage *s, size_t *in_bytes, size_t *out_bytes) {
    if (!s) return LUTR_ERR_INVALID_ARG;
    if (in_bytes) *in_bytes = s->in_bytes;
    if (out_bytes) *out_bytes = s->out_bytes;
    return LUTR_OK;
}

int lutr_stage_packed_planes(const lutr_stage *s, int32_t which, void *base, lutr_planes *planes) {
    if (!s || !planes || !base) return LUTR_ERR_INVALID_ARG;
    packed_planes(which ? s->d.out_fmt : s->d.in_fmt, s->d.width, s->d.height, (uint8_t *)base, planes);
    return LUTR_OK;
}

int lutr_stage_run(lutr_stage *s, const lutr_lut *lut, const lutr_planes *in, const lutr_planes *out, int32_t nframes,
                   void *cuda_stream) {
    if (!s || (!lut && !s->convert_only) || !in || !out || nframes <= 0) {
        set_error("lutr_stage_run: NULL argument or nframes <= 0");
        return LUTR_ERR_INVALID_ARG;
    }
    if (lut && lut->ctx->device != s->ctx->device) {
        set_error("lutr_stage_run: LUT lives on device %d, stage on device %d", lut->ctx->device, s->ctx->device);
        return LUTR_ERR_INVALID_ARG;
    }
    int rc = check_planes(s, in, s->d.in_fmt, nframes, "input");
    if (rc == LUTR_OK) rc = check_planes(s, out, s->d.out_fmt, nframes, "output");
    if (rc != LUTR_OK) return rc;
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    return enqueue(s, lut, in, out, nframes, cuda_stream ? (cudaStream_t)cuda_stream : s->compute);
}

int lutr_stage_run_host(lutr_stage *s, const lutr_lut *lut, const void *in_host, void *out_host, int32_t nframes) {
    if (!s || (!lut && !s->convert_only) || !in_host || !out_host || nframes <= 0) {
        set_error("lutr_stage_run_host: NULL argument or nframes <= 0");
        return LUTR_ERR_INVALID_ARG;
    }
    if (lut && lut->ctx->device != s->ctx->device) {
        set_error("lutr_stage_run_host: LUT and stage are on different devices");
        return LUTR_ERR_INVALID_ARG;
    }
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    // chunk = frames per pipeline slot: ~64 MB of input per slot keeps PCIe busy without hoarding HBM
    if (s->chunk == 0) {
        size_t per = s->in_bytes > s->out_bytes ? s->in_bytes : s->out_bytes;
        int chunk = (int)(((size_t)s->chunk_mb << 20) / per);
        if (chunk < 1) chunk = 1;
        if (chunk > 64) chunk = 64;
        for (int i = 0; i < lutr_stage::NBUF; ++i) {
            cudaError_t e = cudaMalloc((void **)&s->dev_in[i], (size_t)chunk * s->in_bytes);
            if (e == cudaSuccess) e = cudaMalloc((void **)&s->dev_out[i], (size_t)chunk * s->out_bytes);
            if (e != cudaSuccess) {
                set_error("stage: staging allocation failed: %s", cudaGetErrorString(e));
                free_host_pipeline(s);
                return LUTR_ERR_CUDA;
            }
        }
        s->chunk = chunk;
    }
    const uint8_t *src = (const uint8_t *)in_host;
    uint8_t *dst = (uint8_t *)out_host;
    int slot = 0;
    bool used[lutr_stage::NBUF] = {false, false, false};
    for (int f0 = 0; f0 < nframes; f0 += s->chunk, slot = (slot + 1) % lutr_stage::NBUF) {
        const int n = nframes - f0 < s->chunk ? nframes - f0 : s->chunk;
        // slot reuse: its previous D2H must have drained before we overwrite dev_in / dev_out
        if (used[slot]) {
            CUDA_TRY(cudaStreamWaitEvent(s->h2d, s->ev_out[slot], 0));
            CUDA_TRY(cudaStreamWaitEvent(s->compute, s->ev_out[slot], 0));
        }
        CUDA_TRY(cudaMemcpyAsync(s->dev_in[slot], src + (size_t)f0 * s->in_bytes, (size_t)n * s->in_bytes, cudaMemcpyHostToDevice, s->h2d));
        CUDA_TRY(cudaEventRecord(s->ev_in[slot], s->h2d));
        CUDA_TRY(cudaStreamWaitEvent(s->compute, s->ev_in[slot], 0));
        lutr_planes pin{}, pout{};
        packed_planes(s->d.in_fmt, s->d.width, s->d.height, s->dev_in[slot], &pin);
        packed_planes(s->d.out_fmt, s->d.width, s->d.height, s->dev_out[slot], &pout);
        int rc = enqueue(s, lut, &pin, &pout, n, s->compute);
        if (rc != LUTR_OK) return rc;
        CUDA_TRY(cudaEventRecord(s->ev_done[slot], s->compute));
        CUDA_TRY(cudaStreamWaitEvent(s->d2h, s->ev_done[slot], 0));
        CUDA_TRY(cudaMemcpyAsync(dst + (size_t)f0 * s->out_bytes, s->dev_out[slot], (size_t)n * s->out_bytes, cudaMemcpyDeviceToHost, s->d2h));
        CUDA_TRY(cudaEventRecord(s->ev_out[slot], s->d2h));
        used[slot] = true;
    }
    CUDA_TRY(cudaStreamSynchronize(s->d2h));
    CUDA_TRY(cudaStreamSynchronize(s->compute));
    return LUTR_OK;
}

int lutr_stage_sync(lutr_stage *s) {
    if (!s) return LUTR_ERR_INVALID_ARG;
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(s->compute));
    return LUTR_OK;
}

int64_t lutr_stage_launch_count(const lutr_stage *s) { return s ? s->launches.load() : 0; }

const char *lutr_stage_kernel_name(const lutr_stage *s) { return s ? s->last_kernel : ""; }

int lutr_stage_set_option(lutr_stage *s, const char *name, int32_t value) {
    if (!s || !name) return LUTR_ERR_INVALID_ARG;
    if (!std::strcmp(name, "rows_per_cta")) s->rows_per_cta = value;
    else if (!std::strcmp(name, "band_rows")) { /* accepted, ignored: band kernels are gone */ }
    else if (!std::strcmp(name, "fast_path")) s->use_fast = value;
    else if (!std::strcmp(name, "any_kernel")) s->force_any = value;
    else if (!std::strcmp(name, "grid_ctas")) s->grid_ctas = value;
    else if (!std::strcmp(name, "idx_mode")) { /* accepted, ignored (see include/lutr.h) */ }
    else if (!std::strcmp(name, "host_chunk_mb")) {
        if (value < 1 || value > 4096) {
            set_error("host_chunk_mb must be in [1, 4096]");
            return LUTR_ERR_INVALID_ARG;
        }
        if (s->chunk) free_host_pipeline(s);
        s->chunk_mb = value;
    } else if (!std::strcmp(name, "wk_pk")) s->wk_pk = value;
    else if (!std::strcmp(name, "wk_var")) s->wk_var = value;
    else if (!std::strcmp(name, "wk_tma")) s->wk_tma = value;
    else if (!std::strcmp(name, "wk_carveout")) s->wk_carveout = value;
    else if (!std::strcmp(name, "wk_unroll")) s->wk_unroll = value, s->wk_nk_forced = 1;
    else if (!std::strcmp(name, "wk_minb")) s->wk_minb = value >= 2 && value <= 4 ? value : 3, s->wk_minb_set = true;
    else if (!std::strcmp(name, "min_blocks")) { /* accepted, ignored */ }
    else if (!std::strcmp(name, "fast_threads") || !std::strcmp(name, "l2_persist")) {
        // accepted for older callers; the kernel families they tuned are gone (the LUT stays L2-resident by itself)
    } else {
        set_error("unknown option '%s'", name);
        return LUTR_ERR_INVALID_ARG;
    }
    return LUTR_OK;
}

int lutr_tables_yuv2rgb(int32_t matrix, int32_t full_range, int32_t out[7]) {
    Yuv2RgbK k;
    if (!out || !make_yuv2rgb(matrix, full_range != 0, k)) return LUTR_ERR_INVALID_ARG;
    out[0] = k.yC, out[1] = k.yO8, out[2] = k.vr, out[3] = k.ub, out[4] = k.ug, out[5] = k.vg, out[6] = k.yOff10;
    return LUTR_OK;
}

int lutr_tables_rgb2yuv(int32_t matrix, int32_t out[9]) {
    Rgb2YuvK k;
    if (!out || !make_rgb2yuv(matrix, k)) return LUTR_ERR_INVALID_ARG;
    out[0] = k.ry, out[1] = k.gy, out[2] = k.by, out[3] = k.ru, out[4] = k.gu, out[5] = k.bu, out[6] = k.rv, out[7] = k.gv, out[8] = k.bv;
    return LUTR_OK;
}

}  // extern "C"

// ---- raw-pipe pump: decoder process -> fd_in -> [pinned ring -> H2D -> kernel -> D2H -> pinned ring] -> fd_out -> encoder process --------
// Replaces the blocking `Popen(ffmpeg ... -vf lut3d ...)` of the reference's TaskRunner._run_stage (reference
// src/lut_renderer/task_manager.py:134-190) for a LUT stage.  Three threads: a reader fills pinned input slots straight from the
// pipe, the calling thread enqueues copy-in / kernel / copy-out on the stage's three streams (fully asynchronous), a writer waits for
// a slot's copy-out event and drains it into the pipe.  No interpreter lock, no intermediate buffers, 1 MB+ pipe buffers.
namespace {

struct PumpQueue {  // tiny blocking queue of slot indices (-1 = end of stream)
    std::mutex m;
    std::condition_variable cv;
    std::deque<int> q;
    void push(int v) {
        { std::lock_guard<std::mutex> g(m); q.push_back(v); }
        cv.notify_one();
    }
    int pop() {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [&] { return !q.empty(); });
        const int v = q.front();
        q.pop_front();
        return v;
    }
};

struct PumpSlot {
    uint8_t *hin = nullptr, *hout = nullptr, *din = nullptr, *dout = nullptr;
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    int frames = 0;
    bool last = false;
};

void grow_pipe(int fd) {
    // F_SETPIPE_SZ: as large as the kernel allows (the default is 64 KB, a 4K 10-bit frame is 33 MB): fewer wake-ups on both sides
    // (measured on the B200 hosts: 1 / 16 / 64 MB all move 2.0 GB/s each way, 4K 10-bit; a bare generator -> sink pipe does 3.3 GB/s)
    int top = 64;
    if (const char *e = std::getenv("LUTR_PIPE_MB")) top = std::atoi(e) > 0 ? std::atoi(e) : top;  // measurement knob
    for (int mb = top; mb >= 1; mb /= 2)
        if (fcntl(fd, F_SETPIPE_SZ, mb << 20) >= 0) return;
}

bool wait_fd(int fd, short events, const volatile int32_t *cancel, const std::atomic<bool> &stop) {
    for (;;) {
        if (stop.load(std::memory_order_relaxed) || (cancel && *cancel)) return false;
        struct pollfd p = {fd, events, 0};
        const int r = poll(&p, 1, 100);
        if (r > 0) return true;  // readable / writable, hang-up or error: the read / write call reports which
        if (r < 0 && errno != EINTR) return true;
    }
}

}  // namespace

extern "C" int lutr_stage_pump(lutr_stage *s, const lutr_lut *lut, int fd_in, int fd_out, int32_t frames_per_batch,
                               const volatile int32_t *cancel, lutr_progress_fn progress, void *user, lutr_pump_stats *stats) {
    if (!s || (!lut && !s->convert_only) || fd_in < 0 || fd_out < 0) {
        set_error("lutr_stage_pump: NULL stage / LUT or bad file descriptor");
        return LUTR_ERR_INVALID_ARG;
    }
    if (lut && lut->ctx->device != s->ctx->device) {
        set_error("lutr_stage_pump: LUT and stage are on different devices");
        return LUTR_ERR_INVALID_ARG;
    }
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    constexpr int NS = 5;  // reader, copy-in / kernel / copy-out, writer, and the batch whose pages the pipe still references
    const size_t per = s->in_bytes > s->out_bytes ? s->in_bytes : s->out_bytes;
    int batch = frames_per_batch > 0 ? frames_per_batch : (int)(((size_t)64 << 20) / per);
    batch = batch < 1 ? 1 : (batch > 64 ? 64 : batch);
    PumpSlot slot[NS];
    std::atomic<bool> stop{false};
    std::string rerr, werr;  // error text of the reader / writer thread (set_error is thread-local)
    int rc = LUTR_OK;
    auto release = [&]() {
        for (PumpSlot &t : slot) {
            if (t.hin) cudaFreeHost(t.hin);
            if (t.hout) cudaFreeHost(t.hout);
            if (t.din) cudaFree(t.din);
            if (t.dout) cudaFree(t.dout);
            if (t.ev_in) cudaEventDestroy(t.ev_in);
            if (t.ev_k) cudaEventDestroy(t.ev_k);
            if (t.ev_out) cudaEventDestroy(t.ev_out);
        }
    };
    for (PumpSlot &t : slot) {
        cudaError_t e = cudaHostAlloc((void **)&t.hin, (size_t)batch * s->in_bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&t.hout, (size_t)batch * s->out_bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc((void **)&t.din, (size_t)batch * s->in_bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&t.dout, (size_t)batch * s->out_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t.ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t.ev_k, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t.ev_out, cudaEventDisableTiming | cudaEventBlockingSync);
        if (e != cudaSuccess) {
            set_error("lutr_stage_pump: ring allocation failed: %s", cudaGetErrorString(e));
            release();
            return LUTR_ERR_CUDA;
        }
    }
    grow_pipe(fd_in);
    grow_pipe(fd_out);
    PumpQueue free_q, filled_q, out_q;
    for (int i = 0; i < NS; ++i) free_q.push(i);
    int64_t frames_in = 0, frames_out = 0;
    const auto t0 = std::chrono::steady_clock::now();

    std::thread reader([&]() {
        for (;;) {
            const int i = free_q.pop();
            if (i < 0) break;
            PumpSlot &t = slot[i];
            const size_t want = (size_t)batch * s->in_bytes;
            size_t got = 0;
            bool eof = false;
            while (got < want) {
                if (!wait_fd(fd_in, POLLIN, cancel, stop)) { eof = true; break; }
                const ssize_t n = read(fd_in, t.hin + got, want - got);
                if (n > 0) got += (size_t)n;
                else if (n == 0) { eof = true; break; }
                else if (errno != EINTR && errno != EAGAIN) {
                    rerr = std::string("read from the decoder pipe failed: ") + std::strerror(errno);
                    eof = true;
                    break;
                }
            }
            t.frames = (int)(got / s->in_bytes);  // a trailing partial frame is dropped, as ffmpeg's rawvideo demuxer does
            t.last = eof;
            filled_q.push(i);
            if (eof) break;
        }
    });
    std::thread writer([&]() {
        cudaSetDevice(s->ctx->device);
        sigset_t pipe_set, old_set;
        sigemptyset(&pipe_set);
        sigaddset(&pipe_set, SIGPIPE);
        pthread_sigmask(SIG_BLOCK, &pipe_set, &old_set);  // a vanished encoder is an EPIPE error here, not a fatal signal
        bool broken = false;
        // Output side without a copy: vmsplice() lends the pinned pages of a batch to the pipe instead of copying them into it
        // (+14 % on the B200 hosts).  The pipe then REFERENCES those pages until the encoder has read them, so a slot must not be
        // refilled while any of its pages is still queued: the writer keeps the slot it queued last (`held`) and releases it only
        // when no more than the bytes queued after it remain unread (FIONREAD) -- with batches larger than the pipe that is already
        // true when the next batch has gone in.  Anything that is not a pipe (EBADF / EINVAL on the first call) falls back to write().
        bool splice = std::getenv("LUTR_PUMP_NO_SPLICE") == nullptr, spliced_any = false;
        int held = -1;
        auto unread_at_most = [&](size_t limit) {  // false: gave up (cancelled, stopped, encoder gone)
            for (;;) {
                int n = 0;
                if (ioctl(fd_out, FIONREAD, &n) < 0 || (size_t)n <= limit) return true;
                if (stop.load(std::memory_order_relaxed) || (cancel && *cancel)) return false;
                struct pollfd p = {fd_out, 0, 0};
                if (poll(&p, 1, 0) > 0 && (p.revents & (POLLERR | POLLHUP | POLLNVAL))) {  // nobody reads the pipe any more
                    if (werr.empty()) werr = "write to the encoder pipe failed: the encoder closed its end";
                    stop.store(true);
                    return false;
                }
                struct timespec nap = {0, 200000};
                nanosleep(&nap, nullptr);
            }
        };
        for (;;) {
            const int i = out_q.pop();
            if (i < 0) break;
            PumpSlot &t = slot[i];
            if (cudaEventSynchronize(t.ev_out) != cudaSuccess && werr.empty()) werr = "copy-out of a batch failed", broken = true;
            const size_t want = (size_t)t.frames * s->out_bytes;
            size_t put = 0;
            while (!broken && put < want) {
                if (!wait_fd(fd_out, POLLOUT, cancel, stop)) { broken = true; break; }
                ssize_t n;
                if (splice) {
                    struct iovec iov = {t.hout + put, want - put};
                    n = vmsplice(fd_out, &iov, 1, SPLICE_F_NONBLOCK);
                    if (n < 0 && !spliced_any && (errno == EBADF || errno == EINVAL || errno == ENOSYS)) {
                        splice = false;  // not a pipe
                        continue;
                    }
                    if (n > 0) spliced_any = true;
                } else {
                    n = write(fd_out, t.hout + put, want - put);
                }
                if (n > 0) put += (size_t)n;
                else if (n < 0 && errno != EINTR && errno != EAGAIN) {
                    werr = std::string("write to the encoder pipe failed: ") + std::strerror(errno);
                    if (errno == EPIPE) {
                        struct timespec zero = {0, 0};
                        sigtimedwait(&pipe_set, nullptr, &zero);  // consume the pending SIGPIPE before the mask is restored
                    }
                    broken = true;
                    stop.store(true);
                }
            }
            if (!broken) frames_out += t.frames;
            if (spliced_any && !broken) {
                if (held >= 0) {
                    if (!unread_at_most(want)) broken = true;
                    free_q.push(held);
                }
                held = i;
            } else {
                free_q.push(i);
            }
        }
        if (held >= 0) {
            if (!broken) unread_at_most(0);  // the last batch: its pages go back to the allocator when this call returns
            free_q.push(held);
        }
        pthread_sigmask(SIG_SETMASK, &old_set, nullptr);
    });

    // dispatcher (this thread)
    for (;;) {
        const int i = filled_q.pop();
        PumpSlot &t = slot[i];
        const bool last = t.last;
        if (t.frames > 0 && rc == LUTR_OK && !stop.load() && !(cancel && *cancel)) {
            lutr_planes pin{}, pout{};
            packed_planes(s->d.in_fmt, s->d.width, s->d.height, t.din, &pin);
            packed_planes(s->d.out_fmt, s->d.width, s->d.height, t.dout, &pout);
            cudaError_t e = cudaMemcpyAsync(t.din, t.hin, (size_t)t.frames * s->in_bytes, cudaMemcpyHostToDevice, s->h2d);
            if (e == cudaSuccess) e = cudaEventRecord(t.ev_in, s->h2d);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s->compute, t.ev_in, 0);
            if (e == cudaSuccess) {
                rc = enqueue(s, lut, &pin, &pout, t.frames, s->compute);
                if (rc == LUTR_OK) {
                    e = cudaEventRecord(t.ev_k, s->compute);
                    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->d2h, t.ev_k, 0);
                    if (e == cudaSuccess) e = cudaMemcpyAsync(t.hout, t.dout, (size_t)t.frames * s->out_bytes, cudaMemcpyDeviceToHost, s->d2h);
                    if (e == cudaSuccess) e = cudaEventRecord(t.ev_out, s->d2h);
                }
            }
            if (e != cudaSuccess) {
                set_error("lutr_stage_pump: CUDA error %s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
                rc = LUTR_ERR_CUDA;
            }
            if (rc == LUTR_OK) {
                frames_in += t.frames;
                out_q.push(i);
                if (progress) progress(user, frames_in);
            } else {
                stop.store(true);
                free_q.push(i);
            }
        } else {
            free_q.push(i);
        }
        if (last) break;
    }
    out_q.push(-1);
    writer.join();
    stop.store(true);
    free_q.push(-1);  // a reader still waiting for a slot
    reader.join();
    cudaStreamSynchronize(s->d2h);
    cudaStreamSynchronize(s->compute);
    release();
    if (stats) {
        stats->frames_in = frames_in, stats->frames_out = frames_out;
        stats->bytes_in = frames_in * (int64_t)s->in_bytes, stats->bytes_out = frames_out * (int64_t)s->out_bytes;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (rc != LUTR_OK) return rc;
    if (cancel && *cancel) {
        set_error("lutr_stage_pump: cancelled");
        return LUTR_ERR_CANCELLED;
    }
    if (!werr.empty() || !rerr.empty()) {
        set_error("lutr_stage_pump: %s", (!werr.empty() ? werr : rerr).c_str());
        return LUTR_ERR_IO;
    }
    return LUTR_OK;
}
