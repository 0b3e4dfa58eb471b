/*
 * lutr.h -- C-ABI of the B200-native 3D-LUT colour stage (drop-in for the per-pixel work LUT-renderer
 * delegates to FFmpeg).
 *
 * The reference has no FFI: its boundary for this path is the `-vf` sub-chain that
 * build_command() assembles (reference src/lut_renderer/ffmpeg.py:195-247, 304-310) and that
 * TaskRunner._run_stage() hands to an `ffmpeg` subprocess (task_manager.py:145-151):
 *
 *     [scale=in_range=pc:out_range=R[:in_color_matrix=M:out_color_matrix=M],format=yuv4xxp,]   ffmpeg.py:212-233
 *     [scale=in_color_matrix=M:out_color_matrix=M,]                                            ffmpeg.py:234-236
 *     lut3d=file='X.cube':interp=I                                                             ffmpeg.py:242-247
 *     [,format=PIXFMT]   -pix_fmt PIXFMT                                                        ffmpeg.py:304-310
 *
 * Each entry point below replaces one node of that chain (cited per function).  Plain pointers and
 * sizes only; no torch / C++ types.  All functions returning int return LUTR_OK (0) or a negative
 * lutr_status; the message for the last failure on the calling thread is lutr_last_error().
 *
 * Threading: a lutr_ctx and a lutr_lut may be shared by threads that run different stages on
 * different streams; a lutr_stage is single-threaded.  The caller owns frame buffers; the library
 * owns LUT device memory and staging buffers.  There is NO CPU fallback: every compute entry point
 * fails with LUTR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef LUTR_H
#define LUTR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LUTR_API __declspec(dllexport)
#else
#define LUTR_API __attribute__((visibility("default")))
#endif

#define LUTR_VERSION 100 /* 0.1.0, tracks reference src/lut_renderer/__init__.py:2 */

typedef enum lutr_status {
    LUTR_OK = 0,
    LUTR_ERR_INVALID_ARG = -1,  /* NULL handle, bad enum, bad geometry                                   */
    LUTR_ERR_INVALID_DATA = -2, /* .cube: garbled data line / unexpected EOF / empty LUT  (lut3d: AVERROR_INVALIDDATA) */
    LUTR_ERR_LUT_SIZE = -3,     /* .cube: LUT_3D_SIZE < 2 or > 256  (lut3d: "Too large or invalid 3D LUT size", EINVAL) */
    LUTR_ERR_IO = -4,           /* .cube file cannot be opened / read                                    */
    LUTR_ERR_NOMEM = -5,
    LUTR_ERR_CUDA = -6,         /* no device, wrong architecture, or a CUDA runtime error                */
    LUTR_ERR_UNSUPPORTED = -7,  /* valid request outside what the kernels implement      */
    LUTR_ERR_CANCELLED = -8     /* lutr_stage_pump: the caller's cancel flag was raised (task_manager.py:38-44: "canceled") */
} lutr_status;

/* Pixel formats of the stage input / output (`format=` / `-pix_fmt`, ffmpeg.py:233, 287-310). */
typedef enum lutr_pix_fmt {
    LUTR_PIX_YUV420P = 0,     /* 8-bit 4:2:0 planar  (fast-delivery default / force_8bit)                 */
    LUTR_PIX_YUV422P10LE = 1, /* 10-bit 4:2:2 planar in LE u16 (pro-mode ProRes master, ffmpeg.py:417-433) */
    LUTR_PIX_YUV420P10LE = 2, /* 10-bit 4:2:0 planar in LE u16 (10-bit camera / HEVC sources; libx265 'preserve' output, ffmpeg.py:297-298) */
    LUTR_PIX_YUV422P = 3,     /* 8-bit 4:2:2 planar (GUI pix_fmt choice main_window.py:767-771; range-pass intermediate ffmpeg.py:137-143) */
    LUTR_PIX_YUV444P = 4,     /* 8-bit 4:4:4 planar (same two uses) */
    LUTR_PIX_YUV444P10LE = 5  /* 10-bit 4:4:4 planar in LE u16 (10-bit 4:4:4 sources) */
} lutr_pix_fmt;

typedef enum lutr_range { LUTR_RANGE_TV = 0, LUTR_RANGE_PC = 1 } lutr_range;

/* YUV<->RGB matrices (`in_color_matrix` / `out_color_matrix`, ffmpeg.py:113-126, 201-209). */
typedef enum lutr_matrix {
    LUTR_MATRIX_BT709 = 0,
    LUTR_MATRIX_BT601 = 1,    /* smpte170m / bt470bg / FFmpeg's default when nothing is forced ("none") */
    LUTR_MATRIX_BT2020 = 2,   /* bt2020nc / bt2020c */
    LUTR_MATRIX_FCC = 3,
    LUTR_MATRIX_SMPTE240M = 4
} lutr_matrix;

/* lut3d `interp=` (ffmpeg.py:242-246; GUI offers tetrahedral / trilinear, main_window.py:649-652). */
typedef enum lutr_interp {
    LUTR_INTERP_NEAREST = 0,
    LUTR_INTERP_TRILINEAR = 1,
    LUTR_INTERP_TETRAHEDRAL = 2,
    LUTR_INTERP_PYRAMID = 3, /* accepted by the reference's whitelist (ffmpeg.py:243); generic kernel only */
    LUTR_INTERP_PRISM = 4,
    LUTR_INTERP_NONE = 5     /* no lut3d at all: pro mode stage 2 (ffmpeg.py:459-471), the encoding ffmpeg's own yuv422p10le ->
                                yuv420p | yuv420p10le | yuv422p conversion (auto-inserted swscale pass).  The `lut` argument of the
                                run / pump calls is ignored (may be NULL); matrices and ranges must be equal on both sides. */
} lutr_interp;

/* Float evaluation order of the LUT blend.  C = FFmpeg's C code (separate mul and add; what packed
 * rgb24 always uses and the documented semantics of this library).  FMA = contracted multiply-adds
 * in the same operand order (what FFmpeg's x86 AVX2 assembly does for planar formats). */
typedef enum lutr_float_mode { LUTR_FLOAT_C = 0, LUTR_FLOAT_FMA = 1 } lutr_float_mode;

/* One colour stage = everything between the decoder's raw frame and the encoder's raw frame. */
typedef struct lutr_stage_desc {
    uint32_t struct_size; /* sizeof(lutr_stage_desc), for ABI growth */
    int32_t width, height;
    int32_t in_fmt;       /* lutr_pix_fmt of the decoded source frames                                      */
    int32_t out_fmt;      /* lutr_pix_fmt after `format=`                                                    */
    int32_t in_range;     /* lutr_range of the source samples (VideoInfo.color_range / yuvj*, ffmpeg.py:129-134) */
    int32_t norm_range;   /* range after the reference's `scale=in_range=pc:out_range=R` pass (ffmpeg.py:213-225);
                             == in_range when the source is tv.  pc -> tv re-quantises to 8 bit like format=yuv4xxp */
    int32_t in_matrix;    /* lutr_matrix for YUV->RGB in front of lut3d                                      */
    int32_t out_matrix;   /* lutr_matrix for RGB->YUV behind lut3d (output tag policy, ffmpeg.py:348-383)    */
    int32_t out_range;    /* lutr_range of the output (LUTR_RANGE_TV for tags=bt709)                          */
    int32_t interp;       /* lutr_interp                                                                     */
    int32_t float_mode;   /* lutr_float_mode, 0 = C semantics                                                */
    int32_t via_8bit;     /* full-range (in_range = pc) source deeper than 8 bit: 1 = go through the 8-bit yuv4xxp intermediate the
                             reference's `scale=in_range=pc:out_range=R,format=yuv4xxp` produces (ffmpeg.py:211-233; one swscale pass:
                             range conversion + ordered dither for R = tv, dithered planar copy for R = pc), as plan_colour_stage
                             asks for.  0 with norm_range = pc: the samples are used at their own depth (beyond the reference). */
    int32_t reserved[3];
} lutr_stage_desc;

/* A batch of planar frames: plane p of frame f starts at data[p] + f * frame_step, rows `stride[p]` bytes apart. */
typedef struct lutr_planes {
    void *data[3];      /* Y, U, V */
    int32_t stride[3];  /* bytes per row */
    int64_t frame_step; /* bytes between consecutive frames (ignored when nframes == 1) */
} lutr_planes;

typedef struct lutr_ctx lutr_ctx;
typedef struct lutr_lut lutr_lut;
typedef struct lutr_stage lutr_stage;

/* ---- library ------------------------------------------------------------------------------------- */
LUTR_API int lutr_version(void);
LUTR_API const char *lutr_status_string(int status);
LUTR_API const char *lutr_last_error(void); /* thread-local; replaces ffmpeg's stderr text (task_manager.py:105-112) */

/* ---- .cube parsing: replaces vf_lut3d's parse_cube() behind `lut3d=file=...` (ffmpeg.py:246) ---------
 * Host only (usable without a GPU).  On success *table is a malloc'd float[size^3*3] in lut3d's memory
 * order [r][g][b] (release with lutr_free), scale[c] = clip(1/(DOMAIN_MAX[c]-DOMAIN_MIN[c]), 0, 1). */
LUTR_API int lutr_cube_parse(const char *text, size_t len, int32_t *size, float scale[3], float **table);
LUTR_API int lutr_cube_parse_file(const char *path, int32_t *size, float scale[3], float **table);
/* The other file dialects lut3d reads, chosen by the file extension as lut3d's init does (case-insensitive: .cube .dat .3dl .m3d
 * .csp; no / unknown extension -> LUTR_ERR_INVALID_ARG).  The reference's GUI only offers *.cube (lut_manager.py:121); a task built
 * through the model can carry any of them to `lut3d=file=...` (ffmpeg.py:246).  Restated from memory of vf_lut3d.c, unpinned
 * (csrc/lut_dialects.cpp has the rules); a .csp with a 1D shaper in front of the cube is LUTR_ERR_UNSUPPORTED. */
typedef enum lutr_lut_dialect { LUTR_LUT_CUBE = 0, LUTR_LUT_DAT = 1, LUTR_LUT_3DL = 2, LUTR_LUT_M3D = 3, LUTR_LUT_CSP = 4 } lutr_lut_dialect;
LUTR_API int lutr_lut_parse_text(const char *text, size_t len, int32_t dialect, int32_t *size, float scale[3], float **table);
LUTR_API int lutr_lut_parse_file(const char *path, int32_t *size, float scale[3], float **table);
LUTR_API void lutr_free(void *p);

/* ---- device context ------------------------------------------------------------------------------ */
LUTR_API int lutr_device_count(int32_t *count);
LUTR_API int lutr_ctx_create(int32_t device, lutr_ctx **out);
LUTR_API void lutr_ctx_destroy(lutr_ctx *ctx);
LUTR_API int lutr_ctx_info(const lutr_ctx *ctx, int32_t *device, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);
/* pinned host memory for the raw-pipe ring (decode -> GPU -> encode re-plumb of task_manager.py:134-190) */
LUTR_API int lutr_host_alloc(lutr_ctx *ctx, size_t bytes, void **out);
LUTR_API void lutr_host_free(lutr_ctx *ctx, void *p);

/* ---- LUT upload: once per device and per distinct LUT --------------------------------------------- */
LUTR_API int lutr_lut_load_cube(lutr_ctx *ctx, const char *path, lutr_lut **out);
LUTR_API int lutr_lut_load_cube_memory(lutr_ctx *ctx, const char *text, size_t len, lutr_lut **out);
LUTR_API int lutr_lut_load_file(lutr_ctx *ctx, const char *path, lutr_lut **out); /* any dialect, by extension (see lutr_lut_parse_file) */
LUTR_API int lutr_lut_from_table(lutr_ctx *ctx, int32_t size, const float *rgb, const float scale[3], lutr_lut **out);
LUTR_API int lutr_lut_size(const lutr_lut *lut);
LUTR_API void lutr_lut_destroy(lutr_lut *lut);

/* ---- the stage: replaces scale/format/lut3d/format between decode and encode ----------------------- */
LUTR_API int lutr_stage_create(lutr_ctx *ctx, const lutr_stage_desc *desc, lutr_stage **out);
LUTR_API void lutr_stage_destroy(lutr_stage *stage);
/* bytes of one tightly packed frame (Y plane, then U, then V; rows unpadded) on each side */
LUTR_API int lutr_stage_frame_bytes(const lutr_stage *stage, size_t *in_bytes, size_t *out_bytes);
/* fill a lutr_planes for `base` holding tightly packed frames (input side: which=0, output side: which=1) */
LUTR_API int lutr_stage_packed_planes(const lutr_stage *stage, int32_t which, void *base, lutr_planes *planes);

/* Device-resident frames: enqueue the fused kernel for `nframes` frames on `cuda_stream`
 * (a cudaStream_t passed as void*, NULL = the stage's own stream).  Asynchronous. */
LUTR_API int lutr_stage_run(lutr_stage *stage, const lutr_lut *lut, const lutr_planes *in, const lutr_planes *out,
                            int32_t nframes, void *cuda_stream);
/* Host frames (tightly packed, pinned or pageable): H2D -> kernel -> D2H, chunked and overlapped on the
 * stage's copy/compute streams.  Synchronous: `out_host` is complete on return. */
LUTR_API int lutr_stage_run_host(lutr_stage *stage, const lutr_lut *lut, const void *in_host, void *out_host,
                                 int32_t nframes);
LUTR_API int lutr_stage_sync(lutr_stage *stage);
/* number of kernel launches this stage has enqueued so far (bench.py's gpu_launches) */
LUTR_API int64_t lutr_stage_launch_count(const lutr_stage *stage);
/* name of the kernel family the most recent enqueue of this stage launched ("" before the first one); diagnostics */
LUTR_API const char *lutr_stage_kernel_name(const lutr_stage *stage);
/* Raw-pipe pump: the whole data path of one LUT stage of the reference's runner (task_manager.py:134-190 blocks on ONE ffmpeg
 * whose filter graph holds lut3d; the re-plumbed runner starts a decoding and an encoding ffmpeg and calls this between them).
 * Reads tightly packed frames of the stage's input layout from `fd_in` until end of file, runs them through the stage, writes the
 * output frames to `fd_out` in order.  Pinned host ring, H2D / kernel / D2H on the stage's streams, a reader and a writer thread
 * inside the call; the pipes are grown to the kernel's maximum (F_SETPIPE_SZ).  `frames_per_batch` <= 0 picks ~64 MB batches.
 * `cancel` (may be NULL) is polled at least every 100 ms: non-zero ends the call with LUTR_ERR_CANCELLED.  `progress` (may be NULL)
 * is called on the calling thread after each batch with the number of frames handed to the GPU so far.  A trailing partial frame is
 * dropped.  Returns LUTR_ERR_IO when a pipe fails (e.g. the encoder went away: EPIPE) -- the text is in lutr_last_error(). */
typedef struct lutr_pump_stats {
    int64_t frames_in, frames_out; /* frames read and processed / frames written */
    int64_t bytes_in, bytes_out;
    double seconds;
} lutr_pump_stats;
typedef void (*lutr_progress_fn)(void *user, int64_t frames_done);
LUTR_API int lutr_stage_pump(lutr_stage *stage, const lutr_lut *lut, int fd_in, int fd_out, int32_t frames_per_batch,
                             const volatile int32_t *cancel, lutr_progress_fn progress, void *user, lutr_pump_stats *stats);

/* run-time tunables, for benchmarking and testing: "grid_ctas" (CTAs launched by the warp-window kernels, 0 = the library's choice),
 * "rows_per_cta" (rows of one work unit of the 4:2:2 10-bit kernels), "host_chunk_mb" (input megabytes per slot of the host pipeline's
 * 3-slot device ring, default 64), "any_kernel" (1: route every launch through the generic tile kernel), "fast_path" (0: the same,
 * older name), "wk_minb" (CTAs per SM the warp-window kernels are built for: 2..4), "wk_unroll" (4:2:0 8-bit kernel: window width in
 * 64-column steps), "wk_tma" (4:2:2 10-bit stage: 1 = row-staging kernel with a cp.async ring, 2 = with TMA bulk copies; default 0 =
 * register prefetch, the fastest of the three).  "idx_mode", "min_blocks", "fast_threads", "l2_persist" are accepted and ignored
 * (tunables of kernel generations that are no longer in the library). */
LUTR_API int lutr_stage_set_option(lutr_stage *stage, const char *name, int32_t value);

/* ---- coefficient tables (exposed so the parity tests can pin them against libswscale) --------------- */
/* yuv2rgb: out[7] = yC, yO8, vr, ub, ug, vg, yOff10;   rgb2yuv: out[9] = ry gy by ru gu bu rv gv bv */
LUTR_API int lutr_tables_yuv2rgb(int32_t matrix, int32_t full_range, int32_t out[7]);
LUTR_API int lutr_tables_rgb2yuv(int32_t matrix, int32_t out[9]);

#ifdef __cplusplus
}
#endif
#endif /* LUTR_H */
