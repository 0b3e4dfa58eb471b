// The other 3D LUT file dialects FFmpeg's lut3d filter reads besides .cube: .dat (DaVinci), .3dl (Lustre / Flame), .m3d (Pandora),
// .csp (cineSpace).  The reference passes whatever path the task carries to `lut3d=file='...'` (src/lut_renderer/ffmpeg.py:246) and
// lut3d picks the reader from the file extension; the reference's GUI only offers *.cube (lut_manager.py:121, main_window.py:2324),
// so these are reachable through the task model / API only.
//
// STATUS: restated from memory of libavfilter/vf_lut3d.c (parse_dat, parse_3dl, parse_m3d, parse_cinespace, and the extension
// dispatch in its init), NOT pinned against a real FFmpeg -- this image has no libavfilter.  tests/test_lut_dialects.py checks the
// readers against an independent restatement (oracle/lut_dialects.py) and documents the rules below, so that a maintainer with an
// ffmpeg can confirm or correct them.  Rules, as upstream has them:
//   * lines come from a 512-byte fgets(); a line is skipped when it is blank or its first non-blank character is '#';
//     running out of lines while entries are missing is "Unexpected EOF" -> invalid data;
//   * .dat   optional first line "3DLUTSIZE n" (default 33), then n^3 lines "%f %f %f", file order = table order [r][g][b]
//            (blue fastest); the first data line is the line that followed the size directive (or the first line itself);
//   * .3dl   17 points, the first non-skipped line (the shaper / mesh line) is dropped, then 17^3 lines "%d %d %d" divided by
//            4096, blue fastest;
//   * .m3d   header lines "in N", "out M", "values r g b" (channel order of the columns; reading stops at that line); size = the
//            smallest s with s^3 >= N; values scaled by 1 / (M - 1); blue fastest; data lines are NOT filtered for comments;
//   * .csp   "CSPLUTV100", "3D", optional "BEGIN METADATA" .. "END METADATA"; per channel a pre-LUT block (point count, inputs,
//            outputs), then "nr ng nb" (all equal), then n^3 lines, RED fastest, each channel multiplied by (out_max - out_min).
//            Two-point blocks are a plain input / output range: scale = clip(1 / (in_max - in_min), 0, 1).  Blocks with more
//            than two points are a 1D shaper in front of the cube: not implemented here (LUTR_ERR_UNSUPPORTED).
// Non-finite entries are rejected as in the .cube reader (the kernels rely on 0 * entry == 0).
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "lutr_internal.h"

namespace lutr {

namespace {

// the pieces a 512-byte fgets() delivers
class Lines {
  public:
    Lines(const char *text, size_t len) : p_(text), end_(text + len) {}
    bool next(char (&buf)[512]) {
        if (p_ >= end_) return false;
        size_t n = 0;
        while (p_ < end_ && n < 511) {
            const char ch = *p_++;
            buf[n++] = ch;
            if (ch == '\n') break;
        }
        buf[n] = 0;
        return true;
    }
    // NEXT_LINE(skip_line(line)): read until a line that is neither blank nor a comment
    bool next_content(char (&buf)[512]) {
        do {
            if (!next(buf)) return false;
        } while (skip(buf));
        return true;
    }
    static bool skip(const char *p) {
        while (*p && std::isspace((unsigned char)*p)) ++p;
        return !*p || *p == '#';
    }

  private:
    const char *p_, *end_;
};

int eof_error(const char *what, size_t have, size_t want) {
    set_error("%s: Unexpected EOF after %zu of %zu entries", what, have, want);
    return LUTR_ERR_INVALID_DATA;
}

bool alloc_table(const char *what, long size, CubeLut &out) {
    if (size < 2 || size > 256) {
        set_error("%s: Too large or invalid 3D LUT size (%ld)", what, size);
        return false;
    }
    out.size = (int)size;
    out.rgb.assign((size_t)size * size * size * 3, 0.f);
    return true;
}

bool finite3(const float v[3]) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

int parse_dat(const char *text, size_t len, CubeLut &out) {
    Lines rd(text, len);
    char line[512];
    long size = 33;
    if (!rd.next_content(line)) return eof_error("dat", 0, 33 * 33 * 33);
    if (!std::strncmp(line, "3DLUTSIZE ", 10)) {
        size = std::strtol(line + 10, nullptr, 0);
        if (!rd.next_content(line)) return eof_error("dat", 0, 0);
    }
    if (!alloc_table("dat", size, out)) return LUTR_ERR_LUT_SIZE;
    const size_t total = (size_t)size * size * size;
    for (size_t e = 0; e < total; ++e) {  // table order [r][g][b] is the file order
        if (e && !rd.next_content(line)) return eof_error("dat", e, total);
        float v[3];
        if (std::sscanf(line, "%f %f %f", &v[0], &v[1], &v[2]) != 3 || !finite3(v)) {
            set_error("dat: invalid data line at entry %zu", e);
            return LUTR_ERR_INVALID_DATA;
        }
        std::memcpy(&out.rgb[3 * e], v, sizeof(v));
    }
    return LUTR_OK;
}

int parse_3dl(const char *text, size_t len, CubeLut &out) {
    Lines rd(text, len);
    char line[512];
    const size_t total = 17 * 17 * 17;
    if (!alloc_table("3dl", 17, out)) return LUTR_ERR_LUT_SIZE;
    if (!rd.next_content(line)) return eof_error("3dl", 0, total);  // the shaper line: dropped
    for (size_t e = 0; e < total; ++e) {
        if (!rd.next_content(line)) return eof_error("3dl", e, total);
        int r, g, b;
        if (std::sscanf(line, "%d %d %d", &r, &g, &b) != 3) {
            set_error("3dl: invalid data line at entry %zu", e);
            return LUTR_ERR_INVALID_DATA;
        }
        out.rgb[3 * e] = (float)r / 4096.f, out.rgb[3 * e + 1] = (float)g / 4096.f, out.rgb[3 * e + 2] = (float)b / 4096.f;
    }
    return LUTR_OK;
}

int parse_m3d(const char *text, size_t len, CubeLut &out) {
    Lines rd(text, len);
    char line[512];
    long in = -1, outv = -1;
    int map[3] = {0, 1, 2};
    while (rd.next(line)) {
        if (!std::strncmp(line, "in", 2)) in = std::strtol(line + 2, nullptr, 0);
        else if (!std::strncmp(line, "out", 3)) outv = std::strtol(line + 3, nullptr, 0);
        else if (!std::strncmp(line, "values", 6)) {
            const char *p = line + 6;
            for (int id = 0; id < 3; ++id) {
                while (std::isspace((unsigned char)*p)) ++p;
                if (*p == 'r') map[id] = 0;
                else if (*p == 'g') map[id] = 1;
                else if (*p == 'b') map[id] = 2;
                while (*p && !std::isspace((unsigned char)*p)) ++p;
            }
            break;
        }
    }
    if (in == -1 || outv == -1) {
        set_error("m3d: in and out must be defined");
        return LUTR_ERR_INVALID_DATA;
    }
    const long lim = 256l * 256 * 256;
    if (in < 2 || outv < 2 || in > lim || outv > lim) {
        set_error("m3d: invalid in (%ld) or out (%ld)", in, outv);
        return LUTR_ERR_INVALID_DATA;
    }
    long size = 1;
    while (size * size * size < in) ++size;
    if (!alloc_table("m3d", size, out)) return LUTR_ERR_LUT_SIZE;
    const float scale = (float)(1. / (double)(outv - 1));
    const size_t total = (size_t)size * size * size;
    for (size_t e = 0; e < total; ++e) {
        if (!rd.next(line)) return eof_error("m3d", e, total);  // NEXT_LINE(0): comments are not skipped here
        float val[3];
        if (std::sscanf(line, "%f %f %f", &val[0], &val[1], &val[2]) != 3) {
            set_error("m3d: invalid data line at entry %zu", e);
            return LUTR_ERR_INVALID_DATA;
        }
        const float v[3] = {val[map[0]] * scale, val[map[1]] * scale, val[map[2]] * scale};
        if (!finite3(v)) {
            set_error("m3d: non-finite value at entry %zu", e);
            return LUTR_ERR_INVALID_DATA;
        }
        std::memcpy(&out.rgb[3 * e], v, sizeof(v));
    }
    return LUTR_OK;
}

float clip01(float s) {
    const float m = s > 0.f ? s : 0.f;
    return m > 1.f ? 1.f : m;
}

int parse_csp(const char *text, size_t len, CubeLut &out) {
    Lines rd(text, len);
    char line[512];
    float in_min[3] = {0.f, 0.f, 0.f}, in_max[3] = {1.f, 1.f, 1.f}, out_min[3] = {0.f, 0.f, 0.f}, out_max[3] = {1.f, 1.f, 1.f};
    auto eof = [&]() { return eof_error("csp", 0, 0); };
    if (!rd.next_content(line)) return eof();
    if (std::strncmp(line, "CSPLUTV100", 10)) {
        set_error("csp: Not cineSpace LUT format");
        return LUTR_ERR_INVALID_DATA;
    }
    if (!rd.next_content(line)) return eof();
    if (std::strncmp(line, "3D", 2)) {
        set_error("csp: Not 3D LUT format");
        return LUTR_ERR_INVALID_DATA;
    }
    bool inside_metadata = false;
    for (;;) {
        if (!rd.next_content(line)) return eof();
        if (!std::strncmp(line, "BEGIN METADATA", 14)) {
            inside_metadata = true;
            continue;
        }
        if (!std::strncmp(line, "END METADATA", 12)) {
            inside_metadata = false;
            continue;
        }
        if (inside_metadata) continue;
        for (int c = 0; c < 3; ++c) {
            const long npoints = std::strtol(line, nullptr, 0);
            if (npoints > 2) {
                set_error("csp: a %ld-point pre-LUT (1D shaper in front of the cube) is not supported", npoints);
                return LUTR_ERR_UNSUPPORTED;
            }
            if (!rd.next_content(line)) return eof();
            if (std::sscanf(line, "%f %f", &in_min[c], &in_max[c]) != 2) {
                set_error("csp: invalid input range line");
                return LUTR_ERR_INVALID_DATA;
            }
            if (!rd.next_content(line)) return eof();
            if (std::sscanf(line, "%f %f", &out_min[c], &out_max[c]) != 2) {
                set_error("csp: invalid output range line");
                return LUTR_ERR_INVALID_DATA;
            }
            if (!rd.next_content(line)) return eof();
        }
        int sr, sg, sb;
        if (std::sscanf(line, "%d %d %d", &sr, &sg, &sb) != 3) {
            set_error("csp: invalid size line");
            return LUTR_ERR_INVALID_DATA;
        }
        if (sr != sg || sr != sb) {
            set_error("csp: Unsupported size combination: %dx%dx%d", sr, sg, sb);
            return LUTR_ERR_INVALID_DATA;
        }
        if (!alloc_table("csp", sr, out)) return LUTR_ERR_LUT_SIZE;
        const size_t n = (size_t)sr, n2 = n * n, total = n2 * n;
        const float gain[3] = {out_max[0] - out_min[0], out_max[1] - out_min[1], out_max[2] - out_min[2]};
        size_t e = 0;
        for (size_t b = 0; b < n; ++b)          // red fastest in the file, as in .cube
            for (size_t g = 0; g < n; ++g)
                for (size_t r = 0; r < n; ++r, ++e) {
                    if (!rd.next_content(line)) return eof_error("csp", e, total);
                    float v[3];
                    if (std::sscanf(line, "%f %f %f", &v[0], &v[1], &v[2]) != 3) {
                        set_error("csp: invalid data line at entry %zu", e);
                        return LUTR_ERR_INVALID_DATA;
                    }
                    for (int c = 0; c < 3; ++c) v[c] *= gain[c];
                    if (!finite3(v)) {
                        set_error("csp: non-finite value at entry %zu", e);
                        return LUTR_ERR_INVALID_DATA;
                    }
                    std::memcpy(&out.rgb[3 * (r * n2 + g * n + b)], v, sizeof(v));
                }
        break;
    }
    for (int c = 0; c < 3; ++c) out.scale[c] = clip01((float)(1. / ((double)in_max[c] - (double)in_min[c])));
    return LUTR_OK;
}

bool ieq(const char *a, const char *b) {
    for (; *a && *b; ++a, ++b)
        if (std::tolower((unsigned char)*a) != std::tolower((unsigned char)*b)) return false;
    return *a == *b;
}

}  // namespace

int lut_dialect_of_path(const char *path) {
    const char *ext = path ? std::strrchr(path, '.') : nullptr;
    if (!ext) {
        set_error("lut: Unable to guess the format from the extension");
        return LUTR_ERR_INVALID_ARG;
    }
    ++ext;
    if (ieq(ext, "cube")) return LUTR_LUT_CUBE;
    if (ieq(ext, "dat")) return LUTR_LUT_DAT;
    if (ieq(ext, "3dl")) return LUTR_LUT_3DL;
    if (ieq(ext, "m3d")) return LUTR_LUT_M3D;
    if (ieq(ext, "csp")) return LUTR_LUT_CSP;
    set_error("lut: Unrecognized '.%s' file type", ext);
    return LUTR_ERR_INVALID_ARG;
}

int parse_lut_text(const char *text, size_t len, int dialect, CubeLut &out) {
    if (!text) {
        set_error("lut: NULL text");
        return LUTR_ERR_INVALID_ARG;
    }
    out = CubeLut();
    switch (dialect) {
        case LUTR_LUT_CUBE: return parse_cube(text, len, out);
        case LUTR_LUT_DAT: return parse_dat(text, len, out);
        case LUTR_LUT_3DL: return parse_3dl(text, len, out);
        case LUTR_LUT_M3D: return parse_m3d(text, len, out);
        case LUTR_LUT_CSP: return parse_csp(text, len, out);
    }
    set_error("lut: unknown dialect %d", dialect);
    return LUTR_ERR_INVALID_ARG;
}

int parse_lut_file(const char *path, CubeLut &out) {
    const int dialect = lut_dialect_of_path(path);
    if (dialect < 0) return dialect;
    if (dialect == LUTR_LUT_CUBE) return parse_cube_file(path, out);
    FILE *f = std::fopen(path, "rb");
    if (!f) {
        set_error("lut: cannot open '%s': %s", path, std::strerror(errno));
        return LUTR_ERR_IO;
    }
    std::string buf;
    char chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, got);
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) {
        set_error("lut: read error on '%s'", path);
        return LUTR_ERR_IO;
    }
    return parse_lut_text(buf.data(), buf.size(), dialect, out);
}

}  // namespace lutr
