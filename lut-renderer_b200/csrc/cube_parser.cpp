// .cube reader for the device LUT upload.  Behaviour follows what the reference gets from
// `lut3d=file='X.cube'` (reference src/lut_renderer/ffmpeg.py:246), i.e. FFmpeg vf_lut3d's parse_cube():
//   * everything before the first line starting with "LUT_3D_SIZE" is ignored (so TITLE / DOMAIN_* lines
//     placed before it have no effect);
//   * size = strtol(base 0) of the text after the keyword; 2 <= size <= 256 or "invalid 3D LUT size";
//   * then size^3 data lines, red index fastest in the file, stored [r][g][b];
//     while reading data: blank / whitespace-only / '#' lines are skipped, "TITLE..." lines are skipped,
//     "DOMAIN_MIN " / "DOMAIN_MAX " update the domain, any other "DOMAIN_*" is invalid data,
//     any other line must scan as three floats or the file is invalid; running out of lines is invalid;
//   * lines are consumed in pieces of at most 511 bytes, as a 512-byte fgets() buffer would deliver them;
//   * scale[c] = clip(1 / (max[c] - min[c]), 0, 1); the minimum itself is not subtracted.
// One deliberate deviation: non-finite table entries are rejected (the kernels rely on 0 * entry == 0).
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <thread>

#include "lutr_internal.h"

namespace lutr {

namespace {

thread_local char g_error[512] = "";

// Hands out the pieces a 512-byte fgets() would: up to 511 bytes, ending after '\n' if one is found.
class LineReader {
  public:
    LineReader(const char *text, size_t len) : p_(text), end_(text + len) {}
    bool next(char (&buf)[512]) {
        if (p_ >= end_) return false;
        size_t n = 0;
        while (p_ < end_ && n < 511) {
            char ch = *p_++;
            buf[n++] = ch;
            if (ch == '\n') break;
        }
        buf[n] = 0;
        return true;
    }
    const char *pos() const { return p_; }

  private:
    const char *p_, *end_;
};

bool starts_with(const char *s, const char *prefix) { return std::strncmp(s, prefix, std::strlen(prefix)) == 0; }

bool is_blank_or_comment(const char *s) {
    while (*s == ' ' || *s == '\f' || *s == '\n' || *s == '\r' || *s == '\t' || *s == '\v') ++s;
    return *s == 0 || *s == '#';
}

// One "%f" conversion.  Plain decimal tokens ([+-]digits[.digits][e[+-]digits], what every .cube writer emits) go through
// std::from_chars: correctly rounded like strtof, ~5x faster (a 65-point table has 824 k numbers).  Anything else (hex floats,
// inf / nan, locale oddities) takes strtof itself, so the accepted syntax is exactly strtof's.
const char *scan_float(const char *s, float &out) {
    const char *p = s;
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\v' || *p == '\f' || *p == '\r') ++p;
    const char *q = *p == '+' ? p + 1 : p;
    const char *t = *q == '-' ? q + 1 : q;
    const bool plain = (*q != '-' || p == q) && ((*t >= '0' && *t <= '9') || (*t == '.' && t[1] >= '0' && t[1] <= '9')) &&
                       !(*t == '0' && (t[1] == 'x' || t[1] == 'X'));
    if (plain) {
        const char *end = q;
        while (*end && *end != ' ' && *end != '\t' && *end != '\n' && *end != '\r' && *end != '\v' && *end != '\f') ++end;
        float f;
        const auto r = std::from_chars(q, end, f, std::chars_format::general);
        if (r.ec == std::errc() && r.ptr != q) {
            out = f;
            return r.ptr;
        }
    }
    char *e = nullptr;
    errno = 0;
    const float f = std::strtof(s, &e);
    if (e == s) return s;
    out = f;
    return e;
}

// "%f %f %f": three floats separated by optional whitespace; returns how many were converted.
int scan3(const char *s, float v[3]) {
    int got = 0;
    for (; got < 3; ++got) {
        const char *e = scan_float(s, v[got]);
        if (e == s) break;
        s = e;
    }
    return got;
}

float clip01_recip(float mn, float mx) {
    float s = (float)(1. / ((double)mx - (double)mn));  // double division of the float difference promoted, as upstream
    // av_clipf(x, 0, 1) = min(max(x, 0), 1); NaN stays out of the way of max()
    float m = s > 0.f ? s : 0.f;
    return m > 1.f ? 1.f : m;
}

// ---- parallel data section ---------------------------------------------------------------------------------------------------
// A 65-point .cube is 7 MB of text (824 k numbers); one core parses it in ~20 ms, which is as long as the GPU needs for 24 4K frames
// (BASELINE cfg 5: per-clip LUTs).  The data section is cut at line boundaries into one piece per thread; every piece applies the
// sequential rules line by line.  Anything the sequential reader would treat specially in a position-dependent way (a line longer
// than the 511-byte fgets window, a malformed or non-finite data line, too few entries) makes the caller fall back to the
// sequential reader, so accepted files give identical tables and rejected files identical errors.
struct DomainLine {
    size_t entries_before;  // data entries of this piece in front of the line
    bool is_max;
    int count;
    float v[3];
};
struct Piece {
    const char *begin = nullptr, *end = nullptr;
    std::vector<float> vals;  // 3 per entry, file order
    std::vector<DomainLine> domains;
    bool anomaly = false;
};

void parse_piece(Piece &pc) {
    const char *p = pc.begin;
    pc.vals.reserve((size_t)(pc.end - pc.begin) / 8);
    char buf[512];
    while (p < pc.end) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(pc.end - p));
        const char *le = nl ? nl + 1 : pc.end;
        const size_t len = (size_t)(le - p);
        if (len > 511) {
            pc.anomaly = true;
            return;
        }
        std::memcpy(buf, p, len);
        buf[len] = 0;
        p = le;
        if (buf[0] == 'D' && starts_with(buf, "DOMAIN_")) {
            DomainLine d{};
            if (starts_with(buf + 7, "MIN ")) d.is_max = false;
            else if (starts_with(buf + 7, "MAX ")) d.is_max = true;
            else {
                pc.anomaly = true;
                return;
            }
            d.entries_before = pc.vals.size() / 3;
            d.count = scan3(buf + 11, d.v);
            pc.domains.push_back(d);
            continue;
        }
        if (buf[0] == 'T' && starts_with(buf, "TITLE")) continue;
        if (is_blank_or_comment(buf)) continue;
        float v[3];
        if (scan3(buf, v) != 3 || !std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) {
            pc.anomaly = true;
            return;
        }
        pc.vals.push_back(v[0]), pc.vals.push_back(v[1]), pc.vals.push_back(v[2]);
    }
}

// Fills file_order (n^3 entries, file order) and the domain from [begin, end); false = let the sequential reader decide.
bool parse_data_parallel(const char *begin, const char *end, size_t entries, std::vector<float> &file_order, float dmin[3], float dmax[3]) {
    const size_t bytes = (size_t)(end - begin);
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = hw > 8 ? 8 : hw;
    if (bytes < (256u << 10) || nt < 2) return false;
    std::vector<Piece> pieces(nt);
    const char *p = begin;
    for (unsigned i = 0; i < nt; ++i) {
        pieces[i].begin = p;
        const char *q = i + 1 == nt ? end : begin + bytes * (i + 1) / nt;
        if (q < p) q = p;
        if (q < end) {
            const char *nl = (const char *)std::memchr(q, '\n', (size_t)(end - q));
            q = nl ? nl + 1 : end;
        }
        pieces[i].end = q;
        p = q;
    }
    std::vector<std::thread> threads;
    for (unsigned i = 1; i < nt; ++i) threads.emplace_back(parse_piece, std::ref(pieces[i]));
    parse_piece(pieces[0]);
    for (std::thread &t : threads) t.join();
    size_t total = 0;
    for (const Piece &pc : pieces) {
        if (pc.anomaly) return false;
        total += pc.vals.size() / 3;
    }
    if (total < entries) return false;  // unexpected EOF: the sequential reader words the error
    size_t base = 0;
    for (const Piece &pc : pieces) {
        // DOMAIN lines the sequential reader would still have read: those in front of the last entry it needs
        for (const DomainLine &d : pc.domains)
            if (base + d.entries_before < entries)
                for (int k = 0; k < d.count; ++k) (d.is_max ? dmax : dmin)[k] = d.v[k];
        if (base < entries) {
            const size_t take = std::min(pc.vals.size() / 3, entries - base);
            std::memcpy(&file_order[3 * base], pc.vals.data(), take * 3 * sizeof(float));
        }
        base += pc.vals.size() / 3;
    }
    return true;
}

}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

const char *get_error() { return g_error; }

int parse_cube(const char *text, size_t len, CubeLut &out) {
    if (!text) {
        set_error("cube: NULL text");
        return LUTR_ERR_INVALID_ARG;
    }
    LineReader rd(text, len);
    char line[512];
    float dmin[3] = {0.f, 0.f, 0.f}, dmax[3] = {1.f, 1.f, 1.f};
    out = CubeLut();

    bool found = false;
    while (rd.next(line)) {
        if (starts_with(line, "LUT_3D_SIZE")) {
            found = true;
            break;
        }
    }
    if (!found) {
        set_error("cube: 3D LUT is empty (no LUT_3D_SIZE line)");
        return LUTR_ERR_INVALID_DATA;
    }
    const long size = std::strlen(line) > 12 ? std::strtol(line + 12, nullptr, 0) : 0;
    if (size < 2 || size > 256) {
        set_error("cube: Too large or invalid 3D LUT size (%ld)", size);
        return LUTR_ERR_LUT_SIZE;
    }
    const size_t n = (size_t)size, n2 = n * n;
    out.size = (int)size;
    out.rgb.assign(n2 * n * 3, 0.f);
    // the file runs red-fastest, lut3d's table is [r][g][b]: collect in file order, transpose afterwards (a scattered store per
    // entry costs a cache miss each on a 65-point table)
    std::vector<float> file_order(n2 * n * 3);

    float pmin[3] = {0.f, 0.f, 0.f}, pmax[3] = {1.f, 1.f, 1.f};
    const bool parallel_ok = parse_data_parallel(rd.pos(), text + len, n2 * n, file_order, pmin, pmax);
    if (parallel_ok)
        for (int c = 0; c < 3; ++c) dmin[c] = pmin[c], dmax[c] = pmax[c];
    for (size_t b = 0; b < n && !parallel_ok; ++b)
        for (size_t g = 0; g < n; ++g)
            for (size_t r = 0; r < n; ++r) {
                for (;;) {
                    if (!rd.next(line)) {
                        set_error("cube: Unexpected EOF after %zu of %zu entries", (b * n + g) * n + r, n2 * n);
                        return LUTR_ERR_INVALID_DATA;
                    }
                    if (line[0] == 'D' && starts_with(line, "DOMAIN_")) {
                        float *dst = nullptr;
                        if (starts_with(line + 7, "MIN ")) dst = dmin;
                        else if (starts_with(line + 7, "MAX ")) dst = dmax;
                        if (!dst) {
                            set_error("cube: invalid DOMAIN_ keyword");
                            return LUTR_ERR_INVALID_DATA;
                        }
                        float v[3];
                        int k = scan3(line + 11, v);
                        for (int i = 0; i < k; ++i) dst[i] = v[i];
                        continue;
                    }
                    if (line[0] == 'T' && starts_with(line, "TITLE")) continue;
                    if (!is_blank_or_comment(line)) break;
                }
                float v[3];
                if (scan3(line, v) != 3) {
                    set_error("cube: invalid data line at entry %zu", (b * n + g) * n + r);
                    return LUTR_ERR_INVALID_DATA;
                }
                if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) {
                    set_error("cube: non-finite value at entry %zu", (b * n + g) * n + r);
                    return LUTR_ERR_INVALID_DATA;
                }
                float *dstv = &file_order[3 * ((b * n + g) * n + r)];
                dstv[0] = v[0], dstv[1] = v[1], dstv[2] = v[2];
            }
    constexpr size_t T = 16;  // tile of the (b, r) transpose, per g
    for (size_t g = 0; g < n; ++g)
        for (size_t b0 = 0; b0 < n; b0 += T)
            for (size_t r0 = 0; r0 < n; r0 += T)
                for (size_t r = r0; r < r0 + T && r < n; ++r)
                    for (size_t b = b0; b < b0 + T && b < n; ++b) {
                        const float *src = &file_order[3 * ((b * n + g) * n + r)];
                        float *dstv = &out.rgb[3 * (r * n2 + g * n + b)];
                        dstv[0] = src[0], dstv[1] = src[1], dstv[2] = src[2];
                    }
    for (int c = 0; c < 3; ++c) out.scale[c] = clip01_recip(dmin[c], dmax[c]);
    return LUTR_OK;
}

int parse_cube_file(const char *path, CubeLut &out) {
    if (!path) {
        set_error("cube: NULL path");
        return LUTR_ERR_INVALID_ARG;
    }
    FILE *f = std::fopen(path, "rb");
    if (!f) {
        set_error("cube: cannot open '%s': %s", path, std::strerror(errno));
        return LUTR_ERR_IO;
    }
    std::string buf;
    char chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, got);
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) {
        set_error("cube: read error on '%s'", path);
        return LUTR_ERR_IO;
    }
    return parse_cube(buf.data(), buf.size(), out);
}

}  // namespace lutr
