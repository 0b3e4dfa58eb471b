"""TEST INFRASTRUCTURE ONLY (see oracle/README): an independent, line-by-line Python restatement of the LUT file readers FFmpeg's
lut3d has besides .cube -- parse_dat, parse_3dl, parse_m3d, parse_cinespace in libavfilter/vf_lut3d.c -- and of the extension
dispatch in its init.  Reference call site: src/lut_renderer/ffmpeg.py:246 (`lut3d=file='...'`; the GUI only offers *.cube,
lut_manager.py:121).  PARITY UNPINNED: written from memory of upstream, no libavfilter in this image to run against.
Only tests/ may import this; the product readers are lut-renderer_b200/csrc/lut_dialects.cpp.

Each reader returns (size, scale[3], table float32 [r][g][b][3]) or raises ValueError (invalid data) / NotImplementedError."""
import re

import numpy as np

_F = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf(?:inity)?|nan)"


def _fgets_lines(data: bytes):
    """The pieces a 512-byte fgets() delivers: at most 511 bytes, ending after a newline."""
    pos = 0
    while pos < len(data):
        end = data.find(b"\n", pos, pos + 511)
        end = min(pos + 511, len(data)) if end < 0 else end + 1
        yield data[pos:end].decode("latin-1")
        pos = end


def _skip(line: str) -> bool:
    s = line.lstrip(" \t\n\v\f\r")
    return not s or s[0] == "#"


class _Reader:
    def __init__(self, data: bytes):
        self.it = _fgets_lines(data)

    def line(self, filtered=True) -> str:
        for ln in self.it:
            if not filtered or not _skip(ln):
                return ln
        raise ValueError("Unexpected EOF")


def _scan(line: str, kinds: str):
    """sscanf of "%f %f %f" / "%d %d %d" / "%f %f": as many leading conversions as match."""
    out, pos = [], 0
    for k in kinds:
        m = re.compile(r"\s*(" + (_F if k == "f" else r"[-+]?\d+") + ")", re.I).match(line, pos)
        if not m:
            break
        out.append(np.float32(m.group(1)) if k == "f" else int(m.group(1)))
        pos = m.end()
    return out


def _strtol0(s: str) -> int:
    m = re.match(r"\s*([-+]?)(0[xX][0-9a-fA-F]+|0[0-7]*|[1-9]\d*)", s)
    if not m:
        return 0
    t = m.group(2)
    v = int(t, 16) if t[:2].lower() == "0x" else (int(t, 8) if t.startswith("0") and len(t) > 1 else int(t))
    return -v if m.group(1) == "-" else v


def _alloc(size: int) -> np.ndarray:
    if size < 2 or size > 256:
        raise ValueError("Too large or invalid 3D LUT size")
    return np.zeros((size, size, size, 3), np.float32)


def parse_dat(data: bytes):
    rd = _Reader(data)
    size = 33
    line = rd.line()
    if line.startswith("3DLUTSIZE "):
        size = _strtol0(line[10:])
        line = rd.line()
    t = _alloc(size)
    flat = t.reshape(-1, 3)          # [r][g][b]: the file order
    for e in range(size ** 3):
        if e:
            line = rd.line()
        v = _scan(line, "fff")
        if len(v) != 3:
            raise ValueError("invalid data")
        flat[e] = v
    return size, [1.0, 1.0, 1.0], t


def parse_3dl(data: bytes):
    rd = _Reader(data)
    t = _alloc(17)
    flat = t.reshape(-1, 3)
    rd.line()                        # the shaper line
    for e in range(17 ** 3):
        v = _scan(rd.line(), "ddd")
        if len(v) != 3:
            raise ValueError("invalid data")
        flat[e] = [np.float32(x) / np.float32(4096) for x in v]
    return 17, [1.0, 1.0, 1.0], t


def parse_m3d(data: bytes):
    rd = _Reader(data)
    n_in = n_out = -1
    order = [0, 1, 2]
    try:
        while True:
            line = rd.line(filtered=False)
            if line.startswith("in"):
                n_in = _strtol0(line[2:])
            elif line.startswith("out"):
                n_out = _strtol0(line[3:])
            elif line.startswith("values"):
                for i, tok in enumerate(line[6:].split()[:3]):
                    if tok[0] in "rgb":
                        order[i] = "rgb".index(tok[0])
                break
    except ValueError:
        pass
    if n_in == -1 or n_out == -1:
        raise ValueError("in and out must be defined")
    if n_in < 2 or n_out < 2 or n_in > 256 ** 3 or n_out > 256 ** 3:
        raise ValueError("invalid in / out")
    size = 1
    while size ** 3 < n_in:
        size += 1
    t = _alloc(size)
    flat = t.reshape(-1, 3)
    scale = np.float32(1.0 / (n_out - 1))
    for e in range(size ** 3):
        v = _scan(rd.line(filtered=False), "fff")
        if len(v) != 3:
            raise ValueError("invalid data")
        flat[e] = [v[order[0]] * scale, v[order[1]] * scale, v[order[2]] * scale]
    return size, [1.0, 1.0, 1.0], t


def parse_csp(data: bytes):
    rd = _Reader(data)
    if not rd.line().startswith("CSPLUTV100"):
        raise ValueError("Not cineSpace LUT format")
    if not rd.line().startswith("3D"):
        raise ValueError("Not 3D LUT format")
    lo_in, hi_in, lo_out, hi_out = [0.0] * 3, [1.0] * 3, [0.0] * 3, [1.0] * 3
    inside = False
    while True:
        line = rd.line()
        if line.startswith("BEGIN METADATA"):
            inside = True
            continue
        if line.startswith("END METADATA"):
            inside = False
            continue
        if inside:
            continue
        for c in range(3):
            if _strtol0(line) > 2:
                raise NotImplementedError("pre-LUT")
            v = _scan(rd.line(), "ff")
            if len(v) != 2:
                raise ValueError("invalid input range")
            lo_in[c], hi_in[c] = v
            v = _scan(rd.line(), "ff")
            if len(v) != 2:
                raise ValueError("invalid output range")
            lo_out[c], hi_out[c] = v
            line = rd.line()
        sz = _scan(line, "ddd")
        if len(sz) != 3 or sz[0] != sz[1] or sz[0] != sz[2]:
            raise ValueError("size line")
        size = sz[0]
        t = _alloc(size)
        gain = [np.float32(hi_out[c]) - np.float32(lo_out[c]) for c in range(3)]
        for b in range(size):
            for g in range(size):
                for r in range(size):   # red fastest in the file
                    v = _scan(rd.line(), "fff")
                    if len(v) != 3:
                        raise ValueError("invalid data")
                    t[r, g, b] = [v[c] * gain[c] for c in range(3)]
        break
    scale = []
    for c in range(3):
        with np.errstate(divide="ignore"):
            s = np.float32(np.float64(1.0) / (np.float64(np.float32(hi_in[c])) - np.float64(np.float32(lo_in[c]))))
        scale.append(float(min(max(s, np.float32(0)), np.float32(1))))
    return size, scale, t


READERS = {"dat": parse_dat, "3dl": parse_3dl, "m3d": parse_m3d, "csp": parse_csp}


def dialect_of(path: str) -> str:
    """lut3d's init: the text after the last '.', case-insensitive; ValueError when there is none or it is unknown."""
    if "." not in path:
        raise ValueError("Unable to guess the format from the extension")
    ext = path.rsplit(".", 1)[1].lower()
    if ext not in ("cube", "dat", "3dl", "m3d", "csp"):
        raise ValueError(f"Unrecognized '.{ext}' file type")
    return ext
