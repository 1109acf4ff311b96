"""oracle/sws_oracle.py -- CPU restatement (numpy) of the reference's legacy libswscale path.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg as a CHECKER; never by the product (librempeg_b200/).

It restates, function by function, the reference algorithm for the hot path
(hscale -> [range convert] -> vscale -> pixel pack) the way the reference
computes it -- including the byte LUTs of yuv2rgb.c, which the product replaces
by a closed form -- so that product and oracle share no code and no shortcuts.

Pinning (tests/test_oracle_cpu.py):
  * reference FATE golden CRCs  tests/ref/fate/filter-scalechroma (15 frames) and
    tests/ref/fate/sws-yuv-range, on vsynth1 input made by the reference's tests/videogen.c;
  * outputs of the real reference built by oracle/build_ref.py (oracle/_ref), live when
    the .so is present and through the committed fixtures in tests/golden/.

All file:line citations are relative to /root/reference/libswscale/.
"""
import math

import numpy as np

# ---- flags (swscale.h:131-208) ----
SWS_FAST_BILINEAR = 1 << 0
SWS_BILINEAR = 1 << 1
SWS_BICUBIC = 1 << 2
SWS_X = 1 << 3
SWS_POINT = 1 << 4
SWS_AREA = 1 << 5
SWS_BICUBLIN = 1 << 6
SWS_GAUSS = 1 << 7
SWS_SINC = 1 << 8
SWS_LANCZOS = 1 << 9
SWS_SPLINE = 1 << 10
SWS_FULL_CHR_H_INT = 1 << 13
SWS_FULL_CHR_H_INP = 1 << 14
SWS_ACCURATE_RND = 1 << 18
SWS_BITEXACT = 1 << 19
BX = SWS_ACCURATE_RND | SWS_BITEXACT
SWS_PARAM_DEFAULT = 123456
SCALER_MASK = (SWS_POINT | SWS_AREA | SWS_BILINEAR | SWS_FAST_BILINEAR | SWS_BICUBIC | SWS_X |
               SWS_GAUSS | SWS_LANCZOS | SWS_SINC | SWS_SPLINE | SWS_BICUBLIN)

# yuv2rgb.c:47-59
YUV2RGB_COEFFS = {
    0: (104597, 132201, 25675, 53279), 1: (117489, 138438, 13975, 34925),
    2: (104597, 132201, 25675, 53279), 3: (104597, 132201, 25675, 53279),
    4: (104448, 132798, 24759, 53109), 5: (104597, 132201, 25675, 53279),
    6: (104597, 132201, 25675, 53279), 7: (117579, 136230, 16907, 35559),
    9: (110013, 140363, 12277, 42626), 10: (110013, 140363, 12277, 42626),
}

# swscale.c:42-52
DITHER_8x8_128 = np.array([
    [36, 68, 60, 92, 34, 66, 58, 90], [100, 4, 124, 28, 98, 2, 122, 26],
    [52, 84, 44, 76, 50, 82, 42, 74], [116, 20, 108, 12, 114, 18, 106, 10],
    [32, 64, 56, 88, 38, 70, 62, 94], [96, 0, 120, 24, 102, 6, 126, 30],
    [48, 80, 40, 72, 54, 86, 46, 78], [112, 16, 104, 8, 118, 22, 110, 14]], np.int64)


def _fmt(name):
    """(kind, depth, log2_chroma_w, log2_chroma_h); kind in planar/semi/rgb8/rgb32/rgb16."""
    if name in ("rgb24", "bgr24"):
        return ("rgb8", 8, 0, 0)
    if name in ("rgba", "bgra", "argb", "abgr"):
        return ("rgb32", 8, 0, 0)
    if name in ("rgb48le", "bgr48le"):
        return ("rgb16", 16, 0, 0)
    if name in ("rgb565le", "bgr565le", "rgb555le", "bgr555le"):
        return ("rgb565", 5, 0, 0)          # 15/16 bpp, destinations only
    if name in ("nv12", "nv21"):
        return ("semi", 8, 1, 1)
    if name == "p010le":
        return ("semi", 10, 1, 1)
    if name == "grayf32le":
        return ("grayf", 32, 0, 0)          # destinations only
    if name == "gbrpf32le":
        return ("rgbpf", 32, 0, 0)          # destinations only; planar RGB always takes the full-chroma path
    for key, (cw, ch) in {"420": (1, 1), "422": (1, 0), "444": (0, 0)}.items():
        for pre in ("yuvj", "yuv"):
            head = pre + key + "p"
            if name.startswith(head):
                rest = name[len(head):]
                return ("planar", 8 if rest == "" else int(rest[:-2]), cw, ch)
    raise ValueError("oracle: unsupported format " + name)


def _cdiv_shift(a, s):
    return -((-a) >> s)


def _cdiv(a, b):
    """C integer division (truncation toward zero) on Python ints."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _rounded_div(a, b):
    """ROUNDED_DIV, libavutil/common.h:58."""
    return _cdiv(a + (b >> 1), b) if a >= 0 else _cdiv(a - (b >> 1), b)


def _wrap32(x):
    """Reinterpret int64 values as the int32 the C code would hold after unsigned wrap."""
    return ((np.asarray(x, np.int64) + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)


# --------------------------------------------------------------------------- initFilter
def _spline(a, b, c, d, dist):
    """getSplineCoeff, utils.c:155-166."""
    if dist <= 1.0:
        return ((d * dist + c) * dist + b) * dist + a
    return _spline(0.0, b + 2.0 * c + 3.0 * d, c + 3.0 * d, -b - 3.0 * c - 6.0 * d, dist - 1.0)


def init_filter(x_inc, src_w, dst_w, one, scaler, flags, param, src_pos, dst_pos, filter_align=1,
                src_vec=None, dst_vec=None):
    """initFilter, utils.c:197-612; src_vec / dst_vec are the SwsFilter vectors of this bank.  Returns (coef[dst_w][fs], pos[dst_w])
    or None when the reference would ask for a cascade (RETCODE_USE_CASCADE)."""
    ratio = src_w // dst_w
    fone = 1 << (54 - min(int(math.log2(ratio | 1)), 8))
    pos = [0] * dst_w

    if abs(x_inc - 0x10000) < 10 and src_pos == dst_pos:                 # :219 unscaled
        fs = 1
        filt = [[fone] for _ in range(dst_w)]
        pos = list(range(dst_w))
    elif scaler == SWS_POINT:                                           # :229
        fs = 1
        x = ((dst_pos * x_inc) >> 8) - ((src_pos * 0x8000) >> 7)
        filt = []
        for i in range(dst_w):
            pos[i] = (x - ((fs - 1) << 15) + (1 << 15)) >> 16
            filt.append([fone])
            x += x_inc
    elif (x_inc <= (1 << 16) and scaler == SWS_AREA) or scaler == SWS_FAST_BILINEAR:   # :244
        fs = 2
        x = ((dst_pos * x_inc) >> 8) - ((src_pos * 0x8000) >> 7)
        filt = []
        for i in range(dst_w):
            xx = (x - ((fs - 1) << 15) + (1 << 15)) >> 16
            pos[i] = xx
            row = []
            for j in range(fs):
                c = fone - abs(xx * (1 << 16) - x) * (fone >> 16)
                row.append(max(c, 0))
                xx += 1
            filt.append(row)
            x += x_inc
    else:                                                               # :268 general
        size_factor = {SWS_AREA: 1, SWS_BICUBIC: 4, SWS_BILINEAR: 2, SWS_GAUSS: 8, SWS_SINC: 20,
                       SWS_SPLINE: 20, SWS_X: 8}.get(scaler, -1)
        if scaler == SWS_LANCZOS:
            size_factor = int(math.ceil(2 * param[0])) if param[0] != SWS_PARAM_DEFAULT else 6
        assert 0 < size_factor <= 50
        if x_inc <= 1 << 16:
            fs = 1 + size_factor
        else:
            fs = 1 + (size_factor * src_w + dst_w - 1) // dst_w
        fs = max(min(fs, src_w - 2), 1)
        x = ((dst_pos * x_inc) >> 7) - ((src_pos * 0x10000) >> 7)
        filt = []
        for i in range(dst_w):
            xx = _cdiv(x - (fs - 2) * (1 << 16), 1 << 17)
            pos[i] = xx
            row = []
            for j in range(fs):
                d = abs((xx * (1 << 17)) - x) << 13
                if x_inc > 1 << 16:
                    d = _cdiv(d * dst_w, src_w)
                fd = d * (1.0 / (1 << 30))
                if scaler == SWS_BICUBIC:
                    B = int((param[0] if param[0] != SWS_PARAM_DEFAULT else 0) * (1 << 24))
                    C = int((param[1] if param[1] != SWS_PARAM_DEFAULT else 0.6) * (1 << 24))
                    if d >= 1 << 31:
                        coeff = 0
                    else:
                        dd = (d * d) >> 30
                        ddd = (dd * d) >> 30
                        if d < 1 << 30:
                            coeff = ((12 * (1 << 24) - 9 * B - 6 * C) * ddd +
                                     (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                                     (6 * (1 << 24) - 2 * B) * (1 << 30))
                        else:
                            coeff = ((-B - 6 * C) * ddd + (6 * B + 30 * C) * dd +
                                     (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1 << 30))
                    coeff = _cdiv(coeff, (1 << 54) // fone)
                elif scaler == SWS_X:
                    A = param[0] if param[0] != SWS_PARAM_DEFAULT else 1.0
                    c = math.cos(fd * math.pi) if fd < 1.0 else -1.0
                    c = -math.pow(-c, A) if c < 0.0 else math.pow(c, A)
                    coeff = int((c * 0.5 + 0.5) * fone)
                elif scaler == SWS_AREA:
                    d2 = d - (1 << 29)
                    if d2 * x_inc < -(1 << (29 + 16)):
                        coeff = 1 << (30 + 16)
                    elif d2 * x_inc < (1 << (29 + 16)):
                        coeff = -d2 * x_inc + (1 << (29 + 16))
                    else:
                        coeff = 0
                    coeff *= fone >> (30 + 16)
                elif scaler == SWS_GAUSS:
                    p = param[0] if param[0] != SWS_PARAM_DEFAULT else 3.0
                    coeff = int(math.pow(2.0, -p * fd * fd) * fone)
                elif scaler == SWS_SINC:
                    coeff = int((math.sin(fd * math.pi) / (fd * math.pi) if d else 1.0) * fone)
                elif scaler == SWS_LANCZOS:
                    p = param[0] if param[0] != SWS_PARAM_DEFAULT else 3.0
                    coeff = int((math.sin(fd * math.pi) * math.sin(fd * math.pi / p) /
                                 (fd * fd * math.pi * math.pi / p) if d else 1.0) * fone)
                    if fd > p:
                        coeff = 0
                elif scaler == SWS_BILINEAR:
                    coeff = max((1 << 30) - d, 0) * (fone >> 30)
                elif scaler == SWS_SPLINE:
                    p = -2.196152422706632
                    coeff = int(_spline(1.0, 0.0, p, -p - 1.0, fd) * fone)
                else:
                    raise AssertionError("bad scaler")
                row.append(coeff)
                xx += 1
            filt.append(row)
            x += 2 * x_inc

    # :385-413 convolve every row with the source-side vector; the destination-side vector only widens the rows
    # ("FIXME dstFilter").  int64 += double * int64 truncates toward zero after every addition.
    f2 = fs + (len(src_vec) - 1 if src_vec is not None else 0) + (len(dst_vec) - 1 if dst_vec is not None else 0)
    if src_vec is not None or dst_vec is not None:
        new = []
        for i in range(dst_w):
            row = [0] * f2
            if src_vec is not None:
                for k, ck in enumerate(src_vec):
                    for j in range(fs):
                        row[k + j] = int(float(row[k + j]) + float(ck) * float(filt[i][j]))
            else:
                row[:fs] = filt[i]
            new.append(row)
            pos[i] += (fs - 1) // 2 - (f2 - 1) // 2
        filt, fs = new, f2

    # :417-457 reduce: shift near-zero taps out on the left, count them on the right
    f2 = fs
    cutoff = 0.002 * fone
    min_fs = 0
    for i in range(dst_w - 1, -1, -1):
        row = filt[i]
        acc = 0
        for _ in range(f2):
            acc += abs(row[0])
            if acc > cutoff:
                break
            if i < dst_w - 1 and pos[i] >= pos[i + 1]:
                break
            row.pop(0)
            row.append(0)
            pos[i] += 1
        acc = 0
        keep = f2
        for j in range(f2 - 1, 0, -1):
            acc += abs(row[j])
            if acc > cutoff:
                break
            keep -= 1
        min_fs = max(min_fs, keep)
    assert min_fs > 0
    fs = (min_fs + (filter_align - 1)) & ~(filter_align - 1)
    if fs >= 256 * 16 // 16:                                            # :492 (generic APCK_SIZE == 16)
        return None
    # :504-515 copy/truncate, zero the alignment padding under BITEXACT
    out = []
    for i in range(dst_w):
        row = [(filt[i][j] if j < f2 else 0) for j in range(fs)]
        if flags & SWS_BITEXACT:
            for j in range(min_fs, fs):
                row[j] = 0
        out.append(row)
    # :520-560 fold taps outside the image onto the border
    for i in range(dst_w):
        row = out[i]
        if pos[i] < 0:
            for j in range(1, fs):
                left = max(j + pos[i], 0)
                row[left] += row[j]
                row[j] = 0
            pos[i] = 0
        if pos[i] + fs > src_w:
            shift = pos[i] + min(fs - src_w, 0)
            acc = 0
            for j in range(fs - 1, -1, -1):
                if pos[i] + j >= src_w:
                    acc += row[j]
                    row[j] = 0
            for j in range(fs - 1, -1, -1):
                row[j] = 0 if j < shift else row[j - shift]
            pos[i] -= shift
            row[src_w - 1 - pos[i]] += acc
        assert 0 <= pos[i] < src_w
    # :569-588 normalise with error diffusion
    coef = np.zeros((dst_w, fs), np.int16)
    for i in range(dst_w):
        total = sum(out[i])
        total = _cdiv(total + one // 2, one)
        if not total:
            total = 1
        err = 0
        for j in range(fs):
            v = out[i][j] + err
            iv = _rounded_div(v, total)
            coef[i, j] = iv
            err = v - iv * total
    return coef, np.array(pos, np.int32)


# --------------------------------------------------------------------------- yuv2rgb tables
def _tile(t):
    t = np.array(t, np.int64)
    return np.tile(t, (8 // t.shape[0], 8 // t.shape[1]))


# the ordered-dither matrices of planarCopyWrapper (`dithers[shift - 1]`, swscale_unscaled.c:39-112) as their
# periodic tiles: 2x2 (1 and 2 bits), 4x4 (3, 4), 8x8 (5, 6 = 7, and 8 = ff_dither_8x8_128)
_D6 = [[18, 34, 30, 46, 17, 33, 29, 45], [50, 2, 62, 14, 49, 1, 61, 13], [26, 42, 22, 38, 25, 41, 21, 37],
       [58, 10, 54, 6, 57, 9, 53, 5], [16, 32, 28, 44, 19, 35, 31, 47], [48, 0, 60, 12, 51, 3, 63, 15],
       [24, 40, 20, 36, 27, 43, 23, 39], [56, 8, 52, 4, 59, 11, 55, 7]]
DEPTH_DITHER = [
    _tile([[0, 1], [1, 0]]),
    _tile([[1, 2], [3, 0]]),
    _tile([[2, 4, 3, 5], [6, 0, 7, 1], [3, 5, 2, 4], [7, 1, 6, 0]]),
    _tile([[4, 8, 7, 11], [12, 0, 15, 3], [6, 10, 5, 9], [14, 2, 13, 1]]),
    _tile([[9, 17, 15, 23, 8, 16, 14, 22], [25, 1, 31, 7, 24, 0, 30, 6], [13, 21, 11, 19, 12, 20, 10, 18],
           [29, 5, 27, 3, 28, 4, 26, 2], [8, 16, 14, 22, 9, 17, 15, 23], [24, 0, 30, 6, 25, 1, 31, 7],
           [12, 20, 10, 18, 13, 21, 11, 19], [28, 4, 26, 2, 29, 5, 27, 3]]),
    _tile(_D6),
    _tile(_D6),
    _tile([[36, 68, 60, 92, 34, 66, 58, 90], [100, 4, 124, 28, 98, 2, 122, 26], [52, 84, 44, 76, 50, 82, 42, 74],
           [116, 20, 108, 12, 114, 18, 106, 10], [32, 64, 56, 88, 38, 70, 62, 94], [96, 0, 120, 24, 102, 6, 126, 30],
           [48, 80, 40, 72, 54, 86, 46, 78], [112, 16, 104, 8, 118, 22, 110, 14]]),
]


def _round_i16(f):
    """roundToInt16, yuv2rgb.c:705-715 (as the int16_t the caller stores)."""
    r = (f + (1 << 15)) >> 16
    if r < -0x7FFF:
        return -0x8000
    if r > 0x7FFF:
        return 0x7FFF
    return r


def yuv2rgb_tables(inv_table, full_range, brightness, contrast, saturation):
    """ff_yuv2rgb_c_init_tables for bpp 24/48 (yuv2rgb.c:717-914) + fill_table/fill_gv_table (:680-703).
    Returns dict with the byte LUT, the four index tables and the six 16-bit path coefficients."""
    crv, cbu, cgu, cgv = inv_table[0], inv_table[1], -inv_table[2], -inv_table[3]
    cy, oy = 1 << 16, 0
    if not full_range:
        cy = (cy * 255) // 219
        oy = 16 << 16
    else:
        crv = _cdiv(crv * 224, 255); cbu = _cdiv(cbu * 224, 255)
        cgu = _cdiv(cgu * 224, 255); cgv = _cdiv(cgv * 224, 255)
    cy = (cy * contrast) >> 16
    crv = (crv * contrast * saturation) >> 32
    cbu = (cbu * contrast * saturation) >> 32
    cgu = (cgu * contrast * saturation) >> 32
    cgv = (cgv * contrast * saturation) >> 32
    oy -= 256 * brightness
    t = {"y_coeff": _round_i16(cy * (1 << 13)), "y_offset": _round_i16(oy * (1 << 9)),
         "v2r": _round_i16(crv * (1 << 13)), "v2g": _round_i16(cgv * (1 << 13)),
         "u2g": _round_i16(cgu * (1 << 13)), "u2b": _round_i16(cbu * (1 << 13))}
    div = max(cy, 1)
    crv = _cdiv(crv * (1 << 16) + 0x8000, div); cbu = _cdiv(cbu * (1 << 16) + 0x8000, div)
    cgu = _cdiv(cgu * (1 << 16) + 0x8000, div); cgv = _cdiv(cgv * (1 << 16) + 0x8000, div)
    headroom = 512
    yoffs = (384 if full_range else 326) + headroom
    size = 1024 + 2 * headroom
    y_table = np.zeros(size, np.uint8)
    yb = -(384 << 16) - headroom * cy - oy
    for i in range(size):
        y_table[i] = min(max((yb + 0x8000) >> 16, 0), 255)
        yb += cy

    def fill(inc, base):                       # fill_table: pointer into y_table -> index
        start = base - (inc >> 9)
        return np.array([start + ((min(max(i - 512, 0), 255) * inc) >> 16) for i in range(1280)], np.int64)

    t["y_table"] = y_table
    t["rV"] = fill(crv, yoffs)
    t["gU"] = fill(cgu, yoffs)
    t["bU"] = fill(cbu, yoffs)
    off = -(cgv >> 9)                          # fill_gv_table
    t["gV"] = np.array([off + ((min(max(i - 512, 0), 255) * cgv) >> 16) for i in range(1280)], np.int64)
    return t


# --------------------------------------------------------------------------- the context
def rgb2yuv_table(table):
    """fill_rgb2yuv_table (utils.c:614-706): {ry,gy,by,ru,gu,bu,rv,gv,bv} at 15 bits from the
    destination matrix, always limited range; BT.601 takes the classic rounded constants."""
    one, sh = 65536, 1 << 15
    vr, ub, ug, vg = table[0], table[1], -table[2], -table[3]
    cy = _cdiv(one * 255, 219)
    w = _rounded_div(one * one * ug, ub)
    v = _rounded_div(one * one * vg, vr)
    z = one * one - w - v
    Cy, Cu, Cv = _rounded_div(cy * z, one), _rounded_div(ub * z, one), _rounded_div(vr * z, one)
    out = [-_rounded_div(sh * v, Cy), _rounded_div(sh * one * one, Cy), -_rounded_div(sh * w, Cy),
           _rounded_div(sh * v, Cu), -_rounded_div(sh * one * one, Cu), _rounded_div(sh * (z + w), Cu),
           _rounded_div(sh * (v + z), Cv), -_rounded_div(sh * one * one, Cv), _rounded_div(sh * w, Cv)]
    if tuple(table) == YUV2RGB_COEFFS[5]:
        f = lambda k, span: int(k * span / 255 * (1 << 15) + 0.5)      # noqa: E731
        out = [f(0.299, 219), f(0.587, 219), f(0.114, 219), -f(0.169, 224), -f(0.331, 224), f(0.500, 224),
               f(0.500, 224), -f(0.419, 224), -f(0.081, 224)]
    return out


class OracleContext:
    """Geometry and path selection of ff_sws_init_single_context (utils.c:1137-1835) for the
    formats the hot path covers, then whole-frame conversion with the per-line kernels of
    swscale.c / output.c restated on numpy arrays."""

    def __init__(self, sw, sh, sfmt, dw, dh, dfmt, flags, param=None, src_range=0, dst_range=0,
                 chr_pos=(-513, -513, -513, -513), dither=1, colorspace=None, src_filter=None, dst_filter=None):
        self.sw, self.sh, self.dw, self.dh = sw, sh, dw, dh
        param = list(param) if param is not None else [SWS_PARAM_DEFAULT, SWS_PARAM_DEFAULT]
        if sfmt.startswith("yuvj"):                                   # handle_jpeg, utils.c:773
            sfmt, src_range = "yuv" + sfmt[4:], 1
        if dfmt.startswith("yuvj"):
            dfmt, dst_range = "yuv" + dfmt[4:], 1
        self.sfmt, self.dfmt = sfmt, dfmt
        self.skind, self.sdepth, shs, svs = _fmt(sfmt)
        self.dkind, self.ddepth, dhs, dvs = _fmt(dfmt)
        dst_rgb = self.dkind.startswith("rgb")
        src_rgb = self.skind.startswith("rgb")
        if self.skind in ("rgb16", "rgb565"):
            raise NotImplementedError("16-bit and 15/16 bpp RGB sources are not restated")
        # formats that are neither YUV nor gray carry no range, utils.c:844-880
        self.src_range, self.dst_range = (0 if src_rgb else src_range), (0 if dst_rgb else dst_range)
        src_range = self.src_range
        scaler = flags & SCALER_MASK
        if not scaler:                                                # :1209
            scaler = SWS_BICUBIC
            flags |= scaler
        assert scaler & (scaler - 1) == 0
        if scaler == SWS_FAST_BILINEAR and (sw < 8 or dw <= 8):       # :1224-1230
            scaler = SWS_BILINEAR
            flags ^= SWS_FAST_BILINEAR | SWS_BILINEAR
        lum_scaler = SWS_BICUBIC if scaler == SWS_BICUBLIN else scaler
        chr_scaler = SWS_BILINEAR if scaler == SWS_BICUBLIN else scaler
        unscaled = sw == dw and sh == dh
        # SwsFilter: dicts of vectors {"lumH": [...], "lumV": ..., "chrH": ..., "chrV": ...}; any vector longer than
        # one tap rules the unscaled special converters out (utils.c:1256-1263,1624)
        sfv, dfv = src_filter or {}, dst_filter or {}
        if any(v is not None and len(v) > 1 for v in list(sfv.values()) + list(dfv.values())):
            unscaled = False
        if dst_rgb and not (flags & SWS_FULL_CHR_H_INT):              # :1270-1286
            if dw & 1 or (shs == 0 and svs == 0 and dither != 2 and not flags & SWS_FAST_BILINEAR):
                flags |= SWS_FULL_CHR_H_INT
        if self.dkind == "rgbpf":                                     # planar RGB: full chroma forced, :1298-1306
            flags |= SWS_FULL_CHR_H_INT
        if self.dkind == "rgb565":                                    # no full-chroma writer, :1329-1357
            flags &= ~SWS_FULL_CHR_H_INT
        if dst_rgb and not (flags & SWS_FULL_CHR_H_INT):              # :1359
            dhs = 1
        # packed RGB sources: chroma from summed pixel pairs (the *_half readers), utils.c:1367-1390
        if src_rgb and not (sw & 1) and not (flags & SWS_FULL_CHR_H_INP) and \
                ((dw >> dhs) <= (sw >> 1) or flags & SWS_FAST_BILINEAR):
            shs = 1
        self.flags = flags
        self.shs, self.svs, self.dhs, self.dvs = shs, svs, dhs, dvs
        self.csw, self.csh = _cdiv_shift(sw, shs), _cdiv_shift(sh, svs)
        self.cdw, self.cdh = _cdiv_shift(dw, dhs), _cdiv_shift(dh, dvs)
        self.src_bpc, self.dst_bpc = max(self.sdepth, 8), max(self.ddepth, 8)
        if src_rgb:
            self.src_bpc = 16                                          # utils.c:1407-1408
        lum_xinc = ((sw << 16) + (dw >> 1)) // dw
        lum_yinc = ((sh << 16) + (dh >> 1)) // dh
        chr_xinc = ((self.csw << 16) + (self.cdw >> 1)) // self.cdw
        chr_yinc = ((self.csh << 16) + (self.cdh >> 1)) // self.cdh
        cs = colorspace or (5, src_range, 5, dst_range, 0, 1 << 16, 1 << 16)
        self.rgb = None
        if dst_rgb:
            self.rgb = yuv2rgb_tables(YUV2RGB_COEFFS[cs[0]], (0 if src_rgb else cs[1]) if colorspace else src_range,
                                      cs[4], cs[5], cs[6])
        if src_rgb:
            self.rgb2yuv = rgb2yuv_table(YUV2RGB_COEFFS[cs[2]])
        # `colorspace` restates a sws_setColorspaceDetails() call made AFTER init (utils.c:849-905): the special
        # converter below was chosen with the init-time ranges and stays; everything computed per frame sees the
        # new ranges (formats that are neither YUV nor gray carry none)
        sel_ranges = (self.src_range, self.dst_range)
        if colorspace and not src_rgb and not dst_rgb and YUV2RGB_COEFFS[cs[0]] != YUV2RGB_COEFFS[cs[2]]:
            raise NotImplementedError("YUV->YUV with two matrices cascades through RGB (utils.c:915-984): not restated")
        if colorspace:
            self.src_range = 0 if src_rgb else cs[1]
            self.dst_range = 0 if dst_rgb else cs[3]

        # unscaled special converters, swscale_unscaled.c:2392-2731 (only the ones that differ)
        self.unscaled_lut = False
        self.special = None
        if unscaled and (sel_ranges[0] == sel_ranges[1] or dst_rgb):
            if src_rgb and dst_rgb and self.dkind not in ("rgb16", "rgb565", "rgbpf"):
                # rgbToRgbWrapper (findRgbConvFn, swscale_unscaled.c:1843-2060,2463-2466), packedCopyWrapper
                # for identical formats; with SWS_BITEXACT 24 -> rgba/bgra is left to the scaler (:1992-1996)
                s32, d32 = self.skind == "rgb32", self.dkind == "rgb32"
                if sfmt == dfmt or s32 or not d32 or not (flags & SWS_BITEXACT) or dfmt in ("argb", "abgr"):
                    self.special = "shuffle"
                    return
            if src_rgb and self.dkind == "rgb565" and flags & (SWS_FAST_BILINEAR | SWS_POINT):
                # rgb24to16 / rgb32tobgr15 ... (findRgbConvFn): truncation, only when the scaler flag says that no
                # dither is wanted (swscale_unscaled.c:2459-2466)
                self.special = "rgb16pack"
                return
            if sfmt == "bgr24" and dfmt == "yuv420p" and not (flags & SWS_ACCURATE_RND) and not (dw & 1):
                self.special = "bgr24_yv12"                            # swscale_unscaled.c:2062-2077,2453-2457
                return
            if sfmt in ("yuv420p", "yuv422p") and dst_rgb and not (flags & SWS_ACCURATE_RND) \
                    and dither in (1, 2) and not (dh & 1):
                self.unscaled_lut = True                               # yuv2rgb_c_* (yuv2rgb.c:137-236)
                return
            if (not src_rgb and not dst_rgb and self.sdepth == 8 and self.ddepth == 8 and
                    (((shs, svs) == (1, 1) == (dhs, dvs) and (self.skind == "semi") != (self.dkind == "semi")) or
                     ((shs, svs) == (dhs, dvs) and self.skind == self.dkind and (sfmt == "nv21") == (dfmt == "nv21")))):
                # planarCopyWrapper / planarToNv12Wrapper / nv12ToPlanarWrapper (swscale_unscaled.c:147-215,2405-2420,
                # 2675-2700): plain copies, whatever the chroma siting options say
                self.special = "copy8"
                return
            if dfmt == "p010le" and self.skind == "planar" and (shs, svs) == (1, 1) and self.sdepth != 9:
                self.special = "p01x"                                  # swscale_unscaled.c:273-375,2432-2444
                return
            if (not dst_rgb and self.skind in ("planar", "semi") and self.dkind in ("planar", "semi")
                    and (shs, svs) == (dhs, dvs) and (self.skind == "semi") == (self.dkind == "semi")
                    and (self.sdepth != self.ddepth or self.sdepth > 8)):
                if self.skind == "semi" and ((self.sdepth != 8 and self.sdepth != self.ddepth)
                                             or (sfmt == "nv21") != (dfmt == "nv21")):
                    raise NotImplementedError("p010 -> 8-bit semi-planar copies are not restated")
                self.special = "depthcopy"                             # planarCopyWrapper, swscale_unscaled.c:2220-2384
                self.dither = dither
                return
        if self.skind == "rgb32" and self.dkind == "rgb32":
            raise NotImplementedError("the alpha plane (alpToYV12 -> yuv2packedX with alpha) is not restated")
        self.full_chr = bool(dst_rgb and flags & SWS_FULL_CHR_H_INT)

        def lpos(sub, pos):                                            # get_local_pos, utils.c:168
            if pos == -1 or pos <= -513:
                pos = (128 << sub) - 128
            return (pos + 128) >> sub

        shp, svp, dhp, dvp = chr_pos
        self.h_lum = init_filter(lum_xinc, sw, dw, 1 << 14, lum_scaler, flags, param, lpos(0, 0), lpos(0, 0),
                                 src_vec=sfv.get("lumH"), dst_vec=dfv.get("lumH"))
        self.h_chr = init_filter(chr_xinc, self.csw, self.cdw, 1 << 14, chr_scaler, flags, param,
                                 lpos(shs, shp), lpos(dhs, dhp), src_vec=sfv.get("chrH"), dst_vec=dfv.get("chrH"))
        self.v_lum = init_filter(lum_yinc, sh, dh, 1 << 12, lum_scaler, flags, param, lpos(0, 0), lpos(0, 0),
                                 src_vec=sfv.get("lumV"), dst_vec=dfv.get("lumV"))
        self.v_chr = init_filter(chr_yinc, self.csh, self.cdh, 1 << 12, chr_scaler, flags, param,
                                 lpos(svs, svp), lpos(dvs, dvp), src_vec=sfv.get("chrV"), dst_vec=dfv.get("chrV"))
        if None in (self.h_lum, self.h_chr, self.v_lum, self.v_chr):
            raise NotImplementedError("cascaded contexts are not restated")
        # one-tap vertical filters go through yuv2plane1 / yuv2packed1, which ignore the coefficient
        # (vscale.c:135-143,296-316); initFilter's edge fix-up can leave 4095 there
        vl, vc = self.v_lum[0], self.v_chr[0]
        if vl.shape[1] == 1:
            one = np.ones(vl.shape[0], bool)
            if dst_rgb and vc.shape[1] == 2:
                c0, c1 = vc[:vl.shape[0], 0].astype(np.int64), vc[:vl.shape[0], 1].astype(np.int64)
                one = (c0 + c1 == 4096) & (c1 >= 0) & (c1 <= 4096)
            elif dst_rgb and vc.shape[1] != 1:
                one[:] = False
            vl[one, 0] = 4096
        if vc.shape[1] == 1 and ((vl.shape[1] == 1) if dst_rgb else self.dkind != "semi"):
            vc[:, 0] = 4096
        # SWS_FAST_BILINEAR on 8-bit sources with <= 14-bit destinations: hyscale_fast / hcscale_fast
        # replace the horizontal FIR (swscale.c:675-681, hscale.c:54,188)
        self.fast_h = bool(flags & SWS_FAST_BILINEAR) and self.src_bpc == 8 and self.dst_bpc <= 14
        self.lum_xinc, self.chr_xinc = lum_xinc, chr_xinc

    # ---- stage 0: planes as integer sample arrays (input.c:926-941 for nv12/nv21)
    def _unpack(self, planes):
        if self.skind.startswith("rgb"):
            return self._unpack_rgb(planes[0])
        dt = np.uint8 if self.sdepth == 8 else np.dtype("<u2")
        lum = np.ascontiguousarray(planes[0]).view(dt)[:self.sh, :self.sw].astype(np.int64)
        if self.skind == "semi":
            uv = np.ascontiguousarray(planes[1]).view(dt)[:self.csh, :2 * self.csw].astype(np.int64)
            a, b = uv[:, 0::2], uv[:, 1::2]
            u, v = (b, a) if self.sfmt == "nv21" else (a, b)
            if self.sfmt == "p010le":                 # p010LEToY_c / p010LEToUV_c (input.c:950-1006): >> 6
                lum, u, v = lum >> 6, u >> 6, v >> 6
        else:
            u = np.ascontiguousarray(planes[1]).view(dt)[:self.csh, :self.csw].astype(np.int64)
            v = np.ascontiguousarray(planes[2]).view(dt)[:self.csh, :self.csw].astype(np.int64)
        return lum, u, v

    # packed 8-bit RGB readers: rgb24ToY_c / rgb24ToUV_c / rgb24ToUV_half_c (input.c:1068-1180) for
    # 3-byte pixels, the rgb16_32To{Y,UV,UV_half}_c_template instances (input.c:264-345,391-394) for
    # 4-byte pixels (coefficients << 8, unsigned rounding term, logical shift); int16 lines that the
    # horizontal scaler reads back as uint16
    def _unpack_rgb(self, plane):
        bpp = 3 if self.skind == "rgb8" else 4
        ro, go, bo = {"rgb24": (0, 1, 2), "bgr24": (2, 1, 0), "rgba": (0, 1, 2), "bgra": (2, 1, 0),
                      "argb": (1, 2, 3), "abgr": (3, 2, 1)}[self.sfmt]
        px = np.ascontiguousarray(plane)[:self.sh, :self.sw * bpp].astype(np.int64).reshape(self.sh, self.sw, bpp)
        r, g, b = px[:, :, ro], px[:, :, go], px[:, :, bo]
        ry, gy, by, ru, gu, bu, rv, gv, bv = self.rgb2yuv
        half = self.shs == 1

        def store(x):                       # int16 store, uint16 load
            return np.asarray(x, np.int64) & 0xFFFF

        if bpp == 3:
            lum = store(_wrap32(ry * r + gy * g + by * b + (32 << 14) + (1 << 8)) >> 9)
        else:
            lum = store(((ry * r + gy * g + by * b) * 256 + (32 << 22) + (1 << 16) & 0xFFFFFFFF) >> 17)
        if half:
            r, g, b = r[:, 0::2] + r[:, 1::2], g[:, 0::2] + g[:, 1::2], b[:, 0::2] + b[:, 1::2]
        su, sv = ru * r + gu * g + bu * b, rv * r + gv * g + bv * b
        if bpp == 3:
            rnd, sh = ((256 << 15) + (1 << 9), 10) if half else ((256 << 14) + (1 << 8), 9)
            u, v = store(_wrap32(su + rnd) >> sh), store(_wrap32(sv + rnd) >> sh)
        else:
            rnd, sh = ((256 << 23) + (1 << 17), 18) if half else ((256 << 22) + (1 << 16), 17)
            u, v = store((su * 256 + rnd & 0xFFFFFFFF) >> sh), store((sv * 256 + rnd & 0xFFFFFFFF) >> sh)
        return lum, u, v

    # ---- stage H: hScale8To15_c / hScale16To15_c / hScale8To19_c / hScale16To19_c (swscale.c:69-159)
    def _hscale(self, src, bank):
        coef, pos = bank
        inter19 = self.dst_bpc > 14
        if self.skind.startswith("rgb"):
            sh = 9 if inter19 else 13                                  # swscale.c:80-81,108-109
        elif self.src_bpc == 8:
            sh = 3 if inter19 else 7
        else:
            sh = self.sdepth - 1 - 4 if inter19 else self.sdepth - 1
        acc = np.zeros((src.shape[0], len(pos)), np.int64)
        for j in range(coef.shape[1]):
            idx = np.minimum(pos.astype(np.int64) + j, src.shape[1] - 1)
            acc += src[:, idx] * coef[:, j].astype(np.int64)[None, :]
        acc = _wrap32(acc) >> sh
        out = np.minimum(acc, (1 << 19) - 1 if inter19 else (1 << 15) - 1)
        return out if inter19 else out.astype(np.int16).astype(np.int64)

    # ---- ff_hyscale_fast_c / ff_hcscale_fast_c (hscale_fast_bilinear.c:23-55): 16.16 stepping, 7-bit
    # blend factor; the chroma variant weighs the left sample with (xalpha ^ 127), i.e. the taps sum to 127
    @staticmethod
    def _hscale_fast(src, dst_w, x_inc, chroma):
        src_w = src.shape[1]
        i = np.arange(dst_w, dtype=np.int64)
        xpos = (i * x_inc) & 0xFFFFFFFF                                # unsigned int xpos
        xx = xpos >> 16
        xa = (xpos & 0xFFFF) >> 9
        left, right = src[:, np.minimum(xx, src_w - 1)], src[:, np.minimum(xx + 1, src_w - 1)]
        if chroma:
            out = left * (xa ^ 127)[None, :] + right * xa[None, :]
        else:
            out = (left << 7) + (right - left) * xa[None, :]
        tail = ((i * x_inc) >> 16) >= src_w - 1                         # the fix-up loop at the right edge
        out[:, tail] = src[:, src_w - 1:src_w] * 128
        return out.astype(np.int16).astype(np.int64)

    # ---- range conversion on the h-scaled lines (swscale.c:163-255, constants :577-624)
    def _range(self, lum, u, v):
        if self.src_range == self.dst_range or self.dkind.startswith("rgb") or self.dst_bpc >= 32:
            return lum, u, v                          # swscale.c:610: no range conversion into float samples
        bd = min(self.dst_bpc, 16)
        src_bits = 15 if bd <= 14 else 19
        src_shift, mult_shift = src_bits - bd, (14 if bd <= 14 else 18)
        mpeg_min, mpeg_lum, mpeg_chr, jpeg_max = 16 << (bd - 8), 235 << (bd - 8), 240 << (bd - 8), (1 << bd) - 1

        def solve(smin, smax, dmin, dmax):
            total = mult_shift + src_shift
            coeff = _cdiv_shift(((dmax - dmin) << total) // (smax - smin), src_shift)
            offset = (dmax << total) - (smax << src_shift) * coeff + (1 << (mult_shift - 1))
            return coeff, offset

        if self.src_range:
            lc, lo = solve(0, jpeg_max, mpeg_min, mpeg_lum)
            cc, co = solve(0, jpeg_max, mpeg_min, mpeg_chr)
        else:
            lc, lo = solve(mpeg_min, mpeg_lum, 0, jpeg_max)
            cc, co = solve(mpeg_min, mpeg_chr, 0, jpeg_max)
        to_jpeg = not self.src_range
        if bd <= 14:
            lc &= 0xFFFF; cc &= 0xFFFF
            lo = int(_wrap32(lo)); co = int(_wrap32(co))

            def conv(x, c, o):
                y = _wrap32(x * c + o) >> 14
                if to_jpeg:
                    y = np.minimum(y, (1 << 15) - 1)
                return y.astype(np.int16).astype(np.int64)
        else:
            def conv(x, c, o):
                y = (x * c + o) >> 18
                y = _wrap32(y)
                if to_jpeg:
                    y = np.minimum(y, (1 << 19) - 1)
                return y
        return conv(lum, lc, lo), conv(u, cc, co), conv(v, cc, co)

    @staticmethod
    def _vsum(lines, bank, rows=None):
        """sum_j lines[pos[y]+j] * filter[y][j] for every output row, as 32-bit wrapped values."""
        coef, pos = bank
        n = len(pos) if rows is None else rows
        acc = np.zeros((n, lines.shape[1]), np.int64)
        for j in range(coef.shape[1]):
            idx = np.clip(pos[:n].astype(np.int64) + j, 0, lines.shape[0] - 1)
            acc += lines[idx] * coef[:n, j].astype(np.int64)[:, None]
        return acc

    # ---- the conversion
    _ORDER = {"rgb24": "rgb", "bgr24": "bgr", "rgba": "rgba", "bgra": "bgra", "argb": "argb", "abgr": "abgr"}

    def _rgb16pack(self, plane):
        """rgb24to16, rgb24tobgr15, rgb32to16 ... (rgb2rgb_template.c through rgbToRgbWrapper): channel truncation."""
        ro, go, bo, bpp = {"rgb24": (0, 1, 2, 3), "bgr24": (2, 1, 0, 3), "rgba": (0, 1, 2, 4), "bgra": (2, 1, 0, 4),
                           "argb": (1, 2, 3, 4), "abgr": (3, 2, 1, 4)}[self.sfmt]
        px = np.ascontiguousarray(plane)[:self.sh, :self.sw * bpp].reshape(self.sh, self.sw, bpp).astype(np.uint16)
        is565 = self.dfmt in ("rgb565le", "bgr565le")
        r, g, b = px[..., ro] >> 3, px[..., go] >> (2 if is565 else 3), px[..., bo] >> 3
        hi = 11 if is565 else 10
        out = (r << hi) | (g << 5) | b if self.dfmt.startswith("rgb") else (b << hi) | (g << 5) | r
        return [out.astype("<u2").view(np.uint8).reshape(self.sh, -1)]

    def _shuffle(self, plane):
        """Byte permutation per pixel; alpha carried when both sides have it, else 255."""
        so, do = self._ORDER[self.sfmt], self._ORDER[self.dfmt]
        px = np.ascontiguousarray(plane)[:self.sh, :self.sw * len(so)].reshape(self.sh, self.sw, len(so))
        out = np.empty((self.sh, self.sw, len(do)), np.uint8)
        for k, comp in enumerate(do):
            out[:, :, k] = px[:, :, so.index(comp)] if comp in so else 255
        return [out.reshape(self.sh, -1)]

    def _bgr24_yv12(self, plane):
        """ff_rgb24toyv12_c (rgb2rgb_template.c:580-641): truncating 15-bit matrix, 2x2 box chroma,
        an odd last row is its own partner."""
        px = np.ascontiguousarray(plane)[:self.sh, :self.sw * 3].astype(np.int64).reshape(self.sh, self.sw, 3)
        b, g, r = px[:, :, 0], px[:, :, 1], px[:, :, 2]
        ry, gy, by, ru, gu, bu, rv, gv, bv = self.rgb2yuv
        Y = (((ry * r + gy * g + by * b) >> 15) + 16) & 0xFF
        ra = np.arange(0, self.sh, 2)
        rb = np.minimum(ra + 1, self.sh - 1)

        def box(c):
            return (c[ra][:, 0::2] + c[ra][:, 1::2] + c[rb][:, 0::2] + c[rb][:, 1::2]) >> 2
        rx, gx, bx = box(r), box(g), box(b)
        U = (((ru * rx + gu * gx + bu * bx) >> 15) + 128) & 0xFF
        V = (((rv * rx + gv * gx + bv * bx) >> 15) + 128) & 0xFF
        return [Y.astype(np.uint8), U.astype(np.uint8), V.astype(np.uint8)]

    def _depthcopy(self, planes):
        """planarCopyWrapper between planar YUV depths (swscale_unscaled.c:2160-2218,2249-2346), little-endian
        formats: ordered dither down (DITHER_COPY, the matrix of level `shift`), shift or bit replication up.
        Chroma and limited-range luma are `shiftonly`."""
        sd, dd = self.sdepth, self.ddepth
        out = []
        shapes = [(self.sw, self.sh), (self.csw, self.csh), (self.csw, self.csh)]
        if self.skind == "semi":                       # nv12 -> p010: one interleaved plane of 2 cw samples per row
            shapes = [(self.sw, self.sh), (2 * self.csw, self.csh)]
        dst_shift = 6 if self.dfmt == "p010le" else 0
        for i, (w, h) in enumerate(shapes):
            sdt = np.uint8 if sd == 8 else np.dtype("<u2")
            v = np.ascontiguousarray(planes[i]).view(sdt)[:h, :w].astype(np.int64)
            shiftonly = i > 0 or not self.src_range
            if sd > dd:
                shift = sd - dd
                d = DEPTH_DITHER[shift - 1][(np.arange(h) & 7)[:, None], (np.arange(w) & 7)[None, :]]
                if self.dither == 0:                                   # SWS_DITHER_NONE
                    t = (v + (1 << (shift - 1))) >> shift
                    o = t - (t >> dd)
                elif shiftonly:
                    t = (v + d) >> shift
                    o = t - (t >> dd)
                else:
                    o = (v - (v >> dd) + d) >> shift
            else:
                shift = dd - sd
                src_shift = 6 if self.sfmt == "p010le" else 0
                raw, v = v, v >> src_shift
                o = (v << shift if shiftonly else (v << shift) | (v >> (2 * sd - dd))) << dst_shift
                if shiftonly and src_shift:
                    # the 64-bit fast loop (swscale_unscaled.c:2291-2296) shifts four samples as one word: only
                    # the first sample of every full group of four loses its low bits
                    j = np.arange(w)[None, :]
                    keep = ((j & 3) != 0) & ((j & ~3) + 3 < w)
                    o = np.where(keep, raw, o)
            out.append(o.astype(np.uint8) if dd == 8 else o.astype("<u2").view(np.uint8))
        return out

    def _p01x(self, planes):
        """planar8ToP01xleWrapper (<< 8) / planarToP01xWrapper (<< 16 - depth), swscale_unscaled.c:273-375:
        chroma interleaved U first, src_w / 2 pairs per row (an odd last column is left untouched)."""
        sdt = np.uint8 if self.sdepth == 8 else np.dtype("<u2")
        shift = 8 if self.sdepth == 8 else 16 - self.sdepth
        y = np.ascontiguousarray(planes[0]).view(sdt)[:self.sh, :self.sw].astype(np.int64)
        u = np.ascontiguousarray(planes[1]).view(sdt)[:self.csh, :self.csw].astype(np.int64)
        v = np.ascontiguousarray(planes[2]).view(sdt)[:self.csh, :self.csw].astype(np.int64)
        uv = np.zeros((self.cdh, 2 * self.cdw), np.int64)
        n = self.sw // 2
        uv[:, 0:2 * n:2], uv[:, 1:2 * n:2] = u[:, :n] << shift, v[:, :n] << shift
        return [((y << shift) & 0xFFFF).astype("<u2").view(np.uint8), (uv & 0xFFFF).astype("<u2").view(np.uint8)]

    def scale(self, planes):
        if self.special == "p01x":
            return self._p01x(planes)
        if self.special == "depthcopy":
            return self._depthcopy(planes)
        if self.special == "copy8":
            lum, u, v = self._unpack(planes)
            y8, u8, v8 = (np.asarray(a).astype(np.uint8) for a in (lum, u, v))
            if self.dkind == "semi":
                uv = np.zeros((self.cdh, 2 * self.cdw), np.uint8)
                a, b = (v8, u8) if self.dfmt == "nv21" else (u8, v8)
                uv[:, 0::2], uv[:, 1::2] = a[:self.cdh, :self.cdw], b[:self.cdh, :self.cdw]
                return [y8, uv]
            return [y8, u8[:self.cdh, :self.cdw], v8[:self.cdh, :self.cdw]]
        if self.special == "rgb16pack":
            return self._rgb16pack(planes[0])
        if self.special == "shuffle":
            return self._shuffle(planes[0])
        if self.special == "bgr24_yv12":
            return self._bgr24_yv12(planes[0])
        lum, u, v = self._unpack(planes)
        if self.unscaled_lut:
            return self._unscaled_lut(lum, u, v)
        if self.fast_h:
            hl = self._hscale_fast(lum, self.dw, self.lum_xinc, False)
            hu = self._hscale_fast(u, self.cdw, self.chr_xinc, True)
            hv = self._hscale_fast(v, self.cdw, self.chr_xinc, True)
        else:
            hl = self._hscale(lum, self.h_lum)
            hu = self._hscale(u, self.h_chr)
            hv = self._hscale(v, self.h_chr)
        hl, hu, hv = self._range(hl, hu, hv)
        if self.dkind in ("planar", "semi"):
            return self._planar_out(hl, hu, hv)
        if self.dkind == "grayf":
            return self._grayf_out(hl)
        if self.dkind == "rgbpf":
            return self._gbrpf32_out(hl, hu, hv)
        if self.full_chr:
            return self._rgb_full_out(hl, hu, hv)
        if self.dkind == "rgb16":
            return self._rgb16_out(hl, hu, hv)
        return self._rgb8_out(hl, hu, hv)

    # yuv2rgb_full_{X,1,2}_c_template + yuv2rgb_write_full (output.c:1998-2051,2160-2330);
    # yuv2rgba64_full_X_c_template (output.c:1373-1430); chooser vscale.c:135-163
    def _rgb_full_out(self, hl, hu, hv):
        t = self.rgb
        n, w = self.dh, self.dw
        lcoef, ccoef = self.v_lum[0], self.v_chr[0]
        Y = self._vsum(hl, self.v_lum)[:, :w]
        U = self._vsum(hu, self.v_chr, n)[:, :w]
        V = self._vsum(hv, self.v_chr, n)[:, :w]
        if self.dkind == "rgb16":
            Y = (_wrap32(Y - 0x40000000) >> 14) + 0x10000
            U = _wrap32(U - (128 << 23))
            V = _wrap32(V - (128 << 23))
            if lcoef.shape[1] == 1 and ccoef.shape[1] == 2:
                # yuv2rgba64_full_1_c_template with uvalpha != 0 (output.c:1537-1545) keeps U and V in SUINT, so its
                # >> 14 is a LOGICAL shift: chroma below 128 wraps.  Rows the chooser (vscale.c:138-143) sends there:
                c0, c1 = ccoef[:n, 0].astype(np.int64), ccoef[:n, 1].astype(np.int64)
                quirk = ((c0 + c1 == 4096) & (c1 > 0) & (c1 <= 4096))[:, None]
                U = np.where(quirk, _wrap32((U & 0xFFFFFFFF) >> 14), U >> 14)
                V = np.where(quirk, _wrap32((V & 0xFFFFFFFF) >> 14), V >> 14)
            else:
                U, V = U >> 14, V >> 14
            Y = _wrap32(_wrap32((Y - t["y_offset"]) * t["y_coeff"]) + (1 << 13) - (1 << 29))
            R = _wrap32(V * t["v2r"]); G = _wrap32(V * t["v2g"] + U * t["u2g"]); B = _wrap32(U * t["u2b"])
            comp = lambda c: np.clip((_wrap32(c + Y) >> 14) + (1 << 15), 0, 65535).astype("<u2")
            r, g, b = comp(R), comp(G), comp(B)
            out = np.zeros((n, w * 3), "<u2")
            a, c = (r, b) if self.dfmt == "rgb48le" else (b, r)
            out[:, 0::3], out[:, 1::3], out[:, 2::3] = a, g, c
            return [out.view(np.uint8).reshape(n, -1)]
        lb = np.full((n, 1), 1 << 9, np.int64)
        cb = np.full((n, 1), 1 << 9, np.int64)
        if ccoef.shape[1] == 2:
            c0, c1 = ccoef[:n, 0].astype(np.int64), ccoef[:n, 1].astype(np.int64)
            cok = (c0 + c1 == 4096) & (c1 >= 0) & (c1 <= 4096)
            if lcoef.shape[1] == 1:
                cb[cok, 0] = 0
            elif lcoef.shape[1] == 2:
                l0, l1 = lcoef[:n, 0].astype(np.int64), lcoef[:n, 1].astype(np.int64)
                both = cok & (l0 + l1 == 4096) & (l1 >= 0) & (l1 <= 4096)
                lb[both, 0] = 0
                cb[both, 0] = 0
        Y = _wrap32(Y + lb) >> 10
        U = _wrap32(U + cb - (128 << 19)) >> 10
        V = _wrap32(V + cb - (128 << 19)) >> 10
        Y = _wrap32(_wrap32((Y - t["y_offset"]) * t["y_coeff"]) + (1 << 21))
        clip30 = lambda x: np.clip(_wrap32(x), 0, (1 << 30) - 1) >> 22
        R = clip30(Y + V * t["v2r"]).astype(np.uint8)
        G = clip30(Y + V * t["v2g"] + U * t["u2g"]).astype(np.uint8)
        B = clip30(Y + U * t["u2b"]).astype(np.uint8)
        order = {"rgb24": (R, G, B), "bgr24": (B, G, R), "rgba": (R, G, B, None), "bgra": (B, G, R, None),
                 "argb": (None, R, G, B), "abgr": (None, B, G, R)}[self.dfmt]
        bpp = len(order)
        out = np.zeros((n, w * bpp), np.uint8)
        for k, comp in enumerate(order):
            out[:, k::bpp] = 255 if comp is None else comp
        return [out]

    # yuv2plane1_float_c_template / yuv2planeX_float_c_template (output.c:219-268): the 16-bit writers' integer
    # value times 1.0f / 65535.0f in single precision; one-tap filters go through yuv2plane1 (vscale.c:56-61)
    def _grayf_out(self, hl):
        coef, pos = self.v_lum
        if coef.shape[1] == 1:
            idx = np.clip(pos.astype(np.int64), 0, hl.shape[0] - 1)
            val = np.clip(_wrap32(hl[idx] + 4) >> 3, 0, 65535)
        else:
            val = _wrap32(self._vsum(hl, self.v_lum) + (1 << 14) - 0x40000000) >> 15
            val = 0x8000 + np.clip(val, -32768, 32767)
        out = np.float32(1.0) / np.float32(65535.0) * val[:, :self.dw].astype(np.float32)
        return [out.astype("<f4").view(np.uint8).reshape(out.shape[0], -1)]

    # yuv2gbrpf32_full_X_c (output.c:2536-2610): the 16-bit full-chroma arithmetic, planes in G, B, R order
    def _gbrpf32_out(self, hl, hu, hv):
        t = self.rgb
        n, w = self.dh, self.dw
        Y = (_wrap32(self._vsum(hl, self.v_lum)[:, :w] - 0x40000000) >> 14) + 0x10000
        U = _wrap32(self._vsum(hu, self.v_chr, n)[:, :w] - (128 << 23)) >> 14
        V = _wrap32(self._vsum(hv, self.v_chr, n)[:, :w] - (128 << 23)) >> 14
        Y = _wrap32(_wrap32((Y - t["y_offset"]) * t["y_coeff"]) + (1 << 13) - (1 << 29))
        R = _wrap32(V * t["v2r"]); G = _wrap32(V * t["v2g"] + U * t["u2g"]); B = _wrap32(U * t["u2b"])
        mult = np.float32(1.0) / np.float32(65535.0)

        def comp(c):
            v = np.clip((_wrap32(c + Y) >> 14) + (1 << 15), 0, 65535).astype(np.float32) * mult
            return v.astype("<f4").view(np.uint8).reshape(n, -1)
        return [comp(G), comp(B), comp(R)]

    # yuv2planeX_8_c / yuv2planeX_10_c / yuv2planeX_16_c / yuv2nv12cX_c (output.c:163-187,340-357,468-528)
    def _planar_out(self, hl, hu, hv):
        bits = self.ddepth
        yl = self._vsum(hl, self.v_lum)
        yu = self._vsum(hu, self.v_chr)
        yv = self._vsum(hv, self.v_chr)

        def dither_plane(shape, offset):
            h, w = shape
            if self.sdepth > 8:                       # isNBPS||is16BPS source, swscale.c:291-292,519-522
                return DITHER_8x8_128[(np.arange(h) & 7)[:, None], ((np.arange(w) + offset) & 7)[None, :]]
            return np.full(shape, 64, np.int64)

        def out(acc, offset):
            if bits == 8:
                val = _wrap32(acc + (dither_plane(acc.shape, offset) << 12)) >> 19
                return np.clip(val, 0, 255).astype(np.uint8)
            if bits < 16:
                shift = 11 + 16 - bits
                val = _wrap32(acc + (1 << (shift - 1))) >> shift
                return np.clip(val, 0, (1 << bits) - 1).astype("<u2")
            val = _wrap32(acc + (1 << 14) - 0x40000000) >> 15
            return (0x8000 + np.clip(val, -32768, 32767)).astype("<u2")

        py, pu, pv = out(yl, 0), out(yu, 0), out(yv, 3)
        if self.dfmt == "p010le":                     # yuv2p010l1/lX/cX_c (output.c:538-589): 10 bits << 6
            py, pu, pv = (py << 6).astype("<u2"), (pu << 6).astype("<u2"), (pv << 6).astype("<u2")
        if self.dkind == "semi":
            uv = np.zeros((pu.shape[0], pu.shape[1] * 2), pu.dtype)
            a, b = (pv, pu) if self.dfmt == "nv21" else (pu, pv)
            uv[:, 0::2], uv[:, 1::2] = a, b
            return [py.view(np.uint8), uv.view(np.uint8)]
        return [py.view(np.uint8).reshape(py.shape[0], -1), pu.view(np.uint8).reshape(pu.shape[0], -1),
                pv.view(np.uint8).reshape(pv.shape[0], -1)]

    # yuv2rgb_X_c_template / _2_c_template + yuv2rgb_write (output.c:1662-1880); chooser vscale.c:135-169
    def _rgb8_out(self, hl, hu, hv):
        lcoef, _ = self.v_lum
        ccoef, _ = self.v_chr
        n = self.dh
        Y = self._vsum(hl, self.v_lum)
        U = self._vsum(hu, self.v_chr, n)
        V = self._vsum(hv, self.v_chr, n)
        bias = np.full((n, 1), 1 << 18, np.int64)
        if lcoef.shape[1] == 2 and ccoef.shape[1] == 2:        # yuv2packed2: no rounding bias
            l0, l1 = lcoef[:n, 0].astype(np.int64), lcoef[:n, 1].astype(np.int64)
            c0, c1 = ccoef[:n, 0].astype(np.int64), ccoef[:n, 1].astype(np.int64)
            two = (l0 + l1 == 4096) & (l1 >= 0) & (l1 <= 4096) & (c0 + c1 == 4096) & (c1 >= 0) & (c1 <= 4096)
            bias[two, 0] = 0
        Y = _wrap32(Y + bias) >> 19
        U = _wrap32(U + bias) >> 19
        V = _wrap32(V + bias) >> 19
        return [self._write_rgb8(Y, U, V)]

    def _write_rgb8(self, Y, U, V):
        t = self.rgb
        pairs = self.dw >> 1 if self.unscaled_lut else (self.dw + 1) >> 1
        U, V = U[:, :pairs], V[:, :pairs]
        r_idx = t["rV"][V + 512]
        g_idx = t["gU"][U + 512] + t["gV"][V + 512]
        b_idx = t["bU"][U + 512]
        rows = Y.shape[0]
        px = min(pairs * 2, self.dw)       # 15/16 bpp, odd width: the second pixel of the last pair falls off the row
        Yp = Y[:, :px]
        rep = lambda a: np.repeat(a, 2, axis=1)[:, :px]
        R = t["y_table"][Yp + rep(r_idx)]
        G = t["y_table"][Yp + rep(g_idx)]
        B = t["y_table"][Yp + rep(b_idx)]
        if self.dkind == "rgb565":
            # yuv2rgb_write 15/16 bpp (output.c:1714-1747), tables yuv2rgb.c:878-900; the unscaled converters
            # (yuv2rgb.c:371-398) index the same 2x2 tiles by row and column parity
            d8 = np.array([[6, 2], [0, 4]], np.int64)                  # ff_dither_2x2_8, output.c:46-50
            d4 = np.array([[1, 3], [2, 0]], np.int64)                  # ff_dither_2x2_4, output.c:40-44
            yy = (np.arange(rows) & 1)[:, None]
            xx = (np.arange(px) & 1)[None, :]
            is565 = self.dfmt in ("rgb565le", "bgr565le")
            dr = d8[yy, xx]
            db = d8[yy ^ 1, xx]
            dg = d4[yy, xx] if is565 else d8[yy, xx ^ 1]
            R = t["y_table"][Yp + rep(r_idx) + dr].astype(np.uint16) >> 3
            G = t["y_table"][Yp + rep(g_idx) + dg].astype(np.uint16) >> (2 if is565 else 3)
            B = t["y_table"][Yp + rep(b_idx) + db].astype(np.uint16) >> 3
            hi = 11 if is565 else 10
            pix = (R << hi) | (G << 5) | B if self.dfmt.startswith("rgb") else (B << hi) | (G << 5) | R
            out = np.zeros((rows, self.dw), "<u2")
            out[:, :px] = pix
            return out.view(np.uint8).reshape(rows, -1)
        order = {"rgb24": (R, G, B), "bgr24": (B, G, R), "rgba": (R, G, B, None), "bgra": (B, G, R, None),
                 "argb": (None, R, G, B), "abgr": (None, B, G, R),
                 "rgb48le": (R, G, B), "bgr48le": (B, G, R)}[self.dfmt]
        bpp = len(order) * (2 if self.dkind == "rgb16" else 1)
        out = np.zeros((rows, self.dw * bpp), np.uint8)
        for k, comp in enumerate(order):
            val = np.full((rows, px), 255, np.uint8) if comp is None else comp
            if self.dkind == "rgb16":                 # PUTRGB48 (yuv2rgb.c:107-115): byte replicated
                out[:, 2 * k:px * bpp:bpp] = val
                out[:, 2 * k + 1:px * bpp:bpp] = val
            else:
                out[:, k:px * bpp:bpp] = val
        return out

    # yuv2rgba64_X_c_template, hasAlpha=0, eightbytes=0 (output.c:1115-1196)
    def _rgb16_out(self, hl, hu, hv):
        t = self.rgb
        n = self.dh
        Y = _wrap32(self._vsum(hl, self.v_lum) - 0x40000000)
        U = _wrap32(self._vsum(hu, self.v_chr, n) - (128 << 23))
        V = _wrap32(self._vsum(hv, self.v_chr, n) - (128 << 23))
        Y = (Y >> 14) + 0x10000
        U = U >> 14
        V = V >> 14
        Y = _wrap32(_wrap32((Y - t["y_offset"]) * t["y_coeff"]) + (1 << 13) - (1 << 29))
        R = _wrap32(V * t["v2r"])
        G = _wrap32(V * t["v2g"] + U * t["u2g"])
        B = _wrap32(U * t["u2b"])
        pairs = (self.dw + 1) >> 1
        rep = lambda a: np.repeat(a[:, :pairs], 2, axis=1)[:, :self.dw]
        comp = lambda c: np.clip((_wrap32(rep(c) + Y[:, :self.dw]) >> 14) + (1 << 15), 0, 65535).astype("<u2")
        r, g, b = comp(R), comp(G), comp(B)
        out = np.zeros((n, self.dw * 3), "<u2")
        a, c = (r, b) if self.dfmt == "rgb48le" else (b, r)
        out[:, 0::3], out[:, 1::3], out[:, 2::3] = a, g, c
        return [out.view(np.uint8).reshape(n, -1)]

    # yuv2rgb_c_24_rgb & friends: nearest chroma, byte LUTs (yuv2rgb.c:68-236,520-559)
    def _unscaled_lut(self, lum, u, v):
        rows = np.arange(self.dh) >> self.svs
        U = np.repeat(u[rows], 1, axis=0)
        V = v[rows]
        return [self._write_rgb8(lum, U, V)]


def convert(sw, sh, sfmt, dw, dh, dfmt, flags, planes, **kw):
    return OracleContext(sw, sh, sfmt, dw, dh, dfmt, flags, **kw).scale(planes)
