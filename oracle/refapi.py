"""ctypes binding for oracle/_ref/libswsref.so (the REAL reference, built by build_ref.py).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(librempeg_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libswsref.so")

# SwsFlags (reference libswscale/swscale.h:131-208)
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

_lib = None


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO_PATH, mode=os.RTLD_LOCAL)
        L.swsref_pix_fmt.restype = C.c_int
        L.swsref_pix_fmt.argtypes = [C.c_char_p]
        L.swsref_create.restype = C.c_void_p
        L.swsref_create.argtypes = [C.c_int] * 6 + [C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.swsref_free.argtypes = [C.c_void_p]
        L.swsref_set_colorspace.restype = C.c_int
        L.swsref_set_colorspace.argtypes = [C.c_void_p] + [C.c_int] * 7
        L.swsref_scale.restype = C.c_int
        L.swsref_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.swsref_frame_alloc.restype = C.c_void_p
        L.swsref_frame_alloc.argtypes = [C.c_int] * 3
        L.swsref_frame_free.argtypes = [C.c_void_p]
        L.swsref_frame_data.restype = C.c_void_p
        L.swsref_frame_data.argtypes = [C.c_void_p, C.c_int]
        L.swsref_frame_linesize.restype = C.c_int
        L.swsref_frame_linesize.argtypes = [C.c_void_p, C.c_int]
        L.swsref_scale_frame.restype = C.c_int
        L.swsref_scale_frame.argtypes = [C.c_void_p] * 3
        L.swsref_bench_frame.restype = C.c_double
        L.swsref_bench_frame.argtypes = [C.c_void_p] * 3 + [C.c_int]
        L.swsref_filter.restype = C.c_int
        L.swsref_filter.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_int16)),
                                    C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int)]
        L.swsref_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        if hasattr(L, "swsref_rgb2yuv"):
            L.swsref_rgb2yuv.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.swsref_rgb_tables.restype = C.c_int
        L.swsref_rgb_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.swsref_lfg_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_int]
        L.swsref_md5.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.swsref_quiet.argtypes = [C.c_int]
        L.swsref_quiet(16)  # AV_LOG_ERROR
        _lib = L
    return _lib


def pix_fmt(name):
    v = lib().swsref_pix_fmt(name.encode())
    if v < 0:
        raise ValueError("unknown pixel format " + name)
    return v


def lfg_fill(arr, seed, bits=8):
    """Fill a contiguous uint8/uint16 numpy array exactly like SURVEY.md App. B."""
    assert arr.flags["C_CONTIGUOUS"]
    lib().swsref_lfg_fill(arr.ctypes.data, arr.nbytes, seed, bits)
    return arr


def md5(arr):
    out = (C.c_uint8 * 16)()
    a = np.ascontiguousarray(arr)
    lib().swsref_md5(out, a.ctypes.data, a.nbytes)
    return bytes(out).hex()


class RefContext:
    """A reference SwsContext (legacy API, sws_init_context'd)."""

    def __init__(self, sw, sh, sfmt, dw, dh, dfmt, flags, param=None, threads=1,
                 src_range=0, dst_range=0, chr_pos=(-513, -513, -513, -513), dither=1,
                 src_filter=None, dst_filter=None):
        self.sw, self.sh, self.dw, self.dh = sw, sh, dw, dh
        self.sfmt = pix_fmt(sfmt) if isinstance(sfmt, str) else sfmt
        self.dfmt = pix_fmt(dfmt) if isinstance(dfmt, str) else dfmt
        p = None
        if param is not None:
            p = (C.c_double * 2)(*param)
        opts = (C.c_int * 8)(threads, src_range, dst_range, chr_pos[0], chr_pos[1], chr_pos[2], chr_pos[3], dither)
        if src_filter or dst_filter:
            # SwsFilter vectors: dicts {"lumH": [...], "lumV": ..., "chrH": ..., "chrV": ...}
            L = lib()
            L.swsref_create_filtered.restype = C.c_void_p
            keep, ptrs, lens = [], (C.POINTER(C.c_double) * 8)(), (C.c_int * 8)()
            for fi, f in enumerate((src_filter or {}, dst_filter or {})):
                for ki, k in enumerate(("lumH", "lumV", "chrH", "chrV")):
                    v = f.get(k)
                    if v is not None:
                        arr = (C.c_double * len(v))(*v)
                        keep.append(arr)
                        ptrs[4 * fi + ki] = C.cast(arr, C.POINTER(C.c_double))
                        lens[4 * fi + ki] = len(v)
            L.swsref_create_filtered.argtypes = [C.c_int] * 6 + [C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            self.h = L.swsref_create_filtered(sw, sh, self.sfmt, dw, dh, self.dfmt, flags, p, opts, ptrs, lens)
        else:
            self.h = lib().swsref_create(sw, sh, self.sfmt, dw, dh, self.dfmt, flags, p, opts)
        if not self.h:
            raise RuntimeError("reference sws_init_context failed")

    def close(self):
        if self.h:
            lib().swsref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_colorspace(self, src_cs, src_range, dst_cs, dst_range, brightness=0, contrast=1 << 16,
                       saturation=1 << 16):
        return lib().swsref_set_colorspace(self.h, src_cs, src_range, dst_cs, dst_range,
                                           brightness, contrast, saturation)

    def scale(self, src_planes, src_strides, dst_planes, dst_strides, y=0, h=None):
        """planes: lists of numpy arrays (or None); strides in bytes."""
        h = self.sh if h is None else h
        sp = (C.c_void_p * 4)(*[(p.ctypes.data if p is not None else None) for p in _pad4(src_planes)])
        ss = (C.c_int * 4)(*_pad4i(src_strides))
        dp = (C.c_void_p * 4)(*[(p.ctypes.data if p is not None else None) for p in _pad4(dst_planes)])
        ds = (C.c_int * 4)(*_pad4i(dst_strides))
        return lib().swsref_scale(self.h, sp, ss, y, h, dp, ds)

    def filter(self, which):
        coef = C.POINTER(C.c_int16)()
        pos = C.POINTER(C.c_int32)()
        n = C.c_int()
        fs = lib().swsref_filter(self.h, which, C.byref(coef), C.byref(pos), C.byref(n))
        if fs <= 0:
            return None
        co = np.ctypeslib.as_array(coef, shape=(n.value * fs,)).copy().reshape(n.value, fs)
        po = np.ctypeslib.as_array(pos, shape=(n.value,)).copy()
        return co, po

    def info(self):
        out = (C.c_int * 16)()
        lib().swsref_info(self.h, out)
        keys = ["y_offset", "y_coeff", "v2r", "v2g", "u2g", "u2b", "unscaled", "cascaded",
                "chrSrcW", "chrSrcH", "chrDstW", "chrDstH", "srcBpc", "dstBpc", "flags"]
        return dict(zip(keys, list(out)))

    def rgb2yuv(self):
        out = (C.c_int * 9)()
        lib().swsref_rgb2yuv(self.h, out)
        return list(out)

    def rgb_tables(self):
        y = np.zeros(2048, np.uint8)
        t = [np.zeros(1280, np.int32) for _ in range(4)]
        r = lib().swsref_rgb_tables(self.h, y.ctypes.data, *[a.ctypes.data for a in t])
        if r < 0:
            return None
        return y, t[0], t[1], t[2], t[3]


def _pad4(lst):
    lst = list(lst)
    return lst + [None] * (4 - len(lst))


def _pad4i(lst):
    lst = [int(x) for x in lst]
    return lst + [0] * (4 - len(lst))


class RefFrame:
    """Refcounted AVFrame owned by the reference (for the threaded sws_scale_frame path)."""

    def __init__(self, w, h, fmt):
        self.fmt = pix_fmt(fmt) if isinstance(fmt, str) else fmt
        self.w, self.h = w, h
        self.f = lib().swsref_frame_alloc(w, h, self.fmt)
        if not self.f:
            raise MemoryError("av_frame_get_buffer failed")

    def plane(self, i, rows, dtype=np.uint8):
        ptr = lib().swsref_frame_data(self.f, i)
        ls = lib().swsref_frame_linesize(self.f, i)
        buf = (C.c_uint8 * (ls * rows)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint8).reshape(rows, ls), ls

    def close(self):
        if self.f:
            lib().swsref_frame_free(self.f)
            self.f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
