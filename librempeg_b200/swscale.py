"""ctypes binding of libswscale_b200.so -- the thin Python mirror of the C ABI in
include/swscale_b200.h / swscale_b200_cuda.h, used by tests/ and bench.py.

Nothing is computed here: every pixel goes through the C-ABI calls into the
hand-written sm_100a kernels.  If the shared object is missing the import
raises (no fallback of any kind).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWS_B200_LIB: another build of the same library (A/B runs of tools/*.sh)
SO_PATH = os.environ.get("SWS_B200_LIB") or os.path.join(_HERE, "libswscale_b200.so")

# enum AVPixelFormat (values are ABI; reference libavutil/pixfmt.h)
PIX_FMT = {
    "yuv420p": 0, "rgb24": 2, "bgr24": 3, "yuv422p": 4, "yuv444p": 5, "gray": 8,
    "yuvj420p": 12, "yuvj422p": 13, "yuvj444p": 14, "nv12": 23, "nv21": 24,
    "argb": 25, "rgba": 26, "abgr": 27, "bgra": 28, "rgb48le": 35,
    "rgb565le": 37, "rgb555le": 39, "bgr565le": 41, "bgr555le": 43,
    "yuv420p16le": 45, "yuv422p16le": 47, "yuv444p16le": 49, "bgr48le": 58,
    "yuv420p9le": 60, "yuv420p10le": 62, "yuv422p10le": 64, "yuv444p9le": 66,
    "yuv444p10le": 68, "yuv422p9le": 70, "yuv420p12le": 123, "yuv420p14le": 125,
    "yuv422p12le": 127, "yuv422p14le": 129, "yuv444p12le": 131, "yuv444p14le": 133, "p010le": 158,
    "gbrpf32le": 175, "grayf32le": 183,
}

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
SWS_PRINT_INFO = 1 << 12
SWS_FULL_CHR_H_INT = 1 << 13
SWS_FULL_CHR_H_INP = 1 << 14
SWS_ACCURATE_RND = 1 << 18
SWS_BITEXACT = 1 << 19
BX = SWS_ACCURATE_RND | SWS_BITEXACT


class SwsContextStruct(C.Structure):
    """Public SwsContext fields, ABI order (reference swscale.h:227-315)."""
    _fields_ = [
        ("av_class", C.c_void_p), ("opaque", C.c_void_p), ("flags", C.c_uint),
        ("scaler_params", C.c_double * 2), ("threads", C.c_int), ("dither", C.c_int),
        ("alpha_blend", C.c_int), ("gamma_flag", C.c_int),
        ("src_w", C.c_int), ("src_h", C.c_int), ("dst_w", C.c_int), ("dst_h", C.c_int),
        ("src_format", C.c_int), ("dst_format", C.c_int),
        ("src_range", C.c_int), ("dst_range", C.c_int),
        ("src_v_chr_pos", C.c_int), ("src_h_chr_pos", C.c_int),
        ("dst_v_chr_pos", C.c_int), ("dst_h_chr_pos", C.c_int),
        ("intent", C.c_int), ("scaler", C.c_int), ("scaler_sub", C.c_int), ("backends", C.c_int),
    ]


_lib = None


class SwsVectorStruct(C.Structure):
    _fields_ = [("coeff", C.POINTER(C.c_double)), ("length", C.c_int)]


class SwsFilterStruct(C.Structure):
    _fields_ = [("lumH", C.POINTER(SwsVectorStruct)), ("lumV", C.POINTER(SwsVectorStruct)),
                ("chrH", C.POINTER(SwsVectorStruct)), ("chrV", C.POINTER(SwsVectorStruct))]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("libswscale_b200.so is not built (run `python -c 'import __graft_entry__ as g; "
                          "g.build()'`); there is no fallback path")
    L = C.CDLL(SO_PATH, mode=os.RTLD_LOCAL)
    P = C.POINTER
    ctxp = P(SwsContextStruct)
    L.sws_alloc_context.restype = ctxp
    L.sws_init_context.restype = C.c_int
    L.sws_init_context.argtypes = [ctxp, C.c_void_p, C.c_void_p]
    L.sws_freeContext.argtypes = [ctxp]
    L.sws_getContext.restype = ctxp
    L.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p, P(C.c_double)]
    L.sws_scale.restype = C.c_int
    L.sws_scale.argtypes = [ctxp, P(C.c_void_p), P(C.c_int), C.c_int, C.c_int, P(C.c_void_p), P(C.c_int)]
    L.sws_getCoefficients.restype = P(C.c_int)
    L.sws_getCoefficients.argtypes = [C.c_int]
    L.sws_setColorspaceDetails.restype = C.c_int
    L.sws_setColorspaceDetails.argtypes = [ctxp, P(C.c_int), C.c_int, P(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
    L.sws_isSupportedInput.argtypes = [C.c_int]
    L.sws_isSupportedOutput.argtypes = [C.c_int]
    L.sws_cuda_device_count.restype = C.c_int
    L.sws_cuda_scale_batch.restype = C.c_int
    L.sws_cuda_scale_batch.argtypes = [ctxp, P(C.c_void_p), P(C.c_int), P(C.c_int64),
                                       P(C.c_void_p), P(C.c_int), P(C.c_int64), C.c_int]
    L.sws_cuda_scale_batch_host.restype = C.c_int
    L.sws_cuda_scale_batch_host.argtypes = [ctxp, P(C.c_void_p), P(C.c_int), P(C.c_int64),
                                            P(C.c_void_p), P(C.c_int), P(C.c_int64), C.c_int, C.c_int, C.c_int]
    L.sws_cuda_bind_thread_to_device.restype = C.c_int
    L.sws_cuda_bind_thread_to_device.argtypes = [C.c_int]
    L.sws_cuda_device_numa_node.restype = C.c_int
    L.sws_cuda_device_numa_node.argtypes = [C.c_int]
    L.sws_cuda_sync.restype = C.c_int
    L.sws_cuda_sync.argtypes = [ctxp]
    L.sws_cuda_stream.restype = C.c_void_p
    L.sws_cuda_stream.argtypes = [ctxp]
    L.sws_cuda_launch_count.restype = C.c_long
    L.sws_cuda_launch_count.argtypes = [ctxp]
    L.sws_cuda_kernel_name.restype = C.c_char_p
    L.sws_cuda_kernel_name.argtypes = [ctxp]
    L.sws_cuda_last_error.restype = C.c_char_p
    L.sws_cuda_last_error.argtypes = [ctxp]
    L.sws_cuda_host_alloc.restype = C.c_void_p
    L.sws_cuda_host_alloc.argtypes = [C.c_size_t]
    L.sws_cuda_host_free.argtypes = [C.c_void_p]
    L.sws_b200_plan_only.restype = C.c_int
    L.sws_b200_plan_only.argtypes = [ctxp]
    L.sws_b200_plan_only_filtered.restype = C.c_int
    L.sws_b200_plan_only_filtered.argtypes = [ctxp, C.c_void_p, C.c_void_p]
    L.sws_b200_get_filter.restype = C.c_int
    L.sws_b200_get_filter.argtypes = [ctxp, C.c_int, P(P(C.c_int16)), P(P(C.c_int32)), P(C.c_int)]
    L.sws_b200_get_rgb2yuv.restype = C.c_int
    L.sws_b200_get_rgb2yuv.argtypes = [ctxp, P(C.c_int)]
    L.sws_b200_get_info.restype = C.c_int
    L.sws_b200_get_info.argtypes = [ctxp, P(C.c_int)]
    _lib = L
    return L


def pix_fmt(name):
    return PIX_FMT[name] if isinstance(name, str) else int(name)


def _arr4(ctype, vals):
    vals = list(vals) + [0] * (4 - len(vals))
    return (ctype * 4)(*vals)


def _ptr_of(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):       # torch tensor (device or host)
        return x.data_ptr()
    return int(x)


class SwsContext:
    """sws_alloc_context() + field writes + sws_init_context(), like reference callers do
    (e.g. libswscale/tests/swscale.c:252-257)."""

    def __init__(self, src_w, src_h, src_fmt, dst_w, dst_h, dst_fmt, flags, param=None,
                 src_range=0, dst_range=0, chr_pos=None, dither=None, plan_only=False,
                 scaler=None, scaler_sub=None, src_filter=None, dst_filter=None):
        L = lib()
        self._L = L
        self.p = L.sws_alloc_context()
        if not self.p:
            raise MemoryError("sws_alloc_context failed")
        s = self.p.contents
        s.flags = flags
        s.src_w, s.src_h, s.src_format = src_w, src_h, pix_fmt(src_fmt)
        s.dst_w, s.dst_h, s.dst_format = dst_w, dst_h, pix_fmt(dst_fmt)
        s.src_range, s.dst_range = src_range, dst_range
        if param is not None:
            s.scaler_params[0], s.scaler_params[1] = param
        if chr_pos is not None:
            s.src_h_chr_pos, s.src_v_chr_pos, s.dst_h_chr_pos, s.dst_v_chr_pos = chr_pos
        if dither is not None:
            s.dither = dither
        if scaler is not None:
            s.scaler = scaler
        if scaler_sub is not None:
            s.scaler_sub = scaler_sub
        fptr = [None, None]
        self._filter_keep = []
        for fi, f in enumerate((src_filter, dst_filter)):
            if f:           # SwsFilter from a dict of vectors {"lumH": [...], "lumV": ..., "chrH": ..., "chrV": ...}
                flt = SwsFilterStruct()
                for k in ("lumH", "lumV", "chrH", "chrV"):
                    v = f.get(k)
                    if v is not None:
                        arr = (C.c_double * len(v))(*v)
                        vec = SwsVectorStruct(C.cast(arr, C.POINTER(C.c_double)), len(v))
                        self._filter_keep += [arr, vec]
                        setattr(flt, k, C.pointer(vec))
                self._filter_keep.append(flt)
                fptr[fi] = C.byref(flt)
        if plan_only and (src_filter or dst_filter):
            ret = L.sws_b200_plan_only_filtered(self.p, fptr[0], fptr[1])
        else:
            ret = L.sws_b200_plan_only(self.p) if plan_only else L.sws_init_context(self.p, fptr[0], fptr[1])
        if ret < 0:
            msg = L.sws_cuda_last_error(self.p).decode()
            L.sws_freeContext(self.p)
            self.p = None
            raise RuntimeError("sws_init_context failed (%d): %s" % (ret, msg))

    # -- life cycle
    def close(self):
        if self.p:
            self._L.sws_freeContext(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def fields(self):
        return self.p.contents

    # -- reference API
    def set_colorspace(self, src_cs, src_range, dst_cs, dst_range, brightness=0,
                       contrast=1 << 16, saturation=1 << 16):
        L = self._L
        return L.sws_setColorspaceDetails(self.p, L.sws_getCoefficients(src_cs), src_range,
                                          L.sws_getCoefficients(dst_cs), dst_range,
                                          brightness, contrast, saturation)

    def scale(self, src_planes, src_strides, dst_planes, dst_strides, y=0, h=None):
        """sws_scale() with HOST buffers (numpy arrays or raw addresses)."""
        h = self.fields.src_h if h is None else h
        sp = _arr4(C.c_void_p, [_ptr_of(p) for p in src_planes])
        ss = _arr4(C.c_int, src_strides)
        dp = _arr4(C.c_void_p, [_ptr_of(p) for p in dst_planes])
        ds = _arr4(C.c_int, dst_strides)
        return self._L.sws_scale(self.p, sp, ss, y, h, dp, ds)

    # -- CUDA extension
    def scale_batch_device(self, src_ptrs, src_strides, src_fstrides, dst_ptrs, dst_strides,
                           dst_fstrides, nb_frames):
        sp = _arr4(C.c_void_p, [_ptr_of(p) for p in src_ptrs])
        ss = _arr4(C.c_int, src_strides)
        sf = _arr4(C.c_int64, src_fstrides)
        dp = _arr4(C.c_void_p, [_ptr_of(p) for p in dst_ptrs])
        ds = _arr4(C.c_int, dst_strides)
        df = _arr4(C.c_int64, dst_fstrides)
        return self._L.sws_cuda_scale_batch(self.p, sp, ss, sf, dp, ds, df, nb_frames)

    def scale_batch_host(self, src_ptrs, src_strides, src_fstrides, dst_ptrs, dst_strides,
                         dst_fstrides, nb_frames, nb_devices=0, depth=0):
        """sws_cuda_scale_batch_host(): HOST frames, several in flight, spread over the visible devices."""
        sp = _arr4(C.c_void_p, [_ptr_of(p) for p in src_ptrs])
        ss = _arr4(C.c_int, src_strides)
        sf = _arr4(C.c_int64, src_fstrides)
        dp = _arr4(C.c_void_p, [_ptr_of(p) for p in dst_ptrs])
        ds = _arr4(C.c_int, dst_strides)
        df = _arr4(C.c_int64, dst_fstrides)
        return self._L.sws_cuda_scale_batch_host(self.p, sp, ss, sf, dp, ds, df, nb_frames, nb_devices, depth)

    def sync(self):
        return self._L.sws_cuda_sync(self.p)

    @property
    def stream(self):
        return self._L.sws_cuda_stream(self.p)

    @property
    def launch_count(self):
        return self._L.sws_cuda_launch_count(self.p)

    @property
    def kernel_name(self):
        return self._L.sws_cuda_kernel_name(self.p).decode()

    @property
    def last_error(self):
        return self._L.sws_cuda_last_error(self.p).decode()

    # -- diagnostics
    def filter(self, which):
        coef = C.POINTER(C.c_int16)()
        pos = C.POINTER(C.c_int32)()
        n = C.c_int()
        fs = self._L.sws_b200_get_filter(self.p, which, C.byref(coef), C.byref(pos), C.byref(n))
        if fs <= 0:
            return None
        co = np.ctypeslib.as_array(coef, shape=(n.value * fs,)).copy().reshape(n.value, fs)
        po = np.ctypeslib.as_array(pos, shape=(n.value,)).copy()
        return co, po

    def rgb2yuv(self):
        out = (C.c_int * 9)()
        if self._L.sws_b200_get_rgb2yuv(self.p, out) < 0:
            return None
        return list(out)

    def info(self):
        out = (C.c_int * 32)()
        if self._L.sws_b200_get_info(self.p, out) < 0:
            return None
        keys = {0: "y_offset", 1: "y_coeff", 2: "v2r", 3: "v2g", 4: "u2g", 5: "u2b", 6: "unscaled",
                8: "chrSrcW", 9: "chrSrcH", 10: "chrDstW", 11: "chrDstH", 12: "srcBpc", 13: "dstBpc",
                14: "flags", 15: "cy", 16: "yb", 17: "crv", 18: "cbu", 19: "cgu", 20: "cgv",
                21: "base_r", 22: "base_g", 23: "base_b", 24: "range_mode", 25: "lum_rc_coeff",
                26: "chr_rc_coeff", 27: "lum_identity", 28: "chr_h_identity", 29: "h_shift",
                30: "inter_bits", 31: "dst_kind"}
        return {v: out[k] for k, v in keys.items()}


class PinnedBuffer:
    """Page-locked host memory from sws_cuda_host_alloc(), viewed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self._L = lib()
        self.ptr = self._L.sws_cuda_host_alloc(nbytes)
        if not self.ptr:
            raise MemoryError("sws_cuda_host_alloc(%d) failed" % nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            self._L.sws_cuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return lib().sws_cuda_device_count()


def bind_thread_to_device(device):
    """Run the calling thread next to `device` (its NUMA node); -1 if the topology is unknown."""
    return lib().sws_cuda_bind_thread_to_device(device)


def device_numa_node(device):
    return lib().sws_cuda_device_numa_node(device)
