"""Shared helpers of the test-suite: synthetic frames, running the oracle and the
CUDA library on identical inputs, comparing planes.

The oracle here is oracle/_ref/libswsref.so -- the REAL reference libswscale C
path compiled from /root/reference by oracle/build_ref.py (it travels to the GPU
box as a prebuilt .so) -- plus oracle/sws_oracle (our own C restatement).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refapi as R          # noqa: E402
from librempeg_b200 import swscale as S  # noqa: E402

# format -> (bytes per sample, [(w_shift, h_shift, samples_per_pixel)] per plane)
_YUV = {"420": (1, 1), "422": (1, 0), "444": (0, 0)}


def plane_layout(fmt, w, h):
    """[(rows, row_bytes)] for every plane of a tightly described frame."""
    def cdiv(a, s):
        return -((-a) >> s)
    if fmt in ("rgb24", "bgr24"):
        return [(h, w * 3)]
    if fmt in ("rgba", "bgra", "argb", "abgr"):
        return [(h, w * 4)]
    if fmt in ("rgb48le", "bgr48le"):
        return [(h, w * 6)]
    if fmt in ("rgb565le", "bgr565le", "rgb555le", "bgr555le"):
        return [(h, w * 2)]
    if fmt in ("nv12", "nv21"):
        return [(h, w), (cdiv(h, 1), cdiv(w, 1) * 2)]
    if fmt == "p010le":
        return [(h, w * 2), (cdiv(h, 1), cdiv(w, 1) * 4)]
    if fmt == "gray":
        return [(h, w)]
    if fmt == "gbrp":
        return [(h, w)] * 3
    if fmt == "grayf32le":
        return [(h, w * 4)]
    if fmt == "gbrpf32le":
        return [(h, w * 4)] * 3
    for key, (cw, ch) in _YUV.items():
        for pre in ("yuv", "yuvj"):
            if fmt.startswith(pre + key + "p"):
                suffix = fmt[len(pre + key + "p"):]
                bps = 1 if suffix == "" else 2
                return [(h, w * bps), (cdiv(h, ch), cdiv(w, cw) * bps), (cdiv(h, ch), cdiv(w, cw) * bps)]
    raise ValueError(fmt)


def depth_of(fmt):
    if fmt == "p010le":
        return 16            # fill the whole 16-bit container: the readers drop the low six bits
    for d in (9, 10, 12, 14, 16):
        if fmt.endswith("p%dle" % d):
            return d
    return 8


def det_values(seed, n, bits):
    """Deterministic pseudo-random samples (splitmix64 of the index): identical on every
    machine and numpy version, so golden md5s in tests/golden/ stay valid forever."""
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15 + 0x1234567) & 0xFFFFFFFFFFFFFFFF)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z & np.uint64((1 << bits) - 1))


class Frame:
    """Planes as 2-D uint8 arrays [rows, stride] with row_bytes valid bytes per row."""

    def __init__(self, fmt, w, h, pad=0, fill=None):
        self.fmt, self.w, self.h = fmt, w, h
        self.layout = plane_layout(fmt, w, h)
        self.planes, self.strides = [], []
        for rows, rb in self.layout:
            stride = ((rb + 63) & ~63) + pad
            a = np.zeros((rows, stride), np.uint8)
            if fill is not None:
                a[:] = fill
            self.planes.append(a)
            self.strides.append(stride)

    def randomize(self, seed, mode="noise"):
        rng = np.random.default_rng(seed)
        d = depth_of(self.fmt)
        for pi, ((rows, rb), a) in enumerate(zip(self.layout, self.planes)):
            if mode == "noise":
                if d == 8:
                    a[:, :rb] = det_values(seed * 4 + pi, rows * rb, 8).astype(np.uint8).reshape(rows, rb)
                else:
                    v = det_values(seed * 4 + pi, rows * (rb // 2), d).astype(np.uint16).reshape(rows, rb // 2)
                    a[:, :rb] = v.view(np.uint8).reshape(rows, rb)
            elif mode == "smooth":
                n = rb if d == 8 else rb // 2
                yy, xx = np.mgrid[0:rows, 0:n]
                v = (np.sin(xx / 17.0 + seed) * np.cos(yy / 23.0) * 0.5 + 0.5) * ((1 << d) - 1)
                v = v + rng.integers(0, 3, (rows, n))
                v = np.clip(v, 0, (1 << d) - 1)
                if d == 8:
                    a[:, :rb] = v.astype(np.uint8)
                else:
                    a[:, :rb] = v.astype(np.uint16).view(np.uint8).reshape(rows, rb)
            elif mode == "extreme":   # saturated checker: exercises every clip
                n = rb if d == 8 else rb // 2
                yy, xx = np.mgrid[0:rows, 0:n]
                v = (((xx // 3 + yy // 2) & 1) * ((1 << d) - 1)).astype(np.uint16)
                if d == 8:
                    a[:, :rb] = v.astype(np.uint8)
                else:
                    a[:, :rb] = v.view(np.uint8).reshape(rows, rb)
        return self

    def valid(self):
        """Tight copies of the valid bytes of each plane."""
        return [a[:, :rb].copy() for (rows, rb), a in zip(self.layout, self.planes)]


def run_reference(sw, sh, sf, dw, dh, df, flags, src, dst_pad=0, param=None, ctx_kwargs=None,
                  colorspace=None, slices=None):
    kw = dict(ctx_kwargs or {})
    c = R.RefContext(sw, sh, sf, dw, dh, df, flags, param=param, **kw)
    if colorspace:
        c.set_colorspace(*colorspace)
    dst = Frame(df, dw, dh, pad=dst_pad, fill=0)
    _drive(c, src, dst, sh, slices)
    info = c.info()
    c.close()
    return dst, info


def run_cuda(sw, sh, sf, dw, dh, df, flags, src, dst_pad=0, param=None, ctx_kwargs=None,
             colorspace=None, slices=None):
    kw = dict(ctx_kwargs or {})
    c = S.SwsContext(sw, sh, sf, dw, dh, df, flags, param=param, **kw)
    if colorspace and c.set_colorspace(*colorspace) < 0:
        err = c.last_error
        c.close()
        raise NotImplementedError("sws_setColorspaceDetails refused: %s" % err)
    dst = Frame(df, dw, dh, pad=dst_pad, fill=0)
    _drive(c, src, dst, sh, slices)
    name = c.kernel_name
    c.close()
    return dst, name


def _drive(c, src, dst, sh, slices):
    """Call sws_scale() once for the whole frame or once per slice [(y, h), ...]."""
    if not slices:
        ret = c.scale(src.planes, src.strides, dst.planes, dst.strides, 0, sh)
        assert ret == dst.h, "sws_scale returned %d, expected %d" % (ret, dst.h)
        return
    total = 0
    vs = {"420": 1}.get(_subs(src.fmt), 0)
    for (y, h) in slices:
        planes = []
        for i, a in enumerate(src.planes):
            yy = y >> (vs if i else 0) if not src.fmt.startswith(("rgb", "bgr")) else y
            planes.append(a[yy:])
        ret = c.scale(planes, src.strides, dst.planes, dst.strides, y, h)
        assert ret >= 0, "sws_scale failed: %d" % ret
        total += ret
    assert total == dst.h, "slices produced %d rows, expected %d" % (total, dst.h)


def _subs(fmt):
    if fmt in ("nv12", "nv21", "p010le"):
        return "420"
    for k in _YUV:
        if k + "p" in fmt:
            return k
    return "444"


def run_oracle(sw, sh, sf, dw, dh, df, flags, src, param=None, ctx_kwargs=None, colorspace=None, **_):
    """The numpy restatement (oracle/sws_oracle.py) on the same frame; returns tight planes."""
    from oracle import sws_oracle as O
    kw = dict(ctx_kwargs or {})
    c = O.OracleContext(sw, sh, sf, dw, dh, df, flags, param=param, colorspace=colorspace, **kw)
    out = c.scale(src.valid())
    lay = plane_layout(df, dw, dh)
    return [np.ascontiguousarray(o).view(np.uint8).reshape(rows, -1)[:, :rb] for o, (rows, rb) in zip(out, lay)]


def md5_planes(planes):
    import hashlib
    return [hashlib.md5(np.ascontiguousarray(p).tobytes()).hexdigest() for p in planes]


def run_case_both(sw, sh, sf, dw, dh, df, flags, seed=1, mode="noise", src_pad=0, dst_pad=0, **kw):
    src = Frame(sf, sw, sh, pad=src_pad).randomize(seed, mode)
    want, _ = run_reference(sw, sh, sf, dw, dh, df, flags, src, dst_pad=dst_pad, **kw)
    got, name = run_cuda(sw, sh, sf, dw, dh, df, flags, src, dst_pad=dst_pad, **kw)
    return got.valid(), want.valid(), name


def first_diff(got, want):
    for p, (g, w) in enumerate(zip(got, want)):
        if g.shape != w.shape:
            return "plane %d shape %s vs %s" % (p, g.shape, w.shape)
        bad = np.argwhere(g != w)
        if len(bad):
            y, x = bad[0]
            return "plane %d: %d bytes differ, first at row %d byte %d: got %d want %d" % (
                p, len(bad), y, x, g[y, x], w[y, x])
    return None
