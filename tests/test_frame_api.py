"""AVFrame entry points: sws_scale_frame / sws_frame_setup / sws_is_noop and the
slice-wise frame API, against the reference's own sws_scale_frame() in dynamic mode."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S
from oracle import refapi as R


class AVRational(C.Structure):
    _fields_ = [("num", C.c_int), ("den", C.c_int)]


class AVFrame(C.Structure):
    """ctypes twin of the AVFrame prefix declared in include/swscale_b200_frame.h."""
    _fields_ = [
        ("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
        ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int),
        ("pict_type", C.c_int), ("sample_aspect_ratio", AVRational), ("pts", C.c_int64),
        ("pkt_dts", C.c_int64), ("time_base", AVRational), ("quality", C.c_int), ("opaque", C.c_void_p),
        ("repeat_pict", C.c_int), ("sample_rate", C.c_int), ("buf", C.c_void_p * 8),
        ("extended_buf", C.c_void_p), ("nb_extended_buf", C.c_int), ("side_data", C.c_void_p),
        ("nb_side_data", C.c_int), ("flags", C.c_int), ("color_range", C.c_int),
        ("color_primaries", C.c_int), ("color_trc", C.c_int), ("colorspace", C.c_int),
        ("chroma_location", C.c_int), ("best_effort_timestamp", C.c_int64), ("metadata", C.c_void_p),
        ("decode_error_flags", C.c_int), ("hw_frames_ctx", C.c_void_p),
    ]


def _bind():
    L = S.lib()
    P = C.POINTER
    ctxp = P(S.SwsContextStruct)
    L.sws_scale_frame.restype = C.c_int
    L.sws_scale_frame.argtypes = [ctxp, P(AVFrame), P(AVFrame)]
    L.sws_frame_setup.restype = C.c_int
    L.sws_frame_setup.argtypes = [ctxp, P(AVFrame), P(AVFrame)]
    L.sws_is_noop.restype = C.c_int
    L.sws_is_noop.argtypes = [P(AVFrame), P(AVFrame)]
    L.sws_frame_start.restype = C.c_int
    L.sws_frame_start.argtypes = [ctxp, P(AVFrame), P(AVFrame)]
    L.sws_frame_end.argtypes = [ctxp]
    L.sws_send_slice.restype = C.c_int
    L.sws_send_slice.argtypes = [ctxp, C.c_uint, C.c_uint]
    L.sws_receive_slice.restype = C.c_int
    L.sws_receive_slice.argtypes = [ctxp, C.c_uint, C.c_uint]
    L.sws_receive_slice_alignment.restype = C.c_uint
    L.sws_receive_slice_alignment.argtypes = [ctxp]
    return L


def make_avframe(frame, fmt, props=(0, 2, 0)):
    f = AVFrame()
    for i, (pl, st) in enumerate(zip(frame.planes, frame.strides)):
        f.data[i] = pl.ctypes.data
        f.linesize[i] = st
    f.width, f.height, f.format = frame.w, frame.h, S.pix_fmt(fmt)
    f.color_range, f.colorspace, f.chroma_location = props
    return f


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
def test_avframe_mirror_matches_reference_layout():
    """Every field offset of the declared AVFrame prefix equals the reference's struct (libavutil/frame.h)."""
    out = (C.c_int * 24)()
    R.lib().swsref_frame_offsets(out)
    names = ["data", "linesize", "extended_data", "width", "height", "nb_samples", "format", "pict_type",
             "sample_aspect_ratio", "pts", "pkt_dts", "time_base", "quality", "opaque", "repeat_pict",
             "sample_rate", "buf", "flags", "color_range", "color_primaries", "color_trc", "colorspace",
             "chroma_location", "hw_frames_ctx"]
    for n, off in zip(names, out):
        assert getattr(AVFrame, n).offset == off, n


def test_sws_is_noop():
    L = _bind()
    a = T.Frame("yuv420p", 64, 48)
    b = T.Frame("yuv420p", 64, 48)
    fa, fb = make_avframe(a, "yuv420p"), make_avframe(b, "yuv420p")
    assert L.sws_is_noop(C.byref(fa), C.byref(fb)) == 1
    fb.color_range = 2
    assert L.sws_is_noop(C.byref(fa), C.byref(fb)) == 0
    fb.color_range = 0
    fb.chroma_location = 1
    assert L.sws_is_noop(C.byref(fa), C.byref(fb)) == 0
    c = T.Frame("rgb24", 64, 48)
    fc = make_avframe(c, "rgb24")
    assert L.sws_is_noop(C.byref(fc), C.byref(fa)) == 0
    # RGB frames: range/colorspace/location are irrelevant and sanitised away (format.c:305-339)
    fd = make_avframe(c, "rgb24", props=(1, 5, 3))
    assert L.sws_is_noop(C.byref(fc), C.byref(fd)) == 1


# (src_range, src_csp, src_loc, dst_range, dst_csp, dst_loc)
PROPS = [(0, 2, 0, 0, 2, 0), (1, 1, 1, 0, 2, 0), (2, 5, 3, 0, 0, 0), (1, 9, 2, 0, 2, 0), (2, 7, 5, 0, 2, 0)]


@pytest.mark.gpu
@pytest.mark.parametrize("props", PROPS)
@pytest.mark.parametrize("geom", [("yuv420p", 320, 240, "rgb24", 320, 240, S.SWS_BICUBIC | S.BX),
                                  ("yuv420p", 320, 240, "rgb24", 320, 240, S.SWS_BICUBIC),
                                  ("yuv420p", 320, 240, "bgra", 480, 360, S.SWS_BILINEAR | S.BX),
                                  ("yuv420p10le", 320, 240, "rgb48le", 320, 240, S.SWS_LANCZOS | S.BX),
                                  ("yuvj420p", 320, 240, "rgb24", 200, 150, S.SWS_BICUBIC | S.BX)])
def test_scale_frame_dynamic_matches_reference(props, geom):
    """Dynamic mode: everything comes from the frames (vf_scale's call, libavfilter/vf_scale.c:852)."""
    sf, sw, sh, df, dw, dh, flags = geom
    L = _bind()
    src = T.Frame(sf, sw, sh).randomize(61)
    # reference: refcounted frames owned by libavutil
    rs, rd = R.RefFrame(sw, sh, sf), R.RefFrame(dw, dh, df)
    for i, (rows, rb) in enumerate(src.layout):
        a, ls = rs.plane(i, rows)
        a[:, :rb] = src.planes[i][:, :rb]
    pr = (C.c_int * 6)(*props)
    ret = R.lib().swsref_scale_frame_dynamic(flags, 1, C.c_void_p(rd.f), C.c_void_p(rs.f), pr)
    assert ret >= 0
    want = [rd.plane(i, rows)[0][:, :rb].copy() for i, (rows, rb) in enumerate(T.plane_layout(df, dw, dh))]
    # ours
    dst = T.Frame(df, dw, dh, fill=0)
    fs, fd = make_avframe(src, sf, props[:3]), make_avframe(dst, df, props[3:])
    ctx = L.sws_alloc_context()
    ctx.contents.flags = flags
    for _ in range(2):                       # second call reuses the planned inner context
        assert L.sws_scale_frame(ctx, C.byref(fd), C.byref(fs)) >= 0, L.sws_cuda_last_error(ctx)
    L.sws_freeContext(ctx)
    assert T.first_diff(dst.valid(), want) is None


@pytest.mark.gpu
def test_scale_frame_replans_when_frames_change():
    L = _bind()
    ctx = L.sws_alloc_context()
    ctx.contents.flags = S.SWS_BICUBIC | S.BX
    for (sf, sw, sh, df, dw, dh) in [("yuv420p", 320, 240, "rgb24", 320, 240), ("yuv420p", 160, 120, "rgb24", 320, 240),
                                     ("nv12", 320, 240, "yuv420p", 160, 120), ("yuv420p", 320, 240, "rgb24", 320, 240)]:
        src = T.Frame(sf, sw, sh).randomize(62)
        dst = T.Frame(df, dw, dh, fill=0)
        fs, fd = make_avframe(src, sf), make_avframe(dst, df)
        assert L.sws_scale_frame(ctx, C.byref(fd), C.byref(fs)) >= 0
        want, _ = T.run_reference(sw, sh, sf, dw, dh, df, S.SWS_BICUBIC | S.BX, src)
        assert T.first_diff(dst.valid(), want.valid()) is None
    L.sws_freeContext(ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [("yuv420p", 640, 480, "rgb24", 640, 480, S.SWS_BICUBIC | S.BX),
                                  ("yuv420p", 640, 480, "yuv420p", 320, 200, S.SWS_BICUBIC | S.BX),
                                  ("yuv420p", 320, 240, "rgb24", 640, 480, S.SWS_BICUBIC)])
def test_legacy_frame_slice_api(geom):
    """sws_frame_start / sws_send_slice / sws_receive_slice on a legacy context: output requested in
    bands (the reference's threaded path does exactly this per slice thread, swscale.c:1361-1403)."""
    sf, sw, sh, df, dw, dh, flags = geom
    L = _bind()
    src = T.Frame(sf, sw, sh).randomize(63)
    want, _ = T.run_reference(sw, sh, sf, dw, dh, df, flags, src)
    c = S.SwsContext(sw, sh, sf, dw, dh, df, flags)
    dst = T.Frame(df, dw, dh, fill=0)
    fs, fd = make_avframe(src, sf), make_avframe(dst, df)
    assert L.sws_frame_start(c.p, C.byref(fd), C.byref(fs)) == 0
    assert L.sws_receive_slice(c.p, 0, dh) == -11          # AVERROR(EAGAIN): no input signalled yet
    assert L.sws_send_slice(c.p, 0, sh // 2) == 0
    assert L.sws_send_slice(c.p, sh // 2, sh - sh // 2) == 0
    al = L.sws_receive_slice_alignment(c.p)
    band = 56 // al * al
    y = 0
    while y < dh:
        h = min(band, dh - y)
        assert L.sws_receive_slice(c.p, y, h) == h
        y += h
    L.sws_frame_end(c.p)
    assert T.first_diff(dst.valid(), want.valid()) is None
    # whole-frame convenience call on the same legacy context
    dst2 = T.Frame(df, dw, dh, fill=0)
    fd2 = make_avframe(dst2, df)
    assert L.sws_scale_frame(c.p, C.byref(fd2), C.byref(fs)) >= 0
    assert T.first_diff(dst2.valid(), want.valid()) is None
