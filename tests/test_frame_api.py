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


class AVBufferRef(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("data", C.c_void_p), ("size", C.c_size_t)]


class AVHWDeviceContext(C.Structure):
    _fields_ = [("av_class", C.c_void_p), ("type", C.c_int), ("hwctx", C.c_void_p)]


class AVHWFramesContext(C.Structure):
    _fields_ = [("av_class", C.c_void_p), ("device_ref", C.POINTER(AVBufferRef)), ("device_ctx", C.c_void_p),
                ("hwctx", C.c_void_p), ("free", C.c_void_p), ("user_opaque", C.c_void_p), ("pool", C.c_void_p),
                ("initial_pool_size", C.c_int), ("format", C.c_int), ("sw_format", C.c_int),
                ("width", C.c_int), ("height", C.c_int)]


AV_PIX_FMT_CUDA = 117


class CudaFramesCtx:
    """What libavutil's hwcontext_cuda would hand out: device ctx + frames ctx behind AVBufferRefs."""

    def __init__(self, sw_format, w, h, dev_type=2, device=None):
        self.dev = device or AVHWDeviceContext(None, dev_type, None)
        self.dev_ref = AVBufferRef(None, C.addressof(self.dev), C.sizeof(self.dev))
        self.fc = AVHWFramesContext()
        self.fc.device_ref = C.pointer(self.dev_ref)
        self.fc.format, self.fc.sw_format = AV_PIX_FMT_CUDA, S.pix_fmt(sw_format)
        self.fc.width, self.fc.height = w, h
        self.ref = AVBufferRef(None, C.addressof(self.fc), C.sizeof(self.fc))


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


@pytest.mark.skipif(not R.available() or not hasattr(R.lib(), "swsref_hw_offsets"),
                    reason="oracle/_ref/libswsref.so not built")
def test_hw_context_mirrors_match_reference_layout():
    out = (C.c_int * 16)()
    R.lib().swsref_hw_offsets(out)
    got = [AVBufferRef.data.offset, AVBufferRef.size.offset, AVHWDeviceContext.type.offset,
           AVHWDeviceContext.hwctx.offset, AVHWFramesContext.device_ref.offset, AVHWFramesContext.device_ctx.offset,
           AVHWFramesContext.hwctx.offset, AVHWFramesContext.pool.offset, AVHWFramesContext.initial_pool_size.offset,
           AVHWFramesContext.format.offset, AVHWFramesContext.sw_format.offset, AVHWFramesContext.width.offset,
           AVHWFramesContext.height.offset, 2, AV_PIX_FMT_CUDA]
    assert got == list(out)[:15]


def test_hw_frame_rules_without_a_device():
    """sws_frame_setup's hardware-frame checks (swscale.c:1511-1538) are host logic: mixed hw/sw frames,
    unallocated frames, different devices and non-CUDA devices are refused before any device work."""
    L = _bind()
    assert L.sws_test_hw_format(AV_PIX_FMT_CUDA) == 1 and L.sws_test_hw_format(-1) == 1
    assert L.sws_test_hw_format(S.pix_fmt("yuv420p")) == 0
    a, b = T.Frame("yuv420p", 64, 48), T.Frame("rgb24", 64, 48)
    fa, fb = make_avframe(a, "yuv420p"), make_avframe(b, "rgb24")
    ca, cb = CudaFramesCtx("yuv420p", 64, 48), CudaFramesCtx("rgb24", 64, 48)
    L.sws_test_frame.argtypes = [C.POINTER(AVFrame), C.c_int]
    assert L.sws_test_frame(C.byref(fa), 0) == 0                         # primaries/transfer 0 are reserved values
    fa.color_primaries = fa.color_trc = 2                                # what av_frame_alloc() sets: unspecified
    assert L.sws_test_frame(C.byref(fa), 0) == 1
    ctx = L.sws_alloc_context()
    try:
        fa.format, fa.hw_frames_ctx = AV_PIX_FMT_CUDA, C.addressof(ca.ref)
        assert L.sws_test_frame(C.byref(fa), 0) == 1                     # judged by its sw_format
        assert L.sws_frame_setup(ctx, C.byref(fb), C.byref(fa)) == -95   # ENOTSUP: only one side is a hw frame
        fb.format, fb.hw_frames_ctx = AV_PIX_FMT_CUDA, C.addressof(cb.ref)
        assert L.sws_frame_setup(ctx, C.byref(fb), C.byref(fa)) == -22   # EINVAL: different devices
        cb2 = CudaFramesCtx("rgb24", 64, 48, device=ca.dev)
        cb2.fc.device_ref = ca.fc.device_ref
        fb.hw_frames_ctx = C.addressof(cb2.ref)
        d0 = fb.data[0]
        fb.data[0] = None
        assert L.sws_frame_setup(ctx, C.byref(fb), C.byref(fa)) == -22   # EINVAL: not allocated
        fb.data[0] = d0
        ca.dev.type = 11                                                  # AV_HWDEVICE_TYPE_VULKAN
        assert L.sws_frame_setup(ctx, C.byref(fb), C.byref(fa)) == -95   # ENOTSUP: not a CUDA device
        assert L.sws_is_noop(C.byref(fa), C.byref(fa)) == 1 and L.sws_is_noop(C.byref(fb), C.byref(fa)) == 0
    finally:
        S.lib().sws_freeContext(ctx)


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


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [("yuv420p", 1280, 720, "rgb24", 1280, 720, S.SWS_BICUBIC | S.BX),
                                  ("nv12", 1280, 720, "bgra", 1280, 720, S.SWS_BICUBIC | S.BX),
                                  ("yuv420p", 640, 360, "rgb24", 1280, 720, S.SWS_BICUBIC | S.BX),
                                  ("rgb24", 1280, 720, "yuv420p", 1280, 720, S.SWS_BICUBIC | S.BX),
                                  ("nv12", 1920, 1080, "yuv420p", 640, 360, S.SWS_BICUBIC | S.BX)])
def test_scale_frame_cuda_hw_frames(geom):
    """AV_PIX_FMT_CUDA frames (device pointers in data[], sw_format in the frames
    context) convert in place in device memory and equal the host-frame result of the same call."""
    torch = pytest.importorskip("torch")
    sf, sw, sh, df, dw, dh, flags = geom
    L = _bind()
    src = T.Frame(sf, sw, sh).randomize(71)
    want = T.Frame(df, dw, dh, fill=0)
    props, dprops = (1, 1, 1), (0, 1, 0)        # same matrix on both sides (YUV->YUV matrix changes cascade)
    fs, fd = make_avframe(src, sf, props), make_avframe(want, df, dprops)
    ctx = L.sws_alloc_context()
    ctx.contents.flags = flags
    try:
        assert L.sws_scale_frame(ctx, C.byref(fd), C.byref(fs)) >= 0

        dev = torch.device("cuda", 0)
        d_src = [torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in src.planes]
        d_dst = [torch.zeros(p.shape, dtype=torch.uint8, device=dev) for p in want.planes]
        cs, cd = CudaFramesCtx(sf, sw, sh), CudaFramesCtx(df, dw, dh)
        cd.fc.device_ref = cs.fc.device_ref
        hs, hd = make_avframe(src, sf, props), make_avframe(want, df, dprops)
        for f, planes, fc in ((hs, d_src, cs), (hd, d_dst, cd)):
            for i in range(8):
                f.data[i] = planes[i].data_ptr() if i < len(planes) else None
            f.format, f.hw_frames_ctx = AV_PIX_FMT_CUDA, C.addressof(fc.ref)
        torch.cuda.synchronize()
        ctx2 = L.sws_alloc_context()
        ctx2.contents.flags = flags
        try:
            assert L.sws_scale_frame(ctx2, C.byref(hd), C.byref(hs)) >= 0
        finally:
            S.lib().sws_freeContext(ctx2)
        for got, w, (rows, rb) in zip(d_dst, want.planes, want.layout):
            assert np.array_equal(got.cpu().numpy()[:rows, :rb], w[:rows, :rb])
    finally:
        S.lib().sws_freeContext(ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dynamic", "legacy"])
@pytest.mark.parametrize("case", [(640, 360, "yuv420p", 640, 360, "rgb24"), (640, 360, "yuv420p", 320, 180, "yuv420p"),
                                  (322, 242, "nv12", 400, 300, "bgra")])
def test_scale_frame_allocates_destination(mode, case, tmp_path):
    """swscale.c:1316-1330,1437-1467: a destination frame without buffers is allocated by sws_scale_frame().  The
    library finds libavutil's allocator in the calling process; here that process is a small C program that
    links the reference's own libavutil objects."""
    import glob
    import subprocess
    from librempeg_b200 import build as native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    objs = sorted(glob.glob(os.path.join(root, "oracle", "_ref", "obj", "avu_*.o")))
    if not objs:
        pytest.skip("oracle/_ref/obj (reference libavutil objects) not built")
    exe = os.path.join(root, "tests", "native", "_frame_alloc_probe")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(exe[:-len("_frame_alloc_probe")] + "frame_alloc_probe.c"):
        cmd = ["gcc", "-O1", "-rdynamic", "-o", exe, os.path.join(root, "tests", "native", "frame_alloc_probe.c"),
               "-I", os.path.join(root, "include")] + objs + \
              ["-L", os.path.dirname(native.SO_PATH), "-lswscale_b200", "-Wl,-rpath," + os.path.dirname(native.SO_PATH),
               "-lm", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    sw, sh, sf, dw, dh, df = case
    flags = S.SWS_BICUBIC | S.BX
    src = T.Frame(sf, sw, sh).randomize(77)
    # the dynamic mode reads the frames' (unspecified) colour properties: same defaults as a legacy context
    want, _ = T.run_reference(sw, sh, sf, dw, dh, df, flags, src)
    sp, dp = tmp_path / "src.bin", tmp_path / "dst.bin"
    sp.write_bytes(b"".join(p.tobytes() for p in src.valid()))
    r = subprocess.run([exe, mode, str(sw), str(sh), str(S.pix_fmt(sf)), str(dw), str(dh), str(S.pix_fmt(df)),
                        str(flags), str(sp), str(dp)], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stderr[-500:])
    got = np.frombuffer(dp.read_bytes(), np.uint8)
    exp = np.concatenate([p.reshape(-1) for p in want.valid()])
    assert got.size == exp.size and np.array_equal(got, exp)


class AVCUDADeviceContext(C.Structure):
    _fields_ = [("cuda_ctx", C.c_void_p), ("stream", C.c_void_p), ("internal", C.c_void_p)]


@pytest.mark.gpu
def test_cuda_frames_of_a_real_device_context():
    """frames of a libavutil CUDA device carry a CUcontext and a stream (hwcontext_cuda.h).  The primary
    context of the device is accepted and the producer's stream is honoured (a fill queued on it just before the
    call must be seen); a context made by cuCtxCreate() -- libavutil's default -- is refused, not guessed."""
    torch = pytest.importorskip("torch")
    cu = C.CDLL("libcuda.so.1")
    dev = torch.device("cuda", 0)
    torch.zeros(1, device=dev)                    # runtime initialised: the primary context exists
    primary = C.c_void_p()
    assert cu.cuDevicePrimaryCtxRetain(C.byref(primary), 0) == 0
    other = C.c_void_p()
    create = getattr(cu, "cuCtxCreate_v2")
    assert create(C.byref(other), 0, 0) == 0      # pushes it: pop again
    popped = C.c_void_p()
    assert cu.cuCtxPopCurrent_v2(C.byref(popped)) == 0
    L = _bind()
    sw, sh = 640, 360
    src = T.Frame("nv12", sw, sh).randomize(5)
    want = T.Frame("rgb24", sw, sh, fill=0)
    props, dprops = (1, 1, 1), (0, 1, 0)
    try:
        # reference result through the host path
        fs, fd = make_avframe(src, "nv12", props), make_avframe(want, "rgb24", dprops)
        ctx = L.sws_alloc_context()
        ctx.contents.flags = S.SWS_BICUBIC | S.BX
        assert L.sws_scale_frame(ctx, C.byref(fd), C.byref(fs)) >= 0
        S.lib().sws_freeContext(ctx)

        producer = torch.cuda.Stream(device=dev)
        for cuctx, expect_ok in ((primary, True), (other, False)):
            hwdev = AVCUDADeviceContext(cuctx.value, producer.cuda_stream, None)
            device = AVHWDeviceContext(None, 2, C.addressof(hwdev))
            cs, cd = CudaFramesCtx("nv12", sw, sh, device=device), CudaFramesCtx("rgb24", sw, sh, device=device)
            cd.fc.device_ref = cs.fc.device_ref
            d_src = [torch.zeros(p.shape, dtype=torch.uint8, device=dev) for p in src.planes]
            d_dst = [torch.zeros(p.shape, dtype=torch.uint8, device=dev) for p in want.planes]
            h_src = [torch.from_numpy(np.ascontiguousarray(p)).pin_memory() for p in src.planes]
            torch.cuda.synchronize()
            with torch.cuda.stream(producer):     # the "decoder": a long spin, then the upload, all async
                torch.cuda._sleep(200_000_000)
                for d, h in zip(d_src, h_src):
                    d.copy_(h, non_blocking=True)
            hs, hd = make_avframe(src, "nv12", props), make_avframe(want, "rgb24", dprops)
            for f, planes, fc in ((hs, d_src, cs), (hd, d_dst, cd)):
                for i in range(8):
                    f.data[i] = planes[i].data_ptr() if i < len(planes) else None
                f.format, f.hw_frames_ctx = AV_PIX_FMT_CUDA, C.addressof(fc.ref)
            ctx2 = L.sws_alloc_context()
            ctx2.contents.flags = S.SWS_BICUBIC | S.BX
            ret = L.sws_scale_frame(ctx2, C.byref(hd), C.byref(hs))
            S.lib().sws_freeContext(ctx2)
            torch.cuda.synchronize()
            if expect_ok:
                assert ret >= 0
                for got, w, (rows, rb) in zip(d_dst, want.planes, want.layout):
                    assert np.array_equal(got.cpu().numpy()[:rows, :rb], w[:rows, :rb])
            else:
                assert ret == -95                 # ENOTSUP
    finally:
        cu.cuCtxDestroy_v2(other)
        cu.cuDevicePrimaryCtxRelease_v2(0)
