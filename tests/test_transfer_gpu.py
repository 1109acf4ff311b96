"""How host frames reach the kernels (csrc/sws_xfer.cuh): pageable and page-locked memory, the banded
pipeline for every kernel of the dispatch chain, negative strides, bottom-up slices, and the multi-lane
host batch.  Every case is held bit-for-bit against the reference C path."""
import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu

BX = S.BX

# one case per kernel family of the dispatch chain, large enough (>= 8 MB) to take the banded pipeline
BANDED_CASES = [
    dict(sw=1920, sh=1080, sf="yuv420p", dw=1920, dh=1080, df="rgb24", flags=S.SWS_BICUBIC | BX),        # fast420
    dict(sw=1920, sh=1080, sf="yuv420p10le", dw=1920, dh=1080, df="rgb48le", flags=S.SWS_LANCZOS | BX),  # fast16
    dict(sw=1920, sh=1080, sf="yuv420p10le", dw=1920, dh=1080, df="bgra", flags=S.SWS_BICUBIC | BX),     # hi8
    dict(sw=3840, sh=2160, sf="nv12", dw=1280, dh=720, df="yuv420p", flags=S.SWS_BICUBIC | BX),          # scale8
    dict(sw=1280, sh=720, sf="yuv420p", dw=1920, dh=1080, df="rgb24", flags=S.SWS_BICUBIC | BX),         # scale8 rgb
    dict(sw=1920, sh=1080, sf="rgb24", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC | BX),        # rgb420
    dict(sw=1920, sh=1080, sf="bgra", dw=1280, dh=720, df="nv12", flags=S.SWS_BICUBIC | BX),             # tile15
    dict(sw=1920, sh=1080, sf="yuv420p10le", dw=1280, dh=720, df="yuv420p16le", flags=S.SWS_BILINEAR | BX),  # generic
    dict(sw=1920, sh=1080, sf="yuv444p", dw=1920, dh=1080, df="rgb24", flags=S.SWS_BICUBIC | BX),        # full444
    dict(sw=1920, sh=1080, sf="nv12", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC | BX),         # copy8
    dict(sw=1920, sh=1080, sf="yuv420p10le", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC | BX),  # depthcopy (dither rows)
    dict(sw=1920, sh=1080, sf="rgba", dw=1920, dh=1080, df="bgra", flags=S.SWS_BICUBIC | BX),            # shuffle
    dict(sw=1920, sh=1080, sf="bgr24", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC),             # bgr24->yv12
    dict(sw=1918, sh=1078, sf="yuv420p", dw=1000, dh=562, df="rgb565le", flags=S.SWS_BICUBIC | BX),      # odd sizes
]


def _ids(c):
    return "%s_%dx%d_%s_%dx%d" % (c["sf"], c["sw"], c["sh"], c["df"], c["dw"], c["dh"])


@pytest.mark.parametrize("case", BANDED_CASES, ids=_ids)
@pytest.mark.parametrize("memory", ["pageable", "pinned_src", "pinned_both"])
@pytest.mark.parametrize("bands", ["4", "7"])
def test_whole_frame_paths(case, memory, bands, monkeypatch):
    monkeypatch.setenv("SWS_B200_E2E_BANDS", bands)
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(41)
    want, _ = T.run_reference(src=src, **case)
    ctx = S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
    dst = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
    keep = []

    def pinned_copy(frame):
        planes = []
        for a in frame.planes:
            b = S.PinnedBuffer(a.nbytes)
            b.array[:] = a.reshape(-1)
            keep.append(b)
            planes.append(b.array.reshape(a.shape))
        return planes

    sp = pinned_copy(src) if memory != "pageable" else src.planes
    dp = pinned_copy(dst) if memory == "pinned_both" else dst.planes
    for _ in range(2):          # the second call re-uses rings, events and streams
        assert ctx.scale(sp, src.strides, dp, dst.strides, 0, case["sh"]) == case["dh"], ctx.last_error
    got = [np.array(p[:, :rb]) for p, (rows, rb) in zip(dp, dst.layout)]
    name = ctx.kernel_name
    ctx.close()
    for b in keep:
        b.close()
    assert T.first_diff(got, want.valid()) is None, name


@pytest.mark.parametrize("case", [BANDED_CASES[0], BANDED_CASES[3], BANDED_CASES[6], BANDED_CASES[10],
                                  dict(sw=322, sh=242, sf="yuv422p", dw=400, dh=300, df="bgra", flags=S.SWS_BICUBIC | BX)],
                         ids=_ids)
@pytest.mark.parametrize("flip", ["src", "dst", "both"])
def test_negative_strides(case, flip):
    """Bottom-up frames: pointers at the last row in memory, negative line sizes (vflip, BMP-style buffers)."""
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(43)
    want, _ = T.run_reference(src=src, **case)
    ctx = S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
    dst = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
    sp, ss = list(src.planes), list(src.strides)
    dp, ds = list(dst.planes), list(dst.strides)
    if flip in ("src", "both"):      # store the picture upside down in memory, describe it with negative strides
        flipped = [np.ascontiguousarray(a[::-1]) for a in src.planes]
        sp = [a[-1:].ctypes.data for a in flipped]
        ss = [-s for s in src.strides]
    if flip in ("dst", "both"):
        dp = [a[-1:].ctypes.data for a in dst.planes]
        ds = [-s for s in dst.strides]
    assert ctx.scale(sp, ss, dp, ds, 0, case["sh"]) == case["dh"], ctx.last_error
    got = dst.valid()
    if flip in ("dst", "both"):
        got = [g[::-1] for g in got]
    ctx.close()
    assert T.first_diff(got, want.valid()) is None


@pytest.mark.parametrize("slices", [[(120, 120), (0, 120)], [(176, 64), (112, 64), (0, 112)], [(238, 2), (0, 238)]])
@pytest.mark.parametrize("df,dw,dh", [("rgb24", 320, 240), ("yuv420p", 200, 150), ("nv12", 640, 480)])
def test_bottom_up_slices(slices, df, dw, dh):
    """sws_scale() fed from the last slice to the first (reference swscale.c:1141-1159)."""
    case = dict(sw=320, sh=240, sf="yuv420p", dw=dw, dh=dh, df=df, flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p", 320, 240).randomize(17)
    want, _ = T.run_reference(src=src, slices=slices, **case)
    got, name = T.run_cuda(src=src, slices=slices, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("case", [BANDED_CASES[0], BANDED_CASES[3], BANDED_CASES[7]], ids=_ids)
def test_batch_host_equals_per_frame(case, pinned):
    """sws_cuda_scale_batch_host(): 7 distinct host frames, three in flight on every visible device, must equal
    frame-by-frame sws_scale()."""
    frames = 7
    ctx = S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
    sl = T.plane_layout(case["sf"], case["sw"], case["sh"])
    dl = T.plane_layout(case["df"], case["dw"], case["dh"])
    srcs = [T.Frame(case["sf"], case["sw"], case["sh"]).randomize(300 + f) for f in range(frames)]
    keep = []

    def alloc(n):
        if pinned:
            b = S.PinnedBuffer(n)
            keep.append(b)
            return b.array
        return np.zeros(n, np.uint8)

    sbuf = [alloc(frames * rows * rb).reshape(frames, rows, rb) for rows, rb in sl]
    dbuf = [alloc(frames * rows * rb).reshape(frames, rows, rb) for rows, rb in dl]
    for f in range(frames):
        for i, (rows, rb) in enumerate(sl):
            sbuf[i][f] = srcs[f].planes[i][:, :rb]
    for d in dbuf:
        d[:] = 0
    r = ctx.scale_batch_host([a.ctypes.data for a in sbuf], [rb for _, rb in sl], [rows * rb for rows, rb in sl],
                             [a.ctypes.data for a in dbuf], [rb for _, rb in dl], [rows * rb for rows, rb in dl],
                             frames, 0, 0)
    assert r == case["dh"], ctx.last_error
    for f in range(frames):
        one = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
        assert ctx.scale(srcs[f].planes, srcs[f].strides, one.planes, one.strides, 0, case["sh"]) == case["dh"]
        got = [np.array(dbuf[i][f]) for i in range(len(dl))]
        assert T.first_diff(got, one.valid()) is None, "frame %d" % f
    ctx.close()
    for b in keep:
        b.close()


def test_batch_validation():
    """sws_cuda_scale_batch(): NULL stride arrays are an error, not a crash."""
    import ctypes as C
    ctx = S.SwsContext(64, 48, "yuv420p", 64, 48, "rgb24", S.SWS_BICUBIC | BX)
    L = S.lib()
    sp = (C.c_void_p * 4)(1 << 20, 1 << 20, 1 << 20, 0)
    dp = (C.c_void_p * 4)(1 << 20, 0, 0, 0)
    st = (C.c_int * 4)(64, 32, 32, 0)
    fs = (C.c_int64 * 4)(0, 0, 0, 0)
    EINVAL = -22
    assert L.sws_cuda_scale_batch(ctx.p, sp, None, fs, dp, st, fs, 1) == EINVAL
    assert L.sws_cuda_scale_batch(ctx.p, sp, st, None, dp, st, fs, 2) == EINVAL
    assert L.sws_cuda_scale_batch(ctx.p, sp, st, fs, dp, st, None, 2) == EINVAL
    ctx.close()
