"""The in-tree hook on the GPU: the REFERENCE's own sws_scale() / sws_scale_frame() (its libswscale
compiled with ff_sws_init_swscale_cuda(), integration/build_hooked.py) must hand the work to the B200
kernels -- the launch counter moves -- and return the bytes of the un-hooked reference build.
Conversions the CUDA path does not take must fall through to the reference's C kernels."""
import ctypes as C
import zlib

import numpy as np
import pytest

from tests import sws_testlib as T
from tests.test_oracle_cpu import FATE_SCALECHROMA, FATE_YUV_RANGE, H, W, vsynth1  # noqa: F401
from tests.test_parity_gpu import BASELINE_CASES
from oracle import refapi as R
from librempeg_b200 import swscale as S
from integration import hookedapi as HK

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not HK.available(), reason="integration/_build/libswsref_hooked.so not built")]


def _hooked_ctx(case, **kw):
    return HK.H.RefContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"], **kw)


def _run_hooked(case, src, slices=None, colorspace=None, expect_hooked=True, **kw):
    c = _hooked_ctx(case, **kw)
    if colorspace:
        c.set_colorspace(*colorspace)
    before = HK.launches(c)
    dst = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
    T._drive(c, src, dst, case["sh"], slices)
    after = HK.launches(c)
    name = HK.kernel(c)
    c.close()
    if expect_hooked:
        assert before >= 0, "the hook did not claim %r" % (case,)
        assert after > before, "no kernel was launched for %r" % (case,)
    else:
        assert after == -1, "%r should have stayed on the C kernels (kernel %s)" % (case, name)
    return dst, name


@pytest.mark.parametrize("case", BASELINE_CASES, ids=lambda c: "%dx%d_%s_to_%dx%d_%s_%x" % (
    c["sw"], c["sh"], c["sf"], c["dw"], c["dh"], c["df"], c["flags"]))
def test_reference_sws_scale_runs_on_gpu(case):
    """BASELINE configs C1..C5 through the reference's sws_scale(): GPU ran, bytes equal the C path."""
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(4321)
    want, _ = T.run_reference(src=src, **case)
    got, name = _run_hooked(case, src)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("case", BASELINE_CASES[1:5], ids=lambda c: "%s_%dx%d" % (c["sf"], c["dw"], c["dh"]))
def test_reference_sws_scale_frame_legacy_runs_on_gpu(case):
    """sws_scale_frame() on a legacy (sws_init_context'd) context: sws_frame_start / send / receive."""
    def run(api, hooked):
        c = api.RefContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
        s = api.RefFrame(case["sw"], case["sh"], case["sf"])
        d = api.RefFrame(case["dw"], case["dh"], case["df"])
        lay = T.plane_layout(case["sf"], case["sw"], case["sh"])
        src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(99)
        for i, (rows, rb) in enumerate(lay):
            a, ls = s.plane(i, rows)
            a[:, :rb] = src.planes[i][:, :rb]
        assert api.lib().swsref_scale_frame(c.h, d.f, s.f) >= 0
        out = []
        for i, (rows, rb) in enumerate(T.plane_layout(case["df"], case["dw"], case["dh"])):
            a, ls = d.plane(i, rows)
            out.append(a[:, :rb].copy())
        if hooked:
            assert HK.launches(c) > 0
        c.close(); s.close(); d.close()
        return out
    assert T.first_diff(run(HK.H, True), run(R, False)) is None


def test_reference_sws_scale_frame_dynamic_runs_on_gpu():
    """The frame-described mode (vf_scale's call): the graph's legacy pass is a hooked context and the
    whole frame is ONE slice (align = 0)."""
    w, h, dw, dh = 1280, 720, 640, 360
    outs = []
    for api in (R, HK.H):
        L = api.lib()
        L.swsref_scale_frame_dynamic.restype = C.c_int
        L.swsref_scale_frame_dynamic.argtypes = [C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        s = api.RefFrame(w, h, "yuv420p")
        d = api.RefFrame(dw, dh, "rgb24")
        src = T.Frame("yuv420p", w, h).randomize(7)
        for i, (rows, rb) in enumerate(src.layout):
            a, ls = s.plane(i, rows)
            a[:, :rb] = src.planes[i][:, :rb]
        props = (C.c_int * 6)(1, 1, 1, 2, 0, 0)      # bt709 limited, left-sited chroma -> full-range RGB
        before = HK.slices_total() if api is HK.H else 0
        assert L.swsref_scale_frame_dynamic(S.SWS_BICUBIC | S.BX, 4, d.f, s.f, props) >= 0
        if api is HK.H:
            assert HK.slices_total() == before + 1      # four threads requested, one launch
        a, ls = d.plane(0, dh)
        outs.append(a[:, :dw * 3].copy())
        s.close(); d.close()
    assert np.array_equal(outs[0], outs[1])


def test_fate_scalechroma_through_hooked_reference(vsynth1):  # noqa: F811
    """tests/ref/fate/filter-scalechroma: the reference's golden CRCs out of the reference's own sws_scale(),
    computed by the B200 kernel."""
    c = HK.H.RefContext(W, H, "yuv444p", W, H, "yuv420p", S.SWS_BICUBIC | S.SWS_BITEXACT, chr_pos=(-513, -513, 0, 256))
    fs = W * H * 3
    crcs = []
    for i in range(15):
        fr = vsynth1[i * fs:(i + 1) * fs]
        planes = [np.ascontiguousarray(fr[k * W * H:(k + 1) * W * H]) for k in range(3)]
        dst = [np.zeros(W * H, np.uint8), np.zeros(W * H // 4, np.uint8), np.zeros(W * H // 4, np.uint8)]
        assert c.scale(planes, [W, W, W], dst, [W, W // 2, W // 2]) == H
        crcs.append(zlib.adler32(b"".join(d.tobytes() for d in dst), 0))
    assert HK.launches(c) >= 15
    c.close()
    assert crcs == FATE_SCALECHROMA


def test_fate_yuv_range_through_hooked_reference(vsynth1):  # noqa: F811
    c = HK.H.RefContext(W, H, "yuv420p", W, H, "yuv420p", S.SWS_BICUBIC | S.BX, src_range=0, dst_range=1)
    fr = vsynth1[:W * H * 3 // 2]
    planes = [np.ascontiguousarray(fr[:W * H]), np.ascontiguousarray(fr[W * H:W * H * 5 // 4]),
              np.ascontiguousarray(fr[W * H * 5 // 4:])]
    dst = [np.zeros(W * H, np.uint8), np.zeros(W * H // 4, np.uint8), np.zeros(W * H // 4, np.uint8)]
    assert c.scale(planes, [W, W // 2, W // 2], dst, [W, W // 2, W // 2]) == H
    assert HK.launches(c) > 0
    c.close()
    assert zlib.adler32(b"".join(d.tobytes() for d in dst), 0) == FATE_YUV_RANGE


def test_set_colorspace_is_forwarded():
    """sws_setColorspaceDetails() after init must reach the B200 context (BT.709 full-range tables)."""
    case = dict(sw=640, sh=360, sf="yuv420p", dw=640, dh=360, df="bgra", flags=S.SWS_BICUBIC | S.BX)
    cs = (1, 1, 1, 0, 3 << 10, (1 << 16) + 5000, (1 << 16) - 7000)
    src = T.Frame("yuv420p", 640, 360).randomize(3)
    want, _ = T.run_reference(src=src, colorspace=cs, **case)
    got, name = _run_hooked(case, src, colorspace=cs)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("slices", [[(0, 120), (120, 120)], [(0, 64), (64, 64), (128, 112)],
                                    [(120, 120), (0, 120)], [(176, 64), (112, 64), (0, 112)]])
@pytest.mark.parametrize("df,dw,dh", [("rgb24", 320, 240), ("yuv420p", 200, 150)])
def test_slices_both_directions(slices, df, dw, dh):
    """Top-down and bottom-up slice sequences: the reference flips bottom-up slices into negative strides
    (swscale.c:1141-1159) before the hook sees them."""
    case = dict(sw=320, sh=240, sf="yuv420p", dw=dw, dh=dh, df=df, flags=S.SWS_BICUBIC | S.BX)
    src = T.Frame("yuv420p", 320, 240).randomize(17)
    want, _ = T.run_reference(src=src, slices=slices, **case)
    got, name = _run_hooked(case, src, slices=slices)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("case", [
    # what the B200 library declines at init stays on the reference's C kernels, silently and correctly
    dict(sw=320, sh=240, sf="gray", dw=160, dh=120, df="rgb24", flags=S.SWS_BICUBIC | S.BX),                   # gray source
    dict(sw=320, sh=240, sf="pal8", dw=320, dh=240, df="rgb24", flags=S.SWS_BICUBIC | S.BX),                   # palette source
    dict(sw=320, sh=240, sf="yuv420p", dw=160, dh=120, df="gbrp", flags=S.SWS_BICUBIC | S.BX),
])
def test_unsupported_conversions_fall_through(case):
    kw = {}
    c0 = dict(case)
    if "dither" in c0:
        kw["dither"] = c0.pop("dither")
    sf = c0["sf"]
    if sf == "pal8":
        pytest.skip("palette frames need a second plane the test helpers do not build")
    src = T.Frame(sf, c0["sw"], c0["sh"]).randomize(21)
    want, _ = T.run_reference(src=src, ctx_kwargs=kw, **c0)
    got, name = _run_hooked(c0, src, expect_hooked=False, **kw)
    assert T.first_diff(got.valid(), want.valid()) is None


def test_caller_filter_keeps_hook_out():
    """A caller-supplied SwsFilter changes the FIR banks: the hook sees banks that differ from the ones the
    B200 library would build from the options alone and must step aside."""
    case = dict(sw=320, sh=240, sf="yuv420p", dw=320, dh=240, df="yuv420p", flags=S.SWS_BICUBIC | S.BX)
    blur = {"lumH": [0.25, 0.5, 0.25], "lumV": [0.25, 0.5, 0.25]}
    c = _hooked_ctx(case, src_filter=blur)
    assert HK.launches(c) == -1
    c.close()
