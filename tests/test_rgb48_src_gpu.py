"""rgb48le / bgr48le sources: rgb48ToY_c / rgb48ToUV_c / rgb48ToUV_half_c (input.c:111-196) in front of
hScale16To15_c / hScale16To19_c with the plain 16-bit shifts (swscale.c:69-125), and the unscaled 48 -> 48
converters (packedCopyWrapper, rgb48tobgr48_nobswap): bit-exact against the real reference build."""
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.BX


def _run(case, seed=81, mode="noise", **kw):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(seed, mode)
    want, _ = T.run_reference(src=src, **case, **kw)
    got, name = T.run_cuda(src=src, **case, **kw)
    assert T.first_diff(got.valid(), want.valid()) is None, (name, case)
    return name


@pytest.mark.parametrize("sf", ["rgb48le", "bgr48le"])
@pytest.mark.parametrize("df", ["yuv420p", "yuv444p", "nv12", "yuv420p10le", "yuv444p16le", "yuv422p16le", "p010le",
                                "rgb24", "bgra", "rgb565le", "grayf32le", "gbrpf32le"])
@pytest.mark.parametrize("geom,flags", [((322, 182, 322, 182), S.SWS_BICUBIC), ((323, 181, 323, 181), S.SWS_BICUBIC),
                                        ((640, 360, 320, 180), S.SWS_BICUBIC), ((322, 242, 400, 300), S.SWS_BILINEAR),
                                        ((642, 362, 161, 91), S.SWS_LANCZOS), ((320, 240, 333, 240), S.SWS_POINT)])
def test_rgb48_sources(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        _run(case, mode=mode)


@pytest.mark.parametrize("sf,df", [("rgb48le", "rgb48le"), ("rgb48le", "bgr48le"), ("bgr48le", "rgb48le"), ("bgr48le", "bgr48le")])
@pytest.mark.parametrize("geom", [(322, 182), (17, 9), (1920, 1080)])
@pytest.mark.parametrize("flags", [S.SWS_BICUBIC, S.SWS_BICUBIC | BX, S.SWS_POINT])
def test_rgb48_unscaled_copy_and_swap(sf, df, geom, flags):
    w, h = geom
    name = _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags))
    assert name == "rgb48", name
    slices = [(y, min(32, h - y)) for y in range(0, h, 32)]
    _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags), slices=slices)


@pytest.mark.parametrize("df", ["rgb48le", "bgr48le"])
def test_rgb48_to_rgb48_scaled(df):
    _run(dict(sw=322, sh=182, sf="rgb48le", dw=400, dh=300, df=df, flags=S.SWS_BICUBIC | BX))
    _run(dict(sw=322, sh=182, sf="bgr48le", dw=160, dh=90, df=df, flags=S.SWS_BILINEAR | BX))


def test_rgb48_source_odd_stride_is_rejected():
    import numpy as np
    c = S.SwsContext(64, 16, "rgb48le", 64, 16, "yuv420p", S.SWS_BICUBIC | BX)
    src = np.zeros((16, 64 * 6 + 1), np.uint8)
    dst = T.Frame("yuv420p", 64, 16, fill=0)
    assert c.scale([src], [64 * 6 + 1], dst.planes, dst.strides, 0, 16) < 0
    c.close()
