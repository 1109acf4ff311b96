"""Cascaded contexts (reference utils.c:1803-1832 and 915-984, swscale.c:992-1020): conversions the reference
splits into two passes with an intermediate picture -- filters too long for one pass, and YUV -> YUV with two
different matrices (through packed RGB) -- must give the reference's bytes, with the intermediate in HBM."""
import zlib

import numpy as np
import pytest

from tests import sws_testlib as T
from tests.test_oracle_cpu import H, W, vsynth1  # noqa: F401
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.BX
FATE_YUV_COLORSPACE = 0xa0ea32ac      # tests/ref/fate/sws-yuv-colorspace


def test_fate_sws_yuv_colorspace(vsynth1):  # noqa: F811
    """fate-sws-yuv-colorspace: bt709 limited -> bt601 full, yuv420p 352x288, frame 1 of vsynth1."""
    c = S.SwsContext(W, H, "yuv420p", W, H, "yuv420p", S.SWS_BICUBIC | BX)
    assert c.set_colorspace(1, 0, 5, 1) >= 0, c.last_error
    assert c.kernel_name == "cascade"
    fr = vsynth1[:W * H * 3 // 2]
    planes = [np.ascontiguousarray(fr[:W * H]), np.ascontiguousarray(fr[W * H:W * H * 5 // 4]),
              np.ascontiguousarray(fr[W * H * 5 // 4:])]
    dst = [np.zeros(W * H, np.uint8), np.zeros(W * H // 4, np.uint8), np.zeros(W * H // 4, np.uint8)]
    assert c.scale(planes, [W, W // 2, W // 2], dst, [W, W // 2, W // 2]) == H
    assert zlib.adler32(b"".join(d.tobytes() for d in dst), 0) == FATE_YUV_COLORSPACE


@pytest.mark.parametrize("sf,df", [("yuv420p", "yuv420p"), ("nv12", "yuv444p"), ("yuv422p", "nv12"), ("yuv420p10le", "yuv420p"),
                                   # > 8-bit destinations go through bgr48le (utils.c:928-934): the 16-bit RGB readers
                                   ("yuv420p10le", "yuv420p10le"), ("yuv420p", "yuv444p16le"), ("yuv422p12le", "p010le"),
                                   # a float gray destination is neither isNBPS nor is16BPS: bgr24 again (fuzz seeds 402, 403)
                                   ("nv21", "grayf32le"), ("yuv420p12le", "grayf32le")])
@pytest.mark.parametrize("geom", [(352, 288, 352, 288), (640, 360, 320, 180), (320, 180, 500, 300)])
@pytest.mark.parametrize("cs", [(1, 0, 5, 1), (5, 1, 1, 0, 2000, 70000, 60000), (9, 0, 1, 0)])
def test_yuv_matrix_change(sf, df, geom, cs):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=S.SWS_BICUBIC | BX)
    src = T.Frame(sf, sw, sh).randomize(23)
    want, _ = T.run_reference(src=src, colorspace=cs, **case)
    got, name = T.run_cuda(src=src, colorspace=cs, **case)
    assert name == "cascade"
    assert T.first_diff(got.valid(), want.valid()) is None


@pytest.mark.parametrize("case", [
    dict(sw=4096, sh=64, sf="yuv420p", dw=8, dh=32, df="yuv420p", flags=S.SWS_BICUBIC | BX),       # 512:1 horizontally
    dict(sw=64, sh=4096, sf="yuv420p", dw=64, dh=8, df="rgb24", flags=S.SWS_BICUBIC | BX),         # vertically
    dict(sw=3000, sh=2000, sf="nv12", dw=10, dh=10, df="bgra", flags=S.SWS_LANCZOS | BX),
    dict(sw=2048, sh=2048, sf="rgb24", dw=8, dh=8, df="yuv420p", flags=S.SWS_GAUSS | BX),
])
def test_filters_too_long_for_one_pass(case):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(29, "smooth")
    want, info = T.run_reference(src=src, **case)
    assert info["cascaded"], "the reference did not cascade this case: pick a harder one"
    got, name = T.run_cuda(src=src, **case)
    assert name == "cascade"
    assert T.first_diff(got.valid(), want.valid()) is None


@pytest.mark.parametrize("slices", [[(0, 1000), (1000, 1048)], [(1024, 1024), (0, 1024)]])
def test_cascade_with_slices(slices):
    case = dict(sw=64, sh=2048, sf="yuv420p", dw=64, dh=4, df="yuv420p", flags=S.SWS_BICUBIC | BX)
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(31, "smooth")
    want, info = T.run_reference(src=src, **case)
    assert info["cascaded"]
    c = S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
    dst = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
    total = 0
    for (y, h) in slices:
        planes = [a[(y >> (1 if i else 0)):] for i, a in enumerate(src.planes)]
        r = c.scale(planes, src.strides, dst.planes, dst.strides, y, h)
        assert r >= 0, c.last_error
        total += r
    assert total == case["dh"]
    assert T.first_diff(dst.valid(), want.valid()) is None
