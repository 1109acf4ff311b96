"""An alpha channel carried through the scaler (reference input.c:455-471 rgbaToA_c / abgrToA_c, the hasAlpha
branches of yuv2rgb_{1,2,X}_c_template and yuv2rgb_full_{1,2,X}_c_template, output.c:1788-1939,2160-2330):
packed 32-bit RGB on both sides, scaled."""
import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.BX
RGB32 = ["rgba", "bgra", "argb", "abgr"]


@pytest.mark.parametrize("sf", RGB32)
@pytest.mark.parametrize("df", RGB32)
@pytest.mark.parametrize("geom,flags", [
    ((320, 240, 400, 300), S.SWS_BICUBIC),        # yuv2packedX
    ((320, 240, 160, 120), S.SWS_BICUBIC),
    ((320, 240, 640, 480), S.SWS_BILINEAR),       # yuv2packed2 rows
    ((320, 240, 640, 240), S.SWS_BICUBIC),        # horizontal only: yuv2packed1
    ((320, 240, 333, 251), S.SWS_LANCZOS),        # odd width: full-chroma writers
    ((321, 241, 160, 120), S.SWS_POINT),
    ((320, 240, 640, 480), S.SWS_FAST_BILINEAR),
])
@pytest.mark.parametrize("bx", [0, BX])
def test_rgb32_alpha_scaled(sf, df, geom, flags, bx):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | bx)
    src = T.Frame(sf, sw, sh).randomize(201, "noise")
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("mode", ["extreme", "smooth"])
@pytest.mark.parametrize("flags", [S.SWS_BICUBIC | S.SWS_FULL_CHR_H_INT | BX, S.SWS_SPLINE | BX, S.SWS_SINC])
def test_rgb32_alpha_overshoot(mode, flags):
    """saturated alpha edges under filters with negative lobes: the conditional clip of the X writers"""
    case = dict(sw=352, sh=288, sf="rgba", dw=500, dh=300, df="bgra", flags=flags)
    src = T.Frame("rgba", 352, 288).randomize(5, mode)
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
    assert len(np.unique(got.valid()[0].reshape(-1, 4)[:, 3])) > 1      # alpha really varies


def test_alpha_of_packed1_with_a_4096_0_chroma_row():
    """A vertical chroma offset turns the 1-tap chroma filter into {4096, 0} rows: yuv2packed1 is then called with
    uvalpha == 0 and takes its a * 255 alpha form, not (a + 64) >> 7 (output.c:1904-1905 vs 1929-1930; fuzz seed 301)."""
    case = dict(sw=512, sh=98, sf="argb", dw=230, dh=98, df="rgba", flags=S.SWS_FAST_BILINEAR | BX)
    src = T.Frame("argb", 512, 98).randomize(978701, "smooth")
    kw = dict(ctx_kwargs=dict(src_range=1, dst_range=1, chr_pos=(256, 256, 0, 128)))
    want, _ = T.run_reference(src=src, **case, **kw)
    got, name = T.run_cuda(src=src, **case, **kw)
    assert T.first_diff(got.valid(), want.valid()) is None, name
