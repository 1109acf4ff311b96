"""32-bit float destinations (north_star: the float path): grayf32le through yuv2plane1/X_float
(reference output.c:219-316) and gbrpf32le through yuv2gbrpf32_full_X_c (output.c:2536-2610).  Both are the
16-bit integer result of the 19-bit pipeline times 1.0f / 65535.0f -- one IEEE multiply -- so the bar here is
tighter than the stated 1 ULP: the bytes must be identical."""
import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.BX


@pytest.mark.parametrize("sf", ["yuv420p", "yuv422p", "yuv444p", "nv12", "yuv420p10le", "yuv444p16le", "p010le",
                                "rgb24", "bgra"])
@pytest.mark.parametrize("df", ["grayf32le", "gbrpf32le"])
@pytest.mark.parametrize("geom,flags", [((320, 240, 320, 240), S.SWS_BICUBIC), ((322, 242, 400, 300), S.SWS_BICUBIC),
                                        ((642, 362, 161, 91), S.SWS_LANCZOS), ((320, 240, 320, 240), S.SWS_POINT),
                                        ((320, 240, 333, 240), S.SWS_BILINEAR)])
@pytest.mark.parametrize("bx", [0, BX])
def test_float_destinations(sf, df, geom, flags, bx):
    sw, sh, dw, dh = geom
    if sf in ("rgb24", "bgra") and df == "gbrpf32le":
        pytest.skip("packed RGB -> planar RGB is a different converter family (rgbToPlanarRgbWrapper)")
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | bx)
    src = T.Frame(sf, sw, sh).randomize(101, "noise")
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
    f = np.concatenate([p.reshape(-1) for p in got.valid()]).view(np.float32)
    assert np.all(np.isfinite(f)) and f.min() >= 0.0 and f.max() <= 1.0


@pytest.mark.parametrize("df", ["grayf32le", "gbrpf32le"])
def test_float_range_and_colourspace(df):
    """full-range bt709 source with picture controls: the colour constants reach the float writers"""
    case = dict(sw=352, sh=288, sf="yuv420p", dw=352, dh=288, df=df, flags=S.SWS_BICUBIC | BX)
    cs = (1, 1, 1, 0, 2 << 10, (1 << 16) + 3000, (1 << 16) - 4000)
    src = T.Frame("yuv420p", 352, 288).randomize(7, "extreme")
    want, _ = T.run_reference(src=src, colorspace=cs, **case)
    got, name = T.run_cuda(src=src, colorspace=cs, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name


@pytest.mark.parametrize("df", ["grayf32le", "yuv444p16le", "yuv420p16le", "rgb48le", "bgr48le", "gbrpf32le"])
@pytest.mark.parametrize("sf", ["yuvj444p", "yuvj420p"])
@pytest.mark.parametrize("flags", [S.SWS_SINC, S.SWS_LANCZOS])
def test_one_tap_rows_do_not_wrap_on_overshoot(sf, df, flags):
    """Horizontal sinc / lanczos upscale of full-range noise drives 19-bit lines below -2^19 (hScale*To19 only clips
    upwards); with one vertical tap the reference's _1 writers shift the line itself (yuv2plane1_16 / _float,
    yuv2rgba64[_full]_1), so nothing may wrap through a x 4096 (found by the fuzz: seed 301)."""
    case = dict(sw=53, sh=274, sf=sf, dw=266, dh=274, df=df, flags=flags | BX)
    src = T.Frame(sf, 53, 274).randomize(571967, "noise")
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
