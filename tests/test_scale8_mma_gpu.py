"""The tensor-pipe horizontal stage of the scaling kernel (sws_scale8_kernel<KS, RGB, MMA = true>): every
source layout, destination kind, K-step count and edge geometry against the reference C path, and against
the dot-product variant of the same kernel (SWS_B200_DISABLE=s8mma)."""
import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.BX


@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "nv21", "yuv422p", "yuv444p"])
@pytest.mark.parametrize("df", ["yuv420p", "nv12", "yuv444p", "rgb24", "bgra"])
@pytest.mark.parametrize("geom,flags", [
    ((1280, 720, 320, 180), S.SWS_BICUBIC),      # 4:1, 16 taps: K = 64
    ((642, 362, 320, 180), S.SWS_BICUBIC),       # ~2:1, odd sizes
    ((320, 180, 1000, 562), S.SWS_BICUBIC),      # upscale: K = 32
    ((1000, 600, 130, 70), S.SWS_BILINEAR),      # ~8:1 bilinear, 16 taps, K = 128
    ((700, 400, 333, 190), S.SWS_LANCZOS),       # 2.1:1 lanczos
    ((352, 288, 350, 290), S.SWS_SPLINE),        # ~1:1 with real taps
    ((64, 48, 24, 20), S.SWS_AREA),              # tiles narrower than one group row
])
def test_mma_hstage_matches_reference(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    src = T.Frame(sf, sw, sh).randomize(61, "noise")
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
    # a 4:2:2 / 4:4:4 source into a vertically subsampled destination doubles the vertical chroma ratio: past 20 taps
    # the bank needs a second record per row, which only the dot-product variant carries
    vsub_s = sf in ("yuv420p", "nv12", "nv21")
    vsub_d = df in ("yuv420p", "nv12")
    if name.startswith("scale8") and not (vsub_d and not vsub_s):
        assert name == "scale8_mma", name


@pytest.mark.parametrize("mode", ["noise", "extreme"])
def test_c4_full_size_uses_tensor_pipe(mode):
    case = dict(sw=7680, sh=4320, sf="nv12", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("nv12", 7680, 4320).randomize(5, mode)
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert name == "scale8_mma"
    assert T.first_diff(got.valid(), want.valid()) is None


@pytest.mark.parametrize("disable,expect", [("s8mma", "scale8_dp4a")])
def test_older_variants_still_selectable(disable, expect, monkeypatch):
    monkeypatch.setenv("SWS_B200_DISABLE", disable)
    case = dict(sw=1280, sh=720, sf="nv12", dw=320, dh=180, df="yuv420p", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("nv12", 1280, 720).randomize(3)
    want, _ = T.run_reference(src=src, **case)
    got, name = T.run_cuda(src=src, **case)
    assert name == expect
    assert T.first_diff(got.valid(), want.valid()) is None


@pytest.mark.parametrize("slices", [[(0, 128), (128, 232)], [(0, 64), (64, 64), (128, 232)]])
def test_slices_take_the_row_aligned_paths(slices):
    """Destination row ranges that do not start on a 16-row block fall back to the dot-product V stage."""
    case = dict(sw=640, sh=360, sf="yuv420p", dw=320, dh=180, df="yuv420p", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p", 640, 360).randomize(9)
    want, _ = T.run_reference(src=src, slices=slices, **case)
    got, name = T.run_cuda(src=src, slices=slices, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
