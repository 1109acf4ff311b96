"""GPU parity: sws_scale() through the C ABI of libswscale_b200.so versus the real
reference C path (oracle/_ref), on identical seeded host frames.  Bit-exact or fail.

Run with `pytest -m gpu` on the B200 box.
"""
import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu

BX = S.BX


def _check(**case):
    got, want, name = T.run_case_both(**case)
    diff = T.first_diff(got, want)
    assert diff is None, "%r via %s: %s" % (case, name, diff)
    return name


# ---- the five BASELINE.json configurations (C1..C5; C5 is one frame of the batch) ----
BASELINE_CASES = [
    dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=S.SWS_POINT | S.SWS_BITEXACT),
    dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=S.SWS_POINT | BX),
    dict(sw=1920, sh=1080, sf="yuv420p", dw=1920, dh=1080, df="rgb24", flags=S.SWS_BICUBIC | BX),
    dict(sw=3840, sh=2160, sf="yuv420p10le", dw=3840, dh=2160, df="rgb48le", flags=S.SWS_LANCZOS | BX),
    dict(sw=7680, sh=4320, sf="nv12", dw=1920, dh=1080, df="yuv420p", flags=S.SWS_BICUBIC | BX),
    dict(sw=3840, sh=2160, sf="yuv420p", dw=3840, dh=2160, df="rgb24", flags=S.SWS_BICUBIC | BX),
    # default flags: the reference takes the unscaled LUT converter
    dict(sw=1920, sh=1080, sf="yuv420p", dw=1920, dh=1080, df="rgb24", flags=S.SWS_BICUBIC),
]


@pytest.mark.parametrize("case", BASELINE_CASES, ids=lambda c: "%dx%d_%s_to_%dx%d_%s_%x" % (
    c["sw"], c["sh"], c["sf"], c["dw"], c["dh"], c["df"], c["flags"]))
@pytest.mark.parametrize("mode", ["noise", "extreme"])
def test_baseline_configs(case, mode):
    _check(mode=mode, seed=1234, **case)


# ---- scaling ratios, scalers, odd sizes ----
SCALERS = [S.SWS_POINT, S.SWS_BILINEAR, S.SWS_BICUBIC, S.SWS_AREA, S.SWS_GAUSS, S.SWS_SINC,
           S.SWS_LANCZOS, S.SWS_SPLINE, S.SWS_X, S.SWS_BICUBLIN]


@pytest.mark.parametrize("scaler", SCALERS)
@pytest.mark.parametrize("geom", [(352, 288, 200, 100), (352, 288, 704, 576), (320, 240, 333, 251),
                                  (642, 362, 320, 180)])
def test_yuv420p_to_yuv420p_scalers(scaler, geom):
    sw, sh, dw, dh = geom
    _check(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df="yuv420p", flags=scaler | BX, seed=3)


@pytest.mark.parametrize("scaler", [S.SWS_POINT, S.SWS_BILINEAR, S.SWS_BICUBIC, S.SWS_LANCZOS])
@pytest.mark.parametrize("geom", [(352, 288, 200, 100), (352, 288, 704, 576), (320, 240, 334, 251),
                                  (1920, 1080, 1280, 720)])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr", "rgb48le", "bgr48le"])
def test_yuv420p_to_rgb_scaled(scaler, geom, df):
    sw, sh, dw, dh = geom
    _check(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df=df, flags=scaler | BX, seed=5)


@pytest.mark.parametrize("sf", ["yuv420p", "yuv422p", "nv12", "nv21", "yuv420p10le", "yuv422p10le",
                                "yuv420p9le", "yuv420p12le", "yuv420p14le", "yuv420p16le", "yuvj420p"])
@pytest.mark.parametrize("df", ["rgb24", "rgb48le", "yuv420p", "yuv422p", "yuv444p", "nv12", "nv21",
                                "yuv420p10le", "yuv444p12le", "yuv420p16le"])
def test_format_matrix_scaled(sf, df):
    _check(sw=322, sh=242, sf=sf, dw=400, dh=300, df=df, flags=S.SWS_BICUBIC | BX, seed=11)
    _check(sw=322, sh=242, sf=sf, dw=160, dh=120, df=df, flags=S.SWS_BILINEAR | BX, seed=12)


@pytest.mark.parametrize("sf,df", [("yuv420p", "rgb24"), ("yuv420p", "bgra"), ("yuv422p", "rgb24"),
                                   ("yuv420p", "rgb48le"), ("yuv420p", "nv12"), ("nv12", "yuv420p"),
                                   ("yuv420p", "yuv420p"), ("yuv420p10le", "yuv420p10le"),
                                   ("yuv420p", "yuv422p"), ("yuv420p10le", "rgb48le"),
                                   ("yuv420p10le", "rgb24")])
@pytest.mark.parametrize("flags", [S.SWS_BICUBIC, S.SWS_BICUBIC | BX, S.SWS_POINT, S.SWS_BILINEAR | BX])
def test_same_size(sf, df, flags):
    _check(sw=642, sh=362, sf=sf, dw=642, dh=362, df=df, flags=flags, seed=21)


def test_strided_buffers_and_padding_untouched():
    """Padded strides on both sides; bytes outside the image must not be written."""
    case = dict(sw=318, sh=200, sf="yuv420p", dw=318, dh=200, df="rgb24", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p", 318, 200, pad=32).randomize(4)
    want, _ = T.run_reference(src=src, dst_pad=64, **case)
    got, _ = T.run_cuda(src=src, dst_pad=64, **case)
    for g, w in zip(got.planes, want.planes):
        assert np.array_equal(g, w)      # includes the (zero) padding columns


@pytest.mark.parametrize("cs", [(1, 0, 1, 0), (5, 1, 5, 0), (9, 0, 9, 0), (7, 1, 7, 1)])
@pytest.mark.parametrize("df", ["rgb24", "rgb48le"])
def test_colorspace_details(cs, df):
    """BT.709 / full range / BT.2020 / SMPTE240M through sws_setColorspaceDetails()."""
    src_cs, src_range, dst_cs, dst_range = cs
    colorspace = (src_cs, src_range, dst_cs, dst_range, 0, 1 << 16, 1 << 16)
    _check(sw=320, sh=240, sf="yuv420p", dw=320, dh=240, df=df, flags=S.SWS_BICUBIC | BX,
           seed=8, colorspace=colorspace)
    _check(sw=320, sh=240, sf="yuv420p", dw=480, dh=360, df=df, flags=S.SWS_BICUBIC | BX,
           seed=9, colorspace=colorspace)


def test_brightness_contrast_saturation():
    colorspace = (5, 0, 5, 0, 3000, int(1.2 * 65536), int(0.8 * 65536))
    _check(sw=320, sh=240, sf="yuv420p", dw=320, dh=240, df="rgb24", flags=S.SWS_BICUBIC | BX,
           seed=8, colorspace=colorspace)


@pytest.mark.parametrize("ranges", [(0, 1), (1, 0)])
@pytest.mark.parametrize("df", ["yuv420p", "yuv420p10le", "yuv420p16le"])
def test_range_conversion(ranges, df):
    """limited<->full conversion on the h-scaled lines (reference swscale.c:163-255)."""
    _check(sw=320, sh=240, sf="yuv420p", dw=400, dh=300, df=df, flags=S.SWS_BICUBIC | BX, seed=13,
           ctx_kwargs=dict(src_range=ranges[0], dst_range=ranges[1]))
    _check(sw=320, sh=240, sf="yuv420p", dw=320, dh=240, df=df, flags=S.SWS_BICUBIC | BX, seed=14,
           ctx_kwargs=dict(src_range=ranges[0], dst_range=ranges[1]))


@pytest.mark.parametrize("chr_pos", [(0, 128, -513, -513), (0, 0, -513, -513), (128, 128, 0, 0)])
def test_chroma_siting(chr_pos):
    """left / top-left chroma siting makes the horizontal chroma FIR a half-sample shift (§3.3)."""
    _check(sw=320, sh=240, sf="yuv420p", dw=320, dh=240, df="rgb24", flags=S.SWS_BICUBIC | BX,
           seed=15, ctx_kwargs=dict(chr_pos=chr_pos))
    _check(sw=320, sh=240, sf="yuv420p", dw=200, dh=150, df="yuv420p", flags=S.SWS_BICUBIC | BX,
           seed=16, ctx_kwargs=dict(chr_pos=chr_pos))


@pytest.mark.parametrize("case", [
    dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=S.SWS_BICUBIC | BX),
    dict(sw=640, sh=480, sf="yuv420p", dw=320, dh=200, df="yuv420p", flags=S.SWS_BICUBIC | BX),
    dict(sw=320, sh=240, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=S.SWS_LANCZOS | BX),
    dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=S.SWS_BICUBIC),
])
@pytest.mark.parametrize("slice_h", [2, 16, 64, 200])
def test_slices_top_down(case, slice_h):
    """The legacy slice API: feed the source in horizontal bands (reference swscale.h:556-585)."""
    sh = case["sh"]
    slices = [(y, min(slice_h, sh - y)) for y in range(0, sh, slice_h)]
    _check(seed=17, slices=slices, **case)


def test_tiny_and_ragged_sizes():
    for (sw, sh, dw, dh) in [(16, 16, 16, 16), (18, 10, 34, 22), (1920, 2, 1920, 2), (4, 1080, 4, 1080),
                             (130, 66, 62, 30), (34, 34, 1280, 720)]:
        _check(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df="rgb24", flags=S.SWS_BICUBIC | BX, seed=19)
        _check(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df="yuv420p", flags=S.SWS_BICUBIC | BX, seed=20)


def test_unsupported_fails_loudly():
    """No silent CPU fallback: what the CUDA path does not implement must fail at init."""
    with pytest.raises(RuntimeError):
        S.SwsContext(320, 240, "yuv420p", 320, 240, "rgb24", S.SWS_BICUBIC | (1 << 16))          # vertical chroma drop
    with pytest.raises(RuntimeError):
        S.SwsContext(320, 240, "rgb565le", 320, 240, "yuv420p", S.SWS_BICUBIC | BX)      # 15/16 bpp RGB readers


@pytest.mark.parametrize("sf", ["yuv444p", "yuv420p", "yuv422p", "yuv444p10le", "yuv420p12le", "nv12"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr", "rgb48le", "bgr48le"])
@pytest.mark.parametrize("geom,flags", [((322, 242, 323, 242), S.SWS_BICUBIC | BX),
                                        ((322, 242, 401, 301), S.SWS_BILINEAR | BX),
                                        ((322, 242, 161, 121), S.SWS_LANCZOS | BX),
                                        ((322, 242, 322, 300), S.SWS_BILINEAR | BX | S.SWS_FULL_CHR_H_INT),
                                        ((322, 242, 322, 242), S.SWS_POINT | BX | S.SWS_FULL_CHR_H_INT)])
def test_full_chroma_rgb(sf, df, geom, flags):
    """SWS_FULL_CHR_H_INT (forced by odd widths and 4:4:4 sources, utils.c:1270-1286): per-pixel chroma and
    the arithmetic colour step of yuv2rgb_write_full (output.c:1998-2051) incl. the bias-free 2-tap variants."""
    sw, sh, dw, dh = geom
    _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags, seed=41)


# ---- every instantiation of the same-size 8-bit 4:2:0 kernel (3 chroma layouts x 6 byte orders) ----
@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "nv21", "yuv422p", "yuvj422p"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("geom,flags", [((644, 366), S.SWS_BICUBIC | BX),          # ragged right/bottom tiles
                                        ((640, 200), S.SWS_BICUBIC | BX),          # 128 x 64 tile shape, ragged bottom
                                        ((384, 70), S.SWS_BILINEAR | BX),          # 128 x 64 tile shape
                                        ((256, 34), S.SWS_BILINEAR | BX),
                                        ((1280, 720), S.SWS_POINT | BX),
                                        ((648, 360), S.SWS_BICUBIC)])
def test_fast420_instantiations(sf, df, geom, flags):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=77, mode=mode)
        assert name.startswith("fast420"), name


# ---- packed 8-bit RGB sources: rgb24ToY/UV[_half] and the 32-bit template readers (input.c:264-345,1068-1180) ----
RGB_SRC = ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"]


@pytest.mark.parametrize("sf", RGB_SRC)
@pytest.mark.parametrize("df", ["yuv420p", "yuv422p", "yuv444p", "nv12", "nv21", "yuv420p10le", "yuv444p16le",
                                "yuvj420p", "rgb24", "bgr24", "rgb48le"])
@pytest.mark.parametrize("geom,flags", [((322, 182, 322, 182), S.SWS_BICUBIC | BX),
                                        ((323, 181, 323, 181), S.SWS_BICUBIC | BX),
                                        ((322, 182, 160, 90), S.SWS_BILINEAR | BX),
                                        ((322, 182, 500, 300), S.SWS_LANCZOS | BX),
                                        ((322, 182, 200, 182), S.SWS_BICUBIC | BX | S.SWS_FULL_CHR_H_INP),
                                        ((322, 182, 322, 182), S.SWS_POINT)])
def test_rgb_sources(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags, seed=91, mode=mode)


@pytest.mark.parametrize("cs", [1, 7, 9])
@pytest.mark.parametrize("sf", ["rgb24", "abgr"])
def test_rgb_sources_other_matrices(cs, sf):
    _check(sw=322, sh=182, sf=sf, dw=322, dh=182, df="yuv420p", flags=S.SWS_BICUBIC | BX, seed=93,
           colorspace=(cs, 0, cs, 0, 0, 1 << 16, 1 << 16))


@pytest.mark.parametrize("case", [dict(sw=1920, sh=1080, sf="rgb24", dw=1920, dh=1080, df="yuv420p"),
                                  dict(sw=3840, sh=2160, sf="bgra", dw=1920, dh=1080, df="nv12"),
                                  dict(sw=1280, sh=720, sf="rgba", dw=1920, dh=1080, df="yuv420p10le")],
                         ids=lambda c: "%s_%d_to_%s_%d" % (c["sf"], c["sw"], c["df"], c["dw"]))
def test_rgb_sources_full_size(case):
    _check(flags=S.SWS_BICUBIC | BX, seed=95, **case)


# ---- unscaled special converters for packed RGB (swscale_unscaled.c:1843-2077,2453-2466) ----
@pytest.mark.parametrize("sf", RGB_SRC)
@pytest.mark.parametrize("df", RGB_SRC)
@pytest.mark.parametrize("geom,flags", [((322, 182), S.SWS_BICUBIC | BX), ((323, 181), S.SWS_BICUBIC),
                                        ((1030, 64), S.SWS_POINT), ((2050, 31), S.SWS_BICUBIC | S.SWS_BITEXACT)])
def test_rgb_to_rgb_unscaled(sf, df, geom, flags):
    """Byte shuffles (alpha carried or set to 255); with SWS_BITEXACT 24-bit -> rgba/bgra goes through the
    scaler instead, exactly as findRgbConvFn decides."""
    w, h = geom
    name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=97, src_pad=5, dst_pad=3)
    via_scaler = flags & S.SWS_BITEXACT and sf in ("rgb24", "bgr24") and df in ("rgba", "bgra")
    assert name in (("generic_tile", "scale_rgb_dp2a") if via_scaler else ("rgb_shuffle",)), name


@pytest.mark.parametrize("geom", [(322, 182), (322, 181), (1920, 1080), (2, 2), (2, 1)])
@pytest.mark.parametrize("flags", [S.SWS_BICUBIC, S.SWS_POINT | S.SWS_BITEXACT])
def test_bgr24_to_yuv420p_unscaled_box_converter(geom, flags):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf="bgr24", dw=w, dh=h, df="yuv420p", flags=flags, seed=99, mode=mode)
        assert name == "bgr24_to_yv12", name


# ---- dispatch order: every conversion must give the same bytes whichever kernel serves it ----
@pytest.mark.parametrize("case", [
    dict(sw=644, sh=366, sf="yuv420p", dw=644, dh=366, df="rgb24", flags=S.SWS_BICUBIC | BX),      # fast420
    dict(sw=644, sh=366, sf="yuv420p10le", dw=644, dh=366, df="rgb48le", flags=S.SWS_LANCZOS | BX),  # fast16
    dict(sw=1280, sh=720, sf="nv12", dw=640, dh=360, df="yuv420p", flags=S.SWS_BICUBIC | BX),        # scale8
    dict(sw=640, sh=360, sf="yuv420p", dw=1280, dh=720, df="bgra", flags=S.SWS_BICUBIC | BX),        # tile15
    dict(sw=644, sh=366, sf="rgb24", dw=644, dh=366, df="yuv420p", flags=S.SWS_BICUBIC | BX),        # tile15
    dict(sw=644, sh=366, sf="yuv420p10le", dw=320, dh=180, df="yuv420p", flags=S.SWS_BILINEAR | BX),  # tile15
], ids=lambda c: "%s_%d_to_%s_%d" % (c["sf"], c["sw"], c["df"], c["dw"]))
def test_kernel_fallback_chain_is_bit_identical(case, monkeypatch):
    first = _check(seed=111, **case)
    seen = [first]
    for off in ("fast420,fast16,scale8", "fast420,fast16,scale8,tile15"):
        monkeypatch.setenv("SWS_B200_DISABLE", off)
        seen.append(_check(seed=111, **case))
    assert seen[-1] == "generic_tile", seen
    assert first != "generic_tile", seen


# ---- identity 8-bit yuv -> yuv conversions: the copy / (de)interleave kernel (swscale_unscaled.c:147-215) ----
@pytest.mark.parametrize("sf,df", [("nv12", "yuv420p"), ("yuv420p", "nv12"), ("nv21", "nv12"), ("yuv420p", "nv21"),
                                   ("nv21", "yuv420p"), ("yuv420p", "yuv420p"), ("yuv422p", "yuv422p")])
@pytest.mark.parametrize("geom", [(644, 366), (34, 18), (1920, 1080), (322, 243)])
def test_copy8_kernel(sf, df, geom):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=S.SWS_BICUBIC | BX, seed=91, mode=mode)
        assert name == "copy8", name
    slices = [(y, min(32, h - y)) for y in range(0, h, 32)]
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=S.SWS_BILINEAR, seed=92, slices=slices)


# ---- the dot-product scaling kernel with packed RGB output (yuv2rgb_X/_1/_2 rounding, output.c:1662-1939) ----
@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "nv21", "yuv422p", "yuv444p"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("geom,flags", [((320, 180, 640, 360), S.SWS_BICUBIC | BX),
                                        ((640, 360, 320, 180), S.SWS_BILINEAR | BX),      # 2-tap rows: yuv2packed2
                                        ((322, 242, 400, 300), S.SWS_BILINEAR | BX),
                                        ((1280, 720, 642, 362), S.SWS_LANCZOS | BX),       # 13 taps
                                        ((322, 242, 644, 242), S.SWS_POINT | BX)])
def test_scale8_rgb_output(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags, seed=93, mode=mode)
        if sf != "yuv444p":           # 4:4:4 sources force full-chroma RGB: another kernel
            assert name.startswith("scale8"), name


# ---- packed RGB -> 4:2:0 of the same size: the fused reader + vertical-chroma kernel ----
@pytest.mark.parametrize("sf", RGB_SRC)
@pytest.mark.parametrize("df", ["yuv420p", "nv12", "nv21"])
@pytest.mark.parametrize("geom,flags", [((640, 360), S.SWS_BICUBIC | BX), ((1920, 1080), S.SWS_BILINEAR | BX),
                                        ((144, 37), S.SWS_LANCZOS | BX), ((320, 182), S.SWS_BICUBIC),
                                        ((656, 366), S.SWS_AREA | BX), ((48, 10), S.SWS_POINT | BX)])
def test_rgb420_kernel(sf, df, geom, flags):
    w, h = geom
    if sf == "bgr24" and df == "yuv420p" and not (flags & S.SWS_ACCURATE_RND):
        pytest.skip("bgr24ToYv12Wrapper special converter, covered elsewhere")
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=95, mode=mode)
        assert name == "rgb420", name
    slices = [(y, min(34, h - y)) for y in range(0, h, 34)]
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=96, slices=slices)


def test_rgb420_colorspace():
    """BT.709 limited-range destination matrix through the fused RGB reader (a full-range destination
    adds range conversion: test_range_switched_on_after_init)."""
    colorspace = (1, 0, 1, 0, 0, 1 << 16, 1 << 16)
    for sf in ("rgb24", "bgra"):
        name = _check(sw=640, sh=360, sf=sf, dw=640, dh=360, df="yuv420p", flags=S.SWS_BICUBIC | BX, seed=97,
                      colorspace=colorspace)
        assert name == "rgb420", name


@pytest.mark.parametrize("ranges", [(0, 1), (1, 0)])
@pytest.mark.parametrize("case", [dict(sw=640, sh=360, sf="yuv420p", dw=320, dh=180, df="yuv420p"),
                                  dict(sw=640, sh=360, sf="nv12", dw=640, dh=360, df="yuv420p"),
                                  dict(sw=640, sh=360, sf="rgb24", dw=640, dh=360, df="yuv420p"),
                                  dict(sw=640, sh=360, sf="bgra", dw=640, dh=360, df="nv12")])
def test_range_switched_on_after_init(case, ranges):
    """sws_setColorspaceDetails() after init turns range conversion on: the kernels chosen at init that
    do not implement it must step aside (reference swscale.c:626-660, utils.c:849-1005)."""
    colorspace = (5, ranges[0], 5, ranges[1], 0, 1 << 16, 1 << 16)
    _check(flags=S.SWS_BICUBIC | BX, seed=98, colorspace=colorspace, **case)


# ---- SWS_FAST_BILINEAR: ff_hyscale_fast_c / ff_hcscale_fast_c as 2-tap banks (hscale_fast_bilinear.c:23-55) ----
@pytest.mark.parametrize("sf,df", [("yuv420p", "yuv420p"), ("yuv420p", "rgb24"), ("nv12", "bgra"), ("yuv422p", "nv12"),
                                   ("yuv420p", "yuv420p10le"), ("yuv420p", "yuv420p16le"), ("yuv420p10le", "yuv420p"),
                                   ("yuv420p10le", "rgb48le"), ("rgb24", "yuv420p"), ("bgra", "nv12"),
                                   ("yuv444p", "rgb24"), ("nv21", "yuv444p")])
@pytest.mark.parametrize("geom", [(352, 288, 200, 100), (176, 144, 352, 288), (322, 242, 333, 251), (6, 16, 40, 30),
                                  (640, 360, 640, 200), (640, 360, 1280, 360), (1920, 1080, 1280, 720)])
@pytest.mark.parametrize("extra", [0, BX])
def test_fast_bilinear(sf, df, geom, extra):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=S.SWS_FAST_BILINEAR | extra, seed=99, mode=mode)


# ---- unscaled planar depth conversion: planarCopyWrapper's dithered / replicated copies (swscale_unscaled.c:2220-2384) ----
DEPTHS = ["yuv420p", "yuv420p9le", "yuv420p10le", "yuv420p12le", "yuv420p14le", "yuv420p16le"]


@pytest.mark.parametrize("sf", DEPTHS)
@pytest.mark.parametrize("df", DEPTHS)
@pytest.mark.parametrize("opts", [dict(), dict(src_range=1, dst_range=1), dict(dither=0)])
def test_depthcopy(sf, df, opts):
    if sf == df:
        pytest.skip("same format")
    for (w, h) in [(644, 366), (35, 19)]:
        for mode in ("noise", "extreme"):
            name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=S.SWS_BICUBIC | BX, seed=101, mode=mode,
                          ctx_kwargs=opts)
            assert name == "depthcopy", name
    if not opts:
        slices = [(y, min(16, 366 - y)) for y in range(0, 366, 16)]         # dither rows count from the slice start
        _check(sw=644, sh=366, sf=sf, dw=644, dh=366, df=df, flags=S.SWS_BICUBIC, seed=102, slices=slices)


@pytest.mark.parametrize("sf,df", [("yuv422p10le", "yuv422p"), ("yuv444p", "yuv444p12le"), ("yuv444p16le", "yuv444p10le")])
def test_depthcopy_other_subsamplings(sf, df):
    name = _check(sw=322, sh=243, sf=sf, dw=322, dh=243, df=df, flags=S.SWS_BILINEAR, seed=103)
    assert name == "depthcopy", name


# ---- p010le: p010LEToY/UV readers (input.c:950-1006), yuv2p010l1/lX/cX writers (output.c:538-589), the
# unscaled planarToP01xWrapper / planar8ToP01xleWrapper (swscale_unscaled.c:273-375) ----
@pytest.mark.parametrize("df", ["yuv420p", "nv12", "rgb24", "bgra", "yuv420p10le", "yuv444p16le", "rgb48le", "p010le"])
@pytest.mark.parametrize("geom,flags", [((322, 242, 400, 300), S.SWS_BICUBIC | BX), ((322, 242, 160, 120), S.SWS_BILINEAR | BX),
                                        ((322, 242, 322, 242), S.SWS_BICUBIC | BX), ((323, 241, 323, 241), S.SWS_POINT)])
def test_p010_source(df, geom, flags):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        _check(sw=sw, sh=sh, sf="p010le", dw=dw, dh=dh, df=df, flags=flags, seed=105, mode=mode)


@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "yuv422p", "yuv420p10le", "yuv444p12le", "yuv420p16le", "rgb24", "bgra"])
@pytest.mark.parametrize("geom,flags", [((322, 242, 400, 300), S.SWS_BICUBIC | BX), ((322, 242, 160, 120), S.SWS_LANCZOS | BX),
                                        ((322, 242, 322, 242), S.SWS_BICUBIC | BX)])
def test_p010_destination(sf, geom, flags):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df="p010le", flags=flags, seed=106, mode=mode)


@pytest.mark.parametrize("sf", ["yuv420p", "yuv420p10le", "yuv420p12le", "yuv420p14le", "yuv420p16le"])
@pytest.mark.parametrize("geom", [(644, 366), (35, 19), (34, 18)])
def test_p01x_unscaled_wrappers(sf, geom):
    w, h = geom
    name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="p010le", flags=S.SWS_BICUBIC, seed=107)
    assert name == "p01x", name
    slices = [(y, min(16, h - y)) for y in range(0, h, 16)]
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="p010le", flags=S.SWS_BICUBIC, seed=108, slices=slices)


@pytest.mark.parametrize("opts", [dict(), dict(src_range=1, dst_range=1)])
@pytest.mark.parametrize("geom", [(644, 366), (35, 19)])
def test_nv12_to_p010_unscaled(geom, opts):
    """planarCopyWrapper's COPY816 on semi-planar planes (swscale_unscaled.c:2266-2284)."""
    w, h = geom
    name = _check(sw=w, sh=h, sf="nv12", dw=w, dh=h, df="p010le", flags=S.SWS_BICUBIC, seed=109, ctx_kwargs=opts)
    assert name == "depthcopy", name


@pytest.mark.parametrize("opts", [dict(), dict(src_range=1, dst_range=1), dict(dither=0)])
@pytest.mark.parametrize("geom", [(644, 366), (35, 19), (640, 360), (1287, 33)])
def test_p010_to_nv12_unscaled(geom, opts):
    """planarCopyWrapper's DITHER_COPY on semi-planar planes, including its scalar tail (the last width & 7 samples of a
    row), which forgets the >> 6 of the p010 container (swscale_unscaled.c:2174-2176,2193-2195,2212-2214)."""
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf="p010le", dw=w, dh=h, df="nv12", flags=S.SWS_BICUBIC, seed=110, mode=mode, ctx_kwargs=opts)
        assert name == "depthcopy", name
    slices = [(y, min(16, h - y)) for y in range(0, h, 16)]
    _check(sw=w, sh=h, sf="p010le", dw=w, dh=h, df="nv12", flags=S.SWS_BICUBIC, seed=111, slices=slices, ctx_kwargs=opts)


# ---- 9..16-bit planar -> packed 8-bit RGB of the same size: the TMA kernel for 10-bit video to display RGB ----
@pytest.mark.parametrize("sf", ["yuv420p10le", "yuv420p9le", "yuv420p12le", "yuv420p14le", "yuv420p16le", "p010le",
                                "yuv422p10le", "yuv422p12le"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("geom,flags", [((644, 366), S.SWS_BICUBIC | BX), ((256, 34), S.SWS_BILINEAR | BX),
                                        ((1280, 720), S.SWS_LANCZOS | BX), ((648, 360), S.SWS_BICUBIC),
                                        ((132, 66), S.SWS_POINT)])
def test_fast420_hi8_instantiations(sf, df, geom, flags):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=111, mode=mode)
        assert name.startswith("fast420_hi8"), name


def test_fast420_hi8_colorspace_and_slices():
    colorspace = (9, 1, 9, 0, 0, 1 << 16, 1 << 16)              # BT.2020, full-range source
    _check(sw=640, sh=360, sf="yuv420p10le", dw=640, dh=360, df="bgra", flags=S.SWS_BICUBIC | BX, seed=112,
           colorspace=colorspace)
    slices = [(y, min(64, 360 - y)) for y in range(0, 360, 64)]
    _check(sw=640, sh=360, sf="yuv420p10le", dw=640, dh=360, df="rgb24", flags=S.SWS_BICUBIC | BX, seed=113, slices=slices)


# ---- SwsFilter pre-filters (vf_smartblur-style): blur / sharpen vectors convolved into the banks ----
_GAUSS5 = [0.06136, 0.24477, 0.38774, 0.24477, 0.06136]
_SHARP3 = [-0.25, 1.5, -0.25]


@pytest.mark.parametrize("filt", [dict(lumH=_GAUSS5, lumV=_GAUSS5), dict(lumH=_SHARP3, lumV=_SHARP3, chrH=_GAUSS5, chrV=_GAUSS5)])
@pytest.mark.parametrize("case", [dict(sw=640, sh=360, sf="yuv420p", dw=640, dh=360, df="yuv420p"),
                                  dict(sw=640, sh=360, sf="yuv420p", dw=640, dh=360, df="rgb24"),
                                  dict(sw=640, sh=360, sf="nv12", dw=400, dh=300, df="bgra"),
                                  dict(sw=322, sh=242, sf="yuv420p10le", dw=322, dh=242, df="yuv420p10le"),
                                  dict(sw=322, sh=242, sf="rgb24", dw=322, dh=242, df="yuv420p")])
def test_swsfilter_prefilter(case, filt):
    _check(flags=S.SWS_BICUBIC | BX, seed=115, ctx_kwargs=dict(src_filter=filt), **case)
    _check(flags=S.SWS_BILINEAR, seed=116, ctx_kwargs=dict(src_filter=filt, dst_filter=dict(lumH=_GAUSS5)), **case)


# ---- 8-bit 4:4:4 -> packed RGB of the same size: yuv2rgb_full_1_c + yuv2rgb_write_full with identity filters ----
@pytest.mark.parametrize("sf", ["yuv444p", "yuvj444p"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("geom,flags", [((644, 366), S.SWS_BICUBIC | BX), ((35, 19), S.SWS_BILINEAR), ((1280, 720), S.SWS_POINT | BX)])
def test_full444_kernel(sf, df, geom, flags):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=117, mode=mode)
        assert name == "full444", name
    slices = [(y, min(20, h - y)) for y in range(0, h, 20)]
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=118, slices=slices)
    colorspace = (1, 1, 1, 0, 0, 1 << 16, 1 << 16)
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags, seed=119, colorspace=colorspace)


# ---- packed RGB -> 8-bit planar 4:4:4 of the same size: full-resolution readers, identity filters ----
@pytest.mark.parametrize("sf", RGB_SRC)
@pytest.mark.parametrize("geom,flags", [((644, 366), S.SWS_BICUBIC | BX), ((35, 19), S.SWS_BILINEAR), ((1280, 720), S.SWS_POINT | BX)])
def test_rgb444_kernel(sf, geom, flags):
    w, h = geom
    for mode in ("noise", "extreme"):
        name = _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="yuv444p", flags=flags, seed=121, mode=mode)
        assert name == "rgb444", name
    slices = [(y, min(20, h - y)) for y in range(0, h, 20)]
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="yuv444p", flags=flags, seed=122, slices=slices)
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="yuv444p", flags=flags, seed=123, colorspace=(1, 0, 1, 0, 0, 1 << 16, 1 << 16))
    _check(sw=w, sh=h, sf=sf, dw=w, dh=h, df="yuvj444p", flags=flags, seed=124)       # range conversion: another kernel


# ---- 15/16 bpp packed RGB destinations: 2x2 ordered dither in the pair writer and the unscaled LUT converters ----
@pytest.mark.parametrize("df", ["rgb565le", "bgr565le", "rgb555le", "bgr555le"])
@pytest.mark.parametrize("sf,geom,flags", [
    ("yuv420p", (644, 366, 644, 366), S.SWS_BICUBIC | BX), ("yuv420p", (644, 366, 644, 366), S.SWS_BICUBIC),
    ("yuv422p", (322, 182, 322, 182), S.SWS_POINT), ("yuv420p", (322, 182, 400, 300), S.SWS_BICUBIC | BX),
    ("yuv444p", (322, 182, 160, 90), S.SWS_BILINEAR | BX), ("nv12", (322, 182, 400, 300), S.SWS_LANCZOS | BX),
    ("yuv420p10le", (322, 182, 322, 182), S.SWS_BICUBIC | BX), ("p010le", (322, 182, 160, 90), S.SWS_BILINEAR),
    ("bgra", (322, 182, 400, 300), S.SWS_BICUBIC | BX), ("yuv420p", (176, 144, 352, 288), S.SWS_FAST_BILINEAR)])
def test_rgb16bpp_destinations(df, sf, geom, flags):
    sw, sh, dw, dh = geom
    for mode in ("noise", "extreme"):
        _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags, seed=131, mode=mode)
    slices = [(y, min(24, sh - y)) for y in range(0, sh, 24)]
    _check(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags, seed=132, slices=slices)


@pytest.mark.parametrize("sf", RGB_SRC)
@pytest.mark.parametrize("df", ["rgb565le", "bgr565le", "rgb555le", "bgr555le"])
def test_rgb16bpp_from_unscaled_rgb(sf, df):
    """SWS_POINT / SWS_FAST_BILINEAR pick the truncating rgb24to16 family (swscale_unscaled.c:2459-2466); every other
    scaler flag leaves the conversion to the dithering scaler."""
    for g in [(644, 366), (35, 19)]:
        for fl in (S.SWS_POINT, S.SWS_FAST_BILINEAR | BX):
            assert _check(sw=g[0], sh=g[1], sf=sf, dw=g[0], dh=g[1], df=df, flags=fl, seed=161) == "rgb16pack"
        assert _check(sw=g[0], sh=g[1], sf=sf, dw=g[0], dh=g[1], df=df, flags=S.SWS_BICUBIC | BX, seed=162) != "rgb16pack"
    slices = [(y, min(30, 366 - y)) for y in range(0, 366, 30)]
    _check(sw=644, sh=366, sf=sf, dw=644, dh=366, df=df, flags=S.SWS_POINT, seed=163, slices=slices)


def test_rgb16bpp_is_output_only():
    L = S.lib()
    assert not L.sws_isSupportedInput(S.PIX_FMT["rgb565le"]) and L.sws_isSupportedOutput(S.PIX_FMT["rgb565le"])


def test_fuzz_differential_smoke():
    """A fixed-seed slice of tools/fuzz_parity.py (random formats, sizes, scalers, flags, ranges, strides) against the
    real reference, in a subprocess: the full tool found the odd-width LUT, full_1 and same-depth copy cases above."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "--seed", "11", "--cases", "600",
                        "--seconds", "60", "--no-slices"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert "0 mismatches" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("df", ["yuv420p16le", "rgb48le", "yuv444p"])
def test_very_long_vertical_filter(df):
    """12x vertical sinc downscale: ~240 vertical taps; the generic kernel narrows its tile to keep the rows in shared memory."""
    _check(sw=200, sh=1200, sf="yuv420p", dw=200, dh=100, df=df, flags=S.SWS_SINC | BX, seed=141)
    _check(sw=322, sh=1100, sf="yuv420p10le", dw=322, dh=92, df=df, flags=S.SWS_SINC | BX, seed=142, mode="extreme")


@pytest.mark.parametrize("df", ["rgb565le", "bgr565le", "rgb555le", "bgr555le"])
def test_rgb16bpp_odd_widths(df):
    """No full-chroma writer exists for these formats: an odd width keeps the pair writer and its last pair holds one pixel."""
    for sf, g, fl in [("yuv420p", (322, 182, 321, 182), S.SWS_BICUBIC | BX), ("yuv420p", (323, 181, 323, 181), S.SWS_BICUBIC | BX),
                      ("yuv420p", (323, 182, 323, 182), S.SWS_BICUBIC), ("yuv444p", (161, 90, 387, 201), S.SWS_BILINEAR | BX),
                      ("bgra", (163, 91, 257, 131), S.SWS_LANCZOS | BX), ("yuv420p10le", (322, 182, 129, 71), S.SWS_FAST_BILINEAR)]:
        _check(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl, seed=151)


def test_two_live_contexts_with_different_tiles():
    """The dynamic shared-memory limit is a property of the kernel function, not of a context: a second context
    using the same kernel with a smaller tile must not lower the limit of the first one."""
    big = dict(sw=4096, sh=64, sf="yuv420p10le", dw=181, dh=45, df="rgb48le", flags=S.SWS_BICUBIC | BX)
    small = dict(sw=64, sh=48, sf="yuv420p10le", dw=32, dh=24, df="rgb48le", flags=S.SWS_BICUBIC | BX)
    ctxs, srcs = [], []
    for case in (big, small):
        ctxs.append(S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"]))
        srcs.append(T.Frame(case["sf"], case["sw"], case["sh"]).randomize(3))
    for case, c, src in zip((big, small), ctxs, srcs):
        want, _ = T.run_reference(src=src, **case)
        dst = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
        assert c.scale(src.planes, src.strides, dst.planes, dst.strides, 0, case["sh"]) == case["dh"], c.last_error
        assert T.first_diff(dst.valid(), want.valid()) is None, c.kernel_name
    for c in ctxs:
        c.close()
