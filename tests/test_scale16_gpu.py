"""The scaling kernel on 9..16-bit planar sources (sws_scale8_kernel<KS, RGB, MMA = false, S16 = true>:
hScale16To15_c as IDP.2A over sample pairs, swscale.c:99-125), range conversion of the h-scaled lines
(swscale.c:163-216) inside the kernel, and its 9..14-bit planar / dithered 8-bit writers
(output.c:340-357,468-528): bit-exact against the real reference build, with the kernel that ran asserted."""
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu
BX = S.SWS_BITEXACT | S.SWS_ACCURATE_RND

RGB_GEOMS_EARLY = [
    ((640, 360, 320, 180), 4),       # SWS_BICUBIC 2:1
    ((644, 366, 1288, 732), 4),      # 1:2 upscale, ragged tiles
    ((322, 242, 400, 300), 2),       # SWS_BILINEAR
    ((350, 130, 350, 260), 2),       # vertical only
    ((642, 362, 322, 182), 0x200),   # SWS_LANCZOS
]

GEOMS = [
    ((640, 360, 320, 180), S.SWS_BICUBIC),       # 2:1, 8 taps
    ((644, 366, 1288, 732), S.SWS_BICUBIC),      # 1:2 upscale, 4 taps, ragged tiles
    ((1920, 1080, 480, 270), S.SWS_BICUBIC),     # 4:1, 16 taps
    ((1280, 720, 642, 362), S.SWS_LANCZOS),      # 13 taps
    ((322, 242, 400, 300), S.SWS_BILINEAR),
    ((350, 130, 350, 260), S.SWS_BILINEAR),      # vertical only
    ((64, 48, 24, 20), S.SWS_AREA),              # tiles narrower than one warp row
    ((3000, 64, 400, 64), S.SWS_BILINEAR),       # 7.5:1: rows beyond 1 KB, 64-bit TMA elements
]


def _sub(fmt):
    return (1, 1) if "420" in fmt or fmt.startswith("nv") else (1, 0) if "422" in fmt else (0, 0)


def _run(case, seed=71, mode="noise", **kw):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(seed, mode)
    want, _ = T.run_reference(src=src, **case, **kw)
    got, name = T.run_cuda(src=src, **case, **kw)
    assert T.first_diff(got.valid(), want.valid()) is None, (name, case)
    return name


@pytest.mark.parametrize("sf", ["yuv420p10le", "yuv422p10le", "yuv444p12le", "yuv420p9le", "yuv420p14le", "yuv420p16le"])
@pytest.mark.parametrize("df", ["yuv420p10le", "yuv420p", "yuv444p12le", "nv12", "yuv422p9le", "yuv420p14le"])
@pytest.mark.parametrize("geom,flags", GEOMS)
def test_high_depth_sources(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    # a less subsampled source doubles the chroma ratio: its filter can exceed the kernel's 16 taps, and the conversion
    # then stays on the general kernels; with equal subsampling every geometry above fits
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if _sub(sf) == _sub(df) and sw <= 7 * dw:      # (7.5:1 rows of 4:4:4 samples exceed the 2 KB TMA box row)
            assert name == "scale16_dp2a", name


@pytest.mark.parametrize("sf", ["yuv420p10le", "yuv422p12le", "yuv420p16le"])
@pytest.mark.parametrize("df", ["rgb24", "bgra", "abgr"])
@pytest.mark.parametrize("geom,flags", GEOMS[:5])
def test_high_depth_to_packed_rgb(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    name = _run(dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX))
    assert name == "scale16_dp2a", name


@pytest.mark.parametrize("sf,df", [("yuvj420p", "yuv420p"), ("yuv420p", "yuvj420p"), ("yuv420p10le", "yuvj420p"),
                                   ("yuvj422p", "yuv420p10le"), ("nv12", "yuvj444p"), ("yuv444p16le", "yuvj420p"),
                                   ("yuvj420p", "nv21"), ("yuv420p12le", "yuv420p12le")])
@pytest.mark.parametrize("geom,flags", GEOMS[:6])
@pytest.mark.parametrize("ranges", [(0, 1), (1, 0)])
def test_range_conversion_inside_the_scaler(sf, df, geom, flags, ranges):
    """lum/chrRangeToJpeg_c and FromJpeg_c between the two FIR stages, both kernels' horizontal stages."""
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    name = _run(case, ctx_kwargs=dict(src_range=ranges[0], dst_range=ranges[1]))
    if _sub(sf) == _sub(df):
        assert name.startswith("scale8") or name == "scale16_dp2a", name


@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "yuv422p"])
@pytest.mark.parametrize("df", ["yuv420p10le", "yuv444p12le", "yuv422p9le"])
@pytest.mark.parametrize("geom,flags", GEOMS[:5])
def test_8bit_sources_to_high_depth_planar(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    name = _run(dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX))
    if _sub(sf) == _sub(df):
        assert name.startswith("scale8"), name


# ---- 16-bit planar destinations: 19-bit lines (hScale8To19_c / hScale16To19_c, swscale.c:60-97,144-159), the range
# conversion's 16-bit twins (swscale.c:218-255) and yuv2planeX_16_c / yuv2plane1_16_c (output.c:163-187) ----
@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "nv21", "yuv422p", "yuv444p", "yuv420p10le", "yuv422p12le", "yuv444p16le",
                                "yuv420p16le", "yuv420p9le"])
@pytest.mark.parametrize("df", ["yuv420p16le", "yuv422p16le", "yuv444p16le"])
@pytest.mark.parametrize("geom,flags", GEOMS)
def test_16bit_planar_destinations(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if _sub(sf) == _sub(df) and sw <= 7 * dw:
            assert name == ("scale16_i19" if "le" in sf else "scale8_i19"), name


@pytest.mark.parametrize("sf", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("df", ["yuv420p16le", "yuv444p16le", "yuv422p16le"])
@pytest.mark.parametrize("geom,flags", RGB_GEOMS_EARLY)
def test_packed_rgb_sources_to_16bit_planar(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if "420" not in df:          # (a 4:2:0 destination doubles the vertical chroma ratio: its bank can exceed the kernel's taps)
            assert name in ("scale_rgb_i19", "generic_tile"), name
    for rng in ((0, 1), (1, 0)):
        _run(case, ctx_kwargs=dict(src_range=rng[0], dst_range=rng[1]))


@pytest.mark.parametrize("sf,df", [("yuvj420p", "yuv420p16le"), ("yuv420p", "yuv444p16le"), ("yuv420p10le", "yuv420p16le"),
                                   ("yuv444p16le", "yuv444p16le"), ("nv12", "yuv420p16le")])
@pytest.mark.parametrize("geom,flags", GEOMS[:6])
@pytest.mark.parametrize("ranges", [(0, 1), (1, 0)])
def test_16bit_planar_destinations_range_conversion(sf, df, geom, flags, ranges):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode, ctx_kwargs=dict(src_range=ranges[0], dst_range=ranges[1]))
        if _sub(sf) == _sub(df):
            assert name.endswith("_i19"), name


@pytest.mark.parametrize("flags", [S.SWS_SINC, S.SWS_LANCZOS, S.SWS_SPLINE, S.SWS_GAUSS])
def test_16bit_planar_destinations_long_banks(flags):
    """Banks of 17..38 taps (plain vertical banks: no second record), overshooting kernels on extreme input (lines below
    -2^19 / above the 19-bit clip), one-tap rows."""
    for sf, g in [("yuv420p", (1280, 720, 300, 170)), ("yuv420p10le", (642, 1000, 642, 210)), ("yuv444p16le", (900, 100, 120, 100)),
                  ("yuv420p16le", (322, 242, 644, 242))]:
        for mode in ("noise", "extreme"):
            _run(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=sf[:7] + "16le", flags=flags | BX), mode=mode)


# ---- rgb48le / bgr48le through the scaler: yuv2rgba64_{X,2,1} and yuv2rgba64_full_{X,2,1} over 19-bit lines ----
@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "yuv422p", "yuv444p", "yuv420p10le", "yuv444p12le", "yuv422p16le", "p010le",
                                "rgb24", "bgra"])
@pytest.mark.parametrize("df", ["rgb48le", "bgr48le"])
@pytest.mark.parametrize("geom,flags", GEOMS[:7])
def test_rgb48_destinations_through_the_scaler(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if sf not in ("rgb24", "bgra"):    # (wide raw RGB rows stay on the general kernel here)
            assert name.endswith("_i19"), name


@pytest.mark.parametrize("sf", ["yuv420p", "yuv444p", "yuv422p10le", "yuv444p16le", "nv12"])
@pytest.mark.parametrize("df", ["rgb48le", "bgr48le"])
@pytest.mark.parametrize("geom", [(644, 366), (321, 243), (322, 243), (64, 48)])
@pytest.mark.parametrize("flags", [S.SWS_BICUBIC | BX, S.SWS_BILINEAR | BX | S.SWS_FULL_CHR_H_INT, S.SWS_POINT | BX,
                                   S.SWS_LANCZOS | S.SWS_ACCURATE_RND])
def test_rgb48_destinations_same_size(sf, df, geom, flags):
    """Same-size conversions outside fast420_rgb16's shapes (4:4:4 / nv12 sources, odd widths = full chroma, the _1 and _2
    writers of one- and two-tap rows)."""
    w, h = geom
    for mode in ("noise", "extreme"):
        _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags), mode=mode)
    _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags), ctx_kwargs=dict(src_range=1, dst_range=0))
    _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=2 * h, df=df, flags=flags))          # vertical only: one horizontal tap
    _run(dict(sw=w, sh=h, sf=sf, dw=w, dh=h, df=df, flags=flags), ctx_kwargs=dict(chr_pos=(0, 64, 128, 128)))


@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "yuv444p10le", "yuv422p16le", "p010le", "bgra", "yuvj420p"])
@pytest.mark.parametrize("geom,flags", GEOMS[:7])
def test_grayf32_destination_luma_only(sf, geom, flags):
    """grayf32le: yuv2plane1/X_float over the 19-bit luma lines; the chroma stages (and planes) are skipped."""
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df="grayf32le", flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if sf != "bgra":
            assert name.endswith("_i19"), name
    _run(case, ctx_kwargs=dict(src_range=1, dst_range=0))


def test_16bit_planar_destination_slices_and_strides():
    case = dict(sw=644, sh=366, sf="yuv420p", dw=400, dh=222, df="yuv420p16le", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p", 644, 366, pad=16).randomize(5)
    slices = [(y, min(64, 366 - y)) for y in range(0, 366, 64)]
    want, _ = T.run_reference(src=src, slices=slices, dst_pad=6, **case)
    got, name = T.run_cuda(src=src, slices=slices, dst_pad=6, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name
    assert name == "scale8_i19", name


# ---- p010le on either side of the scaler: p010LEToY/UV_c readers (container >> 6, interleaved chroma) and the
# yuv2p010l1/lX_c, yuv2p010cX_c writers (output.c:538-589) ----
@pytest.mark.parametrize("df", ["yuv420p10le", "yuv420p", "nv12", "p010le", "yuv444p12le", "rgb24", "bgra", "yuvj420p",
                                "yuv420p16le", "gbrpf32le"])
@pytest.mark.parametrize("geom,flags", GEOMS)
def test_p010_sources(df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf="p010le", dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if (_sub(df) == (1, 1) or df in ("rgb24", "bgra")) and sw <= 7 * dw:
            assert name == ("scale16_i19" if "16le" in df else "scale16_dp2a"), name


@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "nv21", "yuv420p10le", "yuv420p16le", "yuv422p", "rgb24", "bgra", "yuvj420p"])
@pytest.mark.parametrize("geom,flags", GEOMS[:7])
def test_p010_destinations(sf, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df="p010le", flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        if _sub(sf) == (1, 1):
            assert name.startswith("scale"), name


def test_p010_4k_downscale_full_size():
    case = dict(sw=3840, sh=2160, sf="p010le", dw=1920, dh=1080, df="p010le", flags=S.SWS_BICUBIC | BX)
    assert _run(case) == "scale16_dp2a"
    case = dict(sw=3840, sh=2160, sf="p010le", dw=1920, dh=1080, df="nv12", flags=S.SWS_BICUBIC | BX)
    assert _run(case) == "scale16_dp2a"


def test_high_depth_slices_and_strides():
    case = dict(sw=644, sh=366, sf="yuv420p10le", dw=400, dh=222, df="yuv420p10le", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p10le", 644, 366, pad=16).randomize(5)
    slices = [(y, min(64, 366 - y)) for y in range(0, 366, 64)]
    want, _ = T.run_reference(src=src, slices=slices, dst_pad=6, **case)
    got, name = T.run_cuda(src=src, slices=slices, dst_pad=6, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name


def test_4k_10bit_downscale_full_size():
    case = dict(sw=3840, sh=2160, sf="yuv420p10le", dw=1920, dh=1080, df="yuv420p10le", flags=S.SWS_BICUBIC | BX)
    assert _run(case) == "scale16_dp2a"


# ---- packed 8-bit RGB sources through the scaler: reader stage (input.c:264-345,1068-1180) + IDP.2A FIR stages ----
RGB_GEOMS = [
    ((640, 360, 320, 180), S.SWS_BICUBIC),       # 2:1: chroma from pixel pairs (the *_half readers)
    ((644, 366, 1288, 732), S.SWS_BICUBIC),      # 1:2 upscale: full-resolution chroma readers, ragged tiles
    ((1280, 720, 480, 270), S.SWS_BICUBIC),      # 2.7:1, 12 taps
    ((322, 242, 400, 300), S.SWS_BILINEAR),
    ((350, 130, 350, 260), S.SWS_BILINEAR),      # vertical only
    ((642, 362, 322, 182), S.SWS_LANCZOS),
    ((66, 50, 24, 20), S.SWS_AREA),
]


@pytest.mark.parametrize("sf", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("df", ["yuv420p", "nv12", "yuv444p", "yuv422p10le", "yuvj420p", "nv21"])
@pytest.mark.parametrize("geom,flags", RGB_GEOMS)
def test_packed_rgb_sources(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        # wider raw rows than the 2 KB TMA box row, and vertical chroma filters beyond 16 taps (RGB rows are not
        # subsampled: a 4:2:0 destination doubles the vertical chroma ratio), stay on the general kernels
        vsub = "420" in df or df.startswith("nv")
        if sw <= 2 * dw and not (vsub and flags == S.SWS_LANCZOS and sh > dh):
            assert name == "scale_rgb_dp2a", name


@pytest.mark.parametrize("sf,df", [("rgb24", "bgr24"), ("bgra", "rgb24"), ("rgb24", "bgra"), ("argb", "abgr")])
@pytest.mark.parametrize("geom,flags", RGB_GEOMS[:4])
def test_packed_rgb_to_packed_rgb_scaled(sf, df, geom, flags):
    """RGB -> YUV -> RGB, as the reference does when it has to scale (an even destination width keeps shared chroma)."""
    sw, sh, dw, dh = geom
    _run(dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX))


def test_packed_rgb_source_slices_strides_colorspace():
    case = dict(sw=644, sh=366, sf="bgra", dw=400, dh=222, df="nv12", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("bgra", 644, 366, pad=16).randomize(5)
    slices = [(y, min(64, 366 - y)) for y in range(0, 366, 64)]
    for cs in (None, (1, 0, 1, 0, 0, 1 << 16, 1 << 16), (5, 0, 7, 1, 0, 1 << 16, 1 << 16)):
        want, _ = T.run_reference(src=src, slices=slices, dst_pad=6, colorspace=cs, **case)
        got, name = T.run_cuda(src=src, slices=slices, dst_pad=6, colorspace=cs, **case)
        assert T.first_diff(got.valid(), want.valid()) is None, (name, cs)


def test_4k_bgra_to_1080p_nv12_full_size():
    case = dict(sw=3840, sh=2160, sf="bgra", dw=1920, dh=1080, df="nv12", flags=S.SWS_BICUBIC | BX)
    assert _run(case) == "scale_rgb_dp2a"


# ---- 17..32 horizontal taps (eight tap groups) and 21..40 vertical taps (a second record per row) ----
LONG_GEOMS = [
    ((1920, 1080, 320, 180), S.SWS_BICUBIC),     # 6:1, 24 taps both ways
    ((1280, 720, 160, 90), S.SWS_BICUBIC),       # 8:1, 32 taps
    ((1280, 720, 426, 240), S.SWS_LANCZOS),      # 3:1 lanczos, 19 taps
    ((640, 1080, 640, 200), S.SWS_BICUBIC),      # vertical only, 5.4:1
    ((1920, 360, 300, 360), S.SWS_BILINEAR),     # horizontal only, 6.4:1
    ((1000, 600, 130, 70), S.SWS_GAUSS),
]


@pytest.mark.parametrize("sf,df", [("yuv420p", "yuv420p"), ("nv12", "yuv420p"), ("yuv420p", "rgb24"), ("yuv422p", "yuv420p"),
                                   ("yuvj420p", "yuv420p"), ("yuv420p", "yuv420p10le"),
                                   ("yuv420p10le", "yuv420p10le"), ("yuv422p10le", "yuv420p"), ("yuv444p12le", "nv12"),
                                   ("bgra", "yuv420p"), ("rgb24", "nv12")])
@pytest.mark.parametrize("geom,flags", LONG_GEOMS)
def test_long_filters(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    names = set()
    for mode in ("noise", "extreme"):
        names.add(_run(case, mode=mode))
    # whatever kernel took it, the bytes are the reference's; the family is expected where its limits allow
    # (32 horizontal / 38 vertical taps per bank, 2 KB of staged row)
    if sf in ("yuv420p", "nv12", "yuvj420p") and (sw, sh, dw, dh) != (1280, 720, 160, 90):
        assert names <= {"scale8_mma", "scale8_dp4a"}, names


def test_long_filters_slices():
    case = dict(sw=1280, sh=720, sf="yuv420p", dw=213, dh=120, df="yuv420p", flags=S.SWS_BICUBIC | BX)
    src = T.Frame("yuv420p", 1280, 720).randomize(7)
    slices = [(y, min(180, 720 - y)) for y in range(0, 720, 180)]
    want, _ = T.run_reference(src=src, slices=slices, **case)
    got, name = T.run_cuda(src=src, slices=slices, **case)
    assert T.first_diff(got.valid(), want.valid()) is None, name


# ---- 15/16 bpp packed RGB out of the scaling kernel (2 x 2 ordered dither of yuv2rgb_write, output.c:1714-1747) ----
@pytest.mark.parametrize("df", ["rgb565le", "bgr565le", "rgb555le", "bgr555le"])
@pytest.mark.parametrize("sf", ["yuv420p", "nv12", "yuv422p", "yuv420p10le", "bgra", "rgb24"])
@pytest.mark.parametrize("geom,flags", [((322, 182, 400, 300), S.SWS_BICUBIC), ((640, 360, 320, 180), S.SWS_BILINEAR),
                                        ((322, 182, 322, 182), S.SWS_BICUBIC), ((323, 181, 401, 301), S.SWS_LANCZOS),
                                        ((176, 144, 352, 288), S.SWS_BICUBIC)])
def test_rgb16bpp_out_of_the_scaling_kernel(df, sf, geom, flags):
    sw, sh, dw, dh = geom
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        assert name.startswith("scale"), name


# ---- SWS_FULL_CHR_H_INT packed RGB out of the scaling kernel (odd widths, 4:4:4 sources, explicit flag) ----
@pytest.mark.parametrize("sf", ["yuv444p", "yuv420p", "nv12", "yuv422p10le", "yuv444p12le", "bgra", "rgb24"])
@pytest.mark.parametrize("df", ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"])
@pytest.mark.parametrize("geom,flags", [((322, 242, 323, 243), S.SWS_BICUBIC),                          # odd width forces it
                                        ((322, 242, 401, 301), S.SWS_BILINEAR),                         # 2-tap rows: _2 writer
                                        ((644, 362, 321, 181), S.SWS_LANCZOS),
                                        ((322, 242, 322, 300), S.SWS_BILINEAR | S.SWS_FULL_CHR_H_INT),  # vertical only
                                        ((322, 242, 400, 242), S.SWS_BICUBIC | S.SWS_FULL_CHR_H_INT),   # one vertical tap: _1 writer
                                        ((176, 144, 352, 288), S.SWS_BICUBIC | S.SWS_FULL_CHR_H_INT)])
def test_full_chroma_rgb_out_of_the_scaling_kernel(sf, df, geom, flags):
    sw, sh, dw, dh = geom
    if sf in ("bgra", "rgb24") and len(df) == 4 and sf == "bgra":
        pytest.skip("alpha on both sides travels as a fourth line set: the general kernel")
    case = dict(sw=sw, sh=sh, sf=sf, dw=dw, dh=dh, df=df, flags=flags | BX)
    for mode in ("noise", "extreme"):
        name = _run(case, mode=mode)
        assert name.startswith("scale"), name


def test_full_chroma_vertical_chroma_offset():
    """{4096 - a, a} two-tap chroma rows under a one-tap luma: yuv2rgb_full_1 with uvalpha != 0 (no chroma rounding bias)"""
    case = dict(sw=322, sh=242, sf="yuv444p", dw=400, dh=242, df="rgb24", flags=S.SWS_BICUBIC | BX)
    _run(case, ctx_kwargs=dict(chr_pos=(64, 0, 0, 0)))
    _run(case, ctx_kwargs=dict(chr_pos=(256, 0, 0, 128)))
