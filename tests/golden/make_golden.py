#!/usr/bin/env python3
"""Generate tests/golden/golden_md5.json from the REAL reference (oracle/_ref/libswsref.so,
built from /root/reference by oracle/build_ref.py).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are produced by tests/sws_testlib.det_values (splitmix64 of the sample index), so
they can be regenerated bit-for-bit anywhere; only the md5 of every output plane is stored.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import sws_testlib as T  # noqa: E402
from oracle import refapi as R      # noqa: E402

BX = R.BX


def cases():
    out = []
    # the five BASELINE.json configurations at full size ("large": skipped by the CPU suite)
    out += [
        dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=R.SWS_POINT | R.SWS_BITEXACT),
        dict(sw=640, sh=480, sf="yuv420p", dw=640, dh=480, df="rgb24", flags=R.SWS_POINT | BX),
        dict(sw=1920, sh=1080, sf="yuv420p", dw=1920, dh=1080, df="rgb24", flags=R.SWS_BICUBIC | BX, large=1),
        dict(sw=3840, sh=2160, sf="yuv420p10le", dw=3840, dh=2160, df="rgb48le", flags=R.SWS_LANCZOS | BX, large=1),
        dict(sw=7680, sh=4320, sf="nv12", dw=1920, dh=1080, df="yuv420p", flags=R.SWS_BICUBIC | BX, large=1),
        dict(sw=3840, sh=2160, sf="yuv420p", dw=3840, dh=2160, df="rgb24", flags=R.SWS_BICUBIC | BX, large=1),
        dict(sw=3840, sh=2160, sf="yuv420p", dw=3840, dh=2160, df="rgb24", flags=R.SWS_BICUBIC, large=1),
        dict(sw=1920, sh=1080, sf="yuv420p", dw=3840, dh=2160, df="rgb24", flags=R.SWS_BICUBIC | BX, large=1),
    ]
    # small cases across formats / scalers / options
    scalers = [R.SWS_POINT, R.SWS_BILINEAR, R.SWS_BICUBIC, R.SWS_AREA, R.SWS_GAUSS, R.SWS_SINC,
               R.SWS_LANCZOS, R.SWS_SPLINE, R.SWS_X, R.SWS_BICUBLIN]
    for i, sc in enumerate(scalers):
        out.append(dict(sw=352, sh=288, sf="yuv420p", dw=200, dh=100, df="yuv420p", flags=sc | BX))
        out.append(dict(sw=176, sh=144, sf="yuv420p", dw=352, dh=288, df="rgb24", flags=sc | BX))
    for sf in ["yuv420p", "yuv422p", "nv12", "nv21", "yuv420p10le", "yuv422p10le", "yuv420p9le",
               "yuv420p12le", "yuv420p14le", "yuv420p16le", "yuvj420p"]:
        for df in ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr", "rgb48le", "bgr48le", "yuv420p",
                   "yuv422p", "yuv444p", "nv12", "nv21", "yuv420p10le", "yuv444p12le", "yuv420p16le"]:
            out.append(dict(sw=162, sh=122, sf=sf, dw=200, dh=150, df=df, flags=R.SWS_BICUBIC | BX))
    for sf, df in [("yuv420p", "rgb24"), ("yuv420p", "bgra"), ("yuv422p", "rgb24"), ("yuv420p", "rgb48le"),
                   ("yuv420p", "nv12"), ("nv12", "yuv420p"), ("yuv420p10le", "rgb48le")]:
        for fl in [R.SWS_BICUBIC, R.SWS_BICUBIC | BX, R.SWS_POINT]:
            out.append(dict(sw=322, sh=182, sf=sf, dw=322, dh=182, df=df, flags=fl))
    for cs in [(1, 0, 1, 0), (5, 1, 5, 0), (9, 0, 9, 0), (7, 1, 7, 1)]:
        out.append(dict(sw=160, sh=120, sf="yuv420p", dw=160, dh=120, df="rgb24", flags=R.SWS_BICUBIC | BX,
                        colorspace=list(cs) + [0, 1 << 16, 1 << 16]))
    out.append(dict(sw=160, sh=120, sf="yuv420p", dw=160, dh=120, df="rgb24", flags=R.SWS_BICUBIC | BX,
                    colorspace=[5, 0, 5, 0, 3000, 78643, 52428]))
    for rng in [(0, 1), (1, 0)]:
        for df in ["yuv420p", "yuv420p10le", "yuv420p16le"]:
            out.append(dict(sw=160, sh=120, sf="yuv420p", dw=200, dh=150, df=df, flags=R.SWS_BICUBIC | BX,
                            ctx_kwargs=dict(src_range=rng[0], dst_range=rng[1])))
    for pos in [(0, 128, -513, -513), (0, 0, -513, -513), (128, 128, 0, 0)]:
        out.append(dict(sw=160, sh=120, sf="yuv420p", dw=160, dh=120, df="rgb24", flags=R.SWS_BICUBIC | BX,
                        ctx_kwargs=dict(chr_pos=list(pos))))
        out.append(dict(sw=160, sh=120, sf="yuv420p", dw=100, dh=74, df="yuv420p", flags=R.SWS_BICUBIC | BX,
                        ctx_kwargs=dict(chr_pos=list(pos))))
    for (sw, sh, dw, dh) in [(16, 16, 16, 16), (18, 10, 34, 22), (1920, 2, 1920, 2), (4, 1080, 4, 1080),
                             (130, 66, 62, 30), (34, 34, 1280, 720)]:
        out.append(dict(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df="rgb24", flags=R.SWS_BICUBIC | BX))
        out.append(dict(sw=sw, sh=sh, sf="yuv420p", dw=dw, dh=dh, df="yuv420p", flags=R.SWS_BICUBIC | BX))
    # full-chroma RGB (odd width, 4:4:4 sources, explicit flag)
    for sf in ["yuv444p", "yuv420p", "yuv444p10le"]:
        for df in ["rgb24", "bgra", "argb", "rgb48le"]:
            out.append(dict(sw=162, sh=122, sf=sf, dw=201, dh=150, df=df, flags=R.SWS_BICUBIC | BX))
            out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=160, df=df,
                            flags=R.SWS_BILINEAR | BX | R.SWS_FULL_CHR_H_INT))
    # packed 8-bit RGB sources: pair-summed and full-resolution chroma readers,
    # 3- and 4-byte pixels, planar / semi-planar / high-depth / full-range / RGB destinations
    for sf in ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"]:
        for df in ["yuv420p", "yuv444p", "nv12", "yuv420p10le", "yuv444p16le", "yuvj420p", "rgb24"]:
            out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=122, df=df, flags=R.SWS_BICUBIC | BX))
            out.append(dict(sw=162, sh=122, sf=sf, dw=100, dh=75, df=df, flags=R.SWS_BILINEAR | BX))
        out.append(dict(sw=163, sh=121, sf=sf, dw=163, dh=121, df="yuv420p", flags=R.SWS_BICUBIC | BX))
        out.append(dict(sw=162, sh=122, sf=sf, dw=200, dh=150, df="yuv422p", flags=R.SWS_LANCZOS | BX))
        out.append(dict(sw=162, sh=122, sf=sf, dw=120, dh=122, df="yuv420p",
                        flags=R.SWS_BICUBIC | BX | R.SWS_FULL_CHR_H_INP))
        out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=122, df="yuv420p", flags=R.SWS_POINT))
    for cs in [1, 7, 9]:
        out.append(dict(sw=160, sh=120, sf="rgb24", dw=160, dh=120, df="yuv420p", flags=R.SWS_BICUBIC | BX,
                        colorspace=[cs, 0, cs, 0, 0, 1 << 16, 1 << 16]))
    out.append(dict(sw=1920, sh=1080, sf="rgb24", dw=1920, dh=1080, df="yuv420p", flags=R.SWS_BICUBIC | BX, large=1))
    out.append(dict(sw=3840, sh=2160, sf="bgra", dw=1920, dh=1080, df="nv12", flags=R.SWS_BICUBIC | BX, large=1))
    # SWS_FAST_BILINEAR (appended last so that earlier seeds stay put): hyscale_fast / hcscale_fast on 8-bit
    # sources with <= 14-bit destinations, the 2-tap initFilter banks elsewhere, the srcW < 8 fallback
    FB = R.SWS_FAST_BILINEAR
    for sf, df, g in [("yuv420p", "yuv420p", (352, 288, 200, 100)), ("yuv420p", "rgb24", (176, 144, 352, 288)),
                      ("nv12", "bgra", (162, 122, 200, 150)), ("yuv422p", "nv12", (162, 122, 100, 75)),
                      ("yuv420p", "yuv420p10le", (162, 122, 200, 150)), ("yuv420p", "yuv420p16le", (162, 122, 200, 150)),
                      ("yuv420p10le", "yuv420p", (162, 122, 200, 150)), ("yuv420p10le", "rgb48le", (162, 122, 100, 75)),
                      ("rgb24", "yuv420p", (162, 122, 200, 150)), ("bgra", "nv12", (162, 122, 100, 75)),
                      ("yuv420p", "rgb24", (6, 16, 40, 30)), ("yuv420p", "yuv420p", (40, 30, 8, 6)),
                      ("yuv420p", "rgb24", (162, 122, 162, 200)), ("yuv420p", "yuv444p", (162, 122, 323, 122))]:
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=FB | BX))
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=FB))
    # unscaled planar depth conversion (planarCopyWrapper): every depth pair class, both luma rules, no dither
    for sf, df in [("yuv420p10le", "yuv420p"), ("yuv420p", "yuv420p10le"), ("yuv422p10le", "yuv422p12le"),
                   ("yuv444p16le", "yuv444p"), ("yuv420p16le", "yuv420p10le"), ("yuv420p9le", "yuv420p"),
                   ("yuv420p", "yuv420p16le"), ("yuv444p12le", "yuv444p10le"), ("yuv420p14le", "yuv420p12le"),
                   ("yuv420p12le", "yuv420p")]:
        out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=122, df=df, flags=R.SWS_BICUBIC | BX))
        out.append(dict(sw=163, sh=121, sf=sf, dw=163, dh=121, df=df, flags=R.SWS_BICUBIC,
                        ctx_kwargs=dict(src_range=1, dst_range=1)))
        out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=122, df=df, flags=R.SWS_POINT, ctx_kwargs=dict(dither=0)))
    # p010le (the 10-bit surface format of hardware decoders / encoders): reader, writer, unscaled wrappers
    for sf, df, g in [("p010le", "yuv420p", (162, 122, 200, 150)), ("p010le", "rgb24", (162, 122, 162, 122)),
                      ("p010le", "nv12", (162, 122, 100, 75)), ("p010le", "yuv420p10le", (162, 122, 162, 122)),
                      ("p010le", "rgb48le", (162, 122, 200, 150)), ("p010le", "p010le", (162, 122, 100, 75)),
                      ("yuv420p", "p010le", (162, 122, 200, 150)), ("nv12", "p010le", (162, 122, 100, 75)),
                      ("yuv420p10le", "p010le", (162, 122, 200, 150)), ("rgb24", "p010le", (162, 122, 162, 122)),
                      ("yuv420p", "p010le", (162, 122, 162, 122)), ("yuv420p10le", "p010le", (163, 121, 163, 121)),
                      ("yuv420p12le", "p010le", (162, 122, 162, 122)), ("yuv420p16le", "p010le", (162, 122, 162, 122)),
                      ("yuv420p9le", "p010le", (162, 122, 162, 122)), ("yuv422p10le", "p010le", (162, 122, 162, 122))]:
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=R.SWS_BICUBIC | BX))
    out.append(dict(sw=162, sh=122, sf="p010le", dw=200, dh=150, df="yuv420p", flags=R.SWS_BILINEAR | BX,
                    ctx_kwargs=dict(src_range=0, dst_range=1)))
    # 15/16 bpp packed RGB destinations: 2x2 ordered dither in the pair writer
    # (output.c:1714-1747) and in the unscaled LUT converters (yuv2rgb.c:371-398)
    for df in ["rgb565le", "bgr565le", "rgb555le", "bgr555le"]:
        for sf, g, fl in [("yuv420p", (162, 122, 162, 122), R.SWS_BICUBIC | BX), ("yuv420p", (162, 122, 162, 122), R.SWS_BICUBIC),
                          ("yuv422p", (162, 122, 162, 122), R.SWS_POINT), ("yuv420p", (162, 122, 200, 150), R.SWS_BICUBIC | BX),
                          ("yuv444p", (162, 122, 100, 76), R.SWS_BILINEAR | BX), ("nv12", (162, 122, 200, 150), R.SWS_LANCZOS | BX),
                          ("yuv420p10le", (162, 122, 162, 122), R.SWS_BICUBIC | BX), ("p010le", (162, 122, 100, 76), R.SWS_BILINEAR),
                          ("rgb24", (162, 122, 200, 150), R.SWS_BICUBIC | BX), ("yuvj420p", (162, 122, 162, 122), R.SWS_POINT | BX),
                          ("yuv420p", (176, 144, 352, 288), R.SWS_FAST_BILINEAR)]:
            out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl))
    out.append(dict(sw=160, sh=120, sf="yuv420p", dw=160, dh=120, df="rgb565le", flags=R.SWS_BICUBIC | BX,
                    colorspace=[1, 1, 1, 0, 0, 1 << 16, 1 << 16]))
    # found by tools/fuzz_parity.py: the unscaled LUT converter on an odd width (last pixel untouched), the logical
    # shift of yuv2rgba64_full_1_c_template (1 luma tap, 2 chroma taps, 16-bit full-chroma RGB), same-depth copies
    # (p010 -> p010 keeps the low bits of three samples out of four; a range change after init keeps the copy)
    out.append(dict(sw=391, sh=16, sf="yuv420p", dw=391, dh=16, df="abgr", flags=R.SWS_BILINEAR | R.SWS_FULL_CHR_H_INT))
    out.append(dict(sw=163, sh=62, sf="yuv422p", dw=163, dh=62, df="rgb24", flags=R.SWS_POINT))
    out.append(dict(sw=163, sh=61, sf="yuv420p14le", dw=163, dh=61, df="rgb48le", flags=R.SWS_AREA))
    out.append(dict(sw=163, sh=61, sf="yuv420p", dw=163, dh=61, df="bgr48le", flags=R.SWS_BILINEAR | BX))
    out.append(dict(sw=162, sh=61, sf="yuv420p10le", dw=162, dh=61, df="rgb48le", flags=R.SWS_FAST_BILINEAR | R.SWS_FULL_CHR_H_INT))
    out.append(dict(sw=117, sh=57, sf="p010le", dw=117, dh=57, df="p010le", flags=R.SWS_POINT, mode="extreme"))
    out.append(dict(sw=118, sh=57, sf="p010le", dw=118, dh=57, df="p010le", flags=R.SWS_BICUBIC | BX, mode="noise",
                    ctx_kwargs=dict(src_range=1, dst_range=1)))
    out.append(dict(sw=163, sh=61, sf="yuv420p12le", dw=163, dh=61, df="yuv420p12le", flags=R.SWS_BILINEAR,
                    colorspace=[5, 0, 5, 1, 0, 1 << 16, 1 << 16]))
    out.append(dict(sw=163, sh=61, sf="yuv444p16le", dw=163, dh=61, df="yuv444p16le", flags=R.SWS_BICUBIC | BX,
                    colorspace=[5, 1, 5, 0, 0, 1 << 16, 1 << 16]))
    out.append(dict(sw=163, sh=61, sf="yuv422p10le", dw=163, dh=61, df="yuv422p10le", flags=R.SWS_BICUBIC | BX))
    for df in ["rgb565le", "bgr555le"]:       # odd widths of 15/16 bpp destinations: pair writer, last pair of one pixel
        out.append(dict(sw=162, sh=122, sf="yuv420p", dw=161, dh=122, df=df, flags=R.SWS_BICUBIC | BX))
        out.append(dict(sw=163, sh=122, sf="yuv420p", dw=163, dh=122, df=df, flags=R.SWS_BICUBIC))
        out.append(dict(sw=81, sh=61, sf="yuv444p", dw=201, dh=151, df=df, flags=R.SWS_BILINEAR | BX))
    for sf, df in [("rgb24", "rgb565le"), ("bgra", "bgr555le"), ("argb", "bgr565le"), ("bgr24", "rgb555le")]:
        out.append(dict(sw=163, sh=61, sf=sf, dw=163, dh=61, df=df, flags=R.SWS_POINT))            # rgb24to16 & co.
        out.append(dict(sw=163, sh=61, sf=sf, dw=163, dh=61, df=df, flags=R.SWS_BICUBIC | BX))     # dithering scaler
    # one-tap vertical filters ignore their coefficient (yuv2plane1 / yuv2packed1); initFilter leaves 4095 there when a
    # large chroma offset meets a source of two or three rows.  Unscaled copies ignore the chroma siting options.
    out.append(dict(sw=128, sh=2, sf="rgb24", dw=96, dh=24, df="yuv420p14le", flags=R.SWS_LANCZOS | BX,
                    ctx_kwargs=dict(chr_pos=[0, 256, 128, -513])))
    out.append(dict(sw=69, sh=3, sf="bgra", dw=184, dh=36, df="yuv420p12le", flags=R.SWS_SPLINE | BX,
                    ctx_kwargs=dict(src_range=0, dst_range=1, chr_pos=[64, 256, 64, 64])))
    out.append(dict(sw=128, sh=2, sf="yuv444p", dw=96, dh=24, df="rgb24", flags=R.SWS_BICUBIC | BX,
                    ctx_kwargs=dict(chr_pos=[0, 256, 128, -513])))
    out.append(dict(sw=245, sh=229, sf="yuvj420p", dw=245, dh=229, df="yuvj420p", flags=R.SWS_AREA | R.SWS_ACCURATE_RND,
                    ctx_kwargs=dict(chr_pos=[256, 64, 128, 64])))
    out.append(dict(sw=87, sh=198, sf="yuv422p", dw=87, dh=198, df="nv12", flags=R.SWS_BICUBIC | BX,
                    ctx_kwargs=dict(chr_pos=[0, 64, 128, 128])))
    # ---- round 2 (appended: earlier seeds stay put) ----
    # float destinations (yuv2plane1/X_float, yuv2gbrpf32_full_X_c)
    for df in ["grayf32le", "gbrpf32le"]:
        for sf, g, fl in [("yuv420p", (162, 122, 200, 150), R.SWS_BICUBIC | BX), ("yuv444p10le", (162, 122, 100, 76), R.SWS_LANCZOS | BX),
                          ("nv12", (162, 122, 162, 122), R.SWS_BICUBIC), ("yuvj420p", (53, 74, 266, 74), R.SWS_SINC | BX)]:
            out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl))
    # the scaling kernel family: 10-/12-/16-bit sources, range conversion, 9..14-bit and dithered 8-bit writers,
    # packed RGB sources, 17..32-tap banks, 15/16 bpp and full-chroma RGB out of it
    for sf, df, g, fl in [("yuv420p10le", "yuv420p10le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("yuv422p12le", "yuv420p", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("yuv444p16le", "nv12", (322, 182, 200, 100), R.SWS_BILINEAR | BX),
                          ("yuv420p", "yuv444p12le", (322, 182, 400, 300), R.SWS_BICUBIC | BX),
                          ("yuv420p10le", "bgra", (322, 182, 400, 300), R.SWS_BICUBIC | BX),
                          ("bgra", "nv12", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("rgb24", "yuv420p10le", (322, 182, 400, 300), R.SWS_BILINEAR | BX),
                          ("argb", "yuv444p", (323, 181, 200, 100), R.SWS_LANCZOS | BX),
                          ("yuv420p", "yuv420p", (1280, 720, 160, 90), R.SWS_BICUBIC | BX),
                          ("nv12", "rgb24", (1280, 720, 212, 120), R.SWS_BICUBIC | BX),
                          ("yuv420p10le", "yuv420p", (960, 540, 160, 90), R.SWS_LANCZOS | BX),
                          ("yuv420p", "rgb565le", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("rgb24", "bgr555le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("yuv444p", "bgr24", (322, 182, 401, 301), R.SWS_BICUBIC | BX),
                          ("yuv422p10le", "argb", (322, 182, 161, 91), R.SWS_BILINEAR | BX)]:
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl))
    for rng in [(0, 1), (1, 0)]:
        for sf, df in [("yuv420p", "yuv420p"), ("yuv420p10le", "yuv422p10le"), ("nv12", "yuv444p")]:
            out.append(dict(sw=322, sh=182, sf=sf, dw=200, dh=120, df=df, flags=R.SWS_BICUBIC | BX,
                            ctx_kwargs=dict(src_range=rng[0], dst_range=rng[1])))
    # what the numpy restatement does not cover (the CPU suite skips these; the GPU suite holds the CUDA path to them):
    # rgb48 sources, alpha through the scaler, cascades, the unscaled p010 -> nv12 copy
    for sf, df, g, fl in [("rgb48le", "yuv420p", (162, 122, 200, 150), R.SWS_BICUBIC | BX), ("bgr48le", "yuv444p10le", (162, 122, 162, 122), R.SWS_BICUBIC | BX),
                          ("rgb48le", "bgr48le", (163, 61, 163, 61), R.SWS_BICUBIC), ("bgr48le", "rgb24", (162, 122, 100, 76), R.SWS_BILINEAR | BX),
                          ("rgba", "bgra", (162, 122, 200, 150), R.SWS_BICUBIC | BX), ("argb", "abgr", (163, 121, 100, 121), R.SWS_BILINEAR | BX),
                          ("p010le", "nv12", (163, 61, 163, 61), R.SWS_BICUBIC), ("p010le", "nv12", (162, 122, 162, 122), R.SWS_POINT | BX),
                          ("yuv420p", "yuv420p", (2048, 64, 8, 32), R.SWS_BICUBIC | BX)]:
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl))
    for sf, df in [("yuv420p", "yuv420p"), ("yuv420p10le", "yuv420p10le"), ("nv12", "yuv444p16le"), ("yuv420p", "grayf32le")]:
        out.append(dict(sw=162, sh=122, sf=sf, dw=162, dh=122, df=df, flags=R.SWS_BICUBIC | BX,
                        colorspace=[1, 0, 5, 1, 0, 1 << 16, 1 << 16]))
    # the 19-bit lines of the scaling kernel family: 16-bit planar, rgb48 (pair and full-chroma writers, one-tap rows) and
    # gbrpf32le destinations out of 8-bit / 10-bit / 16-bit / nv12 / p010 / packed RGB sources; p010le on both sides
    for sf, df, g, fl in [("yuv420p", "yuv420p16le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("nv12", "yuv444p16le", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("yuv420p10le", "yuv420p16le", (322, 182, 200, 100), R.SWS_BICUBIC | BX),
                          ("yuv444p16le", "yuv444p16le", (322, 182, 400, 182), R.SWS_SPLINE | BX),
                          ("yuv422p12le", "yuv422p16le", (322, 182, 322, 364), R.SWS_BILINEAR | BX),
                          ("bgra", "yuv444p16le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("rgb24", "yuv422p16le", (322, 182, 400, 300), R.SWS_BILINEAR | BX),
                          ("yuv420p", "yuv420p16le", (1280, 720, 212, 120), R.SWS_LANCZOS | BX),
                          ("yuv420p", "rgb48le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("yuv420p10le", "bgr48le", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("yuv444p10le", "rgb48le", (322, 182, 322, 182), R.SWS_BICUBIC | BX),
                          ("yuv444p", "bgr48le", (321, 181, 200, 100), R.SWS_BILINEAR | BX),
                          ("nv12", "rgb48le", (322, 182, 322, 364), R.SWS_BILINEAR | BX),
                          ("yuv422p16le", "rgb48le", (322, 182, 400, 182), R.SWS_BICUBIC | BX),
                          ("yuv420p", "rgb48le", (323, 181, 323, 181), R.SWS_BILINEAR | BX | R.SWS_FULL_CHR_H_INT),
                          ("bgra", "rgb48le", (322, 182, 200, 100), R.SWS_BICUBIC | BX),
                          ("yuv420p", "gbrpf32le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("yuv422p10le", "gbrpf32le", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("p010le", "p010le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("p010le", "nv12", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("p010le", "yuv420p10le", (322, 182, 200, 100), R.SWS_BILINEAR | BX),
                          ("p010le", "bgra", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("p010le", "rgb48le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("p010le", "yuv420p16le", (322, 182, 400, 300), R.SWS_BICUBIC | BX),
                          ("yuv420p", "p010le", (322, 182, 160, 90), R.SWS_BICUBIC | BX),
                          ("nv12", "p010le", (322, 182, 400, 300), R.SWS_LANCZOS | BX),
                          ("yuv420p16le", "p010le", (322, 182, 200, 100), R.SWS_BILINEAR | BX),
                          ("rgb24", "p010le", (322, 182, 160, 90), R.SWS_BICUBIC | BX)]:
        out.append(dict(sw=g[0], sh=g[1], sf=sf, dw=g[2], dh=g[3], df=df, flags=fl))
    for rng in [(0, 1), (1, 0)]:
        for sf, df in [("yuv420p", "yuv420p16le"), ("yuv420p10le", "rgb48le"), ("p010le", "p010le")]:
            out.append(dict(sw=322, sh=182, sf=sf, dw=200, dh=120, df=df, flags=R.SWS_BICUBIC | BX,
                            ctx_kwargs=dict(src_range=rng[0], dst_range=rng[1])))
    for i, c in enumerate(out):
        c.setdefault("seed", 100 + i)
        c.setdefault("mode", "extreme" if i % 7 == 3 else "noise")
    return out


def main():
    res = []
    for c in cases():
        kw = {k: v for k, v in c.items() if k not in ("large", "seed", "mode")}
        if "ctx_kwargs" in kw and "chr_pos" in kw["ctx_kwargs"]:
            kw["ctx_kwargs"] = dict(kw["ctx_kwargs"], chr_pos=tuple(kw["ctx_kwargs"]["chr_pos"]))
        src = T.Frame(c["sf"], c["sw"], c["sh"]).randomize(c["seed"], c["mode"])
        want, _ = T.run_reference(src=src, **kw)
        c["md5"] = T.md5_planes(want.valid())
        res.append(c)
        print(c["sw"], c["sh"], c["sf"], "->", c["dw"], c["dh"], c["df"], hex(c["flags"]), c["md5"][0])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_md5.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "oracle": "oracle/_ref/libswsref.so (real reference C path)",
                   "cases": res}, f, indent=0)
    print(len(res), "cases written")


if __name__ == "__main__":
    main()
