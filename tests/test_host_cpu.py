"""CPU suite, part 2: the host side of the product (no GPU, no compute calls).

 * the C-ABI library loads and exports every symbol include/*.h declares;
 * the host-built FIR banks and colour constants equal the oracle's (numpy restatement, and the
   real reference when present);
 * the closed-form RGB LUT evaluation used by the kernels equals the reference's byte LUTs;
 * no device => context creation fails loudly (no CPU fallback); the product never touches oracle/.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import sws_testlib as T
from librempeg_b200 import swscale as S
from oracle import sws_oracle as O
from oracle import refapi as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"\b((?:sws|swscale)_[A-Za-z0-9_]+)\s*\(", txt):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(S.SO_PATH, mode=os.RTLD_LOCAL)
    decl = _declared_symbols()
    assert len(decl) >= 30
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_public_struct_layout_matches_reference_abi():
    """Field order of the public SwsContext is ABI (reference swscale.h:222-315)."""
    c = S.lib().sws_alloc_context()
    s = c.contents
    assert s.flags == S.SWS_BICUBIC and s.threads == 1 and s.dither == 1
    assert s.scaler_params[0] == 123456 and s.src_h_chr_pos == -513 and s.dst_v_chr_pos == -513
    assert (s.src_w, s.src_h, s.dst_w, s.dst_h) == (16, 16, 16, 16) and s.intent == 1
    S.lib().sws_freeContext(c)
    assert ctypes.sizeof(S.SwsContextStruct) == 120
    assert S.SwsContextStruct.src_w.offset == 56 and S.SwsContextStruct.backends.offset == 116


def test_version_and_queries():
    L = S.lib()
    L.swscale_version.restype = ctypes.c_uint
    assert L.swscale_version() >> 16 == 10          # LIBSWSCALE_VERSION_MAJOR, version_major.h:27
    assert L.sws_isSupportedInput(S.PIX_FMT["yuv420p"]) and L.sws_isSupportedOutput(S.PIX_FMT["rgb24"])
    assert L.sws_isSupportedInput(S.PIX_FMT["rgb24"])        # packed 8-bit RGB input
    assert L.sws_isSupportedInput(S.PIX_FMT["rgb48le"])      # 16-bit RGB input (rgb48ToY_c & co., input.c:111-196)
    for f in ("rgb565le", "bgr565le", "rgb555le", "bgr555le"):   # 15/16 bpp RGB: output-only
        assert L.sws_isSupportedOutput(S.PIX_FMT[f]) and not L.sws_isSupportedInput(S.PIX_FMT[f])
    assert L.sws_isSupportedInput(S.PIX_FMT["p010le"]) and L.sws_isSupportedOutput(S.PIX_FMT["p010le"])
    if R.available():                                        # the ABI value of the id, libavutil/pixfmt.h
        assert R.pix_fmt("p010le") == S.PIX_FMT["p010le"]
    assert not L.sws_isSupportedOutput(9999)
    co = L.sws_getCoefficients(1)
    assert [co[i] for i in range(4)] == [117489, 138438, 13975, 34925]
    assert [L.sws_getCoefficients(8)[i] for i in range(4)] == [104597, 132201, 25675, 53279]


GEOMS = [
    (640, 480, "yuv420p", 640, 480, "rgb24", S.SWS_POINT | S.BX),
    (1920, 1080, "yuv420p", 1920, 1080, "rgb24", S.SWS_BICUBIC | S.BX),
    (3840, 2160, "yuv420p10le", 3840, 2160, "rgb48le", S.SWS_LANCZOS | S.BX),
    (7680, 4320, "nv12", 1920, 1080, "yuv420p", S.SWS_BICUBIC | S.BX),
    (1920, 1080, "yuv420p", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    (1280, 720, "yuv420p", 854, 480, "yuv420p", S.SWS_LANCZOS | S.BX),
    (1280, 720, "yuv420p", 1000, 700, "yuv420p", S.SWS_SPLINE | S.BX),
    (1280, 720, "yuv420p", 333, 211, "yuv420p", S.SWS_GAUSS | S.BX),
    (1280, 720, "yuv420p", 1921, 1081, "yuv420p", S.SWS_SINC | S.BX),
    (1280, 720, "yuv420p", 640, 360, "yuv420p", S.SWS_AREA | S.BX),
    (640, 360, "yuv420p", 1280, 720, "yuv420p", S.SWS_AREA | S.BX),
    (640, 360, "yuv420p", 1280, 720, "rgb24", S.SWS_BILINEAR),
    (640, 360, "yuv420p", 1280, 720, "yuv420p", S.SWS_X | S.BX),
    (640, 360, "yuv420p", 17, 9, "yuv420p", S.SWS_BICUBIC | S.BX),
    (64, 36, "yuv420p", 1280, 720, "yuv420p", S.SWS_BICUBLIN),
    (640, 360, "yuv422p", 1280, 720, "yuv444p", S.SWS_POINT),
    (640, 360, "yuv420p", 640, 360, "rgb24", S.SWS_BICUBIC),
]


@pytest.mark.parametrize("g", GEOMS, ids=lambda g: "%dx%d_%s_%dx%d_%s_%x" % g)
def test_host_fir_banks_match_oracle(g):
    """ff_b200_build_fir (sws_filter.c) vs initFilter restated in numpy vs the real reference."""
    sw, sh, sf, dw, dh, df, fl = g
    mine = S.SwsContext(sw, sh, sf, dw, dh, df, fl, plan_only=True)
    info = mine.info()
    orc = O.OracleContext(sw, sh, sf, dw, dh, df, fl)
    assert bool(info["unscaled"]) == orc.unscaled_lut
    ref = R.RefContext(sw, sh, sf, dw, dh, df, fl) if R.available() else None
    if ref:
        ri = ref.info()
        for k in ("y_offset", "y_coeff", "v2r", "v2g", "u2g", "u2b", "unscaled", "chrSrcW", "chrSrcH",
                  "chrDstW", "chrDstH", "srcBpc", "dstBpc"):
            assert ri[k] == info[k], k
    if orc.unscaled_lut:
        return
    for which, bank in enumerate((orc.h_lum, orc.h_chr, orc.v_lum, orc.v_chr)):
        co, po = mine.filter(which)
        assert np.array_equal(co, bank[0]) and np.array_equal(po, bank[1]), "bank %d vs numpy oracle" % which
        if ref:
            rc, rp = ref.filter(which)
            assert np.array_equal(co, rc) and np.array_equal(po, rp), "bank %d vs reference" % which


@pytest.mark.parametrize("cs", [(5, 0, 0, 1 << 16, 1 << 16), (1, 0, 0, 1 << 16, 1 << 16), (5, 1, 0, 1 << 16, 1 << 16),
                                (9, 1, 0, 1 << 16, 1 << 16), (7, 0, 0, 1 << 16, 1 << 16),
                                (5, 0, 3000, 78643, 52428), (1, 1, -2000, 60000, 70000)])
def test_rgb_closed_form_equals_reference_luts(cs):
    """The kernels evaluate r = clip_u8((yb + (Y + base_r + ((V8*crv)>>16))*cy) >> 16) instead of reading
    y_table[...] through table_rV (yuv2rgb.c:680-703,901-914).  Check every (Y, U8, V8) the LUT index can see."""
    csp, full, br, co, sa = cs
    mine = S.SwsContext(64, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC | S.BX, plan_only=True)
    assert mine.set_colorspace(csp, full, csp, 0, br, co, sa) == 0
    mine.close()
    # plan_only contexts apply the colourspace at (re)planning time
    mine = S.SwsContext(64, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC | S.BX, plan_only=True,
                        src_range=full)
    L = S.lib()
    L.sws_setColorspaceDetails(mine.p, L.sws_getCoefficients(csp), full, L.sws_getCoefficients(csp), 0, br, co, sa)
    assert L.sws_b200_plan_only(mine.p) == 0
    k = mine.info()
    t = O.yuv2rgb_tables(O.YUV2RGB_COEFFS[csp], full, br, co, sa)
    for key in ("y_offset", "y_coeff", "v2r", "v2g", "u2g", "u2b"):
        assert k[key] == t[key], key
    Y = np.arange(-40, 300, dtype=np.int64)[:, None]
    C8 = np.arange(256, dtype=np.int64)[None, :]

    def closed(idx):
        return np.clip((k["yb"] + idx * k["cy"]) >> 16, 0, 255)

    r = closed(Y + k["base_r"] + ((C8 * k["crv"]) >> 16))
    b = closed(Y + k["base_b"] + ((C8 * k["cbu"]) >> 16))
    assert np.array_equal(r, t["y_table"][Y + t["rV"][C8 + 512]])
    assert np.array_equal(b, t["y_table"][Y + t["bU"][C8 + 512]])
    for u8 in range(0, 256, 5):
        g = closed(Y + k["base_g"] + ((u8 * k["cgu"]) >> 16) + ((C8 * k["cgv"]) >> 16))
        assert np.array_equal(g, t["y_table"][Y + t["gU"][u8 + 512] + t["gV"][C8 + 512]])
    if R.available():
        ref = R.RefContext(64, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC | S.BX, src_range=full)
        ref.set_colorspace(csp, full, csp, 0, br, co, sa)
        y_table, rV, gU, bU, gV = ref.rgb_tables()
        assert np.array_equal(y_table, t["y_table"])
        assert np.array_equal(rV, t["rV"]) and np.array_equal(gU, t["gU"])
        assert np.array_equal(bU, t["bU"]) and np.array_equal(gV, t["gV"])


@pytest.mark.parametrize("cs", [0, 1, 4, 5, 7, 9])
@pytest.mark.parametrize("sf", ["rgb24", "bgra"])
def test_rgb2yuv_matrix_matches_oracle_and_reference(cs, sf):
    """ff_b200_rgb2yuv_table (sws_colorspace.c) vs fill_rgb2yuv_table restated in numpy vs the real
    reference (utils.c:614-706), for every matrix sws_getCoefficients() knows."""
    mine = S.SwsContext(64, 64, sf, 64, 64, "yuv420p", S.SWS_BICUBIC | S.BX, plan_only=True)
    L = S.lib()
    L.sws_setColorspaceDetails(mine.p, L.sws_getCoefficients(cs), 0, L.sws_getCoefficients(cs), 0, 0, 1 << 16, 1 << 16)
    assert L.sws_b200_plan_only(mine.p) == 0
    got = mine.rgb2yuv()
    assert got == O.rgb2yuv_table(O.YUV2RGB_COEFFS.get(cs, O.YUV2RGB_COEFFS[5]))
    info = mine.info()
    assert info["h_shift"] == 13 and info["srcBpc"] == 16 and info["chrSrcW"] == 32
    if R.available() and hasattr(R.lib(), "swsref_rgb2yuv"):
        ref = R.RefContext(64, 64, sf, 64, 64, "yuv420p", S.SWS_BICUBIC | S.BX)
        ref.set_colorspace(cs, 0, cs, 0, 0, 1 << 16, 1 << 16)
        assert ref.rgb2yuv() == got
        ri = ref.info()
        assert ri["srcBpc"] == 16 and ri["chrSrcW"] == 32
    # a YUV source has no such matrix
    yuv = S.SwsContext(64, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC | S.BX, plan_only=True)
    assert yuv.rgb2yuv() is None


def test_rgb_source_chroma_reader_selection():
    """*_half readers only when the source width is even, SWS_FULL_CHR_H_INP is absent and the chroma
    output is at most half the source width (utils.c:1367-1390)."""
    for (sw, dw, df, flags, want) in [(64, 64, "yuv420p", 0, 32), (64, 64, "yuv444p", 0, 64),
                                      (65, 65, "yuv420p", 0, 65), (64, 64, "yuv420p", S.SWS_FULL_CHR_H_INP, 64),
                                      (64, 32, "yuv444p", 0, 32), (64, 200, "yuv420p", 0, 64)]:
        mine = S.SwsContext(sw, 48, "rgb24", dw, 48, df, S.SWS_BICUBIC | S.BX | flags, plan_only=True)
        assert mine.info()["chrSrcW"] == want, (sw, dw, df, flags)
        if R.available():
            assert R.RefContext(sw, 48, "rgb24", dw, 48, df, S.SWS_BICUBIC | S.BX | flags).info()["chrSrcW"] == want


def test_no_device_means_loud_failure_not_cpu_fallback():
    """On a machine without CUDA, sws_init_context must fail; nothing converts on the CPU."""
    if S.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        S.SwsContext(64, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC)
    L = S.lib()
    assert not L.sws_getContext(64, 64, 0, 64, 64, 2, S.SWS_BICUBIC, None, None, None)


def test_init_rejects_what_the_hot_path_does_not_cover():
    for kw in (dict(src_fmt="yuv420p", dst_fmt="rgb24", flags=S.SWS_BICUBIC | S.SWS_BILINEAR),
               dict(src_fmt="yuv420p", dst_fmt="rgb24", flags=S.SWS_BICUBIC | (1 << 16))):    # chroma drop
        with pytest.raises(RuntimeError):
            S.SwsContext(64, 64, kw["src_fmt"], 64, 64, kw["dst_fmt"], kw["flags"], plan_only=True)
    with pytest.raises(RuntimeError):
        S.SwsContext(0, 64, "yuv420p", 64, 64, "rgb24", S.SWS_BICUBIC, plan_only=True)
    # odd width / 4:4:4 sources force full-chroma interpolation (utils.c:1270-1286): covered
    c = S.SwsContext(64, 64, "yuv420p", 65, 64, "rgb24", S.SWS_BICUBIC | S.BX, plan_only=True)
    assert c.fields.flags & S.SWS_FULL_CHR_H_INT
    c = S.SwsContext(64, 64, "yuv444p", 64, 64, "rgb24", S.SWS_BICUBIC | S.BX, plan_only=True)
    assert c.fields.flags & S.SWS_FULL_CHR_H_INT


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under librempeg_b200/ may import, link or mention it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "librempeg_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"\boracle\b|libswsref|refapi", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    out = os.popen("ldd %s" % S.SO_PATH).read()
    assert "swsref" not in out and "avutil" not in out


def test_partition_round_robin():
    from librempeg_b200 import partition as P
    assert P.frames_for_rank(512, 3, 8) == list(range(3, 512, 8))
    assert sum(P.frames_per_rank(513, 8)) == 513 and P.frames_per_rank(512, 8) == [64] * 8
    allf = sorted(sum((P.frames_for_rank(37, r, 4) for r in range(4)), []))
    assert allf == list(range(37))
    with pytest.raises(ValueError):
        P.frames_for_rank(8, 8, 8)


def test_reference_export_list_is_complete():
    """the reference's libswscale exports exactly these 40 symbols (libswscale.v)."""
    want = """sws_alloc_context sws_free_context sws_init_context sws_freeContext sws_getContext
    sws_getCachedContext sws_scale sws_scale_frame sws_frame_setup sws_frame_start sws_frame_end sws_send_slice
    sws_receive_slice sws_receive_slice_alignment sws_is_noop sws_test_format sws_test_hw_format
    sws_test_colorspace sws_test_primaries sws_test_transfer sws_test_frame sws_isSupportedInput
    sws_isSupportedOutput sws_isSupportedEndiannessConversion sws_setColorspaceDetails sws_getColorspaceDetails
    sws_getCoefficients sws_get_class sws_allocVec sws_getGaussianVec sws_scaleVec sws_normalizeVec sws_freeVec
    sws_getDefaultFilter sws_freeFilter sws_convertPalette8ToPacked24 sws_convertPalette8ToPacked32
    swscale_version swscale_configuration swscale_license""".split()
    assert len(want) == 40
    L = S.lib()
    missing = [s for s in want if not hasattr(L, s)]
    assert not missing, missing


def test_colour_property_queries_and_default_filter():
    L = S.lib()
    assert [L.sws_test_colorspace(c, 0) for c in range(12)] == [1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 0, 0]
    assert [L.sws_test_primaries(p, 0) for p in (0, 1, 2, 3, 4, 12, 22, 23, 256, 257)] == [0, 1, 1, 0, 1, 1, 1, 0, 1, 0]
    assert [L.sws_test_transfer(t, 1) for t in (0, 1, 2, 3, 8, 9, 10, 13, 16, 18, 19)] == [0, 1, 1, 0, 1, 0, 0, 1, 1, 1, 0]

    class Vec(ctypes.Structure):
        _fields_ = [("coeff", ctypes.POINTER(ctypes.c_double)), ("length", ctypes.c_int)]

    class Filt(ctypes.Structure):
        _fields_ = [(n, ctypes.POINTER(Vec)) for n in ("lumH", "lumV", "chrH", "chrV")]

    L.sws_getDefaultFilter.restype = ctypes.POINTER(Filt)
    L.sws_getDefaultFilter.argtypes = [ctypes.c_float] * 6 + [ctypes.c_int]
    L.sws_freeFilter.argtypes = [ctypes.POINTER(Filt)]
    f = L.sws_getDefaultFilter(2.0, 0.0, 0.5, 0.0, 1.0, 0.0, 0)
    lum = f.contents.lumH.contents
    co = [lum.coeff[i] for i in range(lum.length)]
    assert lum.length == 7 and abs(sum(co) - 1.0) < 1e-12 and co[3] == max(co) and co[0] < 0   # blur, then unsharp
    chr_h = f.contents.chrH.contents
    assert [chr_h.coeff[i] for i in range(chr_h.length)] == [1.0, 0.0, 0.0]                      # identity shifted left
    assert f.contents.chrV.contents.length == 1
    if R.available():
        RL = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libswsref.so"))
        if hasattr(RL, "swsref_default_filter"):
            out = (ctypes.c_double * 64)()
            n = RL.swsref_default_filter(ctypes.c_float(2.0), ctypes.c_float(0.0), ctypes.c_float(0.5),
                                         ctypes.c_float(0.0), ctypes.c_float(1.0), ctypes.c_float(0.0), out)
            assert n == 7 and [out[i] for i in range(7)] == co
    L.sws_freeFilter(f)
    # the CUDA path refuses to run with one (no silent ignore)
    pal = (ctypes.c_uint8 * 1024)(*([10, 20, 30, 40] * 256))
    src = (ctypes.c_uint8 * 4)(0, 1, 2, 3)
    d24, d32 = (ctypes.c_uint8 * 12)(), (ctypes.c_uint8 * 16)()
    L.sws_convertPalette8ToPacked24(src, d24, 4, pal)
    L.sws_convertPalette8ToPacked32(src, d32, 4, pal)
    assert list(d24) == [10, 20, 30] * 4 and list(d32) == [10, 20, 30, 40] * 4


@pytest.mark.parametrize("g", [(352, 288, "yuv420p", 200, 100, "yuv420p"), (176, 144, "yuv420p", 352, 288, "rgb24"),
                               (640, 360, "nv12", 1000, 360, "yuv420p"), (640, 360, "yuv420p", 640, 200, "rgb24")],
                         ids=lambda g: "%dx%d_%s_%dx%d_%s" % g)
def test_fast_bilinear_banks_restate_hyscale_fast(g):
    """SWS_FAST_BILINEAR: the horizontal banks must reproduce ff_hyscale_fast_c / ff_hcscale_fast_c
    (hscale_fast_bilinear.c:23-55) sample for sample through the ordinary (sum src*coef) >> 7 stage, and the
    vertical banks are the reference's 2-tap initFilter branch (compared with the real reference)."""
    sw, sh, sf, dw, dh, df = g
    fl = S.SWS_FAST_BILINEAR | S.BX
    mine = S.SwsContext(sw, sh, sf, dw, dh, df, fl, plan_only=True)
    orc = O.OracleContext(sw, sh, sf, dw, dh, df, fl)
    assert orc.fast_h
    rng = np.random.default_rng(7)
    for which, (src_w, dst_w, xinc, chroma) in enumerate([(sw, dw, orc.lum_xinc, False),
                                                          (orc.csw, orc.cdw, orc.chr_xinc, True)]):
        co, po = mine.filter(which)
        src = rng.integers(0, 256, (3, src_w)).astype(np.int64)
        want = O.OracleContext._hscale_fast(src, dst_w, xinc, chroma)
        acc = np.zeros((3, dst_w), np.int64)
        for j in range(co.shape[1]):
            acc += src[:, np.minimum(po.astype(np.int64) + j, src_w - 1)] * co[:, j].astype(np.int64)[None, :]
        assert np.array_equal(np.minimum(acc >> 7, 32767), want), "bank %d" % which
        assert (po >= 0).all() and (po + co.shape[1] <= src_w).all()
    if R.available():
        ref = R.RefContext(sw, sh, sf, dw, dh, df, fl)
        for which in (2, 3):
            co, po = mine.filter(which)
            rc, rp = ref.filter(which)
            assert np.array_equal(co, rc) and np.array_equal(po, rp), "vertical bank %d vs reference" % which


def test_special_converter_selection_mirrors_the_reference():
    """Which convert_unscaled hook the reference would install (swscale_unscaled.c:2392-2731) decides the kernel:
    the plan must make the same choice (special id in info()['special'] when exported, else via the error path)."""
    ok = [("nv12", "yuv420p"), ("yuv420p", "nv21"), ("yuv420p", "yuv420p"), ("yuv420p10le", "yuv420p"),
          ("yuv420p", "yuv420p16le"), ("yuv420p", "p010le"), ("yuv420p12le", "p010le"), ("nv12", "p010le"),
          ("rgba", "bgra"), ("bgr24", "yuv420p"), ("p010le", "nv12"), ("rgb48le", "bgr48le"), ("bgr48le", "bgr48le")]
    for sf, df in ok:
        S.SwsContext(128, 64, sf, 128, 64, df, S.SWS_BICUBIC, plan_only=True)
    with pytest.raises(RuntimeError):                      # 15/16 bpp RGB readers are not on the path
        S.SwsContext(128, 64, "rgb565le", 128, 64, "nv12", S.SWS_BICUBIC, plan_only=True)


GAUSS5 = [0.06136, 0.24477, 0.38774, 0.24477, 0.06136]
SHARP3 = [-0.25, 1.5, -0.25]


@pytest.mark.parametrize("src_filter,dst_filter", [
    (dict(lumH=GAUSS5, lumV=GAUSS5), None),
    (dict(lumH=SHARP3, lumV=SHARP3, chrH=GAUSS5, chrV=GAUSS5), None),
    (None, dict(lumH=GAUSS5, chrV=SHARP3)),
    (dict(chrH=[0.5, 0.5]), dict(lumV=[1.0])),
])
@pytest.mark.parametrize("g", [(352, 288, "yuv420p", 352, 288, "rgb24", S.SWS_BICUBIC | S.BX),
                               (352, 288, "yuv420p", 200, 100, "yuv420p", S.SWS_BICUBIC | S.BX),
                               (176, 144, "yuv420p", 352, 288, "yuv420p", S.SWS_BILINEAR)],
                         ids=lambda g: "%dx%d_%s_%dx%d_%s_%x" % g)
def test_swsfilter_vectors_are_convolved_like_initfilter(g, src_filter, dst_filter):
    """SwsFilter pre/post vectors (swscale.h:699-723): the source-side vector of every bank is convolved into
    the rows with the reference's int64 += double * int64 truncation, the destination-side vector only widens
    them (utils.c:385-413); vectors longer than one tap disable the unscaled special converters (:1256,1624)."""
    sw, sh, sf, dw, dh, df, fl = g
    mine = S.SwsContext(sw, sh, sf, dw, dh, df, fl, plan_only=True, src_filter=src_filter, dst_filter=dst_filter)
    orc = O.OracleContext(sw, sh, sf, dw, dh, df, fl, src_filter=src_filter, dst_filter=dst_filter)
    ref = R.RefContext(sw, sh, sf, dw, dh, df, fl, src_filter=src_filter, dst_filter=dst_filter) if R.available() else None
    assert not orc.unscaled_lut and not orc.special
    for which, bank in enumerate((orc.h_lum, orc.h_chr, orc.v_lum, orc.v_chr)):
        co, po = mine.filter(which)
        assert np.array_equal(co, bank[0]) and np.array_equal(po, bank[1]), "bank %d vs numpy oracle" % which
        if ref:
            rc, rp = ref.filter(which)
            assert np.array_equal(co, rc) and np.array_equal(po, rp), "bank %d vs reference" % which


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
def test_host_plan_matches_reference_on_random_cases():
    """Plan-only contexts (no device) over the case generator of tools/fuzz_parity.py: chroma geometry, intermediate
    depths, the six colour constants and all four FIR banks must equal the reference's, bit for bit.  (A 60 s run of
    the same loop compared 110 k cases.)  SWS_FAST_BILINEAR's horizontal banks are exact 2-tap restatements of
    ff_hyscale_fast_c and are checked elsewhere; contexts the reference runs through a special converter keep no banks."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity as F
    rng = random.Random(77)
    compared = 0
    for _ in range(4000):
        c = F.make_case(rng)
        kw = dict(c.get("ctx_kwargs") or {})
        args = (c["sw"], c["sh"], c["sf"], c["dw"], c["dh"], c["df"], c["flags"])
        try:
            ref = R.RefContext(*args, param=c.get("param"), **kw)
        except Exception:
            continue
        try:
            mine = S.SwsContext(*args, param=c.get("param"), plan_only=True, **kw)
        except RuntimeError:
            ref.close()
            continue                      # refused at init
        ri, mi = ref.info(), mine.info()
        if not ri["cascaded"]:
            keys = ["chrSrcW", "chrSrcH", "chrDstH", "srcBpc", "dstBpc"]
            if not ri["unscaled"]:
                keys.append("chrDstW")    # the unscaled LUT converters ignore it (odd widths force the full-chroma flag)
            if c["df"].startswith(("rgb", "bgr", "argb", "abgr")):
                keys += ["y_offset", "y_coeff", "v2r", "v2g", "u2g", "u2b"]
            for k in keys:
                assert ri[k] == mi[k], "%s: reference %r, here %r for %r" % (k, ri[k], mi[k], c)
            if not ri["unscaled"]:
                for which in range(4):
                    if which < 2 and (c["flags"] & S.SWS_FAST_BILINEAR):
                        continue
                    a, b = ref.filter(which), mine.filter(which)
                    assert a is not None and b is not None and a[0].shape == b[0].shape and \
                        np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), "bank %d of %r" % (which, c)
            compared += 1
        ref.close()
        mine.close()
    assert compared > 2500
