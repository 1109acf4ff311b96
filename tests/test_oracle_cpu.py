"""CPU suite, part 1: pin the oracle.

 * the reference's own FATE golden CRCs (tests/ref/fate/filter-scalechroma, sws-yuv-range in
   /root/reference) against the numpy restatement and, when built, the real-reference oracle;
 * the committed golden md5 fixtures (tests/golden/golden_md5.json, made from the real
   reference by tests/golden/make_golden.py) against the numpy restatement;
 * numpy restatement vs the real reference, live, when oracle/_ref is present.
"""
import json
import os
import subprocess
import zlib

import numpy as np
import pytest

from tests import sws_testlib as T
from oracle import refapi as R
from oracle import sws_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VIDEOGEN = os.path.join(ROOT, "oracle", "_ref", "videogen")

# /root/reference/tests/ref/fate/filter-scalechroma: framecrc (adler32, initial value 0) of the 15
# output frames of `-s 352x288 -pix_fmt yuv444p -i vsynth1.yuv -pix_fmt yuv420p -sws_flags +bitexact
# -vf scale=out_chroma_loc=bottomleft` (tests/fate/filter-video.mak:533-535)
FATE_SCALECHROMA = [0x77bb80f8, 0x3a21f6e8, 0xcc0907b0, 0xaa5cd87b, 0x410bd74d, 0x7a763b14, 0x3e4020d4,
                    0x46be8b5d, 0x8021f16c, 0x82ca033d, 0xa76ca6ca, 0x49019bb7, 0x3590adf5, 0xf21235dc,
                    0x6b5f93a9]
# /root/reference/tests/ref/fate/sws-yuv-range (tests/fate/libswscale.mak:29-33): yuv420p limited ->
# full range, bt601, flags=+accurate_rnd+bitexact, frame 0
FATE_YUV_RANGE = 0xbc7a0fa2
W, H = 352, 288


@pytest.fixture(scope="module")
def vsynth1(tmp_path_factory):
    """tests/data/vsynth1.yuv of the reference's FATE suite, made by its tests/videogen.c
    (compiled from where it lies into oracle/_ref/ by oracle/build_ref.py)."""
    if not os.path.exists(VIDEOGEN):
        pytest.skip("oracle/_ref/videogen not built (needs /root/reference at build time)")
    out = tmp_path_factory.mktemp("vsynth") / "vsynth1.yuv"
    subprocess.run([VIDEOGEN, str(out)], check=True)
    return np.fromfile(out, np.uint8)


def _fate_scalechroma(convert, data, frames):
    fs = W * H * 3
    crcs = []
    for i in range(frames):
        fr = data[i * fs:(i + 1) * fs]
        planes = [fr[k * W * H:(k + 1) * W * H].reshape(H, W) for k in range(3)]
        out = convert(planes)
        crcs.append(zlib.adler32(b"".join(np.ascontiguousarray(p).tobytes() for p in out), 0))
    return crcs


def test_fate_scalechroma_numpy_oracle(vsynth1):
    # vf_scale maps out_chroma_loc=bottomleft to dst_h_chr_pos=0, dst_v_chr_pos=256
    # (libswscale/format.c:554-590); 4:4:4 input sitings are stripped (graph.c:622-629)
    ctx = O.OracleContext(W, H, "yuv444p", W, H, "yuv420p", O.SWS_BICUBIC | O.SWS_BITEXACT,
                          chr_pos=(-513, -513, 0, 256))
    assert _fate_scalechroma(ctx.scale, vsynth1, 15) == FATE_SCALECHROMA


def test_fate_yuv_range_numpy_oracle(vsynth1):
    ctx = O.OracleContext(W, H, "yuv420p", W, H, "yuv420p", O.SWS_BICUBIC | O.BX, src_range=0, dst_range=1)
    fr = vsynth1[:W * H * 3 // 2]
    planes = [fr[:W * H].reshape(H, W), fr[W * H:W * H * 5 // 4].reshape(H // 2, W // 2),
              fr[W * H * 5 // 4:].reshape(H // 2, W // 2)]
    out = ctx.scale(planes)
    assert zlib.adler32(b"".join(np.ascontiguousarray(p).tobytes() for p in out), 0) == FATE_YUV_RANGE


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
def test_fate_goldens_real_reference_build(vsynth1):
    """The hand-rolled build of the reference (oracle/build_ref.py) reproduces the reference's own goldens."""
    c = R.RefContext(W, H, "yuv444p", W, H, "yuv420p", R.SWS_BICUBIC | R.SWS_BITEXACT, chr_pos=(-513, -513, 0, 256))

    def conv(planes):
        dst = [np.zeros((H, W), np.uint8), np.zeros((H // 2, W // 2), np.uint8), np.zeros((H // 2, W // 2), np.uint8)]
        c.scale([np.ascontiguousarray(p) for p in planes], [W, W, W], dst, [W, W // 2, W // 2])
        return dst
    assert _fate_scalechroma(conv, vsynth1, 15) == FATE_SCALECHROMA


def _golden_cases(large):
    with open(os.path.join(HERE, "golden", "golden_md5.json")) as f:
        cases = json.load(f)["cases"]
    return [c for c in cases if bool(c.get("large")) == large]


def case_kwargs(c):
    kw = {k: v for k, v in c.items() if k not in ("large", "seed", "mode", "md5")}
    if "ctx_kwargs" in kw and "chr_pos" in kw["ctx_kwargs"]:
        kw["ctx_kwargs"] = dict(kw["ctx_kwargs"], chr_pos=tuple(kw["ctx_kwargs"]["chr_pos"]))
    if "colorspace" in kw:
        kw["colorspace"] = tuple(kw["colorspace"])
    return kw


def case_id(c):
    return "%dx%d_%s_%dx%d_%s_%x_s%d" % (c["sw"], c["sh"], c["sf"], c["dw"], c["dh"], c["df"], c["flags"], c["seed"])


@pytest.mark.parametrize("case", _golden_cases(False), ids=case_id)
def test_numpy_oracle_matches_golden_fixture(case):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(case["seed"], case["mode"])
    try:
        got = T.run_oracle(src=src, **case_kwargs(case))
    except NotImplementedError as e:
        pytest.skip("not restated in numpy (the fixture comes from the real reference; the GPU suite uses it): %s" % e)
    assert T.md5_planes(got) == case["md5"]


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
@pytest.mark.parametrize("idx", range(0, 270, 9))
def test_golden_fixture_still_matches_real_reference(idx):
    """Guards the fixtures themselves: regenerate a sample of them from the live reference."""
    case = _golden_cases(False)[idx]
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(case["seed"], case["mode"])
    want, _ = T.run_reference(src=src, **case_kwargs(case))
    assert T.md5_planes(want.valid()) == case["md5"]


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
def test_survey_md5_recipe_c1_c2():
    """SURVEY.md App. B md5s (reference built with its own configure during the survey) reproduce
    with our hand-rolled reference build: av_lfg(1234) fill, whole-allocation md5."""
    for (w, h, flags, md5) in [(640, 480, R.SWS_POINT | R.BX, "ddc944f9c1efceb168567c4033aa5d97"),
                               (1920, 1080, R.SWS_BICUBIC | R.BX, "4c47750c6a28ca8f4d245ea8d9f59a33")]:
        buf = np.zeros(w * h * 3 // 2, np.uint8)
        R.lfg_fill(buf, 1234, 8)
        planes = [buf[:w * h], buf[w * h:w * h * 5 // 4], buf[w * h * 5 // 4:]]
        dst = np.zeros(w * h * 3, np.uint8)
        c = R.RefContext(w, h, "yuv420p", w, h, "rgb24", flags)
        c.scale(planes, [w, w // 2, w // 2], [dst], [w * 3])
        assert R.md5(dst) == md5


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libswsref.so not built")
def test_oracle_matches_reference_on_random_cases():
    """The case generator of tools/fuzz_parity.py (formats, sizes, scalers and their parameters, flags, ranges,
    colourspace, chroma siting, dither modes) on the CPU: the numpy restatement against the live reference build.
    Cases neither side restates (cascades, alpha through the scaler, ...) are skipped; nothing may differ."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity as F
    rng = random.Random(2024)
    compared = 0
    for _ in range(6000):
        c = F.make_case(rng)
        if max(c["sw"], c["sh"], c["dw"], c["dh"]) > 200:
            continue
        kw = {k: v for k, v in c.items() if k not in ("seed", "mode", "src_pad", "dst_pad")}
        src = T.Frame(c["sf"], c["sw"], c["sh"]).randomize(c["seed"], c["mode"])
        try:
            want = T.run_reference(src=src, **kw)[0].valid()
            got = T.run_oracle(src=src, **kw)
        except NotImplementedError:
            continue
        except RuntimeError as e:
            if "reference" in str(e):
                continue
            raise
        assert T.first_diff(got, want) is None, "%r: %s" % (c, T.first_diff(got, want))
        compared += 1
    assert compared > 800
