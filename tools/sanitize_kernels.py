#!/usr/bin/env python3
"""One small conversion per kernel family, for compute-sanitizer (racecheck / synccheck / memcheck) runs:

    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py

Each case is also compared with the reference build, and the kernel that ran is printed, so a report names what it covered."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from librempeg_b200 import swscale as S  # noqa: E402
from tests import sws_testlib as T      # noqa: E402

BX = S.SWS_BITEXACT | S.SWS_ACCURATE_RND
CASES = [
    dict(sw=644, sh=366, sf="yuv420p", dw=644, dh=366, df="rgb24", flags=S.SWS_BICUBIC | BX),            # fast420 (TMA + mbarrier ring)
    dict(sw=640, sh=200, sf="nv12", dw=640, dh=200, df="bgra", flags=S.SWS_BICUBIC | BX),                # fast420 narrow tiles, nv12
    dict(sw=644, sh=366, sf="yuvj422p", dw=644, dh=366, df="rgb24", flags=S.SWS_BICUBIC | BX),           # fast420 4:2:2
    dict(sw=644, sh=366, sf="yuv420p10le", dw=644, dh=366, df="rgb48le", flags=S.SWS_LANCZOS | BX),      # fast420_rgb16
    dict(sw=644, sh=366, sf="yuv420p10le", dw=644, dh=366, df="rgb24", flags=S.SWS_BICUBIC | BX),        # fast420_hi8
    dict(sw=644, sh=366, sf="p010le", dw=644, dh=366, df="bgra", flags=S.SWS_BICUBIC | BX),              # fast420_hi8 semi-planar
    dict(sw=1280, sh=720, sf="nv12", dw=320, dh=180, df="yuv420p", flags=S.SWS_BICUBIC | BX),            # scale8_mma
    dict(sw=640, sh=360, sf="yuv420p", dw=400, dh=224, df="nv12", flags=S.SWS_BICUBIC | BX),             # scale8_mma planar
    dict(sw=320, sh=180, sf="yuv420p", dw=640, dh=360, df="rgb24", flags=S.SWS_BICUBIC | BX),            # scale8 packed RGB out
    dict(sw=640, sh=360, sf="yuvj420p", dw=400, dh=224, df="yuv420p", flags=S.SWS_BICUBIC | BX),         # scale8_dp4a + range
    dict(sw=640, sh=360, sf="yuv420p10le", dw=400, dh=224, df="yuv420p10le", flags=S.SWS_BICUBIC | BX),  # scale16_dp2a
    dict(sw=640, sh=360, sf="yuv422p12le", dw=800, dh=450, df="yuv420p", flags=S.SWS_LANCZOS | BX),      # scale16_dp2a, dither
    dict(sw=640, sh=360, sf="yuv420p", dw=400, dh=224, df="yuv420p16le", flags=S.SWS_BICUBIC | BX),      # scale8_i19 (19-bit lines)
    dict(sw=640, sh=360, sf="yuv422p10le", dw=800, dh=450, df="yuv422p16le", flags=S.SWS_LANCZOS | BX),  # scale16_i19
    dict(sw=640, sh=360, sf="yuv420p10le", dw=400, dh=224, df="rgb48le", flags=S.SWS_BICUBIC | BX),      # scale16_i19, rgb48 pair writer
    dict(sw=321, sh=243, sf="yuv444p", dw=321, dh=243, df="bgr48le", flags=S.SWS_BICUBIC | BX),          # scale8_i19, rgb48 full chroma
    dict(sw=644, sh=366, sf="bgra", dw=322, dh=182, df="yuv444p16le", flags=S.SWS_BICUBIC | BX),         # scale_rgb_i19
    dict(sw=640, sh=360, sf="p010le", dw=400, dh=224, df="p010le", flags=S.SWS_BICUBIC | BX),            # scale16_dp2a, p010 both sides
    dict(sw=644, sh=366, sf="rgb24", dw=644, dh=366, df="yuv420p", flags=S.SWS_BICUBIC | BX),            # rgb420
    dict(sw=644, sh=366, sf="bgra", dw=322, dh=182, df="nv12", flags=S.SWS_BICUBIC | BX),                # RGB source, scaled
    dict(sw=644, sh=366, sf="yuv444p", dw=644, dh=366, df="rgb24", flags=S.SWS_BICUBIC | BX),            # full444
    dict(sw=644, sh=366, sf="rgb24", dw=644, dh=366, df="yuv444p", flags=S.SWS_BICUBIC | BX),            # rgb444
    dict(sw=644, sh=366, sf="rgb48le", dw=400, dh=300, df="yuv420p", flags=S.SWS_BICUBIC | BX),          # generic (rgb48 source)
    dict(sw=644, sh=366, sf="nv12", dw=644, dh=366, df="yuv420p", flags=S.SWS_BICUBIC | BX),             # copy8
    dict(sw=644, sh=366, sf="yuv420p10le", dw=644, dh=366, df="yuv420p", flags=S.SWS_BICUBIC | BX),      # depthcopy
    dict(sw=644, sh=366, sf="rgba", dw=644, dh=366, df="bgra", flags=S.SWS_BICUBIC | BX),                # rgb_shuffle
    dict(sw=322, sh=182, sf="rgba", dw=400, dh=300, df="bgra", flags=S.SWS_BICUBIC | BX),                # alpha through the scaler
    dict(sw=322, sh=182, sf="yuv420p", dw=400, dh=300, df="gbrpf32le", flags=S.SWS_BICUBIC | BX),        # float writer
]


def main():
    bad = 0
    for case in CASES:
        got, want, name = T.run_case_both(seed=3, **case)
        diff = T.first_diff(got, want)
        print("%-16s %s %dx%d -> %s %dx%d  %s" % (name, case["sf"], case["sw"], case["sh"], case["df"], case["dw"],
                                                  case["dh"], "ok" if diff is None else diff), flush=True)
        bad += diff is not None
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
