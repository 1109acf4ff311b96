#!/usr/bin/env python3
"""Randomised differential test: the CUDA path (through the C ABI) against the real reference build
(oracle/_ref/libswsref.so) on random formats, sizes, scalers, flags, ranges, strides and slicings.

    python tools/fuzz_parity.py [--cases 3000] [--seed 1] [--seconds 150]

A case the CUDA path refuses at init (ENOTSUP, documented in DESIGN.md section 7) or the reference refuses
counts as skipped.  Every mismatch is printed as a dict that can be pasted into tests/test_parity_gpu.py.
Exit status 1 if anything differed.
"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from librempeg_b200 import swscale as S  # noqa: E402
from tests import sws_testlib as T       # noqa: E402

SRC = ["yuv420p", "yuv422p", "yuv444p", "yuvj420p", "yuvj422p", "yuvj444p", "nv12", "nv21", "p010le",
       "yuv420p9le", "yuv420p10le", "yuv422p10le", "yuv444p10le", "yuv420p12le", "yuv422p12le", "yuv444p12le",
       "yuv420p14le", "yuv420p16le", "yuv422p16le", "yuv444p16le", "rgb24", "bgr24", "rgba", "bgra", "argb", "abgr",
       "rgb48le", "bgr48le"]
DST = SRC + ["rgb565le", "bgr565le", "rgb555le", "bgr555le", "grayf32le", "gbrpf32le"]
SCALERS = [S.SWS_FAST_BILINEAR, S.SWS_BILINEAR, S.SWS_BICUBIC, S.SWS_X, S.SWS_POINT, S.SWS_AREA, S.SWS_BICUBLIN,
           S.SWS_GAUSS, S.SWS_SINC, S.SWS_LANCZOS, S.SWS_SPLINE]


def rand_dim(rng):
    r = rng.random()
    if r < 0.15:
        return rng.randint(1, 16)
    if r < 0.75:
        return rng.randint(17, 400)
    return rng.choice([128, 256, 320, 352, 512, 640, 644, 720, 1280]) + rng.choice([0, 0, 0, 1, 2, -2, 4])


RGB8 = ["rgb24", "bgr24", "rgba", "bgra", "argb", "abgr"]
FAST = [  # (sources, destinations, same size?) that the specialised kernels serve
    (["yuv420p", "yuvj420p", "nv12", "nv21", "yuv422p", "yuvj422p"], RGB8, True),                 # fast420
    (["yuv420p10le", "yuv420p9le", "yuv420p12le", "yuv420p16le", "p010le", "yuv422p10le"], RGB8, True),   # fast420_hi8
    (["yuv420p10le", "yuv420p", "yuv420p16le", "yuv422p10le"], ["rgb48le", "bgr48le"], True),      # fast420_rgb16
    (RGB8, ["yuv420p", "nv12", "nv21", "yuvj420p"], True),                                         # rgb420
    (RGB8, ["yuv444p", "yuvj444p"], True),                                                         # rgb444
    (["yuv444p", "yuvj444p"], RGB8, True),                                                         # full444
    (["yuv420p", "nv12", "nv21", "yuv422p", "yuv444p", "yuvj420p"],
     ["yuv420p", "nv12", "nv21", "yuv422p", "yuv444p"] + RGB8, False),                              # scale8
    (["yuv420p", "nv12", "yuv422p", "yuv444p", "yuv420p10le", "yuv444p12le", "yuv422p16le", "p010le"] + RGB8,
     ["yuv420p16le", "yuv422p16le", "yuv444p16le", "rgb48le", "bgr48le", "gbrpf32le", "grayf32le", "p010le"], False),  # 19-bit lines, p010
    (["p010le"], ["yuv420p", "nv12", "yuv420p10le", "p010le", "yuv444p"] + RGB8, False),             # p010 sources
]


def make_fast_case(rng):
    srcs, dsts, same = rng.choice(FAST)
    sw = rng.choice([2, 4, 8, 16]) * rng.randint(8, 90)
    sh = 2 * rng.randint(4, 200)
    if same:
        dw, dh = sw, sh
    else:
        dw = max(2, 2 * int(sw * rng.choice([0.25, 0.5, 0.75, 1.5, 2.0, rng.uniform(0.2, 3.0)]) / 2))
        dh = max(2, 2 * int(sh * rng.choice([0.25, 0.5, 0.75, 1.5, 2.0, rng.uniform(0.2, 3.0)]) / 2))
    flags = rng.choice([S.SWS_BICUBIC, S.SWS_BILINEAR, S.SWS_POINT, S.SWS_LANCZOS, S.SWS_AREA, S.SWS_FAST_BILINEAR])
    if rng.random() < 0.7:
        flags |= S.BX
    case = dict(sw=sw, sh=sh, sf=rng.choice(srcs), dw=dw, dh=dh, df=rng.choice(dsts), flags=flags,
                seed=rng.randint(1, 10 ** 6), mode=rng.choice(["noise", "noise", "smooth", "extreme"]))
    if rng.random() < 0.15:
        cs = rng.choice([1, 5, 7, 9])
        case["colorspace"] = (cs, rng.randint(0, 1), cs, rng.randint(0, 1), 0, 1 << 16, 1 << 16)
    if rng.random() < 0.2:
        case["src_pad"], case["dst_pad"] = rng.choice([0, 16, 64, 2]), rng.choice([0, 16, 64, 6])
    return case


def make_case(rng):
    if rng.random() < 0.3:
        return make_fast_case(rng)
    sw, sh = rand_dim(rng), rand_dim(rng)
    r = rng.random()
    if r < 0.35:
        dw, dh = sw, sh
    elif r < 0.45:
        dw, dh = sw, rand_dim(rng)
    elif r < 0.55:
        dw, dh = rand_dim(rng), sh
    else:
        dw, dh = rand_dim(rng), rand_dim(rng)
    # keep the filters within what one context serves (no cascades): ratio <= 12 either way
    dw = min(max(dw, (sw + 11) // 12), sw * 12)
    dh = min(max(dh, (sh + 11) // 12), sh * 12)
    flags = rng.choice(SCALERS)
    r = rng.random()
    if r < 0.55:
        flags |= S.BX
    elif r < 0.65:
        flags |= S.SWS_ACCURATE_RND
    elif r < 0.75:
        flags |= S.SWS_BITEXACT
    if rng.random() < 0.12:
        flags |= S.SWS_FULL_CHR_H_INT
    if rng.random() < 0.08:
        flags |= S.SWS_FULL_CHR_H_INP
    case = dict(sw=sw, sh=sh, sf=rng.choice(SRC), dw=dw, dh=dh, df=rng.choice(DST), flags=flags,
                seed=rng.randint(1, 10 ** 6), mode=rng.choice(["noise", "noise", "smooth", "extreme"]))
    if rng.random() < 0.2:
        case["ctx_kwargs"] = dict(src_range=rng.randint(0, 1), dst_range=rng.randint(0, 1))
    if rng.random() < 0.1:
        cs = rng.choice([1, 5, 7, 9])
        # YUV -> YUV with two different matrices makes the reference cascade through RGB (the reference itself
        # crashes on some tiny sizes, hence the size guard below)
        is_rgb = lambda f: f.startswith(("rgb", "bgr", "argb", "abgr", "gbr"))
        other = rng.choice([cs, 5]) if is_rgb(case["sf"]) or is_rgb(case["df"]) or rng.random() < 0.4 else cs
        if min(case["sw"], case["sh"], case["dw"], case["dh"]) < 16:
            other = cs
        # an unscaled cascade of an odd width without SWS_ACCURATE_RND: the reference's first stage is the yuv2rgb LUT
        # converter, which leaves the last pixel of every row untouched -- in a buffer av_image_alloc() never
        # initialised -- and its second stage reads it: undefined output in the last column (seed 702)
        if (other != cs and (case["sw"] & 1) and (case["sw"], case["sh"]) == (case["dw"], case["dh"])
                and not (case["flags"] & S.SWS_ACCURATE_RND)):
            other = cs
        case["colorspace"] = (cs, rng.randint(0, 1), other, rng.randint(0, 1), 0, 1 << 16, 1 << 16)
    # less common options: scaler parameters, chroma siting, dither mode, picture controls
    if rng.random() < 0.1:
        base = flags & 0x7FF
        if base == S.SWS_BICUBIC:
            case["param"] = (rng.choice([0.0, 1 / 3.0, 1.0, 0.5]), rng.choice([0.5, 1 / 3.0, 0.0, 0.6]))
        elif base == S.SWS_GAUSS:
            case["param"] = (rng.choice([2.0, 3.0, 4.5]), 123456.0)
        elif base == S.SWS_LANCZOS:
            case["param"] = (float(rng.choice([2, 3, 4, 5])), 123456.0)
    if rng.random() < 0.1:
        kw = case.setdefault("ctx_kwargs", {})
        kw["chr_pos"] = tuple(rng.choice([-513, 0, 64, 128, 256]) for _ in range(4))
    if rng.random() < 0.1:
        kw = case.setdefault("ctx_kwargs", {})
        kw["dither"] = rng.choice([0, 1, 2, 3, 4, 5])
    if "colorspace" in case and rng.random() < 0.3:
        cs = list(case["colorspace"])
        cs[4:] = [rng.choice([0, 3000, -5000]), rng.choice([1 << 16, 78643, 52000]), rng.choice([1 << 16, 52428, 90000])]
        case["colorspace"] = tuple(cs)
    if rng.random() < 0.25:
        # 16-bit samples need even strides (the reference reads them through uint16_t pointers)
        case["src_pad"] = rng.choice([0, 1, 3, 16, 64] if T.depth_of(case["sf"]) == 8 and "48" not in case["sf"] else [0, 2, 6, 16, 64])
        case["dst_pad"] = rng.choice([0, 1, 5, 16, 64] if T.depth_of(case["df"]) == 8 and "48" not in case["df"]
                                     and "5le" not in case["df"] and "f32" not in case["df"] else
                                     [0, 4, 16, 64] if "f32" in case["df"] else [0, 2, 6, 16, 64])
    return case


def fmt_class(f):
    if f.endswith("f32le"):
        return "f32"
    if "48" in f:
        return "rgb48"
    if f.endswith(("565le", "555le")):
        return "rgb16bpp"
    if f in ("rgba", "bgra", "argb", "abgr"):
        return "rgb32"
    if f in ("rgb24", "bgr24"):
        return "rgb24"
    if f == "p010le":
        return "p010"
    if f.startswith("nv"):
        return "nv"
    d = T.depth_of(f)
    return "yuv8" if d == 8 else "yuv16" if d == 16 else "yuvN"


def case_class(case):
    scaled = (case["sw"], case["sh"]) != (case["dw"], case["dh"])
    extra = ""
    if case["flags"] & S.SWS_FULL_CHR_H_INT:
        extra += "+fullchr"
    if case.get("colorspace") and case["colorspace"][0] != case["colorspace"][2]:
        extra += "+matrix"
    return "%s -> %s %s%s" % (fmt_class(case["sf"]), fmt_class(case["df"]), "scaled" if scaled else "same size", extra)


def check_batch(case, frames=3):
    """The device-resident batched entry point (sws_cuda_scale_batch, frames strided in HBM) against the host
    path of the same context, which the caller has just compared with the reference.  Returns a diff string or None."""
    import numpy as np
    import torch
    dev = torch.device("cuda", 0)
    kw = dict(case.get("ctx_kwargs") or {})
    c = S.SwsContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"],
                     param=case.get("param"), **kw)
    try:
        if case.get("colorspace") and c.set_colorspace(*case["colorspace"]) < 0:
            return None
        srcs = [T.Frame(case["sf"], case["sw"], case["sh"], pad=case.get("src_pad", 0)).randomize(case["seed"] + 7 * f, case["mode"])
                for f in range(frames)]
        d0 = T.Frame(case["df"], case["dw"], case["dh"], pad=case.get("dst_pad", 0), fill=0)
        st = [torch.from_numpy(np.stack([s.planes[i].reshape(-1) for s in srcs])).to(dev) for i in range(len(srcs[0].planes))]
        dt = [torch.zeros((frames, a.size), dtype=torch.uint8, device=dev) for a in d0.planes]
        torch.cuda.synchronize()
        r = c.scale_batch_device(st, srcs[0].strides, [a.size for a in srcs[0].planes], dt, d0.strides,
                                 [a.size for a in d0.planes], frames)
        if r != case["dh"] or c.sync() != 0:
            return "sws_cuda_scale_batch returned %d: %s" % (r, c.last_error)
        for f in range(frames):
            want = T.Frame(case["df"], case["dw"], case["dh"], pad=case.get("dst_pad", 0), fill=0)
            assert c.scale(srcs[f].planes, srcs[f].strides, want.planes, want.strides, 0, case["sh"]) == case["dh"]
            got = [t[f].cpu().numpy().reshape(a.shape)[:, :rb] for t, a, (rows, rb) in zip(dt, d0.planes, d0.layout)]
            diff = T.first_diff(got, want.valid())
            if diff is not None:
                return "frame %d of the batch: %s" % (f, diff)
        return None
    finally:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--start", type=int, default=0, help="skip the first N cases (resume after a reference abort)")
    ap.add_argument("--no-slices", action="store_true")
    ap.add_argument("--batch-prob", type=float, default=0.2, help="share of cases also run as a 3-frame device batch")
    ap.add_argument("--cursor", default="", help="file that receives the index of the case being run")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    t0 = time.time()
    ran = skipped = bad = batches = 0
    kernels = {}
    reasons = {}
    classes = {}
    for i in range(args.cases):
        if time.time() - t0 > args.seconds:
            break
        case = make_case(rng)
        # (the reference asserts, swscale.c:474, when slices meet a strong vertical downscale: stay below 4x)
        if rng.random() < 0.2 and case["sh"] >= 8 and case["sh"] <= 4 * case["dh"] and not args.no_slices:
            # top-down slices; chroma rows stay aligned (multiples of 2 rows except the last)
            step = 2 * rng.randint(1, max(1, case["sh"] // 4))
            case["slices"] = [(y, min(step, case["sh"] - y)) for y in range(0, case["sh"], step)]
            # bottom-up: the last slice first (swscale.c:1141-1159).  Even heights only: with an odd dst_h the reference
            # writes the last chroma row in front of the plane (heap underflow), with an odd src_h it reads one there
            if rng.random() < 0.25 and not ((case["sh"] | case["dh"]) & 1):
                case["slices"].reverse()
        if i < args.start:
            continue
        if args.cursor:
            with open(args.cursor, "w") as f:
                f.write("%d %r\n" % (i, case))
        try:
            got, want, name = T.run_case_both(**case)
        except Exception as e:  # refused at init by either side
            skipped += 1
            why = str(e)[:90]
            reasons[why] = reasons.get(why, 0) + 1
            if "--verbose" in sys.argv:
                print("skip", case, repr(e)[:100])
            continue
        ran += 1
        kernels[name] = kernels.get(name, 0) + 1
        if name in ("generic_tile", "tile15"):
            k = name + ": " + case_class(case)
            classes[k] = classes.get(k, 0) + 1
        diff = T.first_diff(got, want)
        if diff is not None:
            bad += 1
            print("MISMATCH via %s: %r\n    %s" % (name, case, diff), flush=True)
        elif rng.random() < args.batch_prob and "slices" not in case:
            bdiff = check_batch(case)
            batches += 1
            if bdiff is not None:
                bad += 1
                print("MISMATCH (device batch) via %s: %r\n    %s" % (name, case, bdiff), flush=True)
    print("fuzz: seed %d cases %d..%d: %d compared (+%d as device batches), %d refused, %d mismatches in %.0f s; kernels %s"
          % (args.seed, args.start, i, ran, batches, skipped, bad, time.time() - t0, dict(sorted(kernels.items()))))
    for k, n in sorted(classes.items(), key=lambda kv: -kv[1])[:40]:
        print("  general kernels %5d x %s" % (n, k))
    for why, n in sorted(reasons.items(), key=lambda kv: -kv[1]):
        print("  refused %5d x %s" % (n, why))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
