#!/usr/bin/env python3
"""Device-resident throughput of every BASELINE.json configuration (not the headline bench line;
bench.py reports the metric configuration).  For each config: frames resident in HBM, one batched
launch per step through sws_cuda_scale_batch(), CUDA events on the library stream.

    python tools/bench_configs.py [--frames 0] [--steps 20]

--frames 0 (default) sizes every batch to about --gbytes of algorithmic traffic (at least 16 frames,
at most 1024), so that every configuration is measured in steady state on a working set far larger
than the 126 MB L2; --frames N forces N frames per launch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from librempeg_b200 import swscale as S  # noqa: E402
from tests import sws_testlib as T  # noqa: E402

CONFIGS = [
    ("C1 640x480 yuv420p->rgb24 point", 640, 480, "yuv420p", 640, 480, "rgb24", S.SWS_POINT | S.BX),
    ("C2 1080p yuv420p->rgb24 bicubic", 1920, 1080, "yuv420p", 1920, 1080, "rgb24", S.SWS_BICUBIC | S.BX),
    ("C3 4K yuv420p10le->rgb48le lanczos", 3840, 2160, "yuv420p10le", 3840, 2160, "rgb48le", S.SWS_LANCZOS | S.BX),
    ("C3b 4K yuv420p10le->rgb24 bicubic", 3840, 2160, "yuv420p10le", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("C3c 4K yuv420p10le->bgra bicubic", 3840, 2160, "yuv420p10le", 3840, 2160, "bgra", S.SWS_BICUBIC | S.BX),
    ("C3q 4K yuv422p10le->bgra bicubic (ProRes-style source)", 3840, 2160, "yuv422p10le", 3840, 2160, "bgra", S.SWS_BICUBIC | S.BX),
    ("C3p 4K p010le->rgb24 bicubic", 3840, 2160, "p010le", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("C4 8K nv12->1080p yuv420p bicubic", 7680, 4320, "nv12", 1920, 1080, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("C5 4K yuv420p->rgb24 bicubic", 3840, 2160, "yuv420p", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("C5' 4K yuv420p->rgb24 default flags (LUT path)", 3840, 2160, "yuv420p", 3840, 2160, "rgb24", S.SWS_BICUBIC),
    ("C5a 4K yuv420p->rgba bicubic", 3840, 2160, "yuv420p", 3840, 2160, "rgba", S.SWS_BICUBIC | S.BX),
    ("C5n 4K nv12->rgb24 bicubic", 3840, 2160, "nv12", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("C5b 4K nv12->bgra bicubic", 3840, 2160, "nv12", 3840, 2160, "bgra", S.SWS_BICUBIC | S.BX),
    ("F1 4K yuv444p->rgb24 bicubic (full chroma)", 3840, 2160, "yuv444p", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("F2 4K yuv444p->bgra bicubic (full chroma)", 3840, 2160, "yuv444p", 3840, 2160, "bgra", S.SWS_BICUBIC | S.BX),
    ("F3 4K rgb24->yuv444p bicubic", 3840, 2160, "rgb24", 3840, 2160, "yuv444p", S.SWS_BICUBIC | S.BX),
    ("F4 4K bgra->yuv444p bicubic", 3840, 2160, "bgra", 3840, 2160, "yuv444p", S.SWS_BICUBIC | S.BX),
    ("C5j 4K yuvj422p->rgb24 bicubic (MJPEG-style source)", 3840, 2160, "yuvj422p", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("E1 4K rgb24->yuv420p bicubic", 3840, 2160, "rgb24", 3840, 2160, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("E2 4K bgra->1080p nv12 bicubic", 3840, 2160, "bgra", 1920, 1080, "nv12", S.SWS_BICUBIC | S.BX),
    ("E3 4K bgr24->yuv420p default flags (box converter)", 3840, 2160, "bgr24", 3840, 2160, "yuv420p", S.SWS_BICUBIC),
    ("E4 4K rgba->bgra shuffle", 3840, 2160, "rgba", 3840, 2160, "bgra", S.SWS_BICUBIC | S.BX),
    ("U1 4K nv12->yuv420p (de-interleave copy)", 3840, 2160, "nv12", 3840, 2160, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("U2 4K yuv420p->nv12 (interleave copy)", 3840, 2160, "yuv420p", 3840, 2160, "nv12", S.SWS_BICUBIC | S.BX),
    ("D1 4K yuv420p10le->yuv420p (dithered depth copy)", 3840, 2160, "yuv420p10le", 3840, 2160, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("D2 4K yuv420p10le->p010le (planar to p010)", 3840, 2160, "yuv420p10le", 3840, 2160, "p010le", S.SWS_BICUBIC | S.BX),
    ("X1 1080p->4K yuv420p->rgb24 bicubic", 1920, 1080, "yuv420p", 3840, 2160, "rgb24", S.SWS_BICUBIC | S.BX),
    ("X2 4K->1080p yuv420p->yuv420p bicubic", 3840, 2160, "yuv420p", 1920, 1080, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("X3 4K->1080p yuv420p10le->yuv420p10le bicubic", 3840, 2160, "yuv420p10le", 1920, 1080, "yuv420p10le", S.SWS_BICUBIC | S.BX),
    ("X4 4K->1080p yuv420p10le->yuv420p bicubic", 3840, 2160, "yuv420p10le", 1920, 1080, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("X5 1080p yuvj420p->720p yuv420p bicubic (range)", 1920, 1080, "yuvj420p", 1280, 720, "yuv420p", S.SWS_BICUBIC | S.BX),
    ("X8 4K->1080p p010le->p010le bicubic", 3840, 2160, "p010le", 1920, 1080, "p010le", S.SWS_BICUBIC | S.BX),
    ("X9 4K->1080p p010le->nv12 bicubic", 3840, 2160, "p010le", 1920, 1080, "nv12", S.SWS_BICUBIC | S.BX),
    ("X10 4K->1080p yuv420p10le->rgb48le bicubic (19-bit lines)", 3840, 2160, "yuv420p10le", 1920, 1080, "rgb48le", S.SWS_BICUBIC | S.BX),
    ("X11 4K yuv444p10le->rgb48le bicubic (same size, full chroma)", 3840, 2160, "yuv444p10le", 3840, 2160, "rgb48le", S.SWS_BICUBIC | S.BX),
    ("X6 4K->1080p yuv420p10le->yuv420p16le bicubic (19-bit lines)", 3840, 2160, "yuv420p10le", 1920, 1080, "yuv420p16le", S.SWS_BICUBIC | S.BX),
    ("X7 1080p->4K yuv420p->yuv420p16le bicubic (19-bit lines)", 1920, 1080, "yuv420p", 3840, 2160, "yuv420p16le", S.SWS_BICUBIC | S.BX),
]


def measure(cfg, dev, peak, frames=0, gbytes=1.5, steps=20):
    """One configuration, frames resident in HBM, one batched launch per step, CUDA events on the library
    stream.  Returns a dict (kernel, ms per launch, Mpixel/s in and out, algorithmic GB/s, fraction of peak)."""
    (name, sw, sh, sf, dw, dh, df, flags) = cfg
    ctx = S.SwsContext(sw, sh, sf, dw, dh, df, flags)
    sl, dl = T.plane_layout(sf, sw, sh), T.plane_layout(df, dw, dh)
    F = frames
    if F <= 0:
        per_frame = sum(rows * rb for rows, rb in sl) + sum(rows * rb for rows, rb in dl)
        F = max(16, min(1024, int(gbytes * 1e9 / per_frame)))
    src = [torch.randint(0, 256, (F, rows * rb), dtype=torch.uint8, device=dev) for rows, rb in sl]
    if "10le" in sf and sf != "p010le":   # keep 10-bit samples in range
        for t in src:
            v = t.view(torch.int16)
            v &= 0x3FF
    dst = [torch.zeros((F, rows * rb), dtype=torch.uint8, device=dev) for rows, rb in dl]
    sstr, dstr = [rb for _, rb in sl], [rb for _, rb in dl]
    sfs, dfs = [rows * rb for rows, rb in sl], [rows * rb for rows, rb in dl]
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step():
        r = ctx.scale_batch_device(src, sstr, sfs, dst, dstr, dfs, F)
        assert r == dh, ctx.last_error
    for _ in range(3):
        step()
    ctx.sync()
    n0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    launches = ctx.launch_count - n0
    ms = e0.elapsed_time(e1) / steps
    by = (sum(sfs) + sum(dfs)) * F
    gbs = by / (ms * 1e-3) / 1e9
    out = {"name": name, "kernel": ctx.kernel_name, "frames": F, "ms": ms, "launches_per_step": launches / steps,
           "mpix_in": F * sw * sh / ms / 1e3, "mpix_out": F * dw * dh / ms / 1e3,
           "algorithmic_bytes": by, "gbs": gbs, "frac": gbs / peak, "peak": peak}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--gbytes", type=float, default=1.5)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6449.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for cfg in CONFIGS:
        name = cfg[0]
        if args.only and not any(name.split()[0] == k or (len(k) > 3 and k in name) for k in args.only.split(",")):
            continue
        r = measure(cfg, dev, peak, args.frames, args.gbytes, args.steps)
        print("%-48s %-18s %8.3f ms/%d frames  in %8.1f Mpix/s  out %8.1f Mpix/s  %7.1f GB/s = %5.1f%% of %.0f"
              % (name, r["kernel"], r["ms"], r["frames"], r["mpix_in"], r["mpix_out"], r["gbs"], 100 * r["frac"], peak))


if __name__ == "__main__":
    main()
