#!/usr/bin/env python3
"""Host-frame (e2e) throughput of the headline conversion under the knobs of the transfer layer:
pageable vs page-locked memory, copy threads, bands, per-frame sws_scale() vs sws_cuda_scale_batch_host().
    python tools/e2e_sweep.py [--frames 16] [--reps 4]
Each configuration runs in a fresh subprocess (the knobs are read from the environment at context creation)."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W, H = 3840, 2160


def child(args):
    import numpy as np
    from librempeg_b200 import swscale as S
    ctx = S.SwsContext(W, H, "yuv420p", W, H, "rgb24", S.SWS_BICUBIC | S.BX)
    ysz, csz, osz = W * H, W * H // 4, W * H * 3
    F = args.frames
    sizes = (ysz, csz, csz, osz)
    if args.memory == "pinned":
        bufs = [S.PinnedBuffer(F * n) for n in sizes]
        arrs = [b.array.reshape(F, n) for b, n in zip(bufs, sizes)]
    else:
        arrs = [np.empty((F, n), np.uint8) for n in sizes]
    rng = np.random.default_rng(1)
    for a in arrs[:3]:
        a[:] = rng.integers(0, 256, a.shape, dtype=np.uint8)

    def per_frame():
        for f in range(F):
            r = ctx.scale([a[f].ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [arrs[3][f].ctypes.data], [W * 3], 0, H)
            assert r == H, ctx.last_error

    def batch():
        r = ctx.scale_batch_host([a.ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [ysz, csz, csz],
                                 [arrs[3].ctypes.data], [W * 3], [osz], F, args.devices, args.depth)
        assert r == H, ctx.last_error
    fn = batch if args.call == "batch" else per_frame
    fn()
    best = 0.0
    for _ in range(args.reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = max(best, F * W * H / dt / 1e6)
    print(json.dumps({"mpix": best}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--memory", default="pinned")
    ap.add_argument("--call", default="frame")
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--depth", type=int, default=0)
    args = ap.parse_args()
    if args.child:
        return child(args)
    runs = []
    for mem in ("pinned", "pageable"):
        for call in ("frame", "batch"):
            envs = [{}]
            if mem == "pageable":
                envs = [{"SWS_B200_COPY_THREADS": t} for t in ("1", "2", "4", "8", "12", "16")]
            if mem == "pinned" and call == "frame":
                envs = [{"SWS_B200_E2E_MODE": m, "SWS_B200_E2E_BANDS": b} for m in ("0", "2", "3") for b in ("2", "4", "8")]
            if mem == "pinned" and call == "batch":
                envs = [{"SWS_B200_E2E_MODE": m} for m in ("2", "3")]
            for e in envs:
                for depth in ((1, 2, 3) if (call == "batch" and mem == "pageable") else (0,)):
                    env = dict(os.environ)
                    env.update(e)
                    cmd = [sys.executable, __file__, "--child", "--memory", mem, "--call", call, "--frames", str(args.frames),
                           "--reps", str(args.reps), "--depth", str(depth)]
                    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
                    try:
                        v = json.loads(r.stdout.strip().splitlines()[-1])["mpix"]
                    except Exception:
                        v = None
                        print(r.stderr[-500:], file=sys.stderr)
                    print("%-9s %-6s depth=%d %-60s %s Mpix/s" % (mem, call, depth, e, "%.0f" % v if v else "FAILED"))
                    sys.stdout.flush()


if __name__ == "__main__":
    main()
