#!/usr/bin/env python3
"""sws_cuda_scale_batch_host(): one process, page-locked 4K frames, every visible device -- Mpixel/s for
1, 2, 4, ... devices (the library-level multi-GPU driver: frames round-robin, no collective)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from librempeg_b200 import swscale as S  # noqa: E402

W, H = 3840, 2160


def main():
    nd = S.device_count()
    F = 16 * nd
    ysz, csz, osz = W * H, W * H // 4, W * H * 3
    ctx = S.SwsContext(W, H, "yuv420p", W, H, "rgb24", S.SWS_BICUBIC | S.BX)
    bufs = [S.PinnedBuffer(F * n) for n in (ysz, csz, csz, osz)]
    arrs = [b.array.reshape(F, n) for b, n in zip(bufs, (ysz, csz, csz, osz))]
    rng = np.random.default_rng(3)
    for a in arrs[:3]:
        a[:] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    ref = np.empty(osz, np.uint8)
    assert ctx.scale([a[F - 1].ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [ref.ctypes.data], [W * 3], 0, H) == H
    n = 1
    while n <= nd:
        frames = 16 * n
        def run():
            r = ctx.scale_batch_host([a.ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [ysz, csz, csz],
                                     [arrs[3].ctypes.data], [W * 3], [osz], frames, n, 0)
            assert r == H, ctx.last_error
        run()
        best = 0.0
        for _ in range(4):
            t0 = time.perf_counter()
            run()
            best = max(best, frames * W * H / (time.perf_counter() - t0) / 1e6)
        print("batch_host: %d device(s), %3d frames: %9.0f Mpixel/s" % (n, frames, best))
        n *= 2
    arrs[3][F - 1][:] = 0
    r = ctx.scale_batch_host([a.ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [ysz, csz, csz],
                             [arrs[3].ctypes.data], [W * 3], [osz], F, nd, 0)
    assert r == H and np.array_equal(arrs[3][F - 1], ref), "multi-device batch differs from sws_scale()"
    print("parity of the last frame on %d devices: ok" % nd)


if __name__ == "__main__":
    main()
