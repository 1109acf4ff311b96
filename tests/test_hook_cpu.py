"""The in-tree hook without a device: the reference built WITH
ff_sws_init_swscale_cuda() loads, exports its driver API, and -- because the B200 library refuses to
create a context when no CUDA device exists -- leaves every context on the reference's own C kernels,
bit-identical to the un-hooked build."""
import os

import numpy as np
import pytest

from tests import sws_testlib as T
from oracle import refapi as R
from librempeg_b200 import swscale as S
from integration import hookedapi as HK

pytestmark = pytest.mark.skipif(not HK.available(), reason="integration/_build/libswsref_hooked.so not built")


def test_hooked_library_exports():
    L = HK.lib()
    for name in ("swsref_create", "swsref_scale", "swsref_scale_frame", "swsref_hook_launches",
                 "swsref_hook_kernel", "swsref_hook_slices_total"):
        assert hasattr(L, name), name
    # the B200 library inside is the prefixed build: none of the reference's names may leak out of it
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", HK.SO_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1].split("@")[0] for l in syms.splitlines() if l.strip()]
    assert all(s.startswith("swsref_") or s == "SWSREF" for s in exported), exported


def test_patch_anchors_are_documented():
    from integration import hook_patch
    files = hook_patch.patched_files()
    assert files == ["graph.c", "swscale.c", "swscale_internal.h", "swscale_unscaled.c", "utils.c"]
    added = sum(p[3].count("\n") for p in hook_patch.PATCHES)
    assert added < 80          # the hook is a few dozen lines in the reference tree, not a fork


@pytest.mark.skipif(S.device_count() > 0, reason="a CUDA device is present: covered by tests/test_hook_gpu.py")
@pytest.mark.parametrize("case", [
    dict(sw=64, sh=48, sf="yuv420p", dw=64, dh=48, df="rgb24", flags=S.SWS_POINT | S.BX),
    dict(sw=64, sh=48, sf="yuv420p", dw=96, dh=80, df="rgb24", flags=S.SWS_BICUBIC | S.BX),
    dict(sw=64, sh=48, sf="nv12", dw=32, dh=24, df="yuv420p", flags=S.SWS_BICUBIC | S.BX),
    dict(sw=64, sh=48, sf="yuv420p", dw=64, dh=48, df="rgb24", flags=S.SWS_BICUBIC),
])
def test_falls_through_to_c_kernels_without_device(case):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(5)
    want, _ = T.run_reference(src=src, **case)
    c = HK.H.RefContext(case["sw"], case["sh"], case["sf"], case["dw"], case["dh"], case["df"], case["flags"])
    assert HK.launches(c) == -1          # not hooked: no device
    got = T.Frame(case["df"], case["dw"], case["dh"], fill=0)
    T._drive(c, src, got, case["sh"], None)
    c.close()
    assert T.first_diff(got.valid(), want.valid()) is None
