"""GPU suite, part 2: the CUDA path against COMMITTED goldens (no oracle binary needed on the box),
against the reference's FATE CRCs, against the numpy restatement, and size-independent properties
at BASELINE.json's full sizes."""
import hashlib
import os
import zlib

import numpy as np
import pytest

from tests import sws_testlib as T
from tests.test_oracle_cpu import (FATE_SCALECHROMA, FATE_YUV_RANGE, H, W, _golden_cases, case_id,
                                   case_kwargs, vsynth1)  # noqa: F401  (vsynth1 is a fixture)
from librempeg_b200 import swscale as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", _golden_cases(False) + _golden_cases(True), ids=case_id)
def test_cuda_matches_golden_fixture(case):
    """tests/golden/golden_md5.json was produced by the real reference; includes the five BASELINE
    configurations at full size."""
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(case["seed"], case["mode"])
    got, name = T.run_cuda(src=src, **case_kwargs(case))
    assert T.md5_planes(got.valid()) == case["md5"], name


def test_fate_scalechroma_cuda(vsynth1):  # noqa: F811
    """Reference golden: tests/ref/fate/filter-scalechroma (15 frames), through sws_scale()."""
    c = S.SwsContext(W, H, "yuv444p", W, H, "yuv420p", S.SWS_BICUBIC | S.SWS_BITEXACT,
                     chr_pos=(-513, -513, 0, 256))
    # chr_pos order of the binding: (src_h, src_v, dst_h, dst_v); vf_scale strips 4:4:4 source sitings
    fs = W * H * 3
    crcs = []
    for i in range(15):
        fr = vsynth1[i * fs:(i + 1) * fs]
        planes = [np.ascontiguousarray(fr[k * W * H:(k + 1) * W * H]) for k in range(3)]
        dst = [np.zeros(W * H, np.uint8), np.zeros(W * H // 4, np.uint8), np.zeros(W * H // 4, np.uint8)]
        assert c.scale(planes, [W, W, W], dst, [W, W // 2, W // 2]) == H
        crcs.append(zlib.adler32(b"".join(d.tobytes() for d in dst), 0))
    assert crcs == FATE_SCALECHROMA


def test_fate_yuv_range_cuda(vsynth1):  # noqa: F811
    c = S.SwsContext(W, H, "yuv420p", W, H, "yuv420p", S.SWS_BICUBIC | S.BX, src_range=0, dst_range=1)
    fr = vsynth1[:W * H * 3 // 2]
    planes = [np.ascontiguousarray(fr[:W * H]), np.ascontiguousarray(fr[W * H:W * H * 5 // 4]),
              np.ascontiguousarray(fr[W * H * 5 // 4:])]
    dst = [np.zeros(W * H, np.uint8), np.zeros(W * H // 4, np.uint8), np.zeros(W * H // 4, np.uint8)]
    assert c.scale(planes, [W, W // 2, W // 2], dst, [W, W // 2, W // 2]) == H
    assert zlib.adler32(b"".join(d.tobytes() for d in dst), 0) == FATE_YUV_RANGE


@pytest.mark.parametrize("case", [
    dict(sw=352, sh=288, sf="yuv420p", dw=352, dh=288, df="rgb24", flags=S.SWS_BICUBIC | S.BX),
    dict(sw=352, sh=288, sf="yuv420p", dw=352, dh=288, df="bgr24", flags=S.SWS_LANCZOS),
    dict(sw=352, sh=288, sf="yuv420p10le", dw=352, dh=288, df="rgb48le", flags=S.SWS_LANCZOS | S.BX),
    dict(sw=704, sh=576, sf="nv12", dw=176, dh=144, df="yuv420p", flags=S.SWS_BICUBIC | S.BX),
    dict(sw=352, sh=288, sf="yuv420p", dw=500, dh=300, df="bgra", flags=S.SWS_SPLINE | S.BX),
])
def test_cuda_matches_numpy_restatement(case):
    src = T.Frame(case["sf"], case["sw"], case["sh"]).randomize(77, "smooth")
    got, name = T.run_cuda(src=src, **case)
    want = T.run_oracle(src=src, **case)
    assert T.first_diff(got.valid(), want) is None, name


def _device_batch(ctx, torch, src_planes, frames, w, h):
    """Run sws_cuda_scale_batch() on `frames` copies/variants resident in HBM; returns the output tensor."""
    dev = torch.device("cuda", 0)
    ysz, csz, osz = w * h, (w // 2) * (h // 2), w * h * 3
    sy = torch.empty((frames, ysz), dtype=torch.uint8, device=dev)
    su = torch.empty((frames, csz), dtype=torch.uint8, device=dev)
    sv = torch.empty((frames, csz), dtype=torch.uint8, device=dev)
    for f in range(frames):
        sy[f] = torch.from_numpy(src_planes[f][0].reshape(-1)).to(dev)
        su[f] = torch.from_numpy(src_planes[f][1].reshape(-1)).to(dev)
        sv[f] = torch.from_numpy(src_planes[f][2].reshape(-1)).to(dev)
    dst = torch.zeros((frames, osz), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    r = ctx.scale_batch_device([sy, su, sv], [w, w // 2, w // 2], [ysz, csz, csz], [dst], [w * 3], [osz], frames)
    assert r == h, ctx.last_error
    assert ctx.sync() == 0
    return dst


def test_device_batch_equals_per_frame_host_path_4k():
    """BASELINE configs[4] shape: a batch of distinct 4K frames through the device-resident entry point
    must equal frame-by-frame sws_scale() on host buffers (which the golden tests pin to the reference);
    also checks frame independence (a checksum of per-frame checksums)."""
    torch = pytest.importorskip("torch")
    w, h, frames = 3840, 2160, 6
    ctx = S.SwsContext(w, h, "yuv420p", w, h, "rgb24", S.SWS_BICUBIC | S.BX)
    srcs = [T.Frame("yuv420p", w, h).randomize(900 + f, "noise" if f % 2 else "smooth") for f in range(frames)]
    planes = [[np.ascontiguousarray(p[:, :rb]) for p, (rows, rb) in zip(s.planes, s.layout)] for s in srcs]
    out = _device_batch(ctx, torch, planes, frames, w, h).cpu().numpy()
    assert ctx.kernel_name.startswith("fast420")
    sums = []
    for f in range(frames):
        dst = T.Frame("rgb24", w, h, fill=0)
        assert ctx.scale(srcs[f].planes, srcs[f].strides, dst.planes, dst.strides, 0, h) == h
        want = dst.valid()[0].reshape(-1)
        assert np.array_equal(out[f], want), "frame %d differs between device batch and host path" % f
        sums.append(hashlib.md5(want.tobytes()).hexdigest())
    assert len(set(sums)) == frames


def test_idempotent_and_deterministic():
    """Same context, same input, many calls: identical bytes every time (no state leaks between frames)."""
    w, h = 1920, 1080
    ctx = S.SwsContext(w, h, "yuv420p", w, h, "rgb24", S.SWS_BICUBIC | S.BX)
    src = T.Frame("yuv420p", w, h).randomize(5)
    ref = None
    for _ in range(5):
        dst = T.Frame("rgb24", w, h, fill=0x55)
        assert ctx.scale(src.planes, src.strides, dst.planes, dst.strides, 0, h) == h
        m = T.md5_planes(dst.valid())
        ref = ref or m
        assert m == ref


def test_gray_ramp_property_full_size():
    """Size-independent property at 4K: with U=V=128 and BT.601 limited range every output pixel is grey
    (R==G==B) and equals the reference LUT value of its luma, whatever the scaler's chroma taps are."""
    w, h = 3840, 2160
    src = T.Frame("yuv420p", w, h)
    src.planes[0][:, :w] = (np.arange(w, dtype=np.int64)[None, :] + np.arange(h)[:, None]) % 256
    src.planes[1][:] = 128
    src.planes[2][:] = 128
    dst, _ = T.run_cuda(w, h, "yuv420p", w, h, "rgb24", S.SWS_BICUBIC | S.BX, src)
    rgb = dst.valid()[0].reshape(h, w, 3)
    assert np.array_equal(rgb[..., 0], rgb[..., 1]) and np.array_equal(rgb[..., 1], rgb[..., 2])
    from oracle import sws_oracle as O
    t = O.yuv2rgb_tables(O.YUV2RGB_COEFFS[5], 0, 0, 1 << 16, 1 << 16)
    lut = t["y_table"][np.arange(256) + t["rV"][128 + 512]]
    assert np.array_equal(rgb[..., 0], lut[src.planes[0][:, :w]])


@pytest.mark.parametrize("mode", ["0", "1", "2", "3"])
@pytest.mark.parametrize("geom", [(3840, 2160), (1920, 1080), (640, 480), (1280, 722)])
def test_pinned_host_frames_all_transfer_modes(mode, geom, monkeypatch):
    """sws_scale() on page-locked frames: serial, zero-copy, banded H2D|kernel|D2H and banded with direct
    host stores must all give the bytes of the pageable path (which the goldens pin to the reference)."""
    w, h = geom
    monkeypatch.setenv("SWS_B200_E2E_MODE", mode)
    monkeypatch.setenv("SWS_B200_E2E_BANDS", "5")
    ctx = S.SwsContext(w, h, "yuv420p", w, h, "rgb24", S.SWS_BICUBIC | S.BX)
    src = T.Frame("yuv420p", w, h).randomize(31)
    want = T.Frame("rgb24", w, h, fill=0)
    assert ctx.scale(src.planes, src.strides, want.planes, want.strides, 0, h) == h
    ysz, csz, osz = w * h, (w // 2) * (h // 2), w * h * 3
    bufs = [S.PinnedBuffer(ysz), S.PinnedBuffer(csz), S.PinnedBuffer(csz), S.PinnedBuffer(osz)]
    for b, pl, (rows, rb) in zip(bufs, src.planes, src.layout):
        b.array[:] = pl[:, :rb].reshape(-1)
    bufs[3].array[:] = 0
    for _ in range(2):
        assert ctx.scale([b.ptr for b in bufs[:3]], [w, w // 2, w // 2], [bufs[3].ptr], [w * 3], 0, h) == h
    assert np.array_equal(bufs[3].array.reshape(h, w * 3), want.valid()[0])
    for b in bufs:
        b.close()
