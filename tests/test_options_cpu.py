"""SwsContext.av_class is a real AVClass: the reference's libavutil (its own objects, linked into a test-only
binary together with libswscale_b200.so) sets, reads and resets the options of a context of THIS library by
name -- what libavfilter/vf_scale.c:273,368 does -- and the option table is entry-for-entry the reference's
(libswscale/options.c:34-118), offsets included."""
import ctypes as C
import glob
import os
import subprocess

import pytest

from oracle import refapi as R
from librempeg_b200 import build as native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
PROBE = os.path.join(ROOT, "tests", "native", "_opt_probe")

pytestmark = pytest.mark.skipif(not glob.glob(os.path.join(OBJ, "avu_opt.o")),
                                reason="oracle/_ref/obj (reference libavutil objects) not built")


@pytest.fixture(scope="module")
def probe_output():
    objs = sorted(glob.glob(os.path.join(OBJ, "avu_*.o")))
    cmd = ["gcc", "-O1", "-o", PROBE, os.path.join(ROOT, "tests", "native", "opt_probe.c"),
           "-I", os.path.join(ROOT, "include")] + objs + \
          ["-L", os.path.dirname(native.SO_PATH), "-lswscale_b200", "-Wl,-rpath," + os.path.dirname(native.SO_PATH),
           "-lm", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([PROBE], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.splitlines(), r.stderr


def test_option_table_equals_reference(probe_output):
    out, _ = probe_output
    ours = [l for l in out if l.startswith("opt ")]
    buf = C.create_string_buffer(1 << 16)
    L = R.lib()
    L.swsref_dump_options.restype = C.c_int
    L.swsref_dump_options.argtypes = [C.c_char_p, C.c_int]
    n = L.swsref_dump_options(buf, len(buf))
    ref = buf.raw[:n].decode().splitlines()
    assert len(ref) > 70
    assert ours == ref


def test_libavutil_sets_options_by_name(probe_output):
    out, err = probe_output
    ctx = [l for l in out if l.startswith("ctx ")]
    default = ("ctx flags=4 dither=1 alpha=0 gamma=0 src=16x16/0 dst=16x16/0 range=0,0 chr=-513,-513,-513,-513 "
               "threads=1 intent=1 scaler=0,0 backends=0 param=123456,123456")
    assert ctx[0] == default                      # sws_alloc_context()
    assert ctx[1] == default                      # av_opt_set_defaults() after scribbling over the fields
    assert all(l.endswith(" 0") for l in out if l.startswith("set ")), out
    SWS_BICUBIC, ACC, BITEXACT = 4, 1 << 18, 1 << 19
    assert ctx[2] == ("ctx flags=%d dither=2 alpha=2 gamma=0 src=3840x2160/62 dst=1920x1080/35 range=1,0 "
                      "chr=-513,-513,128,-513 threads=0 intent=1 scaler=7,0 backends=0 param=0.5,123456"
                      % (SWS_BICUBIC | ACC | BITEXACT))
    assert "bad flag 1" in out and "bad range 1" in out and "bad name 1" in out
    assert "get sws_flags 0x000C0004" in out
    assert "get dstw 1920" in out
    assert "plan chr 1920x1080 -> 960x1080 bpc 10,16" in out
    assert "[swscaler @ 0x" in err and "av_log reaches the class" in err
