"""librempeg_b200 -- B200-native libswscale hot path (hscale -> vscale -> pixel pack).

The product is the C-ABI shared library `libswscale_b200.so` built from `csrc/`
(host C + hand-written sm_100a CUDA).  `swscale.py` is a ctypes mirror of that
ABI for the test-suite and bench.py; it computes nothing itself.
"""
from . import swscale  # noqa: F401

__all__ = ["swscale"]
