"""ctypes binding for integration/_build/libswsref_hooked.so: the reference's libswscale compiled WITH
ff_sws_init_swscale_cuda() (integration/swscale_cuda.c, integration/build_hooked.py).

TEST INFRASTRUCTURE.  It is a second instance of oracle/refapi.py pointed at the hooked library, so tests
drive the un-hooked and the hooked reference through the same RefContext / RefFrame classes.
"""
import ctypes as C
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_build", "libswsref_hooked.so")

_spec = importlib.util.spec_from_file_location("hooked_refapi", os.path.join(os.path.dirname(_HERE), "oracle", "refapi.py"))
H = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(H)
H.SO_PATH = SO_PATH


def available():
    return os.path.exists(SO_PATH)


def lib():
    L = H.lib()
    if not getattr(L, "_hook_bound", False):
        L.swsref_hook_launches.restype = C.c_long
        L.swsref_hook_launches.argtypes = [C.c_void_p]
        L.swsref_hook_kernel.restype = C.c_char_p
        L.swsref_hook_kernel.argtypes = [C.c_void_p]
        L.swsref_hook_slices_total.restype = C.c_long
        L._hook_bound = True
    return L


def launches(ctx):
    """Kernels launched by the B200 context behind a hooked RefContext; -1 when the C kernels run."""
    return lib().swsref_hook_launches(ctx.h)


def kernel(ctx):
    return lib().swsref_hook_kernel(ctx.h).decode()


def slices_total():
    return lib().swsref_hook_slices_total()
