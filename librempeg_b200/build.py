"""Build librempeg_b200/libswscale_b200.so (host C + sm_100a CUDA) in-tree.

gcc compiles the C host sources, nvcc compiles the single CUDA translation unit
for sm_100a only (-gencode arch=compute_100a,code=sm_100a -lineinfo) and links
the shared object with the reference's symbol-version node (LIBSWSCALE_10,
libswscale/libswscale.v).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(HERE, "libswscale_b200.so")

C_SOURCES = ["sws_context.c", "sws_filter.c", "sws_colorspace.c", "sws_pixfmt.c", "sws_frame.c", "sws_compat.c",
             "sws_hook.c", "sws_options.c", "sws_numa.c"]
CU_SOURCES = ["sws_cuda.cu"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout + r.stderr


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force=False, verbose=False):
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    headers = [os.path.join(ROOT, "include", h) for h in os.listdir(os.path.join(ROOT, "include"))]
    headers += [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh", ".ver"))]
    objs = []
    log = ""
    for s in C_SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(CSRC, s[:-2] + ".o")
        if force or _newer(obj, [src] + headers):
            log += _run(["gcc", "-std=c11", "-O2", "-fPIC", "-Wall", "-Wextra",
                         "-Wno-unused-parameter", "-D_GNU_SOURCE"] + inc + ["-c", src, "-o", obj])
        objs.append(obj)
    for s in CU_SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s[:-3] + ".o")
        if force or _newer(obj, [src] + headers):
            extra = os.environ.get("SWS_B200_NVCC_DEFS", "").split()
            log += _run(["nvcc"] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                                                "-Xptxas", "-v"] + extra + inc + ["-c", src, "-o", obj])
        objs.append(obj)
    if force or _newer(SO_PATH, objs):
        log += _run(["nvcc"] + NVCC_ARCH + ["-shared", "-o", SO_PATH] + objs +
                    ["-Xlinker", "--version-script=" + os.path.join(CSRC, "libswscale_b200.ver"),
                     "-Xlinker", "-Bsymbolic", "-lm", "-ldl"])
    if verbose:
        print(log)
    return SO_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
