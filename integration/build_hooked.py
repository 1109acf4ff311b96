#!/usr/bin/env python3
"""Build the reference's libswscale WITH the CUDA hook into integration/_build/libswsref_hooked.so.

TEST INFRASTRUCTURE (the hook source itself, integration/swscale_cuda.c, is product code written for the
reference tree).  What this does:

* creates a scratch tree under a temporary directory: libswscale/ is a directory of symlinks to the files of
  /root/reference/libswscale, except the five files integration/hook_patch.py patches, which are written
  out with the hook lines inserted; libswscale/cuda/swscale_cuda.c links to integration/swscale_cuda.c.
  /root/reference is never written, nothing of it is copied into the repository (the scratch tree is deleted);
* compiles that libswscale with the same synthesised generic-C configuration as oracle/build_ref.py plus
  CONFIG_SWSCALE_CUDA=1, re-using the libavutil objects oracle/build_ref.py already built;
* compiles the B200 host sources with -DSWS_B200_PREFIX=b200_ (include/swscale_b200_prefix.h) so that the
  40 entry points both libraries define do not clash, and re-uses the CUDA object of the product build;
* links everything plus oracle/ref_shim.c (built with -DSWSREF_HOOKED) with nvcc (static CUDA runtime).

The result exports the same swsref_* driver API as oracle/_ref/libswsref.so plus swsref_hook_launches /
swsref_hook_kernel, so tests run the SAME cases through the un-hooked reference and through the reference
calling into the B200 kernels, and compare bytes.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import build_ref            # noqa: E402  (the reference build recipe: source lists, flags, config synthesis)
import hook_patch           # noqa: E402

REF = build_ref.REF
OUT = os.path.join(HERE, "_build")
OBJ = os.path.join(OUT, "obj")
SO = os.path.join(OUT, "libswsref_hooked.so")
CSRC = os.path.join(ROOT, "librempeg_b200", "csrc")
B200_C = ["sws_context.c", "sws_filter.c", "sws_colorspace.c", "sws_pixfmt.c", "sws_frame.c", "sws_compat.c",
          "sws_hook.c", "sws_options.c", "sws_numa.c"]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stderr


def make_scratch(tmp):
    d = os.path.join(tmp, "libswscale")
    os.makedirs(os.path.join(d, "cuda"))
    src = os.path.join(REF, "libswscale")
    patched = set(hook_patch.patched_files())
    for name in os.listdir(src):
        full = os.path.join(src, name)
        if name in patched:
            with open(full, "r") as f:
                text = f.read()
            with open(os.path.join(d, name), "w") as f:
                f.write(hook_patch.apply(name, text))
        else:
            os.symlink(full, os.path.join(d, name))
    os.symlink(os.path.join(HERE, "swscale_cuda.c"), os.path.join(d, "cuda", "swscale_cuda.c"))
    return d


def main():
    if not os.path.isdir(os.path.join(REF, "libswscale")):
        if os.path.exists(SO):
            print("reference tree absent; keeping prebuilt", SO)
            return 0
        print("reference tree absent and no prebuilt", SO, file=sys.stderr)
        return 1
    # the un-hooked build provides config.h and the libavutil objects
    if not os.path.exists(os.path.join(build_ref.OBJ, "avu_opt.o")):
        if build_ref.main() != 0:
            return 1
    # the product build provides the CUDA object (it exports only ff_b200_* / sws_cuda_* names)
    cuda_obj = os.path.join(CSRC, "sws_cuda.o")
    if not os.path.exists(cuda_obj):
        from librempeg_b200 import build as native
        native.build_native()
    os.makedirs(OBJ, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="sws_hooked_")
    try:
        make_scratch(tmp)
        jobs = []
        flags = build_ref.CFLAGS + ["-DCONFIG_SWSCALE_CUDA=1", "-I", build_ref.GEN, "-I", tmp, "-I", REF,
                                    "-I", os.path.join(ROOT, "include")]
        for s in build_ref.SWS_SRCS:
            jobs.append(["gcc"] + flags + ["-DBUILDING_swscale", "-c", os.path.join(tmp, "libswscale", s + ".c"),
                                           "-o", os.path.join(OBJ, "sws_" + s + ".o")])
        jobs.append(["gcc"] + flags + ["-DBUILDING_swscale", "-c", os.path.join(tmp, "libswscale", "cuda", "swscale_cuda.c"),
                                       "-o", os.path.join(OBJ, "sws_cuda_hook.o")])
        jobs.append(["gcc"] + flags + ["-DSWSREF_HOOKED", "-c", os.path.join(ROOT, "oracle", "ref_shim.c"),
                                       "-o", os.path.join(OBJ, "ref_shim.o")])
        for s in B200_C:
            if not os.path.exists(os.path.join(CSRC, s)):
                continue
            jobs.append(["gcc", "-std=c11", "-O2", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter", "-D_GNU_SOURCE",
                         "-DSWS_B200_PREFIX=b200_", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                         "-c", os.path.join(CSRC, s), "-o", os.path.join(OBJ, "b200_" + s[:-2] + ".o")])
        ok = True
        with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            for cmd, (rc, err) in zip(jobs, ex.map(run, jobs)):
                if rc != 0:
                    ok = False
                    print("FAILED", " ".join(cmd[-4:]), "\n", err[-3000:], file=sys.stderr)
        if not ok:
            return 1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    objs = [j[-1] for j in jobs] + [cuda_obj]
    objs += [os.path.join(build_ref.OBJ, "avu_" + s + ".o") for s in build_ref.AVU_SRCS]
    ver = os.path.join(OUT, "swsref.ver")
    with open(ver, "w") as f:
        f.write("SWSREF { global: swsref_*; local: *; };\n")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs + \
          ["-Xlinker", "--version-script=" + ver, "-Xlinker", "-Bsymbolic", "-Xlinker", "--no-undefined",
           "-lm", "-lpthread"]
    rc, err = run(cmd)
    if rc != 0:
        print(err[-6000:], file=sys.stderr)
        return 1
    print("built", SO)
    return 0


if __name__ == "__main__":
    sys.exit(main())
