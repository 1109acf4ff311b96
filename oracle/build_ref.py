#!/usr/bin/env python3
"""Build the REAL reference libswscale (pure-C path) into oracle/_ref/libswsref.so.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is ever linked, imported or
executed by the product library (librempeg_b200/); only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may use it.

What this does
--------------
* compiles the reference's own libswscale/*.c and the libavutil/*.c files it
  needs **from where they lie** under /root/reference with plain gcc (the
  reference's configure/Makefile are NOT run; no reference source is copied
  into this repo);
* the few headers configure would normally generate (config.h,
  config_components.h, libavutil/avconfig.h, libavutil/ffversion.h) are
  synthesised here: every ARCH_*/HAVE_*/CONFIG_* token that occurs in the
  compiled sources is defined to 0 except the short allow-list below, which
  describes "generic little-endian 64-bit Linux, C only, no asm, pthreads";
  this is the same configuration `configure --arch=generic --disable-asm`
  would give, i.e. exactly the bit-exact golden C path;
* links everything plus oracle/ref_shim.c (our own thin wrapper that lets
  Python drive sws_scale()/sws_scale_frame() with plain pointers) into
  oracle/_ref/libswsref.so, exporting only swsref_* symbols.

oracle/_ref/ is git-ignored (no reference-derived binaries in history) but is
shipped to the GPU box by gpurun, where /root/reference does not exist.
"""
import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SWS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
OBJ = os.path.join(OUT, "obj")

SWS_SRCS = """alphablend cms csputils hscale hscale_fast_bilinear filters format
framepool gamma graph input jit lut3d options output rgb2rgb slice swscale
swscale_unscaled utils version yuv2rgb vscale""".split()

AVU_SRCS = """avstring avsscanf bprint buffer channel_layout cpu crc csp dict display
error eval fifo film_grain_params frame hwcontext hwcontext_stub imgutils intmath lfg
log log2_tab mathematics mastering_display_metadata mem opt parseutils pixdesc
random_seed rational refstruct reverse samplefmt side_data slicethread time
utils half2float sha sha512 md5 spherical stereo3d dovi_meta downmix_info
ambient_viewing_environment detection_bbox hdr_dynamic_metadata
hdr_dynamic_vivid_metadata video_enc_params video_hint tdrdi iamf
encryption_info timecode timecode_internal raw_color_params executor threadmessage
container_fifo base64 aes aes_ctr file_open""".split()

# tokens that are 1 in a generic C-only x86-64/Linux build
ON = set("""
HAVE_THREADS HAVE_PTHREADS HAVE_FAST_UNALIGNED HAVE_FAST_64BIT HAVE_FAST_CLZ
HAVE_LOCAL_ALIGNED HAVE_SIMD_ALIGN_16 HAVE_SIMD_ALIGN_32 HAVE_SIMD_ALIGN_64
HAVE_ATOMIC_CAS_PTR HAVE_MACH_ABSOLUTE_TIME_DISABLED
HAVE_UNISTD_H HAVE_SYS_TIME_H HAVE_SYS_PARAM_H HAVE_SYS_RESOURCE_H HAVE_SYS_SELECT_H
HAVE_GETTIMEOFDAY HAVE_CLOCK_GETTIME HAVE_NANOSLEEP HAVE_USLEEP HAVE_SCHED_GETAFFINITY
HAVE_SYSCONF HAVE_POSIX_MEMALIGN HAVE_MEMALIGN HAVE_ISATTY HAVE_GMTIME_R HAVE_LOCALTIME_R
HAVE_MKSTEMP HAVE_MMAP HAVE_MPROTECT HAVE_STRERROR_R HAVE_GETENV HAVE_FCNTL HAVE_LSTAT
HAVE_ACCESS HAVE_GETRUSAGE HAVE_ARC4RANDOM_BUF_DISABLED HAVE_MALLOC_H HAVE_DIRENT_H
HAVE_IO_H_DISABLED HAVE_SYMVER HAVE_SYMVER_GNU_ASM_DISABLED HAVE_PRAGMA_DEPRECATED
HAVE_INLINE_ASM_LABELS_DISABLED HAVE_ATTRIBUTE_MAY_ALIAS HAVE_ATTRIBUTE_PACKED
HAVE_BUILTIN_VECTOR_DISABLED HAVE_STRUCT_POLLFD HAVE_POLL_H
HAVE_CBRT HAVE_CBRTF HAVE_COPYSIGN HAVE_COSF HAVE_ERF HAVE_EXP2 HAVE_EXP2F HAVE_EXPF
HAVE_HYPOT HAVE_ISFINITE HAVE_ISINF HAVE_ISNAN HAVE_LDEXPF HAVE_LLRINT HAVE_LLRINTF
HAVE_LOG2 HAVE_LOG2F HAVE_LOG10F HAVE_LRINT HAVE_LRINTF HAVE_POWF HAVE_RINT HAVE_ROUND
HAVE_ROUNDF HAVE_SINF HAVE_TRUNC HAVE_TRUNCF HAVE_ATANF HAVE_ATAN2F HAVE_FMINF HAVE_FMAXF
HAVE_FMIN HAVE_FMAX HAVE_LIBC_MSVCRT_DISABLED
HAVE_GETAUXVAL HAVE_ELF_AUX_INFO_DISABLED HAVE_ASM_TYPES_H HAVE_LINUX_PERF_EVENT_H_DISABLED
CONFIG_SWSCALE CONFIG_AVUTIL CONFIG_SWSCALE_ALPHA CONFIG_STATIC CONFIG_PIC CONFIG_GPL
CONFIG_VERSION3 CONFIG_SAFE_BITSTREAM_READER CONFIG_FAST_UNALIGNED_DISABLED
""".split())
ON = {t for t in ON if not t.endswith("_DISABLED")}

CFLAGS = ["-std=c17", "-O3", "-fPIC", "-fno-math-errno", "-fno-signed-zeros",
          "-fno-tree-vectorize", "-fomit-frame-pointer", "-pthread", "-w",
          "-D_ISOC11_SOURCE", "-D_FILE_OFFSET_BITS=64", "-D_LARGEFILE_SOURCE",
          "-D_POSIX_C_SOURCE=200112", "-D_XOPEN_SOURCE=600", "-D_DEFAULT_SOURCE",
          "-DHAVE_AV_CONFIG_H", "-DPIC"]


def scan_tokens(paths):
    pat = re.compile(r"\b(?:HAVE|CONFIG|ARCH)_[A-Z0-9_]+\b")
    toks = set()
    for p in paths:
        try:
            with open(p, "r", errors="replace") as f:
                toks.update(pat.findall(f.read()))
        except OSError:
            pass
    return toks


def gen_headers():
    os.makedirs(os.path.join(GEN, "libavutil"), exist_ok=True)
    scan = []
    for d in ("libswscale", "libavutil", "compat", "libavutil/x86", "libswscale/x86"):
        full = os.path.join(REF, d)
        if os.path.isdir(full):
            scan += [os.path.join(full, f) for f in os.listdir(full)
                     if f.endswith((".c", ".h"))]
    toks = scan_tokens(scan)
    # names the sources build by token pasting (HAVE_ ## ext ## suffix, cpu_internal.h)
    exts = """armv5te armv6 armv6t2 armv8 arm_crc dotprod i8mm pmull eor3 neon vfp vfpv3 setend
    sve sve2 sme sme_i16i64 sme2 altivec dcbzl ldbrx power8 ppc4xx vec_xl vsx rv rvv rv_zicbop
    rv_zvbb simd128 aesni clmul amd3dnow amd3dnowext avx avx2 avx512 avx512icl fma3 fma4 mmx
    mmxext sse sse2 sse3 sse4 sse42 ssse3 xop i686 mipsfpu mips32r2 mips32r5 mips64r2 mips32r6
    mips64r6 mipsdsp mipsdspr2 msa loongson2 loongson3 mmi lsx lasx""".split()
    for e in exts:
        for suf in ("", "_EXTERNAL", "_INLINE"):
            toks.add("HAVE_%s%s" % (e.upper(), suf))
    toks = sorted(toks)
    lines = ["/* synthesised by oracle/build_ref.py: generic C-only configuration */",
             "#ifndef FFMPEG_CONFIG_H", "#define FFMPEG_CONFIG_H",
             '#define FFMPEG_CONFIGURATION "--arch=generic --disable-asm (oracle/build_ref.py)"',
             '#define FFMPEG_LICENSE "GPL version 3 or later"',
             "#define CONFIG_THIS_YEAR 2026", '#define FFMPEG_DATADIR "/nonexistent"',
             '#define AVCONV_DATADIR "/nonexistent"', '#define CC_IDENT "gcc"',
             "#define OS_NAME linux", '#define EXTERN_PREFIX ""', "#define EXTERN_ASM",
             '#define BUILDSUF ""', '#define SLIBSUF ".so"',
             "#define SWS_MAX_FILTER_SIZE 256", "#define av_restrict restrict"]
    skip = {"CONFIG_THIS_YEAR", "HAVE_AV_CONFIG_H", "HAVE_MMX2"}
    for t in toks:
        if t in skip:
            continue
        lines.append("#define %s %d" % (t, 1 if t in ON else 0))
    lines += ["#define HAVE_MMX2 HAVE_MMXEXT", "#endif"]
    with open(os.path.join(GEN, "config.h"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(GEN, "config_components.h"), "w") as f:
        f.write("#ifndef FFMPEG_CONFIG_COMPONENTS_H\n#define FFMPEG_CONFIG_COMPONENTS_H\n#endif\n")
    with open(os.path.join(GEN, "libavutil", "avconfig.h"), "w") as f:
        f.write("#ifndef AVUTIL_AVCONFIG_H\n#define AVUTIL_AVCONFIG_H\n"
                "#define AV_HAVE_BIGENDIAN 0\n#define AV_HAVE_FAST_UNALIGNED 1\n#endif\n")
    with open(os.path.join(GEN, "libavutil", "ffversion.h"), "w") as f:
        f.write("#ifndef AVUTIL_FFVERSION_H\n#define AVUTIL_FFVERSION_H\n"
                '#define FFMPEG_VERSION "oracle-ref"\n#endif\n')


def cc(src, obj, extra):
    cmd = ["gcc"] + CFLAGS + extra + ["-I", GEN, "-I", REF, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stderr


def main():
    if not os.path.isdir(os.path.join(REF, "libswscale")):
        if os.path.exists(os.path.join(OUT, "libswsref.so")):
            print("reference tree absent; keeping prebuilt", os.path.join(OUT, "libswsref.so"))
            return 0
        print("reference tree absent and no prebuilt libswsref.so", file=sys.stderr)
        return 1
    os.makedirs(OBJ, exist_ok=True)
    gen_headers()
    jobs = []
    for s in SWS_SRCS:
        jobs.append((os.path.join(REF, "libswscale", s + ".c"),
                     os.path.join(OBJ, "sws_" + s + ".o"), ["-DBUILDING_swscale"]))
    for s in AVU_SRCS:
        jobs.append((os.path.join(REF, "libavutil", s + ".c"),
                     os.path.join(OBJ, "avu_" + s + ".o"), ["-DBUILDING_avutil"]))
    jobs.append((os.path.join(HERE, "ref_shim.c"), os.path.join(OBJ, "ref_shim.o"), []))
    ok = True
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for src, rc, err in ex.map(lambda j: cc(*j), jobs):
            if rc != 0:
                ok = False
                print("FAILED", src, "\n", err[-3000:], file=sys.stderr)
    if not ok:
        return 1
    ver = os.path.join(OUT, "swsref.ver")
    with open(ver, "w") as f:
        f.write("SWSREF { global: swsref_*; local: *; };\n")
    so = os.path.join(OUT, "libswsref.so")
    cmd = ["gcc", "-shared", "-o", so] + [j[1] for j in jobs] + \
          ["-Wl,--version-script=" + ver, "-Wl,-Bsymbolic", "-Wl,--no-undefined",
           "-lm", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr[-6000:], file=sys.stderr)
        return 1
    print("built", so)
    # the FATE input generator (tests/videogen.c): lets the suite check the reference's own
    # golden CRCs (tests/ref/fate/filter-scalechroma, sws-yuv-range) on vsynth1
    vg = os.path.join(OUT, "videogen")
    r = subprocess.run(["gcc", "-O2", "-w", "-I", GEN, "-I", REF, os.path.join(REF, "tests", "videogen.c"),
                        "-o", vg, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr[-3000:], file=sys.stderr)
        return 1
    print("built", vg)
    return 0


if __name__ == "__main__":
    sys.exit(main())
