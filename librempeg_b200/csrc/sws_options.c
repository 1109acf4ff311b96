/*
 * sws_options.c -- the AVClass of SwsContext and its AVOption table.
 *
 * Mirrors reference libswscale/options.c:34-131: the same option names, units, types, ranges and
 * defaults, each bound to offsetof(SwsContext, field) -- the public structure has the reference's field
 * order (include/swscale_b200.h), so libavutil's generic option code (av_opt_set & co.) reads and writes a
 * context of this library exactly like one of the reference's.  What libavfilter/vf_scale.c sets by name
 * (sws_flags, threads, src/dst_format, srcw..., param0/1, *_chr_pos, sws_dither, alphablend, intent) all
 * resolves here.  Host-only, no device.
 */
#include <limits.h>
#include <stddef.h>

#include "sws_internal.h"
#include "swscale_b200_opt.h"

#define OFF(x) (int)offsetof(SwsContext, x)
#define VE (AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_ENCODING_PARAM)

#define OPT_FLAGS(nm, hp, field, dflt, un)      { nm, hp, OFF(field), AV_OPT_TYPE_FLAGS, { .i64 = dflt }, 0, UINT_MAX, VE, un }
#define OPT_INT(nm, hp, field, dflt, lo, hi, un) { nm, hp, OFF(field), AV_OPT_TYPE_INT,   { .i64 = dflt }, lo, hi, VE, un }
#define OPT_BOOL(nm, hp, field, dflt)           { nm, hp, OFF(field), AV_OPT_TYPE_BOOL,  { .i64 = dflt }, 0, 1, VE, NULL }
#define OPT_DBL(nm, hp, field, dflt)            { nm, hp, OFF(field), AV_OPT_TYPE_DOUBLE, { .dbl = dflt }, INT_MIN, INT_MAX, VE, NULL }
#define OPT_PIXFMT(nm, hp, field)               { nm, hp, OFF(field), AV_OPT_TYPE_PIXEL_FMT, { .i64 = 0 }, 0, INT_MAX, VE, NULL }
#define OPT_CONST(nm, hp, val, un)              { nm, hp, 0, AV_OPT_TYPE_CONST, { .i64 = val }, 0, 0, VE, un }

/* backend bits of the reference (swscale.h:110-128); this library is one backend and ignores the mask */
enum { BK_LEGACY = 1, BK_C = 2, BK_MEMCPY = 4, BK_X86 = 8, BK_AARCH64 = 16, BK_SPIRV = 32 };

static const AVOption sws_b200_options[] = {
    OPT_FLAGS("sws_flags", "swscale flags", flags, SWS_BICUBIC, "sws_flags"),
    OPT_CONST("fast_bilinear",   "fast bilinear",                  SWS_FAST_BILINEAR,   "sws_flags"),
    OPT_CONST("bilinear",        "bilinear",                       SWS_BILINEAR,        "sws_flags"),
    OPT_CONST("bicubic",         "bicubic",                        SWS_BICUBIC,         "sws_flags"),
    OPT_CONST("experimental",    "experimental",                   SWS_X,               "sws_flags"),
    OPT_CONST("neighbor",        "nearest neighbor",               SWS_POINT,           "sws_flags"),
    OPT_CONST("area",            "averaging area",                 SWS_AREA,            "sws_flags"),
    OPT_CONST("bicublin",        "luma bicubic, chroma bilinear",  SWS_BICUBLIN,        "sws_flags"),
    OPT_CONST("gauss",           "gaussian approximation",         SWS_GAUSS,           "sws_flags"),
    OPT_CONST("sinc",            "sinc",                           SWS_SINC,            "sws_flags"),
    OPT_CONST("lanczos",         "lanczos (sinc/sinc)",            SWS_LANCZOS,         "sws_flags"),
    OPT_CONST("spline",          "natural bicubic spline",         SWS_SPLINE,          "sws_flags"),
    OPT_CONST("print_info",      "print info",                     SWS_PRINT_INFO,      "sws_flags"),
    OPT_CONST("accurate_rnd",    "accurate rounding",              SWS_ACCURATE_RND,    "sws_flags"),
    OPT_CONST("full_chroma_int", "full chroma interpolation",      SWS_FULL_CHR_H_INT,  "sws_flags"),
    OPT_CONST("full_chroma_inp", "full chroma input",              SWS_FULL_CHR_H_INP,  "sws_flags"),
    OPT_CONST("bitexact",        "bit-exact mode",                 SWS_BITEXACT,        "sws_flags"),
    OPT_CONST("error_diffusion", "error diffusion dither",         SWS_ERROR_DIFFUSION, "sws_flags"),
    OPT_CONST("unstable",        "allow experimental new code",    SWS_UNSTABLE,        "sws_flags"),
    OPT_CONST("strict",          "require all metadata to be set", SWS_STRICT,          "sws_flags"),

    OPT_INT("scaler",     "set scaling algorithm",     scaler,     SWS_SCALE_AUTO, 0, SWS_SCALE_NB - 1, "sws_scaler"),
    OPT_INT("scaler_sub", "set subsampling algorithm", scaler_sub, SWS_SCALE_AUTO, 0, SWS_SCALE_NB - 1, "sws_scaler"),
    OPT_CONST("auto",     "automatic selection",          SWS_SCALE_AUTO,     "sws_scaler"),
    OPT_CONST("bilinear", "bilinear filtering",           SWS_SCALE_BILINEAR, "sws_scaler"),
    OPT_CONST("bicubic",  "2-tap cubic B-spline",         SWS_SCALE_BICUBIC,  "sws_scaler"),
    OPT_CONST("point",    "point sampling",               SWS_SCALE_POINT,    "sws_scaler"),
    OPT_CONST("neighbor", "nearest neighbor",             SWS_SCALE_POINT,    "sws_scaler"),
    OPT_CONST("area",     "area averaging",               SWS_SCALE_AREA,     "sws_scaler"),
    OPT_CONST("gaussian", "2-tap gaussian approximation", SWS_SCALE_GAUSSIAN, "sws_scaler"),
    OPT_CONST("sinc",     "unwindowed sinc",              SWS_SCALE_SINC,     "sws_scaler"),
    OPT_CONST("lanczos",  "3-tap sinc/sinc",              SWS_SCALE_LANCZOS,  "sws_scaler"),
    OPT_CONST("spline",   "2-tap cubic BC spline",        SWS_SCALE_SPLINE,   "sws_scaler"),

    OPT_DBL("param0", "scaler param 0", scaler_params[0], SWS_PARAM_DEFAULT),
    OPT_DBL("param1", "scaler param 1", scaler_params[1], SWS_PARAM_DEFAULT),

    OPT_INT("srcw", "source width",       src_w, 16, 1, INT_MAX, NULL),
    OPT_INT("srch", "source height",      src_h, 16, 1, INT_MAX, NULL),
    OPT_INT("dstw", "destination width",  dst_w, 16, 1, INT_MAX, NULL),
    OPT_INT("dsth", "destination height", dst_h, 16, 1, INT_MAX, NULL),
    OPT_PIXFMT("src_format", "source format",      src_format),
    OPT_PIXFMT("dst_format", "destination format", dst_format),
    OPT_BOOL("src_range", "source is full range",      src_range,  0),
    OPT_BOOL("dst_range", "destination is full range", dst_range,  0),
    OPT_BOOL("gamma",     "gamma correct scaling",     gamma_flag, 0),

    OPT_INT("src_v_chr_pos", "source vertical chroma position in luma grid/256",        src_v_chr_pos, -513, -513, 1024, NULL),
    OPT_INT("src_h_chr_pos", "source horizontal chroma position in luma grid/256",      src_h_chr_pos, -513, -513, 1024, NULL),
    OPT_INT("dst_v_chr_pos", "destination vertical chroma position in luma grid/256",   dst_v_chr_pos, -513, -513, 1024, NULL),
    OPT_INT("dst_h_chr_pos", "destination horizontal chroma position in luma grid/256", dst_h_chr_pos, -513, -513, 1024, NULL),

    OPT_INT("sws_dither", "set dithering algorithm", dither, SWS_DITHER_AUTO, 0, SWS_DITHER_NB - 1, "sws_dither"),
    OPT_CONST("auto",     "automatic selection",        SWS_DITHER_AUTO,     "sws_dither"),
    OPT_CONST("none",     "no dithering",               SWS_DITHER_NONE,     "sws_dither"),
    OPT_CONST("bayer",    "ordered matrix dither",      SWS_DITHER_BAYER,    "sws_dither"),
    OPT_CONST("ed",       "full error diffusion",       SWS_DITHER_ED,       "sws_dither"),
    OPT_CONST("a_dither", "arithmetic addition dither", SWS_DITHER_A_DITHER, "sws_dither"),
    OPT_CONST("x_dither", "arithmetic xor dither",      SWS_DITHER_X_DITHER, "sws_dither"),

    OPT_INT("alphablend", "mode for alpha -> non alpha", alpha_blend, SWS_ALPHA_BLEND_NONE, 0, SWS_ALPHA_BLEND_NB - 1, "alphablend"),
    OPT_CONST("none",          "ignore alpha",               SWS_ALPHA_BLEND_NONE,         "alphablend"),
    OPT_CONST("uniform_color", "blend onto a uniform color", SWS_ALPHA_BLEND_UNIFORM,      "alphablend"),
    OPT_CONST("checkerboard",  "blend onto a checkerboard",  SWS_ALPHA_BLEND_CHECKERBOARD, "alphablend"),

    OPT_INT("threads", "number of threads", threads, 1, 0, INT_MAX, "threads"),
    OPT_CONST("auto", "automatic selection", 0, "threads"),

    OPT_INT("intent", "color mapping intent", intent, 1, 0, 3, "intent"),
    OPT_CONST("perceptual",            "perceptual tone mapping",        0, "intent"),
    OPT_CONST("relative_colorimetric", "relative colorimetric clipping", 1, "intent"),
    OPT_CONST("saturation",            "saturation mapping",             2, "intent"),
    OPT_CONST("absolute_colorimetric", "absolute colorimetric clipping", 3, "intent"),

    OPT_FLAGS("sws_backends", "set allowed swscale backends", backends, 0, "sws_backend"),
    OPT_CONST("auto",     "automatic selection",           0,                                         "sws_backend"),
    OPT_CONST("stable",   "All stable backends",           BK_LEGACY,                                 "sws_backend"),
    OPT_CONST("unstable", "All unstable backends",         BK_C | BK_MEMCPY | BK_X86 | BK_AARCH64 | BK_SPIRV, "sws_backend"),
    OPT_CONST("all",      "All available backends",        BK_LEGACY | BK_C | BK_MEMCPY | BK_X86 | BK_AARCH64 | BK_SPIRV, "sws_backend"),
    OPT_CONST("legacy",   "legacy swscale code",           BK_LEGACY,                                 "sws_backend"),
    OPT_CONST("c",        "template-based reference code", BK_C,                                      "sws_backend"),
    OPT_CONST("memcpy",   "fast path using libc memcpy",   BK_MEMCPY,                                 "sws_backend"),
    OPT_CONST("x86",      "x86 SIMD kernels",              BK_X86,                                    "sws_backend"),
    OPT_CONST("aarch64",  "AArch64 NEON kernels",          BK_AARCH64,                                "sws_backend"),
    OPT_CONST("spirv",    "Vulkan SPIR-V backend",         BK_SPIRV,                                  "sws_backend"),

    { NULL, NULL, 0, 0, { 0 }, 0, 0, 0, NULL },
};

static const char *sws_b200_item_name(void *ctx)
{
    (void)ctx;
    return "swscaler";
}

/* LIBAVUTIL_VERSION_INT of the reference tree (libavutil/version.h:81-83): the class layout version
 * libavutil checks before reading the newer AVClass members */
#define SWS_B200_LAVU_VERSION ((61 << 16) | (5 << 8) | 100)

static const AVClass sws_b200_class = {
    .class_name = "SWScaler",
    .item_name  = sws_b200_item_name,
    .option     = sws_b200_options,
    .version    = SWS_B200_LAVU_VERSION,
    .category   = AV_CLASS_CATEGORY_SWSCALER,
};

const struct AVClass *sws_get_class(void)
{
    return &sws_b200_class;
}

/* option defaults for sws_alloc_context() (the reference runs av_opt_set_defaults() there, utils.c:1042) */
void ff_b200_option_defaults(SwsContext *s)
{
    for (const AVOption *o = sws_b200_options; o->name; o++) {
        void *dst = (uint8_t *)s + o->offset;
        switch (o->type) {
        case AV_OPT_TYPE_FLAGS: case AV_OPT_TYPE_INT: case AV_OPT_TYPE_BOOL: case AV_OPT_TYPE_PIXEL_FMT:
            *(int *)dst = (int)o->default_val.i64;
            break;
        case AV_OPT_TYPE_DOUBLE:
            *(double *)dst = o->default_val.dbl;
            break;
        default:
            break;
        }
    }
}
