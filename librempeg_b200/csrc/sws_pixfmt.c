/*
 * sws_pixfmt.c -- the slice of libavutil's pixel-format descriptor table
 * (reference libavutil/pixdesc.c) that the CUDA hot path understands.
 * Values (depth, chroma shifts, bits per pixel) are properties of the formats
 * themselves; the numeric format ids are ABI (libavutil/pixfmt.h).
 */
#include <stddef.h>
#include "sws_internal.h"

#define YUVP(id, nm, d, cw, ch, bpp) \
    { id, nm, SWSPF_PLANAR, d, cw, ch, bpp, 3, 0, 1, 1, 0 }

static const SwsPixDesc table[] = {
    YUVP(AV_PIX_FMT_YUV420P,     "yuv420p",     8, 1, 1, 12),
    YUVP(AV_PIX_FMT_YUV422P,     "yuv422p",     8, 1, 0, 16),
    YUVP(AV_PIX_FMT_YUV444P,     "yuv444p",     8, 0, 0, 24),
    { AV_PIX_FMT_YUVJ420P, "yuvj420p", SWSPF_PLANAR | SWSPF_JPEG, 8, 1, 1, 12, 3, 0, 1, 1, 0 },
    { AV_PIX_FMT_YUVJ422P, "yuvj422p", SWSPF_PLANAR | SWSPF_JPEG, 8, 1, 0, 16, 3, 0, 1, 1, 0 },
    { AV_PIX_FMT_YUVJ444P, "yuvj444p", SWSPF_PLANAR | SWSPF_JPEG, 8, 0, 0, 24, 3, 0, 1, 1, 0 },
    YUVP(AV_PIX_FMT_YUV420P9LE,  "yuv420p9le",  9, 1, 1, 13),
    YUVP(AV_PIX_FMT_YUV422P9LE,  "yuv422p9le",  9, 1, 0, 18),
    YUVP(AV_PIX_FMT_YUV444P9LE,  "yuv444p9le",  9, 0, 0, 27),
    YUVP(AV_PIX_FMT_YUV420P10LE, "yuv420p10le", 10, 1, 1, 15),
    YUVP(AV_PIX_FMT_YUV422P10LE, "yuv422p10le", 10, 1, 0, 20),
    YUVP(AV_PIX_FMT_YUV444P10LE, "yuv444p10le", 10, 0, 0, 30),
    YUVP(AV_PIX_FMT_YUV420P12LE, "yuv420p12le", 12, 1, 1, 18),
    YUVP(AV_PIX_FMT_YUV422P12LE, "yuv422p12le", 12, 1, 0, 24),
    YUVP(AV_PIX_FMT_YUV444P12LE, "yuv444p12le", 12, 0, 0, 36),
    YUVP(AV_PIX_FMT_YUV420P14LE, "yuv420p14le", 14, 1, 1, 21),
    YUVP(AV_PIX_FMT_YUV422P14LE, "yuv422p14le", 14, 1, 0, 28),
    YUVP(AV_PIX_FMT_YUV444P14LE, "yuv444p14le", 14, 0, 0, 42),
    YUVP(AV_PIX_FMT_YUV420P16LE, "yuv420p16le", 16, 1, 1, 24),
    YUVP(AV_PIX_FMT_YUV422P16LE, "yuv422p16le", 16, 1, 0, 32),
    YUVP(AV_PIX_FMT_YUV444P16LE, "yuv444p16le", 16, 0, 0, 48),
    { AV_PIX_FMT_NV12, "nv12", SWSPF_SEMI, 8, 1, 1, 12, 2, 0, 1, 1, 0 },
    { AV_PIX_FMT_NV21, "nv21", SWSPF_SEMI, 8, 1, 1, 12, 2, 1, 1, 1, 0 },
    { AV_PIX_FMT_P010LE, "p010le", SWSPF_SEMI, 10, 1, 1, 15, 2, 0, 1, 1, 6 },
    { AV_PIX_FMT_RGB24,   "rgb24",   SWSPF_RGB, 8, 0, 0, 24, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_BGR24,   "bgr24",   SWSPF_RGB, 8, 0, 0, 24, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_RGBA,    "rgba",    SWSPF_RGB, 8, 0, 0, 32, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_BGRA,    "bgra",    SWSPF_RGB, 8, 0, 0, 32, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_ARGB,    "argb",    SWSPF_RGB, 8, 0, 0, 32, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_ABGR,    "abgr",    SWSPF_RGB, 8, 0, 0, 32, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_RGB48LE, "rgb48le", SWSPF_RGB, 16, 0, 0, 48, 1, 0, 1, 1, 0 },
    { AV_PIX_FMT_BGR48LE, "bgr48le", SWSPF_RGB, 16, 0, 0, 48, 1, 0, 1, 1, 0 },
    /* 32-bit float destinations (output.c:219-316 yuv2plane1/X_float, output.c:2536-2610 yuv2gbrpf32_full_X_c): the
     * 16-bit integer result of the 19-bit pipeline times 1.0f / 65535.0f */
    { AV_PIX_FMT_GRAYF32LE, "grayf32le", SWSPF_PLANAR | SWSPF_GRAY, 32, 0, 0, 32, 1, 0, 0, 1, 0 },
    { AV_PIX_FMT_GBRPF32LE, "gbrpf32le", SWSPF_PLANAR | SWSPF_RGB, 32, 0, 0, 96, 3, 0, 0, 1, 0 },
    /* 16 bpp packed RGB: destinations only (the rgb16/15 readers are not on the CUDA hot path) */
    { AV_PIX_FMT_RGB565LE, "rgb565le", SWSPF_RGB, 5, 0, 0, 16, 1, 0, 0, 1, 0 },
    { AV_PIX_FMT_BGR565LE, "bgr565le", SWSPF_RGB, 5, 0, 0, 16, 1, 0, 0, 1, 0 },
    { AV_PIX_FMT_RGB555LE, "rgb555le", SWSPF_RGB, 5, 0, 0, 15, 1, 0, 0, 1, 0 },
    { AV_PIX_FMT_BGR555LE, "bgr555le", SWSPF_RGB, 5, 0, 0, 15, 1, 0, 0, 1, 0 },
};

const SwsPixDesc *ff_b200_pix_desc(int fmt)
{
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++)
        if (table[i].fmt == fmt)
            return &table[i];
    return NULL;
}

int sws_isSupportedInput(enum AVPixelFormat pix_fmt)
{
    const SwsPixDesc *d = ff_b200_pix_desc(pix_fmt);
    return d ? d->as_input : 0;
}

int sws_isSupportedOutput(enum AVPixelFormat pix_fmt)
{
    const SwsPixDesc *d = ff_b200_pix_desc(pix_fmt);
    return d ? d->as_output : 0;
}

int sws_isSupportedEndiannessConversion(enum AVPixelFormat pix_fmt)
{
    (void)pix_fmt;
    return 0; /* big-endian twins are not part of the CUDA hot path */
}

int sws_test_format(enum AVPixelFormat format, int output)
{
    return output ? sws_isSupportedOutput(format) : sws_isSupportedInput(format);
}

/* the reference answers 1 for NONE and Vulkan (format.c:616-625); here the hardware format is CUDA */
int sws_test_hw_format(enum AVPixelFormat format)
{
    return format == AV_PIX_FMT_NONE || format == AV_PIX_FMT_CUDA;
}
