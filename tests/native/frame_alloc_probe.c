/* Test-only (GPU): sws_scale_frame() of libswscale_b200.so with a destination AVFrame that carries no buffers.
 * The frames are REAL libavutil AVFrames (the reference's libavutil objects are linked into this binary with
 * -rdynamic so the library finds av_frame_get_buffer() in the running process, as it would in ffmpeg).
 * Prototypes are declared by hand: the file builds on a box without /root/reference.
 * usage: frame_alloc_probe <mode: dynamic|legacy> <sw> <sh> <sfmt> <dw> <dh> <dfmt> <flags> <src.bin> <dst.bin> */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swscale_b200.h"
#include "swscale_b200_frame.h"
#include "swscale_b200_cuda.h"

AVFrame *av_frame_alloc(void);
void av_frame_free(AVFrame **frame);
int av_frame_get_buffer(AVFrame *frame, int align);
int av_pix_fmt_count_planes(int pix_fmt);
int av_image_fill_linesizes(int linesizes[4], int pix_fmt, int width);

static int plane_rows(int fmt, int plane, int h)
{
    const int sub = (fmt == AV_PIX_FMT_YUV420P || fmt == AV_PIX_FMT_NV12 || fmt == AV_PIX_FMT_YUV420P10LE) && plane > 0;
    return sub ? (h + 1) / 2 : h;
}

int main(int argc, char **argv)
{
    if (argc < 11)
        return 2;
    const int legacy = !strcmp(argv[1], "legacy");
    const int sw = atoi(argv[2]), sh = atoi(argv[3]), sf = atoi(argv[4]);
    const int dw = atoi(argv[5]), dh = atoi(argv[6]), df = atoi(argv[7]);
    const unsigned flags = (unsigned)strtoul(argv[8], NULL, 0);
    AVFrame *src = av_frame_alloc(), *dst = av_frame_alloc();
    int ls[4] = { 0 }, ret;
    src->width = sw; src->height = sh; src->format = sf;
    if (av_frame_get_buffer(src, 0) < 0)
        return 3;
    FILE *f = fopen(argv[9], "rb");
    if (!f)
        return 4;
    av_image_fill_linesizes(ls, sf, sw);
    for (int p = 0; p < av_pix_fmt_count_planes(sf); p++)
        for (int y = 0; y < plane_rows(sf, p, sh); y++)
            if (fread(src->data[p] + (size_t)y * src->linesize[p], 1, ls[p], f) != (size_t)ls[p])
                return 5;
    fclose(f);

    SwsContext *c;
    if (legacy) {
        c = sws_getContext(sw, sh, sf, dw, dh, df, (int)flags, NULL, NULL, NULL);
        if (!c) {
            fprintf(stderr, "sws_getContext failed\n");
            return 6;
        }
    } else {
        c = sws_alloc_context();
        c->flags = flags;
        dst->width = dw; dst->height = dh; dst->format = df;       /* described, not allocated */
    }
    if (dst->data[0] || dst->buf[0])
        return 7;
    ret = sws_scale_frame(c, dst, src);
    if (ret < 0) {
        fprintf(stderr, "sws_scale_frame: %d (%s)\n", ret, sws_cuda_last_error(c));
        return 8;
    }
    if (!dst->data[0] || !dst->buf[0] || dst->width != dw || dst->height != dh || dst->format != df)
        return 9;
    f = fopen(argv[10], "wb");
    av_image_fill_linesizes(ls, df, dw);
    for (int p = 0; p < av_pix_fmt_count_planes(df); p++)
        for (int y = 0; y < plane_rows(df, p, dh); y++)
            fwrite(dst->data[p] + (size_t)y * dst->linesize[p], 1, ls[p], f);
    fclose(f);
    sws_free_context(&c);
    av_frame_free(&src);
    av_frame_free(&dst);          /* libavutil frees what libavutil allocated */
    printf("ok\n");
    return 0;
}
