/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin plain-pointer wrapper (ours, not reference code) around the REAL
 * reference libswscale that oracle/build_ref.py compiles from /root/reference.
 * It exists so Python (ctypes) can drive the reference without knowing the
 * layouts of SwsContext/AVFrame, and so tests can read back the tables the
 * reference computed (filters, yuv2rgb constants) to pin our own restatements.
 *
 * Reference entry points wrapped here:
 *   sws_alloc_context/sws_init_context   libswscale/utils.c:1032,1884
 *   sws_scale                            libswscale/swscale.c:1626
 *   sws_scale_frame                      libswscale/swscale.c:1405
 *   sws_setColorspaceDetails             libswscale/utils.c:849
 */
#include <stdint.h>
#include <string.h>
#include <time.h>

#include "config.h"
#include "libavutil/frame.h"
#include "libavutil/buffer.h"
#include "libavutil/hwcontext.h"
#include "libavutil/imgutils.h"
#include "libavutil/pixdesc.h"
#include "libavutil/log.h"
#include "libavutil/lfg.h"
#include "libavutil/md5.h"
#include "libswscale/swscale.h"
#include "libswscale/swscale_internal.h"

#define API __attribute__((visibility("default")))

API int swsref_pix_fmt(const char *name) { return av_get_pix_fmt(name); }
API void swsref_quiet(int level) { av_log_set_level(level); }

/* opts: [0]=threads [1]=src_range [2]=dst_range [3]=src_h_chr_pos [4]=src_v_chr_pos
 *       [5]=dst_h_chr_pos [6]=dst_v_chr_pos [7]=dither   (chr_pos: -513 = default) */
API void *swsref_create(int sw, int sh, int sfmt, int dw, int dh, int dfmt,
                        unsigned flags, const double *param, const int *opts)
{
    SwsContext *s = sws_alloc_context();
    if (!s)
        return NULL;
    s->flags = flags;
    s->src_w = sw; s->src_h = sh; s->src_format = sfmt;
    s->dst_w = dw; s->dst_h = dh; s->dst_format = dfmt;
    if (param) {
        s->scaler_params[0] = param[0];
        s->scaler_params[1] = param[1];
    }
    s->threads = 1;
    if (opts) {
        s->threads       = opts[0];
        s->src_range     = opts[1];
        s->dst_range     = opts[2];
        s->src_h_chr_pos = opts[3];
        s->src_v_chr_pos = opts[4];
        s->dst_h_chr_pos = opts[5];
        s->dst_v_chr_pos = opts[6];
        s->dither        = opts[7];
    }
    if (sws_init_context(s, NULL, NULL) < 0) {
        sws_freeContext(s);
        return NULL;
    }
    return s;
}

/* the same with SwsFilter vectors: vec[0..3] = srcFilter lumH, lumV, chrH, chrV; vec[4..7] = dstFilter (NULL = none) */
API void *swsref_create_filtered(int sw, int sh, int sfmt, int dw, int dh, int dfmt,
                                 unsigned flags, const double *param, const int *opts,
                                 const double *const vec[8], const int len[8])
{
    SwsVector v[8];
    SwsFilter f[2];
    SwsVector **slot[8] = { &f[0].lumH, &f[0].lumV, &f[0].chrH, &f[0].chrV, &f[1].lumH, &f[1].lumV, &f[1].chrH, &f[1].chrV };
    int used[2] = { 0, 0 };
    SwsContext *s = sws_alloc_context();
    if (!s)
        return NULL;
    for (int i = 0; i < 8; i++) {
        *slot[i] = NULL;
        if (vec[i] && len[i] > 0) {
            v[i].coeff = (double *)vec[i];
            v[i].length = len[i];
            *slot[i] = &v[i];
            used[i / 4] = 1;
        }
    }
    s->flags = flags;
    s->src_w = sw; s->src_h = sh; s->src_format = sfmt;
    s->dst_w = dw; s->dst_h = dh; s->dst_format = dfmt;
    if (param) {
        s->scaler_params[0] = param[0];
        s->scaler_params[1] = param[1];
    }
    s->threads = 1;
    if (opts) {
        s->threads       = opts[0];
        s->src_range     = opts[1];
        s->dst_range     = opts[2];
        s->src_h_chr_pos = opts[3];
        s->src_v_chr_pos = opts[4];
        s->dst_h_chr_pos = opts[5];
        s->dst_v_chr_pos = opts[6];
        s->dither        = opts[7];
    }
    if (sws_init_context(s, used[0] ? &f[0] : NULL, used[1] ? &f[1] : NULL) < 0) {
        sws_freeContext(s);
        return NULL;
    }
    return s;
}

API void swsref_free(void *ctx) { sws_freeContext(ctx); }

API int swsref_set_colorspace(void *ctx, int src_cs, int src_range, int dst_cs,
                              int dst_range, int brightness, int contrast, int saturation)
{
    return sws_setColorspaceDetails(ctx, sws_getCoefficients(src_cs), src_range,
                                    sws_getCoefficients(dst_cs), dst_range,
                                    brightness, contrast, saturation);
}

API int swsref_scale(void *ctx, const uint8_t *const src[4], const int src_stride[4],
                     int y, int h, uint8_t *const dst[4], const int dst_stride[4])
{
    return sws_scale(ctx, src, src_stride, y, h, dst, dst_stride);
}

/* ---- refcounted frames so the threaded sws_scale_frame() path does no copies ---- */
API void *swsref_frame_alloc(int w, int h, int fmt)
{
    AVFrame *f = av_frame_alloc();
    if (!f)
        return NULL;
    f->width = w; f->height = h; f->format = fmt;
    if (av_frame_get_buffer(f, 64) < 0) {
        av_frame_free(&f);
        return NULL;
    }
    return f;
}
API void swsref_frame_free(void *f) { AVFrame *fr = f; av_frame_free(&fr); }
API uint8_t *swsref_frame_data(void *f, int p) { return ((AVFrame *)f)->data[p]; }
API int swsref_frame_linesize(void *f, int p) { return ((AVFrame *)f)->linesize[p]; }
API int swsref_scale_frame(void *ctx, void *dst, void *src)
{
    return sws_scale_frame(ctx, dst, src);
}

/* run `iters` conversions back to back; returns seconds (CLOCK_MONOTONIC) or <0 */
API double swsref_bench_frame(void *ctx, void *dst, void *src, int iters)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < iters; i++)
        if (sws_scale_frame(ctx, dst, src) < 0)
            return -1.0;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* the same over a ring of `nb` distinct frame pairs (working set beyond the last-level cache, like a real
 * stream of frames): conversion i uses pair i mod nb */
API double swsref_bench_frames(void *ctx, void *const dst[], void *const src[], int nb, int iters)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < iters; i++)
        if (sws_scale_frame(ctx, dst[i % nb], src[i % nb]) < 0)
            return -1.0;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- introspection: what did the reference compute at init? ---- */
static SwsInternal *first_ctx(void *ctx)
{
    SwsInternal *c = sws_internal(ctx);
    if (c->nb_slice_ctx)
        c = sws_internal(c->slice_ctx[0]);
    return c;
}

/* which: 0=hLum 1=hChr 2=vLum 3=vChr; returns filterSize (0 if unscaled converter in use) */
API int swsref_filter(void *ctx, int which, const int16_t **coef, const int32_t **pos, int *len)
{
    SwsInternal *c = first_ctx(ctx);
    if (c->convert_unscaled || c->cascaded_context[0])
        return 0;
    switch (which) {
    case 0: *coef = c->hLumFilter; *pos = c->hLumFilterPos; *len = c->opts.dst_w;  return c->hLumFilterSize;
    case 1: *coef = c->hChrFilter; *pos = c->hChrFilterPos; *len = c->chrDstW;     return c->hChrFilterSize;
    case 2: *coef = c->vLumFilter; *pos = c->vLumFilterPos; *len = c->opts.dst_h;  return c->vLumFilterSize;
    case 3: *coef = c->vChrFilter; *pos = c->vChrFilterPos; *len = c->chrDstH;     return c->vChrFilterSize;
    }
    return -1;
}

/* out[0..5] = y_offset y_coeff v2r v2g u2g u2b; out[6]=unscaled? out[7]=cascaded?
 * out[8..11] = chrSrcW chrSrcH chrDstW chrDstH; out[12]=srcBpc out[13]=dstBpc; out[14]=flags */
API void swsref_info(void *ctx, int out[16])
{
    SwsInternal *c = first_ctx(ctx);
    out[0] = c->yuv2rgb_y_offset;  out[1] = c->yuv2rgb_y_coeff;
    out[2] = c->yuv2rgb_v2r_coeff; out[3] = c->yuv2rgb_v2g_coeff;
    out[4] = c->yuv2rgb_u2g_coeff; out[5] = c->yuv2rgb_u2b_coeff;
    out[6] = c->convert_unscaled != NULL;
    out[7] = c->cascaded_context[0] != NULL;
    out[8] = c->chrSrcW; out[9] = c->chrSrcH; out[10] = c->chrDstW; out[11] = c->chrDstH;
    out[12] = c->srcBpc; out[13] = c->dstBpc; out[14] = c->opts.flags; out[15] = 0;
}

/* the C entries of input_rgb2yuv_table: ry gy by ru gu bu rv gv bv */
API void swsref_rgb2yuv(void *ctx, int out[9])
{
    SwsInternal *c = first_ctx(ctx);
    for (int i = 0; i < 9; i++)
        out[i] = c->input_rgb2yuv_table[i];
}

/* copy the rgb24/48 LUTs the reference built: y_table[2048]; rV/gU/bU as offsets into it */
API int swsref_rgb_tables(void *ctx, uint8_t y_table[2048], int rV[1280], int gU[1280],
                          int bU[1280], int gV[1280])
{
    SwsInternal *c = first_ctx(ctx);
    if (!c->yuvTable || (c->dstFormatBpp != 24 && c->dstFormatBpp != 48))
        return -1;
    memcpy(y_table, c->yuvTable, 2048);
    for (int i = 0; i < 1280; i++) {
        rV[i] = (int)(c->table_rV[i] - (uint8_t *)c->yuvTable);
        gU[i] = (int)(c->table_gU[i] - (uint8_t *)c->yuvTable);
        bU[i] = (int)(c->table_bU[i] - (uint8_t *)c->yuvTable);
        gV[i] = c->table_gV[i];
    }
    return 0;
}

/* the survey's synthetic-input recipe: av_lfg seeded, sequential fill */
API void swsref_lfg_fill(uint8_t *buf, size_t n, unsigned seed, int bits)
{
    AVLFG r;
    av_lfg_init(&r, seed);
    if (bits <= 8) {
        for (size_t i = 0; i < n; i++)
            buf[i] = (uint8_t)av_lfg_get(&r);
    } else {
        uint16_t *p = (uint16_t *)buf;
        unsigned mask = (1u << bits) - 1;
        for (size_t i = 0; i < n / 2; i++)
            p[i] = av_lfg_get(&r) & mask;
    }
}

API void swsref_md5(uint8_t out[16], const uint8_t *buf, size_t n) { av_md5_sum(out, buf, n); }

/* ---- AVFrame ABI facts and the dynamic (frame-described) mode of sws_scale_frame() ---- */
#include <stddef.h>
API void swsref_frame_offsets(int out[24])
{
    int i = 0;
    out[i++] = offsetof(AVFrame, data);           out[i++] = offsetof(AVFrame, linesize);
    out[i++] = offsetof(AVFrame, extended_data);  out[i++] = offsetof(AVFrame, width);
    out[i++] = offsetof(AVFrame, height);         out[i++] = offsetof(AVFrame, nb_samples);
    out[i++] = offsetof(AVFrame, format);         out[i++] = offsetof(AVFrame, pict_type);
    out[i++] = offsetof(AVFrame, sample_aspect_ratio); out[i++] = offsetof(AVFrame, pts);
    out[i++] = offsetof(AVFrame, pkt_dts);        out[i++] = offsetof(AVFrame, time_base);
    out[i++] = offsetof(AVFrame, quality);        out[i++] = offsetof(AVFrame, opaque);
    out[i++] = offsetof(AVFrame, repeat_pict);    out[i++] = offsetof(AVFrame, sample_rate);
    out[i++] = offsetof(AVFrame, buf);            out[i++] = offsetof(AVFrame, flags);
    out[i++] = offsetof(AVFrame, color_range);    out[i++] = offsetof(AVFrame, color_primaries);
    out[i++] = offsetof(AVFrame, color_trc);      out[i++] = offsetof(AVFrame, colorspace);
    out[i++] = offsetof(AVFrame, chroma_location); out[i++] = offsetof(AVFrame, hw_frames_ctx);
}

/* props: [0..2] = src color_range, colorspace, chroma_location; [3..5] = dst */
API int swsref_scale_frame_dynamic(unsigned flags, int threads, void *dst, void *src, const int props[6])
{
    AVFrame *s = src, *d = dst;
    SwsContext *c = sws_alloc_context();
    int ret;
    if (!c)
        return AVERROR(ENOMEM);
    c->flags = flags;
    c->threads = threads;
    s->color_range = props[0]; s->colorspace = props[1]; s->chroma_location = props[2];
    d->color_range = props[3]; d->colorspace = props[4]; d->chroma_location = props[5];
    ret = sws_scale_frame(c, d, s);
    sws_free_context(&c);
    return ret;
}

API int swsref_is_noop(void *dst, void *src, const int props[6])
{
    AVFrame *s = src, *d = dst;
    s->color_range = props[0]; s->colorspace = props[1]; s->chroma_location = props[2];
    d->color_range = props[3]; d->colorspace = props[4]; d->chroma_location = props[5];
    return sws_is_noop(d, s);
}

/* lumH of sws_getDefaultFilter(): pins our SwsFilter builder; returns the vector length */
API int swsref_default_filter(float lgb, float cgb, float ls, float cs, float chs, float cvs, double out[64])
{
    SwsFilter *f = sws_getDefaultFilter(lgb, cgb, ls, cs, chs, cvs, 0);
    int n;
    if (!f)
        return -1;
    n = f->lumH->length;
    for (int i = 0; i < n && i < 64; i++)
        out[i] = f->lumH->coeff[i];
    sws_freeFilter(f);
    return n;
}

/* ABI facts of the structures an AV_PIX_FMT_CUDA frame carries (include/swscale_b200_frame.h mirrors them) */
API void swsref_hw_offsets(int out[16])
{
    int i = 0;
    out[i++] = offsetof(AVBufferRef, data);             out[i++] = offsetof(AVBufferRef, size);
    out[i++] = offsetof(AVHWDeviceContext, type);       out[i++] = offsetof(AVHWDeviceContext, hwctx);
    out[i++] = offsetof(AVHWFramesContext, device_ref); out[i++] = offsetof(AVHWFramesContext, device_ctx);
    out[i++] = offsetof(AVHWFramesContext, hwctx);      out[i++] = offsetof(AVHWFramesContext, pool);
    out[i++] = offsetof(AVHWFramesContext, initial_pool_size);
    out[i++] = offsetof(AVHWFramesContext, format);     out[i++] = offsetof(AVHWFramesContext, sw_format);
    out[i++] = offsetof(AVHWFramesContext, width);      out[i++] = offsetof(AVHWFramesContext, height);
    out[i++] = AV_HWDEVICE_TYPE_CUDA;                   out[i++] = AV_PIX_FMT_CUDA;
}

/* the reference's AVOption table, one line per entry (tests/test_options_cpu.py holds ours against it) */
#include "libavutil/opt.h"
#include <inttypes.h>
#include <stdio.h>
API int swsref_dump_options(char *buf, int size)
{
    SwsContext *s = sws_alloc_context();
    const AVOption *o = NULL;
    int n = 0;
    if (!s)
        return -1;
    while ((o = av_opt_next(s, o)) && n < size)
        n += snprintf(buf + n, size - n, "opt %s|%s|%d|%d|%" PRId64 "|%g|%g|%g|%d|%s\n", o->name, o->help ? o->help : "",
                      o->offset, (int)o->type, o->type == AV_OPT_TYPE_DOUBLE ? 0 : o->default_val.i64,
                      o->type == AV_OPT_TYPE_DOUBLE ? o->default_val.dbl : 0.0, o->min, o->max, o->flags,
                      o->unit ? o->unit : "");
    sws_free_context(&s);
    return n;
}

#ifdef SWSREF_HOOKED
/* ---- integration/build_hooked.py only: the reference built with ff_sws_init_swscale_cuda() ---- */
/* kernels launched by the B200 context behind this reference context (-1: the hook is not installed,
 * the C kernels run); cascaded sub-contexts are summed */
API long swsref_hook_launches(void *ctx)
{
    SwsInternal *c = first_ctx(ctx);
    long n = ff_sws_cuda_launches(c), any = n >= 0;
    if (n < 0)
        n = 0;
    for (int i = 0; i < 3; i++)
        if (c->cascaded_context[i]) {
            long k = swsref_hook_launches(c->cascaded_context[i]);
            if (k >= 0) {
                n += k;
                any = 1;
            }
        }
    return any ? n : -1;
}
API const char *swsref_hook_kernel(void *ctx) { return ff_sws_cuda_kernel(first_ctx(ctx)); }
/* slices any context of this process handed to the GPU so far (covers the contexts sws_scale_frame() builds internally) */
API long swsref_hook_slices_total(void) { return ff_sws_cuda_slices_total(); }
#endif
