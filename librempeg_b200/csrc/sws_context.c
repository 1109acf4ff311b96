/*
 * sws_context.c -- host side of the B200 libswscale hot path (plain C).
 *
 * Mirrors the legacy API of the reference (libswscale/utils.c, swscale.c):
 *   sws_alloc_context / sws_init_context / sws_getContext / sws_scale / ...
 * Path selection (unscaled LUT converter vs. FIR pipeline), chroma geometry
 * and flag fix-ups follow ff_sws_init_single_context() (utils.c:1137-1835);
 * the per-line machinery it sets up is replaced by a device "plan" handed to
 * the CUDA shim -- the B200 counterpart of ff_sws_init_swscale_<arch>()
 * (swscale.c:697-714), installed at frame granularity like c->convert_unscaled
 * (swscale_internal.h:99-101, swscale.c:1185).
 */
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#include "sws_internal.h"

#define LIBSWSCALE_VERSION_MAJOR 10
#define LIBSWSCALE_VERSION_MINOR 2
#define LIBSWSCALE_VERSION_MICRO 100

unsigned swscale_version(void)
{
    return (LIBSWSCALE_VERSION_MAJOR << 16) | (LIBSWSCALE_VERSION_MINOR << 8) | LIBSWSCALE_VERSION_MICRO;
}
const char *swscale_configuration(void) { return "b200-native sm_100a (no CPU fallback)"; }
const char *swscale_license(void) { return "LGPL version 2.1 or later"; }

static void set_error(SwsInternal *c, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->last_error, sizeof(c->last_error), fmt, ap);
    va_end(ap);
    if (c->opts.flags & SWS_PRINT_INFO)
        fprintf(stderr, "[swscaler-b200] %s\n", c->last_error);
}

/* ------------------------------------------------------------ life cycle */

SwsContext *sws_alloc_context(void)
{
    SwsInternal *c = calloc(1, sizeof(*c));
    if (!c)
        return NULL;
    /* av_class first, then the defaults of the AVOption table (reference utils.c:1032-1045) */
    c->opts.av_class = sws_get_class();
    ff_b200_option_defaults(&c->opts);
    return &c->opts;
}

static void release_lanes(SwsInternal *c)
{
    for (int i = 0; i < c->nb_lanes; i++)
        if (c->lanes[i])
            ff_b200_cuda_destroy(c->lanes[i]);
    memset(c->lanes, 0, sizeof(c->lanes));
    c->nb_lanes = 0;
}

static void release_cascade(SwsInternal *c)
{
    for (int i = 0; i < 2; i++) {
        if (c->cascade[i])
            sws_freeContext(c->cascade[i]);
        c->cascade[i] = NULL;
    }
    for (int i = 0; i < 4; i++) {
        if (c->cascade_tmp[i])
            ff_b200_cuda_free(c->cascade_tmp[i]);
        c->cascade_tmp[i] = NULL;
        c->cascade_tmp_stride[i] = 0;
    }
    c->cascade_main = 0;
    c->cascade_frame_done = 0;
}

static void release_tables(SwsInternal *c)
{
    release_cascade(c);
    release_lanes(c);
    if (c->cuda)
        ff_b200_cuda_destroy(c->cuda);
    c->cuda = NULL;
    ff_b200_free_fir(&c->h_lum);
    ff_b200_free_fir(&c->h_chr);
    ff_b200_free_fir(&c->v_lum);
    ff_b200_free_fir(&c->v_chr);
    c->initialized = 0;
    c->planned = 0;
}

void sws_freeContext(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    if (!c)
        return;
    release_tables(c);
    if (c->dyn)
        sws_freeContext(c->dyn);
    free(c->dyn_key);
    free(c);
}

void sws_free_context(SwsContext **pctx)
{
    if (!pctx || !*pctx)
        return;
    sws_freeContext(*pctx);
    *pctx = NULL;
}

/* ------------------------------------------------------------ helpers */

static int is_rgb(int fmt)
{
    const SwsPixDesc *d = ff_b200_pix_desc(fmt);
    return d && (d->flags & SWSPF_RGB);
}

static int ceil_rshift(int a, int b) { return -((-a) >> b); }

/* deprecated yuvj* ids -> plain id + full range (utils.c:773-809) */
static int fold_jpeg_format(int *fmt)
{
    switch (*fmt) {
    case AV_PIX_FMT_YUVJ420P: *fmt = AV_PIX_FMT_YUV420P; return 1;
    case AV_PIX_FMT_YUVJ422P: *fmt = AV_PIX_FMT_YUV422P; return 1;
    case AV_PIX_FMT_YUVJ444P: *fmt = AV_PIX_FMT_YUV444P; return 1;
    }
    return 0;
}

/* chroma siting -> position relative to the plane's own grid (utils.c:168-175) */
static int local_chroma_pos(int sub, int pos)
{
    if (pos == -1 || pos <= -513)
        pos = (128 << sub) - 128;
    pos += 128;
    return pos >> sub;
}

static int scaler_enum_to_flag(SwsScaler s, int fallback)
{
    switch (s) {
    case SWS_SCALE_BILINEAR: return SWS_BILINEAR;
    case SWS_SCALE_BICUBIC:  return SWS_BICUBIC;
    case SWS_SCALE_POINT:    return SWS_POINT;
    case SWS_SCALE_AREA:     return SWS_AREA;
    case SWS_SCALE_GAUSSIAN: return SWS_GAUSS;
    case SWS_SCALE_SINC:     return SWS_SINC;
    case SWS_SCALE_LANCZOS:  return SWS_LANCZOS;
    case SWS_SCALE_SPLINE:   return SWS_SPLINE;
    default:                 return fallback;
    }
}

/* limited<->full constants for the h-scaled lines (swscale.c:577-624) */
static void solve_range(unsigned src_min, unsigned src_max, unsigned dst_min, unsigned dst_max,
                        int src_shift, int mult_shift, uint32_t *coeff, int64_t *offset)
{
    const unsigned src_range = (uint16_t)(src_max - src_min);
    const unsigned dst_range = (uint16_t)(dst_max - dst_min);
    const int total = mult_shift + src_shift;
    const uint64_t q = ((uint64_t)dst_range << total) / src_range;
    *coeff  = (uint32_t)((q + (1ULL << src_shift) - 1) >> src_shift);
    *offset = ((int64_t)dst_max << total) - ((int64_t)src_max << src_shift) * *coeff +
              (1U << (mult_shift - 1));
}

static void plan_range_convert(SwsInternal *c)
{
    SwsCudaPlan *p = &c->plan;
    p->range_mode = 0;
    p->src_full_range = c->opts.src_range;
    p->dither_none = c->opts.dither == SWS_DITHER_NONE;
    if (c->opts.src_range == c->opts.dst_range || is_rgb(c->opts.dst_format) || c->dst_bpc >= 32)
        return;
    {
        const int bd = c->dst_bpc ? (c->dst_bpc < 16 ? c->dst_bpc : 16) : 8;
        const int src_bits = bd <= 14 ? 15 : 19;
        const int src_shift = src_bits - bd;
        const int mult_shift = bd <= 14 ? 14 : 18;
        const unsigned mpeg_min = 16U << (bd - 8), mpeg_lum = 235U << (bd - 8);
        const unsigned mpeg_chr = 240U << (bd - 8), jpeg_max = (1U << bd) - 1;
        if (c->opts.src_range) {   /* full -> limited */
            solve_range(0, jpeg_max, mpeg_min, mpeg_lum, src_shift, mult_shift, &p->lum_rc_coeff, &p->lum_rc_offset);
            solve_range(0, jpeg_max, mpeg_min, mpeg_chr, src_shift, mult_shift, &p->chr_rc_coeff, &p->chr_rc_offset);
            p->range_mode = 2;
        } else {                   /* limited -> full, clipped */
            solve_range(mpeg_min, mpeg_lum, 0, jpeg_max, src_shift, mult_shift, &p->lum_rc_coeff, &p->lum_rc_offset);
            solve_range(mpeg_min, mpeg_chr, 0, jpeg_max, src_shift, mult_shift, &p->chr_rc_coeff, &p->chr_rc_offset);
            p->range_mode = 1;
        }
        if (bd <= 14) { /* the 15-bit kernels take uint16 coeff / int32 offset (swscale.c:166-216) */
            p->lum_rc_coeff = (uint16_t)p->lum_rc_coeff;
            p->chr_rc_coeff = (uint16_t)p->chr_rc_coeff;
            p->lum_rc_offset = (int32_t)p->lum_rc_offset;
            p->chr_rc_offset = (int32_t)p->chr_rc_offset;
        }
    }
}

static int plan_colorspace(SwsInternal *c)
{
    if (is_rgb(c->opts.src_format))
        ff_b200_rgb2yuv_table(c->plan.rgb2yuv, c->dst_colorspace);
    if (!is_rgb(c->opts.dst_format))
        return 0;
    return ff_b200_rgb_consts(&c->plan.rgb, c->src_colorspace, c->opts.src_range,
                              c->brightness, c->contrast, c->saturation);
}

/* ------------------------------------------------------------ cascaded contexts
 * The reference splits a conversion it cannot do in one pass into two contexts with an intermediate picture
 * (utils.c:1803-1832: a filter longer than the per-line kernels take; utils.c:915-984: YUV -> YUV with two
 * different matrices goes through packed RGB) and runs them back to back (scale_cascaded(), swscale.c:992-1020).
 * Same here, with the intermediate picture in HBM: stage 0 stores to device planes, stage 1 reads them. */

static int cascade_alloc_tmp(SwsInternal *c, int w, int h, int fmt)
{
    const SwsPixDesc *d = ff_b200_pix_desc(fmt);
    if (!d)
        return AVERROR(EINVAL);
    for (int i = 0; i < d->nb_planes; i++) {
        const int chroma = i == 1 || i == 2;
        const int pw = chroma ? ceil_rshift(w, d->log2_cw) : w, ph = chroma ? ceil_rshift(h, d->log2_ch) : h;
        int rowbytes;
        if (d->flags & SWSPF_RGB)
            rowbytes = pw * (d->bpp / 8);
        else
            rowbytes = pw * (d->depth > 8 ? 2 : 1) * ((d->flags & SWSPF_SEMI) && i == 1 ? 2 : 1);
        c->cascade_tmp_stride[i] = (rowbytes + 63) & ~63;          /* av_image_alloc(..., 64) */
        c->cascade_tmp[i] = ff_b200_cuda_alloc((size_t)c->cascade_tmp_stride[i] * ph);
        if (!c->cascade_tmp[i])
            return AVERROR(ENOMEM);
    }
    return 0;
}

static SwsContext *cascade_stage(const SwsContext *parent, int sw, int sh, int sf, int dw, int dh, int df)
{
    SwsContext *s = sws_alloc_context();                            /* alloc_set_opts(), utils.c:1076-1102 */
    if (!s)
        return NULL;
    s->flags = parent->flags;
    s->src_w = sw; s->src_h = sh; s->src_format = sf;
    s->dst_w = dw; s->dst_h = dh; s->dst_format = df;
    s->scaler_params[0] = parent->scaler_params[0];
    s->scaler_params[1] = parent->scaler_params[1];
    return s;
}

/* utils.c:1803-1832 */
static int cascade_long_filter(SwsContext *sws, SwsFilter *srcFilter, SwsFilter *dstFilter)
{
    SwsInternal *c = sws_internal(sws);
    const int srcW = sws->src_w, srcH = sws->src_h, dstW = sws->dst_w, dstH = sws->dst_h;
    const int tmpW = (int)sqrt((double)(srcW * (int64_t)dstW)), tmpH = (int)sqrt((double)(srcH * (int64_t)dstH));
    int ret;
    if (srcW * (int64_t)srcH <= 4LL * dstW * dstH)
        return AVERROR(EINVAL);
    if ((ret = cascade_alloc_tmp(c, tmpW, tmpH, AV_PIX_FMT_YUV420P)) < 0)
        return ret;
    c->cascade[0] = cascade_stage(sws, srcW, srcH, sws->src_format, tmpW, tmpH, AV_PIX_FMT_YUV420P);
    c->cascade[1] = cascade_stage(sws, tmpW, tmpH, AV_PIX_FMT_YUV420P, dstW, dstH, sws->dst_format);
    if (!c->cascade[0] || !c->cascade[1])
        return AVERROR(ENOMEM);
    if ((ret = sws_init_context(c->cascade[0], srcFilter, NULL)) < 0 ||
        (ret = sws_init_context(c->cascade[1], NULL, dstFilter)) < 0) {
        set_error(c, "cascade stage failed: %s | %s", sws_internal(c->cascade[0])->last_error,
                  sws_internal(c->cascade[1])->last_error);
        return ret;
    }
    c->cascade_main = 0;
    return 0;
}

/* utils.c:915-984: YUV (or gray) on both sides and two different matrices */
static int cascade_matrix_change(SwsContext *sws, const int inv_table[4], int srcRange, const int table[4],
                                 int dstRange, int brightness, int contrast, int saturation)
{
    SwsInternal *c = sws_internal(sws);
    const int srcW = sws->src_w, srcH = sws->src_h, dstW = sws->dst_w, dstH = sws->dst_h;
    /* isNBPS(dst) || is16BPS(dst) (utils.c:928): depths 9..16; a 32-bit float gray destination is neither and goes
     * through bgr24 like the 8-bit formats */
    const int tmp_fmt = c->dst_bpc > 8 && c->dst_bpc <= 16 ? AV_PIX_FMT_BGR48LE : AV_PIX_FMT_BGR24;
    const int small = srcW * (int64_t)srcH > dstW * (int64_t)dstH;
    const int tmpW = small ? dstW : srcW, tmpH = small ? dstH : srcH;
    int ret;
    if (!ff_b200_pix_desc(tmp_fmt)->as_input) {
        set_error(c, "YUV->YUV matrix change into a >8-bit destination needs a 16-bit RGB source reader");
        return AVERROR(ENOTSUP);
    }
    if ((ret = cascade_alloc_tmp(c, tmpW, tmpH, tmp_fmt)) < 0)
        return ret;
    c->cascade[0] = cascade_stage(sws, srcW, srcH, sws->src_format, tmpW, tmpH, tmp_fmt);
    c->cascade[1] = cascade_stage(sws, tmpW, tmpH, tmp_fmt, dstW, dstH, sws->dst_format);
    if (!c->cascade[0] || !c->cascade[1])
        return AVERROR(ENOMEM);
    c->cascade[0]->alpha_blend = sws->alpha_blend;
    if ((ret = sws_init_context(c->cascade[0], NULL, NULL)) < 0) {
        set_error(c, "cascade stage 0: %s", sws_internal(c->cascade[0])->last_error);
        return ret;
    }
    /* both sides are set although the RGB side of each stage is ignored */
    sws_setColorspaceDetails(c->cascade[0], inv_table, srcRange, table, dstRange, brightness, contrast, saturation);
    c->cascade[1]->src_range = srcRange;
    c->cascade[1]->dst_range = dstRange;
    if ((ret = sws_init_context(c->cascade[1], NULL, NULL)) < 0) {
        set_error(c, "cascade stage 1: %s", sws_internal(c->cascade[1])->last_error);
        return ret;
    }
    sws_setColorspaceDetails(c->cascade[1], inv_table, srcRange, table, dstRange, 0, 1 << 16, 1 << 16);
    c->cascade_main = 0;
    return 0;
}

/* ------------------------------------------------------------ colourspace API */

int sws_setColorspaceDetails(SwsContext *sws, const int inv_table[4], int srcRange,
                             const int table[4], int dstRange,
                             int brightness, int contrast, int saturation)
{
    SwsInternal *c = sws_internal(sws);
    int changed;
    if (!c || !inv_table || !table)
        return AVERROR(EINVAL);

    /* formats that are neither YUV nor gray carry no range (utils.c:844-880) */
    if (is_rgb(sws->dst_format))
        dstRange = 0;
    if (is_rgb(sws->src_format))
        srcRange = 0;

    changed = !c->colorspace_set || sws->src_range != srcRange || sws->dst_range != dstRange ||
              c->brightness != brightness || c->contrast != contrast || c->saturation != saturation ||
              memcmp(c->src_colorspace, inv_table, sizeof(int) * 4) ||
              memcmp(c->dst_colorspace, table, sizeof(int) * 4);

    memmove(c->src_colorspace, inv_table, sizeof(int) * 4);
    memmove(c->dst_colorspace, table, sizeof(int) * 4);
    c->brightness = brightness;
    c->contrast   = contrast;
    c->saturation = saturation;
    sws->src_range = srcRange;
    sws->dst_range = dstRange;
    c->colorspace_set = 1;

    if (c->cascade[c->cascade_main])        /* utils.c:909-910 */
        return sws_setColorspaceDetails(c->cascade[c->cascade_main], inv_table, srcRange, table, dstRange,
                                        brightness, contrast, saturation);
    if (!changed || !c->initialized)
        return 0;

    if (!is_rgb(sws->dst_format) && !is_rgb(sws->src_format) &&
        memcmp(c->dst_colorspace, c->src_colorspace, sizeof(int) * 4)) {
        /* the reference cascades through an RGB intermediate here (utils.c:915-984) */
        int ret = cascade_matrix_change(sws, inv_table, srcRange, table, dstRange, brightness, contrast, saturation);
        if (ret < 0) {
            release_cascade(c);
            c->refused = 1;         /* the plan no longer describes what the reference would compute: scaling fails */
            return -1;
        }
        c->refused = 0;
        return 0;
    }
    c->refused = 0;
    release_lanes(c);               /* lanes carry copies of the plan: rebuilt on the next batch call */
    plan_range_convert(c);
    if (plan_colorspace(c) < 0)
        return -1;
    return ff_b200_cuda_update_plan(c->cuda, &c->plan) < 0 ? -1 : 0;
}

int sws_getColorspaceDetails(SwsContext *sws, int **inv_table, int *srcRange,
                             int **table, int *dstRange,
                             int *brightness, int *contrast, int *saturation)
{
    SwsInternal *c = sws_internal(sws);
    if (!c)
        return -1;
    *inv_table  = c->src_colorspace;
    *table      = c->dst_colorspace;
    *srcRange   = is_rgb(sws->src_format) ? 1 : sws->src_range;
    *dstRange   = is_rgb(sws->dst_format) ? 1 : sws->dst_range;
    *brightness = c->brightness;
    *contrast   = c->contrast;
    *saturation = c->saturation;
    return 0;
}

/* ------------------------------------------------------------ init */

static int dst_kind_of(int fmt, const SwsPixDesc *d)
{
    switch (fmt) {
    case AV_PIX_FMT_RGB24:   return SWSC_DST_RGB24;
    case AV_PIX_FMT_BGR24:   return SWSC_DST_BGR24;
    case AV_PIX_FMT_RGBA:    return SWSC_DST_RGBA;
    case AV_PIX_FMT_BGRA:    return SWSC_DST_BGRA;
    case AV_PIX_FMT_ARGB:    return SWSC_DST_ARGB;
    case AV_PIX_FMT_ABGR:    return SWSC_DST_ABGR;
    case AV_PIX_FMT_RGB48LE: return SWSC_DST_RGB48;
    case AV_PIX_FMT_BGR48LE: return SWSC_DST_BGR48;
    case AV_PIX_FMT_RGB565LE: return SWSC_DST_RGB565;
    case AV_PIX_FMT_BGR565LE: return SWSC_DST_BGR565;
    case AV_PIX_FMT_RGB555LE: return SWSC_DST_RGB555;
    case AV_PIX_FMT_BGR555LE: return SWSC_DST_BGR555;
    case AV_PIX_FMT_NV12:    return SWSC_DST_NV12;
    case AV_PIX_FMT_NV21:    return SWSC_DST_NV21;
    case AV_PIX_FMT_P010LE:  return SWSC_DST_P010;
    }
    if ((d->flags & SWSPF_PLANAR) && (d->flags & SWSPF_RGB))
        return SWSC_DST_GBRP;
    if (d->depth == 32)
        return SWSC_DST_PLANARF32;
    if (d->flags & SWSPF_PLANAR)
        return d->depth == 8 ? SWSC_DST_PLANAR8 : d->depth == 16 ? SWSC_DST_PLANAR16 : SWSC_DST_PLANARN;
    return -1;
}

static int is_identity_bank(const SwsFirBank *b, int one)
{
    if (b->size != 1)
        return 0;
    for (int i = 0; i < b->len; i++)
        if (b->coef[i] != one || b->pos[i] != i)
            return 0;
    return 1;
}

/* ff_hyscale_fast_c / ff_hcscale_fast_c (hscale_fast_bilinear.c:23-55) as a 2-tap bank for the 8 -> 15 bit
 * horizontal stage, which computes min((sum src * coef) >> 7, 32767):
 *   luma    dst = (src[xx] << 7) + (src[xx+1] - src[xx]) * xalpha = src[xx] * (128 - xalpha) + src[xx+1] * xalpha
 *   chroma  dst = src[xx] * (xalpha ^ 127) + src[xx+1] * xalpha                       (the taps sum to 127)
 * with xx = xpos >> 16, xalpha = (xpos & 0xFFFF) >> 9, unsigned xpos += xInc; outputs whose
 * (i * xInc) >> 16 reaches the last source sample are src[srcW-1] * 128.  Coefficients are the weights
 * times 128, so the >> 7 is exact and the clip (<= 255 * 128) never triggers. */
static int fast_bilinear_bank(SwsFirBank *b, int dst_len, int src_len, unsigned x_inc, int chroma)
{
    if (src_len < 2)
        return AVERROR(EINVAL);
    ff_b200_free_fir(b);
    b->coef = malloc(sizeof(*b->coef) * (size_t)dst_len * 2);
    b->pos  = malloc(sizeof(*b->pos) * (size_t)dst_len);
    if (!b->coef || !b->pos) {
        ff_b200_free_fir(b);
        return AVERROR(ENOMEM);
    }
    b->size = 2;
    b->len  = dst_len;
    unsigned xpos = 0;
    for (int i = 0; i < dst_len; i++) {
        const unsigned xx = xpos >> 16, xalpha = (xpos & 0xFFFF) >> 9;
        if ((((int64_t)i * x_inc) >> 16) >= src_len - 1) {
            b->pos[i] = src_len - 2;
            b->coef[2 * i] = 0;
            b->coef[2 * i + 1] = 128 * 128;
        } else {
            b->pos[i] = (int)xx;
            b->coef[2 * i] = (int16_t)((chroma ? (xalpha ^ 127) : 128 - xalpha) * 128);
            b->coef[2 * i + 1] = (int16_t)(xalpha * 128);
        }
        xpos += x_inc;
    }
    return 0;
}

/* 1-tap bank: output i reads source sample i >> shift */
static int nearest_bank(SwsFirBank *b, int len, int one, int shift)
{
    ff_b200_free_fir(b);
    b->coef = malloc(sizeof(*b->coef) * (size_t)len);
    b->pos  = malloc(sizeof(*b->pos) * (size_t)len);
    if (!b->coef || !b->pos) {
        ff_b200_free_fir(b);
        return AVERROR(ENOMEM);
    }
    for (int i = 0; i < len; i++) {
        b->coef[i] = (int16_t)one;
        b->pos[i]  = i >> shift;
    }
    b->size = 1;
    b->len  = len;
    return 0;
}

static int init_single(SwsContext *sws, int with_device)
{
    SwsInternal *c = sws_internal(sws);
    SwsCudaPlan *p = &c->plan;
    const int srcW = sws->src_w, srcH = sws->src_h, dstW = sws->dst_w, dstH = sws->dst_h;
    const SwsPixDesc *sd, *dd;
    unsigned flags = sws->flags;
    int scaler, lum_scaler, chr_scaler, unscaled, ret;
    int64_t lum_xinc, lum_yinc, chr_xinc, chr_yinc;
    SwsFirSpec spec;

    if (!c->colorspace_set)
        sws_setColorspaceDetails(sws, sws_getCoefficients(SWS_CS_DEFAULT), sws->src_range,
                                 sws_getCoefficients(SWS_CS_DEFAULT), sws->dst_range, 0, 1 << 16, 1 << 16);

    sd = ff_b200_pix_desc(sws->src_format);
    dd = ff_b200_pix_desc(sws->dst_format);
    if (!sd || !sd->as_input) {
        set_error(c, "pixel format %d is not supported as input", sws->src_format);
        return AVERROR(EINVAL);
    }
    if (!dd || !dd->as_output) {
        set_error(c, "pixel format %d is not supported as output", sws->dst_format);
        return AVERROR(EINVAL);
    }

    /* exactly one scaler; bicubic by default (utils.c:1196-1234) */
    scaler = flags & (SWS_POINT | SWS_AREA | SWS_BILINEAR | SWS_FAST_BILINEAR | SWS_BICUBIC | SWS_X |
                      SWS_GAUSS | SWS_LANCZOS | SWS_SINC | SWS_SPLINE | SWS_BICUBLIN);
    if (!scaler) {
        scaler = SWS_BICUBIC;
        flags |= scaler;
        sws->flags = flags;
    } else if (scaler & (scaler - 1)) {
        set_error(c, "exactly one scaler algorithm must be chosen, got %X", scaler);
        return AVERROR(EINVAL);
    }
    if (scaler == SWS_FAST_BILINEAR && (srcW < 8 || dstW <= 8)) {
        scaler = SWS_BILINEAR;                         /* utils.c:1224-1230 */
        flags ^= SWS_FAST_BILINEAR | SWS_BILINEAR;
        sws->flags = flags;
    }
    lum_scaler = scaler_enum_to_flag(sws->scaler, scaler == SWS_BICUBLIN ? SWS_BICUBIC : scaler);
    chr_scaler = scaler_enum_to_flag(sws->scaler_sub ? sws->scaler_sub : sws->scaler,
                                     scaler == SWS_BICUBLIN ? SWS_BILINEAR : scaler);

    if (srcW < 1 || srcH < 1 || dstW < 1 || dstH < 1) {
        set_error(c, "%dx%d -> %dx%d is invalid scaling dimension", srcW, srcH, dstW, dstH);
        return AVERROR(EINVAL);
    }
    if (sws->gamma_flag || (flags & SWS_SRC_V_CHR_DROP_MASK)) {
        set_error(c, "gamma / chroma-drop are outside the CUDA hot path");
        return AVERROR(ENOTSUP);
    }
    /* Error diffusion (utils.c:1288-1291).  The diffusing writers exist only for the 1- to 8-bit-per-pixel
     * destinations (output.c:692-830,2085-2158: monoblack/white, rgb8, bgr8, rgb4_byte, bgr4_byte), none of which
     * is on this path; for every other destination the setting changes path selection only (the unscaled
     * yuv2rgb LUT converter steps aside, swscale_unscaled.c:2428), which init_single() reproduces below. */
    if (sws->dither == SWS_DITHER_AUTO && (flags & SWS_ERROR_DIFFUSION))
        sws->dither = SWS_DITHER_ED;

    unscaled = srcW == dstW && srcH == dstH;
    {
        /* SwsFilter vectors longer than one tap keep the unscaled special converters away (utils.c:1256-1263,1624) */
        const SwsFilter *sf = c->src_filter_tmp, *df = c->dst_filter_tmp;
        const SwsVector *v[8] = { sf ? sf->lumH : NULL, sf ? sf->chrH : NULL, df ? df->lumH : NULL, df ? df->chrH : NULL,
                                  sf ? sf->lumV : NULL, sf ? sf->chrV : NULL, df ? df->lumV : NULL, df ? df->chrV : NULL };
        for (int i = 0; i < 8; i++)
            if (v[i] && v[i]->length > 1)
                unscaled = 0;
    }
    lum_xinc = (((int64_t)srcW << 16) + (dstW >> 1)) / dstW;
    lum_yinc = (((int64_t)srcH << 16) + (dstH >> 1)) / dstH;

    c->chr_src_hsub = sd->log2_cw; c->chr_src_vsub = sd->log2_ch;
    c->chr_dst_hsub = dd->log2_cw; c->chr_dst_vsub = dd->log2_ch;
    c->dst_slice_align = 1 << c->chr_dst_vsub;

    /* packed RGB shares one chroma pair between two pixels unless forced (utils.c:1270-1360) */
    if (is_rgb(sws->dst_format) && !(flags & SWS_FULL_CHR_H_INT)) {
        if (dstW & 1)
            flags |= SWS_FULL_CHR_H_INT;
        if (c->chr_src_hsub == 0 && c->chr_src_vsub == 0 && sws->dither != SWS_DITHER_BAYER &&
            !(flags & SWS_FAST_BILINEAR))                      /* utils.c:1278-1285 */
            flags |= SWS_FULL_CHR_H_INT;
        sws->flags = flags;
    }
    if (dd->flags & SWSPF_PLANAR && is_rgb(sws->dst_format) && !(flags & SWS_FULL_CHR_H_INT)) {
        flags |= SWS_FULL_CHR_H_INT;                   /* planar RGB has no half-chroma writer (utils.c:1317-1325) */
        sws->flags = flags;
    }
    if (is_rgb(sws->dst_format) && dd->bpp <= 16) {
        /* 15/16 bpp destinations have no full-chroma writer: the reference drops the flag again
         * (utils.c:1329-1357).  Its pair writer then stores one pixel past an odd width (into the row
         * padding); the kernel stops at the last valid pixel */
        flags &= ~SWS_FULL_CHR_H_INT;
        sws->flags = flags;
    }
    if (is_rgb(sws->dst_format) && !(flags & SWS_FULL_CHR_H_INT))
        c->chr_dst_hsub = 1;
    /* packed RGB sources: chroma is read from horizontally summed pixel pairs unless the caller
     * asks for full chroma input or the output needs the resolution (utils.c:1367-1390) */
    if (is_rgb(sws->src_format) && !(srcW & 1) && !(flags & SWS_FULL_CHR_H_INP) &&
        ((dstW >> c->chr_dst_hsub) <= (srcW >> 1) || (flags & SWS_FAST_BILINEAR)))
        c->chr_src_hsub = 1;

    c->chr_src_w = ceil_rshift(srcW, c->chr_src_hsub);
    c->chr_src_h = ceil_rshift(srcH, c->chr_src_vsub);
    c->chr_dst_w = ceil_rshift(dstW, c->chr_dst_hsub);
    c->chr_dst_h = ceil_rshift(dstH, c->chr_dst_vsub);
    c->src_bpc = sd->depth < 8 ? 8 : sd->depth;
    c->dst_bpc = dd->depth < 8 ? 8 : dd->depth;
    if (is_rgb(sws->src_format))
        c->src_bpc = 16;          /* the RGB readers emit 14-bit samples in 16-bit lines (utils.c:1407-1408) */

    chr_xinc = (((int64_t)c->chr_src_w << 16) + (c->chr_dst_w >> 1)) / c->chr_dst_w;
    chr_yinc = (((int64_t)c->chr_src_h << 16) + (c->chr_dst_h >> 1)) / c->chr_dst_h;
    if (chr_xinc < 10 || chr_xinc > INT32_MAX || chr_yinc < 10 || chr_yinc > INT32_MAX ||
        lum_xinc < 10 || lum_xinc > INT32_MAX || lum_yinc < 10 || lum_yinc > INT32_MAX)
        return AVERROR_PATCHWELCOME;

    /* which special converter would the reference install? (utils.c:1624-1637,
     * swscale_unscaled.c:2392-2731) */
    c->unscaled_lut = 0;
    c->special = SWSC_SPECIAL_NONE;
    if (unscaled && (sws->src_range == sws->dst_range || is_rgb(sws->dst_format))) {
        /* the plane-copy wrappers need isFloat(src) == isFloat(dst) (swscale_unscaled.c:2656-2663): float destinations
         * always go through the scaler */
        const int planar_yuv_pair = !is_rgb(sws->src_format) && !is_rgb(sws->dst_format) && dd->depth <= 16;
        if (is_rgb(sws->src_format) && is_rgb(sws->dst_format) && sd->depth == 8 && dd->depth == 8) {
            /* rgbToRgbWrapper (findRgbConvFn, swscale_unscaled.c:1843-2060,2463-2466) and
             * packedCopyWrapper for identical formats (:2689-2707).  With SWS_BITEXACT the reference
             * refuses 24 -> rgba/bgra and lets the scaler do it (:1992-1996). */
            const int s32 = sd->bpp == 32, d32 = dd->bpp == 32;
            if (sws->src_format == sws->dst_format || s32 || !d32 || !(flags & SWS_BITEXACT) ||
                sws->dst_format == AV_PIX_FMT_ARGB || sws->dst_format == AV_PIX_FMT_ABGR)
                c->special = SWSC_SPECIAL_SHUFFLE;
        }
        if (is_rgb(sws->src_format) && is_rgb(sws->dst_format) && sd->depth == 16 && dd->depth == 16 && dd->bpp == 48)
            /* identical formats: packedCopyWrapper; rgb48le <-> bgr48le: rgb48tobgr48_nobswap through rgbToRgbWrapper
             * (findRgbConvFn, swscale_unscaled.c:1869-1873; no dither is ever needed between 48-bit layouts) */
            c->special = SWSC_SPECIAL_RGB48;
        if (is_rgb(sws->src_format) && is_rgb(sws->dst_format) && sd->depth == 8 && dd->bpp <= 16 &&
            (flags & (SWS_FAST_BILINEAR | SWS_POINT)))
            /* rgb24to16, rgb32tobgr15, ... (findRgbConvFn): plain truncation, chosen only when the caller's
             * scaler says no dither is wanted (swscale_unscaled.c:2459-2466); otherwise the dithering scaler runs */
            c->special = SWSC_SPECIAL_RGB16PACK;
        if (sws->src_format == AV_PIX_FMT_BGR24 && sws->dst_format == AV_PIX_FMT_YUV420P &&
            !(flags & SWS_ACCURATE_RND) && !(dstW & 1))
            c->special = SWSC_SPECIAL_BGR24_YV12;     /* swscale_unscaled.c:2062-2077,2453-2457 */
        if (planar_yuv_pair && sd->depth == 8 && dd->depth == 8) {
            /* planarToNv12Wrapper, nv12ToPlanarWrapper and the plain plane copy (swscale_unscaled.c:147-215,
             * 2405-2420,2675-2700): chosen at init and, like every convert_unscaled hook, kept even if
             * sws_setColorspaceDetails() later makes the ranges differ (utils.c:849-905) */
            const int ssemi = !!(sd->flags & SWSPF_SEMI), dsemi = !!(dd->flags & SWSPF_SEMI);
            const int s420 = sd->log2_cw == 1 && sd->log2_ch == 1, d420 = dd->log2_cw == 1 && dd->log2_ch == 1;
            if ((s420 && d420 && ssemi != dsemi) ||
                (sd->log2_cw == dd->log2_cw && sd->log2_ch == dd->log2_ch && ssemi == dsemi &&
                 sd->swap_uv == dd->swap_uv))
                c->special = SWSC_SPECIAL_COPY8;
        }
        if ((sws->src_format == AV_PIX_FMT_YUV420P || sws->src_format == AV_PIX_FMT_YUV422P) &&
            is_rgb(sws->dst_format) && !(dd->flags & SWSPF_PLANAR) && !(flags & SWS_ACCURATE_RND) &&
            (sws->dither == SWS_DITHER_BAYER || sws->dither == SWS_DITHER_AUTO) && !(dstH & 1)) {
            c->unscaled_lut = 1;        /* yuv2rgb_c_* nearest-chroma LUT converter (a13) */
            c->dst_slice_align = 2;
            /* one chroma sample per pixel pair whatever the flags say (an odd width forced SWS_FULL_CHR_H_INT
             * above; the converter ignores it and leaves the last pixel of the row untouched) */
            c->chr_dst_hsub = 1;
            c->chr_dst_w = ceil_rshift(dstW, 1);
        } else if (sws->dst_format == AV_PIX_FMT_P010LE && !(sd->flags & SWSPF_SEMI) && planar_yuv_pair &&
                   sd->log2_cw == 1 && sd->log2_ch == 1 && sd->depth != 9) {
            /* planar8ToP01xleWrapper (8-bit: << 8) and planarToP01xWrapper (10/12/14/16-bit: << 16 - depth),
             * swscale_unscaled.c:273-375,2432-2444 */
            c->special = SWSC_SPECIAL_P01X;
        } else if (planar_yuv_pair && (sd->depth != dd->depth || sd->depth > 8) && (sd->flags & SWSPF_SEMI) &&
                   (dd->flags & SWSPF_SEMI) && sd->swap_uv == dd->swap_uv) {
            /* nv12 -> p010: COPY816 (swscale_unscaled.c:2266-2284); p010 -> nv12: DITHER_COPY, whose scalar tail drops the
             * source shift for the last width & 7 samples of a row (:2174-2176) -- reproduced by the kernel */
            c->special = SWSC_SPECIAL_DEPTHCOPY;
        } else if (planar_yuv_pair && c->chr_src_hsub == c->chr_dst_hsub &&
                   c->chr_src_vsub == c->chr_dst_vsub && (sd->depth != dd->depth || sd->depth > 8) &&
                   !(sd->flags & SWSPF_SEMI) && !(dd->flags & SWSPF_SEMI)) {
            /* planarCopyWrapper between depths (swscale_unscaled.c:2220-2384): ordered dither down,
             * bit replication (full-range luma) or plain shift up.  Equal depths above 8 bits are the same
             * wrapper with a zero shift: a copy that, like every convert_unscaled hook, stays in place when
             * sws_setColorspaceDetails() later makes the ranges differ */
            c->special = SWSC_SPECIAL_DEPTHCOPY;
            if (sws->dither != SWS_DITHER_NONE)
                c->dst_slice_align = 8 << c->chr_dst_vsub;      /* :2694-2696: the dither rows count from the slice */
        }
    }
    /* an alpha channel on both sides travels through the scaler as a fourth line set (needAlpha, utils.c:1427) */
    const int need_alpha = !c->special && is_rgb(sws->src_format) && sd->bpp == 32 && is_rgb(sws->dst_format) &&
                           dd->bpp == 32 && !(dd->flags & SWSPF_PLANAR);
    /* ---- FIR banks: horizontal 1<<14, vertical 1<<12 (utils.c:1681-1729) ---- */
    memset(&spec, 0, sizeof(spec));
    spec.flags = flags;
    spec.param[0] = sws->scaler_params[0];
    spec.param[1] = sws->scaler_params[1];

    spec.one = 1 << 14;
    spec.inc = lum_xinc; spec.src_len = srcW; spec.dst_len = dstW; spec.scaler = lum_scaler;
    spec.src_pos = local_chroma_pos(0, 0); spec.dst_pos = local_chroma_pos(0, 0);
    spec.src_vec = c->src_filter_tmp && c->src_filter_tmp->lumH ? c->src_filter_tmp->lumH->coeff : NULL;
    spec.src_vec_len = spec.src_vec ? c->src_filter_tmp->lumH->length : 0;
    spec.dst_vec_len = c->dst_filter_tmp && c->dst_filter_tmp->lumH ? c->dst_filter_tmp->lumH->length : 0;
    if ((ret = ff_b200_build_fir(&c->h_lum, &spec)) < 0)
        goto fir_fail;
    spec.inc = chr_xinc; spec.src_len = c->chr_src_w; spec.dst_len = c->chr_dst_w; spec.scaler = chr_scaler;
    spec.src_pos = local_chroma_pos(c->chr_src_hsub, sws->src_h_chr_pos);
    spec.dst_pos = local_chroma_pos(c->chr_dst_hsub, sws->dst_h_chr_pos);
    spec.src_vec = c->src_filter_tmp && c->src_filter_tmp->chrH ? c->src_filter_tmp->chrH->coeff : NULL;
    spec.src_vec_len = spec.src_vec ? c->src_filter_tmp->chrH->length : 0;
    spec.dst_vec_len = c->dst_filter_tmp && c->dst_filter_tmp->chrH ? c->dst_filter_tmp->chrH->length : 0;
    if ((ret = ff_b200_build_fir(&c->h_chr, &spec)) < 0)
        goto fir_fail;

    spec.one = 1 << 12;
    spec.inc = lum_yinc; spec.src_len = srcH; spec.dst_len = dstH; spec.scaler = lum_scaler;
    spec.src_pos = local_chroma_pos(0, 0); spec.dst_pos = local_chroma_pos(0, 0);
    spec.src_vec = c->src_filter_tmp && c->src_filter_tmp->lumV ? c->src_filter_tmp->lumV->coeff : NULL;
    spec.src_vec_len = spec.src_vec ? c->src_filter_tmp->lumV->length : 0;
    spec.dst_vec_len = c->dst_filter_tmp && c->dst_filter_tmp->lumV ? c->dst_filter_tmp->lumV->length : 0;
    if ((ret = ff_b200_build_fir(&c->v_lum, &spec)) < 0)
        goto fir_fail;
    spec.inc = chr_yinc; spec.src_len = c->chr_src_h; spec.dst_len = c->chr_dst_h; spec.scaler = chr_scaler;
    spec.src_pos = local_chroma_pos(c->chr_src_vsub, sws->src_v_chr_pos);
    spec.dst_pos = local_chroma_pos(c->chr_dst_vsub, sws->dst_v_chr_pos);
    spec.src_vec = c->src_filter_tmp && c->src_filter_tmp->chrV ? c->src_filter_tmp->chrV->coeff : NULL;
    spec.src_vec_len = spec.src_vec ? c->src_filter_tmp->chrV->length : 0;
    spec.dst_vec_len = c->dst_filter_tmp && c->dst_filter_tmp->chrV ? c->dst_filter_tmp->chrV->length : 0;
    if ((ret = ff_b200_build_fir(&c->v_chr, &spec)) < 0)
        goto fir_fail;

    if ((flags & SWS_FAST_BILINEAR) && c->src_bpc == 8 && c->dst_bpc <= 14) {
        /* ff_hyscale_fast_c / ff_hcscale_fast_c replace the horizontal FIR (swscale.c:675-681,
         * hscale_fast_bilinear.c:23-55): the same numbers as a 2-tap bank */
        /* a luma step of exactly 1.0 makes xalpha 0 everywhere: the identity bank already says that.  Not so
         * for chroma, whose left weight is xalpha ^ 127 = 127 (the samples come out as 127/128 of the source) */
        if (lum_xinc != 0x10000 && (ret = fast_bilinear_bank(&c->h_lum, dstW, srcW, (unsigned)lum_xinc, 0)) < 0)
            goto fir_fail;
        if ((ret = fast_bilinear_bank(&c->h_chr, c->chr_dst_w, c->chr_src_w, (unsigned)chr_xinc, 1)) < 0)
            goto fir_fail;
    }
    if (c->unscaled_lut) {
        /* a13 samples chroma at column x>>1, row y>>vsub with no filtering and ignores
         * chroma siting (yuv2rgb.c:137-236): express that as 1-tap banks so the same kernels
         * serve it.  Arithmetic is identical: ((u<<7)*4096 + 2^18) >> 19 == u. */
        if ((ret = nearest_bank(&c->h_chr, c->chr_dst_w, 1 << 14, 0)) < 0 ||
            (ret = nearest_bank(&c->v_chr, c->chr_dst_h, 1 << 12, c->chr_src_vsub)) < 0)
            goto fir_fail;
    }

    /* ---- device plan ---- */
    memset(p, 0, sizeof(*p));
    p->src_w = srcW; p->src_h = srcH; p->dst_w = dstW; p->dst_h = dstH;
    p->chr_src_w = c->chr_src_w; p->chr_src_h = c->chr_src_h;
    p->chr_dst_w = c->chr_dst_w; p->chr_dst_h = c->chr_dst_h;
    p->chr_src_hsub = c->chr_src_hsub; p->chr_src_vsub = c->chr_src_vsub;
    p->chr_dst_hsub = c->chr_dst_hsub; p->chr_dst_vsub = c->chr_dst_vsub;
    p->src_layout = (sd->flags & SWSPF_SEMI) ? (sd->swap_uv ? SWSC_SRC_NV21 : SWSC_SRC_NV12) : SWSC_SRC_PLANAR;
    p->dst_kind = dst_kind_of(sws->dst_format, dd);
    p->src_bits = c->src_bpc;
    p->src_shift = sd->shift;
    p->dst_shift = dd->shift;
    p->dst_bits = c->dst_bpc;
    p->has_chroma = !(sd->flags & SWSPF_GRAY) && !(dd->flags & SWSPF_GRAY);   /* a gray side: luma only */
    p->dst_has_chroma = !(dd->flags & SWSPF_GRAY);
    p->unscaled_lut = c->unscaled_lut;
    p->special = c->special;
    p->full_chr = is_rgb(sws->dst_format) && (flags & SWS_FULL_CHR_H_INT) && !c->unscaled_lut;
    /* hScale selection (swscale.c:675-688) and its shift (swscale.c:69-159) */
    p->inter_bits = c->dst_bpc > 14 ? 19 : 15;
    if (is_rgb(sws->src_format) && sd->depth < 16)
        p->h_shift = p->inter_bits == 15 ? 13 : 9;         /* swscale.c:80-81,108-109 */
    else if (c->src_bpc == 8)
        p->h_shift = p->inter_bits == 15 ? 7 : 3;
    else
        p->h_shift = p->inter_bits == 15 ? c->src_bpc - 1 : c->src_bpc - 1 - 4;
    /* 8-bit planar output of >8-bit sources is dithered (swscale.c:291-292,385-387,519-522);
     * packed 8-bit RGB is neither isNBPS nor is16BPS, so it is not */
    p->dither_bayer = sd->depth > 8;
    if (is_rgb(sws->src_format)) {
        static const struct { int fmt, bpp, r, g, b; } order[] = {
            { AV_PIX_FMT_RGB24, 3, 0, 1, 2 }, { AV_PIX_FMT_BGR24, 3, 2, 1, 0 },
            { AV_PIX_FMT_RGBA,  4, 0, 1, 2 }, { AV_PIX_FMT_BGRA,  4, 2, 1, 0 },
            { AV_PIX_FMT_ARGB,  4, 1, 2, 3 }, { AV_PIX_FMT_ABGR,  4, 3, 2, 1 },
            /* 16 bits per component: indices of the components, not bytes (rgb48ToY_c & co., input.c:111-196) */
            { AV_PIX_FMT_RGB48LE, 6, 0, 1, 2 }, { AV_PIX_FMT_BGR48LE, 6, 2, 1, 0 },
        };
        p->src_layout = SWSC_SRC_RGB;
        p->src_bits = sd->depth == 16 ? 16 : 8;
        for (size_t i = 0; i < sizeof(order) / sizeof(order[0]); i++)
            if (order[i].fmt == sws->src_format) {
                p->src_bpp = order[i].bpp;
                p->src_ro = order[i].r; p->src_go = order[i].g; p->src_bo = order[i].b;
            }
        if (!p->src_bpp) {
            set_error(c, "RGB source format %d is not on the CUDA hot path", sws->src_format);
            return AVERROR(ENOTSUP);
        }
        p->src_rgb_half = c->chr_src_hsub;
        if (need_alpha) {
            p->src_alpha = p->dst_alpha = 1;
            p->src_ao = 6 - p->src_ro - p->src_go - p->src_bo;
        }
        if (c->special == SWSC_SPECIAL_SHUFFLE) {
            /* destination byte k <- source byte of the same component; a missing alpha becomes 255 */
            int so[4] = { -1, -1, -1, -1 }, dorder[4] = { 4, 4, 4, 4 };   /* component (r,g,b,a) -> byte */
            so[0] = p->src_ro; so[1] = p->src_go; so[2] = p->src_bo;
            if (p->src_bpp == 4)
                so[3] = 6 - p->src_ro - p->src_go - p->src_bo;
            p->dst_bpp = dd->bpp / 8;
            for (size_t i = 0; i < sizeof(order) / sizeof(order[0]); i++)
                if (order[i].fmt == sws->dst_format) {
                    dorder[order[i].r] = 0; dorder[order[i].g] = 1; dorder[order[i].b] = 2;
                    if (order[i].bpp == 4)
                        dorder[6 - order[i].r - order[i].g - order[i].b] = 3;
                }
            for (int k = 0; k < 4; k++)
                p->shuf_map[k] = dorder[k] < 4 && so[dorder[k]] >= 0 ? so[dorder[k]] : 4;
        }
    }
    plan_range_convert(c);
    if ((ret = plan_colorspace(c)) < 0) {
        set_error(c, "colourspace constants overflow the int32 kernel arithmetic");
        return ret;
    }
    p->lum_identity = is_identity_bank(&c->h_lum, 1 << 14) && is_identity_bank(&c->v_lum, 1 << 12);
    p->chr_h_identity = is_identity_bank(&c->h_chr, 1 << 14);
    p->chr_v_identity = is_identity_bank(&c->v_chr, 1 << 12);

    if (!with_device) {
        c->planned = 1;
        return 0;
    }
    /* One-tap vertical filters run through yuv2plane1 / yuv2packed1, which never look at the coefficient
     * (vscale.c:135-143,296-316).  It is 4096 except where initFilter's edge fix-up left 4095 behind (a few
     * source rows with a large chroma offset): make the device banks say what those writers compute. */
    if (!((dd->flags & SWSPF_PLANAR) && is_rgb(sws->dst_format))) {     /* planar RGB always runs yuv2anyX with the real taps (vscale.c:173-215) */
        const int packed = is_rgb(sws->dst_format);
        if (c->v_lum.size == 1) {
            for (int y = 0; y < c->v_lum.len; y++) {
                int one = !packed || c->v_chr.size == 1;
                if (packed && c->v_chr.size == 2 && y < c->v_chr.len) {
                    const int c0 = c->v_chr.coef[2 * y], c1 = c->v_chr.coef[2 * y + 1];
                    one = c0 + c1 == 4096 && (unsigned)c1 <= 4096u;
                }
                if (one)
                    c->v_lum.coef[y] = 4096;
            }
        }
        if (c->v_chr.size == 1 && (packed ? c->v_lum.size == 1 : !(dd->flags & SWSPF_SEMI)))
            for (int y = 0; y < c->v_chr.len; y++)
                c->v_chr.coef[y] = 4096;
    }
    ret = ff_b200_cuda_create(&c->cuda, p, &c->h_lum, &c->h_chr, &c->v_lum, &c->v_chr);
    if (ret < 0) {
        set_error(c, "CUDA initialisation failed (%d): no CPU fallback exists on this path", ret);
        return ret;
    }
    c->dst_y = 0;
    c->slice_dir = 0;
    c->rows_received = 0;
    c->initialized = 1;
    return 0;

fir_fail:
    if (ret == SWS_B200_USE_CASCADE) {
        if (!with_device) {
            set_error(c, "filter too long: the reference cascades two contexts (utils.c:1803-1832)");
            return AVERROR(ENOTSUP);
        }
        ret = SWS_B200_USE_CASCADE;          /* sws_init_context() builds the two stages */
    } else {
        set_error(c, "filter construction failed (%d)", ret);
    }
    return ret;
}

int sws_init_context(SwsContext *sws, SwsFilter *srcFilter, SwsFilter *dstFilter)
{
    SwsInternal *c = sws_internal(sws);
    int ret;
    if (!c)
        return AVERROR(EINVAL);
    release_tables(c);
    sws->src_range |= fold_jpeg_format(&sws->src_format);
    sws->dst_range |= fold_jpeg_format(&sws->dst_format);
    c->src_filter_tmp = srcFilter;          /* caller-owned, only read while the banks are built */
    c->dst_filter_tmp = dstFilter;
    ret = init_single(sws, 1);
    c->src_filter_tmp = c->dst_filter_tmp = NULL;
    if (ret == SWS_B200_USE_CASCADE) {
        release_tables(c);
        ret = cascade_long_filter(sws, srcFilter, dstFilter);
        if (ret >= 0) {
            c->initialized = 1;
            c->dst_y = 0;
            c->slice_dir = 0;
            return 0;
        }
    }
    if (ret < 0)
        release_tables(c);
    return ret;
}

/* ------------------------------------------------------------ diagnostics
 * Host-side planning without touching a device, so the table builders can be
 * pinned against the reference on machines that have no GPU. */
int sws_b200_plan_only_filtered(SwsContext *sws, SwsFilter *srcFilter, SwsFilter *dstFilter)
{
    SwsInternal *c = sws_internal(sws);
    int ret;
    if (!c)
        return AVERROR(EINVAL);
    release_tables(c);
    sws->src_range |= fold_jpeg_format(&sws->src_format);
    sws->dst_range |= fold_jpeg_format(&sws->dst_format);
    c->src_filter_tmp = srcFilter;
    c->dst_filter_tmp = dstFilter;
    ret = init_single(sws, 0);
    c->src_filter_tmp = c->dst_filter_tmp = NULL;
    if (ret < 0)
        release_tables(c);
    return ret;
}

int sws_b200_plan_only(SwsContext *sws)
{
    return sws_b200_plan_only_filtered(sws, NULL, NULL);
}

int sws_b200_get_filter(SwsContext *sws, int which, const int16_t **coef, const int32_t **pos, int *len)
{
    SwsInternal *c = sws_internal(sws);
    const SwsFirBank *b;
    if (!c || (!c->planned && !c->initialized))
        return AVERROR(EINVAL);
    b = which == 0 ? &c->h_lum : which == 1 ? &c->h_chr : which == 2 ? &c->v_lum : which == 3 ? &c->v_chr : NULL;
    if (!b || !b->coef)
        return AVERROR(EINVAL);
    *coef = b->coef;
    *pos  = b->pos;
    *len  = b->len;
    return b->size;
}

int sws_b200_get_rgb2yuv(SwsContext *sws, int out[9])
{
    SwsInternal *c = sws_internal(sws);
    if (!c || (!c->planned && !c->initialized) || c->plan.src_layout != SWSC_SRC_RGB)
        return AVERROR(EINVAL);
    for (int i = 0; i < 9; i++)
        out[i] = c->plan.rgb2yuv[i];
    return 0;
}

int sws_b200_get_info(SwsContext *sws, int out[32])
{
    SwsInternal *c = sws_internal(sws);
    const SwsCudaPlan *p;
    if (!c || (!c->planned && !c->initialized))
        return AVERROR(EINVAL);
    p = &c->plan;
    memset(out, 0, sizeof(int) * 32);
    out[0] = p->rgb.y_offset; out[1] = p->rgb.y_coeff; out[2] = p->rgb.v2r; out[3] = p->rgb.v2g;
    out[4] = p->rgb.u2g;      out[5] = p->rgb.u2b;     out[6] = c->unscaled_lut;
    out[8] = c->chr_src_w;    out[9] = c->chr_src_h;   out[10] = c->chr_dst_w; out[11] = c->chr_dst_h;
    out[12] = c->src_bpc;     out[13] = c->dst_bpc;    out[14] = (int)sws->flags;
    out[15] = p->rgb.cy;      out[16] = p->rgb.yb;     out[17] = p->rgb.crv;   out[18] = p->rgb.cbu;
    out[19] = p->rgb.cgu;     out[20] = p->rgb.cgv;    out[21] = p->rgb.base_r; out[22] = p->rgb.base_g;
    out[23] = p->rgb.base_b;  out[24] = p->range_mode; out[25] = (int)p->lum_rc_coeff;
    out[26] = (int)p->chr_rc_coeff; out[27] = p->lum_identity; out[28] = p->chr_h_identity;
    out[29] = p->h_shift;     out[30] = p->inter_bits; out[31] = p->dst_kind;
    return 0;
}

SwsContext *sws_getContext(int srcW, int srcH, enum AVPixelFormat srcFormat,
                           int dstW, int dstH, enum AVPixelFormat dstFormat,
                           int flags, SwsFilter *srcFilter,
                           SwsFilter *dstFilter, const double *param)
{
    SwsContext *sws = sws_alloc_context();
    if (!sws)
        return NULL;
    sws->flags = (unsigned)flags;
    sws->src_w = srcW; sws->src_h = srcH; sws->src_format = srcFormat;
    sws->dst_w = dstW; sws->dst_h = dstH; sws->dst_format = dstFormat;
    for (int i = 0; param && i < SWS_NUM_SCALER_PARAMS; i++)
        sws->scaler_params[i] = param[i];
    if (sws_init_context(sws, srcFilter, dstFilter) < 0) {
        if (flags & SWS_PRINT_INFO)
            fprintf(stderr, "[swscaler-b200] sws_getContext failed: %s\n", sws_internal(sws)->last_error);
        sws_freeContext(sws);
        return NULL;
    }
    return sws;
}

SwsContext *sws_getCachedContext(SwsContext *prev, int srcW, int srcH,
                                 enum AVPixelFormat srcFormat, int dstW, int dstH,
                                 enum AVPixelFormat dstFormat, int flags,
                                 SwsFilter *srcFilter, SwsFilter *dstFilter,
                                 const double *param)
{
    static const double default_param[SWS_NUM_SCALER_PARAMS] = { SWS_PARAM_DEFAULT, SWS_PARAM_DEFAULT };
    if (!param)
        param = default_param;
    if (prev && (prev->src_w != srcW || prev->src_h != srcH || prev->src_format != (int)srcFormat ||
                 prev->dst_w != dstW || prev->dst_h != dstH || prev->dst_format != (int)dstFormat ||
                 prev->flags != (unsigned)flags || prev->scaler_params[0] != param[0] ||
                 prev->scaler_params[1] != param[1])) {
        sws_freeContext(prev);
        prev = NULL;
    }
    if (!prev)
        return sws_getContext(srcW, srcH, srcFormat, dstW, dstH, dstFormat, flags,
                              srcFilter, dstFilter, param);
    return prev;
}

/* ------------------------------------------------------------ sws_scale */

/* last source rows (luma, chroma) output row y needs; mirrors swscale.c:412-425 */
static void rows_needed(const SwsInternal *c, int y, int *last_lum, int *last_chr)
{
    const int vsub_mask = (1 << c->chr_dst_vsub) - 1;
    int y2 = y | vsub_mask;
    int first_l, first_c;
    if (y2 > c->opts.dst_h - 1)
        y2 = c->opts.dst_h - 1;
    first_l = c->v_lum.pos[y2];
    if (first_l < 1 - c->v_lum.size)
        first_l = 1 - c->v_lum.size;
    first_c = c->v_chr.pos[y >> c->chr_dst_vsub];
    if (first_c < 1 - c->v_chr.size)
        first_c = 1 - c->v_chr.size;
    *last_lum = first_l + c->v_lum.size - 1;
    if (*last_lum > c->opts.src_h - 1)
        *last_lum = c->opts.src_h - 1;
    *last_chr = first_c + c->v_chr.size - 1;
    if (*last_chr > c->chr_src_h - 1)
        *last_chr = c->chr_src_h - 1;
}

/* one slice through one (non-cascaded) context; mem: SWS_MEM_* for device-resident sides */
static int scale_slice(SwsContext *sws, const uint8_t *const srcSlice[], const int srcStride[],
                       int srcSliceY, int srcSliceH, uint8_t *const dst[], const int dstStride[], int mem)
{
    SwsInternal *c = sws_internal(sws);
    int macro_src, y0, y1, ret, avail_l, avail_c;
    const uint8_t *src2[4];
    uint8_t *dst2[4];
    int sstride[4], dstride[4];

    if (!c || !c->initialized)
        return AVERROR(EINVAL);
    if (!srcStride || !dstStride || !dst || !srcSlice) {
        set_error(c, "one of the input parameters to sws_scale() is NULL");
        return AVERROR(EINVAL);
    }
    if (c->refused) {
        set_error(c, "the last sws_setColorspaceDetails() asked for a conversion that is not on the CUDA hot path");
        return AVERROR(ENOTSUP);
    }
    if ((c->src_bpc > 8 && (!is_rgb(sws->src_format) || ff_b200_pix_desc(sws->src_format)->depth == 16) &&
         ((srcStride[0] | srcStride[1] | srcStride[2]) & 1)) ||
        (c->dst_bpc > 8 && ((dstStride[0] | dstStride[1] | dstStride[2]) & 1))) {
        set_error(c, "16-bit samples need even strides");
        return AVERROR(EINVAL);
    }
    if (c->dst_bpc == 32 && ((dstStride[0] | dstStride[1] | dstStride[2]) & 3)) {
        set_error(c, "32-bit float samples need strides that are multiples of 4");
        return AVERROR(EINVAL);
    }
    macro_src = 1 << c->chr_src_vsub;
    if ((srcSliceY & (macro_src - 1)) ||
        ((srcSliceH & (macro_src - 1)) && srcSliceY + srcSliceH != sws->src_h) ||
        srcSliceY + srcSliceH > sws->src_h || srcSliceY < 0 || srcSliceH < 0) {
        set_error(c, "slice parameters %d, %d are invalid", srcSliceY, srcSliceH);
        return AVERROR(EINVAL);
    }
    if (!srcSlice[0] || !dst[0]) {
        set_error(c, "bad image pointers");
        return AVERROR(EINVAL);
    }
    if (srcSliceH == 0)
        return 0;

    if (!c->slice_dir) {
        if (srcSliceY != 0 && srcSliceY + srcSliceH != sws->src_h) {
            set_error(c, "slices start in the middle");
            return AVERROR(EINVAL);
        }
        c->slice_dir = srcSliceY == 0 ? 1 : -1;
        c->dst_y = 0;
    }

    for (int i = 0; i < 4; i++) {
        src2[i] = srcSlice[i]; dst2[i] = dst[i];
        sstride[i] = srcStride[i]; dstride[i] = dstStride[i];
    }
    if (c->slice_dir != 1) {
        /* slices arrive bottom to top: flip the picture internally, exactly like the reference
         * (swscale.c:1141-1159) -- every stride negated, pointers moved to the last row, the slice
         * position mirrored -- and convert it top-down */
        const int csh = (srcSliceH >> c->chr_src_vsub) - 1;
        for (int i = 0; i < 4; i++) {
            sstride[i] = -sstride[i];
            dstride[i] = -dstride[i];
        }
        if (src2[0]) src2[0] += (ptrdiff_t)(srcSliceH - 1) * srcStride[0];
        if (src2[1]) src2[1] += (ptrdiff_t)csh * srcStride[1];
        if (src2[2]) src2[2] += (ptrdiff_t)csh * srcStride[2];
        if (src2[3]) src2[3] += (ptrdiff_t)(srcSliceH - 1) * srcStride[3];
        if (dst2[0]) dst2[0] += (ptrdiff_t)(sws->dst_h - 1) * dstStride[0];
        if (dst2[1]) dst2[1] += (ptrdiff_t)((sws->dst_h >> c->chr_dst_vsub) - 1) * dstStride[1];
        if (dst2[2]) dst2[2] += (ptrdiff_t)((sws->dst_h >> c->chr_dst_vsub) - 1) * dstStride[2];
        if (dst2[3]) dst2[3] += (ptrdiff_t)(sws->dst_h - 1) * dstStride[3];
        srcSliceY = sws->src_h - srcSliceY - srcSliceH;
        /* with an odd dst_h the reference stores the last chroma row one row BEFORE the plane (a heap underflow the
         * fuzz ran into): that row is dropped here instead */
        mem |= SWS_MEM_DST_FLIPPED;
    }
    if (srcSliceY == 0)
        c->dst_y = 0;

    /* how far can the output advance with the rows received so far?  Same rule as the
     * reference's enough_lines test (swscale.c:462-470). */
    avail_l = srcSliceY + srcSliceH;
    avail_c = ceil_rshift(srcSliceY + srcSliceH, c->chr_src_vsub);
    y0 = c->dst_y;
    if (c->unscaled_lut || c->special) {
        /* the unscaled converters map slice rows 1:1 (swscale.c:1161-1187) */
        y0 = srcSliceY;
        y1 = srcSliceY + srcSliceH;
    } else {
        for (y1 = y0; y1 < sws->dst_h; y1++) {
            int ll, lc;
            rows_needed(c, y1, &ll, &lc);
            if (!(ll < avail_l && lc < avail_c))
                break;
        }
    }

    ret = ff_b200_cuda_scale_host(c->cuda, src2, sstride, srcSliceY, srcSliceH, 1,
                                  dst2, dstride, y0, y1, mem);
    if (ret < 0) {
        set_error(c, "CUDA conversion failed (%d)", ret);
        c->slice_dir = 0;
        return ret;
    }
    c->dst_y = y1;
    if (srcSliceY + srcSliceH == sws->src_h)
        c->slice_dir = 0;
    return y1 - y0;
}

/* scale_cascaded() (swscale.c:992-1020): the first stage assembles the intermediate picture slice by slice; the
 * second runs once the whole source has been consumed.  The intermediate never leaves the device. */
static int scale_cascaded(SwsInternal *c, const uint8_t *const srcSlice[], const int srcStride[],
                          int srcSliceY, int srcSliceH, uint8_t *const dst[], const int dstStride[])
{
    SwsInternal *c0 = sws_internal(c->cascade[0]);
    int ret = scale_slice(c->cascade[0], srcSlice, srcStride, srcSliceY, srcSliceH, c->cascade_tmp,
                          c->cascade_tmp_stride, SWS_MEM_DST_DEVICE);
    if (ret < 0) {
        set_error(c, "cascade stage 0: %s", c0->last_error);
        return ret;
    }
    if (c0->slice_dir != 0)
        return 0;                       /* intermediate incomplete: no output lines yet */
    ret = scale_slice(c->cascade[1], (const uint8_t *const *)c->cascade_tmp, c->cascade_tmp_stride, 0,
                      c->cascade[0]->dst_h, dst, dstStride, SWS_MEM_SRC_DEVICE);
    if (ret < 0)
        set_error(c, "cascade stage 1: %s", sws_internal(c->cascade[1])->last_error);
    return ret;
}

int sws_scale(SwsContext *sws, const uint8_t *const srcSlice[], const int srcStride[],
              int srcSliceY, int srcSliceH, uint8_t *const dst[], const int dstStride[])
{
    SwsInternal *c = sws_internal(sws);
    if (!c || !c->initialized)
        return AVERROR(EINVAL);
    if (c->cascade[0]) {
        if (!srcStride || !dstStride || !dst || !srcSlice)
            return AVERROR(EINVAL);
        if (c->refused)
            return AVERROR(ENOTSUP);
        return scale_cascaded(c, srcSlice, srcStride, srcSliceY, srcSliceH, dst, dstStride);
    }
    return scale_slice(sws, srcSlice, srcStride, srcSliceY, srcSliceH, dst, dstStride, 0);
}

/* Whole source frame in, destination rows [y0, y1) out (dst[] addresses frame row 0): sws_receive_slice()
 * and the destination-slice mode of the in-tree hook (reference swscale.c:371-375).  `upload` = 0: the source
 * was already moved (and, cascaded, the intermediate built) by an earlier call for the same frame. */
int ff_b200_scale_frame_rows(SwsInternal *c, const uint8_t *const src[4], const int srcStride[4], int upload,
                             uint8_t *const dst[4], const int dstStride[4], int y0, int y1)
{
    int ret;
    if (c->cascade[0]) {
        SwsInternal *c0 = sws_internal(c->cascade[0]), *c1 = sws_internal(c->cascade[1]);
        if (upload) {
            ret = ff_b200_cuda_scale_host(c0->cuda, src, srcStride, 0, c0->opts.src_h, 1, c->cascade_tmp,
                                          c->cascade_tmp_stride, 0, c0->opts.dst_h, SWS_MEM_DST_DEVICE);
            if (ret < 0) {
                set_error(c, "cascade stage 0 failed (%d)", ret);
                return ret;
            }
        }
        ret = ff_b200_cuda_scale_host(c1->cuda, (const uint8_t *const *)c->cascade_tmp, c->cascade_tmp_stride, 0,
                                      c1->opts.src_h, 0, dst, dstStride, y0, y1, SWS_MEM_SRC_DEVICE);
        if (ret < 0)
            set_error(c, "cascade stage 1 failed (%d)", ret);
        return ret;
    }
    ret = ff_b200_cuda_scale_host(c->cuda, src, srcStride, 0, c->opts.src_h, upload, dst, dstStride, y0, y1, 0);
    if (ret < 0)
        set_error(c, "CUDA conversion failed (%d)", ret);
    return ret;
}

/* ------------------------------------------------------------ CUDA extension */

int sws_cuda_scale_batch(SwsContext *sws, const uint8_t *const src[4], const int srcStride[4],
                         const int64_t srcFrameStride[4], uint8_t *const dst[4],
                         const int dstStride[4], const int64_t dstFrameStride[4], int nb_frames)
{
    SwsInternal *c = sws_internal(sws);
    int ret;
    if (!c || !c->initialized || !src || !dst || nb_frames < 1)
        return AVERROR(EINVAL);
    if (!srcStride || !dstStride || (nb_frames > 1 && (!srcFrameStride || !dstFrameStride))) {
        set_error(c, "sws_cuda_scale_batch(): stride arrays must not be NULL");
        return AVERROR(EINVAL);
    }
    if (!src[0] || !dst[0] || srcStride[0] <= 0 || dstStride[0] <= 0) {
        set_error(c, "sws_cuda_scale_batch(): bad image pointers or strides");
        return AVERROR(EINVAL);
    }
    if (c->refused) {
        set_error(c, "the last sws_setColorspaceDetails() asked for a conversion that is not on the CUDA hot path");
        return AVERROR(ENOTSUP);
    }
    if (c->cascade[0]) {
        /* frame by frame through both stages, the intermediate stays in HBM */
        SwsInternal *c0 = sws_internal(c->cascade[0]), *c1 = sws_internal(c->cascade[1]);
        for (int f = 0; f < nb_frames; f++) {
            const uint8_t *s[4];
            uint8_t *d[4];
            for (int i = 0; i < 4; i++) {
                s[i] = src[i] ? src[i] + (srcFrameStride ? srcFrameStride[i] : 0) * f : NULL;
                d[i] = dst[i] ? dst[i] + (dstFrameStride ? dstFrameStride[i] : 0) * f : NULL;
            }
            ret = ff_b200_cuda_launch(c0->cuda, s, srcStride, NULL, c->cascade_tmp, c->cascade_tmp_stride, NULL, 1, 0,
                                      c0->opts.dst_h);
            if (ret >= 0)
                ret = ff_b200_cuda_sync(c0->cuda);
            if (ret >= 0)
                ret = ff_b200_cuda_launch(c1->cuda, (const uint8_t *const *)c->cascade_tmp, c->cascade_tmp_stride, NULL,
                                          d, dstStride, NULL, 1, 0, sws->dst_h);
            if (ret >= 0)
                ret = ff_b200_cuda_sync(c1->cuda);
            if (ret < 0)
                return ret;
        }
        return sws->dst_h;
    }
    ret = ff_b200_cuda_launch(c->cuda, src, srcStride, srcFrameStride, dst, dstStride,
                              dstFrameStride, nb_frames, 0, sws->dst_h);
    return ret < 0 ? ret : sws->dst_h;
}

/* ---- host frames, several in flight, all visible devices (frames round-robin) ---- */

typedef struct LaneJob {
    SwsInternal *c;
    SwsCudaState *st;
    int lane, nb_lanes, nb_frames, own_thread;
    const uint8_t *const *src; const int *src_stride; const int64_t *src_fstride;
    uint8_t *const *dst; const int *dst_stride; const int64_t *dst_fstride;
    int ret;
} LaneJob;

static void *lane_main(void *arg)
{
    LaneJob *j = arg;
    const SwsInternal *c = j->c;
    if (j->own_thread)              /* never touch the affinity of the caller's thread */
        sws_cuda_bind_thread_to_device(ff_b200_cuda_device_of(j->st));
    /* frame f belongs to lane f mod nb_lanes, and lane l to device l mod nb_devices: consecutive frames land
     * on different devices, consecutive frames of one device on different lanes */
    for (int f = j->lane; f < j->nb_frames; f += j->nb_lanes) {
        const uint8_t *s[4];
        uint8_t *d[4];
        for (int i = 0; i < 4; i++) {
            s[i] = j->src[i] ? j->src[i] + (j->src_fstride ? j->src_fstride[i] : 0) * f : NULL;
            d[i] = j->dst[i] ? j->dst[i] + (j->dst_fstride ? j->dst_fstride[i] : 0) * f : NULL;
        }
        j->ret = ff_b200_cuda_scale_host(j->st, s, j->src_stride, 0, c->opts.src_h, 1, d, j->dst_stride, 0, c->opts.dst_h, 0);
        if (j->ret < 0)
            break;
    }
    return NULL;
}

int sws_cuda_scale_batch_host(SwsContext *sws, const uint8_t *const src[4], const int srcStride[4],
                              const int64_t srcFrameStride[4], uint8_t *const dst[4],
                              const int dstStride[4], const int64_t dstFrameStride[4],
                              int nb_frames, int nb_devices, int depth)
{
    SwsInternal *c = sws_internal(sws);
    LaneJob jobs[SWS_B200_MAX_LANES];
    pthread_t tid[SWS_B200_MAX_LANES];
    int started[SWS_B200_MAX_LANES] = { 0 };
    int visible, lanes, pinned, ret = 0;
    if (!c || !c->initialized || !src || !dst || nb_frames < 1 || !srcStride || !dstStride ||
        (nb_frames > 1 && (!srcFrameStride || !dstFrameStride)))
        return AVERROR(EINVAL);
    if (c->refused) {
        set_error(c, "the last sws_setColorspaceDetails() asked for a conversion that is not on the CUDA hot path");
        return AVERROR(ENOTSUP);
    }
    if (c->cascade[0]) {                /* cascaded conversions go frame by frame through sws_scale() */
        for (int f = 0; f < nb_frames; f++) {
            const uint8_t *s[4];
            uint8_t *d[4];
            for (int i = 0; i < 4; i++) {
                s[i] = src[i] ? src[i] + (srcFrameStride ? srcFrameStride[i] : 0) * f : NULL;
                d[i] = dst[i] ? dst[i] + (dstFrameStride ? dstFrameStride[i] : 0) * f : NULL;
            }
            ret = sws_scale(sws, s, srcStride, 0, sws->src_h, d, dstStride);
            if (ret < 0)
                return ret;
        }
        return sws->dst_h;
    }
    visible = sws_cuda_device_count();
    if (nb_devices <= 0 || nb_devices > visible)
        nb_devices = visible;
    if (depth <= 0)
        depth = 3;                      /* H2D of one frame, kernel of another, D2H of a third */
    if (nb_devices * depth > SWS_B200_MAX_LANES)
        depth = SWS_B200_MAX_LANES / nb_devices;
    {
        /* page-locked frames need no host work at all: one lane per device, everything enqueued from this thread */
        const uint8_t *last_s[4];
        const uint8_t *last_d[4];
        for (int i = 0; i < 4; i++) {
            last_s[i] = src[i] ? src[i] + (srcFrameStride ? srcFrameStride[i] : 0) * (nb_frames - 1) : NULL;
            last_d[i] = dst[i] ? dst[i] + (dstFrameStride ? dstFrameStride[i] : 0) * (nb_frames - 1) : NULL;
        }
        pinned = ff_b200_cuda_frame_is_pinned(c->cuda, src, 0) && ff_b200_cuda_frame_is_pinned(c->cuda, last_s, 0) &&
                 ff_b200_cuda_frame_is_pinned(c->cuda, (const uint8_t *const *)dst, 1) &&
                 ff_b200_cuda_frame_is_pinned(c->cuda, last_d, 1);
        for (int i = 0; i < 4; i++)
            if ((src[i] && srcStride[i] <= 0) || (dst[i] && dstStride[i] <= 0))
                pinned = 0;
    }
    lanes = pinned ? nb_devices : nb_devices * depth;
    if (lanes > nb_frames)
        lanes = nb_frames;
    if (c->nb_lanes && c->lanes_devices != nb_devices)
        release_lanes(c);
    /* every lane is a clone of the context's device state with its own streams and staging
     * (tables < 1 MB per lane) */
    for (int l = c->nb_lanes; l < lanes; l++) {
        SwsCudaPlan plan = c->plan;
        /* the rotation starts at the context's own device: lane l -> device (base + l mod nb_devices) */
        const int dev = (ff_b200_cuda_device_of(c->cuda) + l % nb_devices) % visible;
        int r = ff_b200_cuda_create_on(dev, &c->lanes[l], &plan, &c->h_lum, &c->h_chr, &c->v_lum, &c->v_chr);
        if (r < 0) {
            if (c->lanes[l])
                ff_b200_cuda_destroy(c->lanes[l]);
            c->lanes[l] = NULL;
            set_error(c, "could not create lane %d of the host batch (%d)", l, r);
            return r;
        }
        c->nb_lanes = l + 1;
    }
    c->lanes_devices = nb_devices;
    c->lanes_depth = depth;

    if (pinned) {
        for (int l = 0; l < lanes && ret >= 0; l++)
            ret = ff_b200_cuda_frames_enqueue(c->lanes[l], src, srcStride, srcFrameStride, dst, dstStride,
                                              dstFrameStride, l, lanes, nb_frames);
        for (int l = 0; l < lanes; l++) {
            int r = ff_b200_cuda_frames_wait(c->lanes[l]);
            if (r < 0 && ret >= 0)
                ret = r;
        }
        if (ret < 0) {
            set_error(c, "host batch conversion failed (%d)", ret);
            return ret;
        }
        return sws->dst_h;
    }

    for (int l = 0; l < lanes; l++) {
        LaneJob *j = &jobs[l];
        j->c = c; j->st = c->lanes[l]; j->lane = l; j->nb_lanes = lanes; j->nb_frames = nb_frames;
        j->src = src; j->src_stride = srcStride; j->src_fstride = srcFrameStride;
        j->dst = dst; j->dst_stride = dstStride; j->dst_fstride = dstFrameStride;
        j->ret = 0;
        j->own_thread = l != lanes - 1;
        if (l == lanes - 1) {
            lane_main(j);               /* the calling thread works too */
        } else if (pthread_create(&tid[l], NULL, lane_main, j) == 0) {
            started[l] = 1;
        } else {
            j->own_thread = 0;
            lane_main(j);
        }
    }
    for (int l = 0; l < lanes; l++) {
        if (started[l])
            pthread_join(tid[l], NULL);
        if (jobs[l].ret < 0 && ret >= 0)
            ret = jobs[l].ret;
    }
    if (ret < 0) {
        set_error(c, "host batch conversion failed (%d)", ret);
        return ret;
    }
    return sws->dst_h;
}

int sws_cuda_sync(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    if (c && c->cascade[1])
        return sws_cuda_sync(c->cascade[1]);
    return c && c->cuda ? ff_b200_cuda_sync(c->cuda) : AVERROR(EINVAL);
}

void *sws_cuda_stream(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    if (c && c->cascade[1])
        return sws_cuda_stream(c->cascade[1]);
    return c && c->cuda ? ff_b200_cuda_stream(c->cuda) : NULL;
}

long sws_cuda_launch_count(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    if (c && c->cascade[0])
        return sws_cuda_launch_count(c->cascade[0]) + sws_cuda_launch_count(c->cascade[1]);
    return c && c->cuda ? ff_b200_cuda_launch_count(c->cuda) : 0;
}

const char *sws_cuda_kernel_name(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    if (c && c->cascade[1])
        return "cascade";
    return c && c->cuda ? ff_b200_cuda_kernel_name(c->cuda) : "";
}

const char *sws_cuda_last_error(SwsContext *sws)
{
    SwsInternal *c = sws_internal(sws);
    return c ? c->last_error : "";
}

/* ------------------------------------------------------------ SwsVector helpers
 * (utils.c:1956-2248; only the subset callers of the legacy API use) */

SwsVector *sws_allocVec(int length)
{
    SwsVector *v;
    if (length <= 0 || length > (int)(INT32_MAX / sizeof(double)))
        return NULL;
    v = malloc(sizeof(*v));
    if (!v)
        return NULL;
    v->length = length;
    v->coeff = malloc(sizeof(double) * (size_t)length);
    if (!v->coeff) {
        free(v);
        return NULL;
    }
    return v;
}

void sws_scaleVec(SwsVector *a, double scalar)
{
    for (int i = 0; i < a->length; i++)
        a->coeff[i] *= scalar;
}

void sws_normalizeVec(SwsVector *a, double height)
{
    double sum = 0;
    for (int i = 0; i < a->length; i++)
        sum += a->coeff[i];
    sws_scaleVec(a, height / sum);
}

SwsVector *sws_getGaussianVec(double variance, double quality)
{
    const int length = (int)(variance * quality + 0.5) | 1;
    const double middle = (length - 1) * 0.5;
    SwsVector *v;
    if (variance < 0 || quality < 0)
        return NULL;
    v = sws_allocVec(length);
    if (!v)
        return NULL;
    for (int i = 0; i < length; i++) {
        double dist = i - middle;
        v->coeff[i] = exp(-dist * dist / (2 * variance * variance)) / sqrt(2 * variance * M_PI);
    }
    sws_normalizeVec(v, 1.0);
    return v;
}

void sws_freeVec(SwsVector *a)
{
    if (!a)
        return;
    free(a->coeff);
    free(a);
}
