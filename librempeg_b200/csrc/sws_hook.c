/*
 * sws_hook.c -- the narrow entry the in-tree hook binds (include/swscale_b200_hook.h).
 *
 * integration/swscale_cuda.c (ff_sws_init_swscale_cuda(), the CUDA sibling of
 * ff_sws_init_swscale_x86() & co., reference libswscale/swscale.c:697-714) is compiled against the
 * reference's headers and reaches this library through plain ints and pointers only.  Everything here is a
 * thin adapter over the public API of this library: a hooked reference context owns ONE B200 context
 * created from the same option values.
 */
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "sws_internal.h"
#include "swscale_b200_frame.h"
#include "swscale_b200_hook.h"

void *sws_b200_hook_open(const SwsB200HookParams *p, int *err)
{
    SwsContext *s;
    int ret;
    if (err)
        *err = 0;
    if (!p) {
        if (err)
            *err = AVERROR(EINVAL);
        return NULL;
    }
    s = sws_alloc_context();
    if (!s) {
        if (err)
            *err = AVERROR(ENOMEM);
        return NULL;
    }
    s->flags = p->flags;
    s->scaler_params[0] = p->scaler_params[0];
    s->scaler_params[1] = p->scaler_params[1];
    s->dither = (SwsDither)p->dither;
    s->alpha_blend = (SwsAlphaBlend)p->alpha_blend;
    s->gamma_flag = p->gamma_flag;
    s->src_w = p->src_w; s->src_h = p->src_h; s->dst_w = p->dst_w; s->dst_h = p->dst_h;
    s->src_format = p->src_format; s->dst_format = p->dst_format;
    s->src_range = p->src_range; s->dst_range = p->dst_range;
    s->src_v_chr_pos = p->src_v_chr_pos; s->src_h_chr_pos = p->src_h_chr_pos;
    s->dst_v_chr_pos = p->dst_v_chr_pos; s->dst_h_chr_pos = p->dst_h_chr_pos;
    s->scaler = (SwsScaler)p->scaler; s->scaler_sub = (SwsScaler)p->scaler_sub;
    ret = sws_init_context(s, NULL, NULL);
    if (ret < 0) {
        if (err)
            *err = ret;
        sws_free_context(&s);
        return NULL;
    }
    return s;
}

void sws_b200_hook_close(void *h)
{
    SwsContext *s = h;
    sws_free_context(&s);
}

int sws_b200_hook_colorspace(void *h, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                             int brightness, int contrast, int saturation)
{
    return sws_setColorspaceDetails(h, inv_table, srcRange, table, dstRange, brightness, contrast, saturation);
}

int sws_b200_hook_bank(void *h, int which, const int16_t **coef, const int32_t **pos, int *len)
{
    return sws_b200_get_filter(h, which, coef, pos, len);
}

int sws_b200_hook_is_unscaled(void *h)
{
    const SwsInternal *c = sws_internal(h);
    return c && (c->unscaled_lut || c->special);
}

int sws_b200_hook_dst_slice_align(void *h)
{
    return (int)sws_receive_slice_alignment(h);
}

int sws_b200_hook_scale(void *h, const uint8_t *const src[4], const int srcStride[4], int srcSliceY, int srcSliceH,
                        uint8_t *const dst[4], const int dstStride[4])
{
    return sws_scale(h, src, srcStride, srcSliceY, srcSliceH, dst, dstStride);
}

int sws_b200_hook_scale_rows(void *h, const uint8_t *const src[4], const int srcStride[4],
                             uint8_t *const dst[4], const int dstStride[4], int dstY, int dstH)
{
    SwsInternal *c = sws_internal(h);
    uint8_t *base[4] = { NULL, NULL, NULL, NULL };
    int ret;
    if (!c || !c->initialized || !src || !dst || !srcStride || !dstStride)
        return AVERROR(EINVAL);
    if (c->refused)
        return AVERROR(ENOTSUP);
    if (dstY < 0 || dstH < 0 || dstY + dstH > c->opts.dst_h)
        return AVERROR(EINVAL);
    if (!dstH)
        return 0;
    /* the shim addresses destination planes by frame row: step back from the slice's first row */
    for (int i = 0; i < 4; i++)
        if (dst[i])
            base[i] = dst[i] - (ptrdiff_t)(dstY >> ((i == 1 || i == 2) ? c->chr_dst_vsub : 0)) * dstStride[i];
    ret = ff_b200_scale_frame_rows(c, src, srcStride, 1, base, dstStride, dstY, dstY + dstH);
    return ret < 0 ? ret : dstH;
}

long sws_b200_hook_launches(void *h)
{
    return sws_cuda_launch_count(h);
}

const char *sws_b200_hook_kernel(void *h)
{
    return sws_cuda_kernel_name(h);
}

const char *sws_b200_hook_error(void *h)
{
    return sws_cuda_last_error(h);
}
