/*
 * libswscale/cuda/swscale_cuda.c -- ff_sws_init_swscale_cuda(): the B200 sibling of
 * ff_sws_init_swscale_x86() / _aarch64() / ... (libswscale/swscale.c:697-714, prototypes
 * swscale_internal.h:1034-1040) and of ff_get_unscaled_swscale_aarch64() & co.
 * (swscale_unscaled.c:2699-2705).
 *
 * THIS FILE IS WRITTEN FOR THE REFERENCE TREE: it includes the reference's own swscale_internal.h and is
 * compiled by integration/build_hooked.py together with the reference's libswscale sources (read where
 * they lie, five of them patched on the fly with the few lines listed in integration/hook_patch.py).  It
 * reaches the B200 library only through include/swscale_b200_hook.h (plain ints and pointers).
 *
 * The per-arch hooks of the reference install per-LINE function pointers (hyScale, yuv2packedX, ...).
 * A GPU wants a whole slice per call, so this hook claims the context at the granularity of
 * c->convert_unscaled / ff_swscale() (swscale.c:1163-1192): when it succeeds, c->cuda_priv is set and
 * scale_internal() / the legacy graph pass hand the slice to ff_sws_cuda_scale().  When the B200 library
 * declines (no device, conversion outside its path, FIR banks that differ from the ones the reference
 * computed -- e.g. a caller-supplied SwsFilter) nothing is installed and the C kernels run as before.
 */
#include <stdatomic.h>
#include <string.h>

#include "libavutil/attributes.h"
#include "libavutil/log.h"
#include "libavutil/mem.h"
#include "libavutil/refstruct.h"
#include "libswscale/swscale_internal.h"

#include "swscale_b200_hook.h"

typedef struct SwsCudaHook {
    void *b200;                 /* context of the B200 library with the same options */
} SwsCudaHook;

static atomic_long cuda_slices_total;       /* slices handed to the GPU by any context (diagnostics) */

static void hook_free(AVRefStructOpaque opaque, void *obj)
{
    SwsCudaHook *h = obj;
    sws_b200_hook_close(h->b200);
}

static void fill_params(SwsB200HookParams *p, const SwsContext *o)
{
    memset(p, 0, sizeof(*p));
    p->flags = o->flags;
    p->scaler_params[0] = o->scaler_params[0];
    p->scaler_params[1] = o->scaler_params[1];
    p->dither = o->dither; p->alpha_blend = o->alpha_blend; p->gamma_flag = o->gamma_flag;
    p->src_w = o->src_w; p->src_h = o->src_h; p->dst_w = o->dst_w; p->dst_h = o->dst_h;
    p->src_format = o->src_format; p->dst_format = o->dst_format;
    p->src_range = o->src_range; p->dst_range = o->dst_range;
    p->src_v_chr_pos = o->src_v_chr_pos; p->src_h_chr_pos = o->src_h_chr_pos;
    p->dst_v_chr_pos = o->dst_v_chr_pos; p->dst_h_chr_pos = o->dst_h_chr_pos;
    p->scaler = o->scaler; p->scaler_sub = o->scaler_sub;
}

/* Do the two libraries agree on one FIR bank?  Rows may differ in width (filterAlign padding): compare
 * tap by tap over the union of both windows. */
static int bank_equal(void *b200, int which, const int16_t *coef, const int32_t *pos, int size, int len)
{
    const int16_t *bc;
    const int32_t *bp;
    int blen = 0;
    const int bsize = sws_b200_hook_bank(b200, which, &bc, &bp, &blen);
    if (bsize < 0 || blen != len || !coef || !pos)
        return 0;
    for (int i = 0; i < len; i++) {
        const int lo = FFMIN(pos[i], bp[i]);
        const int hi = FFMAX(pos[i] + size, bp[i] + bsize);
        for (int x = lo; x < hi; x++) {
            const int a = x >= pos[i] && x < pos[i] + size ? coef[i * size + x - pos[i]] : 0;
            const int b = x >= bp[i] && x < bp[i] + bsize ? bc[i * bsize + x - bp[i]] : 0;
            if (a != b)
                return 0;
        }
    }
    return 1;
}

static av_cold int hook_install(SwsInternal *c, int unscaled)
{
    SwsB200HookParams p;
    SwsCudaHook *h;
    void *b200;
    int err = 0;

    if (c->cuda_priv || c->parent)          /* slice-thread clones of a legacy context stay on the C path */
        return 0;
    fill_params(&p, &c->opts);
    b200 = sws_b200_hook_open(&p, &err);
    if (!b200) {
        av_log(c, AV_LOG_DEBUG, "CUDA path not used (%d)\n", err);
        return 0;
    }
    /* both libraries must have taken the same fork: special converter vs. FIR pipeline */
    if (!!sws_b200_hook_is_unscaled(b200) != !!unscaled)
        goto decline;
    if (!unscaled &&
        (!bank_equal(b200, 0, c->hLumFilter, c->hLumFilterPos, c->hLumFilterSize, c->opts.dst_w) ||
         !bank_equal(b200, 1, c->hChrFilter, c->hChrFilterPos, c->hChrFilterSize, c->chrDstW)     ||
         !bank_equal(b200, 2, c->vLumFilter, c->vLumFilterPos, c->vLumFilterSize, c->opts.dst_h) ||
         !bank_equal(b200, 3, c->vChrFilter, c->vChrFilterPos, c->vChrFilterSize, c->chrDstH)))
        goto decline;
    /* colour tables the context already carries (sws_init_context() set the defaults, utils.c:1186-1194) */
    if (sws_b200_hook_colorspace(b200, c->srcColorspaceTable, c->opts.src_range, c->dstColorspaceTable,
                                 c->opts.dst_range, c->brightness, c->contrast, c->saturation) < 0)
        goto decline;

    h = av_refstruct_alloc_ext(sizeof(*h), 0, NULL, hook_free);
    if (!h)
        goto decline;
    h->b200 = b200;
    c->cuda_priv = h;
    if (c->opts.flags & SWS_PRINT_INFO)
        av_log(c, AV_LOG_INFO, "using the B200 CUDA path (%s)\n", sws_b200_hook_kernel(b200));
    return 1;

decline:
    sws_b200_hook_close(b200);
    return 0;
}

/* called at the end of ff_sws_init_scale() (swscale.c:713), after the C function pointers are set */
av_cold void ff_sws_init_swscale_cuda(SwsInternal *c)
{
    hook_install(c, 0);
}

/* called at the end of ff_get_unscaled_swscale() (swscale_unscaled.c:2705): only claims conversions for
 * which the reference found a special converter, so path selection stays the reference's */
av_cold void ff_get_unscaled_swscale_cuda(SwsInternal *c)
{
    if (c->convert_unscaled)
        hook_install(c, 1);
}

/* called from sws_setColorspaceDetails() (utils.c:849) with the caller's arguments.  A refusal (e.g. YUV->YUV with
 * two different matrices, which the reference cascades through RGB) hands the context back to the C kernels. */
void ff_sws_cuda_set_colorspace(SwsInternal *c, const int inv_table[4], int srcRange, const int table[4],
                                int dstRange, int brightness, int contrast, int saturation)
{
    SwsCudaHook *h = c->cuda_priv;
    if (!h)
        return;
    if (sws_b200_hook_colorspace(h->b200, inv_table, srcRange, table, dstRange, brightness, contrast, saturation) < 0)
        av_refstruct_unref(&c->cuda_priv);
}

/* The slice entry: same arguments as ff_swscale() (swscale.c:263), called from scale_internal() in its place.
 * Strides are negative when the caller feeds bottom-up slices (swscale.c:1141-1159). */
int ff_sws_cuda_scale(SwsInternal *c, const uint8_t *const src[], const int srcStride[], int srcSliceY, int srcSliceH,
                      uint8_t *const dst[], const int dstStride[], int dstSliceY, int dstSliceH)
{
    SwsCudaHook *h = c->cuda_priv;
    const int scale_dst = dstSliceY > 0 || dstSliceH < c->opts.dst_h;
    int ret;
    if (scale_dst)
        ret = sws_b200_hook_scale_rows(h->b200, src, srcStride, dst, dstStride, dstSliceY, dstSliceH);
    else
        ret = sws_b200_hook_scale(h->b200, src, srcStride, srcSliceY, srcSliceH, dst, dstStride);
    if (ret < 0)
        av_log(c, AV_LOG_ERROR, "CUDA conversion failed: %s\n", sws_b200_hook_error(h->b200));
    else
        atomic_fetch_add_explicit(&cuda_slices_total, 1, memory_order_relaxed);
    return ret;
}

/* diagnostics for the test-suite: kernels launched so far / the kernel variant in use */
long ff_sws_cuda_launches(const SwsInternal *c)
{
    const SwsCudaHook *h = c->cuda_priv;
    return h ? sws_b200_hook_launches(h->b200) : -1;
}

long ff_sws_cuda_slices_total(void)
{
    return atomic_load_explicit(&cuda_slices_total, memory_order_relaxed);
}

const char *ff_sws_cuda_kernel(const SwsInternal *c)
{
    const SwsCudaHook *h = c->cuda_priv;
    return h ? sws_b200_hook_kernel(h->b200) : "";
}
