/*
 * sws_compat.c -- the remaining entry points of the reference's export list
 * (libswscale/libswscale.v; SURVEY.md §8b lists the 40 symbols) that are not part of the
 * conversion itself: colour-property queries used by libavfilter/vf_scale.c:328-478, the
 * SwsFilter builder of the legacy API and the two palette helpers.  Host-only C.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sws_internal.h"
#include "swscale_b200_frame.h"

/* ---- colour property queries (reference libswscale/format.c:627-691) ---- */

int sws_test_colorspace(int csp, int output)
{
    (void)output;
    switch (csp) {
    case AVCOL_SPC_UNSPECIFIED:
    case AVCOL_SPC_RGB:
    case AVCOL_SPC_BT709:
    case AVCOL_SPC_BT470BG:
    case AVCOL_SPC_SMPTE170M:
    case AVCOL_SPC_FCC:
    case AVCOL_SPC_SMPTE240M:
    case AVCOL_SPC_BT2020_NCL:
        return 1;
    default:
        return 0;
    }
}

/* AVColorPrimaries: 1..12 and 22 are defined, 3 is reserved, 256 is the first extension value
 * (libavutil/pixfmt.h:643-665).  The CUDA path converts between equal primaries only
 * (sws_frame_setup refuses the rest), but the query answers like the reference. */
int sws_test_primaries(int prim, int output)
{
    (void)output;
    return ((prim > 0 && prim < 23) || prim == 256) && prim != 3;
}

/* AVColorTransferCharacteristic: every ITU value with an EOTF in libavutil/csp.c:660-720, i.e. all
 * of 1..18 except reserved (3) and the two logarithmic curves (9, 10); unspecified (2) passes. */
int sws_test_transfer(int trc, int output)
{
    (void)output;
    if (trc == 2)
        return 1;
    return trc >= 1 && trc <= 18 && trc != 3 && trc != 9 && trc != 10;
}

int sws_test_frame(const AVFrame *frame, int output)
{
    if (!frame || frame->width <= 0 || frame->height <= 0)
        return 0;
    int format = frame->format;
    if (frame->hw_frames_ctx) {          /* ff_fmt_from_frame: test the software layout of a hardware frame */
        const AVHWFramesContext *fc = (const AVHWFramesContext *)((const AVBufferRef *)frame->hw_frames_ctx)->data;
        if (!fc || !sws_test_hw_format(frame->format))
            return 0;
        format = fc->sw_format;
    }
    return sws_test_format(format, output) &&
           sws_test_colorspace(frame->colorspace, output) &&
           sws_test_primaries(frame->color_primaries, output) &&
           sws_test_transfer(frame->color_trc, output) &&
           (unsigned)frame->color_range < 3u &&          /* AVCOL_RANGE_NB */
           (unsigned)frame->chroma_location < 7u;        /* AVCHROMA_LOC_NB */
}

/* ---- SwsFilter builder (reference libswscale/utils.c:1956-2248) ----
 * sws_init_context() convolves these vectors into the FIR banks exactly like initFilter does (sws_filter.c,
 * stage 1b); callers that construct one (libavfilter/vf_smartblur.c:144-163) get the same vectors and the
 * same pictures. */

static SwsVector *const_vec(double c, int length)
{
    SwsVector *v = sws_allocVec(length);
    if (v)
        for (int i = 0; i < length; i++)
            v->coeff[i] = c;
    return v;
}

static void make_nan(SwsVector *a)
{
    for (int i = 0; i < a->length; i++)
        a->coeff[i] = NAN;
}

static int has_nan(const SwsVector *a)
{
    for (int i = 0; i < a->length; i++)
        if (isnan(a->coeff[i]))
            return 1;
    return 0;
}

/* a += b, both centred */
static void add_vec(SwsVector *a, const SwsVector *b)
{
    const int length = a->length > b->length ? a->length : b->length;
    SwsVector *sum = const_vec(0.0, length);
    if (!sum) {
        make_nan(a);
        return;
    }
    for (int i = 0; i < a->length; i++)
        sum->coeff[i + (length - 1) / 2 - (a->length - 1) / 2] += a->coeff[i];
    for (int i = 0; i < b->length; i++)
        sum->coeff[i + (length - 1) / 2 - (b->length - 1) / 2] += b->coeff[i];
    free(a->coeff);
    a->coeff = sum->coeff;
    a->length = sum->length;
    free(sum);
}

/* shift left, or right if `shift` is negative */
static void shift_vec(SwsVector *a, int shift)
{
    const int length = a->length + abs(shift) * 2;
    SwsVector *s = const_vec(0.0, length);
    if (!s) {
        make_nan(a);
        return;
    }
    for (int i = 0; i < a->length; i++)
        s->coeff[i + (length - 1) / 2 - (a->length - 1) / 2 - shift] = a->coeff[i];
    free(a->coeff);
    a->coeff = s->coeff;
    a->length = s->length;
    free(s);
}

void sws_freeFilter(SwsFilter *filter)
{
    if (!filter)
        return;
    sws_freeVec(filter->lumH);
    sws_freeVec(filter->lumV);
    sws_freeVec(filter->chrH);
    sws_freeVec(filter->chrV);
    free(filter);
}

SwsFilter *sws_getDefaultFilter(float lumaGBlur, float chromaGBlur, float lumaSharpen, float chromaSharpen,
                                float chromaHShift, float chromaVShift, int verbose)
{
    SwsFilter *f = calloc(1, sizeof(*f));
    (void)verbose;
    if (!f)
        return NULL;
    f->lumH = lumaGBlur != 0.0 ? sws_getGaussianVec(lumaGBlur, 3.0) : const_vec(1.0, 1);
    f->lumV = lumaGBlur != 0.0 ? sws_getGaussianVec(lumaGBlur, 3.0) : const_vec(1.0, 1);
    f->chrH = chromaGBlur != 0.0 ? sws_getGaussianVec(chromaGBlur, 3.0) : const_vec(1.0, 1);
    f->chrV = chromaGBlur != 0.0 ? sws_getGaussianVec(chromaGBlur, 3.0) : const_vec(1.0, 1);
    if (!f->lumH || !f->lumV || !f->chrH || !f->chrV)
        goto fail;

    if (chromaSharpen != 0.0) {
        SwsVector *id = const_vec(1.0, 1);
        if (!id)
            goto fail;
        sws_scaleVec(f->chrH, -chromaSharpen);
        sws_scaleVec(f->chrV, -chromaSharpen);
        add_vec(f->chrH, id);
        add_vec(f->chrV, id);
        sws_freeVec(id);
    }
    if (lumaSharpen != 0.0) {
        SwsVector *id = const_vec(1.0, 1);
        if (!id)
            goto fail;
        sws_scaleVec(f->lumH, -lumaSharpen);
        sws_scaleVec(f->lumV, -lumaSharpen);
        add_vec(f->lumH, id);
        add_vec(f->lumV, id);
        sws_freeVec(id);
    }
    if (chromaHShift != 0.0)
        shift_vec(f->chrH, (int)(chromaHShift + 0.5));
    if (chromaVShift != 0.0)
        shift_vec(f->chrV, (int)(chromaVShift + 0.5));

    sws_normalizeVec(f->chrH, 1.0);
    sws_normalizeVec(f->chrV, 1.0);
    sws_normalizeVec(f->lumH, 1.0);
    sws_normalizeVec(f->lumV, 1.0);
    if (has_nan(f->chrH) || has_nan(f->chrV) || has_nan(f->lumH) || has_nan(f->lumV))
        goto fail;
    return f;

fail:
    sws_freeFilter(f);
    return NULL;
}

/* ---- palette helpers (reference libswscale/swscale_unscaled.c:2733-2760) ---- */

/* palette entries and destination pixels are the same packed 32-bit format */
void sws_convertPalette8ToPacked32(const uint8_t *src, uint8_t *dst, int num_pixels, const uint8_t *palette)
{
    for (int i = 0; i < num_pixels; i++)
        memcpy(dst + 4 * (size_t)i, palette + 4 * (size_t)src[i], 4);
}

/* palette entries ABCD -> destination pixels ABC */
void sws_convertPalette8ToPacked24(const uint8_t *src, uint8_t *dst, int num_pixels, const uint8_t *palette)
{
    for (int i = 0; i < num_pixels; i++)
        memcpy(dst + 3 * (size_t)i, palette + 4 * (size_t)src[i], 3);
}
