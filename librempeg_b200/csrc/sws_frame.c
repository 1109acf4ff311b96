/*
 * sws_frame.c -- AVFrame entry points: sws_scale_frame(), sws_frame_setup(),
 * sws_is_noop() and the slice-wise frame API, as vf_scale uses them
 * (reference libswscale/swscale.c:1219-1480,1500-1620; graph.c:560-661; format.c:305-339,554-590,693).
 *
 * Dynamic mode (context never passed to sws_init_context): the conversion is described by the two
 * frames; we keep ONE inner legacy context, re-planned only when the description changes, exactly
 * the role of add_legacy_sws_pass() (graph.c:560) minus the pass graph: format/size from the frames,
 * range from color_range, matrix from colorspace, chroma siting from chroma_location, scaler flags
 * and dither from the outer context.  A frame is always ONE launch (the reference splits it over
 * slice threads, graph.c:228-235; a GPU pass wants num_slices = 1 as for error diffusion, :475-476).
 */
#include <dlfcn.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sws_internal.h"
#include "swscale_b200_frame.h"
#include "swscale_b200_cuda.h"

/* Destination buffers (reference swscale.c:1316-1330,1437-1467: av_frame_get_buffer() on a legacy context, the
 * frame pool in the dynamic mode).  AVBufferRef objects can only be made by libavutil itself, and this library
 * does not link it: every caller that owns an AVFrame has it loaded, so the allocator is looked up in the
 * running process.  Without it (a caller with hand-built frame structs) the old answer stands: ENOTSUP. */
typedef int (*get_buffer_fn)(AVFrame *frame, int align);

static int alloc_dst(AVFrame *dst, int width, int height, int format)
{
    static get_buffer_fn get_buffer;
    static int looked_up;
    if (!looked_up) {
        get_buffer = (get_buffer_fn)dlsym(RTLD_DEFAULT, "av_frame_get_buffer");
        looked_up = 1;
    }
    if (!get_buffer)
        return AVERROR(ENOTSUP);
    dst->width  = width;
    dst->height = height;
    dst->format = format;
    return get_buffer(dst, 0);
}

typedef struct FrameDesc {
    int format, width, height, range, csp, loc;
    int hw;                 /* 1: AV_PIX_FMT_CUDA frame, `format` is its sw_format */
} FrameDesc;

static const AVHWFramesContext *frames_ctx(const AVFrame *f)
{
    const AVBufferRef *ref = f->hw_frames_ctx;
    return ref ? (const AVHWFramesContext *)ref->data : NULL;
}

/* sanitize_fmt + the fields ff_fmt_from_frame keeps (format.c:305-339,345-380) */
static int describe(FrameDesc *d, const AVFrame *f)
{
    const SwsPixDesc *pd;
    d->format = f->format;
    d->hw = 0;
    if (f->format == AV_PIX_FMT_CUDA) {          /* ff_fmt_from_frame: hw frames describe their sw_format */
        const AVHWFramesContext *fc = frames_ctx(f);
        if (!fc || fc->format != AV_PIX_FMT_CUDA)
            return AVERROR(EINVAL);
        d->format = fc->sw_format;
        d->hw = 1;
    }
    pd = ff_b200_pix_desc(d->format);
    d->width  = f->width;
    d->height = f->height;
    d->range  = f->color_range;
    d->csp    = f->colorspace;
    d->loc    = f->chroma_location;
    if (!pd)
        return AVERROR(ENOTSUP);
    if (pd->flags & SWSPF_RGB) {
        d->csp   = AVCOL_SPC_RGB;
        d->range = AVCOL_RANGE_JPEG;
    }
    if (pd->flags & SWSPF_JPEG)
        d->range = AVCOL_RANGE_JPEG;
    if (!pd->log2_cw && !pd->log2_ch)
        d->loc = AVCHROMA_LOC_UNSPECIFIED;
    return 0;
}

static int desc_equal(const FrameDesc *a, const FrameDesc *b)
{
    return !memcmp(a, b, sizeof(*a));
}

/* ff_sws_chroma_pos (format.c:554-590) for progressive frames */
static void chroma_pos(const FrameDesc *d, int *x_pos, int *y_pos)
{
    const SwsPixDesc *pd = ff_b200_pix_desc(d->format);
    int loc = d->loc == AVCHROMA_LOC_UNSPECIFIED ? AVCHROMA_LOC_CENTER : d->loc;
    const int pos = loc - 1;                     /* av_chroma_location_enum_to_pos */
    int x = (pos & 1) * 128;
    int y = ((pos >> 1) ^ (pos < 4)) * 128;
    x *= (1 << pd->log2_cw) - 1;
    y *= (1 << pd->log2_ch) - 1;
    *x_pos = pd->log2_cw ? x : -513;             /* graph.c:622-629 */
    *y_pos = pd->log2_ch ? y : -513;
}

int sws_is_noop(const AVFrame *dst, const AVFrame *src)
{
    FrameDesc a, b;
    if (!dst || !src || describe(&a, dst) < 0 || describe(&b, src) < 0)
        return 0;
    if ((dst->flags ^ src->flags) & AV_FRAME_FLAG_INTERLACED)
        return 0;
    return desc_equal(&a, &b) && dst->color_primaries == src->color_primaries &&
           dst->color_trc == src->color_trc;
}

typedef struct DynKey {
    FrameDesc src, dst;
    unsigned flags;
    int dither, scaler, scaler_sub;
    int chr_pos[4];
    int device;             /* CUDA device of hardware frames, -1: the caller's current device */
    double param[2];
} DynKey;

int sws_frame_setup(SwsContext *ctx, const AVFrame *dst, const AVFrame *src)
{
    SwsInternal *c = sws_internal(ctx);
    DynKey key;
    SwsContext *in;
    int ret, sx, sy, dx, dy, hw_device = -1, prev_device;

    if (!c || !src || !dst)
        return AVERROR(EINVAL);
    if (c->initialized)
        return AVERROR(EINVAL);                 /* legacy contexts use sws_frame_start() & co. */
    /* if a single frame has a frames context, then both need one (swscale.c:1511-1538) */
    if (!!src->hw_frames_ctx != !!dst->hw_frames_ctx)
        return AVERROR(ENOTSUP);
    if (src->hw_frames_ctx) {
        const AVHWFramesContext *sf = frames_ctx(src), *df = frames_ctx(dst);
        if (!src->data[0] || !dst->data[0])
            return AVERROR(EINVAL);             /* hardware frames must already be allocated */
        if (!sf || !df || !sf->device_ref || !df->device_ref ||
            sf->device_ref->data != df->device_ref->data)
            return AVERROR(EINVAL);             /* both frames must live on the same device */
        if (((const AVHWDeviceContext *)sf->device_ref->data)->type != AV_HWDEVICE_TYPE_CUDA ||
            src->format != AV_PIX_FMT_CUDA || dst->format != AV_PIX_FMT_CUDA)
            return AVERROR(ENOTSUP);            /* only CUDA devices are supported */
        {
            /* libavutil creates its own CUcontext unless told to use the primary one (hwcontext_cuda.c:740,
             * AV_CUDA_USE_PRIMARY_CONTEXT): the kernels run in the primary context of the frames' device */
            const AVCUDADeviceContext *cu = ((const AVHWDeviceContext *)sf->device_ref->data)->hwctx;
            hw_device = -1;
            if (cu && cu->cuda_ctx) {
                hw_device = ff_b200_cuda_hwctx_device(cu->cuda_ctx);
                if (hw_device < 0) {
                    snprintf(c->last_error, sizeof(c->last_error),
                             "AV_PIX_FMT_CUDA frames must live in the primary context of their device "
                             "(create the device with AV_CUDA_USE_PRIMARY_CONTEXT)");
                    return hw_device == AVERROR(ENOTSUP) ? AVERROR(ENOTSUP) : AVERROR(EINVAL);
                }
            }
        }
    } else if (src->format == AV_PIX_FMT_CUDA || dst->format == AV_PIX_FMT_CUDA) {
        return AVERROR(EINVAL);
    }
    if ((src->flags | dst->flags) & AV_FRAME_FLAG_INTERLACED)
        return AVERROR(ENOTSUP);
    if (src->width < 1 || src->height < 1 || dst->width < 1 || dst->height < 1)
        return AVERROR(EINVAL);

    memset(&key, 0, sizeof(key));
    if ((ret = describe(&key.src, src)) < 0 || (ret = describe(&key.dst, dst)) < 0)
        return ret;
    if (!sws_isSupportedInput(key.src.format) || !sws_isSupportedOutput(key.dst.format))
        return AVERROR(ENOTSUP);
    /* differing primaries / transfer need the 3DLUT pass of the ops engine: outside the hot path */
    if (src->color_primaries != dst->color_primaries && src->color_primaries > 2 && dst->color_primaries > 2)
        return AVERROR(ENOTSUP);
    if (src->color_trc != dst->color_trc && src->color_trc > 2 && dst->color_trc > 2)
        return AVERROR(ENOTSUP);
    key.device = hw_device;
    key.flags = ctx->flags; key.dither = ctx->dither; key.scaler = ctx->scaler; key.scaler_sub = ctx->scaler_sub;
    key.param[0] = ctx->scaler_params[0]; key.param[1] = ctx->scaler_params[1];
    chroma_pos(&key.src, &sx, &sy);
    chroma_pos(&key.dst, &dx, &dy);
    /* legacy overrides (graph.c:616-620) */
    key.chr_pos[0] = ctx->src_h_chr_pos != -513 ? ctx->src_h_chr_pos : sx;
    key.chr_pos[1] = ctx->src_v_chr_pos != -513 ? ctx->src_v_chr_pos : sy;
    key.chr_pos[2] = ctx->dst_h_chr_pos != -513 ? ctx->dst_h_chr_pos : dx;
    key.chr_pos[3] = ctx->dst_v_chr_pos != -513 ? ctx->dst_v_chr_pos : dy;
    if (!ff_b200_pix_desc(key.src.format)->log2_cw) key.chr_pos[0] = -513;
    if (!ff_b200_pix_desc(key.src.format)->log2_ch) key.chr_pos[1] = -513;
    if (!ff_b200_pix_desc(key.dst.format)->log2_cw) key.chr_pos[2] = -513;
    if (!ff_b200_pix_desc(key.dst.format)->log2_ch) key.chr_pos[3] = -513;

    if (c->dyn && c->dyn_key && !memcmp(c->dyn_key, &key, sizeof(key)))
        return 0;

    sws_free_context(&c->dyn);
    in = sws_alloc_context();
    if (!in)
        return AVERROR(ENOMEM);
    in->flags = ctx->flags; in->dither = ctx->dither; in->alpha_blend = ctx->alpha_blend;
    in->gamma_flag = ctx->gamma_flag; in->scaler = ctx->scaler; in->scaler_sub = ctx->scaler_sub;
    in->src_w = key.src.width; in->src_h = key.src.height; in->src_format = key.src.format;
    in->dst_w = key.dst.width; in->dst_h = key.dst.height; in->dst_format = key.dst.format;
    in->src_range = key.src.range == AVCOL_RANGE_JPEG;
    in->dst_range = key.dst.range == AVCOL_RANGE_JPEG;
    in->src_h_chr_pos = key.chr_pos[0]; in->src_v_chr_pos = key.chr_pos[1];
    in->dst_h_chr_pos = key.chr_pos[2]; in->dst_v_chr_pos = key.chr_pos[3];
    in->scaler_params[0] = key.param[0]; in->scaler_params[1] = key.param[1];
    prev_device = ff_b200_cuda_current_device();
    if (hw_device >= 0)
        ff_b200_cuda_use_device(hw_device);       /* the inner context binds to the device of the frames */
    ret = sws_init_context(in, NULL, NULL);
    if (hw_device >= 0)
        ff_b200_cuda_use_device(prev_device);
    if (ret < 0) {
        memcpy(c->last_error, sws_internal(in)->last_error, sizeof(c->last_error));
        sws_free_context(&in);
        return ret;
    }
    {   /* colour matrices from the frames (graph.c:642-658) */
        int *inv, *tab, in_full, out_full, br, co, sa;
        sws_getColorspaceDetails(in, &inv, &in_full, &tab, &out_full, &br, &co, &sa);
        if (sws_setColorspaceDetails(in, sws_getCoefficients(key.src.csp), in->src_range,
                                     sws_getCoefficients(key.dst.csp), in->dst_range, br, co, sa) < 0) {
            memcpy(c->last_error, sws_internal(in)->last_error, sizeof(c->last_error));
            sws_free_context(&in);
            return AVERROR(ENOTSUP);
        }
    }
    if (!c->dyn_key)
        c->dyn_key = malloc(sizeof(DynKey));
    if (!c->dyn_key) {
        sws_free_context(&in);
        return AVERROR(ENOMEM);
    }
    memcpy(c->dyn_key, &key, sizeof(key));
    c->dyn = in;
    return 0;
}

static int frame_matches(const SwsContext *s, const AVFrame *dst, const AVFrame *src)
{
    int sf = src->format, df = dst->format;
    /* a legacy context folded yuvj* into yuv* + range at init (utils.c:1901-1907) */
    if (ff_b200_pix_desc(sf) && (ff_b200_pix_desc(sf)->flags & SWSPF_JPEG))
        sf = sf == AV_PIX_FMT_YUVJ420P ? AV_PIX_FMT_YUV420P : sf == AV_PIX_FMT_YUVJ422P ? AV_PIX_FMT_YUV422P : AV_PIX_FMT_YUV444P;
    if (ff_b200_pix_desc(df) && (ff_b200_pix_desc(df)->flags & SWSPF_JPEG))
        df = df == AV_PIX_FMT_YUVJ420P ? AV_PIX_FMT_YUV420P : df == AV_PIX_FMT_YUVJ422P ? AV_PIX_FMT_YUV422P : AV_PIX_FMT_YUV444P;
    return src->width == s->src_w && src->height == s->src_h && sf == s->src_format &&
           dst->width == s->dst_w && dst->height == s->dst_h && df == s->dst_format;
}

int sws_scale_frame(SwsContext *ctx, AVFrame *dst, const AVFrame *src)
{
    SwsInternal *c = sws_internal(ctx);
    int ret;
    if (!c || !src || !dst)
        return AVERROR(EINVAL);

    if (c->initialized) {
        /* legacy behaviour: sws_frame_start / send_slice / receive_slice / frame_end (swscale.c:1412-1426) */
        ret = sws_frame_start(ctx, dst, src);
        if (ret < 0)
            return ret;
        ret = sws_send_slice(ctx, 0, src->height);
        if (ret >= 0)
            ret = sws_receive_slice(ctx, 0, dst->height);
        sws_frame_end(ctx);
        return ret;
    }

    ret = sws_frame_setup(ctx, dst, src);
    if (ret < 0)
        return ret;
    if (!src->data[0])
        return 0;
    if (!dst->data[0]) {
        /* user did not provide buffers: allocate them like the reference (swscale.c:1437-1467); hardware frames
         * must come allocated (sws_frame_setup() checked that) */
        ret = alloc_dst(dst, dst->width, dst->height, dst->format);
        if (ret < 0)
            return ret;
    }
    if (src->format == AV_PIX_FMT_CUDA) {
        /* device-resident planes: one launch on the context stream, complete on return.  Work the producer of the
         * source frame queued on the device context's stream (NVDEC, hwupload, a previous filter) comes first. */
        const AVHWFramesContext *sf = frames_ctx(src);
        const AVCUDADeviceContext *cu = sf && sf->device_ref ? ((const AVHWDeviceContext *)sf->device_ref->data)->hwctx : NULL;
        SwsInternal *in = sws_internal(c->dyn);
        if (cu && cu->cuda_ctx && in->cuda && (ret = ff_b200_cuda_wait_stream(in->cuda, cu->stream)) < 0)
            return ret;
        ret = sws_cuda_scale_batch(c->dyn, (const uint8_t *const *)src->data, src->linesize, NULL,
                                   dst->data, dst->linesize, NULL, 1);
        if (ret >= 0)
            ret = sws_cuda_sync(c->dyn);
        if (ret < 0)
            memcpy(c->last_error, sws_internal(c->dyn)->last_error, sizeof(c->last_error));
        return ret < 0 ? ret : 0;
    }
    ret = sws_scale(c->dyn, (const uint8_t *const *)src->data, src->linesize, 0, src->height,
                    dst->data, dst->linesize);
    if (ret < 0)
        memcpy(c->last_error, sws_internal(c->dyn)->last_error, sizeof(c->last_error));
    return ret < 0 ? ret : 0;
}

/* ---- slice-wise frame API of a legacy context (swscale.c:1219-1403) ---- */

int sws_frame_start(SwsContext *ctx, AVFrame *dst, const AVFrame *src)
{
    SwsInternal *c = sws_internal(ctx);
    if (!c || !c->initialized || !dst || !src)
        return AVERROR(EINVAL);
    if (!dst->data[0]) {                    /* swscale.c:1316-1330: the context describes the destination */
        int ret = alloc_dst(dst, ctx->dst_w, ctx->dst_h, ctx->dst_format);
        if (ret < 0)
            return ret;
    }
    if (!frame_matches(ctx, dst, src))
        return AVERROR(EINVAL);
    c->frame_src = src;
    c->frame_dst = dst;
    c->frame_rows_sent = 0;
    c->frame_uploaded = 0;
    return 0;
}

void sws_frame_end(SwsContext *ctx)
{
    SwsInternal *c = sws_internal(ctx);
    if (!c)
        return;
    c->frame_src = NULL;
    c->frame_dst = NULL;
    c->frame_rows_sent = 0;
    c->frame_uploaded = 0;
}

int sws_send_slice(SwsContext *ctx, unsigned int slice_start, unsigned int slice_height)
{
    SwsInternal *c = sws_internal(ctx);
    if (!c || !c->initialized || !c->frame_src)
        return AVERROR(EINVAL);
    if (slice_start + slice_height > (unsigned)ctx->src_h)
        return AVERROR(EINVAL);
    /* slices may arrive in any order but may not overlap (swscale.h:626-638): count rows */
    c->frame_rows_sent += (int)slice_height;
    return 0;
}

unsigned int sws_receive_slice_alignment(const SwsContext *ctx)
{
    const SwsInternal *c = sws_internal(ctx);
    return c && c->dst_slice_align ? (unsigned)c->dst_slice_align : 1;
}

int sws_receive_slice(SwsContext *ctx, unsigned int slice_start, unsigned int slice_height)
{
    SwsInternal *c = sws_internal(ctx);
    const unsigned align = sws_receive_slice_alignment(ctx);
    const AVFrame *src;
    AVFrame *dst;
    int ret;
    if (!c || !c->initialized || !c->frame_src)
        return AVERROR(EINVAL);
    if (c->refused)
        return AVERROR(ENOTSUP);
    if (c->frame_rows_sent < ctx->src_h)
        return AVERROR(EAGAIN);                 /* wait until the whole input was signalled */
    if (slice_start + slice_height > (unsigned)ctx->dst_h)
        return AVERROR(EINVAL);
    if ((slice_start > 0 || slice_height < (unsigned)ctx->dst_h) &&
        (slice_start % align || (slice_height % align && slice_start + slice_height != (unsigned)ctx->dst_h)))
        return AVERROR(EINVAL);
    src = (const AVFrame *)c->frame_src;
    dst = (AVFrame *)c->frame_dst;
    /* whole source is available: upload it once, then convert exactly the requested rows
     * (the reference's scale_dst mode, swscale.c:371-375) */
    ret = ff_b200_scale_frame_rows(c, (const uint8_t *const *)src->data, src->linesize, !c->frame_uploaded,
                                   dst->data, dst->linesize, (int)slice_start, (int)(slice_start + slice_height));
    if (ret < 0)
        return ret;
    c->frame_uploaded = 1;
    return (int)slice_height;
}
