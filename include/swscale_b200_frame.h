/*
 * swscale_b200_frame.h -- AVFrame entry points of the libswscale ABI.
 *
 * Mirrors reference libswscale/swscale.h:405,415,439,613-669 (sws_frame_setup, sws_is_noop,
 * sws_scale_frame, sws_frame_start/end, sws_send_slice, sws_receive_slice[_alignment]) -- the calls
 * libavfilter/vf_scale.c:840-852 makes.  libavutil is not part of this repository, so the slice of
 * its public ABI these functions read is declared here (guarded: an in-tree build includes the real
 * libavutil/frame.h first and these declarations vanish).  tests/test_frame_api_cpu.py checks every
 * offset against the reference's own struct through oracle/_ref.
 */
#ifndef SWSCALE_B200_FRAME_H
#define SWSCALE_B200_FRAME_H

#include <stdint.h>
#include <stddef.h>
#include "swscale_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef AVUTIL_FRAME_H
/* reference libavutil/pixfmt.h:707-809 */
enum AVColorRange { AVCOL_RANGE_UNSPECIFIED = 0, AVCOL_RANGE_MPEG = 1, AVCOL_RANGE_JPEG = 2 };
enum AVColorSpace {
    AVCOL_SPC_RGB = 0, AVCOL_SPC_BT709 = 1, AVCOL_SPC_UNSPECIFIED = 2, AVCOL_SPC_RESERVED = 3,
    AVCOL_SPC_FCC = 4, AVCOL_SPC_BT470BG = 5, AVCOL_SPC_SMPTE170M = 6, AVCOL_SPC_SMPTE240M = 7,
    AVCOL_SPC_YCGCO = 8, AVCOL_SPC_BT2020_NCL = 9, AVCOL_SPC_BT2020_CL = 10,
};
enum AVChromaLocation {
    AVCHROMA_LOC_UNSPECIFIED = 0, AVCHROMA_LOC_LEFT = 1, AVCHROMA_LOC_CENTER = 2, AVCHROMA_LOC_TOPLEFT = 3,
    AVCHROMA_LOC_TOP = 4, AVCHROMA_LOC_BOTTOMLEFT = 5, AVCHROMA_LOC_BOTTOM = 6,
};
#define AV_NUM_DATA_POINTERS 8
#define AV_FRAME_FLAG_INTERLACED (1 << 3)
typedef struct AVRational { int num, den; } AVRational;

/* Leading part of AVFrame, field for field (reference libavutil/frame.h:493-769).  Never allocate
 * this type by value: the real structure continues after hw_frames_ctx. */
typedef struct AVFrame {
    uint8_t *data[AV_NUM_DATA_POINTERS];
    int linesize[AV_NUM_DATA_POINTERS];
    uint8_t **extended_data;
    int width, height;
    int nb_samples;
    int format;
    int pict_type;
    AVRational sample_aspect_ratio;
    int64_t pts;
    int64_t pkt_dts;
    AVRational time_base;
    int quality;
    void *opaque;
    int repeat_pict;
    int sample_rate;
    void *buf[AV_NUM_DATA_POINTERS];
    void **extended_buf;
    int nb_extended_buf;
    void **side_data;
    int nb_side_data;
    int flags;
    int color_range;        /* enum AVColorRange */
    int color_primaries;
    int color_trc;
    int colorspace;         /* enum AVColorSpace */
    int chroma_location;    /* enum AVChromaLocation */
    int64_t best_effort_timestamp;
    void *metadata;
    int decode_error_flags;
    void *hw_frames_ctx;
} AVFrame;
#endif /* AVUTIL_FRAME_H */

#ifndef AVUTIL_HWCONTEXT_H
/* What an AV_PIX_FMT_CUDA frame carries (reference libavutil/buffer.h:82-95, hwcontext.h:27-44,63-101,
 * 118-221): frame->hw_frames_ctx is an AVBufferRef whose data points at an AVHWFramesContext;
 * data[]/linesize[] of the frame are device pointers / pitches in the layout of sw_format. */
typedef struct AVBufferRef {
    void *buffer;
    uint8_t *data;
    size_t size;
} AVBufferRef;
enum AVHWDeviceType { AV_HWDEVICE_TYPE_NONE = 0, AV_HWDEVICE_TYPE_VDPAU = 1, AV_HWDEVICE_TYPE_CUDA = 2 };
typedef struct AVHWDeviceContext {
    const void *av_class;
    int type;                        /* enum AVHWDeviceType */
    void *hwctx;
} AVHWDeviceContext;
/* what AVHWDeviceContext.hwctx points at for AV_HWDEVICE_TYPE_CUDA (reference libavutil/hwcontext_cuda.h:42-46) */
typedef struct AVCUDADeviceContext {
    void *cuda_ctx;                  /* CUcontext */
    void *stream;                    /* CUstream the producer / consumer of the frames works on */
    void *internal;
} AVCUDADeviceContext;
typedef struct AVHWFramesContext {
    const void *av_class;
    AVBufferRef *device_ref;         /* -> AVHWDeviceContext */
    AVHWDeviceContext *device_ctx;
    void *hwctx;
    void (*free)(struct AVHWFramesContext *ctx);
    void *user_opaque;
    void *pool;
    int initial_pool_size;
    int format;                      /* AV_PIX_FMT_CUDA */
    int sw_format;                   /* layout of the device planes */
    int width, height;
} AVHWFramesContext;
#endif /* AVUTIL_HWCONTEXT_H */

/* reference swscale.h:405 / swscale.c:1500.  Dynamic mode: (re)plans the conversion described by the
 * two frames (format, size, color_range, colorspace, chroma_location) with the context's flags.
 * Hardware frames: the reference accepts Vulkan frames only (swscale.c:1511-1538); this library
 * accepts AV_PIX_FMT_CUDA frames instead under the same rules -- both frames
 * carry a frames context, both are allocated, both live on the same CUDA device -- and converts
 * them in place in device memory, no PCIe transfer. */
int sws_frame_setup(SwsContext *ctx, const AVFrame *dst, const AVFrame *src);

/* reference swscale.h:415 / format.c:693 */
int sws_is_noop(const AVFrame *dst, const AVFrame *src);
/* reference swscale.h:392 / format.c:680-691: format, colour properties, range and siting all supported */
int sws_test_frame(const AVFrame *frame, int output);

/* reference swscale.h:439 / swscale.c:1405.  A dst without buffers is allocated like the reference does
 * (swscale.c:1316-1330,1437-1467) through the av_frame_get_buffer() of the libavutil loaded in the calling
 * process (looked up at run time: this library does not link libavutil); ENOTSUP if there is none.
 * Returns >= 0 or a negative AVERROR. */
int sws_scale_frame(SwsContext *c, AVFrame *dst, const AVFrame *src);

/* reference swscale.h:613-669 / swscale.c:1219-1403: the slice-wise frame API of a legacy context */
int sws_frame_start(SwsContext *c, AVFrame *dst, const AVFrame *src);
void sws_frame_end(SwsContext *c);
int sws_send_slice(SwsContext *c, unsigned int slice_start, unsigned int slice_height);
int sws_receive_slice(SwsContext *c, unsigned int slice_start, unsigned int slice_height);
unsigned int sws_receive_slice_alignment(const SwsContext *c);

#ifdef __cplusplus
}
#endif
#endif
