/*
 * swscale_b200.h -- public C ABI of the B200-native libswscale hot path.
 *
 * This header mirrors the reference's libswscale/swscale.h for the legacy
 * (stateful) scaling API so that existing callers compile and link unchanged:
 * same type names, same SwsContext public field order (ABI per
 * libswscale/swscale.h:222-315), same flag values (swscale.h:131-208), same
 * function names and argument meaning.  Every entry point cites the reference
 * interface it replaces.  Only plain pointers and sizes cross this boundary;
 * no CUDA or torch types appear in any signature.
 *
 * The implementation (librempeg_b200/csrc) is host C + hand-written sm_100a
 * CUDA.  There is NO CPU fallback on the data path: if no CUDA device is
 * usable, context creation fails (sws_getContext() returns NULL,
 * sws_init_context() returns AVERROR(ENOSYS)); a conversion the CUDA path does
 * not implement fails with AVERROR(ENOTSUP)/NULL instead of silently running
 * on the CPU.
 */
#ifndef SWSCALE_B200_H
#define SWSCALE_B200_H

#include <stdint.h>
#include <stddef.h>
#include "swscale_b200_prefix.h"   /* no-op unless an in-tree build defines SWS_B200_PREFIX */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- pixel formats: numeric values are ABI (reference libavutil/pixfmt.h) ---- */
#ifndef AVUTIL_PIXFMT_H
enum AVPixelFormat {
    AV_PIX_FMT_NONE        = -1,
    AV_PIX_FMT_YUV420P     = 0,
    AV_PIX_FMT_RGB24       = 2,
    AV_PIX_FMT_BGR24       = 3,
    AV_PIX_FMT_YUV422P     = 4,
    AV_PIX_FMT_YUV444P     = 5,
    AV_PIX_FMT_GRAY8       = 8,
    AV_PIX_FMT_YUVJ420P    = 12,
    AV_PIX_FMT_YUVJ422P    = 13,
    AV_PIX_FMT_YUVJ444P    = 14,
    AV_PIX_FMT_NV12        = 23,
    AV_PIX_FMT_NV21        = 24,
    AV_PIX_FMT_ARGB        = 25,
    AV_PIX_FMT_RGBA        = 26,
    AV_PIX_FMT_ABGR        = 27,
    AV_PIX_FMT_BGRA        = 28,
    AV_PIX_FMT_RGB48LE     = 35,
    AV_PIX_FMT_RGB565LE    = 37,   /* output only */
    AV_PIX_FMT_RGB555LE    = 39,   /* output only */
    AV_PIX_FMT_BGR565LE    = 41,   /* output only */
    AV_PIX_FMT_BGR555LE    = 43,   /* output only */
    AV_PIX_FMT_YUV420P16LE = 45,
    AV_PIX_FMT_YUV422P16LE = 47,
    AV_PIX_FMT_YUV444P16LE = 49,
    AV_PIX_FMT_BGR48LE     = 58,
    AV_PIX_FMT_YUV420P9LE  = 60,
    AV_PIX_FMT_YUV420P10LE = 62,
    AV_PIX_FMT_YUV422P10LE = 64,
    AV_PIX_FMT_YUV444P9LE  = 66,
    AV_PIX_FMT_YUV444P10LE = 68,
    AV_PIX_FMT_YUV422P9LE  = 70,
    AV_PIX_FMT_CUDA        = 117,
    AV_PIX_FMT_YUV420P12LE = 123,
    AV_PIX_FMT_YUV420P14LE = 125,
    AV_PIX_FMT_YUV422P12LE = 127,
    AV_PIX_FMT_YUV422P14LE = 129,
    AV_PIX_FMT_YUV444P12LE = 131,
    AV_PIX_FMT_YUV444P14LE = 133,
    AV_PIX_FMT_P010LE      = 158,
    AV_PIX_FMT_GBRPF32LE   = 175,   /* output only */
    AV_PIX_FMT_GRAYF32LE   = 183,   /* output only */
};
#endif

/* AVERROR codes are negated errno values (reference libavutil/error.h) */
#ifndef AVERROR
#define AVERROR(e) (-(e))
#endif
#ifndef AVERROR_PATCHWELCOME
#define AVERROR_PATCHWELCOME (-(int)(('P') | ('A' << 8) | ('W' << 16) | ((unsigned)'E' << 24)))
#endif

/* ---- enums and flags: reference libswscale/swscale.h:77-216 ---- */
typedef enum SwsDither {
    SWS_DITHER_NONE = 0,
    SWS_DITHER_AUTO,
    SWS_DITHER_BAYER,
    SWS_DITHER_ED,
    SWS_DITHER_A_DITHER,
    SWS_DITHER_X_DITHER,
    SWS_DITHER_NB,
    SWS_DITHER_MAX_ENUM = 0x7FFFFFFF,
} SwsDither;

typedef enum SwsAlphaBlend {
    SWS_ALPHA_BLEND_NONE = 0,
    SWS_ALPHA_BLEND_UNIFORM,
    SWS_ALPHA_BLEND_CHECKERBOARD,
    SWS_ALPHA_BLEND_NB,
    SWS_ALPHA_BLEND_MAX_ENUM = 0x7FFFFFFF,
} SwsAlphaBlend;

typedef enum SwsScaler {
    SWS_SCALE_AUTO = 0,
    SWS_SCALE_BILINEAR,
    SWS_SCALE_BICUBIC,
    SWS_SCALE_POINT,
    SWS_SCALE_AREA,
    SWS_SCALE_GAUSSIAN,
    SWS_SCALE_SINC,
    SWS_SCALE_LANCZOS,
    SWS_SCALE_SPLINE,
    SWS_SCALE_NB,
    SWS_SCALE_MAX_ENUM = 0x7FFFFFFF,
} SwsScaler;

typedef enum SwsBackend {
    SWS_BACKEND_LEGACY   = (1 << 0),
    SWS_BACKEND_STABLE   = SWS_BACKEND_LEGACY,
    SWS_BACKEND_MAX_ENUM = 0x7FFFFFFF,
} SwsBackend;

typedef enum SwsFlags {
    SWS_FAST_BILINEAR  = 1 <<  0,
    SWS_BILINEAR       = 1 <<  1,
    SWS_BICUBIC        = 1 <<  2,
    SWS_X              = 1 <<  3,
    SWS_POINT          = 1 <<  4,
    SWS_AREA           = 1 <<  5,
    SWS_BICUBLIN       = 1 <<  6,
    SWS_GAUSS          = 1 <<  7,
    SWS_SINC           = 1 <<  8,
    SWS_LANCZOS        = 1 <<  9,
    SWS_SPLINE         = 1 << 10,
    SWS_STRICT         = 1 << 11,
    SWS_PRINT_INFO     = 1 << 12,
    SWS_FULL_CHR_H_INT = 1 << 13,
    SWS_FULL_CHR_H_INP = 1 << 14,
    SWS_DIRECT_BGR     = 1 << 15,
    SWS_ACCURATE_RND   = 1 << 18,
    SWS_BITEXACT       = 1 << 19,
    SWS_UNSTABLE       = 1 << 20,
    SWS_ERROR_DIFFUSION = 1 << 23,
} SwsFlags;

#define SWS_SRC_V_CHR_DROP_MASK  0x30000
#define SWS_SRC_V_CHR_DROP_SHIFT 16
#define SWS_PARAM_DEFAULT        123456
#define SWS_MAX_REDUCE_CUTOFF    0.002

#define SWS_CS_ITU709     1
#define SWS_CS_FCC        4
#define SWS_CS_ITU601     5
#define SWS_CS_ITU624     5
#define SWS_CS_SMPTE170M  5
#define SWS_CS_SMPTE240M  7
#define SWS_CS_DEFAULT    5
#define SWS_CS_BT2020     9

struct AVClass;

/* Public context: field order is ABI, sizeof is not (reference swscale.h:227-315). */
typedef struct SwsContext {
    const struct AVClass *av_class;
    void *opaque;
    unsigned flags;
#define SWS_NUM_SCALER_PARAMS 2
    double scaler_params[SWS_NUM_SCALER_PARAMS];
    int threads;
    SwsDither dither;
    SwsAlphaBlend alpha_blend;
    int gamma_flag;
    int src_w, src_h;
    int dst_w, dst_h;
    int src_format;
    int dst_format;
    int src_range;
    int dst_range;
    int src_v_chr_pos;
    int src_h_chr_pos;
    int dst_v_chr_pos;
    int dst_h_chr_pos;
    int intent;
    SwsScaler scaler;
    SwsScaler scaler_sub;
    SwsBackend backends;
} SwsContext;

typedef struct SwsVector {
    double *coeff;
    int length;
} SwsVector;

typedef struct SwsFilter {
    SwsVector *lumH;
    SwsVector *lumV;
    SwsVector *chrH;
    SwsVector *chrV;
} SwsFilter;

/* ---- library identification: reference swscale.h:53-71, version.c ---- */
unsigned    swscale_version(void);
const char *swscale_configuration(void);
const char *swscale_license(void);
const struct AVClass *sws_get_class(void);

/* ---- context life cycle ---- */
/* reference swscale.h:320 / utils.c:1032 */
SwsContext *sws_alloc_context(void);
/* reference swscale.h:326 / utils.c:2313 */
void sws_free_context(SwsContext **ctx);
/* reference swscale.h:522 / utils.c:1884.  srcFilter/dstFilter must be NULL (ENOTSUP otherwise). */
int sws_init_context(SwsContext *ctx, SwsFilter *srcFilter, SwsFilter *dstFilter);
/* reference swscale.h:528 / utils.c:2250 */
void sws_freeContext(SwsContext *ctx);
/* reference swscale.h:551 / utils.c:1919 */
SwsContext *sws_getContext(int srcW, int srcH, enum AVPixelFormat srcFormat,
                           int dstW, int dstH, enum AVPixelFormat dstFormat,
                           int flags, SwsFilter *srcFilter,
                           SwsFilter *dstFilter, const double *param);
/* reference swscale.h:737 / utils.c:2331 */
SwsContext *sws_getCachedContext(SwsContext *context, int srcW, int srcH,
                                 enum AVPixelFormat srcFormat, int dstW, int dstH,
                                 enum AVPixelFormat dstFormat, int flags,
                                 SwsFilter *srcFilter, SwsFilter *dstFilter,
                                 const double *param);

/* ---- the hot entry point: reference swscale.h:583 / swscale.c:1626 ----
 * Host pointers in, host pointers out; synchronous; returns the height of the
 * output slice (>= 0) or a negative AVERROR. Slices must arrive top to bottom. */
int sws_scale(SwsContext *c, const uint8_t *const srcSlice[],
              const int srcStride[], int srcSliceY, int srcSliceH,
              uint8_t *const dst[], const int dstStride[]);

/* ---- colourspace: reference swscale.h:474,684,692 / yuv2rgb.c:61, utils.c:849,1007 ---- */
const int *sws_getCoefficients(int colorspace);
int sws_setColorspaceDetails(SwsContext *c, const int inv_table[4], int srcRange,
                             const int table[4], int dstRange,
                             int brightness, int contrast, int saturation);
int sws_getColorspaceDetails(SwsContext *c, int **inv_table, int *srcRange,
                             int **table, int *dstRange,
                             int *brightness, int *contrast, int *saturation);

/* ---- format queries: reference swscale.h:342,352,495,501,508 / format.c ---- */
int sws_isSupportedInput(enum AVPixelFormat pix_fmt);
int sws_isSupportedOutput(enum AVPixelFormat pix_fmt);
int sws_isSupportedEndiannessConversion(enum AVPixelFormat pix_fmt);
int sws_test_format(enum AVPixelFormat format, int output);
int sws_test_hw_format(enum AVPixelFormat format);
/* colour-property queries: reference swscale.h:363,374,385 / format.c:627-656.  The arguments are
 * libavutil's AVColorSpace / AVColorPrimaries / AVColorTransferCharacteristic values (ints here;
 * include/swscale_b200_frame.h names the AVColorSpace ones). */
int sws_test_colorspace(int colorspace, int output);
int sws_test_primaries(int primaries, int output);
int sws_test_transfer(int trc, int output);

/* ---- SwsVector helpers: reference swscale.h:699-717 / utils.c:1956-2248 ---- */
SwsVector *sws_allocVec(int length);
SwsVector *sws_getGaussianVec(double variance, double quality);
void sws_scaleVec(SwsVector *a, double scalar);
void sws_normalizeVec(SwsVector *a, double height);
void sws_freeVec(SwsVector *a);
/* SwsFilter builder: reference swscale.h:719-723 / utils.c:2155-2248.  sws_init_context() convolves the
 * source-side vectors into the FIR banks exactly like initFilter (utils.c:385-413; like the reference it only
 * widens the rows for the destination-side vectors). */
SwsFilter *sws_getDefaultFilter(float lumaGBlur, float chromaGBlur, float lumaSharpen, float chromaSharpen,
                                float chromaHShift, float chromaVShift, int verbose);
void sws_freeFilter(SwsFilter *filter);

/* ---- palette helpers: reference swscale.h:753,765 / swscale_unscaled.c:2733-2760 ---- */
void sws_convertPalette8ToPacked32(const uint8_t *src, uint8_t *dst, int num_pixels, const uint8_t *palette);
void sws_convertPalette8ToPacked24(const uint8_t *src, uint8_t *dst, int num_pixels, const uint8_t *palette);

#ifdef __cplusplus
}
#endif
#endif /* SWSCALE_B200_H */
