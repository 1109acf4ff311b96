/*
 * sws_internal.h -- private state of the B200 libswscale hot path.
 *
 * Plays the role of the reference's SwsInternal (libswscale/swscale_internal.h:337-706)
 * but holds only what a frame-level GPU converter needs: geometry, the four
 * FIR banks, colour constants and an opaque device state.  The per-line
 * function pointers, ring-buffer slices and descriptor chains of the reference
 * (swscale_internal.h:563-672, slice.c) have no equivalent: a whole frame is
 * resident in HBM and one fused kernel replaces ff_swscale()'s line-pull loop.
 */
#ifndef SWS_B200_INTERNAL_H
#define SWS_B200_INTERNAL_H

#include <stdint.h>
#include "swscale_b200.h"
#include "swscale_b200_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ pixfmt */
#define SWSPF_RGB     1   /* packed RGB family                        */
#define SWSPF_PLANAR  2   /* one plane per component                  */
#define SWSPF_SEMI    4   /* luma plane + interleaved chroma plane    */
#define SWSPF_GRAY    8
#define SWSPF_JPEG   16   /* deprecated yuvj*: implies full range     */

typedef struct SwsPixDesc {
    int fmt;
    const char *name;
    int flags;
    int depth;        /* bits per component                         */
    int log2_cw;      /* horizontal chroma shift                    */
    int log2_ch;      /* vertical chroma shift                      */
    int bpp;          /* bits per pixel (av_get_bits_per_pixel)     */
    int nb_planes;
    int swap_uv;      /* nv21                                        */
    int as_input, as_output;
    int shift;        /* samples sit in the high bits of their 16-bit container (p010: 6) */
} SwsPixDesc;

const SwsPixDesc *ff_b200_pix_desc(int fmt);

/* ------------------------------------------------------------------ filters */
typedef struct SwsFirBank {
    int16_t *coef;   /* [len][size]                               */
    int32_t *pos;    /* [len] first source sample of each output  */
    int size;        /* taps per output                           */
    int len;         /* number of outputs                         */
} SwsFirBank;

typedef struct SwsFirSpec {
    int64_t inc;           /* 16.16 source step per output sample                */
    int src_len, dst_len;
    int one;               /* fixed-point 1.0 of the stored coefficients         */
    int scaler;            /* one SWS_* scaler flag                              */
    unsigned flags;        /* full flag word (BITEXACT / ACCURATE_RND matter)    */
    double param[2];
    int src_pos, dst_pos;  /* chroma siting, 1/256 sample units, already local   */
    /* SwsFilter vectors convolved into every row (utils.c:385-413); NULL = none.  Only the length of the
     * destination-side vector matters: the reference widens the rows for it but never applies it */
    const double *src_vec; int src_vec_len;
    int dst_vec_len;
} SwsFirSpec;

#define SWS_B200_USE_CASCADE (-12345)
#define SWS_B200_MAX_LANES 64

int  ff_b200_build_fir(SwsFirBank *out, const SwsFirSpec *spec);
void ff_b200_free_fir(SwsFirBank *b);

/* --------------------------------------------------------------- colourspace */
typedef struct SwsRgbConsts {
    /* closed form of the 8-bit LUT chain (reference yuv2rgb.c:680-703,901-914) */
    int32_t cy;        /* luma slope, 16.16                                   */
    int32_t yb;        /* LUT intercept incl. rounding: value(i)=(yb+i*cy)>>16 */
    int32_t crv, cbu, cgu, cgv;         /* chroma slopes rescaled by cy        */
    int32_t base_r, base_g, base_b;     /* LUT index bases (g: gU base + gV base) */
    /* 16-bit arithmetic path (reference yuv2rgb.c:786-791, output.c:1115-1196) */
    int32_t y_offset, y_coeff, v2r, v2g, u2g, u2b;
} SwsRgbConsts;

int ff_b200_rgb_consts(SwsRgbConsts *k, const int inv_table[4], int full_range,
                       int brightness, int contrast, int saturation);
void ff_b200_rgb2yuv_table(int32_t out[9], const int table[4]);

/* ------------------------------------------------------------ device plan */
enum {
    SWSC_SRC_PLANAR = 0,   /* Y, U, V planes (or Y only for gray)     */
    SWSC_SRC_NV12   = 1,   /* Y plane + interleaved UV                */
    SWSC_SRC_NV21   = 2,   /* Y plane + interleaved VU                */
    SWSC_SRC_RGB    = 3,   /* one packed plane of 8-bit R,G,B[,A] pixels */
};

enum {
    SWSC_DST_PLANAR8 = 0,
    SWSC_DST_PLANARN,      /* 9..14 bit little-endian                  */
    SWSC_DST_PLANAR16,
    SWSC_DST_NV12,
    SWSC_DST_NV21,
    SWSC_DST_P010,         /* semi-planar, 10 bits in the high bits of 16  */
    SWSC_DST_RGB24,
    SWSC_DST_BGR24,
    SWSC_DST_RGBA,
    SWSC_DST_BGRA,
    SWSC_DST_ARGB,
    SWSC_DST_ABGR,
    SWSC_DST_RGB48,
    SWSC_DST_BGR48,
    SWSC_DST_RGB565,       /* 16 bpp, ordered 2x2 dither (output.c:1714-1747) */
    SWSC_DST_BGR565,
    SWSC_DST_RGB555,
    SWSC_DST_BGR555,
    SWSC_DST_RGBA64,       /* 16-bit packed RGB with alpha (output.c:1115-1196, hasAlpha) */
    SWSC_DST_BGRA64,
    SWSC_DST_GBRP,         /* planar RGB, plane order G, B, R [, A] (output.c:2378-2610) */
    SWSC_DST_PLANARF32,    /* 32-bit float planes (output.c:219-316) */
};

/* unscaled converters the reference installs instead of the scaler (swscale_unscaled.c) */
enum {
    SWSC_SPECIAL_NONE = 0,
    SWSC_SPECIAL_SHUFFLE,        /* rgbToRgbWrapper / packedCopyWrapper between 8-bit packed RGB layouts */
    SWSC_SPECIAL_BGR24_YV12,     /* bgr24ToYv12Wrapper: 2x2 box chroma, truncating 15-bit matrix */
    SWSC_SPECIAL_COPY8,          /* planarCopyWrapper / planarToNv12Wrapper / nv12ToPlanarWrapper, 8-bit */
    SWSC_SPECIAL_P01X,           /* planarToP01xWrapper / planar8ToP01xleWrapper: shift + chroma interleave */
    SWSC_SPECIAL_DEPTHCOPY,      /* planarCopyWrapper between planar YUV depths (dithered down, replicated up) */
    SWSC_SPECIAL_RGB16PACK,      /* rgb24to16 / rgb32tobgr15 & co.: 8-bit packed RGB truncated into 15/16 bpp */
    SWSC_SPECIAL_RGB48,          /* packedCopyWrapper / rgb48tobgr48_nobswap between rgb48le and bgr48le */
};

/* POD description of one conversion; passed by value to the kernels. */
typedef struct SwsCudaPlan {
    int src_w, src_h, dst_w, dst_h;
    int chr_src_w, chr_src_h, chr_dst_w, chr_dst_h;
    int chr_src_hsub, chr_src_vsub, chr_dst_hsub, chr_dst_vsub;
    int src_layout, dst_kind;
    int src_bits, dst_bits;      /* component depth                           */
    int src_shift;               /* right shift of 16-bit source samples (p010: 6) */
    int dst_shift;               /* left shift of 16-bit destination samples (p010: 6) */
    int inter_bits;              /* 15 or 19: width of the h-scaled lines     */
    int h_shift;                 /* right shift applied after the H FIR       */
    int has_chroma;              /* 0 for gray sources                        */
    int dst_has_chroma;          /* 0 for gray destinations                   */
    int src_alpha, dst_alpha;    /* an alpha plane / channel is read / written through the scaler */
    int src_ao;                  /* packed RGB32 sources: byte offset of A inside a pixel */
    int unscaled_lut;            /* 1: reference would take convert_unscaled  */
    int special;                 /* whole-frame special converter, SWSC_SPECIAL_* */
    int shuf_map[4];             /* SHUFFLE: source byte of every destination byte, 4 = constant 255 */
    int dst_bpp;                 /* SHUFFLE: destination pixel stride in bytes */
    int full_chr;                /* 1: SWS_FULL_CHR_H_INT packed RGB (per-pixel chroma, arithmetic) */
    int dither_bayer;            /* 1: ff_dither_8x8_128 rows, 0: constant 64 */
    /* range conversion on the h-scaled lines (reference swscale.c:163-255,577-660) */
    int range_mode;              /* 0 none, 1 to-jpeg, 2 from-jpeg            */
    int src_full_range;          /* sws->src_range, read by the depth-copy converter at call time */
    int dither_none;             /* sws->dither == SWS_DITHER_NONE            */
    uint32_t lum_rc_coeff, chr_rc_coeff;
    int64_t  lum_rc_offset, chr_rc_offset;
    SwsRgbConsts rgb;
    /* packed-RGB sources (reference input.c:264-345,1068-1180): pixel stride, byte offsets of R,G,B (48-bit pixels,
     * src_bpp = 6: component indices; readers input.c:111-196),
     * chroma taken from horizontally summed pixel pairs (the *_half readers), and the 15-bit matrix */
    int src_bpp, src_ro, src_go, src_bo, src_rgb_half;
    int32_t rgb2yuv[9];
    /* device pointers to the four FIR banks */
    const int16_t *hl_coef, *hc_coef, *vl_coef, *vc_coef;
    const int32_t *hl_pos,  *hc_pos,  *vl_pos,  *vc_pos;
    int hl_size, hc_size, vl_size, vc_size;
    /* fast-path eligibility, decided on the host at init */
    int lum_identity;            /* h and v luma FIRs are the identity        */
    int chr_h_identity;          /* horizontal chroma FIR is the identity     */
    int chr_v_identity;          /* vertical chroma FIR is the identity       */
} SwsCudaPlan;

typedef struct SwsCudaState SwsCudaState;

/* C-ABI shim implemented in sws_cuda.cu (the only translation unit that sees CUDA). */
int  ff_b200_cuda_probe(void);    /* >=0: device ordinal in use, <0: AVERROR       */
int  ff_b200_cuda_create(SwsCudaState **st, SwsCudaPlan *plan,
                         const SwsFirBank *hl, const SwsFirBank *hc,
                         const SwsFirBank *vl, const SwsFirBank *vc);
/* the same on an explicit device ordinal (the caller's current device is left alone) */
int  ff_b200_cuda_create_on(int device, SwsCudaState **st, SwsCudaPlan *plan,
                            const SwsFirBank *hl, const SwsFirBank *hc,
                            const SwsFirBank *vl, const SwsFirBank *vc);
int  ff_b200_cuda_device_of(SwsCudaState *st);
void ff_b200_cuda_destroy(SwsCudaState *st);
int  ff_b200_cuda_update_plan(SwsCudaState *st, const SwsCudaPlan *plan);
/* device-resident frames; rows [y0,y1) of every frame; async on the context stream */
int  ff_b200_cuda_launch(SwsCudaState *st,
                         const uint8_t *const src[4], const int src_stride[4], const int64_t src_fstride[4],
                         uint8_t *const dst[4], const int dst_stride[4], const int64_t dst_fstride[4],
                         int nb_frames, int y0, int y1);
/* host frame in, host rows [y0,y1) out; synchronous */
#define SWS_MEM_SRC_DEVICE 1   /* src[] are device pointers to whole planes */
#define SWS_MEM_DST_DEVICE 2   /* dst[] are device pointers to whole planes */
#define SWS_MEM_DST_FLIPPED 4  /* bottom-up frame: dst[1], dst[2] address row (dst_h >> vsub) - 1 (swscale.c:1153-1154), so with an
                                * odd dst_h the last chroma row has no memory behind it: it is converted but never stored */
int  ff_b200_cuda_scale_host(SwsCudaState *st,
                             const uint8_t *const src[4], const int src_stride[4],
                             int src_y, int src_h, int upload,
                             uint8_t *const dst[4], const int dst_stride[4], int y0, int y1, int mem);
int  ff_b200_cuda_hwctx_device(void *cuda_ctx);
int  ff_b200_cuda_current_device(void);
void ff_b200_cuda_use_device(int dev);
int  ff_b200_cuda_wait_stream(SwsCudaState *st, void *producer_stream);
void *ff_b200_cuda_alloc(size_t bytes);
void ff_b200_cuda_free(void *p);
/* page-locked host frames first, first+step, ... through a ring of staging sets; enqueue only, then wait */
int  ff_b200_cuda_frames_enqueue(SwsCudaState *st,
                                 const uint8_t *const src[4], const int src_stride[4], const int64_t src_fstride[4],
                                 uint8_t *const dst[4], const int dst_stride[4], const int64_t dst_fstride[4],
                                 int first, int step, int nb_frames);
int  ff_b200_cuda_frames_wait(SwsCudaState *st);
int  ff_b200_cuda_frame_is_pinned(SwsCudaState *st, const uint8_t *const planes[4], int dst_side);
int  ff_b200_cuda_sync(SwsCudaState *st);
void *ff_b200_cuda_stream(SwsCudaState *st);
long ff_b200_cuda_launch_count(SwsCudaState *st);
const char *ff_b200_cuda_kernel_name(SwsCudaState *st);

/* ----------------------------------------------------------------- context */
typedef struct SwsInternal {
    SwsContext opts;                 /* must be first (reference swscale_internal.h:79-82) */
    int initialized;
    int refused;                     /* a post-init option change was refused: scaling fails until it is undone */
    int planned;                     /* tables built by sws_b200_plan_only() (no device) */
    int src_colorspace[4], dst_colorspace[4];
    int brightness, contrast, saturation;
    int colorspace_set;
    int chr_src_hsub, chr_src_vsub, chr_dst_hsub, chr_dst_vsub;
    int chr_src_w, chr_src_h, chr_dst_w, chr_dst_h;
    int src_bpc, dst_bpc;
    int unscaled_lut;                /* reference would use c->convert_unscaled (a13) */
    int special;                     /* SWSC_SPECIAL_*: rows map 1:1 like the unscaled LUT converter */
    int dst_slice_align;
    SwsFirBank h_lum, h_chr, v_lum, v_chr;
    const SwsFilter *src_filter_tmp, *dst_filter_tmp;   /* caller-owned, only read during init (swscale.h:699-723) */
    SwsCudaPlan plan;
    SwsCudaState *cuda;
    /* slice state of the legacy API (reference swscale.c:296-298,562-564) */
    int dst_y;
    int slice_dir;
    int rows_received;
    /* AVFrame entry points (sws_frame.c) */
    SwsContext *dyn;                 /* inner legacy context of the dynamic sws_scale_frame() mode */
    void *dyn_key;                   /* description it was planned for */
    const void *frame_src, *frame_dst;   /* sws_frame_start() .. sws_frame_end() */
    int frame_rows_sent, frame_uploaded;
    /* cascaded contexts (reference utils.c:1803-1832 filters too long for one pass; utils.c:915-984 YUV -> YUV with
     * two different matrices goes through RGB): two inner contexts and an intermediate picture kept in HBM */
    SwsContext *cascade[2];
    int cascade_main;                /* which inner context sws_setColorspaceDetails() is forwarded to */
    uint8_t *cascade_tmp[4];
    int cascade_tmp_stride[4];
    int cascade_frame_done;          /* the first stage has consumed the whole source frame */
    /* lanes of sws_cuda_scale_batch_host(): extra device states (own stream, own staging) spread over the
     * visible devices so that several host frames are in flight at once */
    SwsCudaState *lanes[SWS_B200_MAX_LANES];
    int nb_lanes, lanes_devices, lanes_depth;
    char last_error[256];
} SwsInternal;

static inline SwsInternal *sws_internal(const SwsContext *s) { return (SwsInternal *)s; }

void ff_b200_option_defaults(SwsContext *s);
/* NUMA placement of page-locked host memory and of the calling thread (sws_numa.c) */
int  ff_b200_numa_node_of_pci(const char *bus_id);
int  ff_b200_numa_prefer(int node);
void ff_b200_numa_restore(void);
int  ff_b200_numa_bind_thread(int node);
int ff_b200_scale_frame_rows(SwsInternal *c, const uint8_t *const src[4], const int srcStride[4], int upload,
                             uint8_t *const dst[4], const int dstStride[4], int y0, int y1);

#ifdef __cplusplus
}
#endif
#endif
