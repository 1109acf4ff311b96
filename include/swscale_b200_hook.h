/*
 * swscale_b200_hook.h -- the entry points the in-tree hook binds.
 *
 * integration/swscale_cuda.c -- the ff_sws_init_swscale_cuda() a maintainer adds next to
 * ff_sws_init_swscale_x86() & co. (reference libswscale/swscale.c:697-714,
 * swscale_internal.h:1034-1040) -- is compiled against the REFERENCE's headers, so it cannot include
 * swscale_b200.h (both define SwsContext, SwsFlags, ...).  This header is the whole of what it needs:
 * plain ints, pointers and one POD struct, no type of either library.
 *
 * Every function replaces, for a hooked context, one step of the reference:
 *   sws_b200_hook_open        the kernel selection of ff_sws_init_scale()/ff_get_unscaled_swscale()
 *                             (swscale.c:697, swscale_unscaled.c:2392)
 *   sws_b200_hook_colorspace  the table rebuild of sws_setColorspaceDetails() (utils.c:849-1005)
 *   sws_b200_hook_bank        read-back of the FIR banks, so the hook can refuse to install itself when they
 *                             differ from the reference's c->hLumFilter & co. (e.g. a caller-supplied SwsFilter)
 *   sws_b200_hook_scale       ff_swscale() / c->convert_unscaled() for a source slice (swscale.c:1163-1192)
 *   sws_b200_hook_scale_rows  the same for a destination slice (sws_receive_slice(), swscale.c:371-375)
 *   sws_b200_hook_close       the frees of sws_freeContext() (utils.c:2250)
 */
#ifndef SWSCALE_B200_HOOK_H
#define SWSCALE_B200_HOOK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The fields of the reference's public SwsContext (swscale.h:227-315) that describe a conversion;
 * enums travel as ints (the numeric values are ABI in both libraries). */
typedef struct SwsB200HookParams {
    unsigned flags;
    double scaler_params[2];
    int dither, alpha_blend, gamma_flag;
    int src_w, src_h, dst_w, dst_h;
    int src_format, dst_format;          /* enum AVPixelFormat */
    int src_range, dst_range;
    int src_v_chr_pos, src_h_chr_pos, dst_v_chr_pos, dst_h_chr_pos;
    int scaler, scaler_sub;              /* SwsScaler */
} SwsB200HookParams;

/* NULL + *err < 0 when the conversion is not on the CUDA path (or no device): keep the C kernels. */
void *sws_b200_hook_open(const SwsB200HookParams *p, int *err);
void  sws_b200_hook_close(void *h);
int   sws_b200_hook_colorspace(void *h, const int inv_table[4], int srcRange, const int table[4], int dstRange,
                               int brightness, int contrast, int saturation);
/* which: 0 hLum, 1 hChr, 2 vLum, 3 vChr; returns taps per output (coef is [len][taps]) or < 0 */
int   sws_b200_hook_bank(void *h, int which, const int16_t **coef, const int32_t **pos, int *len);
/* 1 when rows map 1:1 like the reference's unscaled converters */
int   sws_b200_hook_is_unscaled(void *h);
int   sws_b200_hook_dst_slice_align(void *h);
/* same arguments and return value as sws_scale() (swscale.h:583); strides may be negative */
int   sws_b200_hook_scale(void *h, const uint8_t *const src[4], const int srcStride[4], int srcSliceY, int srcSliceH,
                          uint8_t *const dst[4], const int dstStride[4]);
/* whole source frame at src[], destination rows [dstY, dstY+dstH) stored at dst[] (= first row of the slice) */
int   sws_b200_hook_scale_rows(void *h, const uint8_t *const src[4], const int srcStride[4],
                               uint8_t *const dst[4], const int dstStride[4], int dstY, int dstH);
long  sws_b200_hook_launches(void *h);
const char *sws_b200_hook_kernel(void *h);
const char *sws_b200_hook_error(void *h);

#ifdef __cplusplus
}
#endif
#endif
