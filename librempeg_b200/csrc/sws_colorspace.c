/*
 * sws_colorspace.c -- YUV->RGB constants for the kernels.
 *
 * The reference turns {crv,cbu,cgu,cgv} + range + brightness/contrast/saturation
 * into byte LUTs (ff_yuv2rgb_c_init_tables, libswscale/yuv2rgb.c:717-973) and
 * into six 13-bit coefficients for the 16-bit arithmetic path
 * (yuv2rgb.c:786-791).  Random byte-LUT gathers are the wrong shape for a GPU
 * (bank-conflicted shared-memory loads would cap throughput far below HBM), so
 * we keep the LUTs in closed form instead.  For 24/48 bpp the reference fills
 *
 *     y_table[i]  = clip_u8((yb0 + i*cy + 0x8000) >> 16)          yuv2rgb.c:905-909
 *     table_rV[k] = &y_table[yoffs - (crv>>9) + ((clip_u8(k-512)*crv) >> 16)]
 *     table_gU/bU likewise, table_gV[k] = -(cgv>>9) + ((clip_u8(k-512)*cgv)>>16)
 *                                                                  yuv2rgb.c:680-703
 *
 * so r = y_table[Y + base_r + ((V8*crv)>>16)] is an exact integer expression in
 * (Y, V8); the kernels evaluate it directly.  tests/test_host_tables.py checks
 * the closed form against the LUTs the real reference builds, entry by entry.
 */
#include <errno.h>
#include <string.h>

#include "sws_internal.h"

/* ITU/SMPTE matrix constants, 16.16 (public standard values; same table layout
 * as sws_getCoefficients() in the reference, yuv2rgb.c:47-66). */
static const int yuv2rgb_coeffs[11][4] = {
    { 104597, 132201, 25675, 53279 }, /* 0: default (BT.601)      */
    { 117489, 138438, 13975, 34925 }, /* 1: BT.709                */
    { 104597, 132201, 25675, 53279 }, /* 2: unspecified           */
    { 104597, 132201, 25675, 53279 }, /* 3: reserved              */
    { 104448, 132798, 24759, 53109 }, /* 4: FCC                   */
    { 104597, 132201, 25675, 53279 }, /* 5: BT.470BG / BT.601     */
    { 104597, 132201, 25675, 53279 }, /* 6: SMPTE 170M            */
    { 117579, 136230, 16907, 35559 }, /* 7: SMPTE 240M            */
    {      0,      0,     0,     0 }, /* 8: YCgCo (unsupported)   */
    { 110013, 140363, 12277, 42626 }, /* 9: BT.2020 NCL           */
    { 110013, 140363, 12277, 42626 }, /* 10: BT.2020 CL           */
};

const int *sws_getCoefficients(int colorspace)
{
    if (colorspace > 10 || colorspace < 0 || colorspace == 8)
        colorspace = SWS_CS_DEFAULT;
    return yuv2rgb_coeffs[colorspace];
}

static int16_t round_q16_to_i16(int64_t f)
{
    int r = (int)((f + (1 << 15)) >> 16);
    if (r < -0x7FFF)
        return (int16_t)0x8000;
    if (r > 0x7FFF)
        return 0x7FFF;
    return (int16_t)r;
}

static int fits_i32(int64_t v) { return v >= INT32_MIN && v <= INT32_MAX; }

int ff_b200_rgb_consts(SwsRgbConsts *k, const int inv_table[4], int full_range,
                       int brightness, int contrast, int saturation)
{
    const int luma_headroom = 512;                       /* YUVRGB_TABLE_LUMA_HEADROOM */
    const int yoffs = (full_range ? 384 : 326) + luma_headroom;
    int64_t crv =  inv_table[0];
    int64_t cbu =  inv_table[1];
    int64_t cgu = -inv_table[2];
    int64_t cgv = -inv_table[3];
    int64_t cy  = 1 << 16;
    int64_t oy  = 0;
    int64_t yb, div;

    memset(k, 0, sizeof(*k));

    if (!full_range) {
        cy = (cy * 255) / 219;
        oy = 16 << 16;
    } else {
        crv = (crv * 224) / 255;
        cbu = (cbu * 224) / 255;
        cgu = (cgu * 224) / 255;
        cgv = (cgv * 224) / 255;
    }

    cy  = (cy  * contrast)              >> 16;
    crv = (crv * contrast * saturation) >> 32;
    cbu = (cbu * contrast * saturation) >> 32;
    cgu = (cgu * contrast * saturation) >> 32;
    cgv = (cgv * contrast * saturation) >> 32;
    oy -= 256LL * brightness;

    /* 16-bit arithmetic path */
    k->y_coeff  = round_q16_to_i16(cy  * (1 << 13));
    k->y_offset = round_q16_to_i16(oy  * (1 <<  9));
    k->v2r      = round_q16_to_i16(crv * (1 << 13));
    k->v2g      = round_q16_to_i16(cgv * (1 << 13));
    k->u2g      = round_q16_to_i16(cgu * (1 << 13));
    k->u2b      = round_q16_to_i16(cbu * (1 << 13));

    /* 8-bit LUT path: chroma slopes are re-expressed in LUT-index units */
    div = cy > 1 ? cy : 1;
    crv = ((crv * (1 << 16)) + 0x8000) / div;
    cbu = ((cbu * (1 << 16)) + 0x8000) / div;
    cgu = ((cgu * (1 << 16)) + 0x8000) / div;
    cgv = ((cgv * (1 << 16)) + 0x8000) / div;

    yb = -(384 << 16) - luma_headroom * cy - oy + 0x8000;

    /* the kernels evaluate yb + i*cy (0 <= i < 2048) and V8*slope in int32 */
    if (!fits_i32(cy) || !fits_i32(yb) || !fits_i32(yb + 2048 * cy) ||
        !fits_i32(255 * crv) || !fits_i32(255 * cbu) ||
        !fits_i32(255 * cgu) || !fits_i32(255 * cgv))
        return AVERROR(ENOTSUP);

    k->cy  = (int32_t)cy;
    k->yb  = (int32_t)yb;
    k->crv = (int32_t)crv;
    k->cbu = (int32_t)cbu;
    k->cgu = (int32_t)cgu;
    k->cgv = (int32_t)cgv;
    k->base_r = yoffs - (int32_t)(crv >> 9);
    k->base_b = yoffs - (int32_t)(cbu >> 9);
    k->base_g = yoffs - (int32_t)(cgu >> 9) - (int32_t)(cgv >> 9);
    return 0;
}
