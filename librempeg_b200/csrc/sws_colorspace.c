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
 * (Y, V8); the kernels evaluate it directly.  tests/test_host_cpu.py
 * (test_rgb_closed_form_equals_reference_luts) checks the closed form against the LUTs the real reference
 * builds, entry by entry.
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

/* ---------------------------------------------------------------- RGB -> YUV
 * Nine 15-bit coefficients {ry,gy,by, ru,gu,bu, rv,gv,bv} for packed-RGB sources, derived from
 * the DESTINATION matrix exactly as the reference does (fill_rgb2yuv_table, utils.c:614-706):
 * the luma weights follow from inverting {crv,cbu,cgu,cgv}; always limited range (the full-range
 * case is handled by the range conversion of the h-scaled lines).  BT.601 uses the classic
 * rounded constants instead of the derived ones (utils.c:692-702). */
static int64_t rounded_div(int64_t a, int64_t b)
{
    return (a >= 0 ? a + (b >> 1) : a - (b >> 1)) / b;
}

void ff_b200_rgb2yuv_table(int32_t out[9], const int table[4])
{
    const int64_t one = 65536, sh = 1 << 15;
    const int64_t vr = table[0], ub = table[1], ug = -table[2], vg = -table[3];
    const int64_t cy = one * 255 / 219;
    const int64_t w = rounded_div(one * one * ug, ub);
    const int64_t v = rounded_div(one * one * vg, vr);
    const int64_t z = one * one - w - v;
    const int64_t Cy = rounded_div(cy * z, one);
    const int64_t Cu = rounded_div(ub * z, one);
    const int64_t Cv = rounded_div(vr * z, one);

    out[0] = (int32_t)-rounded_div(sh * v, Cy);
    out[1] = (int32_t) rounded_div(sh * one * one, Cy);
    out[2] = (int32_t)-rounded_div(sh * w, Cy);
    out[3] = (int32_t) rounded_div(sh * v, Cu);
    out[4] = (int32_t)-rounded_div(sh * one * one, Cu);
    out[5] = (int32_t) rounded_div(sh * (z + w), Cu);
    out[6] = (int32_t) rounded_div(sh * (v + z), Cv);
    out[7] = (int32_t)-rounded_div(sh * one * one, Cv);
    out[8] = (int32_t) rounded_div(sh * w, Cv);

    if (!memcmp(table, yuv2rgb_coeffs[SWS_CS_DEFAULT], sizeof(yuv2rgb_coeffs[0]))) {
        out[0] =  (int)(0.299 * 219 / 255 * (1 << 15) + 0.5);
        out[1] =  (int)(0.587 * 219 / 255 * (1 << 15) + 0.5);
        out[2] =  (int)(0.114 * 219 / 255 * (1 << 15) + 0.5);
        out[3] = -(int)(0.169 * 224 / 255 * (1 << 15) + 0.5);
        out[4] = -(int)(0.331 * 224 / 255 * (1 << 15) + 0.5);
        out[5] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        out[6] =  (int)(0.500 * 224 / 255 * (1 << 15) + 0.5);
        out[7] = -(int)(0.419 * 224 / 255 * (1 << 15) + 0.5);
        out[8] = -(int)(0.081 * 224 / 255 * (1 << 15) + 0.5);
    }
}
