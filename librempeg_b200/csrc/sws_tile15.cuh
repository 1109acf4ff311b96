/*
 * sws_tile15.cuh -- the general scaling kernel for 15-bit intermediate lines: every conversion
 * whose destination is at most 14 bits deep (8-bit planar / semi-planar, 9..14-bit planar, packed
 * 8-bit RGB with shared chroma) and whose filters have at most 16 horizontal taps, that is not
 * served by one of the specialised kernels.  Sources: 8-bit planar / nv12 / nv21, 9..16-bit planar,
 * packed 8-bit RGB.  It replaces the table-driven sws_generic_tile_kernel on these paths (which
 * stays as the fallback) with the same arithmetic at a fraction of the instructions:
 *
 *   readers   nv12ToUV_c, rgb24ToY/UV[_half]_c, rgb16_32To*_c_template   libswscale/input.c:264-345,926-941,1068-1180
 *   H         hScale8To15_c / hScale16To15_c                             libswscale/swscale.c:99-142
 *   range     lum/chrRangeTo/FromJpeg_c                                  libswscale/swscale.c:163-216
 *   V + pack  yuv2planeX_8_c / yuv2plane1_8_c / yuv2nv12cX_c             libswscale/output.c:468-528
 *             yuv2planeX_10_c_template                                   libswscale/output.c:340-357
 *             yuv2rgb_X/_1/_2_c_template + yuv2rgb_write (closed form)   libswscale/output.c:1662-1939
 *
 * Shape: one CTA = 128 x TH output pixels, three CTAs per SM.  Source rows are staged in shared memory as samples
 * (RGB is converted to its 14-bit Y/U/V samples while staging, so the matrix is applied once per
 * source pixel); in the H stage a thread owns one output column and keeps that column's taps in
 * registers across all rows; the 15-bit lines stay in shared memory exactly as the reference stores
 * them (int16, clipped); in the V stage a warp owns an output row, vertical taps are warp-uniform.
 * H and V are never merged.
 */
#pragma once

#define T15_TW 128
#define T15_OUT_BYTES 4096   /* 8 warps x 512 B of packed-RGB row staging (reuses the sample staging area) */

enum { T15_SRC_U8 = 0, T15_SRC_U16 = 1, T15_SRC_RGB = 2 };
enum { T15_OUT_PLANAR8 = 0, T15_OUT_PLANARN = 1, T15_OUT_RGB8 = 2 };

struct Tile15Args {
    const uint8_t *src[3];
    uint8_t *dst[3];
    long long src_fstride[3], dst_fstride[3];
    int src_stride[3], dst_stride[3];
    int y0, y1, tile_h;
    int nl_cap, nc_cap;          /* rows of h-scaled luma / chroma lines kept per tile */
    int seg_l, seg_c;            /* staged source samples per row */
    int srl, src_rows;           /* source rows staged per pass (luma, chroma) */
};

__device__ __forceinline__ int t15_range(int val, int mode, int coeff, int offset)
{
    /* (x * coeff + offset) >> 14, clipped only towards full range (swscale.c:166-216) */
    val = (val * coeff + offset) >> 14;
    if (mode == 1)
        val = min(val, (1 << 15) - 1);
    return (int16_t)val;
}

template <int SRCK, int HT, int OUTK>
__global__ void __launch_bounds__(256, 3)
sws_tile15_kernel(const __grid_constant__ SwsCudaPlan P, const __grid_constant__ Tile15Args A)
{
    typedef typename std::conditional<SRCK == T15_SRC_U8, uint8_t, uint16_t>::type samp_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int f = blockIdx.z;
    const uint8_t *src0 = A.src[0] + f * A.src_fstride[0];
    const uint8_t *src1 = A.src[1] ? A.src[1] + f * A.src_fstride[1] : nullptr;
    const uint8_t *src2 = A.src[2] ? A.src[2] + f * A.src_fstride[2] : nullptr;
    uint8_t *dst0 = A.dst[0] + f * A.dst_fstride[0];
    uint8_t *dst1 = A.dst[1] ? A.dst[1] + f * A.dst_fstride[1] : nullptr;
    uint8_t *dst2 = A.dst[2] ? A.dst[2] + f * A.dst_fstride[2] : nullptr;

    const int TH = A.tile_h;
    const int x0 = blockIdx.x * T15_TW;
    const int ry0 = A.y0 + blockIdx.y * TH;
    const int ry1 = min(ry0 + TH, A.y1);
    const int tw = min(T15_TW, P.dst_w - x0), th = ry1 - ry0;
    const int hs = P.chr_dst_hsub, vs = P.chr_dst_vsub;
    const int CW = T15_TW >> hs;
    const int cx0 = x0 >> hs;
    const int cw = min(CW, P.chr_dst_w - cx0);
    const int cy0 = ry0 >> vs;
    const int cy1 = (ry1 == P.dst_h) ? P.chr_dst_h : (ry1 >> vs);
    const int ch = cy1 - cy0;

    int16_t *hb_l = reinterpret_cast<int16_t *>(smem_raw);
    int16_t *hb_u = hb_l + (size_t)A.nl_cap * T15_TW;
    int16_t *hb_v = hb_u + (size_t)A.nc_cap * CW;
    samp_t *stage = reinterpret_cast<samp_t *>(hb_v + (size_t)A.nc_cap * CW);

    /* source row windows of the tile; the host checked that positions are monotonic and inside the image */
    const int lo_l = P.vl_pos[ry0];
    const int nl = P.vl_pos[ry1 - 1] + P.vl_size - lo_l;
    const int lo_c = ch > 0 ? P.vc_pos[cy0] : 0;
    const int nc = ch > 0 ? P.vc_pos[cy1 - 1] + P.vc_size - lo_c : 0;
    const int sh = P.h_shift;
    const int rmode = P.range_mode;

    /* Source rows are staged as samples, srl luma rows and src chroma rows (x2 planes) per pass; the
     * first pass of both is requested before anything is computed so that a tile exposes one global
     * load latency, hidden by the other CTAs resident on the SM.  warp = row, lane = samples. */
    const int srl = A.srl, src_rows = A.src_rows;
    samp_t *stage_l = stage;
    samp_t *stage_c = stage + (size_t)srl * A.seg_l;
    const int a0 = P.hl_pos[x0];
    const int ca0 = ch > 0 ? P.hc_pos[cx0] : 0;
    const int layout = P.src_layout;
    const int uo = layout == SWSC_SRC_NV21 ? 1 : 0;

    auto load_luma = [&](int r) {
        const int nr = min(srl, nl - r), seg = A.seg_l;
        for (int row = warp; row < nr; row += 8) {
            const uint8_t *srow = src0 + (size_t)(lo_l + r + row) * A.src_stride[0];
            samp_t *d = stage_l + row * seg;
#pragma unroll 4
            for (int i = lane; i < seg; i += 32) {
                const int sx = min(a0 + i, P.src_w - 1);
                if (SRCK == T15_SRC_U8)
                    d[i] = srow[sx];
                else if (SRCK == T15_SRC_U16)
                    d[i] = reinterpret_cast<const uint16_t *>(srow)[sx];
                else
                    d[i] = (samp_t)rgb_luma14(P, srow, sx);
            }
        }
    };
    auto load_chroma = [&](int r) {
        const int nr = min(src_rows, nc - r), seg = A.seg_c;
        for (int row = warp; row < nr; row += 8) {
            const int sy = lo_c + r + row;
            samp_t *du = stage_c + row * seg, *dv = stage_c + (src_rows + row) * seg;
#pragma unroll 4
            for (int i = lane; i < seg; i += 32) {
                const int sx = min(ca0 + i, P.chr_src_w - 1);
                if (SRCK == T15_SRC_RGB) {
                    int u, v;
                    rgb_chroma14(P, src0 + (size_t)sy * A.src_stride[0], sx, u, v);
                    du[i] = (samp_t)u; dv[i] = (samp_t)v;
                } else if (layout == SWSC_SRC_PLANAR) {
                    const uint8_t *ru = src1 + (size_t)sy * A.src_stride[1];
                    const uint8_t *rv = src2 + (size_t)sy * A.src_stride[2];
                    if (SRCK == T15_SRC_U8) {
                        du[i] = ru[sx]; dv[i] = rv[sx];
                    } else {
                        du[i] = reinterpret_cast<const uint16_t *>(ru)[sx];
                        dv[i] = reinterpret_cast<const uint16_t *>(rv)[sx];
                    }
                } else {        /* nv12 / nv21 (8-bit): nv12ToUV_c */
                    const uint8_t *ruv = src1 + (size_t)sy * A.src_stride[1];
                    du[i] = ruv[2 * sx + uo]; dv[i] = ruv[2 * sx + 1 - uo];
                }
            }
        }
    };
    load_luma(0);
    if (ch > 0)
        load_chroma(0);
    __syncthreads();

    /* ================= stage H, luma: thread = (column, row parity) ================= */
    {
        const int x = tid & (T15_TW - 1), g = tid >> 7;
        const int gx = min(x0 + x, P.dst_w - 1);
        const int off = P.hl_pos[gx] - a0;
        const int fs = P.hl_size;
        int c[HT];
#pragma unroll
        for (int j = 0; j < HT; j++)
            c[j] = j < fs ? (int)P.hl_coef[(size_t)gx * fs + j] : 0;
        const int seg = A.seg_l;
        for (int r = 0; r < nl; r += srl) {
            const int nr = min(srl, nl - r);
            if (r) {
                __syncthreads();
                load_luma(r);
                __syncthreads();
            }
            for (int rr = g; rr < nr; rr += 2) {
                const samp_t *s = stage_l + rr * seg + off;
                int acc = 0;
#pragma unroll
                for (int j = 0; j < HT; j++)
                    if (j < fs)
                        acc += (int)s[j] * c[j];
                int val = min(acc >> sh, (1 << 15) - 1);
                if (rmode)
                    val = t15_range(val, rmode, (int)P.lum_rc_coeff, (int)P.lum_rc_offset);
                if (x < tw)
                    hb_l[(size_t)(r + rr) * T15_TW + x] = (int16_t)val;
            }
        }
    }
    /* ================= stage H, chroma: thread = (column, plane, row group) ================= */
    if (ch > 0) {
        const int x = tid & (CW - 1), pl = (tid / CW) & 1, g = tid / (2 * CW);
        const int ngroups = 256 / (2 * CW);
        const int gx = min(cx0 + x, P.chr_dst_w - 1);
        const int off = P.hc_pos[gx] - ca0;
        const int fs = P.hc_size;
        int c[HT];
#pragma unroll
        for (int j = 0; j < HT; j++)
            c[j] = j < fs ? (int)P.hc_coef[(size_t)gx * fs + j] : 0;
        const int seg = A.seg_c;
        int16_t *hb = pl ? hb_v : hb_u;
        for (int r = 0; r < nc; r += src_rows) {
            const int nr = min(src_rows, nc - r);
            if (r) {
                __syncthreads();
                load_chroma(r);
                __syncthreads();
            }
            for (int rr = g; rr < nr; rr += ngroups) {
                const samp_t *s = stage_c + (pl * src_rows + rr) * seg + off;
                int acc = 0;
#pragma unroll
                for (int j = 0; j < HT; j++)
                    if (j < fs)
                        acc += (int)s[j] * c[j];
                int val = min(acc >> sh, (1 << 15) - 1);
                if (rmode)
                    val = t15_range(val, rmode, (int)P.chr_rc_coeff, (int)P.chr_rc_offset);
                if (x < cw)
                    hb[(size_t)(r + rr) * CW + x] = (int16_t)val;
            }
        }
    }
    __syncthreads();

    const int kind = P.dst_kind;
    const int lfs = P.vl_size, cfs = P.vc_size;

    if (OUTK == T15_OUT_RGB8) {
        /* ============ packed RGB, one chroma pair per two pixels: warp = row, lane = pairs lane, lane+32 ============ */
        const int bpp = kind >= SWSC_DST_RGBA ? 4 : 3;
        unsigned char *orow = reinterpret_cast<unsigned char *>(stage) + warp * 512;
        const int cy = P.rgb.cy, yb = P.rgb.yb;
        for (int ty = warp; ty < th; ty += 8) {
            const int y = ry0 + ty;
            const int16_t *lf = P.vl_coef + (size_t)y * lfs;
            const int16_t *cf = P.vc_coef + (size_t)y * cfs;
            const int16_t *pl = hb_l + (size_t)(P.vl_pos[y] - lo_l) * T15_TW + 2 * lane;
            const int16_t *pu = hb_u + (size_t)(P.vc_pos[y] - lo_c) * CW + lane;
            const int16_t *pv = hb_v + (size_t)(P.vc_pos[y] - lo_c) * CW + lane;
            unsigned Y1[2] = { 0, 0 }, Y2[2] = { 0, 0 }, U[2] = { 0, 0 }, V[2] = { 0, 0 };
            for (int j = 0; j < lfs; j++) {
                const unsigned c = (unsigned)(int)__ldg(lf + j);
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(pl + (size_t)j * T15_TW + 64 * k);
                    Y1[k] += (unsigned)(int)(int16_t)(w & 0xFFFF) * c;
                    Y2[k] += (unsigned)((int)w >> 16) * c;
                }
            }
            for (int j = 0; j < cfs; j++) {
                const unsigned c = (unsigned)(int)__ldg(cf + j);
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    U[k] += (unsigned)(int)pu[(size_t)j * CW + 32 * k] * c;
                    V[k] += (unsigned)(int)pv[(size_t)j * CW + 32 * k] * c;
                }
            }
            /* yuv2packed2 (both filters bilinear 2-tap) has no rounding bias (vscale.c:148-163, output.c:1861-1864) */
            unsigned bias = 1u << 18;
            if (lfs == 2 && cfs == 2) {
                const int l0 = lf[0], l1 = lf[1], c0 = cf[0], c1 = cf[1];
                if (l0 + l1 == 4096 && (unsigned)l1 <= 4096u && c0 + c1 == 4096 && (unsigned)c1 <= 4096u)
                    bias = 0;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int y1v = (int)(Y1[k] + bias) >> 19, y2v = (int)(Y2[k] + bias) >> 19;
                const int u8 = clamp_u8((int)(U[k] + bias) >> 19), v8 = clamp_u8((int)(V[k] + bias) >> 19);
                const int pR = (P.rgb.base_r + ((v8 * P.rgb.crv) >> 16)) * cy + yb;
                const int pG = (P.rgb.base_g + ((u8 * P.rgb.cgu) >> 16) + ((v8 * P.rgb.cgv) >> 16)) * cy + yb;
                const int pB = (P.rgb.base_b + ((u8 * P.rgb.cbu) >> 16)) * cy + yb;
                const uint32_t tRa = y1v * cy + pR, tGa = y1v * cy + pG, tBa = y1v * cy + pB;
                const uint32_t tRb = y2v * cy + pR, tGb = y2v * cy + pG, tBb = y2v * cy + pB;
                constexpr uint32_t FF = 0x00FF0000u;
                const int p = lane + 32 * k;
                if (bpp == 3) {
                    uint32_t h0, h1, h2;
                    if (kind == SWSC_DST_RGB24) {
                        h0 = clamp_u8x2(prmt(tRa, tGa, 0x7632)); h1 = clamp_u8x2(prmt(tBa, tRb, 0x7632));
                        h2 = clamp_u8x2(prmt(tGb, tBb, 0x7632));
                    } else {
                        h0 = clamp_u8x2(prmt(tBa, tGa, 0x7632)); h1 = clamp_u8x2(prmt(tRa, tBb, 0x7632));
                        h2 = clamp_u8x2(prmt(tGb, tRb, 0x7632));
                    }
                    uint16_t *o = reinterpret_cast<uint16_t *>(orow + 6 * p);
                    o[0] = (uint16_t)prmt(h0, 0u, 0x4420); o[1] = (uint16_t)prmt(h1, 0u, 0x4420);
                    o[2] = (uint16_t)prmt(h2, 0u, 0x4420);
                } else {
                    uint32_t a0, a1, b0, b1;
                    if (kind == SWSC_DST_RGBA) {
                        a0 = prmt(tRa, tGa, 0x7632); a1 = prmt(tBa, FF, 0x7632); b0 = prmt(tRb, tGb, 0x7632); b1 = prmt(tBb, FF, 0x7632);
                    } else if (kind == SWSC_DST_BGRA) {
                        a0 = prmt(tBa, tGa, 0x7632); a1 = prmt(tRa, FF, 0x7632); b0 = prmt(tBb, tGb, 0x7632); b1 = prmt(tRb, FF, 0x7632);
                    } else if (kind == SWSC_DST_ARGB) {
                        a0 = prmt(FF, tRa, 0x7632); a1 = prmt(tGa, tBa, 0x7632); b0 = prmt(FF, tRb, 0x7632); b1 = prmt(tGb, tBb, 0x7632);
                    } else {
                        a0 = prmt(FF, tBa, 0x7632); a1 = prmt(tGa, tRa, 0x7632); b0 = prmt(FF, tBb, 0x7632); b1 = prmt(tGb, tRb, 0x7632);
                    }
                    a0 = clamp_u8x2(a0); a1 = clamp_u8x2(a1); b0 = clamp_u8x2(b0); b1 = clamp_u8x2(b1);
                    *reinterpret_cast<uint2 *>(orow + 8 * p) = make_uint2(prmt(a0, a1, 0x6420), prmt(b0, b1, 0x6420));
                }
            }
            __syncwarp();
            /* copy the finished row out: 16-byte stores when the destination row allows it */
            uint8_t *d = dst0 + (size_t)y * A.dst_stride[0] + (size_t)x0 * bpp;
            const int nbytes = tw * bpp;
            int done = 0;
            if (((uintptr_t)d & 15) == 0) {
                for (int i = lane; i < (nbytes >> 4); i += 32)
                    reinterpret_cast<uint4 *>(d)[i] = reinterpret_cast<const uint4 *>(orow)[i];
                done = nbytes & ~15;
            }
            for (int i = done + lane; i < nbytes; i += 32)
                d[i] = orow[i];
        }
        return;
    }

    /* ============ planar / semi-planar YUV: warp = row, lane = columns lane + 32k ============ */
    const int bits = P.dst_bits;
    for (int ty = warp; ty < th; ty += 8) {
        const int y = ry0 + ty;
        const int16_t *lf = P.vl_coef + (size_t)y * lfs;
        const int16_t *pl = hb_l + (size_t)(P.vl_pos[y] - lo_l) * T15_TW + lane;
        unsigned acc[4] = { 0, 0, 0, 0 };
        for (int j = 0; j < lfs; j++) {
            const unsigned c = (unsigned)(int)__ldg(lf + j);
#pragma unroll
            for (int k = 0; k < 4; k++)
                acc[k] += (unsigned)(int)pl[(size_t)j * T15_TW + 32 * k] * c;
        }
        uint8_t *d = dst0 + (size_t)y * A.dst_stride[0];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int col = lane + 32 * k, gx = x0 + col;
            if (col >= tw)
                continue;
            if (OUTK == T15_OUT_PLANAR8) {
                const int dz = P.dither_bayer ? c_dither_8x8_128[y & 7][gx & 7] : 64;
                d[gx] = (uint8_t)clamp_u8((int)(acc[k] + ((unsigned)dz << 12)) >> 19);
            } else {
                const int shift = 27 - bits;
                reinterpret_cast<uint16_t *>(d)[gx] =
                    (uint16_t)clip_uintp2((int)(acc[k] + (1u << (shift - 1))) >> shift, bits);
            }
        }
    }
    for (int ty = warp; ty < ch; ty += 8) {
        const int y = cy0 + ty;
        const int16_t *cf = P.vc_coef + (size_t)y * cfs;
        const int16_t *pu = hb_u + (size_t)(P.vc_pos[y] - lo_c) * CW + lane;
        const int16_t *pv = hb_v + (size_t)(P.vc_pos[y] - lo_c) * CW + lane;
        unsigned au[4] = { 0, 0, 0, 0 }, av[4] = { 0, 0, 0, 0 };
        for (int j = 0; j < cfs; j++) {
            const unsigned c = (unsigned)(int)__ldg(cf + j);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (32 * k < CW) {
                    au[k] += (unsigned)(int)pu[(size_t)j * CW + 32 * k] * c;
                    av[k] += (unsigned)(int)pv[(size_t)j * CW + 32 * k] * c;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int col = lane + 32 * k, gx = cx0 + col;
            if (col >= cw)
                continue;
            if (OUTK == T15_OUT_PLANAR8) {
                const int du = P.dither_bayer ? c_dither_8x8_128[y & 7][gx & 7] : 64;
                const int dv = P.dither_bayer ? c_dither_8x8_128[y & 7][(gx + 3) & 7] : 64;
                const int u8 = clamp_u8((int)(au[k] + ((unsigned)du << 12)) >> 19);
                const int v8 = clamp_u8((int)(av[k] + ((unsigned)dv << 12)) >> 19);
                if (kind == SWSC_DST_PLANAR8) {
                    dst1[(size_t)y * A.dst_stride[1] + gx] = (uint8_t)u8;
                    dst2[(size_t)y * A.dst_stride[2] + gx] = (uint8_t)v8;
                } else {
                    uint8_t *d = dst1 + (size_t)y * A.dst_stride[1] + 2 * gx;
                    d[0] = (uint8_t)(kind == SWSC_DST_NV12 ? u8 : v8);
                    d[1] = (uint8_t)(kind == SWSC_DST_NV12 ? v8 : u8);
                }
            } else {
                const int shift = 27 - bits;
                reinterpret_cast<uint16_t *>(dst1 + (size_t)y * A.dst_stride[1])[gx] =
                    (uint16_t)clip_uintp2((int)(au[k] + (1u << (shift - 1))) >> shift, bits);
                reinterpret_cast<uint16_t *>(dst2 + (size_t)y * A.dst_stride[2])[gx] =
                    (uint16_t)clip_uintp2((int)(av[k] + (1u << (shift - 1))) >> shift, bits);
            }
        }
    }
}
