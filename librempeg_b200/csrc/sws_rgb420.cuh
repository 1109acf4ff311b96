/*
 * sws_rgb420.cuh -- packed 8-bit RGB -> 8-bit 4:2:0 YUV of the same size through the scaler path
 * (the encode-side mirror of the headline conversion): rgb24 / bgr24 / rgba /
 * bgra / argb / abgr -> yuv420p / nv12 / nv21 with SWS_ACCURATE_RND or any scaler flag that does not
 * select a special converter.  Fuses
 *   rgb24ToY_c / rgb24ToUV_half_c, rgb16_32ToY/UV_half_c_template     libswscale/input.c:264-345,1068-1180
 *   hScale16To15_c with the identity filter (x << 1, clipped)         libswscale/swscale.c:99-125
 *   yuv2plane1_8_c (luma) / yuv2planeX_8_c (chroma, 2:1 vertical FIR)  libswscale/output.c:468-493
 *   with the constant dither 64 of 8-bit source formats               libswscale/swscale.c:54-56,385-387
 *
 * Arithmetic, exactly the reference's:
 *   Y14 = (ry r + gy g + by b + (32 << 14) + (1 << 8)) >> 9               (the 32-bit readers' unsigned
 *   U14 = (ru (r0+r1) + gu (g0+g1) + bu (b0+b1) + (256 << 15) + (1 << 9)) >> 10   form is the same number)
 *   line15 = min(2 * x14, 32767);  Y = clip_u8((line15 + 64) >> 7);  U = clip_u8((sum_j c_j line15_j + (64 << 12)) >> 19)
 * The host admits only matrices for which 0 <= x14 < 16384 for every input (true for every colourspace
 * libswscale can produce; checked, not assumed), so the uint16 wrap and the clip at 32767 cannot trigger and
 *   Y = min(255, (sY + 16384) >> 15)          with sY the biased luma dot product (floor of floor = floor),
 *   U = clip_u8((sum_j c_j U14_j + (64 << 11)) >> 18)     with the 14-bit samples kept in shared memory.
 * A pixel is one 32-bit word (3-byte pixels are cut out of the row with funnel shifts), so the matrix
 * row is two IDP.2A with the 16-bit coefficients arranged per byte order on the host.
 *
 * Shape: one CTA = 128 pixels x CR chroma rows; phase 1: a thread converts 16 pixels of one source row
 * (16-byte loads, one 16-byte luma store, 8 U + 8 V 15-bit samples into shared memory); phase 2: a
 * thread filters 8 chroma columns of one output row vertically.
 */
#pragma once

#define R420_TW 128
#define R420_MAXROWS 64          /* source rows of 15-bit chroma lines kept per tile */

struct Rgb420Args {
    const uint8_t *src;
    uint8_t *dst[3];
    long long src_fstride, dst_fstride[3];
    int src_stride, dst_stride[3];
    int w;                       /* luma width (a multiple of 16) */
    int cy_begin, cy_end;        /* chroma rows of this launch */
    int y_end;                   /* luma rows end (exclusive) */
    int cr;                      /* chroma rows per tile */
    int vc_size;
    const int16_t *vc_coef;
    const int32_t *vc_pos;
    uint32_t ylo, yhi, ulo, uhi, vlo, vhi;   /* matrix rows as 16-bit pairs in pixel byte order */
    int dst_kind;
};

template <int BPP>
__global__ void __launch_bounds__(256)
sws_rgb420_kernel(const __grid_constant__ Rgb420Args A)
{
    __shared__ __align__(16) uint16_t s_u[R420_MAXROWS][R420_TW / 2];
    __shared__ __align__(16) uint16_t s_v[R420_MAXROWS][R420_TW / 2];
    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * R420_TW;
    const int cy0 = A.cy_begin + blockIdx.y * A.cr;
    const int cy1 = min(cy0 + A.cr, A.cy_end);
    const int nchunk = min(R420_TW, A.w - x0) >> 4;            /* 16-pixel chunks in this tile */
    /* source rows: the chroma window of the tile, widened to its luma rows if needed */
    const int ly0 = 2 * cy0, ly1 = min(2 * cy1, A.y_end);
    const int lo_c = __ldg(A.vc_pos + cy0);
    const int hi_c = __ldg(A.vc_pos + cy1 - 1) + A.vc_size;
    const int r0 = min(lo_c, ly0), r1 = max(hi_c, ly1);
    const uint8_t *src = A.src + f * A.src_fstride + (size_t)x0 * BPP;
    uint8_t *dy = A.dst[0] + f * A.dst_fstride[0] + x0;

    /* ---------------- phase 1: item = (source row, 16-pixel chunk) ----------------
     * the loads of both rounds of a full tile are issued before anything is computed */
    const int items = (r1 - r0) * 8;
    constexpr int NW = BPP == 4 ? 16 : 12;
    uint32_t w[2][NW];
    for (int base = 0; base < items; base += 512) {
#pragma unroll
        for (int rd = 0; rd < 2; rd++) {
            const int it = base + rd * 256 + tid;
            const int row = r0 + (it >> 3), k = it & 7;
            if (it < items && k < nchunk) {
                const uint4 *q = reinterpret_cast<const uint4 *>(src + (size_t)row * A.src_stride + (size_t)k * 16 * BPP);
#pragma unroll
                for (int i = 0; i < NW / 4; i++) {
                    const uint4 v = __ldg(q + i);
                    w[rd][4 * i] = v.x; w[rd][4 * i + 1] = v.y; w[rd][4 * i + 2] = v.z; w[rd][4 * i + 3] = v.w;
                }
            }
        }
#pragma unroll
        for (int rd = 0; rd < 2; rd++) {
            const int it = base + rd * 256 + tid;
            const int row = r0 + (it >> 3), k = it & 7;
            if (it >= items || k >= nchunk)
                continue;
            uint32_t px[16];
            if (BPP == 4) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    px[i] = w[rd][i];
            } else {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    px[4 * g] = w[rd][3 * g];
                    px[4 * g + 1] = __funnelshift_r(w[rd][3 * g], w[rd][3 * g + 1], 24);
                    px[4 * g + 2] = __funnelshift_r(w[rd][3 * g + 1], w[rd][3 * g + 2], 16);
                    px[4 * g + 3] = w[rd][3 * g + 2] >> 8;
                }
            }
            if (row >= ly0 && row < ly1) {
                uint32_t yo[4];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    uint32_t y4[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int sy = dp2a_hi_su(A.yhi, px[i + j], dp2a_lo_su(A.ylo, px[i + j], (32 << 14) + (1 << 8) + 16384));
                        y4[j] = (uint32_t)min(sy >> 15, 255);
                    }
                    yo[i >> 2] = prmt(prmt(y4[0], y4[1], 0x0040), prmt(y4[2], y4[3], 0x0040), 0x5410);
                }
                __stcs(reinterpret_cast<uint4 *>(dy + (size_t)row * A.dst_stride[0] + 16 * k), make_uint4(yo[0], yo[1], yo[2], yo[3]));
            }
            if (row >= lo_c && row < hi_c) {
                uint32_t uo[4], vo[4];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    int su[2], sv[2];
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const uint32_t a = px[2 * (i + j)], b = px[2 * (i + j) + 1];
                        su[j] = dp2a_hi_su(A.uhi, a, dp2a_lo_su(A.ulo, a, (256 << 15) + (1 << 9)));
                        su[j] = dp2a_hi_su(A.uhi, b, dp2a_lo_su(A.ulo, b, su[j]));
                        sv[j] = dp2a_hi_su(A.vhi, a, dp2a_lo_su(A.vlo, a, (256 << 15) + (1 << 9)));
                        sv[j] = dp2a_hi_su(A.vhi, b, dp2a_lo_su(A.vlo, b, sv[j]));
                    }
                    uo[i >> 1] = prmt((uint32_t)(su[0] >> 10), (uint32_t)(su[1] >> 10), 0x5410);
                    vo[i >> 1] = prmt((uint32_t)(sv[0] >> 10), (uint32_t)(sv[1] >> 10), 0x5410);
                }
                *reinterpret_cast<uint4 *>(&s_u[row - lo_c][8 * k]) = make_uint4(uo[0], uo[1], uo[2], uo[3]);
                *reinterpret_cast<uint4 *>(&s_v[row - lo_c][8 * k]) = make_uint4(vo[0], vo[1], vo[2], vo[3]);
            }
        }
    }
    __syncthreads();

    /* ---------------- phase 2: item = (chroma output row, 8 chroma columns) ---------------- */
    const int fs = A.vc_size;
    const int citems = (cy1 - cy0) * 8;
    for (int it = tid; it < citems; it += 256) {
        const int cy = cy0 + (it >> 3), g = it & 7;
        if (g >= nchunk)
            continue;
        const int pos = __ldg(A.vc_pos + cy) - lo_c;
        const int16_t *cf = A.vc_coef + (size_t)cy * fs;
        unsigned au[8], av[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            au[i] = av[i] = 64u << 11;
        for (int j = 0; j < fs; j++) {
            const unsigned c = (unsigned)(int)__ldg(cf + j);
            const uint4 u = *reinterpret_cast<const uint4 *>(&s_u[pos + j][8 * g]);
            const uint4 v = *reinterpret_cast<const uint4 *>(&s_v[pos + j][8 * g]);
            au[0] += (u.x & 0xFFFFu) * c; au[1] += (u.x >> 16) * c; au[2] += (u.y & 0xFFFFu) * c; au[3] += (u.y >> 16) * c;
            au[4] += (u.z & 0xFFFFu) * c; au[5] += (u.z >> 16) * c; au[6] += (u.w & 0xFFFFu) * c; au[7] += (u.w >> 16) * c;
            av[0] += (v.x & 0xFFFFu) * c; av[1] += (v.x >> 16) * c; av[2] += (v.y & 0xFFFFu) * c; av[3] += (v.y >> 16) * c;
            av[4] += (v.z & 0xFFFFu) * c; av[5] += (v.z >> 16) * c; av[6] += (v.w & 0xFFFFu) * c; av[7] += (v.w >> 16) * c;
        }
        uint32_t ub[8], vb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ub[i] = (uint32_t)clip_u8((int)au[i] >> 18);
            vb[i] = (uint32_t)clip_u8((int)av[i] >> 18);
        }
        const int cx = (x0 >> 1) + 8 * g;
        if (A.dst_kind == SWSC_DST_PLANAR8) {
            const uint2 uw = make_uint2(ub[0] | ub[1] << 8 | ub[2] << 16 | ub[3] << 24, ub[4] | ub[5] << 8 | ub[6] << 16 | ub[7] << 24);
            const uint2 vw = make_uint2(vb[0] | vb[1] << 8 | vb[2] << 16 | vb[3] << 24, vb[4] | vb[5] << 8 | vb[6] << 16 | vb[7] << 24);
            __stcs(reinterpret_cast<uint2 *>(A.dst[1] + f * A.dst_fstride[1] + (size_t)cy * A.dst_stride[1] + cx), uw);
            __stcs(reinterpret_cast<uint2 *>(A.dst[2] + f * A.dst_fstride[2] + (size_t)cy * A.dst_stride[2] + cx), vw);
        } else {
            const uint32_t *e = A.dst_kind == SWSC_DST_NV12 ? ub : vb, *o = A.dst_kind == SWSC_DST_NV12 ? vb : ub;
            const uint4 w = make_uint4(e[0] | o[0] << 8 | e[1] << 16 | o[1] << 24, e[2] | o[2] << 8 | e[3] << 16 | o[3] << 24,
                                       e[4] | o[4] << 8 | e[5] << 16 | o[5] << 24, e[6] | o[6] << 8 | e[7] << 16 | o[7] << 24);
            __stcs(reinterpret_cast<uint4 *>(A.dst[1] + f * A.dst_fstride[1] + (size_t)cy * A.dst_stride[1] + 2 * cx), w);
        }
    }
}
