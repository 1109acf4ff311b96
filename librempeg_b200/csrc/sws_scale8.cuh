/*
 * sws_scale8.cuh -- the scaling kernel for 8-bit YUV -> 8-bit planar / semi-planar YUV
 * (BASELINE config C4: 7680x4320 nv12 -> 1920x1080 yuv420p bicubic, 16x16 taps; every
 * yuv -> yuv resize of 8-bit material).  Fuses per 128 x TH output tile
 *   nv12ToUV_c (de-interleave, done while staging rows)        libswscale/input.c:926-941
 *   hScale8To15_c                                               libswscale/swscale.c:128-142
 *   yuv2planeX_8_c / yuv2plane1_8_c / yuv2nv12cX_c              libswscale/output.c:468-528
 *
 * This path is integer-MAC bound, not HBM bound (C4: ~250 M MAC per frame), so the work is shaped
 * for the dot-product units, exactly (no approximation):
 *  H: source rows are staged in shared memory with 16-byte loads; each thread owns one output column,
 *     keeps its taps as packed byte words (c = 256*ch + cl) in registers and evaluates four taps per
 *     IDP.4A pair:  sum src*c = 256*dp4a(src, ch) + dp4a(src, cl); unaligned windows come from
 *     funnel shifts of aligned 32-bit shared-memory words.  The 15-bit result is stored TRANSPOSED
 *     (column major) so that two vertically adjacent samples share a 32-bit word.
 *  V: IDP.2A consumes those pairs: two taps per instruction pair, again with split coefficients;
 *     an odd first source row is absorbed by a leading zero tap prepared on the host.
 * H and V stay separate stages with the reference's 15-bit clip in between.
 *
 * MMA = true (round 2): the horizontal stage runs on the tensor pipe instead.  For 8 neighbouring output
 * columns the FIR is a banded 8-column matrix B[k][n] = coef[n][window + k - pos[n]] over a K = 32*KS byte
 * window of the staged rows, so 16 source rows x 8 columns are KS x 2 warp-level integer MMAs
 *   mma.sync.m16n8k32  u8 (rows) x u8 (low coefficient bytes)  and  u8 x s8 (high bytes),  s32 accumulate
 * -- exact: sum src*c = 256*sum src*ch + sum src*cl, the 15-bit clip stays after it (swscale.c:128-142).
 * A fragments come straight out of the TMA ring with ldmatrix (MMA row i = source row 2i, row 8+i = source row
 * 2i+1, so a thread's accumulator pair is the vertically adjacent sample pair the transposed line buffer
 * stores as one word); B fragments are prepared on the host per column group and live in registers for the
 * whole tile.  Interleaved nv12 / nv21 chroma is de-interleaved with two PRMT per fragment (the K order of a
 * contraction is free: the host permutes B to match).  Measured: the legacy integer MMA sustains 2044 MAC /
 * clk / SM on B200, 8x the IDP.4A pipe (profiles/microbench/mma_i8_bench.cu), and the dot-product pipe is
 * left to the vertical stage.
 *
 * The same tile machinery serves a family of instantiations (template parameters of sws_scale8_kernel):
 *   SRCK  S8_SRC_U8   8-bit planar / nv12 / nv21 (hScale8To15_c; MMA or IDP.4A horizontal stage)
 *         S8_SRC_U16  9..16-bit planar (hScale16To15_c as IDP.2A over sample pairs, swscale.c:99-125)
 *         S8_SRC_P010 p010le: containers >> 6 after the funnel shift, interleaved 16-bit chroma cut apart with PRMT
 *         S8_SRC_RGB  packed 8-bit RGB: a reader stage per ring slot (input.c:264-345,1068-1180) feeding the 16-bit stage
 *   RGBK  0 planar / semi-planar YUV of 8..14 bits and p010le (output.c:340-357,468-589)
 *         1 packed 8-bit / 15-16 bpp RGB, one chroma sample per pixel pair (yuv2rgb_{X,2,1} + yuv2rgb_write)
 *         2 packed 8-bit RGB with full horizontal chroma (yuv2rgb_full_{X,2,1} + yuv2rgb_write_full)
 *         3 19-bit lines (hScale8To19_c / hScale16To19_c, one int32 sample per word) into 16-bit planar YUV
 *           (yuv2planeX_16_c / yuv2plane1_16_c), rgb48le / bgr48le (yuv2rgba64[_full]_{X,2,1}), gbrpf32le
 *           (yuv2gbrpf32_full_X_c) and grayf32le (yuv2plane1/X_float_c, luma only)
 * with the range conversion of swscale.c:163-255 between the two FIR stages.  DESIGN.md section 4.3 has the measurements.
 */
#pragma once

#define S8_TW 128
#ifndef S8_ROWS
#define S8_ROWS 16       /* source rows per ring slot (one TMA box): 8 or 16 */
#endif
#ifndef S8_STAGES
#define S8_STAGES 2      /* ring depth */
#endif
#define S8_MAX_STAGES 6  /* the ring depth is a launch parameter: as many slots as the shared-memory budget leaves */
#define S8_THREADS 288   /* 8 filtering warps + 1 TMA producer warp */
#ifndef S8_LIGHT_FS4
#define S8_LIGHT_FS4 2    /* planar variants with at most this many tap groups are also compiled for S8_RGB_CTAS CTAs per SM */
#endif
#ifndef S8_RGB_CTAS
#define S8_RGB_CTAS 4     /* CTAs per SM the packed-RGB variant is compiled for (register budget of its V + colour stage) */
#endif
#define S8_VF4 5         /* vertical taps: up to 5 groups of 4 (16 taps + parity pad) */
enum { S8_SRC_U8 = 0, S8_SRC_U16 = 1, S8_SRC_RGB = 2, S8_SRC_P010 = 3 };

struct S8VRow {          /* per destination row, 48 bytes */
    int pos_even;        /* first source row, rounded down to even; luma bank, RGB output: bit 0 = this row takes
                            the bias-free yuv2packed2 rounding (vscale.c:148-163, output.c:1861-1864) */
    int n4;              /* groups of four taps in use */
    uint32_t cl[S8_VF4]; /* low bytes of the taps, four per word */
    uint32_t ch[S8_VF4]; /* high (signed) bytes */
};

struct Scale8Args {
    uint8_t *dst[3];
    long long dst_fstride[3];
    int dst_stride[3];
    int src_w, src_h, chr_src_w, chr_src_h, dst_w, dst_h, chr_dst_w, chr_dst_h;
    int hs, vs;
    int src_layout, dst_kind;
    int y0, y1, tile_h;
    int nl_cap, nc_cap;
    int cy, yb, base_r, base_g, base_b, crv, cgu, cgv, cbu;   /* packed RGB output: closed-form LUT constants */
    /* packed RGB output with full horizontal chroma (SWS_FULL_CHR_H_INT: odd widths, 4:4:4 sources): one U, V pair per
     * pixel and the arithmetic colour step of yuv2rgb_write_full (output.c:1998-2051) */
    int full_chr, y_offset, y_coeff, v2r, v2g, u2g, u2b;
    int vl_n4, vc_n4;        /* vertical tap groups of four in use (max over rows) */
    int seg_l, seg_c;        /* staged bytes per luma row / chroma samples per chroma row */
    int slot_bytes;          /* one ring slot: max(8 luma rows, 8 rows of both chroma planes), 128-byte multiple */
    int stages;              /* ring depth, 2 .. S8_MAX_STAGES */
    int bps;                 /* log2 bytes per source sample: 0 (8-bit) or 1 (9..16-bit little-endian planar) */
    int elt_shift;           /* log2 bytes per TMA element of the source maps (2: u32, 3: u64 for rows beyond 1 KB) */
    int h_shift;             /* 16-bit sources: right shift after the horizontal FIR (depth - 1, swscale.c:99-125) */
    int range_mode;          /* 0 none, 1 to full range, 2 to limited range (swscale.c:163-216), on the h-scaled lines */
    int lum_rc_coeff, lum_rc_offset, chr_rc_coeff, chr_rc_offset;
    /* packed 8-bit RGB sources: bytes per pixel, chroma from summed pixel pairs (the *_half readers), the matrix rows
     * as 16-bit pairs in the byte order of a pixel word, bytes per row of the luma / chroma sample buffers */
    int src_bpp, rgb_half, seg_sy, seg_sc;
    uint32_t ylo, yhi, ulo, uhi, vlo, vhi;
    /* 19-bit lines (RGBK = 3: 16-bit planar destinations): range constants in full width, the plain vertical banks */
    long long lum_rc_offset64, chr_rc_offset64;
    const int16_t *vl_coef16, *vc_coef16;
    const int32_t *vl_pos32, *vc_pos32;
    int vl_size, vc_size;
    int out_bits;            /* planar destinations: 8, or 9..14 (16-bit little-endian samples) */
    int no_chroma;           /* the destination has no chroma planes (grayf32le) */
    int out_lshift;          /* p010le destination: 10-bit samples shifted up by 6, chroma interleaved U first */
    int dither_bayer;        /* 8-bit planar output of > 8-bit sources: ff_dither_8x8_128 instead of the constant 64 */
    const int *hl_pos, *hc_pos;
    const uint32_t *hl_cl, *hl_ch, *hc_cl, *hc_ch;
    const S8VRow *vl, *vc;
    /* vertical banks of 21..40 taps (incl. the parity pad): the taps beyond the first 20 of every row, as a second
     * record per row (first row = the first record's + 20; n4 = 0 and the same first row where a row has none);
     * nullptr for banks that fit one record */
    const S8VRow *vl2, *vc2;
    /* MMA horizontal stage: per group of 8 output columns the byte offset of its K window inside the staged
     * row, and the B fragments [group][KS][lo0 lo1 hi0 hi1][lane] */
    const int *hl_goff, *hc_goff;
    const uint32_t *hl_B, *hc_B;
};

/* ---- warp-level integer MMA helpers ---- */
__device__ __forceinline__ void s8_ldmatrix4(uint32_t addr, uint32_t (&a)[4])
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
__device__ __forceinline__ void s8_mma_uu(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void s8_mma_us(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS>
struct S8Bfrag {
    uint32_t r[KS][4];       /* lo0 lo1 hi0 hi1 per K step */
};

template <int KS>
__device__ __forceinline__ void s8_load_bfrag(const uint32_t *tab, int group, int lane, S8Bfrag<KS> &b)
{
    const uint32_t *p = tab + (size_t)group * (KS * 4 * 32) + lane;
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            b.r[ks][j] = __ldg(p + (ks * 4 + j) * 32);
}

/* Which staged row feeds which MMA row.  A thread's accumulators c0/c1 (MMA row g) and c2/c3 (MMA row g + 8) must
 * be the vertically adjacent pair (2g, 2g+1) that one word of the transposed line buffer holds, and the eight
 * rows of each ldmatrix matrix must fall into eight different 16-byte bank groups (row pitch = odd multiple of
 * 16 bytes): MMA rows 0..3 take the even row of their pair, rows 4..7 the odd one (rows 0 2 4 6 9 11 13 15 are
 * all different mod 8), MMA rows 8..15 the other one.  S8_PAIR_SEL(lane) orders the two halves of the word. */
__device__ __forceinline__ uint32_t s8_ldm_row(int lane)
{
    const int i = lane & 7;
    return 2 * i + (((lane >> 3) & 1) ^ (i >> 2));
}
#define S8_PAIR_SEL(lane) (((lane) & 16) ? 0x1054u : 0x5410u)

/* 16 staged rows x 8 output columns: accumulators to two words of vertically adjacent 15-bit samples
 * (columns 2t and 2t+1 of the group, source rows 2g and 2g+1 of the slot; t = lane & 3, g = lane >> 2) */
struct S8Range {         /* range conversion of the h-scaled lines; mode 0 = none */
    int mode, coeff, offset;
};

__device__ __forceinline__ int s8_range(int val, int mode, int coeff, int offset);

__device__ __forceinline__ void s8_mma_pack(const int (&lo)[4], const int (&hi)[4], uint32_t sel, const S8Range &rc,
                                            uint32_t &wa, uint32_t &wb)
{
    int v0 = min(((hi[0] << 8) + lo[0]) >> 7, (1 << 15) - 1);
    int v1 = min(((hi[1] << 8) + lo[1]) >> 7, (1 << 15) - 1);
    int v2 = min(((hi[2] << 8) + lo[2]) >> 7, (1 << 15) - 1);
    int v3 = min(((hi[3] << 8) + lo[3]) >> 7, (1 << 15) - 1);
    if (rc.mode) {
        v0 = s8_range(v0, rc.mode, rc.coeff, rc.offset); v1 = s8_range(v1, rc.mode, rc.coeff, rc.offset);
        v2 = s8_range(v2, rc.mode, rc.coeff, rc.offset); v3 = s8_range(v3, rc.mode, rc.coeff, rc.offset);
    }
    wa = prmt((uint32_t)v0, (uint32_t)v2, sel);
    wb = prmt((uint32_t)v1, (uint32_t)v3, sel);
}

/* plain rows (luma, planar chroma): KS ldmatrix + 2 KS MMAs */
template <int KS>
__device__ __forceinline__ void s8_mma_rows(uint32_t addr, const S8Bfrag<KS> &b, uint32_t sel, const S8Range &rc,
                                            uint32_t &wa, uint32_t &wb)
{
    int lo[4] = { 0, 0, 0, 0 }, hi[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        uint32_t a[4];
        s8_ldmatrix4(addr + 32 * ks, a);
        s8_mma_uu(lo, a, b.r[ks][0], b.r[ks][1]);
        s8_mma_us(hi, a, b.r[ks][2], b.r[ks][3]);
    }
    s8_mma_pack(lo, hi, sel, rc, wa, wb);
}

/* interleaved chroma rows (nv12 / nv21): two ldmatrix per K step of 32 chroma samples, U and V fragments cut
 * out of them with PRMT; `even` = the plane stored first in memory */
template <int KS>
__device__ __forceinline__ void s8_mma_rows_uv(uint32_t addr, const S8Bfrag<KS> &b, uint32_t sel, const S8Range &rc,
                                               uint32_t &ea, uint32_t &eb, uint32_t &oa, uint32_t &ob)
{
    int elo[4] = { 0, 0, 0, 0 }, ehi[4] = { 0, 0, 0, 0 }, olo[4] = { 0, 0, 0, 0 }, ohi[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
        uint32_t p[4], q[4], e[4], o[4];
        s8_ldmatrix4(addr + 64 * ks, p);
        s8_ldmatrix4(addr + 64 * ks + 32, q);
        e[0] = prmt(p[0], p[2], 0x6420); e[1] = prmt(p[1], p[3], 0x6420);
        e[2] = prmt(q[0], q[2], 0x6420); e[3] = prmt(q[1], q[3], 0x6420);
        o[0] = prmt(p[0], p[2], 0x7531); o[1] = prmt(p[1], p[3], 0x7531);
        o[2] = prmt(q[0], q[2], 0x7531); o[3] = prmt(q[1], q[3], 0x7531);
        s8_mma_uu(elo, e, b.r[ks][0], b.r[ks][1]);
        s8_mma_us(ehi, e, b.r[ks][2], b.r[ks][3]);
        s8_mma_uu(olo, o, b.r[ks][0], b.r[ks][1]);
        s8_mma_us(ohi, o, b.r[ks][2], b.r[ks][3]);
    }
    s8_mma_pack(elo, ehi, sel, rc, ea, eb);
    s8_mma_pack(olo, ohi, sel, rc, oa, ob);
}

__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* horizontal FIR of one staged row for one output column: FS4 groups of four taps */
template <int FS4, bool I19 = false>
__device__ __forceinline__ int s8_hfir(const unsigned char *srow, int sh, const uint32_t (&cl)[FS4],
                                       const uint32_t (&ch)[FS4])
{
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(srow);
    uint32_t w0 = wp[0];
    int acc_l = 0, acc_h = 0;
#pragma unroll
    for (int k = 0; k < FS4; k++) {
        const uint32_t w1 = wp[k + 1];
        const uint32_t v = __funnelshift_r(w0, w1, sh);
        acc_l = dp4a_uu(v, cl[k], acc_l);
        acc_h = dp4a_us(v, ch[k], acc_h);
        w0 = w1;
    }
    /* hScale8To15_c: >> 7, 15 bits; hScale8To19_c (swscale.c:144-159): >> 3, 19 bits */
    return I19 ? min(((acc_h << 8) + acc_l) >> 3, (1 << 19) - 1) : min(((acc_h << 8) + acc_l) >> 7, (1 << 15) - 1);
}

/* lumRangeToJpeg_c / lumRangeFromJpeg_c and the chroma twins (swscale.c:163-216) on one 15-bit sample:
 * (x * coeff + offset) >> 14, clipped towards full range only, stored as int16 */
__device__ __forceinline__ int s8_range(int val, int mode, int coeff, int offset)
{
    val = (val * coeff + offset) >> 14;
    if (mode == 1)
        val = min(val, (1 << 15) - 1);
    return (int)(int16_t)val;
}

/* the 19-bit twins (lumRangeToJpeg16_c & co., swscale.c:218-255): 64-bit product, >> 18 */
__device__ __forceinline__ int s19_range(int val, int mode, uint32_t coeff, long long offset)
{
    val = (int)(((long long)val * coeff + offset) >> 18);
    if (mode == 1)
        val = min(val, (1 << 19) - 1);
    return val;
}

/* yuv2planeX_16_c / yuv2plane1_16_c (output.c:163-187) for NC columns of 19-bit lines (one sample per word, cstep
 * words apart), executed by a whole warp: the row's taps sit in two registers per lane and are broadcast per tap.
 * One tap never multiplies (a line below -2^19 must not wrap through the x 4096 of the general form). */
template <int NC>
__device__ __forceinline__ void s19_vrow(const uint32_t *col, int cstep, const int16_t *coef, int size, int lane,
                                         int (&out)[NC])
{
    if (size == 1) {
#pragma unroll
        for (int c = 0; c < NC; c++)
            out[c] = min(max(((int)col[c * cstep] + 4) >> 3, 0), 65535);
        return;
    }
    const int c0 = lane < size ? (int)__ldg(coef + lane) : 0;
    const int c1 = lane + 32 < size ? (int)__ldg(coef + lane + 32) : 0;
    unsigned acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++)
        acc[c] = (1u << 14) - 0x40000000u;
    for (int j = 0; j < size; j++) {
        const unsigned cj = (unsigned)__shfl_sync(0xffffffffu, j < 32 ? c0 : c1, j & 31);
#pragma unroll
        for (int c = 0; c < NC; c++)
            acc[c] += col[c * cstep + j] * cj;
    }
#pragma unroll
    for (int c = 0; c < NC; c++)
        out[c] = 0x8000 + min(max((int)acc[c] >> 15, -32768), 32767);
}

__device__ __forceinline__ int dp2a_lo_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* hScale16To15_c (swscale.c:99-125) of one staged row of 16-bit samples for one output column: two samples per
 * word, an odd first sample is a 16-bit funnel shift; IDP.2A against the split coefficient bytes, everything
 * modulo 2^32 like the C code's int accumulator */
template <int FS4, bool I19 = false, int SSH = 0>
__device__ __forceinline__ int s16_hfir(const unsigned char *srow, int sh, int hshift, const uint32_t (&cl)[FS4],
                                        const uint32_t (&ch)[FS4])
{
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(srow);
    uint32_t w0 = wp[0];
    int acc_l = 0, acc_h = 0;
#pragma unroll
    for (int k = 0; k < FS4; k++) {
        const uint32_t w1 = wp[2 * k + 1], w2 = wp[2 * k + 2];
        uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
        if (SSH) {               /* p010LEToY_c (input.c): the sample is the container >> 6 */
            v0 = (v0 >> SSH) & (0x00010001u * (0xFFFFu >> SSH));
            v1 = (v1 >> SSH) & (0x00010001u * (0xFFFFu >> SSH));
        }
        acc_l = dp2a_lo_uu(v0, cl[k], acc_l);
        acc_h = dp2a_lo_us(v0, ch[k], acc_h);
        acc_l = dp2a_hi_uu(v1, cl[k], acc_l);
        acc_h = dp2a_hi_us(v1, ch[k], acc_h);
        w0 = w2;
    }
    return min(((acc_h << 8) + acc_l) >> hshift, I19 ? (1 << 19) - 1 : (1 << 15) - 1);     /* hScale16To15_c / To19_c */
}

/* one staged row of interleaved 16-bit chroma (p010le: U in the low, V in the high half of every word; p010LEToUV_c,
 * input.c) for one output column: sample pairs of each plane are cut out of two words with PRMT, shifted down by 6 */
template <int FS4, bool I19 = false>
__device__ __forceinline__ void s16_hfir_uv(const unsigned char *srow, int hshift, const uint32_t (&cl)[FS4],
                                            const uint32_t (&ch)[FS4], int &u, int &v)
{
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(srow);
    int ul = 0, uh = 0, vl = 0, vh = 0;
    constexpr uint32_t M = 0x03FF03FFu;
#pragma unroll
    for (int k = 0; k < FS4; k++) {
        const uint32_t w0 = wp[4 * k], w1 = wp[4 * k + 1], w2 = wp[4 * k + 2], w3 = wp[4 * k + 3];
        const uint32_t u0 = (prmt(w0, w1, 0x5410) >> 6) & M, u1 = (prmt(w2, w3, 0x5410) >> 6) & M;
        const uint32_t v0 = (prmt(w0, w1, 0x7632) >> 6) & M, v1 = (prmt(w2, w3, 0x7632) >> 6) & M;
        ul = dp2a_lo_uu(u0, cl[k], ul); uh = dp2a_lo_us(u0, ch[k], uh);
        ul = dp2a_hi_uu(u1, cl[k], ul); uh = dp2a_hi_us(u1, ch[k], uh);
        vl = dp2a_lo_uu(v0, cl[k], vl); vh = dp2a_lo_us(v0, ch[k], vh);
        vl = dp2a_hi_uu(v1, cl[k], vl); vh = dp2a_hi_us(v1, ch[k], vh);
    }
    u = min(((uh << 8) + ul) >> hshift, I19 ? (1 << 19) - 1 : (1 << 15) - 1);
    v = min(((vh << 8) + vl) >> hshift, I19 ? (1 << 19) - 1 : (1 << 15) - 1);
}

/* vertical FIR for NC columns (transposed 15-bit lines, cstep words apart): bias + sum of taps, before
 * the >> 19.  G = tap groups of four in use by this row (warp-uniform): the body is compiled once per group count
 * and picked with one switch -- the former `if (k < n4)` cascade cost 10 instructions per call and kept the
 * compiler from hoisting the column addresses.  Columns are the inner loop: 2 NC independent accumulator chains. */
template <int NC, int G>
__device__ __forceinline__ void s8_vsum_g(const uint32_t *hp, int cstep, const S8VRow &vr, int bias, int (&out)[NC])
{
    int acc_l[NC], acc_h[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        acc_l[c] = bias;
        acc_h[c] = 0;
    }
#pragma unroll
    for (int k = 0; k < G; k++) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const uint32_t w0 = hp[c * cstep + 2 * k], w1 = hp[c * cstep + 2 * k + 1];
            acc_l[c] = dp2a_lo_su(w0, vr.cl[k], acc_l[c]);
            acc_h[c] = dp2a_lo_ss(w0, vr.ch[k], acc_h[c]);
            acc_l[c] = dp2a_hi_su(w1, vr.cl[k], acc_l[c]);
            acc_h[c] = dp2a_hi_ss(w1, vr.ch[k], acc_h[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++)
        out[c] = (acc_h[c] << 8) + acc_l[c];
}

template <int NC>
__device__ __forceinline__ void s8_vsum(const uint32_t *hp, int cstep, const S8VRow &vr, int n4, int bias,
                                        int (&out)[NC])
{
    switch (n4) {
    case 1:  s8_vsum_g<NC, 1>(hp, cstep, vr, bias, out); break;
    case 2:  s8_vsum_g<NC, 2>(hp, cstep, vr, bias, out); break;
    case 3:  s8_vsum_g<NC, 3>(hp, cstep, vr, bias, out); break;
    case 4:  s8_vsum_g<NC, 4>(hp, cstep, vr, bias, out); break;
    default: s8_vsum_g<NC, S8_VF4>(hp, cstep, vr, bias, out); break;
    }
}

/* planar 8-bit output: dither 64 for 8-bit sources (swscale.c:54-56,385-387), clip */
template <int NC>
__device__ __forceinline__ void s8_vfir(const uint32_t *hp, int cstep, const S8VRow &vr, int n4, int (&out)[NC])
{
    s8_vsum<NC>(hp, cstep, vr, n4, 64 << 12, out);
#pragma unroll
    for (int c = 0; c < NC; c++)
        out[c] = clip_u8(out[c] >> 19);
}

__device__ __forceinline__ S8VRow s8_load_vrow(const S8VRow *p);

/* taps 21..40 of a row: the second record's sum added to `out` (rows without such taps have n4 = 0) */
template <int NC>
__device__ __forceinline__ void s8_vsum_more(const S8VRow *rec, const uint32_t *col0, int lo, int cstep, int (&out)[NC])
{
    const S8VRow v2 = s8_load_vrow(rec);
    const int m4 = __shfl_sync(0xffffffffu, v2.n4, 0);
    if (m4) {
        int t[NC];
        s8_vsum<NC>(col0 + (((v2.pos_even & ~1) - lo) >> 1), cstep, v2, m4, 0, t);
#pragma unroll
        for (int c = 0; c < NC; c++)
            out[c] += t[c];
    }
}

__device__ __forceinline__ S8VRow s8_load_vrow(const S8VRow *p)
{
    S8VRow r;
    const int4 *q = reinterpret_cast<const int4 *>(p);
    const int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    r.pos_even = a.x; r.n4 = a.y;
    r.cl[0] = a.z; r.cl[1] = a.w; r.cl[2] = b.x; r.cl[3] = b.y; r.cl[4] = b.z;
    r.ch[0] = b.w; r.ch[1] = c.x; r.ch[2] = c.y; r.ch[3] = c.z; r.ch[4] = c.w;
    return r;
}



/* mbarrier / TMA helpers on raw shared-window addresses (computed once per kernel) */
__device__ __forceinline__ void s8_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "S8_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra S8_DONE;\n"
        "bra S8_WAIT;\n"
        "S8_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void s8_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void s8_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s8_tma_load(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void s8_tma_prefetch(const CUtensorMap *map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}

/* horizontal FIR of one staged row of interleaved chroma (nv12 / nv21) for one output column: the
 * de-interleave of nv12ToUV_c (input.c:926-941) is two byte permutes per four taps */
template <int FS4, bool I19 = false>
__device__ __forceinline__ void s8_hfir_uv(const unsigned char *srow, int sh, const uint32_t (&cl)[FS4],
                                           const uint32_t (&ch)[FS4], int &even, int &odd)
{
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(srow);
    uint32_t w0 = wp[0];
    int el = 0, eh = 0, ol = 0, oh = 0;
#pragma unroll
    for (int k = 0; k < FS4; k++) {
        const uint32_t w1 = wp[2 * k + 1], w2 = wp[2 * k + 2];
        const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh);
        const uint32_t e = prmt(s0, s1, 0x6420), o = prmt(s0, s1, 0x7531);
        el = dp4a_uu(e, cl[k], el);
        eh = dp4a_us(e, ch[k], eh);
        ol = dp4a_uu(o, cl[k], ol);
        oh = dp4a_us(o, ch[k], oh);
        w0 = w2;
    }
    even = I19 ? min(((eh << 8) + el) >> 3, (1 << 19) - 1) : min(((eh << 8) + el) >> 7, (1 << 15) - 1);
    odd = I19 ? min(((oh << 8) + ol) >> 3, (1 << 19) - 1) : min(((oh << 8) + ol) >> 7, (1 << 15) - 1);
}

/*
 * Kernel shape.  One CTA = 128 x TH outputs, 8 filtering warps + 1 producer warp, three CTAs per SM.
 *  staging: the source rows a tile needs stream through a shared-memory ring of S8_STAGES slots,
 *           S8_ROWS rows per slot, one TMA tensor load (cp.async.bulk.tensor) per plane and slot
 *           issued by the producer warp; full/empty mbarriers, no CTA-wide barrier while filtering.
 *           Rows and columns outside the image are zero-filled by TMA; only zero taps ever read
 *           them.  The luma slots of a tile are followed by its chroma slots in the same ring.
 *  H:       thread = (output column, row group).  A thread filters PAIRS of vertically adjacent rows
 *           and stores both 15-bit results with one 32-bit shared-memory store into the transposed
 *           line buffer (column stride is an odd number of words: conflict-free for H stores and V
 *           loads).  Chroma: one thread filters U and V of its column (interleaved nv12 rows are
 *           de-interleaved on the fly).
 *  V:       warp = output row, lane = columns lane + 32k.
 */
/* MINB = 3: the same kernel compiled for three CTAs per SM (72 registers instead of 56) -- picked when the tile's
 * shared memory allows no more than three anyway (C4: 45.5 -> 47.5 % of the HBM peak; X1 / X2, which fit four,
 * lose 0.7 points with it and keep the 56-register build) */
/* RGBK: 0 planar / semi-planar YUV output, 1 packed RGB with one chroma sample per pixel pair, 2 packed RGB with full
 * horizontal chroma (its own instantiation: carried as a run-time branch it cost X1 10 %) */
template <int FS4, int RGBK, bool MMA, int SRCK, int MINB = 0>
__global__ void __launch_bounds__(S8_THREADS, MINB ? MINB : (SRCK != S8_SRC_U8 || (MMA && FS4 > 2) || RGBK == 3) ? 3 : (RGBK || FS4 <= S8_LIGHT_FS4) ? S8_RGB_CTAS : 3)
sws_scale8_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_u,
                  const __grid_constant__ CUtensorMap map_v, const __grid_constant__ Scale8Args A)
{
    constexpr bool RGB = RGBK == 1 || RGBK == 2;
    constexpr bool I19 = RGBK == 3;               /* int32 lines of 19 bits, one per word, for 16-bit planar destinations */
    constexpr bool P10 = SRCK == S8_SRC_P010;     /* p010le: 16-bit containers >> 6, chroma interleaved U first */
    constexpr bool S16 = SRCK == S8_SRC_U16 || P10;   /* 16-bit samples straight from the ring */
    constexpr int SSH = P10 ? 6 : 0;
    constexpr bool RGBS = SRCK == S8_SRC_RGB;     /* packed 8-bit RGB rows in the ring, converted to 14-bit Y/U/V samples per slot */
    extern __shared__ __align__(128) unsigned char s8_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[S8_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[S8_MAX_STAGES];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int f = blockIdx.z;

    const int TH = A.tile_h;
    const int x0 = blockIdx.x * S8_TW;
    const int ry0 = A.y0 + blockIdx.y * TH;
    const int ry1 = min(ry0 + TH, A.y1);
    const int cs = 7 - A.hs;                 /* log2 of the chroma tile width */
    const int CW = 1 << cs;
    const int cx0 = x0 >> A.hs;
    const int cy0 = ry0 >> A.vs;
    /* (a destination without chroma planes -- grayf32le -- has no chroma rows: every chroma stage below is skipped) */
    const int cy1 = A.no_chroma ? cy0 : (ry1 == A.dst_h) ? A.chr_dst_h : (ry1 >> A.vs);
    const int ch = cy1 - cy0;
    const int slot = A.slot_bytes;

    if (tid == 0) {
        for (int s = 0; s < A.stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    /* source row windows of the tile (first rows are even by construction): lane = output row */
    int lo_l = INT_MAX, hi_l = 0, lo_c = INT_MAX, hi_c = 0;
    if (I19) {
        /* plain banks: the window of a row is [pos & ~1, pos + size) (the horizontal stage works on row pairs) */
        if (ry0 + lane < ry1) {
            const int ps = __ldg(A.vl_pos32 + ry0 + lane);
            lo_l = ps & ~1;
            hi_l = ps + A.vl_size;
        }
        if (cy0 + lane < cy1) {
            const int ps = __ldg(A.vc_pos32 + cy0 + lane);
            lo_c = ps & ~1;
            hi_c = ps + A.vc_size;
        }
    }
    if (!I19 && ry0 + lane < ry1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vl + ry0 + lane));
        lo_l = pn.x & ~1;
        hi_l = lo_l + 4 * pn.y;
        if (FS4 == 8 && A.vl2) {
            const int2 p2 = __ldg(reinterpret_cast<const int2 *>(A.vl2 + ry0 + lane));
            hi_l = max(hi_l, (p2.x & ~1) + 4 * p2.y);
        }
    }
    if (!I19 && cy0 + lane < cy1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vc + cy0 + lane));
        lo_c = pn.x & ~1;            /* (bit 0 of an RGB row: a rounding flag, see the V stage) */
        hi_c = lo_c + 4 * pn.y;
        if (FS4 == 8 && A.vc2) {
            const int2 p2 = __ldg(reinterpret_cast<const int2 *>(A.vc2 + cy0 + lane));
            hi_c = max(hi_c, (p2.x & ~1) + 4 * p2.y);
        }
    }
    lo_l = __reduce_min_sync(0xffffffffu, lo_l);
    hi_l = __reduce_max_sync(0xffffffffu, hi_l);
    lo_c = __reduce_min_sync(0xffffffffu, lo_c);
    hi_c = __reduce_max_sync(0xffffffffu, hi_c);
    const int nl = min(min(hi_l, A.src_h) - lo_l, A.nl_cap);
    const int nc = ch > 0 ? min(min(hi_c, A.chr_src_h) - lo_c, A.nc_cap) : 0;
    const int npl = (nl + S8_ROWS - 1) / S8_ROWS, npc = (nc + S8_ROWS - 1) / S8_ROWS;

    /* first staged sample of the tile's rows: 16-byte aligned */
    const int a0l = __ldg(A.hl_pos + x0) & ~(15 >> A.bps);
    const int a0c = ch > 0 ? __ldg(A.hc_pos + cx0) & ~(15 >> A.bps) : 0;
    const bool planar = A.src_layout == SWSC_SRC_PLANAR;
    const uint32_t ring_a = smem_u32(s8_smem_raw), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
    /* packed RGB: luma and chroma come out of the same rows (no vertical subsampling), so ONE window of raw rows is
     * streamed: the union of both row windows (both start on even rows), from the first pixel either bank reads */
    int a0r = 0, lo_r = 0, npr = 0;
    if (RGBS) {
        a0r = __ldg(A.hl_pos + x0);
        lo_r = lo_l;
        int hi_r = lo_l + nl;
        if (ch > 0) {
            a0r = min(a0r, __ldg(A.hc_pos + cx0) << A.rgb_half);
            lo_r = min(lo_l, lo_c);
            hi_r = max(hi_r, lo_c + nc);
        }
        a0r &= ~15;
        npr = (hi_r - lo_r + S8_ROWS - 1) / S8_ROWS;
    }
    __syncthreads();

    if (warp == 8) {
        /* ===== producer: one thread keeps the ring full ===== */
        if (lane == 0) {
            int b = 0;
            uint32_t par = 1;              /* parity of the previous use of slot b */
            for (int q = 0; q < (RGBS ? npr : npl + npc); q++) {
                if (q >= A.stages)
                    s8_wait(empty_a + 8 * b, par);          /* all 8 warps released the slot */
                const uint32_t d = ring_a + b * slot, bar = full_a + 8 * b;
                if (RGBS) {
                    s8_expect_tx(bar, S8_ROWS * A.seg_l);
                    s8_tma_load(d, &map_y, bar, (a0r * A.src_bpp) >> A.elt_shift, lo_r + S8_ROWS * q, f);
                } else if (q < npl) {
                    s8_expect_tx(bar, S8_ROWS * A.seg_l);
                    s8_tma_load(d, &map_y, bar, (a0l << A.bps) >> A.elt_shift, lo_l + S8_ROWS * q, f);
                } else {
                    const int row = lo_c + S8_ROWS * (q - npl);
                    s8_expect_tx(bar, 2 * S8_ROWS * A.seg_c);
                    if (planar) {
                        const int cx = (a0c << A.bps) >> A.elt_shift;
                        s8_tma_load(d, &map_u, bar, cx, row, f);
                        s8_tma_load(d + S8_ROWS * A.seg_c, &map_v, bar, cx, row, f);
                    } else {
                        s8_tma_load(d, &map_u, bar, (a0c << (1 + A.bps)) >> A.elt_shift, row, f);   /* two samples per chroma position */
                    }
                }
                if (++b == A.stages) {
                    b = 0;
                    par ^= 1;
                }
            }
        }
        return;
    }

    uint8_t *dst0 = A.dst[0] + f * A.dst_fstride[0];
    uint8_t *dst1 = A.dst[1] + f * A.dst_fstride[1];
    uint8_t *dst2 = A.dst[2] ? A.dst[2] + f * A.dst_fstride[2] : nullptr;
    const int tw = min(S8_TW, A.dst_w - x0), th = ry1 - ry0;
    const int cw = min(CW, A.chr_dst_w - cx0);
    /* words per line-buffer column: row pairs (odd by construction), or for 19-bit lines single rows + 1 (odd) */
    const int lstride_w = I19 ? A.nl_cap + 1 : A.nl_cap >> 1, cstride_w = I19 ? A.nc_cap + 1 : A.nc_cap >> 1;
    const unsigned char *ring = s8_smem_raw;
    uint32_t *hb_l = reinterpret_cast<uint32_t *>(s8_smem_raw + A.stages * slot);
    uint32_t *hb_u = hb_l + S8_TW * lstride_w;
    uint32_t *hb_v = hb_u + CW * cstride_w;
    /* full-chroma RGB output: chroma columns are stored like the luma columns of the RGB variant, even columns in
     * slots 0..63, odd ones in 64..127, so that a lane of the V stage finds the chroma of its four pixels at the
     * luma's conflict-free stride */
    constexpr bool fullc = RGBK == 2;
    auto cslot = [&](int x) { return fullc ? (x >> 1) + 64 * (x & 1) : x; };

    /* ring position of the pass being filtered: slot sb, parity sphase */
    int sb = 0;
    uint32_t sphase = 0;
    auto release = [&]() {
        __syncwarp();
        if (lane == 0)
            s8_arrive(empty_a + 8 * sb);
        if (++sb == A.stages) {
            sb = 0;
            sphase ^= 1;
        }
    };

    /* The tensor-pipe variants are compiled without range conversion, 9..14-bit output and ordered dither (the host
     * keeps those conversions on the dot-product variants): C4 pays 12 % in instructions for the checks.  Vertical
     * banks of more than 20 taps (a second record per row) run the eight-group variants only: X3 paid 6 % for
     * carrying that code. */
    constexpr bool GEN = !MMA;
    constexpr bool LONGV = FS4 == 8;      /* second vertical records: only the eight-group variants carry the code */
    const S8Range rcl = { GEN ? A.range_mode : 0, A.lum_rc_coeff, A.lum_rc_offset };
    const S8Range rcc = { GEN ? A.range_mode : 0, A.chr_rc_coeff, A.chr_rc_offset };

    if (RGBS) {
        /* ================= packed RGB source: per ring slot  reader -> samples -> H luma + H chroma =================
         * rgb24ToY_c / rgb24ToUV[_half]_c and the 32-bit template readers (input.c:264-345,1068-1180) as two IDP.2A
         * per matrix row and pixel (the host admits only matrices whose samples stay inside 14 bits, where the
         * readers' 16-bit store and the 24- / 32-bit forms of the rounding are the same number); then
         * hScale16To15_c with shift 13 over the sample rows exactly like the 16-bit planar sources. */
        unsigned char *smp_y = reinterpret_cast<unsigned char *>(hb_v + CW * cstride_w);
        unsigned char *smp_u = smp_y + S8_ROWS * A.seg_sy;
        unsigned char *smp_v = smp_u + S8_ROWS * A.seg_sc;
        const int bpp = A.src_bpp, half = A.rgb_half;
        const int units = A.seg_l / (4 * bpp);                       /* groups of four pixels per staged row */
        /* H stage roles, fixed for the tile */
        const int xl = tid & (S8_TW - 1), gl = tid >> 7;
        const int gxl = min(x0 + xl, A.dst_w - 1);
        const int offl = __ldg(A.hl_pos + gxl) - a0r;
        const int lslot = RGB ? (xl >> 1) + 64 * (xl & 1) : xl;
        uint32_t lcl[FS4], lch[FS4], ccl[FS4], cch[FS4];
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            lcl[k] = __ldg(A.hl_cl + (size_t)gxl * FS4 + k);
            lch[k] = __ldg(A.hl_ch + (size_t)gxl * FS4 + k);
        }
        const int xc = tid & (CW - 1), gc = tid >> cs;
        const int npair = A.hs ? S8_ROWS / 8 : S8_ROWS / 4;
        const int gxc = min(cx0 + xc, A.chr_dst_w - 1);
        const int offc = ch > 0 ? __ldg(A.hc_pos + gxc) - (a0r >> half) : 0;
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            ccl[k] = ch > 0 ? __ldg(A.hc_cl + (size_t)gxc * FS4 + k) : 0;
            cch[k] = ch > 0 ? __ldg(A.hc_ch + (size_t)gxc * FS4 + k) : 0;
        }
        const int ybias = (32 << 14) + (1 << 8);
        const int cbias = half ? (256 << 15) + (1 << 9) : (256 << 14) + (1 << 8);
        const int cshift = half ? 10 : 9;
        for (int q = 0; q < npr; q++) {
            s8_wait(full_a + 8 * sb, sphase);
            const int rbase = lo_r + S8_ROWS * q;
            /* ---- reader: warp = rows warp and warp + 8 of the slot, lane = groups of four pixels ---- */
#pragma unroll 1
            for (int rr = warp; rr < S8_ROWS; rr += 8) {
                const unsigned char *raw = ring + sb * slot + rr * A.seg_l;
                uint2 *yo = reinterpret_cast<uint2 *>(smp_y + rr * A.seg_sy);
                for (int k = lane; k < units; k += 32) {
                    uint32_t px[4];
                    if (bpp == 4) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(raw + 16 * k);
                        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
                    } else {
                        const uint32_t *wq = reinterpret_cast<const uint32_t *>(raw + 12 * k);
                        const uint32_t w0 = wq[0], w1 = wq[1], w2 = wq[2];
                        px[0] = w0;
                        px[1] = __funnelshift_r(w0, w1, 24);
                        px[2] = __funnelshift_r(w1, w2, 16);
                        px[3] = w2 >> 8;
                    }
                    int y[4], u[4], v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        y[j] = dp2a_hi_su(A.yhi, px[j], dp2a_lo_su(A.ylo, px[j], ybias)) >> 9;
                        u[j] = dp2a_hi_su(A.uhi, px[j], dp2a_lo_su(A.ulo, px[j], (j & 1) && half ? u[j - 1] : cbias));
                        v[j] = dp2a_hi_su(A.vhi, px[j], dp2a_lo_su(A.vlo, px[j], (j & 1) && half ? v[j - 1] : cbias));
                    }
                    yo[k] = make_uint2(prmt((uint32_t)y[0], (uint32_t)y[1], 0x5410), prmt((uint32_t)y[2], (uint32_t)y[3], 0x5410));
                    if (half) {
                        reinterpret_cast<uint32_t *>(smp_u + rr * A.seg_sc)[k] =
                            prmt((uint32_t)(u[1] >> cshift), (uint32_t)(u[3] >> cshift), 0x5410);
                        reinterpret_cast<uint32_t *>(smp_v + rr * A.seg_sc)[k] =
                            prmt((uint32_t)(v[1] >> cshift), (uint32_t)(v[3] >> cshift), 0x5410);
                    } else {
                        reinterpret_cast<uint2 *>(smp_u + rr * A.seg_sc)[k] =
                            make_uint2(prmt((uint32_t)(u[0] >> cshift), (uint32_t)(u[1] >> cshift), 0x5410),
                                       prmt((uint32_t)(u[2] >> cshift), (uint32_t)(u[3] >> cshift), 0x5410));
                        reinterpret_cast<uint2 *>(smp_v + rr * A.seg_sc)[k] =
                            make_uint2(prmt((uint32_t)(v[0] >> cshift), (uint32_t)(v[1] >> cshift), 0x5410),
                                       prmt((uint32_t)(v[2] >> cshift), (uint32_t)(v[3] >> cshift), 0x5410));
                    }
                }
            }
            release();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            /* ---- H luma: thread = (column, half of the slot's row pairs) ---- */
            {
                const int shl = (offl & 1) * 16;
                const unsigned char *sp = smp_y + (offl >> 1) * 4;
#pragma unroll
                for (int m = 0; m < S8_ROWS / 4; m++) {
                    const int r = 2 * ((S8_ROWS / 4) * gl + m), li = rbase + r - lo_l;
                    if (li >= 0 && li < nl) {
                        int va = s16_hfir<FS4, I19>(sp + r * A.seg_sy, shl, I19 ? 9 : 13, lcl, lch);
                        int vb = s16_hfir<FS4, I19>(sp + (r + 1) * A.seg_sy, shl, I19 ? 9 : 13, lcl, lch);
                        if (I19) {
                            if (rcl.mode) {
                                va = s19_range(va, rcl.mode, (uint32_t)rcl.coeff, A.lum_rc_offset64);
                                vb = s19_range(vb, rcl.mode, (uint32_t)rcl.coeff, A.lum_rc_offset64);
                            }
                            hb_l[lslot * lstride_w + li] = (uint32_t)va;
                            hb_l[lslot * lstride_w + li + 1] = (uint32_t)vb;
                        } else {
                            if (rcl.mode) {
                                va = s8_range(va, rcl.mode, rcl.coeff, rcl.offset);
                                vb = s8_range(vb, rcl.mode, rcl.coeff, rcl.offset);
                            }
                            hb_l[lslot * lstride_w + (li >> 1)] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                        }
                    }
                }
            }
            /* ---- H chroma: thread = (column, row-pair group), both planes ---- */
            if (ch > 0) {
                const int shc = (offc & 1) * 16;
                const unsigned char *su = smp_u + (offc >> 1) * 4, *sv = smp_v + (offc >> 1) * 4;
                for (int m = 0; m < npair; m++) {
                    const int r = 2 * (npair * gc + m), ci = rbase + r - lo_c;
                    if (ci >= 0 && ci < nc) {
                        int ua = s16_hfir<FS4, I19>(su + r * A.seg_sc, shc, I19 ? 9 : 13, ccl, cch);
                        int ub = s16_hfir<FS4, I19>(su + (r + 1) * A.seg_sc, shc, I19 ? 9 : 13, ccl, cch);
                        int va = s16_hfir<FS4, I19>(sv + r * A.seg_sc, shc, I19 ? 9 : 13, ccl, cch);
                        int vb = s16_hfir<FS4, I19>(sv + (r + 1) * A.seg_sc, shc, I19 ? 9 : 13, ccl, cch);
                        if (I19) {
                            if (rcc.mode) {
                                ua = s19_range(ua, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                                ub = s19_range(ub, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                                va = s19_range(va, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                                vb = s19_range(vb, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                            }
                            hb_u[xc * cstride_w + ci] = (uint32_t)ua; hb_u[xc * cstride_w + ci + 1] = (uint32_t)ub;
                            hb_v[xc * cstride_w + ci] = (uint32_t)va; hb_v[xc * cstride_w + ci + 1] = (uint32_t)vb;
                        } else {
                            if (rcc.mode) {
                                ua = s8_range(ua, rcc.mode, rcc.coeff, rcc.offset); ub = s8_range(ub, rcc.mode, rcc.coeff, rcc.offset);
                                va = s8_range(va, rcc.mode, rcc.coeff, rcc.offset); vb = s8_range(vb, rcc.mode, rcc.coeff, rcc.offset);
                            }
                            hb_u[cslot(xc) * cstride_w + (ci >> 1)] = prmt((uint32_t)ua, (uint32_t)ub, 0x5410);
                            hb_v[cslot(xc) * cstride_w + (ci >> 1)] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");      /* the sample rows are rewritten by the next slot */
        }
    } else {
    /* ================= stage H, luma ================= */
    if (MMA) {
        /* warp = 16 output columns (two groups of 8), all 16 rows of a slot per pass: FS4 = K steps of 32 bytes */
        const int t = lane & 3, g = lane >> 2;
        const int grp = (x0 >> 3) + 2 * warp;
        S8Bfrag<FS4> b0, b1;
        s8_load_bfrag<FS4>(A.hl_B, grp, lane, b0);
        s8_load_bfrag<FS4>(A.hl_B, grp + 1, lane, b1);
        /* ldmatrix row address of this lane: matrix = lane >> 3 (bit 0: odd source rows, bit 1: bytes 16..31) */
        const uint32_t lrow = s8_ldm_row(lane) * A.seg_l + (lane >> 4) * 16;
        const uint32_t sel = S8_PAIR_SEL(lane);
        const uint32_t o0 = lrow + __ldg(A.hl_goff + grp), o1 = lrow + __ldg(A.hl_goff + grp + 1);
        const int c0 = 16 * warp + 2 * t, c1 = c0 + 8;           /* tile columns of the two words per group */
        uint32_t *h00 = hb_l + (RGB ? (c0 >> 1) : c0) * lstride_w + g;
        uint32_t *h01 = hb_l + (RGB ? (c0 >> 1) + 64 : c0 + 1) * lstride_w + g;
        uint32_t *h10 = hb_l + (RGB ? (c1 >> 1) : c1) * lstride_w + g;
        uint32_t *h11 = hb_l + (RGB ? (c1 >> 1) + 64 : c1 + 1) * lstride_w + g;
        int left = nl - 2 * g;                                   /* rows of this lane's pair still inside the window */
        for (int q = 0; q < npl; q++) {
            s8_wait(full_a + 8 * sb, sphase);
            const uint32_t base = ring_a + sb * slot;
            uint32_t wa, wb, wc, wd;
            s8_mma_rows<FS4>(base + o0, b0, sel, rcl, wa, wb);
            s8_mma_rows<FS4>(base + o1, b1, sel, rcl, wc, wd);
            if (left > 0) {
                h00[0] = wa; h01[0] = wb; h10[0] = wc; h11[0] = wd;
            }
            h00 += S8_ROWS / 2; h01 += S8_ROWS / 2; h10 += S8_ROWS / 2; h11 += S8_ROWS / 2;
            left -= S8_ROWS;
            release();
        }
    } else {
        /* thread = (column, half of the slot's rows) */
        constexpr int NP = S8_ROWS / 4;          /* row pairs per thread and pass */
        const int x = tid & (S8_TW - 1), g = tid >> 7;
        const int gx = min(x0 + x, A.dst_w - 1);
        const int off = __ldg(A.hl_pos + gx) - a0l;
        const int sh = S16 ? (off & 1) * 16 : (off & 3) * 8;
        uint32_t cl[FS4], chh[FS4];
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            cl[k] = __ldg(A.hl_cl + (size_t)gx * FS4 + k);
            chh[k] = __ldg(A.hl_ch + (size_t)gx * FS4 + k);
        }
        const int seg = A.seg_l;
        const int so = 2 * NP * g * seg + (S16 ? (off >> 1) * 4 : (off & ~3));
        /* RGB output: even columns in slots 0..63, odd columns in 64..127, so that a lane of the V stage
         * finds both pixels of its pair at a conflict-free stride */
        const int lslot = RGB ? (x >> 1) + 64 * (x & 1) : x;
        uint32_t *hp = hb_l + lslot * lstride_w + (I19 ? 2 : 1) * NP * g;
        int left = nl - 2 * NP * g;              /* rows of this thread's group still inside the window */
        for (int q = 0; q < npl; q++) {
            s8_wait(full_a + 8 * sb, sphase);
            const unsigned char *sp = ring + sb * slot + so;
#pragma unroll
            for (int m = 0; m < NP; m++) {
                if (2 * m < left) {
                    int va, vb;
                    if (S16) {
                        va = s16_hfir<FS4, I19, SSH>(sp + (2 * m) * seg, sh, A.h_shift, cl, chh);
                        vb = s16_hfir<FS4, I19, SSH>(sp + (2 * m + 1) * seg, sh, A.h_shift, cl, chh);
                    } else {
                        va = s8_hfir<FS4, I19>(sp + (2 * m) * seg, sh, cl, chh);
                        vb = s8_hfir<FS4, I19>(sp + (2 * m + 1) * seg, sh, cl, chh);
                    }
                    if (I19) {
                        if (rcl.mode) {
                            va = s19_range(va, rcl.mode, (uint32_t)rcl.coeff, A.lum_rc_offset64);
                            vb = s19_range(vb, rcl.mode, (uint32_t)rcl.coeff, A.lum_rc_offset64);
                        }
                        hp[2 * m] = (uint32_t)va;
                        hp[2 * m + 1] = (uint32_t)vb;
                    } else {
                        if (rcl.mode) {
                            va = s8_range(va, rcl.mode, rcl.coeff, rcl.offset);
                            vb = s8_range(vb, rcl.mode, rcl.coeff, rcl.offset);
                        }
                        hp[m] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                    }
                }
            }
            hp += I19 ? S8_ROWS : S8_ROWS / 2;
            left -= S8_ROWS;
            release();
        }
    }
    /* ================= stage H, chroma ================= */
    if (MMA) {
        if (npc > 0) {
            /* hs = 1: 8 groups per plane, warp = group `warp` of U and of V; hs = 0: 16 groups, warp = two of them */
            const int t = lane & 3, g = lane >> 2;
            const int ng = A.hs ? 1 : 2;
            const int grp = (cx0 >> 3) + ng * warp;
            const bool vfirst = A.src_layout == SWSC_SRC_NV21;
            const int rowbytes = planar ? A.seg_c : 2 * A.seg_c;
            const uint32_t lrow = s8_ldm_row(lane) * rowbytes + (lane >> 4) * 16;
            const uint32_t sel = S8_PAIR_SEL(lane);
            S8Bfrag<FS4> b0, b1;
            s8_load_bfrag<FS4>(A.hc_B, grp, lane, b0);
            uint32_t o0 = lrow + __ldg(A.hc_goff + grp), o1 = 0;
            if (ng == 2) {
                s8_load_bfrag<FS4>(A.hc_B, grp + 1, lane, b1);
                o1 = lrow + __ldg(A.hc_goff + grp + 1);
            }
            const int c0 = 8 * ng * warp + 2 * t;
            uint32_t *hu = hb_u + g, *hv = hb_v + g;
            /* line-buffer slots (in words) of columns c0, c0 + 1, c0 + 8, c0 + 9 */
            const int s0 = cslot(c0) * cstride_w, s1 = cslot(c0 + 1) * cstride_w;
            const int s8 = cslot(c0 + 8) * cstride_w, s9 = cslot(c0 + 9) * cstride_w;
            int left = nc - 2 * g;
            for (int qc = 0; qc < npc; qc++) {
                s8_wait(full_a + 8 * sb, sphase);
                const uint32_t base = ring_a + sb * slot;
                uint32_t ua, ub, va, vb;
                if (planar) {
                    s8_mma_rows<FS4>(base + o0, b0, sel, rcc, ua, ub);
                    s8_mma_rows<FS4>(base + S8_ROWS * A.seg_c + o0, b0, sel, rcc, va, vb);
                } else if (vfirst) {
                    s8_mma_rows_uv<FS4>(base + o0, b0, sel, rcc, va, vb, ua, ub);
                } else {
                    s8_mma_rows_uv<FS4>(base + o0, b0, sel, rcc, ua, ub, va, vb);
                }
                if (left > 0) {
                    hu[s0] = ua; hu[s1] = ub; hv[s0] = va; hv[s1] = vb;
                }
                if (ng == 2) {
                    if (planar) {
                        s8_mma_rows<FS4>(base + o1, b1, sel, rcc, ua, ub);
                        s8_mma_rows<FS4>(base + S8_ROWS * A.seg_c + o1, b1, sel, rcc, va, vb);
                    } else if (vfirst) {
                        s8_mma_rows_uv<FS4>(base + o1, b1, sel, rcc, va, vb, ua, ub);
                    } else {
                        s8_mma_rows_uv<FS4>(base + o1, b1, sel, rcc, ua, ub, va, vb);
                    }
                    if (left > 0) {
                        hu[s8] = ua; hu[s9] = ub; hv[s8] = va; hv[s9] = vb;
                    }
                }
                hu += S8_ROWS / 2; hv += S8_ROWS / 2;
                left -= S8_ROWS;
                release();
            }
        }
    } else if (npc > 0) {
        const int x = tid & (CW - 1), g = tid >> cs;
        const int npair = A.hs ? S8_ROWS / 8 : S8_ROWS / 4;          /* row pairs per thread and pass */
        const int gx = min(cx0 + x, A.chr_dst_w - 1);
        const int off = __ldg(A.hc_pos + gx) - a0c;
        uint32_t cl[FS4], chh[FS4];
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            cl[k] = __ldg(A.hc_cl + (size_t)gx * FS4 + k);
            chh[k] = __ldg(A.hc_ch + (size_t)gx * FS4 + k);
        }
        const int seg = A.seg_c;
        const bool vfirst = A.src_layout == SWSC_SRC_NV21;
        uint32_t *hpu = hb_u + cslot(x) * cstride_w + (I19 ? 2 : 1) * npair * g;
        uint32_t *hpv = hb_v + cslot(x) * cstride_w + (I19 ? 2 : 1) * npair * g;
        const int rowbytes = planar ? seg : 2 * seg;
        const int so = 2 * npair * g * rowbytes + (P10 ? off * 4 : S16 ? (off >> 1) * 4 : planar ? (off & ~3) : ((2 * off) & ~3));
        const int sh = S16 ? (off & 1) * 16 : planar ? (off & 3) * 8 : (off & 1) * 16;
        int left = nc - 2 * npair * g;
        for (int qc = 0; qc < npc; qc++) {
            s8_wait(full_a + 8 * sb, sphase);
            const unsigned char *sp = ring + sb * slot + so;
            for (int m = 0; m < npair; m++) {
                if (2 * m < left) {
                    int ua, ub, va, vb;
                    if (P10) {               /* interleaved 16-bit chroma */
                        s16_hfir_uv<FS4, I19>(sp, A.h_shift, cl, chh, ua, va);
                        s16_hfir_uv<FS4, I19>(sp + rowbytes, A.h_shift, cl, chh, ub, vb);
                    } else if (S16) {        /* planar 16-bit chroma */
                        ua = s16_hfir<FS4, I19>(sp, sh, A.h_shift, cl, chh);
                        ub = s16_hfir<FS4, I19>(sp + seg, sh, A.h_shift, cl, chh);
                        va = s16_hfir<FS4, I19>(sp + S8_ROWS * seg, sh, A.h_shift, cl, chh);
                        vb = s16_hfir<FS4, I19>(sp + (S8_ROWS + 1) * seg, sh, A.h_shift, cl, chh);
                    } else if (planar) {
                        ua = s8_hfir<FS4, I19>(sp, sh, cl, chh);
                        ub = s8_hfir<FS4, I19>(sp + seg, sh, cl, chh);
                        va = s8_hfir<FS4, I19>(sp + S8_ROWS * seg, sh, cl, chh);
                        vb = s8_hfir<FS4, I19>(sp + (S8_ROWS + 1) * seg, sh, cl, chh);
                    } else {
                        s8_hfir_uv<FS4, I19>(sp, sh, cl, chh, ua, va);
                        s8_hfir_uv<FS4, I19>(sp + 2 * seg, sh, cl, chh, ub, vb);
                        if (vfirst) {
                            int t = ua; ua = va; va = t;
                            t = ub; ub = vb; vb = t;
                        }
                    }
                    if (I19) {
                        if (rcc.mode) {
                            ua = s19_range(ua, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                            ub = s19_range(ub, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                            va = s19_range(va, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                            vb = s19_range(vb, rcc.mode, (uint32_t)rcc.coeff, A.chr_rc_offset64);
                        }
                        hpu[2 * m] = (uint32_t)ua; hpu[2 * m + 1] = (uint32_t)ub;
                        hpv[2 * m] = (uint32_t)va; hpv[2 * m + 1] = (uint32_t)vb;
                    } else {
                        if (rcc.mode) {
                            ua = s8_range(ua, rcc.mode, rcc.coeff, rcc.offset); ub = s8_range(ub, rcc.mode, rcc.coeff, rcc.offset);
                            va = s8_range(va, rcc.mode, rcc.coeff, rcc.offset); vb = s8_range(vb, rcc.mode, rcc.coeff, rcc.offset);
                        }
                        hpu[m] = prmt((uint32_t)ua, (uint32_t)ub, 0x5410);
                        hpv[m] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                    }
                }
                sp += 2 * rowbytes;
            }
            hpu += I19 ? S8_ROWS : S8_ROWS / 2;
            hpv += I19 ? S8_ROWS : S8_ROWS / 2;
            left -= S8_ROWS;
            release();
        }
    }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");      /* the 8 filtering warps: all h-scaled lines are in place */

    if (RGB) {
        /* ============ stage V + yuv2rgb, packed RGB with one chroma sample per pixel pair: warp = row,
         * lane = pairs lane and lane + 32 (yuv2rgb_X/_1/_2_c_template + yuv2rgb_write, output.c:1662-1939;
         * the byte LUTs in closed form as in the other RGB kernels) ============ */
        const int kind = A.dst_kind;
        const bool rgb16 = kind >= SWSC_DST_RGB565;              /* 15/16 bpp: rgb565le, bgr565le, rgb555le, bgr555le */
        const int bpp = rgb16 ? 2 : kind >= SWSC_DST_RGBA ? 4 : 3;
        unsigned char *orow = s8_smem_raw + warp * 512;          /* the ring is idle now: 512 B of row staging per warp */
        const int cy = A.cy, yb = A.yb;
        S8VRow vl, vc;
        if (warp < th) {
            vl = s8_load_vrow(A.vl + ry0 + warp);
            vc = s8_load_vrow(A.vc + ry0 + warp);
        }
        for (int ty = warp; ty < th; ty += 8) {
            const int y = ry0 + ty;
            S8VRow nl_, nc_;
            if (ty + 8 < th) {
                nl_ = s8_load_vrow(A.vl + y + 8);
                nc_ = s8_load_vrow(A.vc + y + 8);
            }
            if (fullc) {
                /* ---- yuv2rgb_full_{X,2,1}_c_template + yuv2rgb_write_full (output.c:1998-2051,2160-2330): the sums
                 * keep 10 fraction bits less; the 2-tap writers drop the rounding bias 1 << 9 (luma and chroma for
                 * _2: bit 0 of the luma row; chroma only for _1 with two chroma taps: bit 0 of the chroma row) ---- */
                const int lb = (vl.pos_even & 1) ? 0 : 1 << 9;
                const int cb = ((vl.pos_even | vc.pos_even) & 1) ? 0 : 1 << 9;
                const int ln4f = __shfl_sync(0xffffffffu, vl.n4, 0), cn4f = __shfl_sync(0xffffffffu, vc.n4, 0);
                const int coff = ((vc.pos_even & ~1) - lo_c) >> 1;
                int Yf[4], Uf[4], Vf[4];
                s8_vsum<4>(hb_l + lane * lstride_w + (((vl.pos_even & ~1) - lo_l) >> 1), 32 * lstride_w, vl, ln4f, lb, Yf);
                s8_vsum<4>(hb_u + lane * cstride_w + coff, 32 * cstride_w, vc, cn4f, cb - (128 << 19), Uf);
                s8_vsum<4>(hb_v + lane * cstride_w + coff, 32 * cstride_w, vc, cn4f, cb - (128 << 19), Vf);
                if (LONGV && A.vl2)
                    s8_vsum_more<4>(A.vl2 + y, hb_l + lane * lstride_w, lo_l, 32 * lstride_w, Yf);
                if (LONGV && A.vc2) {
                    s8_vsum_more<4>(A.vc2 + y, hb_u + lane * cstride_w, lo_c, 32 * cstride_w, Uf);
                    s8_vsum_more<4>(A.vc2 + y, hb_v + lane * cstride_w, lo_c, 32 * cstride_w, Vf);
                }
                __syncwarp();
                uint32_t px[4];                 /* B | G << 8 | R << 16 of pixels 2 lane, 2 (lane + 32), 2 lane + 1, 2 (lane + 32) + 1 */
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int Yi = Yf[c] >> 10, Ui = Uf[c] >> 10, Vi = Vf[c] >> 10;
                    const unsigned Yu = (unsigned)(Yi - A.y_offset) * (unsigned)A.y_coeff + (1u << 21);
                    const int R = (int)(Yu + (unsigned)Vi * (unsigned)A.v2r);
                    const int G = (int)(Yu + (unsigned)Vi * (unsigned)A.v2g + (unsigned)Ui * (unsigned)A.u2g);
                    const int B = (int)(Yu + (unsigned)Ui * (unsigned)A.u2b);
                    const uint32_t r = (uint32_t)min(max(R, 0), (1 << 30) - 1) >> 22, gq = (uint32_t)min(max(G, 0), (1 << 30) - 1) >> 22;
                    const uint32_t bq = (uint32_t)min(max(B, 0), (1 << 30) - 1) >> 22;
                    px[c] = bq | (gq << 8) | (r << 16);
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const uint32_t a = px[k], b = px[k + 2];       /* the two pixels of pair lane + 32 k */
                    const int p = lane + 32 * k;
                    if (bpp == 3) {
                        uint16_t *o = reinterpret_cast<uint16_t *>(orow + 6 * p);
                        if (kind == SWSC_DST_RGB24) {           /* R G B R G B */
                            o[0] = (uint16_t)prmt(a, 0u, 0x4412); o[1] = (uint16_t)prmt(a, b, 0x4460); o[2] = (uint16_t)prmt(b, 0u, 0x4401);
                        } else {                                /* B G R B G R */
                            o[0] = (uint16_t)prmt(a, 0u, 0x4410); o[1] = (uint16_t)prmt(a, b, 0x4442); o[2] = (uint16_t)prmt(b, 0u, 0x4421);
                        }
                    } else {
                        uint32_t wa, wb;
                        if (kind == SWSC_DST_RGBA) {            /* R G B 255 */
                            wa = prmt(a, 0xFFu, 0x4012); wb = prmt(b, 0xFFu, 0x4012);
                        } else if (kind == SWSC_DST_BGRA) {     /* B G R 255 */
                            wa = prmt(a, 0xFFu, 0x4210); wb = prmt(b, 0xFFu, 0x4210);
                        } else if (kind == SWSC_DST_ARGB) {     /* 255 R G B */
                            wa = prmt(a, 0xFFu, 0x0124); wb = prmt(b, 0xFFu, 0x0124);
                        } else {                                /* 255 B G R */
                            wa = prmt(a, 0xFFu, 0x2104); wb = prmt(b, 0xFFu, 0x2104);
                        }
                        *reinterpret_cast<uint2 *>(orow + 8 * p) = make_uint2(wa, wb);
                    }
                }
            } else {
            const int bias = (vl.pos_even & 1) ? 0 : 1 << 18;
            int Y[4], U[2], V[2];
            /* this row's own tap-group counts (a leading zero tap pads odd first rows), made warp-uniform */
            const int ln4 = __shfl_sync(0xffffffffu, vl.n4, 0), cn4 = __shfl_sync(0xffffffffu, vc.n4, 0);
            s8_vsum<4>(hb_l + lane * lstride_w + (((vl.pos_even & ~1) - lo_l) >> 1), 32 * lstride_w, vl, ln4, bias, Y);
            s8_vsum<2>(hb_u + lane * cstride_w + ((vc.pos_even - lo_c) >> 1), 32 * cstride_w, vc, cn4, bias, U);
            s8_vsum<2>(hb_v + lane * cstride_w + ((vc.pos_even - lo_c) >> 1), 32 * cstride_w, vc, cn4, bias, V);
            if (LONGV && A.vl2)
                s8_vsum_more<4>(A.vl2 + y, hb_l + lane * lstride_w, lo_l, 32 * lstride_w, Y);
            if (LONGV && A.vc2) {
                s8_vsum_more<2>(A.vc2 + y, hb_u + lane * cstride_w, lo_c, 32 * cstride_w, U);
                s8_vsum_more<2>(A.vc2 + y, hb_v + lane * cstride_w, lo_c, 32 * cstride_w, V);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int y1v = Y[k] >> 19, y2v = Y[k + 2] >> 19;
                const int u8 = clamp_u8(U[k] >> 19), v8 = clamp_u8(V[k] >> 19);
                const int pR = (A.base_r + ((v8 * A.crv) >> 16)) * cy + yb;
                const int pG = (A.base_g + ((u8 * A.cgu) >> 16) + ((v8 * A.cgv) >> 16)) * cy + yb;
                const int pB = (A.base_b + ((u8 * A.cbu) >> 16)) * cy + yb;
                const uint32_t tRa = y1v * cy + pR, tGa = y1v * cy + pG, tBa = y1v * cy + pB;
                const uint32_t tRb = y2v * cy + pR, tGb = y2v * cy + pG, tBb = y2v * cy + pB;
                constexpr uint32_t FF = 0x00FF0000u;
                const int p = lane + 32 * k;
                if (rgb16) {
                    /* yuv2rgb_write, 15/16 bpp (output.c:1714-1747; tables yuv2rgb.c:878-900): the 2 x 2 ordered-dither
                     * offsets move the LUT index (ff_dither_2x2_8 = {6,2 / 0,4}, ff_dither_2x2_4 = {1,3 / 2,0}), then
                     * the bytes are truncated into their fields */
                    const int odd = y & 1;
                    const bool is565 = kind <= SWSC_DST_BGR565;
                    const int dr1 = odd ? 0 : 6, dr2 = odd ? 4 : 2, db1 = odd ? 6 : 0, db2 = odd ? 2 : 4;
                    const int dg1 = is565 ? (odd ? 2 : 1) : dr2, dg2 = is565 ? (odd ? 0 : 3) : dr1;
                    const int gsh = is565 ? 2 : 3, hi = is565 ? 11 : 10;
                    const int r1 = clamp_u8((int)(tRa + dr1 * cy) >> 16) >> 3, r2 = clamp_u8((int)(tRb + dr2 * cy) >> 16) >> 3;
                    const int g1 = clamp_u8((int)(tGa + dg1 * cy) >> 16) >> gsh, g2 = clamp_u8((int)(tGb + dg2 * cy) >> 16) >> gsh;
                    const int b1 = clamp_u8((int)(tBa + db1 * cy) >> 16) >> 3, b2 = clamp_u8((int)(tBb + db2 * cy) >> 16) >> 3;
                    const bool rfirst = kind == SWSC_DST_RGB565 || kind == SWSC_DST_RGB555;   /* R in the high bits */
                    const uint32_t p1 = rfirst ? (r1 << hi) | (g1 << 5) | b1 : (b1 << hi) | (g1 << 5) | r1;
                    const uint32_t p2 = rfirst ? (r2 << hi) | (g2 << 5) | b2 : (b2 << hi) | (g2 << 5) | r2;
                    *reinterpret_cast<uint32_t *>(orow + 4 * p) = p1 | (p2 << 16);
                } else if (bpp == 3) {
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
            vl = nl_;
            vc = nc_;
        }
        return;
    }

    if (I19 && (A.dst_kind == SWSC_DST_RGB48 || A.dst_kind == SWSC_DST_BGR48 || A.dst_kind == SWSC_DST_GBRP)) {
        /* ============ stage V + colour step over 19-bit lines, rgb48le / bgr48le: yuv2rgba64_{X,2,1}_c_template
         * (output.c:1115-1300; one chroma sample per pixel pair) and yuv2rgba64_full_{X,2,1}_c_template (:1373-1560; one per
         * pixel), 32-bit wrap-around exactly like the C templates.  warp = row; the row's taps sit in two registers per
         * lane and are broadcast per tap.  Rows the reference hands to the _1 writers (one luma tap, one chroma tap or a
         * bilinear chroma pair: vscale.c:135-147) shift the line itself instead of multiplying by 4096. ============ */
        const bool swap = A.dst_kind == SWSC_DST_BGR48;
        /* gbrpf32le: yuv2gbrpf32_full_X_c (output.c:2536-2610) -- always the X form with the real taps, the 16-bit result
         * times 1.0f / 65535.0f (one IEEE multiply: bit-exact), planes G, B, R */
        const bool gbrpf = A.dst_kind == SWSC_DST_GBRP;
        const int lfs = A.vl_size, cfs = A.vc_size;
        const unsigned yofs = (unsigned)A.y_offset, ycf = (unsigned)A.y_coeff;
        const unsigned v2r = (unsigned)A.v2r, v2g = (unsigned)A.v2g, u2g = (unsigned)A.u2g, u2b = (unsigned)A.u2b;
        auto px16 = [&](unsigned yu, unsigned uu, unsigned vu, int &r, int &g, int &b) {
            yu = (yu - yofs) * ycf + (1u << 13) - (1u << 29);
            r = clip_uintp2(((int)(vu * v2r + yu) >> 14) + (1 << 15), 16);
            g = clip_uintp2(((int)(vu * v2g + uu * u2g + yu) >> 14) + (1 << 15), 16);
            b = clip_uintp2(((int)(uu * u2b + yu) >> 14) + (1 << 15), 16);
        };
        for (int ty = warp; ty < th; ty += 8) {
            const int y = ry0 + ty;
            const int16_t *lf = A.vl_coef16 + (size_t)y * lfs, *cf = A.vc_coef16 + (size_t)y * cfs;
            const int l0 = lane < lfs ? (int)__ldg(lf + lane) : 0, l1 = lane + 32 < lfs ? (int)__ldg(lf + lane + 32) : 0;
            const int k0 = lane < cfs ? (int)__ldg(cf + lane) : 0, k1 = lane + 32 < cfs ? (int)__ldg(cf + lane + 32) : 0;
            const int cf0 = __shfl_sync(0xffffffffu, k0, 0), cf1 = __shfl_sync(0xffffffffu, k0, 1);
            const bool chr2 = cfs == 2 && cf0 + cf1 == 4096 && (unsigned)cf1 <= 4096u;
            const bool one = !gbrpf && lfs == 1 && (cfs == 1 || chr2);          /* the _1 writers */
            const bool one_c = one && (cfs == 1 || cf1 == 0);         /* ... with uvalpha == 0: chroma from the line too */
            const int rl = __ldg(A.vl_pos32 + y) - lo_l, rc = __ldg(A.vc_pos32 + y) - lo_c;
            uint8_t *drow = dst0 + (size_t)y * A.dst_stride[0];
            if (!A.full_chr) {
                /* lane = pixel pairs lane and lane + 32: luma columns 2 lane, 2 lane + 1, 64 + 2 lane, 65 + 2 lane */
                const uint32_t *pl = hb_l + 2 * lane * lstride_w + rl;
                const uint32_t *pu = hb_u + lane * cstride_w + rc, *pv = hb_v + lane * cstride_w + rc;
                unsigned Y[4] = { 0, 0, 0, 0 }, U[2] = { 0, 0 }, V[2] = { 0, 0 };
                for (int j = 0; j < lfs; j++) {
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, j < 32 ? l0 : l1, j & 31);
                    Y[0] += pl[j] * c; Y[1] += pl[lstride_w + j] * c;
                    Y[2] += pl[64 * lstride_w + j] * c; Y[3] += pl[65 * lstride_w + j] * c;
                }
                for (int j = 0; j < cfs; j++) {
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, j < 32 ? k0 : k1, j & 31);
                    U[0] += pu[j] * c; U[1] += pu[32 * cstride_w + j] * c;
                    V[0] += pv[j] * c; V[1] += pv[32 * cstride_w + j] * c;
                }
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    unsigned y1 = (unsigned)((int)(Y[2 * k] - 0x40000000u) >> 14) + 0x10000u;
                    unsigned y2 = (unsigned)((int)(Y[2 * k + 1] - 0x40000000u) >> 14) + 0x10000u;
                    unsigned uu = (unsigned)((int)(U[k] - (128u << 23)) >> 14), vu = (unsigned)((int)(V[k] - (128u << 23)) >> 14);
                    if (one) {
                        y1 = (unsigned)((int)pl[64 * k * lstride_w] >> 2);
                        y2 = (unsigned)((int)pl[(64 * k + 1) * lstride_w] >> 2);
                        if (one_c) {
                            uu = (unsigned)(((int)pu[32 * k * cstride_w] - (128 << 11)) >> 2);
                            vu = (unsigned)(((int)pv[32 * k * cstride_w] - (128 << 11)) >> 2);
                        }
                    }
                    int r1, g1, b1, r2, g2, b2;
                    px16(y1, uu, vu, r1, g1, b1);
                    px16(y2, uu, vu, r2, g2, b2);
                    const int xp = 2 * (lane + 32 * k);                  /* first pixel of the pair inside the tile */
                    if (xp < tw) {
                        uint16_t *w = reinterpret_cast<uint16_t *>(drow) + 3 * (x0 + xp);
                        const int f1 = swap ? b1 : r1, t1 = swap ? r1 : b1, f2 = swap ? b2 : r2, t2 = swap ? r2 : b2;
                        if (((uintptr_t)w & 3) == 0) {
                            uint32_t *w32 = reinterpret_cast<uint32_t *>(w);
                            w32[0] = (uint32_t)f1 | ((uint32_t)g1 << 16);
                            w32[1] = (uint32_t)t1 | ((uint32_t)f2 << 16);
                            w32[2] = (uint32_t)g2 | ((uint32_t)t2 << 16);
                        } else {
                            w[0] = (uint16_t)f1; w[1] = (uint16_t)g1; w[2] = (uint16_t)t1;
                            w[3] = (uint16_t)f2; w[4] = (uint16_t)g2; w[5] = (uint16_t)t2;
                        }
                    }
                }
            } else {
                /* lane = pixels lane + 32 c, every one with its own chroma sample */
                const uint32_t *pl = hb_l + lane * lstride_w + rl;
                const uint32_t *pu = hb_u + lane * cstride_w + rc, *pv = hb_v + lane * cstride_w + rc;
                /* yuv2rgba64_full_1_c_template with uvalpha != 0 keeps U and V unsigned: its >> 14 is a logical shift */
                const bool ulog = !gbrpf && lfs == 1 && cfs == 2 && cf0 + cf1 == 4096 && cf1 > 0 && cf1 <= 4096;
                unsigned Y[4] = { 0, 0, 0, 0 }, U[4] = { 0, 0, 0, 0 }, V[4] = { 0, 0, 0, 0 };
                for (int j = 0; j < lfs; j++) {
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, j < 32 ? l0 : l1, j & 31);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        Y[q] += pl[32 * q * lstride_w + j] * c;
                }
                for (int j = 0; j < cfs; j++) {
                    const unsigned c = (unsigned)__shfl_sync(0xffffffffu, j < 32 ? k0 : k1, j & 31);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        U[q] += pu[32 * q * cstride_w + j] * c;
                        V[q] += pv[32 * q * cstride_w + j] * c;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    unsigned yv = (unsigned)((int)(Y[q] - 0x40000000u) >> 14) + 0x10000u;
                    unsigned uu = ulog ? (U[q] - (128u << 23)) >> 14 : (unsigned)((int)(U[q] - (128u << 23)) >> 14);
                    unsigned vu = ulog ? (V[q] - (128u << 23)) >> 14 : (unsigned)((int)(V[q] - (128u << 23)) >> 14);
                    if (one) {
                        yv = (unsigned)((int)pl[32 * q * lstride_w] >> 2);
                        if (one_c) {
                            uu = (unsigned)(((int)pu[32 * q * cstride_w] - (128 << 11)) >> 2);
                            vu = (unsigned)(((int)pv[32 * q * cstride_w] - (128 << 11)) >> 2);
                        }
                    }
                    int r, g, b;
                    px16(yv, uu, vu, r, g, b);
                    if (lane + 32 * q < tw) {
                        if (gbrpf) {
                            const int gx = x0 + lane + 32 * q;
                            reinterpret_cast<float *>(drow)[gx] = __fmul_rn(1.0f / 65535.0f, (float)g);
                            reinterpret_cast<float *>(dst1 + (size_t)y * A.dst_stride[1])[gx] = __fmul_rn(1.0f / 65535.0f, (float)b);
                            reinterpret_cast<float *>(dst2 + (size_t)y * A.dst_stride[2])[gx] = __fmul_rn(1.0f / 65535.0f, (float)r);
                        } else {
                            uint16_t *w = reinterpret_cast<uint16_t *>(drow) + 3 * (x0 + lane + 32 * q);
                            w[0] = (uint16_t)(swap ? b : r); w[1] = (uint16_t)g; w[2] = (uint16_t)(swap ? r : b);
                        }
                    }
                }
            }
        }
        return;
    }

    if (I19) {
        /* ============ stage V over 19-bit lines: yuv2planeX_16_c / yuv2plane1_16_c (output.c:163-187), warp = row (luma)
         * or (plane, row) (chroma), lane = columns lane + 32 k ============ */
        for (int ty = warp; ty < th; ty += 8) {
            const int y = ry0 + ty;
            const uint32_t *col = hb_l + lane * lstride_w + (__ldg(A.vl_pos32 + y) - lo_l);
            int v[S8_TW / 32];
            s19_vrow<S8_TW / 32>(col, 32 * lstride_w, A.vl_coef16 + (size_t)y * A.vl_size, A.vl_size, lane, v);
            if (A.dst_kind == SWSC_DST_PLANARF32) {
                /* yuv2plane1_float_c / yuv2planeX_float_c (output.c:219-263): the 16-bit value times 1.0f / 65535.0f */
                float *d = reinterpret_cast<float *>(dst0 + (size_t)y * A.dst_stride[0]) + x0 + lane;
#pragma unroll
                for (int c = 0; c < S8_TW / 32; c++)
                    if (lane + 32 * c < tw)
                        d[32 * c] = __fmul_rn(1.0f / 65535.0f, (float)v[c]);
            } else {
                uint16_t *d = reinterpret_cast<uint16_t *>(dst0 + (size_t)y * A.dst_stride[0]) + x0 + lane;
#pragma unroll
                for (int c = 0; c < S8_TW / 32; c++)
                    if (lane + 32 * c < tw)
                        d[32 * c] = (uint16_t)v[c];
            }
        }
        for (int task = warp; task < 2 * ch; task += 8) {
            const int pl = task & 1, y = cy0 + (task >> 1);
            const uint32_t *col = (pl ? hb_v : hb_u) + lane * cstride_w + (__ldg(A.vc_pos32 + y) - lo_c);
            uint16_t *d = reinterpret_cast<uint16_t *>((pl ? dst2 : dst1) + (size_t)y * A.dst_stride[pl ? 2 : 1]) + cx0 + lane;
            for (int c = 0; 32 * c < CW; c += 2) {
                int v[2];
                s19_vrow<2>(col + 32 * c * cstride_w, 32 * cstride_w, A.vc_coef16 + (size_t)y * A.vc_size, A.vc_size, lane, v);
                if (lane + 32 * c < cw)
                    d[32 * c] = (uint16_t)v[0];
                if (lane + 32 * c + 32 < cw)
                    d[32 * (c + 1)] = (uint16_t)v[1];
            }
        }
        return;
    }

    /* ================= stage V, luma: warp = row, lane = columns lane, lane+32, ... ================= */
    const int obits = GEN ? A.out_bits : 8, oshift = 27 - obits;
    const bool bayer = GEN && A.dither_bayer != 0;
    const int olsh = GEN ? A.out_lshift : 0;       /* p010le: yuv2p010l1/lX_c, yuv2p010cX_c (output.c:538-589) */
    /* Both loops below keep the destination pointer of the warp's row in registers and step it (recomputing it from the
     * frame, row and tile indices cost 20 instructions per row), test the ragged-tile guards once, compute before they
     * test (predicated stores instead of four reconvergence regions) and alternate between two tap records, one in use
     * and one in flight (a single record copied at the end of every row cost 12 moves). */
    {
        const bool in0 = lane < tw, in1 = lane + 32 < tw, in2 = lane + 64 < tw, in3 = lane + 96 < tw;
        const size_t rstep = (size_t)8 * A.dst_stride[0];
        uint8_t *drow = dst0 + (size_t)(ry0 + warp) * A.dst_stride[0] + ((size_t)(x0 + lane) << (obits == 8 ? 0 : 1));
        auto vrow = [&](const S8VRow &vr, int y, uint8_t *dp) {
            const int n4 = __shfl_sync(0xffffffffu, vr.n4, 0);     /* this row's own group count, warp-uniform */
            const uint32_t *hp = hb_l + lane * lstride_w + ((vr.pos_even - lo_l) >> 1);
            int v[S8_TW / 32];
            s8_vsum<S8_TW / 32>(hp, 32 * lstride_w, vr, n4, 0, v);
            if (LONGV && A.vl2)
                s8_vsum_more<S8_TW / 32>(A.vl2 + y, hb_l + lane * lstride_w, lo_l, 32 * lstride_w, v);
            if (obits == 8) {
                /* lane + 32 c == lane (mod 8): one dither value per lane and row */
                const int dz = (bayer ? c_dither_8x8_128[y & 7][lane & 7] : 64) << 12;
                const int o0 = clip_u8((v[0] + dz) >> 19), o1 = clip_u8((v[1] + dz) >> 19);
                const int o2 = clip_u8((v[2] + dz) >> 19), o3 = clip_u8((v[3] + dz) >> 19);
                if (in0) dp[0] = (uint8_t)o0;
                if (in1) dp[32] = (uint8_t)o1;
                if (in2) dp[64] = (uint8_t)o2;
                if (in3) dp[96] = (uint8_t)o3;
            } else {                                   /* yuv2planeX_10_c_template / yuv2plane1_10 (output.c:340-357) */
                uint16_t *d = reinterpret_cast<uint16_t *>(dp);
                const int rnd = 1 << (oshift - 1);
                const int o0 = clip_uintp2((v[0] + rnd) >> oshift, obits), o1 = clip_uintp2((v[1] + rnd) >> oshift, obits);
                const int o2 = clip_uintp2((v[2] + rnd) >> oshift, obits), o3 = clip_uintp2((v[3] + rnd) >> oshift, obits);
                if (in0) d[0] = (uint16_t)(o0 << olsh);
                if (in1) d[32] = (uint16_t)(o1 << olsh);
                if (in2) d[64] = (uint16_t)(o2 << olsh);
                if (in3) d[96] = (uint16_t)(o3 << olsh);
            }
        };
        S8VRow ra, rb;
        if (warp < th)
            ra = s8_load_vrow(A.vl + ry0 + warp);
        for (int ty = warp; ty < th; ty += 16) {
            const bool second = ty + 8 < th;
            if (second)
                rb = s8_load_vrow(A.vl + ry0 + ty + 8);     /* the next row's taps are in flight while this row is filtered */
            vrow(ra, ry0 + ty, drow);
            if (second) {
                if (ty + 16 < th)
                    ra = s8_load_vrow(A.vl + ry0 + ty + 16);
                vrow(rb, ry0 + ty + 8, drow + rstep);
            }
            drow += 2 * rstep;
        }
    }
    /* ================= stage V, chroma: task = (plane, row); 8 warps, so a warp keeps its plane ================= */
    if (ch > 0) {
        const bool semi = A.dst_kind == SWSC_DST_NV12 || A.dst_kind == SWSC_DST_NV21 || A.dst_kind == SWSC_DST_P010;
        const int first = A.dst_kind == SWSC_DST_NV21 ? 1 : 0;   /* nv12, p010le: U first */
        const int pl = warp & 1;
        const uint32_t *hb_p = pl ? hb_v : hb_u;
        const int cstr = A.dst_stride[semi ? 1 : pl ? 2 : 1];
        const size_t rstep = (size_t)4 * cstr;
        uint8_t *drow = (semi ? dst1 + ((2 * (cx0 + lane) + (pl ^ first)) << (obits == 8 ? 0 : 1))
                              : (pl ? dst2 : dst1) + ((size_t)(cx0 + lane) << (obits == 8 ? 0 : 1))) +
                        (size_t)(cy0 + (warp >> 1)) * cstr;
        const int dstep = semi ? 64 : 32;
        const int dcol = (cx0 + lane + 3 * pl) & 7;      /* chroma dither: U reads column x, V column x + 3 of the row (swscale.c:519-522, output.c:468-528) */
        auto crow = [&](const S8VRow &vr, int y, uint8_t *dp) {
            const int n4 = __shfl_sync(0xffffffffu, vr.n4, 0);
            const uint32_t *hp = hb_p + lane * cstride_w + ((vr.pos_even - lo_c) >> 1);
            const int dz = (bayer ? c_dither_8x8_128[y & 7][dcol] : 64) << 12;
            for (int c = 0; 32 * c < CW; c += 2) {
                int v[2];
                s8_vsum<2>(hp + 32 * c * cstride_w, 32 * cstride_w, vr, n4, 0, v);
                if (LONGV && A.vc2)
                    s8_vsum_more<2>(A.vc2 + y, hb_p + (lane + 32 * c) * cstride_w, lo_c, 32 * cstride_w, v);
                const bool i0 = lane + 32 * c < cw, i1 = lane + 32 * c + 32 < cw;
                if (obits == 8) {
                    const int o0 = clip_u8((v[0] + dz) >> 19), o1 = clip_u8((v[1] + dz) >> 19);
                    if (i0) dp[dstep * c] = (uint8_t)o0;
                    if (i1) dp[dstep * (c + 1)] = (uint8_t)o1;
                } else {
                    uint16_t *d = reinterpret_cast<uint16_t *>(dp);
                    const int rnd = 1 << (oshift - 1);
                    const int o0 = clip_uintp2((v[0] + rnd) >> oshift, obits), o1 = clip_uintp2((v[1] + rnd) >> oshift, obits);
                    if (i0) d[dstep * c] = (uint16_t)(o0 << olsh);
                    if (i1) d[dstep * (c + 1)] = (uint16_t)(o1 << olsh);
                }
            }
        };
        /* tasks warp, warp + 8, ...: rows cy0 + (warp >> 1) + 4 k */
        S8VRow ra, rb;
        const int r0 = warp >> 1;
        if (r0 < ch)
            ra = s8_load_vrow(A.vc + cy0 + r0);
        for (int r = r0; r < ch; r += 8) {
            const bool second = r + 4 < ch;
            if (second)
                rb = s8_load_vrow(A.vc + cy0 + r + 4);
            crow(ra, cy0 + r, drow);
            if (second) {
                if (r + 8 < ch)
                    ra = s8_load_vrow(A.vc + cy0 + r + 8);
                crow(rb, cy0 + r + 4, drow + rstep);
            }
            drow += 2 * rstep;
        }
    }
}
