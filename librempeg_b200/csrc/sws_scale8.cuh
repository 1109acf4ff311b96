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
 */
#pragma once

#define S8_TW 128
#define S8_CH 16         /* source rows staged per pass */
#define S8_PRE 4         /* 16-byte loads a thread keeps in flight for the next pass */
#define S8_VF4 5         /* vertical taps: up to 5 groups of 4 (16 taps + parity pad) */

struct S8VRow {          /* per destination row, 48 bytes */
    int pos_even;        /* first source row, rounded down to even */
    int n4;              /* groups of four taps in use */
    uint32_t cl[S8_VF4]; /* low bytes of the taps, four per word */
    uint32_t ch[S8_VF4]; /* high (signed) bytes */
};

struct Scale8Args {
    const uint8_t *src[3];
    uint8_t *dst[3];
    long long src_fstride[3], dst_fstride[3];
    int src_stride[3], dst_stride[3];
    int src_w, src_h, chr_src_w, chr_src_h, dst_w, dst_h, chr_dst_w, chr_dst_h;
    int hs, vs;
    int src_layout, dst_kind;
    int y0, y1, tile_h;
    int nl_cap, nc_cap;
    int seg_l, seg_c;
    const int *hl_pos, *hc_pos;
    const uint32_t *hl_cl, *hl_ch, *hc_cl, *hc_ch;
    const S8VRow *vl, *vc;
};

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
template <int FS4>
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
    return min(((acc_h << 8) + acc_l) >> 7, (1 << 15) - 1);
}

/* vertical FIR for one column (transposed 15-bit lines), result before the >> 19 */
__device__ __forceinline__ int s8_vfir(const uint32_t *hp, const S8VRow &vr)
{
    int acc_l = 64 << 12, acc_h = 0;       /* dither 64 for 8-bit sources (swscale.c:54-56,385-387) */
#pragma unroll
    for (int k = 0; k < S8_VF4; k++) {
        if (k < vr.n4) {
            const uint32_t w0 = hp[2 * k], w1 = hp[2 * k + 1];
            acc_l = dp2a_lo_su(w0, vr.cl[k], acc_l);
            acc_h = dp2a_lo_ss(w0, vr.ch[k], acc_h);
            acc_l = dp2a_hi_su(w1, vr.cl[k], acc_l);
            acc_h = dp2a_hi_ss(w1, vr.ch[k], acc_h);
        }
    }
    return acc_h * 256 + acc_l;
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


/*
 * Kernel shape.  One CTA = 128 x TH outputs, 256 threads, three CTAs per SM.
 *  staging: 16 source rows per pass; warp w brings in rows 2w and 2w+1 (16-byte loads, one row segment
 *           per warp instruction), held in registers while the previous pass is being filtered.
 *  H:       thread = (output column, row group).  A thread filters PAIRS of vertically adjacent rows
 *           and stores both 15-bit results with one 32-bit shared-memory store into the transposed
 *           line buffer (column stride is an odd number of words: conflict-free for H stores and V loads).
 *  V:       warp = output row, lane = columns lane + 32k.
 */
template <int FS4>
__global__ void __launch_bounds__(256, 3)
sws_scale8_kernel(const __grid_constant__ Scale8Args A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int f = blockIdx.z;
    const uint8_t *src0 = A.src[0] + f * A.src_fstride[0];
    const uint8_t *src1 = A.src[1] + f * A.src_fstride[1];
    const uint8_t *src2 = A.src[2] ? A.src[2] + f * A.src_fstride[2] : nullptr;
    uint8_t *dst0 = A.dst[0] + f * A.dst_fstride[0];
    uint8_t *dst1 = A.dst[1] + f * A.dst_fstride[1];
    uint8_t *dst2 = A.dst[2] ? A.dst[2] + f * A.dst_fstride[2] : nullptr;

    const int TH = A.tile_h;
    const int x0 = blockIdx.x * S8_TW;
    const int ry0 = A.y0 + blockIdx.y * TH;
    const int ry1 = min(ry0 + TH, A.y1);
    const int tw = min(S8_TW, A.dst_w - x0), th = ry1 - ry0;
    const int cs = 7 - A.hs;                 /* log2 of the chroma tile width */
    const int CW = 1 << cs;
    const int cx0 = x0 >> A.hs;
    const int cw = min(CW, A.chr_dst_w - cx0);
    const int cy0 = ry0 >> A.vs;
    const int cy1 = (ry1 == A.dst_h) ? A.chr_dst_h : (ry1 >> A.vs);
    const int ch = cy1 - cy0;
    const int lstride_w = A.nl_cap >> 1, cstride_w = A.nc_cap >> 1;   /* odd by construction */

    uint32_t *hb_l = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *hb_u = hb_l + S8_TW * lstride_w;
    uint32_t *hb_v = hb_u + CW * cstride_w;
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb_v + CW * cstride_w);

    /* source row windows of the tile (first rows are even by construction): lane = output row */
    int lo_l = INT_MAX, hi_l = 0, lo_c = INT_MAX, hi_c = 0;
    if (ry0 + lane < ry1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vl + ry0 + lane));
        lo_l = pn.x;
        hi_l = pn.x + 4 * pn.y;
    }
    if (cy0 + lane < cy1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vc + cy0 + lane));
        lo_c = pn.x;
        hi_c = pn.x + 4 * pn.y;
    }
    lo_l = __reduce_min_sync(0xffffffffu, lo_l);
    hi_l = __reduce_max_sync(0xffffffffu, hi_l);
    lo_c = __reduce_min_sync(0xffffffffu, lo_c);
    hi_c = __reduce_max_sync(0xffffffffu, hi_c);
    const int nl = min(min(hi_l, A.src_h) - lo_l, A.nl_cap);
    const int nc = ch > 0 ? min(min(hi_c, A.chr_src_h) - lo_c, A.nc_cap) : 0;

    /* ================= stage H, luma: thread = (column, row group of 8) ================= */
    {
        const int x = tid & (S8_TW - 1), g = tid >> 7;
        const int gx = min(x0 + x, A.dst_w - 1);
        const int a0 = __ldg(A.hl_pos + x0) & ~15;
        const int off = __ldg(A.hl_pos + gx) - a0;
        const int sh = (off & 3) * 8;
        uint32_t cl[FS4], chh[FS4];
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            cl[k] = __ldg(A.hl_cl + (size_t)gx * FS4 + k);
            chh[k] = __ldg(A.hl_ch + (size_t)gx * FS4 + k);
        }
        const int seg = A.seg_l, nchunk = seg >> 4;
        const int last16 = (A.src_stride[0] - 16) & ~15;
        const int go0 = min(a0 + 16 * lane, last16), go1 = min(a0 + 16 * lane + 512, last16);
        const bool v0 = lane < nchunk, v1 = lane + 32 < nchunk;
        uint4 pre[4];
        auto fetch = [&](int r) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int sr = min(lo_l + r + 2 * warp + j, A.src_h - 1);
                const uint8_t *rp = src0 + (size_t)sr * A.src_stride[0];
                if (v0) pre[2 * j] = __ldg(reinterpret_cast<const uint4 *>(rp + go0));
                if (v1) pre[2 * j + 1] = __ldg(reinterpret_cast<const uint4 *>(rp + go1));
            }
        };
        unsigned char *sd = stage + 2 * warp * seg + 16 * lane;
        const unsigned char *sp = stage + 8 * g * seg + (off & ~3);
        uint32_t *hp = hb_l + x * lstride_w + 4 * g;
        if (nl > 0)
            fetch(0);
        for (int r = 0; r < nl; r += S8_CH) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (v0) *reinterpret_cast<uint4 *>(sd + j * seg) = pre[2 * j];
                if (v1) *reinterpret_cast<uint4 *>(sd + j * seg + 512) = pre[2 * j + 1];
            }
            __syncthreads();
            if (r + S8_CH < nl)
                fetch(r + S8_CH);
            const int left = nl - r - 8 * g;         /* rows of this group still inside the window */
#pragma unroll
            for (int m = 0; m < 4; m++) {
                if (2 * m < left) {
                    const int va = s8_hfir<FS4>(sp + (2 * m) * seg, sh, cl, chh);
                    const int vb = s8_hfir<FS4>(sp + (2 * m + 1) * seg, sh, cl, chh);
                    hp[(r >> 1) + m] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                }
            }
        }
    }
    /* ================= stage H, chroma: thread = (column, plane, row group) ================= */
    if (ch > 0) {
        const int x = tid & (CW - 1), pl = (tid >> cs) & 1, g = tid >> (cs + 1);
        const int npair = A.hs ? 4 : 8;          /* row pairs per thread and pass */
        const int gx = min(cx0 + x, A.chr_dst_w - 1);
        const int a0 = __ldg(A.hc_pos + cx0) & ~15;
        const int off = __ldg(A.hc_pos + gx) - a0;
        const int sh = (off & 3) * 8;
        uint32_t cl[FS4], chh[FS4];
#pragma unroll
        for (int k = 0; k < FS4; k++) {
            cl[k] = __ldg(A.hc_cl + (size_t)gx * FS4 + k);
            chh[k] = __ldg(A.hc_ch + (size_t)gx * FS4 + k);
        }
        const int seg = A.seg_c;
        const bool planar = A.src_layout == SWSC_SRC_PLANAR;
        const int uo = A.src_layout == SWSC_SRC_NV21 ? 1 : 0;     /* nv21: V first */
        const int last16 = (A.src_stride[1] - 16) & ~15;
        /* planar: slot q = (plane q>>1, row 2w + (q&1)), chunk = lane;  nv12: slot q = (row 2w + (q>>1), chunk lane + 32(q&1)) */
        const int nchunk = planar ? seg >> 4 : seg >> 3;
        const int go0 = planar ? min(a0 + 16 * lane, last16) : min(2 * a0 + 16 * lane, last16);
        const int go1 = planar ? go0 : min(2 * a0 + 16 * lane + 512, last16);
        const bool v0 = lane < nchunk, v1 = planar ? v0 : lane + 32 < nchunk;
        uint4 pre[4];
        auto fetch = [&](int r) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int sr = min(lo_c + r + 2 * warp + j, A.chr_src_h - 1);
                if (planar) {
                    if (v0) {
                        pre[j] = __ldg(reinterpret_cast<const uint4 *>(src1 + (size_t)sr * A.src_stride[1] + go0));
                        pre[2 + j] = __ldg(reinterpret_cast<const uint4 *>(src2 + (size_t)sr * A.src_stride[2] + go0));
                    }
                } else {
                    const uint8_t *rp = src1 + (size_t)sr * A.src_stride[1];
                    if (v0) pre[2 * j] = __ldg(reinterpret_cast<const uint4 *>(rp + go0));
                    if (v1) pre[2 * j + 1] = __ldg(reinterpret_cast<const uint4 *>(rp + go1));
                }
            }
        };
        const unsigned char *sp = stage + (pl * S8_CH + 2 * npair * g) * seg + (off & ~3);
        uint32_t *hp = (pl ? hb_v : hb_u) + x * cstride_w + npair * g;
        if (nc > 0)
            fetch(0);
        for (int r = 0; r < nc; r += S8_CH) {
            __syncthreads();
            if (planar) {
                unsigned char *sd = stage + 2 * warp * seg + 16 * lane;
                if (v0) {
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        *reinterpret_cast<uint4 *>(sd + j * seg) = pre[j];
                        *reinterpret_cast<uint4 *>(sd + (S8_CH + j) * seg) = pre[2 + j];
                    }
                }
            } else {
                /* nv12 / nv21: 16 interleaved bytes -> 8 U + 8 V (input.c:926-941) */
                unsigned char *se = stage + (uo * S8_CH + 2 * warp) * seg + 8 * lane;
                unsigned char *so = stage + ((1 - uo) * S8_CH + 2 * warp) * seg + 8 * lane;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if ((q & 1) ? v1 : v0) {
                        const uint4 v = pre[q];
                        const int o = (q >> 1) * seg + 256 * (q & 1);
                        *reinterpret_cast<uint2 *>(se + o) = make_uint2(prmt(v.x, v.y, 0x6420), prmt(v.z, v.w, 0x6420));
                        *reinterpret_cast<uint2 *>(so + o) = make_uint2(prmt(v.x, v.y, 0x7531), prmt(v.z, v.w, 0x7531));
                    }
                }
            }
            __syncthreads();
            if (r + S8_CH < nc)
                fetch(r + S8_CH);
            const int left = nc - r - 2 * npair * g;
            for (int mm = 0; mm < npair; mm += 4) {
#pragma unroll
                for (int m4 = 0; m4 < 4; m4++) {
                    const int m = mm + m4;
                    if (2 * m < left) {
                        const int va = s8_hfir<FS4>(sp + (2 * m) * seg, sh, cl, chh);
                        const int vb = s8_hfir<FS4>(sp + (2 * m + 1) * seg, sh, cl, chh);
                        hp[(r >> 1) + m] = prmt((uint32_t)va, (uint32_t)vb, 0x5410);
                    }
                }
            }
        }
    }
    __syncthreads();

    /* ================= stage V, luma: warp = row, lane = columns lane, lane+32, ... ================= */
    for (int ty = warp; ty < th; ty += 8) {
        const int y = ry0 + ty;
        const S8VRow vr = s8_load_vrow(A.vl + y);
        const uint32_t *hp = hb_l + lane * lstride_w + ((vr.pos_even - lo_l) >> 1);
        uint8_t *d = dst0 + (size_t)y * A.dst_stride[0] + x0 + lane;
#pragma unroll
        for (int c = 0; c < S8_TW / 32; c++) {
            const int v = clip_u8(s8_vfir(hp + 32 * c * lstride_w, vr) >> 19);
            if (lane + 32 * c < tw)
                d[32 * c] = (uint8_t)v;
        }
    }
    /* ================= stage V, chroma: task = (plane, row) ================= */
    if (ch > 0) {
        const bool semi = A.dst_kind == SWSC_DST_NV12 || A.dst_kind == SWSC_DST_NV21;
        const int first = A.dst_kind == SWSC_DST_NV21 ? 1 : 0;   /* nv12: U first */
        for (int task = warp; task < 2 * ch; task += 8) {
            const int pl = task & 1, y = cy0 + (task >> 1);
            const S8VRow vr = s8_load_vrow(A.vc + y);
            const uint32_t *hp = (pl ? hb_v : hb_u) + lane * cstride_w + ((vr.pos_even - lo_c) >> 1);
            uint8_t *d = semi ? dst1 + (size_t)y * A.dst_stride[1] + 2 * (cx0 + lane) + (pl ^ first)
                              : (pl ? dst2 : dst1) + (size_t)y * A.dst_stride[pl ? 2 : 1] + cx0 + lane;
            const int dstep = semi ? 64 : 32;
            for (int c = 0; 32 * c < CW; c++) {
                const int v = clip_u8(s8_vfir(hp + 32 * c * cstride_w, vr) >> 19);
                if (lane + 32 * c < cw)
                    d[dstep * c] = (uint8_t)v;
            }
        }
    }
}
