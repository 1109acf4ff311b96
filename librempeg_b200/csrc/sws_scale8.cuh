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
    return min((acc_h * 256 + acc_l) >> 7, (1 << 15) - 1);
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
    const int CW = S8_TW >> A.hs;
    const int cx0 = x0 >> A.hs;
    const int cw = min(CW, A.chr_dst_w - cx0);
    const int cy0 = ry0 >> A.vs;
    const int cy1 = (ry1 == A.dst_h) ? A.chr_dst_h : (ry1 >> A.vs);
    const int ch = cy1 - cy0;

    int16_t *hb_l = reinterpret_cast<int16_t *>(smem_raw);
    int16_t *hb_u = hb_l + (size_t)S8_TW * A.nl_cap;
    int16_t *hb_v = hb_u + (size_t)CW * A.nc_cap;
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb_v + (size_t)CW * A.nc_cap);

    /* source row windows of the tile (first rows are even by construction) */
    int lo_l = INT_MAX, hi_l = 0, lo_c = INT_MAX, hi_c = 0;
    for (int y = ry0; y < ry1; y++) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vl + y));
        lo_l = min(lo_l, pn.x);
        hi_l = max(hi_l, pn.x + 4 * pn.y);
    }
    for (int y = cy0; y < cy1; y++) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vc + y));
        lo_c = min(lo_c, pn.x);
        hi_c = max(hi_c, pn.x + 4 * pn.y);
    }
    const int nl = min(min(hi_l, A.src_h) - lo_l, A.nl_cap);
    const int nc = ch > 0 ? min(min(hi_c, A.chr_src_h) - lo_c, A.nc_cap) : 0;

    /* ================= stage H, luma: thread = (column, row parity) ================= */
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
        const int nchunk = A.seg_l >> 4;
        const int last16 = (A.src_stride[0] - 16) & ~15;
        const int per_pass = S8_CH * nchunk;
        uint4 pre[S8_PRE];
        /* software pipeline: the loads of pass k+1 are in flight while pass k is filtered */
        auto fetch = [&](int r) {
#pragma unroll
            for (int q = 0; q < S8_PRE; q++) {
                const int i = tid + q * 256;
                if (i < per_pass) {
                    const int row = i / nchunk, c = i - row * nchunk;
                    const int sr = min(lo_l + r + row, A.src_h - 1);
                    pre[q] = __ldg(reinterpret_cast<const uint4 *>(src0 + (size_t)sr * A.src_stride[0] +
                                                                    min(a0 + 16 * c, last16)));
                }
            }
        };
        if (nl > 0)
            fetch(0);
        for (int r = 0; r < nl; r += S8_CH) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < S8_PRE; q++) {
                const int i = tid + q * 256;
                if (i < per_pass) {
                    const int row = i / nchunk, c = i - row * nchunk;
                    *reinterpret_cast<uint4 *>(stage + row * A.seg_l + 16 * c) = pre[q];
                }
            }
            __syncthreads();
            if (r + S8_CH < nl)
                fetch(r + S8_CH);
            for (int rr = g; rr < S8_CH && r + rr < nl; rr += 2) {
                const int val = s8_hfir<FS4>(stage + rr * A.seg_l + (off & ~3), sh, cl, chh);
                if (x < tw)
                    hb_l[(size_t)x * A.nl_cap + r + rr] = (int16_t)val;
            }
        }
    }
    /* ================= stage H, chroma: thread = (column, plane, row parity) ================= */
    if (ch > 0) {
        const int x = tid & (CW - 1), pl = (tid / CW) & 1, g = tid / (2 * CW);
        const int ngroups = 256 / (2 * CW);
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
        int16_t *hb = pl ? hb_v : hb_u;
        const bool planar = A.src_layout == SWSC_SRC_PLANAR;
        const int uo = A.src_layout == SWSC_SRC_NV21 ? 1 : 0;     /* nv21: V first */
        const int nchunk = planar ? A.seg_c >> 4 : A.seg_c >> 3;
        const int per_pass = (planar ? 2 : 1) * S8_CH * nchunk;
        const int last16 = (A.src_stride[1] - 16) & ~15;
        uint4 pre[S8_PRE];
        auto fetch = [&](int r) {
#pragma unroll
            for (int q = 0; q < S8_PRE; q++) {
                const int i = tid + q * 256;
                if (i < per_pass) {
                    if (planar) {
                        const int p = i / (S8_CH * nchunk), j = i - p * (S8_CH * nchunk);
                        const int row = j / nchunk, c = j - row * nchunk;
                        const int sr = min(lo_c + r + row, A.chr_src_h - 1);
                        const uint8_t *base = p ? src2 + (size_t)sr * A.src_stride[2]
                                                : src1 + (size_t)sr * A.src_stride[1];
                        pre[q] = __ldg(reinterpret_cast<const uint4 *>(base + min(a0 + 16 * c, last16)));
                    } else {
                        const int row = i / nchunk, c = i - row * nchunk;
                        const int sr = min(lo_c + r + row, A.chr_src_h - 1);
                        pre[q] = __ldg(reinterpret_cast<const uint4 *>(src1 + (size_t)sr * A.src_stride[1] +
                                                                        min(2 * a0 + 16 * c, last16)));
                    }
                }
            }
        };
        if (nc > 0)
            fetch(0);
        for (int r = 0; r < nc; r += S8_CH) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < S8_PRE; q++) {
                const int i = tid + q * 256;
                if (i < per_pass) {
                    const uint4 v = pre[q];
                    if (planar) {
                        const int p = i / (S8_CH * nchunk), j = i - p * (S8_CH * nchunk);
                        const int row = j / nchunk, c = j - row * nchunk;
                        *reinterpret_cast<uint4 *>(stage + (p * S8_CH + row) * A.seg_c + 16 * c) = v;
                    } else {
                        /* nv12 / nv21: 16 interleaved bytes -> 8 U + 8 V (input.c:926-941) */
                        const int row = i / nchunk, c = i - row * nchunk;
                        const uint2 e = make_uint2(prmt(v.x, v.y, 0x6420), prmt(v.z, v.w, 0x6420));
                        const uint2 o = make_uint2(prmt(v.x, v.y, 0x7531), prmt(v.z, v.w, 0x7531));
                        *reinterpret_cast<uint2 *>(stage + (uo * S8_CH + row) * A.seg_c + 8 * c) = e;
                        *reinterpret_cast<uint2 *>(stage + ((1 - uo) * S8_CH + row) * A.seg_c + 8 * c) = o;
                    }
                }
            }
            __syncthreads();
            if (r + S8_CH < nc)
                fetch(r + S8_CH);
            for (int rr = g; rr < S8_CH && r + rr < nc; rr += ngroups) {
                const int val = s8_hfir<FS4>(stage + (pl * S8_CH + rr) * A.seg_c + (off & ~3), sh, cl, chh);
                if (x < cw)
                    hb[(size_t)x * A.nc_cap + r + rr] = (int16_t)val;
            }
        }
    }
    __syncthreads();

    /* ================= stage V, luma: warp = row, lane = columns lane, lane+32, ... ================= */
    for (int ty = warp; ty < th; ty += 8) {
        const int y = ry0 + ty;
        const S8VRow vr = s8_load_vrow(A.vl + y);
        const int p2 = (vr.pos_even - lo_l) >> 1;
        uint8_t *d = dst0 + (size_t)y * A.dst_stride[0] + x0;
#pragma unroll
        for (int c = 0; c < S8_TW / 32; c++) {
            const int col = lane + 32 * c;
            const uint32_t *hp = reinterpret_cast<const uint32_t *>(hb_l + (size_t)col * A.nl_cap) + p2;
            const int v = clip_u8(s8_vfir(hp, vr) >> 19);
            if (col < tw)
                d[col] = (uint8_t)v;
        }
    }
    /* ================= stage V, chroma: task = (plane, row) ================= */
    if (ch > 0) {
        const bool semi = A.dst_kind == SWSC_DST_NV12 || A.dst_kind == SWSC_DST_NV21;
        for (int task = warp; task < 2 * ch; task += 8) {
            const int pl = task & 1, y = cy0 + (task >> 1);
            const S8VRow vr = s8_load_vrow(A.vc + y);
            const int p2 = (vr.pos_even - lo_c) >> 1;
            const int16_t *hb = pl ? hb_v : hb_u;
            for (int col = lane; col < CW; col += 32) {
                const uint32_t *hp = reinterpret_cast<const uint32_t *>(hb + (size_t)col * A.nc_cap) + p2;
                const int v = clip_u8(s8_vfir(hp, vr) >> 19);
                if (col < cw) {
                    if (!semi) {
                        uint8_t *d = (pl ? dst2 : dst1) + (size_t)y * A.dst_stride[pl ? 2 : 1];
                        d[cx0 + col] = (uint8_t)v;
                    } else {
                        const int first = A.dst_kind == SWSC_DST_NV12 ? 0 : 1;   /* nv12: U first */
                        dst1[(size_t)y * A.dst_stride[1] + 2 * (cx0 + col) + (pl ^ first)] = (uint8_t)v;
                    }
                }
            }
        }
    }
}
