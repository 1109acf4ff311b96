/*
 * sws_scale8t.cuh -- 8-bit YUV -> 8-bit planar / semi-planar YUV scaling with BOTH FIR stages on the tensor
 * pipe (BASELINE config C4, every yuv -> yuv resize of 8-bit material).  Same fused pipeline and the same
 * arithmetic as sws_scale8.cuh
 *   nv12ToUV_c -> hScale8To15_c -> yuv2planeX_8_c / yuv2plane1_8_c / yuv2nv12cX_c
 *   (libswscale/input.c:926-941, swscale.c:128-142, output.c:468-528)
 * but the vertical stage no longer spends one IDP.2A per tap and output:
 *
 *  H  as in sws_scale8.cuh<MMA>: 16 staged rows x 8 output columns = KS x 2 integer MMAs over a banded
 *     coefficient block.  The 15-bit results (clip at 32767 included) are stored as TWO BYTE PLANES,
 *     lo[col][row] and hi[col][row] (hi signed), four vertically adjacent rows per 32-bit word: lanes g and
 *     g ^ 1 swap one packed pair with a butterfly shuffle so that each owns four rows of one column.
 *  V  out[y][x] = (sum_r coef[y][r] * line[r][x] + dither) >> 19 for 16 output rows x 8 columns is
 *     A (coefficient bytes of the 16 rows over a K = 32 KV source-row window, host-prepared fragments)
 *     x B (the byte planes, one LDS.32 per fragment register), four sign combinations:
 *        sum c*l = LL + 256 (LH + HL) + 65536 HH,   c = 256 ch + cl,  l = 256 lh + ll
 *     -- exact in 32-bit two's complement, the same number the C code's int accumulator holds.
 *     Finished 16 x 32 blocks are staged in the (idle) ring and leave with 16-byte stores.
 *
 * Measured on B200 (profiles/): the legacy integer MMA sustains 2044 MAC/clk/SM, 8x the dot-product pipe;
 * at 12.5 % (H) and 4-17 % (V) band density that is still the cheaper pipe, and it frees the ALU pipes
 * for the epilogues.
 */
#pragma once

struct Scale8TArgs {
    Scale8Args b;                  /* geometry, ring layout, horizontal fragments (sws_scale8.cuh) */
    const uint32_t *vl_A, *vc_A;   /* vertical A fragments [block of 16 rows][K step][cl0..3 ch0..3][lane] */
    const int *vl_ws, *vc_ws;      /* first source row of every block's K window (multiple of 4) */
    int vl_ks, vc_ks;              /* K steps of 32 source rows per block */
    int lS, cS;                    /* words per column of a byte plane (== 4 mod 8: conflict-free stores and loads) */
    int vec_ok;                    /* destination planes and strides allow 16-byte row stores */
};

__device__ __forceinline__ void s8t_mma(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int kind)
{
    /* kind: 0 u8*u8, 1 u8*s8, 2 s8*u8, 3 s8*s8 (A type, B type); resolved at compile time by the callers */
    if (kind == 0)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else if (kind == 1)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else if (kind == 2)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

/* accumulators of one 16 x 8 block: LL, LH + HL, HH */
struct S8TAcc {
    int ll[4], mid[4], hh[4];
};

__device__ __forceinline__ void s8t_zero(S8TAcc &a)
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        a.ll[i] = a.mid[i] = a.hh[i] = 0;
}

/* one K step of one block: B fragments straight from the byte planes */
__device__ __forceinline__ void s8t_vstep(S8TAcc &acc, const uint32_t (&acl)[4], const uint32_t (&ach)[4],
                                          const uint32_t *lo, const uint32_t *hi)
{
    const uint32_t bl0 = lo[0], bl1 = lo[4], bh0 = hi[0], bh1 = hi[4];
    s8t_mma(acc.ll, acl, bl0, bl1, 0);
    s8t_mma(acc.mid, acl, bh0, bh1, 1);
    s8t_mma(acc.mid, ach, bl0, bl1, 2);
    s8t_mma(acc.hh, ach, bh0, bh1, 3);
}

/* (sum + dither 64 << 12) >> 19, clipped: yuv2planeX_8_c / yuv2plane1_8_c / yuv2nv12cX_c for 8-bit sources */
__device__ __forceinline__ uint32_t s8t_out(const S8TAcc &a, int i)
{
    const uint32_t v = (uint32_t)a.ll[i] + ((uint32_t)a.mid[i] << 8) + ((uint32_t)a.hh[i] << 16) + (64u << 12);
    return (uint32_t)clip_u8((int)v >> 19);
}

/* 15-bit sample pairs of two MMA accumulator sets -> byte-plane words (four rows of one column) */
__device__ __forceinline__ void s8t_store4(uint32_t *lo, uint32_t *hi, uint32_t wa, uint32_t wb, int lane, bool ok)
{
    /* wa = rows (2g, 2g+1) of column 2t, wb = the same rows of column 2t+1.  Even g keeps column 2t and needs
     * rows (2g+2, 2g+3) from lane g+1; odd g keeps column 2t+1 and needs rows (2g-2, 2g-1) from lane g-1. */
    const bool odd = lane & 4;
    const uint32_t give = odd ? wa : wb, keep = odd ? wb : wa;
    const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 4);
    const uint32_t first = odd ? got : keep, second = odd ? keep : got;     /* rows 4i, 4i+1 | rows 4i+2, 4i+3 */
    if (ok) {
        *lo = prmt(first, second, 0x6420);
        *hi = prmt(first, second, 0x7531);
    }
}

template <int KS>
__global__ void __launch_bounds__(S8_THREADS, 3)
sws_scale8t_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_u,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ Scale8TArgs T)
{
    extern __shared__ __align__(128) unsigned char s8_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[S8_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[S8_MAX_STAGES];
    const Scale8Args &A = T.b;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int f = blockIdx.z;

    const int TH = A.tile_h;                   /* multiple of 16 */
    const int x0 = blockIdx.x * S8_TW;
    const int ry0 = A.y0 + blockIdx.y * TH;    /* multiple of 16 */
    const int ry1 = min(ry0 + TH, A.y1);
    const int cs = 7 - A.hs;
    const int CW = 1 << cs;
    const int cx0 = x0 >> A.hs;
    const int cy0 = ry0 >> A.vs;
    const int cy1 = (ry1 == A.dst_h) ? A.chr_dst_h : (ry1 >> A.vs);
    const int ch = cy1 - cy0;
    const int slot = A.slot_bytes;

    if (tid == 0) {
        for (int s = 0; s < A.stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    /* source row windows of the tile: first row = the first block's window start (multiple of 4) */
    const int nbl = (ry1 - ry0 + 15) >> 4, nbc = ch > 0 ? (ch + 15) >> 4 : 0;
    const int bl0 = ry0 >> 4, bc0 = cy0 >> 4;
    int lo_l = INT_MAX, hi_l = 0, lo_c = INT_MAX, hi_c = 0;
    if (lane < nbl)
        lo_l = __ldg(T.vl_ws + bl0 + lane);
    if (lane < nbc)
        lo_c = __ldg(T.vc_ws + bc0 + lane);
    if (ry0 + lane < ry1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vl + ry0 + lane));
        hi_l = (pn.x & ~1) + 4 * pn.y;
    }
    if (cy0 + lane < cy1) {
        const int2 pn = __ldg(reinterpret_cast<const int2 *>(A.vc + cy0 + lane));
        hi_c = pn.x + 4 * pn.y;
    }
    lo_l = __reduce_min_sync(0xffffffffu, lo_l);
    hi_l = __reduce_max_sync(0xffffffffu, hi_l);
    lo_c = __reduce_min_sync(0xffffffffu, lo_c);
    hi_c = __reduce_max_sync(0xffffffffu, hi_c);
    const int nl = min(min(hi_l, A.src_h) - lo_l, A.nl_cap);
    const int nc = ch > 0 ? min(min(hi_c, A.chr_src_h) - lo_c, A.nc_cap) : 0;
    const int npl = (nl + S8_ROWS - 1) / S8_ROWS, npc = (nc + S8_ROWS - 1) / S8_ROWS;

    const int a0l = __ldg(A.hl_pos + x0) & ~15;
    const int a0c = ch > 0 ? __ldg(A.hc_pos + cx0) & ~15 : 0;
    const bool planar = A.src_layout == SWSC_SRC_PLANAR;
    const uint32_t ring_a = smem_u32(s8_smem_raw), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
    __syncthreads();

    if (warp == 8) {
        /* ===== producer: one thread keeps the ring full ===== */
        if (lane == 0) {
            int b = 0;
            uint32_t par = 1;
            for (int q = 0; q < npl + npc; q++) {
                if (q >= A.stages)
                    s8_wait(empty_a + 8 * b, par);
                const uint32_t d = ring_a + b * slot, bar = full_a + 8 * b;
                if (q < npl) {
                    s8_expect_tx(bar, S8_ROWS * A.seg_l);
                    s8_tma_load(d, &map_y, bar, a0l >> 2, lo_l + S8_ROWS * q, f);
                } else {
                    const int row = lo_c + S8_ROWS * (q - npl);
                    s8_expect_tx(bar, 2 * S8_ROWS * A.seg_c);
                    if (planar) {
                        s8_tma_load(d, &map_u, bar, a0c >> 2, row, f);
                        s8_tma_load(d + S8_ROWS * A.seg_c, &map_v, bar, a0c >> 2, row, f);
                    } else {
                        s8_tma_load(d, &map_u, bar, a0c >> 1, row, f);
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

    const int tw = min(S8_TW, A.dst_w - x0);
    const int cw = min(CW, A.chr_dst_w - cx0);
    const int lS = T.lS, cS = T.cS;
    /* byte planes behind the ring: luma lo, luma hi, U lo, U hi, V lo, V hi (+ tail padding, see host) */
    uint32_t *pl_lo = reinterpret_cast<uint32_t *>(s8_smem_raw + A.stages * slot);
    uint32_t *pl_hi = pl_lo + S8_TW * lS;
    uint32_t *pu_lo = pl_hi + S8_TW * lS;
    uint32_t *pu_hi = pu_lo + CW * cS;
    uint32_t *pv_lo = pu_hi + CW * cS;
    uint32_t *pv_hi = pv_lo + CW * cS;

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

    const int t = lane & 3, g = lane >> 2;
    const uint32_t sel = S8_PAIR_SEL(lane);
    /* ================= stage H, luma: warp = 16 output columns, all 16 rows of a slot per pass ================= */
    {
        const int grp = (x0 >> 3) + 2 * warp;
        S8Bfrag<KS> b0, b1;
        s8_load_bfrag<KS>(A.hl_B, grp, lane, b0);
        s8_load_bfrag<KS>(A.hl_B, grp + 1, lane, b1);
        const uint32_t lrow = s8_ldm_row(lane) * A.seg_l + (lane >> 4) * 16;
        const uint32_t o0 = lrow + __ldg(A.hl_goff + grp), o1 = lrow + __ldg(A.hl_goff + grp + 1);
        /* after the pair swap this lane owns column 2t + (g & 1) of each group, rows 4 (g >> 1) .. + 3 of the slot */
        const int col = 16 * warp + 2 * t + (g & 1);
        const int wofs = col * lS + (g >> 1);
        int left = nl - 4 * (g >> 1);
        for (int q = 0; q < npl; q++) {
            s8_wait(full_a + 8 * sb, sphase);
            const uint32_t base = ring_a + sb * slot;
            uint32_t wa, wb, wc, wd;
            s8_mma_rows<KS>(base + o0, b0, sel, wa, wb);
            s8_mma_rows<KS>(base + o1, b1, sel, wc, wd);
            const bool ok = left > 0;
            s8t_store4(pl_lo + wofs + 4 * q, pl_hi + wofs + 4 * q, wa, wb, lane, ok);
            s8t_store4(pl_lo + wofs + 8 * lS + 4 * q, pl_hi + wofs + 8 * lS + 4 * q, wc, wd, lane, ok);
            left -= S8_ROWS;
            release();
        }
    }
    /* ================= stage H, chroma ================= */
    if (npc > 0) {
        const int ng = A.hs ? 1 : 2;
        const int grp = (cx0 >> 3) + ng * warp;
        const bool vfirst = A.src_layout == SWSC_SRC_NV21;
        const int rowbytes = planar ? A.seg_c : 2 * A.seg_c;
        const uint32_t lrow = s8_ldm_row(lane) * rowbytes + (lane >> 4) * 16;
        S8Bfrag<KS> b0, b1;
        s8_load_bfrag<KS>(A.hc_B, grp, lane, b0);
        uint32_t o0 = lrow + __ldg(A.hc_goff + grp), o1 = 0;
        if (ng == 2) {
            s8_load_bfrag<KS>(A.hc_B, grp + 1, lane, b1);
            o1 = lrow + __ldg(A.hc_goff + grp + 1);
        }
        const int col = 8 * ng * warp + 2 * t + (g & 1);
        const int wofs = col * cS + (g >> 1);
        int left = nc - 4 * (g >> 1);
        for (int qc = 0; qc < npc; qc++) {
            s8_wait(full_a + 8 * sb, sphase);
            const uint32_t base = ring_a + sb * slot;
            const bool ok = left > 0;
            uint32_t ua, ub, va, vb;
            if (planar) {
                s8_mma_rows<KS>(base + o0, b0, sel, ua, ub);
                s8_mma_rows<KS>(base + S8_ROWS * A.seg_c + o0, b0, sel, va, vb);
            } else if (vfirst) {
                s8_mma_rows_uv<KS>(base + o0, b0, sel, va, vb, ua, ub);
            } else {
                s8_mma_rows_uv<KS>(base + o0, b0, sel, ua, ub, va, vb);
            }
            s8t_store4(pu_lo + wofs + 4 * qc, pu_hi + wofs + 4 * qc, ua, ub, lane, ok);
            s8t_store4(pv_lo + wofs + 4 * qc, pv_hi + wofs + 4 * qc, va, vb, lane, ok);
            if (ng == 2) {
                if (planar) {
                    s8_mma_rows<KS>(base + o1, b1, sel, ua, ub);
                    s8_mma_rows<KS>(base + S8_ROWS * A.seg_c + o1, b1, sel, va, vb);
                } else if (vfirst) {
                    s8_mma_rows_uv<KS>(base + o1, b1, sel, va, vb, ua, ub);
                } else {
                    s8_mma_rows_uv<KS>(base + o1, b1, sel, ua, ub, va, vb);
                }
                s8t_store4(pu_lo + wofs + 8 * cS + 4 * qc, pu_hi + wofs + 8 * cS + 4 * qc, ua, ub, lane, ok);
                s8t_store4(pv_lo + wofs + 8 * cS + 4 * qc, pv_hi + wofs + 8 * cS + 4 * qc, va, vb, lane, ok);
            }
            left -= S8_ROWS;
            release();
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");      /* the 8 filtering warps: all h-scaled lines are in place */

    uint8_t *dst0 = A.dst[0] + f * A.dst_fstride[0];
    uint8_t *dst1 = A.dst[1] + f * A.dst_fstride[1];
    uint8_t *dst2 = A.dst[2] ? A.dst[2] + f * A.dst_fstride[2] : nullptr;
    unsigned char *stage = s8_smem_raw + warp * 512;     /* the ring is idle now: 16 rows x 32 bytes per warp */

    /* ================= stage V, luma: unit = (block of 16 rows, 32 columns) ================= */
    for (int unit = warp; unit < 4 * nbl; unit += 8) {
        const int yb = unit >> 2, gq = unit & 3;
        const int blk = bl0 + yb;
        const int wbase = (__ldg(T.vl_ws + blk) - lo_l) >> 2;
        S8TAcc acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            s8t_zero(acc[i]);
        const uint32_t *afrag = T.vl_A + (size_t)blk * T.vl_ks * 256 + lane;
        const uint32_t *plo = pl_lo + (32 * gq + g) * lS + wbase + t;
        const uint32_t *phi = pl_hi + (32 * gq + g) * lS + wbase + t;
        for (int ks = 0; ks < T.vl_ks; ks++) {
            uint32_t acl[4], ach[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                acl[j] = __ldg(afrag + (ks * 8 + j) * 32);
                ach[j] = __ldg(afrag + (ks * 8 + 4 + j) * 32);
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                s8t_vstep(acc[i], acl, ach, plo + 8 * i * lS + 8 * ks, phi + 8 * i * lS + 8 * ks);
        }
        /* rows g and g + 8 of the block, columns 8 i + 2t, + 1 of the 32-column unit */
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            *reinterpret_cast<uint16_t *>(stage + g * 32 + 8 * i + 2 * t) =
                (uint16_t)(s8t_out(acc[i], 0) | (s8t_out(acc[i], 1) << 8));
            *reinterpret_cast<uint16_t *>(stage + (g + 8) * 32 + 8 * i + 2 * t) =
                (uint16_t)(s8t_out(acc[i], 2) | (s8t_out(acc[i], 3) << 8));
        }
        __syncwarp();
        {
            const int row = lane >> 1, half = lane & 1;
            const int y = ry0 + 16 * yb + row, xc = 32 * gq + 16 * half;
            if (y < ry1 && xc < tw) {
                uint8_t *d = dst0 + (size_t)y * A.dst_stride[0] + x0 + xc;
                const unsigned char *s = stage + row * 32 + 16 * half;
                if (T.vec_ok && xc + 16 <= tw) {
                    *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
                } else {
                    for (int i = 0; i < 16 && xc + i < tw; i++)
                        d[i] = s[i];
                }
            }
        }
    }
    /* ================= stage V, chroma: unit = (block of 16 rows, 16 columns), both planes ================= */
    if (ch > 0) {
        const bool semi = A.dst_kind == SWSC_DST_NV12 || A.dst_kind == SWSC_DST_NV21;
        const int vfirst_out = A.dst_kind == SWSC_DST_NV21 ? 1 : 0;
        const int upb = CW >> 4;                       /* units per block */
        for (int unit = warp; unit < upb * nbc; unit += 8) {
            const int yb = unit / upb, gp = unit - yb * upb;
            const int blk = bc0 + yb;
            const int wbase = (__ldg(T.vc_ws + blk) - lo_c) >> 2;
            S8TAcc acc[4];                             /* [U group 0, U group 1, V group 0, V group 1] */
#pragma unroll
            for (int i = 0; i < 4; i++)
                s8t_zero(acc[i]);
            const uint32_t *afrag = T.vc_A + (size_t)blk * T.vc_ks * 256 + lane;
            const int cofs = (16 * gp + g) * cS + wbase + t;
            for (int ks = 0; ks < T.vc_ks; ks++) {
                uint32_t acl[4], ach[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    acl[j] = __ldg(afrag + (ks * 8 + j) * 32);
                    ach[j] = __ldg(afrag + (ks * 8 + 4 + j) * 32);
                }
                s8t_vstep(acc[0], acl, ach, pu_lo + cofs + 8 * ks, pu_hi + cofs + 8 * ks);
                s8t_vstep(acc[1], acl, ach, pu_lo + cofs + 8 * cS + 8 * ks, pu_hi + cofs + 8 * cS + 8 * ks);
                s8t_vstep(acc[2], acl, ach, pv_lo + cofs + 8 * ks, pv_hi + cofs + 8 * ks);
                s8t_vstep(acc[3], acl, ach, pv_lo + cofs + 8 * cS + 8 * ks, pv_hi + cofs + 8 * cS + 8 * ks);
            }
            __syncwarp();
            if (semi) {
                /* 16 rows x 32 interleaved bytes: sample x of the unit at byte 2x + (plane ^ vfirst_out) */
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int pl = (i >> 1) ^ vfirst_out, xs = 8 * (i & 1) + 2 * t;
                    stage[g * 32 + 2 * xs + pl] = (unsigned char)s8t_out(acc[i], 0);
                    stage[g * 32 + 2 * xs + 2 + pl] = (unsigned char)s8t_out(acc[i], 1);
                    stage[(g + 8) * 32 + 2 * xs + pl] = (unsigned char)s8t_out(acc[i], 2);
                    stage[(g + 8) * 32 + 2 * xs + 2 + pl] = (unsigned char)s8t_out(acc[i], 3);
                }
                __syncwarp();
                const int row = lane >> 1, half = lane & 1;
                const int y = cy0 + 16 * yb + row, xc = 16 * gp + 8 * half;      /* chroma samples */
                if (y < cy1 && xc < cw) {
                    uint8_t *d = dst1 + (size_t)y * A.dst_stride[1] + 2 * (cx0 + xc);
                    const unsigned char *s = stage + row * 32 + 16 * half;
                    if (T.vec_ok && xc + 8 <= cw) {
                        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
                    } else {
                        for (int i = 0; i < 16 && xc + (i >> 1) < cw; i++)
                            d[i] = s[i];
                    }
                }
            } else {
                /* two planes of 16 rows x 16 bytes: U at stage, V at stage + 256 */
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    unsigned char *sp = stage + 256 * (i >> 1) + 8 * (i & 1) + 2 * t;
                    *reinterpret_cast<uint16_t *>(sp + g * 16) = (uint16_t)(s8t_out(acc[i], 0) | (s8t_out(acc[i], 1) << 8));
                    *reinterpret_cast<uint16_t *>(sp + (g + 8) * 16) = (uint16_t)(s8t_out(acc[i], 2) | (s8t_out(acc[i], 3) << 8));
                }
                __syncwarp();
                const int pl = lane >> 4, row = lane & 15;
                const int y = cy0 + 16 * yb + row, xc = 16 * gp;
                if (y < cy1 && xc < cw) {
                    uint8_t *d = (pl ? dst2 : dst1) + (size_t)y * A.dst_stride[pl ? 2 : 1] + cx0 + xc;
                    const unsigned char *s = stage + 256 * pl + row * 16;
                    if (T.vec_ok && xc + 16 <= cw) {
                        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
                    } else {
                        for (int i = 0; i < 16 && xc + i < cw; i++)
                            d[i] = s[i];
                    }
                }
            }
            __syncwarp();
        }
    }
}
