/*
 * sws_fast420.cuh -- the headline kernel: planar 8-bit 4:2:0 (any 2:1 horizontal chroma whose vertical window fits) -> packed 8-bit RGB
 * when the luma path is the identity and chroma is only filtered vertically
 * (BASELINE configs C1, C2, C5 and the default-flags LUT converter).
 *
 * Fuses, per 256x32 output tile, what the reference runs per line as
 *   hScale8To15_c (identity: x<<7)            libswscale/swscale.c:128-142
 *   yuv2rgb24_X_c / _1_c (vertical chroma FIR) libswscale/output.c:1788-1939
 *   yuv2rgb_write + LUTs (closed form)         libswscale/output.c:1697-1713, yuv2rgb.c:680-914
 *
 * Data movement: TMA (cp.async.bulk.tensor) loads the Y/U/V tiles into shared
 * memory behind an mbarrier, double-buffered across the tiles of a persistent
 * CTA; results are staged in shared memory and leave through one TMA tensor
 * store per tile, so every HBM transaction is a full coalesced line.
 *
 * Arithmetic (exact): with identity H the 15-bit line is u<<7, so
 *   U = (2^18 + sum_j (u_j<<7) c_j) >> 19 = (2048 + sum_j u_j c_j) >> 12.
 * Split c_j = 256*ch_j + cl_j (cl_j in 0..255): S = 256*T + L, and
 *   U = (T + ((L + 2048) >> 8)) >> 4                      -- two IDP.4A + two shifts.
 * The four vertical taps of one chroma column live in one register (a sliding
 * byte window advanced with one PRMT per column per source row).
 */
#pragma once

#include <cuda.h>

#define F420_TW 256          /* tile width  (luma pixels)  */
#define F420_TH 32           /* tile height (luma rows)    */
#define F420_CROWS 20        /* chroma source rows staged per tile */
#define F420_CROWS_NARROW 36 /* the same for the 128 x 64 tile shape */
#define F420_CWARPS 8        /* consumer warps, 4 rows each */
#define F420_THREADS (32 * (F420_CWARPS + 1))   /* + one producer warp */
#define F420_Y_BYTES (F420_TW * F420_TH)                                     /*  8192 */
#define F420_C_BYTES ((F420_TW / 2) * F420_CROWS)                            /*  2560 */
#define F420_META_BYTES (F420_TH * 16)                                       /*   512 */
#define F420_IN_BYTES (F420_Y_BYTES + 2 * F420_C_BYTES + F420_META_BYTES)    /* 13824 */
#define F420_OUT_BYTES(bpp) (F420_TW * (bpp) * F420_TH)                      /* 24576 / 32768 */
#ifndef F420_STAGES
#define F420_STAGES 3          /* input ring depth (measured: 3 stages x 3 CTAs/SM beats 2 x 4 by 1.3 %) */
#endif
#ifndef F420_CTAS_PER_SM
#define F420_CTAS_PER_SM 3
#endif
#define F420_SMEM(bpp) (F420_STAGES * F420_IN_BYTES + F420_OUT_BYTES(bpp))
/* sources without vertical chroma subsampling (4:2:2: yuv422p / yuvj422p, the MJPEG decode format) stage TH + 4
 * chroma rows per tile instead of 20 (36 for the narrow shape): the ring slot grows, the kernel is the same */
#define F420_IN_BYTES_CROWS(tw, th, crows) ((tw) * (th) + 2 * ((tw) / 2) * (crows) + (th) * 16)
#define F420_SMEM_CROWS(bpp, tw, th, crows) (F420_STAGES * F420_IN_BYTES_CROWS(tw, th, crows) + F420_OUT_BYTES(bpp))

/* destination byte orders and source chroma layouts the kernel is instantiated for */
enum { F420_RGB24 = 0, F420_BGR24, F420_RGBA, F420_BGRA, F420_ARGB, F420_ABGR };
enum { F420_PLANAR = 0, F420_NV12, F420_NV21 };

struct Fast420Args {
    int tiles_x, tiles_y, frames;
    int ty_first;                  /* first tile row of this launch (row-range launches) */
    int dst_h;
    int cy, yb;                    /* LUT closed form (sws_colorspace.c) */
    int crv, cbu, cgu, cgv;
    int kr, kg, kb;                /* index bases << 16 */
    /* per output row: {chroma row relative to the tile's first chroma row, cl pack, ch pack,
     * absolute chroma row}; padded to a multiple of F420_TH rows */
    const int4 *rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
#ifdef F420_WAIT_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z) : "memory");
}

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ int clamp_u8(int v)
{
    int d;
    asm("min.relu.s32 %0, %1, 255;" : "=r"(d) : "r"(v));
    return d;
}

__device__ __forceinline__ uint32_t clamp_u8x2(uint32_t v)
{
    uint32_t d;
    asm("min.relu.s16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0x00FF00FFu));
    return d;
}

/* 4x4 byte transpose of four row words (4 chroma columns each) into four column words
 * whose byte j is row j: the vertical taps of one chroma column, ready for IDP.4A. */
__device__ __forceinline__ void transpose4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t (&w)[4])
{
    const uint32_t a = prmt(r0, r1, 0x5140), b = prmt(r2, r3, 0x5140);   /* cols 0,1 */
    const uint32_t c = prmt(r0, r1, 0x7362), d = prmt(r2, r3, 0x7362);   /* cols 2,3 */
    w[0] = prmt(a, b, 0x5410); w[1] = prmt(a, b, 0x7632);
    w[2] = prmt(c, d, 0x5410); w[3] = prmt(c, d, 0x7632);
}

/* NARROW: tile 128 x 64 instead of 256 x 32 (the same bytes per stage): frames whose width is not a multiple of
 * 256 waste less of their right-most tile column (1920 = 15 x 128, 640 = 5 x 128).  A warp then owns 8 rows,
 * lanes 0-15 the first four, lanes 16-31 the last four. */
template <int FMT, int SRC, bool NARROW, bool V422>
__global__ void __launch_bounds__(F420_THREADS, F420_CTAS_PER_SM)
sws_fast420_rgb8_kernel(const __grid_constant__ CUtensorMap map_y,
                        const __grid_constant__ CUtensorMap map_u,
                        const __grid_constant__ CUtensorMap map_v,
                        const __grid_constant__ CUtensorMap map_o,
                        const __grid_constant__ Fast420Args A)
{
    constexpr int BPP = FMT >= F420_RGBA ? 4 : 3;
    constexpr int TW = NARROW ? F420_TW / 2 : F420_TW, TH = NARROW ? 2 * F420_TH : F420_TH;
    constexpr int RPW = TH / F420_CWARPS;                  /* rows per warp: 4 or 8 */
    constexpr int Y_BYTES = TW * TH, META_BYTES = TH * 16;
    /* staged chroma rows: compile-time constants so that the headline instantiation keeps immediate offsets */
    constexpr int CROWS = V422 ? TH + 4 : (NARROW ? F420_CROWS_NARROW : F420_CROWS);
    constexpr int C_BYTES = (TW / 2) * CROWS, IN_BYTES = Y_BYTES + 2 * C_BYTES + META_BYTES;
    static_assert(V422 || IN_BYTES == F420_IN_BYTES, "both tile shapes fill one ring slot for 4:2:0 sources");
    /* [stage: Y | U | V (or interleaved UV) | row meta] x STAGES, then [out: 8 warps x RPW rows] */
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[F420_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F420_STAGES];
    __shared__ __align__(16) int4 tile_info[F420_STAGES];     /* {x tile, first row, frame, -} */

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      /* warp-uniform for the compiler */
    const int tiles_per_frame = A.tiles_x * A.tiles_y;
    const int total = tiles_per_frame * A.frames;

    if (tid == 0) {
        for (int s = 0; s < F420_STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], F420_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == F420_CWARPS) {
        /* ===== producer: one thread feeds the ring with TMA loads ===== */
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_u) : "memory");
            if (SRC == F420_PLANAR)
                asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
            int i = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
                const int stage = i % F420_STAGES, k = i / F420_STAGES;
                if (k > 0)
                    mbar_wait(&empty_bar[stage], (k - 1) & 1);   /* all 8 warps released the slot */
                const int f = tile / tiles_per_frame;
                const int t = tile - f * tiles_per_frame;
                const int ty = t / A.tiles_x, tx = t - ty * A.tiles_x;
                const int y0 = (A.ty_first + ty) * TH;
                const int c_lo = __ldg(&A.rows[y0]).w;
                unsigned char *b = smem_dyn + stage * IN_BYTES;
                tile_info[stage] = make_int4(tx, y0, f, 0);
                mbar_expect_tx(&full_bar[stage], IN_BYTES);
                tma_load_3d(b, &map_y, &full_bar[stage], tx * TW, y0, f);
                if (SRC == F420_PLANAR) {
                    tma_load_3d(b + Y_BYTES, &map_u, &full_bar[stage], tx * (TW / 2), c_lo, f);
                    tma_load_3d(b + Y_BYTES + C_BYTES, &map_v, &full_bar[stage], tx * (TW / 2), c_lo, f);
                } else {   /* nv12 / nv21: one box of interleaved UV rows, same bytes */
                    tma_load_3d(b + Y_BYTES, &map_u, &full_bar[stage], tx * TW, c_lo, f);
                }
                bulk_load_1d(b + Y_BYTES + 2 * C_BYTES, A.rows + y0, META_BYTES, &full_bar[stage]);
            }
        }
        return;
    }

    /* ===== consumers: each warp converts 4 rows of every tile, 8 pixels per lane ===== */
    if (lane == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    const int cy = A.cy, yb = A.yb;
    const int crv = A.crv, cbu = A.cbu, cgu = A.cgu, cgv = A.cgv;
    const int kr = A.kr, kg = A.kg, kb = A.kb;
    const int r0 = warp * RPW;
    const int g = NARROW ? lane & 15 : lane;               /* 8-pixel column group of this lane */
    const int rb = NARROW ? r0 + 4 * (lane >> 4) : r0;     /* first of this lane's four rows */
    unsigned char *so_warp = smem_dyn + F420_STAGES * IN_BYTES + r0 * (TW * BPP);
    unsigned char *so = so_warp + (rb - r0) * (TW * BPP) + g * (8 * BPP);
    /* chroma rows: planar = TW/2-byte U row + TW/2-byte V row (one word each per lane);
     * semi-planar = one TW-byte UV row (two words per lane) */
    constexpr int CSTRIDE = SRC == F420_PLANAR ? TW / 2 : TW;

    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
        const int stage = i % F420_STAGES;
        const unsigned char *sb = smem_dyn + stage * IN_BYTES;
        mbar_wait(&full_bar[stage], (i / F420_STAGES) & 1);

        const int4 ti = tile_info[stage];
        const int4 *mrow = reinterpret_cast<const int4 *>(sb + Y_BYTES + 2 * C_BYTES) + rb;
        const unsigned char *sy = sb + rb * TW + g * 8;
        const unsigned char *sp = sb + Y_BYTES + (SRC == F420_PLANAR ? g * 4 : g * 8);
        const unsigned char *sq = SRC == F420_PLANAR ? sp + C_BYTES : sp + 4;

        /* this warp's previous TMA store must have finished READING its staging rows */
        if (lane == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();

        /* P/Q: byte j of P[k] = source row j of byte lane k of the first / second chroma word */
        uint32_t P[4], Q[4];
        int wpos = -64;                        /* chroma row (tile relative) held in window byte 0 */
#pragma unroll
        for (int rr = 0; rr < 4; rr++) {
            const int4 meta = mrow[rr];
            const int pos = meta.x;
            int d = pos - wpos;
            if (d < 0 || d >= 4) {             /* (re)fill: transpose rows pos..pos+3 */
                const unsigned char *pp = sp + pos * CSTRIDE, *pq = sq + pos * CSTRIDE;
                transpose4(*reinterpret_cast<const uint32_t *>(pp), *reinterpret_cast<const uint32_t *>(pp + CSTRIDE),
                           *reinterpret_cast<const uint32_t *>(pp + 2 * CSTRIDE),
                           *reinterpret_cast<const uint32_t *>(pp + 3 * CSTRIDE), P);
                transpose4(*reinterpret_cast<const uint32_t *>(pq), *reinterpret_cast<const uint32_t *>(pq + CSTRIDE),
                           *reinterpret_cast<const uint32_t *>(pq + 2 * CSTRIDE),
                           *reinterpret_cast<const uint32_t *>(pq + 3 * CSTRIDE), Q);
            } else {
#pragma unroll 1
                for (int nr = wpos + 4; d > 0; d--, nr++) {   /* slide down one source row */
                    const uint32_t np = *reinterpret_cast<const uint32_t *>(sp + nr * CSTRIDE);
                    const uint32_t nq = *reinterpret_cast<const uint32_t *>(sq + nr * CSTRIDE);
                    P[0] = prmt(P[0], np, 0x4321); P[1] = prmt(P[1], np, 0x5321);
                    P[2] = prmt(P[2], np, 0x6321); P[3] = prmt(P[3], np, 0x7321);
                    Q[0] = prmt(Q[0], nq, 0x4321); Q[1] = prmt(Q[1], nq, 0x5321);
                    Q[2] = prmt(Q[2], nq, 0x6321); Q[3] = prmt(Q[3], nq, 0x7321);
                }
            }
            wpos = pos;
            const uint32_t clp = (uint32_t)meta.y, chp = (uint32_t)meta.z;
            const uint2 yw = *reinterpret_cast<const uint2 *>(sy + rr * TW);
            uint32_t h[BPP == 3 ? 12 : 16];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                /* which window registers hold U and V of chroma column c */
                const uint32_t wU = SRC == F420_PLANAR ? P[c]
                                  : SRC == F420_NV12 ? (c < 2 ? P[2 * c] : Q[2 * c - 4])
                                                     : (c < 2 ? P[2 * c + 1] : Q[2 * c - 3]);
                const uint32_t wV = SRC == F420_PLANAR ? Q[c]
                                  : SRC == F420_NV12 ? (c < 2 ? P[2 * c + 1] : Q[2 * c - 3])
                                                     : (c < 2 ? P[2 * c] : Q[2 * c - 4]);
                /* S + 2048 = 256*T + (L + 2048); both dot products are independent */
                const int U = clamp_u8((dp4a_us(wU, chp, 0) * 256 + dp4a_uu(wU, clp, 2048)) >> 12);
                const int V = clamp_u8((dp4a_us(wV, chp, 0) * 256 + dp4a_uu(wV, clp, 2048)) >> 12);
                const int pr = ((V * crv + kr) >> 16) * cy + yb;
                const int pb = ((U * cbu + kb) >> 16) * cy + yb;
                const int pg = (((U * cgu) >> 16) + ((V * cgv + kg) >> 16)) * cy + yb;
                const uint32_t w = (c < 2) ? yw.x : yw.y;
                const int ya = prmt(w, 0u, 0x4440 + 2 * (c & 1));
                const int yc = prmt(w, 0u, 0x4441 + 2 * (c & 1));
                const uint32_t tRa = ya * cy + pr, tGa = ya * cy + pg, tBa = ya * cy + pb;
                const uint32_t tRb = yc * cy + pr, tGb = yc * cy + pg, tBb = yc * cy + pb;
                /* high halves are (t >> 16) as s16: pack pairs, clamp both lanes at once */
                constexpr uint32_t FF = 0x00FF0000u;          /* high half = 255: the opaque alpha */
                if (FMT == F420_RGB24) {
                    h[3 * c + 0] = clamp_u8x2(prmt(tRa, tGa, 0x7632));
                    h[3 * c + 1] = clamp_u8x2(prmt(tBa, tRb, 0x7632));
                    h[3 * c + 2] = clamp_u8x2(prmt(tGb, tBb, 0x7632));
                } else if (FMT == F420_BGR24) {
                    h[3 * c + 0] = clamp_u8x2(prmt(tBa, tGa, 0x7632));
                    h[3 * c + 1] = clamp_u8x2(prmt(tRa, tBb, 0x7632));
                    h[3 * c + 2] = clamp_u8x2(prmt(tGb, tRb, 0x7632));
                } else if (FMT == F420_RGBA) {
                    h[4 * c + 0] = clamp_u8x2(prmt(tRa, tGa, 0x7632)); h[4 * c + 1] = clamp_u8x2(prmt(tBa, FF, 0x7632));
                    h[4 * c + 2] = clamp_u8x2(prmt(tRb, tGb, 0x7632)); h[4 * c + 3] = clamp_u8x2(prmt(tBb, FF, 0x7632));
                } else if (FMT == F420_BGRA) {
                    h[4 * c + 0] = clamp_u8x2(prmt(tBa, tGa, 0x7632)); h[4 * c + 1] = clamp_u8x2(prmt(tRa, FF, 0x7632));
                    h[4 * c + 2] = clamp_u8x2(prmt(tBb, tGb, 0x7632)); h[4 * c + 3] = clamp_u8x2(prmt(tRb, FF, 0x7632));
                } else if (FMT == F420_ARGB) {
                    h[4 * c + 0] = clamp_u8x2(prmt(FF, tRa, 0x7632)); h[4 * c + 1] = clamp_u8x2(prmt(tGa, tBa, 0x7632));
                    h[4 * c + 2] = clamp_u8x2(prmt(FF, tRb, 0x7632)); h[4 * c + 3] = clamp_u8x2(prmt(tGb, tBb, 0x7632));
                } else {
                    h[4 * c + 0] = clamp_u8x2(prmt(FF, tBa, 0x7632)); h[4 * c + 1] = clamp_u8x2(prmt(tGa, tRa, 0x7632));
                    h[4 * c + 2] = clamp_u8x2(prmt(FF, tBb, 0x7632)); h[4 * c + 3] = clamp_u8x2(prmt(tGb, tRb, 0x7632));
                }
            }
            /* every h[] holds two bytes (at bits 0 and 16): gather four of them per output word */
            if (BPP == 3) {
                uint2 *o = reinterpret_cast<uint2 *>(so + rr * (TW * 3));
                o[0] = make_uint2(prmt(h[0], h[1], 0x6420), prmt(h[2], h[3], 0x6420));
                o[1] = make_uint2(prmt(h[4], h[5], 0x6420), prmt(h[6], h[7], 0x6420));
                o[2] = make_uint2(prmt(h[8], h[9], 0x6420), prmt(h[10], h[11], 0x6420));
            } else {
                uint4 *o = reinterpret_cast<uint4 *>(so + rr * (TW * 4));
                o[0] = make_uint4(prmt(h[0], h[1], 0x6420), prmt(h[2], h[3], 0x6420),
                                  prmt(h[4], h[5], 0x6420), prmt(h[6], h[7], 0x6420));
                o[1] = make_uint4(prmt(h[8], h[9], 0x6420), prmt(h[10], h[11], 0x6420),
                                  prmt(h[12], h[13], 0x6420), prmt(h[14], h[15], 0x6420));
            }
        }

        /* publish this warp's 4 rows: generic-proxy writes -> async proxy, one TMA store per warp;
         * the input slot is released at the same point (all lanes have finished reading it) */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty_bar[stage]);
            tma_store_3d(&map_o, so_warp, ti.x * (TW * BPP / 4), ti.y + r0, ti.z);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
