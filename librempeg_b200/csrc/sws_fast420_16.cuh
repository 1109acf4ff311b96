/*
 * sws_fast420_16.cuh -- planar 9..16-bit 4:2:0 -> rgb48le / bgr48le when the luma path is
 * the identity and chroma is only filtered vertically (BASELINE config C3: 4K yuv420p10le ->
 * rgb48le, lanczos => 6 vertical chroma taps).
 *
 * Same skeleton as sws_fast420.cuh (TMA producer warp, 2-stage mbarrier ring, 8 consumer warps x
 * 4 rows, per-warp TMA tensor stores); tile 128 x 32 pixels so one output row of the tile is the
 * same 768 bytes.  Arithmetic restates, with the C code's 32-bit unsigned wrap-around,
 *   hScale16To19_c (identity: x << (19 - depth), clipped to 2^19-1)   libswscale/swscale.c:69-97
 *   yuv2rgba64_X_c_template / _1 (hasAlpha = 0, eightbytes = 0)         libswscale/output.c:1115-1371
 * With a 1-tap 4096 luma filter  Y1 = ((l19*4096 - 2^30) >> 14) + 2^16 = l19 >> 2, and
 *   out = clip_u16((int)(C + (Y1 - y_off)*y_coeff + 2^13 - 2^29) >> 14) + 2^15)
 * becomes one IMAD per pixel and channel: t = Y1*y_coeff + P_c with the per-pair constant
 * P_c = C + 2^13 - 2^29 - y_off*y_coeff (all mod 2^32, as in the reference).
 */
#pragma once

#include "sws_fast420.cuh"

#define F16_TW 128
#define F16_TH 32
#define F16_CROWS 24
#define F16_Y_BYTES (F16_TW * 2 * F16_TH)                                    /*  8192 */
#define F16_C_BYTES ((F16_TW / 2) * 2 * F16_CROWS)                           /*  3072 */
#define F16_META_BYTES (F16_TH * 32)                                         /*  1024 */
#define F16_IN_BYTES (F16_Y_BYTES + 2 * F16_C_BYTES + F16_META_BYTES)        /* 15360 */
#define F16_OUT_BYTES (F16_TW * 6 * F16_TH)                                  /* 24576 */
#define F16_SMEM (F420_STAGES * F16_IN_BYTES + F16_OUT_BYTES)

struct Fast16Row {            /* per output row, 32 bytes */
    int pos_rel, pos_abs;     /* first chroma source row: relative to the tile's first, absolute */
    int c01, c23, c45, c67;   /* eight int16 vertical chroma taps */
    int pad0, pad1;
};

struct Fast16Args {
    int tiles_x, tiles_y, frames, ty_first, dst_h;
    int s19;                  /* 19 - source depth */
    int bgr;
    unsigned ycoef, kconst;   /* y_coeff, 2^13 - 2^29 - y_offset*y_coeff */
    unsigned v2r, v2g, u2g, u2b;
    const Fast16Row *rows;
};

__device__ __forceinline__ uint32_t pack_sat_s16(int hi, int lo)
{
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

template <int TAPS, bool BGR>
__global__ void __launch_bounds__(F420_THREADS, F420_CTAS_PER_SM)
sws_fast420_rgb16_kernel(const __grid_constant__ CUtensorMap map_y,
                         const __grid_constant__ CUtensorMap map_u,
                         const __grid_constant__ CUtensorMap map_v,
                         const __grid_constant__ CUtensorMap map_o,
                         const __grid_constant__ Fast16Args A)
{
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[F420_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F420_STAGES];
    __shared__ __align__(16) int4 tile_info[F420_STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
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
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_u) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
            int i = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
                const int stage = i % F420_STAGES, k = i / F420_STAGES;
                if (k > 0)
                    mbar_wait(&empty_bar[stage], (k - 1) & 1);
                const int f = tile / tiles_per_frame;
                const int t = tile - f * tiles_per_frame;
                const int ty = t / A.tiles_x, tx = t - ty * A.tiles_x;
                const int y0 = (A.ty_first + ty) * F16_TH;
                const int c_lo = __ldg(&A.rows[y0].pos_abs);
                unsigned char *b = smem_dyn + stage * F16_IN_BYTES;
                tile_info[stage] = make_int4(tx, y0, f, 0);
                mbar_expect_tx(&full_bar[stage], F16_IN_BYTES);
                tma_load_3d(b, &map_y, &full_bar[stage], tx * F16_TW, y0, f);
                tma_load_3d(b + F16_Y_BYTES, &map_u, &full_bar[stage], tx * (F16_TW / 2), c_lo, f);
                tma_load_3d(b + F16_Y_BYTES + F16_C_BYTES, &map_v, &full_bar[stage], tx * (F16_TW / 2), c_lo, f);
                bulk_load_1d(b + F16_Y_BYTES + 2 * F16_C_BYTES, A.rows + y0, F16_META_BYTES, &full_bar[stage]);
            }
        }
        return;
    }

    if (lane == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    const int s19 = A.s19;
    const uint32_t smask = 0xFFFFu << s19;
    const unsigned ycoef = A.ycoef, kconst = A.kconst;
    const unsigned v2r = A.v2r, v2g = A.v2g, u2g = A.u2g, u2b = A.u2b;
    const int r0 = warp * (F16_TH / F420_CWARPS);
    unsigned char *so_warp = smem_dyn + F420_STAGES * F16_IN_BYTES + r0 * (F16_TW * 6);
    unsigned char *so = so_warp + lane * 24;

    /* 16-bit sample (either half of a packed word) -> 19-bit h-scaled line value */
    auto lo19 = [&](uint32_t w) { return min((int)((w << s19) & smask), (1 << 19) - 1); };
    auto hi19 = [&](uint32_t w) { return min((int)((w >> (16 - s19)) & smask), (1 << 19) - 1); };

    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
        const int stage = i % F420_STAGES;
        const unsigned char *sb = smem_dyn + stage * F16_IN_BYTES;
        mbar_wait(&full_bar[stage], (i / F420_STAGES) & 1);

        const int4 ti = tile_info[stage];
        const int4 *mrow = reinterpret_cast<const int4 *>(sb + F16_Y_BYTES + 2 * F16_C_BYTES) + 2 * r0;
        const unsigned char *sy = sb + r0 * (F16_TW * 2) + lane * 8;
        const unsigned char *su = sb + F16_Y_BYTES + lane * 4;
        const unsigned char *sv = su + F16_C_BYTES;

        if (lane == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();

        int wu[TAPS][2], wv[TAPS][2];          /* unpacked 19-bit samples of TAPS source rows */
        int wpos = -64;
#pragma unroll
        for (int rr = 0; rr < F16_TH / F420_CWARPS; rr++) {
            const int4 m0 = mrow[2 * rr], m1 = mrow[2 * rr + 1];
            const int pos = m0.x;
            int d = pos - wpos;
            if (d < 0 || d >= TAPS) {
#pragma unroll
                for (int j = 0; j < TAPS; j++) {
                    const uint32_t nu = *reinterpret_cast<const uint32_t *>(su + (pos + j) * F16_TW);
                    const uint32_t nv = *reinterpret_cast<const uint32_t *>(sv + (pos + j) * F16_TW);
                    wu[j][0] = lo19(nu); wu[j][1] = hi19(nu);
                    wv[j][0] = lo19(nv); wv[j][1] = hi19(nv);
                }
            } else {
#pragma unroll 1
                for (int nr = wpos + TAPS; d > 0; d--, nr++) {
                    const uint32_t nu = *reinterpret_cast<const uint32_t *>(su + nr * F16_TW);
                    const uint32_t nv = *reinterpret_cast<const uint32_t *>(sv + nr * F16_TW);
#pragma unroll
                    for (int j = 0; j < TAPS - 1; j++) {
                        wu[j][0] = wu[j + 1][0]; wu[j][1] = wu[j + 1][1];
                        wv[j][0] = wv[j + 1][0]; wv[j][1] = wv[j + 1][1];
                    }
                    wu[TAPS - 1][0] = lo19(nu); wu[TAPS - 1][1] = hi19(nu);
                    wv[TAPS - 1][0] = lo19(nv); wv[TAPS - 1][1] = hi19(nv);
                }
            }
            wpos = pos;
            int cf[8];
            cf[0] = (int)(short)(m0.z & 0xFFFF); cf[1] = m0.z >> 16;
            cf[2] = (int)(short)(m0.w & 0xFFFF); cf[3] = m0.w >> 16;
            cf[4] = (int)(short)(m1.x & 0xFFFF); cf[5] = m1.x >> 16;
            cf[6] = (int)(short)(m1.y & 0xFFFF); cf[7] = m1.y >> 16;

            const uint2 yw = *reinterpret_cast<const uint2 *>(sy + rr * (F16_TW * 2));
            unsigned v[12];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                unsigned U = 0u - (128u << 23), V = 0u - (128u << 23);
#pragma unroll
                for (int j = 0; j < TAPS; j++) {
                    U += (unsigned)wu[j][c] * (unsigned)cf[j];
                    V += (unsigned)wv[j][c] * (unsigned)cf[j];
                }
                const unsigned Ui = (unsigned)((int)U >> 14), Vi = (unsigned)((int)V >> 14);
                const unsigned pr = Vi * v2r + kconst;
                const unsigned pg = Vi * v2g + (Ui * u2g + kconst);
                const unsigned pb = Ui * u2b + kconst;
                const uint32_t w = c ? yw.y : yw.x;
                /* Y1 = l19 >> 2 = min(y << (s19 - 2), 2^17 - 1) */
                const unsigned ya = (unsigned)min((int)((w << (s19 - 2)) & (smask >> 2)), (1 << 17) - 1);
                const unsigned yc = (unsigned)min((int)((w >> (18 - s19)) & (smask >> 2)), (1 << 17) - 1);
                const unsigned p0 = BGR ? pb : pr, p2 = BGR ? pr : pb;
                v[6 * c + 0] = ya * ycoef + p0; v[6 * c + 1] = ya * ycoef + pg; v[6 * c + 2] = ya * ycoef + p2;
                v[6 * c + 3] = yc * ycoef + p0; v[6 * c + 4] = yc * ycoef + pg; v[6 * c + 5] = yc * ycoef + p2;
            }
            uint32_t o[6];
#pragma unroll
            for (int k = 0; k < 6; k++)     /* clip_u16(t + 2^15) == sat_s16(t) ^ 0x8000 */
                o[k] = pack_sat_s16((int)v[2 * k + 1] >> 14, (int)v[2 * k] >> 14) ^ 0x80008000u;
            uint2 *op = reinterpret_cast<uint2 *>(so + rr * (F16_TW * 6));
            op[0] = make_uint2(o[0], o[1]);
            op[1] = make_uint2(o[2], o[3]);
            op[2] = make_uint2(o[4], o[5]);
        }

        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty_bar[stage]);
            tma_store_3d(&map_o, so_warp, ti.x * (F16_TW * 6 / 4), ti.y + r0, ti.z);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
