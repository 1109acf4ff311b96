/*
 * sws_fast420_hi8.cuh -- planar 9..16-bit 4:2:0 or p010le -> packed 8-bit RGB (rgb24 / bgr24 / rgba / bgra /
 * argb / abgr) of the same size: the 10-bit-video-to-display conversion (p010le is what hardware decoders emit:
 * samples in the high 10 bits, chroma interleaved; p010LEToY/UV_c, input.c:950-1006).  Luma path = identity, chroma only
 * filtered vertically (<= 8 taps), one chroma sample per pixel pair.
 *
 * Same skeleton as sws_fast420_16.cuh (TMA producer warp, mbarrier ring, 8 consumer warps x 4 rows, per-warp
 * TMA tensor stores, tile 128 x 32, a lane owns 4 pixels = 2 chroma columns); the arithmetic restates
 *   hScale16To15_c with the identity filter: min((x << 14) >> (depth - 1), 32767)      libswscale/swscale.c:99-125
 *   yuv2rgb_X_c_template + yuv2rgb_write (8-bit targets), also what vscale.c:135-147     libswscale/output.c:1662-1840
 *   hands to the algebraically identical _1 variants when chroma has 1 or 2 taps
 * With a 1-tap 4096 luma filter  Y = (l15 * 4096 + 2^18) >> 19 = (l15 + 64) >> 7 (not clipped: the byte LUTs
 * have headroom), U = clip_u8((2^18 + sum u15_j c_j) >> 19), and the LUT chain in closed form as in
 * sws_fast420.cuh: out = clip_u8((Y * cy + (base_c + chroma term) * cy + yb) >> 16).
 */
#pragma once

#include "sws_fast420_16.cuh"

#define H8_OUT_BYTES(bpp) (F16_TW * (bpp) * F16_TH)                          /* 12288 / 16384 */
/* chroma rows staged per tile are a launch parameter: 24 for 4:2:0 (16 + the vertical taps), 36 for 4:2:2 sources
 * whose chroma is not subsampled vertically (32 + taps: yuv422p10le, the ProRes / DNxHR decode format) */
#define H8_C_BYTES(crows) ((F16_TW / 2) * 2 * (crows))
#define H8_IN_BYTES(crows) (F16_Y_BYTES + 2 * H8_C_BYTES(crows) + F16_META_BYTES)
#define H8_SMEM(bpp, crows) (F420_STAGES * H8_IN_BYTES(crows) + H8_OUT_BYTES(bpp))

struct FastHi8Args {
    int tiles_x, tiles_y, frames, ty_first, dst_h;
    int sdown;                /* source depth - 1: the right shift of the identity horizontal filter */
    int sshift;               /* position of the samples in their 16-bit containers (p010: 6) */
    int crows;                /* chroma source rows staged per tile (24 or 36) */
    int cy, yb;               /* LUT closed form (sws_colorspace.c) */
    int crv, cbu, cgu, cgv;
    int base_r, base_g, base_b;
    const Fast16Row *rows;
};

__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int TAPS, int FMT, bool SEMI>
__global__ void __launch_bounds__(F420_THREADS, F420_CTAS_PER_SM)
sws_fast420_hi8_kernel(const __grid_constant__ CUtensorMap map_y,
                       const __grid_constant__ CUtensorMap map_u,
                       const __grid_constant__ CUtensorMap map_v,
                       const __grid_constant__ CUtensorMap map_o,
                       const __grid_constant__ FastHi8Args A)
{
    constexpr int BPP = FMT >= F420_RGBA ? 4 : 3;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[F420_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[F420_STAGES];
    __shared__ __align__(16) int4 tile_info[F420_STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int tiles_per_frame = A.tiles_x * A.tiles_y;
    const int total = tiles_per_frame * A.frames;
    const int c_bytes = H8_C_BYTES(A.crows), in_bytes = H8_IN_BYTES(A.crows);

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
            if (!SEMI)
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
                unsigned char *b = smem_dyn + stage * in_bytes;
                tile_info[stage] = make_int4(tx, y0, f, 0);
                mbar_expect_tx(&full_bar[stage], in_bytes);
                tma_load_3d(b, &map_y, &full_bar[stage], tx * F16_TW, y0, f);
                if (SEMI) {          /* one box of interleaved UV rows: the same bytes as the two planar boxes */
                    tma_load_3d(b + F16_Y_BYTES, &map_u, &full_bar[stage], tx * F16_TW, c_lo, f);
                } else {
                    tma_load_3d(b + F16_Y_BYTES, &map_u, &full_bar[stage], tx * (F16_TW / 2), c_lo, f);
                    tma_load_3d(b + F16_Y_BYTES + c_bytes, &map_v, &full_bar[stage], tx * (F16_TW / 2), c_lo, f);
                }
                bulk_load_1d(b + F16_Y_BYTES + 2 * c_bytes, A.rows + y0, F16_META_BYTES, &full_bar[stage]);
            }
        }
        return;
    }

    /* ===== consumers: each warp converts 4 rows of every tile, 4 pixels per lane ===== */
    if (lane == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_o) : "memory");
    const int sdown = A.sdown, sshift = A.sshift;
    const int cy = A.cy, yb = A.yb;
    const int crv = A.crv, cbu = A.cbu, cgu = A.cgu, cgv = A.cgv;
    const int r0 = warp * (F16_TH / F420_CWARPS);
    unsigned char *so_warp = smem_dyn + F420_STAGES * in_bytes + r0 * (F16_TW * BPP);
    unsigned char *so = so_warp + lane * (4 * BPP);

    /* packed form: both 16-bit samples of a word -> two 15-bit line values.  x is clipped to 2^depth first, so the
     * left shift stays inside its half and lands on 2^15 exactly where the reference clips: min(., 32767) finishes it.
     * Three instructions per sample pair instead of five per sample. */
    const int kshift = 14 - sdown;                       /* 15 - depth: 1..6; -1 for 16-bit samples */
    const uint32_t lim2 = kshift >= 0 ? 0x00010001u << (sdown + 1) : 0u;
    const uint32_t smask2 = 0x00010001u * (0xFFFFu >> sshift);
    auto l15x2 = [&](uint32_t w) -> uint32_t {
        if (kshift < 0)
            return (w >> 1) & 0x7FFF7FFFu;
        if (SEMI)
            w = (w >> sshift) & smask2;
        return min_u16x2(min_u16x2(w, lim2) << kshift, 0x7FFF7FFFu);
    };
    /* chroma words of one source row: two U and two V samples of this lane's columns */
    auto chroma = [&](const unsigned char *su, const unsigned char *sv, int row, uint32_t &nu, uint32_t &nv) {
        if (SEMI) {
            const uint2 q = *reinterpret_cast<const uint2 *>(su + row * (2 * F16_TW));
            nu = prmt(q.x, q.y, 0x5410);
            nv = prmt(q.x, q.y, 0x7632);
        } else {
            nu = *reinterpret_cast<const uint32_t *>(su + row * F16_TW);
            nv = *reinterpret_cast<const uint32_t *>(sv + row * F16_TW);
        }
    };

    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, i++) {
        const int stage = i % F420_STAGES;
        const unsigned char *sb = smem_dyn + stage * in_bytes;
        mbar_wait(&full_bar[stage], (i / F420_STAGES) & 1);

        const int4 ti = tile_info[stage];
        const int4 *mrow = reinterpret_cast<const int4 *>(sb + F16_Y_BYTES + 2 * c_bytes) + 2 * r0;
        const unsigned char *sy = sb + r0 * (F16_TW * 2) + lane * 8;
        const unsigned char *su = sb + F16_Y_BYTES + lane * (SEMI ? 8 : 4);
        const unsigned char *sv = su + c_bytes;

        /* this warp's previous TMA store must have finished READING its staging rows */
        if (lane == 0)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();

        int wu[TAPS][2], wv[TAPS][2];          /* unpacked 15-bit samples of TAPS source rows */
        int wpos = -64;
#pragma unroll
        for (int rr = 0; rr < F16_TH / F420_CWARPS; rr++) {
            const int4 m0 = mrow[2 * rr], m1 = mrow[2 * rr + 1];
            const int pos = m0.x;
            int d = pos - wpos;
            if (d < 0 || d >= TAPS) {
#pragma unroll
                for (int j = 0; j < TAPS; j++) {
                    uint32_t nu, nv;
                    chroma(su, sv, pos + j, nu, nv);
                    const uint32_t pu = l15x2(nu), pv = l15x2(nv);
                    wu[j][0] = (int)(pu & 0xFFFFu); wu[j][1] = (int)(pu >> 16);
                    wv[j][0] = (int)(pv & 0xFFFFu); wv[j][1] = (int)(pv >> 16);
                }
            } else {
#pragma unroll 1
                for (int nr = wpos + TAPS; d > 0; d--, nr++) {
                    uint32_t nu, nv;
                    chroma(su, sv, nr, nu, nv);
#pragma unroll
                    for (int j = 0; j < TAPS - 1; j++) {
                        wu[j][0] = wu[j + 1][0]; wu[j][1] = wu[j + 1][1];
                        wv[j][0] = wv[j + 1][0]; wv[j][1] = wv[j + 1][1];
                    }
                    const uint32_t pu = l15x2(nu), pv = l15x2(nv);
                    wu[TAPS - 1][0] = (int)(pu & 0xFFFFu); wu[TAPS - 1][1] = (int)(pu >> 16);
                    wv[TAPS - 1][0] = (int)(pv & 0xFFFFu); wv[TAPS - 1][1] = (int)(pv >> 16);
                }
            }
            wpos = pos;
            int cf[8];
            cf[0] = (int)(short)(m0.z & 0xFFFF); cf[1] = m0.z >> 16;
            cf[2] = (int)(short)(m0.w & 0xFFFF); cf[3] = m0.w >> 16;
            cf[4] = (int)(short)(m1.x & 0xFFFF); cf[5] = m1.x >> 16;
            cf[6] = (int)(short)(m1.y & 0xFFFF); cf[7] = m1.y >> 16;

            const uint2 yw = *reinterpret_cast<const uint2 *>(sy + rr * (F16_TW * 2));
            uint32_t tR[4], tG[4], tB[4];      /* per pixel: Y * cy + P_channel, result = bits 16..23 after the clip */
#pragma unroll
            for (int c = 0; c < 2; c++) {
                unsigned U = 1u << 18, V = 1u << 18;
#pragma unroll
                for (int j = 0; j < TAPS; j++) {
                    U += (unsigned)wu[j][c] * (unsigned)cf[j];
                    V += (unsigned)wv[j][c] * (unsigned)cf[j];
                }
                const int u8 = clamp_u8((int)U >> 19), v8 = clamp_u8((int)V >> 19);
                const int pR = (A.base_r + ((v8 * crv) >> 16)) * cy + yb;
                const int pG = (A.base_g + ((u8 * cgu) >> 16) + ((v8 * cgv) >> 16)) * cy + yb;
                const int pB = (A.base_b + ((u8 * cbu) >> 16)) * cy + yb;
                const uint32_t w = c ? yw.y : yw.x;
                const uint32_t yp = ((l15x2(w) + 0x00400040u) >> 7) & 0x01FF01FFu;     /* (l15 + 64) >> 7 in both halves */
                const int ya = (int)(yp & 0xFFFFu), yc = (int)(yp >> 16);
                tR[2 * c] = ya * cy + pR; tG[2 * c] = ya * cy + pG; tB[2 * c] = ya * cy + pB;
                tR[2 * c + 1] = yc * cy + pR; tG[2 * c + 1] = yc * cy + pG; tB[2 * c + 1] = yc * cy + pB;
            }
            constexpr uint32_t FF = 0x00FF0000u;          /* high half = 255: the opaque alpha */
            if (BPP == 3) {
                /* six half-word pairs per two pixels, then four bytes per output word */
                uint32_t h[6];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const uint32_t x0 = FMT == F420_RGB24 ? tR[2 * c] : tB[2 * c], x1 = FMT == F420_RGB24 ? tB[2 * c] : tR[2 * c];
                    const uint32_t z0 = FMT == F420_RGB24 ? tR[2 * c + 1] : tB[2 * c + 1], z1 = FMT == F420_RGB24 ? tB[2 * c + 1] : tR[2 * c + 1];
                    h[3 * c + 0] = clamp_u8x2(prmt(x0, tG[2 * c], 0x7632));
                    h[3 * c + 1] = clamp_u8x2(prmt(x1, z0, 0x7632));
                    h[3 * c + 2] = clamp_u8x2(prmt(tG[2 * c + 1], z1, 0x7632));
                }
                uint32_t *op = reinterpret_cast<uint32_t *>(so + rr * (F16_TW * 3));
                op[0] = prmt(h[0], h[1], 0x6420);
                op[1] = prmt(h[2], h[3], 0x6420);
                op[2] = prmt(h[4], h[5], 0x6420);
            } else {
                uint32_t o[4];
#pragma unroll
                for (int px = 0; px < 4; px++) {
                    uint32_t a0, a1;
                    if (FMT == F420_RGBA) {
                        a0 = prmt(tR[px], tG[px], 0x7632); a1 = prmt(tB[px], FF, 0x7632);
                    } else if (FMT == F420_BGRA) {
                        a0 = prmt(tB[px], tG[px], 0x7632); a1 = prmt(tR[px], FF, 0x7632);
                    } else if (FMT == F420_ARGB) {
                        a0 = prmt(FF, tR[px], 0x7632); a1 = prmt(tG[px], tB[px], 0x7632);
                    } else {
                        a0 = prmt(FF, tB[px], 0x7632); a1 = prmt(tG[px], tR[px], 0x7632);
                    }
                    o[px] = prmt(clamp_u8x2(a0), clamp_u8x2(a1), 0x6420);
                }
                *reinterpret_cast<uint4 *>(so + rr * (F16_TW * 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }

        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty_bar[stage]);
            tma_store_3d(&map_o, so_warp, ti.x * (F16_TW * BPP / 4), ti.y + r0, ti.z);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
