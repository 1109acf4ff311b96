/*
 * sws_cuda.cu -- sm_100a kernels + the C-ABI shim of the B200 libswscale hot path.
 *
 * One fused kernel per output tile replaces the reference's per-line pipeline
 *   lumToYV12/chrToYV12  -> hyScale/hcScale -> [lum/chrConvertRange]
 *   -> yuv2plane1/X | yuv2nv12cX | yuv2packed1/2/X
 * (libswscale/input.c:926-941, swscale.c:69-255, output.c:163-528,1115-1196,
 *  1788-1939; scheduling in swscale.c:263-567 and vscale.c:41-171).
 *
 * Bit-exactness rules honoured here:
 *  - the h-scaled lines are materialised exactly as the reference stores them
 *    (int16 clipped to 15 bits, or int32 clipped to 19 bits) in shared memory;
 *    H and V filters are never merged algebraically;
 *  - all accumulations wrap in 32 bits exactly where the C code's `unsigned`
 *    casts make them wrap; right shifts of signed values are arithmetic;
 *  - the 8-bit RGB LUT chain is evaluated in closed form (sws_colorspace.c).
 */
#include <cuda_runtime.h>
#include <climits>
#include <type_traits>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <errno.h>

#include "sws_internal.h"
#include "sws_fast420.cuh"
#include "sws_fast420_16.cuh"

#define CUDA_OK(call)                                                           \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            fprintf(stderr, "[swscaler-b200] %s failed: %s (%s:%d)\n", #call,   \
                    cudaGetErrorString(e_), __FILE__, __LINE__);                \
            return AVERROR(EIO);                                                \
        }                                                                       \
    } while (0)

/* NVTX ranges (header-only NVTX 3: free when no tool is attached) around the three things a timeline of this library
 * shows: host -> device rows, the conversion launch, device -> host rows; plus the whole-frame pipelines */
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

/* Every entry point runs on the device the context was created on, whatever device the calling thread
 * has current, and leaves the caller's current device as it found it. */
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
            switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched)
            cudaSetDevice(prev);
    }
};

/* cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION (per device), not of a launch: two
 * live contexts that use the same kernel instantiation with different tile geometries would lower each other's
 * limit.  Keep the largest size ever asked for. */
#include <mutex>
static cudaError_t set_max_smem(const void *func, size_t bytes)
{
    struct Entry { const void *func; int dev; size_t bytes; };
    static Entry table[512];
    static int used;
    static std::mutex lock;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(lock);
    Entry *e = nullptr;
    for (int i = 0; i < used; i++)
        if (table[i].func == func && table[i].dev == dev)
            e = &table[i];
    if (e && e->bytes >= bytes)
        return cudaSuccess;
    cudaError_t r = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (r != cudaSuccess)
        return r;
    if (!e && used < 512)
        e = &table[used++];
    if (e) {
        e->func = func; e->dev = dev; e->bytes = bytes;
    }
    return cudaSuccess;
}

/* ordered-dither rows for 8-bit planar output of >8-bit sources; same matrix as
 * ff_dither_8x8_128 (reference swscale.c:42-52) */
__constant__ __align__(8) uint8_t c_dither_8x8_128[8][8] = {
    {  36, 68,  60, 92,  34, 66,  58, 90 },
    { 100,  4, 124, 28,  98,  2, 122, 26 },
    {  52, 84,  44, 76,  50, 82,  42, 74 },
    { 116, 20, 108, 12, 114, 18, 106, 10 },
    {  32, 64,  56, 88,  38, 70,  62, 94 },
    {  96,  0, 120, 24, 102,  6, 126, 30 },
    {  48, 80,  40, 72,  54, 86,  46, 78 },
    { 112, 16, 104,  8, 118, 22, 110, 14 },
};

struct FrameArgs {
    const uint8_t *src[4];
    uint8_t *dst[4];
    long long src_fstride[4];
    long long dst_fstride[4];
    int src_stride[4];
    int dst_stride[4];
    int y0, y1;          /* destination row range */
    int tile_w, tile_h;  /* luma output tile */
    int rows_l_cap, rows_c_cap; /* shared-memory row capacity */
};

__device__ __forceinline__ int clip_u8(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int clip_uintp2(int v, int bits) { return min(max(v, 0), (1 << bits) - 1); }
__device__ __forceinline__ int clip_i16(int v) { return min(max(v, -32768), 32767); }

template <bool SRC16>
__device__ __forceinline__ int fetch(const uint8_t *row, int idx, int shift = 0)
{
    if (SRC16)      /* p010: the sample sits in the high bits (p010LEToY_c / p010LEToUV_c, input.c:950-1006) */
        return reinterpret_cast<const uint16_t *>(row)[idx] >> shift;
    return row[idx];
}

/* Packed 8-bit RGB readers: one 14-bit luma or chroma sample (8-bit value << 6) from pixel x of a
 * row, restating rgb24ToY_c / rgb24ToUV_c / rgb24ToUV_half_c (reference input.c:1068-1180) for
 * 3-byte pixels and the rgb16_32To{Y,UV,UV_half}_c_template instances for 4-byte pixels
 * (input.c:264-345,391-394: coefficients << 8, unsigned rounding constant, logical shift).  Both
 * store into int16 lines that the horizontal scaler reads back as uint16 (swscale.c:99-125). */
/* 48-bit pixels: rgb48ToY_c / rgb48ToUV_c / rgb48ToUV_half_c (input.c:111-196): 16-bit components, the matrix product in
 * unsigned arithmetic (int32 coefficient x unsigned sample), rounding 0x2001 << 14 (luma) / 0x10001 << 14 (chroma), a
 * logical >> 15 and a 16-bit store; the *_half reader averages the pixel pair with rounding first */
__device__ __forceinline__ int rgb48_luma16(const SwsCudaPlan &P, const uint8_t *row, int x)
{
    const uint16_t *px = reinterpret_cast<const uint16_t *>(row) + 3 * x;
    const unsigned r = px[P.src_ro], g = px[P.src_go], b = px[P.src_bo];
    const unsigned s = (unsigned)P.rgb2yuv[0] * r + (unsigned)P.rgb2yuv[1] * g + (unsigned)P.rgb2yuv[2] * b;
    return (uint16_t)((s + (0x2001u << 14)) >> 15);
}

__device__ __forceinline__ void rgb48_chroma16(const SwsCudaPlan &P, const uint8_t *row, int x, int &u, int &v)
{
    unsigned r, g, b;
    if (P.src_rgb_half) {
        const uint16_t *px = reinterpret_cast<const uint16_t *>(row) + 6 * x;
        r = (px[P.src_ro] + px[3 + P.src_ro] + 1u) >> 1;
        g = (px[P.src_go] + px[3 + P.src_go] + 1u) >> 1;
        b = (px[P.src_bo] + px[3 + P.src_bo] + 1u) >> 1;
    } else {
        const uint16_t *px = reinterpret_cast<const uint16_t *>(row) + 3 * x;
        r = px[P.src_ro]; g = px[P.src_go]; b = px[P.src_bo];
    }
    const unsigned su = (unsigned)P.rgb2yuv[3] * r + (unsigned)P.rgb2yuv[4] * g + (unsigned)P.rgb2yuv[5] * b;
    const unsigned sv = (unsigned)P.rgb2yuv[6] * r + (unsigned)P.rgb2yuv[7] * g + (unsigned)P.rgb2yuv[8] * b;
    u = (uint16_t)((su + (0x10001u << 14)) >> 15);
    v = (uint16_t)((sv + (0x10001u << 14)) >> 15);
}

__device__ __forceinline__ int rgb_luma14(const SwsCudaPlan &P, const uint8_t *row, int x)
{
    if (P.src_bpp == 6)
        return rgb48_luma16(P, row, x);
    const uint8_t *px = row + x * P.src_bpp;
    const int r = px[P.src_ro], g = px[P.src_go], b = px[P.src_bo];
    const int s = P.rgb2yuv[0] * r + P.rgb2yuv[1] * g + P.rgb2yuv[2] * b;
    if (P.src_bpp == 3)
        return (uint16_t)((s + (32 << 14) + (1 << 8)) >> 9);
    return (uint16_t)(((unsigned)s * 256u + (32u << 22) + (1u << 16)) >> 17);
}

__device__ __forceinline__ void rgb_chroma14(const SwsCudaPlan &P, const uint8_t *row, int x, int &u, int &v)
{
    if (P.src_bpp == 6) {
        rgb48_chroma16(P, row, x, u, v);
        return;
    }
    int r, g, b;
    if (P.src_rgb_half) {
        const uint8_t *px = row + 2 * x * P.src_bpp, *qx = px + P.src_bpp;
        r = px[P.src_ro] + qx[P.src_ro]; g = px[P.src_go] + qx[P.src_go]; b = px[P.src_bo] + qx[P.src_bo];
    } else {
        const uint8_t *px = row + x * P.src_bpp;
        r = px[P.src_ro]; g = px[P.src_go]; b = px[P.src_bo];
    }
    const int su = P.rgb2yuv[3] * r + P.rgb2yuv[4] * g + P.rgb2yuv[5] * b;
    const int sv = P.rgb2yuv[6] * r + P.rgb2yuv[7] * g + P.rgb2yuv[8] * b;
    if (P.src_bpp == 3) {
        if (P.src_rgb_half) {
            u = (uint16_t)((su + (256 << 15) + (1 << 9)) >> 10);
            v = (uint16_t)((sv + (256 << 15) + (1 << 9)) >> 10);
        } else {
            u = (uint16_t)((su + (256 << 14) + (1 << 8)) >> 9);
            v = (uint16_t)((sv + (256 << 14) + (1 << 8)) >> 9);
        }
    } else if (P.src_rgb_half) {
        u = (uint16_t)(((unsigned)su * 256u + (256u << 23) + (1u << 17)) >> 18);
        v = (uint16_t)(((unsigned)sv * 256u + (256u << 23) + (1u << 17)) >> 18);
    } else {
        u = (uint16_t)(((unsigned)su * 256u + (256u << 22) + (1u << 16)) >> 17);
        v = (uint16_t)(((unsigned)sv * 256u + (256u << 22) + (1u << 16)) >> 17);
    }
}

#include "sws_scale8.cuh"
#include "sws_tile15.cuh"
#include "sws_rgb420.cuh"
#include "sws_fast420_hi8.cuh"


/* ------------------------------------------------------------------------
 * Unscaled special converters (reference swscale_unscaled.c), pure HBM-bound byte work.
 * ------------------------------------------------------------------------ */
struct ShuffleArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long src_fstride, dst_fstride;
    int src_stride, dst_stride;
    int w, y0;
    int map[4];            /* destination byte k of a pixel <- source byte map[k]; 4 = constant 255 */
};

/* packedCopyWrapper (identical 48-bit formats) and rgb48tobgr48_nobswap (rgb2rgb_template.c via rgbToRgbWrapper):
 * components 0 and 2 of every pixel change places, or nothing does.  One pixel per thread. */
struct Rgb48Args {
    const uint8_t *src;
    uint8_t *dst;
    long long src_fstride, dst_fstride;
    int src_stride, dst_stride;
    int w, y0, rows, swap;
};

__global__ void __launch_bounds__(256)
sws_rgb48_kernel(const Rgb48Args A)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)A.w * A.rows)
        return;
    const int y = A.y0 + (int)(idx / A.w), x = (int)(idx % A.w);
    const uint16_t *s = reinterpret_cast<const uint16_t *>(A.src + blockIdx.z * A.src_fstride + (size_t)y * A.src_stride) + 3 * x;
    uint16_t *d = reinterpret_cast<uint16_t *>(A.dst + blockIdx.z * A.dst_fstride + (size_t)y * A.dst_stride) + 3 * x;
    const uint16_t c0 = s[0], c1 = s[1], c2 = s[2];
    d[0] = A.swap ? c2 : c0;
    d[1] = c1;
    d[2] = A.swap ? c0 : c2;
}

/* rgb24to16 / rgb24tobgr16 / rgb32to15 ... (rgb2rgb_template.c via rgbToRgbWrapper): every 8-bit channel is truncated
 * into its 5- or 6-bit field.  Two pixels per thread. */
struct Rgb16PackArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long src_fstride, dst_fstride;
    int src_stride, dst_stride;
    int w, y0, rows;
    int bpp, ro, go, bo;       /* source pixel size and channel byte offsets */
    int gbits, rgb;            /* green field width (6 / 5); 1 when red sits in the high bits */
};

__global__ void __launch_bounds__(256)
sws_rgb16pack_kernel(const __grid_constant__ Rgb16PackArgs A)
{
    const int pairs = (A.w + 1) >> 1;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = (int)(idx / pairs), i = (int)(idx - (long long)row * pairs);
    if (row >= A.rows)
        return;
    const uint8_t *s = A.src + blockIdx.z * A.src_fstride + (size_t)(A.y0 + row) * A.src_stride + (size_t)2 * i * A.bpp;
    uint16_t *d = reinterpret_cast<uint16_t *>(A.dst + blockIdx.z * A.dst_fstride + (size_t)(A.y0 + row) * A.dst_stride) + 2 * i;
    const int hi = 5 + A.gbits;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (2 * i + k >= A.w)
            break;
        const uint8_t *px = s + k * A.bpp;
        const unsigned r = px[A.ro] >> 3, g = px[A.go] >> (8 - A.gbits), b = px[A.bo] >> 3;
        d[k] = (uint16_t)(A.rgb ? (r << hi) | (g << 5) | b : (b << hi) | (g << 5) | r);
    }
}

#define SHUF_PX 1024       /* pixels of one row per block */

/* rgbToRgbWrapper / packedCopyWrapper between 8-bit packed RGB layouts (swscale_unscaled.c:2001-2060):
 * a byte permutation per pixel; alpha is carried when both sides have it, 255 when only the
 * destination has it.  Rows are staged through shared memory so that global loads and stores are
 * whole words whenever the row addresses allow it. */
template <int SBPP, int DBPP>
__global__ void __launch_bounds__(256)
sws_rgb_shuffle_kernel(const __grid_constant__ ShuffleArgs A)
{
    __shared__ __align__(16) uint32_t sm[SHUF_PX * SBPP / 4];
    uint8_t *sb = reinterpret_cast<uint8_t *>(sm);
    const int x0 = blockIdx.x * SHUF_PX;
    const int y = A.y0 + blockIdx.y;
    const int n = min(SHUF_PX, A.w - x0);
    const uint8_t *s = A.src + blockIdx.z * A.src_fstride + (size_t)y * A.src_stride + (size_t)x0 * SBPP;
    uint8_t *d = A.dst + blockIdx.z * A.dst_fstride + (size_t)y * A.dst_stride + (size_t)x0 * DBPP;
    const int nin = n * SBPP, nout = n * DBPP;

    if (((uintptr_t)s & 3) == 0) {
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
        for (int i = threadIdx.x; i < (nin >> 2); i += blockDim.x)
            sm[i] = __ldg(sw + i);
        for (int i = (nin & ~3) + threadIdx.x; i < nin; i += blockDim.x)
            sb[i] = s[i];
    } else {
        for (int i = threadIdx.x; i < nin; i += blockDim.x)
            sb[i] = s[i];
    }
    __syncthreads();

    auto out_byte = [&](int ob) -> uint32_t {
        const int px = ob / DBPP, k = ob - px * DBPP;
        const int m = A.map[k];
        return m == 4 ? 255u : sb[px * SBPP + m];
    };
    if (((uintptr_t)d & 3) == 0) {
        uint32_t *dw = reinterpret_cast<uint32_t *>(d);
        for (int i = threadIdx.x; i < (nout >> 2); i += blockDim.x) {
            const int ob = 4 * i;
            dw[i] = out_byte(ob) | out_byte(ob + 1) << 8 | out_byte(ob + 2) << 16 | out_byte(ob + 3) << 24;
        }
        for (int i = (nout & ~3) + threadIdx.x; i < nout; i += blockDim.x)
            d[i] = (uint8_t)out_byte(i);
    } else {
        for (int i = threadIdx.x; i < nout; i += blockDim.x)
            d[i] = (uint8_t)out_byte(i);
    }
}

/* The same permutation for 16-byte aligned rows: a thread owns 16 pixels (three or four 16-byte
 * loads and stores).  Per group of four pixels every destination word is gathered from at most three
 * consecutive source words with two byte permutes; the selectors only depend on the byte map and
 * are prepared on the host (ShuffleVecArgs). */
struct ShuffleVecArgs {
    const uint8_t *src;
    uint8_t *dst;
    long long src_fstride, dst_fstride;
    int src_stride, dst_stride;
    int w, y0, rows;
    int chunks;            /* 16-pixel chunks per row, the last one may be partial */
    uint32_t sel1[4], sel2[4], keep[4], fill[4];   /* per destination word of a four-pixel group */
    int map[4];
};

template <int SBPP, int DBPP>
__global__ void __launch_bounds__(256)
sws_rgb_shuffle_vec_kernel(const __grid_constant__ ShuffleVecArgs A)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = (int)(idx / A.chunks), c = (int)(idx - (long long)row * A.chunks);
    if (row >= A.rows)
        return;
    const uint8_t *s = A.src + blockIdx.z * A.src_fstride + (size_t)(A.y0 + row) * A.src_stride + (size_t)c * 16 * SBPP;
    uint8_t *d = A.dst + blockIdx.z * A.dst_fstride + (size_t)(A.y0 + row) * A.dst_stride + (size_t)c * 16 * DBPP;
    const int n = min(16, A.w - 16 * c);
    if (n < 16) {                       /* partial last chunk of a row: byte by byte */
        for (int px = 0; px < n; px++)
#pragma unroll
            for (int k = 0; k < DBPP; k++) {
                const int m = A.map[k];
                d[px * DBPP + k] = m == 4 ? 255 : s[px * SBPP + m];
            }
        return;
    }
    uint32_t w[4 * SBPP + 2], o[4 * DBPP];
#pragma unroll
    for (int i = 0; i < SBPP; i++) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(s) + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    w[4 * SBPP] = w[4 * SBPP + 1] = 0;
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int j = 0; j < DBPP; j++) {
            const int lo = g * SBPP + ((4 * j) / DBPP * SBPP) / 4;
            const uint32_t t = prmt(w[lo], w[lo + 1], A.sel1[j]);
            o[g * DBPP + j] = (prmt(t, w[lo + 2], A.sel2[j]) & A.keep[j]) | A.fill[j];
        }
#pragma unroll
    for (int i = 0; i < DBPP; i++)
        __stcs(reinterpret_cast<uint4 *>(d) + i, make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
}

struct Bgr24Yv12Args {
    const uint8_t *src;
    uint8_t *dst[3];
    long long src_fstride, dst_fstride[3];
    int src_stride, dst_stride[3];
    int cw;                /* chroma width = width >> 1 */
    int y0, h;             /* slice rows [y0, y0 + h) */
    int ry, gy, by, ru, gu, bu, rv, gv, bv;
};

/* bgr24ToYv12Wrapper -> ff_rgb24toyv12_c (rgb2rgb_template.c:580-641): per 2x2 block four luma
 * samples ((ry*r+gy*g+by*b) >> 15) + 16 and one chroma pair from the truncated box average of the
 * four pixels; an odd last row is paired with itself.  One thread per chroma sample. */
__global__ void __launch_bounds__(256)
sws_bgr24_to_yv12_kernel(const __grid_constant__ Bgr24Yv12Args A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pr = blockIdx.y;                       /* row pair inside the slice */
    if (i >= A.cw)
        return;
    const int ya = 2 * pr, yb = min(2 * pr + 1, A.h - 1);
    const uint8_t *s = A.src + blockIdx.z * A.src_fstride;
    const uint8_t *s1 = s + (size_t)(A.y0 + ya) * A.src_stride + 6 * i;
    const uint8_t *s2 = s + (size_t)(A.y0 + yb) * A.src_stride + 6 * i;
    uint8_t *d0 = A.dst[0] + blockIdx.z * A.dst_fstride[0];
    uint8_t *y1p = d0 + (size_t)(A.y0 + ya) * A.dst_stride[0] + 2 * i;
    uint8_t *y2p = d0 + (size_t)(A.y0 + yb) * A.dst_stride[0] + 2 * i;
    const unsigned b11 = s1[0], g11 = s1[1], r11 = s1[2], b12 = s1[3], g12 = s1[4], r12 = s1[5];
    const unsigned b21 = s2[0], g21 = s2[1], r21 = s2[2], b22 = s2[3], g22 = s2[4], r22 = s2[5];
    auto luma = [&](unsigned r, unsigned g, unsigned b) {
        return (uint8_t)(((A.ry * (int)r + A.gy * (int)g + A.by * (int)b) >> 15) + 16);
    };
    y1p[0] = luma(r11, g11, b11); y1p[1] = luma(r12, g12, b12);
    if (yb != ya) {                                  /* an odd last row is its own partner */
        y2p[0] = luma(r21, g21, b21); y2p[1] = luma(r22, g22, b22);
    }
    const int bx = (b11 + b12 + b21 + b22) >> 2, gx = (g11 + g12 + g21 + g22) >> 2, rx = (r11 + r12 + r21 + r22) >> 2;
    const size_t crow = (size_t)((A.y0 >> 1) + pr);
    (A.dst[1] + blockIdx.z * A.dst_fstride[1])[crow * A.dst_stride[1] + i] =
        (uint8_t)(((A.ru * rx + A.gu * gx + A.bu * bx) >> 15) + 128);
    (A.dst[2] + blockIdx.z * A.dst_fstride[2])[crow * A.dst_stride[2] + i] =
        (uint8_t)(((A.rv * rx + A.gv * gx + A.bv * bx) >> 15) + 128);
}

/* The same converter for 16-byte aligned rows: a thread owns 16 pixels of a row pair (three 16-byte
 * loads per row, one 16-byte luma store per row, 8 bytes of U and of V). */
__global__ void __launch_bounds__(256)
sws_bgr24_to_yv12_vec_kernel(const __grid_constant__ Bgr24Yv12Args A, int chunks, int pairs)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int pr = (int)(idx / chunks), c = (int)(idx - (long long)pr * chunks);
    if (pr >= pairs)
        return;
    const int ya = 2 * pr, yb = min(2 * pr + 1, A.h - 1);
    const uint8_t *s = A.src + blockIdx.z * A.src_fstride + (size_t)c * 48;
    uint32_t w[2][12];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint4 *q = reinterpret_cast<const uint4 *>(s + (size_t)(A.y0 + (r ? yb : ya)) * A.src_stride);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint4 v = __ldcs(q + i);
            w[r][4 * i] = v.x; w[r][4 * i + 1] = v.y; w[r][4 * i + 2] = v.z; w[r][4 * i + 3] = v.w;
        }
    }
    auto byte_at = [&](int r, int k) -> int { return (int)((w[r][k >> 2] >> (8 * (k & 3))) & 0xFFu); };
    uint32_t yw[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } }, uw[2] = { 0, 0 }, vw[2] = { 0, 0 };
#pragma unroll
    for (int p2 = 0; p2 < 8; p2++) {                 /* pixel pairs = chroma samples */
        int bs = 0, gs = 0, rs = 0;
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int px = 2 * p2 + q;
                const int b = byte_at(r, 3 * px), g = byte_at(r, 3 * px + 1), rr = byte_at(r, 3 * px + 2);
                bs += b; gs += g; rs += rr;
                const uint32_t yv = (uint32_t)(((A.ry * rr + A.gy * g + A.by * b) >> 15) + 16) & 0xFFu;
                yw[r][px >> 2] |= yv << (8 * (px & 3));
            }
        bs >>= 2; gs >>= 2; rs >>= 2;
        const uint32_t u = (uint32_t)(((A.ru * rs + A.gu * gs + A.bu * bs) >> 15) + 128) & 0xFFu;
        const uint32_t v = (uint32_t)(((A.rv * rs + A.gv * gs + A.bv * bs) >> 15) + 128) & 0xFFu;
        uw[p2 >> 2] |= u << (8 * (p2 & 3));
        vw[p2 >> 2] |= v << (8 * (p2 & 3));
    }
    uint8_t *d0 = A.dst[0] + blockIdx.z * A.dst_fstride[0] + (size_t)c * 16;
    __stcs(reinterpret_cast<uint4 *>(d0 + (size_t)(A.y0 + ya) * A.dst_stride[0]), make_uint4(yw[0][0], yw[0][1], yw[0][2], yw[0][3]));
    if (yb != ya)                                    /* an odd last row is its own partner */
        __stcs(reinterpret_cast<uint4 *>(d0 + (size_t)(A.y0 + yb) * A.dst_stride[0]), make_uint4(yw[1][0], yw[1][1], yw[1][2], yw[1][3]));
    const size_t crow = (size_t)((A.y0 >> 1) + pr);
    __stcs(reinterpret_cast<uint2 *>(A.dst[1] + blockIdx.z * A.dst_fstride[1] + crow * A.dst_stride[1] + (size_t)c * 8), make_uint2(uw[0], uw[1]));
    __stcs(reinterpret_cast<uint2 *>(A.dst[2] + blockIdx.z * A.dst_fstride[2] + crow * A.dst_stride[2] + (size_t)c * 8), make_uint2(vw[0], vw[1]));
}

/* 8-bit planar 4:4:4 -> packed 8-bit RGB of the same size (yuv444p / yuvj444p -> rgb24 ... abgr): 4:4:4 sources force
 * SWS_FULL_CHR_H_INT (utils.c:1278-1285), every filter is the identity and vscale.c:135 hands the rows to
 * yuv2rgb_full_1_c_template with uvalpha = 0 (output.c:2258-2290) + yuv2rgb_write_full (output.c:1998-2051):
 *   Y = l15 * 4 = y << 9,  U = (u15 - (128 << 7)) * 4 = (u - 128) << 9,
 *   Y' = (Y - y_offset) * y_coeff + 2^21,  R = Y' + V * v2r,  G = Y' + V * v2g + U * u2g,  B = Y' + U * u2b
 * (32-bit unsigned wrap-around), clip to 30 bits, >> 22.  With the constants folded on the host that is one IMAD
 * per term: R = y * ky + v * kvr + cr, ...  A thread owns 16 pixels (three 16-byte loads, 48 or 64 bytes out). */
struct Full444Args {
    const uint8_t *src[3];
    uint8_t *dst;
    long long src_fstride[3], dst_fstride;
    int src_stride[3], dst_stride;
    int w, y0, rows, chunks;
    unsigned ky, kvr, kvg, kug, kub, cr, cg, cb;
    int dst_kind;
    int vec;
};

template <int KIND>
__global__ void __launch_bounds__(256)
sws_full444_kernel(const __grid_constant__ Full444Args A)
{
    constexpr int BPP = KIND == SWSC_DST_RGB24 || KIND == SWSC_DST_BGR24 ? 3 : 4;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = (int)(idx / A.chunks), c = (int)(idx - (long long)row * A.chunks);
    if (row >= A.rows)
        return;
    const int y = A.y0 + row, f = blockIdx.z;
    const uint8_t *sy = A.src[0] + f * A.src_fstride[0] + (size_t)y * A.src_stride[0] + 16 * c;
    const uint8_t *su = A.src[1] + f * A.src_fstride[1] + (size_t)y * A.src_stride[1] + 16 * c;
    const uint8_t *sv = A.src[2] + f * A.src_fstride[2] + (size_t)y * A.src_stride[2] + 16 * c;
    uint8_t *d = A.dst + f * A.dst_fstride + (size_t)y * A.dst_stride + (size_t)16 * c * BPP;
    const int n = min(16, A.w - 16 * c);
    const bool vec = n == 16 && A.vec;
    uint32_t wy[4], wu[4], wv[4];
    if (vec) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(sy)), b = __ldcs(reinterpret_cast<const uint4 *>(su)),
                    e = __ldcs(reinterpret_cast<const uint4 *>(sv));
        wy[0] = a.x; wy[1] = a.y; wy[2] = a.z; wy[3] = a.w;
        wu[0] = b.x; wu[1] = b.y; wu[2] = b.z; wu[3] = b.w;
        wv[0] = e.x; wv[1] = e.y; wv[2] = e.z; wv[3] = e.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            wy[i] = wu[i] = wv[i] = 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < n) {
                wy[i >> 2] |= (uint32_t)sy[i] << (8 * (i & 3));
                wu[i >> 2] |= (uint32_t)su[i] << (8 * (i & 3));
                wv[i >> 2] |= (uint32_t)sv[i] << (8 * (i & 3));
            }
    }
    constexpr int kind = KIND;
    /* byte position of R, G, B inside a pixel and the constant alpha word of the 32-bit layouts */
    constexpr int pr = kind == SWSC_DST_RGB24 || kind == SWSC_DST_RGBA ? 0 : kind == SWSC_DST_BGR24 || kind == SWSC_DST_BGRA ? 2
                 : kind == SWSC_DST_ARGB ? 1 : 3;
    constexpr int pb = kind == SWSC_DST_RGB24 || kind == SWSC_DST_RGBA ? 2 : kind == SWSC_DST_BGR24 || kind == SWSC_DST_BGRA ? 0
                 : kind == SWSC_DST_ARGB ? 3 : 1;
    constexpr int pg = kind == SWSC_DST_ARGB || kind == SWSC_DST_ABGR ? 2 : 1;
    constexpr uint32_t alpha = BPP == 3 ? 0u : (kind == SWSC_DST_ARGB || kind == SWSC_DST_ABGR ? 0x000000FFu : 0xFF000000u);
    uint32_t out[4 * BPP];
#pragma unroll
    for (int g4 = 0; g4 < 4; g4++) {            /* four pixels at a time: three or four output words */
        uint32_t p4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned yv = (wy[g4] >> (8 * j)) & 0xFFu, uv = (wu[g4] >> (8 * j)) & 0xFFu, vv = (wv[g4] >> (8 * j)) & 0xFFu;
            const unsigned yy = yv * A.ky;
            const int R = (int)(yy + vv * A.kvr + A.cr);
            const int G = (int)(yy + vv * A.kvg + uv * A.kug + A.cg);
            const int B = (int)(yy + uv * A.kub + A.cb);
            const uint32_t r = (uint32_t)clip_uintp2(R, 30) >> 22, g = (uint32_t)clip_uintp2(G, 30) >> 22,
                           b = (uint32_t)clip_uintp2(B, 30) >> 22;
            p4[j] = (r << (8 * pr)) | (g << (8 * pg)) | (b << (8 * pb)) | alpha;
        }
        if (BPP == 4) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                out[4 * g4 + j] = p4[j];
        } else {                                /* 4 x 3 bytes -> 3 words */
            out[3 * g4 + 0] = p4[0] | (p4[1] << 24);
            out[3 * g4 + 1] = (p4[1] >> 8) | (p4[2] << 16);
            out[3 * g4 + 2] = (p4[2] >> 16) | (p4[3] << 8);
        }
    }
    if (vec) {
#pragma unroll
        for (int k = 0; k < BPP; k++)
            __stcs(reinterpret_cast<uint4 *>(d) + k, make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]));
    } else {            /* ragged last chunk / unaligned planes: byte stores, indices known at compile time */
#pragma unroll
        for (int k = 0; k < 4 * BPP; k++)
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (4 * k + t < n * BPP)
                    d[4 * k + t] = (uint8_t)(out[k] >> (8 * t));
    }
}

/* packed 8-bit RGB -> 8-bit planar 4:4:4 of the same size (rgb24 ... abgr -> yuv444p): the full-resolution readers
 * rgb24ToY_c / rgb24ToUV_c and the 32-bit templates (input.c:264-345,1068-1180), identity hScale16To15_c, yuv2plane1_8_c
 * with the constant dither 64.  As in sws_rgb420.cuh a pixel is one word and a matrix row two IDP.2A; with
 * S = dot + bias the chain (2 * (S >> 9) + 64) >> 7 folds to (S + 16384) >> 15 (floor of floor), the host admits only
 * matrices whose 14-bit samples stay below 16384 (no uint16 wrap, no clip at 32767).  16 pixels per thread. */
struct Rgb444Args {
    const uint8_t *src;
    uint8_t *dst[3];
    long long src_fstride, dst_fstride[3];
    int src_stride, dst_stride[3];
    int w, y0, rows, chunks;
    uint32_t ylo, yhi, ulo, uhi, vlo, vhi;
    int vec;
};

template <int BPP>
__global__ void __launch_bounds__(256)
sws_rgb444_kernel(const __grid_constant__ Rgb444Args A)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int row = (int)(idx / A.chunks), c = (int)(idx - (long long)row * A.chunks);
    if (row >= A.rows)
        return;
    const int y = A.y0 + row, f = blockIdx.z;
    const uint8_t *s = A.src + f * A.src_fstride + (size_t)y * A.src_stride + (size_t)16 * c * BPP;
    const int n = min(16, A.w - 16 * c);
    const bool vec = n == 16 && A.vec;
    uint32_t px[16];
    if (vec) {
        constexpr int NW = 4 * BPP;
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < BPP; i++) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(s) + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        if (BPP == 4) {
#pragma unroll
            for (int i = 0; i < 16; i++)
                px[i] = w[i];
        } else {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                px[4 * g] = w[3 * g];
                px[4 * g + 1] = __funnelshift_r(w[3 * g], w[3 * g + 1], 24);
                px[4 * g + 2] = __funnelshift_r(w[3 * g + 1], w[3 * g + 2], 16);
                px[4 * g + 3] = w[3 * g + 2] >> 8;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            px[i] = 0;
            if (i < n) {
#pragma unroll
                for (int k = 0; k < BPP; k++)
                    px[i] |= (uint32_t)s[i * BPP + k] << (8 * k);
            }
        }
    }
    uint32_t oy[4], ou[4], ov[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t yw = 0, uw = 0, vw = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t p = px[4 * g + j];
            const int sy = dp2a_hi_su(A.yhi, p, dp2a_lo_su(A.ylo, p, (32 << 14) + (1 << 8) + 16384));
            const int su = dp2a_hi_su(A.uhi, p, dp2a_lo_su(A.ulo, p, (256 << 14) + (1 << 8) + 16384));
            const int sv = dp2a_hi_su(A.vhi, p, dp2a_lo_su(A.vlo, p, (256 << 14) + (1 << 8) + 16384));
            yw |= (uint32_t)min(sy >> 15, 255) << (8 * j);
            uw |= (uint32_t)min(su >> 15, 255) << (8 * j);
            vw |= (uint32_t)min(sv >> 15, 255) << (8 * j);
        }
        oy[g] = yw; ou[g] = uw; ov[g] = vw;
    }
    uint8_t *d0 = A.dst[0] + f * A.dst_fstride[0] + (size_t)y * A.dst_stride[0] + 16 * c;
    uint8_t *d1 = A.dst[1] + f * A.dst_fstride[1] + (size_t)y * A.dst_stride[1] + 16 * c;
    uint8_t *d2 = A.dst[2] + f * A.dst_fstride[2] + (size_t)y * A.dst_stride[2] + 16 * c;
    if (vec) {
        __stcs(reinterpret_cast<uint4 *>(d0), make_uint4(oy[0], oy[1], oy[2], oy[3]));
        __stcs(reinterpret_cast<uint4 *>(d1), make_uint4(ou[0], ou[1], ou[2], ou[3]));
        __stcs(reinterpret_cast<uint4 *>(d2), make_uint4(ov[0], ov[1], ov[2], ov[3]));
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < n) {
                d0[i] = (uint8_t)(oy[i >> 2] >> (8 * (i & 3)));
                d1[i] = (uint8_t)(ou[i >> 2] >> (8 * (i & 3)));
                d2[i] = (uint8_t)(ov[i >> 2] >> (8 * (i & 3)));
            }
    }
}

/* 8-bit YUV -> 8-bit YUV of the same geometry with identity filters (planarToNv12Wrapper,
 * nv12ToPlanarWrapper, planar copies: swscale_unscaled.c:147-215): the scaler arithmetic collapses to
 * ((x << 7) * 4096 + (64 << 12)) >> 19 == x, so the conversion is a copy with chroma (de)interleaving.
 * 16-byte aligned planes; blockIdx.y selects the luma or the chroma half of the work. */
struct Copy8Args {
    const uint8_t *src[3];
    uint8_t *dst[3];
    long long src_fstride[3], dst_fstride[3];
    int src_stride[3], dst_stride[3];
    int w, cw, y0, rows, cy0, crows;
    int lchunks, cchunks;          /* 16-byte luma chunks / 16-sample chroma chunks per row */
    int src_layout, dst_kind;
    int vec;                       /* every plane and stride is 16-byte aligned */
};

__global__ void __launch_bounds__(256)
sws_copy8_kernel(const __grid_constant__ Copy8Args A)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int f = blockIdx.z;
    if (blockIdx.y == 0) {
        const int row = (int)(idx / A.lchunks), c = (int)(idx - (long long)row * A.lchunks);
        if (row >= A.rows)
            return;
        const uint8_t *s = A.src[0] + f * A.src_fstride[0] + (size_t)(A.y0 + row) * A.src_stride[0] + 16 * (size_t)c;
        uint8_t *d = A.dst[0] + f * A.dst_fstride[0] + (size_t)(A.y0 + row) * A.dst_stride[0] + 16 * (size_t)c;
        const int n = min(16, A.w - 16 * c);
        if (n == 16 && A.vec)
            __stcs(reinterpret_cast<uint4 *>(d), __ldcs(reinterpret_cast<const uint4 *>(s)));
        else
            for (int i = 0; i < n; i++)
                d[i] = s[i];
        return;
    }
    const int row = (int)(idx / A.cchunks), c = (int)(idx - (long long)row * A.cchunks);
    if (row >= A.crows)
        return;
    const int y = A.cy0 + row, n = min(16, A.cw - 16 * c);
    const bool splanar = A.src_layout == SWSC_SRC_PLANAR, dplanar = A.dst_kind == SWSC_DST_PLANAR8;
    const int sswap = A.src_layout == SWSC_SRC_NV21, dswap = A.dst_kind == SWSC_DST_NV21;
    const uint8_t *s1 = A.src[1] + f * A.src_fstride[1] + (size_t)y * A.src_stride[1];
    const uint8_t *s2 = splanar ? A.src[2] + f * A.src_fstride[2] + (size_t)y * A.src_stride[2] : nullptr;
    uint8_t *d1 = A.dst[1] + f * A.dst_fstride[1] + (size_t)y * A.dst_stride[1];
    uint8_t *d2 = dplanar ? A.dst[2] + f * A.dst_fstride[2] + (size_t)y * A.dst_stride[2] : nullptr;
    if (n < 16 || !A.vec) {             /* partial last chunk of a chroma row, or unaligned planes */
        for (int i = 0; i < n; i++) {
            const int x = 16 * c + i;
            const uint8_t u = splanar ? s1[x] : s1[2 * x + sswap], v = splanar ? s2[x] : s1[2 * x + 1 - sswap];
            if (dplanar) {
                d1[x] = u; d2[x] = v;
            } else {
                d1[2 * x + dswap] = u; d1[2 * x + 1 - dswap] = v;
            }
        }
        return;
    }
    uint4 u, v;
    if (splanar) {
        u = __ldcs(reinterpret_cast<const uint4 *>(s1) + c);
        v = __ldcs(reinterpret_cast<const uint4 *>(s2) + c);
    } else {
        const uint4 a = __ldcs(reinterpret_cast<const uint4 *>(s1) + 2 * c), b = __ldcs(reinterpret_cast<const uint4 *>(s1) + 2 * c + 1);
        const uint4 e = make_uint4(prmt(a.x, a.y, 0x6420), prmt(a.z, a.w, 0x6420), prmt(b.x, b.y, 0x6420), prmt(b.z, b.w, 0x6420));
        const uint4 o = make_uint4(prmt(a.x, a.y, 0x7531), prmt(a.z, a.w, 0x7531), prmt(b.x, b.y, 0x7531), prmt(b.z, b.w, 0x7531));
        u = sswap ? o : e;
        v = sswap ? e : o;
    }
    if (dplanar) {
        __stcs(reinterpret_cast<uint4 *>(d1) + c, u);
        __stcs(reinterpret_cast<uint4 *>(d2) + c, v);
    } else {
        const uint4 e = dswap ? v : u, o = dswap ? u : v;
        __stcs(reinterpret_cast<uint4 *>(d1) + 2 * c, make_uint4(prmt(e.x, o.x, 0x5140), prmt(e.x, o.x, 0x7362), prmt(e.y, o.y, 0x5140), prmt(e.y, o.y, 0x7362)));
        __stcs(reinterpret_cast<uint4 *>(d1) + 2 * c + 1, make_uint4(prmt(e.z, o.z, 0x5140), prmt(e.z, o.z, 0x7362), prmt(e.w, o.w, 0x5140), prmt(e.w, o.w, 0x7362)));
    }
}

/* planarCopyWrapper between planar YUV depths (swscale_unscaled.c:2160-2218,2249-2346), little-endian:
 *   down, dither none:        t = (v + (1 << (shift - 1))) >> shift;  out = t - (t >> dst_depth)
 *   down, chroma / MPEG luma: t = (v + d) >> shift;                   out = t - (t >> dst_depth)
 *   down, full-range luma:    out = (v - (v >> dst_depth) + d) >> shift
 *   up,   chroma / MPEG luma: out = v << shift
 *   up,   full-range luma:    out = (v << shift) | (v >> (2 * src_depth - dst_depth))
 * d = ordered-dither matrix of level `shift` (1..8), row = plane row & 7, column = x & 7.  The matrices are
 * periodic tiles (2x2 for levels 1-2, 4x4 for 3-4, 8x8 above; level 7 repeats level 6, level 8 is
 * ff_dither_8x8_128); the host expands them once into constant memory. */
__constant__ __align__(8) uint8_t c_depth_dither[8][8][8];

static const uint8_t depth_tile_1[2][2] = { { 0, 1 }, { 1, 0 } };
static const uint8_t depth_tile_2[2][2] = { { 1, 2 }, { 3, 0 } };
static const uint8_t depth_tile_3[4][4] = { { 2, 4, 3, 5 }, { 6, 0, 7, 1 }, { 3, 5, 2, 4 }, { 7, 1, 6, 0 } };
static const uint8_t depth_tile_4[4][4] = { { 4, 8, 7, 11 }, { 12, 0, 15, 3 }, { 6, 10, 5, 9 }, { 14, 2, 13, 1 } };
static const uint8_t depth_tile_5[8][8] = {
    { 9, 17, 15, 23, 8, 16, 14, 22 }, { 25, 1, 31, 7, 24, 0, 30, 6 }, { 13, 21, 11, 19, 12, 20, 10, 18 },
    { 29, 5, 27, 3, 28, 4, 26, 2 },   { 8, 16, 14, 22, 9, 17, 15, 23 }, { 24, 0, 30, 6, 25, 1, 31, 7 },
    { 12, 20, 10, 18, 13, 21, 11, 19 }, { 28, 4, 26, 2, 29, 5, 27, 3 } };
static const uint8_t depth_tile_6[8][8] = {
    { 18, 34, 30, 46, 17, 33, 29, 45 }, { 50, 2, 62, 14, 49, 1, 61, 13 }, { 26, 42, 22, 38, 25, 41, 21, 37 },
    { 58, 10, 54, 6, 57, 9, 53, 5 },    { 16, 32, 28, 44, 19, 35, 31, 47 }, { 48, 0, 60, 12, 51, 3, 63, 15 },
    { 24, 40, 20, 36, 27, 43, 23, 39 }, { 56, 8, 52, 4, 59, 11, 55, 7 } };
static const uint8_t depth_tile_8[8][8] = {            /* == ff_dither_8x8_128 */
    { 36, 68, 60, 92, 34, 66, 58, 90 },   { 100, 4, 124, 28, 98, 2, 122, 26 }, { 52, 84, 44, 76, 50, 82, 42, 74 },
    { 116, 20, 108, 12, 114, 18, 106, 10 }, { 32, 64, 56, 88, 38, 70, 62, 94 }, { 96, 0, 120, 24, 102, 6, 126, 30 },
    { 48, 80, 40, 72, 54, 86, 46, 78 },   { 112, 16, 104, 8, 118, 22, 110, 14 } };

static cudaError_t upload_depth_dither(void)
{
    uint8_t t[8][8][8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            t[0][y][x] = depth_tile_1[y & 1][x & 1];
            t[1][y][x] = depth_tile_2[y & 1][x & 1];
            t[2][y][x] = depth_tile_3[y & 3][x & 3];
            t[3][y][x] = depth_tile_4[y & 3][x & 3];
            t[4][y][x] = depth_tile_5[y][x];
            t[5][y][x] = t[6][y][x] = depth_tile_6[y][x];
            t[7][y][x] = depth_tile_8[y][x];
        }
    return cudaMemcpyToSymbol(c_depth_dither, t, sizeof(t));
}

/* planarToP01xWrapper / planar8ToP01xleWrapper (swscale_unscaled.c:273-375): planar 4:2:0 -> p010le, every
 * sample shifted left into the 16-bit container (8-bit sources by 8, N-bit sources by 16 - N), chroma
 * interleaved U first.  blockIdx.y: 0 = luma, 1 = chroma; a thread owns 8 luma samples or 8 chroma pairs. */
struct P01xArgs {
    const uint8_t *src[3];
    uint8_t *dst[2];
    long long src_fstride[3], dst_fstride[2];
    int src_stride[3], dst_stride[2];
    int w, cw, y0, rows, cy0, crows;
    int shift;
    int vec;                       /* every plane and stride allows 16-byte accesses */
};

/* eight samples of a row as 32-bit values (8- or 16-bit containers; one 8- or 16-byte load when `vec`) */
template <typename SrcT>
__device__ __forceinline__ void load8(const SrcT *s, int n, bool vec, unsigned (&v)[8])
{
    if (vec && n == 8) {
        if (sizeof(SrcT) == 1) {
            const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(s));
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] = ((i < 4 ? q.x : q.y) >> (8 * (i & 3))) & 0xFFu;
        } else {
            const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(s));
            const unsigned w4[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] = (w4[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            v[i] = i < n ? s[i] : 0;
    }
}

template <typename SrcT>
__global__ void __launch_bounds__(256)
sws_p01x_kernel(const __grid_constant__ P01xArgs A)
{
    /* one index space: the luma chunks of all rows, then the chroma chunks (8 samples / 8 pairs per thread) */
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int f = blockIdx.z;
    const int lchunks = (A.w + 7) / 8, cchunks = (A.cw + 7) / 8;
    const long long lwork = (long long)lchunks * A.rows;
    if (idx < lwork) {
        const int row = (int)(idx / lchunks), c = (int)(idx - (long long)row * lchunks);
        const SrcT *s = reinterpret_cast<const SrcT *>(A.src[0] + f * A.src_fstride[0] + (size_t)(A.y0 + row) * A.src_stride[0]) + 8 * c;
        uint16_t *d = reinterpret_cast<uint16_t *>(A.dst[0] + f * A.dst_fstride[0] + (size_t)(A.y0 + row) * A.dst_stride[0]) + 8 * c;
        const int n = min(8, A.w - 8 * c);
        unsigned v[8];
        load8<SrcT>(s, n, A.vec, v);
        if (A.vec && n == 8) {
            __stcs(reinterpret_cast<uint4 *>(d), make_uint4(((v[0] << A.shift) & 0xFFFFu) | (v[1] << (A.shift + 16)),
                                                            ((v[2] << A.shift) & 0xFFFFu) | (v[3] << (A.shift + 16)),
                                                            ((v[4] << A.shift) & 0xFFFFu) | (v[5] << (A.shift + 16)),
                                                            ((v[6] << A.shift) & 0xFFFFu) | (v[7] << (A.shift + 16))));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < n)
                    d[i] = (uint16_t)(v[i] << A.shift);
        }
        return;
    }
    const long long cidx = idx - lwork;
    const int row = (int)(cidx / cchunks), c = (int)(cidx - (long long)row * cchunks);
    if (row >= A.crows)
        return;
    const int y = A.cy0 + row;
    const SrcT *su = reinterpret_cast<const SrcT *>(A.src[1] + f * A.src_fstride[1] + (size_t)y * A.src_stride[1]) + 8 * c;
    const SrcT *sv = reinterpret_cast<const SrcT *>(A.src[2] + f * A.src_fstride[2] + (size_t)y * A.src_stride[2]) + 8 * c;
    uint16_t *d = reinterpret_cast<uint16_t *>(A.dst[1] + f * A.dst_fstride[1] + (size_t)y * A.dst_stride[1]) + 16 * c;
    const int n = min(8, A.cw - 8 * c);
    unsigned u[8], v[8];
    load8<SrcT>(su, n, A.vec, u);
    load8<SrcT>(sv, n, A.vec, v);
    if (A.vec && n == 8) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            w[i] = ((u[i] << A.shift) & 0xFFFFu) | (v[i] << (A.shift + 16));
        __stcs(reinterpret_cast<uint4 *>(d), make_uint4(w[0], w[1], w[2], w[3]));
        __stcs(reinterpret_cast<uint4 *>(d) + 1, make_uint4(w[4], w[5], w[6], w[7]));
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < n) {
                d[2 * i] = (uint16_t)(u[i] << A.shift);
                d[2 * i + 1] = (uint16_t)(v[i] << A.shift);
            }
    }
}

struct DepthCopyArgs {
    const uint8_t *src[3];
    uint8_t *dst[3];
    long long src_fstride[3], dst_fstride[3];
    int src_stride[3], dst_stride[3];
    int w[3], y0[3], rows[3];      /* per plane: width, first row, row count of this launch */
    int chunks[3];                 /* 8-sample chunks per row */
    int src_depth, dst_depth;
    int src_shift, dst_shift;      /* position of the samples inside 16-bit containers (p010: 6) */
    int luma_shiftonly;            /* limited-range source: luma is shifted like chroma */
    int dither_none;
    int vec;                       /* planes and strides allow 8/16-byte accesses */
    int nplanes;
};

template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(256, 8)       /* 32 registers: full occupancy is what the streaming path lives on */
sws_depthcopy_kernel(const __grid_constant__ DepthCopyArgs A)
{
    /* one index space over the planes: no block is launched only to find its plane already done */
    long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    int plane = 0;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        const long long work = (long long)A.chunks[plane] * A.rows[plane];
        if (plane == pl && idx >= work) {
            idx -= work;
            plane = pl + 1;
        }
    }
    if (plane >= A.nplanes)
        return;
    /* 10..16 -> 8 bit moves 16 samples per thread (two 16-byte loads, one 16-byte store); the others 8 */
    constexpr int CH = (sizeof(SrcT) == 2 && sizeof(DstT) == 1) ? 16 : 8;
    const unsigned i32 = (unsigned)idx, nch = (unsigned)A.chunks[plane];       /* a plane has < 2^31 chunks */
    const int row = (int)(i32 / nch), c = (int)(i32 - (unsigned)row * nch);
    if (row >= A.rows[plane])
        return;
    const int y = A.y0[plane] + row;
    const SrcT *s = reinterpret_cast<const SrcT *>(A.src[plane] + blockIdx.z * A.src_fstride[plane] + (size_t)y * A.src_stride[plane]) + CH * c;
    DstT *d = reinterpret_cast<DstT *>(A.dst[plane] + blockIdx.z * A.dst_fstride[plane] + (size_t)y * A.dst_stride[plane]) + CH * c;
    const int n = min(CH, A.w[plane] - CH * c);
    const bool shiftonly = plane != 0 || A.luma_shiftonly;
    const int sd = A.src_depth, dd = A.dst_depth;
    if (sizeof(SrcT) == 2 && sd > dd && sd <= 15 && A.vec && n == CH && A.src_shift == 0) {
        /* two samples per 32-bit word: with at most 15 source bits v + dither cannot carry into the upper half,
         * and t - (t >> dst_depth) cannot borrow.  Same numbers as the scalar path below. */
        const int shift = sd - dd;
        uint32_t w[CH / 2];
#pragma unroll
        for (int g = 0; g < CH / 8; g++) {
            const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(s) + g);
            w[4 * g] = q.x; w[4 * g + 1] = q.y; w[4 * g + 2] = q.z; w[4 * g + 3] = q.w;
        }
        const uint2 dq = *reinterpret_cast<const uint2 *>(c_depth_dither[shift - 1][row & 7]);
        const uint32_t half = 0x00010001u;
        const uint32_t hm = (0xFFFFu >> shift) * half;
        uint32_t r[CH / 2];
#pragma unroll
        for (int k = 0; k < CH / 2; k++) {
            const uint32_t dw = (k & 3) < 2 ? dq.x : dq.y;
            const uint32_t d2 = A.dither_none ? (1u << (shift - 1)) * half : prmt(dw, 0u, (k & 1) ? 0x4342 : 0x4140);
            if (A.dither_none || shiftonly) {
                const uint32_t t = ((w[k] + d2) >> shift) & hm;
                r[k] = t - ((t >> dd) & half);
            } else {
                const uint32_t x = w[k] - ((w[k] >> dd) & ((0xFFFFu >> dd) * half));
                r[k] = ((x + d2) >> shift) & hm;
            }
        }
        if (sizeof(DstT) == 1)
            __stcs(reinterpret_cast<uint4 *>(d), make_uint4(prmt(r[0], r[1], 0x6420), prmt(r[2], r[3], 0x6420),
                                                            prmt(r[CH / 2 - 4], r[CH / 2 - 3], 0x6420), prmt(r[CH / 2 - 2], r[CH / 2 - 1], 0x6420)));
        else
            __stcs(reinterpret_cast<uint4 *>(d), make_uint4(r[0], r[1], r[2], r[3]));
        return;
    }
#pragma unroll
    for (int g = 0; g < CH / 8; g++) {
        const int ng = min(8, n - 8 * g);
        if (ng <= 0)
            break;
        const SrcT *sg = s + 8 * g;
        DstT *dg = d + 8 * g;
        unsigned v[8], o[8];
        load8<SrcT>(sg, ng, A.vec, v);
        if (sd > dd && A.src_shift) {
            /* DITHER_COPY's scalar tail (the last width & 7 samples of a row, swscale_unscaled.c:2174-2176,2193-2195,
             * 2212-2214) forgets the source shift: p010 samples go in with their six low bits, and the 8-bit store
             * keeps the low byte of what comes out */
            const int j0 = CH * c + 8 * g, body = A.w[plane] & ~7;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (j0 + i < body)
                    v[i] >>= A.src_shift;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] >>= A.src_shift;
        }
        if (sd > dd) {
            const int shift = sd - dd;
            const uint2 dq = *reinterpret_cast<const uint2 *>(c_depth_dither[shift - 1][row & 7]);   /* rows count from the slice */
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const unsigned dz = ((i < 4 ? dq.x : dq.y) >> (8 * (i & 3))) & 0xFFu;
                if (A.dither_none) {
                    const unsigned t = (v[i] + (1u << (shift - 1))) >> shift;
                    o[i] = t - (t >> dd);
                } else if (shiftonly) {
                    const unsigned t = (v[i] + dz) >> shift;
                    o[i] = t - (t >> dd);
                } else {
                    o[i] = (v[i] - (v[i] >> dd) + dz) >> shift;
                }
            }
        } else {
            const int shift = dd - sd, rep = 2 * sd - dd;
#pragma unroll
            for (int i = 0; i < 8; i++)
                o[i] = (shiftonly ? v[i] << shift : (v[i] << shift) | (v[i] >> rep)) << A.dst_shift;
            if (shiftonly && A.src_shift && sizeof(SrcT) == 2) {
                /* p010 -> p010: the reference shifts four samples at a time as one 64-bit word
                 * (swscale_unscaled.c:2291-2296), so only the first of each full group of four loses the
                 * bits below its sample; the scalar tail clears them for every sample */
                const int j0 = CH * c + 8 * g, len = A.w[plane];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (i < ng && (i & 3) && ((j0 + i) & ~3) + 3 < len)
                        o[i] = (unsigned)sg[i];
            }
        }
        if (ng == 8 && A.vec) {
            if (sizeof(DstT) == 1) {
                __stcs(reinterpret_cast<uint2 *>(dg), make_uint2((o[0] & 0xFF) | (o[1] & 0xFF) << 8 | (o[2] & 0xFF) << 16 | o[3] << 24,
                                                                 (o[4] & 0xFF) | (o[5] & 0xFF) << 8 | (o[6] & 0xFF) << 16 | o[7] << 24));
            } else {
                __stcs(reinterpret_cast<uint4 *>(dg), make_uint4((o[0] & 0xFFFF) | o[1] << 16, (o[2] & 0xFFFF) | o[3] << 16,
                                                                 (o[4] & 0xFFFF) | o[5] << 16, (o[6] & 0xFFFF) | o[7] << 16));
            }
        } else {
            for (int i = 0; i < ng; i++)
                dg[i] = (DstT)o[i];
        }
    }
}

/* ------------------------------------------------------------------------
 * Generic fused tile kernel: table-driven H FIR -> (range) -> V FIR -> pack.
 *   SRC16   : source samples are 16-bit containers (9..16 bit depths)
 *   INTER32 : h-scaled lines are 19-bit int32 (dst depth > 14) else 15-bit int16
 * grid = (tiles_x, tiles_y, frames), block = 256 threads, dynamic smem.
 * ------------------------------------------------------------------------ */
template <bool SRC16, bool INTER32>
__global__ void __launch_bounds__(256)
sws_generic_tile_kernel(const __grid_constant__ SwsCudaPlan P, const __grid_constant__ FrameArgs A)
{
    typedef typename std::conditional<INTER32, int32_t, int16_t>::type inter_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int f = blockIdx.z;
    const uint8_t *src0 = A.src[0] + f * A.src_fstride[0];
    const uint8_t *src1 = A.src[1] ? A.src[1] + f * A.src_fstride[1] : nullptr;
    const uint8_t *src2 = A.src[2] ? A.src[2] + f * A.src_fstride[2] : nullptr;
    uint8_t *dst0 = A.dst[0] + f * A.dst_fstride[0];
    uint8_t *dst1 = A.dst[1] ? A.dst[1] + f * A.dst_fstride[1] : nullptr;
    uint8_t *dst2 = A.dst[2] ? A.dst[2] + f * A.dst_fstride[2] : nullptr;

    const int TW = A.tile_w, TH = A.tile_h;
    const int x0 = blockIdx.x * TW;
    const int ry0 = A.y0 + blockIdx.y * TH;
    const int ry1 = min(ry0 + TH, A.y1);
    const int tw = min(TW, P.dst_w - x0);            /* luma columns in this tile */
    const int th = ry1 - ry0;

    /* chroma output window of the tile */
    const int hs = P.chr_dst_hsub, vs = P.chr_dst_vsub;
    const int CW = TW >> hs;
    const int cx0 = x0 >> hs;
    const int cw = min(CW, P.chr_dst_w - cx0);
    const int cy0 = ry0 >> vs;
    const int cy1 = (ry1 == P.dst_h) ? P.chr_dst_h : (ry1 >> vs);
    const int ch = cy1 - cy0;

    inter_t *hb_l = reinterpret_cast<inter_t *>(smem_raw);
    inter_t *hb_u = hb_l + (size_t)A.rows_l_cap * TW;
    inter_t *hb_v = hb_u + (size_t)A.rows_c_cap * CW;
    inter_t *hb_a = hb_v + (size_t)A.rows_c_cap * CW;      /* alpha lines (P.src_alpha), laid out like luma */

    /* source row windows (positions are monotonic; take min/max defensively) */
    int lo_l = INT_MAX, hi_l = 0, lo_c = INT_MAX, hi_c = 0;
    for (int y = ry0; y < ry1; y++) {
        int p = P.vl_pos[y];
        lo_l = min(lo_l, p);
        hi_l = max(hi_l, p + P.vl_size);
    }
    for (int y = cy0; y < cy1; y++) {
        int p = P.vc_pos[y];
        lo_c = min(lo_c, p);
        hi_c = max(hi_c, p + P.vc_size);
    }
    lo_l = max(lo_l, 0); lo_c = max(lo_c, 0);
    const int nl = min(hi_l - lo_l, A.rows_l_cap);
    const int nc = (ch > 0) ? min(hi_c - lo_c, A.rows_c_cap) : 0;

    const int h_max = INTER32 ? (1 << 19) - 1 : (1 << 15) - 1;
    const int sh = P.h_shift;

    /* ---- stage H, luma ---- */
    {
        const int fs = P.hl_size;
        for (int idx = threadIdx.x; idx < nl * TW; idx += blockDim.x) {
            const int r = idx / TW, x = idx - r * TW;
            if (x >= tw)
                continue;
            const int sy = min(lo_l + r, P.src_h - 1);
            const uint8_t *row = src0 + (size_t)sy * A.src_stride[0];
            const int gx = x0 + x;
            const int pos = P.hl_pos[gx];
            const int16_t *co = P.hl_coef + (size_t)gx * fs;
            int val = 0;
            if (P.src_layout == SWSC_SRC_RGB) {
                for (int j = 0; j < fs; j++)
                    val += rgb_luma14(P, row, min(pos + j, P.src_w - 1)) * (int)co[j];
            } else {
                for (int j = 0; j < fs; j++)
                    val += fetch<SRC16>(row, min(pos + j, P.src_w - 1), P.src_shift) * (int)co[j];
            }
            val = min(val >> sh, h_max);
            if (P.range_mode) {
                if (!INTER32) {
                    val = (val * (int)P.lum_rc_coeff + (int)P.lum_rc_offset) >> 14;
                    if (P.range_mode == 1)
                        val = min(val, (1 << 15) - 1);
                    val = (int16_t)val;
                } else {
                    val = (int)(((long long)val * P.lum_rc_coeff + P.lum_rc_offset) >> 18);
                    if (P.range_mode == 1)
                        val = min(val, (1 << 19) - 1);
                }
            }
            hb_l[idx] = (inter_t)val;
        }
    }
    /* ---- stage H, alpha: rgbaToA_c / abgrToA_c (input.c:455-471) through the luma bank; no range conversion ---- */
    if (P.src_alpha) {
        const int fs = P.hl_size;
        const int ao = P.src_ao;
        for (int idx = threadIdx.x; idx < nl * TW; idx += blockDim.x) {
            const int r = idx / TW, x = idx - r * TW;
            if (x >= tw)
                continue;
            const int sy = min(lo_l + r, P.src_h - 1);
            const uint8_t *row = src0 + (size_t)sy * A.src_stride[0] + ao;
            const int gx = x0 + x;
            const int pos = P.hl_pos[gx];
            const int16_t *co = P.hl_coef + (size_t)gx * fs;
            int val = 0;
            for (int j = 0; j < fs; j++) {
                const int a8 = row[4 * min(pos + j, P.src_w - 1)];
                val += ((a8 << 6) | (a8 >> 2)) * (int)co[j];
            }
            hb_a[idx] = (inter_t)min(val >> sh, h_max);
        }
    }
    /* ---- stage H, chroma (input unpack of nv12/nv21 fused into the fetch) ---- */
    if (P.has_chroma && ch > 0) {
        const int fs = P.hc_size;
        const int layout = P.src_layout;
        for (int idx = threadIdx.x; idx < nc * CW; idx += blockDim.x) {
            const int r = idx / CW, x = idx - r * CW;
            if (x >= cw)
                continue;
            const int sy = min(lo_c + r, P.chr_src_h - 1);
            const int gx = cx0 + x;
            const int pos = P.hc_pos[gx];
            const int16_t *co = P.hc_coef + (size_t)gx * fs;
            int u = 0, v = 0;
            if (layout == SWSC_SRC_PLANAR) {
                const uint8_t *ru = src1 + (size_t)sy * A.src_stride[1];
                const uint8_t *rv = src2 + (size_t)sy * A.src_stride[2];
                for (int j = 0; j < fs; j++) {
                    const int sx = min(pos + j, P.chr_src_w - 1);
                    const int cj = co[j];
                    u += fetch<SRC16>(ru, sx) * cj;
                    v += fetch<SRC16>(rv, sx) * cj;
                }
            } else if (layout == SWSC_SRC_RGB) {
                const uint8_t *row = src0 + (size_t)sy * A.src_stride[0];
                for (int j = 0; j < fs; j++) {
                    int uu, vv;
                    rgb_chroma14(P, row, min(pos + j, P.chr_src_w - 1), uu, vv);
                    u += uu * (int)co[j];
                    v += vv * (int)co[j];
                }
            } else {
                const uint8_t *ruv = src1 + (size_t)sy * A.src_stride[1];
                const int uo = layout == SWSC_SRC_NV12 ? 0 : 1;
                for (int j = 0; j < fs; j++) {
                    const int sx = min(pos + j, P.chr_src_w - 1);
                    const int cj = co[j];
                    u += fetch<SRC16>(ruv, 2 * sx + uo, P.src_shift) * cj;
                    v += fetch<SRC16>(ruv, 2 * sx + 1 - uo, P.src_shift) * cj;
                }
            }
            u = min(u >> sh, h_max);
            v = min(v >> sh, h_max);
            if (P.range_mode) {
                if (!INTER32) {
                    u = (u * (int)P.chr_rc_coeff + (int)P.chr_rc_offset) >> 14;
                    v = (v * (int)P.chr_rc_coeff + (int)P.chr_rc_offset) >> 14;
                    if (P.range_mode == 1) {
                        u = min(u, (1 << 15) - 1);
                        v = min(v, (1 << 15) - 1);
                    }
                    u = (int16_t)u; v = (int16_t)v;
                } else {
                    u = (int)(((long long)u * P.chr_rc_coeff + P.chr_rc_offset) >> 18);
                    v = (int)(((long long)v * P.chr_rc_coeff + P.chr_rc_offset) >> 18);
                    if (P.range_mode == 1) {
                        u = min(u, (1 << 19) - 1);
                        v = min(v, (1 << 19) - 1);
                    }
                }
            }
            hb_u[idx] = (inter_t)u;
            hb_v[idx] = (inter_t)v;
        }
    }
    __syncthreads();

    const int kind = P.dst_kind;
    const int lfs = P.vl_size, cfs = P.vc_size;

    /* ---- stage V + pack ---- */
    if (kind == SWSC_DST_GBRP) {
        /* planar RGB, 32-bit float: yuv2gbrpf32_full_X_c (output.c:2536-2610) -- always the X writer with the real
         * taps, 32-bit wrap-around in Y + R, the 16-bit result times 1.0f / 65535.0f (one IEEE multiply: bit-exact) */
        const float float_mult = 1.0f / 65535.0f;
        for (int idx = threadIdx.x; idx < th * TW; idx += blockDim.x) {
            const int ty = idx / TW, x = idx - ty * TW;
            if (x >= tw)
                continue;
            const int y = ry0 + ty;
            const int16_t *lf = P.vl_coef + (size_t)y * lfs;
            const int16_t *cf = P.vc_coef + (size_t)y * cfs;
            const int rl = max(P.vl_pos[y], 0) - lo_l;
            const int rc = max(P.vc_pos[y], 0) - lo_c;
            const inter_t *pl = hb_l + (size_t)rl * TW + x;
            const inter_t *pu = hb_u + (size_t)rc * CW + x;
            const inter_t *pv = hb_v + (size_t)rc * CW + x;
            unsigned Yv = 0u - 0x40000000u, U = 0u - (128u << 23), V = 0u - (128u << 23);
            for (int j = 0; j < lfs; j++)
                Yv += (unsigned)(int)pl[(size_t)(min(rl + j, nl - 1) - rl) * TW] * (unsigned)(int)lf[j];
            for (int j = 0; j < cfs; j++) {
                const int r = min(rc + j, nc - 1) - rc;
                const unsigned c = (unsigned)(int)cf[j];
                U += (unsigned)(int)pu[(size_t)r * CW] * c;
                V += (unsigned)(int)pv[(size_t)r * CW] * c;
            }
            int Yi = ((int)Yv >> 14) + 0x10000;
            const int Ui = (int)U >> 14, Vi = (int)V >> 14;
            Yi -= P.rgb.y_offset;
            const unsigned Yu = (unsigned)Yi * (unsigned)P.rgb.y_coeff + (1u << 13) - (1u << 29);
            const unsigned R = (unsigned)Vi * (unsigned)P.rgb.v2r;
            const unsigned G = (unsigned)Vi * (unsigned)P.rgb.v2g + (unsigned)Ui * (unsigned)P.rgb.u2g;
            const unsigned B = (unsigned)Ui * (unsigned)P.rgb.u2b;
            const int r = clip_uintp2(((int)(Yu + R) >> 14) + (1 << 15), 16);
            const int g = clip_uintp2(((int)(Yu + G) >> 14) + (1 << 15), 16);
            const int b = clip_uintp2(((int)(Yu + B) >> 14) + (1 << 15), 16);
            const int gx = x0 + x;
            reinterpret_cast<float *>(dst0 + (size_t)y * A.dst_stride[0])[gx] = __fmul_rn(float_mult, (float)g);
            reinterpret_cast<float *>(dst1 + (size_t)y * A.dst_stride[1])[gx] = __fmul_rn(float_mult, (float)b);
            reinterpret_cast<float *>(dst2 + (size_t)y * A.dst_stride[2])[gx] = __fmul_rn(float_mult, (float)r);
        }
    } else if (kind >= SWSC_DST_RGB24 && kind <= SWSC_DST_BGRA64 && P.full_chr) {
        /* packed RGB with full horizontal chroma (odd width / 4:4:4 source): every pixel has its own
         * U,V and the colour step is arithmetic, not LUT based: yuv2rgb_full_{X,1,2}_c_template +
         * yuv2rgb_write_full (output.c:1998-2051,2160-2330), yuv2rgba64_full_X_c_template (:1373-1430) */
        const bool is16 = kind == SWSC_DST_RGB48 || kind == SWSC_DST_BGR48;
        for (int idx = threadIdx.x; idx < th * TW; idx += blockDim.x) {
            const int ty = idx / TW, x = idx - ty * TW;
            if (x >= tw)
                continue;
            const int y = ry0 + ty;
            const int16_t *lf = P.vl_coef + (size_t)y * lfs;
            const int16_t *cf = P.vc_coef + (size_t)y * cfs;
            const int rl = max(P.vl_pos[y], 0) - lo_l;
            const int rc = max(P.vc_pos[y], 0) - lo_c;
            const inter_t *pl = hb_l + (size_t)rl * TW + x;
            const inter_t *pu = hb_u + (size_t)rc * CW + x;
            const inter_t *pv = hb_v + (size_t)rc * CW + x;
            unsigned Yv = 0, U = 0, V = 0;
            for (int j = 0; j < lfs; j++)
                Yv += (unsigned)(int)pl[(size_t)(min(rl + j, nl - 1) - rl) * TW] * (unsigned)(int)lf[j];
            for (int j = 0; j < cfs; j++) {
                const int r = min(rc + j, nc - 1) - rc;
                const unsigned c = (unsigned)(int)cf[j];
                U += (unsigned)(int)pu[(size_t)r * CW] * c;
                V += (unsigned)(int)pv[(size_t)r * CW] * c;
            }
            const int gx = x0 + x;
            uint8_t *d = dst0 + (size_t)y * A.dst_stride[0];
            if (!is16) {
                /* rounding bias 1<<9, dropped by the 2-tap fast paths exactly as vscale.c:135-163 picks them */
                unsigned lb = 1u << 9, cb = 1u << 9;
                if (cfs == 2) {
                    const int c0 = cf[0], c1 = cf[1];
                    const bool cok = c0 + c1 == 4096 && (unsigned)c1 <= 4096u;
                    if (lfs == 1 && cok)
                        cb = 0;
                    else if (lfs == 2 && cok && lf[0] + lf[1] == 4096 && (unsigned)(int)lf[1] <= 4096u)
                        lb = cb = 0;
                }
                int Yi = (int)(Yv + lb) >> 10;
                const int Ui = (int)(U + cb - (128u << 19)) >> 10;
                const int Vi = (int)(V + cb - (128u << 19)) >> 10;
                unsigned Yu = (unsigned)(Yi - P.rgb.y_offset) * (unsigned)P.rgb.y_coeff + (1u << 21);
                int R = (int)(Yu + (unsigned)Vi * (unsigned)P.rgb.v2r);
                int G = (int)(Yu + (unsigned)Vi * (unsigned)P.rgb.v2g + (unsigned)Ui * (unsigned)P.rgb.u2g);
                int B = (int)(Yu + (unsigned)Ui * (unsigned)P.rgb.u2b);
                R = clip_uintp2(R, 30) >> 22; G = clip_uintp2(G, 30) >> 22; B = clip_uintp2(B, 30) >> 22;
                int Av = 255;
                if (P.src_alpha) {
                    /* yuv2rgb_full_{X,2,1}_c_template (output.c:2193-2201,2241-2245,2278-2282): the same three
                     * writers vscale.c:135-169 picks for Y, U, V */
                    const inter_t *pa = hb_a + (size_t)rl * TW + x;
                    const bool chr2 = cfs == 2 && cf[0] + cf[1] == 4096 && (unsigned)(int)cf[1] <= 4096u;
                    if (lfs == 1 && (cfs == 1 || chr2)) {
                        Av = ((int)pa[0] + 64) >> 7;
                    } else {
                        unsigned acc = 1u << 18;       /* _2 and _X both round */
                        for (int j = 0; j < lfs; j++)
                            acc += (unsigned)(int)pa[(size_t)(min(rl + j, nl - 1) - rl) * TW] * (unsigned)(int)lf[j];
                        Av = (int)acc >> 19;
                    }
                    if (Av & 0x100)
                        Av = clip_u8(Av);
                    Av &= 0xFF;
                }
                switch (kind) {
                case SWSC_DST_RGB24: d += 3 * gx; d[0] = R; d[1] = G; d[2] = B; break;
                case SWSC_DST_BGR24: d += 3 * gx; d[0] = B; d[1] = G; d[2] = R; break;
                case SWSC_DST_RGBA: d += 4 * gx; d[0] = R; d[1] = G; d[2] = B; d[3] = Av; break;
                case SWSC_DST_BGRA: d += 4 * gx; d[0] = B; d[1] = G; d[2] = R; d[3] = Av; break;
                case SWSC_DST_ARGB: d += 4 * gx; d[0] = Av; d[1] = R; d[2] = G; d[3] = B; break;
                case SWSC_DST_ABGR: d += 4 * gx; d[0] = Av; d[1] = B; d[2] = G; d[3] = R; break;
                }
            } else {
                int Yi = ((int)(Yv - 0x40000000u) >> 14) + 0x10000;
                int Ui = (int)(U - (128u << 23)) >> 14;
                int Vi = (int)(V - (128u << 23)) >> 14;
                if (lfs == 1 && cfs == 2 && cf[0] + cf[1] == 4096 && cf[1] > 0 && cf[1] <= 4096) {
                    /* yuv2rgba64_full_1_c_template with uvalpha != 0 keeps U and V unsigned: its >> 14 is a
                     * logical shift (output.c:1537-1545; chooser vscale.c:138-143) */
                    Ui = (int)((U - (128u << 23)) >> 14);
                    Vi = (int)((V - (128u << 23)) >> 14);
                }
                if (lfs == 1 && (cfs == 1 || (cfs == 2 && cf[0] + cf[1] == 4096 && (unsigned)(int)cf[1] <= 4096u))) {
                    /* the _1 writers shift the line itself (output.c:1497-1500,1278-1281): no x 4096 that could
                     * wrap for 19-bit samples below -2^19 (overshoot is only clipped upwards) */
                    Yi = (int)pl[0] >> 2;
                    if (cfs == 1 || cf[1] == 0) {
                        Ui = ((int)pu[0] - (128 << 11)) >> 2;
                        Vi = ((int)pv[0] - (128 << 11)) >> 2;
                    }
                }
                const unsigned Yu = (unsigned)(Yi - P.rgb.y_offset) * (unsigned)P.rgb.y_coeff + (1u << 13) - (1u << 29);
                const unsigned R = (unsigned)Vi * (unsigned)P.rgb.v2r;
                const unsigned G = (unsigned)Vi * (unsigned)P.rgb.v2g + (unsigned)Ui * (unsigned)P.rgb.u2g;
                const unsigned B = (unsigned)Ui * (unsigned)P.rgb.u2b;
                const int r = clip_uintp2(((int)(R + Yu) >> 14) + (1 << 15), 16);
                const int g = clip_uintp2(((int)(G + Yu) >> 14) + (1 << 15), 16);
                const int b = clip_uintp2(((int)(B + Yu) >> 14) + (1 << 15), 16);
                uint16_t *w = reinterpret_cast<uint16_t *>(d) + 3 * gx;
                w[0] = kind == SWSC_DST_RGB48 ? r : b; w[1] = g; w[2] = kind == SWSC_DST_RGB48 ? b : r;
            }
        }
    } else if (kind >= SWSC_DST_RGB24 && kind <= SWSC_DST_BGRA64) {
        /* packed RGB: one chroma pair per two pixels (output.c:1788-1840 / 1115-1196) */
        /* pairs in this tile; widths are even here except for 15/16 bpp destinations (no full-chroma writer), whose
         * last pair keeps only its first pixel, and the unscaled LUT converters, which leave the odd pixel alone */
        const int pw = (tw + (kind >= SWSC_DST_RGB565 && !P.unscaled_lut ? 1 : 0)) >> 1;
        const bool is16 = kind == SWSC_DST_RGB48 || kind == SWSC_DST_BGR48;
        for (int idx = threadIdx.x; idx < th * (TW >> 1); idx += blockDim.x) {
            const int ty = idx / (TW >> 1), i = idx - ty * (TW >> 1);
            if (i >= pw)
                continue;
            const int y = ry0 + ty;
            const int16_t *lf = P.vl_coef + (size_t)y * lfs;
            const int16_t *cf = P.vc_coef + (size_t)y * cfs;   /* vs == 0 for RGB */
            const int rl = max(P.vl_pos[y], 0) - lo_l;
            const int rc = max(P.vc_pos[y], 0) - lo_c;
            const inter_t *pl = hb_l + (size_t)rl * TW + 2 * i;
            const inter_t *pu = hb_u + (size_t)rc * CW + i;
            const inter_t *pv = hb_v + (size_t)rc * CW + i;
            unsigned Y1 = 0, Y2 = 0, U = 0, V = 0;
            for (int j = 0; j < lfs; j++) {
                const int r = min(rl + j, nl - 1) - rl;
                const unsigned c = (unsigned)(int)lf[j];
                Y1 += (unsigned)(int)pl[(size_t)r * TW]     * c;
                Y2 += (unsigned)(int)pl[(size_t)r * TW + 1] * c;
            }
            for (int j = 0; j < cfs; j++) {
                const int r = min(rc + j, nc - 1) - rc;
                const unsigned c = (unsigned)(int)cf[j];
                U += (unsigned)(int)pu[(size_t)r * CW] * c;
                V += (unsigned)(int)pv[(size_t)r * CW] * c;
            }
            if (!is16 || P.unscaled_lut) {
                int y1v, y2v, uv, vv;
                if (!INTER32) {
                    /* yuv2packed2 (bilinear both ways) has no rounding bias (output.c:1861-1864) */
                    unsigned bias = 1u << 18;
                    if (lfs == 2 && cfs == 2) {
                        const int l0 = lf[0], l1 = lf[1], c0 = cf[0], c1 = cf[1];
                        if (l0 + l1 == 4096 && (unsigned)l1 <= 4096u &&
                            c0 + c1 == 4096 && (unsigned)c1 <= 4096u)
                            bias = 0;
                    }
                    y1v = (int)(Y1 + bias) >> 19; y2v = (int)(Y2 + bias) >> 19;
                    uv  = (int)(U  + bias) >> 19; vv  = (int)(V  + bias) >> 19;
                } else {
                    /* unscaled LUT converter on a 16-bit destination: identity taps, 19-bit lines */
                    y1v = (int)Y1 >> 23; y2v = (int)Y2 >> 23; uv = (int)U >> 23; vv = (int)V >> 23;
                }
                const int u8 = clip_u8(uv), v8 = clip_u8(vv);
                const int oR = P.rgb.base_r + ((v8 * P.rgb.crv) >> 16);
                const int oG = P.rgb.base_g + ((u8 * P.rgb.cgu) >> 16) + ((v8 * P.rgb.cgv) >> 16);
                const int oB = P.rgb.base_b + ((u8 * P.rgb.cbu) >> 16);
                const int cy = P.rgb.cy, yb = P.rgb.yb;
                if (kind >= SWSC_DST_RGB565) {
                    /* 15/16 bpp: the 2x2 ordered-dither offsets move the LUT index, then the bytes are
                     * truncated into their fields (output.c:1714-1747; tables yuv2rgb.c:878-900;
                     * unscaled twin yuv2rgb.c:371-398).  ff_dither_2x2_8 = {6,2 / 0,4}, 2x2_4 = {1,3 / 2,0}. */
                    const int odd = y & 1;
                    const bool is565 = kind <= SWSC_DST_BGR565;
                    const int dr1 = odd ? 0 : 6, dr2 = odd ? 4 : 2;
                    const int db1 = odd ? 6 : 0, db2 = odd ? 2 : 4;
                    const int dg1 = is565 ? (odd ? 2 : 1) : dr2, dg2 = is565 ? (odd ? 0 : 3) : dr1;
                    const int gsh = is565 ? 2 : 3, hi = is565 ? 11 : 10;
                    const bool rgb = kind == SWSC_DST_RGB565 || kind == SWSC_DST_RGB555;   /* R in the high bits */
                    const int r1 = clip_u8((yb + (y1v + oR + dr1) * cy) >> 16) >> 3;
                    const int g1 = clip_u8((yb + (y1v + oG + dg1) * cy) >> 16) >> gsh;
                    const int b1 = clip_u8((yb + (y1v + oB + db1) * cy) >> 16) >> 3;
                    const int r2 = clip_u8((yb + (y2v + oR + dr2) * cy) >> 16) >> 3;
                    const int g2 = clip_u8((yb + (y2v + oG + dg2) * cy) >> 16) >> gsh;
                    const int b2 = clip_u8((yb + (y2v + oB + db2) * cy) >> 16) >> 3;
                    const uint32_t p1 = rgb ? (r1 << hi) | (g1 << 5) | b1 : (b1 << hi) | (g1 << 5) | r1;
                    const uint32_t p2 = rgb ? (r2 << hi) | (g2 << 5) | b2 : (b2 << hi) | (g2 << 5) | r2;
                    uint16_t *w = reinterpret_cast<uint16_t *>(dst0 + (size_t)y * A.dst_stride[0]) + 2 * ((x0 >> 1) + i);
                    w[0] = (uint16_t)p1;
                    if (2 * i + 1 < tw)
                        w[1] = (uint16_t)p2;
                    continue;
                }
                const int r1 = clip_u8((yb + (y1v + oR) * cy) >> 16);
                const int g1 = clip_u8((yb + (y1v + oG) * cy) >> 16);
                const int b1 = clip_u8((yb + (y1v + oB) * cy) >> 16);
                const int r2 = clip_u8((yb + (y2v + oR) * cy) >> 16);
                const int g2 = clip_u8((yb + (y2v + oG) * cy) >> 16);
                const int b2 = clip_u8((yb + (y2v + oB) * cy) >> 16);
                int a1 = 255, a2 = 255;
                if (P.src_alpha) {
                    /* alpha of yuv2rgb_{1,2,X}_c_template (output.c:1818-1829,1870-1875,1903-1932), chosen per row
                     * like the colour channels (vscale.c:135-169) */
                    const inter_t *pa = hb_a + (size_t)rl * TW + 2 * i;
                    const bool chr2 = cfs == 2 && cf[0] + cf[1] == 4096 && (unsigned)(int)cf[1] <= 4096u;
                    const bool lum2 = lfs == 2 && lf[0] + lf[1] == 4096 && (unsigned)(int)lf[1] <= 4096u;
                    if (lfs == 1 && (cfs == 1 || (chr2 && cf[1] == 0))) {   /* yuv2packed1, uvalpha == 0 (a 2-tap chroma
                                                                              * row of {4096, 0} lands here too) */
                        a1 = clip_u8(((int)pa[0] * 255 + 16384) >> 15);
                        a2 = clip_u8(((int)pa[1] * 255 + 16384) >> 15);
                    } else if (lfs == 1 && chr2) {              /* yuv2packed1, uvalpha != 0 */
                        a1 = clip_u8(((int)pa[0] + 64) >> 7);
                        a2 = clip_u8(((int)pa[1] + 64) >> 7);
                    } else {
                        const bool two = lum2 && chr2;          /* yuv2packed2: no rounding bias */
                        unsigned s1 = two ? 0u : 1u << 18, s2 = s1;
                        for (int j = 0; j < lfs; j++) {
                            const int r = min(rl + j, nl - 1) - rl;
                            const unsigned c = (unsigned)(int)lf[j];
                            s1 += (unsigned)(int)pa[(size_t)r * TW] * c;
                            s2 += (unsigned)(int)pa[(size_t)r * TW + 1] * c;
                        }
                        a1 = (int)s1 >> 19; a2 = (int)s2 >> 19;
                        if (two || ((a1 | a2) & 0x100)) {
                            a1 = clip_u8(a1); a2 = clip_u8(a2);
                        }
                        a1 &= 0xFF; a2 &= 0xFF;
                    }
                }
                uint8_t *d = dst0 + (size_t)y * A.dst_stride[0];
                const int gx = (x0 >> 1) + i;          /* pair index in the row */
                switch (kind) {
                case SWSC_DST_RGB24: d += 6 * gx;
                    d[0] = r1; d[1] = g1; d[2] = b1; d[3] = r2; d[4] = g2; d[5] = b2; break;
                case SWSC_DST_BGR24: d += 6 * gx;
                    d[0] = b1; d[1] = g1; d[2] = r1; d[3] = b2; d[4] = g2; d[5] = r2; break;
                case SWSC_DST_RGBA: d += 8 * gx;
                    d[0] = r1; d[1] = g1; d[2] = b1; d[3] = a1; d[4] = r2; d[5] = g2; d[6] = b2; d[7] = a2; break;
                case SWSC_DST_BGRA: d += 8 * gx;
                    d[0] = b1; d[1] = g1; d[2] = r1; d[3] = a1; d[4] = b2; d[5] = g2; d[6] = r2; d[7] = a2; break;
                case SWSC_DST_ARGB: d += 8 * gx;
                    d[0] = a1; d[1] = r1; d[2] = g1; d[3] = b1; d[4] = a2; d[5] = r2; d[6] = g2; d[7] = b2; break;
                case SWSC_DST_ABGR: d += 8 * gx;
                    d[0] = a1; d[1] = b1; d[2] = g1; d[3] = r1; d[4] = a2; d[5] = b2; d[6] = g2; d[7] = r2; break;
                case SWSC_DST_RGB48: { uint16_t *w = reinterpret_cast<uint16_t *>(d) + 6 * gx;
                    w[0] = r1 * 257; w[1] = g1 * 257; w[2] = b1 * 257; w[3] = r2 * 257; w[4] = g2 * 257; w[5] = b2 * 257; break; }
                case SWSC_DST_BGR48: { uint16_t *w = reinterpret_cast<uint16_t *>(d) + 6 * gx;
                    w[0] = b1 * 257; w[1] = g1 * 257; w[2] = r1 * 257; w[3] = b2 * 257; w[4] = g2 * 257; w[5] = r2 * 257; break; }
                }
            } else {
                /* 16-bit arithmetic path, wrapping exactly like the C template */
                unsigned y1u = Y1 - 0x40000000u, y2u = Y2 - 0x40000000u;
                unsigned uu = U - (128u << 23), vu = V - (128u << 23);
                y1u = (unsigned)((int)y1u >> 14) + 0x10000u;
                y2u = (unsigned)((int)y2u >> 14) + 0x10000u;
                uu = (unsigned)((int)uu >> 14);
                vu = (unsigned)((int)vu >> 14);
                if (lfs == 1 && (cfs == 1 || (cfs == 2 && cf[0] + cf[1] == 4096 && (unsigned)(int)cf[1] <= 4096u))) {
                    /* yuv2rgba64_1_c_template (output.c:1278-1281): the line itself >> 2, nothing that could wrap */
                    y1u = (unsigned)((int)pl[0] >> 2);
                    y2u = (unsigned)((int)pl[1] >> 2);
                    if (cfs == 1 || cf[1] == 0) {
                        uu = (unsigned)(((int)pu[0] - (128 << 11)) >> 2);
                        vu = (unsigned)(((int)pv[0] - (128 << 11)) >> 2);
                    }
                }
                y1u -= (unsigned)P.rgb.y_offset; y2u -= (unsigned)P.rgb.y_offset;
                y1u *= (unsigned)P.rgb.y_coeff;  y2u *= (unsigned)P.rgb.y_coeff;
                y1u += (1u << 13) - (1u << 29);  y2u += (1u << 13) - (1u << 29);
                const unsigned R = vu * (unsigned)P.rgb.v2r;
                const unsigned G = vu * (unsigned)P.rgb.v2g + uu * (unsigned)P.rgb.u2g;
                const unsigned B = uu * (unsigned)P.rgb.u2b;
                const int r1 = clip_uintp2(((int)(R + y1u) >> 14) + (1 << 15), 16);
                const int g1 = clip_uintp2(((int)(G + y1u) >> 14) + (1 << 15), 16);
                const int b1 = clip_uintp2(((int)(B + y1u) >> 14) + (1 << 15), 16);
                const int r2 = clip_uintp2(((int)(R + y2u) >> 14) + (1 << 15), 16);
                const int g2 = clip_uintp2(((int)(G + y2u) >> 14) + (1 << 15), 16);
                const int b2 = clip_uintp2(((int)(B + y2u) >> 14) + (1 << 15), 16);
                uint16_t *w = reinterpret_cast<uint16_t *>(dst0 + (size_t)y * A.dst_stride[0]) + 6 * ((x0 >> 1) + i);
                if (kind == SWSC_DST_RGB48) {
                    w[0] = r1; w[1] = g1; w[2] = b1; w[3] = r2; w[4] = g2; w[5] = b2;
                } else {
                    w[0] = b1; w[1] = g1; w[2] = r1; w[3] = b2; w[4] = g2; w[5] = r2;
                }
            }
        }
    } else {
        /* planar / semi-planar YUV (output.c:163-187,340-357,468-528; vscale.c:41-107) */
        const int bits = P.dst_bits;
        for (int idx = threadIdx.x; idx < th * TW; idx += blockDim.x) {
            const int ty = idx / TW, x = idx - ty * TW;
            if (x >= tw)
                continue;
            const int y = ry0 + ty;
            const int16_t *lf = P.vl_coef + (size_t)y * lfs;
            const int rl = max(P.vl_pos[y], 0) - lo_l;
            const inter_t *pl = hb_l + (size_t)rl * TW + x;
            unsigned acc = 0;
            for (int j = 0; j < lfs; j++) {
                const int r = min(rl + j, nl - 1) - rl;
                acc += (unsigned)(int)pl[(size_t)r * TW] * (unsigned)(int)lf[j];
            }
            const int gx = x0 + x;
            uint8_t *d = dst0 + (size_t)y * A.dst_stride[0];
            if (kind == SWSC_DST_PLANAR8 || kind == SWSC_DST_NV12 || kind == SWSC_DST_NV21) {
                const int dz = P.dither_bayer ? c_dither_8x8_128[y & 7][gx & 7] : 64;
                d[gx] = clip_u8((int)(acc + ((unsigned)dz << 12)) >> 19);
            } else if (kind == SWSC_DST_PLANARN) {
                const int shift = 27 - bits;
                reinterpret_cast<uint16_t *>(d)[gx] = clip_uintp2((int)(acc + (1u << (shift - 1))) >> shift, bits);
            } else if (kind == SWSC_DST_P010) {     /* yuv2p010l1_c / yuv2p010lX_c (output.c:538-566): 10 bits << 6 */
                const int shift = 27 - 10;
                reinterpret_cast<uint16_t *>(d)[gx] = clip_uintp2((int)(acc + (1u << (shift - 1))) >> shift, 10) << 6;
            } else {
                /* yuv2planeX_16_c (output.c:163-187).  One tap goes through yuv2plane1_16_c, which never multiplies:
                 * a 19-bit line below -2^19 (hScale*To19 only clips upwards; sinc / lanczos overshoot of noise gets
                 * there) must not wrap through the x 4096 of the general form */
                const int v16 = lfs == 1 ? clip_uintp2(((int)pl[0] + 4) >> 3, 16)
                                         : 0x8000 + clip_i16((int)(acc + (1u << 14) - 0x40000000u) >> 15);
                if (kind == SWSC_DST_PLANARF32)     /* yuv2plane1_float / yuv2planeX_float (output.c:219-263): x 1/65535 */
                    reinterpret_cast<float *>(d)[gx] = __fmul_rn(1.0f / 65535.0f, (float)v16);
                else
                    reinterpret_cast<uint16_t *>(d)[gx] = (uint16_t)v16;
            }
        }
        if (P.has_chroma && P.dst_has_chroma && ch > 0) {
            for (int idx = threadIdx.x; idx < ch * CW; idx += blockDim.x) {
                const int ty = idx / CW, x = idx - ty * CW;
                if (x >= cw)
                    continue;
                const int y = cy0 + ty;
                const int16_t *cf = P.vc_coef + (size_t)y * cfs;
                const int rc = max(P.vc_pos[y], 0) - lo_c;
                const inter_t *pu = hb_u + (size_t)rc * CW + x;
                const inter_t *pv = hb_v + (size_t)rc * CW + x;
                unsigned au = 0, av = 0;
                for (int j = 0; j < cfs; j++) {
                    const int r = min(rc + j, nc - 1) - rc;
                    const unsigned c = (unsigned)(int)cf[j];
                    au += (unsigned)(int)pu[(size_t)r * CW] * c;
                    av += (unsigned)(int)pv[(size_t)r * CW] * c;
                }
                const int gx = cx0 + x;
                if (kind == SWSC_DST_PLANAR8) {
                    const int du = P.dither_bayer ? c_dither_8x8_128[y & 7][gx & 7] : 64;
                    const int dv = P.dither_bayer ? c_dither_8x8_128[y & 7][(gx + 3) & 7] : 64;
                    dst1[(size_t)y * A.dst_stride[1] + gx] = clip_u8((int)(au + ((unsigned)du << 12)) >> 19);
                    dst2[(size_t)y * A.dst_stride[2] + gx] = clip_u8((int)(av + ((unsigned)dv << 12)) >> 19);
                } else if (kind == SWSC_DST_NV12 || kind == SWSC_DST_NV21) {
                    const int du = P.dither_bayer ? c_dither_8x8_128[y & 7][gx & 7] : 64;
                    const int dv = P.dither_bayer ? c_dither_8x8_128[y & 7][(gx + 3) & 7] : 64;
                    const int u8 = clip_u8((int)(au + ((unsigned)du << 12)) >> 19);
                    const int v8 = clip_u8((int)(av + ((unsigned)dv << 12)) >> 19);
                    uint8_t *d = dst1 + (size_t)y * A.dst_stride[1] + 2 * gx;
                    d[0] = kind == SWSC_DST_NV12 ? u8 : v8;
                    d[1] = kind == SWSC_DST_NV12 ? v8 : u8;
                } else if (kind == SWSC_DST_P010) { /* yuv2p010cX_c (output.c:568-589): interleaved, U first */
                    const int shift = 27 - 10;
                    uint16_t *d = reinterpret_cast<uint16_t *>(dst1 + (size_t)y * A.dst_stride[1]) + 2 * gx;
                    d[0] = clip_uintp2((int)(au + (1u << (shift - 1))) >> shift, 10) << 6;
                    d[1] = clip_uintp2((int)(av + (1u << (shift - 1))) >> shift, 10) << 6;
                } else if (kind == SWSC_DST_PLANARN) {
                    const int shift = 27 - bits;
                    reinterpret_cast<uint16_t *>(dst1 + (size_t)y * A.dst_stride[1])[gx] =
                        clip_uintp2((int)(au + (1u << (shift - 1))) >> shift, bits);
                    reinterpret_cast<uint16_t *>(dst2 + (size_t)y * A.dst_stride[2])[gx] =
                        clip_uintp2((int)(av + (1u << (shift - 1))) >> shift, bits);
                } else {
                    const int u = cfs == 1 ? clip_uintp2(((int)pu[0] + 4) >> 3, 16)
                                           : 0x8000 + clip_i16((int)(au + (1u << 14) - 0x40000000u) >> 15);
                    const int v = cfs == 1 ? clip_uintp2(((int)pv[0] + 4) >> 3, 16)
                                           : 0x8000 + clip_i16((int)(av + (1u << 14) - 0x40000000u) >> 15);
                    reinterpret_cast<uint16_t *>(dst1 + (size_t)y * A.dst_stride[1])[gx] = (uint16_t)u;
                    reinterpret_cast<uint16_t *>(dst2 + (size_t)y * A.dst_stride[2])[gx] = (uint16_t)v;
                }
            }
        }
    }
}

/* ======================================================================== host side */

struct SwsCudaState {
    int device;
    cudaStream_t stream;
    SwsCudaPlan plan;
    void *tables;            /* one allocation holding the four FIR banks */
    /* host copies of the vertical banks for launch planning */
    int32_t *h_vl_pos, *h_vc_pos;
    /* staging frames for host-pointer sws_scale() */
    uint8_t *d_src[4], *d_dst[4];
    uint8_t *h_src[4], *h_dst[4];   /* page-locked twins of the staging planes: bounce ring for pageable frames */
    /* frame ring of sws_cuda_scale_batch_host(): RING_DEPTH staging sets so that the upload of one frame, the
     * kernel of another and the download of a third overlap (slot 0 is d_src/d_dst) */
    uint8_t *ring_src[4][4], *ring_dst[4][4];
    int ring_ready;
    cudaEvent_t rev_in[4], rev_k[4], rev_out[4];
    uint8_t *d_flip;                /* scratch for mirroring the rows of bottom-up frames */
    size_t flip_bytes;
    int staging_ready, bounce_src_ready, bounce_dst_ready;
    int d_src_stride[4], d_dst_stride[4];
    int src_rows[4], dst_rows[4];
    int src_rowbytes[4], dst_rowbytes[4];
    long launches;
    const char *kernel_name;
    /* fast420 path */
    int fast_ok;
    int r420_ok, r420_cr;
    int fast_v422;               /* the source has one chroma row per luma row: TH + 4 staged chroma rows */
    int fast_narrow;             /* the 128 x 64 tile shape of the fast420 kernel is set up and preferred */
    int4 *d_fast_rows_narrow;
    int fasthi8_ok, fasthi8_crows;
    int s8_stages, s8_srck, s8_elt_shift, s8_seg_sy, s8_seg_sc, s8_wide;
    int s8_ok, s8_fs4, s8_tile_h, s8_nl_cap, s8_nc_cap, s8_seg_l, s8_seg_c, s8_slot, s8_vl_n4, s8_vc_n4;
    size_t s8_smem;
    void *s8_tables;
    int *s8_hl_pos, *s8_hc_pos;
    uint32_t *s8_hl_cl, *s8_hl_ch, *s8_hc_cl, *s8_hc_ch;
    int s8_mma;                  /* horizontal stage on the tensor pipe: s8_fs4 counts K steps of 32 samples */
    int *s8_hl_goff, *s8_hc_goff;
    uint32_t *s8_hl_B, *s8_hc_B;
    S8VRow *s8_vl, *s8_vc, *s8_vl2, *s8_vc2;
    int fast16_ok, fast16_taps;
    /* tile15 path */
    int t15_ok, t15_srck, t15_ht, t15_outk, t15_tile_h, t15_nl_cap, t15_nc_cap, t15_seg_l, t15_seg_c, t15_srl, t15_src;
    size_t t15_smem;
    unsigned disabled;       /* SWS_B200_DISABLE: bit set of kernels switched off for A/B runs */
    Fast16Row *d_fast16_rows;
    int e2e_mode, e2e_bands; /* how sws_scale() moves page-locked host frames (see scale_host)  */
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[16], ev_k[16], ev_out[16];
    int4 *d_fast_rows;
    int num_sms;
    int tile_w, tile_h, rows_l_cap, rows_c_cap;
    size_t smem_bytes;
};

extern "C" int sws_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

/* NUMA node the PCIe root of a device hangs off (-1: unknown or single-node box) */
static int device_numa_node(int dev)
{
    static int cache[64];
    static bool known[64];
    if (dev < 0 || dev >= 64)
        return -1;
    if (!known[dev]) {
        char id[32] = { 0 };
        cache[dev] = -1;
        if (cudaDeviceGetPCIBusId(id, sizeof(id), dev) == cudaSuccess)
            cache[dev] = ff_b200_numa_node_of_pci(id);
        else
            cudaGetLastError();
        known[dev] = true;
    }
    return cache[dev];
}

/* page-locked memory on the NUMA node of the current device: DMA does not cross the inter-socket link */
static cudaError_t host_alloc_local(void **p, size_t size)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        dev = -1;
    const int node = device_numa_node(dev);
    const bool bound = node >= 0 && ff_b200_numa_prefer(node) == 0;
    cudaError_t e = cudaHostAlloc(p, size, cudaHostAllocDefault);
    if (bound)
        ff_b200_numa_restore();
    return e;
}

extern "C" void *sws_cuda_host_alloc(size_t size)
{
    void *p = nullptr;
    if (host_alloc_local(&p, size) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

/* Run the calling thread on the CPUs next to `device` (its NUMA node); returns the number of CPUs in the
 * new affinity set, or -1 when the topology is unknown / SWS_B200_NUMA=0 (nothing changed). */
extern "C" int sws_cuda_bind_thread_to_device(int device)
{
    return ff_b200_numa_bind_thread(device_numa_node(device));
}

extern "C" int sws_cuda_device_numa_node(int device)
{
    return device_numa_node(device);
}

extern "C" void sws_cuda_host_free(void *ptr)
{
    if (ptr)
        cudaFreeHost(ptr);
}

extern "C" int ff_b200_cuda_probe(void)
{
    int n = sws_cuda_device_count(), dev = 0;
    if (n <= 0)
        return AVERROR(ENOSYS);
    if (cudaGetDevice(&dev) != cudaSuccess)
        return AVERROR(ENOSYS);
    return dev;
}

/* most h-scaled source rows any window of `th` consecutive output rows can need
 * (tiles may start at any row when the legacy slice API is used) */
static int max_rows_needed(const int32_t *pos, int fs, int n_out, int th)
{
    int worst = fs;
    for (int y = 0; y < n_out; y++) {
        int y1 = y + th < n_out ? y + th : n_out;
        int lo = INT32_MAX, hi = 0;
        for (int k = y; k < y1; k++) {
            int p = pos[k] < 0 ? 0 : pos[k];
            if (p < lo) lo = p;
            if (pos[k] + fs > hi) hi = pos[k] + fs;
        }
        if (hi - lo > worst)
            worst = hi - lo;
    }
    return worst;
}

static int plan_tiles(SwsCudaState *st)
{
    const SwsCudaPlan *p = &st->plan;
    const int isz = p->inter_bits == 19 ? 4 : 2;
    const int vs = p->chr_dst_vsub, hs = p->chr_dst_hsub;
    const size_t budget = 96 * 1024;
    int tw = 128;
    while (tw > 32 && tw / 2 >= p->dst_w)
        tw /= 2;
    /* very long vertical filters (strong downscales into 19-bit lines) trade tile width for rows */
    for (; tw >= 32; tw >>= 1) {
        for (int th = 32; th >= (1 << vs); th >>= 1) {
            int rl = max_rows_needed(st->h_vl_pos, p->vl_size, p->dst_h, th);
            int cth = th >> vs ? th >> vs : 1;
            int rc = max_rows_needed(st->h_vc_pos, p->vc_size, p->chr_dst_h, cth);
            size_t need = ((size_t)rl * tw * (p->src_alpha ? 2 : 1) + 2 * (size_t)rc * (tw >> hs)) * isz;
            if (need <= budget || th == (1 << vs)) {
                if (need > 200 * 1024)
                    break;
                st->tile_w = tw; st->tile_h = th; st->rows_l_cap = rl; st->rows_c_cap = rc;
                st->smem_bytes = need;
                return 0;
            }
        }
    }
    return AVERROR(ENOTSUP);
}


static int tile15_setup(SwsCudaState *st, const SwsFirBank *hl, const SwsFirBank *hc,
                        const SwsFirBank *vl, const SwsFirBank *vc);
static int rgb420_setup(SwsCudaState *st, const SwsFirBank *vc);

/* ---------------------------------------------------------------- fast420 host side */

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_tiled(void)
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}

static int make_map_3d(CUtensorMap *m, CUtensorMapDataType dt, const void *base, uint64_t d0, uint64_t d1,
                       uint64_t d2, uint64_t stride1, uint64_t stride2, uint32_t b0, uint32_t b1)
{
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc)
        return AVERROR(ENOSYS);
    cuuint64_t dims[3] = { d0, d1, d2 };
    cuuint64_t strides[2] = { stride1, stride2 };
    cuuint32_t box[3] = { b0, b1, 1 };
    cuuint32_t es[3] = { 1, 1, 1 };
    CUresult r = enc(m, dt, 3, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[swscaler-b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
        return AVERROR(EIO);
    }
    return 0;
}

typedef void (*fast420_kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                 const Fast420Args);

static int fast420_fmt(int dst_kind)
{
    switch (dst_kind) {
    case SWSC_DST_RGB24: return F420_RGB24;
    case SWSC_DST_BGR24: return F420_BGR24;
    case SWSC_DST_RGBA:  return F420_RGBA;
    case SWSC_DST_BGRA:  return F420_BGRA;
    case SWSC_DST_ARGB:  return F420_ARGB;
    default:             return F420_ABGR;
    }
}

template <int FMT, bool NARROW, bool V422>
static fast420_kernel_t pick_fast420_src(int layout)
{
    return layout == SWSC_SRC_PLANAR ? sws_fast420_rgb8_kernel<FMT, F420_PLANAR, NARROW, V422>
         : layout == SWSC_SRC_NV12   ? sws_fast420_rgb8_kernel<FMT, F420_NV12, NARROW, V422>
                                     : sws_fast420_rgb8_kernel<FMT, F420_NV21, NARROW, V422>;
}

template <bool NARROW, bool V422>
static fast420_kernel_t pick_fast420_shape(int fmt, int layout)
{
    switch (fmt) {
    case F420_RGB24: return pick_fast420_src<F420_RGB24, NARROW, V422>(layout);
    case F420_BGR24: return pick_fast420_src<F420_BGR24, NARROW, V422>(layout);
    case F420_RGBA:  return pick_fast420_src<F420_RGBA, NARROW, V422>(layout);
    case F420_BGRA:  return pick_fast420_src<F420_BGRA, NARROW, V422>(layout);
    case F420_ARGB:  return pick_fast420_src<F420_ARGB, NARROW, V422>(layout);
    default:         return pick_fast420_src<F420_ABGR, NARROW, V422>(layout);
    }
}

/* narrow = 128 x 64 tiles; v422 = TH + 4 staged chroma rows (sources without vertical chroma subsampling; wide shape only) */
static fast420_kernel_t pick_fast420(int fmt, int layout, bool narrow = false, bool v422 = false)
{
    if (v422)
        return pick_fast420_shape<false, true>(fmt, layout);
    return narrow ? pick_fast420_shape<true, false>(fmt, layout) : pick_fast420_shape<false, false>(fmt, layout);
}

/* per-row metadata of the same-size 8-bit kernel for tiles of `th` rows that stage `crows` chroma rows: returns a
 * device array, or nullptr with *fits = 0 when some tile's chroma window does not fit */
static int fast420_rows(const SwsCudaPlan *p, const SwsFirBank *vc, int th, int crows, int4 **d_rows, int *fits)
{
    *fits = 0;
    *d_rows = nullptr;
    const int padded = ((vc->len + th - 1) / th + 1) * th;
    int4 *rows = (int4 *)calloc(padded, sizeof(int4));
    if (!rows)
        return AVERROR(ENOMEM);
    for (int y = 0; y < vc->len; y++) {
        uint32_t cl = 0, ch = 0;
        for (int j = 0; j < vc->size; j++) {
            const int c = vc->coef[(size_t)y * vc->size + j];
            cl |= (uint32_t)(c & 0xFF) << (8 * j);
            ch |= (uint32_t)((c >> 8) & 0xFF) << (8 * j);
        }
        int pos = vc->pos[y];
        /* keep the 4-row window inside the plane: taps beyond vc->size are zero */
        if (pos + 4 > p->chr_src_h && p->chr_src_h >= 4) {
            const int shift = pos + 4 - p->chr_src_h;
            if (shift * 8 < 32 && (cl >> (32 - 8 * shift)) == 0 && (ch >> (32 - 8 * shift)) == 0) {
                cl <<= 8 * shift; ch <<= 8 * shift; pos -= shift;
            }
        }
        rows[y] = make_int4(0, (int)cl, (int)ch, pos);
    }
    for (int y = vc->len; y < padded; y++)
        rows[y] = rows[vc->len - 1];
    /* the kernel's window only slides down: positions must be monotonic, and every tile must
     * fit the staged chroma rows */
    for (int y = 0; y < padded; y++) {
        const int base = rows[y - y % th].w;
        rows[y].x = rows[y].w - base;
        if ((y > 0 && rows[y].w < rows[y - 1].w) || rows[y].x < 0 || rows[y].x + 4 > crows) {
            free(rows);
            return 0;
        }
    }
    cudaError_t e = cudaMalloc(d_rows, sizeof(int4) * padded);
    if (e == cudaSuccess)
        e = cudaMemcpy(*d_rows, rows, sizeof(int4) * padded, cudaMemcpyHostToDevice);
    free(rows);
    CUDA_OK(e);
    *fits = 1;
    return 0;
}

static int fast420_setup(SwsCudaState *st, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->fast_ok = 0;
    if (p->src_bits != 8 || p->inter_bits != 15 || p->src_layout > SWSC_SRC_NV21)
        return 0;
    if (p->dst_kind < SWSC_DST_RGB24 || p->dst_kind > SWSC_DST_ABGR || p->full_chr)
        return 0;
    if (!p->lum_identity || !p->chr_h_identity || p->chr_src_hsub != 1 || p->chr_dst_hsub != 1)
        return 0;
    if (vc->size > 4 || p->range_mode || !get_encode_tiled())
        return 0;
    if ((p->dst_kind == SWSC_DST_RGB24 || p->dst_kind == SWSC_DST_BGR24) && (p->dst_w & 3))
        return 0;                     /* the store tensor map counts 32-bit words */
    if (p->unscaled_lut && (p->dst_w & 1))
        return 0;                     /* the LUT converters leave the last pixel of an odd row untouched */
    /* staged chroma rows per tile: 20 (36 for the narrow shape) cover vertically subsampled chroma; 4:2:2 sources
     * need one chroma row per luma row, TH + 4 (a separate instantiation of the wide shape) */
    int fits = 0, ret;
    st->fast_v422 = 0;
    if ((ret = fast420_rows(p, vc, F420_TH, F420_CROWS, &st->d_fast_rows, &fits)) < 0)
        return ret;
    if (!fits) {
        if ((ret = fast420_rows(p, vc, F420_TH, F420_TH + 4, &st->d_fast_rows, &fits)) < 0)
            return ret;
        st->fast_v422 = fits;
    }
    if (!fits)
        return 0;
    const int fmt = fast420_fmt(p->dst_kind);
    const int bpp = fmt >= F420_RGBA ? 4 : 3;
    CUDA_OK(set_max_smem((const void *)pick_fast420(fmt, p->src_layout, false, st->fast_v422), (size_t)(F420_SMEM_CROWS(bpp, F420_TW, F420_TH, st->fast_v422 ? F420_TH + 4 : F420_CROWS))));
    /* 128 x 64 tiles when they cover the frame width with less waste than 256 x 32 tiles (1920, 640, ...) */
    st->fast_narrow = 0;
    if (!st->fast_v422 && (p->dst_w + 127) / 128 * 128 < (p->dst_w + 255) / 256 * 256 && !getenv("SWS_B200_NO_NARROW")) {
        if ((ret = fast420_rows(p, vc, 2 * F420_TH, F420_CROWS_NARROW, &st->d_fast_rows_narrow, &fits)) < 0)
            return ret;
        if (fits) {
            CUDA_OK(set_max_smem((const void *)pick_fast420(fmt, p->src_layout, true), (size_t)(F420_SMEM(bpp))));
            st->fast_narrow = 1;
        }
    }
    st->fast_ok = 1;
    st->kernel_name = "fast420_rgb8_tma";
    return 0;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

/* returns 1 if launched, 0 if the arguments do not qualify (caller falls back to the generic kernel) */
static int fast420_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                          const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                          const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    /* row ranges must start on a tile row; the end is clipped by the store tensor map */
    if (!st->fast_ok || (st->disabled & 1) || (y0 % F420_TH) || y1 <= y0 || y1 > p->dst_h)
        return 0;
    const bool planar = p->src_layout == SWSC_SRC_PLANAR;
    const int fmt = fast420_fmt(p->dst_kind);
    const int bpp = fmt >= F420_RGBA ? 4 : 3;
    for (int i = 0; i < (planar ? 3 : 2); i++)
        if (!src[i] || !aligned16(src[i]) || (src_stride[i] & 15) || src_stride[i] <= 0 ||
            (nb_frames > 1 && (src_fstride[i] & 15 || src_fstride[i] <= 0)))
            return 0;
    if (!aligned16(dst[0]) || (dst_stride[0] & 15) || dst_stride[0] <= 0 ||
        (nb_frames > 1 && (dst_fstride[0] & 15 || dst_fstride[0] <= 0)))
        return 0;
    /* tile shape: 128 x 64 when that wastes less of the right-most tile column and the row range allows it */
    const bool narrow = st->fast_narrow && !(y0 % (2 * F420_TH));
    const int TW = narrow ? F420_TW / 2 : F420_TW, TH = narrow ? 2 * F420_TH : F420_TH;
    const int CROWS = st->fast_v422 ? F420_TH + 4 : narrow ? F420_CROWS_NARROW : F420_CROWS;
    CUtensorMap my, mu, mv, mo;
    const uint64_t fs_y = nb_frames > 1 ? src_fstride[0] : (uint64_t)src_stride[0] * p->src_h;
    const uint64_t fs_u = nb_frames > 1 ? src_fstride[1] : (uint64_t)src_stride[1] * p->chr_src_h;
    const uint64_t fs_o = nb_frames > 1 ? dst_fstride[0] : (uint64_t)dst_stride[0] * p->dst_h;
    int ret;
    if ((ret = make_map_3d(&my, CU_TENSOR_MAP_DATA_TYPE_UINT8, src[0], p->src_w, p->src_h, nb_frames,
                           src_stride[0], fs_y, TW, TH)) < 0)
        return ret;
    if (planar) {
        const uint64_t fs_v = nb_frames > 1 ? src_fstride[2] : (uint64_t)src_stride[2] * p->chr_src_h;
        if ((ret = make_map_3d(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT8, src[1], p->chr_src_w, p->chr_src_h, nb_frames,
                               src_stride[1], fs_u, TW / 2, CROWS)) < 0 ||
            (ret = make_map_3d(&mv, CU_TENSOR_MAP_DATA_TYPE_UINT8, src[2], p->chr_src_w, p->chr_src_h, nb_frames,
                               src_stride[2], fs_v, TW / 2, CROWS)) < 0)
            return ret;
    } else {
        /* interleaved UV plane: 2 bytes per chroma sample, one TW-byte box row per tile row */
        if ((ret = make_map_3d(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT8, src[1], 2 * (uint64_t)p->chr_src_w,
                               p->chr_src_h, nb_frames, src_stride[1], fs_u, TW, CROWS)) < 0)
            return ret;
        mv = mu;
    }
    if ((ret = make_map_3d(&mo, CU_TENSOR_MAP_DATA_TYPE_UINT32, dst[0], (uint64_t)p->dst_w * bpp / 4, y1,
                           nb_frames, dst_stride[0], fs_o, TW * bpp / 4, TH / F420_CWARPS)) < 0)
        return ret;
    Fast420Args a;
    a.tiles_x = (p->dst_w + TW - 1) / TW;
    a.tiles_y = (y1 - y0 + TH - 1) / TH;
    a.ty_first = y0 / TH;
    a.frames = nb_frames;
    a.dst_h = p->dst_h;
    a.cy = p->rgb.cy; a.yb = p->rgb.yb;
    a.crv = p->rgb.crv; a.cbu = p->rgb.cbu; a.cgu = p->rgb.cgu; a.cgv = p->rgb.cgv;
    a.kr = p->rgb.base_r << 16; a.kg = p->rgb.base_g << 16; a.kb = p->rgb.base_b << 16;
    a.rows = narrow ? st->d_fast_rows_narrow : st->d_fast_rows;
    const long long total = (long long)a.tiles_x * a.tiles_y * nb_frames;
    const int grid = (int)(total < (long long)st->num_sms * F420_CTAS_PER_SM ? total : (long long)st->num_sms * F420_CTAS_PER_SM);
    pick_fast420(fmt, p->src_layout, narrow, st->fast_v422)<<<grid, F420_THREADS, F420_SMEM_CROWS(bpp, TW, TH, CROWS), stream>>>(my, mu, mv, mo, a);
    st->kernel_name = "fast420_rgb8_tma";     /* the name always reports the last kernel launched */
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}


/* ---------------------------------------------------------------- fast420 16-bit host side */

typedef void (*fast16_kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                const Fast16Args);

static fast16_kernel_t pick_fast16(int taps, bool bgr)
{
    switch (taps) {
    case 4:  return bgr ? sws_fast420_rgb16_kernel<4, true> : sws_fast420_rgb16_kernel<4, false>;
    case 6:  return bgr ? sws_fast420_rgb16_kernel<6, true> : sws_fast420_rgb16_kernel<6, false>;
    default: return bgr ? sws_fast420_rgb16_kernel<8, true> : sws_fast420_rgb16_kernel<8, false>;
    }
}

/* per-row metadata shared by the two high-depth same-size kernels: first chroma source row (tile relative and
 * absolute) and the eight int16 vertical chroma taps.  Returns 1 when uploaded, 0 when the geometry does not
 * fit the kernels' 24-row chroma window, < 0 on error. */
static int fast16_rows(SwsCudaState *st, const SwsFirBank *vc, int taps, int crows = F16_CROWS)
{
    if (st->d_fast16_rows)
        return 1;
    const int padded = ((vc->len + F16_TH - 1) / F16_TH + 1) * F16_TH;
    Fast16Row *rows = (Fast16Row *)calloc(padded, sizeof(Fast16Row));
    if (!rows)
        return AVERROR(ENOMEM);
    for (int y = 0; y < padded; y++) {
        const int yy = y < vc->len ? y : vc->len - 1;
        int c[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int j = 0; j < vc->size; j++)
            c[j] = vc->coef[(size_t)yy * vc->size + j];
        rows[y].pos_abs = vc->pos[yy];
        rows[y].c01 = (c[0] & 0xFFFF) | (c[1] << 16);
        rows[y].c23 = (c[2] & 0xFFFF) | (c[3] << 16);
        rows[y].c45 = (c[4] & 0xFFFF) | (c[5] << 16);
        rows[y].c67 = (c[6] & 0xFFFF) | (c[7] << 16);
    }
    for (int y = 0; y < padded; y++) {
        const int base = rows[y & ~(F16_TH - 1)].pos_abs;
        rows[y].pos_rel = rows[y].pos_abs - base;
        if ((y > 0 && rows[y].pos_abs < rows[y - 1].pos_abs) || rows[y].pos_rel < 0 ||
            rows[y].pos_rel + taps > crows) {
            free(rows);
            return 0;
        }
    }
    cudaError_t e = cudaMalloc(&st->d_fast16_rows, sizeof(Fast16Row) * padded);
    if (e == cudaSuccess)
        e = cudaMemcpy(st->d_fast16_rows, rows, sizeof(Fast16Row) * padded, cudaMemcpyHostToDevice);
    free(rows);
    CUDA_OK(e);
    return 1;
}

static int fast16_setup(SwsCudaState *st, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->fast16_ok = 0;
    if (p->src_layout != SWSC_SRC_PLANAR || p->src_bits <= 8 || p->src_bits > 16 || p->inter_bits != 19)
        return 0;
    if (p->dst_kind != SWSC_DST_RGB48 && p->dst_kind != SWSC_DST_BGR48)
        return 0;
    if (!p->lum_identity || !p->chr_h_identity || p->chr_src_hsub != 1 || p->chr_dst_hsub != 1)
        return 0;
    if (vc->size > 8 || p->range_mode || p->full_chr || p->unscaled_lut || (p->dst_w & 1) || !get_encode_tiled())
        return 0;
    const int taps = vc->size <= 4 ? 4 : vc->size <= 6 ? 6 : 8;
    int ret = fast16_rows(st, vc, taps);
    if (ret <= 0)
        return ret;
    CUDA_OK(set_max_smem((const void *)pick_fast16(taps, false), (size_t)(F16_SMEM)));
    CUDA_OK(set_max_smem((const void *)pick_fast16(taps, true), (size_t)(F16_SMEM)));
    st->fast16_ok = 1;
    st->fast16_taps = taps;
    st->kernel_name = "fast420_rgb16_tma";
    return 0;
}

static int fast16_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                         const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                         const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    if (!st->fast16_ok || (st->disabled & 2) || (y0 % F16_TH) || y1 <= y0 || y1 > p->dst_h)
        return 0;
    for (int i = 0; i < 3; i++)
        if (!aligned16(src[i]) || (src_stride[i] & 15) || src_stride[i] <= 0 ||
            (nb_frames > 1 && (src_fstride[i] & 15 || src_fstride[i] <= 0)))
            return 0;
    if (!aligned16(dst[0]) || (dst_stride[0] & 15) || dst_stride[0] <= 0 ||
        (nb_frames > 1 && (dst_fstride[0] & 15 || dst_fstride[0] <= 0)))
        return 0;
    CUtensorMap my, mu, mv, mo;
    const uint64_t fs_y = nb_frames > 1 ? src_fstride[0] : (uint64_t)src_stride[0] * p->src_h;
    const uint64_t fs_u = nb_frames > 1 ? src_fstride[1] : (uint64_t)src_stride[1] * p->chr_src_h;
    const uint64_t fs_v = nb_frames > 1 ? src_fstride[2] : (uint64_t)src_stride[2] * p->chr_src_h;
    const uint64_t fs_o = nb_frames > 1 ? dst_fstride[0] : (uint64_t)dst_stride[0] * p->dst_h;
    int ret;
    if ((ret = make_map_3d(&my, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[0], p->src_w, p->src_h, nb_frames,
                           src_stride[0], fs_y, F16_TW, F16_TH)) < 0 ||
        (ret = make_map_3d(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[1], p->chr_src_w, p->chr_src_h, nb_frames,
                           src_stride[1], fs_u, F16_TW / 2, F16_CROWS)) < 0 ||
        (ret = make_map_3d(&mv, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[2], p->chr_src_w, p->chr_src_h, nb_frames,
                           src_stride[2], fs_v, F16_TW / 2, F16_CROWS)) < 0 ||
        (ret = make_map_3d(&mo, CU_TENSOR_MAP_DATA_TYPE_UINT32, dst[0], (uint64_t)p->dst_w * 3 / 2, y1,
                           nb_frames, dst_stride[0], fs_o, F16_TW * 6 / 4, F16_TH / F420_CWARPS)) < 0)
        return ret;
    Fast16Args a;
    memset(&a, 0, sizeof(a));
    a.tiles_x = (p->dst_w + F16_TW - 1) / F16_TW;
    a.tiles_y = (y1 - y0 + F16_TH - 1) / F16_TH;
    a.ty_first = y0 / F16_TH;
    a.frames = nb_frames;
    a.dst_h = p->dst_h;
    a.s19 = 19 - p->src_bits;
    a.bgr = p->dst_kind == SWSC_DST_BGR48;
    a.ycoef = (unsigned)p->rgb.y_coeff;
    a.kconst = (1u << 13) - (1u << 29) - (unsigned)p->rgb.y_offset * (unsigned)p->rgb.y_coeff;
    a.v2r = (unsigned)p->rgb.v2r; a.v2g = (unsigned)p->rgb.v2g;
    a.u2g = (unsigned)p->rgb.u2g; a.u2b = (unsigned)p->rgb.u2b;
    a.rows = st->d_fast16_rows;
    const long long total = (long long)a.tiles_x * a.tiles_y * nb_frames;
    const int grid = (int)(total < (long long)st->num_sms * F420_CTAS_PER_SM ? total : (long long)st->num_sms * F420_CTAS_PER_SM);
    pick_fast16(st->fast16_taps, a.bgr)<<<grid, F420_THREADS, F16_SMEM, stream>>>(my, mu, mv, mo, a);
    st->kernel_name = "fast420_rgb16_tma";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}


/* ---------------------------------------------------------------- fast420 high-depth -> 8-bit RGB host side */

typedef void (*fasthi8_kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                 const FastHi8Args);

template <int TAPS, bool SEMI>
static fasthi8_kernel_t pick_fasthi8_fmt(int fmt)
{
    switch (fmt) {
    case F420_RGB24: return sws_fast420_hi8_kernel<TAPS, F420_RGB24, SEMI>;
    case F420_BGR24: return sws_fast420_hi8_kernel<TAPS, F420_BGR24, SEMI>;
    case F420_RGBA:  return sws_fast420_hi8_kernel<TAPS, F420_RGBA, SEMI>;
    case F420_BGRA:  return sws_fast420_hi8_kernel<TAPS, F420_BGRA, SEMI>;
    case F420_ARGB:  return sws_fast420_hi8_kernel<TAPS, F420_ARGB, SEMI>;
    default:         return sws_fast420_hi8_kernel<TAPS, F420_ABGR, SEMI>;
    }
}

static fasthi8_kernel_t pick_fasthi8(int taps, int fmt, bool semi)
{
    if (semi)
        return taps == 4 ? pick_fasthi8_fmt<4, true>(fmt) : taps == 6 ? pick_fasthi8_fmt<6, true>(fmt)
                                                                      : pick_fasthi8_fmt<8, true>(fmt);
    return taps == 4 ? pick_fasthi8_fmt<4, false>(fmt) : taps == 6 ? pick_fasthi8_fmt<6, false>(fmt)
                                                                   : pick_fasthi8_fmt<8, false>(fmt);
}

static int fasthi8_setup(SwsCudaState *st, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->fasthi8_ok = 0;
    /* planar 9..16-bit, or p010le (semi-planar, samples in the high bits) */
    const bool semi = p->src_layout == SWSC_SRC_NV12;
    if ((p->src_layout != SWSC_SRC_PLANAR && !semi) || p->src_bits <= 8 || p->src_bits > 16 || p->inter_bits != 15 ||
        (!semi && p->src_shift))
        return 0;
    if (p->dst_kind < SWSC_DST_RGB24 || p->dst_kind > SWSC_DST_ABGR || p->full_chr || p->special || p->unscaled_lut)
        return 0;
    if (!p->lum_identity || !p->chr_h_identity || p->chr_src_hsub != 1 || p->chr_dst_hsub != 1)
        return 0;
    if (vc->size > 8 || (p->dst_w & 1) || !get_encode_tiled())
        return 0;
    if ((p->dst_kind == SWSC_DST_RGB24 || p->dst_kind == SWSC_DST_BGR24) && (p->dst_w & 3))
        return 0;                     /* the store tensor map counts 32-bit words */
    const int taps = vc->size <= 4 ? 4 : vc->size <= 6 ? 6 : 8;
    /* 24 staged chroma rows cover 4:2:0; sources without vertical chroma subsampling (4:2:2) need 32 + taps */
    int crows = F16_CROWS;
    int ret = fast16_rows(st, vc, taps, crows);
    if (ret == 0) {
        crows = 36;
        ret = fast16_rows(st, vc, taps, crows);
    }
    if (ret <= 0)
        return ret;
    st->fasthi8_crows = crows;
    const int fmt = fast420_fmt(p->dst_kind);
    CUDA_OK(set_max_smem((const void *)pick_fasthi8(taps, fmt, semi), (size_t)(H8_SMEM(fmt >= F420_RGBA ? 4 : 3, crows))));
    st->fasthi8_ok = 1;
    st->fast16_taps = taps;
    st->kernel_name = "fast420_hi8_tma";
    return 0;
}

static int fasthi8_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                          const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                          const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    if (!st->fasthi8_ok || (st->disabled & 128) || (y0 % F16_TH) || y1 <= y0 || y1 > p->dst_h)
        return 0;
    const bool semi = p->src_layout == SWSC_SRC_NV12;
    for (int i = 0; i < (semi ? 2 : 3); i++)
        if (!src[i] || !aligned16(src[i]) || (src_stride[i] & 15) || src_stride[i] <= 0 ||
            (nb_frames > 1 && (src_fstride[i] & 15 || src_fstride[i] <= 0)))
            return 0;
    if (!aligned16(dst[0]) || (dst_stride[0] & 15) || dst_stride[0] <= 0 ||
        (nb_frames > 1 && (dst_fstride[0] & 15 || dst_fstride[0] <= 0)))
        return 0;
    const int fmt = fast420_fmt(p->dst_kind);
    const int bpp = fmt >= F420_RGBA ? 4 : 3;
    const int crows = st->fasthi8_crows;
    CUtensorMap my, mu, mv, mo;
    const uint64_t fs_y = nb_frames > 1 ? src_fstride[0] : (uint64_t)src_stride[0] * p->src_h;
    const uint64_t fs_u = nb_frames > 1 ? src_fstride[1] : (uint64_t)src_stride[1] * p->chr_src_h;
    const uint64_t fs_o = nb_frames > 1 ? dst_fstride[0] : (uint64_t)dst_stride[0] * p->dst_h;
    int ret;
    if ((ret = make_map_3d(&my, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[0], p->src_w, p->src_h, nb_frames,
                           src_stride[0], fs_y, F16_TW, F16_TH)) < 0 ||
        (ret = make_map_3d(&mo, CU_TENSOR_MAP_DATA_TYPE_UINT32, dst[0], (uint64_t)p->dst_w * bpp / 4, y1,
                           nb_frames, dst_stride[0], fs_o, F16_TW * bpp / 4, F16_TH / F420_CWARPS)) < 0)
        return ret;
    if (semi) {
        if ((ret = make_map_3d(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[1], 2 * (uint64_t)p->chr_src_w, p->chr_src_h,
                               nb_frames, src_stride[1], fs_u, F16_TW, crows)) < 0)
            return ret;
        mv = mu;
    } else {
        const uint64_t fs_v = nb_frames > 1 ? src_fstride[2] : (uint64_t)src_stride[2] * p->chr_src_h;
        if ((ret = make_map_3d(&mu, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[1], p->chr_src_w, p->chr_src_h, nb_frames,
                               src_stride[1], fs_u, F16_TW / 2, crows)) < 0 ||
            (ret = make_map_3d(&mv, CU_TENSOR_MAP_DATA_TYPE_UINT16, src[2], p->chr_src_w, p->chr_src_h, nb_frames,
                               src_stride[2], fs_v, F16_TW / 2, crows)) < 0)
            return ret;
    }
    FastHi8Args a;
    memset(&a, 0, sizeof(a));
    a.tiles_x = (p->dst_w + F16_TW - 1) / F16_TW;
    a.tiles_y = (y1 - y0 + F16_TH - 1) / F16_TH;
    a.ty_first = y0 / F16_TH;
    a.frames = nb_frames;
    a.dst_h = p->dst_h;
    a.sdown = p->src_bits - 1;
    a.sshift = p->src_shift;
    a.crows = crows;
    a.cy = p->rgb.cy; a.yb = p->rgb.yb;
    a.crv = p->rgb.crv; a.cbu = p->rgb.cbu; a.cgu = p->rgb.cgu; a.cgv = p->rgb.cgv;
    a.base_r = p->rgb.base_r; a.base_g = p->rgb.base_g; a.base_b = p->rgb.base_b;
    a.rows = st->d_fast16_rows;
    const long long total = (long long)a.tiles_x * a.tiles_y * nb_frames;
    const int grid = (int)(total < (long long)st->num_sms * F420_CTAS_PER_SM ? total : (long long)st->num_sms * F420_CTAS_PER_SM);
    pick_fasthi8(st->fast16_taps, fmt, semi)<<<grid, F420_THREADS, H8_SMEM(bpp, crows), stream>>>(my, mu, mv, mo, a);
    st->kernel_name = "fast420_hi8_tma";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}


/* ---------------------------------------------------------------- scale8 host side */

typedef void (*scale8_kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Scale8Args);
static scale8_kernel_t pick_scale8(int fs4, int rgbk, bool mma, int srck, bool wide = false)
{
    /* rgbk: 0 planar output, 1 packed RGB with shared chroma, 2 packed RGB with full chroma, 3 16-bit planar output
     * over 19-bit lines.
     * `wide`: the tile's shared memory admits at most three CTAs per SM: take the 72-register build */
    if (wide && mma && !rgbk && srck == S8_SRC_U8 && fs4 <= 2)
        return fs4 == 1 ? sws_scale8_kernel<1, 0, true, S8_SRC_U8, 3> : sws_scale8_kernel<2, 0, true, S8_SRC_U8, 3>;
#define S8_PICK(R, M, K) (fs4 == 1 ? sws_scale8_kernel<1, R, M, K> : fs4 == 2 ? sws_scale8_kernel<2, R, M, K> \
                          : fs4 == 4 ? sws_scale8_kernel<4, R, M, K> : sws_scale8_kernel<8, R, M, K>)
#define S8_PICK_MMA(R) (fs4 == 1 ? sws_scale8_kernel<1, R, true, S8_SRC_U8> : fs4 == 2 ? sws_scale8_kernel<2, R, true, S8_SRC_U8> \
                        : sws_scale8_kernel<4, R, true, S8_SRC_U8>)
#define S8_PICK_K(M, K) (rgbk == 2 ? S8_PICK(2, M, K) : rgbk == 1 ? S8_PICK(1, M, K) : S8_PICK(0, M, K))
    if (rgbk == 3)              /* 19-bit lines into 16-bit planar destinations: dot-product horizontal stage only */
        return srck == S8_SRC_U16 ? S8_PICK(3, false, S8_SRC_U16) : srck == S8_SRC_RGB ? S8_PICK(3, false, S8_SRC_RGB)
               : srck == S8_SRC_P010 ? S8_PICK(3, false, S8_SRC_P010) : S8_PICK(3, false, S8_SRC_U8);
    if (srck == S8_SRC_RGB)     /* packed 8-bit RGB sources: reader stage + IDP.2A horizontal stage */
        return S8_PICK_K(false, S8_SRC_RGB);
    if (srck == S8_SRC_U16)     /* 9..16-bit planar sources: IDP.2A horizontal stage */
        return S8_PICK_K(false, S8_SRC_U16);
    if (srck == S8_SRC_P010)    /* p010le sources */
        return rgbk == 3 ? S8_PICK(3, false, S8_SRC_P010) : S8_PICK_K(false, S8_SRC_P010);
    if (mma)                    /* fs4 = K steps of the tensor-pipe horizontal stage */
        return rgbk == 2 ? S8_PICK_MMA(2) : rgbk == 1 ? S8_PICK_MMA(1) : S8_PICK_MMA(0);
    return S8_PICK_K(false, S8_SRC_U8);
#undef S8_PICK
#undef S8_PICK_MMA
#undef S8_PICK_K
}

/* ---- tensor-pipe horizontal stage: per group of 8 output columns, the K window and the banded B fragments ----
 * Window of group G of a tile starting at column x0: first sample = min pos of the group rounded down to a
 * 16-byte boundary of the staged row (the tile's rows start at a0 = pos[x0] & ~15).  `inter` = interleaved
 * nv12 / nv21 chroma: two bytes per sample, and the kernel cuts the planes out of the fragments with PRMT, which
 * reorders the K axis as s8_mma_perm() says.  Returns the K steps needed (0: not expressible). */
static int s8_mma_perm(int k, bool inter)
{
    if (!inter)
        return k;
    const int half = k >> 4, t = (k & 15) >> 2, j = k & 3;
    static const int base[4] = { 0, 1, 8, 9 };
    return 16 * half + 2 * t + base[j];
}

static void s8_mma_window(const SwsFirBank *b, int x0, int G, bool inter, int *wstart, int *wend)
{
    const int xa = x0 + 8 * G;
    int lo = INT32_MAX, hi = 0;
    for (int x = xa; x < xa + 8 && x < b->len; x++) {
        if (b->pos[x] < lo) lo = b->pos[x];
        if (b->pos[x] + b->size > hi) hi = b->pos[x] + b->size;
    }
    if (lo == INT32_MAX)
        lo = hi = b->pos[x0];
    if (lo < 0)
        lo = 0;
    *wstart = lo & ~(inter ? 7 : 15);
    *wend = hi;
}

static int s8_mma_ksteps(const SwsFirBank *b, int tile_cols, bool inter)
{
    int ks = 1;
    for (int x0 = 0; x0 < b->len; x0 += tile_cols)
        for (int G = 0; G < tile_cols / 8; G++) {
            int ws, we;
            s8_mma_window(b, x0, G, inter, &ws, &we);
            const int need = (we - ws + 31) / 32;
            if (need > ks) ks = need;
        }
    return ks;
}

/* staged samples per row so that every window of every tile is inside the TMA box */
static int s8_mma_seg(const SwsFirBank *b, int tile_cols, bool inter, int ks)
{
    int worst = 16;
    for (int x0 = 0; x0 < b->len; x0 += tile_cols) {
        const int a0 = b->pos[x0] & ~15;
        for (int G = 0; G < tile_cols / 8; G++) {
            int ws, we;
            s8_mma_window(b, x0, G, inter, &ws, &we);
            if (ws < a0)
                return -1;
            if (ws - a0 + 32 * ks > worst) worst = ws - a0 + 32 * ks;
        }
    }
    return worst;
}

static void s8_mma_tables(const SwsFirBank *b, int tile_cols, bool inter, int ks, int *goff, uint32_t *B)
{
    const int gpt = tile_cols / 8;
    const int tiles = (b->len + tile_cols - 1) / tile_cols;
    for (int tx = 0; tx < tiles; tx++) {
        const int x0 = tx * tile_cols, a0 = b->pos[x0] & ~15;
        for (int G = 0; G < gpt; G++) {
            int ws, we;
            s8_mma_window(b, x0, G, inter, &ws, &we);
            const int gi = tx * gpt + G;
            goff[gi] = (ws - a0) * (inter ? 2 : 1);
            uint32_t *out = B + (size_t)gi * ks * 4 * 32;
            for (int step = 0; step < ks; step++)
                for (int lane = 0; lane < 32; lane++) {
                    const int t = lane & 3, n = lane >> 2, x = x0 + 8 * G + n;
                    uint32_t r[4] = { 0, 0, 0, 0 };
                    for (int reg = 0; reg < 2; reg++)
                        for (int j = 0; j < 4; j++) {
                            const int smp = ws + 32 * step + s8_mma_perm(16 * reg + 4 * t + j, inter);
                            int c = 0;
                            if (x < b->len) {
                                const int tap = smp - b->pos[x];
                                if (tap >= 0 && tap < b->size)
                                    c = b->coef[(size_t)x * b->size + tap];
                            }
                            r[reg] |= (uint32_t)(c & 0xFF) << (8 * j);
                            r[2 + reg] |= (uint32_t)((c >> 8) & 0xFF) << (8 * j);
                        }
                    for (int q = 0; q < 4; q++)
                        out[(step * 4 + q) * 32 + lane] = r[q];
                }
        }
    }
}

/* pack one horizontal bank: per output, fs4 words of low bytes and fs4 words of high bytes */
static void s8_pack_h(const SwsFirBank *b, int fs4, uint32_t *cl, uint32_t *ch)
{
    for (int x = 0; x < b->len; x++)
        for (int k = 0; k < fs4; k++) {
            uint32_t l = 0, h = 0;
            for (int t = 0; t < 4; t++) {
                const int j = 4 * k + t;
                const int c = j < b->size ? b->coef[(size_t)x * b->size + j] : 0;
                l |= (uint32_t)(c & 0xFF) << (8 * t);
                h |= (uint32_t)((c >> 8) & 0xFF) << (8 * t);
            }
            cl[(size_t)x * fs4 + k] = l;
            ch[(size_t)x * fs4 + k] = h;
        }
}

/* vertical bank: even first row, leading zero tap when the true first row is odd; taps 21..40 of a row go to a second
 * record (rows2) whose first row is 20 further down.  Returns the number of records in use (1 or 2), -1: not expressible */
static int s8_pack_v(const SwsFirBank *b, S8VRow *rows, S8VRow *rows2)
{
    int parts = 1;
    for (int y = 0; y < b->len; y++) {
        const int par = b->pos[y] & 1;
        const int n = b->size + par;
        S8VRow *r = &rows[y], *r2 = &rows2[y];
        memset(r, 0, sizeof(*r));
        memset(r2, 0, sizeof(*r2));
        r->pos_even = b->pos[y] - par;
        const int n4 = (n + 3) / 4;
        if (n4 > 2 * S8_VF4 || b->pos[y] < 0)
            return -1;
        r->n4 = n4 < S8_VF4 ? n4 : S8_VF4;
        r2->n4 = n4 - r->n4;
        r2->pos_even = r2->n4 ? r->pos_even + 4 * S8_VF4 : r->pos_even;
        if (r2->n4)
            parts = 2;
        for (int j = 0; j < b->size; j++) {
            const int c = b->coef[(size_t)y * b->size + j];
            int t = j + par;
            S8VRow *d = r;
            if (t >= 4 * S8_VF4) {
                t -= 4 * S8_VF4;
                d = r2;
            }
            d->cl[t >> 2] |= (uint32_t)(c & 0xFF) << (8 * (t & 3));
            d->ch[t >> 2] |= (uint32_t)((c >> 8) & 0xFF) << (8 * (t & 3));
        }
    }
    return parts;
}

/* rows of transposed h-scaled lines any window of th output rows needs; == 2 (mod 4) for bank spread */
static int s8_rows_cap(const S8VRow *rows, const S8VRow *rows2, int n, int th)
{
    int worst = 4, n4 = 0;
    for (int y = 0; y < n; y++)            /* the bank's largest group count stands for every row */
        if (rows[y].n4 + rows2[y].n4 > n4) n4 = rows[y].n4 + rows2[y].n4;
    for (int y = 0; y < n; y++) {
        const int y1 = y + th < n ? y + th : n;
        int lo = INT32_MAX, hi = 0;
        for (int k = y; k < y1; k++) {
            const int pe = rows[k].pos_even & ~1;
            if (pe < lo) lo = pe;
            if (pe + 4 * n4 > hi) hi = pe + 4 * n4;
        }
        if (hi - lo > worst) worst = hi - lo;
    }
    while ((worst & 3) != 2)
        worst++;
    return worst;
}

static int s8_seg_bytes(const SwsFirBank *b, int fs4, int tile_cols)
{
    int worst = 16;
    for (int x0 = 0; x0 < b->len; x0 += tile_cols) {
        const int x1 = x0 + tile_cols - 1 < b->len - 1 ? x0 + tile_cols - 1 : b->len - 1;
        const int a0 = b->pos[x0] & ~15;
        int need = 0;
        for (int x = x0; x <= x1; x++) {      /* positions are monotonic, but be safe */
            if (b->pos[x] < b->pos[x0])
                return -1;
            const int e = ((b->pos[x] - a0) & ~3) + 4 * fs4 + 4;
            if (e > need) need = e;
        }
        need = (need + 15) & ~15;
        if (need > worst) worst = need;
    }
    return worst;
}

/* 16-bit samples: rows start at a multiple of 8 samples, a column reads 2 fs4 + 1 words from its even first sample */
static int s16_seg_bytes(const SwsFirBank *b, int fs4, int tile_cols)
{
    int worst = 16;
    for (int x0 = 0; x0 < b->len; x0 += tile_cols) {
        const int x1 = x0 + tile_cols - 1 < b->len - 1 ? x0 + tile_cols - 1 : b->len - 1;
        const int a0 = b->pos[x0] & ~7;
        int need = 0;
        for (int x = x0; x <= x1; x++) {
            if (b->pos[x] < b->pos[x0])
                return -1;
            const int e = ((b->pos[x] - a0) >> 1) * 4 + 8 * fs4 + 4;
            if (e > need) need = e;
        }
        need = (need + 15) & ~15;
        if (need > worst) worst = need;
    }
    return worst;
}

/* interleaved 16-bit chroma (p010le), bytes per row of ONE plane (a staged row is twice that): rows start at a multiple
 * of 8 chroma positions, a column reads 4 fs4 words (one per chroma position) from its first position */
static int p10_seg_bytes(const SwsFirBank *b, int fs4, int tile_cols)
{
    int worst = 16;
    for (int x0 = 0; x0 < b->len; x0 += tile_cols) {
        const int x1 = x0 + tile_cols - 1 < b->len - 1 ? x0 + tile_cols - 1 : b->len - 1;
        const int a0 = b->pos[x0] & ~7;
        int need = 0;
        for (int x = x0; x <= x1; x++) {
            if (b->pos[x] < b->pos[x0])
                return -1;
            const int e = (b->pos[x] - a0 + 4 * fs4) * 2;
            if (e > need) need = e;
        }
        need = (need + 15) & ~15;
        if (need > worst) worst = need;
    }
    return worst;
}

static bool rgb420_matrix_ok(const SwsCudaPlan *p);
static bool rgb444_matrix_ok(const SwsCudaPlan *p);

/* packed RGB sources: bytes of raw row a tile stages (first pixel = the lower of both banks' first reads, rounded down to
 * 16 pixels), and the bytes per row of the luma / chroma sample buffers the horizontal stage reads (2 fs4 + 1 words from
 * every column's even first sample).  Returns -1 when a bank is not monotonic. */
static int s8_rgb_segs(const SwsFirBank *hl, const SwsFirBank *hc, int hs, int half, int bpp, int fs4, int *seg_raw,
                       int *seg_sy, int *seg_sc)
{
    int raw = 16, sy = 16, sc = 16;
    const int cw = S8_TW >> hs;
    for (int x0 = 0, cx0 = 0; x0 < hl->len; x0 += S8_TW, cx0 += cw) {
        const int x1 = x0 + S8_TW < hl->len ? x0 + S8_TW : hl->len;
        const int c1 = cx0 + cw < hc->len ? cx0 + cw : hc->len;
        int a0 = hl->pos[x0];
        if (cx0 < hc->len && (hc->pos[cx0] << half) < a0)
            a0 = hc->pos[cx0] << half;
        a0 &= ~15;
        int end = 0;
        for (int x = x0; x < x1; x++) {
            if (hl->pos[x] < hl->pos[x0])
                return -1;
            const int off = hl->pos[x] - a0;
            if (hl->pos[x] + hl->size > end) end = hl->pos[x] + hl->size;
            const int need = ((off >> 1) + 2 * fs4 + 1) * 4;
            if (need > sy) sy = need;
        }
        for (int x = cx0; x < c1; x++) {
            if (hc->pos[x] < hc->pos[cx0])
                return -1;
            const int off = hc->pos[x] - (a0 >> half);
            if (((hc->pos[x] + hc->size) << half) > end) end = (hc->pos[x] + hc->size) << half;
            const int need = ((off >> 1) + 2 * fs4 + 1) * 4;
            if (need > sc) sc = need;
        }
        const int bytes = (end - a0) * bpp;
        if (bytes > raw) raw = bytes;
    }
    const int unit = bpp == 3 ? 48 : 16;
    raw = (raw + unit - 1) / unit * unit;
    const int npx = raw / bpp;                  /* pixels the reader converts per row: a multiple of four */
    if (2 * npx > sy) sy = 2 * npx;
    if (2 * (npx >> half) > sc) sc = 2 * (npx >> half);
    *seg_raw = raw;
    *seg_sy = (sy + 15) & ~15;
    *seg_sc = (sc + 15) & ~15;
    return 0;
}

static int scale8_setup(SwsCudaState *st, const SwsFirBank *hl, const SwsFirBank *hc,
                        const SwsFirBank *vl, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->s8_ok = 0;
    const bool rgbs = p->src_layout == SWSC_SRC_RGB;
    /* 19-bit lines (hScale8To19_c / hScale16To19_c, swscale.c:60-97,144-159): 16-bit planar YUV destinations only */
    const bool i19 = p->inter_bits == 19;
    /* (gbrpf32le rides along: always full chroma, always the X form, float stores) */
    const bool rgb48 = p->dst_kind == SWSC_DST_RGB48 || p->dst_kind == SWSC_DST_BGR48 ||
                       (p->dst_kind == SWSC_DST_GBRP && p->full_chr && !p->dst_alpha);
    /* grayf32le: the 16-bit luma writer with a float store, no chroma stages at all */
    const bool grayf = p->dst_kind == SWSC_DST_PLANARF32 && !p->dst_has_chroma;
    if (i19 && ((p->dst_kind != SWSC_DST_PLANAR16 && !rgb48 && !grayf) || p->dst_shift))
        return 0;
    /* rgb48le / bgr48le: yuv2rgba64_{X,2,1} (one chroma sample per pixel pair) or, with SWS_FULL_CHR_H_INT, the _full_ forms */
    if (rgb48 && (!i19 || p->chr_dst_hsub != (p->full_chr ? 0 : 1) || p->chr_dst_vsub != 0 || p->special || p->unscaled_lut ||
                  !p->has_chroma))
        return 0;
    if ((p->inter_bits != 15 && !i19) || (p->src_layout > SWSC_SRC_NV21 && !rgbs) || p->src_alpha || p->dst_alpha)
        return 0;
    /* packed 8-bit RGB sources: samples the readers keep inside 14 bits (checked again at every launch: the matrix can
     * change), luma and chroma out of the same rows */
    if (rgbs && (p->h_shift != (i19 ? 9 : 13) || p->chr_src_h != p->src_h || (p->src_bpp != 3 && p->src_bpp != 4) ||
                 !(p->src_rgb_half ? rgb420_matrix_ok(p) : rgb444_matrix_ok(p))))
        return 0;
    /* sources: 8-bit planar / nv12 / nv21, or 9..16-bit little-endian planar (hScale16To15_c) */
    const bool s16 = !rgbs && p->src_bits > 8;
    /* p010le (p010LEToY/UV_c: the 16-bit containers >> 6, chroma interleaved U first) */
    const bool p10 = s16 && p->src_layout == SWSC_SRC_NV12 && p->src_shift == 6 && p->src_bits == 10;
    const int srck = rgbs ? S8_SRC_RGB : p10 ? S8_SRC_P010 : s16 ? S8_SRC_U16 : S8_SRC_U8;
    if (s16 && !p10 && (p->src_layout != SWSC_SRC_PLANAR || p->src_shift || p->src_bits > 16))
        return 0;
    /* destinations: 8-bit planar / semi-planar YUV, 9..14-bit planar YUV, or packed 8-bit RGB with one chroma
     * sample per pixel pair */
    const bool rgb = (p->dst_kind >= SWSC_DST_RGB24 && p->dst_kind <= SWSC_DST_ABGR) ||
                     (p->dst_kind >= SWSC_DST_RGB565 && p->dst_kind <= SWSC_DST_BGR555);
    /* one chroma sample per pixel pair (yuv2rgb_{X,2,1} + yuv2rgb_write), or SWS_FULL_CHR_H_INT: one per pixel and the
     * arithmetic colour step (yuv2rgb_full_{X,2,1} + yuv2rgb_write_full) */
    if (rgb && (p->chr_dst_hsub != (p->full_chr ? 0 : 1) || p->chr_dst_vsub != 0 || p->special || p->unscaled_lut ||
                !p->has_chroma || (p->full_chr && p->dst_kind > SWSC_DST_ABGR)))
        return 0;
    if (!rgb && !i19 && p->dst_kind != SWSC_DST_PLANAR8 && p->dst_kind != SWSC_DST_NV12 && p->dst_kind != SWSC_DST_NV21 &&
        !(p->dst_kind == SWSC_DST_PLANARN && p->dst_bits >= 9 && p->dst_bits <= 14 && !p->dst_shift) &&
        !(p->dst_kind == SWSC_DST_P010 && p->dst_bits == 10 && p->dst_shift == 6))
        return 0;
    if ((!p->has_chroma || !p->dst_has_chroma) && !grayf)
        return 0;
    if (p->special || p->unscaled_lut)
        return 0;
    if (hl->size > 32 || hc->size > 32 || vl->size > 38 || vc->size > 38)
        return 0;                     /* horizontal: 8 tap groups of four; vertical: two records of 20 taps (39 + parity pad) */
    if (hl->size > p->src_w || hc->size > p->chr_src_w || p->chr_dst_hsub > 1 || p->chr_dst_vsub > 1)
        return 0;
    const int fs = hl->size > hc->size ? hl->size : hc->size;
    int fs4 = fs <= 4 ? 1 : fs <= 8 ? 2 : fs <= 16 ? 4 : 8;
    const int cw = S8_TW >> p->chr_dst_hsub;

    /* [0, len): first records, [len, 2 len): second records */
    S8VRow *hvl = (S8VRow *)malloc(sizeof(S8VRow) * vl->len * 2);
    S8VRow *hvc = (S8VRow *)malloc(sizeof(S8VRow) * vc->len * 2);
    int ret = 0, lparts = 1, cparts = 1;
    if (!hvl || !hvc || (lparts = s8_pack_v(vl, hvl, hvl + vl->len)) < 0 || (cparts = s8_pack_v(vc, hvc, hvc + vc->len)) < 0)
        ret = 1;
    if ((lparts == 2 || cparts == 2) && !i19)
        fs4 = 8;                      /* only the eight-group variants are compiled with the second vertical record */
    if (!ret && rgb && p->full_chr && vl->size == 1 && vc->size == 2) {
        /* yuv2rgb_full_1 with two chroma taps (vscale.c:138-143) blends chroma without the rounding bias:
         * flagged in bit 0 of the chroma row */
        for (int y = 0; y < vc->len; y++) {
            const int c0 = vc->coef[2 * y], c1 = vc->coef[2 * y + 1];
            if (c0 + c1 == 4096 && (unsigned)c1 <= 4096u)
                hvc[y].pos_even |= 1;
        }
    }
    if (!ret && rgb && vl->size == 2 && vc->size == 2) {
        /* rows the reference hands to yuv2packed2 (both filters 2-tap bilinear) round without a bias
         * (vscale.c:148-163, output.c:1861-1864): flagged in bit 0 of the even first row */
        for (int y = 0; y < vl->len && y < vc->len; y++) {
            const int l0 = vl->coef[2 * y], l1 = vl->coef[2 * y + 1], c0 = vc->coef[2 * y], c1 = vc->coef[2 * y + 1];
            if (l0 + l1 == 4096 && (unsigned)l1 <= 4096u && c0 + c1 == 4096 && (unsigned)c1 <= 4096u)
                hvl[y].pos_even |= 1;
        }
    }
    int seg_l = ret ? -1 : s16 ? s16_seg_bytes(hl, fs4, S8_TW) : s8_seg_bytes(hl, fs4, S8_TW);
    int seg_c = ret ? -1 : p10 ? p10_seg_bytes(hc, fs4, cw) : s16 ? s16_seg_bytes(hc, fs4, cw) : s8_seg_bytes(hc, fs4, cw);
    int seg_sy = 0, seg_sc = 0;
    if (!ret && rgbs) {
        if (s8_rgb_segs(hl, hc, p->chr_dst_hsub, p->src_rgb_half, p->src_bpp, fs4, &seg_l, &seg_sy, &seg_sc) < 0)
            seg_l = -1;
        seg_c = 0;
    }
    /* tensor-pipe horizontal stage when every 8-column window fits K <= 128 samples (SWS_B200_DISABLE=s8mma: A/B) */
    const bool inter = p->src_layout != SWSC_SRC_PLANAR;
    int mma = 0, ks = 0;
    /* what the MMA variants are compiled for: no range conversion, 8-bit output, vertical banks of one record */
    const bool plain = !p->range_mode && p->dst_kind != SWSC_DST_PLANARN && p->dst_kind != SWSC_DST_P010 && !i19 &&
                       lparts == 1 && cparts == 1;
    if (!ret && !s16 && !rgbs && plain && seg_l >= 0 && seg_c >= 0 &&
        !(getenv("SWS_B200_DISABLE") && strstr(getenv("SWS_B200_DISABLE"), "s8mma"))) {
        const int kl = s8_mma_ksteps(hl, S8_TW, false), kc = s8_mma_ksteps(hc, cw, inter);
        ks = kl > kc ? kl : kc;
        ks = ks <= 1 ? 1 : ks == 2 ? 2 : ks <= 4 ? 4 : 0;
        if (ks) {
            int ml = s8_mma_seg(hl, S8_TW, false, ks), mc = s8_mma_seg(hc, cw, inter, ks);
            if (ml > 0 && mc > 0) {
                /* row pitch = odd multiple of 16 bytes: the eight rows of an ldmatrix tile hit eight bank groups */
                ml = (ml + 15) & ~15;
                if (!((ml >> 4) & 1)) ml += 16;
                if (inter) {
                    mc = (mc + 7) & ~7;
                    if (!((mc >> 3) & 1)) mc += 8;       /* 2 * mc bytes per row */
                } else {
                    mc = (mc + 15) & ~15;
                    if (!((mc >> 4) & 1)) mc += 16;
                }
                seg_l = ml; seg_c = mc;
                mma = 1;
            }
        }
    }
    if (seg_l < 0 || seg_c < 0)
        ret = 1;
    else if (seg_l > 2048 || seg_c > (s16 && !p10 ? 2048 : 1024) || !get_encode_tiled())
        ret = 1;                      /* one ring-slot row is one TMA box row of at most 256 elements of 4 or 8 bytes */
    /* bytes per row of one chroma plane; interleaved rows are 2 seg_c bytes */
    const int rowc = p->src_layout == SWSC_SRC_PLANAR ? seg_c : 2 * seg_c;
    const int elt_shift = (seg_l > 1024 || rowc > 1024) ? 3 : 2;
    int th = 0, nl_cap = 0, nc_cap = 0, slot = 0, stages = 0;
    size_t smem = 0;
    if (!ret) {
        ret = 1;
        /* largest tile that still leaves three CTAs per SM (about 74 KB each) with a ring of two slots; failing that,
         * two CTAs.  SWS_B200_S8_STAGES raises the ring depth up to what the budget leaves: measured on C4 / X1 / X2
         * with 2..5 slots, no difference (the kernel does not wait on TMA), and a persistent grid with the ring
         * running ahead across tiles was slower (profiles/r02/scale8_persistent_r02.txt). */
        size_t budget[2] = { 74 * 1024 + 512, 100 * 1024 };
        const char *th_env = getenv("SWS_B200_S8_TH"), *st_env = getenv("SWS_B200_S8_STAGES"), *kb_env = getenv("SWS_B200_S8_KB");
        const int th_max = th_env && atoi(th_env) >= 2 ? atoi(th_env) : 32;
        const int st_max = st_env && atoi(st_env) >= 2 && atoi(st_env) <= S8_MAX_STAGES ? atoi(st_env) : S8_STAGES;
        if (kb_env && atoi(kb_env) >= 32 && atoi(kb_env) <= 200)
            budget[0] = budget[1] = (size_t)atoi(kb_env) * 1024;
        /* Every (CTAs per SM, tile height) pair that fits is a candidate; the cost of one is the h-scaled samples it
         * computes per output row (the rows of vertical halo are recomputed by the neighbouring tile), with 10 % added
         * for running two CTAs per SM instead of three.  Conversions whose staging is small (every BASELINE
         * configuration) end up with 32-row tiles at three CTAs as before; wide raw rows (packed RGB sources) trade
         * occupancy for taller tiles instead of shrinking to 8 rows. */
        budget[1] = kb_env ? budget[1] : 112 * 1024;
        double best = 0;
        int best_th = 0, best_pass = 0;
        slot = (S8_ROWS * (seg_l > 2 * seg_c ? seg_l : 2 * seg_c) + 127) & ~127;
        if (rgb && S8_STAGES * slot < 8 * 512)
            slot = 8 * 512 / S8_STAGES;               /* the idle ring stages the packed RGB rows of the 8 warps */
        for (int pass = 0; pass < 2 && slot <= 48 * 1024; pass++)
            for (int t = th_max; t >= 2; t >>= 1) {
                const int cth = t >> p->chr_dst_vsub ? t >> p->chr_dst_vsub : 1;
                const int nl = s8_rows_cap(hvl, hvl + vl->len, vl->len, t), nc = s8_rows_cap(hvc, hvc + vc->len, vc->len, cth);
                /* 15-bit lines: two rows per word; 19-bit lines: one row per word, columns nl + 1 words apart */
                const size_t lines = (i19 ? ((size_t)S8_TW * (nl + 1) + 2 * (size_t)cw * (nc + 1)) * 4
                                          : ((size_t)S8_TW * nl + 2 * (size_t)cw * nc) * 2) +
                                     (size_t)S8_ROWS * (seg_sy + 2 * seg_sc);     /* + the RGB readers' sample rows */
                if (lines + 2 * (size_t)slot > budget[pass])
                    continue;
                const double cost = ((double)S8_TW * nl + 2.0 * cw * nc) / t * (pass ? 1.1 : 1.0);
                if (!best_th || cost < best) {
                    best = cost; best_th = t; best_pass = pass;
                }
            }
        if (best_th) {
            th = best_th;
            const int cth = th >> p->chr_dst_vsub ? th >> p->chr_dst_vsub : 1;
            nl_cap = s8_rows_cap(hvl, hvl + vl->len, vl->len, th);
            nc_cap = s8_rows_cap(hvc, hvc + vc->len, vc->len, cth);
            const size_t lines = (i19 ? ((size_t)S8_TW * (nl_cap + 1) + 2 * (size_t)cw * (nc_cap + 1)) * 4
                                      : ((size_t)S8_TW * nl_cap + 2 * (size_t)cw * nc_cap) * 2) + (size_t)S8_ROWS * (seg_sy + 2 * seg_sc);
            stages = (int)((budget[best_pass] - lines) / slot);
            if (stages > st_max) stages = st_max;
            smem = lines + (size_t)stages * slot;
            ret = 0;
        }
    }
    if (ret) {
        free(hvl); free(hvc);
        return 0;                     /* not eligible: the generic kernel handles it */
    }
    /* one device allocation: positions, packed H taps, V rows */
    size_t off = 0, o_hlp, o_hcp, o_hlcl, o_hlch, o_hccl, o_hcch, o_vl, o_vc;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    o_hlp = take(sizeof(int) * hl->len); o_hcp = take(sizeof(int) * hc->len);
    o_hlcl = take(4 * (size_t)hl->len * fs4); o_hlch = take(4 * (size_t)hl->len * fs4);
    o_hccl = take(4 * (size_t)hc->len * fs4); o_hcch = take(4 * (size_t)hc->len * fs4);
    o_vl = take(sizeof(S8VRow) * vl->len * 2); o_vc = take(sizeof(S8VRow) * vc->len * 2);
    const int ngl = (hl->len + S8_TW - 1) / S8_TW * (S8_TW / 8) + 2, ngc = (hc->len + cw - 1) / cw * (cw / 8) + 2;
    size_t o_gl = 0, o_gc = 0, o_bl = 0, o_bc = 0;
    if (mma) {
        o_gl = take(sizeof(int) * ngl); o_gc = take(sizeof(int) * ngc);
        o_bl = take((size_t)ngl * ks * 4 * 32 * 4); o_bc = take((size_t)ngc * ks * 4 * 32 * 4);
    }
    uint8_t *host = (uint8_t *)calloc(1, off);
    if (!host) {
        free(hvl); free(hvc);
        return AVERROR(ENOMEM);
    }
    memcpy(host + o_hlp, hl->pos, sizeof(int) * hl->len);
    memcpy(host + o_hcp, hc->pos, sizeof(int) * hc->len);
    s8_pack_h(hl, fs4, (uint32_t *)(host + o_hlcl), (uint32_t *)(host + o_hlch));
    s8_pack_h(hc, fs4, (uint32_t *)(host + o_hccl), (uint32_t *)(host + o_hcch));
    st->s8_vl_n4 = st->s8_vc_n4 = 0;
    for (int y = 0; y < vl->len; y++)
        if (hvl[y].n4 > st->s8_vl_n4) st->s8_vl_n4 = hvl[y].n4;
    for (int y = 0; y < vc->len; y++)
        if (hvc[y].n4 > st->s8_vc_n4) st->s8_vc_n4 = hvc[y].n4;
    memcpy(host + o_vl, hvl, sizeof(S8VRow) * vl->len * 2);
    memcpy(host + o_vc, hvc, sizeof(S8VRow) * vc->len * 2);
    if (mma) {
        s8_mma_tables(hl, S8_TW, false, ks, (int *)(host + o_gl), (uint32_t *)(host + o_bl));
        s8_mma_tables(hc, cw, inter, ks, (int *)(host + o_gc), (uint32_t *)(host + o_bc));
    }
    free(hvl); free(hvc);
    cudaError_t e = cudaMalloc(&st->s8_tables, off);
    if (e == cudaSuccess)
        e = cudaMemcpy(st->s8_tables, host, off, cudaMemcpyHostToDevice);
    free(host);
    CUDA_OK(e);
    uint8_t *t = (uint8_t *)st->s8_tables;
    st->s8_hl_pos = (int *)(t + o_hlp); st->s8_hc_pos = (int *)(t + o_hcp);
    st->s8_hl_cl = (uint32_t *)(t + o_hlcl); st->s8_hl_ch = (uint32_t *)(t + o_hlch);
    st->s8_hc_cl = (uint32_t *)(t + o_hccl); st->s8_hc_ch = (uint32_t *)(t + o_hcch);
    st->s8_vl = (S8VRow *)(t + o_vl); st->s8_vc = (S8VRow *)(t + o_vc);
    st->s8_vl2 = lparts == 2 ? st->s8_vl + vl->len : nullptr;
    st->s8_vc2 = cparts == 2 ? st->s8_vc + vc->len : nullptr;
    st->s8_fs4 = mma ? ks : fs4; st->s8_tile_h = th; st->s8_nl_cap = nl_cap; st->s8_nc_cap = nc_cap;
    st->s8_seg_l = seg_l; st->s8_seg_c = seg_c; st->s8_smem = smem; st->s8_slot = slot; st->s8_stages = stages; st->s8_srck = srck; st->s8_elt_shift = elt_shift;
    st->s8_wide = 4 * (smem + 1024) > 227 * 1024;
    st->s8_seg_sy = seg_sy; st->s8_seg_sc = seg_sc;
    st->s8_mma = mma;
    if (mma) {
        st->s8_hl_goff = (int *)(t + o_gl); st->s8_hc_goff = (int *)(t + o_gc);
        st->s8_hl_B = (uint32_t *)(t + o_bl); st->s8_hc_B = (uint32_t *)(t + o_bc);
    }
    CUDA_OK(set_max_smem((const void *)pick_scale8(st->s8_fs4, i19 ? 3 : rgb ? (p->full_chr ? 2 : 1) : 0, mma, srck, st->s8_wide), (size_t)((int)smem)));
    st->s8_ok = 1;
    if (getenv("SWS_B200_DEBUG"))
        fprintf(stderr, "[swscaler-b200] scale8: %s fs4=%d tile_h=%d nl_cap=%d nc_cap=%d seg_l=%d seg_c=%d slot=%d stages=%d smem=%zu\n",
                rgbs ? "rgb" : i19 ? "i19" : s16 ? "dp2a16" : mma ? "mma" : "dp4a", st->s8_fs4, th, nl_cap, nc_cap, seg_l, seg_c, slot, stages, smem);
    if (!st->fast_ok && !st->fast16_ok)
        st->kernel_name = rgbs ? (i19 ? "scale_rgb_i19" : "scale_rgb_dp2a") : i19 ? (s16 ? "scale16_i19" : "scale8_i19") : s16 ? "scale16_dp2a" : mma ? "scale8_mma" : "scale8_dp4a";
    return 0;
}


static int scale8_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                         const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                         const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    /* a range change after init (sws_setColorspaceDetails) finds the tensor-pipe variant compiled without it */
    if (!st->s8_ok || (st->disabled & 4) || (st->s8_mma && p->range_mode))
        return 0;
    const bool rgbs = st->s8_srck == S8_SRC_RGB;
    if (rgbs && !(p->src_rgb_half ? rgb420_matrix_ok(p) : rgb444_matrix_ok(p)))
        return 0;                     /* sws_setColorspaceDetails() installed a matrix the 14-bit readers cannot take */
    const bool luma_only = !p->dst_has_chroma;        /* grayf32le: the chroma planes are never read (nor uploaded) */
    const int nsrc = rgbs || luma_only ? 1 : p->src_layout == SWSC_SRC_PLANAR ? 3 : 2;
    for (int i = 0; i < nsrc; i++)
        if (!src[i] || !aligned16(src[i]) || (src_stride[i] & 15) || src_stride[i] < 16 ||
            (nb_frames > 1 && (src_fstride[i] & 15)))
            return 0;
    /* source planes as tensors of 4- or 8-byte elements {row elements, rows, frames}: a ring slot is one box of S8_ROWS rows */
    CUtensorMap my, mu, mv;
    const bool planar = p->src_layout == SWSC_SRC_PLANAR;
    const int es = st->s8_elt_shift, bps = st->s8_srck == S8_SRC_U16 || st->s8_srck == S8_SRC_P010 ? 1 : 0;
    const CUtensorMapDataType edt = es == 3 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    const uint64_t emask = (1u << es) - 1;
    const uint64_t fs_y = nb_frames > 1 ? src_fstride[0] : (uint64_t)src_stride[0] * p->src_h;
    const uint64_t fs_u = nb_frames > 1 ? src_fstride[1] : (uint64_t)src_stride[1] * p->chr_src_h;
    const uint64_t ybytes = rgbs ? (uint64_t)p->src_w * p->src_bpp : (uint64_t)p->src_w << bps;
    const uint64_t cbytes = (planar ? (uint64_t)p->chr_src_w : 2 * (uint64_t)p->chr_src_w) << bps;
    int ret;
    if ((ret = make_map_3d(&my, edt, src[0], (ybytes + emask) >> es, p->src_h,
                           nb_frames, src_stride[0], fs_y, st->s8_seg_l >> es, S8_ROWS)) < 0)
        return ret;
    mu = my;
    if (!rgbs && !luma_only && (ret = make_map_3d(&mu, edt, src[1], (cbytes + emask) >> es, p->chr_src_h,
                                    nb_frames, src_stride[1], fs_u, (planar ? st->s8_seg_c : 2 * st->s8_seg_c) >> es,
                                    S8_ROWS)) < 0)
        return ret;
    mv = mu;
    if (planar && !rgbs && !luma_only) {
        const uint64_t fs_v = nb_frames > 1 ? src_fstride[2] : (uint64_t)src_stride[2] * p->chr_src_h;
        if ((ret = make_map_3d(&mv, edt, src[2], (cbytes + emask) >> es, p->chr_src_h,
                               nb_frames, src_stride[2], fs_v, st->s8_seg_c >> es, S8_ROWS)) < 0)
            return ret;
    }
    Scale8Args a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 3; i++) {
        a.dst[i] = dst[i];
        a.dst_stride[i] = dst_stride[i];
        a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
    }
    a.src_w = p->src_w; a.src_h = p->src_h; a.chr_src_w = p->chr_src_w; a.chr_src_h = p->chr_src_h;
    a.dst_w = p->dst_w; a.dst_h = p->dst_h; a.chr_dst_w = p->chr_dst_w; a.chr_dst_h = p->chr_dst_h;
    a.hs = p->chr_dst_hsub; a.vs = p->chr_dst_vsub;
    a.src_layout = p->src_layout; a.dst_kind = p->dst_kind;
    a.y0 = y0; a.y1 = y1; a.tile_h = st->s8_tile_h;
    a.nl_cap = st->s8_nl_cap; a.nc_cap = st->s8_nc_cap; a.seg_l = st->s8_seg_l; a.seg_c = st->s8_seg_c;
    a.slot_bytes = st->s8_slot;
    a.stages = st->s8_stages;
    a.bps = bps; a.elt_shift = es; a.h_shift = p->h_shift;
    a.range_mode = p->range_mode;
    a.lum_rc_coeff = (int)p->lum_rc_coeff; a.lum_rc_offset = (int)p->lum_rc_offset;
    a.chr_rc_coeff = (int)p->chr_rc_coeff; a.chr_rc_offset = (int)p->chr_rc_offset;
    a.out_bits = p->dst_kind == SWSC_DST_PLANARN || p->dst_kind == SWSC_DST_P010 ? p->dst_bits : 8;
    a.out_lshift = p->dst_kind == SWSC_DST_P010 ? p->dst_shift : 0;
    a.no_chroma = !p->dst_has_chroma;
    if (rgbs) {
        /* matrix rows as 16-bit pairs in the byte order of a pixel word (unused bytes get a zero coefficient) */
        int ky[4] = { 0, 0, 0, 0 }, ku[4] = { 0, 0, 0, 0 }, kv[4] = { 0, 0, 0, 0 };
        ky[p->src_ro] = p->rgb2yuv[0]; ky[p->src_go] = p->rgb2yuv[1]; ky[p->src_bo] = p->rgb2yuv[2];
        ku[p->src_ro] = p->rgb2yuv[3]; ku[p->src_go] = p->rgb2yuv[4]; ku[p->src_bo] = p->rgb2yuv[5];
        kv[p->src_ro] = p->rgb2yuv[6]; kv[p->src_go] = p->rgb2yuv[7]; kv[p->src_bo] = p->rgb2yuv[8];
        auto pair = [](int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); };
        a.ylo = pair(ky[0], ky[1]); a.yhi = pair(ky[2], ky[3]);
        a.ulo = pair(ku[0], ku[1]); a.uhi = pair(ku[2], ku[3]);
        a.vlo = pair(kv[0], kv[1]); a.vhi = pair(kv[2], kv[3]);
        a.src_bpp = p->src_bpp; a.rgb_half = p->src_rgb_half;
        a.seg_sy = st->s8_seg_sy; a.seg_sc = st->s8_seg_sc;
    }
    a.dither_bayer = p->dither_bayer;
    a.vl_n4 = st->s8_vl_n4; a.vc_n4 = st->s8_vc_n4;
    a.cy = p->rgb.cy; a.yb = p->rgb.yb; a.base_r = p->rgb.base_r; a.base_g = p->rgb.base_g; a.base_b = p->rgb.base_b;
    a.crv = p->rgb.crv; a.cgu = p->rgb.cgu; a.cgv = p->rgb.cgv; a.cbu = p->rgb.cbu;
    a.full_chr = p->full_chr; a.y_offset = p->rgb.y_offset; a.y_coeff = p->rgb.y_coeff;
    a.v2r = p->rgb.v2r; a.v2g = p->rgb.v2g; a.u2g = p->rgb.u2g; a.u2b = p->rgb.u2b;
    const bool rgb = (p->dst_kind >= SWSC_DST_RGB24 && p->dst_kind <= SWSC_DST_ABGR) ||
                     (p->dst_kind >= SWSC_DST_RGB565 && p->dst_kind <= SWSC_DST_BGR555);
    const bool i19 = p->inter_bits == 19;
    a.lum_rc_offset64 = p->lum_rc_offset; a.chr_rc_offset64 = p->chr_rc_offset;
    a.vl_coef16 = p->vl_coef; a.vc_coef16 = p->vc_coef; a.vl_pos32 = p->vl_pos; a.vc_pos32 = p->vc_pos;
    a.vl_size = p->vl_size; a.vc_size = p->vc_size;
    a.hl_pos = st->s8_hl_pos; a.hc_pos = st->s8_hc_pos;
    a.hl_cl = st->s8_hl_cl; a.hl_ch = st->s8_hl_ch; a.hc_cl = st->s8_hc_cl; a.hc_ch = st->s8_hc_ch;
    a.vl = st->s8_vl; a.vc = st->s8_vc; a.vl2 = st->s8_vl2; a.vc2 = st->s8_vc2;
    a.hl_goff = st->s8_hl_goff; a.hc_goff = st->s8_hc_goff; a.hl_B = st->s8_hl_B; a.hc_B = st->s8_hc_B;
    dim3 grid((p->dst_w + S8_TW - 1) / S8_TW, (y1 - y0 + st->s8_tile_h - 1) / st->s8_tile_h, nb_frames);
    pick_scale8(st->s8_fs4, i19 ? 3 : rgb ? (p->full_chr ? 2 : 1) : 0, st->s8_mma, st->s8_srck, st->s8_wide)<<<grid, S8_THREADS, st->s8_smem, stream>>>(my, mu, mv, a);
    st->kernel_name = rgbs ? (i19 ? "scale_rgb_i19" : "scale_rgb_dp2a") : i19 ? (st->s8_srck == S8_SRC_U16 || st->s8_srck == S8_SRC_P010 ? "scale16_i19" : "scale8_i19")
                           : st->s8_srck == S8_SRC_U16 || st->s8_srck == S8_SRC_P010 ? "scale16_dp2a" : st->s8_mma ? "scale8_mma" : "scale8_dp4a";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}

typedef void (*generic_kernel_t)(const SwsCudaPlan, const FrameArgs);

static generic_kernel_t pick_generic(const SwsCudaPlan *p)
{
    const bool src16 = p->src_bits > 8, i32 = p->inter_bits == 19;
    if (src16)
        return i32 ? sws_generic_tile_kernel<true, true> : sws_generic_tile_kernel<true, false>;
    return i32 ? sws_generic_tile_kernel<false, true> : sws_generic_tile_kernel<false, false>;
}

extern "C" int ff_b200_cuda_create_on(int device, SwsCudaState **out, SwsCudaPlan *plan,
                                      const SwsFirBank *hl, const SwsFirBank *hc,
                                      const SwsFirBank *vl, const SwsFirBank *vc)
{
    if (device < 0 || device >= sws_cuda_device_count())
        return AVERROR(EINVAL);
    DeviceGuard guard(device);
    return ff_b200_cuda_create(out, plan, hl, hc, vl, vc);
}

extern "C" int ff_b200_cuda_device_of(SwsCudaState *st) { return st->device; }

extern "C" int ff_b200_cuda_create(SwsCudaState **out, SwsCudaPlan *plan,
                                   const SwsFirBank *hl, const SwsFirBank *hc,
                                   const SwsFirBank *vl, const SwsFirBank *vc)
{
    int dev = ff_b200_cuda_probe();
    if (dev < 0)
        return dev;
    SwsCudaState *st = (SwsCudaState *)calloc(1, sizeof(*st));
    if (!st)
        return AVERROR(ENOMEM);
    st->device = dev;
    *out = st;
    CUDA_OK(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    if (plan->special == SWSC_SPECIAL_DEPTHCOPY)
        CUDA_OK(upload_depth_dither());

    /* upload the four banks into one allocation: [coef|pos] x4, 16-byte aligned */
    const SwsFirBank *banks[4] = { hl, hc, vl, vc };
    size_t off[8], total = 0;
    for (int i = 0; i < 4; i++) {
        off[2 * i] = total;
        total += (((size_t)banks[i]->len * banks[i]->size * sizeof(int16_t)) + 15) & ~(size_t)15;
        off[2 * i + 1] = total;
        total += (((size_t)banks[i]->len * sizeof(int32_t)) + 15) & ~(size_t)15;
    }
    CUDA_OK(cudaMalloc(&st->tables, total));
    uint8_t *host = (uint8_t *)calloc(1, total);
    if (!host)
        return AVERROR(ENOMEM);
    for (int i = 0; i < 4; i++) {
        memcpy(host + off[2 * i], banks[i]->coef, (size_t)banks[i]->len * banks[i]->size * sizeof(int16_t));
        memcpy(host + off[2 * i + 1], banks[i]->pos, (size_t)banks[i]->len * sizeof(int32_t));
    }
    cudaError_t e = cudaMemcpy(st->tables, host, total, cudaMemcpyHostToDevice);
    free(host);
    CUDA_OK(e);
    uint8_t *t = (uint8_t *)st->tables;
    plan->hl_coef = (const int16_t *)(t + off[0]); plan->hl_pos = (const int32_t *)(t + off[1]); plan->hl_size = hl->size;
    plan->hc_coef = (const int16_t *)(t + off[2]); plan->hc_pos = (const int32_t *)(t + off[3]); plan->hc_size = hc->size;
    plan->vl_coef = (const int16_t *)(t + off[4]); plan->vl_pos = (const int32_t *)(t + off[5]); plan->vl_size = vl->size;
    plan->vc_coef = (const int16_t *)(t + off[6]); plan->vc_pos = (const int32_t *)(t + off[7]); plan->vc_size = vc->size;
    st->plan = *plan;

    st->h_vl_pos = (int32_t *)malloc(sizeof(int32_t) * vl->len);
    st->h_vc_pos = (int32_t *)malloc(sizeof(int32_t) * vc->len);
    if (!st->h_vl_pos || !st->h_vc_pos)
        return AVERROR(ENOMEM);
    memcpy(st->h_vl_pos, vl->pos, sizeof(int32_t) * vl->len);
    memcpy(st->h_vc_pos, vc->pos, sizeof(int32_t) * vc->len);

    {
        const char *e = getenv("SWS_B200_E2E_MODE"), *b = getenv("SWS_B200_E2E_BANDS");
        st->e2e_mode = e ? atoi(e) : 3;
        st->e2e_bands = b ? atoi(b) : 4;
        if (st->e2e_bands < 1) st->e2e_bands = 1;
        if (st->e2e_bands > 16) st->e2e_bands = 16;
    }
    int ret = plan_tiles(st);
    if (ret < 0)
        return ret;
    st->kernel_name = "generic_tile";
    CUDA_OK(set_max_smem((const void *)pick_generic(&st->plan), (size_t)((int)st->smem_bytes)));
    CUDA_OK(cudaDeviceGetAttribute(&st->num_sms, cudaDevAttrMultiProcessorCount, dev));
    ret = fast420_setup(st, vc);
    if (ret < 0)
        return ret;
    ret = fast16_setup(st, vc);
    if (ret < 0)
        return ret;
    ret = fasthi8_setup(st, vc);
    if (ret < 0)
        return ret;
    ret = scale8_setup(st, hl, hc, vl, vc);
    if (ret < 0)
        return ret;
    ret = rgb420_setup(st, vc);
    if (ret < 0)
        return ret;
    ret = tile15_setup(st, hl, hc, vl, vc);
    if (ret < 0)
        return ret;
    {
        /* SWS_B200_DISABLE=fast420,fast16,scale8,tile15: force the next kernel in the dispatch order (A/B runs, tests) */
        const char *d = getenv("SWS_B200_DISABLE");
        st->disabled = 0;
        if (d) {
            if (strstr(d, "fast420")) st->disabled |= 1;
            if (strstr(d, "fast16"))  st->disabled |= 2;
            if (strstr(d, "scale8"))  st->disabled |= 4;
            if (strstr(d, "copy8"))   st->disabled |= 32;
            if (strstr(d, "rgb420"))  st->disabled |= 64;
            if (strstr(d, "hi8"))     st->disabled |= 128;
            if (strstr(d, "full444")) st->disabled |= 256;
            if (strstr(d, "rgb444"))  st->disabled |= 512;
            if (strstr(d, "tile15"))  st->disabled |= 16;
        }
    }
    return 0;
}

static void free_staging(SwsCudaState *st);

extern "C" void ff_b200_cuda_destroy(SwsCudaState *st)
{
    if (!st)
        return;
    DeviceGuard guard(st->device);
    if (st->stream) {
        cudaStreamSynchronize(st->stream);
        cudaStreamDestroy(st->stream);
    }
    cudaFree(st->tables);
    cudaFree(st->d_fast_rows);
    cudaFree(st->d_fast_rows_narrow);
    cudaFree(st->d_fast16_rows);
    cudaFree(st->s8_tables);
    if (st->s_in) {
        cudaStreamDestroy(st->s_in);
        cudaStreamDestroy(st->s_out);
        for (int k = 0; k < 16; k++) {
            cudaEventDestroy(st->ev_in[k]);
            cudaEventDestroy(st->ev_k[k]);
            cudaEventDestroy(st->ev_out[k]);
        }
    }
    free_staging(st);
    free(st->h_vl_pos);
    free(st->h_vc_pos);
    free(st);
}

extern "C" int ff_b200_cuda_update_plan(SwsCudaState *st, const SwsCudaPlan *plan)
{
    /* keep the device table pointers, refresh the scalar constants */
    SwsCudaPlan n = *plan;
    n.hl_coef = st->plan.hl_coef; n.hl_pos = st->plan.hl_pos; n.hl_size = st->plan.hl_size;
    n.hc_coef = st->plan.hc_coef; n.hc_pos = st->plan.hc_pos; n.hc_size = st->plan.hc_size;
    n.vl_coef = st->plan.vl_coef; n.vl_pos = st->plan.vl_pos; n.vl_size = st->plan.vl_size;
    n.vc_coef = st->plan.vc_coef; n.vc_pos = st->plan.vc_pos; n.vc_size = st->plan.vc_size;
    st->plan = n;
    return 0;
}

/* ---------------------------------------------------------------- tile15 host side */

typedef void (*tile15_kernel_t)(const SwsCudaPlan, const Tile15Args);

template <int SRCK, int HT>
static tile15_kernel_t pick_tile15_out(int outk)
{
    return outk == T15_OUT_PLANAR8 ? sws_tile15_kernel<SRCK, HT, T15_OUT_PLANAR8>
         : outk == T15_OUT_PLANARN ? sws_tile15_kernel<SRCK, HT, T15_OUT_PLANARN>
                                   : sws_tile15_kernel<SRCK, HT, T15_OUT_RGB8>;
}

template <int SRCK>
static tile15_kernel_t pick_tile15_ht(int ht, int outk)
{
    return ht == 4 ? pick_tile15_out<SRCK, 4>(outk) : ht == 8 ? pick_tile15_out<SRCK, 8>(outk)
                                                              : pick_tile15_out<SRCK, 16>(outk);
}

static tile15_kernel_t pick_tile15(int srck, int ht, int outk)
{
    return srck == T15_SRC_U8 ? pick_tile15_ht<T15_SRC_U8>(ht, outk)
         : srck == T15_SRC_U16 ? pick_tile15_ht<T15_SRC_U16>(ht, outk)
                               : pick_tile15_ht<T15_SRC_RGB>(ht, outk);
}

/* positions non-decreasing and every tap inside [0, src_len): what initFilter guarantees */
static bool bank_regular(const SwsFirBank *b, int src_len)
{
    for (int i = 0; i < b->len; i++) {
        if (b->pos[i] < 0 || b->pos[i] + b->size > src_len)
            return false;
        if (i && b->pos[i] < b->pos[i - 1])
            return false;
    }
    return true;
}

/* widest source span (in samples) any tile of `tw` output columns reads */
static int bank_span(const SwsFirBank *b, int tw)
{
    int worst = 0;
    for (int x0 = 0; x0 < b->len; x0 += tw) {
        const int x1 = x0 + tw - 1 < b->len - 1 ? x0 + tw - 1 : b->len - 1;
        const int span = b->pos[x1] + b->size - b->pos[x0];
        if (span > worst)
            worst = span;
    }
    return worst;
}

static int tile15_setup(SwsCudaState *st, const SwsFirBank *hl, const SwsFirBank *hc,
                        const SwsFirBank *vl, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->t15_ok = 0;
    if (p->inter_bits != 15 || p->full_chr || p->special || !p->has_chroma || (p->unscaled_lut && (p->dst_w & 1)) ||
        p->src_alpha)
        return 0;
    int outk;
    if (p->dst_kind == SWSC_DST_PLANAR8 || p->dst_kind == SWSC_DST_NV12 || p->dst_kind == SWSC_DST_NV21)
        outk = T15_OUT_PLANAR8;
    else if (p->dst_kind == SWSC_DST_PLANARN)
        outk = T15_OUT_PLANARN;
    else if (p->dst_kind >= SWSC_DST_RGB24 && p->dst_kind <= SWSC_DST_ABGR && p->chr_dst_hsub == 1)
        outk = T15_OUT_RGB8;
    else
        return 0;
    const int srck = p->src_layout == SWSC_SRC_RGB ? T15_SRC_RGB : p->src_bits == 8 ? T15_SRC_U8 : T15_SRC_U16;
    if (srck == T15_SRC_U16 && p->src_layout != SWSC_SRC_PLANAR)
        return 0;
    if (srck == T15_SRC_RGB && p->src_bpp == 6)
        return 0;                     /* 48-bit pixels: the general kernel's readers */
    const int fs = hl->size > hc->size ? hl->size : hc->size;
    if (fs > 16)
        return 0;
    if (!bank_regular(hl, p->src_w) || !bank_regular(hc, p->chr_src_w) ||
        !bank_regular(vl, p->src_h) || !bank_regular(vc, p->chr_src_h))
        return 0;
    const int ht = fs <= 4 ? 4 : fs <= 8 ? 8 : 16;
    const int cw = T15_TW >> p->chr_dst_hsub;
    const int seg_l = (bank_span(hl, T15_TW) + 7) & ~7;
    const int seg_c = (bank_span(hc, cw) + 7) & ~7;
    const size_t ss = srck == T15_SRC_U8 ? 1 : 2;
    const int vs = p->chr_dst_vsub;
    const size_t budget = 72 * 1024;               /* three CTAs per SM */
    int found = 0;
    for (int th = 32; th >= (1 << vs) && !found; th >>= 1) {
        const int rl = max_rows_needed(st->h_vl_pos, p->vl_size, p->dst_h, th);
        const int cth = th >> vs ? th >> vs : 1;
        const int rc = max_rows_needed(st->h_vc_pos, p->vc_size, p->chr_dst_h, cth);
        const size_t lines = ((size_t)rl * T15_TW + 2 * (size_t)rc * cw) * 2;
        /* stage whole tiles when they fit, else as many rows per pass as the budget allows (>= 4) */
        for (int div = 1; div <= 64 && !found; div *= 2) {
            int srl = (rl + div - 1) / div, src = (rc + div - 1) / div;
            if (div > 1 && (srl < 4 || src < 4)) {
                srl = srl < 4 ? (rl < 4 ? rl : 4) : srl;
                src = src < 4 ? (rc < 4 ? rc : 4) : src;
            }
            size_t stage = ((size_t)srl * seg_l + 2 * (size_t)src * seg_c) * ss;
            if (stage < T15_OUT_BYTES)
                stage = T15_OUT_BYTES;
            stage = (stage + 15) & ~(size_t)15;
            if (lines + stage <= budget || (th == (1 << vs) && div == 64 && lines + stage <= 200 * 1024)) {
                st->t15_tile_h = th; st->t15_nl_cap = rl; st->t15_nc_cap = rc;
                st->t15_srl = srl; st->t15_src = src;
                st->t15_smem = lines + stage;
                found = 1;
            }
        }
    }
    if (!found)
        return 0;
    st->t15_srck = srck; st->t15_ht = ht; st->t15_outk = outk;
    st->t15_seg_l = seg_l; st->t15_seg_c = seg_c;
    CUDA_OK(set_max_smem((const void *)pick_tile15(srck, ht, outk), (size_t)((int)st->t15_smem)));
    st->t15_ok = 1;
    return 0;
}

static int tile15_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                         const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                         const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    if (!st->t15_ok || (st->disabled & 16))
        return 0;
    const SwsCudaPlan *p = &st->plan;
    Tile15Args a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 3; i++) {
        a.src[i] = src[i]; a.dst[i] = dst[i];
        a.src_stride[i] = src_stride[i]; a.dst_stride[i] = dst_stride[i];
        a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
        a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
    }
    a.y0 = y0; a.y1 = y1; a.tile_h = st->t15_tile_h;
    a.nl_cap = st->t15_nl_cap; a.nc_cap = st->t15_nc_cap;
    a.seg_l = st->t15_seg_l; a.seg_c = st->t15_seg_c;
    a.srl = st->t15_srl; a.src_rows = st->t15_src;
    dim3 grid((p->dst_w + T15_TW - 1) / T15_TW, (y1 - y0 + a.tile_h - 1) / a.tile_h, nb_frames);
    pick_tile15(st->t15_srck, st->t15_ht, st->t15_outk)<<<grid, 256, st->t15_smem, stream>>>(*p, a);
    st->kernel_name = "tile15";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}

/* ---------------------------------------------------------------- rgb420 host side */

/* the rgb420 kernel keeps 16-bit matrix coefficients and drops the readers' uint16 wrap and the 15-bit clip:
 * admit only matrices whose 14-bit samples stay in [0, 16384) for every input (luma: one pixel, bias
 * (32 << 14) + (1 << 8), >> 9; chroma: a pixel pair, bias (256 << 15) + (1 << 9), >> 10).  Checked at init and
 * at every launch (sws_setColorspaceDetails() may replace the matrix). */
static bool rgb420_matrix_ok(const SwsCudaPlan *p)
{
    for (int i = 0; i < 9; i++)
        if (p->rgb2yuv[i] < -32768 || p->rgb2yuv[i] > 32767)
            return false;
    for (int r = 0; r < 3; r++) {
        long long lo = 0, hi = 0;
        for (int k = 0; k < 3; k++) {
            const long long c = p->rgb2yuv[3 * r + k];
            (c < 0 ? lo : hi) += c * (r ? 510 : 255);
        }
        const long long bias = r ? (256LL << 15) + (1 << 9) : (32LL << 14) + (1 << 8);
        const int sh = r ? 10 : 9;
        if (bias + lo < 0 || ((bias + hi) >> sh) >= 16384)
            return false;
    }
    return true;
}

static int rgb420_setup(SwsCudaState *st, const SwsFirBank *vc)
{
    const SwsCudaPlan *p = &st->plan;
    st->r420_ok = 0;
    if (p->src_layout != SWSC_SRC_RGB || !p->src_rgb_half || p->special || p->range_mode || p->dither_bayer ||
        p->inter_bits != 15 || !p->lum_identity || !p->chr_h_identity || !p->has_chroma)
        return 0;
    if (p->dst_kind != SWSC_DST_PLANAR8 && p->dst_kind != SWSC_DST_NV12 && p->dst_kind != SWSC_DST_NV21)
        return 0;
    if (p->chr_dst_hsub != 1 || p->chr_dst_vsub != 1 || p->src_w != p->dst_w || p->src_h != p->dst_h ||
        (p->src_w & 15) || p->chr_src_w != p->chr_dst_w || p->chr_src_h != p->src_h || p->h_shift != 13)
        return 0;
    if (vc->size > 16 || !bank_regular(vc, p->chr_src_h))
        return 0;
    if (!rgb420_matrix_ok(p))
        return 0;
    /* chroma rows per tile: as many as keep the source-row window (and the luma rows) inside the line buffer */
    int cr;
    for (cr = 32; cr >= 1; cr--) {
        int worst = 0;
        for (int c0 = 0; c0 < vc->len; c0 += cr) {
            const int c1 = c0 + cr < vc->len ? c0 + cr : vc->len;
            const int n = vc->pos[c1 - 1] + vc->size - vc->pos[c0];
            if (n > worst) worst = n;
        }
        if (worst <= R420_MAXROWS)
            break;
    }
    if (cr < 4)
        return 0;
    st->r420_cr = cr;
    st->r420_ok = 1;
    return 0;
}

static int rgb420_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                         const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                         const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    /* range conversion can be switched on after init by sws_setColorspaceDetails(): another kernel's job */
    if (!st->r420_ok || (st->disabled & 64) || p->range_mode || (y0 & 1) || (y1 != p->dst_h && (y1 & 1)) ||
        !rgb420_matrix_ok(p))
        return 0;
    const int ndst = p->dst_kind == SWSC_DST_PLANAR8 ? 3 : 2;
    if (!src[0] || !aligned16(src[0]) || (src_stride[0] & 15) || src_stride[0] <= 0 ||
        (nb_frames > 1 && (src_fstride[0] & 15)))
        return 0;
    for (int i = 0; i < ndst; i++)
        if (!dst[i] || !aligned16(dst[i]) || (dst_stride[i] & 15) || dst_stride[i] <= 0 ||
            (nb_frames > 1 && (dst_fstride[i] & 15)))
            return 0;
    Rgb420Args a;
    memset(&a, 0, sizeof(a));
    a.src = src[0]; a.src_stride = src_stride[0]; a.src_fstride = src_fstride ? src_fstride[0] : 0;
    for (int i = 0; i < 3; i++) {
        a.dst[i] = dst[i]; a.dst_stride[i] = dst_stride[i];
        a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
    }
    a.w = p->dst_w;
    a.cy_begin = y0 >> 1;
    a.cy_end = y1 == p->dst_h ? p->chr_dst_h : y1 >> 1;
    a.y_end = y1;
    a.cr = st->r420_cr;
    a.vc_size = p->vc_size; a.vc_coef = p->vc_coef; a.vc_pos = p->vc_pos;
    a.dst_kind = p->dst_kind;
    /* matrix rows as 16-bit pairs in the byte order of a pixel word (unused bytes get a zero coefficient) */
    int ky[4] = { 0, 0, 0, 0 }, ku[4] = { 0, 0, 0, 0 }, kv[4] = { 0, 0, 0, 0 };
    ky[p->src_ro] = p->rgb2yuv[0]; ky[p->src_go] = p->rgb2yuv[1]; ky[p->src_bo] = p->rgb2yuv[2];
    ku[p->src_ro] = p->rgb2yuv[3]; ku[p->src_go] = p->rgb2yuv[4]; ku[p->src_bo] = p->rgb2yuv[5];
    kv[p->src_ro] = p->rgb2yuv[6]; kv[p->src_go] = p->rgb2yuv[7]; kv[p->src_bo] = p->rgb2yuv[8];
    auto pair = [](int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); };
    a.ylo = pair(ky[0], ky[1]); a.yhi = pair(ky[2], ky[3]);
    a.ulo = pair(ku[0], ku[1]); a.uhi = pair(ku[2], ku[3]);
    a.vlo = pair(kv[0], kv[1]); a.vhi = pair(kv[2], kv[3]);
    if (a.cy_end <= a.cy_begin)
        return 0;
    dim3 grid((p->dst_w + R420_TW - 1) / R420_TW, (a.cy_end - a.cy_begin + a.cr - 1) / a.cr, nb_frames);
    if (p->src_bpp == 3)
        sws_rgb420_kernel<3><<<grid, 256, 0, stream>>>(a);
    else
        sws_rgb420_kernel<4><<<grid, 256, 0, stream>>>(a);
    st->kernel_name = "rgb420";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}

/* sws_rgb444_kernel drops the readers' uint16 wrap and the 15-bit clip: admit only matrices whose 14-bit samples stay in
 * [0, 16384) for every pixel (luma bias (32 << 14) + (1 << 8), chroma bias (256 << 14) + (1 << 8), both >> 9) */
static bool rgb444_matrix_ok(const SwsCudaPlan *p)
{
    for (int i = 0; i < 9; i++)
        if (p->rgb2yuv[i] < -32768 || p->rgb2yuv[i] > 32767)
            return false;
    for (int r = 0; r < 3; r++) {
        long long lo = 0, hi = 0;
        for (int k = 0; k < 3; k++) {
            const long long c = p->rgb2yuv[3 * r + k];
            (c < 0 ? lo : hi) += c * 255;
        }
        const long long bias = r ? (256LL << 14) + (1 << 8) : (32LL << 14) + (1 << 8);
        if (bias + lo < 0 || ((bias + hi) >> 9) >= 16384)
            return false;
    }
    return true;
}

/* whole-frame special converters; returns 1 if launched */
static int special_launch(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                          const int64_t src_fstride[4], uint8_t *const dst[4], const int dst_stride[4],
                          const int64_t dst_fstride[4], int nb_frames, int y0, int y1, cudaStream_t stream)
{
    const SwsCudaPlan *p = &st->plan;
    /* identity 8-bit yuv -> yuv: the reference's copy wrappers (chosen at init, range changes after init
     * do not unseat them) and every conversion whose four filters are the identity without range conversion */
    if (p->special == SWSC_SPECIAL_COPY8 ||
        (!p->special && p->src_bits == 8 && p->src_layout <= SWSC_SRC_NV21 && p->has_chroma && !p->range_mode &&
         !p->dither_bayer && p->lum_identity && p->chr_h_identity && p->chr_v_identity &&
         (p->dst_kind == SWSC_DST_PLANAR8 || p->dst_kind == SWSC_DST_NV12 || p->dst_kind == SWSC_DST_NV21) &&
         p->src_w == p->dst_w && p->src_h == p->dst_h && p->chr_src_w == p->chr_dst_w && p->chr_src_h == p->chr_dst_h &&
         !(st->disabled & 32))) {
        const int nsrc = p->src_layout == SWSC_SRC_PLANAR ? 3 : 2, ndst = p->dst_kind == SWSC_DST_PLANAR8 ? 3 : 2;
        bool vec = true, ok = true;
        for (int i = 0; i < nsrc; i++) {
            ok = ok && src[i];
            vec = vec && aligned16(src[i]) && !(src_stride[i] & 15) && !(nb_frames > 1 && (src_fstride[i] & 15));
        }
        for (int i = 0; i < ndst; i++) {
            ok = ok && dst[i];
            vec = vec && aligned16(dst[i]) && !(dst_stride[i] & 15) && !(nb_frames > 1 && (dst_fstride[i] & 15));
        }
        if (!ok)
            return AVERROR(EINVAL);
        {
            Copy8Args a;
            memset(&a, 0, sizeof(a));
            for (int i = 0; i < 3; i++) {
                a.src[i] = src[i]; a.dst[i] = dst[i];
                a.src_stride[i] = src_stride[i]; a.dst_stride[i] = dst_stride[i];
                a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
                a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
            }
            a.w = p->dst_w; a.cw = p->chr_dst_w; a.y0 = y0; a.rows = y1 - y0;
            a.cy0 = y0 >> p->chr_dst_vsub;
            a.crows = (y1 == p->dst_h ? p->chr_dst_h : y1 >> p->chr_dst_vsub) - a.cy0;
            a.lchunks = (a.w + 15) / 16; a.cchunks = (a.cw + 15) / 16;
            a.src_layout = p->src_layout; a.dst_kind = p->dst_kind;
            a.vec = vec;
            const long long lw = (long long)a.lchunks * a.rows, cwk = (long long)a.cchunks * a.crows;
            const long long mx = lw > cwk ? lw : cwk;
            dim3 grid((unsigned)((mx + 255) / 256), 2, nb_frames);
            sws_copy8_kernel<<<grid, 256, 0, stream>>>(a);
            st->kernel_name = "copy8";
            CUDA_OK(cudaGetLastError());
            st->launches++;
            return 1;
        }
    }
    /* packed 8-bit RGB -> 8-bit planar 4:4:4 of the same size: full-resolution readers, identity filters */
    if (!p->special && p->src_layout == SWSC_SRC_RGB && !p->src_rgb_half && p->dst_kind == SWSC_DST_PLANAR8 &&
        p->chr_dst_hsub == 0 && p->chr_dst_vsub == 0 && p->has_chroma && !p->range_mode && !p->dither_bayer &&
        p->inter_bits == 15 && p->h_shift == 13 && p->lum_identity && p->chr_h_identity && p->chr_v_identity &&
        p->src_w == p->dst_w && p->src_h == p->dst_h && p->chr_src_w == p->dst_w && !(st->disabled & 512) &&
        rgb444_matrix_ok(p)) {
        Rgb444Args a;
        memset(&a, 0, sizeof(a));
        if (!src[0])
            return AVERROR(EINVAL);
        bool vec = aligned16(src[0]) && !(src_stride[0] & 15) && !(nb_frames > 1 && (src_fstride[0] & 15));
        a.src = src[0]; a.src_stride = src_stride[0]; a.src_fstride = src_fstride ? src_fstride[0] : 0;
        for (int i = 0; i < 3; i++) {
            if (!dst[i])
                return AVERROR(EINVAL);
            a.dst[i] = dst[i]; a.dst_stride[i] = dst_stride[i]; a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
            vec = vec && aligned16(dst[i]) && !(dst_stride[i] & 15) && !(a.dst_fstride[i] & 15);
        }
        a.w = p->dst_w; a.y0 = y0; a.rows = y1 - y0; a.chunks = (p->dst_w + 15) / 16; a.vec = vec;
        int ky[4] = { 0, 0, 0, 0 }, ku[4] = { 0, 0, 0, 0 }, kv[4] = { 0, 0, 0, 0 };
        ky[p->src_ro] = p->rgb2yuv[0]; ky[p->src_go] = p->rgb2yuv[1]; ky[p->src_bo] = p->rgb2yuv[2];
        ku[p->src_ro] = p->rgb2yuv[3]; ku[p->src_go] = p->rgb2yuv[4]; ku[p->src_bo] = p->rgb2yuv[5];
        kv[p->src_ro] = p->rgb2yuv[6]; kv[p->src_go] = p->rgb2yuv[7]; kv[p->src_bo] = p->rgb2yuv[8];
        auto pair = [](int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); };
        a.ylo = pair(ky[0], ky[1]); a.yhi = pair(ky[2], ky[3]);
        a.ulo = pair(ku[0], ku[1]); a.uhi = pair(ku[2], ku[3]);
        a.vlo = pair(kv[0], kv[1]); a.vhi = pair(kv[2], kv[3]);
        const long long work = (long long)a.chunks * a.rows;
        dim3 grid((unsigned)((work + 255) / 256), 1, nb_frames);
        if (p->src_bpp == 3)
            sws_rgb444_kernel<3><<<grid, 256, 0, stream>>>(a);
        else
            sws_rgb444_kernel<4><<<grid, 256, 0, stream>>>(a);
        st->kernel_name = "rgb444";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    /* 8-bit 4:4:4 -> packed RGB of the same size: full-chroma arithmetic path with identity filters */
    if (!p->special && p->full_chr && p->src_bits == 8 && p->src_layout == SWSC_SRC_PLANAR && p->has_chroma &&
        p->chr_src_hsub == 0 && p->chr_src_vsub == 0 && p->dst_kind >= SWSC_DST_RGB24 && p->dst_kind <= SWSC_DST_ABGR &&
        p->lum_identity && p->chr_h_identity && p->chr_v_identity && p->src_w == p->dst_w && p->src_h == p->dst_h &&
        p->chr_dst_w == p->dst_w && !(st->disabled & 256)) {
        Full444Args a;
        memset(&a, 0, sizeof(a));
        bool vec = aligned16(dst[0]) && !(dst_stride[0] & 15) && !(nb_frames > 1 && (dst_fstride[0] & 15));
        for (int i = 0; i < 3; i++) {
            if (!src[i])
                return AVERROR(EINVAL);
            a.src[i] = src[i]; a.src_stride[i] = src_stride[i]; a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
            vec = vec && aligned16(src[i]) && !(src_stride[i] & 15) && !(a.src_fstride[i] & 15);
        }
        a.dst = dst[0]; a.dst_stride = dst_stride[0]; a.dst_fstride = dst_fstride ? dst_fstride[0] : 0;
        a.w = p->dst_w; a.y0 = y0; a.rows = y1 - y0; a.chunks = (p->dst_w + 15) / 16;
        a.dst_kind = p->dst_kind; a.vec = vec;
        /* R = ((y << 9) - y_off) * y_coeff + 2^21 + ((v - 128) << 9) * v2r, all mod 2^32 */
        const unsigned yc = (unsigned)p->rgb.y_coeff, base = (1u << 21) - (unsigned)p->rgb.y_offset * yc;
        a.ky = yc << 9;
        a.kvr = (unsigned)p->rgb.v2r << 9; a.kvg = (unsigned)p->rgb.v2g << 9;
        a.kug = (unsigned)p->rgb.u2g << 9; a.kub = (unsigned)p->rgb.u2b << 9;
        a.cr = base - 128u * a.kvr;
        a.cg = base - 128u * a.kvg - 128u * a.kug;
        a.cb = base - 128u * a.kub;
        const long long work = (long long)a.chunks * a.rows;
        dim3 grid((unsigned)((work + 255) / 256), 1, nb_frames);
        switch (p->dst_kind) {
        case SWSC_DST_RGB24: sws_full444_kernel<SWSC_DST_RGB24><<<grid, 256, 0, stream>>>(a); break;
        case SWSC_DST_BGR24: sws_full444_kernel<SWSC_DST_BGR24><<<grid, 256, 0, stream>>>(a); break;
        case SWSC_DST_RGBA:  sws_full444_kernel<SWSC_DST_RGBA><<<grid, 256, 0, stream>>>(a); break;
        case SWSC_DST_BGRA:  sws_full444_kernel<SWSC_DST_BGRA><<<grid, 256, 0, stream>>>(a); break;
        case SWSC_DST_ARGB:  sws_full444_kernel<SWSC_DST_ARGB><<<grid, 256, 0, stream>>>(a); break;
        default:             sws_full444_kernel<SWSC_DST_ABGR><<<grid, 256, 0, stream>>>(a); break;
        }
        st->kernel_name = "full444";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    if (p->special == SWSC_SPECIAL_P01X) {
        P01xArgs a;
        memset(&a, 0, sizeof(a));
        for (int i = 0; i < 3; i++) {
            if (!src[i] || (i < 2 && !dst[i]))
                return AVERROR(EINVAL);
            a.src[i] = src[i]; a.src_stride[i] = src_stride[i]; a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
            if (i < 2) {
                a.dst[i] = dst[i]; a.dst_stride[i] = dst_stride[i]; a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
            }
        }
        a.w = p->dst_w; a.cw = p->dst_w / 2;                 /* the reference converts src_w / 2 pairs per row */
        a.y0 = y0; a.rows = y1 - y0;
        a.cy0 = (y0 + 1) >> 1;                               /* chroma rows at even luma rows (:311,:358) */
        a.crows = ((y1 + 1) >> 1) - a.cy0;
        a.shift = p->src_bits == 8 ? 8 : 16 - p->src_bits;
        a.vec = 1;
        for (int i = 0; i < 3; i++)
            a.vec = a.vec && aligned16(src[i]) && !(src_stride[i] & 15) && !(a.src_fstride[i] & 15);
        for (int i = 0; i < 2; i++)
            a.vec = a.vec && aligned16(dst[i]) && !(dst_stride[i] & 15) && !(a.dst_fstride[i] & 15);
        const long long work = (long long)((a.w + 7) / 8) * a.rows + (long long)((a.cw + 7) / 8) * a.crows;
        dim3 grid((unsigned)((work + 255) / 256), 1, nb_frames);
        if (p->src_bits == 8)
            sws_p01x_kernel<uint8_t><<<grid, 256, 0, stream>>>(a);
        else
            sws_p01x_kernel<uint16_t><<<grid, 256, 0, stream>>>(a);
        st->kernel_name = "p01x";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    if (p->special == SWSC_SPECIAL_DEPTHCOPY) {
        DepthCopyArgs a;
        memset(&a, 0, sizeof(a));
        const bool semi = p->src_layout != SWSC_SRC_PLANAR;       /* nv12 -> p010: the UV plane is one row of 2 cw samples */
        const int np = semi ? 2 : 3;
        bool vec = true;
        long long mx = 0;
        for (int i = 0; i < np; i++) {
            if (!src[i] || !dst[i])
                return AVERROR(EINVAL);
            a.src[i] = src[i]; a.dst[i] = dst[i];
            a.src_stride[i] = src_stride[i]; a.dst_stride[i] = dst_stride[i];
            a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
            a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
            vec = vec && aligned16(src[i]) && aligned16(dst[i]) && !(src_stride[i] & 15) && !(dst_stride[i] & 15) &&
                  !(a.src_fstride[i] & 15) && !(a.dst_fstride[i] & 15);
            a.w[i] = i ? (semi ? 2 * p->chr_dst_w : p->chr_dst_w) : p->dst_w;
            a.y0[i] = i ? (y0 + (1 << p->chr_dst_vsub) - 1) >> p->chr_dst_vsub : y0;       /* AV_CEIL_RSHIFT, :2229-2230 */
            const int yend = i ? (y1 == p->dst_h ? p->chr_dst_h : (y1 + (1 << p->chr_dst_vsub) - 1) >> p->chr_dst_vsub) : y1;
            a.rows[i] = yend - a.y0[i];
            const int ch = p->src_bits > 8 && p->dst_bits == 8 ? 16 : 8;      /* samples per thread, as in the kernel */
            a.chunks[i] = (a.w[i] + ch - 1) / ch;
            const long long work = (long long)a.chunks[i] * a.rows[i];
            if (work > mx) mx = work;
        }
        a.src_depth = p->src_bits; a.dst_depth = p->dst_bits;
        a.src_shift = p->src_shift; a.dst_shift = p->dst_shift;
        a.luma_shiftonly = !p->src_full_range;
        a.dither_none = p->dither_none;
        a.vec = vec;
        a.nplanes = np;
        long long total = 0;
        for (int i = 0; i < np; i++)
            total += (long long)a.chunks[i] * a.rows[i];
        (void)mx;
        dim3 grid((unsigned)((total + 255) / 256), 1, nb_frames);
        if (p->src_bits == 8)
            sws_depthcopy_kernel<uint8_t, uint16_t><<<grid, 256, 0, stream>>>(a);
        else if (p->dst_bits == 8)
            sws_depthcopy_kernel<uint16_t, uint8_t><<<grid, 256, 0, stream>>>(a);
        else
            sws_depthcopy_kernel<uint16_t, uint16_t><<<grid, 256, 0, stream>>>(a);
        st->kernel_name = "depthcopy";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    if (p->special == SWSC_SPECIAL_RGB16PACK) {
        Rgb16PackArgs a;
        memset(&a, 0, sizeof(a));
        if (!src[0] || !dst[0])
            return AVERROR(EINVAL);
        a.src = src[0]; a.dst = dst[0];
        a.src_fstride = src_fstride ? src_fstride[0] : 0;
        a.dst_fstride = dst_fstride ? dst_fstride[0] : 0;
        a.src_stride = src_stride[0]; a.dst_stride = dst_stride[0];
        a.w = p->src_w; a.y0 = y0; a.rows = y1 - y0;
        a.bpp = p->src_bpp; a.ro = p->src_ro; a.go = p->src_go; a.bo = p->src_bo;
        a.gbits = p->dst_kind == SWSC_DST_RGB565 || p->dst_kind == SWSC_DST_BGR565 ? 6 : 5;
        a.rgb = p->dst_kind == SWSC_DST_RGB565 || p->dst_kind == SWSC_DST_RGB555;
        const long long work = (long long)((a.w + 1) >> 1) * a.rows;
        dim3 grid((unsigned)((work + 255) / 256), 1, nb_frames);
        sws_rgb16pack_kernel<<<grid, 256, 0, stream>>>(a);
        st->kernel_name = "rgb16pack";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    if (p->special == SWSC_SPECIAL_RGB48) {
        if (!src[0] || !dst[0])
            return AVERROR(EINVAL);
        Rgb48Args a;
        a.src = src[0]; a.dst = dst[0];
        a.src_fstride = src_fstride ? src_fstride[0] : 0;
        a.dst_fstride = dst_fstride ? dst_fstride[0] : 0;
        a.src_stride = src_stride[0]; a.dst_stride = dst_stride[0];
        a.w = p->src_w; a.y0 = y0; a.rows = y1 - y0;
        a.swap = (p->src_ro == 0) != (p->dst_kind == SWSC_DST_RGB48);
        const long long work = (long long)a.w * a.rows;
        for (int f0 = 0; f0 < nb_frames; f0 += 65535) {
            Rgb48Args b = a;
            b.src += f0 * a.src_fstride; b.dst += f0 * a.dst_fstride;
            dim3 grid((unsigned)((work + 255) / 256), 1, nb_frames - f0 < 65535 ? nb_frames - f0 : 65535);
            sws_rgb48_kernel<<<grid, 256, 0, stream>>>(b);
        }
        st->kernel_name = "rgb48";
        CUDA_OK(cudaGetLastError());
        st->launches++;
        return 1;
    }
    if (p->special == SWSC_SPECIAL_SHUFFLE) {
        ShuffleArgs a;
        a.src = src[0]; a.dst = dst[0];
        a.src_fstride = src_fstride ? src_fstride[0] : 0;
        a.dst_fstride = dst_fstride ? dst_fstride[0] : 0;
        a.src_stride = src_stride[0]; a.dst_stride = dst_stride[0];
        a.w = p->src_w; a.y0 = y0;
        for (int k = 0; k < 4; k++)
            a.map[k] = p->shuf_map[k];
        const bool vec = aligned16(src[0]) && aligned16(dst[0]) && !(src_stride[0] & 15) && !(dst_stride[0] & 15) &&
                         src_stride[0] > 0 && dst_stride[0] > 0 && !(a.src_fstride & 15) && !(a.dst_fstride & 15);
        if (vec) {
            ShuffleVecArgs v;
            memset(&v, 0, sizeof(v));
            v.src = a.src; v.dst = a.dst; v.src_fstride = a.src_fstride; v.dst_fstride = a.dst_fstride;
            v.src_stride = a.src_stride; v.dst_stride = a.dst_stride;
            v.w = a.w; v.y0 = y0; v.rows = y1 - y0; v.chunks = (a.w + 15) / 16;
            const int sb = p->src_bpp, db = p->dst_bpp;
            for (int j = 0; j < db; j++) {
                const int lo = ((4 * j) / db * sb) / 4;
                uint32_t s1 = 0, s2 = 0, keep = 0, fill = 0;
                for (int b = 0; b < 4; b++) {
                    const int ob = 4 * j + b, px = ob / db, m = p->shuf_map[ob % db];
                    v.map[ob % db] = m;
                    if (m == 4) {
                        fill |= 0xFFu << (8 * b);
                        s2 |= (uint32_t)b << (4 * b);
                        continue;
                    }
                    const int ib = px * sb + m, wi = ib / 4 - lo, bi = ib % 4;
                    keep |= 0xFFu << (8 * b);
                    if (wi <= 1) {
                        s1 |= (uint32_t)(wi * 4 + bi) << (4 * b);
                        s2 |= (uint32_t)b << (4 * b);
                    } else {
                        s2 |= (uint32_t)(4 + bi) << (4 * b);
                    }
                }
                v.sel1[j] = s1; v.sel2[j] = s2; v.keep[j] = keep; v.fill[j] = fill;
            }
            const long long work = (long long)v.chunks * v.rows;
            dim3 vgrid((unsigned)((work + 255) / 256), 1, nb_frames);
            if (sb == 3 && db == 3)
                sws_rgb_shuffle_vec_kernel<3, 3><<<vgrid, 256, 0, stream>>>(v);
            else if (sb == 3)
                sws_rgb_shuffle_vec_kernel<3, 4><<<vgrid, 256, 0, stream>>>(v);
            else if (db == 3)
                sws_rgb_shuffle_vec_kernel<4, 3><<<vgrid, 256, 0, stream>>>(v);
            else
                sws_rgb_shuffle_vec_kernel<4, 4><<<vgrid, 256, 0, stream>>>(v);
            st->kernel_name = "rgb_shuffle";
            CUDA_OK(cudaGetLastError());
            st->launches++;
            return 1;
        }
        dim3 grid((p->src_w + SHUF_PX - 1) / SHUF_PX, y1 - y0, nb_frames);
        if (p->src_bpp == 3 && p->dst_bpp == 3)
            sws_rgb_shuffle_kernel<3, 3><<<grid, 256, 0, stream>>>(a);
        else if (p->src_bpp == 3)
            sws_rgb_shuffle_kernel<3, 4><<<grid, 256, 0, stream>>>(a);
        else if (p->dst_bpp == 3)
            sws_rgb_shuffle_kernel<4, 3><<<grid, 256, 0, stream>>>(a);
        else
            sws_rgb_shuffle_kernel<4, 4><<<grid, 256, 0, stream>>>(a);
        st->kernel_name = "rgb_shuffle";
    } else if (p->special == SWSC_SPECIAL_BGR24_YV12) {
        Bgr24Yv12Args a;
        a.src = src[0];
        a.src_fstride = src_fstride ? src_fstride[0] : 0;
        a.src_stride = src_stride[0];
        for (int i = 0; i < 3; i++) {
            if (!dst[i])
                return AVERROR(EINVAL);
            a.dst[i] = dst[i];
            a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
            a.dst_stride[i] = dst_stride[i];
        }
        a.cw = p->src_w >> 1; a.y0 = y0; a.h = y1 - y0;
        a.ry = p->rgb2yuv[0]; a.gy = p->rgb2yuv[1]; a.by = p->rgb2yuv[2];
        a.ru = p->rgb2yuv[3]; a.gu = p->rgb2yuv[4]; a.bu = p->rgb2yuv[5];
        a.rv = p->rgb2yuv[6]; a.gv = p->rgb2yuv[7]; a.bv = p->rgb2yuv[8];
        bool vec = !(p->src_w & 15) && aligned16(src[0]) && !(src_stride[0] & 15) && !(a.src_fstride & 15) &&
                   aligned16(dst[0]) && !(dst_stride[0] & 15) && !(a.dst_fstride[0] & 15);
        for (int i = 1; i < 3; i++)
            vec = vec && !((uintptr_t)dst[i] & 7) && !(dst_stride[i] & 7) && !(a.dst_fstride[i] & 7);
        if (vec) {
            const int chunks = p->src_w / 16, pairs = (a.h + 1) / 2;
            const long long work = (long long)chunks * pairs;
            dim3 vgrid((unsigned)((work + 255) / 256), 1, nb_frames);
            sws_bgr24_to_yv12_vec_kernel<<<vgrid, 256, 0, stream>>>(a, chunks, pairs);
        } else {
            dim3 grid((a.cw + 255) / 256, (a.h + 1) / 2, nb_frames);
            sws_bgr24_to_yv12_kernel<<<grid, 256, 0, stream>>>(a);
        }
        st->kernel_name = "bgr24_to_yv12";
    } else {
        return 0;
    }
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 1;
}

extern "C" int ff_b200_cuda_launch(SwsCudaState *st,
                                   const uint8_t *const src[4], const int src_stride[4], const int64_t src_fstride[4],
                                   uint8_t *const dst[4], const int dst_stride[4], const int64_t dst_fstride[4],
                                   int nb_frames, int y0, int y1)
{
    if (y1 <= y0)
        return 0;
    DeviceGuard guard(st->device);
    NvtxRange nvtx("sws_b200: convert");
    {
        int r = special_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = fast420_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = fast16_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = fasthi8_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = rgb420_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = scale8_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
        r = tile15_launch(st, src, src_stride, src_fstride, dst, dst_stride, dst_fstride, nb_frames, y0, y1, st->stream);
        if (r != 0)
            return r < 0 ? r : 0;
    }
    FrameArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; i++) {
        a.src[i] = src[i]; a.dst[i] = dst[i];
        a.src_stride[i] = src_stride[i]; a.dst_stride[i] = dst_stride[i];
        a.src_fstride[i] = src_fstride ? src_fstride[i] : 0;
        a.dst_fstride[i] = dst_fstride ? dst_fstride[i] : 0;
    }
    a.y0 = y0; a.y1 = y1;
    a.tile_w = st->tile_w; a.tile_h = st->tile_h;
    a.rows_l_cap = st->rows_l_cap; a.rows_c_cap = st->rows_c_cap;
    dim3 grid((st->plan.dst_w + st->tile_w - 1) / st->tile_w,
              (y1 - y0 + st->tile_h - 1) / st->tile_h, nb_frames);
    pick_generic(&st->plan)<<<grid, 256, st->smem_bytes, st->stream>>>(st->plan, a);
    st->kernel_name = "generic_tile";
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 0;
}

#include "sws_xfer.cuh"

extern "C" int ff_b200_cuda_sync(SwsCudaState *st)
{
    DeviceGuard guard(st->device);
    CUDA_OK(cudaStreamSynchronize(st->stream));
    return 0;
}

extern "C" void *ff_b200_cuda_stream(SwsCudaState *st) { return (void *)st->stream; }
extern "C" long ff_b200_cuda_launch_count(SwsCudaState *st) { return st->launches; }
extern "C" const char *ff_b200_cuda_kernel_name(SwsCudaState *st) { return st->kernel_name; }
