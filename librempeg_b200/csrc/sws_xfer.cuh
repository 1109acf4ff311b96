/*
 * sws_xfer.cuh -- how host frames reach the kernels and come back: the host-pointer half of sws_scale()
 * (reference swscale.c:1626; the reference works on caller memory directly, a GPU path has to move it).
 *
 * Included by sws_cuda.cu after every kernel launcher.  Three ways through, chosen per call:
 *   serial      H2D, kernel, D2H back to back on the context stream: slices of the legacy API, tiny frames,
 *               negative (bottom-up) strides -- rows are mirrored on the device by sws_flip_rows_kernel
 *   banded      a whole frame as row bands over three streams (H2D of band k+1 | kernel on band k | D2H of
 *               band k-1): PCIe runs in both directions at once, for EVERY kernel of the dispatch chain
 *   bounce      the banded pipeline for pageable caller memory (what av_frame_get_buffer() hands out): the
 *               bands pass through a page-locked ring owned by the context, filled and drained by a small
 *               pool of copy threads so that one core's memcpy rate is not the ceiling
 * Page-locked caller frames + the fast420 kernel additionally let the kernel store straight into the host
 * destination through its TMA store map (no D2H copy at all).
 */
#ifndef SWS_B200_XFER_CUH
#define SWS_B200_XFER_CUH

#include <atomic>
#include <condition_variable>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define SWS_B200_NT_COPY 1
#endif
#include <mutex>
#include <thread>
#include <vector>

/* ------------------------------------------------------------------ copy threads */

namespace {

struct CopyJob {
    uint8_t *dst;
    const uint8_t *src;
    ptrdiff_t dpitch, spitch;
    size_t rowbytes;
    int rows;
};

/* A process-wide pool: W workers sleep on a condition variable; a job is a rows x rowbytes copy that the
 * caller and the workers split by row ranges.  Created on first use (never in a process that only handles
 * page-locked or device frames). */
class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool *pool = new CopyPool();      /* leaked on purpose: workers may outlive static dtors */
        return *pool;
    }

    void copy(const CopyJob &job)
    {
        const size_t total = job.rowbytes * (size_t)job.rows;
        if (workers_.empty() || total < (512u << 10) || job.rows < 2) {
            run_rows(job, 0, job.rows);
            return;
        }
        std::unique_lock<std::mutex> user(user_mutex_);          /* one job at a time */
        const int parts = (int)workers_.size() + 1;
        {
            std::lock_guard<std::mutex> lk(mutex_);
            job_ = job;
            parts_ = parts;
            next_part_ = 1;                                      /* part 0 is the caller's */
            pending_ = parts - 1;
            generation_++;
        }
        wake_.notify_all();
        run_part(job, 0, parts);
        std::unique_lock<std::mutex> lk(mutex_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

    int threads() const { return (int)workers_.size() + 1; }

private:
    CopyPool()
    {
        int n = 0;
        const char *e = getenv("SWS_B200_COPY_THREADS");
        if (e) {
            n = atoi(e);
        } else {
            /* the copies are memory-bound: a handful of cores saturate the host; one process per GPU is the
             * common deployment, so share the cores between the devices of the box */
            const unsigned hc = std::thread::hardware_concurrency();
            int devs = sws_cuda_device_count();
            n = (int)hc / (devs > 0 ? devs : 1);
            n = n > 12 ? 12 : n < 2 ? 2 : n;
        }
        if (n > 32)
            n = 32;
        for (int i = 1; i < n; i++)
            workers_.emplace_back([this] { worker(); });
        for (auto &w : workers_)
            w.detach();
    }

    /* Frames are written once and next read by a DMA engine or much later by the caller: streaming
     * (non-temporal) stores skip the read-for-ownership of the destination lines, a third of the traffic. */
    static void stream_copy(uint8_t *d, const uint8_t *s, size_t n)
    {
#ifdef SWS_B200_NT_COPY
        if (n >= 4096) {
            const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
            memcpy(d, s, head);
            d += head; s += head; n -= head;
            size_t blocks = n >> 6;
            for (; blocks; blocks--, d += 64, s += 64) {
                const __m128i a = _mm_loadu_si128((const __m128i *)(s));
                const __m128i b = _mm_loadu_si128((const __m128i *)(s + 16));
                const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32));
                const __m128i e = _mm_loadu_si128((const __m128i *)(s + 48));
                _mm_stream_si128((__m128i *)(d), a);
                _mm_stream_si128((__m128i *)(d + 16), b);
                _mm_stream_si128((__m128i *)(d + 32), c);
                _mm_stream_si128((__m128i *)(d + 48), e);
            }
            n &= 63;
        }
#endif
        memcpy(d, s, n);
    }

    static void run_rows(const CopyJob &j, int r0, int r1)
    {
        if (r1 <= r0)
            return;
        if (j.dpitch == (ptrdiff_t)j.rowbytes && j.spitch == (ptrdiff_t)j.rowbytes)
            stream_copy(j.dst + (size_t)r0 * j.rowbytes, j.src + (size_t)r0 * j.rowbytes, (size_t)(r1 - r0) * j.rowbytes);
        else
            for (int r = r0; r < r1; r++)
                stream_copy(j.dst + r * j.dpitch, j.src + r * j.spitch, j.rowbytes);
#ifdef SWS_B200_NT_COPY
        _mm_sfence();
#endif
    }

    static void run_part(const CopyJob &j, int part, int parts)
    {
        const int per = (j.rows + parts - 1) / parts;
        const int r0 = part * per;
        const int r1 = r0 + per < j.rows ? r0 + per : j.rows;
        run_rows(j, r0, r1);
    }

    void worker()
    {
        unsigned long seen = 0;
        for (;;) {
            CopyJob job;
            int part, parts;
            {
                std::unique_lock<std::mutex> lk(mutex_);
                wake_.wait(lk, [&] { return generation_ != seen && next_part_ < parts_; });
                job = job_;
                parts = parts_;
                part = next_part_++;
                if (next_part_ >= parts_)
                    seen = generation_;
            }
            run_part(job, part, parts);
            {
                std::lock_guard<std::mutex> lk(mutex_);
                if (--pending_ == 0)
                    done_.notify_all();
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mutex_, user_mutex_;
    std::condition_variable wake_, done_;
    CopyJob job_{};
    int parts_ = 0, next_part_ = 0, pending_ = 0;
    unsigned long generation_ = 0;
};

} // namespace

/* ------------------------------------------------------------------ device-side row mirror */

/* dst row r = src row (rows - 1 - r); rowbytes need not be a multiple of anything, the pitches are multiples of 16
 * and both bases 16-byte aligned (staging buffers) */
__global__ void __launch_bounds__(256)
sws_flip_rows_kernel(uint8_t *dst, int dpitch, const uint8_t *src, int spitch, int rowbytes, int rows)
{
    const int vec = (rowbytes + 15) >> 4;
    const long long n = (long long)vec * rows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / vec), v = (int)(i - (long long)r * vec);
        const uint4 w = *reinterpret_cast<const uint4 *>(src + (size_t)(rows - 1 - r) * spitch + 16 * v);
        *reinterpret_cast<uint4 *>(dst + (size_t)r * dpitch + 16 * v) = w;
    }
}

/* ------------------------------------------------------------------ staging */

#define RING_DEPTH 3

static void free_staging(SwsCudaState *st)
{
    for (int i = 0; i < 4; i++) {
        cudaFree(st->d_src[i]);
        cudaFree(st->d_dst[i]);
        st->d_src[i] = st->d_dst[i] = nullptr;
        if (st->h_src[i])
            cudaFreeHost(st->h_src[i]);
        if (st->h_dst[i])
            cudaFreeHost(st->h_dst[i]);
        st->h_src[i] = st->h_dst[i] = nullptr;
    }
    if (st->ring_ready) {
        for (int r = 1; r < RING_DEPTH; r++)
            for (int i = 0; i < 4; i++) {
                cudaFree(st->ring_src[r][i]);
                cudaFree(st->ring_dst[r][i]);
            }
        for (int r = 0; r < RING_DEPTH; r++) {
            cudaEventDestroy(st->rev_in[r]);
            cudaEventDestroy(st->rev_k[r]);
            cudaEventDestroy(st->rev_out[r]);
        }
        memset(st->ring_src, 0, sizeof(st->ring_src));
        memset(st->ring_dst, 0, sizeof(st->ring_dst));
        st->ring_ready = 0;
    }
    cudaFree(st->d_flip);
    st->d_flip = nullptr;
    st->flip_bytes = 0;
    st->staging_ready = 0;
    st->bounce_src_ready = st->bounce_dst_ready = 0;
}

static int ensure_staging(SwsCudaState *st)
{
    if (st->staging_ready)
        return 0;
    const SwsCudaPlan *p = &st->plan;
    const int sb = p->src_bits > 8 ? 2 : 1;
    memset(st->src_rows, 0, sizeof(st->src_rows));
    memset(st->dst_rows, 0, sizeof(st->dst_rows));
    /* source planes */
    st->src_rows[0] = p->src_h; st->src_rowbytes[0] = p->src_w * sb;
    if (p->src_layout == SWSC_SRC_RGB) {
        st->src_rowbytes[0] = p->src_w * p->src_bpp;
    } else if (p->src_layout == SWSC_SRC_PLANAR) {
        if (p->has_chroma) {
            st->src_rows[1] = st->src_rows[2] = p->chr_src_h;
            st->src_rowbytes[1] = st->src_rowbytes[2] = p->chr_src_w * sb;
        }
    } else {
        st->src_rows[1] = p->chr_src_h;
        st->src_rowbytes[1] = p->chr_src_w * 2 * sb;
    }
    if (p->src_alpha && p->src_layout != SWSC_SRC_RGB) {      /* a separate alpha plane (packed RGB carries it in plane 0) */
        st->src_rows[3] = p->src_h;
        st->src_rowbytes[3] = p->src_w * sb;
    }
    /* destination planes */
    switch (p->dst_kind) {
    case SWSC_DST_RGB24: case SWSC_DST_BGR24: st->dst_rowbytes[0] = p->dst_w * 3; break;
    case SWSC_DST_RGBA: case SWSC_DST_BGRA: case SWSC_DST_ARGB: case SWSC_DST_ABGR:
        st->dst_rowbytes[0] = p->dst_w * 4; break;
    case SWSC_DST_RGB48: case SWSC_DST_BGR48: st->dst_rowbytes[0] = p->dst_w * 6; break;
    case SWSC_DST_RGBA64: case SWSC_DST_BGRA64: st->dst_rowbytes[0] = p->dst_w * 8; break;
    case SWSC_DST_RGB565: case SWSC_DST_BGR565: case SWSC_DST_RGB555: case SWSC_DST_BGR555:
        st->dst_rowbytes[0] = p->dst_w * 2; break;
    default: {
        const int db = p->dst_bits > 16 ? 4 : p->dst_bits > 8 ? 2 : 1;
        st->dst_rowbytes[0] = p->dst_w * db;
        if (p->dst_kind == SWSC_DST_NV12 || p->dst_kind == SWSC_DST_NV21 || p->dst_kind == SWSC_DST_P010) {
            st->dst_rows[1] = p->chr_dst_h; st->dst_rowbytes[1] = p->chr_dst_w * 2 * db;
        } else if (p->dst_kind == SWSC_DST_GBRP) {
            /* planar RGB: three (four with alpha) full-size planes */
            st->dst_rows[1] = st->dst_rows[2] = p->dst_h;
            st->dst_rowbytes[1] = st->dst_rowbytes[2] = p->dst_w * db;
        } else if (p->dst_has_chroma) {
            st->dst_rows[1] = st->dst_rows[2] = p->chr_dst_h;
            st->dst_rowbytes[1] = st->dst_rowbytes[2] = p->chr_dst_w * db;
        }
        if (p->dst_alpha) {
            st->dst_rows[3] = p->dst_h;
            st->dst_rowbytes[3] = p->dst_w * db;
        }
    } }
    st->dst_rows[0] = p->dst_h;
    for (int i = 0; i < 4; i++) {
        cudaError_t e = cudaSuccess;
        if (st->src_rows[i]) {
            st->d_src_stride[i] = (st->src_rowbytes[i] + 15) & ~15;
            e = cudaMalloc(&st->d_src[i], (size_t)st->d_src_stride[i] * st->src_rows[i]);
        }
        if (e == cudaSuccess && st->dst_rows[i]) {
            st->d_dst_stride[i] = (st->dst_rowbytes[i] + 15) & ~15;
            e = cudaMalloc(&st->d_dst[i], (size_t)st->d_dst_stride[i] * st->dst_rows[i]);
        }
        if (e != cudaSuccess) {
            fprintf(stderr, "[swscaler-b200] staging allocation failed: %s\n", cudaGetErrorString(e));
            cudaGetLastError();
            free_staging(st);                     /* never leave a half-allocated set behind */
            return e == cudaErrorMemoryAllocation ? AVERROR(ENOMEM) : AVERROR(EIO);
        }
    }
    st->staging_ready = 1;
    return 0;
}

/* page-locked twins of the staging planes (same pitches), for pageable caller memory */
static int ensure_bounce(SwsCudaState *st, bool src_side)
{
    int *ready = src_side ? &st->bounce_src_ready : &st->bounce_dst_ready;
    if (*ready)
        return 0;
    for (int i = 0; i < 4; i++) {
        const int rows = src_side ? st->src_rows[i] : st->dst_rows[i];
        const int pitch = src_side ? st->d_src_stride[i] : st->d_dst_stride[i];
        uint8_t **slot = src_side ? &st->h_src[i] : &st->h_dst[i];
        if (!rows || *slot)
            continue;
        if (host_alloc_local((void **)slot, (size_t)pitch * rows) != cudaSuccess) {
            cudaGetLastError();
            *slot = nullptr;
            return AVERROR(ENOMEM);
        }
    }
    *ready = 1;
    return 0;
}

static int ensure_flip(SwsCudaState *st, size_t bytes)
{
    if (st->flip_bytes >= bytes)
        return 0;
    CUDA_OK(cudaStreamSynchronize(st->stream));
    cudaFree(st->d_flip);
    st->d_flip = nullptr;
    st->flip_bytes = 0;
    CUDA_OK(cudaMalloc(&st->d_flip, bytes));
    st->flip_bytes = bytes;
    return 0;
}

/* rows x rowbytes copy; one contiguous DMA when both pitches equal the row size */
static cudaError_t copy_rows_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t rowbytes,
                                   size_t rows, cudaMemcpyKind kind, cudaStream_t s)
{
    if (!rows || !rowbytes)
        return cudaSuccess;
    if (dpitch == rowbytes && spitch == rowbytes)
        return cudaMemcpyAsync(dst, src, rowbytes * rows, kind, s);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, rowbytes, rows, kind, s);
}

/* Host rows -> staging rows.  `src` addresses the FIRST (top-most in picture order) row; a negative pitch
 * means the following rows lie at lower addresses (bottom-up frames, reference swscale.c:1141-1159). */
static int upload_rows(SwsCudaState *st, uint8_t *d, int dpitch, const uint8_t *src, ptrdiff_t spitch,
                       int rowbytes, int rows, cudaStream_t s)
{
    if (rows <= 0)
        return 0;
    NvtxRange nvtx("sws_b200: upload rows");
    if (spitch >= 0 || rows == 1) {
        CUDA_OK(copy_rows_async(d, dpitch, src, (size_t)spitch, rowbytes, rows, cudaMemcpyHostToDevice, s));
        return 0;
    }
    const int fp = (rowbytes + 15) & ~15;
    int ret = ensure_flip(st, (size_t)fp * rows);
    if (ret < 0)
        return ret;
    /* the block of memory holding the rows starts at the last row */
    CUDA_OK(copy_rows_async(st->d_flip, fp, src + (ptrdiff_t)(rows - 1) * spitch, (size_t)-spitch, rowbytes, rows,
                            cudaMemcpyHostToDevice, s));
    const long long n = (long long)(fp >> 4) * rows;
    sws_flip_rows_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(
        d, dpitch, st->d_flip, fp, rowbytes, rows);
    CUDA_OK(cudaGetLastError());
    st->launches++;
    return 0;
}

/* staging rows -> host rows, the mirror of upload_rows() */
static int download_rows(SwsCudaState *st, uint8_t *dst, ptrdiff_t dpitch, const uint8_t *d, int spitch,
                         int rowbytes, int rows, cudaStream_t s)
{
    if (rows <= 0)
        return 0;
    NvtxRange nvtx("sws_b200: download rows");
    if (dpitch >= 0 || rows == 1) {
        CUDA_OK(copy_rows_async(dst, (size_t)dpitch, d, spitch, rowbytes, rows, cudaMemcpyDeviceToHost, s));
        return 0;
    }
    const int fp = (rowbytes + 15) & ~15;
    int ret = ensure_flip(st, (size_t)fp * rows);
    if (ret < 0)
        return ret;
    const long long n = (long long)(fp >> 4) * rows;
    sws_flip_rows_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(
        st->d_flip, fp, d, spitch, rowbytes, rows);
    CUDA_OK(cudaGetLastError());
    st->launches++;
    CUDA_OK(copy_rows_async(dst + (ptrdiff_t)(rows - 1) * dpitch, (size_t)-dpitch, st->d_flip, fp, rowbytes, rows,
                            cudaMemcpyDeviceToHost, s));
    /* d_flip is reused by the next plane: stream order keeps that safe */
    return 0;
}

/* ------------------------------------------------------------------ banded pipeline */

#define E2E_MAX_BANDS 16

static int ensure_pipeline(SwsCudaState *st)
{
    if (st->s_in)
        return 0;
    CUDA_OK(cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < E2E_MAX_BANDS; k++) {
        CUDA_OK(cudaEventCreateWithFlags(&st->ev_in[k], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&st->ev_k[k], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&st->ev_out[k], cudaEventDisableTiming));
    }
    return 0;
}

/* is every plane of a host frame page-locked and mapped into the device address space? */
static bool planes_pinned(const uint8_t *const ptr[4], const int rows[4], uint8_t *dev[4])
{
    bool ok = true;
    for (int i = 0; i < 4 && ok; i++) {
        if (dev)
            dev[i] = nullptr;
        if (!rows[i])
            continue;
        cudaPointerAttributes at;
        if (!ptr[i] || cudaPointerGetAttributes(&at, ptr[i]) != cudaSuccess ||
            at.type != cudaMemoryTypeHost || !at.devicePointer)
            ok = false;
        else if (dev)
            dev[i] = (uint8_t *)at.devicePointer;
    }
    cudaGetLastError();   /* pageable pointers make older runtimes report an error */
    return ok;
}

/* source rows [0, n) of the luma / chroma planes that destination rows [0, y1) read */
static void source_rows_needed(const SwsCudaState *st, int y1, int *need_l, int *need_c)
{
    const SwsCudaPlan *p = &st->plan;
    int nl, nc;
    if (y1 >= p->dst_h) {
        *need_l = p->src_h;
        *need_c = p->chr_src_h;
        return;
    }
    if (p->special || p->unscaled_lut) {
        /* the unscaled converters map rows 1:1 (swscale.c:1161-1187) */
        nl = y1;
        nc = -((-y1) >> p->chr_src_vsub);
    } else {
        const int cy = (y1 - 1) >> p->chr_dst_vsub;
        nl = st->h_vl_pos[y1 - 1] + p->vl_size;
        nc = st->h_vc_pos[cy < p->chr_dst_h ? cy : p->chr_dst_h - 1] + p->vc_size;
    }
    if (p->src_layout == SWSC_SRC_RGB) {
        /* chroma is read from the same packed rows */
        const int l2 = nc << p->chr_src_vsub;
        if (l2 > nl)
            nl = l2;
    }
    nl += 2; nc += 2;                   /* slack: uploading early is harmless */
    *need_l = nl < p->src_h ? (nl < 0 ? 0 : nl) : p->src_h;
    *need_c = nc < p->chr_src_h ? (nc < 0 ? 0 : nc) : p->chr_src_h;
}

/* One synchronous whole-frame conversion as a pipeline of row bands (all strides positive):
 *   copy threads: caller rows -> page-locked ring          [pageable source only]
 *   stream s_in : H2D of band k+1                          (copy engine)
 *   st->stream  : kernel on band k                         (SMs)
 *   stream s_out: D2H of band k-1                          (second copy engine) [or the kernel stores to host]
 *   copy threads: page-locked ring -> caller rows          [pageable destination only]
 * PCIe is full duplex, so a frame costs ~max(H2D, D2H) instead of their sum. */
static int banded_host_frame(SwsCudaState *st, const uint8_t *const src[4], const int src_stride[4],
                             uint8_t *const dst[4], const int dst_stride[4])
{
    const SwsCudaPlan *p = &st->plan;
    int ret;
    NvtxRange nvtx("sws_b200: banded frame (H2D | convert | D2H)");
    if ((ret = ensure_pipeline(st)) < 0)
        return ret;

    uint8_t *dsrc_map[4], *ddst_map[4];
    const bool src_pinned = planes_pinned(src, st->src_rows, dsrc_map);
    const bool dst_pinned = planes_pinned((const uint8_t *const *)dst, st->dst_rows, ddst_map);
    if (!src_pinned && (ret = ensure_bounce(st, true)) < 0)
        return ret;
    if (!dst_pinned && (ret = ensure_bounce(st, false)) < 0)
        return ret;

    /* where the DMA engines read from / write to on the host side */
    const uint8_t *hs[4];
    uint8_t *hd[4];
    int hs_pitch[4], hd_pitch[4];
    for (int i = 0; i < 4; i++) {
        hs[i] = src_pinned ? src[i] : st->h_src[i];
        hs_pitch[i] = src_pinned ? src_stride[i] : st->d_src_stride[i];
        hd[i] = dst_pinned ? dst[i] : st->h_dst[i];
        hd_pitch[i] = dst_pinned ? dst_stride[i] : st->d_dst_stride[i];
    }
    /* the fast420 kernel can store its rows straight into page-locked host memory */
    const bool fast = st->fast_ok && !(st->disabled & 1);
    bool direct_store = fast && st->e2e_mode == 3 && aligned16(hd[0]) && !(hd_pitch[0] & 15);
    if (direct_store && !dst_pinned)
        planes_pinned((const uint8_t *const *)hd, st->dst_rows, ddst_map);   /* device view of our own ring */
    uint8_t *store_dst[4] = { ddst_map[0], ddst_map[1], ddst_map[2], ddst_map[3] };
    if (direct_store && !store_dst[0])
        direct_store = false;

    int bands = st->e2e_bands;
    const size_t frame_bytes = (size_t)st->src_rowbytes[0] * p->src_h + (size_t)st->dst_rowbytes[0] * p->dst_h;
    if (frame_bytes < (2u << 20))
        bands = 1;
    else if (frame_bytes < (8u << 20) && bands > 2)
        bands = 2;
    const int band_unit = 64;             /* a multiple of every tile height and of the 8-row dither period */
    int band_h = ((p->dst_h + bands - 1) / bands + band_unit - 1) / band_unit * band_unit;
    if (band_h < band_unit)
        band_h = band_unit;
    const int nb = (p->dst_h + band_h - 1) / band_h;
    if (nb > E2E_MAX_BANDS)
        return 1;

    CopyPool *pool = (!src_pinned || !dst_pinned) ? &CopyPool::get() : nullptr;
    int64_t zero[4] = { 0, 0, 0, 0 };
    int up[4] = { 0, 0, 0, 0 };                    /* source rows uploaded so far, per plane */
    int drained = 0;                               /* bands already copied out of the ring */

    auto drain_band = [&](int k) -> int {          /* ring -> pageable destination */
        const int y0 = k * band_h, y1 = y0 + band_h < p->dst_h ? y0 + band_h : p->dst_h;
        CUDA_OK(cudaEventSynchronize(st->ev_out[k]));
        for (int i = 0; i < 4; i++) {
            if (!st->dst_rows[i])
                continue;
            const int vs = (i == 1 || i == 2) && st->dst_rows[i] != p->dst_h ? p->chr_dst_vsub : 0;
            const int r0 = y0 >> vs, r1 = y1 == p->dst_h ? st->dst_rows[i] : y1 >> vs;
            if (r1 <= r0)
                continue;
            CopyJob j = { dst[i] + (size_t)r0 * dst_stride[i], st->h_dst[i] + (size_t)r0 * st->d_dst_stride[i],
                          dst_stride[i], st->d_dst_stride[i], (size_t)st->dst_rowbytes[i], r1 - r0 };
            pool->copy(j);
        }
        return 0;
    };

    for (int k = 0; k < nb; k++) {
        const int y0 = k * band_h, y1 = y0 + band_h < p->dst_h ? y0 + band_h : p->dst_h;
        int need[4], nl, nc;
        source_rows_needed(st, y1, &nl, &nc);
        need[0] = nl; need[1] = need[2] = nc; need[3] = nl;
        for (int i = 0; i < 4; i++) {
            if (!st->src_rows[i])
                continue;
            int n = need[i] < st->src_rows[i] ? need[i] : st->src_rows[i];
            if (y1 == p->dst_h)
                n = st->src_rows[i];
            if (n <= up[i])
                continue;
            if (!src_pinned) {
                CopyJob j = { st->h_src[i] + (size_t)up[i] * st->d_src_stride[i], src[i] + (size_t)up[i] * src_stride[i],
                              st->d_src_stride[i], src_stride[i], (size_t)st->src_rowbytes[i], n - up[i] };
                pool->copy(j);
            }
            CUDA_OK(copy_rows_async(st->d_src[i] + (size_t)up[i] * st->d_src_stride[i], st->d_src_stride[i],
                                    hs[i] + (size_t)up[i] * hs_pitch[i], hs_pitch[i],
                                    src_pinned ? (size_t)st->src_rowbytes[i] : (size_t)st->d_src_stride[i],
                                    n - up[i], cudaMemcpyHostToDevice, st->s_in));
            up[i] = n;
        }
        CUDA_OK(cudaEventRecord(st->ev_in[k], st->s_in));
        CUDA_OK(cudaStreamWaitEvent(st->stream, st->ev_in[k], 0));
        if (direct_store) {
            int r = fast420_launch(st, st->d_src, st->d_src_stride, zero, store_dst, hd_pitch, zero, 1, y0, y1, st->stream);
            if (r < 0)
                return r;
            if (r == 0)
                direct_store = false;              /* not eligible after all: take the staged route from here on */
            else
                CUDA_OK(cudaEventRecord(st->ev_out[k], st->stream));
        }
        if (!direct_store) {
            ret = ff_b200_cuda_launch(st, st->d_src, st->d_src_stride, zero, st->d_dst, st->d_dst_stride, zero, 1, y0, y1);
            if (ret < 0)
                return ret;
            CUDA_OK(cudaEventRecord(st->ev_k[k], st->stream));
            CUDA_OK(cudaStreamWaitEvent(st->s_out, st->ev_k[k], 0));
            for (int i = 0; i < 4; i++) {
                if (!st->dst_rows[i])
                    continue;
                const int vs = (i == 1 || i == 2) && st->dst_rows[i] != p->dst_h ? p->chr_dst_vsub : 0;
                const int r0 = y0 >> vs, r1 = y1 == p->dst_h ? st->dst_rows[i] : y1 >> vs;
                if (r1 <= r0)
                    continue;
                CUDA_OK(copy_rows_async(hd[i] + (size_t)r0 * hd_pitch[i], hd_pitch[i],
                                        st->d_dst[i] + (size_t)r0 * st->d_dst_stride[i], st->d_dst_stride[i],
                                        st->dst_rowbytes[i], r1 - r0, cudaMemcpyDeviceToHost, st->s_out));
            }
            CUDA_OK(cudaEventRecord(st->ev_out[k], st->s_out));
        }
        /* drain finished bands while the device works on this one (keep one band of lag) */
        if (!dst_pinned)
            while (drained < k && cudaEventQuery(st->ev_out[drained]) == cudaSuccess) {
                if ((ret = drain_band(drained)) < 0)
                    return ret;
                drained++;
            }
    }
    if (!dst_pinned) {
        for (; drained < nb; drained++)
            if ((ret = drain_band(drained)) < 0)
                return ret;
    } else {
        CUDA_OK(cudaEventSynchronize(st->ev_out[nb - 1]));
    }
    CUDA_OK(cudaStreamSynchronize(st->stream));
    return 0;
}

/* ------------------------------------------------------------------ frame ring */

static int ensure_ring(SwsCudaState *st)
{
    int ret;
    if (st->ring_ready)
        return 0;
    if ((ret = ensure_staging(st)) < 0 || (ret = ensure_pipeline(st)) < 0)
        return ret;
    for (int i = 0; i < 4; i++) {
        st->ring_src[0][i] = st->d_src[i];
        st->ring_dst[0][i] = st->d_dst[i];
    }
    for (int r = 1; r < RING_DEPTH; r++)
        for (int i = 0; i < 4; i++) {
            if (st->src_rows[i])
                CUDA_OK(cudaMalloc(&st->ring_src[r][i], (size_t)st->d_src_stride[i] * st->src_rows[i]));
            if (st->dst_rows[i])
                CUDA_OK(cudaMalloc(&st->ring_dst[r][i], (size_t)st->d_dst_stride[i] * st->dst_rows[i]));
        }
    for (int r = 0; r < RING_DEPTH; r++) {
        CUDA_OK(cudaEventCreateWithFlags(&st->rev_in[r], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&st->rev_k[r], cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&st->rev_out[r], cudaEventDisableTiming));
    }
    st->ring_ready = 1;
    return 0;
}

/* Page-locked host frames first, first + step, ... < nb_frames through this device, nothing but enqueues:
 *   s_in   H2D of frame i into ring slot i mod RING_DEPTH   (after the kernel that last read the slot)
 *   stream kernel of frame i                                  (after its H2D, and after the D2H that last read the slot)
 *   s_out  D2H of frame i                                     (after its kernel)   [absent when the kernel stores to host]
 * The host never blocks here; ff_b200_cuda_frames_wait() drains the device.  One host thread can feed every
 * device of the box this way. */
extern "C" int ff_b200_cuda_frames_enqueue(SwsCudaState *st,
                                           const uint8_t *const src[4], const int src_stride[4], const int64_t src_fstride[4],
                                           uint8_t *const dst[4], const int dst_stride[4], const int64_t dst_fstride[4],
                                           int first, int step, int nb_frames)
{
    const SwsCudaPlan *p = &st->plan;
    DeviceGuard guard(st->device);
    NvtxRange nvtx("sws_b200: enqueue host frames");
    int ret = ensure_ring(st);
    if (ret < 0)
        return ret;
    int64_t zero[4] = { 0, 0, 0, 0 };
    const bool fast = st->fast_ok && !(st->disabled & 1) && st->e2e_mode == 3;
    int slot = 0;
    for (int f = first; f < nb_frames; f += step, slot = (slot + 1) % RING_DEPTH) {
        const uint8_t *s[4];
        uint8_t *d[4];
        for (int i = 0; i < 4; i++) {
            s[i] = src[i] ? src[i] + (src_fstride ? src_fstride[i] : 0) * f : nullptr;
            d[i] = dst[i] ? dst[i] + (dst_fstride ? dst_fstride[i] : 0) * f : nullptr;
        }
        CUDA_OK(cudaStreamWaitEvent(st->s_in, st->rev_k[slot], 0));
        for (int i = 0; i < 4; i++) {
            if (!st->src_rows[i])
                continue;
            if (!s[i] || src_stride[i] <= 0)
                return AVERROR(EINVAL);
            CUDA_OK(copy_rows_async(st->ring_src[slot][i], st->d_src_stride[i], s[i], src_stride[i],
                                    st->src_rowbytes[i], st->src_rows[i], cudaMemcpyHostToDevice, st->s_in));
        }
        CUDA_OK(cudaEventRecord(st->rev_in[slot], st->s_in));
        CUDA_OK(cudaStreamWaitEvent(st->stream, st->rev_in[slot], 0));
        int direct = 0;
        if (fast && d[0] && aligned16(d[0]) && !(dst_stride[0] & 15)) {
            /* page-locked memory is mapped at the same address under unified addressing */
            direct = fast420_launch(st, st->ring_src[slot], st->d_src_stride, zero, d, dst_stride, zero, 1, 0, p->dst_h, st->stream);
            if (direct < 0)
                return direct;
        }
        if (!direct) {
            CUDA_OK(cudaStreamWaitEvent(st->stream, st->rev_out[slot], 0));
            ret = ff_b200_cuda_launch(st, st->ring_src[slot], st->d_src_stride, zero, st->ring_dst[slot], st->d_dst_stride,
                                      zero, 1, 0, p->dst_h);
            if (ret < 0)
                return ret;
        }
        CUDA_OK(cudaEventRecord(st->rev_k[slot], st->stream));
        if (!direct) {
            CUDA_OK(cudaStreamWaitEvent(st->s_out, st->rev_k[slot], 0));
            for (int i = 0; i < 4; i++) {
                if (!st->dst_rows[i])
                    continue;
                if (!d[i] || dst_stride[i] <= 0)
                    return AVERROR(EINVAL);
                CUDA_OK(copy_rows_async(d[i], dst_stride[i], st->ring_dst[slot][i], st->d_dst_stride[i],
                                        st->dst_rowbytes[i], st->dst_rows[i], cudaMemcpyDeviceToHost, st->s_out));
            }
            CUDA_OK(cudaEventRecord(st->rev_out[slot], st->s_out));
        }
    }
    return 0;
}

extern "C" int ff_b200_cuda_frames_wait(SwsCudaState *st)
{
    DeviceGuard guard(st->device);
    CUDA_OK(cudaStreamSynchronize(st->stream));
    if (st->s_out)
        CUDA_OK(cudaStreamSynchronize(st->s_out));
    return 0;
}

/* 1 if every plane the plan reads (source side) / writes (destination side) is page-locked host memory */
extern "C" int ff_b200_cuda_frame_is_pinned(SwsCudaState *st, const uint8_t *const planes[4], int dst_side)
{
    DeviceGuard guard(st->device);
    if (ensure_staging(st) < 0)
        return 0;
    return planes_pinned(planes, dst_side ? st->dst_rows : st->src_rows, nullptr) ? 1 : 0;
}

/* ---- AV_PIX_FMT_CUDA frames of a real libavutil CUDA device (hwcontext_cuda.h: AVCUDADeviceContext) ----
 * The kernels run through the CUDA runtime, i.e. in the PRIMARY context of a device.  A frame pool created by
 * libavutil may live in any CUcontext: find out which device it belongs to and whether it is that device's
 * primary context (then its allocations are ours to address); anything else is refused, not guessed. */
typedef CUresult (*ctx_push_fn)(CUcontext);
typedef CUresult (*ctx_pop_fn)(CUcontext *);
typedef CUresult (*ctx_getdev_fn)(CUdevice *);
typedef CUresult (*pctx_retain_fn)(CUcontext *, CUdevice);
typedef CUresult (*pctx_release_fn)(CUdevice);

static void *driver_entry(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return p;
}

/* >= 0: device ordinal whose primary context `cuda_ctx` is; AVERROR(ENOTSUP): some other context */
extern "C" int ff_b200_cuda_hwctx_device(void *cuda_ctx)
{
    static ctx_push_fn push = (ctx_push_fn)driver_entry("cuCtxPushCurrent");
    static ctx_pop_fn pop = (ctx_pop_fn)driver_entry("cuCtxPopCurrent");
    static ctx_getdev_fn getdev = (ctx_getdev_fn)driver_entry("cuCtxGetDevice");
    static pctx_retain_fn retain = (pctx_retain_fn)driver_entry("cuDevicePrimaryCtxRetain");
    static pctx_release_fn release = (pctx_release_fn)driver_entry("cuDevicePrimaryCtxRelease");
    if (!cuda_ctx || !push || !pop || !getdev || !retain || !release)
        return AVERROR(ENOSYS);
    CUdevice dev = -1;
    CUcontext dummy = nullptr, primary = nullptr;
    if (push((CUcontext)cuda_ctx) != CUDA_SUCCESS)
        return AVERROR(EINVAL);
    const CUresult r = getdev(&dev);
    pop(&dummy);
    if (r != CUDA_SUCCESS)
        return AVERROR(EINVAL);
    if (retain(&primary, dev) != CUDA_SUCCESS)
        return AVERROR(EIO);
    release(dev);
    return primary == (CUcontext)cuda_ctx ? (int)dev : AVERROR(ENOTSUP);
}

extern "C" int ff_b200_cuda_current_device(void)
{
    int d = -1;
    return cudaGetDevice(&d) == cudaSuccess ? d : -1;
}

extern "C" void ff_b200_cuda_use_device(int dev)
{
    if (dev >= 0)
        cudaSetDevice(dev);
}

/* order the context stream after everything queued so far on the producer's stream (AVCUDADeviceContext.stream) */
extern "C" int ff_b200_cuda_wait_stream(SwsCudaState *st, void *producer_stream)
{
    DeviceGuard guard(st->device);
    cudaEvent_t ev;
    CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, (cudaStream_t)producer_stream);
    if (e == cudaSuccess)
        e = cudaStreamWaitEvent(st->stream, ev, 0);
    cudaEventDestroy(ev);
    CUDA_OK(e);
    return 0;
}

extern "C" void *ff_b200_cuda_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void ff_b200_cuda_free(void *p)
{
    cudaFree(p);
}

/* mem: SWS_MEM_SRC_DEVICE -- src[] are DEVICE pointers to whole planes (frame row 0), nothing is uploaded;
 *      SWS_MEM_DST_DEVICE -- dst[] are DEVICE pointers to whole planes, the kernel stores there, nothing is
 *      downloaded.  Cascaded contexts keep their intermediate picture in HBM this way. */
extern "C" int ff_b200_cuda_scale_host(SwsCudaState *st,
                                       const uint8_t *const src[4], const int src_stride[4],
                                       int src_y, int src_h, int upload,
                                       uint8_t *const dst[4], const int dst_stride[4], int y0, int y1, int mem)
{
    const SwsCudaPlan *p = &st->plan;
    DeviceGuard guard(st->device);
    int ret = ensure_staging(st);
    if (ret < 0)
        return ret;
    const bool flipped = mem & SWS_MEM_DST_FLIPPED;
    mem &= SWS_MEM_SRC_DEVICE | SWS_MEM_DST_DEVICE;
    if (mem) {
        const bool sdev = mem & SWS_MEM_SRC_DEVICE, ddev = mem & SWS_MEM_DST_DEVICE;
        if (upload && !sdev)
            for (int i = 0; i < 4; i++) {
                if (!st->src_rows[i])
                    continue;
                if (!src[i])
                    return AVERROR(EINVAL);
                const int vs = (i == 1 || i == 2) ? p->chr_src_vsub : 0;
                const int r0 = src_y >> vs;
                int r1 = -((-(src_y + src_h)) >> vs);
                if (r1 > st->src_rows[i])
                    r1 = st->src_rows[i];
                ret = upload_rows(st, st->d_src[i] + (size_t)r0 * st->d_src_stride[i], st->d_src_stride[i],
                                  src[i], src_stride[i], st->src_rowbytes[i], r1 - r0, st->stream);
                if (ret < 0)
                    return ret;
            }
        if (y1 > y0) {
            int64_t zero[4] = { 0, 0, 0, 0 };
            /* a bottom-up device destination (the first stage of a cascade fed with bottom-up slices): convert into
             * the staging planes, then mirror the rows into place */
            const bool flip = ddev && dst_stride[0] < 0;
            const bool direct = ddev && !flip;
            ret = ff_b200_cuda_launch(st, sdev ? src : (const uint8_t *const *)st->d_src, sdev ? src_stride : st->d_src_stride, zero,
                                      direct ? dst : (uint8_t *const *)st->d_dst, direct ? dst_stride : st->d_dst_stride, zero,
                                      1, y0, y1);
            if (ret < 0)
                return ret;
            if (flip)
                for (int i = 0; i < 4; i++) {
                    if (!st->dst_rows[i] || !dst[i])
                        continue;
                    const int vs = (i == 1 || i == 2) && st->dst_rows[i] != p->dst_h ? p->chr_dst_vsub : 0;
                    const int r0 = y0 >> vs;
                    int r1 = (y1 == p->dst_h) ? st->dst_rows[i] : (y1 >> vs);
                    if (flipped && vs && r1 > (p->dst_h >> vs))
                        r1 = p->dst_h >> vs;
                    if (r1 <= r0)
                        continue;
                    /* picture row r lives at dst[i] + r * stride (stride < 0): rows r1-1 .. r0 in memory order */
                    const long long n = (long long)((st->dst_rowbytes[i] + 15) >> 4) * (r1 - r0);
                    sws_flip_rows_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st->stream>>>(
                        dst[i] + (ptrdiff_t)(r1 - 1) * dst_stride[i], -dst_stride[i],
                        st->d_dst[i] + (size_t)r0 * st->d_dst_stride[i], st->d_dst_stride[i], st->dst_rowbytes[i], r1 - r0);
                    CUDA_OK(cudaGetLastError());
                    st->launches++;
                }
            if (!ddev)
                for (int i = 0; i < 4; i++) {
                    if (!st->dst_rows[i])
                        continue;
                    if (!dst[i])
                        return AVERROR(EINVAL);
                    const int vs = (i == 1 || i == 2) && st->dst_rows[i] != p->dst_h ? p->chr_dst_vsub : 0;
                    const int r0 = y0 >> vs;
                    int r1 = (y1 == p->dst_h) ? st->dst_rows[i] : (y1 >> vs);
                    if (flipped && vs && r1 > (p->dst_h >> vs))
                        r1 = p->dst_h >> vs;
                    if (r1 <= r0)
                        continue;
                    ret = download_rows(st, dst[i] + (ptrdiff_t)r0 * dst_stride[i], dst_stride[i],
                                        st->d_dst[i] + (size_t)r0 * st->d_dst_stride[i], st->d_dst_stride[i],
                                        st->dst_rowbytes[i], r1 - r0, st->stream);
                    if (ret < 0)
                        return ret;
                }
        }
        CUDA_OK(cudaStreamSynchronize(st->stream));
        return 0;
    }

    bool positive = true;
    for (int i = 0; i < 4; i++) {
        if (st->src_rows[i] && upload) {
            if (!src[i])
                return AVERROR(EINVAL);
            if (src_stride[i] <= 0)
                positive = false;
        }
        if (st->dst_rows[i] && y1 > y0) {
            if (!dst[i])
                return AVERROR(EINVAL);
            if (dst_stride[i] <= 0)
                positive = false;
        }
    }

    const bool whole = upload && src_y == 0 && src_h == p->src_h && y0 == 0 && y1 == p->dst_h;
    if (whole && positive && st->e2e_mode) {
        /* e2e_mode: 0 serial; 1 zero-copy (fast420 reads and writes page-locked caller frames directly);
         * 2 banded with D2H copies; 3 (default) banded, fast420 storing straight into the host destination */
        if (st->e2e_mode == 1 && st->fast_ok && !(st->disabled & 1)) {
            uint8_t *ds[4], *dd[4];
            if (planes_pinned(src, st->src_rows, ds) && planes_pinned((const uint8_t *const *)dst, st->dst_rows, dd)) {
                int r = fast420_launch(st, ds, src_stride, nullptr, dd, dst_stride, nullptr, 1, 0, p->dst_h, st->stream);
                if (r < 0)
                    return r;
                if (r == 1) {
                    CUDA_OK(cudaStreamSynchronize(st->stream));
                    return 0;
                }
            }
        }
        int r = banded_host_frame(st, src, src_stride, dst, dst_stride);
        if (r <= 0)
            return r;
    }

    if (upload) {
        for (int i = 0; i < 4; i++) {
            if (!st->src_rows[i])
                continue;
            const int vs = (i == 1 || i == 2) ? p->chr_src_vsub : 0;
            const int r0 = src_y >> vs;
            int r1 = -((-(src_y + src_h)) >> vs);
            if (r1 > st->src_rows[i])
                r1 = st->src_rows[i];
            /* slice pointers address the first row of the slice (swscale.h:566-576) */
            ret = upload_rows(st, st->d_src[i] + (size_t)r0 * st->d_src_stride[i], st->d_src_stride[i],
                              src[i], src_stride[i], st->src_rowbytes[i], r1 - r0, st->stream);
            if (ret < 0)
                return ret;
        }
    }
    if (y1 > y0) {
        int64_t zero[4] = { 0, 0, 0, 0 };
        ret = ff_b200_cuda_launch(st, st->d_src, st->d_src_stride, zero, st->d_dst, st->d_dst_stride, zero, 1, y0, y1);
        if (ret < 0)
            return ret;
        for (int i = 0; i < 4; i++) {
            if (!st->dst_rows[i])
                continue;
            const int vs = (i == 1 || i == 2) && st->dst_rows[i] != p->dst_h ? p->chr_dst_vsub : 0;
            const int r0 = y0 >> vs;
            int r1 = (y1 == p->dst_h) ? st->dst_rows[i] : (y1 >> vs);
            if (flipped && vs && r1 > (p->dst_h >> vs))
                r1 = p->dst_h >> vs;
            if (r1 <= r0)
                continue;
            ret = download_rows(st, dst[i] + (ptrdiff_t)r0 * dst_stride[i], dst_stride[i],
                                st->d_dst[i] + (size_t)r0 * st->d_dst_stride[i], st->d_dst_stride[i],
                                st->dst_rowbytes[i], r1 - r0, st->stream);
            if (ret < 0)
                return ret;
        }
    }
    CUDA_OK(cudaStreamSynchronize(st->stream));
    return 0;
}

#endif /* SWS_B200_XFER_CUH */
