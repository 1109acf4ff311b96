/*
 * swscale_b200_cuda.h -- device-resident extension of the libswscale ABI.
 *
 * The reference has no device-memory entry point on this path: sws_frame_setup()
 * rejects every hardware frame except Vulkan (libswscale/swscale.c:1511-1538).
 * These calls are what a maintainer would bind for AV_PIX_FMT_CUDA frames
 *  and what bench.py uses to time the kernel with the
 * frames already resident in HBM.  Plain C ABI: pointers, ints, no CUDA types.
 */
#ifndef SWSCALE_B200_CUDA_H
#define SWSCALE_B200_CUDA_H

#include <stdint.h>
#include "swscale_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of usable CUDA devices (0 if none / driver missing). */
int sws_cuda_device_count(void);

/*
 * Convert nb_frames device-resident frames with one launch.  Same argument
 * meaning as sws_scale() (reference swscale.h:583) for one whole frame, except
 * that every pointer is a DEVICE pointer and frame f of plane p lives at
 * src[p] + f * srcFrameStride[p] (likewise for dst).  Asynchronous on the
 * context's stream; call sws_cuda_sync() before reading the result.
 * Returns dst_h on success or a negative AVERROR.
 */
int sws_cuda_scale_batch(SwsContext *c,
                         const uint8_t *const src[4], const int srcStride[4],
                         const int64_t srcFrameStride[4],
                         uint8_t *const dst[4], const int dstStride[4],
                         const int64_t dstFrameStride[4], int nb_frames);

/*
 * The same for HOST frames (pageable or page-locked): frame f of plane p lives at src[p] + f * srcFrameStride[p].
 * Synchronous: every frame is in dst when the call returns.  Frames share nothing, so frame f goes to device
 * f mod nb_devices (nb_devices <= 0: every visible device), and on each device `depth` frames are in flight
 * at once (depth <= 0: 3 -- upload of one, kernel of another, download of a third), each on its own stream
 * and staging set; worker threads run next to their device (NUMA).  One process drives the whole box; no
 * collective is involved (BASELINE.json configs[4]: a batch round-robin across 8 GPUs).
 * Returns dst_h on success or a negative AVERROR.
 */
int sws_cuda_scale_batch_host(SwsContext *c,
                              const uint8_t *const src[4], const int srcStride[4],
                              const int64_t srcFrameStride[4],
                              uint8_t *const dst[4], const int dstStride[4],
                              const int64_t dstFrameStride[4], int nb_frames, int nb_devices, int depth);

/* Block until all work queued on the context's stream has finished. */
int sws_cuda_sync(SwsContext *c);

/* The context's cudaStream_t as an opaque pointer (for event timing by the caller). */
void *sws_cuda_stream(SwsContext *c);

/* Kernels launched by this context so far (bench.py's gpu_launches evidence). */
long sws_cuda_launch_count(SwsContext *c);

/* Name of the kernel variant the context selected at init ("fast420_rgb8", "generic", ...). */
const char *sws_cuda_kernel_name(SwsContext *c);

/* Last error text recorded on the context ("" if none). */
const char *sws_cuda_last_error(SwsContext *c);

/* Page-locked host memory for callers that want DMA-speed sws_scale() (optional). */
void *sws_cuda_host_alloc(size_t size);
void  sws_cuda_host_free(void *ptr);

/* NUMA placement on multi-socket hosts.  sws_cuda_host_alloc() already places its pages on the NUMA node of
 * the current CUDA device; sws_cuda_bind_thread_to_device() moves the calling thread next to a device as well
 * (returns the number of CPUs in the new affinity set, -1 if the topology is unknown: nothing changed).
 * SWS_B200_NUMA=0 disables both. */
int sws_cuda_bind_thread_to_device(int device);
int sws_cuda_device_numa_node(int device);

/* ---- diagnostics (used by the CPU-only test-suite; no device is touched) ----
 * sws_b200_plan_only(): run everything sws_init_context() (reference utils.c:1884)
 * does on the host -- path selection, geometry, FIR banks, colour constants --
 * but skip the device upload.  sws_b200_get_filter(): which = 0 hLum, 1 hChr,
 * 2 vLum, 3 vChr; returns the tap count, coef is [len][taps] int16, pos [len]
 * (the tables of reference swscale_internal.h:437-448).  sws_b200_get_info():
 * out[0..5] = yuv2rgb y_offset,y_coeff,v2r,v2g,u2g,u2b (yuv2rgb.c:786-791), ...
 * sws_b200_get_rgb2yuv(): {ry,gy,by,ru,gu,bu,rv,gv,bv}, the C entries of the reference's
 * input_rgb2yuv_table (swscale_internal.h:467-476, utils.c:614-706); RGB sources only. */
int sws_b200_plan_only(SwsContext *c);
int sws_b200_plan_only_filtered(SwsContext *c, SwsFilter *srcFilter, SwsFilter *dstFilter);
int sws_b200_get_filter(SwsContext *c, int which, const int16_t **coef, const int32_t **pos, int *len);
int sws_b200_get_info(SwsContext *c, int out[32]);
int sws_b200_get_rgb2yuv(SwsContext *c, int out[9]);

#ifdef __cplusplus
}
#endif
#endif
