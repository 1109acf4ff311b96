#!/usr/bin/env python3
"""bench.py -- the libswscale hot-path benchmark (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): 3840x2160 yuv420p -> rgb24, SWS_BICUBIC |
SWS_BITEXACT | SWS_ACCURATE_RND -- BASELINE.json configs[4], the configuration
the metric is quoted on; one 4K frame is 37.3 MB so it fits one GPU trivially.
A step = one pass of the hot path over FRAMES_PER_GPU distinct synthetic frames
(64 per GPU => 512 frames at 8 GPUs, the reference batch of configs[4]); frames
are partitioned across ranks with no data-path collective (weak scaling).

value   : Mpixels/s, whole job, frames already resident in HBM (one batched
          launch per step through sws_cuda_scale_batch()).
e2e     : Mpixels/s through the reference-facing call sws_scale() with HOST
          (page-locked) buffers -- H2D + kernel + D2H inside the timed region.
e2e_pageable   : the same call on plain pageable numpy buffers (what av_frame_get_buffer() hands a caller).
e2e_batch_host : sws_cuda_scale_batch_host(), several host frames in flight per device.
configs : BASELINE.json configs[0..3] (C1..C4) and the scaler's widening rows (X1..X3, X6, X8, E1, E2, C3b) device-resident,
          same method as `value` (N=1 only).
roofline: algorithmic bytes (4.5 B/pixel) / average kernel
          duration, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline: the real reference C path (oracle/_ref) on the host cores, bounded sample.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 3840, 2160
FRAMES_PER_GPU = 64          # 64 x 37.3 MB = 2.39 GB per step per GPU  (>> 126 MB L2)
E2E_FRAMES = 16
ALG_BYTES_PER_PIXEL = 4.5    # 1.5 B in (yuv420p) + 3 B out (rgb24)
WORKLOAD = "3840x2160 yuv420p->rgb24 SWS_BICUBIC|SWS_BITEXACT|SWS_ACCURATE_RND"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref, built from the
    unmodified /root/reference sources) with all host threads, on the same config."""
    if rank != 0:
        return
    from oracle import refapi as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libswsref.so not built"}))
        return
    import numpy as np
    cores = host_cores()
    ctx = R.RefContext(W, H, "yuv420p", W, H, "rgb24", R.SWS_BICUBIC | R.BX, threads=cores)
    ring, dsts, srcs = _ref_ring(R, np)
    frames_per_step = 8          # bounded sample of the 64-frame step
    L = R.lib()
    for _ in range(max(args.warmup, 1)):
        L.swsref_bench_frames(ctx.h, dsts, srcs, ring, ring)
    t = 0.0
    for _ in range(args.steps):
        dt = L.swsref_bench_frames(ctx.h, dsts, srcs, ring, frames_per_step)
        if dt < 0:
            raise RuntimeError("reference sws_scale_frame failed")
        t += dt
    mpix = frames_per_step * args.steps * W * H / t / 1e6
    line = {
        "impl": "reference", "metric": "Mpixels/s", "value": mpix, "unit": "Mpixels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": FRAMES_PER_GPU,
                   "bytes_per_step_per_gpu": int(ALG_BYTES_PER_PIXEL * FRAMES_PER_GPU * W * H),
                   "note": "reference C path (sws_scale_frame, slice-threaded) on the host cores; each timed step "
                           "is a bounded sample (%d frames out of a ring of %d distinct ones) of the %d-frame step; "
                           "generic C build (no hand-written SIMD is selected on the bit-exact accurate_rnd path)"
                           % (frames_per_step, ring, FRAMES_PER_GPU)},
        "cpu_baseline": {"value": mpix, "unit": "Mpixels/s", "cores": cores, "kind": "reference",
                         "sample": "%d frames x %d steps over a ring of %d distinct 4K frames, threads=%d"
                                   % (frames_per_step, args.steps, ring, cores)},
        "e2e": {"value": mpix, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


_REF_KEEP = []


def _ref_ring(R, np, ring=8):
    """`ring` distinct 4K source/destination AVFrame pairs (300 MB: beyond the last-level cache) for the
    reference arm, as ctypes arrays of frame pointers."""
    rng = np.random.default_rng(1234)
    srcs, dsts = [], []
    for _ in range(ring):
        s, d = R.RefFrame(W, H, "yuv420p"), R.RefFrame(W, H, "rgb24")
        for i, rows in enumerate((H, H // 2, H // 2)):
            a, ls = s.plane(i, rows)
            a[:] = rng.integers(0, 256, a.shape, dtype=np.uint8)
        srcs.append(s)
        dsts.append(d)
    _REF_KEEP.extend(srcs + dsts)
    L = R.lib()
    L.swsref_bench_frames.restype = ctypes.c_double
    L.swsref_bench_frames.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]
    return ring, (ctypes.c_void_p * ring)(*[f.f for f in dsts]), (ctypes.c_void_p * ring)(*[f.f for f in srcs])


def cpu_baseline_sample(seconds=12.0):
    """rank 0, N=1: the reference on the host cores, ~10-30 s of CPU work."""
    from oracle import refapi as R
    if not R.available():
        return None
    import numpy as np
    cores = host_cores()
    out = {}
    ring, dsts, srcs = _ref_ring(R, np)
    L = R.lib()
    for label, threads in (("threads_all", cores), ("threads_1", 1)):
        ctx = R.RefContext(W, H, "yuv420p", W, H, "rgb24", R.SWS_BICUBIC | R.BX, threads=threads)
        L.swsref_bench_frames(ctx.h, dsts, srcs, ring, 1)
        n, t = 0, 0.0
        budget = seconds * (0.7 if threads > 1 else 0.3)
        while t < budget:
            k = 8 if threads > 1 else 1
            t += L.swsref_bench_frames(ctx.h, dsts, srcs, ring, k)
            n += k
        out[label] = (n * W * H / t / 1e6, n, t)
        ctx.close()
    return {"value": out["threads_all"][0], "unit": "Mpixels/s", "cores": cores, "kind": "reference",
            "sample": "%d 4K frames (ring of %d distinct ones) in %.1f s with %d threads (sws_scale_frame, "
                      "bitexact+accurate_rnd generic-C path: no hand-written SIMD is selected there); "
                      "1 thread: %.1f Mpixels/s" % (out["threads_all"][1], ring, out["threads_all"][2], cores,
                                                    out["threads_1"][0])}


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- our arm
def run_b200_arm(args, rank, local_rank, world):
    import numpy as np
    import torch
    from librempeg_b200 import swscale as S

    if not torch.cuda.is_available() or S.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # run next to the GPU: this rank's thread and every page-locked frame it allocates live on the NUMA node
    # the device hangs off (SWS_B200_NUMA=0 switches it off for A/B runs)
    numa = {"node": S.device_numa_node(local_rank), "cpus_bound": S.bind_thread_to_device(local_rank)}
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    flags = S.SWS_BICUBIC | S.BX
    ctx = S.SwsContext(W, H, "yuv420p", W, H, "rgb24", flags)
    F = FRAMES_PER_GPU
    ysz, csz, osz = W * H, (W // 2) * (H // 2), W * H * 3
    # distinct synthetic frames, seed depends on the global frame index
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    src_y = torch.randint(0, 256, (F, ysz), dtype=torch.uint8, device=dev, generator=g)
    src_u = torch.randint(0, 256, (F, csz), dtype=torch.uint8, device=dev, generator=g)
    src_v = torch.randint(0, 256, (F, csz), dtype=torch.uint8, device=dev, generator=g)
    dst = torch.zeros((F, osz), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step():
        r = ctx.scale_batch_device([src_y, src_u, src_v], [W, W // 2, W // 2], [ysz, csz, csz],
                                   [dst], [W * 3], [osz], F)
        if r < 0:
            raise RuntimeError("sws_cuda_scale_batch failed: %d %s" % (r, ctx.last_error))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # parity spot-check before any number is reported: frame 0 vs the oracle (rank 0)
    parity = None
    if rank == 0:
        try:
            from oracle import refapi as R
            if R.available():
                rc = R.RefContext(W, H, "yuv420p", W, H, "rgb24", R.SWS_BICUBIC | R.BX)
                hy, hu, hv = src_y[0].cpu().numpy(), src_u[0].cpu().numpy(), src_v[0].cpu().numpy()
                want = np.zeros(osz, np.uint8)
                rc.scale([hy, hu, hv], [W, W // 2, W // 2], [want], [W * 3])
                parity = bool(np.array_equal(want, dst[0].cpu().numpy()))
                rc.close()
                if not parity:
                    raise AssertionError("bench: CUDA output differs from the reference; refusing to time")
        except ImportError:
            pass

    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- e2e: sws_scale() on HOST frames, H2D + kernel + D2H timed ----
    EF = E2E_FRAMES
    host_src = src_y[:EF].cpu().numpy(), src_u[:EF].cpu().numpy(), src_v[:EF].cpu().numpy()
    sizes = (ysz, csz, csz, osz)

    def host_buffers(pinned):
        if pinned:          # sws_cuda_host_alloc(): page-locked, on the NUMA node of this rank's device
            bufs = [S.PinnedBuffer(EF * n) for n in sizes]
            arrs = [b.array.reshape(EF, n) for b, n in zip(bufs, sizes)]
        else:               # plain pageable memory, what av_frame_get_buffer() / malloc hand a caller
            bufs = None
            arrs = [np.empty((EF, n), dtype=np.uint8) for n in sizes]
        for a, h in zip(arrs[:3], host_src):
            a[:] = h
        arrs[3][:] = 0
        return bufs, arrs

    def time_e2e(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        n = max(1, min(args.steps, 5))
        for _ in range(n):
            fn()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * EF * n * W * H / float(te.item()) / 1e6

    def per_frame(arrs):
        ptrs = [[a[f].ctypes.data for a in arrs] for f in range(EF)]

        def fn():
            for f in range(EF):
                r = ctx.scale(ptrs[f][:3], [W, W // 2, W // 2], [ptrs[f][3]], [W * 3], 0, H)
                if r != H:
                    raise RuntimeError("sws_scale failed: %d %s" % (r, ctx.last_error))
        return fn

    def batched(arrs):
        def fn():
            r = ctx.scale_batch_host([a.ctypes.data for a in arrs[:3]], [W, W // 2, W // 2], [ysz, csz, csz],
                                     [arrs[3].ctypes.data], [W * 3], [osz], EF, 1, 0)
            if r != H:
                raise RuntimeError("sws_cuda_scale_batch_host failed: %d %s" % (r, ctx.last_error))
        return fn

    want0 = dst[0].cpu().numpy() if (rank == 0 and parity is not None) else None
    pin_bufs, pin_arrs = host_buffers(True)
    e2e_mpix = time_e2e(per_frame(pin_arrs))
    if want0 is not None:       # the host path must give the same bytes as the device path
        assert np.array_equal(pin_arrs[3][0], want0), "host and device paths disagree"
    pin_arrs[3][:] = 0
    e2e_batch = time_e2e(batched(pin_arrs))
    if want0 is not None:
        assert np.array_equal(pin_arrs[3][0], want0), "batched host path and device path disagree"
    _, pag_arrs = host_buffers(False)
    e2e_pageable = time_e2e(per_frame(pag_arrs))
    if want0 is not None:
        assert np.array_equal(pag_arrs[3][0], want0), "pageable host path and device path disagree"
    pin_arrs = None
    for b in pin_bufs:
        b.close()

    # ---- what the box allows: page-locked frames copied in and out at the same time on every rank at once, no kernel.
    # The host-frame numbers above cannot exceed  min(H2D / bytes in, D2H / bytes out)  frames per second. ----
    def host_ceiling(seconds=0.6):
        hin = torch.empty(ysz + 2 * csz, dtype=torch.uint8).pin_memory()
        hout = torch.empty(osz, dtype=torch.uint8).pin_memory()
        din = torch.empty(ysz + 2 * csz, dtype=torch.uint8, device=dev)
        dout = torch.empty(osz, dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        barrier()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                with torch.cuda.stream(s_in):
                    din.copy_(hin, non_blocking=True)
                with torch.cuda.stream(s_out):
                    hout.copy_(dout, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            n += 8
        dt = time.perf_counter() - t0
        t = torch.tensor([n * hin.numel() / dt / 1e9, n * hout.numel() / dt / 1e9, n * W * H / dt / 1e6],
                         dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {"value": float(t[2].item()), "unit": "Mpixels/s", "h2d_gbs": float(t[0].item()), "d2h_gbs": float(t[1].item()),
                "how": "%d rank(s) at once: page-locked 4K frames copied in (12.4 MB) and out (24.9 MB) concurrently for "
                       "%.1f s, no kernel; sum over ranks" % (world, seconds)}
    ceiling = host_ceiling()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    step_ms = ms_max / args.steps
    pixels_per_step = F * W * H
    value = world * pixels_per_step / (step_ms * 1e-3) / 1e6
    kernel_ms = ms / max(launches, 1)          # rank-0 kernel: one launch per step
    bytes_per_launch = ALG_BYTES_PER_PIXEL * pixels_per_step * (args.steps / max(launches, 1))
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9

    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(ctx.kernel_name)
    except Exception:
        pass

    # ---- the other BASELINE configurations, device-resident, same method (N=1: the box is otherwise idle) ----
    configs = None
    if world == 1 and not args.no_configs:
        from tools import bench_configs as BC
        configs = {}
        for cfg in BC.CONFIGS:
            tag = cfg[0].split()[0]
            # C1..C4: BASELINE.json configs[0..3]; the rest: the scaler's widening rows (up / down scaling, 10-bit
            # and packed-RGB sources) and the encode-side conversion, so that they carry driver-timed numbers too
            if tag not in ("C1", "C2", "C3", "C4", "C3b", "X1", "X2", "X3", "X6", "X8", "E1", "E2"):
                continue
            r = BC.measure(cfg, dev, peak, steps=max(5, min(args.steps, 20)))
            configs[tag] = {"workload": r["name"], "kernel": r["kernel"], "frames_per_step": r["frames"],
                            "ms_per_step": r["ms"], "value": r["mpix_in"], "unit": "Mpixels/s (source pixels)",
                            "mpix_out": r["mpix_out"],
                            "roofline": {"bound": "hbm", "achieved": r["gbs"], "peak": peak, "unit": "GB/s",
                                         "frac": r["frac"], "algorithmic_bytes_per_launch": r["algorithmic_bytes"]}}

    line = {
        "metric": "Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F,
                   "bytes_per_step_per_gpu": int(ALG_BYTES_PER_PIXEL * pixels_per_step),
                   "l2_policy": "inputs larger than L2: %d distinct frames (%.2f GB) cycled every step"
                                % (F, ALG_BYTES_PER_PIXEL * pixels_per_step / 1e9),
                   "partition": "frames round-robin over ranks, no collective",
                   "kernel": ctx.kernel_name, "parity_checked_vs_reference": parity},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                       "ncu --set full capture of this launch shape; not measured in this run)",
                     "peak_source": peak_src,
                     "kernel": ctx.kernel_name, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": bytes_per_launch},
        "e2e": {"value": e2e_mpix, "unit": "Mpixels/s",
                "h2d_bytes_per_step": int(EF * (ysz + 2 * csz)), "d2h_bytes_per_step": int(EF * osz),
                "frames_per_step": EF,
                "call": "sws_scale() per frame, page-locked host buffers (sws_cuda_host_alloc)"},
        "e2e_pageable": {"value": e2e_pageable, "unit": "Mpixels/s", "frames_per_step": EF,
                         "call": "sws_scale() per frame, pageable numpy buffers (bounce ring + copy threads)"},
        "e2e_batch_host": {"value": e2e_batch, "unit": "Mpixels/s", "frames_per_step": EF,
                           "call": "sws_cuda_scale_batch_host(), page-locked buffers, 3 frames in flight"},
        "e2e_ceiling": ceiling,
        "numa": numa,
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if configs:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline_sample()
        if cb:
            line["cpu_baseline"] = cb
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
