#!/usr/bin/env python
"""bench.py -- scans/sec of the lidar registration hot path (scanRegistration -> laserOdometry ->
laserMapping) on synthetic HDL-64E sweeps against a planted ~1M-point cube map (BASELINE.json
config C3: "HDL-64E scan-to-map laserMapping against ~1M-point local cube map").

One "step" = one sweep through the whole per-frame chain (MAIN.cpp:143-144, 186-190).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference [...]                       the CPU restatement of the reference
                                                               (oracle/, KD-tree mode) on the host cores

Under torchrun (N > 1) every rank replays its own independent sequence (weak scaling, no data-path
collective); poses and timings are gathered with one NCCL all_gather at the end.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SENSOR = 1  # HDL-64E
WORKLOAD = "C3: synthetic HDL-64E sweeps (~118k pts), full scanRegistration+laserOdometry+laserMapping per sweep, planted ~1M-point cube map"


def make_sequence(pkg, seq_id, n_frames, world_kind=1):
    synth = pkg.synth
    world = synth.World(1234, world_kind, 190.0)
    traj = synth.trajectory(n_frames, seed=77 + seq_id)
    scans = [world.scan(SENSOR, traj[k], 1000 + 7919 * seq_id + k) for k in range(n_frames)]
    corner = world.plant(0, 0.4, seed=99)
    surf = world.plant(1, 0.8, seed=98)
    cblob, sblob = synth.cubes_blob(corner, 0.4), synth.cubes_blob(surf, 0.8)
    return scans, traj, cblob, sblob


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions.  In-process NVML (nvidia_ml_py): a
    spawned `nvidia-smi -lms` stalls CUDA calls for tens of ms every time it polls, which on a ~100 ms
    timed region is the difference between 900 and 450 scans/s; NVML calls from a thread cost ~0.1 ms."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.thread, self.nv = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        period = float(os.environ.get("BENCH_SMI_MS", "50")) / 1e3
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, reasons))
            except Exception:
                pass
            time.sleep(period)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.nv or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        seen = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.sm_max), "reasons": seen,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def run_reference(args, rank, world_size):
    """CPU arm: the oracle in baseline mode (KD-tree kNN, std::sort voxel grids, restated Ceres LM),
    one single-threaded replica per host core like the reference (MAIN.cpp:277-282)."""
    if rank != 0:
        return
    pkg = importlib.import_module("vloam-noted_b200")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as op
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    replicas = max(1, min(cores, 16))
    budget_s = 200.0
    n_frames = args.warmup + args.steps + 1
    scans, traj, cblob, sblob = make_sequence(pkg, 0, min(n_frames, 64))

    def replica(out, idx):
        o = op.Oracle(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, knn_backend=1)
        o.set("lm.cornerMap", cblob); o.set("lm.surfMap", sblob)
        times = []
        t_start = time.perf_counter()
        for k in range(n_frames):
            s = scans[k % len(scans)]
            t0 = time.perf_counter()
            o.process(s)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_start > budget_s:
                break
        out[idx] = times

    results = [None] * replicas
    th = [threading.Thread(target=replica, args=(results, i)) for i in range(replicas)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    done = min(len(r) for r in results)
    w = min(args.warmup + 1, max(done - 1, 0))
    timed = done - w
    per_rep = [sum(r[w:done]) for r in results]
    wall = max(per_rep)
    value = replicas * timed / wall if wall > 0 else 0.0
    single = timed / float(np.mean(per_rep)) if timed else 0.0
    line = {
        "impl": "reference", "metric": "scans/sec", "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": timed, "warmup": w,
        "ms_per_step": 1e3 * wall / max(timed, 1) / replicas, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "map_points": int(pkg.synth.blob_counts(cblob).sum() + pkg.synth.blob_counts(sblob).sum())},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": replicas, "kind": "port",
                         "sample": "%d frames x %d single-threaded replicas of oracle/ (KD-tree mode); one replica alone: %.3f scans/s; p50 %.1f ms/frame"
                                   % (timed, replicas, single, 1e3 * float(np.median(np.concatenate([r[w:done] for r in results])))) if timed else "none"},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(pkg, scans, cblob, sblob, frames=4):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as op
    o = op.Oracle(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, knn_backend=1)
    o.set("lm.cornerMap", cblob); o.set("lm.surfMap", sblob)
    times, stages = [], []
    for k in range(frames + 1):
        t0 = time.perf_counter()
        o.process(scans[k])
        times.append(time.perf_counter() - t0)
        stages.append(o.get("timing")[:3])
    t = times[1:]
    st = np.mean(np.array(stages[1:]), axis=0)
    return {"value": len(t) / sum(t), "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": "%d frames of the same workload after 1 warm-up frame, single thread like the reference; mean ms SR/LO/LM = %.1f/%.1f/%.1f"
                      % (len(t), st[0], st[1], st[2])}



def _time_frames(ctx, torch, ext, frames, warm, fn):
    """Device time of `fn(k)` over frames[warm:], CUDA events on the context's stream; returns ms per frame."""
    for k in range(warm):
        fn(k)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(warm, frames):
        fn(k)
    ctx.synchronize()
    e1.record(ext)
    ctx.synchronize(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / max(frames - warm, 1)


def other_workloads(pkg, torch, local_rank, frames=70, warm=10):
    """BASELINE.json's other single-GPU configurations, reported beside the headline (north_star: "throughput on
    synthetic 16/64/128-beam scans is reported at 1 GPU"): device-resident sweeps, device time per sweep.
      C1  VLP-16 (~25k kept points), scan-to-scan odometry + mapping from an empty map
      C2  HDL-64E, scanRegistration + laserOdometry only (no mapping call)
      C4  OS1-128 (~257k points) against a planted ~2.2M-point map, full chain"""
    synth = pkg.synth
    out = {}
    traj = synth.trajectory(frames, seed=78)

    def run(name, world, sensor, ctx, blobs, mapping):
        scans = [world.scan(sensor, traj[k], 5000 + k) for k in range(frames)]
        if blobs:
            ctx.set("lm.cornerMap", blobs[0]); ctx.set("lm.surfMap", blobs[1])
        d = [torch.from_numpy(x).cuda(local_rank) for x in scans]
        ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
        if mapping:
            fn = lambda k: ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
        else:
            def fn(k):
                ctx.begin_frame(); ctx.scan_registration_device(d[k].data_ptr(), d[k].shape[0], 4); ctx.laser_odometry(want_pose=False)
        ms = _time_frames(ctx, torch, ext, frames, warm, fn)
        out[name] = {"scans_per_s": 1e3 / ms, "ms_per_sweep": ms, "points_per_sweep": int(np.mean([len(x) for x in scans])),
                     "map_points": int(sum(synth.blob_counts(b).sum() for b in blobs)) if blobs else 0, "stages": "SR+LO+LM" if mapping else "SR+LO"}
        ctx.close()

    run("C1_vlp16", synth.World(1234, 0, 160.0), 0, pkg.Context(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4, device=local_rank), None, True)
    run("C2_hdl64_sr_lo", synth.World(1234, 1, 190.0), 1, pkg.Context(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, device=local_rank), None, False)
    w4 = synth.World(1234, 2, 190.0)
    blobs4 = (synth.cubes_blob(w4.plant(0, 0.4, seed=99), 0.4), synth.cubes_blob(w4.plant(1, 0.8, seed=98), 0.8))
    run("C4_os1_128", w4, 2, pkg.Context(n_scans=128, minimum_range=0.3, line_res=0.4, plane_res=0.8, device=local_rank), blobs4, True)
    return out


def batched_sequences(pkg, torch, local_rank, cblob, sblob, nseq=4, frames=70, warm=10, first_seq=0, seq_stride=1):
    """BASELINE config C5 on one GPU: `nseq` independent sequences, one context (4 streams + its helper thread) and
    one host thread each, replayed concurrently.  A single sequence leaves the GPU mostly idle (the frame is a
    chain of short dependent kernels), so concurrent sequences overlap almost freely."""
    seqs = []
    for q in range(nseq):
        world = pkg.synth.World(1234, 1, 190.0)
        sid = first_seq + q * seq_stride
        traj = pkg.synth.trajectory(frames, seed=177 + sid)
        scans = [world.scan(SENSOR, traj[k], 9000 + 7919 * sid + k) for k in range(frames)]
        ctx = pkg.Context(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, device=local_rank)
        ctx.set("lm.cornerMap", cblob); ctx.set("lm.surfMap", sblob)
        seqs.append((ctx, [torch.from_numpy(x).cuda(local_rank) for x in scans]))
    lat = [[] for _ in range(nseq)]
    barrier = threading.Barrier(nseq + 1)

    def worker(q):
        ctx, d = seqs[q]
        torch.cuda.set_device(local_rank)
        pose = np.zeros(14)
        for k in range(warm):
            ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4, pose.ctypes.data)
        ctx.synchronize()
        barrier.wait(); barrier.wait()
        for k in range(warm, frames):
            t1 = time.perf_counter()
            ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4, pose.ctypes.data)
            lat[q].append(time.perf_counter() - t1)
        ctx.synchronize()
        barrier.wait()

    th = [threading.Thread(target=worker, args=(q,)) for q in range(nseq)]
    for t in th: t.start()
    barrier.wait()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    barrier.wait()
    barrier.wait()
    e1.record(); torch.cuda.synchronize()
    for t in th: t.join()
    ms = e0.elapsed_time(e1)
    for ctx, _ in seqs: ctx.close()
    allat = np.concatenate([np.array(x) for x in lat]) * 1e3
    return {"sequences_per_gpu": nseq, "device_ms": ms, "scans_per_s": nseq * (frames - warm) / (ms * 1e-3), "p50_ms_per_frame": float(np.median(allat)),
            "p99_ms_per_frame": float(np.percentile(allat, 99)), "frames_per_sequence": frames - warm,
            "note": "every frame returns its pose to its own host thread (sync per frame); timed with CUDA events on the default stream around all threads"}


def cold_l2_frames(ctx, torch, local_rank, dscans, first, n=40):
    """Same chain with the 126 MB L2 flushed before every sweep (a 512 MB fill): per-sweep device time, median."""
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:%d" % local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    ms = []
    for k in range(first, first + n):
        ctx.synchronize()
        flush.fill_(k & 255); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        ctx.process_frame_device(dscans[k].data_ptr(), dscans[k].shape[0], 4)
        ctx.synchronize()
        e1.record(ext); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return {"ms_per_step_median": float(np.median(ms)), "frames": n, "flush": "512 MB device fill before every sweep; includes the map update and the speculative sub-map build"}


def run_ours(args, rank, world_size, local_rank):
    import torch
    pkg = importlib.import_module("vloam-noted_b200")
    pkg.load_lib()  # raises if the CUDA library is missing: there is no fallback
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)
    n_frames = W + K + 1
    scans, traj, cblob, sblob = make_sequence(pkg, rank, n_frames)
    map_points = int(pkg.synth.blob_counts(cblob).sum() + pkg.synth.blob_counts(sblob).sum())

    def fresh():
        ctx = pkg.Context(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, device=local_rank)
        ctx.set("lm.cornerMap", cblob); ctx.set("lm.surfMap", sblob)
        return ctx

    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("BENCH_SMI_MS", "50") != "0":
        sampler.start()
    # ---- leg 1: device-resident inputs (value) ------------------------------------------------
    ctx = fresh()
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dscans = [torch.from_numpy(s).cuda(local_rank) for s in scans]
    torch.cuda.synchronize()
    # a replay knows the next sweep: registering it lets its scan registration run underneath this sweep's odometry
    # and mapping (vloam_b200_prefetch_scan_device); the sweep timed first was registered during the warm-up
    for k in range(W + 1):
        ctx.prefetch_device(dscans[k + 1].data_ptr(), dscans[k + 1].shape[0], 4)
        ctx.process_frame_device(dscans[k].data_ptr(), dscans[k].shape[0], 4)
    ctx.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(W + 1, W + 1 + K):
        if k + 1 < len(dscans): ctx.prefetch_device(dscans[k + 1].data_ptr(), dscans[k + 1].shape[0], 4)
        ctx.process_frame_device(dscans[k].data_ptr(), dscans[k].shape[0], 4)
    ctx.synchronize()  # the last frame's map update runs on a side stream: include it
    e1.record(ext)
    ctx.synchronize(); torch.cuda.synchronize()
    if dist: dist.barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    # dominant-kernel timing (CUDA events around that kernel's launches, on the launching stream)
    roof, ktable = profile_dominant(ctx, dscans, W + 1, K, map_points)
    extras = {}
    if world_size == 1 and not args.no_extras:
        extras["cold_l2"] = cold_l2_frames(ctx, torch, local_rank, dscans, W + 1, min(K, 40))
    ctx.close()
    if world_size == 1 and not args.no_extras:
        extras["batched"] = batched_sequences(pkg, torch, local_rank, cblob, sblob)
        extras["workloads"] = other_workloads(pkg, torch, local_rank)

    # ---- leg 2: end to end through the C ABI with host buffers (e2e) --------------------------
    ctx = fresh()
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]
    pose = np.zeros(14)
    lat = []
    for k in range(W + 1):
        ctx.prefetch_ptr(pinned[k + 1].data_ptr(), pinned[k + 1].shape[0], 4)
        ctx.process_frame_ptr(pinned[k].data_ptr(), pinned[k].shape[0], 4, pose.ctypes.data)
    if dist: dist.barrier()
    t0 = time.perf_counter()
    poses = np.zeros((K, 14))
    for i, k in enumerate(range(W + 1, W + 1 + K)):
        t1 = time.perf_counter()
        if k + 1 < len(pinned): ctx.prefetch_ptr(pinned[k + 1].data_ptr(), pinned[k + 1].shape[0], 4)  # upload + scan registration of the next sweep overlap this one
        ctx.process_frame_ptr(pinned[k].data_ptr(), pinned[k].shape[0], 4, pose.ctypes.data)
        lat.append(time.perf_counter() - t1)
        poses[i] = pose
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ctx.close()
    h2d = float(np.mean([s.shape[0] * 16 for s in scans[W + 1:W + 1 + K]]))

    # trajectory sanity against the generator's ground truth (frame W+K)
    err = float(np.linalg.norm(poses[-1, 11:14] - traj[W + K][:3]))

    # ---- max over ranks; one collective: gather of poses + timings ----------------------------
    lat_ms = np.array(lat) * 1e3
    slow = [(int(i), round(float(lat_ms[i]), 2)) for i in np.argsort(-lat_ms)[:5]]  # frame index within the timed region, ms
    times = torch.tensor([dev_ms, e2e_s * 1e3, float(np.median(lat)) * 1e3, err], dtype=torch.float64, device="cuda")
    if dist:
        allt = [torch.zeros_like(times) for _ in range(world_size)]
        dist.all_gather(allt, times)
        gp = [torch.zeros(K, 14, dtype=torch.float64, device="cuda") for _ in range(world_size)]
        dist.all_gather(gp, torch.from_numpy(poses).cuda())
        allt = torch.stack(allt).cpu().numpy()
    else:
        allt = times.cpu().numpy()[None]
    if rank != 0:
        if dist: dist.destroy_process_group()
        return
    dev_ms_max, e2e_ms_max = float(allt[:, 0].max()), float(allt[:, 1].max())
    value = world_size * K / (dev_ms_max * 1e-3)
    e2e_value = world_size * K / (e2e_ms_max * 1e-3)
    cpu = cpu_baseline_sample(pkg, scans, cblob, sblob) if world_size == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": "scans/sec", "value": value, "unit": "scans/s", "n_gpus": world_size, "steps": K, "warmup": W,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "map_points": map_points, "points_per_sweep": int(np.mean([len(s) for s in scans])),
                   "l2": "every sweep is a new 1.9 MB input (the %d device-resident sweeps total > 3x L2); the persistent state (~16 MB sub-map + grids) "
                         "stays L2-resident across sweeps as it does in deployment; `cold_l2` repeats the measurement with L2 flushed before every sweep" % len(scans),
                   "parallelism": "independent sequences, one per GPU",
                   "lookahead": "replay mode: the next sweep is registered with vloam_b200_prefetch_scan[_device] before each process_frame call, so its upload "
                                "and scan registration run underneath the current sweep; every sweep's own H2D copy and pose read-back stay inside the timed region"},
        "p50_ms_per_frame_e2e": float(np.max(allt[:, 2])), "p99_ms_per_frame_e2e": float(np.percentile(lat_ms, 99)),
        "slowest_frames_e2e": slow, "final_pose_error_m": float(allt[:, 3].max()),
        "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 14 * 8 + 352 + 720},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extras)
    if roof:
        roof["whole_frame"] = {"algorithmic_bytes": 58e6, "achieved_gbs": 58e6 / (dev_ms_max / K * 1e-3) / 1e9, "frac": 58e6 / (dev_ms_max / K * 1e-3) / 1e9 / roof["peak"]}
        line["roofline"] = roof
    if ktable: line["kernels"] = ktable
    if cpu: line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist: dist.destroy_process_group()


def profile_dominant(ctx, dscans, first, K, map_points):
    """Every launch of a 20-frame replay is bracketed by CUDA events on the stream it is launched on
    (vloam_b200_profile_kernel("*")); the kernel with the largest summed time is the `roofline` kernel,
    the per-kernel table goes into `kernels`.  Event pairs around every launch serialise the streams a
    little, so these runs are separate from the timed legs above."""
    import ctypes
    L = ctx.L
    L.vloam_b200_profile_kernel.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.vloam_b200_profile_table.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    n = min(K, 20)
    L.vloam_b200_profile_kernel(ctx.h, b"*")
    for k in range(first, first + n):
        ctx.process_frame_device(dscans[k].data_ptr(), dscans[k].shape[0], 4)
    ctx.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    if L.vloam_b200_profile_table(ctx.h, buf, len(buf)) <= 0:
        L.vloam_b200_profile_kernel(ctx.h, None)
        return None, None
    L.vloam_b200_profile_kernel(ctx.h, None)
    rows = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms, byt = line.split()
        rows.append((name, int(cnt), float(ms), float(byt)))
    total = sum(r[2] for r in rows)
    rows.sort(key=lambda r: -r[2])
    peak, how = measured_peak()
    table = [{"kernel": nm, "launches_per_frame": cnt / n, "us_per_frame": 1e3 * ms / n, "share": ms / total,
              "achieved_gbs": (byt / (ms * 1e-3) / 1e9) if byt > 0 and ms > 0 else None} for nm, cnt, ms, byt in rows[:12]]
    nm, cnt, ms, byt = rows[0]
    achieved = byt / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(nm)
    roof = {"bound": "hbm", "kernel": nm, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": how, "launches_timed": cnt, "avg_us": 1e3 * ms / cnt, "algorithmic_bytes_per_launch": byt / cnt,
            "share_of_kernel_time": ms / total,
            "note": "the frame is dependency-latency bound (SURVEY 8d): ~60 MB of algorithmic traffic per sweep, i.e. ~9 us at the HBM peak, "
                    "spread over ~60 short dependent kernels; per-kernel times here come from event pairs around every launch (serialised)"}
    return roof, table



def run_c5(args, rank, world_size, local_rank):
    """BASELINE config C5: 8 independent HDL-64E sequences batched across the GPUs of one box (strong scaling): sequence
    s runs on rank s mod G, each rank replays its 8/G sequences concurrently (one context + host thread each), poses and
    timings are gathered with one NCCL all_gather.  `python bench.py --workload c5 [--frames F]` (F = 1000 in BASELINE;
    the default keeps synthetic-sweep generation to a minute)."""
    import torch
    pkg = importlib.import_module("vloam-noted_b200")
    pkg.load_lib()
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total = 8
    mine = len(range(rank, total, world_size))
    _, _, cblob, sblob = make_sequence(pkg, 0, 1)
    if dist: dist.barrier()
    r = batched_sequences(pkg, torch, local_rank, cblob, sblob, nseq=mine, frames=args.frames, warm=min(10, args.frames // 4),
                          first_seq=rank, seq_stride=world_size)
    t = torch.tensor([r["device_ms"], r["p50_ms_per_frame"], r["p99_ms_per_frame"], float(mine * r["frames_per_sequence"])], dtype=torch.float64, device="cuda")
    if dist:
        allt = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu().numpy()
    else:
        allt = t.cpu().numpy()[None]
    if rank == 0:
        print(json.dumps({"metric": "scans/sec", "value": float(allt[:, 3].sum() / (allt[:, 0].max() * 1e-3)), "unit": "scans/s", "n_gpus": world_size,
                          "higher_is_better": True, "scaling": "strong", "dtype": "f32+f64", "data": "synthetic",
                          "config": {"workload": "C5: 8 independent synthetic HDL-64E sequences, %d timed sweeps each, batched across %d GPU(s)" % (r["frames_per_sequence"], world_size),
                                     "sequences_per_gpu": mine},
                          "p50_ms_per_frame": float(allt[:, 1].max()), "p99_ms_per_frame": float(allt[:, 2].max())}), flush=True)
    if dist: dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cold-L2, batched-sequence and C1/C2/C4 legs")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"], help="c3: the headline (default); c5: 8 sequences batched across the GPUs")
    ap.add_argument("--frames", type=int, default=110, help="sweeps per sequence for --workload c5")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "c5":
        run_c5(args, rank, world, local)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
