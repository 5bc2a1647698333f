"""Ad-hoc: a 1000-sweep C3 replay with the look-ahead (BASELINE config C5 uses 1000-frame sequences): per-100-frame
throughput, latency percentiles, pool use, trajectory error; catches drift, capacity and regrow stalls (not collected by pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = int(os.environ.get("FRAMES", "1000"))
world = pkg.synth.World(1234, 1, 190.0)
traj = pkg.synth.trajectory(N, seed=77, step=float(os.environ.get("STEP", "0.25")))  # 250 m in all: stays inside the synthetic world
_, _, cb, sb = bench.make_sequence(pkg, 0, 1)
ctx = pkg.Context()
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
CH = 100
pose = np.zeros(14)
lat_all = []
t_all = 0.0
for c0 in range(0, N, CH):
    scans = [torch.from_numpy(world.scan(1, traj[k], 1000 + k)).cuda() for k in range(c0, min(c0 + CH, N))]
    torch.cuda.synchronize()
    lat = []
    t0 = time.perf_counter()
    for i, s in enumerate(scans):
        t1 = time.perf_counter()
        if i + 1 < len(scans): ctx.prefetch_device(scans[i + 1].data_ptr(), scans[i + 1].shape[0], 4)
        ctx.process_frame_device(s.data_ptr(), s.shape[0], 4, pose.ctypes.data)
        lat.append((time.perf_counter() - t1) * 1e3)
    dt = time.perf_counter() - t0
    t_all += dt
    lat = np.array(lat); lat_all.append(lat)
    k = min(c0 + CH, N) - 1
    err = np.linalg.norm(pose[11:14] - (np.array(traj[k][:3]) - np.array(traj[0][:3])))
    print("sweeps %4d-%4d: %.0f scans/s, p50 %.3f p99 %.3f max %.3f ms (sweep %d) | mapped t = %s, |t - truth| = %.3f m" % (
        c0, k, len(scans) / dt, np.median(lat), np.percentile(lat, 99), lat.max(), c0 + int(lat.argmax()), np.round(pose[11:14], 2), err), flush=True)
    if c0 == 0: print("   first sweeps, ms:", " ".join("%.2f" % v for v in lat[:8]))
lat = np.concatenate(lat_all)
print("total: %d sweeps, %.0f scans/s, p50 %.3f p99 %.3f p99.9 %.3f max %.3f ms" % (N, N / t_all, np.median(lat), np.percentile(lat, 99), np.percentile(lat, 99.9), lat.max()))
print("map bytes corner/surf:", len(ctx.get("lm.cornerMap")), len(ctx.get("lm.surfMap")))
ctx.close()
