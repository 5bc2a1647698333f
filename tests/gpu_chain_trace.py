"""Ad-hoc: device-side timeline of the pose chain in the benchmarked mode (two sweeps registered ahead), from %globaltimer
stamps taken by the kernels themselves -- no events, no profiler (not collected by pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 40
NAMES = {1: "lm_prepare_fast", 2: "lg_knn", 3: "lm_fit", 4: "lm_solve_cluster(16)", 9: "  solve exit", 5: "lm_transform_update", 6: "mu_keys", 7: "mu_apply",
         11: "lo_assoc_grid_both", 14: "lm_solve_cluster(8)", 24: "  solve(8) exit"}
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
ctx = pkg.Context(**bench.KW)
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
d = [torch.from_numpy(s).cuda() for s in scans]
pose = np.zeros(14)
host = []
marks = []
for k in range(N - 2):
    if k == 20:
        ctx.synchronize(); ctx.get_raw("chain.trace")
    bench.prefetch_ahead(ctx, d, k, True)
    t0 = time.perf_counter()
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4, pose.ctypes.data)
    host.append((t0, time.perf_counter()))
    marks.append(np.frombuffer(ctx.get_raw("timing.host"), np.float64).copy())
raw = np.frombuffer(ctx.get_raw("chain.trace"), np.uint64)
n = int(raw[0]); rec = raw[1:1 + 2 * min(n, 4096)].reshape(-1, 2)
order = np.argsort(rec[:, 1], kind="stable"); rec = rec[order]
prep = [i for i in range(len(rec)) if rec[i, 0] == 1]
print("%d stamps, %d mapping stages" % (n, len(prep)))
per = np.diff([rec[i, 1] for i in prep]) / 1e3
print("period between lm_prepare_fast starts, us: median %.1f  min %.1f  max %.1f" % (np.median(per), per.min(), per.max()))
i0, i1 = prep[len(prep) // 2], prep[len(prep) // 2 + 1]
t0 = rec[i0, 1]
for i in range(i0, i1 + 1):
    print("  %8.1f us  %s" % ((rec[i, 1] - t0) / 1e3, NAMES.get(int(rec[i, 0]), str(rec[i, 0]))))
hp = np.array([b - a for a, b in host[22:]]) * 1e6
gap = np.array([host[i + 1][0] - host[i][1] for i in range(22, len(host) - 1)]) * 1e6
print("host: process_frame call %.1f us median, python between calls %.1f us median" % (np.median(hp), np.median(gap)))
m = np.median(np.array(marks[22:]), 0)
print("host clock inside process_frame, us since entry (median): SR adopted %.0f | odometry adopted %.0f | lm_run %.0f | helper joined %.0f | S2 recorded %.0f | S2 passed %.0f | bookkeeping done %.0f"
      % (m[0], m[1], m[2], m[3], m[4], m[5], m[6]))
print("   in-place path: prepare queued %.0f | first kNN + fit queued %.0f | side work submitted %.0f | passes queued %.0f | update queued %.0f | lm_sync_s2 returned %.0f"
      % (m[7], m[8], m[9], m[10], m[11], m[12]))
