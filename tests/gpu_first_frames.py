"""Ad-hoc: where the first sweeps of a context spend their time (not collected by pytest).
VLOAM_TRACE_ALLOC=1 prints every device buffer (re)allocation with its cudaMalloc time."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 8
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
d = [torch.from_numpy(s).cuda() for s in scans]
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    ctx = pkg.Context(**bench.KW)
    t1 = time.perf_counter()
    ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
    ctx.synchronize()
    t2 = time.perf_counter()
    print("context %d: create %.1f ms, map import %.1f ms, allocations so far %d" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, int(ctx.get("alloc.count")[0])), flush=True)
    pose = np.zeros(14)
    for k in range(N - 1):
        t3 = time.perf_counter()
        ctx.prefetch_device(d[k + 1].data_ptr(), d[k + 1].shape[0], 4)
        ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4, pose.ctypes.data)
        t4 = time.perf_counter()
        ctx.synchronize()
        t5 = time.perf_counter()
        print("  sweep %d: process_frame %.2f ms, + synchronize %.2f ms, allocations %d" % (k, (t4 - t3) * 1e3, (t5 - t4) * 1e3, int(ctx.get("alloc.count")[0])), flush=True)
    ctx.close()
