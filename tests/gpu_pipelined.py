"""Ad-hoc: device-resident replay with and without a per-frame sync (not collected by pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 120
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
d = [torch.from_numpy(s).cuda() for s in scans]
for mode in ("sync", "nosync", "sync", "nosync"):
    ctx = pkg.Context()
    ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
    for k in range(10):
        ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    ctx.synchronize()
    t0 = time.perf_counter()
    per = []
    for k in range(10, N):
        t1 = time.perf_counter()
        ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
        if mode == "sync":
            ctx.synchronize()
        per.append((time.perf_counter() - t1) * 1e3)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    per = np.array(per)
    print(mode, "total ms/frame %.3f" % (dt * 1e3 / (N - 10)), "median %.3f max %.3f" % (np.median(per), per.max()), "n>2ms", int((per > 2).sum()), flush=True)
    ctx.close()
