"""Ad-hoc: per-query-warp cycle counts of the odometry grid association (not collected by pytest)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 30
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
ctx = pkg.Context()
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
d = [torch.from_numpy(s).cuda() for s in scans]
for k in range(N):
    if k == 20:
        ctx.get_raw("lo.trace")
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    ctx.synchronize()
t = np.frombuffer(ctx.get_raw("lo.trace"), np.int32).reshape(-1, 2)
t = t[t[:, 0] > 0]
us = t[:, 0] / 1965.0
for name, m in (("corner", (t[:, 1] & 4) == 0), ("surf", (t[:, 1] & 4) != 0)):
    u, f = us[m], t[m, 1] & 27
    print("%s: n=%d  median %.1f  p90 %.1f  p99 %.1f  max %.1f us | NN pass widened to 5^3: %d, to 9^3: %d; 2nd pass to 5^3: %d, to 9^3: %d" % (
        name, len(u), np.median(u), np.percentile(u, 90), np.percentile(u, 99), u.max(), ((f & 1) != 0).sum(), ((f & 8) != 0).sum(),
        ((f & 2) != 0).sum(), ((f & 16) != 0).sum()))
    for fl in sorted(set(f.tolist())):
        print("   flags %2d: n=%d median %.1f max %.1f" % (fl, (f == fl).sum(), np.median(u[f == fl]), u[f == fl].max()))
