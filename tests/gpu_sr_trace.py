"""Ad-hoc: phase timing of sr_pick / sr_ring_voxel per ring CTA (clock64 stamps; not collected by pytest)."""
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
        ctx.get_raw("sr.trace")
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    ctx.synchronize()
t = np.frombuffer(ctx.get_raw("sr.trace"), np.int64).reshape(256, 8)
for name, base, labels in (("sr_pick", 0, ["init(gap,keys)", "sort", "walk", "barrier", "re-walk"]), ("sr_ring_voxel", 128, ["select+bbox", "keys", "sort", "centroids"])):
    rows = t[base:base + 64]
    rows = rows[rows[:, 1] > 0]
    dt = np.diff(rows[:, :len(labels) + 1], axis=1) / 1965.0
    tot = (rows[:, len(labels)] - rows[:, 0]) / 1965.0
    print("%s: %d CTAs, per-CTA total median %.1f max %.1f us" % (name, len(rows), np.median(tot), tot.max()))
    for i, l in enumerate(labels):
        print("   %-16s median %6.2f  max %6.2f us" % (l, np.median(dt[:, i]), dt[:, i].max()))
    if base == 0:
        a, b, c2, d2 = rows[:, 2], rows[:, 6], rows[:, 7], rows[:, 3]
        ok = (b > a) & (c2 > b)
        print("   walk of sector 0: load+lane sort %.2f | descending walk %.2f | ascending walk %.2f us (medians; stamps of the first walk only when no re-walk overwrote them)" % (
            np.median((b - a)[ok]) / 1965.0, np.median((c2 - b)[ok]) / 1965.0, np.median((d2 - c2)[ok]) / 1965.0))
