"""Ad-hoc: per-stage device times on the C3 workload (not collected by pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = int(os.environ.get('FRAMES', '40'))
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
ctx = pkg.Context()
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
ctx.set_timing(True)
d = [torch.from_numpy(s).cuda() for s in scans]
torch.cuda.synchronize()
rows = []
walls = []
for k in range(N):
    t0 = time.perf_counter()
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    ctx.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3)
    rows.append(ctx.stage_ms())
rows = np.array(rows)[8:]
print("stage ms median SR/LO/LM:", np.median(rows, 0).round(3), "sum", np.median(rows.sum(1)).round(3), "wall median", np.median(walls[8:]).round(3))
print("stage ms mean   SR/LO/LM:", rows.mean(0).round(3))
print("launches/frame", ctx.kernel_launches / N)
allr = np.array(walls)
print("wall ms per frame:", " ".join("%.1f" % v for v in allr))
