"""Ad-hoc: phase timing of the last lm_solve_cluster launch (clock64 stamps of CTA 0, not collected by pytest)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 12
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
ctx = pkg.Context()
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
d = [torch.from_numpy(s).cuda() for s in scans]
for k in range(N):
    if k == 4:
        ctx.get_raw("solver.trace")  # arms the trace
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    ctx.synchronize()
    if k >= 5:
        t = np.frombuffer(ctx.get_raw("solver.trace"), np.int64)
        names = ["load(+it0)", "eval", "reduce", "csync", "gather", "logic", "rest", "exit"]
        dt = np.diff(t[:9])
        print("frame %d mapping solve 2: total %.2f us | " % (k, (t[8] - t[0]) / 1965.0) + " ".join("%s %.2f" % (n, v / 1965.0) for n, v in zip(names, dt)))
        lg = [t[5], t[9], t[10], t[11], t[12], t[13]]
        print("    logic0: " + " ".join("%s %.2f" % (n, v / 1965.0) for n, v in zip(["book", "scaleH", "ldl", "mcc", "plus"], np.diff(lg))))
