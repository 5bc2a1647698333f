"""Ad-hoc: plain replay vs look-ahead replay, per-frame pose differences (not collected by pytest)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import torch
synth = pkg.synth
street = synth.World(1234, 0, 160.0)
KW = dict(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8)
traj = synth.trajectory(8)
scans = [street.scan(1, traj[k], 1000 + k) for k in range(8)]
a = pkg.Context(**KW)
ref = [a.process_frame(s).copy() for s in scans]
a.close()
dev = [torch.from_numpy(s).cuda() for s in scans]
b = pkg.Context(**KW)
pose = np.zeros(14)
for k in range(8):
    if k + 1 < 8: b.prefetch_device(dev[k + 1].data_ptr(), dev[k + 1].shape[0], 4)
    b.process_frame_device(dev[k].data_ptr(), dev[k].shape[0], 4, pose.ctypes.data)
    d = np.abs(pose - ref[k])
    print("frame %d: LO q %.3g t %.3g | LM q %.3g t %.3g" % (k, d[0:4].max(), d[4:7].max(), d[7:11].max(), d[11:14].max()), flush=True)
b.close()
