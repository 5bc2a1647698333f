"""Ad-hoc: a short replay for compute-sanitizer (memcheck / racecheck), not collected by pytest.
    compute-sanitizer --tool memcheck python tests/gpu_sanitize.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
w = pkg.synth.World(1234, 0, 160.0)
traj = pkg.synth.trajectory(5)
for sensor, kw in ((0, dict(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4)), (1, dict(n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8))):
    g = pkg.Context(**kw)
    scans = [w.scan(sensor, traj[k], 1000 + k) for k in range(5)]
    for k in range(5):
        if k + 1 < 5:
            a = np.ascontiguousarray(scans[k + 1], np.float32)
        pose = g.process_frame(scans[k])
    reg = g.register_full_cloud()
    print("sensor", sensor, "ok, mapped t", np.round(pose[11:14], 3), "registered", len(reg), flush=True)
    g.close()
print("done")
