"""Ad-hoc: per-stage device times on the C3 workload (not collected by pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = int(os.environ.get('FRAMES', '40'))
WL = os.environ.get('WORKLOAD', 'c3')
if WL == 'c4':  # OS1-128 sweeps against a ~1.9M-point map
    world = pkg.synth.World(1234, 2, 190.0)
    traj = pkg.synth.trajectory(N)
    scans = [world.scan(2, traj[k], 1000 + k) for k in range(N)]
    cb, sb = pkg.synth.cubes_blob(world.plant(0, 0.4, seed=99), 0.4), pkg.synth.cubes_blob(world.plant(1, 0.8, seed=98), 0.8)
    ctx = pkg.Context(n_scans=128, minimum_range=0.3)
elif WL == 'c1':  # VLP-16 sweeps, no planted map
    world = pkg.synth.World(1234, 0, 160.0)
    traj = pkg.synth.trajectory(N)
    scans = [world.scan(0, traj[k], 1000 + k) for k in range(N)]
    cb = sb = None
    ctx = pkg.Context(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4)
else:
    scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
    ctx = pkg.Context()
if cb is not None:
    ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
ctx.set_timing(True)
d = [torch.from_numpy(s).cuda() for s in scans]
torch.cuda.synchronize()
rows = []
detail = []
walls = []
LOOKAHEAD = os.environ.get("LOOKAHEAD") in ("1", "2")  # 2: two sweeps registered ahead (what bench.py does)
TWO = os.environ.get("LOOKAHEAD") == "2"   # replay mode: register the next sweep, no full synchronize between sweeps
host = []
for k in range(N):
    t0 = time.perf_counter()
    if LOOKAHEAD and k + 1 < N: ctx.prefetch_device(d[k + 1].data_ptr(), d[k + 1].shape[0], 4)
    if TWO and k + 2 < N: ctx.prefetch_device(d[k + 2].data_ptr(), d[k + 2].shape[0], 4)
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
    if not LOOKAHEAD: ctx.synchronize()
    host.append(np.frombuffer(ctx.get_raw("timing.host"), np.float64).copy()) if LOOKAHEAD else None
    walls.append((time.perf_counter() - t0) * 1e3)
    rows.append(ctx.stage_ms())
    detail.append(np.frombuffer(ctx.get_raw("timing.detail"), np.float32).copy())
rows = np.array(rows)[8:]
print("stage ms median SR/LO/LM:", np.median(rows, 0).round(3), "sum", np.median(rows.sum(1)).round(3), "wall median", np.median(walls[8:]).round(3))
print("stage ms mean   SR/LO/LM:", rows.mean(0).round(3))
dm = np.median(np.array(detail)[8:], 0)
print("ms since frame start (median): SR end %.3f | LO end %.3f | sub-map build end %.3f | stacks awaited %.3f | solve 1 end %.3f | LM end %.3f" % (dm[0], dm[1], dm[3], dm[4], dm[5], dm[2]))
print("   side streams: surf stack ready %.3f | corner stack ready %.3f | next LO grid ready %.3f" % (dm[6], dm[7], dm[8]))
if len(dm) >= 14:
    print("   next sweep: sharp/flat features ready %.3f | scan registration complete %.3f | look-ahead odometry starts %.3f" % (dm[11], dm[13], dm[12]))
if len(dm) >= 11:
    print("   look-ahead odometry of the next sweep done %.3f | in-place map update done %.3f  (ms since this sweep's start; events of the PREVIOUS sweep when negative or > 1)" % (dm[9], dm[10]))
if LOOKAHEAD:
    hm = np.median(np.array(host)[8:], 0)
    print("host clock inside process_frame, us (median): SR adopted %.0f | odometry + look-ahead queued %.0f | S1 + side streams queued %.0f | helper joined %.0f | mapping queued %.0f | S2 passed %.0f | update submitted %.0f" % tuple(hm[:7]))
print("launches/frame", ctx.kernel_launches / N)
allr = np.array(walls)
print("wall ms per frame:", " ".join("%.1f" % v for v in allr))
