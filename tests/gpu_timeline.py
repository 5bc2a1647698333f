"""Ad-hoc: per-launch timeline of one C3 frame (events around every launch; not collected by pytest)."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("vloam-noted_b200")
import bench
import torch
N = 14
scans, traj, cb, sb = bench.make_sequence(pkg, 0, N)
ctx = pkg.Context()
ctx.set("lm.cornerMap", cb); ctx.set("lm.surfMap", sb)
d = [torch.from_numpy(s).cuda() for s in scans]
L = ctx.L
L.vloam_b200_profile_kernel.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
L.vloam_b200_profile_timeline.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
LOOKAHEAD = os.environ.get("LOOKAHEAD") in ("1", "2")  # 2: two sweeps registered ahead (what bench.py does)
TWO = os.environ.get("LOOKAHEAD") == "2"
for k in range(N - 2):
    if LOOKAHEAD: ctx.prefetch_device(d[k + 1].data_ptr(), d[k + 1].shape[0], 4)
    if TWO and k + 2 < N: ctx.prefetch_device(d[k + 2].data_ptr(), d[k + 2].shape[0], 4)
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
ctx.synchronize()
L.vloam_b200_profile_kernel(ctx.h, b"*")
for k in range(N - 2, N):
    if LOOKAHEAD and k + 1 < N: ctx.prefetch_device(d[k + 1].data_ptr(), d[k + 1].shape[0], 4)
    if TWO and k + 2 < N: ctx.prefetch_device(d[k + 2].data_ptr(), d[k + 2].shape[0], 4)
    ctx.process_frame_device(d[k].data_ptr(), d[k].shape[0], 4)
ctx.synchronize()
buf = ctypes.create_string_buffer(1 << 18)
n = L.vloam_b200_profile_timeline(ctx.h, buf, len(buf))
rows = [l.split() for l in buf.value.decode().splitlines()]
print("%d launches in 2 frames" % n)
prev_end = {}
for nm, s, t0, t1 in rows:
    t0, t1 = float(t0), float(t1)
    gap = t0 - prev_end.get(s, t0)
    print("s%s %9.1f %9.1f  dur %7.1f  gap %6.1f  %s%s" % (s, t0, t1, t1 - t0, gap, "    " * int(s), nm))
    prev_end[s] = t1
