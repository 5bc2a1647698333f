"""Ad-hoc: bench.py's device leg + per-kernel table, printed compactly (not collected by pytest)."""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", os.environ.get("STEPS", "300"), "--warmup", os.environ.get("WARMUP", "20"), "--no-cpu-baseline"] + ([] if os.environ.get("EXTRAS") else ["--no-extras"]), capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print("value %.1f scans/s  %.4f ms/step | e2e %.1f | p50 %.3f ms | launches/frame %.1f | clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["p50_ms_per_frame_e2e"], d["gpu_launches"] / d["steps"], d["clocks"]))
print("p99 %.3f ms, slowest e2e frames (index, ms): %s" % (d["p99_ms_per_frame_e2e"], d["slowest_frames_e2e"]))
for key in ("cold_l2", "batched", "workloads"):
    if key in d: print(key, json.dumps(d[key]))
for k in d["kernels"]:
    print("  %-22s x%.0f %7.1f us/frame  %5.1f%%  %s GB/s" % (k["kernel"], k["launches_per_frame"], k["us_per_frame"], 100 * k["share"],
          "%.0f" % k["achieved_gbs"] if k["achieved_gbs"] else "-"))
open(os.path.join(ROOT, "gpurun_out", "bench_latest.json"), "w").write(line + "\n")
