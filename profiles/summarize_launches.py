"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time per sweep and share.
With two sweeps registered ahead the launches of three sweeps interleave (scan registration of k+2, odometry of k+1, mapping of
k), so the list is not cut into frames: totals over the capture are divided by the number of mapping stages in it
(one lm_prepare_fast / lm_prepare per sweep)."""
import collections, csv, io, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    rows.append((row["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0], v))
frames = max(1, sum(1 for n, _ in rows if n in ("lm_prepare_fast", "lm_prepare")))
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, v in rows:
    tot[n] += v; cnt[n] += 1
s = sum(tot.values())
print("%s: %d launches over %d sweeps = %.1f launches and %.1f us of kernel time per sweep (ncu: cold cache, serialised)" % (path, len(rows), frames, len(rows) / frames, s / frames))
print("%-28s %10s %12s %7s" % ("kernel", "per sweep", "us / sweep", "share"))
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-28s %10.2f %12.1f %6.1f%%" % (n, cnt[n] / frames, v / frames, 100 * v / s))
