"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the last frame."""
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
    rows.append((row["Kernel Name"].split("(")[0].replace("void ", ""), v))
starts = [i for i, (n, _) in enumerate(rows) if n.startswith("sr_find_bounds")]
last = rows[starts[-2]:starts[-1]] if len(starts) >= 2 else rows
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, v in last:
    tot[n] += v; cnt[n] += 1
s = sum(tot.values())
print("last complete frame of %s: %d launches, %.1f us of kernel time (ncu: cold cache, serialised)" % (path, len(last), s))
print("%-28s %6s %10s %7s" % ("kernel", "count", "us", "share"))
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-28s %6d %10.1f %6.1f%%" % (n, cnt[n], v, 100 * v / s))
