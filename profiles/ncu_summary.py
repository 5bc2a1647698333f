"""Summarise an .ncu-rep captured with `ncu --set full`: per launch duration, DRAM bytes (read + write), achieved
DRAM GB/s against the measured HBM peak, registers, grid, top warp-stall reasons.  Reads the report with `ncu -i`
(no GPU needed).  Usage: python profiles/ncu_summary.py report.ncu-rep [--json traffic.json]"""
import csv, io, json, os, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}

def main():
    path = sys.argv[1]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    def g(r, name):
        try: return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)
        except Exception: return 0.0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 6546.6
    try: peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception: pass
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    print("# %s   (HBM peak %.1f GB/s measured; durations are ncu's: cold cache, serialised, clock-control none)" % (os.path.basename(path), peak))
    print("%-24s %6s %8s %11s %11s %8s %7s %5s  top stall reasons (%% of samples)" % ("kernel", "grid", "us", "dram_rd_B", "dram_wr_B", "GB/s", "of_peak", "regs"))
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        us = g(r, "gpu__time_duration.sum")
        rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
        tot = sum(g(r, h) for h in stall) or 1.0
        top = sorted(((g(r, h) / tot * 100, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall), reverse=True)[:4]
        print("%-24s %6s %8.1f %11.0f %11.0f %8.1f %6.2f%% %5.0f  %s" % (name[:24], r[col["launch__grid_size"]], us, rd, wr, gbs, 100 * gbs / peak,
              g(r, "launch__registers_per_thread"), ", ".join("%s %.0f" % (n, p) for p, n in top)))
        traffic.setdefault(name, []).append(rd + wr)
    if "--json" in sys.argv:
        dst = sys.argv[sys.argv.index("--json") + 1]
        js = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (largest of the captured launches of that kernel) from `ncu --set full --clock-control none`; "
                          "bench.py copies the dominant kernel's entry into roofline.traffic"}
        for k, v in traffic.items(): js[k] = int(max(v))  # the largest launch: skipped (early-exit) launches of the same kernel carry no traffic
        json.dump(js, open(dst, "w"), indent=1)

if __name__ == "__main__":
    main()
