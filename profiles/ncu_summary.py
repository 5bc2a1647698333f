"""Summarise an .ncu-rep (ncu --set full): per launch duration, DRAM bytes, achieved DRAM GB/s, registers, top stall reasons.
Usage: python profiles/ncu_summary.py report.ncu-rep [--src KERNEL_REGEX:ID]  (reads the report with `ncu -i`, no GPU needed)."""
import csv, io, subprocess, sys

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, rows[2:]

def main():
    path = sys.argv[1]
    hdr, rows = raw(path)
    col = {h: i for i, h in enumerate(hdr)}
    def g(r, name, default=0.0):
        try: return float(r[col[name]].replace(",", ""))
        except Exception: return default
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    print("%-26s %8s %10s %10s %9s %5s %6s  top stall reasons (%% of samples)" % ("kernel", "us", "dram_rd_B", "dram_wr_B", "GB/s", "regs", "grid"))
    for r in rows:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        us = g(r, "gpu__time_duration.sum")  # usecond in --csv raw page
        rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
        unit_rd = hdr_units.get("dram__bytes_read.sum", "byte") if False else None
        tot = sum(g(r, h) for h in stall) or 1.0
        top = sorted(((g(r, h) / tot * 100, h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall), reverse=True)[:4]
        print("%-26s %8.1f %10.0f %10.0f %9.1f %5.0f %6s  %s" % (name[:26], us, rd, wr, 0.0, g(r, "launch__registers_per_thread"),
              r[col["launch__grid_size"]] if "launch__grid_size" in col else "-", ", ".join("%s %.0f" % (n, p) for p, n in top)))

hdr_units = {}
if __name__ == "__main__":
    main()
