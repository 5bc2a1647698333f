"""Hottest CUDA source lines of one kernel launch in an .ncu-rep (warp-stall samples per source line).
The CSV source page of ncu carries samples per SASS instruction only; the line of each instruction comes from
`nvdisasm -g` on the cubins inside the shared library (built with -lineinfo).  No GPU needed.
Usage: python profiles/ncu_hot_lines.py report.ncu-rep KERNEL_SUBSTRING LAUNCH_INDEX [TOP_N] [lib.so]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

def line_map(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    m = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin") or "-" in f.split(".sm_")[0]:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        fn, cur = None, None
        for l in txt.splitlines():
            if l.startswith(".text."):
                fn = l[6:].rstrip(":"); cur = None
            elif "//## File" in l:
                mm = re.search(r'File "([^"]+)", line (\d+)', l)
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            elif fn and kernel in fn:
                mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
                if mm: m.setdefault(fn, {})[int(mm.group(1), 16)] = cur
    return m

def main():
    path, kernel, idx = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(root, "vloam-noted_b200", "libvloam_b200.so")
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (kernel, idx)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[h]
    ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    stalls = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    data = []
    for r in rows[h + 1:]:  # the first matching launch only (a regex that matches several kernels yields several sections)
        if len(r) <= max(stalls): continue
        if r[ia] == "Address": break
        data.append(r)
    I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    base = int(data[0][ia], 16)
    maps = line_map(lib, kernel)
    # the template instance / overload whose instruction count matches
    fn = min(maps, key=lambda f: abs(len(maps[f]) - len(data))) if maps else None
    lm = maps.get(fn, {})
    per = collections.defaultdict(lambda: [0, collections.Counter()])
    tot = 0
    for r in data:
        s = I(r[isamp]); tot += s
        key = lm.get(int(r[ia], 16) - base)
        per[key][0] += s
        for i in stalls: per[key][1][hdr[i][6:]] += I(r[i])
    print("%s launch %s: %d samples, %d SASS instructions, line map from %s" % (rows[0][1][:60] if len(rows[0]) > 1 else kernel, idx, tot, len(data), fn))
    src = {}
    for key, (s, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            p = os.path.join(root, "vloam-noted_b200", "csrc", key[0])
            if os.path.exists(p):
                if p not in src: src[p] = open(p).read().splitlines()
                text = src[p][key[1] - 1].strip()[:100] if key[1] - 1 < len(src[p]) else ""
        print("%5.1f%%  %-28s %-100s %s" % (100.0 * s / max(tot, 1), "%s:%d" % key if key else "?", text, " ".join("%s:%d" % kv for kv in st.most_common(3))))

if __name__ == "__main__":
    main()
