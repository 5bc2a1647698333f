"""SASS evidence for the hot kernels (no GPU needed): cuobjdump -sass of libvloam_b200.so.
Writes profiles/<round>_sass_hot_kernels.txt (full listings of the kernels named below) and
profiles/<round>_sass_mnemonics.txt; usage: python profiles/make_sass_listing.py [r2] (per-kernel instruction count and mnemonic histogram of EVERY kernel in the library)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vloam-noted_b200", "libvloam_b200.so")
ROUND = sys.argv[1] if len(sys.argv) > 1 else "r2"
HOT = ["lm_solve_cluster", "lo_assoc_grid_both", "sr_pick", "sr_ring_voxel", "lg_knn", "lm_fit", "mu_keys", "mu_apply", "lo_grid_alloc", "bt_cluster_sort", "vg_centroid"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception: return n
with open(os.path.join(ROOT, "profiles", ROUND + "_sass_mnemonics.txt"), "w") as f:
    f.write("# cuobjdump -sass vloam-noted_b200/libvloam_b200.so (sm_100a): instructions and top mnemonics per kernel\n")
    for name, lines in funcs.items():
        ops = collections.Counter()
        for l in lines:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m: ops[m.group(1).split(".")[0]] += 1
        f.write("%s\n    %d instructions: %s\n" % (demangle(name).split("(")[0], sum(ops.values()), ", ".join("%s %d" % kv for kv in ops.most_common(14))))
with open(os.path.join(ROOT, "profiles", ROUND + "_sass_hot_kernels.txt"), "w") as f:
    f.write("# cuobjdump -sass, hot kernels only (encodings stripped)\n")
    done = set()
    for name, lines in funcs.items():
        short = demangle(name).split("(")[0].replace("void ", "")
        base = short.split("<")[0]
        if base not in HOT or (base in done and base != "lm_solve_cluster"): continue
        if base == "lm_solve_cluster" and "2, 256, 16>" not in short.replace("(int)", ""): continue
        done.add(base)
        f.write("\n======== %s ========\n" % short)
        for l in lines:
            if re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", l): continue
            f.write(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() + "\n")
print("kernels in library:", len(funcs))
