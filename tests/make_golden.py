"""Regenerates tests/golden/vlp16_pair.npz: a small fixed input (two thinned VLP-16 sweeps) and the
oracle's outputs for it.  The reference ships no vectors, so these pin the ORACLE
(and through it the CUDA path) against regressions; they are not outputs of the reference itself.
Run:  python tests/make_golden.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("vloam-noted_b200")
import oracle_py as op

KW = dict(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4)


def make():
    synth = pkg.synth
    w = synth.World(1234, 0, 160.0)
    traj = synth.trajectory(2)
    scans = []
    for k in range(2):
        s = w.scan(synth.VLP16, traj[k], 1000 + k)
        # keep every 3rd azimuth column (16 beams per column, azimuth-major) to keep the fixture small
        col = np.arange(len(s)) // 16
        scans.append(s[col % 3 == 0][:, :3].copy())
    o = op.Oracle(**KW)
    out = {"scan0": scans[0], "scan1": scans[1]}
    for k in range(2):
        o.process(scans[k])
        for nm in ("sr.sharp", "sr.flat", "sr.lessSharp", "sr.lessFlat"):
            out["f%d.%s" % (k, nm)] = o.get(nm)
        out["f%d.label" % k] = o.get("sr.label").astype(np.int8)
        out["f%d.lo.pose" % k] = o.get("lo.pose")
        out["f%d.lm.pose" % k] = o.get("lm.pose")
    out["f1.assoc.corner0"] = o.get("lo.assoc.corner0")
    out["f1.assoc.surf0"] = o.get("lo.assoc.surf0")
    out["f1.knn.cidx0"] = o.get("lm.knn.cidx0")
    out["f1.knn.cd2_0"] = o.get("lm.knn.cd20")
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vlp16_pair.npz"), **make())
    print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "vlp16_pair.npz")), "bytes")
