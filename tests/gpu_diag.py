"""Ad-hoc GPU diagnostic (not collected by pytest): free-running + stage parity printout."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("vloam-noted_b200")
import oracle_py as op
synth = pkg.synth


def cmp(name, a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    if a.shape != b.shape:
        print("  %-22s SHAPE %s vs %s" % (name, a.shape, b.shape)); return False
    if a.dtype == np.float32:
        a = a.view(np.uint32); b = b.view(np.uint32)
    bad = int((a != b).sum())
    print("  %-22s %s n=%d mismatches=%d" % (name, "OK " if bad == 0 else "BAD", a.size, bad))
    if bad:
        idx = np.argwhere(a != b)[:3]
        print("     first:", idx.tolist())
    return bad == 0


def main():
    sensor = int(os.environ.get("SENSOR", "1"))
    nframes = int(os.environ.get("FRAMES", "4"))
    w = synth.World(1234, 0, 160.0)
    traj = synth.trajectory(nframes)
    rng_kw = dict(n_scans=synth.N_SCANS[sensor], minimum_range=5.0 if sensor == 1 else 0.3,
                  line_res=0.4 if sensor == 1 else 0.2, plane_res=0.8 if sensor == 1 else 0.4)
    o = op.Oracle(**rng_kw)
    g = pkg.Context(**rng_kw)
    g.set_capture(True)
    for k in range(nframes):
        scan = w.scan(sensor, traj[k], 1000 + k)
        print("frame", k, "n", len(scan))
        # teacher-force the GPU with the oracle's state from before this frame
        if k > 0:
            g.set_last(o.get("lo.cornerLast"), o.get("lo.surfLast"))
            g.set("lo.pose", o.get("lo.pose"))
            g.set("lm.state", o.get("lm.state")[:4])
            g.set("lm.pose", o.get("lm.pose"))
            g.set("lm.cornerMap", o.get("lm.cornerMap"))
            g.set("lm.surfMap", o.get("lm.surfMap"))
        o.scan_registration(scan)
        g.begin_frame(); g.scan_registration(scan)
        for nm in ("sr.laserCloud", "sr.curvature", "sr.label", "sr.sharp", "sr.lessSharp", "sr.flat", "sr.lessFlat"):
            cmp(nm, o.get(nm), g.get(nm))
        o.laser_odometry(); r = g.laser_odometry()
        if k > 0:
            for nm in ("lo.assoc.corner0", "lo.assoc.surf0", "lo.assoc.corner1", "lo.assoc.surf1"):
                cmp(nm, o.get(nm), g.get(nm))
            print("  lo.costs oracle", o.get("lo.costs"), "gpu", g.get("lo.costs"))
        po, pg = o.get("lo.pose"), g.get("lo.pose")
        print("  lo.pose max|diff| %.3e" % np.abs(po - pg).max())
        o.laser_mapping(); g.laser_mapping()
        for nm in ("lm.cornerStack", "lm.surfStack", "lm.cornerFromMap", "lm.surfFromMap"):
            cmp(nm, o.get(nm), g.get(nm))
        st_o, st_g = o.get("lm.state"), g.get("lm.state")
        print("  lm.state oracle", st_o, "gpu", st_g)
        if st_o[4]:
            for p in (0, 1):
                for kind in ("c", "s"):
                    oi, gi = o.get("lm.knn.%sidx%d" % (kind, p)), g.get("lm.knn.%sidx%d" % (kind, p))
                    od, gd = o.get("lm.knn.%sd2%d" % (kind, p)), g.get("lm.knn.%sd2%d" % (kind, p))
                    ook, gok = o.get("lm.knn.%sok%d" % (kind, p)), g.get("lm.knn.%sok%d" % (kind, p))
                    if oi.shape != gi.shape:
                        print("  knn shape mismatch", oi.shape, gi.shape); continue
                    acc = od[:, 4] < 1.0
                    print("  knn pass%d %s: queries %d accepted-ball %d idx-mismatch %d d2-mismatch %d ok-mismatch %d (ok %d)" % (
                        p, kind, len(oi), acc.sum(), int((oi[acc] != gi[acc]).any(axis=1).sum()),
                        int((od[acc].view(np.uint32) != gd[acc].view(np.uint32)).any(axis=1).sum()), int((ook != gok).sum()), int(ook.sum())))
            print("  lm.costs oracle", o.get("lm.costs"), "gpu", g.get("lm.costs"))
        pmo, pmg = o.get("lm.pose"), g.get("lm.pose")
        print("  lm.pose max|diff| %.3e   t_oracle %s" % (np.abs(pmo - pmg).max(), np.round(pmo[4:7], 4)))
        mo, mg = o.get("lm.cornerMap"), g.get("lm.cornerMap")
        print("  cornerMap equal:", mo == mg, len(mo), len(mg), " surfMap equal:", o.get("lm.surfMap") == g.get("lm.surfMap"))
    print("kernel launches:", g.kernel_launches)


if __name__ == "__main__":
    main()
