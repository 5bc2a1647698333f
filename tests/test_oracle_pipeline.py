"""Oracle pipeline self-checks: structural invariants of scanRegistration, odometry / mapping
recovering known motion, golden fixtures (tests/golden, made by tests/make_golden.py)."""
import os

import numpy as np
import pytest

KW16 = dict(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4)


@pytest.fixture(scope="module")
def golden():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vlp16_pair.npz")
    return dict(np.load(p))


def test_oracle_matches_golden(op, golden):
    o = op.Oracle(**KW16)
    for k in range(2):
        o.process(golden["scan%d" % k])
        for nm in ("sr.sharp", "sr.flat", "sr.lessSharp", "sr.lessFlat"):
            a, b = o.get(nm), golden["f%d.%s" % (k, nm)]
            assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all(), nm
        assert (o.get("sr.label").astype(np.int8) == golden["f%d.label" % k]).all()
        assert np.abs(o.get("lo.pose") - golden["f%d.lo.pose" % k]).max() < 1e-12
        assert np.abs(o.get("lm.pose") - golden["f%d.lm.pose" % k]).max() < 1e-12
    assert (o.get("lo.assoc.corner0") == golden["f1.assoc.corner0"]).all()
    assert (o.get("lo.assoc.surf0") == golden["f1.assoc.surf0"]).all()
    assert (o.get("lm.knn.cidx0") == golden["f1.knn.cidx0"]).all()


@pytest.mark.parametrize("sensor,kw", [(0, KW16), (1, dict(n_scans=64, minimum_range=5.0)), (2, dict(n_scans=128, minimum_range=0.3)),
                                       (3, dict(n_scans=32, minimum_range=0.3))])
def test_scan_registration_invariants(op, synth, street, sensor, kw):
    scan = street.scan(sensor, [0, 0, 0, 0, 0, 0], 1000)
    o = op.Oracle(**kw)
    o.scan_registration(scan)
    R = kw["n_scans"]
    cloud, label, curv = o.get("sr.laserCloud"), o.get("sr.label"), o.get("sr.curvature")
    start, end = o.get("sr.scanStartInd"), o.get("sr.scanEndInd")
    finite = np.isfinite(scan[:, :3]).all(1)
    assert 0 < len(cloud) <= finite.sum()
    ring = np.floor(cloud[:, 3] + 1e-4).astype(int)
    assert (np.diff(ring) >= 0).all() and ring.min() >= 0 and ring.max() < R            # ring-major (SR.cpp:308-315)
    assert (cloud[:, 3] - ring < 0.11).all()                                           # intensity = ring + 0.1 * relTime
    sharp, less, flat, lessflat = (o.get(n) for n in ("sr.sharp", "sr.lessSharp", "sr.flat", "sr.lessFlat"))
    assert len(sharp) <= 2 * 6 * R and len(less) <= 20 * 6 * R and len(flat) <= 4 * 6 * R  # SR.cpp:386-400, 452
    assert (label == 2).sum() == len(sharp) and ((label == 2) | (label == 1)).sum() == len(less) and (label == -1).sum() == len(flat)
    assert (curv[label >= 1] > 0.1).all() and (curv[label == -1] < 0.1).all()
    for r in range(R):  # nothing is picked outside [scanStartInd, scanEndInd)
        lo, hi = start[r] - 5, end[r] + 6
        if hi > lo:
            assert (label[lo:start[r]] == 0).all() and (label[max(end[r], lo):hi] == 0).all()
    assert len(lessflat) < (label <= 0).sum()                                          # the 0.2 m voxel filter thins it


def test_scan_registration_edge_cases(op):
    o = op.Oracle(**KW16)
    o.scan_registration(np.zeros((0, 3), np.float32))
    assert len(o.get("sr.laserCloud")) == 0 and len(o.get("sr.sharp")) == 0
    o.scan_registration(np.full((100, 3), np.nan, np.float32))
    assert len(o.get("sr.laserCloud")) == 0
    near = np.random.RandomState(0).randn(100, 3).astype(np.float32) * 0.05            # all inside minimum_range
    o.scan_registration(near)
    assert len(o.get("sr.laserCloud")) == 0
    few = np.array([[5, 0, 0.1], [5, 0.1, 0.1], [5, 0.2, 0.1]], np.float32)              # rings with < 6 usable points are skipped
    o.scan_registration(few)
    assert len(o.get("sr.laserCloud")) == 3 and len(o.get("sr.lessFlat")) == 0


def test_odometry_and_mapping_track_known_motion(op, synth, street):
    traj = synth.trajectory(8)
    o = op.Oracle(n_scans=64, minimum_range=5.0, knn_backend=1)
    for k in range(8):
        o.process(street.scan(synth.HDL64, traj[k], 1000 + k))
    lo, lm = o.get("lo.pose"), o.get("lm.pose")
    assert np.linalg.norm(lm[4:7] - traj[7][:3]) < 0.05          # mapping within 5 cm after 7 m
    assert abs(np.linalg.norm(lo[11:14]) - 1.0) < 0.05           # frame-to-frame step of 1 m
    q_true, _ = synth.pose_to_qt(traj[7])
    assert min(np.abs(lm[:4] - q_true).max(), np.abs(lm[:4] + q_true).max()) < 2e-3


def test_kdtree_backend_equals_brute_force_pipeline(op, synth, street):
    traj = synth.trajectory(3)
    a, b = op.Oracle(**KW16, knn_backend=0), op.Oracle(**KW16, knn_backend=1)
    for k in range(3):
        s = street.scan(synth.VLP16, traj[k], 1000 + k)
        a.process(s); b.process(s)
    assert (a.get("lo.pose") == b.get("lo.pose")).all() and (a.get("lm.pose") == b.get("lm.pose")).all()
    assert a.get("lm.surfMap") == b.get("lm.surfMap")


def test_mapping_recovers_planted_offset(op, synth):
    """Scan-to-map against a planted map: start the mapper from a perturbed pose, it must pull back."""
    world = synth.World(1234, 0, 160.0)
    corner, surf = world.plant(0, 0.4), world.plant(1, 0.8)
    o = op.Oracle(n_scans=64, minimum_range=5.0, knn_backend=1)
    o.set("lm.cornerMap", synth.cubes_blob(corner, 0.4)); o.set("lm.surfMap", synth.cubes_blob(surf, 0.8))
    pose = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0.15, -0.10, 0.05], np.float64)   # q_wmap_wodom = I, t_wmap_wodom off by 15 cm
    o.set("lm.pose", pose)
    scan = world.scan(synth.HDL64, [0, 0, 0, 0, 0, 0], 1000)
    o.process(scan)
    t = o.get("lm.pose")[4:7]
    assert np.linalg.norm(t) < 0.04, t


def test_cube_roll_keeps_window_centred(op, synth, street):
    """Drive the mapper 200 m away in one step: the six roll loops must re-centre the 21x21x11 window."""
    o = op.Oracle(n_scans=64, minimum_range=5.0, knn_backend=1)
    scan = street.scan(synth.HDL64, [0, 0, 0, 0, 0, 0], 1000)
    o.process(scan)
    n0 = synth.blob_counts(o.get("lm.surfMap")).sum()
    pose = o.get("lm.pose")
    pose[11:14] = [480.0, -470.0, 0.0]  # t_wmap_wodom: the next initial guess lands far away
    o.set("lm.pose", pose)
    o.process(scan)
    st = o.get("lm.state")
    t = o.get("lm.pose")[4:7]
    ci = int((t[0] + 25.0) / 50.0) + st[0] - (1 if t[0] + 25.0 < 0 else 0)
    cj = int((t[1] + 25.0) / 50.0) + st[1] - (1 if t[1] + 25.0 < 0 else 0)
    assert 3 <= ci < 18 and 3 <= cj < 18
    assert synth.blob_counts(o.get("lm.surfMap")).sum() > 0 and n0 > 0
