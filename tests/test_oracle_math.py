"""Self-checks that pin the oracle's restated third-party pieces (the reference ships
no golden vectors, so these stand in): VoxelGrid, exact kNN, eigen / QR, quaternions, Ceres pieces."""
import numpy as np
import pytest


def np_voxel_grid(cloud, leaf):
    """Independent numpy restatement of pcl::VoxelGrid, float32 arithmetic throughout."""
    c = np.ascontiguousarray(cloud, np.float32)
    if len(c) == 0:
        return c
    inv = np.float32(1.0) / np.float32(leaf)
    mn, mx = c[:, :3].min(0), c[:, :3].max(0)
    d = ((mx - mn) * inv).astype(np.int64) + 1
    if int(d[0]) * int(d[1]) * int(d[2]) > 2**31 - 1:
        return c.copy()
    minb = np.floor(mn * inv).astype(np.int64)
    maxb = np.floor(mx * inv).astype(np.int64)
    div = maxb - minb + 1
    ijk = (np.floor(c[:, :3] * inv) - minb.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.lexsort((np.arange(len(c)), idx))
    out = []
    s = 0
    while s < len(order):
        e = s
        acc = np.zeros(4, np.float32)
        while e < len(order) and idx[order[e]] == idx[order[s]]:
            acc = (acc + c[order[e]]).astype(np.float32)
            e += 1
        out.append(acc / np.float32(e - s))
        s = e
    return np.array(out, np.float32)


@pytest.mark.parametrize("n,leaf,scale", [(1, 0.2, 1.0), (50, 0.2, 1.0), (3000, 0.2, 5.0), (3000, 0.4, 30.0), (2000, 0.8, 60.0)])
def test_voxel_grid_matches_numpy(op, n, leaf, scale):
    rng = np.random.RandomState(n)
    c = (rng.randn(n, 4) * scale).astype(np.float32)
    a, b = op.voxel_grid(c, leaf), np_voxel_grid(c, leaf)
    assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()


def test_voxel_grid_edge_cases(op):
    assert len(op.voxel_grid(np.zeros((0, 4), np.float32), 0.2)) == 0
    # leaf too small for the extent: pcl returns the input unchanged
    c = np.array([[0, 0, 0, 1], [1e6, 1e6, 1e6, 2], [5, 5, 5, 3]], np.float32)
    out = op.voxel_grid(c, 0.2)
    assert (out == c).all()
    # duplicates collapse to their mean, output ascending in voxel index
    c = np.array([[0.05, 0.05, 0.05, 1], [0.06, 0.05, 0.05, 3], [-1, 0, 0, 5]], np.float32)
    out = op.voxel_grid(c, 0.2)
    assert len(out) == 2 and out[0, 3] == 5 and out[1, 3] == 2


def test_voxel_grid_idempotent_on_lattice(op):
    rng = np.random.RandomState(3)
    c = (rng.randn(5000, 4) * 10).astype(np.float32)
    once = op.voxel_grid(c, 0.4)
    twice = op.voxel_grid(once, 0.4)
    # almost every centroid stays in its voxel; the filter never grows the cloud
    assert len(twice) <= len(once) <= len(c)
    assert len(twice) >= 0.99 * len(once)


def test_knn_kdtree_equals_brute_force_with_ties(op):
    rng = np.random.RandomState(5)
    # quantised coordinates force equal distances -> exercises the (d2, index) tie rule
    cloud = np.round(rng.rand(4000, 4) * 8).astype(np.float32)
    q = np.round(rng.rand(300, 4) * 8).astype(np.float32)
    ib, db = op.knn(cloud, q, 5, backend=0)
    ik, dk = op.knn(cloud, q, 5, backend=1)
    assert (ib == ik).all() and (db.view(np.uint32) == dk.view(np.uint32)).all()
    # brute force against numpy
    d = ((q[:, None, :3] - cloud[None, :, :3]).astype(np.float32) ** 2)
    d2 = (d[..., 0] + d[..., 1]).astype(np.float32) + d[..., 2]
    for i in range(len(q)):
        order = np.lexsort((np.arange(len(cloud)), d2[i]))[:5]
        assert (order == ib[i]).all()


def test_sym_eig3(op):
    rng = np.random.RandomState(7)
    for _ in range(200):
        a = rng.randn(3, 5)
        A = a @ a.T
        ev, V = op.sym_eig3(A)
        w, U = np.linalg.eigh(A)
        assert np.allclose(ev, w, rtol=1e-11, atol=1e-12)
        assert np.allclose(A @ V, V * ev, atol=1e-10)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-12)
    ev, V = op.sym_eig3(np.diag([3.0, 1.0, 2.0]))
    assert np.allclose(ev, [1, 2, 3])


def test_qr_solve(op):
    rng = np.random.RandomState(9)
    for _ in range(200):
        A = rng.randn(5, 3) * 10
        b = -np.ones(5)
        x, full = op.qr_solve_5x3(A, b)
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert full and np.allclose(x, ref, rtol=1e-9, atol=1e-11)
    # five collinear points: rank deficient, finite answer
    t = np.linspace(0, 1, 5)[:, None]
    A = np.hstack([t, 2 * t, 3 * t]) + 1.0
    x, full = op.qr_solve_5x3(A, -np.ones(5))
    assert np.isfinite(x).all()


def test_quaternion_semantics(op):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(11)
    for _ in range(50):
        a, b = R.random(random_state=rng.randint(1 << 30)), R.random(random_state=rng.randint(1 << 30))
        v = rng.randn(3)
        ab, av, ai = op.quat(a.as_quat(), b.as_quat(), v)
        assert np.allclose(R.from_quat(ab).as_matrix(), (a * b).as_matrix(), atol=1e-12)
        assert np.allclose(av, a.apply(v), atol=1e-12)
        assert np.allclose(R.from_quat(ai).as_matrix(), a.inv().as_matrix(), atol=1e-12)


def make_factors(rng, pose_q, pose_t, n_edge=60, n_plane=60, n_norm=60, noise=0.0):
    """Factors consistent with a known pose: world = R p + t lies on the line / plane."""
    from scipy.spatial.transform import Rotation as R
    Rm = R.from_quat(pose_q).as_matrix()
    f = []
    for _ in range(n_edge):
        p = rng.randn(3) * 10
        w = Rm @ p + pose_t
        d = rng.randn(3); d /= np.linalg.norm(d)
        a, b = w + 0.7 * d + noise * rng.randn(3), w - 0.9 * d + noise * rng.randn(3)
        f.append([0, *p, *a, *b])
    for _ in range(n_plane):
        p = rng.randn(3) * 10
        w = Rm @ p + pose_t
        n = rng.randn(3); n /= np.linalg.norm(n)
        u = np.cross(n, rng.randn(3)); u /= np.linalg.norm(u)
        j = w + 1.3 * u + noise * rng.randn(3)
        f.append([1, *p, *j, *n])
    for _ in range(n_norm):
        p = rng.randn(3) * 10
        w = Rm @ p + pose_t
        n = rng.randn(3); n /= np.linalg.norm(n)
        f.append([2, *p, *n, -float(n @ w) + noise * rng.randn(), 0, 0])
    return np.array(f, np.float64)


def plus(x, d):
    n = np.linalg.norm(d[:3])
    q = x[:4].copy()
    if n > 0:
        s = np.sin(n) / n
        dq = np.array([s * d[0], s * d[1], s * d[2], np.cos(n)])
        a, b = dq, q
        q = np.array([a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                      a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                      a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0],
                      a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]])
    return np.concatenate([q, x[4:] + d[3:]])


def test_gradient_matches_finite_differences(op):
    """g = J^T r of the robustified problem is the derivative of the cost along Plus(x, eps e_k):
    checks the dual-number Jacobians, the plus-Jacobian contraction and the Huber corrector."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(13)
    q = R.from_euler("xyz", [0.02, -0.03, 0.05]).as_quat()
    t = np.array([0.3, -0.2, 0.1])
    f = make_factors(rng, q, t, noise=0.05)
    x = np.concatenate([R.from_euler("xyz", [0.03, -0.01, 0.02]).as_quat(), [0.5, -0.1, 0.3]])
    cost, H, g = op.evaluate(f, x)
    eps = 1e-6
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        cp = op.evaluate(f, plus(x, d))[0]
        cm = op.evaluate(f, plus(x, -d))[0]
        assert abs((cp - cm) / (2 * eps) - g[k]) < 1e-5 * max(1.0, abs(g[k]))
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > -1e-9)


def test_ceres_solve_recovers_known_pose(op):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(17)
    q = R.from_euler("xyz", [0.01, -0.02, 0.03]).as_quat()
    t = np.array([0.8, -0.1, 0.05])
    f = make_factors(rng, q, t)
    x = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    costs = []
    for _ in range(4):  # each call is one ceres::Solve with max_num_iterations = 4
        x, log = op.ceres_solve(f, x)
        assert log[0] <= 4 and log[3] <= log[2] + 1e-15
        costs.append(log[3])
    assert costs[-1] < 1e-16
    assert np.allclose(x[4:], t, atol=1e-7)
    assert np.allclose(R.from_quat(x[:4]).as_matrix(), R.from_quat(q).as_matrix(), atol=1e-7)


def test_ceres_solve_fixed_point_matches_scipy(op):
    """With noise the LM fixed point equals scipy's minimiser of the same Huber cost."""
    from scipy.optimize import minimize
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(19)
    q = R.from_euler("xyz", [0.01, 0.02, -0.01]).as_quat()
    t = np.array([0.2, 0.1, -0.05])
    f = make_factors(rng, q, t, noise=0.03)
    x = np.concatenate([q, t])
    for _ in range(12):
        x, log = op.ceres_solve(f, x)
    x0 = x.copy()
    res = minimize(lambda d: op.evaluate(f, plus(x0, d))[0], np.zeros(6), method="BFGS", options={"gtol": 1e-12})
    assert np.abs(res.x).max() < 5e-6, res.x
    assert res.fun >= op.evaluate(f, x0)[0] - 1e-10


def test_ceres_solve_no_factors_leaves_x(op):
    x0 = np.array([0.1, 0.2, 0.3, 0.9, 1, 2, 3], np.float64)
    x, log = op.ceres_solve(np.zeros((0, 10)), x0)
    assert (x == x0).all()


def _witness(gold, prefix):
    return {k[len(prefix):]: v for k, v in gold.items() if k.startswith(prefix) and k not in (prefix + "sets", prefix + "family")}


def test_fits_match_numpy_fixture(op):
    """tests/golden/fit_sets.npz (numpy eigh / lstsq answers, generator tests/make_golden_fit.py): the oracle's restated
    Eigen algorithms -- tridiagonal QR eigen-solver, column-pivoted Householder -- against an independent LAPACK witness on
    4096 five-point sets per kind, degenerate families included; and the witness fields regenerate identically."""
    import os
    import fit_checks as fc
    import make_golden_fit as mg
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_sets.npz")))
    sets, _ = mg.fit_sets(4096, 4101)
    assert (sets == gold["line_sets"]).all(), "the committed fixture is not what its generator makes"
    ok, prm = op.fit(gold["line_sets"], 0)
    r = fc.check_line(ok, prm, _witness(gold, "line_"))
    assert r["accepted"] > 2000 and r["decided"] > 3000
    ok, prm = op.fit(gold["plane_sets"], 1)
    r = fc.check_plane(ok, prm, _witness(gold, "plane_"))
    assert r["accepted"] > 2000
    # degenerate families never produce NaN parameters or accepted garbage
    assert np.isfinite(prm).all()


def corridor_factors(rng, pose_q, pose_t, n, kind):
    """Ill-conditioned mapping problems: `corridor` = plane-norm factors on two walls and the floor only (translation along
    the corridor axis unobserved but for a few far-away, nearly parallel planes); `single_plane` = one ground plane (x, y,
    yaw unobserved but for rounding-level tilt).  cond(J^T J) >= 1e8 at the solution."""
    from scipy.spatial.transform import Rotation as R
    Rm = R.from_quat(pose_q).as_matrix()
    f = []
    for i in range(n):
        p = np.array([rng.uniform(-40, 40), rng.uniform(-3, 3), rng.uniform(-1.5, 2.0)])
        w = Rm @ p + pose_t
        if kind == "corridor":
            nrm = [np.array([0.0, 1.0, 0.0]), np.array([0.0, -1.0, 0.0]), np.array([0.0, 0.0, 1.0])][i % 3]
            nrm = nrm + np.array([1e-5 * rng.randn(), 0.0, 0.0])   # walls are parallel to x to within 1e-5 rad
        else:
            nrm = np.array([1e-6 * rng.randn(), 1e-6 * rng.randn(), 1.0])
        nrm = nrm / np.linalg.norm(nrm)
        f.append([2, *p, *nrm, -float(nrm @ w) + 0.01 * rng.randn(), 0, 0])
    return np.array(f, np.float64)


def test_corridor_problems_are_ill_conditioned(op):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(31)
    q = R.from_euler("xyz", [0.01, -0.02, 0.03]).as_quat()
    t = np.array([0.8, -0.1, 0.05])
    for kind in ("corridor", "single_plane"):
        f = corridor_factors(rng, q, t, 900, kind)
        _, H, _ = op.evaluate(f, np.concatenate([q, t]))
        ev = np.linalg.eigvalsh(H)
        assert ev[-1] / max(ev[0], 1e-300) >= 1e8, (kind, ev)
        x, log = op.ceres_solve(f, np.concatenate([q, t + [0.05, 0.02, -0.03]]))
        assert np.isfinite(x).all() and log[3] <= log[2]


def test_deskew_functors_slerp_and_gradient(op):
    """DISTORTION == true (LO.cpp:368-372, LF.hpp:29-36): the functors evaluate Identity.slerp(s, q) * p + s t.  With s == 1
    the literal slerp path equals the shortcut exactly (value and Jacobian: SURVEY A.5); with random s the dual-number
    gradient matches central differences of the robust cost; s = 0 removes the pose from the residual altogether."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(3)
    q = R.from_euler("xyz", [0.02, -0.03, 0.05]).as_quat()
    t = np.array([0.9, -0.2, 0.1])
    f = make_factors(rng, q, t, 60, 60, 0, noise=0.05)
    for x in (np.concatenate([R.from_euler("xyz", [0.03, -0.01, 0.02]).as_quat(), [0.5, -0.1, 0.3]]),
              np.concatenate([-R.from_euler("xyz", [0.3, -0.2, 0.4]).as_quat(), [0.5, -0.1, 0.3]]),      # w < 0: slerp flips scale1
              np.array([0, 0, 0, 1.0, 0.4, 0.0, 0.0])):                                                    # |w| >= 1 - eps: linear branch
        c0, H0, g0 = op.evaluate(f, x)
        c1, H1, g1 = op.evaluate(f, x, np.ones(len(f)))
        assert c0 == c1 and (H0 == H1).all() and (g0 == g1).all()
        s = rng.uniform(0, 1, len(f))
        cost, H, g = op.evaluate(f, x, s)
        eps = 1e-6
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            num = (op.evaluate(f, plus(x, d), s)[0] - op.evaluate(f, plus(x, -d), s)[0]) / (2 * eps)
            assert abs(num - g[k]) < 2e-5 * max(1.0, abs(g[k])), (k, num, g[k])
        _, Hz, gz = op.evaluate(f, x, np.zeros(len(f)))
        assert np.abs(Hz).max() == 0 and np.abs(gz).max() == 0


def test_deskew_pipeline_runs_and_differs(op, synth, street):
    """distortion = 1 changes the odometry (s < 1 for most points) but keeps the pipeline sane on a static-snapshot sweep."""
    traj = synth.trajectory(3)
    a = op.Oracle(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4)
    b = op.Oracle(n_scans=16, minimum_range=0.3, line_res=0.2, plane_res=0.4, distortion=1)
    for k in range(3):
        scan = street.scan(0, traj[k], 1000 + k)
        a.process(scan); b.process(scan)
    pa, pb = a.get("lo.pose"), b.get("lo.pose")
    assert np.isfinite(pb).all() and np.abs(pa - pb).max() > 1e-6
    assert (a.get("sr.sharp") == b.get("sr.sharp")).all()   # scan registration is not affected
