"""Five-point neighbourhood sets for the two fits of LaserMapping::solveMapping (LM.cpp:559-603 line test,
LM.cpp:637-680 plane fit) and their answers from an INDEPENDENT witness: numpy's LAPACK eigh / lstsq (neither the
oracle's restated Eigen algorithms nor the CUDA path's Jacobi / Householder code).

    python tests/make_golden_fit.py        regenerates tests/golden/fit_sets.npz (4096 sets per kind)

fit_sets(n, seed) is also imported by the tests to make larger sets on the fly (the GPU test uses >= 1e5).

Witness fields per set:
  line : lam[3] (ascending), accept = lam2 > 3 lam1, decided = the margin |lam2 - 3 lam1| > 1e-9 lam2 (only decided sets
         pin the flag), centre[3], v2[3] (unit eigenvector of the largest eigenvalue, sign free)
  plane: n[3] (unit), d, accept = all |n.p + d| <= 0.2, decided = full column rank (cond < 1e7) and no residual within
         1e-9 of 0.2 (rank-deficient sets are where ColPivHouseholderQR's basic solution and LAPACK's minimum-norm one
         legitimately differ; those only compare the CUDA path with the restated Eigen algorithm)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fit_sets(n, seed):
    """float32[n, 5, 3]: map-like neighbourhoods around centres up to ~120 m from the origin -- noisy lines, noisy planes,
    blobs, and the degenerate families (exactly collinear, exactly coplanar, duplicated points, all five identical)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((n, 5, 3))
    fam = rng.randint(0, 8, n)
    cen = rng.uniform(-120, 120, (n, 3)) * [1, 1, 0.1]
    for i in range(n):
        f = fam[i]
        d = rng.randn(3); d /= np.linalg.norm(d)
        e = np.cross(d, rng.randn(3)); e /= np.linalg.norm(e)
        t = rng.uniform(-0.8, 0.8, 5)
        u = rng.uniform(-0.8, 0.8, 5)
        if f == 0:    # line + small noise (a pole / an edge in the map at 0.4 m voxels)
            p = t[:, None] * d + rng.randn(5, 3) * rng.choice([0.005, 0.02, 0.08])
        elif f == 1:  # plane + small noise
            p = t[:, None] * d + u[:, None] * e + rng.randn(5, 3) * rng.choice([0.005, 0.02, 0.08, 0.15])
        elif f == 2:  # blob
            p = rng.randn(5, 3) * rng.choice([0.05, 0.3])
        elif f == 3:  # exactly collinear on a lattice (f32-exact coordinates)
            p = np.round(t * 4)[:, None] * np.round(d * 2) * 0.25
        elif f == 4:  # exactly coplanar, axis-aligned lattice
            p = np.stack([np.round(t * 4) * 0.25, np.round(u * 4) * 0.25, np.zeros(5)], 1)[:, rng.permutation(3)]
        elif f == 5:  # duplicated points
            p = (t[:, None] * d + rng.randn(5, 3) * 0.02)[[0, 0, 1, 1, 2]]
        elif f == 6:  # all five identical
            p = np.zeros((5, 3))
        else:         # thick line: near the lam2 = 3 lam1 boundary on purpose
            p = t[:, None] * d + u[:, None] * e * rng.uniform(0.3, 0.9)
        out[i] = p + (np.round(cen[i] * 4) * 0.25 if f in (3, 4, 6) else cen[i])
    return out.astype(np.float32), fam


def witness_line(sets):
    P = sets.astype(np.float64)
    cen = P.sum(1) / 5.0
    Z = P - cen[:, None, :]
    cov = np.einsum("nja,njb->nab", Z, Z)
    lam, vec = np.linalg.eigh(cov)
    accept = lam[:, 2] > 3 * lam[:, 1]
    decided = np.abs(lam[:, 2] - 3 * lam[:, 1]) > 1e-9 * np.maximum(lam[:, 2], 1e-300)
    return {"lam": lam, "accept": accept, "decided": decided, "centre": cen, "v2": vec[:, :, 2]}


def witness_plane(sets):
    P = sets.astype(np.float64)
    n = len(P)
    nrm, d, accept, decided = np.zeros((n, 3)), np.zeros(n), np.zeros(n, bool), np.zeros(n, bool)
    for i in range(n):
        A = P[i]
        sv = np.linalg.svd(A, compute_uv=False)
        x = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        nn = np.linalg.norm(x)
        if not (nn > 0) or sv[-1] < 1e-7 * sv[0]:
            continue
        nrm[i], d[i] = x / nn, 1.0 / nn
        res = np.abs(A @ nrm[i] + d[i])
        accept[i] = (res <= 0.2).all()
        decided[i] = (np.abs(res - 0.2) > 1e-9).all()
    return {"n": nrm, "d": d, "accept": accept, "decided": decided}


def make(n=4096):
    ls, lf = fit_sets(n, 4101)
    ps, pf = fit_sets(n, 4102)
    out = {"line_sets": ls, "line_family": lf.astype(np.int8), "plane_sets": ps, "plane_family": pf.astype(np.int8)}
    for k, v in witness_line(ls).items(): out["line_" + k] = v
    for k, v in witness_plane(ps).items(): out["plane_" + k] = v
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "fit_sets.npz")
    np.savez_compressed(path, **make())
    print("written", os.path.getsize(path), "bytes")
