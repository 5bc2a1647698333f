"""Shared assertions for the neighbourhood fits (LM.cpp:559-603, 637-680): an implementation (the oracle's restated Eigen
algorithms, or the CUDA path through vloam_b200_fit) against the numpy / LAPACK witness of tests/make_golden_fit.py."""
import numpy as np


def check_line(ok, prm, w, tol=1e-12):
    """ok[n], prm[n,6] = {a, b} against the eigh witness: flags on every decided set, and on accepted sets the line
    points a, b = centre +- 0.1 v2 (the eigenvector's sign is free: either assignment)."""
    ok = np.asarray(ok).astype(bool)
    dec = w["decided"]
    bad = np.flatnonzero(dec & (ok != w["accept"]))
    assert len(bad) == 0, "line accept flag differs on %d decided sets, first %s" % (len(bad), bad[:5])
    sel = np.flatnonzero(dec & ok)
    a_ref = w["centre"][sel] + 0.1 * w["v2"][sel]
    b_ref = w["centre"][sel] - 0.1 * w["v2"][sel]
    a, b = prm[sel, :3], prm[sel, 3:]
    e1 = np.maximum(np.abs(a - a_ref).max(1), np.abs(b - b_ref).max(1))
    e2 = np.maximum(np.abs(a - b_ref).max(1), np.abs(b - a_ref).max(1))
    err = np.minimum(e1, e2)
    # eigenvector conditioning: error ~ eps * |cov| / gap; the accept test guarantees gap >= 2/3 lam2
    scale = 1.0 + np.abs(w["centre"][sel]).max(1)
    assert (err <= tol * scale * 10).all(), "line factor differs: max %.3g (scaled %.3g)" % (err.max(), (err / scale).max())
    return {"sets": int(len(ok)), "decided": int(dec.sum()), "accepted": int(len(sel)), "max_err": float(err.max()) if len(sel) else 0.0}


def check_plane(ok, prm, w, tol=1e-9):
    ok = np.asarray(ok).astype(bool)
    dec = w["decided"]
    bad = np.flatnonzero(dec & (ok != w["accept"]))
    assert len(bad) == 0, "plane accept flag differs on %d decided sets, first %s" % (len(bad), bad[:5])
    sel = np.flatnonzero(dec & ok)
    en = np.abs(prm[sel, :3] - w["n"][sel]).max(1) if len(sel) else np.zeros(0)
    ed = np.abs(prm[sel, 3] - w["d"][sel]) if len(sel) else np.zeros(0)
    # least squares through points ~100 m from the origin: cond(A) up to ~1e4 amplifies the 1e-16 rounding
    assert (en <= tol).all() and (ed <= tol * (1.0 + np.abs(w["d"][sel]))).all(), "plane factor differs: n %.3g, d %.3g" % (en.max(), ed.max())
    return {"sets": int(len(ok)), "decided": int(dec.sum()), "accepted": int(len(sel)), "max_err_n": float(en.max()) if len(sel) else 0.0,
            "max_err_d": float(ed.max()) if len(sel) else 0.0}


def check_same(ok_a, prm_a, ok_b, prm_b, kind, decided=None):
    """Two implementations against each other on ALL sets (degenerate ones included): number of differing accept flags
    (on the sets `decided` marks -- a knife-edge set such as a lattice with lambda_2 == 3 lambda_1 exactly is decided by the
    last rounding of whichever eigen-solver runs), and the largest scaled parameter difference on sets both accept, up to
    the line's sign freedom."""
    ok_a, ok_b = np.asarray(ok_a).astype(bool), np.asarray(ok_b).astype(bool)
    if decided is not None:
        ok_a = ok_a & decided; ok_b = ok_b & decided
    both = ok_a & ok_b
    if kind == 0:
        e1 = np.maximum(np.abs(prm_a[:, :3] - prm_b[:, :3]).max(1), np.abs(prm_a[:, 3:] - prm_b[:, 3:]).max(1))
        e2 = np.maximum(np.abs(prm_a[:, :3] - prm_b[:, 3:]).max(1), np.abs(prm_a[:, 3:] - prm_b[:, :3]).max(1))
        err = np.minimum(e1, e2)[both]
    else:
        err = np.abs(prm_a - prm_b).max(1)[both]
    scale = 1.0 + np.abs(prm_a[both]).max(1) if both.any() else np.ones(0)
    return int((ok_a != ok_b).sum()), float((err / scale).max()) if both.any() else 0.0
