"""ctypes driver of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs."""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Params(ctypes.Structure):
    _fields_ = [("n_scans", ctypes.c_int), ("minimum_range", ctypes.c_float), ("line_res", ctypes.c_float),
                ("plane_res", ctypes.c_float), ("mapping_skip_frame", ctypes.c_int), ("knn_backend", ctypes.c_int), ("distortion", ctypes.c_int)]


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle.so")
        srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        L.vloam_oracle_create.restype = ctypes.c_void_p
        L.vloam_oracle_create.argtypes = [ctypes.POINTER(Params)]
        L.vloam_oracle_destroy.argtypes = [ctypes.c_void_p]
        for f in ("process", "scan_registration"):
            getattr(L, "vloam_oracle_" + f).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.vloam_oracle_laser_odometry.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.vloam_oracle_laser_mapping.argtypes = [ctypes.c_void_p]
        L.vloam_oracle_lo_associate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.vloam_oracle_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        L.vloam_oracle_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        L.vloam_oracle_voxel_grid.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
        L.vloam_oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.vloam_oracle_sym_eig3.argtypes = [ctypes.c_void_p] * 3
        L.vloam_oracle_qr_solve_5x3.argtypes = [ctypes.c_void_p] * 3
        L.vloam_oracle_fit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.vloam_oracle_ceres_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.vloam_oracle_evaluate.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.vloam_oracle_ceres_solve_s.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.vloam_oracle_evaluate_s.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.vloam_oracle_quat.argtypes = [ctypes.c_void_p] * 6
        _lib = L
    return _lib


_DTYPES = {"sr.curvature": np.float32, "sr.label": np.int32, "sr.scanStartInd": np.int32, "sr.scanEndInd": np.int32,
           "lo.pose": np.float64, "lm.pose": np.float64, "lm.poseHighFreq": np.float64, "lm.state": np.int32, "lm.validInd": np.int32,
           "lo.costs": np.float64, "lm.costs": np.float64, "timing": np.float64}


def decode(name, raw):
    """Shared by the oracle and the CUDA library wrappers: bytes -> numpy by buffer name."""
    if name in ("lm.cornerMap", "lm.surfMap"):
        return raw
    if name in _DTYPES:
        return np.frombuffer(raw, _DTYPES[name]).copy()
    if name.startswith("lo.assoc.corner"):
        return np.frombuffer(raw, np.int32).reshape(-1, 2).copy()
    if name.startswith("lo.assoc.surf"):
        return np.frombuffer(raw, np.int32).reshape(-1, 3).copy()
    if name.startswith("lm.knn."):
        kind = name[len("lm.knn."):-1]
        if kind in ("cidx", "sidx"):
            return np.frombuffer(raw, np.int32).reshape(-1, 5).copy()
        if kind in ("cd2", "sd2"):
            return np.frombuffer(raw, np.float32).reshape(-1, 5).copy()
        return np.frombuffer(raw, np.int32).copy()
    return np.frombuffer(raw, np.float32).reshape(-1, 4).copy()  # clouds


class Oracle:
    def __init__(self, n_scans=64, minimum_range=5.0, line_res=0.4, plane_res=0.8, mapping_skip_frame=1, knn_backend=0, distortion=0):
        self.p = Params(n_scans, minimum_range, line_res, plane_res, mapping_skip_frame, knn_backend, distortion)
        self.h = lib().vloam_oracle_create(ctypes.byref(self.p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().vloam_oracle_destroy(self.h)
            self.h = None

    @staticmethod
    def _xyz(a):
        a = np.ascontiguousarray(a, np.float32)
        return a, a.shape[0], a.shape[1]

    def process(self, scan):
        a, n, s = self._xyz(scan)
        lib().vloam_oracle_process(self.h, a.ctypes.data, n, s)

    def scan_registration(self, scan):
        a, n, s = self._xyz(scan)
        lib().vloam_oracle_scan_registration(self.h, a.ctypes.data, n, s)

    def laser_odometry(self, prior_q=None, prior_t=None):
        if prior_q is None:
            return lib().vloam_oracle_laser_odometry(self.h, None, None, 0)
        q = np.ascontiguousarray(prior_q, np.float64)
        t = np.ascontiguousarray(prior_t, np.float64)
        return lib().vloam_oracle_laser_odometry(self.h, q.ctypes.data, t.ctypes.data, 1)

    def laser_mapping(self):
        lib().vloam_oracle_laser_mapping(self.h)

    def lo_associate(self, x):
        x = np.ascontiguousarray(x, np.float64)
        ns, nf = len(self.get("sr.sharp")), len(self.get("sr.flat"))
        ci = np.full((ns, 2), -1, np.int32)
        si = np.full((nf, 3), -1, np.int32)
        lib().vloam_oracle_lo_associate(self.h, x.ctypes.data, ci.ctypes.data, si.ctypes.data)
        return ci, si

    def get_raw(self, name):
        n = lib().vloam_oracle_get(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        buf = ctypes.create_string_buffer(max(n, 1))
        lib().vloam_oracle_get(self.h, name.encode(), buf, n)
        return buf.raw[:n]

    def get(self, name):
        return decode(name, self.get_raw(name))

    def set(self, name, data):
        raw = data if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data).tobytes()
        r = lib().vloam_oracle_set(self.h, name.encode(), raw, len(raw))
        if r != 0:
            raise KeyError(name)

    def set_last(self, corner, surf):
        c = np.ascontiguousarray(corner, np.float32)
        s = np.ascontiguousarray(surf, np.float32)
        self.set("lo.last", np.array([len(c), len(s)], np.int32).tobytes() + c.tobytes() + s.tobytes())


def voxel_grid(cloud, leaf):
    c = np.ascontiguousarray(cloud, np.float32)
    out = np.empty_like(c)
    n = lib().vloam_oracle_voxel_grid(c.ctypes.data, len(c), leaf, out.ctypes.data)
    return out[:n].copy()


def knn(cloud, queries, k, backend=0):
    c = np.ascontiguousarray(cloud, np.float32)
    q = np.ascontiguousarray(queries, np.float32)
    idx = np.empty((len(q), k), np.int32)
    d2 = np.empty((len(q), k), np.float32)
    lib().vloam_oracle_knn(c.ctypes.data, len(c), q.ctypes.data, len(q), k, backend, idx.ctypes.data, d2.ctypes.data)
    return idx, d2


def sym_eig3(A):
    A = np.ascontiguousarray(A, np.float64)
    ev = np.empty(3)
    V = np.empty((3, 3))
    lib().vloam_oracle_sym_eig3(A.ctypes.data, ev.ctypes.data, V.ctypes.data)
    return ev, V


def qr_solve_5x3(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.empty(3)
    ok = lib().vloam_oracle_qr_solve_5x3(A.ctypes.data, b.ctypes.data, x.ctypes.data)
    return x, bool(ok)


def fit(near, kind):
    """LM.cpp:559-603 (kind 0) / 637-680 (kind 1) on n five-point sets float32[n,5,3] -> (ok[n], params[n,6])."""
    a = np.ascontiguousarray(near, np.float32).reshape(-1, 15)
    ok = np.zeros(len(a), np.int32)
    prm = np.zeros((len(a), 6))
    lib().vloam_oracle_fit(a.ctypes.data, len(a), kind, ok.ctypes.data, prm.ctypes.data)
    return ok, prm


def ceres_solve(factors, x, s=None):
    """s: per-factor interpolation ratio (DISTORTION == true: the functors slerp q by s and scale t by s); None = the s == 1 path."""
    f = np.ascontiguousarray(factors, np.float64)
    x = np.array(x, np.float64)
    log = np.zeros(4)
    if s is None:
        lib().vloam_oracle_ceres_solve(f.ctypes.data, len(f), x.ctypes.data, log.ctypes.data)
    else:
        sv = np.ascontiguousarray(s, np.float64)
        lib().vloam_oracle_ceres_solve_s(f.ctypes.data, sv.ctypes.data, len(f), x.ctypes.data, log.ctypes.data)
    return x, log


def evaluate(factors, x, s=None):
    f = np.ascontiguousarray(factors, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    cost = np.zeros(1)
    H = np.zeros((6, 6))
    g = np.zeros(6)
    if s is None:
        lib().vloam_oracle_evaluate(f.ctypes.data, len(f), x.ctypes.data, cost.ctypes.data, H.ctypes.data, g.ctypes.data)
    else:
        sv = np.ascontiguousarray(s, np.float64)
        lib().vloam_oracle_evaluate_s(f.ctypes.data, sv.ctypes.data, len(f), x.ctypes.data, cost.ctypes.data, H.ctypes.data, g.ctypes.data)
    return cost[0], H, g


def quat(a, b, v):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); v = np.ascontiguousarray(v, np.float64)
    ab = np.empty(4); av = np.empty(3); ai = np.empty(4)
    lib().vloam_oracle_quat(a.ctypes.data, b.ctypes.data, v.ctypes.data, ab.ctypes.data, av.ctypes.data, ai.ctypes.data)
    return ab, av, ai
