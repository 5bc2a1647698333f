"""KITTI on-disk formats either side of the path.

* `read_velodyne_bin`: a KITTI odometry sweep (`velodyne/000000.bin`, float32 x y z reflectance), the layout the
  reference's bag player feeds to `ScanRegistration::input`; pass it to `Context.process_frame` (stride 4).
* `KittiPoseWriter`: the evaluation file the reference writes per frame in `VloamTF::{VO,LO,MO}2Cam0StartFrame`
  (vloam_tf.cpp:84-160): the pose of camera 0 relative to camera 0 at the first written frame,
      cam0_start_T_cam0_last = (cam0_init_T_cam0_start)^-1 * (base_T_cam0^-1 * world_T_base_last * base_T_cam0),
  cast to float and printed as the 12 row-major entries of its 3x4 part with "%f".
"""
import numpy as np


def read_velodyne_bin(path):
    """float32[n, 4] (x, y, z, reflectance) from a KITTI .bin sweep; raises on a truncated file."""
    import os
    if os.path.getsize(path) % 16:
        raise ValueError("%s: size is not a multiple of 4 floats" % path)
    return np.fromfile(path, dtype=np.float32).reshape(-1, 4)


def write_velodyne_bin(path, xyzr):
    np.ascontiguousarray(xyzr, np.float32).reshape(-1, 4).tofile(path)


def read_calib_tr(path):
    """`Tr:` (velodyne -> camera 0, 3x4) of a KITTI odometry calib.txt as a 4x4 matrix."""
    for line in open(path):
        if line.startswith("Tr:"):
            v = np.array([float(x) for x in line.split()[1:13]]).reshape(3, 4)
            return np.vstack([v, [0, 0, 0, 1]])
    raise ValueError("%s: no Tr: line" % path)


def pose_matrix(q_xyzw, t):
    """Eigen::Quaterniond(x, y, z, w) + translation -> 4x4 (toRotationMatrix convention, no normalisation)."""
    x, y, z, w = [float(v) for v in q_xyzw]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    m = np.eye(4)
    m[0, 0] = 1 - (tyy + tzz); m[0, 1] = txy - twz; m[0, 2] = txz + twy
    m[1, 0] = txy + twz; m[1, 1] = 1 - (txx + tzz); m[1, 2] = tyz - twx
    m[2, 0] = txz - twy; m[2, 1] = tyz + twx; m[2, 2] = 1 - (txx + tyy)
    m[:3, 3] = t
    return m


class KittiPoseWriter:
    """One line per frame, as VloamTF::MO2Cam0StartFrame (vloam_tf.cpp:133-160).  base_T_cam0: pose of camera 0
    in the frame the lidar poses are expressed in (for a velodyne-frame odometry: inverse of calib `Tr`)."""

    def __init__(self, path, base_T_cam0=None):
        self.f = open(path, "w") if path else None
        self.base_T_cam0 = np.eye(4) if base_T_cam0 is None else np.asarray(base_T_cam0, float)
        self.cam0_T_base = np.linalg.inv(self.base_T_cam0)
        self.start_inv = None
        self.rows = []

    def write(self, q_xyzw, t):
        cam0_init_T_last = self.cam0_T_base @ pose_matrix(q_xyzw, t) @ self.base_T_cam0
        if self.start_inv is None:  # count == 0
            self.start_inv = np.linalg.inv(cam0_init_T_last)
        m = (self.start_inv @ cam0_init_T_last).astype(np.float32)
        line = " ".join("%f" % float(v) for v in m[:3].reshape(-1))
        self.rows.append(m[:3].copy())
        if self.f:
            self.f.write(line + "\n")
        return line

    def close(self):
        if self.f:
            self.f.close(); self.f = None


def read_poses(path):
    """KITTI pose file -> float64[n, 3, 4]."""
    return np.loadtxt(path).reshape(-1, 3, 4)
