"""KITTI .bin sweeps and the reference's pose-file line (vloam_tf.cpp:84-160); CPU only."""
import os
import subprocess
import numpy as np
import pytest


def test_velodyne_bin_round_trip(pkg, tmp_path):
    k = pkg.kitti_io
    pts = np.random.RandomState(0).randn(1000, 4).astype(np.float32)
    p = str(tmp_path / "000000.bin")
    k.write_velodyne_bin(p, pts)
    assert (k.read_velodyne_bin(p) == pts).all()
    open(p, "ab").write(b"\0\0")
    with pytest.raises(ValueError):
        k.read_velodyne_bin(p)


def test_pose_writer_matches_reference_formula(pkg, tmp_path):
    from scipy.spatial.transform import Rotation as R
    k = pkg.kitti_io
    rng = np.random.RandomState(1)
    base_T_cam0 = np.eye(4)
    base_T_cam0[:3, :3] = R.from_euler("xyz", [-1.57, 0.01, -1.56]).as_matrix()
    base_T_cam0[:3, 3] = [0.27, -0.05, -0.08]
    path = str(tmp_path / "MO.txt")
    w = k.KittiPoseWriter(path, base_T_cam0)
    Ts = []
    for i in range(5):
        q = R.from_euler("xyz", rng.randn(3) * 0.1).as_quat()
        t = rng.randn(3) * 3
        T = np.eye(4); T[:3, :3] = R.from_quat(q).as_matrix(); T[:3, 3] = t
        Ts.append(np.linalg.inv(base_T_cam0) @ T @ base_T_cam0)
        w.write(q, t)
    w.close()
    got = k.read_poses(path)
    assert got.shape == (5, 3, 4)
    assert np.abs(got[0] - np.eye(4)[:3]).max() < 1e-6            # relative to the first written frame
    for i in range(5):
        ref = (np.linalg.inv(Ts[0]) @ Ts[i]).astype(np.float32)[:3]
        assert np.abs(got[i] - ref).max() < 2e-6                  # "%f": six decimals
    assert all(len(l.split()) == 12 for l in open(path))


def test_calib_tr(pkg, tmp_path):
    p = tmp_path / "calib.txt"
    p.write_text("P0: 1 0 0 0 0 1 0 0 0 0 1 0\nTr: 0 -1 0 0.1 0 0 -1 0.2 1 0 0 0.3\n")
    m = pkg.kitti_io.read_calib_tr(str(p))
    assert m.shape == (4, 4) and m[0, 3] == 0.1 and m[2, 0] == 1 and m[3, 3] == 1


def test_cpp_pose_writer_agrees_with_python(pkg, tmp_path):
    """The header-only C++ writer (csrc/vloam_kitti_io.hpp) prints the same lines as the Python one."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"
    src.write_text('''
#include "vloam_kitti_io.hpp"
int main(int argc, char** argv) {
  vloam::Mat4 b = {0, 0, 1, 0.27, -1, 0, 0, -0.05, 0, -1, 0, -0.08, 0, 0, 0, 1};
  vloam::KittiPoseWriter w(argv[1], b);
  const double q[3][4] = {{0, 0, 0, 1}, {0.01, -0.02, 0.03, 0.9993}, {0.1, 0.0, -0.05, 0.99373}};
  const double t[3][3] = {{0, 0, 0}, {1.5, 0.1, -0.02}, {3.25, -0.4, 0.07}};
  for (int i = 0; i < 3; ++i) w.write(q[i], t[i]);
  auto v = vloam::read_velodyne_bin(argv[2]);
  printf("%zu\\n", v.size());
  return 0;
}''')
    exe = str(tmp_path / "t")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "vloam-noted_b200", "csrc"), "-o", exe, str(src)], check=True)
    k = pkg.kitti_io
    binp = str(tmp_path / "s.bin"); k.write_velodyne_bin(binp, np.zeros((7, 4), np.float32))
    out = subprocess.run([exe, str(tmp_path / "cpp.txt"), binp], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "28"
    b = np.array([[0, 0, 1, 0.27], [-1, 0, 0, -0.05], [0, -1, 0, -0.08], [0, 0, 0, 1]], float)
    w = k.KittiPoseWriter(str(tmp_path / "py.txt"), b)
    for q, t in (((0, 0, 0, 1), (0, 0, 0)), ((0.01, -0.02, 0.03, 0.9993), (1.5, 0.1, -0.02)), ((0.1, 0.0, -0.05, 0.99373), (3.25, -0.4, 0.07))):
        w.write(q, t)
    w.close()
    a, c = k.read_poses(str(tmp_path / "cpp.txt")), k.read_poses(str(tmp_path / "py.txt"))
    assert np.abs(a - c).max() <= 1e-6
